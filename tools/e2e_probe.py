import sys, time
import numpy as np
sys.path.insert(0, ".")
from gpirt_b200 import synthetic, ResponseMatrix
import gpirt_b200.sampler as G
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfg = synthetic.WORKLOADS[wl]
d = synthetic.make(cfg["n"], cfg["m"])
y = ResponseMatrix(d["y"])
for rep in range(2):
    t0 = time.perf_counter()
    out = G.gpirtMCMC(y, K, 0, theta_init=d["theta_init"], seed=1)
    el = time.perf_counter() - t0
    print("rep %d: %d sweeps in %.3f s -> %.2f sweeps/s" % (rep, K, el, K / el), flush=True)
t0 = time.perf_counter()
out = G.gpirtMCMC(y, K, 0, theta_init=d["theta_init"], seed=1, store_f=False)
el = time.perf_counter() - t0
print("no f draws: %d sweeps in %.3f s -> %.2f sweeps/s" % (K, el, K / el), flush=True)
