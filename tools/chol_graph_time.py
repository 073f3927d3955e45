"""K build + Cholesky alone on the GPU at a named workload's n: eager launch sequence vs replayed as a CUDA graph."""
import sys
sys.path.insert(0, ".")
from gpirt_b200 import synthetic
import gpirt_b200.sampler as G
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
cfg = synthetic.WORKLOADS[wl]
d = synthetic.make(cfg["n"], 512)
s = G.Sampler(d["y"], d["theta_init"], seed=1)
s.init_draws()
s.sweep(2)
n = cfg["n"]
for rep in range(2):
    for g in (False, True):
        ms = s.time_factorisation(10, g)
        print("n = %d  %-12s %7.3f ms   %6.2f TFLOP/s" % (n, "graph replay" if g else "eager", ms, n ** 3 / 3.0 / ms * 1e-9), flush=True)
