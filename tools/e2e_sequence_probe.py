import sys, time
import numpy as np
sys.path.insert(0, ".")
from gpirt_b200 import synthetic, ResponseMatrix
import gpirt_b200.sampler as G
d = synthetic.make(4096, 10000)
y = ResponseMatrix(d["y"])
G.gpirtMCMC(y, 1, 0, theta_init=d["theta_init"], seed=1)
for name, K, kw in [("full 30", 30, {}), ("thin10+summary", 100, dict(thin=10, f_summary=True)), ("thin10+summary again", 100, dict(thin=10, f_summary=True)), ("no f", 300, dict(store_f=False))]:
    t0 = time.perf_counter()
    out = G.gpirtMCMC(y, K, 0, theta_init=d["theta_init"], seed=1, **kw)
    el = time.perf_counter() - t0
    print("%-22s %4d sweeps in %.3f s -> %.2f sweeps/s" % (name, K, el, K / el), flush=True)
    del out
