"""e2e probe of the draw-storage options: full draws, thin=10 with on-device f summaries, no f draws (GPU box)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from gpirt_b200 import synthetic, ResponseMatrix
import gpirt_b200.sampler as G
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
cfg = synthetic.WORKLOADS[wl]
d = synthetic.make(cfg["n"], cfg["m"])
y = ResponseMatrix(d["y"])
G.gpirtMCMC(y, 4, 0, theta_init=d["theta_init"], seed=1)   # warm: library, pools, bounce buffers
for name, K, kw in [("full draws", 12, {}), ("thin=10 + f summary", 100, dict(thin=10, f_summary=True)),
                    ("thin=10", 100, dict(thin=10)), ("store_f=False", 200, dict(store_f=False))]:
    t0 = time.perf_counter()
    out = G.gpirtMCMC(y, K, 0, theta_init=d["theta_init"], seed=1, **kw)
    el = time.perf_counter() - t0
    print("%-22s %4d sweeps in %.3f s -> %.2f sweeps/s" % (name, K, el, K / el), flush=True)
