"""Developer benchmark: per-step CUDA-event timings of the resident sampler at a named workload."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
from gpirt_b200 import synthetic, _lib
import gpirt_b200.sampler as G

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cfg = synthetic.WORKLOADS[wl]
t0 = time.time()
d = synthetic.make(cfg["n"], cfg["m"])
print("gen %.1fs" % (time.time() - t0), flush=True)
s = G.Sampler(d["y"], d["theta_init"], seed=1, fstar_mode=mode)
s.init_draws()
s.sweep(2)
s.timings(reset=True)
ms = s.sweep(sweeps)
t = s.timings()
n, m, N = cfg["n"], cfg["m"], 1001
flops = {"lz_gemm": n * n * m, "fstar_gemm": 2.0 * n * N * m, "theta_gemm": 2.0 * n * N * m, "chol": n ** 3 / 3.0, "trtri": n ** 3 / 3.0,
         "trsm": 2.0 * n * n * N if mode == 0 else (n * n * N + 2.0 * n * n * m)}  # triangular products: n^2 per RHS each
print("workload %s n=%d m=%d: %.3f ms/sweep (%.1f sweeps/s), launches/sweep %.0f" % (wl, n, m, ms / sweeps, 1000 * sweeps / ms, s.launches() / (sweeps + 3)))
tot = sum(v[0] for v in t.values())
for k, (a, c) in t.items():
    if c:
        per = a / sweeps
        extra = ""
        if k in flops:
            extra = "  %.1f TF/s" % (flops[k] / per * 1e-9)
        print("  %-12s %9.3f ms/sweep  %5.1f%%%s" % (k, per, 100 * a / tot, extra))
nprop = s.get(_lib.ESS_NPROP)
print("  ess proposals/item mean %.2f max %d" % (nprop.mean(), nprop.max()))
th = s.get(_lib.THETA)
print("  corr(theta, truth) = %.3f" % np.corrcoef(th, d["theta_true"])[0, 1])
