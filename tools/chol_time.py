"""Probe: isolated time of the Cholesky (and L^-1) inside the resident sampler at a named workload, pipelining off."""
import sys
sys.path.insert(0, ".")
from gpirt_b200 import synthetic
import gpirt_b200.sampler as G
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
cfg = synthetic.WORKLOADS[wl]
m = int(sys.argv[2]) if len(sys.argv) > 2 else 512
d = synthetic.make(cfg["n"], m)
s = G.Sampler(d["y"], d["theta_init"], seed=1)
s.set_pipeline(False)
s.init_draws()
s.sweep(2)
s.timings(reset=True)
K = 5
s.sweep(K)
t = s.timings()
n = cfg["n"]
for k in ("kbuild", "chol", "trtri", "trsm"):
    if t[k][1]:
        print("%-8s %8.3f ms  %6.2f TF/s" % (k, t[k][0] / K, (n ** 3 / 3.0) / (t[k][0] / K) * 1e-9 if k in ("chol", "trtri") else 0.0))
