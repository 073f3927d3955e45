"""Developer benchmark of the small-n path (C1-shaped, launch-latency-bound): sweeps/s with and without graph replay."""
import os, sys
sys.path.insert(0, ".")
from gpirt_b200 import synthetic
import gpirt_b200.sampler as G
wl = sys.argv[1] if len(sys.argv) > 1 else "c1"
c = synthetic.WORKLOADS[wl]
d = synthetic.make(c["n"], c["m"])
for graph in (0, -1):
    s = G.Sampler(d["y"], d["theta_init"], seed=1, use_graph=graph)
    s.set_timing(False)
    s.init_draws()
    s.sweep(5)
    ms = s.sweep(500)
    print("%s use_graph=%d: %.4f ms/sweep  %.0f sweeps/s  launches/sweep %.1f" % (wl, graph, ms / 500, 500000 / ms, s.launches() / 505), flush=True)
    s.close()
