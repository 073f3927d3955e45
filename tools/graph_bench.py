"""Sweep time of the resident sampler the way gpirt_b200_mcmc runs it (timers off, graph replay) at a named workload:
   [GPIRT_SOLVE_MODE=1 ...] python tools/graph_bench.py c3 20"""
import sys
sys.path.insert(0, ".")
from gpirt_b200 import synthetic
import gpirt_b200.sampler as G

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cfg = synthetic.WORKLOADS[wl]
d = synthetic.make(cfg["n"], cfg["m"])
s = G.Sampler(d["y"], d["theta_init"], seed=1)
s.set_timing(False)
s.init_draws()
s.sweep(4)
ms = s.sweep(K)
print("%s n=%d m=%d route %d: %.3f ms/sweep (%.1f sweeps/s), graph replays %d" % (wl, cfg["n"], cfg["m"], s.uses(4), ms / K, 1000 * K / ms, s.uses(5)))
