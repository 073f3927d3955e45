"""Stress: repeat short chains on a workload and report failures / non-finite state (hunting cross-stream hazards)."""
import sys, os
import numpy as np
sys.path.insert(0, ".")
from gpirt_b200 import synthetic, _lib
import gpirt_b200.sampler as G
wl = sys.argv[1]; reps = int(sys.argv[2]); sweeps = int(sys.argv[3])
cfg = synthetic.WORKLOADS[wl]
d = synthetic.make(cfg["n"], cfg["m"])
fails = 0
ref = None
for r in range(reps):
    try:
        s = G.Sampler(d["y"], d["theta_init"], seed=7)
        s.set_timing(os.environ.get("STRESS_TIMING", "0") == "1")
        s.init_draws()
        s.sweep(sweeps)
        th = s.get(_lib.THETA); f = s.get(_lib.F)
        s.close()
        sig = (float(th.sum()), float(np.abs(f).sum()))
        if ref is None: ref = sig
        ok = np.isfinite(sig[1]) and sig == ref
        if not ok: fails += 1; print("rep", r, "MISMATCH", sig, ref, flush=True)
    except Exception as ex:
        fails += 1; print("rep", r, "EXC", ex, flush=True)
        try:
            th = s.get(_lib.THETA); f = s.get(_lib.F); b = s.get(_lib.BETA); L = s.get(_lib.CHOL); nu = s.get(_lib.NU); fs = s.get(_lib.FSTAR)
            print("   theta finite %s range [%g,%g] ongrid %s | f nan %d | beta nan %d absmax %g | L nan %d | nu nan %d | fstar nan %d absmax %g" % (
                np.isfinite(th).all(), np.nanmin(th), np.nanmax(th), np.allclose(th*100, np.round(th*100)), np.isnan(f).sum(), np.isnan(b).sum(),
                np.nanmax(np.abs(b)), np.isnan(L).sum(), np.isnan(nu).sum(), np.isnan(fs).sum(), np.nanmax(np.abs(fs))), flush=True)
            bad = np.argwhere(np.isnan(L))
            if bad.size: print("   first NaN in L at", bad[0], "count per col head", np.isnan(L).sum(axis=0)[:12], flush=True)
        except Exception as ex2:
            print("   (state dump failed: %s)" % ex2)
print("stress %s: %d/%d failed, signature %r" % (wl, fails, reps, ref))
