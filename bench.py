#!/usr/bin/env python
"""bench.py — Gibbs sweeps/sec of the GP-IRT sampler (BASELINE.json metric) on synthetic data of a named n x m shape.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl reference]

A "step" is one Gibbs sweep (reference loop body, src/gpirtMCMC.cpp:68-78) over the whole response matrix.
  value : sweeps/s with everything resident in HBM, timed by CUDA events on the sampler's own stream inside the C library
          (gpirt_b200_sampler_sweep), max over ranks.
  e2e   : sweeps/s through the public drop-in call gpirtMCMC() -> C-ABI gpirt_b200_mcmc() with HOST buffers: the H2D
          copy of the response matrix and the per-sweep D2H of the theta / beta / f draws are inside the timed region.
  N > 1 : launched under torchrun, one rank per GPU; items are sharded across ranks (strong scaling of the named
          workload) and the per-respondent log-posterior partial sums are all-reduced over NCCL once per sweep.
  --impl reference : the reference's own CPU sampler (its sources compiled against stand-in Armadillo headers,
          oracle/_ref; else the oracle port) on a bounded item sample, extrapolated linearly in m (stated in `sample`).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from gpirt_b200 import synthetic  # noqa: E402

N_GRID = 1001


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return json.load(fh), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's loop body on host cores, on a bounded item sample
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_sweeps(cfg, data, steps, warmup, budget_s, threads):
    """Times `steps` sweeps of the reference CPU path on the first m_s items (all n respondents) and extrapolates
    T(m) = T_fixed + (T(m_s) - T_fixed) * m / m_s  (every step but K/Cholesky and the L^-1 K* solve is linear in m)."""
    from oracle import oracle as O
    n, m = cfg["n"], cfg["m"]
    use_ref = os.path.exists(O.REF_SO)
    O.set_blas_threads(threads)
    per_item = 1.35e-7 * n * N_GRID + 1.2e-8 * n * n       # rough seconds/item (theta grid loop + three O(n^2) passes)
    fixed = 6e-11 * n ** 3 / max(1, min(threads, 8)) + 2e-8 * n * n
    m_s = int(max(4, min(m, 128, (budget_s / max(1, steps + warmup) - fixed) / per_item)))
    y = np.asfortranarray(data["y"][:, :m_s])
    pm, psd, pstep = data["pm"][:, :m_s], data["psd"][:, :m_s], data["pstep"][:, :m_s]
    theta = data["theta_init"].copy()
    ts, prior = O.grid()
    if use_ref:
        O.ref().gpref_seed(12345)
        def chol(th):
            S = O.ref_K(th, th)                                # K(theta, theta), gpirtMCMC.cpp:76
            S[np.diag_indices(n)] += 0.001                     # :77
            return O.ref_chol_lower(S)                         # :78
        draw_f = lambda f, L, mu: O.ref_draw_f(f, y, L, mu)                      # noqa: E731
        draw_fstar = lambda f, th, L, mus: O.ref_draw_fstar(f, th, ts, L, mus)   # noqa: E731
        draw_theta = lambda fs, mus: O.ref_draw_theta(ts, y, prior, fs, mus)     # noqa: E731
        draw_beta = lambda b, th, f: O.ref_draw_beta(b, th, y, f, pm, psd, pstep)  # noqa: E731
        kind = "reference"
    else:
        rng = O.Rng.keyed(12345)
        chol = lambda th: O.build_cholS(th)                                      # noqa: E731
        draw_f = lambda f, L, mu: O.draw_f(f, y, L, mu, rng)[0]                  # noqa: E731
        draw_fstar = lambda f, th, L, mus: O.draw_fstar(f, th, ts, L, mus, rng)[0]  # noqa: E731
        draw_theta = lambda fs, mus: O.draw_theta(ts, y, prior, fs, rng, mode=0)[0]  # noqa: E731
        draw_beta = lambda b, th, f: O.draw_beta(b, th, y, f, pm, psd, pstep, rng)[0]  # noqa: E731
        kind = "port"
    rs = np.random.RandomState(7)
    L = chol(theta)
    f = np.asfortranarray(L @ rs.randn(n, m_s))
    beta = np.asfortranarray(rs.randn(2, m_s) * 3.0)
    t_sweep, t_fixed = [], []
    for it in range(warmup + steps):
        if kind == "port":
            rng.set_sweep(it + 1)
        mu, mus = O.linear_mean(theta, beta), O.linear_mean(ts, beta)
        t0 = time.perf_counter()
        f = draw_f(f, L, mu)                                   # gpirtMCMC.cpp:68
        fs = draw_fstar(f, theta, L, mus)                      # :69
        th_new = draw_theta(fs, mus)                           # :70
        if not np.all(np.isfinite(th_new)):                    # the reference's theta_star[N] read — cannot
            th_new = np.where(np.isfinite(th_new), th_new, theta)  # happen at m_s <= 128; guard keeps the run alive
        theta = th_new
        beta = draw_beta(beta, theta, f)                       # :72
        mu, mus = O.linear_mean(theta, beta), O.linear_mean(ts, beta)   # :74-75
        t1 = time.perf_counter()
        L = chol(theta)                                        # :76-78
        t2 = time.perf_counter()
        if it >= warmup:
            t_sweep.append(t2 - t0); t_fixed.append(t2 - t1)
    # fixed part of draw_fstar (K*, L^-1 K*): time it with a single item
    t0 = time.perf_counter()
    if use_ref:
        O.ref_draw_fstar(f[:, :1], theta, ts, L, O.linear_mean(ts, beta[:, :1]))
    else:
        O.draw_fstar(f[:, :1], theta, ts, L, O.linear_mean(ts, beta[:, :1]), rng)
    t_fs1 = time.perf_counter() - t0
    T_s, T_fix = float(np.mean(t_sweep)), float(np.mean(t_fixed)) + t_fs1
    T_full = T_fix + max(0.0, T_s - T_fix) * (m / m_s)
    sample = ("%d sweep(s) of the %s on the first %d of %d items (all %d respondents), %.2f s/sweep measured; "
              "extrapolated linearly in m to %.1f s/sweep (fixed part K+chol+L^-1K* %.2f s)" %
              (steps, "reference sources (oracle/_ref)" if use_ref else "oracle port", m_s, m, n, T_s, T_full, T_fix))
    if m_s == m:
        sample = "%d full sweep(s) of the %s, %.2f s/sweep" % (steps, "reference sources (oracle/_ref)" if use_ref else "oracle port", T_s)
    return dict(value=1.0 / T_full, unit="sweeps/s", cores=threads if threads > 1 else 1, kind=kind, sample=sample,
                blas_threads=threads, sampler_threads=1, s_per_sweep=T_full, measured_s=float(np.sum(t_sweep)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=sorted(synthetic.WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--ess-samples", type=int, default=300)
    ap.add_argument("--fstar-mode", type=int, default=0)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K, W = max(1, args.steps), max(0, args.warmup)
    cfg = synthetic.WORKLOADS[args.workload]
    n, m = cfg["n"], cfg["m"]
    config = {"workload": "%s (%s), synthetic 2PL responses, seed %d, priors pm=0 psd=3 step=0.1" % (args.workload, cfg["desc"], synthetic.SEED),
              "n": n, "m": m, "n_grid": N_GRID, "cache": "inputs larger than L2 (f, Z, nu are 3 x %.0f MB)" % (n * m * 8 / 1e6),
              "parallelism": "items sharded over %d GPU(s), logP all-reduce per sweep" % world if world > 1 else "single GPU"}
    host_threads = os.cpu_count() or 1

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        data = synthetic.make(n, m)
        res = cpu_reference_sweeps(cfg, data, K, W, budget_s=150.0, threads=host_threads)
        line = {"impl": "reference", "metric": "gibbs_sweeps_per_sec", "value": res["value"], "unit": "sweeps/s", "n_gpus": args.gpus,
                "steps": K, "warmup": W, "ms_per_step": 1000.0 * res["s_per_sweep"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": res["value"], "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import gpirt_b200.sampler as G
    from gpirt_b200 import _lib
    dist = None
    uid = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from gpirt_b200.sharding import share_unique_id

        def fresh_uid():   # one ncclUniqueId per communicator
            return share_unique_id(dist, rank, G.nccl_unique_id, device="cuda")
        uid = fresh_uid()
    data = synthetic.make(n, m)
    # contiguous item block of this rank
    from gpirt_b200.sharding import item_block
    j0, j1 = item_block(m, rank, world)
    y_loc = np.asfortranarray(data["y"][:, j0:j1])
    kw = dict(seed=synthetic.SEED, device=local_rank, fstar_mode=args.fstar_mode)
    if world > 1:
        kw.update(rank=rank, world_size=world, m_global=m, item_offset=j0, nccl_unique_id=uid)
    s = G.Sampler(y_loc, data["theta_init"], data["pm"][:, j0:j1], data["psd"][:, j0:j1], data["pstep"][:, j0:j1], **kw)
    s.init_draws()
    s.sweep(max(3, W))                       # >= 3 untimed warm-up sweeps
    s.timings(reset=True)
    launches0 = s.launches()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    if dist is not None:
        dist.barrier()
    ms = s.sweep(K)                          # CUDA events on the sampler's stream, sync on both sides inside the call
    if dist is not None:
        import torch
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    clk = clocks.finish() if rank == 0 else None
    launches = s.launches() - launches0
    timers = s.timings()
    value = 1000.0 * K / ms

    # roofline of the dominant KERNEL: timer segments are grouped by the kernel that runs them (the fixed-point tensor-core
    # GEMM serves three segments), the group with the largest summed CUDA-event time in the timed region is reported
    m_loc = j1 - j0
    work = {"lz_gemm": ("tensor", float(n) * n * m_loc), "fstar_gemm": ("tensor", 2.0 * n * N_GRID * m_loc),
            "chol": ("tensor", n ** 3 / 3.0), "trtri": ("tensor", n ** 3 / 3.0),
            "trsm": ("tensor", 2.0 * n * n * N_GRID if args.fstar_mode == 0 else n * n * N_GRID + 2.0 * n * n * m_loc),
            "ess": ("hbm", 25.0 * n * m_loc), "beta": ("hbm", 17.0 * n * m_loc), "kbuild": ("hbm", 4.0 * n * n)}
    fixed_point = s.uses(1) == 1
    groups = {k: [k] for k in work}
    names = {"chol": "potrf_lower_rl (k_diag128 + gemm_f64_kernel updates)", "trtri": "trtri_lower (gemm_f64_kernel)",
             "lz_gemm": "gemm_f64_kernel (nu = L Z)", "fstar_gemm": "gemm_f64_kernel (f* product)", "trsm": "gemm_f64_kernel (K* solves)",
             "ess": "k_ess_persist", "beta": "k_beta", "kbuild": "k_se_cov"}
    if fixed_point:
        # 36 exact int8 plane-pair products on tcgen05 (dgemm_i8.cu): the tensor pipe executes 36 x the FP64 product's work
        i8_segs = ["lz_gemm", "fstar_gemm"] + (["trsm"] if args.fstar_mode == 0 and world == 1 else [])
        for k in i8_segs:
            del groups[k]
        groups["k_dgemm_i8"] = i8_segs
        names["k_dgemm_i8"] = "k_dgemm_i8 (%s)" % " + ".join(i8_segs)
    seg_ms = {k: timers[k][0] / K for k in work}              # per sweep, pipelined (a segment may be several launches)
    dom = max(groups, key=lambda g: sum(seg_ms[k] for k in groups[g]))
    segs = groups[dom]
    dom_ms = sum(seg_ms[k] for k in segs)
    fp64_work = sum(work[k][1] for k in segs)
    bound = work[segs[0]][0]
    peaks, peak_src = _peaks()
    dmma, dfma = G.fp64_peak_tflops()
    fp64_src = ("FP64 tensor pipe (DMMA.8x8x4) issue-rate microbenchmark measured in this run; "
                "MEASURED_PEAKS.json has no FP64 figure")
    # the same segments timed WITHOUT sweep pipelining (no co-running kernels): kernel quality, not schedule
    s.set_pipeline(False)
    s.sweep(1)
    s.timings(reset=True)
    Ki = 3
    ms_iso = s.sweep(Ki)
    t_iso = s.timings()
    s.set_pipeline(True)
    iso_ms = sum(t_iso[k][0] for k in segs) / Ki
    i8_note = None
    if dom == "k_dgemm_i8":
        # tcgen05 kind::i8 issues twice the multiply-adds per instruction of kind::f16 (K = 32 vs 16), so the int8
        # ceiling is 2 x the measured dense bf16 figure: sustained inside the long step, burst for the kernel alone
        alg = 36.0 * fp64_work
        peak = 2.0 * peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
        iso_peak = 2.0 * peaks["bf16_tflops"]
        achieved, iso, unit = alg / dom_ms * 1e-9, alg / iso_ms * 1e-9, "TFLOP/s"
        peak_src = "2 x dense bf16 (%s) = int8 tensor ops/s; nominal int8 dense is 4500" % peak_src
        i8_note = {"plane_pair_products": 36, "fp64_tensor_peak_tflops": dmma,
                   "fp64_equivalent_tflops": fp64_work / dom_ms * 1e-9, "fp64_equivalent_tflops_isolated": fp64_work / iso_ms * 1e-9,
                   "frac_of_nominal_int8_isolated": iso / 4500.0,
                   "segments_ms": {k: seg_ms[k] for k in segs}, "segments_ms_isolated": {k: t_iso[k][0] / Ki for k in segs},
                   "note": "segment times include the operand slicing kernels of each product"}
        bound = "tensor"
    elif bound == "tensor":
        alg = fp64_work
        achieved, iso, peak, iso_peak, unit, peak_src = alg / dom_ms * 1e-9, alg / iso_ms * 1e-9, dmma, dmma, "TFLOP/s", fp64_src
    else:
        alg = fp64_work
        achieved, iso, peak, iso_peak, unit = alg / dom_ms * 1e-6, alg / iso_ms * 1e-6, peaks["hbm_gbs"], peaks["hbm_gbs"], "GB/s"
    traffic = None
    tp = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tp):
        try:
            with open(tp) as fh:
                traffic = json.load(fh).get(args.workload, {}).get(dom)
        except Exception:
            traffic = None
    roofline = {"kernel": names.get(dom, dom), "bound": bound, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_work_per_sweep": alg,
                "ms_per_sweep": dom_ms, "share_of_step": dom_ms / (ms / K),
                "note": "timed region runs pipelined: the L Z product and the beta step execute UNDER the Cholesky chain and the K* "
                        "solves beside the ESS, so their event durations include co-running kernels (shares can sum to > 1); "
                        "`isolated` repeats the measurement with pipelining off",
                "isolated": {"achieved": iso, "peak": iso_peak, "frac": iso / iso_peak, "ms_per_sweep": iso_ms,
                             "sweep_ms_unpipelined": ms_iso / Ki},
                "fixed_point": i8_note,
                "per_step_ms": {k: v[0] / K for k, v in timers.items() if v[1]},
                "per_step_ms_isolated": {k: v[0] / Ki for k, v in t_iso.items() if v[1]}}
    s.close()

    line = {"metric": "gibbs_sweeps_per_sec", "value": value, "unit": "sweeps/s", "n_gpus": world, "steps": K, "warmup": max(3, W),
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config, "gpu_launches": int(launches), "clocks": clk, "roofline": roofline}

    # ------------------------------------------------------------------ chain-parallel mode (BASELINE config 4), N > 1 only
    # every rank runs an INDEPENDENT chain of the whole workload (own seed, no communication): aggregate sweeps/s
    if world > 1 and not args.no_extras:
        sc = G.Sampler(data["y"], data["theta_init"], data["pm"], data["psd"], data["pstep"], seed=synthetic.SEED + 1000 + rank,
                       device=local_rank, fstar_mode=args.fstar_mode)
        sc.set_timing(False)
        sc.init_draws()
        sc.sweep(max(3, W))
        dist.barrier()
        ms_c = sc.sweep(K)
        import torch
        t = torch.tensor([ms_c], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sc.close()
        line["chain_parallel"] = {"value": 1000.0 * K * world / float(t.item()), "unit": "sweeps/s (sum over %d independent chains)" % world,
                                  "ms_per_step_per_chain": float(t.item()) / K, "scaling": "weak",
                                  "note": "one full-size chain per GPU, no collective (BASELINE config 4); `value` above is the item-sharded single chain (config 3)"}

    # ------------------------------------------------------------------ end-to-end through the public call (host buffers)
    if not args.no_e2e:
        from gpirt_b200 import ResponseMatrix
        Ke = min(K, 12)   # f draws are n*m*8 bytes per stored sweep on the host; bound the host allocation
        yrm = ResponseMatrix(y_loc)
        common = dict(beta_prior_means=data["pm"][:, j0:j1], beta_prior_sds=data["psd"][:, j0:j1],
                      beta_proposal_sds=data["pstep"][:, j0:j1], theta_init=data["theta_init"], seed=synthetic.SEED,
                      device=local_rank, fstar_mode=args.fstar_mode)
        shard = (rank, world, m, j0, fresh_uid()) if world > 1 else None
        G.gpirtMCMC(yrm, 1, 0, shard=shard, **common)     # warm the call path (allocator, pinning, communicator)
        shard = (rank, world, m, j0, None) if world > 1 else None   # re-use the communicator of the previous call
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        out = G.gpirtMCMC(yrm, Ke, 0, shard=shard, **common)
        el = time.perf_counter() - t0
        if dist is not None:
            import torch
            t = torch.tensor([el], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
        assert np.isfinite(out["theta"]).all()
        h2d = (n * m_loc + n + 6 * m_loc) * 8 / Ke
        d2h = (n * m_loc + 2 * m_loc + n) * 8 * (Ke + 1) / Ke + N_GRID * m_loc * 8 / Ke
        line["e2e"] = {"value": Ke / el, "unit": "sweeps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                       "note": "gpirtMCMC(sample_iterations=%d, burn_iterations=0) wall time incl. setup, initial draws, H2D of y "
                               "and D2H of every theta/beta/f draw (reference output contract)" % Ke}
        del out

    # ------------------------------------------------------------------ theta ESS/sec (second half of BASELINE's metric), N = 1 only
    if rank == 0 and world == 1 and not args.no_extras:
        try:
            from gpirt_b200 import ResponseMatrix
            from gpirt_b200.diagnostics import ess_geyer
            S_ess, B_ess = args.ess_samples, 50
            t0 = time.perf_counter()
            ch = G.gpirtMCMC(ResponseMatrix(y_loc), S_ess, B_ess, beta_prior_means=data["pm"], beta_prior_sds=data["psd"],
                             beta_proposal_sds=data["pstep"], theta_init=data["theta_init"], seed=synthetic.SEED, device=local_rank,
                             store_f=False, fstar_mode=args.fstar_mode)
            el_ess = time.perf_counter() - t0
            ess = ess_geyer(ch["theta"][1:])
            ess = ess[np.isfinite(ess)]
            line["theta_ess"] = {"samples": S_ess, "burn": B_ess, "seconds": el_ess, "ess_median": float(np.median(ess)),
                                 "ess_min": float(ess.min()), "ess_per_sec_median": float(np.median(ess) / el_ess),
                                 "ess_per_sec_min": float(ess.min() / el_ess),
                                 "estimator": "Geyer initial positive sequence per respondent; seconds = wall time of the "
                                              "gpirtMCMC(S, B, store_f=False) call incl. burn-in",
                                 "corr_with_generating_theta": float(abs(np.corrcoef(ch["theta"][1:].mean(axis=0), data["theta_true"])[0, 1]))}
            if "e2e" in line:   # the same public call without the n x m x (S+1) f array: what bounds `e2e` is storing f
                line["e2e"]["without_f_draws"] = {"value": (S_ess + B_ess) / el_ess, "unit": "sweeps/s",
                                                  "note": "gpirtMCMC(%d, %d, store_f=False) wall time: theta, beta and IRFs "
                                                          "still come back to the host every sweep" % (S_ess, B_ess)}
        except Exception as ex:
            line["theta_ess"] = {"error": repr(ex)}

    # ------------------------------------------------------------------ the other single-GPU configs, briefly (N = 1 only)
    if rank == 0 and world == 1 and args.workload == "c3" and not args.no_extras:
        extras = {}
        for wl in ("c1", "c2"):
            try:
                c = synthetic.WORKLOADS[wl]
                d2 = synthetic.make(c["n"], c["m"])
                s2 = G.Sampler(d2["y"], d2["theta_init"], d2["pm"], d2["psd"], d2["pstep"], seed=synthetic.SEED, device=local_rank)
                s2.set_timing(False)
                s2.init_draws()
                s2.sweep(5)
                k2 = 50
                ms2 = s2.sweep(k2)
                extras[wl] = {"n": c["n"], "m": c["m"], "value": 1000.0 * k2 / ms2, "unit": "sweeps/s", "ms_per_step": ms2 / k2, "steps": k2}
                s2.close()
            except Exception as ex:
                extras[wl] = {"error": repr(ex)}
        line["other_workloads"] = extras

    # ------------------------------------------------------------------ CPU baseline beside it (rank 0, N = 1 only)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            res = cpu_reference_sweeps(cfg, data, 1, 0, budget_s=20.0, threads=host_threads)
            line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as ex:  # the baseline is informative; never lose the GPU line over it
            line["cpu_baseline"] = {"value": None, "unit": "sweeps/s", "cores": host_threads, "kind": "port", "sample": "failed: %r" % (ex,)}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
