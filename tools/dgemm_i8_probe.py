"""Development probe for the fixed-point (int8 tensor core) FP64 GEMM: correctness on small shapes, then timing at the
sampler's two product shapes.  Run on the GPU box:  python tools/dgemm_i8_probe.py [quick]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import gpirt_b200.sampler as G  # noqa: E402


def check(M, N, K, ta, lower, seed=0):
    rs = np.random.RandomState(seed)
    A = rs.randn(M, K) * np.exp(rs.randn(M, 1))
    if lower:
        A = np.tril(A)
    B = rs.randn(K, N) * np.exp(rs.randn(1, N))
    Ain = np.asfortranarray(A.T) if ta else np.asfortranarray(A)
    C = G.dgemm_i8(Ain, np.asfortranarray(B), ta=ta, a_lower=lower)
    ref = A @ B
    bound = K * 2.0 ** -51 * np.abs(A).max(axis=1)[:, None] * np.abs(B).max(axis=0)[None, :]
    err = np.abs(C - ref)
    print(f"M={M} N={N} K={K} ta={int(ta)} lower={int(lower)}: max err {err.max():.3e}  max err/bound {np.max(err / bound):.3e}"
          f"  rel-to-|A||B| {np.max(err / (np.abs(A) @ np.abs(B))):.3e}", flush=True)
    assert np.all(err <= bound + 1e-300), "outside the stated bound"


if __name__ == "__main__":
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    if not (len(sys.argv) > 1 and sys.argv[1] == "bench"):
        check(128, 64, 64, False, False)
        check(128, 64, 128, True, False)
        check(100, 37, 50, False, False)
        check(300, 200, 257, True, False)
        check(257, 130, 257, False, True)
        check(1000, 333, 1000, False, True)
        check(1001, 500, 1100, True, False)
    if not quick:
        n, m = 4096, 10000
        rs = np.random.RandomState(1)
        L = np.asfortranarray(np.tril(rs.randn(n, n)) / 64.0)
        Z = np.asfortranarray(rs.randn(n, m))
        t0 = time.time()
        C, ms = G.dgemm_i8(L, Z, a_lower=True, reps=5)
        print(f"L Z  {n}x{n} lower x {n}x{m}: kernel {ms[0]:.3f} ms  slicing {ms[1]:.3f} ms   ({n * n * m / ms[0] / 1e9:.1f} FP64-equivalent TFLOP/s,"
              f" {36 * n * n * m / ms[0] / 1e12:.2f} int8 POP/s)  wall {time.time() - t0:.1f}s", flush=True)
        ref = L[:256] @ Z
        print("   max err (first 256 rows)", np.abs(C[:256] - ref).max(), " last rows", np.abs(C[-64:] - L[-64:] @ Z).max())
        A = np.asfortranarray(rs.randn(n, 1001))
        C, ms = G.dgemm_i8(A, Z, ta=True, reps=5)
        print(f"A^T f  1001x{n} x {n}x{m}: kernel {ms[0]:.3f} ms  slicing {ms[1]:.3f} ms   ({2 * 1001 * n * m / ms[0] / 1e9:.1f} FP64-equivalent TFLOP/s,"
              f" {2 * 36 * 1024 * n * m / ms[0] / 1e12:.2f} int8 POP/s)", flush=True)
        print("   max err", np.abs(C[:64] - A[:, :64].T @ Z).max())
