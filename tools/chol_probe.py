"""Probe: one Cholesky (n from argv) + one big GEMM through the C ABI, for ncu launch lists."""
import sys
import numpy as np
sys.path.insert(0, ".")
import gpirt_b200.sampler as G
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rs = np.random.RandomState(0)
th = np.round(rs.randn(n), 2)
S = G.se_cov(th, th, jitter=1e-3)
L = G.chol_lower(S)
print("chol ok", np.abs(L @ L.T - S).max())
if len(sys.argv) > 2:
    m = int(sys.argv[2])
    Z = rs.randn(n, m)
    C = G.dgemm(L, Z, tri=1)
    print("gemm ok", np.abs(C - L @ Z).max())
