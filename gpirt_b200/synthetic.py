"""Synthetic 2PL response data of a named n x m shape — the generator of the reference's own example
(R/gpirtMCMC.R:49-68): theta_i ~ N(0,1), a_j ~ N(0,1), b_j ~ U(0.5,3), y_ij = +1 w.p. plogis(a_j + b_j theta_i) else -1.
Unanimous items (which response_matrix() would drop, R/response_matrix.R:80-88) are repaired so exactly m items remain."""
import numpy as np

WORKLOADS = {
    "c1": dict(n=100, m=418, desc="senate116-shaped 100 x 418"),
    "c2": dict(n=1024, m=2000, desc="synthetic binary responses n=1024 x m=2000"),
    "c3": dict(n=4096, m=10000, desc="synthetic n=4096 x m=10000"),
    "c5": dict(n=16384, m=20000, desc="synthetic n=16384 x m=20000"),
}
SEED = 20261018


def make(n, m, seed=SEED, missing=0.0):
    rs = np.random.Generator(np.random.Philox(seed))
    theta = rs.standard_normal(n)
    a = rs.standard_normal(m)
    b = rs.uniform(0.5, 3.0, m)
    y = np.empty((n, m), order="F")
    blk = max(1, (1 << 22) // n)
    for j0 in range(0, m, blk):
        j1 = min(m, j0 + blk)
        p = 1.0 / (1.0 + np.exp(-(a[None, j0:j1] + b[None, j0:j1] * theta[:, None])))
        y[:, j0:j1] = np.where(rs.random((n, j1 - j0)) < p, 1.0, -1.0)
        if missing > 0:
            y[:, j0:j1][rs.random((n, j1 - j0)) < missing] = np.nan
    for j in range(m):  # repair unanimous items instead of dropping them (keeps the named shape)
        col = y[:, j]
        obs = np.flatnonzero(~np.isnan(col))
        if n >= 2 and (obs.size < 2 or np.all(col[obs] == col[obs[0]])):
            col[0], col[1] = 1.0, -1.0
    theta_init = rs.standard_normal(n)
    return dict(y=y, theta_true=theta, theta_init=theta_init, pm=np.zeros((2, m)), psd=np.full((2, m), 3.0),
                pstep=np.full((2, m), 0.1))
