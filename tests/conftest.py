import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def make_problem(n, m, seed=0, missing=0.1, grid_theta=False):
    """Synthetic 2PL responses as in the reference's own example (R/gpirtMCMC.R:49-68)."""
    rs = np.random.RandomState(seed)
    theta = rs.randn(n)
    if grid_theta:  # theta already on the 0.01 grid (duplicates!), as after the first sweep
        theta = np.round(np.clip(theta, -5, 5), 2)
    a = rs.randn(m)
    b = rs.uniform(0.5, 3.0, m)
    p = 1.0 / (1.0 + np.exp(-(a[None, :] + b[None, :] * theta[:, None])))
    y = np.where(rs.rand(n, m) < p, 1.0, -1.0)
    if missing > 0:
        y[rs.rand(n, m) < missing] = np.nan
    # no unanimous items
    for j in range(m):
        col = y[:, j]
        obs = col[~np.isnan(col)]
        if n >= 2 and (obs.size < 2 or np.all(obs == obs[0])):
            y[0, j], y[1, j] = 1.0, -1.0
    pm = np.zeros((2, m)); psd = np.full((2, m), 3.0); pstep = np.full((2, m), 0.1)
    return dict(y=np.asfortranarray(y), theta=theta, pm=pm, psd=psd, pstep=pstep)


@pytest.fixture(scope="session")
def O():
    from oracle import oracle
    oracle.lib()
    return oracle
