"""Parity at BASELINE.json's full single-GPU size (n = 4096 respondents, m = 10000 items), where the CPU oracle needs
hours per sweep: size-independent properties of the same kernels the small lock-step tests pin against the oracle.
Everything goes through the C ABI (gpirt_b200/_lib.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N, M = 4096, 10000


@pytest.fixture(scope="module")
def G():
    import gpirt_b200.sampler as G
    from gpirt_b200 import _lib
    if _lib.load().gpirt_b200_device_count() < 1:
        pytest.skip("no CUDA device")
    return G


@pytest.fixture(scope="module")
def data():
    from gpirt_b200 import synthetic
    return synthetic.make(N, M)


def test_cholesky_of_the_full_size_covariance(G, O):
    """K(theta, theta) + 1e-3 I at n = 4096 with theta on the 0.01 grid (duplicated rows: PD only through the jitter)"""
    theta = np.round(np.clip(np.random.RandomState(1).randn(N), -5, 5), 2)
    S = O.K(theta, theta) + 1e-3 * np.eye(N)
    L = G.chol_lower(S)
    assert np.array_equal(np.triu(L, 1), np.zeros_like(L))
    assert np.max(np.abs(L @ L.T - S)) <= 1e-12 * 12
    assert np.all(np.diag(L) > 0) and np.max(np.abs(L)) < 2.0       # the bound the fixed-scale slicing of L relies on


def test_fixed_point_product_at_full_k(G):
    """nu = L Z with the contraction length of the full problem: stated bound, determinism, agreement with FP64 DMMA,
    and linearity L (Z1 + Z2) = L Z1 + L Z2 to the same bound"""
    rs = np.random.RandomState(2)
    L = np.tril(rs.randn(N, N)) / 64.0
    Z1, Z2 = rs.randn(N, 640), rs.randn(N, 640)
    a = G.dgemm_i8(L, Z1, a_lower=1)
    assert np.array_equal(a, G.dgemm_i8(L, Z1, a_lower=1))
    bound = N * 2.0 ** -51 * np.abs(L).max(axis=1)[:, None] * np.abs(Z1).max(axis=0)[None, :]
    assert np.all(np.abs(a - L @ Z1) <= bound)
    assert np.max(np.abs(a - G.dgemm(L, Z1, None, tri=1))) <= 1e-12
    lin = G.dgemm_i8(L, Z1 + Z2, a_lower=1) - a - G.dgemm_i8(L, Z2, a_lower=1)
    assert np.max(np.abs(lin)) <= 3 * N * 2.0 ** -51 * np.abs(L).max() * 12.0


def test_full_size_sweeps_pipelined_and_unpipelined_agree(G, data):
    """three Gibbs sweeps at 4096 x 10000: the pipelined schedule (Cholesky chain over the L Z slices, beta step and K*
    solves) and the plain sequential one draw the same theta and the same f / beta to rounding; draws are well-formed"""
    from gpirt_b200 import _lib
    out = {}
    for pipe in (True, False):
        s = G.Sampler(data["y"], data["theta_init"], data["pm"], data["psd"], data["pstep"], seed=11)
        s.set_pipeline(pipe)
        s.init_draws()
        s.sweep(3)
        out[pipe] = {k: s.get(f) for k, f in (("theta", _lib.THETA), ("beta", _lib.BETA), ("f", _lib.F), ("nprop", _lib.ESS_NPROP))}
        s.close()
    a, b = out[True], out[False]
    assert np.array_equal(a["theta"], b["theta"])
    assert np.max(np.abs(a["beta"] - b["beta"])) <= 1e-10
    assert np.max(np.abs(a["f"] - b["f"])) <= 1e-9
    assert np.array_equal(a["nprop"], b["nprop"])
    grid = np.round(a["theta"] * 100.0)
    assert np.max(np.abs(a["theta"] * 100.0 - grid)) < 1e-9 and a["theta"].min() >= -5.0 and a["theta"].max() <= 5.0
    assert np.isfinite(a["f"]).all() and np.isfinite(a["beta"]).all()
    assert a["nprop"].min() >= 1
