"""The host steps either side of the sampler on the native side: response coding on the device
(R/response_matrix.R:79-98, numeric codes) against the Python mirror of the reference function, and the multi-chain theta
diagnostics (split-R-hat, Geyer ESS) against numpy restatements."""
import warnings

import numpy as np
import pytest


def _split_rhat(chains):
    halves = []
    for c in chains:
        h = c.shape[0] // 2
        halves += [c[:h], c[c.shape[0] - h:]]
    x = np.stack(halves)                       # (2 chains, half, n)
    h = x.shape[1]
    W = x.var(axis=1, ddof=1).mean(axis=0)
    B = h * x.mean(axis=1).var(axis=0, ddof=1)
    return np.sqrt(((h - 1) / h * W + B / h) / W)


def test_theta_diagnostics_match_numpy():
    import gpirt_b200.sampler as G
    from gpirt_b200.diagnostics import ess_geyer
    rs = np.random.RandomState(0)
    chains = []
    for c in range(3):                         # AR(1) chains with different persistence per respondent, on the 0.01 grid
        e = rs.randn(400, 7)
        x = np.zeros_like(e)
        phi = np.linspace(0.0, 0.95, 7)
        for t in range(1, 400):
            x[t] = phi * x[t - 1] + e[t]
        chains.append(np.round(x + 0.1 * c, 2))
    rhat, ess = G.theta_diagnostics(chains)
    assert np.max(np.abs(rhat - _split_rhat(chains))) <= 1e-12
    want = sum(ess_geyer(c) for c in chains)
    assert np.max(np.abs(ess - want) / want) <= 1e-9
    assert ess[0] > 5 * ess[-1]                # the sticky respondent has the small effective sample size
    r1, e1 = G.theta_diagnostics(chains[:1])
    assert np.max(np.abs(e1 - ess_geyer(chains[0])) / e1) <= 1e-9 and np.all(np.isfinite(r1))


@pytest.mark.gpu
def test_native_response_coding_matches_the_reference_function():
    import gpirt_b200
    import gpirt_b200.sampler as G
    codes, _, _ = gpirt_b200.senate116()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = np.asarray(gpirt_b200.response_matrix(codes))
    y, kept, uncoded = G.response_matrix_native(np.asarray(codes, dtype=np.float64))
    assert y.shape == want.shape == (100, 418) and uncoded == 0
    assert np.array_equal(np.isnan(y), np.isnan(want)) and np.array_equal(np.nan_to_num(y), np.nan_to_num(want))
    # random codes: an uncoded value (11), NA cells, unanimous items, an all-missing item (kept: the reference's rule is == 1)
    rs = np.random.RandomState(1)
    c = rs.choice([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 11], size=(60, 40)).astype(float)
    c[rs.rand(60, 40) < 0.05] = np.nan
    c[:, 3] = 1; c[:, 7] = np.where(rs.rand(60) < 0.5, 5, 9); c[:, 11] = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = np.asarray(gpirt_b200.response_matrix(c))
    y, kept, uncoded = G.response_matrix_native(c)
    assert y.shape == want.shape and uncoded == int((c == 11).sum())
    assert 3 not in kept and 7 not in kept and 11 in kept
    assert np.array_equal(np.isnan(y), np.isnan(want)) and np.array_equal(np.nan_to_num(y), np.nan_to_num(want))
    # overlapping code lists: later rules win (yea, then nay, then missing)
    y2, _, _ = G.response_matrix_native(np.array([[1.0, 2.0], [2.0, 1.0], [3.0, 3.0]]), yea=(1, 2), nay=(2, 3), missing=(3,))
    assert np.array_equal(np.nan_to_num(y2, nan=9.0), np.array([[1.0, -1.0], [-1.0, 1.0], [9.0, 9.0]]))
