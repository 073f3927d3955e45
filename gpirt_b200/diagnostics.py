"""Chain diagnostics for the theta draws: effective sample size by Geyer's initial positive sequence estimator
(theta is discrete on the 0.01 grid; the estimator only needs autocovariances)."""
import numpy as np


def ess_geyer(x):
    """x: (draws,) or (draws, chains-as-columns). Returns ESS per column."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    T = x.shape[0]
    xc = x - x.mean(axis=0, keepdims=True)
    var = (xc * xc).mean(axis=0)
    nfft = 1 << int(np.ceil(np.log2(2 * T)))
    f = np.fft.rfft(xc, n=nfft, axis=0)
    acov = np.fft.irfft(f * np.conj(f), n=nfft, axis=0)[:T] / T          # biased autocovariance, lag 0..T-1
    out = np.empty(x.shape[1])
    for c in range(x.shape[1]):
        if not var[c] > 0:
            out[c] = np.nan                                               # constant chain
            continue
        rho = acov[:, c] / acov[0, c]
        # sums of adjacent pairs Gamma_k = rho_{2k} + rho_{2k+1}; stop at the first non-positive pair
        pairs = rho[0:2 * (T // 2):2] + rho[1:2 * (T // 2):2]
        k = np.argmax(pairs <= 0) if np.any(pairs <= 0) else pairs.size
        pairs = np.minimum.accumulate(pairs[:k]) if k > 0 else pairs[:0]   # initial monotone sequence
        tau = -1.0 + 2.0 * pairs.sum()
        out[c] = T / max(tau, 1e-12)
    return out
