"""Synthetic stand-ins for the covariance blobs that are absent from the reference checkout.

The reference's `.MISSING_LARGE_BLOBS` lists the Pantheon+ `covariance_stat_sys.txt` (1701^2) and the
DES-Dovekie `STAT+SYS.npy` (1820^2) as not shipped, so every Pantheon+/DES fit script fails at import
(SURVEY.md D8).  Parity and benchmark runs therefore use a seeded SPD matrix with a realistic diagonal and
a rank-40 "systematics" part.  The same recipe feeds the
reference (golden generation), the oracle and the CUDA engine, so all three see identical inputs.
"""
import numpy as np

__all__ = ["synthetic_sn_covariance", "uniform_theta"]


def synthetic_sn_covariance(sigma_diag, seed=12345, rank=40, amp=0.02):
    """C = diag(sigma^2) + A A^T with A ~ N(0, amp^2), shape (n, rank), from default_rng(seed)."""
    sigma_diag = np.asarray(sigma_diag, dtype=np.float64)
    n = sigma_diag.size
    a = np.random.default_rng(seed).standard_normal((n, rank)) * amp
    return np.diag(sigma_diag**2) + a @ a.T


def uniform_theta(bounds, n, seed=42, shrink=1e-3):
    """Seeded uniform parameter batch strictly inside the open prior box `bounds[d,2]`."""
    bounds = np.asarray(bounds, dtype=np.float64)
    lo, hi = bounds[:, 0], bounds[:, 1]
    w = hi - lo
    return np.random.default_rng(seed).uniform(lo + shrink * w, hi - shrink * w, (n, lo.size))
