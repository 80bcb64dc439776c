"""Profile-likelihood grids with the magnitude offset (and, for SN-only late-time models, H0) handled analytically
(BASELINE.json config 4).  NOT in the reference: parity is unpinned; tests check the closed forms against
brute-force evaluation of the engine's own chi_squared.

For a Cholesky SN block, delta = delta0 - M 1, so with y = L^-1 delta0 and u = L^-1 1 (static):
    chi2(M) = y.y - 2 M y.u + M^2 u.u,   min_M chi2 = y.y - (y.u)^2 / u.u   at  M* = y.u / u.u,
    -2 ln int dM exp(-chi2/2) = y.y - (y.u)^2/u.u + ln(u.u / 2 pi).
In the late-time family D_M is proportional to 1/H0, i.e. H0 only shifts delta by 5 log10(H0/H0_ref): the H0 axis of a grid costs no
GEMM rows at all.  `Engine.sn_moments` returns (y.y, y.u, u.u) per row from the two-dot epilogue of the chi-squared GEMM.
"""
from __future__ import annotations


import numpy as np


def offset_profile(moments, mode="profile"):
    """chi2 with the offset profiled ("profile") or marginalised with a flat prior ("marginal"); also returns M*."""
    yy, yu, uu = moments[..., 0], moments[..., 1], moments[..., 2]
    chi2 = yy - yu * yu / uu
    if mode == "marginal":
        chi2 = chi2 + np.log(uu / (2 * np.pi))
    elif mode != "profile":
        raise ValueError("mode must be 'profile' or 'marginal'")
    return chi2, yu / uu


def chi2_at_offset(moments, M):
    yy, yu, uu = moments[..., 0], moments[..., 1], moments[..., 2]
    return yy - 2 * M * yu + M * M * uu


def sn_profile_grid(engine, axes, mode="profile", chunk=65536, h0_axis=None, h0_ref=70.0):
    """chi2 on the Cartesian grid of `axes` = {theta column: 1-D array} (every column except the offset column; if
    `h0_axis` (values of H0) is given, the H0 column is fixed at `h0_ref` on the GPU and the axis is applied in closed
    form).  Returns (chi2, M_star) with shape [len(a) for a in axes.values()] (+ [len(h0_axis)] last)."""
    sp = engine.spec
    if sp.col_offset < 0:
        raise ValueError("the spec has no offset column")
    cols = sorted(axes)
    fixed = {sp.col_offset: 0.0}
    if h0_axis is not None:
        if sp.col_H0 < 0 or sp.col_H0 in axes:
            raise ValueError("h0_axis needs a free H0 column that is not among `axes`")
        if sp.bao_z is not None or sp.cmb_mode != 0 or sp.cc_z is not None or sp.family != 0:
            raise ValueError("the H0 axis is analytic only for SN-only late-time models")
        if sp.sn_mu_fixed is not None:
            # SH0ES calibrator rows carry a fixed distance modulus (sn/pantheon_and_sh0es.py:63-69) and do not move with H0
            raise ValueError("the H0 axis is not a pure offset when the spec has fixed-distance (calibrator) rows")
        fixed[sp.col_H0] = h0_ref
    if sorted(list(axes) + list(fixed)) != list(range(sp.ndim)):
        raise ValueError("axes must cover every theta column except the offset (and H0 with h0_axis)")
    shape = [len(axes[c]) for c in cols]
    n = int(np.prod(shape))
    mom = np.empty((n, 3))
    ax = [np.asarray(axes[c], dtype=np.float64) for c in cols]
    done = 0
    while done < n:   # arbitrary (non-linspace) axes: theta rows of a chunk by index arithmetic; np.linspace grids go through cl_eval_grid
        m = min(chunk, n - done)
        theta = np.empty((m, sp.ndim))
        for c, v in fixed.items():
            theta[:, c] = v
        idx = np.unravel_index(np.arange(done, done + m), shape)    # C order: the last axis runs fastest (itertools.product order)
        for k, c in enumerate(cols):
            theta[:, c] = ax[k][idx[k]]
        mom[done:done + m] = engine.sn_moments(theta)
        done += m
    mom = mom.reshape(shape + [3])
    if h0_axis is None:
        return offset_profile(mom, mode)
    # H0 only shifts the effective offset: delta(H0) = delta(h0_ref) + 5 log10(H0 / h0_ref)
    shift = 5.0 * np.log10(np.asarray(h0_axis, dtype=np.float64) / h0_ref)
    chi2, mstar = offset_profile(mom, mode)
    chi2 = np.broadcast_to(chi2[..., None], shape + [len(shift)]).copy()  # flat in H0 once M is profiled / marginalised
    return chi2, mstar[..., None] + shift
