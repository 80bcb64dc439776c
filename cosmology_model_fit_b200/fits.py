"""Model specs for the reference's fit scripts (one function per script, same name as the script).

Each function takes the tuples the reference's own loaders return (`yXXXX*/data.py: get_data()`) and does
what the script's import-time section does — Cholesky/inverse of the covariance, the 4000-point z-grid, the
step-template weights — then describes the model as a `LikelihoodSpec` whose theta column order is the
script's own.

SN data tuple  : (z_cmb, z_hel, obs, cov)         e.g. y2022pantheonSHOES/data.py:28-35 without the legend
BAO data tuple : (z, value, quantity, cov)        quantity = strings or CL_BAO_* codes (y2025BAO/data.py)
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import block_diag, cho_factor

from . import spec as S
from .spec import LikelihoodSpec

BBN_SCHONEBERG = (0.02218, 0.00055)  # y2024BBN/prior_lcdm_schoneberg.py:2-3
H0_TRGB = (70.39, 1.80)              # sn/pantheon.py:85
DES_Y6_BAO = (np.array([0.85]), np.array([19.74]), np.array([S.BAO_DM]), np.array([[0.60**2]]))  # y2024DESBAO/data.py
SIXDF_BAO = (np.array([0.106]), np.array([2.9761904762]), np.array([S.BAO_DV]), np.array([[0.0176517796]]))  # y20116dFBAO/data.py


def _qty_codes(q):
    q = np.asarray(q)
    if q.dtype.kind in "US":
        return np.array([S.BAO_QTY_MAP[str(x)] for x in q], dtype=np.int32)
    return q.astype(np.int32)


def concat_bao(*sets):
    """np.concatenate of the tables + block_diag of the covariances (bao/desi_cmb_union3.py:20-21)."""
    z = np.concatenate([np.asarray(s[0], dtype=np.float64) for s in sets])
    v = np.concatenate([np.asarray(s[1], dtype=np.float64) for s in sets])
    q = np.concatenate([_qty_codes(s[2]) for s in sets])
    cov = block_diag(*[np.asarray(s[3], dtype=np.float64) for s in sets])
    return z, v, q, cov


def _cho_lower(cov):
    """cho_factor(cov, lower=True)[0] with the (garbage) upper triangle cleared (sn/pantheon.py:14)."""
    return np.tril(cho_factor(np.asarray(cov, dtype=np.float64), lower=True)[0])


def _grid(*zs):
    return LikelihoodSpec.make_grid(max(float(np.max(z)) for z in zs))


def _sn_block(sp, sn, form, z_turn, col_offset, col_v):
    z_cmb, z_hel, obs, cov = sn
    sp.sn_zcmb, sp.sn_zhel, sp.sn_obs = z_cmb, z_hel, obs
    sp.sn_cov_form = form
    sp.sn_mat = _cho_lower(cov) if form == S.SN_CHOLESKY else np.linalg.inv(cov)
    sp.col_offset = col_offset
    if col_v is not None:
        sp.col_vel = (col_v,)
        sp.sn_vel_weight = LikelihoodSpec.step_weight(z_cmb, z_turn)
    return sp


def _bao_block(sp, bao, dh_mode, rd_mode, rd_fixed=147.09, col_rd=-1):
    z, v, q, cov = bao
    sp.bao_z, sp.bao_value, sp.bao_qty = z, v, _qty_codes(q)
    sp.bao_inv_cov = np.linalg.inv(cov)
    sp.bao_dh_mode, sp.rd_mode, sp.rd_fixed, sp.col_rd = dh_mode, rd_mode, rd_fixed, col_rd
    return sp


# ---------------------------------------------------------------------------------------------- sn/*
def sn_pantheon(sn, z_turn=0.15):
    """sn/pantheon.py (config 0): theta = (M, H0, Om, v); late LCDM; Cholesky chi2; TRGB H0 prior."""
    bounds = np.array([(-20.0, -19.0), (50.0, 90.0), (0.0, 0.7), (-3.0, 3.0)])  # sn/pantheon.py:68-75
    sp = LikelihoodSpec(ndim=4, family=S.FAMILY_LATE, de_model=S.DE_LCDM, col_H0=1, col_Om=2,
                        z_grid=_grid(sn[0]), bounds=bounds, gauss_prior=((1, *H0_TRGB),))
    return _sn_block(sp, sn, S.SN_CHOLESKY, z_turn, 0, 3)


def sn_pantheon_and_sh0es(sn_shoes, z_turn=0.15):
    """sn/pantheon_and_sh0es.py: theta = (M, H0, Om, v); mu of the Cepheid calibrators is their Cepheid distance (:63-69);
    the z_turn step applies to non-calibrators only (:47)."""
    z_cmb, z_hel, mb, ceph, cov = sn_shoes
    ceph_mask = np.asarray(ceph) != -9
    bounds = np.array([(-20.0, -18.5), (60.0, 85.0), (0.1, 0.6), (-3.5, 3.5)])  # :77-84
    sp = LikelihoodSpec(ndim=4, family=S.FAMILY_LATE, de_model=S.DE_LCDM, col_H0=1, col_Om=2, z_grid=_grid(z_cmb), bounds=bounds)
    _sn_block(sp, (z_cmb, z_hel, mb, cov), S.SN_CHOLESKY, z_turn, 0, 3)
    sp.sn_vel_weight = np.where((np.asarray(z_cmb) <= z_turn) & ~ceph_mask, 1.0, -1.0)
    sp.sn_mu_fixed = np.where(ceph_mask, np.asarray(ceph, dtype=np.float64), np.nan)
    return sp


def sn_des5y(sn, z_turn=0.11):
    """sn/des5y.py: theta = (dM, H0, Om, v); late LCDM; step at z_cmb <= 0.11 (sn/des5y.py:44-45)."""
    sp = LikelihoodSpec(ndim=4, family=S.FAMILY_LATE, de_model=S.DE_LCDM, col_H0=1, col_Om=2, z_grid=_grid(sn[0]))
    return _sn_block(sp, sn, S.SN_CHOLESKY, z_turn, 0, 3)


def sn_union3_1(sn, z_turn=0.2):
    """sn/union3_1.py: theta = (dM, Om, v); H0 fixed at 70; chi2 = d @ inv_cov @ d."""
    sp = LikelihoodSpec(ndim=3, family=S.FAMILY_LATE, de_model=S.DE_LCDM, col_H0=-1, H0_fixed=70.0, col_Om=1,
                        z_grid=_grid(sn[0]))
    return _sn_block(sp, sn, S.SN_INVCOV, z_turn, 0, 2)


def _sn_cmb(sn, form, z_turn, consts, bounds=None):
    consts = consts or S.cmb_planck_act()
    sp = LikelihoodSpec(ndim=5, family=S.FAMILY_FULL, de_model=S.DE_LCDM, col_H0=1, col_obh2=2, col_och2=3,
                        cmb_consts=consts, cmb_mode=consts.mode, z_grid=_grid(sn[0]), bounds=bounds)
    return _sn_block(sp, sn, form, z_turn, 0, 4)


def sn_pantheon_cmb(sn, consts=None):
    """sn/pantheon_cmb.py: theta = (M, H0, obh2, och2, v); full LCDM; Pantheon+ (Cholesky) + compressed CMB 3x3; box prior."""
    bounds = np.array([(-20.0, -19.0), (60.0, 75.0), (0.010, 0.030), (0.010, 0.25), (-2.5, 2.5)])  # sn/pantheon_cmb.py:86-94
    return _sn_cmb(sn, S.SN_CHOLESKY, 0.15, consts, bounds)


def sn_des5y_cmb(sn, consts=None):
    """sn/des5y_cmb.py: theta = (dM, H0, obh2, och2, v); DES-Dovekie (Cholesky, step at z_cmb <= 0.11) + CMB 3x3."""
    return _sn_cmb(sn, S.SN_CHOLESKY, 0.11, consts)


def sn_union3_1_cmb(sn, consts=None):
    """sn/union3_1_cmb.py: theta = (dM, H0, obh2, och2, v); Union3.1 (d @ inv_cov @ d, step at 0.2) + CMB 3x3."""
    return _sn_cmb(sn, S.SN_INVCOV, 0.2, consts)


# --------------------------------------------------------------------------------------------- cmb/*
def cmb_cmb(consts=None):
    """cmb/cmb.py: theta = (H0, obh2, och2); compressed CMB only."""
    consts = consts or S.cmb_planck_act()
    bounds = np.array([(60.0, 75.0), (0.020, 0.025), (0.05, 0.25)])  # cmb/cmb.py:30-36
    return LikelihoodSpec(ndim=3, family=S.FAMILY_FULL, de_model=S.DE_LCDM, col_H0=0, col_obh2=1, col_och2=2,
                          cmb_consts=consts, cmb_mode=consts.mode, bounds=bounds,
                          z_grid=LikelihoodSpec.make_grid(2.33))


# --------------------------------------------------------------------------------------------- bao/*
def bao_desi(desi, des_y6=DES_Y6_BAO):
    """bao/desi.py: theta = (h, Om, w0); thawing; r_d = 147.09 fixed; pchip D_H; emcee vectorize=True."""
    bao = concat_bao(desi, des_y6)
    bounds = np.array([(0.50, 0.80), (0.1, 0.5), (-1.0, 0.0)])  # bao/desi.py:67-73
    sp = LikelihoodSpec(ndim=3, family=S.FAMILY_LATE, de_model=S.DE_THAWING, col_H0=0, H0_scale=100.0, col_Om=1,
                        col_w0=2, z_grid=_grid(bao[0]), bounds=bounds)
    return _bao_block(sp, bao, S.DH_PCHIP, S.RD_FIXED, 147.09)


def bao_desi_cmb(desi, consts=None):
    """bao/desi_cmb.py: theta = (H0, obh2, och2, w0); thawing; early-LCDM compressed CMB (theta*, omega_b, omega_m);
    exact D_H; emcee vectorize=True."""
    consts = consts or S.cmb_early_lcdm()
    bounds = np.array([(50.0, 80.0), (0.020, 0.024), (0.05, 0.30), (-1.0, 0.0)])  # bao/desi_cmb.py:105-112
    sp = LikelihoodSpec(ndim=4, family=S.FAMILY_FULL, de_model=S.DE_THAWING, col_H0=0, col_obh2=1, col_och2=2, col_w0=3,
                        cmb_consts=consts, cmb_mode=consts.mode, z_grid=_grid(desi[0]), bounds=bounds)
    return _bao_block(sp, desi, S.DH_EXACT, S.RD_FIT)


def bao_desi_union3_obh2_theta_star(sn, desi, consts=None):
    """bao/desi_union3_obh2_theta_star.py: theta = (dM, H0, obh2, och2, v); LCDM; CMB rows [1:] (l_A, omega_b) weighted by
    the inverse of the 2x2 sub-covariance (:17,124-128)."""
    consts = consts or S.cmb_planck_act()
    w = np.zeros((3, 3))
    w[1:, 1:] = np.linalg.inv(consts.covariance[1:, 1:])
    sp = LikelihoodSpec(ndim=5, family=S.FAMILY_FULL, de_model=S.DE_LCDM, col_H0=1, col_obh2=2, col_och2=3,
                        cmb_consts=consts, cmb_mode=S.CMB_R_LA_WB, cmb_weight=w, z_grid=_grid(sn[0], desi[0]))
    _sn_block(sp, sn, S.SN_INVCOV, 0.2, 0, 4)
    return _bao_block(sp, desi, S.DH_EXACT, S.RD_FIT)


def bao_desi_cmb_union3(sn, desi_fs_lya, consts=None, des_y6=DES_Y6_BAO, sixdf=SIXDF_BAO, de_model=S.DE_LCDM):
    """bao/desi_cmb_union3.py (config 2): theta = (dM, H0, obh2, och2, v [, w0, wa])."""
    consts = consts or S.cmb_planck_act()
    bao = concat_bao(desi_fs_lya, des_y6, sixdf)
    ndim = 5 + (2 if de_model == S.DE_CPL else 1 if de_model in (S.DE_WCDM, S.DE_THAWING) else 0)
    sp = LikelihoodSpec(ndim=ndim, family=S.FAMILY_FULL, de_model=de_model, col_H0=1, col_obh2=2, col_och2=3,
                        col_w0=5 if ndim > 5 else -1, col_wa=6 if ndim > 6 else -1, guard_cpl=de_model == S.DE_CPL,
                        cmb_consts=consts, cmb_mode=consts.mode, z_grid=_grid(sn[0], bao[0]))
    _sn_block(sp, sn, S.SN_INVCOV, 0.2, 0, 4)
    return _bao_block(sp, bao, S.DH_PCHIP, S.RD_FIT)


def bao_desi_fs_lya_cmb(desi_fs_lya, consts=None):
    """bao/desi_fs_lya_cmb.py: theta = (H0, obh2, och2, w0, wa); CPL with the w0+wa>=0 -> -1e8 guard."""
    consts = consts or S.cmb_planck_act()
    sp = LikelihoodSpec(ndim=5, family=S.FAMILY_FULL, de_model=S.DE_CPL, col_H0=0, col_obh2=1, col_och2=2,
                        col_w0=3, col_wa=4, cmb_consts=consts, cmb_mode=consts.mode, guard_cpl=True,
                        z_grid=_grid(desi_fs_lya[0]))
    return _bao_block(sp, (desi_fs_lya[0], desi_fs_lya[1], _qty_codes(desi_fs_lya[2]), desi_fs_lya[3]), S.DH_PCHIP, S.RD_FIT)


def bao_desi_des5y_bbn_theta_star(sn, desi, consts=None, with_v=False, z_turn=0.10563):
    """bao/desi_des5y_bbn_theta_star.py (config 1): theta = (dM, H0, obh2, och2, w0 [, v]); thawing; l_A-only CMB
    term d^2/cov[1,1]; BBN Gaussian in the prior.  `with_v=True` adds the z_turn=0.10563 step of the sibling DES
    scripts (bao/desi_cmb_des5y.py:103-111) as BASELINE.json's config text asks."""
    consts = consts or S.cmb_planck_act()
    bounds = [(-0.5, 0.5), (50.0, 90.0), (0.010, 0.030), (0.05, 0.30), (-1.0, -1 / 3)]  # :122-130
    if with_v:
        bounds.append((-4.5, 4.5))
    w = np.zeros((3, 3))
    w[1, 1] = 1.0 / consts.covariance[1, 1]  # :110-111
    sp = LikelihoodSpec(ndim=len(bounds), family=S.FAMILY_FULL, de_model=S.DE_THAWING, col_H0=1, col_obh2=2, col_och2=3,
                        col_w0=4, cmb_consts=consts, cmb_mode=S.CMB_R_LA_WB, cmb_weight=w,
                        z_grid=_grid(sn[0], desi[0]), bounds=np.array(bounds), gauss_prior=((2, *BBN_SCHONEBERG),))
    _sn_block(sp, sn, S.SN_CHOLESKY, z_turn, 0, 5 if with_v else None)
    return _bao_block(sp, desi, S.DH_EXACT, S.RD_FIT)


def _la_only(consts):
    """chi2 = (l_A - l_A_model)^2 / cov[1, 1] (bao/desi_bbn_theta_star.py:96-99): a 3x3 weight with one non-zero entry."""
    w = np.zeros((3, 3))
    w[1, 1] = 1.0 / consts.covariance[1, 1]
    return w


def bao_desi_bbn_theta_star(desi, consts=None):
    """bao/desi_bbn_theta_star.py: theta = (H0, obh2, och2, w0); full thawing; Planck (PR3) module, l_A-only term; exact D_H;
    BBN Gaussian on omega_b through the nautilus prior (:118)."""
    consts = consts or S.cmb_planck()
    sp = LikelihoodSpec(ndim=4, family=S.FAMILY_FULL, de_model=S.DE_THAWING, col_H0=0, col_obh2=1, col_och2=2, col_w0=3,
                        cmb_consts=consts, cmb_mode=S.CMB_R_LA_WB, cmb_weight=_la_only(consts), z_grid=_grid(desi[0]),
                        gauss_prior=((1, *BBN_SCHONEBERG),))
    return _bao_block(sp, desi, S.DH_EXACT, S.RD_FIT)


def bao_desi_union3_bbn_theta_star(sn, desi_fs_lya, consts=None, des_y6=DES_Y6_BAO):
    """bao/desi_union3_bbn_theta_star.py: theta = (dM, H0, obh2, och2, v); full LCDM; Union3.1 (inv_cov, step at 0.2);
    DESI FS-Lya + DES-Y6 BAO (15 points, F_AP rows, exact D_H); l_A-only term; BBN prior on omega_b (:168)."""
    consts = consts or S.cmb_planck_act()
    bao = concat_bao(desi_fs_lya, des_y6)
    sp = LikelihoodSpec(ndim=5, family=S.FAMILY_FULL, de_model=S.DE_LCDM, col_H0=1, col_obh2=2, col_och2=3,
                        cmb_consts=consts, cmb_mode=S.CMB_R_LA_WB, cmb_weight=_la_only(consts), z_grid=_grid(sn[0], bao[0]),
                        gauss_prior=((2, *BBN_SCHONEBERG),))
    _sn_block(sp, sn, S.SN_INVCOV, 0.2, 0, 4)
    return _bao_block(sp, bao, S.DH_EXACT, S.RD_FIT)


def bao_desi_union3_cc_theta_star(sn, desi, cc, consts=None):
    """bao/desi_union3_cc_theta_star.py: theta = (f_cc, dM, H0, obh2, och2, v); full LCDM; Union3.1 + DESI (exact D_H) + l_A
    + CC (f^2 chi2 with normalisation); nautilus vectorized=True, batch log_likelihood returns float32 (:142-147)."""
    consts = consts or S.cmb_planck_act()
    sp = LikelihoodSpec(ndim=6, family=S.FAMILY_FULL, de_model=S.DE_LCDM, col_H0=2, col_obh2=3, col_och2=4,
                        cmb_consts=consts, cmb_mode=S.CMB_R_LA_WB, cmb_weight=_la_only(consts), z_grid=_grid(sn[0], desi[0]))
    _sn_block(sp, sn, S.SN_INVCOV, 0.2, 1, 5)
    _cc_block(sp, cc, 0)
    return _bao_block(sp, desi, S.DH_EXACT, S.RD_FIT)


def bao_desi_des5y_cc_theta_star(sn, desi, cc, consts=None):
    """bao/desi_des5y_cc_theta_star.py: theta = (f_cc, dM, H0, obh2, och2, w0); full thawing; DES-Dovekie (Cholesky, no
    velocity term) + DESI (exact D_H) + l_A + CC; box prior (:133-151)."""
    consts = consts or S.cmb_planck_act()
    bounds = np.array([(0.5, 2.5), (-0.60, 0.60), (50.0, 85.0), (0.005, 0.035), (0.05, 0.30), (-1.0, -1 / 3)])
    sp = LikelihoodSpec(ndim=6, family=S.FAMILY_FULL, de_model=S.DE_THAWING, col_H0=2, col_obh2=3, col_och2=4, col_w0=5,
                        cmb_consts=consts, cmb_mode=S.CMB_R_LA_WB, cmb_weight=_la_only(consts), z_grid=_grid(sn[0], desi[0]),
                        bounds=bounds)
    _sn_block(sp, sn, S.SN_CHOLESKY, 0.0, 1, None)
    _cc_block(sp, cc, 0)
    return _bao_block(sp, desi, S.DH_EXACT, S.RD_FIT)


def bao_desi_cmb_pantheon(sn, desi, consts=None, z_turn=0.15):
    """bao/desi_cmb_pantheon.py (config 3): theta = (M, H0, obh2, och2, v); LCDM as checked in."""
    consts = consts or S.cmb_planck_act()
    sp = LikelihoodSpec(ndim=5, family=S.FAMILY_FULL, de_model=S.DE_LCDM, col_H0=1, col_obh2=2, col_och2=3,
                        cmb_consts=consts, cmb_mode=consts.mode, z_grid=_grid(sn[0], desi[0]))
    _sn_block(sp, sn, S.SN_CHOLESKY, z_turn, 0, 4)
    return _bao_block(sp, desi, S.DH_EXACT, S.RD_FIT)


def bao_desi_cmb_pantheon_H0trgb(sn, desi, consts=None):
    """bao/desi_cmb_pantheon_H0trgb.py: theta = (M, H0, obh2, och2, v_flow); the flow enters as a linear magnitude
    template 100 v (5/ln 10)/(c z) (:103-106); TRGB H0 term inside chi2 (:124)."""
    consts = consts or S.cmb_planck_act()
    sp = LikelihoodSpec(ndim=5, family=S.FAMILY_FULL, de_model=S.DE_LCDM, col_H0=1, col_obh2=2, col_och2=3,
                        cmb_consts=consts, cmb_mode=consts.mode, z_grid=_grid(sn[0], desi[0]), gauss_chi2=((1, *H0_TRGB),))
    _sn_block(sp, sn, S.SN_CHOLESKY, 0.0, 0, None)
    sp.col_lin = (4,)
    sp.sn_lin_template = 100 * (5 / np.log(10)) / (S.C_KMS * np.asarray(sn[0], dtype=np.float64))
    return _bao_block(sp, desi, S.DH_EXACT, S.RD_FIT)


def bao_desi_cmb_des5y(sn, desi_fs_lya, consts=None, z_turn=0.10563):
    """bao/desi_cmb_des5y.py: theta = (dM, H0, obh2, och2, v); pchip D_H; F_AP rows; step at 0.10563."""
    consts = consts or S.cmb_planck_act()
    sp = LikelihoodSpec(ndim=5, family=S.FAMILY_FULL, de_model=S.DE_LCDM, col_H0=1, col_obh2=2, col_och2=3,
                        cmb_consts=consts, cmb_mode=consts.mode, z_grid=_grid(sn[0], desi_fs_lya[0]))
    _sn_block(sp, sn, S.SN_CHOLESKY, z_turn, 0, 4)
    return _bao_block(sp, desi_fs_lya, S.DH_PCHIP, S.RD_FIT)


def bao_desi_bbn(desi, consts=None):
    """bao/desi_bbn.py: theta = (H0, Om, obh2, w0); late thawing; r_d = r_drag(obh2, Om h^2); pchip D_H."""
    consts = consts or S.cmb_planck()
    bounds = np.array([(55.0, 75.0), (0.17, 0.50), (0.016, 0.030), (-1.0, -1 / 3)])  # bao/desi_bbn.py:70-77
    sp = LikelihoodSpec(ndim=4, family=S.FAMILY_LATE, de_model=S.DE_THAWING, col_H0=0, col_Om=1, col_obh2=2, col_w0=3,
                        cmb_consts=consts, z_grid=_grid(desi[0]), bounds=bounds)
    return _bao_block(sp, desi, S.DH_PCHIP, S.RD_FIT)


def bao_desi_pantheon_cc(sn, desi, cc):
    """bao/desi_pantheon_cc.py: theta = (H0, M, r_d, Om, v, f_cc); late LCDM; uniform flow with the multiplicative shift
    z_cosmo = max((1+z)(1+v/c) - 1, 1e-8) (:83-90); r_d sampled; CC term with its normalisation (:134-137)."""
    bounds = np.array([(40.0, 90.0), (-20.0, -19.0), (115.0, 170.0), (0.0, 1.0), (-1.3, 3.5), (0.4, 2.5)])  # :94-103
    zc, Hc, covc = cc
    sp = LikelihoodSpec(ndim=6, family=S.FAMILY_LATE, de_model=S.DE_LCDM, col_H0=0, col_Om=3, z_grid=_grid(sn[0], desi[0]),
                        bounds=bounds, cc_z=zc, cc_H=Hc, cc_inv_cov=np.linalg.inv(covc), col_fcc=5,
                        cc_logdet=float(np.linalg.slogdet(covc)[1]), cc_norm_sign=1.0)
    _sn_block(sp, sn, S.SN_CHOLESKY, 0.0, 1, 4)
    sp.sn_vel_weight = np.ones_like(np.asarray(sn[0], dtype=np.float64))
    sp.vel_mode = S.VEL_MULTIPLY
    return _bao_block(sp, desi, S.DH_EXACT, S.RD_PARAM, col_rd=2)


def sn_pantheon_dipole_xyz(sn, ra, dec, survey_id, z_c=0.10, dz=0.02, target_ids=(1, 5, 15, 50, 51, 56, 63, 150)):
    """sn/pantheon_dipole_xyz.py: theta = (M, H0, Om, vx, vy, vz); velocity templates n_k * attenuation * survey mask
    (:14-21, 47-55)."""
    z_cmb = np.asarray(sn[0], dtype=np.float64)
    ra_rad, dec_rad = np.deg2rad(ra), np.deg2rad(dec)
    n = np.vstack([np.cos(dec_rad) * np.cos(ra_rad), np.cos(dec_rad) * np.sin(ra_rad), np.sin(dec_rad)])
    att = 0.5 * (1.0 - np.tanh((z_cmb - z_c) / dz))
    mask = np.isin(survey_id, target_ids).astype(int)
    sp = LikelihoodSpec(ndim=6, family=S.FAMILY_LATE, de_model=S.DE_LCDM, col_H0=1, col_Om=2, z_grid=_grid(sn[0]))
    _sn_block(sp, sn, S.SN_CHOLESKY, 0.0, 0, None)
    sp.col_vel = (3, 4, 5)
    sp.sn_vel_weight = n * att * mask
    return sp


def sn_pantheon_dipole(sn, ra, dec, survey_id, ra_fixed_deg=217, dec_fixed_deg=-29, z_c=0.10, dz=0.02,
                       target_ids=(1, 5, 15, 50, 51, 56, 63, 150)):
    """sn/pantheon_dipole.py: theta = (M, H0, Om, v); dipole of fixed direction: weight = cos(angle) * attenuation * mask
    (:14-32, 60-68)."""
    z_cmb = np.asarray(sn[0], dtype=np.float64)
    ra_rad, dec_rad = np.deg2rad(ra), np.deg2rad(dec)
    nx, ny, nz = np.cos(dec_rad) * np.cos(ra_rad), np.cos(dec_rad) * np.sin(ra_rad), np.sin(dec_rad)
    raf, decf = np.deg2rad(ra_fixed_deg), np.deg2rad(dec_fixed_deg)
    cos_angle = nx * (np.cos(decf) * np.cos(raf)) + ny * (np.cos(decf) * np.sin(raf)) + nz * np.sin(decf)
    att = 0.5 * (1.0 - np.tanh((z_cmb - z_c) / dz))
    mask = np.isin(survey_id, target_ids).astype(int)
    sp = LikelihoodSpec(ndim=4, family=S.FAMILY_LATE, de_model=S.DE_LCDM, col_H0=1, col_Om=2, z_grid=_grid(sn[0]))
    _sn_block(sp, sn, S.SN_CHOLESKY, 0.0, 0, None)
    sp.col_vel = (3,)
    sp.sn_vel_weight = cos_angle * att * mask
    return sp


def bao_desi_fs_lya(desi_fs_lya):
    """bao/desi_fs_lya.py: theta = (h, Om, w0); late thawing; DESI DR2 full-shape + Lya table (F_AP rows); r_d = 147.09; pchip D_H."""
    sp = LikelihoodSpec(ndim=3, family=S.FAMILY_LATE, de_model=S.DE_THAWING, col_H0=0, H0_scale=100.0, col_Om=1, col_w0=2,
                        z_grid=_grid(desi_fs_lya[0]))
    return _bao_block(sp, desi_fs_lya, S.DH_PCHIP, S.RD_FIXED, 147.09)


def bao_desi_cc(desi, cc):
    """bao/desi_cc.py: theta = (f_cc, H0, r_d, Om, w0); late thawing; exact D_H; sampled r_d; CC with f^2 chi2 + normalisation."""
    bounds = np.array([(0.5, 2.5), (45.0, 90.0), (120.0, 175.0), (0.1, 0.7), (-1.0, 0.0)])  # bao/desi_cc.py:106-114
    sp = LikelihoodSpec(ndim=5, family=S.FAMILY_LATE, de_model=S.DE_THAWING, col_H0=1, col_Om=3, col_w0=4,
                        z_grid=_grid(desi[0]), bounds=bounds)
    _cc_block(sp, cc, 0)
    return _bao_block(sp, desi, S.DH_EXACT, S.RD_PARAM, col_rd=2)


def _sn_bao_rd(sn, desi, form, z_turn):
    sp = LikelihoodSpec(ndim=5, family=S.FAMILY_LATE, de_model=S.DE_LCDM, col_H0=2, col_Om=3, z_grid=_grid(sn[0], desi[0]),
                        gauss_prior=((1, 147.09, 0.26),))  # nautilus prior norm(147.09, 0.26) on r_d (bao/desi_des5y_rd.py:125)
    _sn_block(sp, sn, form, z_turn, 0, 4)
    return _bao_block(sp, desi, S.DH_PCHIP, S.RD_PARAM, col_rd=1)


def bao_desi_des5y_rd(sn, desi):
    """bao/desi_des5y_rd.py: theta = (dM, r_d, H0, Om, v); late LCDM; DES-Dovekie (Cholesky, step at 0.10563); pchip D_H."""
    return _sn_bao_rd(sn, desi, S.SN_CHOLESKY, 0.10563)


def bao_desi_union3_rd(sn, desi):
    """bao/desi_union3_rd.py: theta = (dM, r_d, H0, Om, v); late LCDM; Union3.1 (inv_cov, step at 0.2); pchip D_H."""
    return _sn_bao_rd(sn, desi, S.SN_INVCOV, 0.2)


def bao_desi_pantheon_rd(sn, desi):
    """bao/desi_pantheon_rd.py: theta = (M, H0, Om, r_d, w0); late thawing; Pantheon+ (Cholesky, no velocity term); exact D_H;
    box prior + Gaussian r_d prior (147.14, 0.29) (:109-113)."""
    bounds = np.array([(-20.0, -19.0), (50.0, 100.0), (0.2, 0.7), (144.0, 150.0), (-1.0, -1 / 3)])  # :76-84
    sp = LikelihoodSpec(ndim=5, family=S.FAMILY_LATE, de_model=S.DE_THAWING, col_H0=1, col_Om=2, col_w0=4,
                        z_grid=_grid(sn[0], desi[0]), bounds=bounds, gauss_prior=((3, 147.14, 0.29),))
    _sn_block(sp, sn, S.SN_CHOLESKY, 0.0, 0, None)
    return _bao_block(sp, desi, S.DH_EXACT, S.RD_PARAM, col_rd=3)


def _sn_bao_omh2(sn, desi, form, z_turn):
    sp = LikelihoodSpec(ndim=5, family=S.FAMILY_LATE, de_model=S.DE_LCDM, col_H0=2, col_Om=3, Om_is_physical=True,
                        z_grid=_grid(sn[0], desi[0]), gauss_prior=((3, 0.1430, 0.0011),))  # nautilus prior on omega_m (:117 / :132)
    _sn_block(sp, sn, form, z_turn, 0, 4)
    return _bao_block(sp, desi, S.DH_EXACT, S.RD_PARAM, col_rd=1)


def bao_desi_union3_omh2(sn, desi):
    """bao/desi_union3_omh2.py: theta = (dM, r_d, H0, omega_m, v); late LCDM with Om = omega_m / h^2; Union3.1 (step at 0.2)."""
    return _sn_bao_omh2(sn, desi, S.SN_INVCOV, 0.2)


def bao_desi_des5y_omh2(sn, desi):
    """bao/desi_des5y_omh2.py: theta = (dM, r_d, H0, omega_m, v); late LCDM with Om = omega_m / h^2; DES-Dovekie (step at 0.10563)."""
    return _sn_bao_omh2(sn, desi, S.SN_CHOLESKY, 0.10563)


def _sub_weight(matrix, rows, invert):
    """3x3 CMB weight from a row/column sub-selection: the inverse of the sub-covariance (invert=True; bao/
    desi_union3_omh2_theta_star.py:17) or the sub-block of the full inverse (bao/desi_des5y_obh2_theta_star.py:104)."""
    w = np.zeros((3, 3))
    sub = np.asarray(matrix)[np.ix_(rows, rows)]
    w[np.ix_(rows, rows)] = np.linalg.inv(sub) if invert else sub
    return w


def bao_desi_union3_omh2_theta_star(sn, desi, consts=None):
    """bao/desi_union3_omh2_theta_star.py: theta = (dM, H0, obh2, och2, v); full LCDM; early-LCDM compression rows
    (theta*, omega_m) with the inverse of their 2x2 sub-covariance; Union3.1 (step at 0.2); exact D_H."""
    consts = consts or S.cmb_early_lcdm()
    sp = LikelihoodSpec(ndim=5, family=S.FAMILY_FULL, de_model=S.DE_LCDM, col_H0=1, col_obh2=2, col_och2=3,
                        cmb_consts=consts, cmb_mode=consts.mode, cmb_weight=_sub_weight(consts.covariance, [0, 2], True),
                        z_grid=_grid(sn[0], desi[0]))
    _sn_block(sp, sn, S.SN_INVCOV, 0.2, 0, 4)
    return _bao_block(sp, desi, S.DH_EXACT, S.RD_FIT)


def bao_desi_pantheon_obh2_theta_star(sn, desi, consts=None):
    """bao/desi_pantheon_obh2_theta_star.py: theta = (M, H0, obh2, och2, w0); full thawing; early-LCDM rows (theta*, omega_b);
    Pantheon+ (Cholesky, no velocity term); exact D_H; box prior."""
    consts = consts or S.cmb_early_lcdm()
    bounds = np.array([(-20.0, -19.0), (50.0, 90.0), (0.0, 0.05), (0.05, 0.30), (-1.0, -1 / 3)])  # :108-116
    sp = LikelihoodSpec(ndim=5, family=S.FAMILY_FULL, de_model=S.DE_THAWING, col_H0=1, col_obh2=2, col_och2=3, col_w0=4,
                        cmb_consts=consts, cmb_mode=consts.mode, cmb_weight=_sub_weight(consts.covariance, [0, 1], True),
                        z_grid=_grid(sn[0], desi[0]), bounds=bounds)
    _sn_block(sp, sn, S.SN_CHOLESKY, 0.0, 0, None)
    return _bao_block(sp, desi, S.DH_EXACT, S.RD_FIT)


def bao_desi_des5y_obh2_theta_star(sn, desi, consts=None):
    """bao/desi_des5y_obh2_theta_star.py: theta = (dM, H0, obh2, och2, w0); full thawing; Planck+ACT rows (l_A, omega_b)
    weighted by the SUB-BLOCK of the full inverse covariance (:104); DES-Dovekie (Cholesky, no velocity term); box prior."""
    consts = consts or S.cmb_planck_act()
    bounds = np.array([(-0.4, 0.4), (50.0, 90.0), (0.010, 0.030), (0.05, 0.30), (-1.0, -1 / 3)])  # :110-118
    sp = LikelihoodSpec(ndim=5, family=S.FAMILY_FULL, de_model=S.DE_THAWING, col_H0=1, col_obh2=2, col_och2=3, col_w0=4,
                        cmb_consts=consts, cmb_mode=S.CMB_R_LA_WB,
                        cmb_weight=_sub_weight(np.linalg.inv(consts.covariance), [1, 2], False),
                        z_grid=_grid(sn[0], desi[0]), bounds=bounds)
    _sn_block(sp, sn, S.SN_CHOLESKY, 0.0, 0, None)
    return _bao_block(sp, desi, S.DH_EXACT, S.RD_FIT)


def bao_desi_union3_bbn(sn, desi, des_y6=DES_Y6_BAO):
    """bao/desi_union3_bbn.py: theta = (H0, Om, obh2, v, dM); late LCDM; DESI + DES-Y6 BAO (14 points, exact D_H) with the
    r_drag fit written out in the script; Union3.1 (inv_cov, step at 0.2); BBN prior on omega_b (:154)."""
    bao = concat_bao(desi, des_y6)
    sp = LikelihoodSpec(ndim=5, family=S.FAMILY_LATE, de_model=S.DE_LCDM, col_H0=0, col_Om=1, col_obh2=2,
                        cmb_consts=S.cmb_rdrag_plain(), z_grid=_grid(sn[0], bao[0]), gauss_prior=((2, *BBN_SCHONEBERG),))
    _sn_block(sp, sn, S.SN_INVCOV, 0.2, 4, 3)
    return _bao_block(sp, bao, S.DH_EXACT, S.RD_FIT)


def bao_desi_des5y_bbn(sn, desi):
    """bao/desi_des5y_bbn.py: theta = (H0, Om, obh2, w0, dM); late thawing; DESI BAO through its Cholesky factor (:17; the
    same quadratic form as inv_cov), exact D_H, in-script r_drag fit; DES-Dovekie (Cholesky, no velocity term); BBN prior."""
    sp = LikelihoodSpec(ndim=5, family=S.FAMILY_LATE, de_model=S.DE_THAWING, col_H0=0, col_Om=1, col_obh2=2, col_w0=3,
                        cmb_consts=S.cmb_rdrag_plain(), z_grid=_grid(sn[0], desi[0]), gauss_prior=((2, *BBN_SCHONEBERG),))
    _sn_block(sp, sn, S.SN_CHOLESKY, 0.0, 4, None)
    return _bao_block(sp, desi, S.DH_EXACT, S.RD_FIT)


def bao_desi_des5y_H0trgb(sn, desi):
    """bao/desi_des5y_H0trgb.py: theta = (dM, H0, r_d, Om, w0); late thawing; sampled r_d; exact D_H; DES-Dovekie (Cholesky, no
    velocity term); box prior with the TRGB H0 Gaussian inside log_prior (:96-101)."""
    bounds = np.array([(-0.5, 0.5), (56.0, 85.0), (120.0, 160.0), (0.1, 0.7), (-1.0, -1 / 3)])  # :84-92
    sp = LikelihoodSpec(ndim=5, family=S.FAMILY_LATE, de_model=S.DE_THAWING, col_H0=1, col_Om=3, col_w0=4,
                        z_grid=_grid(sn[0], desi[0]), bounds=bounds, gauss_prior=((1, *H0_TRGB),))
    _sn_block(sp, sn, S.SN_CHOLESKY, 0.0, 0, None)
    return _bao_block(sp, desi, S.DH_EXACT, S.RD_PARAM, col_rd=2)


def _cmb_sn_bao_H0trgb(sn, desi, form, z_turn, consts, sixdf):
    consts = consts or S.cmb_planck_act()
    bao = concat_bao(desi, sixdf)   # the 6dF point has its own chi2 term in the script: block-diagonal covariance
    sp = LikelihoodSpec(ndim=5, family=S.FAMILY_FULL, de_model=S.DE_LCDM, col_H0=1, col_obh2=2, col_och2=3,
                        cmb_consts=consts, cmb_mode=consts.mode, z_grid=_grid(sn[0], desi[0]), gauss_chi2=((1, *H0_TRGB),))
    _sn_block(sp, sn, form, z_turn, 0, 4)
    return _bao_block(sp, bao, S.DH_EXACT, S.RD_FIT)


def bao_desi_cmb_des5y_H0trgb(sn, desi, consts=None, sixdf=SIXDF_BAO):
    """bao/desi_cmb_des5y_H0trgb.py: theta = (dM, H0, obh2, och2, v); full LCDM; DES-Dovekie (Cholesky, step at 0.11) + DESI +
    6dF (separate chi2 term) + CMB 3x3 + TRGB H0 chi2 term (:150)."""
    return _cmb_sn_bao_H0trgb(sn, desi, S.SN_CHOLESKY, 0.11, consts, sixdf)


def bao_desi_cmb_union3_H0trgb(sn, desi, consts=None, sixdf=SIXDF_BAO):
    """bao/desi_cmb_union3_H0trgb.py: theta = (dM, H0, obh2, och2, v); full LCDM; Union3.1 (inv_cov, step at 0.2) + DESI + 6dF +
    CMB 3x3 + TRGB H0 chi2 term."""
    return _cmb_sn_bao_H0trgb(sn, desi, S.SN_INVCOV, 0.2, consts, sixdf)


def bao_desi_des5y_cc(sn, desi_fs_lya, cc):
    """bao/desi_des5y_cc.py: theta = (f_cc, dM, H0, r_d, Om, v); late LCDM; DES-Dovekie (Cholesky, step at 0.10563) + DESI FS-Lya
    (exact D_H, F_AP rows) + CC through its Cholesky factor with f^2 and the normalisation; box prior."""
    bounds = np.array([(0.5, 2.5), (-0.55, 0.55), (50.0, 80.0), (110.0, 175.0), (0.2, 0.7), (-4.5, 4.5)])  # :108-117
    sp = LikelihoodSpec(ndim=6, family=S.FAMILY_LATE, de_model=S.DE_LCDM, col_H0=2, col_Om=4, z_grid=_grid(sn[0], desi_fs_lya[0]),
                        bounds=bounds)
    _sn_block(sp, sn, S.SN_CHOLESKY, 0.10563, 1, 5)
    _cc_block(sp, cc, 0)
    return _bao_block(sp, desi_fs_lya, S.DH_EXACT, S.RD_PARAM, col_rd=3)


def bao_desi_fs_lya_union3_cc(sn, desi_fs_lya, cc):
    """bao/desi_fs_lya_union3_cc.py: theta = (f_cc, dM, H0, r_d, Om, v); late LCDM; Union3.1 (inv_cov, step at 0.2) + DESI FS-Lya
    (exact D_H, F_AP rows) + CC (f^2 inv_cov) with the normalisation."""
    sp = LikelihoodSpec(ndim=6, family=S.FAMILY_LATE, de_model=S.DE_LCDM, col_H0=2, col_Om=4, z_grid=_grid(sn[0], desi_fs_lya[0]))
    _sn_block(sp, sn, S.SN_INVCOV, 0.2, 1, 5)
    _cc_block(sp, cc, 0)
    return _bao_block(sp, desi_fs_lya, S.DH_EXACT, S.RD_PARAM, col_rd=3)


def bao_desi_omh2(desi):
    """bao/desi_omh2.py: theta = (r_d, H0, omega_m, w0); late thawing with Om = omega_m / h^2; r_d sampled; exact D_H."""
    sp = LikelihoodSpec(ndim=4, family=S.FAMILY_LATE, de_model=S.DE_THAWING, col_H0=1, col_Om=2, Om_is_physical=True, col_w0=3,
                        z_grid=_grid(desi[0]), gauss_prior=((2, 0.1430, 0.0011),))
    return _bao_block(sp, desi, S.DH_EXACT, S.RD_PARAM, col_rd=0)


# --------------------------------------------------------------------------------------------- ohd/*
def _cc_block(sp, cc, col_fcc, norm_sign=1.0):
    z, H, cov = cc
    sp.cc_z, sp.cc_H, sp.cc_inv_cov, sp.col_fcc = z, H, np.linalg.inv(cov), col_fcc
    sp.cc_logdet, sp.cc_norm_sign = float(np.linalg.slogdet(cov)[1]), norm_sign
    return sp


def ohd_cc_des5y(sn, cc):
    """ohd/cc_des5y.py: theta = (f_cc, dM, H0, Om, w0); late wCDM; DES-Dovekie SN (Cholesky) + CC."""
    bounds = np.array([(0.2, 3), (-0.5, 0.5), (50, 85), (0.05, 0.6), (-1, -1 / 3)], dtype=np.float64)  # :53-62
    sp = LikelihoodSpec(ndim=5, family=S.FAMILY_LATE, de_model=S.DE_WCDM, col_H0=2, col_Om=3, col_w0=4, z_grid=_grid(sn[0]),
                        bounds=bounds)
    _sn_block(sp, sn, S.SN_CHOLESKY, 0.0, 1, None)
    return _cc_block(sp, cc, 0)


def ohd_cc_union3(sn, cc, z_turn=0.2):
    """ohd/cc_union3.py: theta = (f_cc, dM, H0, Om, v); v in km/s (no factor 100, :46); the grid ends at max(z_cmb) (:19)."""
    sp = LikelihoodSpec(ndim=5, family=S.FAMILY_LATE, de_model=S.DE_LCDM, col_H0=2, col_Om=3,
                        z_grid=np.linspace(0, np.max(sn[0]), num=4000), vel_scale=1.0)
    _sn_block(sp, sn, S.SN_INVCOV, z_turn, 1, 4)
    return _cc_block(sp, cc, 0)


def ohd_cc(cc):
    """ohd/cc.py: theta = (H0, Om, f); chi2 = f^2 d^T C^-1 d; log L adds N ln 2pi + logdet - 2N ln f."""
    z, H, cov = cc
    return LikelihoodSpec(ndim=3, family=S.FAMILY_LATE, de_model=S.DE_LCDM, col_H0=0, col_Om=1,
                          cc_z=z, cc_H=H, cc_inv_cov=np.linalg.inv(cov), col_fcc=2,
                          cc_logdet=float(np.linalg.slogdet(cov)[1]), cc_norm_sign=1.0,
                          z_grid=LikelihoodSpec.make_grid(float(np.max(z))))


def ohd_cc_cmb(cc, consts=None):
    """ohd/cc_cmb.py: theta = (H0, obh2, och2, f_cc); full LCDM; CC (f^2 chi2 + normalisation) + compressed CMB 3x3."""
    consts = consts or S.cmb_planck_act()
    bounds = np.array([(63.0, 73.0), (0.0210, 0.0235), (0.05, 0.30), (0.30, 2.75)])  # ohd/cc_cmb.py:38-45
    sp = LikelihoodSpec(ndim=4, family=S.FAMILY_FULL, de_model=S.DE_LCDM, col_H0=0, col_obh2=1, col_och2=2,
                        cmb_consts=consts, cmb_mode=consts.mode, bounds=bounds, log_prior_norm=0.0,
                        z_grid=LikelihoodSpec.make_grid(float(np.max(cc[0]))))
    return _cc_block(sp, cc, 3)


def ohd_cc_pantheon(sn, cc):
    """ohd/cc_pantheon.py: theta = (f_cc, H0, M, Om, w0); late thawing; Pantheon+ (Cholesky, no velocity term) + CC with
    f_cc inflating the errors: chi2_cc * f ** -2 and + 2 N ln f in the normalisation (:63, :92)."""
    bounds = np.array([(0.1, 1.5), (55, 80), (-20, -19), (0.15, 0.70), (-1.0, -1 / 3)], dtype=np.float64)  # :68-77
    sp = LikelihoodSpec(ndim=5, family=S.FAMILY_LATE, de_model=S.DE_THAWING, col_H0=1, col_Om=3, col_w0=4, z_grid=_grid(sn[0]),
                        bounds=bounds)
    _sn_block(sp, sn, S.SN_CHOLESKY, 0.0, 2, None)
    return _cc_block(sp, cc, 0, norm_sign=-1.0)
