"""Model specification shared by the CUDA engine (and, in tests, the oracle): the ctypes mirror of `cl_spec`
in include/cosmolike.h plus a small builder that keeps the NumPy operands alive.

The reference has no configuration layer — models are switched by editing the fit scripts (SURVEY.md D7,
section 5 "Config / flags").  `LikelihoodSpec` is the explicit equivalent: E(z) family, dark-energy model, a
role -> theta-column map, the data blocks (SN / BAO / compressed CMB / cosmic chronometers),
Gaussian terms, the prior box and the CPL guard.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

CL_ABI_VERSION = 4
CL_MAX_DIM, CL_MAX_VEL, CL_MAX_GAUSS, CL_MAX_BAO, CL_MAX_GL, CL_MAX_CC, CL_SN_SMALL_MAX = 12, 3, 4, 32, 128, 64, 64

FAMILY_LATE, FAMILY_FULL = 0, 1
DE_LCDM, DE_WCDM, DE_CPL, DE_THAWING = 0, 1, 2, 3
SN_CHOLESKY, SN_INVCOV = 0, 1
VEL_DIVIDE, VEL_MULTIPLY = 0, 1
BAO_DV, BAO_DM, BAO_DH, BAO_FAP = 0, 1, 2, 3
DH_EXACT, DH_PCHIP = 0, 1
RD_FIXED, RD_PARAM, RD_FIT = 0, 1, 2
CMB_NONE, CMB_R_LA_WB, CMB_THETA_WB_WM = 0, 1, 2
OUT_CHI2, OUT_LOGLIKE, OUT_LOGPROB = 0, 1, 2

#: quantity strings of the reference's BAO tables -> codes (bao/desi_cmb_union3.py:72)
BAO_QTY_MAP = {"DV_over_rs": BAO_DV, "DM_over_rs": BAO_DM, "DH_over_rs": BAO_DH, "F_AP": BAO_FAP}

C_KMS = 299792.458  # scipy.constants.c / 1000 (sn/pantheon.py:12)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class ClCmbConsts(C.Structure):
    _fields_ = [
        ("Or_h2", C.c_double), ("Omnu_h2", C.c_double), ("Ogamma_h2", C.c_double),
        ("nu_m0", C.c_double), ("nu_rho0", C.c_double), ("nu_q2", C.c_double * 5), ("nu_w", C.c_double * 5),
        ("zstar_s1", C.c_double), ("zstar_s2", C.c_double), ("zstar_b", C.c_double), ("zstar_m", C.c_double),
        ("rdrag_b", C.c_double), ("rdrag_m", C.c_double),
    ]


class ClSpec(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32), ("ndim", C.c_int32),
        ("family", C.c_int32), ("de_model", C.c_int32), ("col_H0", C.c_int32),
        ("H0_fixed", C.c_double), ("H0_scale", C.c_double),
        ("col_Om", C.c_int32), ("Om_is_physical", C.c_int32), ("col_obh2", C.c_int32), ("col_och2", C.c_int32),
        ("col_w0", C.c_int32), ("col_wa", C.c_int32),
        ("cmbc", ClCmbConsts),
        ("z_grid", _dp), ("n_grid", C.c_int32),
        ("n_sn", C.c_int32), ("sn_zcmb", _dp), ("sn_zhel", _dp), ("sn_obs", _dp),
        ("sn_cov_form", C.c_int32), ("sn_mat", _dp), ("col_offset", C.c_int32), ("n_vel", C.c_int32),
        ("col_vel", C.c_int32 * CL_MAX_VEL), ("sn_vel_weight", _dp), ("vel_scale", C.c_double), ("vel_mode", C.c_int32),
        ("sn_mu_fixed", _dp), ("n_lin", C.c_int32), ("col_lin", C.c_int32 * CL_MAX_VEL), ("sn_lin_template", _dp),
        ("n_bao", C.c_int32), ("bao_z", _dp), ("bao_value", _dp), ("bao_qty", _ip), ("bao_inv_cov", _dp),
        ("bao_dh_mode", C.c_int32), ("rd_mode", C.c_int32), ("rd_fixed", C.c_double), ("col_rd", C.c_int32),
        ("cmb_mode", C.c_int32), ("cmb_prior", C.c_double * 3), ("cmb_weight", C.c_double * 9),
        ("gl_x", _dp), ("gl_w", _dp), ("n_gl", C.c_int32),
        ("n_cc", C.c_int32), ("cc_z", _dp), ("cc_H", _dp), ("cc_inv_cov", _dp), ("col_fcc", C.c_int32),
        ("cc_logdet", C.c_double), ("cc_norm_sign", C.c_double),
        ("n_gauss_chi2", C.c_int32), ("gauss_chi2_col", C.c_int32 * CL_MAX_GAUSS),
        ("gauss_chi2_mean", C.c_double * CL_MAX_GAUSS), ("gauss_chi2_sigma", C.c_double * CL_MAX_GAUSS),
        ("has_bounds", C.c_int32), ("lo", C.c_double * CL_MAX_DIM), ("hi", C.c_double * CL_MAX_DIM),
        ("log_prior_norm", C.c_double),
        ("n_gauss_prior", C.c_int32), ("gauss_prior_col", C.c_int32 * CL_MAX_GAUSS),
        ("gauss_prior_mean", C.c_double * CL_MAX_GAUSS), ("gauss_prior_sigma", C.c_double * CL_MAX_GAUSS),
        ("guard_cpl", C.c_int32), ("guard_value", C.c_double),
    ]


# ------------------------------------------------------------------------------------------------
# compressed-CMB constant sets (the five cmb/data_*_compression.py modules differ only in these)
# ------------------------------------------------------------------------------------------------
_NU_WEIGHTS = (0.0380051, 0.262676, 0.46542, 0.217161, 0.0167379)  # nu_evolution.py:20
_NU_Q_COEFFS = (  # nu_evolution.py:10-16
    (0.51957626, -0.32971882, 61.81645189, 1.63914879),
    (1.44003028, +0.18098045, 55.20830625, 1.62412241),
    (2.98731126, -0.15154978, 38.50221716, 1.54532306),
    (5.51951238, +0.27573416, 27.23306000, 1.54350910),
    (9.82330637, -1.14831159, 14.84585003, 1.55585284),
)


def _nu_nodes(m0):
    """5-node massive-neutrino quadrature: qs and rho0 (nu_evolution.py:5-28)."""
    qs = np.array([a + b / (m0**d + c) for a, b, c, d in _NU_Q_COEFFS], dtype=np.float64)
    rho0 = 0.0
    for i in range(5):
        rho0 += _NU_WEIGHTS[i] * np.sqrt(qs[i] ** 2 + m0**2)
    return qs, float(rho0)


@dataclass
class CmbConstants:
    """Constants of one compressed-CMB module (cmb/data_planck_act_compression.py:15-50,86-138)."""
    name: str
    priors: np.ndarray
    covariance: np.ndarray
    mode: int
    Or_h2: float
    Omnu_h2: float
    Ogamma_h2: float
    nu_m0: float
    nu_rho0: float
    nu_q: np.ndarray
    zstar: tuple
    rdrag: tuple
    zdrag: tuple
    nu_w: tuple = _NU_WEIGHTS

    @property
    def inv_cov(self):
        return np.linalg.inv(self.covariance)

    def to_c(self) -> ClCmbConsts:
        k = ClCmbConsts()
        k.Or_h2, k.Omnu_h2, k.Ogamma_h2 = self.Or_h2, self.Omnu_h2, self.Ogamma_h2
        k.nu_m0, k.nu_rho0 = self.nu_m0, self.nu_rho0
        q2 = np.asarray(self.nu_q, dtype=np.float64) ** 2
        for i in range(5):
            k.nu_q2[i] = q2[i]
            k.nu_w[i] = self.nu_w[i]
        k.zstar_s1, k.zstar_s2, k.zstar_b, k.zstar_m = self.zstar
        k.rdrag_b, k.rdrag_m = self.rdrag
        return k


def _cmb_constants(name, priors, cov, mode, n_eff, ogamma, omnu_form, zstar, rdrag, zdrag, cov_scale=1.0):
    k_B, TCMB, mnu_tot = 8.617333262e-5, 2.7255, 0.06
    T_nu0 = (4 / 11) ** (1 / 3) * (n_eff / 3) ** (1 / 4) * TCMB
    m0 = mnu_tot / (T_nu0 * k_B)
    if omnu_form == "act":
        omnu = mnu_tot / (94.0641 / (n_eff / 3.0) ** 0.75)
    elif omnu_form == "planck":
        omnu = (mnu_tot / 94.07) * (n_eff / 3) ** (3 / 4)
    else:
        omnu = mnu_tot / (94.07 / (n_eff / 3.0) ** 0.75)
    or_h2 = ogamma * (1 + (2 * n_eff / 3) * (7 / 8) * (4 / 11) ** (4 / 3))
    qs, rho0 = _nu_nodes(m0)
    return CmbConstants(name, np.array(priors, dtype=np.float64), cov_scale * np.array(cov, dtype=np.float64), mode,
                        or_h2, omnu, ogamma, m0, rho0, qs, zstar, rdrag, zdrag)


def cmb_planck_act():
    """cmb/data_planck_act_compression.py"""
    return _cmb_constants(
        "planck_act", [1.74795802, 301.803306, 0.0224962530],
        [[1.54911112e-05, 1.03997132e-04, -2.10953275e-07],
         [1.03997132e-04, 5.43880523e-03, -1.53612827e-06],
         [-2.10953275e-07, -1.53612827e-06, 1.23574770e-08]],
        CMB_R_LA_WB, 3.044, 2.472975328714087e-05, "act",
        (0.70130133, 1.00839438, 1.02468387, 1.18438972), (0.99625075, 1.00593295),
        (1.00791144, 1.00585853, 1.05510863, 0.84044899))


def cmb_act():
    """cmb/data_act_compression.py"""
    k = cmb_planck_act()
    k.name = "act"
    k.priors = np.array([1.76114018, 301.858188, 0.0225906400])
    k.covariance = np.array([[4.21173357e-05, 2.72141593e-04, -1.81499538e-07],
                             [2.72141593e-04, 8.16733306e-03, 2.41363324e-07],
                             [-1.81499538e-07, 2.41363324e-07, 2.81508052e-08]])
    return k


def cmb_planck():
    """cmb/data_planck_compression.py"""
    return _cmb_constants(
        "planck", [1.75063846, 301.760701, 0.0223597502],
        [[2.09107356e-05, 1.78419597e-04, -4.46283183e-07],
         [1.78419597e-04, 7.81249750e-03, -4.24834772e-06],
         [-4.46283183e-07, -4.24834772e-06, 2.21402189e-08]],
        CMB_R_LA_WB, 3.046, 2.4729753287140862e-05, "planck",
        (0.73491615, 1.00820929, 1.01709662, 1.17030559), (1.00078696, 1.00128548),
        (1.00044649, 1.00006975, 1.00041899, 1.00135313))


def cmb_planck_lens():
    """cmb/data_planck_lens_compression.py"""
    k = cmb_planck()
    k.name = "planck_lens"
    k.priors = np.array([1.74996427, 301.757385, 0.0223731992])
    k.covariance = np.array([[1.59647091e-05, 1.63009220e-04, -3.62871093e-07],
                             [1.63009220e-04, 7.90694821e-03, -4.51155896e-06],
                             [-3.62871093e-07, -4.51155896e-06, 2.12418149e-08]])
    return k


def cmb_early_lcdm():
    """cmb/data_early_lcdm_compression.py: (theta*, omega_b, omega_m)"""
    return _cmb_constants(
        "early_lcdm", [0.010410274, 0.02223, 0.14208],
        [[0.00662099420, 0.124442058, -1.19287532],
         [0.124442058, 21.3441666, -94.0008323],
         [-1.19287532, -94.0008323, 1488.41714]],
        CMB_THETA_WB_WM, 3.044, 2.472975328714087e-05, "early",
        (0.75717491, 1.00737989, 1.02737182, 1.20432292), (1.00140649, 1.00072621),
        (1.00329735, 0.99968141, 1.00232108, 0.9893333), cov_scale=1e-9)


def cmb_rdrag_plain():
    """The r_drag fit written out inside bao/desi_union3_bbn.py:31-45 and bao/desi_des5y_bbn.py:22-42 (arXiv:2106.00428 eq. 8
    without the per-module rescaling exponents of cmb/data_*_compression.py): omega_b ** 1, omega_m ** 1."""
    k = cmb_planck()
    k.name = "rdrag_plain"
    k.rdrag = (1.0, 1.0)
    return k


CMB_MODULES = {"planck_act": cmb_planck_act, "act": cmb_act, "planck": cmb_planck, "planck_lens": cmb_planck_lens,
               "early_lcdm": cmb_early_lcdm}


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _ptr(a):
    return a.ctypes.data_as(_dp)


@dataclass
class LikelihoodSpec:
    """Python-side model spec; `.c_spec()` gives the ctypes `cl_spec` (arrays are kept alive by this object)."""
    ndim: int
    family: int = FAMILY_LATE
    de_model: int = DE_LCDM
    col_H0: int = -1
    H0_fixed: float = 70.0
    H0_scale: float = 1.0
    col_Om: int = -1
    Om_is_physical: bool = False
    col_obh2: int = -1
    col_och2: int = -1
    col_w0: int = -1
    col_wa: int = -1
    cmb_consts: CmbConstants | None = None
    z_grid: np.ndarray | None = None
    # SN
    sn_zcmb: np.ndarray | None = None
    sn_zhel: np.ndarray | None = None
    sn_obs: np.ndarray | None = None
    sn_cov_form: int = SN_CHOLESKY
    sn_mat: np.ndarray | None = None
    col_offset: int = -1
    col_vel: tuple = ()
    sn_vel_weight: np.ndarray | None = None
    vel_scale: float = 100.0
    vel_mode: int = VEL_DIVIDE
    sn_mu_fixed: np.ndarray | None = None      # [n_sn], NaN where the model applies (SH0ES calibrators)
    col_lin: tuple = ()
    sn_lin_template: np.ndarray | None = None  # [n_lin, n_sn]
    # BAO
    bao_z: np.ndarray | None = None
    bao_value: np.ndarray | None = None
    bao_qty: np.ndarray | None = None
    bao_inv_cov: np.ndarray | None = None
    bao_dh_mode: int = DH_EXACT
    rd_mode: int = RD_FIXED
    rd_fixed: float = 147.09
    col_rd: int = -1
    # CMB
    cmb_mode: int = CMB_NONE
    cmb_prior: np.ndarray | None = None
    cmb_weight: np.ndarray | None = None
    gl_nodes: int = 100
    # CC
    cc_z: np.ndarray | None = None
    cc_H: np.ndarray | None = None
    cc_inv_cov: np.ndarray | None = None
    col_fcc: int = -1
    cc_logdet: float = 0.0
    cc_norm_sign: float = 0.0
    # Gaussian chi2 terms / prior
    gauss_chi2: tuple = ()   # ((col, mean, sigma), ...)
    bounds: np.ndarray | None = None
    log_prior_norm: float | None = None
    gauss_prior: tuple = ()  # ((col, mean, sigma), ...)
    guard_cpl: bool = False
    guard_value: float = -1e8
    _keep: list = field(default_factory=list, repr=False)

    # -- helpers mirroring the reference's import-time section -------------------------------------------------
    @staticmethod
    def make_grid(z_max, num=4000):
        """z_grid = np.linspace(0, z_max + 0.1, num) (sn/pantheon.py:16)"""
        return np.linspace(0, z_max + 0.1, num=num)

    @staticmethod
    def step_weight(z_cmb, z_turn):
        """np.where(z_cmb <= z_turn, 1, -1) (sn/pantheon.py:46)"""
        return np.where(np.asarray(z_cmb) <= z_turn, 1.0, -1.0)

    def c_spec(self) -> ClSpec:
        s = ClSpec()
        # the arrays the struct points to live as long as the struct itself (every call returns an independent struct: an
        # Oracle and an Engine built from one LikelihoodSpec must not invalidate each other's pointers)
        keep = s._keep = []

        def arr(a, dtype=np.float64):
            a = np.ascontiguousarray(np.asarray(a, dtype=dtype))
            keep.append(a)
            return a

        s.abi_version, s.ndim = CL_ABI_VERSION, int(self.ndim)
        if not (1 <= self.ndim <= CL_MAX_DIM):
            raise ValueError("ndim out of range")
        s.family, s.de_model = int(self.family), int(self.de_model)
        s.col_H0, s.H0_fixed, s.H0_scale = int(self.col_H0), float(self.H0_fixed), float(self.H0_scale)
        s.col_Om, s.Om_is_physical = int(self.col_Om), int(bool(self.Om_is_physical))
        s.col_obh2, s.col_och2, s.col_w0, s.col_wa = int(self.col_obh2), int(self.col_och2), int(self.col_w0), int(self.col_wa)
        if self.cmb_consts is not None:
            s.cmbc = self.cmb_consts.to_c()
        if self.z_grid is not None:
            g = arr(self.z_grid)
            s.z_grid, s.n_grid = _ptr(g), g.size
        # SN
        if self.sn_zcmb is not None and len(self.sn_zcmb) > 0:
            zc, zh, ob, m = arr(self.sn_zcmb), arr(self.sn_zhel), arr(self.sn_obs), arr(self.sn_mat)
            n = zc.size
            if zh.size != n or ob.size != n or m.shape != (n, n):
                raise ValueError("inconsistent SN block shapes")
            s.n_sn, s.sn_zcmb, s.sn_zhel, s.sn_obs, s.sn_mat = n, _ptr(zc), _ptr(zh), _ptr(ob), _ptr(m)
            s.sn_cov_form, s.col_offset = int(self.sn_cov_form), int(self.col_offset)
            s.n_vel = len(self.col_vel)
            if s.n_vel > CL_MAX_VEL:
                raise ValueError("too many velocity templates")
            if s.n_vel:
                w = arr(np.asarray(self.sn_vel_weight, dtype=np.float64).reshape(s.n_vel, n))
                s.sn_vel_weight = _ptr(w)
                for i, col in enumerate(self.col_vel):
                    s.col_vel[i] = int(col)
            s.vel_scale, s.vel_mode = float(self.vel_scale), int(self.vel_mode)
            if self.sn_mu_fixed is not None:
                mf = arr(self.sn_mu_fixed)
                if mf.size != n:
                    raise ValueError("sn_mu_fixed must have n_sn entries")
                s.sn_mu_fixed = _ptr(mf)
            s.n_lin = len(self.col_lin)
            if s.n_lin > CL_MAX_VEL:
                raise ValueError("too many linear templates")
            if s.n_lin:
                lt = arr(np.asarray(self.sn_lin_template, dtype=np.float64).reshape(s.n_lin, n))
                s.sn_lin_template = _ptr(lt)
                for i, col in enumerate(self.col_lin):
                    s.col_lin[i] = int(col)
        # BAO
        if self.bao_z is not None and len(self.bao_z) > 0:
            bz, bv, bw = arr(self.bao_z), arr(self.bao_value), arr(self.bao_inv_cov)
            bq = arr(self.bao_qty, np.int32)
            k = bz.size
            if k > CL_MAX_BAO or bv.size != k or bq.size != k or bw.shape != (k, k):
                raise ValueError("inconsistent BAO block shapes")
            s.n_bao, s.bao_z, s.bao_value, s.bao_inv_cov = k, _ptr(bz), _ptr(bv), _ptr(bw)
            s.bao_qty = bq.ctypes.data_as(_ip)
            s.bao_dh_mode, s.rd_mode, s.rd_fixed, s.col_rd = int(self.bao_dh_mode), int(self.rd_mode), float(self.rd_fixed), int(self.col_rd)
        # CMB
        s.cmb_mode = int(self.cmb_mode)
        if self.cmb_mode != CMB_NONE:
            pr = _f64(self.cmb_prior if self.cmb_prior is not None else self.cmb_consts.priors)
            wt = _f64(self.cmb_weight if self.cmb_weight is not None else self.cmb_consts.inv_cov).reshape(9)
            for i in range(3):
                s.cmb_prior[i] = pr[i]
            for i in range(9):
                s.cmb_weight[i] = wt[i]
        if self.cmb_mode != CMB_NONE or self.family == FAMILY_FULL:
            x, w = np.polynomial.legendre.leggauss(int(self.gl_nodes))  # cmb/data_planck_act_compression.py:150
            x, w = arr(x), arr(w)
            s.gl_x, s.gl_w, s.n_gl = _ptr(x), _ptr(w), x.size
        # CC
        if self.cc_z is not None and len(self.cc_z) > 0:
            cz, ch, cw = arr(self.cc_z), arr(self.cc_H), arr(self.cc_inv_cov)
            if cz.size > CL_MAX_CC or cw.shape != (cz.size, cz.size):
                raise ValueError("inconsistent CC block shapes")
            s.n_cc, s.cc_z, s.cc_H, s.cc_inv_cov = cz.size, _ptr(cz), _ptr(ch), _ptr(cw)
            s.col_fcc, s.cc_logdet, s.cc_norm_sign = int(self.col_fcc), float(self.cc_logdet), float(self.cc_norm_sign)
        else:
            s.col_fcc = -1
        # Gaussian terms
        s.n_gauss_chi2 = len(self.gauss_chi2)
        for i, (col, mean, sigma) in enumerate(self.gauss_chi2):
            s.gauss_chi2_col[i], s.gauss_chi2_mean[i], s.gauss_chi2_sigma[i] = int(col), float(mean), float(sigma)
        s.n_gauss_prior = len(self.gauss_prior)
        for i, (col, mean, sigma) in enumerate(self.gauss_prior):
            s.gauss_prior_col[i], s.gauss_prior_mean[i], s.gauss_prior_sigma[i] = int(col), float(mean), float(sigma)
        if self.bounds is not None:
            b = np.asarray(self.bounds, dtype=np.float64)
            if b.shape != (self.ndim, 2):
                raise ValueError("bounds must be [ndim, 2]")
            s.has_bounds = 1
            for j in range(self.ndim):
                s.lo[j], s.hi[j] = b[j, 0], b[j, 1]
            norm = self.log_prior_norm
            if norm is None:
                norm = -np.sum(np.log(b[:, 1] - b[:, 0]))  # sn/pantheon.py:77
            s.log_prior_norm = float(norm)
        else:
            s.log_prior_norm = float(self.log_prior_norm or 0.0)
        s.guard_cpl, s.guard_value = int(bool(self.guard_cpl)), float(self.guard_value)
        return s
