#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py            # all cases
    python tests/golden/make_golden.py sn_pantheon bao_desi

How it works
------------
* One subprocess per reference script, cwd = /root/reference (the loaders use cwd-relative paths and
  `cmb.set_HZ` freezes a module global at first numba compile, so two CMB scripts in one process give NaN:
  SURVEY.md trap T1).
* The Pantheon+ and DES-Dovekie covariance blobs are absent from the checkout (`.MISSING_LARGE_BLOBS`), so
  for scripts that need them the loader module is pre-seeded in `sys.modules` with a stub that serves the
  REAL redshifts/magnitudes from the reference's own data files plus the seeded synthetic SPD covariance of
  `cosmology_model_fit_b200.synthetic`.  Everything else is the reference's code.
* Output: `data_*.npz` (the observational columns a test needs to rebuild the case without the reference
  tree) and `golden_<case>.npz` (theta batch + the reference's chi-squared, components and distances).
"""
import os
import subprocess
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(os.path.dirname(os.path.dirname(HERE)), "cosmology_model_fit_b200", "data")   # observational columns (inputs, not goldens)
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("COSMO_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)

from cosmology_model_fit_b200.synthetic import synthetic_sn_covariance, uniform_theta  # noqa: E402


# ------------------------------------------------------------------------------------------------
# data columns (read with the same parsers as the reference loaders, but only the needed columns)
# ------------------------------------------------------------------------------------------------
def dump_data_columns():
    import pandas as pd

    df = pd.read_csv(f"{REF}/y2022pantheonSHOES/raw-data/distances.txt", sep=" ")
    np.savez_compressed(
        f"{DATA}/data_pantheon_plus.npz",
        zHD=df["zHD"].to_numpy(np.float64),
        zHEL=df["zHEL"].to_numpy(np.float64),
        m_b_corr=df["m_b_corr"].to_numpy(np.float64),
        m_b_corr_err_DIAG=df["m_b_corr_err_DIAG"].to_numpy(np.float64),
        RA=df["RA"].to_numpy(np.float64),
        DEC=df["DEC"].to_numpy(np.float64),
        IDSURVEY=df["IDSURVEY"].to_numpy(np.int32),
        CEPH_DIST=df["CEPH_DIST"].to_numpy(np.float64),
        IS_CALIBRATOR=df["IS_CALIBRATOR"].to_numpy(np.int32),
    )
    df = pd.read_csv(f"{REF}/y2025DESdovekie/raw-data/distances.csv", sep=r"\s+")
    np.savez_compressed(
        f"{DATA}/data_des_dovekie.npz",
        zHD=df["zHD"].to_numpy(np.float64),
        zHEL=df["zHEL"].to_numpy(np.float64),
        MU=df["MU"].to_numpy(np.float64),
        MUERR=df["MUERR"].to_numpy(np.float64),
    )
    df = pd.read_csv(f"{REF}/y2026union3_1/raw-data/bins_union_3_1.csv")
    cov = np.genfromtxt(f"{REF}/y2026union3_1/raw-data/covariance.txt", dtype=np.float64)
    n = df["zcmb"].size
    np.savez_compressed(
        f"{DATA}/data_union3_1.npz",
        zcmb=df["zcmb"].to_numpy(np.float64),
        zhel=df["zhel"].to_numpy(np.float64),
        mb=df["mb"].to_numpy(np.float64),
        cov=cov.reshape(n, n),
    )
    out = {}
    for tag, dfile, cfile in (("dr2", "data.csv", "covariance.txt"), ("fs_lya", "data_fs_lya.csv", "covariance_fs_lya.txt")):
        d = np.genfromtxt(
            f"{REF}/y2025BAO/raw-data/{dfile}",
            dtype=[("z", np.float64), ("value", np.float64), ("quantity", "U10")],
            delimiter=",",
            names=True,
        )
        out[f"{tag}_z"] = d["z"]
        out[f"{tag}_value"] = d["value"]
        out[f"{tag}_quantity"] = d["quantity"]
        out[f"{tag}_cov"] = np.loadtxt(f"{REF}/y2025BAO/raw-data/{cfile}", delimiter=" ", dtype=np.float64)
    np.savez_compressed(f"{DATA}/data_desi_bao.npz", **out)
    # cosmic chronometers: the loader builds the covariance with the reference's own pchip (y2005cc/data.py:5-40)
    cwd = os.getcwd()
    os.chdir(REF)
    sys.path.insert(0, REF)
    try:
        from y2005cc.data import get_data as cc_get
        _, zc, Hc, covc = cc_get()
    finally:
        os.chdir(cwd)
        sys.path.remove(REF)
    np.savez_compressed(f"{DATA}/data_cc.npz", z=np.asarray(zc, dtype=np.float64), H=np.asarray(Hc, dtype=np.float64), cov=np.asarray(covc, dtype=np.float64))


# ------------------------------------------------------------------------------------------------
# loader stubs for the missing blobs
# ------------------------------------------------------------------------------------------------
def _stub_pantheon():
    d = np.load(f"{DATA}/data_pantheon_plus.npz")
    cov_full = synthetic_sn_covariance(d["m_b_corr_err_DIAG"])
    keep = np.where(d["zHD"] > 0.01)[0]  # y2022pantheonSHOES/data.py:25

    def get_data():
        return ("Pantheon+ (2022)", d["zHD"][keep], d["zHEL"][keep], d["m_b_corr"][keep], cov_full[np.ix_(keep, keep)])

    def get_data_with_position():
        return ("Pantheon+ (2022)", d["zHD"][keep], d["zHEL"][keep], d["m_b_corr"][keep], d["RA"][keep], d["DEC"][keep],
                d["IDSURVEY"][keep], cov_full[np.ix_(keep, keep)])

    m = types.ModuleType("y2022pantheonSHOES.data")
    m.get_data = get_data
    m.get_data_with_position = get_data_with_position
    # y2022pantheonSHOES/data_shoes.py:24-39 (calibrators kept regardless of redshift)
    sel = np.where((d["IS_CALIBRATOR"] == 1) | (d["zHD"] > 0.01))[0]

    def get_data_shoes(z_cut_ceph=0.0):
        return ("Pantheon+ and SH0ES", d["zHD"][sel], d["zHEL"][sel], d["m_b_corr"][sel], d["CEPH_DIST"][sel],
                cov_full[np.ix_(sel, sel)])

    ms = types.ModuleType("y2022pantheonSHOES.data_shoes")
    ms.get_data = get_data_shoes
    pkg = types.ModuleType("y2022pantheonSHOES")
    pkg.__path__ = []
    pkg.data = m
    pkg.data_shoes = ms
    sys.modules["y2022pantheonSHOES"] = pkg
    sys.modules["y2022pantheonSHOES.data"] = m
    sys.modules["y2022pantheonSHOES.data_shoes"] = ms


def _stub_des():
    d = np.load(f"{DATA}/data_des_dovekie.npz")
    cov_full = synthetic_sn_covariance(d["MUERR"])
    order = np.argsort(d["zHD"])  # y2025DESdovekie/data.py:25

    def get_data():
        return ("DES-SN5YR", d["zHD"][order], d["zHEL"][order], d["MU"][order], cov_full[order, :][:, order])

    m = types.ModuleType("y2025DESdovekie.data")
    m.get_data = get_data
    m.effective_sample_size = 1820
    pkg = types.ModuleType("y2025DESdovekie")
    pkg.__path__ = []
    pkg.data = m
    sys.modules["y2025DESdovekie"] = pkg
    sys.modules["y2025DESdovekie.data"] = m


def _enter_reference():
    os.chdir(REF)
    sys.path.insert(0, REF)


def _grid_probe_z(z_grid, rng, n=24):
    """Query redshifts: random interior points, exact grid nodes, both ends and slightly outside."""
    zi = rng.uniform(z_grid[0], z_grid[-1], n)
    nodes = z_grid[[1, 7, 1999, len(z_grid) - 2]]
    return np.concatenate([zi, nodes, [0.0, z_grid[-1], z_grid[-1] + 0.05, 1e-4]])


# ------------------------------------------------------------------------------------------------
# cases (each runs in its own subprocess)
# ------------------------------------------------------------------------------------------------
def case_sn_pantheon():
    """Config 0: sn/pantheon.py, theta = (M, H0, Om, v)."""
    _stub_pantheon()
    _enter_reference()
    import sn.pantheon as ref

    theta = uniform_theta(ref.bounds, 48)
    theta = np.vstack([theta, [[-19.35, 72.0, 0.33, 0.0], [-19.3, 70.0, 0.3, -2.5], [-19.5, 55.0, 0.05, 2.9]]])
    chi2 = np.array([ref.chi_squared(t) for t in theta])
    zq = _grid_probe_z(ref.z_grid, np.random.default_rng(1))
    dm = np.array([ref.DM_z(t, zq) for t in theta[:8]])
    dm_zcmb = np.array([ref.DM_z(t, ref.z_cmb) for t in theta[:4]])
    delta = np.array([ref.mb_vals - t[0] - ref.mu_corr(t, ref.DM_z(t, ref.z_cmb)) - ref.mu_theory(ref.DM_z(t, ref.z_cmb)) for t in theta[:4]])
    # log_probability incl. out-of-box rows (sn/pantheon.py:80-97)
    tp = np.vstack([theta[:6], [[-18.5, 70.0, 0.3, 0.0], [-19.3, 70.0, 0.0, 0.0], [-19.3, 95.0, 0.3, 0.0]]])
    logp = np.array([ref.log_probability(t) for t in tp])
    return dict(theta=theta, chi2=chi2, zq=zq, dm=dm, dm_zcmb=dm_zcmb, delta=delta, theta_logp=tp, logp=logp,
                bounds=ref.bounds, z_grid=ref.z_grid, n_sn=np.int64(ref.z_cmb.size))


def case_sn_union3_1():
    """sn/union3_1.py, theta = (dM, Om, v), H0 fixed at 70."""
    _enter_reference()
    import sn.union3_1 as ref

    bounds = np.array([(-1.0, 1.0), (0.1, 0.7), (-9.0, 9.0)])
    theta = uniform_theta(bounds, 40)
    theta = np.vstack([theta, [[-0.05, 0.3, -3.0], [-0.05, 0.3, 0.0], [0.027, 0.335, 0.0],
                               [-0.00544678, 0.29682551, -3.11693562]]])
    chi2 = np.array([ref.chi_squared(t) for t in theta])
    zq = np.concatenate([[0.05, 0.5, 2.26226], _grid_probe_z(ref.z_grid, np.random.default_rng(2))])
    dm = np.array([ref.DM_z(zq, t) for t in theta[:8]])
    return dict(theta=theta, chi2=chi2, zq=zq, dm=dm, bounds=bounds, z_grid=ref.z_grid)


def case_bao_desi():
    """bao/desi.py, theta = (h, Om, w0) thawing, r_d fixed, pchip D_H; batch API returns float32."""
    _enter_reference()
    import bao.desi as ref

    theta = uniform_theta(ref.bounds, 40)
    theta = np.vstack([theta, [[0.691, 0.297, -1.0], [0.666, 0.312, -0.768]]])
    chi2 = np.array([ref.chi_squared(t) for t in theta])
    theory = np.array([ref.bao_theory(ref.data["z"], ref.bao_qty, t) for t in theta])
    batch = np.vstack([theta, [[0.9, 0.3, -0.5], [0.6, 0.3, 0.1]]])
    logp32 = ref.log_probs_vectorized(np.ascontiguousarray(batch))
    logp64 = np.array([ref.log_probability(t) for t in batch])
    return dict(theta=theta, chi2=chi2, theory=theory, batch=batch, logp32=logp32, logp64=logp64, bounds=ref.bounds,
                z_grid=ref.z_grid, bao_z=ref.data["z"], bao_value=ref.data["value"], bao_qty=ref.bao_qty,
                bao_cov=ref.cov_matrix)


def case_bao_desi_cmb_union3():
    """Config 2: bao/desi_cmb_union3.py, theta = (dM, H0, obh2, och2, v)."""
    _enter_reference()
    import bao.desi_cmb_union3 as ref

    bounds = np.array([(-1.0, 1.0), (60.0, 75.0), (0.01, 0.03), (0.01, 0.25), (-8.0, 8.0)])
    theta = uniform_theta(bounds, 40)
    theta = np.vstack([theta, [[-0.0519, 68.42, 0.02257, 0.11738, -3.0]]])
    n = len(theta)
    chi2 = np.empty(n); c_cmb = np.empty(n); c_bao = np.empty(n); c_sn = np.empty(n)
    cmbd = np.empty((n, 3)); zstar = np.empty(n); rd = np.empty(n); bao_th = np.empty((n, ref.bao.size))
    for i, t in enumerate(theta):
        g = ref.DM_grid(t)
        chi2[i] = ref.chi_squared(t)
        c_cmb[i] = ref.chi2_cmb(t); c_bao[i] = ref.chi2_bao(t, g); c_sn[i] = ref.chi2_sn(t, g)
        cmbd[i] = ref.cmb.cmb_distances(t[2], t[3], t)
        wm = t[2] + t[3] + ref.cmb.Omnu_h2
        zstar[i] = ref.cmb.z_star(t[2], wm); rd[i] = ref.cmb.r_drag(t[2], wm)
        bao_th[i] = ref.bao_theory(ref.bao["z"], ref.bao_qty, t, g)
    cm = ref.cmb
    consts = dict(cmb_priors=cm.DISTANCE_PRIORS, cmb_cov=cm.covariance, cmb_inv_cov=cm.inv_cov_mat, Or_h2=cm.Or_h2,
                  Omnu_h2=cm.Omnu_h2, m0=cm.m0, rho0=cm.rho0, qs=cm.qs, ws=cm.ws, O_GAMMA_H2=cm.O_GAMMA_H2,
                  GL_X=cm.GL_X, GL_W=cm.GL_W, c=cm.c)
    return dict(theta=theta, chi2=chi2, chi2_cmb=c_cmb, chi2_bao=c_bao, chi2_sn=c_sn, cmb_distances=cmbd,
                z_star=zstar, r_drag=rd, bao_theory=bao_th, bounds=bounds, z_grid=ref.z_grid,
                bao_z=ref.bao["z"], bao_value=ref.bao["value"], bao_qty=ref.bao_qty, bao_cov=ref.bao_cov_mat,
                **consts)


def case_bao_desi_fs_lya_cmb():
    """bao/desi_fs_lya_cmb.py: CPL w0wa + the w0+wa>=0 guard, theta = (H0, obh2, och2, w0, wa)."""
    _enter_reference()
    import bao.desi_fs_lya_cmb as ref

    bounds = np.array([(60.0, 75.0), (0.01, 0.03), (0.01, 0.25), (-3.0, 1.0), (-3.0, 2.0)])
    theta = uniform_theta(bounds, 48)
    theta = np.vstack([theta, [[64.9, 0.02251, 0.1189, -0.58, -1.26], [64.9, 0.02251, 0.1189, -0.5, 0.6]]])
    n = len(theta)
    loglike = np.array([ref.log_likelihood(t) for t in theta])
    ok = theta[:, 3] + theta[:, 4] < 0.0
    chi2 = np.full(n, np.nan); c_cmb = np.full(n, np.nan); c_bao = np.full(n, np.nan); cmbd = np.full((n, 3), np.nan)
    for i, t in enumerate(theta):
        if ok[i]:
            chi2[i] = ref.chi_squared(t); c_cmb[i] = ref.chi2_cmb(t); c_bao[i] = ref.chi2_bao(t)
            cmbd[i] = ref.cmb.cmb_distances(t[1], t[2], t)
    return dict(theta=theta, loglike=loglike, chi2=chi2, chi2_cmb=c_cmb, chi2_bao=c_bao, cmb_distances=cmbd,
                bounds=bounds, z_grid=ref.z_grid, bao_z=ref.bao["z"], bao_value=ref.bao["value"],
                bao_qty=ref.bao_qty, bao_cov=ref.cov_mat)


def case_cmb_cmb():
    """cmb/cmb.py: CMB-only, theta = (H0, obh2, och2); blobs = (100 theta*, r*, DM*/1000, z*)."""
    _enter_reference()
    import cmb.cmb as ref

    theta = uniform_theta(ref.bounds, 40)
    theta = np.vstack([theta, [[67.61, 0.0225, 0.1193]]])
    ll = np.empty(len(theta)); blobs = np.empty((len(theta), 4))
    for i, t in enumerate(theta):
        ll[i], blobs[i] = ref.log_likelihood(t)
    tp = np.vstack([theta[:4], [[59.0, 0.0225, 0.12], [67.0, 0.026, 0.12]]])
    logp = np.array([ref.log_probability(t)[0] for t in tp])
    return dict(theta=theta, loglike=ll, blobs=blobs, theta_logp=tp, logp=logp, bounds=ref.bounds)


def _cmb_cmb_with_module(modname):
    """cmb/cmb.py with `import cmb.<modname> as cmb` in place of its checked-in `cmb.data_planck_act_compression`: the
    constants module is pre-seeded under the name cmb/cmb.py imports (the reference's author swaps that import line by hand)."""
    _enter_reference()
    import importlib
    mod = importlib.import_module(f"cmb.{modname}")
    sys.modules["cmb.data_planck_act_compression"] = mod
    import cmb as cmb_pkg
    cmb_pkg.data_planck_act_compression = mod
    import cmb.cmb as ref
    assert ref.cmb is mod

    theta = uniform_theta(ref.bounds, 24)
    theta = np.vstack([theta, [[67.61, 0.0225, 0.1193]]])
    ll = np.empty(len(theta)); blobs = np.empty((len(theta), 4))
    for i, t in enumerate(theta):
        ll[i], blobs[i] = ref.log_likelihood(t)
    return dict(theta=theta, loglike=ll, blobs=blobs, priors=np.asarray(mod.DISTANCE_PRIORS), bounds=ref.bounds)


def case_cmb_cmb_act():
    """cmb/data_act_compression.py (ACT DR6 alone) behind cmb/cmb.py."""
    return _cmb_cmb_with_module("data_act_compression")


def case_cmb_cmb_planck_lens():
    """cmb/data_planck_lens_compression.py (Planck PR3 + lensing) behind cmb/cmb.py."""
    return _cmb_cmb_with_module("data_planck_lens_compression")


def case_cmb_cmb_planck():
    """cmb/data_planck_compression.py (Planck PR3) behind cmb/cmb.py."""
    return _cmb_cmb_with_module("data_planck_compression")


def case_bao_desi_des5y_bbn_theta_star():
    """Config 1 (as checked in): thawing w0, l_A-only CMB term, BBN prior, theta = (dM, H0, obh2, och2, w0)."""
    _stub_des()
    _enter_reference()
    import bao.desi_des5y_bbn_theta_star as ref

    theta = uniform_theta(ref.bounds, 32)
    theta = np.vstack([theta, [[-0.05, 67.5, 0.0222, 0.119, -0.85]]])
    chi2 = np.array([ref.chi_squared(t) for t in theta])
    tp = np.vstack([theta[:6], [[-0.6, 67.0, 0.022, 0.12, -0.8], [0.0, 67.0, 0.022, 0.12, -0.2]]])
    logp = np.array([ref.log_probability(t) for t in tp])
    bao_th = np.array([ref.bao_theory(ref.bao_data["z"], ref.quantities, t) for t in theta[:8]])
    mu = np.array([ref.theory_mu(t) for t in theta[:4]])
    return dict(theta=theta, chi2=chi2, theta_logp=tp, logp=logp, bao_theory=bao_th, mu_theory=mu, bounds=ref.bounds,
                z_grid=ref.z_grid, bbn_mean=np.float64(ref.bbn.Obh2), bbn_sigma=np.float64(ref.bbn.Obh2_sigma),
                n_sn=np.int64(ref.z_cmb.size))


def case_bao_desi_cmb_pantheon():
    """Config 3: bao/desi_cmb_pantheon.py, theta = (M, H0, obh2, och2, v)."""
    _stub_pantheon()
    _enter_reference()
    import bao.desi_cmb_pantheon as ref

    bounds = np.array([(-20.0, -19.0), (60.0, 75.0), (0.019, 0.025), (0.01, 0.25), (-3.0, 1.5)])
    theta = uniform_theta(bounds, 32)
    theta = np.vstack([theta, [[-19.4, 68.0, 0.0224, 0.118, -1.0]]])
    chi2 = np.array([ref.chi_squared(t) for t in theta])
    return dict(theta=theta, chi2=chi2, bounds=bounds, z_grid=ref.z_grid)


def case_bao_desi_cmb_des5y():
    """bao/desi_cmb_des5y.py: DES + DESI FS-Lya (pchip D_H, F_AP rows) + CMB 3x3 + v step at 0.10563."""
    _stub_des()
    _enter_reference()
    import bao.desi_cmb_des5y as ref

    bounds = np.array([(-0.5, 0.5), (60.0, 75.0), (0.010, 0.030), (0.01, 0.25), (-4.5, 4.5)])
    theta = uniform_theta(bounds, 32)
    n = len(theta)
    chi2 = np.empty(n); c_cmb = np.empty(n); c_bao = np.empty(n); c_sn = np.empty(n)
    for i, t in enumerate(theta):
        g = ref.DM_grid(t)
        chi2[i] = ref.chi_squared(t); c_cmb[i] = ref.chi2_cmb(t); c_bao[i] = ref.chi2_bao(t, g); c_sn[i] = ref.chi2_sn(t, g)
    return dict(theta=theta, chi2=chi2, chi2_cmb=c_cmb, chi2_bao=c_bao, chi2_sn=c_sn, bounds=bounds, z_grid=ref.z_grid)


def case_sn_des5y():
    """sn/des5y.py: late LCDM, DES-Dovekie N=1820, v step (z<=0.11 == 0.10563 mask), theta=(dM,H0,Om,v)."""
    _stub_des()
    _enter_reference()
    import sn.des5y as ref

    bounds = np.array([(-1.0, 1.0), (60.0, 80.0), (0.0, 0.8), (-5.0, 5.0)])
    theta = uniform_theta(bounds, 32)
    chi2 = np.array([ref.chi_squared(t) for t in theta])
    return dict(theta=theta, chi2=chi2, bounds=bounds, z_grid=ref.z_grid)


def case_ohd_cc():
    """ohd/cc.py: theta = (H0, Om, f); chi2 = f^2 |L^-1 d|^2; log L with the N ln 2pi + logdet - 2N ln f term."""
    _enter_reference()
    import ohd.cc as ref

    bounds = np.array([(30.0, 100.0), (0.0, 1.0), (0.4, 2.5)])
    theta = uniform_theta(bounds, 40)
    chi2 = np.array([ref.chi_squared(t) for t in theta])
    ll = np.array([ref.log_likelihood(t) for t in theta])
    return dict(theta=theta, chi2=chi2, loglike=ll, bounds=bounds)


def case_bao_desi_bbn():
    """bao/desi_bbn.py: late thawing, r_d = r_drag(obh2, Om h^2) with the Planck-module fit, pchip D_H."""
    _enter_reference()
    import bao.desi_bbn as ref

    theta = uniform_theta(ref.bounds, 40)
    chi2 = np.array([ref.chi_squared(t) for t in theta])
    theory = np.array([ref.bao_theory(ref.bao["z"], ref.bao_qty, t) for t in theta[:8]])
    return dict(theta=theta, chi2=chi2, theory=theory, bounds=ref.bounds, z_grid=ref.z_grid)


def case_bao_desi_pantheon_cc():
    """bao/desi_pantheon_cc.py: theta = (H0, M, r_d, Om, v, f_cc); multiplicative z shift with the 1e-8 floor,
    sampled r_d, CC term with normalisation, box prior."""
    _stub_pantheon()
    _enter_reference()
    import bao.desi_pantheon_cc as ref

    theta = uniform_theta(ref.bounds, 24)
    chi2 = np.array([ref.chi_squared(t) for t in theta])
    ll = np.array([ref.log_likelihood(t) for t in theta])
    tp = np.vstack([theta[:4], [[95.0, -19.3, 147.0, 0.3, 0.0, 1.0]]])
    logp = np.array([ref.log_probability(t) for t in tp])
    return dict(theta=theta, chi2=chi2, loglike=ll, theta_logp=tp, logp=logp, bounds=ref.bounds, z_grid=ref.z_grid)


def case_sn_pantheon_dipole_xyz():
    """sn/pantheon_dipole_xyz.py: theta = (M, H0, Om, vx, vy, vz); three velocity templates n_k * attenuation * mask."""
    _stub_pantheon()
    _enter_reference()
    import sn.pantheon_dipole_xyz as ref

    bounds = np.array([(-20.0, -19.0), (60.0, 80.0), (0.1, 0.6), (-8.0, 8.0), (-8.0, 8.0), (-8.0, 8.0)])
    theta = uniform_theta(bounds, 24)
    chi2 = np.array([ref.chi_squared(t) for t in theta])
    att = 0.5 * (1.0 - np.tanh((ref.z_cmb - 0.10) / 0.02))
    w = np.vstack([ref.nx, ref.ny, ref.nz]) * att * ref.survey_mask
    return dict(theta=theta, chi2=chi2, bounds=bounds, z_grid=ref.z_grid, weights=w)


def case_sn_pantheon_and_sh0es():
    """sn/pantheon_and_sh0es.py: Cepheid distances replace mu for the calibrators; step mask excludes them."""
    _stub_pantheon()
    _enter_reference()
    import sn.pantheon_and_sh0es as ref

    theta = uniform_theta(ref.bounds, 24)
    chi2 = np.array([ref.chi_squared(t) for t in theta])
    return dict(theta=theta, chi2=chi2, bounds=ref.bounds, z_grid=ref.z_grid, n_sn=np.int64(ref.z_cmb.size),
                n_ceph=np.int64(ref.ceph_mask.sum()))


def case_bao_desi_cmb_pantheon_H0trgb():
    """bao/desi_cmb_pantheon_H0trgb.py: linear-in-magnitude flow template + TRGB H0 chi2 term."""
    _stub_pantheon()
    _enter_reference()
    import bao.desi_cmb_pantheon_H0trgb as ref

    bounds = np.array([(-20.0, -19.0), (60.0, 75.0), (0.019, 0.025), (0.01, 0.25), (-1.2, 3.2)])
    theta = uniform_theta(bounds, 24)
    chi2 = np.array([ref.chi_squared(t) for t in theta])
    return dict(theta=theta, chi2=chi2, bounds=bounds, z_grid=ref.z_grid)


def case_bao_desi_cmb():
    """bao/desi_cmb.py: early-LCDM compression (theta*, omega_b, omega_m), thawing w0, emcee vectorize=True (float32)."""
    _enter_reference()
    import bao.desi_cmb as ref

    theta = uniform_theta(ref.bounds, 32)
    chi2 = np.array([ref.chi_squared(t) for t in theta])
    cmbd = np.array([ref.cmb.cmb_distances(t[1], t[2], t) for t in theta])
    batch = np.vstack([theta, [[85.0, 0.022, 0.12, -0.5]]])
    logp32 = ref.log_probability_vect(np.ascontiguousarray(batch))
    return dict(theta=theta, chi2=chi2, cmb_distances=cmbd, batch=batch, logp32=logp32, bounds=ref.bounds, z_grid=ref.z_grid)


def case_bao_desi_union3_obh2_theta_star():
    """bao/desi_union3_obh2_theta_star.py: CMB rows [1:] with the inverse of the sub-covariance (no shift parameter R)."""
    _enter_reference()
    import bao.desi_union3_obh2_theta_star as ref

    bounds = np.array([(-1.0, 1.0), (50.0, 90.0), (0.01, 0.03), (0.05, 0.3), (-12.0, 5.0)])
    theta = uniform_theta(bounds, 32)
    chi2 = np.array([ref.chi_squared(t) for t in theta])
    c_cmb = np.array([ref.chi2_cmb(t) for t in theta])
    return dict(theta=theta, chi2=chi2, chi2_cmb=c_cmb, bounds=bounds, z_grid=ref.z_grid)


def _stub_plotting(*names):
    """Plot helpers some scripts import at module level (matplotlib is not installed; they are not on the path)."""
    for name in names:
        m = types.ModuleType(name)
        m.plot_predictions = m.plot_cc_predictions = lambda *a, **k: None
        sys.modules[name] = m


def _generic(module, bounds, extra=lambda ref, theta: {}, n=24, loglike=False, logp=False):
    import importlib
    ref = importlib.import_module(module)
    bounds = np.asarray(bounds if bounds is not None else ref.bounds, dtype=np.float64)
    theta = uniform_theta(bounds, n)
    out = dict(theta=theta, chi2=np.array([ref.chi_squared(t) for t in theta]), bounds=bounds)
    if loglike:
        out["loglike"] = np.array([ref.log_likelihood(t) for t in theta])
    if logp:  # the script's own log_probability incl. rows outside its box prior (first / last parameter past a bound)
        lo, hi = np.asarray(ref.bounds, dtype=np.float64).T
        mid = 0.5 * (lo + hi)
        o1, o2 = mid.copy(), mid.copy()
        o1[0] = hi[0] + 0.1 * (hi[0] - lo[0]); o2[-1] = lo[-1] - 0.1 * (hi[-1] - lo[-1])
        inside = uniform_theta(np.asarray(ref.bounds, dtype=np.float64), 6, seed=77)
        tp = np.vstack([inside, o1, o2])
        out["theta_logp"] = tp
        out["logp"] = np.array([float(ref.log_probability(t)) for t in tp])
    for name in ("z_grid", "grid"):
        if hasattr(ref, name):
            out["z_grid"] = getattr(ref, name)
    out.update(extra(ref, theta))
    return out


def case_sn_pantheon_dipole():
    """sn/pantheon_dipole.py: one velocity template cos(angle) * tanh attenuation * survey mask (not +-1)."""
    _stub_pantheon()
    _enter_reference()
    att = lambda ref, th: dict(weights=ref.cos_angle * 0.5 * (1.0 - np.tanh((ref.z_cmb - 0.10) / 0.02)) * ref.survey_mask)
    return _generic("sn.pantheon_dipole", [(-20, -19), (62, 78), (0.1, 0.7), (-0.5, 4.5)], att)


def case_ohd_cc_des5y():
    """ohd/cc_des5y.py: theta = (f_cc, dM, H0, Om, w0); late wCDM; DES SN + CC."""
    _stub_des()
    _enter_reference()
    return _generic("ohd.cc_des5y", [(0.2, 3), (-0.5, 0.5), (50, 85), (0.05, 0.6), (-1, -1 / 3)], loglike=True)


def case_ohd_cc_union3():
    """ohd/cc_union3.py: theta = (f_cc, dM, H0, Om, v[km/s]); grid ends at max(z_cmb) (no +0.1)."""
    _enter_reference()
    return _generic("ohd.cc_union3", [(0.05, 3.35), (-1.0, 1.0), (40.0, 95.0), (0.1, 0.7), (-900, 900)], loglike=True)


def case_bao_desi_omh2():
    """bao/desi_omh2.py: theta = (r_d, H0, omega_m, w0); Om = omega_m / h^2; thawing."""
    _enter_reference()
    return _generic("bao.desi_omh2", [(120, 160), (50.0, 85.0), (0.138, 0.148), (-1.0, -1 / 3)])


def case_sn_pantheon_cmb():
    """sn/pantheon_cmb.py: theta = (M, H0, obh2, och2, v); Pantheon+ + compressed CMB, box prior."""
    _stub_pantheon()
    _enter_reference()
    return _generic("sn.pantheon_cmb", None, logp=True)


def case_sn_des5y_cmb():
    """sn/des5y_cmb.py: theta = (dM, H0, obh2, och2, v); DES-Dovekie + compressed CMB."""
    _stub_des()
    _enter_reference()
    return _generic("sn.des5y_cmb", [(-0.7, 0.7), (55, 75), (0.01, 0.03), (0.01, 0.25), (-4.5, 4.5)])


def case_sn_union3_1_cmb():
    """sn/union3_1_cmb.py: theta = (dM, H0, obh2, och2, v); Union3.1 + compressed CMB."""
    _enter_reference()
    return _generic("sn.union3_1_cmb", [(-1.0, 1.0), (60.0, 75.0), (0.01, 0.03), (0.01, 0.25), (-9.0, 9.0)])


def case_ohd_cc_cmb():
    """ohd/cc_cmb.py: theta = (H0, obh2, och2, f_cc); CC + compressed CMB; log_prior = 0 inside the box."""
    _enter_reference()
    return _generic("ohd.cc_cmb", None, loglike=True, logp=True)


def case_ohd_cc_pantheon():
    """ohd/cc_pantheon.py: theta = (f_cc, H0, M, Om, w0); thawing; Pantheon+ + CC with f_cc inflating the errors."""
    _stub_pantheon()
    _stub_plotting("sn.plotting", "ohd.plot_predictions")
    _enter_reference()
    return _generic("ohd.cc_pantheon", None, loglike=True, logp=True)


def case_bao_desi_fs_lya():
    """bao/desi_fs_lya.py: theta = (h, Om, w0); thawing; FS+Lya table with F_AP rows; fixed r_d; pchip D_H."""
    _enter_reference()
    return _generic("bao.desi_fs_lya", [(0.5, 0.8), (0.1, 0.8), (-1.0, 0.0)])


def case_bao_desi_cc():
    """bao/desi_cc.py: theta = (f_cc, H0, r_d, Om, w0); BAO + cosmic chronometers."""
    _enter_reference()
    return _generic("bao.desi_cc", None, loglike=True, logp=True)


def case_bao_desi_des5y_rd():
    """bao/desi_des5y_rd.py: theta = (dM, r_d, H0, Om, v)."""
    _stub_des()
    _enter_reference()
    return _generic("bao.desi_des5y_rd", [(-0.4, 0.4), (146.0, 148.2), (50.0, 85.0), (0.1, 0.6), (-4.5, 4.5)])


def case_bao_desi_union3_rd():
    """bao/desi_union3_rd.py: theta = (dM, r_d, H0, Om, v)."""
    _enter_reference()
    return _generic("bao.desi_union3_rd", [(-1.0, 1.0), (146.0, 148.2), (50.0, 85.0), (0.1, 0.6), (-8.0, 8.0)])


def case_bao_desi_pantheon_rd():
    """bao/desi_pantheon_rd.py: theta = (M, H0, Om, r_d, w0); Gaussian r_d prior inside log_prior."""
    _stub_pantheon()
    _enter_reference()
    return _generic("bao.desi_pantheon_rd", None, logp=True)


def case_bao_desi_bbn_theta_star():
    """bao/desi_bbn_theta_star.py: theta = (H0, obh2, och2, w0); thawing; Planck PR3 constants; l_A-only term."""
    _enter_reference()
    return _generic("bao.desi_bbn_theta_star", [(50.0, 90.0), (0.019, 0.025), (0.05, 0.30), (-1.0, 0.0)])


def case_bao_desi_union3_bbn_theta_star():
    """bao/desi_union3_bbn_theta_star.py: theta = (dM, H0, obh2, och2, v); 15 BAO points incl. F_AP, exact D_H."""
    _enter_reference()
    return _generic("bao.desi_union3_bbn_theta_star", [(-1.0, 1.0), (50.0, 90.0), (0.019, 0.025), (0.05, 0.30), (-8.5, 8.5)])


def case_bao_desi_union3_cc_theta_star():
    """bao/desi_union3_cc_theta_star.py: theta = (f_cc, dM, H0, obh2, och2, v); float32 batch log_likelihood."""
    _enter_reference()
    def extra(ref, theta):
        return dict(loglike=np.array([ref.log_likelihood_single(t) for t in theta]),
                    loglike32=ref.log_likelihood(np.ascontiguousarray(theta)))
    return _generic("bao.desi_union3_cc_theta_star", [(0.2, 3.0), (-1.0, 1.0), (50.0, 85.0), (0.003, 0.050), (0.05, 0.30), (-10.5, 4.5)], extra)


def case_bao_desi_des5y_cc_theta_star():
    """bao/desi_des5y_cc_theta_star.py: theta = (f_cc, dM, H0, obh2, och2, w0); thawing; box prior."""
    _stub_des()
    _enter_reference()
    return _generic("bao.desi_des5y_cc_theta_star", None, loglike=True, logp=True)


def case_bao_desi_union3_omh2():
    """bao/desi_union3_omh2.py: theta = (dM, r_d, H0, omega_m, v)."""
    _enter_reference()
    return _generic("bao.desi_union3_omh2", [(-1.0, 1.0), (120, 160), (50.0, 85.0), (0.138, 0.148), (-12.0, 5.0)])


def case_bao_desi_des5y_omh2():
    """bao/desi_des5y_omh2.py: theta = (dM, r_d, H0, omega_m, v)."""
    _stub_des()
    _enter_reference()
    return _generic("bao.desi_des5y_omh2", [(-0.5, 0.5), (120.0, 165.0), (50.0, 90.0), (0.138, 0.148), (-5.5, 2.5)])


def case_bao_desi_union3_omh2_theta_star():
    """bao/desi_union3_omh2_theta_star.py: early-LCDM rows (theta*, omega_m), inverse of the sub-covariance."""
    _enter_reference()
    return _generic("bao.desi_union3_omh2_theta_star", [(-1.0, 1.0), (50.0, 90.0), (0.01, 0.04), (0.05, 0.3), (-8.5, 8.5)])


def case_bao_desi_pantheon_obh2_theta_star():
    """bao/desi_pantheon_obh2_theta_star.py: early-LCDM rows (theta*, omega_b); thawing; box prior."""
    _stub_pantheon()
    _enter_reference()
    return _generic("bao.desi_pantheon_obh2_theta_star", None, logp=True)


def case_bao_desi_des5y_obh2_theta_star():
    """bao/desi_des5y_obh2_theta_star.py: Planck+ACT rows (l_A, omega_b) with the sub-block of the full inverse."""
    _stub_des()
    _enter_reference()
    return _generic("bao.desi_des5y_obh2_theta_star", None, logp=True)


def case_bao_desi_union3_bbn():
    """bao/desi_union3_bbn.py: theta = (H0, Om, obh2, v, dM); r_drag fit written out in the script."""
    _enter_reference()
    return _generic("bao.desi_union3_bbn", [(55, 80), (0.10, 0.65), (0.019, 0.025), (-12.0, 5.0), (-1.0, 1.0)])


def case_bao_desi_des5y_bbn():
    """bao/desi_des5y_bbn.py: theta = (H0, Om, obh2, w0, dM); BAO chi2 through the Cholesky factor."""
    _stub_des()
    _enter_reference()
    return _generic("bao.desi_des5y_bbn", [(55, 80), (0.10, 0.65), (0.019, 0.025), (-1.0, -1 / 3), (-0.5, 0.5)])


def case_bao_desi_des5y_H0trgb():
    """bao/desi_des5y_H0trgb.py: theta = (dM, H0, r_d, Om, w0); TRGB H0 Gaussian inside log_prior."""
    _stub_des()
    _enter_reference()
    return _generic("bao.desi_des5y_H0trgb", None, logp=True)


def case_bao_desi_cmb_des5y_H0trgb():
    """bao/desi_cmb_des5y_H0trgb.py: theta = (dM, H0, obh2, och2, v); separate 6dF term, TRGB H0 term in chi2."""
    _stub_des()
    _enter_reference()
    return _generic("bao.desi_cmb_des5y_H0trgb", [(-0.5, 0.5), (60.0, 75.0), (0.010, 0.030), (0.01, 0.25), (-6.0, 2.0)])


def case_bao_desi_cmb_union3_H0trgb():
    """bao/desi_cmb_union3_H0trgb.py: theta = (dM, H0, obh2, och2, v)."""
    _enter_reference()
    return _generic("bao.desi_cmb_union3_H0trgb", [(-1.0, 1.0), (60.0, 75.0), (0.010, 0.030), (0.01, 0.25), (-9.5, 3.5)])


def case_bao_desi_des5y_cc():
    """bao/desi_des5y_cc.py: theta = (f_cc, dM, H0, r_d, Om, v); box prior."""
    _stub_des()
    _enter_reference()
    return _generic("bao.desi_des5y_cc", None, loglike=True, logp=True)


def case_bao_desi_fs_lya_union3_cc():
    """bao/desi_fs_lya_union3_cc.py: theta = (f_cc, dM, H0, r_d, Om, v)."""
    _enter_reference()
    return _generic("bao.desi_fs_lya_union3_cc", [(0.01, 3.0), (-1, 1), (45, 90), (100, 200), (0.2, 0.50), (-8.5, 8.5)], loglike=True)


def case_interpolator():
    """interpolator.py known answers on non-uniform and monotone/non-monotone data (pchip + hermite)."""
    _enter_reference()
    from interpolator import interp_hermite, interp_pchip, _pchip_slopes

    rng = np.random.default_rng(7)
    out = {}
    # uniform grid, smooth monotone function with analytic derivative
    x = np.linspace(0.0, 2.5, 400)
    y = np.log1p(x) * 3000.0
    yp = 3000.0 / (1.0 + x)
    xq = np.concatenate([rng.uniform(-0.1, 2.7, 60), x[[0, 1, 5, 398, 399]]])
    out.update(u_x=x, u_y=y, u_yp=yp, u_xq=xq, u_herm=interp_hermite(xq, x, y, yp), u_pchip=interp_pchip(xq, x, y),
               u_slopes=_pchip_slopes(x, y))
    # non-uniform grid, non-monotone data (exercises the zero-slope and end-point limiter branches)
    x2 = np.sort(rng.uniform(0.0, 10.0, 57))
    y2 = np.sin(x2) + 0.1 * x2
    y2[10:13] = y2[10]  # flat run
    xq2 = rng.uniform(-0.5, 10.5, 80)
    out.update(n_x=x2, n_y=y2, n_xq=xq2, n_pchip=interp_pchip(xq2, x2, y2), n_slopes=_pchip_slopes(x2, y2))
    # decreasing data like dh_grid
    y3 = 4000.0 / np.sqrt(0.3 * (1 + x) ** 3 + 0.7)
    out.update(d_y=y3, d_pchip=interp_pchip(xq, x, y3), d_slopes=_pchip_slopes(x, y3))
    return out


def case_solve_triangular():
    """solve_triangular.py returns y.y with L y = b."""
    _enter_reference()
    from solve_triangular import solve_triangular
    from scipy.linalg import cho_factor

    rng = np.random.default_rng(11)
    n = 97
    a = rng.standard_normal((n, n))
    cov = a @ a.T + n * np.eye(n)
    L = cho_factor(cov, lower=True)[0]
    L = np.tril(L)
    b = rng.standard_normal((5, n))
    val = np.array([solve_triangular(L, bi) for bi in b])
    return dict(L=L, b=b, value=val)


CASES = {k[5:]: v for k, v in list(globals().items()) if k.startswith("case_")}


def main(argv):
    if len(argv) >= 2 and argv[0] == "--worker":
        name = argv[1]
        res = CASES[name]()
        np.savez_compressed(f"{HERE}/golden_{name}.npz", **{k: np.asarray(v) for k, v in res.items()})
        return 0
    names = argv or list(CASES)
    if not os.path.exists(f"{DATA}/data_pantheon_plus.npz") or not argv:
        dump_data_columns()
    for name in names:
        print(f"[golden] {name} ...", flush=True)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", name], check=True)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
