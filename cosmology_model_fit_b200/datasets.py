"""Observational inputs for the engine.

The reference's loaders (`y20xx*/data.py: get_data()`) are reused as-is when a reference checkout is on
sys.path; this module only (a) reads the compact column files shipped inside the package (cosmology_model_fit_b200/data/, or
$COSMOLIKE_DATA; the GPU box has no reference tree) and (b) applies the same row selection / ordering as the loaders, with the seeded
synthetic covariance standing in for the blobs missing from the reference checkout.
"""
from __future__ import annotations

import os

import numpy as np

from .synthetic import synthetic_sn_covariance

_DATA = os.environ.get("COSMOLIKE_DATA") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def _load(name, root=None):
    return np.load(os.path.join(root or _DATA, name))


def pantheon_plus(cut=True, root=None):
    """(z_cmb, z_hel, m_b, cov).  cut=True applies zHD > 0.01 -> N=1590 (y2022pantheonSHOES/data.py:25);
    cut=False keeps all 1701 rows (the size named in BASELINE.json's metric)."""
    d = _load("data_pantheon_plus.npz", root)
    cov = synthetic_sn_covariance(d["m_b_corr_err_DIAG"])
    keep = np.where(d["zHD"] > 0.01)[0] if cut else np.arange(d["zHD"].size)
    return d["zHD"][keep], d["zHEL"][keep], d["m_b_corr"][keep], cov[np.ix_(keep, keep)]


def des_dovekie(root=None):
    """(z_cmb, z_hel, mu, cov) sorted by zHD, N=1820 (y2025DESdovekie/data.py:25-35)."""
    d = _load("data_des_dovekie.npz", root)
    cov = synthetic_sn_covariance(d["MUERR"])
    o = np.argsort(d["zHD"])
    return d["zHD"][o], d["zHEL"][o], d["MU"][o], cov[o, :][:, o]


def union3_1(root=None):
    """(z_cmb, z_hel, mu, cov) 22 bins (y2026union3_1/data.py)."""
    d = _load("data_union3_1.npz", root)
    return d["zcmb"], d["zhel"], d["mb"], d["cov"]


def desi_dr2(root=None):
    """(z, value, quantity, cov) DESI DR2 BAO, 13 points (y2025BAO/data.py)."""
    d = _load("data_desi_bao.npz", root)
    return d["dr2_z"], d["dr2_value"], d["dr2_quantity"], d["dr2_cov"]


def desi_fs_lya(root=None):
    """(z, value, quantity, cov) DESI DR2 + full-shape Lya, 14 points (y2025BAO/data_fs_lya.py)."""
    d = _load("data_desi_bao.npz", root)
    return d["fs_lya_z"], d["fs_lya_value"], d["fs_lya_quantity"], d["fs_lya_cov"]


def pantheon_plus_positions(cut=True, root=None):
    """(RA, DEC, IDSURVEY) of the Pantheon+ rows (y2022pantheonSHOES/data.py:38-48)."""
    d = _load("data_pantheon_plus.npz", root)
    keep = np.where(d["zHD"] > 0.01)[0] if cut else np.arange(d["zHD"].size)
    return d["RA"][keep], d["DEC"][keep], d["IDSURVEY"][keep]


def cosmic_chronometers(root=None):
    """(z, H, cov) (y2005cc/data.py:5-40)."""
    d = _load("data_cc.npz", root)
    return d["z"], d["H"], d["cov"]


def pantheon_plus_shoes(root=None):
    """(z_cmb, z_hel, m_b, ceph_dist, cov): calibrators kept at any redshift, N=1657 (y2022pantheonSHOES/data_shoes.py:24-39)."""
    d = _load("data_pantheon_plus.npz", root)
    cov = synthetic_sn_covariance(d["m_b_corr_err_DIAG"])
    sel = np.where((d["IS_CALIBRATOR"] == 1) | (d["zHD"] > 0.01))[0]
    return d["zHD"][sel], d["zHEL"][sel], d["m_b_corr"][sel], d["CEPH_DIST"][sel], cov[np.ix_(sel, sel)]
