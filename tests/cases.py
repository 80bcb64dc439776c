"""The golden cases: (spec builder, golden file) for every reference script the fixtures cover."""
import os

import numpy as np

from cosmology_model_fit_b200 import datasets, fits
from cosmology_model_fit_b200 import spec as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

_cache = {}


def golden(name):
    if name not in _cache:
        _cache[name] = dict(np.load(os.path.join(GOLDEN, f"golden_{name}.npz")))
    return _cache[name]


def _memo(fn):
    def wrapper():
        key = "data:" + fn.__name__
        if key not in _cache:
            _cache[key] = fn()
        return _cache[key]
    return wrapper


pantheon = _memo(lambda: datasets.pantheon_plus())
pantheon.__name__ = "pantheon"
pantheon = _memo(datasets.pantheon_plus)
des = _memo(datasets.des_dovekie)
union3 = _memo(datasets.union3_1)
desi = _memo(datasets.desi_dr2)
desi_fs = _memo(datasets.desi_fs_lya)
cc = _memo(datasets.cosmic_chronometers)
pantheon_pos = _memo(datasets.pantheon_plus_positions)
pantheon_shoes = _memo(datasets.pantheon_plus_shoes)

SPECS = {
    "sn_pantheon": lambda: fits.sn_pantheon(pantheon()),
    "sn_union3_1": lambda: fits.sn_union3_1(union3()),
    "sn_des5y": lambda: fits.sn_des5y(des()),
    "bao_desi": lambda: fits.bao_desi(desi()),
    "bao_desi_cmb_union3": lambda: fits.bao_desi_cmb_union3(union3(), desi_fs()),
    "bao_desi_fs_lya_cmb": lambda: fits.bao_desi_fs_lya_cmb(desi_fs()),
    "cmb_cmb": lambda: fits.cmb_cmb(),
    "cmb_cmb_act": lambda: fits.cmb_cmb(S.cmb_act()),
    "cmb_cmb_planck_lens": lambda: fits.cmb_cmb(S.cmb_planck_lens()),
    "cmb_cmb_planck": lambda: fits.cmb_cmb(S.cmb_planck()),
    "bao_desi_des5y_bbn_theta_star": lambda: fits.bao_desi_des5y_bbn_theta_star(des(), desi()),
    "bao_desi_cmb_pantheon": lambda: fits.bao_desi_cmb_pantheon(pantheon(), desi()),
    "bao_desi_cmb_des5y": lambda: fits.bao_desi_cmb_des5y(des(), desi_fs()),
    "ohd_cc": lambda: fits.ohd_cc(cc()),
    "bao_desi_bbn": lambda: fits.bao_desi_bbn(desi()),
    "bao_desi_pantheon_cc": lambda: fits.bao_desi_pantheon_cc(pantheon(), desi(), cc()),
    "sn_pantheon_dipole_xyz": lambda: fits.sn_pantheon_dipole_xyz(pantheon(), *pantheon_pos()),
    "sn_pantheon_and_sh0es": lambda: fits.sn_pantheon_and_sh0es(pantheon_shoes()),
    "bao_desi_cmb_pantheon_H0trgb": lambda: fits.bao_desi_cmb_pantheon_H0trgb(pantheon(), desi()),
    "bao_desi_cmb": lambda: fits.bao_desi_cmb(desi()),
    "sn_pantheon_dipole": lambda: fits.sn_pantheon_dipole(pantheon(), *pantheon_pos()),
    "ohd_cc_des5y": lambda: fits.ohd_cc_des5y(des(), cc()),
    "ohd_cc_union3": lambda: fits.ohd_cc_union3(union3(), cc()),
    "bao_desi_omh2": lambda: fits.bao_desi_omh2(desi()),
    "bao_desi_union3_obh2_theta_star": lambda: fits.bao_desi_union3_obh2_theta_star(union3(), desi()),
    "sn_pantheon_cmb": lambda: fits.sn_pantheon_cmb(pantheon()),
    "sn_des5y_cmb": lambda: fits.sn_des5y_cmb(des()),
    "sn_union3_1_cmb": lambda: fits.sn_union3_1_cmb(union3()),
    "ohd_cc_cmb": lambda: fits.ohd_cc_cmb(cc()),
    "ohd_cc_pantheon": lambda: fits.ohd_cc_pantheon(pantheon(), cc()),
    "bao_desi_fs_lya": lambda: fits.bao_desi_fs_lya(desi_fs()),
    "bao_desi_cc": lambda: fits.bao_desi_cc(desi(), cc()),
    "bao_desi_des5y_rd": lambda: fits.bao_desi_des5y_rd(des(), desi()),
    "bao_desi_union3_rd": lambda: fits.bao_desi_union3_rd(union3(), desi()),
    "bao_desi_pantheon_rd": lambda: fits.bao_desi_pantheon_rd(pantheon(), desi()),
    "bao_desi_bbn_theta_star": lambda: fits.bao_desi_bbn_theta_star(desi()),
    "bao_desi_union3_bbn_theta_star": lambda: fits.bao_desi_union3_bbn_theta_star(union3(), desi_fs()),
    "bao_desi_union3_cc_theta_star": lambda: fits.bao_desi_union3_cc_theta_star(union3(), desi(), cc()),
    "bao_desi_des5y_cc_theta_star": lambda: fits.bao_desi_des5y_cc_theta_star(des(), desi(), cc()),
    "bao_desi_union3_omh2": lambda: fits.bao_desi_union3_omh2(union3(), desi()),
    "bao_desi_des5y_omh2": lambda: fits.bao_desi_des5y_omh2(des(), desi()),
    "bao_desi_union3_omh2_theta_star": lambda: fits.bao_desi_union3_omh2_theta_star(union3(), desi()),
    "bao_desi_pantheon_obh2_theta_star": lambda: fits.bao_desi_pantheon_obh2_theta_star(pantheon(), desi()),
    "bao_desi_des5y_obh2_theta_star": lambda: fits.bao_desi_des5y_obh2_theta_star(des(), desi()),
    "bao_desi_union3_bbn": lambda: fits.bao_desi_union3_bbn(union3(), desi()),
    "bao_desi_des5y_bbn": lambda: fits.bao_desi_des5y_bbn(des(), desi()),
    "bao_desi_des5y_H0trgb": lambda: fits.bao_desi_des5y_H0trgb(des(), desi()),
    "bao_desi_cmb_des5y_H0trgb": lambda: fits.bao_desi_cmb_des5y_H0trgb(des(), desi()),
    "bao_desi_cmb_union3_H0trgb": lambda: fits.bao_desi_cmb_union3_H0trgb(union3(), desi()),
    "bao_desi_des5y_cc": lambda: fits.bao_desi_des5y_cc(des(), desi_fs(), cc()),
    "bao_desi_fs_lya_union3_cc": lambda: fits.bao_desi_fs_lya_union3_cc(union3(), desi_fs(), cc()),
}

#: cases whose golden file has a plain chi2[n] for theta[n]
CHI2_CASES = ["sn_pantheon", "sn_union3_1", "sn_des5y", "bao_desi", "bao_desi_cmb_union3",
              "bao_desi_des5y_bbn_theta_star", "bao_desi_cmb_pantheon", "bao_desi_cmb_des5y",
              "ohd_cc", "bao_desi_bbn", "bao_desi_pantheon_cc", "sn_pantheon_dipole_xyz",
              "sn_pantheon_and_sh0es", "bao_desi_cmb_pantheon_H0trgb", "bao_desi_cmb", "bao_desi_union3_obh2_theta_star",
              "sn_pantheon_dipole", "ohd_cc_des5y", "ohd_cc_union3", "bao_desi_omh2",
              "sn_pantheon_cmb", "sn_des5y_cmb", "sn_union3_1_cmb", "ohd_cc_cmb", "ohd_cc_pantheon",
              "bao_desi_fs_lya", "bao_desi_cc", "bao_desi_des5y_rd", "bao_desi_union3_rd", "bao_desi_pantheon_rd",
              "bao_desi_bbn_theta_star", "bao_desi_union3_bbn_theta_star", "bao_desi_union3_cc_theta_star",
              "bao_desi_des5y_cc_theta_star", "bao_desi_union3_omh2", "bao_desi_des5y_omh2", "bao_desi_union3_omh2_theta_star",
              "bao_desi_pantheon_obh2_theta_star", "bao_desi_des5y_obh2_theta_star",
              "bao_desi_union3_bbn", "bao_desi_des5y_bbn", "bao_desi_des5y_H0trgb", "bao_desi_cmb_des5y_H0trgb",
              "bao_desi_cmb_union3_H0trgb", "bao_desi_des5y_cc", "bao_desi_fs_lya_union3_cc"]

#: cases generated with the generic helper whose golden file also holds log_likelihood / log_probability rows
GENERIC_LOGLIKE_CASES = ["ohd_cc_cmb", "ohd_cc_pantheon", "bao_desi_cc", "bao_desi_union3_cc_theta_star", "bao_desi_des5y_cc_theta_star",
                         "bao_desi_des5y_cc", "bao_desi_fs_lya_union3_cc"]
GENERIC_LOGP_CASES = ["sn_pantheon_cmb", "ohd_cc_cmb", "ohd_cc_pantheon", "bao_desi_cc", "bao_desi_pantheon_rd",
                      "bao_desi_des5y_cc_theta_star", "bao_desi_pantheon_obh2_theta_star", "bao_desi_des5y_obh2_theta_star",
                      "bao_desi_des5y_H0trgb", "bao_desi_des5y_cc"]


def spec(name):
    key = "spec:" + name
    if key not in _cache:
        _cache[key] = SPECS[name]()
    return _cache[key]


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
