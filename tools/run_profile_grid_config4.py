#!/usr/bin/env python3
"""BASELINE.json config 4: profile-likelihood grid over (w0, wa, Om, H0) x Pantheon+ 1701^2 covariance with the magnitude
offset profiled analytically (not in the reference: parity unpinned, closed forms checked in
tests/test_gpu_parity.py::test_profile_grid_closed_forms).  Late-time w0wa (CPL) model, SN only: H0 only shifts the
effective offset, so the H0 axis costs no contraction rows: n_w0 x n_wa x n_Om rows go through the engine's two-dot
epilogue (y.y, y.u), the H0 axis is applied in closed form on the host.

    python tools/run_profile_grid_config4.py [n_per_axis=100] [n_h0=100]      # 100^3 rows x 100 = 1e8 grid points
"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cosmology_model_fit_b200 import Engine, datasets, fits, spec as S
from cosmology_model_fit_b200.profile import offset_profile

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n_h0 = int(sys.argv[2]) if len(sys.argv) > 2 else 100
sn = datasets.pantheon_plus(cut=False)
sp = S.LikelihoodSpec(ndim=5, family=S.FAMILY_LATE, de_model=S.DE_CPL, col_H0=1, col_Om=2, col_w0=3, col_wa=4,
                      z_grid=S.LikelihoodSpec.make_grid(float(np.max(sn[0]))))
fits._sn_block(sp, sn, S.SN_CHOLESKY, 0.0, 0, None)       # theta = (M, H0, Om, w0, wa), no velocity term
w0 = np.linspace(-1.5, -0.3, n); wa = np.linspace(-2.5, 0.2, n); om = np.linspace(0.15, 0.55, n)
h0 = np.linspace(60.0, 80.0, n_h0)
W0, WA, OM = np.meshgrid(w0, wa, om, indexing="ij")
keep = (W0 + WA < 0.0).ravel()                              # w0 + wa < 0 (the CPL guard of bao/desi_fs_lya_cmb.py:119-120)
theta = np.empty((n**3, 5))
theta[:, 0] = 0.0; theta[:, 1] = 70.0; theta[:, 2] = OM.ravel(); theta[:, 3] = W0.ravel(); theta[:, 4] = WA.ravel()
rows = theta[keep]
with Engine(sp) as eng:
    eng.sn_moments(rows[:4096])                              # warm-up: workspace, digit planes of W
    t0 = time.perf_counter()
    mom = np.empty((len(rows), 3))
    chunk = 65536
    for i in range(0, len(rows), chunk):
        mom[i:i + chunk] = eng.sn_moments(rows[i:i + chunk])
    t_gpu = time.perf_counter() - t0
    chi2_rows, mstar = offset_profile(mom, "profile")
    # H0 axis in closed form: the profiled chi2 does not depend on H0, the profiled offset shifts by 5 log10(H0 / 70)
    t1 = time.perf_counter()
    chi2 = np.full(n**3, np.inf); chi2[keep] = chi2_rows
    grid = np.broadcast_to(chi2.reshape(n, n, n, 1), (n, n, n, n_h0))
    m_grid_min = float(np.min(mstar)) + 5.0 * np.log10(h0 / 70.0)
    t_host = time.perf_counter() - t1
    best = int(np.argmin(chi2_rows))
    # spot check of the closed form against direct chi_squared at the best row for three H0 values
    chk = []
    for h in (62.0, 70.0, 78.0):
        t = rows[best].copy(); t[1] = h; t[0] = mstar[best] + 5.0 * np.log10(h / 70.0)
        chk.append(float(eng.chi_squared(t)))
print(json.dumps({"config": "profile grid (w0, wa, Om, H0) x Pantheon+ N=1701, M profiled analytically", "rows": int(len(rows)),
                  "grid_points": int(len(rows)) * n_h0, "gpu_wall_s": t_gpu, "rows_per_s": len(rows) / t_gpu,
                  "grid_points_per_s": len(rows) * n_h0 / (t_gpu + t_host), "chi2_min": float(chi2_rows[best]),
                  "best": {"Om": rows[best][2], "w0": rows[best][3], "wa": rows[best][4], "M_at_H0_70": float(mstar[best])},
                  "closed_form_check_chi2_at_H0_62_70_78": chk, "grid_shape": [n, n, n, n_h0], "broadcast_view": list(grid.shape)}))
