"""Small all-paths run for compute-sanitizer (memcheck): every golden configuration, 600 rows, few stage-1/2 CTAs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from cases import SPECS, golden, spec
from cosmology_model_fit_b200 import Engine
from cosmology_model_fit_b200.synthetic import uniform_theta
for name in SPECS:
    g = golden(name)
    th = uniform_theta(g["bounds"], 300, seed=1)
    with Engine(spec(name)) as e:
        e.set_option("stage12_ctas", 37)
        c = e.chi_squared(th); e.log_probability(th); e.components(th[:50])
        if spec(name).z_grid is not None and (spec(name).sn_zcmb is not None or spec(name).bao_z is not None):
            e.distances(th[:8], np.linspace(-0.01, 2.6, 33))
        if spec(name).sn_zcmb is not None:
            e.sn_residuals(th[:8])
            if len(spec(name).sn_zcmb) > 64 and spec(name).col_offset >= 0:
                e.sn_moments(th[:130])
    print(name, "ok", float(np.nanmax(c)))
