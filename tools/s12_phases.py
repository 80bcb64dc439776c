"""Stage-1/2 time of a configuration with the grid pass / SN pass switched off (dbg bits 1 / 2): a crude phase split."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import golden, spec
from cosmology_model_fit_b200 import Engine
from cosmology_model_fit_b200.synthetic import uniform_theta
B = 65536
for name in sys.argv[1:] or ["bao_desi_cmb_pantheon", "bao_desi_des5y_bbn_theta_star", "bao_desi", "cmb_cmb", "sn_pantheon"]:
    theta = uniform_theta(golden(name)["bounds"], B, seed=5)
    with Engine(spec(name)) as e:
        out = []
        for dbg in (0, 1, 2, 3):
            e.set_option("dbg", dbg)
            for _ in range(3):
                e.chi_squared(theta)
            out.append(e.timing_history(2).mean(0)[0])
    print(f"{name:34s} stage12 ms: full {out[0]:.3f}  no-grid {out[1]:.3f}  no-SN {out[2]:.3f}  neither {out[3]:.3f}", flush=True)
