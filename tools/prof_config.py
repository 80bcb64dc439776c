"""A few full-size evaluations of one golden configuration, for ncu: python tools/prof_config.py <name> [B] [n_evals] [opt=value ...]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import golden, spec
from cosmology_model_fit_b200 import Engine
from cosmology_model_fit_b200.synthetic import uniform_theta

name = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4
theta = uniform_theta(golden(name)["bounds"], B, seed=5)
with Engine(spec(name)) as eng:
    for k, v in (a.split("=") for a in sys.argv[4:]):
        eng.set_option(k, int(v))
    for _ in range(n):
        out = eng.chi_squared(theta)
    print(name, "timing", eng.last_timing(), "sum", float(np.nansum(out)))
