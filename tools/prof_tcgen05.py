"""One full-size evaluation per engine setting, for ncu: python tools/prof_tcgen05.py [slices] [B] [n_evals]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosmology_model_fit_b200 import Engine, datasets, fits
from cosmology_model_fit_b200.synthetic import uniform_theta

S = int(sys.argv[1]) if len(sys.argv) > 1 else 6
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
n = int(sys.argv[3]) if len(sys.argv) > 3 else 3
spec = fits.sn_pantheon(datasets.pantheon_plus(cut=False))
theta = uniform_theta(spec.bounds, B, seed=42)
with Engine(spec, device=0) as eng:
    if S:
        eng.set_option("chi2_engine", 1)
        eng.set_option("chi2_slices", S)
    for k, v in (a.split("=") for a in sys.argv[4:]):
        eng.set_option(k, int(v))
    for _ in range(n):
        out = eng.chi_squared(theta)
    print("slices", S, "timing", eng.last_timing(), "split", eng.stage3_split(), "sum", float(out.sum()))
