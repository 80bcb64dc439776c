#!/usr/bin/env python3
"""Stage timings of every golden configuration at a given batch size (host-buffer path, library CUDA events)."""
import sys, os, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import SPECS, golden, spec
from cosmology_model_fit_b200 import Engine
from cosmology_model_fit_b200.synthetic import uniform_theta

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
names = sys.argv[2:] or list(SPECS)
rows = []
for name in names:
    sp = spec(name)
    g = golden(name)
    theta = uniform_theta(g["bounds"], B, seed=5)
    with Engine(sp) as e:
        for _ in range(2):
            e.chi_squared(theta)
        for _ in range(3):
            e.chi_squared(theta)
        h = e.timing_history(3).mean(0)
    n_sn = 0 if sp.sn_zcmb is None else len(sp.sn_zcmb)
    rows.append(dict(config=name, n_sn=n_sn, family=sp.family, de=sp.de_model, stage12_ms=h[0], stage3_ms=h[1], total_ms=h[3],
                     evals_per_s=B / (h[3] * 1e-3)))
    print(json.dumps(rows[-1]), flush=True)
