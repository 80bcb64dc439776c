"""Bring-up of the tcgen05 digit-plane engine: parity against the FP64 DMMA engine and the oracle, and stage timings.
Run on a B200: python tools/try_tcgen05.py [B]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosmology_model_fit_b200 import Engine, datasets, fits
from cosmology_model_fit_b200.synthetic import uniform_theta
import oracle.oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
spec = fits.sn_pantheon(datasets.pantheon_plus(cut=False))
theta = uniform_theta(spec.bounds, B, seed=42)
small = theta[:777]
with Engine(spec, device=0) as eng:
    ref_small = eng.chi_squared(small)
    want = O.Oracle(spec).chi_squared(small[:64], nthreads=0)
    print("DMMA vs oracle (64 rows): max |d| = %.3e" % np.max(np.abs(ref_small[:64] - want)), flush=True)
    ref = eng.chi_squared(theta)
    for _ in range(3):
        eng.chi_squared(theta)
    print("DMMA timing:", eng.last_timing(), flush=True)
    eng.set_option("chi2_engine", 1)
    for S in (6, 7, 5):
        eng.set_option("chi2_slices", S)
        for trim in (0, 1):
            eng.set_option("gemm_diag_skip", trim)
            got_small = eng.chi_squared(small)
            d = np.abs(got_small - ref_small)
            print(f"S={S} trim={trim} ragged batch 777: max |d chi2| = {d.max():.3e}, max rel = {(d / ref_small).max():.3e}", flush=True)
            got = eng.chi_squared(theta)
            d = np.abs(got - ref)
            print(f"S={S} trim={trim} B={B}: max |d chi2| = {d.max():.3e} (chi2 up to {ref.max():.3e}), max rel = {(d / ref).max():.3e}", flush=True)
            for _ in range(3):
                eng.chi_squared(theta)
            print("   timing:", eng.last_timing(), "split (slice, mma):", eng.stage3_split(), flush=True)
