#!/usr/bin/env python3
"""BASELINE.json config 3: nested-sampling evidence over bao.desi_cmb_pantheon with batched live points on the GPU.
The reference drives nautilus (bao/desi_cmb_pantheon.py:154-170, n_live 6000, ~100 points per likelihood call); nautilus is
not installed here, so the run uses cosmology_model_fit_b200.samplers.NestedSampler with the same uniform priors and
reports ln Z next to the Laplace approximation around the best point.  python tools/run_nested_config3.py [n_live]"""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import spec
from cosmology_model_fit_b200 import Engine
from cosmology_model_fit_b200.samplers import BoxPrior, DeviceProposer, NestedSampler, laplace_log_evidence
from cosmology_model_fit_b200.spec import OUT_LOGLIKE

n_live = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
on_device = (sys.argv[2] if len(sys.argv) > 2 else "device") == "device"   # "host": proposals drawn, transformed and filtered with numpy
bounds = np.array([(-20.0, -19.0), (60.0, 75.0), (0.019, 0.025), (0.09, 0.14), (-3.0, 1.5)])   # M, H0, obh2, och2, v (:146-151, och2 narrowed)
sp = spec("bao_desi_cmb_pantheon")
with Engine(sp) as eng:
    eng.log_likelihood(np.tile(bounds.mean(1), (1024, 1)))   # warm-up: workspace + digit planes of W
    prior = BoxPrior(bounds)
    eng.set_option("max_rows_per_pass", 262144)
    ns = NestedSampler(prior, eng.log_likelihood, n_live=n_live, n_replace=n_live // 4, batch=262144, min_batch=65536, seed=1,
                       proposer=DeviceProposer(prior, eng, OUT_LOGLIKE, seed=1) if on_device else None)
    t0 = time.perf_counter()
    res = ns.run(dlogz=0.01)
    dt = time.perf_counter() - t0
    best = res["samples"][np.argmax(res["log_like"])]
    lz, _ = laplace_log_evidence(eng.log_likelihood, best, step=1e-3, scales=np.array([0.01, 0.3, 1e-4, 1e-3, 0.3]))
    lz -= np.sum(np.log(bounds[:, 1] - bounds[:, 0]))
mean = (res["weights"][:, None] * res["samples"]).sum(0)
std = np.sqrt((res["weights"][:, None] * (res["samples"] - mean) ** 2).sum(0))
print(json.dumps({"config": "bao_desi_cmb_pantheon nested sampling", "proposals": "device (cl_propose_eval)" if on_device else "host (numpy)", "n_live": n_live, "logz": res["logz"], "logz_err": res["logz_err"],
                  "laplace_logz": lz, "information_nats": res["information"], "iterations": res["n_iter"], "likelihood_calls": res["n_calls"],
                  "likelihood_evals": res["n_evals"], "rows_per_call": res["n_evals"] / res["n_calls"], "wall_s": dt,
                  "evals_per_s_wall": res["n_evals"] / dt, "posterior_mean": mean.tolist(), "posterior_std": std.tolist(),
                  "chi2_min": float(-2 * res["log_like"].max())}))
