"""Wall-clock latency of Engine.log_probability for small batches (what emcee / nautilus with stock settings send),
with the ordinary launches and with the CUDA-graph replay of cl_eval (option cuda_graphs)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import golden, spec
from cosmology_model_fit_b200 import Engine
from cosmology_model_fit_b200.synthetic import uniform_theta
for name in ("sn_pantheon", "bao_desi_cmb_pantheon", "bao_desi_cmb_union3"):
    theta = uniform_theta(golden(name)["bounds"], 4096, seed=5)
    with Engine(spec(name)) as e:
        for eng, sl in ((1, 7), (0, 7)):
            for graphs in (0, 1):
                e.set_option("chi2_engine", eng); e.set_option("chi2_slices", sl); e.set_option("cuda_graphs", graphs)
                row = []
                for B in (1, 16, 75, 128, 1024, 4096):
                    t = theta[:B]
                    for _ in range(5):
                        e.log_probability(t)
                    t0 = time.perf_counter()
                    n = 100
                    for _ in range(n):
                        e.log_probability(t)
                    row.append((time.perf_counter() - t0) / n * 1e6)
                print(f"{name:24s} engine={'tcgen05' if eng else 'dmma'} graphs={graphs}: us per call at B=1,16,75,128,1024,4096: " + ", ".join(f"{x:.0f}" for x in row)
                      + f"  {e.graph_info()}", flush=True)
