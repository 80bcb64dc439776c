#!/usr/bin/env python3
"""Two (or more) ranks, one GPU each, under torchrun: the library-owned NCCL path (cl_comm_init / cl_eval_allgather /
cl_eval_grid + cl_grid_allreduce) against single-GPU evaluation of the same rows.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from cosmology_model_fit_b200 import Engine, datasets, fits
from cosmology_model_fit_b200.engine import GRID_PROFILE
from cosmology_model_fit_b200.parallel import ShardedEngine, grid_points, grid_stats_of
from cosmology_model_fit_b200.profile import offset_profile
from cosmology_model_fit_b200.spec import OUT_LOGLIKE, OUT_LOGPROB
from cosmology_model_fit_b200.synthetic import uniform_theta

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")          # only ships the NCCL id: the data path is the library's own communicator
spec = fits.sn_pantheon(datasets.pantheon_plus(cut=False))
sh = ShardedEngine(spec, device=local)
eng = sh.engine
res = {"world": world}
for B in (65536, 1000, 7):
    theta = uniform_theta(spec.bounds, B, seed=77)
    want = eng.log_probability(theta)                       # every rank evaluates everything on its own GPU
    got = sh.log_probability(theta)
    assert np.array_equal(got, want), (rank, B)
    only = sh.log_probability(theta, root=0)
    assert (only is None) == (rank != 0) and (only is None or np.array_equal(only, want))
axes = {2: (0.15, 0.55, 60), 3: (-2.5, 2.5, 50)}
fixed = {0: 0.0, 1: 70.0}
stats, vals, (first, count) = sh.grid(axes, fixed, GRID_PROFILE, want_values=True)
pts = grid_points(axes, fixed, 4, 0, 3000)
full, _ = offset_profile(eng.sn_moments(pts), "profile")
ref = grid_stats_of(full, 0)
assert np.array_equal(vals, full[first:first + count])
assert stats["index"] == ref["index"] and stats["best"] == ref["best"] and abs(stats["log_sum"] - ref["log_sum"]) < 1e-10 and stats["count"] == 3000
# timing of the host-buffer collective path at the headline shape, page-locked buffers, results on rank 0 only / everywhere
B = 65536 * world
theta = eng.pinned_empty((B, 4)); theta[...] = uniform_theta(spec.bounds, B, seed=5)
out = eng.pinned_empty((B,))
for mode, root in (("all ranks receive", None), ("root only", 0)):
    for _ in range(3):
        sh.log_likelihood(theta, out=out if (root is None or rank == 0) else None, root=root)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        sh.log_likelihood(theta, out=out if (root is None or rank == 0) else None, root=root)
    torch.cuda.synchronize(); dist.barrier()
    res[mode] = B * 10 / (time.perf_counter() - t0)
if rank == 0:
    print(json.dumps({"multi_gpu_check": "ok", **res}), flush=True)
dist.destroy_process_group()
