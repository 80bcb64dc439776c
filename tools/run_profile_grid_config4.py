#!/usr/bin/env python3
"""BASELINE.json config 4: profile-likelihood grid over (w0, wa, Om, H0) x Pantheon+ 1701^2 covariance, sharded over the GPUs
of one box, with the magnitude offset profiled analytically (not in the reference: parity unpinned, closed forms
checked in tests/test_gpu_parity.py and tests/test_gpu_multi.py).  Late-time w0wa (CPL) model, SN only.

Every rank takes a contiguous slice of the flattened grid, generates its parameter vectors ON THE DEVICE (cl_eval_grid),
reduces its slice to (min chi2, argmin, log-sum-exp) on the device, and the ranks combine those with one 32-byte all-gather
(cl_grid_allreduce): neither theta nor the chi2 values cross PCIe.

    python tools/run_profile_grid_config4.py [n_per_axis=100] [n_h0=100] [h0=explicit|analytic]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 tools/run_profile_grid_config4.py

h0=explicit (default): H0 is a real grid axis, n^3 x n_h0 contraction rows (1e8 at the defaults).
h0=analytic: in the late-time family H0 only shifts the effective offset, so the profiled chi2 is flat in H0: n^3 rows, the H0
axis applied in closed form (the same 1e8 grid points from 1e6 rows).
"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cosmology_model_fit_b200 import Engine, datasets, fits, spec as S
from cosmology_model_fit_b200.engine import GRID_PROFILE
from cosmology_model_fit_b200.parallel import ShardedEngine, grid_points
from cosmology_model_fit_b200.profile import offset_profile

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n_h0 = int(sys.argv[2]) if len(sys.argv) > 2 else 100
h0_mode = sys.argv[3].split("=")[-1] if len(sys.argv) > 3 else "explicit"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")      # ships the NCCL id; the data path is the library's own communicator

sn = datasets.pantheon_plus(cut=False)
sp = S.LikelihoodSpec(ndim=5, family=S.FAMILY_LATE, de_model=S.DE_CPL, col_H0=1, col_Om=2, col_w0=3, col_wa=4,
                      z_grid=S.LikelihoodSpec.make_grid(float(np.max(sn[0]))))
fits._sn_block(sp, sn, S.SN_CHOLESKY, 0.0, 0, None)       # theta = (M, H0, Om, w0, wa), no velocity term
axes = {3: (-1.5, -0.3, n), 4: (-2.5, 0.2, n), 2: (0.15, 0.55, n)}       # w0 (slowest), wa, Om
fixed = {0: 0.0, 1: 70.0}
if h0_mode == "explicit":
    axes[1] = (60.0, 80.0, n_h0)                                          # H0 fastest
    fixed = {0: 0.0}
sh = ShardedEngine(sp, device=local, rank=rank, world=world) if world == 1 else ShardedEngine(sp, device=local)
eng = sh.engine
warm = eng.make_grid(axes, fixed)
eng.eval_grid(warm, 0, 65536, GRID_PROFILE, allreduce=world > 1)          # workspace at its full size, digit planes of W, NCCL connections
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
stats, _, (first, count) = sh.grid(axes, fixed, GRID_PROFILE)
wall = time.perf_counter() - t0
if world > 1:
    walls = [None] * world
    dist.all_gather_object(walls, wall)
    wall = max(walls)
rows = stats["count"]
points = rows * (n_h0 if h0_mode == "analytic" else 1)
log_sum = stats["log_sum"] + (np.log(n_h0) if h0_mode == "analytic" else 0.0)   # flat in H0 once M is profiled
best = grid_points(axes, fixed, 5, stats["index"], 1)[0]
res = {"config": "profile grid (w0, wa, Om, H0) x Pantheon+ N=1701, M profiled analytically, theta generated and reduced on the device",
       "h0_axis": h0_mode, "n_gpus": world, "rows": int(rows), "grid_points": int(points), "wall_s": wall, "rows_per_s": rows / wall,
       "grid_points_per_s": points / wall, "chi2_min": stats["best"], "argmin_index": int(stats["index"]),
       "log_sum_exp_minus_half_chi2": log_sum,
       "best": {"H0": best[1], "Om": best[2], "w0": best[3], "wa": best[4]}, "grid_shape": [n, n, n, n_h0],
       "rows_this_rank": int(count), "guard_rows_recomputed_this_rank": eng.guard_info()["rows_total"]}
if rank == 0:
    # spot check: the profiled chi2 at the best point against direct chi_squared with M = M*(best) (closed form, profile.py)
    mom = eng.sn_moments(best[None, :])
    chi2p, mstar = offset_profile(mom, "profile")
    t = best.copy(); t[0] = mstar[0]
    res["check_direct_chi2_at_best"] = float(eng.chi_squared(t))
    res["check_profiled_chi2_at_best"] = float(chi2p[0])
    print(json.dumps(res), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
