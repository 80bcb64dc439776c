#!/usr/bin/env python3
"""bench.py — likelihood evals/sec on the headline workload of BASELINE.json.

Workload W0: sn/pantheon.py-shaped likelihood — flat LCDM + z_turn=0.15 velocity step, theta =
(M, H0, Om, v) — on the real Pantheon+ redshifts/magnitudes with the full N x N covariance (synthetic SPD
stand-in: the real blob is not shipped with the reference), N = 1701 (all rows; `--n-sn 1590` gives the z>0.01
cut the reference actually fits), batch B = 65536 parameter vectors per GPU per step.

One "step" = one pass of the whole hot path (stage 1+2 Friedmann distances + residuals, stage 3 chi-squared
contraction, finalize) over one batch.  Stage 3 runs on the library's default engine (tcgen05 kind::i8 digit planes, 7
planes = every bit of the FP64 rows); the FP64 DMMA engine and the 6-plane setting are timed beside it ("engines").  `value` is device-resident throughput (inputs already in HBM), `e2e` is
the same through the C ABI with host buffers (H2D + D2H inside the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one process per GPU); rows are sharded across ranks (weak scaling, B per GPU
fixed) and the per-rank log-likelihoods are all-gathered with NCCL every step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "likelihood evals/sec (Pantheon+ full-cov, batch 65536)"
UNIT = "evals/s"
#: tcgen05.mma kind::i8 issue rate measured on this pool's B200 (profiles/r02_ubench_umma_i8.log), int8 TOP/s
I8_PEAK_TOPS = 4485.0
#: DRAM bytes per launch of k_chi2_ozaki<S> from ncu --set full (profiles/), keyed by (planes, N, B)
#: FP64 pipe of one B200: 64 DFMA per clock and SM (DMMA issue peak measured 37.1 TFLOP/s, profiles/r01_ubench_fp64.log)
FP64_PIPE_TFLOPS = 37.1
#: algorithmic FP64 flops of stage 1+2 per evaluation: 9 per grid node + 40 per supernova
S12_FLOPS_PER_EVAL = lambda n_sn, n_grid: 9.0 * n_grid + 40.0 * n_sn
OZ_DRAM_BYTES = {(7, 1701, 65536): 0.916e9}   # profiles/r05_ncu_full_summary.txt: 0.884e9 read + 0.031e9 written


def build_spec(n_sn):
    from cosmology_model_fit_b200 import datasets, fits
    sn = datasets.pantheon_plus(cut=(n_sn == 1590))
    assert sn[0].size == n_sn, sn[0].size
    return fits.sn_pantheon(sn)


def theta_batch(spec, B, seed):
    from cosmology_model_fit_b200.synthetic import uniform_theta
    return np.ascontiguousarray(uniform_theta(spec.bounds, B, seed=seed))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def dgemm_peak_tflops(torch, dev):
    """FP64 tensor-pipe denominator: cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64
    figure)."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    for _ in range(2):
        a @ b
    torch.cuda.synchronize(dev)
    best = 1e30
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize(dev)
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * n**3 / (best * 1e-3) / 1e12


def cpu_baseline(spec, theta, seconds=12.0):
    """The oracle port (oracle/cosmo_oracle.c) on all host cores, on a bounded sample of the same batch."""
    import oracle.oracle as O
    orc = O.Oracle(spec, grid_builds=2)  # sn/pantheon.py rebuilds the z-grid twice per evaluation (:59-60,49)
    cores = O.max_threads()
    n0 = min(len(theta), 16 * cores)
    t0 = time.perf_counter(); orc.chi_squared(theta[:n0], nthreads=0); dt = time.perf_counter() - t0
    n = int(min(len(theta), max(n0, n0 * seconds / max(dt, 1e-3))))
    t0 = time.perf_counter(); orc.chi_squared(theta[:n], nthreads=0); dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {n} rows of the step-0 batch, oracle/cosmo_oracle.c with {cores} threads, {dt:.1f} s"}


def load_numba_reference(spec, n_sn):
    """The reference's OWN implementation of the path: sn/pantheon.py (numba) imported from oracle/_ref (unmodified sources
    staged by oracle/make_ref.py), with its data loader pre-seeded (the covariance blob is not in the reference checkout:
    same real z / m_b and synthetic covariance as every other leg of this benchmark), wrapped in the reference's own batch
    pattern - `@njit(parallel=True)` + `prange` over rows (bao/desi.py:100-106).  Returns (batch_fn, threads) or raises."""
    import types
    import numba
    from numba import njit, prange
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "sn", "pantheon.py")):
        raise RuntimeError("oracle/_ref is not staged (python oracle/make_ref.py needs the reference checkout)")
    from cosmology_model_fit_b200 import datasets
    z, zh, mb, cov = datasets.pantheon_plus(cut=(n_sn == 1590))
    stub = types.ModuleType("y2022pantheonSHOES.data")
    stub.get_data = lambda: ("Pantheon+ (2022)", z, zh, mb, cov)
    pkg = types.ModuleType("y2022pantheonSHOES")
    pkg.data = stub
    sys.modules["y2022pantheonSHOES"], sys.modules["y2022pantheonSHOES.data"] = pkg, stub
    sys.path.insert(0, ref_dir)
    import importlib
    ref = importlib.import_module("sn.pantheon")
    ref_chi2 = ref.chi_squared

    @njit(parallel=True)
    def chi2_batch(batch):
        n = batch.shape[0]
        out = np.empty(n, dtype=np.float64)
        for i in prange(n):
            out[i] = ref_chi2(batch[i])
        return out

    return chi2_batch, numba.get_num_threads()


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on all host cores, bounded sample per step.  The real
    thing when it can run here - sn/pantheon.py under numba from oracle/_ref (kind "reference") - else the C port of the same
    algorithm (oracle/cosmo_oracle.c, kind "port")."""
    if rank != 0:
        return 0
    import oracle.oracle as O
    spec = build_spec(args.n_sn)
    orc = O.Oracle(spec, grid_builds=2)
    theta = theta_batch(spec, args.batch, 1000)
    kind, why = "port", None
    try:
        fn, cores = load_numba_reference(spec, args.n_sn)
        got = fn(theta[:32])                                 # JIT compilation + parity of the two CPU implementations
        want = orc.chi_squared(theta[:32], nthreads=0)
        if not np.all(np.abs(got - want) <= np.maximum(1e-6, 1e-12 * np.abs(want))):
            raise RuntimeError(f"numba reference and C port disagree: {np.max(np.abs(got - want))}")
        kind = "reference"
        evaluate = lambda th: fn(th)
    except Exception as e:      # numba missing, oracle/_ref not staged, ...: the port stands in and the line says so
        why = f"{type(e).__name__}: {e}"
        cores = O.max_threads()
        evaluate = lambda th: orc.chi_squared(th, nthreads=0)
    n = min(args.batch, max(256, 64 * cores))
    for _ in range(max(1, min(args.warmup, 2))):
        evaluate(theta[:n])
    t0 = time.perf_counter()
    for k in range(args.steps):
        off = (k * n) % max(1, args.batch - n)
        evaluate(theta[off:off + n])
    dt = time.perf_counter() - t0
    val = args.steps * n / dt
    what = ("sn/pantheon.py chi_squared (numba @njit, unmodified, oracle/_ref) under @njit(parallel=True) prange as in bao/desi.py:100-106"
            if kind == "reference" else "oracle/cosmo_oracle.c (C port of the same algorithm)")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, spec, world, sample_rows=n),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{n} rows per step of the same workload (bounded sample), {cores} threads: {what}"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if why:
        line["cpu_baseline"]["fallback_reason"] = why
    if world > 1:
        line["note"] = f"one host ({cores} threads) against {world} GPUs: rank 0 alone runs the CPU arm"
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, spec, world, sample_rows=None):
    cfg = {"workload": f"W0 sn.pantheon shape: flat LCDM + v-step(z_turn=0.15), theta=(M,H0,Om,v), Pantheon+ N={args.n_sn} "
                       f"full covariance (synthetic SPD stand-in), z-grid 4000, batch {args.batch} per GPU",
           "n_sn": args.n_sn, "batch_per_gpu": args.batch, "global_batch": args.batch * world, "n_grid": int(spec.z_grid.size),
           "parallelism": f"theta rows sharded over {world} GPU(s), statics replicated, NCCL all-gather of logL (the library's own communicator)",
           "l2": "no explicit flush: each step writes+reads the int8 digit planes of the residual rows (7 B*N = 0.8 GB at B = 65536) which "
                 "exceed the 126 MB L2; theta batches rotate between steps"}
    if sample_rows is not None:
        cfg["cpu_sample_rows_per_step"] = sample_rows
    return cfg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--n-sn", type=int, default=1701, choices=[1590, 1701])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value (tuning experiments)")
    ap.add_argument("--engine", default="tcgen05", choices=["tcgen05", "dmma"], help="stage-3 engine of the headline run")
    ap.add_argument("--slices", type=int, default=7, choices=[5, 6, 7], help="int8 digit planes of the tcgen05 engine")
    ap.add_argument("--no-alt", action="store_true", help="skip timing the other stage-3 engines")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch rows per GPU (default, the metric's shape); strong: --batch rows in total, split over the GPUs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.global_batch = args.batch * world if args.scaling == "weak" else args.batch
    if args.scaling == "strong":
        if args.batch % world:
            raise SystemExit("--scaling strong needs --batch divisible by the number of GPUs")
        args.batch //= world          # rows per GPU from here on
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from cosmology_model_fit_b200 import Engine
    from cosmology_model_fit_b200.spec import OUT_LOGLIKE

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the engine has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    spec = build_spec(args.n_sn)
    eng = Engine(spec, device=local_rank)
    eng.set_option("chi2_engine", 1 if args.engine == "tcgen05" else 0)
    eng.set_option("chi2_slices", args.slices)
    for kv in args.opt:
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    B, nd = args.batch, spec.ndim
    n_rot = 4  # distinct theta batches rotated between steps
    host_batches = [theta_batch(spec, B, seed=1000 + rank * 16 + i) for i in range(n_rot)]
    d_theta = [torch.from_numpy(h).to(dev) for h in host_batches]
    d_all = torch.empty(B * world, dtype=torch.float64, device=dev)
    d_out = d_all[rank * B:(rank + 1) * B]      # this rank's results, gathered in place
    stream = torch.cuda.Stream(dev)  # a non-default stream: the kernels, the NCCL gather and the timing events share it
    torch.cuda.set_stream(stream)
    sh = None
    if world > 1:   # the library's own NCCL communicator (cl_comm_init); torch.distributed ships the id and keeps the barriers
        from cosmology_model_fit_b200.parallel import ShardedEngine
        sh = ShardedEngine(spec, device=local_rank, engine=eng)

    def step(i):
        if world > 1:
            eng.eval_allgather_device(d_theta[i % n_rot].data_ptr(), B, nd, OUT_LOGLIKE, d_all.data_ptr(), stream.cuda_stream)
        else:
            eng.eval_device(d_theta[i % n_rot].data_ptr(), B, nd, OUT_LOGLIKE, d_out.data_ptr(), stream.cuda_stream)

    peak_tf = dgemm_peak_tflops(torch, dev)
    torch.cuda.synchronize(dev)
    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize(dev)

    # ---- parity spot check of the benchmarked configuration before timing (rank 0, 64 rows vs the oracle) ----
    parity = None
    if rank == 0:
        import oracle.oracle as O
        got = d_out[:64].cpu().numpy()
        want = O.Oracle(spec).log_likelihood(host_batches[(args.warmup - 1) % n_rot][:64])
        parity = float(np.max(np.abs(-2 * got - -2 * want) / np.maximum(1.0, 1e-6 * np.abs(2 * want)) ))
        if not args.opt and args.slices >= 6 and not np.all(np.abs(2 * got - 2 * want) <= np.maximum(1e-6, 1e-12 * np.abs(2 * want))):
            raise SystemExit(f"parity check failed before timing: max |dchi2| = {np.max(np.abs(2*got-2*want))}")

    # ---- timed region: device-resident ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    e1.record(stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - l0
    hist = eng.timing_history(min(args.steps, 64))
    split = np.atleast_2d(np.asarray(eng.stage3_split(min(args.steps, 64))))
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end from HOST buffers: H2D of theta and D2H of logL inside the timed region, every step ----
    # 1 GPU: Engine.log_likelihood(batch, out=) -> cl_eval, page-locked caller buffers (DMA as they are); the same with ordinary
    #        numpy arrays (one staging copy each way inside the library) is reported beside it as e2e_pageable.
    # N GPUs: ShardedEngine.log_likelihood(global batch, root=0) -> cl_eval_allgather: each rank uploads its row shard,
    #        evaluates, one ncclAllGather inside the library, rank 0 downloads the gathered vector (the master / worker shape
    #        of the reference's Pool.map); the variant where EVERY rank downloads it is reported as e2e_all_ranks.
    e2e_steps = max(3, min(args.steps, 10))

    def time_e2e(call):
        for i in range(2):
            call(i)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            call(i)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return B * world * e2e_steps / dt

    e2e_extra = {}
    if world > 1:
        global_batches = [np.concatenate([theta_batch(spec, B, seed=1000 + r * 16 + i) for r in range(world)]) for i in range(2)]
        pinned_global = []
        for gb in global_batches:   # page-locked global batches and result vector (Engine.pinned_empty): DMA without staging copies
            pb = eng.pinned_empty(gb.shape)
            pb[...] = gb
            pinned_global.append(pb)
        pinned_all = eng.pinned_empty((B * world,))
        e2e_value = time_e2e(lambda i: sh.log_likelihood(pinned_global[i % 2], out=pinned_all if rank == 0 else None, root=0))
        e2e_extra["e2e_all_ranks"] = {"value": time_e2e(lambda i: sh.log_likelihood(pinned_global[i % 2], out=pinned_all)), "unit": UNIT,
                                      "d2h_bytes_per_step": B * world * 8 * world,
                                      "note": "every rank downloads the gathered vector (SPMD samplers that all need every value)"}
        e2e_api = ("ShardedEngine.log_likelihood(global batch, out=, root=0) on page-locked host arrays -> cl_eval_allgather: H2D of the row shard on "
                   "every rank, ncclAllGather inside the library, D2H of the gathered vector on rank 0")
        h2d, d2h = B * nd * 8 * world, B * world * 8
    else:
        # theta batches and the result buffer live in page-locked arrays (Engine.pinned_empty): cl_eval moves them by DMA
        pinned_batches = []
        for hb in host_batches:
            pb = eng.pinned_empty(hb.shape)
            pb[...] = hb
            pinned_batches.append(pb)
        pinned_out = eng.pinned_empty((B,))
        e2e_value = time_e2e(lambda i: eng.log_likelihood(pinned_batches[i % n_rot], out=pinned_out))
        e2e_extra["e2e_pageable"] = {"value": time_e2e(lambda i: eng.log_likelihood(host_batches[i % n_rot])), "unit": UNIT,
                                     "note": "ordinary numpy arrays in and out (what emcee / nautilus hand over): one staging copy each way inside cl_eval"}
        e2e_api = "Engine.log_likelihood(batch, out=) on Engine.pinned_empty() host arrays -> cl_eval (DMA from / to the caller's page-locked buffers)"
        h2d, d2h = B * nd * 8, B * 8

    # ---- the other stage-3 engines on the same batches (rank 0, one GPU): device-resident, 20 steps each ----
    engines = None
    if rank == 0 and world == 1 and not args.no_alt and not args.opt:
        engines = {}
        ref_out = d_out.clone()
        for name, e_id, sl in (("dmma_fp64", 0, 7), ("tcgen05_7planes", 1, 7), ("tcgen05_6planes", 1, 6)):
            eng.set_option("chi2_engine", e_id); eng.set_option("chi2_slices", sl)
            eng.set_option("chi2_guard", 1 if sl >= 7 else 0)   # 6 planes: the accuracy guard would send every row to the FP64 engine at its default tolerance
            for i in range(3):
                step(args.steps - 1)
            diff = float((d_out - ref_out).abs().max().item()) * 2.0      # |d chi2| against the headline engine, same batch
            torch.cuda.synchronize(dev)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            for i in range(20):
                step(i)
            a1.record(stream)
            torch.cuda.synchronize(dev)
            h, sp = eng.timing_history(20), eng.stage3_split(20)
            engines[name] = {"evals_per_s": B * 20 / (a0.elapsed_time(a1) * 1e-3), "stage12_ms": float(np.mean(h[:, 0])),
                             "planes_ms": float(np.mean(sp[:, 0])), "contraction_ms": float(np.mean(sp[:, 1])),
                             "max_abs_dchi2_vs_headline": diff}
        eng.set_option("chi2_engine", 1 if args.engine == "tcgen05" else 0); eng.set_option("chi2_slices", args.slices)
        eng.set_option("chi2_guard", 1)

    if rank == 0:
        N = args.n_sn
        value = B * world * args.steps / (ms * 1e-3)
        gemm_ms = float(np.mean(split[:, 1])); planes_ms = float(np.mean(split[:, 0])); s12_ms = float(np.mean(hist[:, 0]))
        flops = B * (N * N + 2.0 * N)  # algorithmic FP64: forward-substitution-equivalent MACs*2
        pk = measured_peaks()
        hbm_peak = pk.get("hbm_gbs", 6650.0)
        s12_bytes = B * (8.0 * nd + (args.slices if args.engine == "tcgen05" else 8.0) * N + 8.0 * 8)  # theta in, digit planes (or the FP64 row) out, aux
        if args.engine == "dmma":
            ach = flops / (gemm_ms * 1e-3) / 1e12
            roofline = {"bound": "tensor", "kernel": "k_chi2_gemm (stage 3, FP64 DMMA)", "achieved": ach, "peak": peak_tf,
                        "unit": "TFLOP/s", "frac": ach / peak_tf,
                        "traffic": 1.20e9 if (N == 1701 and B == 65536) else None,
                        "traffic_note": "DRAM bytes per launch (dram__bytes_read+write, ncu, profiles/r01f_gemm_dram.csv); "
                                        "algorithmic 0.92e9 (R read once + W + partial sums)",
                        "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 entry); "
                                       "DMMA.8x8x4 issue peak measured 37.1 TFLOP/s (profiles/r01_ubench_fp64.log)",
                        "algorithmic_flops_per_eval": N * N + 2.0 * N, "avg_kernel_ms": gemm_ms}
        else:
            S = args.slices
            pairs = S * (S + 1) // 2
            i8_ops = B * pairs * (N * N + N + 0.0)   # algorithmic int8 ops: `pairs` exact digit-plane products of the N(N+1)/2-MAC triangle
            ach = i8_ops / (gemm_ms * 1e-3) / 1e12
            i8_peak = I8_PEAK_TOPS
            roofline = {"bound": "tensor", "kernel": f"k_chi2_ozaki<{S}> (stage 3, tcgen05.mma kind::i8, {pairs} digit-plane products)",
                        "achieved": ach, "peak": i8_peak, "unit": "TFLOP/s", "frac": ach / i8_peak,
                        "unit_note": "int8 tensor TOP/s (2 x MAC); the FP64 tensor pipe is not used by this engine",
                        "traffic": OZ_DRAM_BYTES.get((S, N, B)),
                        "traffic_note": "DRAM bytes per launch (dram__bytes_read+write, ncu --set full, profiles/); algorithmic = the digit planes "
                                        f"read once ({S} B per residual) + the W planes = {(S * B * N + S * N * N / 2) / 1e9:.2f}e9",
                        "peak_source": "tcgen05.mma kind::i8 M=128 N=256 K=32 issue rate measured on this pool's B200 (tools/ubench_umma_i8.cu, "
                                       "profiles/r02_ubench_umma_i8.log: 136.1 cycles per MMA = 7709 MAC/clk/SM -> 4485 TOP/s at 1965 MHz; nominal 4766)",
                        "algorithmic_int8_ops_per_eval": pairs * (N * N + N), "avg_kernel_ms": gemm_ms,
                        "fp64_equivalent": {"achieved_tflops": flops / (gemm_ms * 1e-3) / 1e12, "dgemm_peak_tflops": peak_tf,
                                            "note": "the same contraction counted in FP64 flops (N^2 + 2N per eval) against cuBLAS DGEMM measured in this run"},
                        "bound_note": "shared-memory bandwidth (UTCIMMA operand reads + TMA writes) sits just above the tensor pipe for this shape: "
                                      "DESIGN.md section 4"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64",
            "dtype_note": "all stages FP64" if args.engine == "dmma" else
                          "FP64 throughout; stage 3 multiplies exact int8 digit planes of the FP64 operands (int32 accumulation, FP64 recombination)",
            "data": "synthetic", "config": workload_config(args, spec, world),
            "roofline": roofline,
            "roofline_stage12": {"bound": "fp64", "kernel": "k_friedmann_residuals (stage 1+2: Friedmann grid, SN residuals, digit planes)",
                                 "achieved": B * S12_FLOPS_PER_EVAL(N, int(spec.z_grid.size)) / (s12_ms * 1e-3) / 1e12, "peak": FP64_PIPE_TFLOPS, "unit": "TFLOP/s",
                                 "frac": B * S12_FLOPS_PER_EVAL(N, int(spec.z_grid.size)) / (s12_ms * 1e-3) / 1e12 / FP64_PIPE_TFLOPS,
                                 "avg_kernel_ms": s12_ms, "algorithmic_flops_per_eval": S12_FLOPS_PER_EVAL(N, int(spec.z_grid.size)),
                                 "hbm_gbs": s12_bytes / (s12_ms * 1e-3) / 1e9, "hbm_frac": s12_bytes / (s12_ms * 1e-3) / 1e9 / hbm_peak,
                                 "note": "FP64-ALU / issue bound, not HBM bound (SURVEY.md T5): algorithmic FP64 flops (9 G + 40 N, SURVEY.md 8(d)) against the "
                                         "FP64 pipe (64 DFMA/clk/SM = 37.1 TFLOP/s, profiles/r01_ubench_fp64.log); ncu: FP64 pipe 39 % of cycles, "
                                         "issue slots 57 % (profiles/r05_ncu_full_summary.txt); HBM traffic = theta in + 7 digit planes out"},
            "stage_ms": {"stage12": s12_ms, "stage3_planes": planes_ms, "stage3_contraction": gemm_ms, "finalize": float(np.mean(hist[:, 2])),
                         "total": float(np.mean(hist[:, 3]))},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": e2e_api},
            "gpu_launches": int(launches), "clocks": clocks, "parity_check": "64 rows vs oracle ok",
            "engine": eng.describe() + f", stage 3: {args.engine}" + (f" {args.slices} planes" if args.engine == "tcgen05" else ""),
        }
        line.update(e2e_extra)
        if engines:
            line["engines"] = engines
            d = engines["dmma_fp64"]
            ach = flops / (d["contraction_ms"] * 1e-3) / 1e12
            line["roofline_fp64"] = {"bound": "tensor", "kernel": "k_chi2_gemm (stage 3 on the FP64 tensor pipe, chi2_engine = 0)", "achieved": ach,
                                     "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "avg_kernel_ms": d["contraction_ms"],
                                     "evals_per_s": d["evals_per_s"],
                                     "note": "the north star's FP64 DMMA GEMM, timed in the same run: algorithmic N^2 + 2N flops per eval against cuBLAS DGEMM "
                                             "8192^3 measured in this run (DMMA issue peak 37.1 TFLOP/s)"}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(spec, host_batches[0])
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
