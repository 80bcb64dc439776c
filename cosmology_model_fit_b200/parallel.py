"""Row-sharded evaluation across the GPUs of one node: one process per GPU.

The likelihood of every parameter vector is independent (the reference exploits exactly this with
`multiprocessing.Pool` / numba `prange`: sn/pantheon.py:119-125, bao/desi.py:104-105), so the batch is split into
contiguous row shards, the static operands (W = L^-1, SN vectors, grid, BAO/CMB blocks) are replicated on every
GPU, and the only communication is one all-gather of the per-row results (NCCL over NVLink/NVSwitch).  There is
no data-path collective inside the likelihood itself.

The collective lives in the C ABI (`cl_comm_init`, `cl_eval_allgather`, `cl_eval_grid` + `cl_grid_allreduce`,
include/cosmolike.h): the library holds its own NCCL communicator, so a non-Python caller has the same multi-GPU path.
`torch.distributed` is only used here to ship the 128-byte NCCL id from rank 0 to the other ranks (any transport would
do) and, in the CPU tests, as the gloo stand-in for the collective around an injected evaluator.
"""
from __future__ import annotations

import math

import numpy as np

from .spec import OUT_CHI2, OUT_LOGLIKE, OUT_LOGPROB


def shard_bounds(n_rows: int, rank: int, world: int):
    """Contiguous, balanced row range [lo, hi) of `rank`; the first n_rows % world ranks get one extra row."""
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def grid_stats_of(values, first, larger_is_better=False):
    """Stats of one slice of a flattened grid, as cl_eval_grid forms them: best value, its index (smallest among ties),
    ln sum exp(-v/2) (ln sum exp(v) for log-probabilities), count.  NaN carries no weight and never wins."""
    v = np.asarray(values, dtype=np.float64)
    c = -2.0 * v if larger_is_better else v            # chi2-like: smaller is better
    ok = ~np.isnan(c)
    best, index = (math.inf, -1)
    if v.size:
        cand = np.where(ok, c, math.inf)
        k = int(np.argmin(cand))          # first occurrence: the smallest index among ties
        if cand[k] < math.inf:
            best, index = float(cand[k]), first + k
    lw = np.where(ok, -0.5 * c, -math.inf)
    m = float(np.max(lw)) if lw.size else -math.inf
    log_sum = m + math.log(float(np.sum(np.exp(lw - m)))) if m > -math.inf else -math.inf
    return {"best": (-0.5 * best if larger_is_better else best), "index": index, "log_sum": log_sum, "count": int(v.size)}


def fold_grid_stats(parts, larger_is_better=False):
    """Combines the stats of disjoint slices (cl_grid_allreduce: rank order, ties -> smallest index)."""
    best, index, lmax, s, count = (-math.inf if larger_is_better else math.inf), -1, -math.inf, 0.0, 0
    for p in parts:
        better = p["best"] > best if larger_is_better else p["best"] < best
        if p["index"] >= 0 and (better or (p["best"] == best and (index < 0 or p["index"] < index))):
            best, index = p["best"], p["index"]
        if p["log_sum"] > -math.inf:
            if p["log_sum"] > lmax:
                s = s * math.exp(lmax - p["log_sum"]) + 1.0
                lmax = p["log_sum"]
            else:
                s += math.exp(p["log_sum"] - lmax)
        count += p["count"]
    return {"best": best, "index": index, "log_sum": (lmax + math.log(s) if s > 0.0 else -math.inf), "count": count}


class ShardedEngine:
    """Evaluate a global batch over the ranks of a `torch.distributed` job (or of an explicit (rank, world, uid) triple).

    `evaluator(theta_local[B_r, d], what) -> ndarray[B_r]` may be injected (CPU tests use it with the gloo backend);
    by default it is a CUDA `Engine` bound to `device` whose library-owned NCCL communicator gathers the results.
    """

    def __init__(self, spec, device=None, group=None, evaluator=None, engine=None, rank=None, world=None, uid=None):
        self.spec = spec
        self.evaluator = evaluator
        self.engine = None
        self.group = group
        self.dist = None
        if rank is None or world is None:
            import torch.distributed as dist
            self.dist = dist
            init = dist.is_available() and dist.is_initialized()
            rank = dist.get_rank(group) if init else 0
            world = dist.get_world_size(group) if init else 1
        self.rank, self.world = int(rank), int(world)
        if evaluator is None:
            from .engine import Engine
            if device is None:
                import torch
                device = torch.cuda.current_device()
            self.engine = engine if engine is not None else Engine(spec, device=int(device))
            self._own_engine = engine is None
            if self.world > 1 and self.engine.comm_info()[1] == 0:
                if uid is None:   # rank 0 creates the NCCL id, torch.distributed ships it (any transport would do)
                    box = [Engine.nccl_unique_id() if self.rank == 0 else None]
                    self.dist.broadcast_object_list(box, src=0, group=group)
                    uid = box[0]
                self.engine.comm_init(self.rank, self.world, uid)
            self._pad = None

    # -- batches -------------------------------------------------------------------------------------------------
    def evaluate(self, theta, what=OUT_LOGLIKE, out=None, root=None):
        """theta: the GLOBAL batch [B, d] (identical on every rank).  Returns ndarray[B] on every rank (`out` if given), or -
        with `root` = a rank - only there (None elsewhere: the master/worker shape of the reference's Pool.map, where only
        the sampler's process needs the values).  Page-locked arrays (`self.engine.pinned_empty`) are moved by DMA directly."""
        theta = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
        B = theta.shape[0]
        if out is not None and (out.dtype != np.float64 or out.shape != (B,) or not out.flags.c_contiguous):
            raise ValueError("out must be a contiguous float64 array with one element per row of theta")
        lo, hi = shard_bounds(B, self.rank, self.world)
        rows_max = -(-B // self.world)
        receive = root is None or root == self.rank
        if self.evaluator is not None:
            res = self._evaluate_host(theta, lo, hi, rows_max, what)
            if not receive:
                return None
            if out is None:
                return res
            out[...] = res
            return out
        if self.world == 1:
            return self.engine._eval(theta, what, out)
        local = theta[lo:hi]
        if hi - lo < rows_max:   # every rank passes the same number of rows: the short shards repeat their last row
            if self._pad is None or self._pad.shape[0] < rows_max:
                self._pad = np.empty((rows_max, theta.shape[1]))
            self._pad[: hi - lo] = local
            self._pad[hi - lo: rows_max] = local[-1] if hi > lo else theta[0]
            local = self._pad[:rows_max]
        equal = rows_max * self.world == B
        gathered = out if (receive and out is not None and equal) else None
        gathered = self.engine.eval_allgather(local, what, out_all=gathered, root=-1 if root is None else int(root))
        if not receive:
            return None
        if equal:
            return gathered
        res = self._unpad(gathered, B, rows_max)
        if out is None:
            return res
        out[...] = res
        return out

    def evaluate_local(self, theta_local, what=OUT_LOGLIKE, out=None):
        """This rank's own rows, no collective at all (each rank drives its own chains)."""
        if self.evaluator is not None:
            res = np.asarray(self.evaluator(np.atleast_2d(theta_local), what))
            if out is None:
                return res
            out[...] = res
            return out
        return self.engine._eval(theta_local, what, out)

    def _evaluate_host(self, theta, lo, hi, rows_max, what):
        import torch as t
        local = np.zeros(rows_max)
        if hi > lo:
            local[: hi - lo] = self.evaluator(theta[lo:hi], what)
        if self.world == 1:
            return local[: hi - lo].copy()
        gathered = [t.empty(rows_max, dtype=t.float64) for _ in range(self.world)]
        self.dist.all_gather(gathered, t.from_numpy(local), group=self.group)
        return self._unpad(t.cat(gathered).numpy(), theta.shape[0], rows_max)

    def _unpad(self, flat, B, rows_max):
        if self.world == 1:
            return np.array(flat[:B])
        if rows_max * self.world == B:   # equal shards: the gathered vector is already in row order
            return np.array(flat[:B])
        out = np.empty(B)
        for r in range(self.world):
            lo, hi = shard_bounds(B, r, self.world)
            out[lo:hi] = flat[r * rows_max: r * rows_max + (hi - lo)]
        return out

    def chi_squared(self, theta, out=None, root=None):
        return self.evaluate(theta, OUT_CHI2, out, root)

    def log_likelihood(self, theta, out=None, root=None):
        return self.evaluate(theta, OUT_LOGLIKE, out, root)

    def log_probability(self, theta, out=None, root=None):
        return self.evaluate(theta, OUT_LOGPROB, out, root)

    # -- profile-likelihood grids (BASELINE.json config 4) ----------------------------------------------------------
    def grid(self, axes, fixed, what, want_values=False):
        """A Cartesian grid {theta column: (lo, hi, n)} sharded over the ranks by contiguous slices of the flattened index.
        Every rank generates its own parameter vectors ON THE DEVICE, reduces its slice to (best, index, log-sum-exp) and
        the ranks combine those with one small all-gather: neither theta nor the values cross PCIe (unless `want_values`).
        Returns (global stats, this rank's values or None, (first, count) of this rank's slice)."""
        total = 1
        for (_, _, n) in axes.values():
            total *= int(n)
        first, last = shard_bounds(total, self.rank, self.world)
        larger = what in (OUT_LOGLIKE, OUT_LOGPROB)
        if self.evaluator is not None:
            vals = np.asarray(self.evaluator(grid_points(axes, fixed, self.spec.ndim, first, last - first), what))
            mine = grid_stats_of(vals, first, larger)
            parts = [mine]
            if self.world > 1:
                parts = [None] * self.world
                self.dist.all_gather_object(parts, mine, group=self.group)
            return fold_grid_stats(parts, larger), (vals if want_values else None), (first, last - first)
        g = self.engine.make_grid(axes, fixed)
        stats, vals = self.engine.eval_grid(g, first, last - first, what, want_values=want_values, allreduce=self.world > 1)
        return stats, vals, (first, last - first)

    def close(self):
        if self.engine is not None and getattr(self, "_own_engine", True):
            self.engine.close()


def grid_points(axes, fixed, ndim, first, count):
    """theta rows of the flattened grid points first .. first + count - 1 (last axis fastest; axis = np.linspace(lo, hi, n)):
    the host-side statement of k_grid_theta, for tests and injected evaluators."""
    theta = np.zeros((count, ndim))
    for col, v in (fixed or {}).items():
        theta[:, int(col)] = v
    r = np.arange(first, first + count, dtype=np.int64)
    for col, (lo, hi, n) in reversed(list(axes.items())):
        k = r % n
        r = r // n
        theta[:, int(col)] = np.linspace(lo, hi, n)[k]
    return theta
