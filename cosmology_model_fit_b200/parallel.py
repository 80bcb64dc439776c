"""Row-sharded evaluation across the GPUs of one node: one process per GPU, `torch.distributed` for the plumbing.

The likelihood of every parameter vector is independent (the reference exploits exactly this with
`multiprocessing.Pool` / numba `prange`: sn/pantheon.py:119-125, bao/desi.py:104-105), so the batch is split into
contiguous row shards, the static operands (W = L^-1, SN vectors, grid, BAO/CMB blocks) are replicated on every
GPU, and the only communication is one all-gather of the per-row results (NCCL over NVLink/NVSwitch).  There is
no data-path collective inside the likelihood itself.
"""
from __future__ import annotations

import numpy as np

from .spec import OUT_CHI2, OUT_LOGLIKE, OUT_LOGPROB


def shard_bounds(n_rows: int, rank: int, world: int):
    """Contiguous, balanced row range [lo, hi) of `rank`; the first n_rows % world ranks get one extra row."""
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class ShardedEngine:
    """Evaluate a global batch, every rank returning the full result vector.

    `evaluator(theta_local[B_r, d], what) -> ndarray[B_r]` may be injected (CPU tests use it with the gloo backend);
    by default it is a CUDA `Engine` bound to `device` and results travel GPU -> NCCL all-gather -> host.
    """

    def __init__(self, spec, device=None, group=None, evaluator=None, engine=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.spec = spec
        self.evaluator = evaluator
        self.engine = None
        if evaluator is None:
            from .engine import Engine
            self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
            self.engine = engine if engine is not None else Engine(spec, device=self.device.index)
            self._own_engine = engine is None
            self.stream = torch.cuda.Stream(self.device)
            self._cap = 0

    def _buffers(self, rows_max, total):
        t = self.torch
        if rows_max > self._cap:
            self._cap = rows_max
            self.h_theta = t.empty((rows_max, self.spec.ndim), dtype=t.float64).pin_memory()
            self.d_theta = t.empty((rows_max, self.spec.ndim), dtype=t.float64, device=self.device)
            self.d_out = t.empty(rows_max, dtype=t.float64, device=self.device)
        if getattr(self, "_tot", 0) < total:
            self._tot = total
            self.d_all = t.empty(total, dtype=t.float64, device=self.device)
            self.h_all = t.empty(total, dtype=t.float64).pin_memory()

    def evaluate(self, theta, what=OUT_LOGLIKE, out=None):
        """theta: the GLOBAL batch [B, d] (identical on every rank).  Returns ndarray[B] on every rank (`out` if given).
        Page-locked arrays (`self.engine.pinned_empty`) for theta and `out` are moved by DMA directly: no staging copy of the
        row shard, and - when the shards are equal - the gathered vector lands in `out` without a second host copy."""
        theta = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
        B = theta.shape[0]
        if out is not None and (out.dtype != np.float64 or out.shape != (B,) or not out.flags.c_contiguous):
            raise ValueError("out must be a contiguous float64 array with one element per row of theta")
        lo, hi = shard_bounds(B, self.rank, self.world)
        rows_max = -(-B // self.world)
        if self.evaluator is not None:
            res = self._evaluate_host(theta, lo, hi, rows_max, what)
            if out is None:
                return res
            out[...] = res
            return out
        t, dist = self.torch, self.dist
        self._buffers(rows_max, rows_max * self.world)
        n = hi - lo
        equal = rows_max * self.world == B
        h_out = t.from_numpy(out) if out is not None and equal else None
        direct_out = h_out is not None and h_out.is_pinned()
        with t.cuda.stream(self.stream):
            if n:
                src = t.from_numpy(theta[lo:hi])
                if not src.is_pinned():   # ordinary host memory: stage the row shard in the page-locked buffer
                    self.h_theta[:n].copy_(src)
                    src = self.h_theta[:n]
                self.d_theta[:n].copy_(src, non_blocking=True)
                self.engine.eval_device(self.d_theta.data_ptr(), n, self.spec.ndim, what, self.d_out.data_ptr(),
                                        self.stream.cuda_stream)
            if self.world > 1:
                dist.all_gather_into_tensor(self.d_all[: rows_max * self.world], self.d_out[:rows_max], group=self.group)
                (h_out if direct_out else self.h_all[: rows_max * self.world]).copy_(self.d_all[: rows_max * self.world], non_blocking=True)
            else:
                (h_out if direct_out else self.h_all[:n]).copy_(self.d_out[:n], non_blocking=True)
        self.stream.synchronize()
        if direct_out:
            return out
        res = self._unpad(self.h_all.numpy(), B, rows_max)
        if out is None:
            return res
        out[...] = res
        return out

    def _evaluate_host(self, theta, lo, hi, rows_max, what):
        t, dist = self.torch, self.dist
        local = np.zeros(rows_max)
        if hi > lo:
            local[: hi - lo] = self.evaluator(theta[lo:hi], what)
        if self.world == 1:
            return local[: hi - lo].copy()
        gathered = [t.empty(rows_max, dtype=t.float64) for _ in range(self.world)]
        dist.all_gather(gathered, t.from_numpy(local), group=self.group)
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

    def chi_squared(self, theta, out=None):
        return self.evaluate(theta, OUT_CHI2, out)

    def log_likelihood(self, theta, out=None):
        return self.evaluate(theta, OUT_LOGLIKE, out)

    def log_probability(self, theta, out=None):
        return self.evaluate(theta, OUT_LOGPROB, out)

    def close(self):
        if self.engine is not None and getattr(self, "_own_engine", True):
            self.engine.close()
