"""ctypes binding of libcosmolike_b200.so — the host-side mirror of the reference's per-script callables.

`Engine(spec)` exposes, for one `LikelihoodSpec`, the names every reference fit script defines at module level:
`chi_squared(theta)`, `log_likelihood(theta)`, `log_probability(theta)` (sn/pantheon.py:57-97) and the batch
form `log_probs_vectorized(batch)` (bao/desi.py:100-106), plus the helper exports their `main()`s use
(`DM_z`, `DH_z`, `bao_theory`, `cmb_distances`).  A 1-D theta returns a Python float like the reference's
scalar functions; a 2-D batch returns an ndarray.

There is no CPU fallback: if the shared library is missing or no sm_100 GPU is visible, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .spec import ClSpec, LikelihoodSpec, OUT_CHI2, OUT_LOGLIKE, OUT_LOGPROB

_dp = C.POINTER(C.c_double)
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libcosmolike_b200.so"

#: every symbol include/cosmolike.h declares
ABI_SYMBOLS = (
    "cl_create", "cl_destroy", "cl_last_error", "cl_eval", "cl_eval_device", "cl_eval_components",
    "cl_eval_sn_moments", "cl_distances", "cl_bao_theory", "cl_cmb", "cl_sn_residuals", "cl_last_timing",
    "cl_timing_history", "cl_launch_count", "cl_set_option", "cl_describe", "cl_stage3_split", "cl_host_alloc", "cl_host_free",
    "cl_set_option_f64", "cl_guard_info", "cl_graph_info",
    "cl_comm_unique_id", "cl_comm_init", "cl_comm_destroy", "cl_comm_info", "cl_eval_allgather", "cl_eval_allgather_device",
    "cl_eval_grid", "cl_grid_allreduce", "cl_propose_eval",
)

#: cl_eval_grid selectors beyond OUT_CHI2 / OUT_LOGLIKE / OUT_LOGPROB (include/cosmolike.h CL_GRID_*)
GRID_PROFILE, GRID_MARGINAL = 16, 17
NCCL_UID_BYTES = 128


class ClGrid(C.Structure):
    """ctypes mirror of cl_grid (include/cosmolike.h)."""
    _fields_ = [("n_axes", C.c_int32), ("col", C.c_int32 * 12), ("n", C.c_int64 * 12), ("lo", C.c_double * 12),
                ("hi", C.c_double * 12), ("fixed", C.c_double * 12)]


class ClProposal(C.Structure):
    """ctypes mirror of cl_proposal."""
    _fields_ = [("ndim", C.c_int32), ("gauss", C.c_int32 * 12), ("mu", C.c_double * 12), ("L", C.c_double * 144),
                ("lo", C.c_double * 12), ("hi", C.c_double * 12), ("mean", C.c_double * 12), ("sigma", C.c_double * 12)]


class ClGridStats(C.Structure):
    """ctypes mirror of cl_grid_stats."""
    _fields_ = [("best", C.c_double), ("index", C.c_int64), ("log_sum", C.c_double), ("count", C.c_int64),
                ("larger_is_better", C.c_int32), ("reserved", C.c_int32)]

#: chi-squared engines for large SN blocks (include/cosmolike.h CL_CHI2_ENGINE_*)
CHI2_ENGINE_DMMA, CHI2_ENGINE_TCGEN05 = 0, 1


class EngineError(RuntimeError):
    pass


def library_path():
    return os.environ.get("COSMOLIKE_LIB") or os.path.join(_HERE, _LIB_NAME)


_lib = None


def load_library():
    """dlopen the in-tree CUDA library (built by cosmology_model_fit_b200/build.py or __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise EngineError(f"{path} not found: run `python -m cosmology_model_fit_b200.build` (needs nvcc); "
                          "there is no CPU fallback")
    lib = C.CDLL(path)
    sp, ctxp, i64 = C.POINTER(ClSpec), C.c_void_p, C.c_int64
    lib.cl_create.argtypes = [sp, C.c_int, C.POINTER(ctxp)]
    lib.cl_destroy.argtypes = [ctxp]
    lib.cl_last_error.argtypes = [ctxp]
    lib.cl_last_error.restype = C.c_char_p
    lib.cl_describe.argtypes = [ctxp]
    lib.cl_describe.restype = C.c_char_p
    lib.cl_eval.argtypes = [ctxp, _dp, i64, i64, C.c_int, _dp]
    lib.cl_eval_device.argtypes = [ctxp, C.c_void_p, i64, i64, C.c_int, C.c_void_p, C.c_void_p]
    lib.cl_eval_components.argtypes = [ctxp, _dp, i64, i64, _dp]
    lib.cl_eval_sn_moments.argtypes = [ctxp, _dp, i64, i64, _dp]
    lib.cl_distances.argtypes = [ctxp, _dp, i64, i64, _dp, i64, _dp, _dp]
    lib.cl_bao_theory.argtypes = [ctxp, _dp, i64, i64, _dp]
    lib.cl_cmb.argtypes = [ctxp, _dp, i64, i64, _dp]
    lib.cl_sn_residuals.argtypes = [ctxp, _dp, i64, i64, _dp]
    lib.cl_last_timing.argtypes = [ctxp, C.c_double * 4]
    lib.cl_timing_history.argtypes = [ctxp, C.c_int, _dp]
    lib.cl_stage3_split.argtypes = [ctxp, C.c_int, _dp]
    lib.cl_launch_count.argtypes = [ctxp]
    lib.cl_launch_count.restype = i64
    lib.cl_set_option.argtypes = [ctxp, C.c_char_p, i64]
    lib.cl_set_option_f64.argtypes = [ctxp, C.c_char_p, C.c_double]
    lib.cl_graph_info.argtypes = [ctxp, C.POINTER(C.c_int64)]
    lib.cl_guard_info.argtypes = [ctxp, C.c_double * 4]
    lib.cl_host_alloc.argtypes = [ctxp, C.c_size_t, C.POINTER(C.c_void_p)]
    lib.cl_host_free.argtypes = [ctxp, C.c_void_p]
    lib.cl_comm_unique_id.argtypes = [C.c_void_p]
    lib.cl_comm_init.argtypes = [ctxp, C.c_int, C.c_int, C.c_void_p]
    lib.cl_comm_destroy.argtypes = [ctxp]
    lib.cl_comm_info.argtypes = [ctxp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.cl_eval_allgather.argtypes = [ctxp, _dp, i64, i64, C.c_int, _dp, C.c_int]
    lib.cl_eval_allgather_device.argtypes = [ctxp, C.c_void_p, i64, i64, C.c_int, C.c_void_p, C.c_void_p]
    lib.cl_eval_grid.argtypes = [ctxp, C.POINTER(ClGrid), i64, i64, C.c_int, _dp, C.POINTER(ClGridStats)]
    lib.cl_grid_allreduce.argtypes = [ctxp, C.POINTER(ClGridStats)]
    lib.cl_propose_eval.argtypes = [ctxp, C.POINTER(ClProposal), i64, C.c_uint64, C.c_uint64, C.c_int, C.c_double, i64, _dp, _dp, _dp,
                                    C.POINTER(C.c_int64 * 3)]
    _lib = lib
    return lib


def _p(a):
    return a.ctypes.data_as(_dp)


def _point_at_bundled_nccl():
    """The library binds NCCL at run time (dlopen of libnccl.so.2).  In a process that already imported torch the bundled
    library is loaded and found by its SONAME; otherwise point $COSMOLIKE_NCCL_LIB at the wheel's copy if there is one."""
    if os.environ.get("COSMOLIKE_NCCL_LIB"):
        return
    try:
        import importlib.util
        sp = importlib.util.find_spec("nvidia.nccl")
        for root in (sp.submodule_search_locations or []) if sp else []:
            cand = os.path.join(root, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["COSMOLIKE_NCCL_LIB"] = cand
                return
    except Exception:
        pass


class Engine:
    """One likelihood on one GPU (one process per GPU; see parallel.py for the sharded form)."""

    def __init__(self, spec: LikelihoodSpec, device: int = 0):
        self.lib = load_library()
        self.spec = spec
        self._c_spec = spec.c_spec()
        self._ctx = C.c_void_p()
        rc = self.lib.cl_create(C.byref(self._c_spec), int(device), C.byref(self._ctx))
        if rc != 0:
            msg = self.lib.cl_last_error(None)
            self._ctx = C.c_void_p()
            raise EngineError(f"cl_create failed ({rc}): {msg.decode() if msg else ''}")
        self.ndim = spec.ndim
        self.device = device

    # -- lifetime ---------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self.lib.cl_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc != 0:
            msg = self.lib.cl_last_error(self._ctx)
            raise EngineError(f"cosmolike error {rc}: {msg.decode() if msg else ''}")

    def describe(self):
        return self.lib.cl_describe(self._ctx).decode()

    def set_option(self, name, value):
        if isinstance(value, float) and name in ("chi2_guard_abs", "chi2_guard_rel"):
            self._check(self.lib.cl_set_option_f64(self._ctx, name.encode(), float(value)))
        else:
            self._check(self.lib.cl_set_option(self._ctx, name.encode(), int(value)))

    def guard_info(self):
        """Accuracy guard of the tcgen05 chi-squared engine (cl_guard_info): rows recomputed on the FP64 engine since
        creation / by the last pass, and the static factors of the a-priori bound."""
        v = (C.c_double * 4)()
        self._check(self.lib.cl_guard_info(self._ctx, v))
        return {"rows_total": int(v[0]), "rows_last_pass": int(v[1]), "omega": v[2], "kappa": v[3]}

    def graph_info(self):
        """CUDA graphs of small evaluations (cl_graph_info): graphs held and replays since creation."""
        v = (C.c_int64 * 2)()
        self._check(self.lib.cl_graph_info(self._ctx, v))
        return {"graphs": int(v[0]), "replays": int(v[1])}

    # -- evaluation -------------------------------------------------------------------------------------------
    def _theta(self, theta):
        t = np.asarray(theta, dtype=np.float64)
        scalar = t.ndim == 1
        t = np.ascontiguousarray(np.atleast_2d(t))
        if t.ndim != 2 or t.shape[1] != self.ndim:
            raise ValueError(f"theta must have shape ({self.ndim},) or (B, {self.ndim})")
        return t, scalar

    def pinned_empty(self, shape, dtype=np.float64):
        """Uninitialised array in page-locked host memory (cl_host_alloc).  theta batches and `out=` buffers that live in such
        arrays are moved by DMA directly; ordinary numpy arrays go through the library's staging copy.  The memory is
        released when the array (and every view of it) is garbage-collected."""
        import weakref
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        ptr = C.c_void_p()
        self._check(self.lib.cl_host_alloc(self._ctx, max(n, 1), C.byref(ptr)))
        buf = (C.c_char * max(n, 1)).from_address(ptr.value)
        weakref.finalize(buf, self.lib.cl_host_free, None, C.c_void_p(ptr.value))
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def _eval(self, theta, what, out=None):
        t, scalar = self._theta(theta)
        if out is None:
            out = np.empty(t.shape[0], dtype=np.float64)
        elif out.dtype != np.float64 or out.shape != (t.shape[0],) or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous float64 array with one element per row of theta")
        self._check(self.lib.cl_eval(self._ctx, _p(t), t.shape[0], t.shape[1], what, _p(out)))
        return float(out[0]) if scalar else out

    def chi_squared(self, theta, out=None):
        """chi_squared(params) (sn/pantheon.py:57-61)"""
        return self._eval(theta, OUT_CHI2, out)

    def log_likelihood(self, theta, out=None):
        """log_likelihood(params) (sn/pantheon.py:64-65; guard of bao/desi_fs_lya_cmb.py:117-121)"""
        return self._eval(theta, OUT_LOGLIKE, out)

    def log_probability(self, theta, out=None):
        """log_probability(params): -inf outside the prior box (sn/pantheon.py:88-97)"""
        return self._eval(theta, OUT_LOGPROB, out)

    def log_probs_vectorized(self, batch, dtype=np.float32):
        """Batch log-probability for emcee(vectorize=True); float32 like bao/desi.py:100-106 unless dtype says
        otherwise."""
        return np.asarray(self._eval(np.atleast_2d(batch), OUT_LOGPROB)).astype(dtype, copy=False)

    def log_likelihood_vectorized(self, batch, dtype=np.float32):
        """Batch log-likelihood for nautilus(vectorized=True) (bao/desi_union3_cc_theta_star.py:142-147)."""
        return np.asarray(self._eval(np.atleast_2d(batch), OUT_LOGLIKE)).astype(dtype, copy=False)

    def components(self, theta):
        """chi2 split as (sn, bao, cmb, cc+gaussian) (bao/desi_cmb_union3.py:97-135)"""
        t, scalar = self._theta(theta)
        out = np.empty((t.shape[0], 4))
        self._check(self.lib.cl_eval_components(self._ctx, _p(t), t.shape[0], t.shape[1], _p(out)))
        return out[0] if scalar else out

    def sn_moments(self, theta):
        """(y.y, y.u, u.u) for the analytic treatment of the magnitude offset."""
        t, scalar = self._theta(theta)
        out = np.empty((t.shape[0], 3))
        self._check(self.lib.cl_eval_sn_moments(self._ctx, _p(t), t.shape[0], t.shape[1], _p(out)))
        return out[0] if scalar else out

    def eval_device(self, d_theta: int, B: int, ld: int, what: int, d_out: int, stream: int = 0):
        """Asynchronous evaluation on raw device pointers (ints), e.g. torch tensors' data_ptr()."""
        self._check(self.lib.cl_eval_device(self._ctx, C.c_void_p(d_theta), B, ld, what, C.c_void_p(d_out), C.c_void_p(stream)))

    # -- multi-GPU (one process per GPU): NCCL communicator inside the library ----------------------------------
    @staticmethod
    def nccl_unique_id() -> bytes:
        """The 128-byte NCCL id rank 0 creates and ships to the other ranks (cl_comm_unique_id)."""
        _point_at_bundled_nccl()
        lib = load_library()
        buf = C.create_string_buffer(NCCL_UID_BYTES)
        rc = lib.cl_comm_unique_id(buf)
        if rc != 0:
            msg = lib.cl_last_error(None)
            raise EngineError(f"cl_comm_unique_id failed ({rc}): {msg.decode() if msg else ''}")
        return buf.raw

    def comm_init(self, rank: int, world: int, uid: bytes):
        """Collective over all ranks: binds an NCCL communicator to this context (cl_comm_init)."""
        _point_at_bundled_nccl()
        if len(uid) != NCCL_UID_BYTES:
            raise ValueError("uid must be the 128 bytes of Engine.nccl_unique_id()")
        self._check(self.lib.cl_comm_init(self._ctx, int(rank), int(world), C.create_string_buffer(uid, NCCL_UID_BYTES)))

    def comm_info(self):
        r, w = C.c_int(), C.c_int()
        self._check(self.lib.cl_comm_info(self._ctx, C.byref(r), C.byref(w)))
        return r.value, w.value

    def eval_allgather(self, theta_local, what, out_all=None, root=-1):
        """Collective: this rank's rows -> the results of every rank in rank order (cl_eval_allgather).  root >= 0: only
        that rank receives (returns None elsewhere)."""
        t = np.ascontiguousarray(np.atleast_2d(theta_local), dtype=np.float64)
        rank, world = self.comm_info()
        receive = root < 0 or root == rank
        if receive and out_all is None:
            out_all = np.empty(world * t.shape[0])
        self._check(self.lib.cl_eval_allgather(self._ctx, _p(t), t.shape[0], t.shape[1], int(what),
                                               _p(out_all) if receive else None, int(root)))
        return out_all if receive else None

    def eval_allgather_device(self, d_theta: int, B: int, ld: int, what: int, d_out_all: int, stream: int = 0):
        self._check(self.lib.cl_eval_allgather_device(self._ctx, C.c_void_p(d_theta), B, ld, what, C.c_void_p(d_out_all), C.c_void_p(stream)))

    # -- profile-likelihood grids generated on the device (BASELINE.json config 4) -------------------------------
    def make_grid(self, axes, fixed=None):
        """cl_grid from {theta column: (lo, hi, n)} (insertion order = axis order, last axis fastest) and {column: value}."""
        g = ClGrid()
        g.n_axes = len(axes)
        for a, (col, (lo, hi, n)) in enumerate(axes.items()):
            g.col[a], g.lo[a], g.hi[a], g.n[a] = int(col), float(lo), float(hi), int(n)
        for col, v in (fixed or {}).items():
            g.fixed[int(col)] = float(v)
        return g

    def eval_grid(self, grid, first, count, what, want_values=False, allreduce=False):
        """Evaluates the points first .. first + count - 1 of the flattened grid on the device (cl_eval_grid); returns
        (stats dict, values or None).  allreduce=True combines the stats over the communicator's ranks (cl_grid_allreduce)."""
        st = ClGridStats()
        out = np.empty(count) if want_values else None
        self._check(self.lib.cl_eval_grid(self._ctx, C.byref(grid), int(first), int(count), int(what), _p(out) if want_values else None, C.byref(st)))
        if allreduce:
            self._check(self.lib.cl_grid_allreduce(self._ctx, C.byref(st)))
        return {"best": st.best, "index": st.index, "log_sum": st.log_sum, "count": st.count}, out

    # -- nested-sampling proposals on the device ------------------------------------------------------------------
    def propose_eval(self, mu, L, bounds, n, seed, offset, what, thresh, max_keep, gauss=None):
        """n points uniform in the ellipsoid {mu + L z} of the unit cube, prior-transformed, evaluated and filtered on the
        device (cl_propose_eval).  Returns (u, theta, values) of the accepted rows in draw order (at most max_keep) and the
        counts (inside the cube, accepted, returned)."""
        d = self.ndim
        p = ClProposal()
        p.ndim = d
        L = np.asarray(L, dtype=np.float64)
        for j in range(d):
            p.mu[j], p.lo[j], p.hi[j] = float(mu[j]), float(bounds[j][0]), float(bounds[j][1])
            for k in range(d):
                p.L[j * d + k] = float(L[j, k])
        for col, (mean, sigma) in (gauss or {}).items():
            p.gauss[col], p.mean[col], p.sigma[col] = 1, float(mean), float(sigma)
        u, th, val = np.empty((max_keep, d)), np.empty((max_keep, d)), np.empty(max_keep)
        cnt = (C.c_int64 * 3)()
        self._check(self.lib.cl_propose_eval(self._ctx, C.byref(p), int(n), int(seed), int(offset), int(what), float(thresh), int(max_keep),
                                             _p(u), _p(th), _p(val), C.byref(cnt)))
        k = cnt[2]
        return u[:k], th[:k], val[:k], (cnt[0], cnt[1], cnt[2])

    # -- helper exports ---------------------------------------------------------------------------------------
    def distances(self, theta, zq):
        t, scalar = self._theta(theta)
        zq = np.ascontiguousarray(np.atleast_1d(zq), dtype=np.float64)
        dm = np.empty((t.shape[0], zq.size)); dh = np.empty((t.shape[0], zq.size))
        self._check(self.lib.cl_distances(self._ctx, _p(t), t.shape[0], t.shape[1], _p(zq), zq.size, _p(dm), _p(dh)))
        return (dm[0], dh[0]) if scalar else (dm, dh)

    def DM_z(self, z, theta):
        """DM_z(z, params): trapezoid grid + Hermite (bao/desi_cmb_pantheon.py:66-72)"""
        return self.distances(theta, z)[0]

    def DH_z(self, z, theta):
        """DH_z(z, params) = c / H(z) (bao/desi_cmb_pantheon.py:61-63)"""
        return self.distances(theta, z)[1]

    def bao_theory(self, theta):
        t, scalar = self._theta(theta)
        out = np.empty((t.shape[0], self._c_spec.n_bao))
        self._check(self.lib.cl_bao_theory(self._ctx, _p(t), t.shape[0], t.shape[1], _p(out)))
        return out[0] if scalar else out

    def cmb(self, theta):
        """(v0, v1, v2, z*, r_s*, D_M*, r_drag, 100 theta*)"""
        t, scalar = self._theta(theta)
        out = np.empty((t.shape[0], 8))
        self._check(self.lib.cl_cmb(self._ctx, _p(t), t.shape[0], t.shape[1], _p(out)))
        return out[0] if scalar else out

    def cmb_distances(self, theta):
        """cmb.cmb_distances(...) -> (R, l_A, omega_b) (cmb/data_planck_act_compression.py:200-212)"""
        r = self.cmb(theta)
        return r[..., :3]

    def sn_residuals(self, theta):
        t, scalar = self._theta(theta)
        out = np.empty((t.shape[0], self._c_spec.n_sn))
        self._check(self.lib.cl_sn_residuals(self._ctx, _p(t), t.shape[0], t.shape[1], _p(out)))
        return out[0] if scalar else out

    # -- instrumentation --------------------------------------------------------------------------------------
    def last_timing(self):
        """dict of CUDA-event times (ms) of the last evaluation: stage12, stage3, finalize, total."""
        ms = (C.c_double * 4)()
        self._check(self.lib.cl_last_timing(self._ctx, ms))
        return {"stage12_ms": ms[0], "stage3_ms": ms[1], "finalize_ms": ms[2], "total_ms": ms[3]}

    def stage3_split(self, n=1):
        """[k, 2] array (ms forming the int8 digit planes of the residual rows, ms in the contraction kernel) of the
        last k <= n evaluations, oldest first; the single pair for n == 1."""
        buf = np.zeros((max(n, 1), 2))
        k = self.lib.cl_stage3_split(self._ctx, int(n), _p(buf))
        if k < 0:
            self._check(k)
        return tuple(buf[0]) if n == 1 else buf[:k]

    def timing_history(self, n):
        """[k, 4] array (stage12, stage3, finalize, total ms) of the last k <= n evaluations, oldest first."""
        buf = np.zeros((max(n, 1), 4))
        k = self.lib.cl_timing_history(self._ctx, int(n), _p(buf))
        if k < 0:
            self._check(k)
        return buf[:k]

    def launch_count(self):
        return int(self.lib.cl_launch_count(self._ctx))
