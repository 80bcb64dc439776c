// multi.cuh — what the C ABI needs beyond one GPU and one caller-supplied batch:
//   * NCCL, bound at run time (dlopen of libnccl.so.2: the library has no link-time dependency on it and single-GPU callers
//     never load it) — one communicator per context, used for the all-gather of a sharded batch's results
//     (the reference's Pool.map / prange over rows, sn/pantheon.py:119-125, bao/desi.py:100-106) and for the reduction of a
//     sharded profile-likelihood grid;
//   * parameter vectors generated ON THE DEVICE from a Cartesian grid description, and the running reduction of a grid's
//     values (minimum, its index, log-sum-exp), so that a 1e8-point grid neither uploads theta nor downloads values.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/cosmolike.h"

namespace cosmolike {

// ---- NCCL entry points (the subset used), resolved with dlsym --------------------------------------------------------
struct NcclUid { char internal[128]; };   // ncclUniqueId
typedef struct ncclComm* NcclComm;
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUid*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUid, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  const char* why = nullptr;
};
constexpr int kNcclFloat64 = 8, kNcclInt8 = 0;   // ncclDataType_t values (nccl.h)

inline NcclApi& nccl_api() {
  static NcclApi api;
  if (api.handle || api.why) return api;
  const char* env = getenv("COSMOLIKE_NCCL_LIB");
  const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n || !*n) continue;
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) { api.why = "libnccl.so.2 not found (set COSMOLIKE_NCCL_LIB to its path)"; return api; }
#define CL_NCCL_SYM(field, name)                                              \
  *reinterpret_cast<void**>(&api.field) = dlsym(api.handle, name);            \
  if (!api.field) { api.why = "symbol " name " missing in the NCCL library"; api.handle = nullptr; return api; }
  CL_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  CL_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  CL_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  CL_NCCL_SYM(AllGather, "ncclAllGather")
  CL_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef CL_NCCL_SYM
  return api;
}

// ---- Cartesian grids on the device ------------------------------------------------------------------------------------
struct DevGrid {
  int ndim, n_axes;
  int col[CL_MAX_DIM];
  int64_t n[CL_MAX_DIM];
  double lo[CL_MAX_DIM], step[CL_MAX_DIM], hi[CL_MAX_DIM];
  double fixed[CL_MAX_DIM];
};

// theta[i][*] for the grid points first .. first + rows - 1 (linear index, LAST axis fastest, the order of
// np.meshgrid(..., indexing="ij").ravel()); axis values are np.linspace(lo, hi, n): lo + k * step, the last one hi exactly
__global__ void __launch_bounds__(256) k_grid_theta(const __grid_constant__ DevGrid g, int64_t first, int64_t rows, double* __restrict__ theta) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  double* t = theta + i * g.ndim;
  for (int j = 0; j < g.ndim; j++) t[j] = g.fixed[j];
  int64_t r = first + i;
  for (int a = g.n_axes - 1; a >= 0; a--) {
    const int64_t k = r % g.n[a];
    r /= g.n[a];
    t[g.col[a]] = (k == g.n[a] - 1 && g.n[a] > 1) ? g.hi[a] : __dadd_rn(g.lo[a], __dmul_rn((double)k, g.step[a]));   // numpy's arange * step + start: no FMA
  }
}

// value of a grid point from the engine's output: plain (chi2 / log L / log P as they are) or the SN moments with the
// offset profiled / marginalised: yy - yu^2 / uu (+ ln(uu / 2 pi))
struct GridReduceArgs {
  const double* src;      // [rows] or [rows][3]
  int moments;            // 0 plain, 1 offset profiled, 2 offset marginalised
  int as_chi2;            // 1: the value is a chi2 (weight exp(-v/2), best = minimum); 0: a log-probability (weight exp(v), best = maximum)
  int64_t rows, first;
  double* vals;           // nullable [rows]: the values themselves
  double* part;           // [blocks][4]: best value (as a chi2-like "smaller is better" number), its index, max log-weight, sum exp(log-weight - max)
};
__device__ __forceinline__ double grid_value(const GridReduceArgs& a, int64_t i) {
  if (!a.moments) return a.src[i];
  const double yy = a.src[3 * i], yu = a.src[3 * i + 1], uu = a.src[3 * i + 2];
  double v = yy - yu * yu / uu;
  if (a.moments == 2) v += log(uu / (2.0 * M_PI));
  return v;
}
__global__ void __launch_bounds__(256) k_grid_reduce(const GridReduceArgs a) {
  __shared__ double s_best[8], s_idx[8], s_max[8], s_sum[8];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double best = INFINITY, idx = -1.0, lw = -INFINITY;
  if (i < a.rows) {
    const double v = grid_value(a, i);
    if (a.vals) a.vals[i] = v;
    const double c = a.as_chi2 ? v : -2.0 * v;   // chi2-like: smaller is better; NaN never wins
    if (c < best) { best = c; idx = (double)(a.first + i); }
    lw = -0.5 * c;
    if (!(lw == lw)) lw = -INFINITY;   // NaN carries no weight
  }
  // warp, then block: minimum with its index (ties: the smaller index), and a two-pass log-sum-exp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double mx = lw;
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o), oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ob < best || (ob == best && oi >= 0.0 && (idx < 0.0 || oi < idx))) { best = ob; idx = oi; }
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (lane == 0) { s_best[warp] = best; s_idx[warp] = idx; s_max[warp] = mx; }
  __syncthreads();
  double bmx = s_max[0];
  for (int w = 1; w < 8; w++) bmx = fmax(bmx, s_max[w]);
  double sum = (lw > -INFINITY) ? exp(lw - bmx) : 0.0;
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) s_sum[warp] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++)
      if (s_best[w] < best || (s_best[w] == best && s_idx[w] >= 0.0 && (idx < 0.0 || s_idx[w] < idx))) { best = s_best[w]; idx = s_idx[w]; }
    double tot = 0.0;
    for (int w = 0; w < 8; w++) tot += s_sum[w];
    double* p = a.part + 4 * (int64_t)blockIdx.x;
    p[0] = best; p[1] = idx; p[2] = bmx; p[3] = tot;
  }
}

// host side: fold (best, index, max log-weight, sum) pairs; associative, so chunks, passes and ranks combine in any grouping
struct GridStats { double best = INFINITY, index = -1.0, lmax = -INFINITY, sum = 0.0; int64_t count = 0; };
inline void grid_fold(GridStats& s, double best, double index, double lmax, double sum) {
  if (best < s.best || (best == s.best && index >= 0.0 && (s.index < 0.0 || index < s.index))) { s.best = best; s.index = index; }
  if (lmax > -INFINITY && sum > 0.0) {
    if (lmax > s.lmax) { s.sum = s.sum * exp(s.lmax - lmax) + sum; s.lmax = lmax; }
    else s.sum += sum * exp(lmax - s.lmax);
  }
}

}  // namespace cosmolike
