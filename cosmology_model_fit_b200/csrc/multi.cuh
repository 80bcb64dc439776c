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

// ---- proposals of a nested sampler generated, evaluated and filtered on the device -------------------------------------
// (the sampler front-end sized for the GPU; the reference drives nautilus, bao/desi_cmb_pantheon.py:153-170)
namespace cosmolike {

// Philox4x32-10 (Salmon et al. 2011): counter-based, so row i of a batch always sees the same numbers whatever the launch
// shape; cosmology_model_fit_b200/samplers.py carries the same generator in numpy for the CPU-driven twin of a run.
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
// two uniforms in (0, 1) from one Philox block: (53-bit integer + 1/2) 2^-53
__host__ __device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
  return ((double)(((uint64_t)(hi >> 5) << 26) | (uint64_t)(lo >> 6)) + 0.5) * (1.0 / 9007199254740992.0);
}

struct DevProposal {
  int ndim;
  double mu[CL_MAX_DIM], L[CL_MAX_DIM * CL_MAX_DIM], lo[CL_MAX_DIM], hi[CL_MAX_DIM], mean[CL_MAX_DIM], sigma[CL_MAX_DIM];
  int gauss[CL_MAX_DIM];
};

// row i: z ~ uniform in the unit ball (normal direction by Box-Muller, radius U^(1/d)), u = mu + L z, theta = prior transform(u).
// inside[i] = 1 when u lies in the open unit cube (the prior's support); other rows still get a finite theta (mu) so that the
// likelihood pass can run over the whole batch without special cases - their values are ignored by the selection.
__global__ void __launch_bounds__(256) k_propose(const __grid_constant__ DevProposal p, int64_t n, uint64_t seed, uint64_t offset,
                                                 double* __restrict__ u_out, double* __restrict__ theta, unsigned char* __restrict__ inside) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int d = p.ndim;
  const uint64_t row = offset + (uint64_t)i;
  double z[CL_MAX_DIM];
  double nrm2 = 0.0, ur = 0.5;
  const int pairs = (d + 1) / 2;
  for (int j = 0; j <= pairs; j++) {
    uint32_t c[4] = {(uint32_t)row, (uint32_t)(row >> 32), (uint32_t)j, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double a = u53(c[0], c[1]), b = u53(c[2], c[3]);
    if (j == pairs) { ur = a; break; }
    const double r = sqrt(-2.0 * log(a));
    double sn, cs;
    sincospi(2.0 * b, &sn, &cs);
    z[2 * j] = r * cs; nrm2 += z[2 * j] * z[2 * j];
    if (2 * j + 1 < d) { z[2 * j + 1] = r * sn; nrm2 += z[2 * j + 1] * z[2 * j + 1]; }
  }
  const double scale = pow(ur, 1.0 / (double)d) / sqrt(nrm2);
  bool in = true;
  for (int r = 0; r < d; r++) {
    double u = p.mu[r];
    for (int k = 0; k <= r; k++) u = fma(p.L[r * d + k], z[k] * scale, u);
    in = in && u > 0.0 && u < 1.0;
    u_out[i * d + r] = u;
  }
  inside[i] = in ? 1 : 0;
  for (int r = 0; r < d; r++) {
    const double u = in ? u_out[i * d + r] : p.mu[r];
    theta[i * d + r] = p.gauss[r] ? p.mean[r] + p.sigma[r] * normcdfinv(fmin(fmax(u, 1e-300), 1.0 - 1e-16)) : p.lo[r] + u * (p.hi[r] - p.lo[r]);
  }
}

// ordered selection of the rows with inside && value > thresh: per-block counts, one-block scan, scatter of the first max_keep
__global__ void __launch_bounds__(256) k_select_count(const double* __restrict__ val, const unsigned char* __restrict__ inside, double thresh,
                                                      int64_t n, int* __restrict__ block_count, int* __restrict__ inside_count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = i < n && inside[i];
  const bool ok = in && val[i] > thresh;
  const int c = __syncthreads_count(ok), ci = __syncthreads_count(in);
  if (threadIdx.x == 0) { block_count[blockIdx.x] = c; if (ci) atomicAdd(inside_count, ci); }
}
__global__ void __launch_bounds__(1024) k_select_scan(int* __restrict__ block_count, int n_blocks, int* __restrict__ total) {
  __shared__ int s_carry;
  __shared__ int s_warp[32];
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n_blocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n_blocks ? block_count[i] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if ((threadIdx.x & 31) >= o) inc += t; }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = inc;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) woff += s_warp[w];
    const int excl = s_carry + woff + inc - v;
    if (i < n_blocks) block_count[i] = excl;   // exclusive prefix: first output slot of the block
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_carry;
}
__global__ void __launch_bounds__(256) k_select_scatter(const double* __restrict__ val, const unsigned char* __restrict__ inside, double thresh, int64_t n, int d,
                                                        const int* __restrict__ block_first, const double* __restrict__ u, const double* __restrict__ theta,
                                                        int64_t max_keep, double* __restrict__ u_out, double* __restrict__ theta_out, double* __restrict__ val_out) {
  __shared__ int s_warp[8];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = i < n && inside[i] && val[i] > thresh;
  const unsigned ballot = __ballot_sync(0xffffffffu, ok);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_warp[warp] = __popc(ballot);
  __syncthreads();
  int slot = block_first[blockIdx.x] + __popc(ballot & ((1u << lane) - 1u));
  for (int w = 0; w < warp; w++) slot += s_warp[w];
  if (ok && slot < max_keep) {
    for (int r = 0; r < d; r++) { u_out[(int64_t)slot * d + r] = u[i * d + r]; theta_out[(int64_t)slot * d + r] = theta[i * d + r]; }
    val_out[slot] = val[i];
  }
}

}  // namespace cosmolike
