// chi2_ozaki.cuh — stage 3 on the 5th-generation tensor cores: chi2_sn[b] = | W r_b |^2 with W = L^-1, computed as an
// error-free sliced integer contraction (Ozaki scheme) with tcgen05.mma kind::i8 and int32 accumulators in TMEM.
//
// Same contract as chi2_gemm.cuh (the reference's solve_triangular.py:5-14 forward substitution + y.y, batched), other
// arithmetic: Blackwell's tcgen05 has no FP64 kind, so every residual row and every row of W is written as
//     x = 2^e * sum_i d_i 2^-(6 + 8 i),   d_i in [-128, 127] (balanced base-256 digits of a 6 + 8 (S-1) bit fixed-point number)
// with one exponent e per row.  The S x S slice products are exact in int32; products whose weight lies below the last
// kept digit (i + j > S - 1) are dropped, the rest are accumulated per level l = i + j in S TMEM accumulators and
// recombined in FP64 by the epilogue: y_bn = 2^(eR_b + eW_n - 12) sum_l 2^-8l acc_l.  With S = 6 (46 bits, 21 slice
// products) the result agrees with the FP64 contraction to ~3e-13 relative (|d chi2| ~ 3e-9 at chi2 ~ 4e4).
//
// Kernel: persistent CTAs, 11 warps.  warp 0 = scheduler + TMA producer (3-D boxes {64 B of k, rows, S planes}, 64-byte
// swizzle, separate shared-memory rings for the R planes (A, 128 rows) and the W planes (B, NT rows); the item counter is
// read one tile ahead and the decoded item goes through the queue); warps 1 and 10 = MMA issuers (one thread each, one
// K = 32 half of every block each; one MMA covers up to 256 / NT stacked planes of W); warps 2-9 = epilogue (tcgen05.ld
// in half-level steps, FP64 recombination, row sum of squares; two barrier waits per tile and ONE hand-over arrival).
// Hand-over facts measured with the in-kernel event trace (OZ_PROF=1, `dbg` bit 3, tools/oz_trace_view.py) and
// tools/ubench_mbar.cu: an mbarrier poll costs 13 cycles even with ten other warps polling; a tcgen05.fence costs an
// epilogue warp ~300 cycles, so fences only follow actual waits; the levels of a tile complete in order, so the epilogue
// waits for level 0 and then for the last level, and the issuer polls one barrier (the level read last), not S.
// Triangular structure as in chi2_gemm.cuh: column tiles aligned to the end of the matrix, k stops at the diagonal
// block (the zeros of W above the diagonal inside that block are simply multiplied: ~4 % of the executed products).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

#include "chi2_gemm.cuh"
#include "digits.cuh"

namespace cosmolike {

constexpr int kOzM = 128;      // theta rows per tile (TMEM lanes)
constexpr int kOzKB = 64;      // bytes of k per pipeline block (= TMA inner box = swizzle span), two K=32 MMAs
constexpr int kOzThreads = 352;   // producer warp, MMA warp A, 8 epilogue warps, MMA warp B
constexpr int kOzIssuerB = 10;     // warp index of the second MMA issuer
constexpr int kOzQueue = 4;
constexpr int kOzBStages = 2;

template <int S> struct OzCfg {
  static constexpr int NT = S <= 5 ? 96 : S == 6 ? 80 : 64;  // S levels x NT columns <= 512 TMEM columns
  // R planes per TMA box / ring unit.  SLO = S: one box per k block.  (SLO = (S + 1) / 2 lets the first half of the planes
  // go back to the producer half way through the block, but two 24 KB boxes move measurably slower than one 48 KB box:
  // 2.84 M cycles per CTA instead of 2.61 M at N = 1701, so the split is off.)
  static constexpr int SLO = S;
  static constexpr int PLANE_BYTES = kOzM * kOzKB;
  static constexpr int A_UNIT_BYTES = SLO * PLANE_BYTES;
  static constexpr int B_BYTES = S * NT * kOzKB;
  static constexpr int A_UNITS = (3 * A_UNIT_BYTES + kOzBStages * B_BYTES <= 224 * 1024) ? 3 : 2;   // 7 planes: 3 x 56 KB + 2 x 28 KB = 224 KB, 231296 B with the rest (limit 232448)
  static constexpr int A_RING = A_UNITS * A_UNIT_BYTES, B_RING = kOzBStages * B_BYTES;
  static constexpr int SMEM = 1024 + A_RING + B_RING + 3 * NT * 8 + 384;   // 7 planes: 232320 B (limit 232448)
  static constexpr int FRAC_BITS = 6 + 8 * (S - 1);
  static constexpr int MAX_STACK = 256 / NT;   // digit planes of W one MMA may cover (N <= 256)
};

struct OzArgs {
  int64_t B;               // rows of R in this pass
  int N;                   // SN count
  int T;                   // column tiles = ceil(N / NT)
  int n_rb;                // row blocks = ceil(B / 128)
  double* part;            // [2 T][B] partial sums of squares (one plane per half column tile)
  const double* rowscale;  // [B] 2^eR_b
  const double* colscale;  // [N] 2^eW_n
  double* part_u;          // nullable [2 T][B]: partial sums of y_n u_n (offset moments)
  const double* u;         // [N] u = W 1
  int* counter;            // dynamic scheduling counter (zeroed before the launch)
  int group_rb;            // row blocks per L2 group
  long long* prof;         // nullable [grid][8]: cycle counters of the MMA issuer and of one epilogue warp
  long long* trace;        // nullable [1 + 16384]: event trace of CTA 0 (count, then (code << 44) | clock), dbg bit 3
  int dbg_skip;            // timing experiments (results invalid): bit 0 skip the B loads, bit 1 skip the A loads after the first blocks
};

// OZ_PROF=1 compiles the cycle counters (dbg bit 2) and the event trace (dbg bit 3) in; the production build carries neither
#ifndef OZ_PROF
#define OZ_PROF 0
#endif
__device__ __forceinline__ long long oz_clock() {
#if OZ_PROF
  return clock64();
#else
  return 0;
#endif
}
// event trace of CTA 0: three writers (issuer A, issuer B, epilogue warp 2 lane 0), each with its own region and its own
// running index (plain stores: no round trip on the traced thread)
constexpr int kOzTraceCap = 1400;   // events per warp (11 regions)
__device__ __forceinline__ void oz_trace(const OzArgs& g, int& idx, int code) {
  if (OZ_PROF && g.trace && blockIdx.x == 0 && idx < kOzTraceCap) {
    const int role = threadIdx.x >> 5;
    g.trace[1 + role * kOzTraceCap + idx] = ((long long)code << 44) | (clock64() & ((1LL << 44) - 1));
    idx++;
  }
}
__device__ __forceinline__ void oz_decode_item(const OzArgs& g, int item, int& jt, int& rb) {   // items < 2^31 (checked on the host)
  const int per_group = g.group_rb * g.T;
  const int n_full = g.n_rb / g.group_rb;
  int grp = item / per_group;
  int r, width;
  if (grp < n_full) { r = item - grp * per_group; width = g.group_rb; }
  else { grp = n_full; r = item - n_full * per_group; width = g.n_rb - n_full * g.group_rb; }
  const int q = r / width;
  jt = g.T - 1 - q;
  rb = grp * g.group_rb + (r - q * width);
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// shared-memory matrix descriptor: K-major tile of 64-byte rows, 64-byte swizzle, 8-row groups 512 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
// instruction descriptor: int8 x int8 -> int32, both operands K-major, dense
__device__ __forceinline__ uint32_t umma_idesc_i8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// The same with the descriptors given by their low words (start address >> 4 | LBO field): the high word of every
// descriptor of this kernel is the constant {SBO = 512 B, version 1, 64-byte swizzle}.  One thread issues every MMA of the
// CTA, so the instructions spent per issue bound the kernel: with the address fields precomputed per ring slot an issue is
// two integer adds instead of two shift / mask / or chains.
constexpr uint32_t kDescHiSw64 = (512u >> 4) | (1u << 14) | (4u << 29);
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ void umma_i8_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 da, {%1, %5};\nmov.b64 db, {%2, %5};\n"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %3, p;\n}\n"
               ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHiSw64) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// asynchronous TMEM load: the registers are valid only after tmem_ld_wait() AND tmem_ld_fence() on them
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                 "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, int32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_fence4(int32_t* v) { asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]) :: "memory"); }
__device__ __forceinline__ void tmem_ld_fence8(int32_t* v) {
  asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]) :: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// empty volatile asm that "modifies" 16 loaded registers: pins every use of them after the wait above (volatile asm
// statements keep their order; plain arithmetic on the registers would otherwise be free to move above the wait)
__device__ __forceinline__ void tmem_ld_fence(int32_t* v) {
  asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                    "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]) :: "memory");
}
// bounded spin: a wrong descriptor or a lost arrival traps (the context dies with an error) instead of hanging the GPU
__device__ __forceinline__ void oz_wait(uint32_t bar, uint32_t parity) {
  for (long long i = 0; i < (1LL << 28); i++) if (mbar_try_wait(bar, parity)) return;
  __trap();
}

// long waits (the epilogue's first level of a tile, the producer's ring slots): poll with a pause in between - the waiting
// warps share their SM sub-partitions with the two MMA issuers
__device__ __forceinline__ void oz_wait_relaxed(uint32_t bar, uint32_t parity, unsigned ns) {
  for (long long i = 0; i < (1LL << 28); i++) {
    if (mbar_try_wait(bar, parity)) return;
    __nanosleep(ns);
  }
  __trap();
}
#ifndef OZ_EPI_NS
#define OZ_EPI_NS 100   // pause between the epilogue's polls for the first level of a tile
#endif
#ifndef OZ_SPIN_NS
#define OZ_SPIN_NS 40
#endif
// latency-critical waits of the tile hand-over: non-blocking test_wait in a tight loop (try_wait may park the thread)
__device__ __forceinline__ void oz_spin(uint32_t bar, uint32_t parity) {
  for (long long i = 0; i < (1LL << 28); i++) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    __nanosleep(OZ_SPIN_NS);
  }
  __trap();
}

// ---- slicing: fp64 rows -> S int8 digit planes + one power-of-two scale per row --------------------------------------
// dst[s][row][ld] (int8), scale[row] = 2^e with |x| < 2^e for the whole row.  One warp per row.
template <int S>
__global__ void __launch_bounds__(256) k_oz_slice_rows(const double* __restrict__ src, int64_t ld_src, int64_t rows, int n, int8_t* __restrict__ dst,
                                                       int64_t ld_dst, double* __restrict__ scale) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const double* x = src + row * ld_src;
  double mx = 0.0;
  bool bad = false;   // a NaN / Inf residual (log10 of a non-positive distance, as in the reference) must give chi2 = NaN
  for (int k = lane; k < n; k += 32) { const double a = fabs(x[k]); bad |= !(a <= 1.7e308); mx = fmax(mx, a); }
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  bad = __any_sync(0xffffffffu, bad);
  int e = 0;
  if (mx > 0.0 && mx < 1.7e308) e = ilogb(mx) + 1;
  e = max(e, -900);
  const double up = __longlong_as_double((long long)(1023 + OzCfg<S>::FRAC_BITS - e) << 52);   // 2^(FRAC_BITS - e)
  if (lane == 0) scale[row] = bad ? __longlong_as_double(0x7ff8000000000000LL) : __longlong_as_double((long long)(1023 + e) << 52);
  constexpr double kLim = 1.01 * (double)(1ULL << OzCfg<S>::FRAC_BITS);
  const int64_t plane = rows * ld_dst;
  int8_t* d0 = dst + row * ld_dst;
  // four consecutive k per lane -> one 32-bit store per digit plane
  for (int k0 = 4 * lane; k0 < ld_dst; k0 += 128) {
    uint32_t w[S];
    long long v[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int k = k0 + q;
      v[q] = 0;
      if (k < n) v[q] = __double2ll_rn(fmin(fmax(x[k] * up, -kLim), kLim));   // |v| <= 2^FRAC_BITS for finite rows; garbage rows are clamped
    }
    oz_digits4<S>(v, w);
#pragma unroll
    for (int s = 0; s < S; s++) *reinterpret_cast<uint32_t*>(d0 + s * plane + k0) = w[s];
  }
}

// Same, one pass: the whole row lives in registers (TRIPS x 4 doubles per lane, row length <= 128 TRIPS), so the fp64
// residuals are read from HBM once (streaming loads: they are not needed again).
template <int S, int TRIPS>
__global__ void __launch_bounds__(256) k_oz_slice_rows_reg(const double* __restrict__ src, int64_t ld_src, int64_t rows, int n, int8_t* __restrict__ dst,
                                                           int64_t ld_dst, double* __restrict__ scale) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const double* x = src + row * ld_src;
  double2 xa[TRIPS], xb[TRIPS];
  double mx = 0.0;
  bool bad = false;   // NaN / Inf residual -> chi2 = NaN (see k_oz_slice_rows)
#pragma unroll
  for (int t = 0; t < TRIPS; t++) {
    const int k0 = 4 * lane + 128 * t;
    xa[t] = make_double2(0.0, 0.0); xb[t] = make_double2(0.0, 0.0);
    if (k0 + 3 < n) {   // ld_src >= n rounded up to 2 and 32-byte aligned groups: both halves are in bounds
      xa[t] = __ldcs(reinterpret_cast<const double2*>(x + k0));
      xb[t] = __ldcs(reinterpret_cast<const double2*>(x + k0 + 2));
    } else {
      if (k0 < n) xa[t].x = x[k0];
      if (k0 + 1 < n) xa[t].y = x[k0 + 1];
      if (k0 + 2 < n) xb[t].x = x[k0 + 2];
    }
    const double a0 = fabs(xa[t].x), a1 = fabs(xa[t].y), a2 = fabs(xb[t].x), a3 = fabs(xb[t].y);
    bad |= !(a0 <= 1.7e308) | !(a1 <= 1.7e308) | !(a2 <= 1.7e308) | !(a3 <= 1.7e308);
    mx = fmax(fmax(mx, fmax(a0, a1)), fmax(a2, a3));
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  bad = __any_sync(0xffffffffu, bad);
  int e = 0;
  if (mx > 0.0 && mx < 1.7e308) e = ilogb(mx) + 1;
  e = max(e, -900);
  const double up = __longlong_as_double((long long)(1023 + OzCfg<S>::FRAC_BITS - e) << 52);
  if (lane == 0) scale[row] = bad ? __longlong_as_double(0x7ff8000000000000LL) : __longlong_as_double((long long)(1023 + e) << 52);
  constexpr double kLim = 1.01 * (double)(1ULL << OzCfg<S>::FRAC_BITS);
  const int64_t plane = rows * ld_dst;
  int8_t* d0 = dst + row * ld_dst;
#pragma unroll
  for (int t = 0; t < TRIPS; t++) {
    const int k0 = 4 * lane + 128 * t;
    if (k0 >= ld_dst) break;
    const double xs[4] = {xa[t].x, xa[t].y, xb[t].x, xb[t].y};
    uint32_t w[S];
    long long v[4];
#pragma unroll
    for (int q = 0; q < 4; q++) v[q] = __double2ll_rn(fmin(fmax(xs[q] * up, -kLim), kLim));
    oz_digits4<S>(v, w);
#pragma unroll
    for (int s = 0; s < S; s++) *reinterpret_cast<uint32_t*>(d0 + s * plane + k0) = w[s];
  }
}

// ---- the contraction -----------------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(kOzThreads, 1)
k_chi2_ozaki(const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmW, const OzArgs g) {
  using C = OzCfg<S>;
  constexpr int NT = C::NT;
  // (1.5 * 2^52 + 2^31) * (sum of the weights of the level groups behind the first): see the epilogue's fold
  constexpr double kFoldC = 6755401588539392.0 * (S == 7 ? (0x1p-40 + 0x1p-48) : S == 6 ? 0x1p-40 : 0x1p-32);
  extern __shared__ unsigned char osm_raw[];
  const uint32_t base = (smem_u32(osm_raw) + 1023u) & ~1023u;
  unsigned char* base_ptr = osm_raw + (base - smem_u32(osm_raw));
  const uint32_t sA = base;                       // [A_UNITS][SLO planes][128 rows][64 B]
  const uint32_t sB = base + C::A_RING;           // [kOzBStages][S planes][NT rows][64 B]
  double* s_cs = reinterpret_cast<double*>(base_ptr + C::A_RING + C::B_RING);  // [NT] column scales of the tile
  double* s_u = s_cs + NT;                                                     // [NT] u of the tile's columns (moments)
  double* s_cc = s_u + NT;                                                     // [NT] -(bias constant) * column scale, see the epilogue
  const uint32_t bars = base + C::A_RING + C::B_RING + 3 * NT * 8;
  auto fullA = [&](int s) { return bars + 8u * s; };
  auto emptyA = [&](int s) { return bars + 8u * (6 + s); };
  auto fullB = [&](int s) { return bars + 8u * (12 + s); };
  auto emptyB = [&](int s) { return bars + 8u * (14 + s); };
  auto qfull = [&](int s) { return bars + 8u * (16 + s); };
  auto qempty = [&](int s) { return bars + 8u * (20 + s); };
  auto lvl_full = [&](int l) { return bars + 8u * (24 + l); };    // level l of the tile is complete (tcgen05.commit)
  auto lvl_empty = [&](int l) { return bars + 8u * (31 + l); };   // level l has been read by the eight epilogue warps
  const uint32_t q_items = bars + 8u * 38;   // int[kOzQueue]
  const uint32_t tslot = bars + 8u * 40;
  const uint32_t touched = bars + 8u * 41;   // issuer A has issued the first-touch MMAs of the tile

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < C::A_UNITS; s++) { mbar_init(fullA(s), 1); mbar_init(emptyA(s), 2); }   // released by both issuers' commits
    for (int s = 0; s < kOzBStages; s++) { mbar_init(fullB(s), 1); mbar_init(emptyB(s), 2); }
    mbar_init(touched, 1);
    for (int s = 0; s < kOzQueue; s++) { mbar_init(qfull(s), 1); mbar_init(qempty(s), 10); }   // 2 MMA threads + 8 epilogue warps
    for (int l = 0; l < S; l++) { mbar_init(lvl_full(l), 2); mbar_init(lvl_empty(l), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot) : "memory");

  const int64_t total = (int64_t)g.n_rb * g.T;
  int qslot = 0;
  uint32_t qphase = 0;

  if (warp == 0) {
#if OZ_PROF
    // observer (profiling build): lane 1 records when the LAST level of every tile completes (the end of the tile's MMAs)
    if (lane == 1 && g.trace && blockIdx.x == 0) {
      int tr_i = 0;
      uint32_t p = 0;
      for (int n = 0; n < kOzTraceCap; n++) {
        bool seen = false;
        for (long long i = 0; i < (1LL << 24) && !seen; i++) {
          uint32_t ok;
          asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(lvl_full(S - 1)), "r"(p) : "memory");
          seen = ok != 0;
        }
        if (!seen) break;
        oz_trace(g, tr_i, 70);
        p ^= 1u;
      }
    }
#endif
    // ===================== scheduler + TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      // The item counter is read one tile AHEAD: the atomic's round trip (148 CTAs on one address), the decode and the TMA
      // latency of a new tile's first blocks otherwise follow the previous tile's last load back to back, and the three-slot
      // ring only hides two blocks of that (it showed as the issuers waiting ~20 % of their time for the first blocks of a tile).
      int next = atomicAdd(g.counter, 1);
      for (;;) {
        const int item = next;
        const bool done = item >= total;
        if (!done) next = atomicAdd(g.counter, 1);   // in flight while this tile's loads are issued
        oz_wait(qempty(qslot), qphase ^ 1u);
        int jt = 0, rb = 0;
        if (!done) oz_decode_item(g, item, jt, rb);
        // the decoded item goes through the queue ((jt << 24) | rb): the 64-bit divisions stay off the consumers' paths
        asm volatile("st.shared.s32 [%0], %1;" ::"r"(q_items + 4u * qslot), "r"(done ? -1 : ((jt << 24) | rb)) : "memory");
        mbar_arrive(qfull(qslot));
        if (++qslot == kOzQueue) { qslot = 0; qphase ^= 1u; }
        if (done) break;
        const int c0 = g.N - NT * (g.T - jt);              // first column of the tile (< 0 only for jt == 0: TMA zero-fills)
        const int nk = (c0 + NT + kOzKB - 1) / kOzKB;      // k runs to the end of the diagonal block
        for (int ks = 0; ks < nk; ks++) {
          oz_wait_relaxed(emptyA(sa), pa ^ 1u, 50);
          if ((g.dbg_skip & 2) && ks >= 3) mbar_arrive(fullA(sa));
          else {
            mbar_arrive_expect_tx(fullA(sa), C::A_UNIT_BYTES);
            tma_load_3d(sA + sa * C::A_UNIT_BYTES, &tmR, ks * kOzKB, rb * kOzM, 0, fullA(sa));
          }
          if (++sa == C::A_UNITS) { sa = 0; pa ^= 1u; }
          oz_wait_relaxed(emptyB(sb), pb ^ 1u, 50);
          if ((g.dbg_skip & 1) && ks >= 2) mbar_arrive(fullB(sb));
          else {
            mbar_arrive_expect_tx(fullB(sb), C::B_BYTES);
            tma_load_3d(sB + sb * C::B_BYTES, &tmW, ks * kOzKB, c0, 0, fullB(sb));
          }
          if (++sb == kOzBStages) { sb = 0; pb ^= 1u; }
          if constexpr (C::SLO < S) {
            oz_wait(emptyA(sa), pa ^ 1u);
            mbar_arrive_expect_tx(fullA(sa), C::A_UNIT_BYTES);   // planes past S are out of bounds: zero-filled, still counted
            tma_load_3d(sA + sa * C::A_UNIT_BYTES, &tmR, ks * kOzKB, rb * kOzM, C::SLO, fullA(sa));
            if (++sa == C::A_UNITS) { sa = 0; pa ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1 || warp == kOzIssuerB) {
    // ===================== MMA issuers (two threads) =====================
    // One MMA covers SEVERAL slice products: the digit planes of W are stacked in shared memory ([plane][NT rows]) and the
    // levels are stacked in TMEM ([level][NT columns]), so A_i x [B_j0 ; ... ; B_j1] with N = (j1 - j0 + 1) NT lands exactly
    // on levels i + j0 .. i + j1.  S (S + 1) / 2 products per K = 32 step become ~S + 3 instructions that read the A plane
    // from shared memory once each (shared-memory bandwidth, not the tensor pipe, bounds the narrow form).  Round i completes
    // level i: in the last k block each level is committed on its own barrier as soon as its round has been issued, so the
    // epilogue overlaps the rest of the block.
    // The MMA queue is shallow: an issuing thread is held back until the pipe has taken its previous MMAs, and between two
    // MMAs it spends ~100 cycles of dependent uniform-datapath work while the pipe needs ~90 per instruction of this schedule.
    // The issue is therefore split: thread A (warp 1) issues the first K = 32 half of every block, thread B (warp kOzIssuerB)
    // the second (tensor pipe active 67.8 % -> 70.6 % of the kernel).  Integer accumulation commutes; the only order that matters is that A's
    // first-touch MMAs of a tile (accumulate = 0) enter the pipe before B's first MMAs of that tile: B waits for `touched`,
    // on which A arrives after issuing round 0 of the tile's first block (the pipe executes in issue order).  Every
    // tcgen05.commit only tracks the MMAs of its own thread, so the release barriers count two arrivals.
    // ONE copy of the issue code serves every block, both issuers and every ring position (run-time descriptors): a variant
    // with compile-time ring slots (12 instantiations, plain R2UR instead of the ELECT / R2UR.BROADCAST loop around every
    // tcgen05.mma) has half the instructions per MMA and is 3-4 % SLOWER - the issue rate is set by the pipe's back-pressure
    // (the MMA queue is shallow), not by the instruction count, and the larger code costs instruction-cache misses.
    if (lane == 0) {
      const int H = warp == 1 ? 0 : 1;
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0, pt = 0;
      long long t_full = 0, t_lvl = 0, t_q = 0, n_blk = 0;
      int tr_i = 0;
      const long long t_begin = oz_clock();
      for (;;) {
        long long tq0 = oz_clock();
        oz_wait(qfull(qslot), qphase);
        t_q += oz_clock() - tq0;
        int item;
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(item) : "r"(q_items + 4u * qslot) : "memory");
        mbar_arrive(qempty(qslot));
        if (++qslot == kOzQueue) { qslot = 0; qphase ^= 1u; }
        if (item < 0) break;
        const int jt = item >> 24;
        const int c0 = g.N - NT * (g.T - jt);
        const int nk = (c0 + NT + kOzKB - 1) / kOzKB;
        oz_trace(g, tr_i, 1000 * H + 1);   // tile start (item known)
        for (int ks = 0; ks < nk; ks++) {
          long long tf0 = oz_clock();
          oz_wait(fullA(sa), pa);
          oz_wait(fullB(sb), pb);
          t_full += oz_clock() - tf0; n_blk++;
          const bool first_blk = ks == 0, last_blk = ks == nk - 1;
          if (last_blk) oz_trace(g, tr_i, 1000 * H + 40);   // last block: operands resident
          if (first_blk) {
            // first touch of the tile: the epilogue must have read the previous tile's levels (A polls the barrier of the
            // level that is read last), and A's accumulate = 0 MMAs must be in the pipe before B's first MMAs (B polls `touched`)
            long long tl0 = oz_clock();
            if (H == 0) oz_spin(lvl_empty(S - 1), pt ^ 1u); else oz_spin(touched, pt);
            t_lvl += oz_clock() - tl0;
            oz_trace(g, tr_i, 1000 * H + 10);
          }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          // descriptor low words of this block's ring slots, at this thread's K = 32 half
          const uint32_t aD0 = umma_desc_lo(sA + sa * C::A_UNIT_BYTES + 32 * H), bD0 = umma_desc_lo(sB + sb * C::B_BYTES + 32 * H);
          const bool zero = H == 0 && first_blk;
#pragma unroll
          for (int i = 0; i < S; i++) {
#pragma unroll
            for (int j0 = 0; j0 < S - i; j0 += C::MAX_STACK) {
              const int cnt = (S - i - j0) < C::MAX_STACK ? (S - i - j0) : C::MAX_STACK;
              umma_i8_lo(tmem + (uint32_t)((i + j0) * NT), aD0 + (uint32_t)((i * C::PLANE_BYTES) >> 4),
                         bD0 + (uint32_t)((j0 * (NT * kOzKB)) >> 4), umma_idesc_i8(kOzM, cnt * NT), (zero && i == 0) ? 0u : 1u);
            }
            if (zero && i == 0) mbar_arrive(touched);   // the accumulate = 0 MMAs are in the pipe: B may follow
            if (last_blk) umma_commit(lvl_full(i));     // round i completes level i (a commit after the round's FIRST instruction, which already completes it, measured slower)
          }
          umma_commit(emptyA(sa));
          umma_commit(emptyB(sb));
          if (first_blk) oz_trace(g, tr_i, 1000 * H + 20);   // first block issued
          if (++sa == C::A_UNITS) { sa = 0; pa ^= 1u; }
          if (++sb == kOzBStages) { sb = 0; pb ^= 1u; }
        }
        oz_trace(g, tr_i, 1000 * H + 41);   // tile issued
        pt ^= 1u;
      }
      if (OZ_PROF && g.prof && H == 0) {
        long long* p = g.prof + 8 * blockIdx.x;
        p[0] = oz_clock() - t_begin; p[1] = t_full; p[2] = t_lvl; p[3] = t_q; p[4] = n_blk;
      }
    }
  } else {
    // ===================== epilogue: TMEM -> FP64 recombination -> row sum of squares =====================
    // two warps per TMEM lane group (a warp may only read lanes 32 (warp % 4) ..): each takes half of the tile's columns
    constexpr int NH = NT / 2;
    const int lg = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row_in_tile = lg * 32 + lane;
    const int etid = tid - 64;               // 0..255
    uint32_t pt = 0;
    long long t_wait = 0, t_tiles = 0, t_ld = 0, t_pre = 0;
    int tr_i = 0;
    for (;;) {
      if (lane == 0) oz_wait_relaxed(qfull(qslot), qphase, 100);
      __syncwarp();
      int item;
      asm volatile("ld.shared.s32 %0, [%1];" : "=r"(item) : "r"(q_items + 4u * qslot) : "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(qempty(qslot));
      if (++qslot == kOzQueue) { qslot = 0; qphase ^= 1u; }
      if (item < 0) {
        if (OZ_PROF && g.prof && tid == 64) { g.prof[8 * blockIdx.x + 5] = t_wait; g.prof[8 * blockIdx.x + 6] = t_tiles; g.prof[8 * blockIdx.x + 7] = t_ld; g.prof[8 * blockIdx.x + 4] += t_pre << 32; }
        break;
      }
      t_tiles++;
      const int jt = item >> 24, rb = item & 0xffffff;
      const int c0 = g.N - NT * (g.T - jt);
      long long tp0 = oz_clock();
      // column scales of this tile (the previous tile's readers are past them: both barriers below order the rewrite after their last read)
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (etid < NT) {
        const int col = c0 + etid;
        const double cs = col >= 0 ? g.colscale[col] : 0.0;
        s_cs[etid] = cs;
        s_cc[etid] = -kFoldC * cs;   // exact: the column scale is a power of two
        if (g.part_u) s_u[etid] = col >= 0 ? g.u[col] : 0.0;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      t_pre += oz_clock() - tp0;
      // h_n = sum_l 2^-8l acc_l[n].  The levels are folded THREE AT A TIME in integer arithmetic: a group value
      // g = acc_l 2^16 + acc_(l+1) 2^8 + acc_(l+2) (|g| < 2^47 for n_sn <= 16384) is built by two IMAD.WIDE on top of the
      // 1.5 * 2^52 bias pattern, so one DADD turns it into a double and one DFMA adds it to h: 2 FP64 instructions per
      // column for three levels instead of 6.  The epilogue's time after the tile's last MMA is what the next tile waits for
      // (all of TMEM is in use), and it was bound by the FP64 pipe (14 FP64 instructions per column: 1800 cycles per tile)
      // and by one load-wait-fold round trip per half level; tcgen05.ld itself moves 300 B/clk/SM (tools/ubench_tmem_ld.cu:
      // 750 cycles for the whole tile).  Loads are software-pipelined in chunks of CH columns (all levels of a group per
      // chunk); a group is read as soon as its LAST level is complete - the levels of a tile complete in order.
      constexpr int CH = 8;
      static_assert(NH % CH == 0, "chunks of eight columns");
      constexpr int NCH = NH / CH, NG = (S + 2) / 3;
      // FP64 work per column: ONE DFMA per group.  A group value arrives as the double t = kBias + g (the bias pattern with g
      // in the low mantissa bits).  Group 0 removes its bias inside the DFMA, h = fma(t, w0, -kBias w0) = g w0 exactly; the
      // later groups are added bias and all, h = fma(t, w, h), which carries kBias w along - small constants (1.5 * 2^12 and
      // 1.5 * 2^4 at S = 7) next to which h keeps every bit it would keep anyway: if |h| is below them the sum's ulp is 2^-40
      // (only the lowest digits of level 6 round, an absolute 2^-41 on a quantity whose TERMS carry errors of 2^-39 by the
      // a-priori bound), if |h| is above them the rounding is the usual relative 2^-53.  Their exactly representable sum
      // kFoldC comes off together with the column scale, y = fma(h, cs, -kFoldC cs), one rounding.  5 FP64 instructions per
      // column instead of 8 (the epilogue's ~3.6 k cycles per tile are half FP64-pipe time, and what the next tile waits for).
      double h[NH];
      int32_t buf[2][3][CH];
      const uint32_t trow = tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(half * NH);
      auto issue_chunk = [&](int l0, int cnt, int c, int32_t (&dst)[3][CH]) {
        if (g.dbg_skip & 8) return;   // timing experiment: no TMEM loads (results invalid)
#pragma unroll
        for (int q = 0; q < 3; q++)
          if (q < cnt) tmem_ld8(trow + (uint32_t)((l0 + q) * NT + c * CH), dst[q]);
      };
      auto land_chunk = [&](int cnt, int32_t (&cur)[3][CH]) {
        long long tl0 = oz_clock();
        tmem_ld_wait();
        t_ld += oz_clock() - tl0;
#pragma unroll
        for (int q = 0; q < 3; q++)
          if (q < cnt) tmem_ld_fence8(cur[q]);
      };
#pragma unroll
      for (int n = 0; n < NH; n++) h[n] = 0.0;
      // group boundaries {0,1,2} {3,4,5} {6}: measured against {0,1,2} {3,4} {5,6} (a smaller tail after the tile's last MMA,
      // but one more FP64 pair per column): 1.744 vs 1.756 ms at S = 7
      constexpr int kG0[3] = {0, 3, 6};
      constexpr double kBias = 6755401588539392.0;   // 1.5 * 2^52 + 2^31
#pragma unroll
      for (int gi = 0; gi < NG; gi++) {
        const int l0 = kG0[gi], cnt = (gi + 1 < NG ? kG0[gi + 1] : S) - l0;
        // wait for the group's last level (relaxed poll for the tile's first group, which starts under the MMAs of the last k block)
        {
          long long tw0 = oz_clock();
          if (lane == 0) { if (gi == 0) oz_wait_relaxed(lvl_full(l0 + cnt - 1), pt, OZ_EPI_NS); else oz_spin(lvl_full(l0 + cnt - 1), pt); }
          __syncwarp();
          t_wait += oz_clock() - tw0;
          if (lane == 0 && gi == 0) oz_trace(g, tr_i, 2000 + 100 * (warp - 2));   // first group complete (seen by this warp)
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const double wg = __longlong_as_double((long long)(1023 - 8 * (l0 + cnt - 1)) << 52);   // weight of the group's lowest level
        issue_chunk(l0, cnt, 0, buf[0]);
#pragma unroll
        for (int c = 0; c < NCH; c++) {
          land_chunk(cnt, buf[c & 1]);
          if (c + 1 < NCH) issue_chunk(l0, cnt, c + 1, buf[(c + 1) & 1]);
          if (gi == NG - 1 && c == NCH - 1) {
            // hand-over: every TMEM read of the tile has landed; the issuer polls this one barrier
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(lvl_empty(S - 1));
            if (lane == 0) oz_trace(g, tr_i, 2010 + (S - 1) + 100 * (warp - 2));   // last level in registers
          }
          if ((g.dbg_skip & 4) && gi > 0) continue;   // timing experiment: no folds beyond the first group (results invalid)
#pragma unroll
          for (int n = 0; n < CH; n++) {
            // {low word: lowest level + 2^31, high word: 1.5 * 2^52} + higher levels * 2^8, 2^16
            long long t = ((long long)0x43380000 << 32) | (unsigned long long)(uint32_t)(buf[c & 1][cnt - 1][n] ^ (int)0x80000000);
            if (cnt >= 2) asm("mad.wide.s32 %0, %1, %2, %0;" : "+l"(t) : "r"(buf[c & 1][cnt - 2][n]), "r"(256));
            if (cnt >= 3) asm("mad.wide.s32 %0, %1, %2, %0;" : "+l"(t) : "r"(buf[c & 1][cnt - 3][n]), "r"(65536));
            h[c * CH + n] = fma(__longlong_as_double(t), wg, gi == 0 ? -kBias * wg : h[c * CH + n]);
          }
        }
      }
      if (lane == 0) oz_trace(g, tr_i, 2020 + 100 * (warp - 2));   // all levels recombined
      pt ^= 1u;
      double acc = 0.0;
#pragma unroll
      for (int n = 0; n < NH; n++) {
        const double y = fma(h[n], s_cs[half * NH + n], s_cc[half * NH + n]);
        acc = fma(y, y, acc);
        h[n] = y;
      }
      double acc_u = 0.0;
      if (g.part_u) {
#pragma unroll
        for (int n = 0; n < NH; n++) acc_u = fma(h[n], s_u[half * NH + n], acc_u);
      }
      const int64_t row = (int64_t)rb * kOzM + row_in_tile;
      if (row < g.B) {
        const double rs = g.rowscale[row] * 0.000244140625;   // 2^eR_b * 2^-12 (fixed-point position of the digit products)
        g.part[(int64_t)(2 * jt + half) * g.B + row] = acc * rs * rs;
        if (g.part_u) g.part_u[(int64_t)(2 * jt + half) * g.B + row] = acc_u * rs;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace cosmolike
