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
// Kernel: persistent CTAs, 6 warps.  warp 0 = scheduler + TMA producer (3-D boxes {64 B of k, rows, S slices}, 64-byte
// swizzle, separate shared-memory rings for the R slices (A, 128 rows) and the W slices (B, NT rows)); warp 1 = TMEM
// allocator + single-thread MMA issuer; warps 2-5 = epilogue (tcgen05.ld, FP64 recombination, row sum of squares).
// Triangular structure as in chi2_gemm.cuh: column tiles aligned to the end of the matrix, k stops at the diagonal
// block, and inside the diagonal block the MMA N extent shrinks past the columns that are already complete.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "chi2_gemm.cuh"

namespace cosmolike {

constexpr int kOzM = 128;      // theta rows per tile (TMEM lanes)
constexpr int kOzKB = 64;      // bytes of k per pipeline block (= TMA inner box = swizzle span), two K=32 MMAs
constexpr int kOzThreads = 192;
constexpr int kOzQueue = 4;
constexpr int kOzAStages = 3, kOzBStages = 2;

template <int S> struct OzCfg {
  static constexpr int NT = S <= 5 ? 96 : S == 6 ? 80 : 64;  // S levels x NT columns <= 512 TMEM columns
  static constexpr int A_BYTES = S * kOzM * kOzKB;
  static constexpr int B_BYTES = S * NT * kOzKB;
  static constexpr int A_STAGES = (kOzAStages * A_BYTES + kOzBStages * B_BYTES <= 216 * 1024) ? kOzAStages : 2;
  static constexpr int SMEM = 1024 + A_STAGES * A_BYTES + kOzBStages * B_BYTES + NT * 8 + 256;
  static constexpr int FRAC_BITS = 6 + 8 * (S - 1);
};

struct OzArgs {
  int64_t B;               // rows of R in this pass
  int N;                   // SN count
  int T;                   // column tiles = ceil(N / NT)
  int n_rb;                // row blocks = ceil(B / 128)
  double* part;            // [T][B] partial sums of squares
  const double* rowscale;  // [B] 2^eR_b
  const double* colscale;  // [N] 2^eW_n
  int* counter;            // dynamic scheduling counter (zeroed before the launch)
  int group_rb;            // row blocks per L2 group
  int diag_trim;           // 1: shrink the MMA N extent inside the diagonal block
};

__device__ __forceinline__ void oz_decode_item(const OzArgs& g, int64_t item, int& jt, int& rb) {
  const int per_group = g.group_rb * g.T;
  const int n_full = g.n_rb / g.group_rb;
  int grp = (int)(item / per_group);
  int r, width;
  if (grp < n_full) { r = (int)(item % per_group); width = g.group_rb; }
  else { grp = n_full; r = (int)(item - (int64_t)n_full * per_group); width = g.n_rb - n_full * g.group_rb; }
  jt = g.T - 1 - r / width;
  rb = grp * g.group_rb + r % width;
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
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// load + wait in ONE asm statement: the registers are not valid before tcgen05.wait::ld and nothing else orders
// plain arithmetic on them after a separate wait statement
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
               "tcgen05.wait::ld.sync.aligned;"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                 "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
}
// bounded spin: a wrong descriptor or a lost arrival traps (the context dies with an error) instead of hanging the GPU
__device__ __forceinline__ void oz_wait(uint32_t bar, uint32_t parity) {
  for (long long i = 0; i < (1LL << 28); i++) if (mbar_try_wait(bar, parity)) return;
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
  for (int k = lane; k < n; k += 32) mx = fmax(mx, fabs(x[k]));   // fmax drops NaN: stale rows of skipped thetas stay harmless
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  int e = 0;
  if (mx > 0.0 && mx < 1.7e308) e = ilogb(mx) + 1;
  e = max(e, -900);
  const double up = __longlong_as_double((long long)(1023 + OzCfg<S>::FRAC_BITS - e) << 52);   // 2^(FRAC_BITS - e)
  if (lane == 0) scale[row] = __longlong_as_double((long long)(1023 + e) << 52);
  constexpr double kLim = 1.01 * (double)(1ULL << OzCfg<S>::FRAC_BITS);
  const int64_t plane = rows * ld_dst;
  int8_t* d0 = dst + row * ld_dst;
  // four consecutive k per lane -> one 32-bit store per digit plane
  for (int k0 = 4 * lane; k0 < ld_dst; k0 += 128) {
    uint32_t w[S];
#pragma unroll
    for (int s = 0; s < S; s++) w[s] = 0u;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int k = k0 + q;
      long long v = 0;
      if (k < n) {
        double t = x[k] * up;
        t = fmin(fmax(t, -kLim), kLim);   // |v| <= 2^FRAC_BITS for finite rows; garbage rows are clamped
        v = __double2ll_rn(t);
      }
#pragma unroll
      for (int s = S - 1; s >= 1; s--) {
        const long long d = (long long)(int8_t)(v & 0xff);
        w[s] |= (uint32_t)(uint8_t)d << (8 * q);
        v = (v - d) >> 8;
      }
      w[0] |= (uint32_t)(uint8_t)(int8_t)v << (8 * q);
    }
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
  extern __shared__ unsigned char osm_raw[];
  const uint32_t base = (smem_u32(osm_raw) + 1023u) & ~1023u;
  unsigned char* base_ptr = osm_raw + (base - smem_u32(osm_raw));
  const uint32_t sA = base;                                    // [A_STAGES][S][128 rows][64 B]
  const uint32_t sB = base + C::A_STAGES * C::A_BYTES;         // [kOzBStages][S][NT rows][64 B]
  double* s_cs = reinterpret_cast<double*>(base_ptr + C::A_STAGES * C::A_BYTES + kOzBStages * C::B_BYTES);  // [NT] column scales of the tile
  const uint32_t bars = base + C::A_STAGES * C::A_BYTES + kOzBStages * C::B_BYTES + NT * 8;
  auto fullA = [&](int s) { return bars + 8u * s; };
  auto emptyA = [&](int s) { return bars + 8u * (4 + s); };
  auto fullB = [&](int s) { return bars + 8u * (8 + s); };
  auto emptyB = [&](int s) { return bars + 8u * (10 + s); };
  auto qfull = [&](int s) { return bars + 8u * (12 + s); };
  auto qempty = [&](int s) { return bars + 8u * (16 + s); };
  const uint32_t tmem_full = bars + 8u * 20, tmem_empty = bars + 8u * 21;
  const uint32_t q_items = bars + 8u * 22;   // int[kOzQueue]
  const uint32_t tslot = bars + 8u * 24;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < C::A_STAGES; s++) { mbar_init(fullA(s), 1); mbar_init(emptyA(s), 1); }
    for (int s = 0; s < kOzBStages; s++) { mbar_init(fullB(s), 1); mbar_init(emptyB(s), 1); }
    for (int s = 0; s < kOzQueue; s++) { mbar_init(qfull(s), 1); mbar_init(qempty(s), 5); }   // MMA thread + 4 epilogue warps
    mbar_init(tmem_full, 1); mbar_init(tmem_empty, 4);
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
    // ===================== scheduler + TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (;;) {
        const int64_t item = atomicAdd(g.counter, 1);
        const bool done = item >= total;
        oz_wait(qempty(qslot), qphase ^ 1u);
        asm volatile("st.shared.s32 [%0], %1;" ::"r"(q_items + 4u * qslot), "r"(done ? -1 : (int)item) : "memory");
        mbar_arrive(qfull(qslot));
        if (++qslot == kOzQueue) { qslot = 0; qphase ^= 1u; }
        if (done) break;
        int jt, rb;
        oz_decode_item(g, item, jt, rb);
        const int c0 = g.N - NT * (g.T - jt);              // first column of the tile (< 0 only for jt == 0: TMA zero-fills)
        const int nk = (c0 + NT + kOzKB - 1) / kOzKB;      // k runs to the end of the diagonal block
        for (int ks = 0; ks < nk; ks++) {
          oz_wait(emptyA(sa), pa ^ 1u);
          mbar_arrive_expect_tx(fullA(sa), C::A_BYTES);
          tma_load_3d(sA + sa * C::A_BYTES, &tmR, ks * kOzKB, rb * kOzM, 0, fullA(sa));
          if (++sa == C::A_STAGES) { sa = 0; pa ^= 1u; }
          oz_wait(emptyB(sb), pb ^ 1u);
          mbar_arrive_expect_tx(fullB(sb), C::B_BYTES);
          tma_load_3d(sB + sb * C::B_BYTES, &tmW, ks * kOzKB, c0, 0, fullB(sb));
          if (++sb == kOzBStages) { sb = 0; pb ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0, pt = 0;
      for (;;) {
        oz_wait(qfull(qslot), qphase);
        int item;
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(item) : "r"(q_items + 4u * qslot) : "memory");
        mbar_arrive(qempty(qslot));
        if (++qslot == kOzQueue) { qslot = 0; qphase ^= 1u; }
        if (item < 0) break;
        int jt, rb;
        oz_decode_item(g, item, jt, rb);
        const int c0 = g.N - NT * (g.T - jt);
        const int nk = (c0 + NT + kOzKB - 1) / kOzKB;
        oz_wait(tmem_empty, pt ^ 1u);   // the epilogue has drained the previous tile
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int ks = 0; ks < nk; ks++) {
          oz_wait(fullA(sa), pa);
          oz_wait(fullB(sb), pb);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t aS = sA + sa * C::A_BYTES, bS = sB + sb * C::B_BYTES;
#pragma unroll
          for (int kk = 0; kk < kOzKB; kk += 32) {
            // W[n][k] = 0 for k > n: columns c0 + n < k0 are complete, shrink the N extent (multiples of 16 columns)
            const int k0 = ks * kOzKB + kk;
            int n0 = 0;
            if (g.diag_trim && k0 > c0) n0 = min((k0 - c0) & ~15, NT - 16);
            const uint32_t idesc = umma_idesc_i8(kOzM, NT - n0);
            const uint32_t first = (ks == 0 && kk == 0) ? 0u : 1u;
#pragma unroll
            for (int i = 0; i < S; i++)
#pragma unroll
              for (int j = 0; j + i < S; j++)
                umma_i8(tmem + (uint32_t)((i + j) * NT + n0), umma_desc_sw64(aS + i * (kOzM * kOzKB) + kk),
                        umma_desc_sw64(bS + j * (NT * kOzKB) + n0 * kOzKB + kk), idesc, i == 0 ? first : 1u);
          }
          umma_commit(emptyA(sa));
          umma_commit(emptyB(sb));
          if (++sa == C::A_STAGES) { sa = 0; pa ^= 1u; }
          if (++sb == kOzBStages) { sb = 0; pb ^= 1u; }
        }
        umma_commit(tmem_full);
        pt ^= 1u;
      }
    }
  } else {
    // ===================== epilogue: TMEM -> FP64 recombination -> row sum of squares =====================
    const int lg = warp & 3;                 // TMEM lane group this warp may read
    const int row_in_tile = lg * 32 + lane;
    const int etid = tid - 64;               // 0..127
    uint32_t pt = 0;
    for (;;) {
      oz_wait(qfull(qslot), qphase);
      int item;
      asm volatile("ld.shared.s32 %0, [%1];" : "=r"(item) : "r"(q_items + 4u * qslot) : "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(qempty(qslot));
      if (++qslot == kOzQueue) { qslot = 0; qphase ^= 1u; }
      if (item < 0) break;
      int jt, rb;
      oz_decode_item(g, item, jt, rb);
      const int c0 = g.N - NT * (g.T - jt);
      // column scales of this tile (the previous tile's readers are past them: they arrived on tmem_empty after reading)
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (etid < NT) { const int col = c0 + etid; s_cs[etid] = col >= 0 ? g.colscale[col] : 0.0; }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      oz_wait(tmem_full, pt);
      pt ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      double acc = 0.0;
      const uint32_t trow = tmem + ((uint32_t)(lg * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < NT; c += 16) {
        int32_t v[S][16];
#pragma unroll
        for (int l = 0; l < S; l++) tmem_ld16(trow + (uint32_t)(l * NT + c), v[l]);
#pragma unroll
        for (int q = 0; q < 16; q++) {
          double h = (double)v[S - 1][q];
#pragma unroll
          for (int l = S - 2; l >= 0; l--) h = fma(h, 0.00390625, (double)v[l][q]);
          const double y = h * s_cs[c + q];
          acc = fma(y, y, acc);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
      const int64_t row = (int64_t)rb * kOzM + row_in_tile;
      if (row < g.B) {
        const double rs = g.rowscale[row] * 0.000244140625;   // 2^eR_b * 2^-12 (fixed-point position of the digit products)
        g.part[(int64_t)jt * g.B + row] = acc * rs * rs;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace cosmolike
