// chi2_gemm.cuh — stage 3: chi2_sn[b] = | W r_b |^2 for a whole batch, W = L^-1 lower triangular.
//
// Replaces the reference's per-theta forward substitution (solve_triangular.py:5-14: y = L^-1 delta, return
// y.y) by one batched contraction Y[B,N] = R[B,N] . W^T restricted to the lower triangle, on the FP64 tensor
// pipe (mma.sync f64 -> DMMA.8x8x4 on sm_100a), with a fused row-dot epilogue: Y is never written, only the
// per-row sum of squares of each 128-column tile.
//
// Data movement: TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) brings [128 rows x 16 k] boxes of R and of W
// into a 6-stage shared-memory ring guarded by full/empty mbarriers; one producer warp, eight consumer warps
// (2 along M x 4 along N, warp tile 64x32, accumulators in registers).  Fragments are read with conflict-free
// 128-bit shared loads: the k index inside a k16 step is permuted (lane t owns k = 4t..4t+3) identically for
// both operands, which leaves the product unchanged.
//
// Work decomposition: items (row block, column tile) ordered by decreasing k-extent of the column tile and
// dealt round-robin to a persistent grid.  Column tiles are aligned to the END of the matrix (tile t covers
// columns [N - 128 (T - t), +128), the first tile may start at a negative column that TMA zero-fills), so the
// ragged tile is the cheap one; the k loop of a tile stops at its diagonal block, and inside the diagonal block
// every 8-column MMA tile stops at its own last column.  Each warp owns the interleaved 8-column tiles
// {w, w+4, w+8, w+12} of the 128-column tile, so this skipping stays balanced across warps and sub-partitions.  Executed MACs per row ~ N^2/2 (1 + ~0.05) instead of N^2.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cosmolike {

constexpr int kBM = 128, kBN = 128, kBK = 16;
constexpr int kStages = 6;
constexpr int kConsumerWarps = 8;
constexpr int kProducerWarps = 4;  // one full warpgroup so that setmaxnreg can hand its registers to the consumers
constexpr int kGemmThreads = (kConsumerWarps + kProducerWarps) * 32;
constexpr int kBoxBytes = kBM * kBK * 8;       // 16 KB, A and B boxes have the same shape
constexpr int kStageBytes = 2 * kBoxBytes;     // 32 KB
constexpr int kGemmSmemBytes = 1024 /*align slack*/ + kStages * kStageBytes + 2 * 4 * kBM * 8 /*epilogue*/ + 256 /*barriers + item queue*/;

struct GemmArgs {
  int64_t B;       // rows of R in this pass
  int N;           // SN count
  int T;           // column tiles = ceil(N / 128)
  int n_rb;        // row blocks = ceil(B / 128)
  double* part;    // [T][B] partial sums of squares
  double* part_u;  // nullable [T][B]: partial sums of y_j * u_j (moments mode)
  const double* u; // [N] u = W 1 (moments mode)
  int diag_skip;   // 1: warps stop at their own last column inside the diagonal block
  int* counter;    // dynamic scheduling: global item counter (zeroed before the launch); nullptr = static snake order
  int group_rb;    // dynamic scheduling: row blocks per L2 group
  const int* guard; // nullable; accuracy-guard fallback pass of the tcgen05 engine (dynamic scheduling only): guard[1] = rows
                    // flagged (0: nothing to do), guard[2 + rb] != 0 marks the row blocks to compute; the others are skipped
};

constexpr int kQueueDepth = 4;  // items the producer may run ahead of the consumers

// Item order.  Static: tiles by decreasing k extent, row blocks inside (dealt boustrophedon to the CTAs).
// Dynamic: row blocks are taken in groups whose residual rows (group_rb x 128 x N x 8 B) fit in L2 together with W;
// inside a group the tiles go by decreasing k extent, so the R rows of a group are read from HBM once and re-read
// from L2 for the other T-1 column tiles.  CTAs pull items from an atomic counter (the order of execution does
// not affect the result: every item writes its own partial sums).
__device__ __forceinline__ void decode_item(const GemmArgs& g, int64_t item, int& jt, int& rb) {
  if (g.counter == nullptr) {
    jt = g.T - 1 - (int)(item / g.n_rb);
    rb = (int)(item % g.n_rb);
    return;
  }
  const int per_group = g.group_rb * g.T;
  const int n_full = g.n_rb / g.group_rb;
  int grp = (int)(item / per_group);
  int r, width;
  if (grp < n_full) { r = (int)(item % per_group); width = g.group_rb; }
  else { grp = n_full; r = (int)(item - (int64_t)n_full * per_group); width = g.n_rb - n_full * g.group_rb; }
  jt = g.T - 1 - r / width;
  rb = grp * g.group_rb + r % width;
}

// item visited by CTA `cta` in round `k` of the persistent loop: boustrophedon over the cost-sorted item list, so
// that every CTA gets the same mix of expensive and cheap items (static, deterministic, no atomics)
__device__ __forceinline__ int64_t snake_item(int64_t k, int cta, int ncta) {
  return k * ncta + ((k & 1) ? (ncta - 1 - cta) : cta);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mma_f64_16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
      "{%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
        "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

// one k16 step of a warp tile: 64 rows x the 8-column tiles NT0..3 (tile nt lives at smem rows +32 nt)
template <int NT0>
__device__ __forceinline__ void kstep(double (&acc)[4][4][4], uint32_t aS, uint32_t bS, uint32_t ch0, uint32_t ch1) {
  double bf[4][4];
#pragma unroll
  for (int nt = NT0; nt < 4; nt++) {
    double2 lo = lds128(bS + nt * 4096 + ch0), hi = lds128(bS + nt * 4096 + ch1);
    bf[nt][0] = lo.x; bf[nt][1] = lo.y; bf[nt][2] = hi.x; bf[nt][3] = hi.y;
  }
#pragma unroll
  for (int mt = 0; mt < 4; mt++) {
    const uint32_t ar = aS + mt * 2048;
    double2 r0lo = lds128(ar + ch0), r0hi = lds128(ar + ch1);
    double2 r1lo = lds128(ar + 1024 + ch0), r1hi = lds128(ar + 1024 + ch1);
    // a_i: row = g + 8*(i&1), logical k = t + 4*(i>>1) -> physical k = 4t + (i>>1)
    double af[8] = {r0lo.x, r1lo.x, r0lo.y, r1lo.y, r0hi.x, r1hi.x, r0hi.y, r1hi.y};
#pragma unroll
    for (int nt = NT0; nt < 4; nt++) mma_f64_16816(acc[mt][nt], af, bf[nt]);
  }
}

template <bool MOMENTS>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_chi2_gemm(const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmW, const GemmArgs g) {
  extern __shared__ unsigned char gsm_raw[];
  const uint32_t base = (smem_u32(gsm_raw) + 1023u) & ~1023u;  // 128B swizzle needs 1024-byte aligned tiles
  unsigned char* base_ptr = gsm_raw + (base - smem_u32(gsm_raw));
  const uint32_t sA = base;                                   // [stage][128 rows][128 B]
  const uint32_t sB = base + kStages * kBoxBytes;
  double* s_epi = reinterpret_cast<double*>(base_ptr + kStages * kStageBytes);  // [2][4][128]
  const uint32_t bars = base + kStages * kStageBytes + 2 * 4 * kBM * 8;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  auto qfull_bar = [&](int s) { return bars + 8u * (2 * kStages + s); };
  auto qempty_bar = [&](int s) { return bars + 8u * (2 * kStages + kQueueDepth + s); };
  const uint32_t q_items = bars + 8u * (2 * kStages + 2 * kQueueDepth);  // int[kQueueDepth]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kStages; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), kConsumerWarps); }
    for (int s = 0; s < kQueueDepth; s++) { mbar_init(qfull_bar(s), 1); mbar_init(qempty_bar(s), kConsumerWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  const int64_t total = (int64_t)g.n_rb * g.T;
  int stage = 0, qslot = 0;
  uint32_t phase = 0, qphase = 0;

  if (warp >= kConsumerWarps) {
    // ===================== TMA producer (one lane of the producer warpgroup) =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (warp == kConsumerWarps && lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
      const bool nothing = g.guard != nullptr && g.guard[1] == 0;
      for (int64_t round = 0;; round++) {
        // next item: atomic counter (dynamic) or boustrophedon over the static order; published through the item queue
        int64_t item;
        if (nothing) item = total;
        else if (g.counter) item = atomicAdd(g.counter, 1);
        else { item = snake_item(round, blockIdx.x, gridDim.x); if (item >= total && round * gridDim.x < total) continue; }
        const bool done = item >= total;
        if (!done && g.guard != nullptr) {   // fallback pass: only the flagged row blocks
          int jt_, rb_;
          decode_item(g, item, jt_, rb_);
          if (g.guard[2 + rb_] == 0) continue;
        }
        mbar_wait(qempty_bar(qslot), qphase ^ 1u);
        asm volatile("st.shared.s32 [%0], %1;" ::"r"(q_items + 4u * qslot), "r"(done ? -1 : (int)item) : "memory");
        mbar_arrive(qfull_bar(qslot));
        if (++qslot == kQueueDepth) { qslot = 0; qphase ^= 1u; }
        if (done) break;
        int jt, rb;
        decode_item(g, item, jt, rb);
        const int c0 = g.N - kBN * (g.T - jt);          // first column of the tile (may be < 0 for jt == 0)
        const int nk = (c0 + kBN + kBK - 1) / kBK;      // k runs to the end of the diagonal block
        for (int ks = 0; ks < nk; ks++) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_arrive_expect_tx(full_bar(stage), kStageBytes);
          tma_load_2d(sA + stage * kBoxBytes, &tmR, ks * kBK, rb * kBM, full_bar(stage));
          tma_load_2d(sB + stage * kBoxBytes, &tmW, ks * kBK, c0, full_bar(stage));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    return;
  }

  // ===================== consumers: DMMA + fused row-dot epilogue =====================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 240;");
  const int gq = lane >> 2, t = lane & 3;   // mma "groupID" and "threadID_in_group"
  // warp -> 64 rows (warp_m) x the four interleaved 8-column tiles {warp_n + 4 nt}
  const int warp_m = warp & 1, warp_n = warp >> 1;
  // byte offsets of this lane's two 16-byte chunks inside a 128-byte row, after the 128B swizzle
  const uint32_t ch0 = (uint32_t)(((2 * t) ^ gq) << 4), ch1 = (uint32_t)(((2 * t + 1) ^ gq) << 4);
  const uint32_t a_row0 = (uint32_t)(warp_m * 64 + gq) * 128u;  // + mt*16*128 (+8 rows = +1024)
  const uint32_t b_row0 = (uint32_t)(warp_n * 8 + gq) * 128u;   // + nt*32*128
  int epi_buf = 0;

  for (;;) {
    mbar_wait(qfull_bar(qslot), qphase);
    int item;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(item) : "r"(q_items + 4u * qslot) : "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(qempty_bar(qslot));
    if (++qslot == kQueueDepth) { qslot = 0; qphase ^= 1u; }
    if (item < 0) break;
    int jt, rb;
    decode_item(g, item, jt, rb);
    const int c0 = g.N - kBN * (g.T - jt);
    const int nk = (c0 + kBN + kBK - 1) / kBK;

    double acc[4][4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++)
#pragma unroll
        for (int q = 0; q < 4; q++) acc[i][j][q] = 0.0;

    for (int ks = 0; ks < nk; ks++) {
      mbar_wait(full_bar(stage), phase);
      // W[j][k] = 0 for k > j: the 8-column tile nt (columns c0 + 8 (warp_n + 4 nt) .. +7) needs k16 step ks only
      // if ks*16 <= its last column.  nt0 = first tile still active (0 outside the diagonal block).
      int nt0 = 0;
      if (g.diag_skip) {
        const int d = ks * kBK - c0 - 8 * warp_n - 8;  // tile nt active  <=>  32 nt > d
        nt0 = d < 0 ? 0 : (d >> 5) + 1;
      }
      if (nt0 < 4) {
        const uint32_t aS = sA + stage * kBoxBytes + a_row0;
        const uint32_t bS = sB + stage * kBoxBytes + b_row0;
        // separate unpredicated code paths per number of live tiles (predicated mma.sync serialises on temporaries)
        switch (nt0) {
          case 0: kstep<0>(acc, aS, bS, ch0, ch1); break;
          case 1: kstep<1>(acc, aS, bS, ch0, ch1); break;
          case 2: kstep<2>(acc, aS, bS, ch0, ch1); break;
          default: kstep<3>(acc, aS, bS, ch0, ch1); break;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar(stage));
      if (++stage == kStages) { stage = 0; phase ^= 1u; }
    }

    // ---- epilogue: per-row sum over this warp's 32 columns of y^2 (and y*u) ----
    double* epi = s_epi + epi_buf * (4 * kBM);
    double* epi_u = nullptr;
    (void)epi_u;
    double uu[4][2];
    if (MOMENTS) {
#pragma unroll
      for (int nt = 0; nt < 4; nt++) {
        int col = c0 + (warp_n + 4 * nt) * 8 + 2 * t;
        uu[nt][0] = (col >= 0 && col < g.N) ? g.u[col] : 0.0;
        uu[nt][1] = (col + 1 >= 0 && col + 1 < g.N) ? g.u[col + 1] : 0.0;
      }
    }
#pragma unroll
    for (int mt = 0; mt < 4; mt++) {
      double s0 = 0.0, s1 = 0.0, u0 = 0.0, u1 = 0.0;
#pragma unroll
      for (int nt = 0; nt < 4; nt++) {
        s0 += acc[mt][nt][0] * acc[mt][nt][0] + acc[mt][nt][1] * acc[mt][nt][1];
        s1 += acc[mt][nt][2] * acc[mt][nt][2] + acc[mt][nt][3] * acc[mt][nt][3];
        if (MOMENTS) {
          u0 += acc[mt][nt][0] * uu[nt][0] + acc[mt][nt][1] * uu[nt][1];
          u1 += acc[mt][nt][2] * uu[nt][0] + acc[mt][nt][3] * uu[nt][1];
        }
      }
      s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
      if (MOMENTS) {
        u0 += __shfl_xor_sync(0xffffffffu, u0, 1); u0 += __shfl_xor_sync(0xffffffffu, u0, 2);
        u1 += __shfl_xor_sync(0xffffffffu, u1, 1); u1 += __shfl_xor_sync(0xffffffffu, u1, 2);
      }
      const int r = warp_m * 64 + mt * 16 + gq;
      if (t == 0) { epi[warp_n * kBM + r] = s0; epi[warp_n * kBM + r + 8] = s1; }
      if (MOMENTS && t == 1) {
        // the y.u partials share the epilogue buffer of the other parity: written after the barrier below
        acc[mt][0][0] = u0; acc[mt][0][1] = u1;
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory");
    if (tid < kBM) {
      const int64_t row = (int64_t)rb * kBM + tid;
      if (row < g.B) g.part[(int64_t)jt * g.B + row] = (epi[tid] + epi[kBM + tid]) + (epi[2 * kBM + tid] + epi[3 * kBM + tid]);
    }
    if (MOMENTS) {
      // second pass through the same buffer for y.u
      asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory");
      if (t == 1) {
#pragma unroll
        for (int mt = 0; mt < 4; mt++) {
          const int r = warp_m * 64 + mt * 16 + gq;
          epi[warp_n * kBM + r] = acc[mt][0][0]; epi[warp_n * kBM + r + 8] = acc[mt][0][1];
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory");
      if (tid < kBM) {
        const int64_t row = (int64_t)rb * kBM + tid;
        if (row < g.B) g.part_u[(int64_t)jt * g.B + row] = (epi[tid] + epi[kBM + tid]) + (epi[2 * kBM + tid] + epi[3 * kBM + tid]);
      }
    }
    epi_buf ^= 1;
  }
}

}  // namespace cosmolike
