// tcgen05 kind::i8 micro-benchmark / bring-up for sm_100a (B200).
//  (1) correctness of the UMMA shared-memory descriptors used by the sliced chi-squared kernel: TMA 3-D boxes
//      {KB bytes of k, rows, slices} with 32/64/128-byte swizzle, K-major A (128 rows) and B (NT rows), int32
//      accumulators in TMEM, several "levels" (column ranges) per tile, read back with tcgen05.ld.
//  (2) issue rate of tcgen05.mma M=128 x N x K=32 (int8) for N = 64, 80, 128, 256 with operands resident in smem.
//  (3) chip-wide TMA load bandwidth from L2 for the same boxes (inner extent 32 / 64 / 128 bytes).
//  (4) the contraction's stacked-plane MMA schedule with resident operands: cycles per 64-byte k block against the pipe floor;
//      variants: 1 non-overlapping accumulators, 2 one A plane, 3 / 4 one / two tcgen05.commit per block, 5 random operand data,
//      6 a tcgen05.fence per block (`ubench_umma_i8 sched 6`).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench_umma_i8 ubench_umma_i8.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode_fn() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
    printf("no cuTensorMapEncodeTiled\n"); exit(1);
  }
  return (PFN_encodeTiled)p;
}
// int8 tensor [slices][rows][ld bytes], logical k extent `cols`; box {kb, box_rows, box_slices}
static CUtensorMap make_map(const int8_t* ptr, int64_t cols, int64_t rows, int64_t slices, int64_t ld, int kb, int box_rows, int box_slices) {
  CUtensorMap tm;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)slices};
  cuuint64_t strides[2] = {(cuuint64_t)ld, (cuuint64_t)ld * rows};
  cuuint32_t box[3] = {(cuuint32_t)kb, (cuuint32_t)box_rows, (cuuint32_t)box_slices};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapSwizzle sw = kb == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : kb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = get_encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return tm;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
// bounded wait: returns false on timeout so that a wrong descriptor cannot hang the box
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  for (long long i = 0; i < (1LL << 26); i++) if (mbar_try_wait(bar, parity)) return true;
  return false;
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// K-major operand tile, rows x KB bytes, swizzle = KB bytes; 8-row groups are SBO = 8*KB bytes apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, int kb) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                                  // LBO (unused for swizzled K-major)
  d |= (uint64_t)(((8 * kb) >> 4) & 0x3FFF) << 32;         // SBO
  d |= (uint64_t)1 << 46;                                  // descriptor version (Blackwell)
  d |= (uint64_t)(kb == 32 ? 6 : kb == 64 ? 4 : 2) << 61;  // layout type
  return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_i8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- (1) correctness
// D_level[128][NT] = sum over pairs (i, j) with i + j == level of A_i[128][K] . B_j[NT][K]^T, S slices each.
template <int KB, int NT, int S>
__global__ void __launch_bounds__(128, 1) k_check(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int K, int32_t* out, int* status) {
  extern __shared__ unsigned char raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  constexpr uint32_t A_BYTES = S * 128 * KB, B_BYTES = S * NT * KB;
  const uint32_t sA = base, sB = base + A_BYTES, bars = sB + ((B_BYTES + 1023u) & ~1023u), tslot = bars + 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(bars, 1); mbar_init(bars + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot) : "memory");
  bool ok = true;
  if (tid == 0) {
    constexpr uint32_t idesc = umma_idesc_i8(128, NT);
    uint32_t ph = 0;
    for (int k0 = 0; k0 < K && ok; k0 += KB) {
      mbar_arrive_expect_tx(bars, A_BYTES + B_BYTES);
      tma_load_3d(sA, &tmA, k0, 0, 0, bars);
      tma_load_3d(sB, &tmB, k0, 0, 0, bars);
      ok = mbar_wait(bars, ph);
      if (!ok) { *status = 1; break; }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int kk = 0; kk < KB; kk += 32)
        for (int i = 0; i < S; i++)
          for (int j = 0; j + i < S; j++)
            umma_i8(tmem + (uint32_t)((i + j) * NT), umma_desc(sA + i * 128 * KB + kk, KB), umma_desc(sB + j * NT * KB + kk, KB), idesc,
                    (k0 > 0 || kk > 0 || i > 0) ? 1u : 0u);
      umma_commit(bars + 8);
      ok = mbar_wait(bars + 8, ph);
      if (!ok) { *status = 2; break; }
      ph ^= 1u;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (*status == 0) {
    for (int lvl = 0; lvl < S; lvl++)
      for (int c = 0; c < NT; c += 8) {
        uint32_t v[8];
        tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(lvl * NT + c), v);
        for (int q = 0; q < 8; q++) out[((size_t)lvl * 128 + warp * 32 + lane) * NT + c + q] = (int32_t)v[q];
      }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int KB, int NT, int S>
static bool run_check(int K) {
  const int ld = ((K + 15) / 16) * 16;
  std::vector<int8_t> A((size_t)S * 128 * ld), B((size_t)S * NT * ld);
  srand(1234 + KB + NT);
  for (auto& x : A) x = (int8_t)(rand() % 256 - 128);
  for (auto& x : B) x = (int8_t)(rand() % 256 - 128);
  int8_t *dA, *dB; int32_t* dO; int* dS;
  CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dO, (size_t)S * 128 * NT * 4)); CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dO, 0xff, (size_t)S * 128 * NT * 4)); CK(cudaMemset(dS, 0, 4));
  // logical k extent K - 5: the ragged tail must come back as zeros (TMA out-of-bounds fill)
  const int Kl = K - 5;
  CUtensorMap tA = make_map(dA, Kl, 128, S, ld, KB, 128, S), tB = make_map(dB, Kl, NT, S, ld, KB, NT, S);
  size_t smem = 1024 + S * 128 * KB + ((S * NT * KB + 1023) & ~1023) + 256;
  CK(cudaFuncSetAttribute(k_check<KB, NT, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_check<KB, NT, S><<<1, 128, smem>>>(tA, tB, K, dO, dS);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("check KB=%d NT=%d S=%d: kernel error %s\n", KB, NT, S, cudaGetErrorString(e)); exit(2); }
  int st; CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  std::vector<int32_t> O((size_t)S * 128 * NT);
  CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
  long long bad = 0;
  for (int lvl = 0; lvl < S; lvl++)
    for (int m = 0; m < 128; m++)
      for (int n = 0; n < NT; n++) {
        long long acc = 0;
        for (int i = 0; i <= lvl; i++) {
          int j = lvl - i;
          const int8_t* a = &A[((size_t)i * 128 + m) * ld];
          const int8_t* b = &B[((size_t)j * NT + n) * ld];
          for (int k = 0; k < Kl; k++) acc += (int)a[k] * (int)b[k];
        }
        if ((int32_t)acc != O[((size_t)lvl * 128 + m) * NT + n]) { if (bad < 4) printf("  mismatch lvl %d m %d n %d: want %lld got %d\n", lvl, m, n, acc, O[((size_t)lvl * 128 + m) * NT + n]); bad++; }
      }
  printf("check KB=%3d NT=%3d S=%d K=%d: status %d, mismatches %lld / %zu -> %s\n", KB, NT, S, K, st, bad, O.size(), (st == 0 && bad == 0) ? "OK" : "FAIL");
  cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dS);
  return st == 0 && bad == 0;
}

// ---------------------------------------------------------------- (2) MMA issue rate, operands resident in smem
template <int KB, int NT, int S>
__global__ void __launch_bounds__(128, 1) k_rate(int iters, long long* cycles) {
  extern __shared__ unsigned char raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  constexpr uint32_t A_BYTES = S * 128 * KB, B_BYTES = S * NT * KB;
  const uint32_t sA = base, sB = base + A_BYTES, bars = sB + ((B_BYTES + 1023u) & ~1023u), tslot = bars + 64;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (uint32_t o = tid * 4; o < A_BYTES + B_BYTES; o += 128 * 4) asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + o), "r"(0x01010101u * (o & 3)));
  if (tid == 0) { mbar_init(bars, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot) : "memory");
  constexpr int LV = (512 / NT) < S ? (512 / NT) : S;  // levels that fit in TMEM
  if (tid == 0) {
    constexpr uint32_t idesc = umma_idesc_i8(128, NT);
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int it = 0; it < iters; it++) {
      for (int kk = 0; kk < KB; kk += 32)
        for (int i = 0; i < S; i++)
          for (int j = 0; j + i < S; j++)
            umma_i8(tmem + (uint32_t)(((i + j) % LV) * NT), umma_desc(sA + i * 128 * KB + kk, KB), umma_desc(sB + j * NT * KB + kk, KB), idesc, 1u);
      if ((it & 7) == 7 || it == iters - 1) {  // keep the issue queue bounded
        umma_commit(bars);
        if (!mbar_wait(bars, ph)) { cycles[blockIdx.x] = -1; break; }
        ph ^= 1u;
      }
    }
    long long t1 = clock64();
    if (cycles[blockIdx.x] != -1) cycles[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int KB, int NT, int S>
static void run_rate(int nsm) {
  long long* dC; CK(cudaMalloc(&dC, nsm * 8)); CK(cudaMemset(dC, 0, nsm * 8));
  size_t smem = 1024 + S * 128 * KB + ((S * NT * KB + 1023) & ~1023) + 256;
  CK(cudaFuncSetAttribute(k_rate<KB, NT, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int iters = 400;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_rate<KB, NT, S><<<nsm, 128, smem>>>(10, dC);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k_rate<KB, NT, S><<<nsm, 128, smem>>>(iters, dC);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> c(nsm); CK(cudaMemcpy(c.data(), dC, nsm * 8, cudaMemcpyDeviceToHost));
  const double n_mma = (double)iters * (KB / 32) * (S * (S + 1) / 2);
  const double macs = n_mma * 128.0 * NT * 32 * nsm;
  printf("rate KB=%3d NT=%3d S=%d: %8.1f cycles/MMA (floor %d), %.3f ms, %.1f TMAC/s int8 chip-wide (%.0f TOP/s)%s\n", KB, NT, S, (double)c[0] / n_mma, NT / 2, ms,
         macs / (ms * 1e-3) * 1e-12, 2 * macs / (ms * 1e-3) * 1e-12, c[0] < 0 ? "  TIMEOUT" : "");
  cudaFree(dC);
}

// ---------------------------------------------------------------- (3) TMA bandwidth from L2
template <int KB, int ROWS, int S, int STAGES>
__global__ void __launch_bounds__(64, 1) k_tma(const __grid_constant__ CUtensorMap tm, int n_kblocks, int n_rowblocks, int rounds, int* status) {
  extern __shared__ unsigned char raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  constexpr uint32_t BYTES = S * ROWS * KB;
  const uint32_t bars = base + STAGES * BYTES;
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 8 * (STAGES + s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int rb = blockIdx.x % n_rowblocks;
  if (tid == 0) {
    int stage = 0; uint32_t ph = 0;
    for (int r = 0; r < rounds; r++)
      for (int kb = 0; kb < n_kblocks; kb++) {
        if (!mbar_wait(bars + 8 * (STAGES + stage), ph ^ 1u)) { *status = 1; return; }
        mbar_arrive_expect_tx(bars + 8 * stage, BYTES);
        tma_load_3d(base + stage * BYTES, &tm, ((kb + blockIdx.x) % n_kblocks) * KB, rb * ROWS, 0, bars + 8 * stage);
        if (++stage == STAGES) { stage = 0; ph ^= 1u; }
      }
  } else if (tid == 32) {
    int stage = 0; uint32_t ph = 0;
    for (int r = 0; r < rounds; r++)
      for (int kb = 0; kb < n_kblocks; kb++) {
        if (!mbar_wait(bars + 8 * stage, ph)) { *status = 2; return; }
        mbar_arrive(bars + 8 * (STAGES + stage));
        if (++stage == STAGES) { stage = 0; ph ^= 1u; }
      }
  }
}

template <int KB, int ROWS, int S, int STAGES>
static void run_tma(int nsm, const int8_t* d, int64_t K, int64_t rows_total, int64_t ld, int* dS) {
  CUtensorMap tm = make_map(d, K, rows_total, S, ld, KB, ROWS, S);
  size_t smem = 1024 + (size_t)STAGES * S * ROWS * KB + 256;
  CK(cudaFuncSetAttribute(k_tma<KB, ROWS, S, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nkb = (int)(K / KB), nrb = (int)(rows_total / ROWS), rounds = 40;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_tma<KB, ROWS, S, STAGES><<<nsm, 64, smem>>>(tm, nkb, nrb, 2, dS);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k_tma<KB, ROWS, S, STAGES><<<nsm, 64, smem>>>(tm, nkb, nrb, rounds, dS);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  int st; CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  double bytes = (double)nsm * rounds * nkb * S * ROWS * KB;
  printf("tma  KB=%3d rows=%3d S=%d stages=%d (%3zu KB smem): %.3f ms, %.2f TB/s chip-wide, %.1f B/clk/SM @1.965GHz, status %d\n", KB, ROWS, S, STAGES, smem >> 10, ms,
         bytes / (ms * 1e-3) * 1e-12, bytes / (ms * 1e-3) / nsm / 1.965e9, st);
}


// ---------------------------------------------------------------- (4) the contraction's MMA schedule, operands resident
// One "block" = two K = 32 halves x rounds i = 0..S-1 x stacks of up to 256 / NT planes of W (chi2_ozaki.cuh).
// VARIANT 0: as in the kernel (round i accumulates into levels i..S-1); 1: every MMA gets its own TMEM columns where they
// fit (no accumulator overlap between consecutive MMAs); 2: all MMAs read A plane 0 (same shared-memory lines).
template <int NT, int S, int VARIANT>
__global__ void __launch_bounds__(128, 1) k_sched(int iters, long long* cycles) {
  extern __shared__ unsigned char raw[];
  constexpr int KB = 64;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  constexpr uint32_t A_BYTES = S * 128 * KB, B_BYTES = S * NT * KB;
  const uint32_t sA = base, sB = base + A_BYTES, bars = sB + ((B_BYTES + 1023u) & ~1023u), tslot = bars + 64;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (uint32_t o = tid * 4; o < A_BYTES + B_BYTES; o += 128 * 4) {
    uint32_t v = 0x01010101u * (o & 3);
    if (VARIANT == 5) { v = (o + 12345u) * 2654435761u; v ^= v >> 15; v *= 2246822519u; v ^= v >> 13; }   // random digits: full data toggling
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + o), "r"(v));
  }
  if (tid == 0) { mbar_init(bars, 1); mbar_init(bars + 8, 1); mbar_init(bars + 16, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot) : "memory");
  constexpr int MAXST = 256 / NT;
  if (tid == 0) {
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int it = 0; it < iters; it++) {
      int slot = 0;
      if constexpr (VARIANT == 7 || VARIANT == 8) {
        // level-group-major order (groups from the top level down): 7 = two groups, 8 = three groups
        constexpr int NG = VARIANT == 7 ? 2 : 3;
#pragma unroll
        for (int gq = 0; gq < NG; gq++) {
          const int la = NG == 2 ? (gq == 0 ? (S + 1) / 2 : 0) : (S == 7 ? (gq == 0 ? 5 : gq == 1 ? 3 : 0) : (gq == 0 ? 4 : gq == 1 ? 2 : 0));
          const int lb = gq == 0 ? S - 1 : (NG == 2 ? (S + 1) / 2 - 1 : (S == 7 ? (gq == 1 ? 4 : 2) : (gq == 1 ? 3 : 1)));
#pragma unroll
          for (int i = 0; i <= lb; i++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
              const int jlo = la - i > 0 ? la - i : 0;
              const int jhi = lb - i < S - 1 - i ? lb - i : S - 1 - i;
              if (jhi >= jlo)
                umma_i8(tmem + (uint32_t)((i + jlo) * NT), umma_desc(sA + i * 128 * KB + 32 * h, KB), umma_desc(sB + jlo * NT * KB + 32 * h, KB),
                        umma_idesc_i8(128, (jhi - jlo + 1) * NT), 1u);
            }
        }
      } else
#pragma unroll
      for (int i = 0; i < S; i++)
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
          for (int j0 = 0; j0 < S - i; j0 += MAXST) {
            const int cnt = (S - i - j0) < MAXST ? (S - i - j0) : MAXST;
            uint32_t dcol = (uint32_t)((i + j0) * NT);
            if (VARIANT == 1) { dcol = (uint32_t)((slot * 256) % 512); if (dcol + cnt * NT > 512) dcol = 0; slot++; }
            const int ai = VARIANT == 2 ? 0 : i;
            umma_i8(tmem + dcol, umma_desc(sA + ai * 128 * KB + 32 * h, KB), umma_desc(sB + j0 * NT * KB + 32 * h, KB), umma_idesc_i8(128, cnt * NT), 1u);
          }
      if (VARIANT == 6) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");   // 6: one fence per block
      if (VARIANT == 3 || VARIANT == 4) umma_commit(bars + 8);    // 3: one, 4: two tcgen05.commit per block (nobody waits on them)
      if (VARIANT == 4) umma_commit(bars + 16);
      if ((it & 3) == 3 || it == iters - 1) {
        umma_commit(bars);
        if (!mbar_wait(bars, ph)) { cycles[blockIdx.x] = -1; break; }
        ph ^= 1u;
      }
    }
    long long t1 = clock64();
    if (cycles[blockIdx.x] != -1) cycles[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int NT, int S, int VARIANT>
static void run_sched(int nsm) {
  long long* dC; CK(cudaMalloc(&dC, nsm * 8)); CK(cudaMemset(dC, 0, nsm * 8));
  size_t smem = 1024 + S * 128 * 64 + ((S * NT * 64 + 1023) & ~1023) + 256;
  CK(cudaFuncSetAttribute(k_sched<NT, S, VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int iters = 400;
  k_sched<NT, S, VARIANT><<<nsm, 128, smem>>>(8, dC);
  CK(cudaDeviceSynchronize());
  k_sched<NT, S, VARIANT><<<nsm, 128, smem>>>(iters, dC);
  CK(cudaDeviceSynchronize());
  std::vector<long long> c(nsm); CK(cudaMemcpy(c.data(), dC, nsm * 8, cudaMemcpyDeviceToHost));
  const int pairs = S * (S + 1) / 2;
  fflush(stdout);
  printf("sched NT=%3d S=%d variant %d: %8.1f cycles per K=64 block (tensor floor %d, %d products)%s\n", NT, S, VARIANT, (double)c[0] / iters, pairs * NT, pairs,
         c[0] < 0 ? "  TIMEOUT" : "");
  cudaFree(dC);
}

int main(int argc, char** argv) {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, nsm);
  bool ok = true;
  ok &= run_check<32, 64, 1>(64);
  ok &= run_check<32, 80, 3>(256);
  ok &= run_check<64, 64, 3>(256);
  ok &= run_check<64, 80, 6>(192);
  ok &= run_check<32, 80, 6>(192);
  ok &= run_check<128, 80, 2>(256);
  ok &= run_check<32, 256, 1>(128);
  if (!ok) { printf("descriptor check FAILED; skipping rates\n"); return 3; }
  if (argc > 2) {
    int v = atoi(argv[2]);
    if (v == 6) run_sched<64, 7, 6>(nsm);
    if (v == 7) { run_sched<64, 7, 0>(nsm); run_sched<64, 7, 7>(nsm); run_sched<64, 7, 8>(nsm); run_sched<80, 6, 0>(nsm); run_sched<80, 6, 7>(nsm); run_sched<80, 6, 8>(nsm); }
    fflush(stdout); return 0;
  }
  run_sched<64, 7, 0>(nsm); run_sched<64, 7, 1>(nsm); run_sched<64, 7, 2>(nsm); run_sched<64, 7, 3>(nsm); run_sched<64, 7, 4>(nsm); run_sched<64, 7, 5>(nsm);
  run_sched<80, 6, 0>(nsm); run_sched<80, 6, 1>(nsm); run_sched<80, 6, 2>(nsm);
  fflush(stdout);
  if (argc > 1) return 0;
  run_rate<32, 64, 6>(nsm); run_rate<32, 80, 6>(nsm); run_rate<64, 64, 6>(nsm); run_rate<64, 80, 6>(nsm);
  run_rate<32, 128, 3>(nsm); run_rate<32, 256, 2>(nsm); run_rate<128, 256, 1>(nsm);
  run_rate<32, 80, 6>(1); run_rate<32, 256, 2>(1);
  // L2-resident slices: 6 x 5632 rows x 1728 B = 58 MB
  const int64_t K = 1728, rows = 5632, ld = 1728;
  int8_t* d; CK(cudaMalloc(&d, (size_t)6 * rows * ld)); CK(cudaMemset(d, 1, (size_t)6 * rows * ld));
  int* dS; CK(cudaMalloc(&dS, 4)); CK(cudaMemset(dS, 0, 4));
  run_tma<32, 128, 6, 6>(nsm, d, K, rows, ld, dS);
  run_tma<64, 128, 6, 3>(nsm, d, K, rows, ld, dS);
  run_tma<128, 128, 6, 2>(nsm, d, K, rows, ld, dS);
  run_tma<32, 128, 1, 36>(nsm, d, K, rows, ld, dS);
  run_tma<64, 128, 1, 18>(nsm, d, K, rows, ld, dS);
  run_tma<128, 128, 1, 9>(nsm, d, K, rows, ld, dS);
  run_tma<128, 128, 6, 2>(nsm / 2, d, K, rows, ld, dS);
  return 0;
}
