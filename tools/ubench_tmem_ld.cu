// tcgen05.ld micro-benchmark for sm_100a: how fast can a CTA read its tensor memory back into registers?
// The epilogue of k_chi2_ozaki (csrc/chi2_ozaki.cuh) has to read S levels x NT columns x 128 lanes x 4 B = 229 KB per tile
// at S = 7; this measures bytes per clock for W reading warps (each warp may only touch its own quarter of the lanes, so
// W = 4, 8, 16 means 1, 2, 4 warps per lane quarter) and loads of 16 / 32 / 64 columns per instruction, all in flight
// before one tcgen05.wait::ld (the pattern of the epilogue), with nothing else running on the SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench_tmem_ld ubench_tmem_ld.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t* v);
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t a, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                 "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(a) : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t a, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                 "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                 "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                 "=r"(v[30]), "=r"(v[31]) : "r"(a) : "memory");
}

// each warp reads `cols` columns of its lane quarter, X columns per instruction, `reps` times; returns cycles of warp 0
template <int X>
__global__ void __launch_bounds__(512, 1) k_ld(int n_warps, int cols, int reps, long long* out, uint32_t* sink) {
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < n_warps) {
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16);   // every warp of a lane quarter reads the same 448 columns
    __syncwarp();
    t0 = clock64();
    for (int r = 0; r < reps; r++) {
      for (int c0 = 0; c0 < cols; c0 += 2 * X) {   // two loads in flight per wait, as the epilogue does with its half levels
        uint32_t v0[X], v1[X];
        tmem_ld<X>(base + (uint32_t)c0, v0);
        tmem_ld<X>(base + (uint32_t)(c0 + X), v1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < X; i++) acc += v0[i] ^ v1[i];
      }
    }
    t1 = clock64();
  }
  if (acc == 0x12345678u) sink[threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* d_out; uint32_t* d_sink;
  cudaMalloc(&d_out, 64); cudaMalloc(&d_sink, 4096);
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  printf("# %s, %d SMs; bytes per clock and SM read with tcgen05.ld.32x32b (two loads per tcgen05.wait::ld)\n", prop.name, prop.multiProcessorCount);
  printf("# tile of k_chi2_ozaki<7>: 7 levels x 64 columns x 128 lanes x 4 B = 229376 B per drain\n");
  const int reps = 64, cols = 448;   // 448 = 14 x 32 = 7 x 64
  for (int nw : {4, 8, 16}) {
    for (int x : {16, 32}) {
      long long best = 1LL << 60;
      for (int it = 0; it < 3; it++) {
        if (x == 16) k_ld<16><<<prop.multiProcessorCount, 512>>>(nw, cols, reps, d_out, d_sink);
        else k_ld<32><<<prop.multiProcessorCount, 512>>>(nw, cols, reps, d_out, d_sink);
        long long h = 0;
        cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
        if (cudaGetLastError() != cudaSuccess) { printf("launch failed\n"); return 1; }
        if (h < best) best = h;
      }
      const double bytes = (double)nw * 32 * cols * 4.0 * reps;   // all reading warps together
      printf("warps %2d  x%-2d  %8lld cycles  %7.1f B/clk/SM   -> one 229376-byte drain at this rate: %6.0f cycles\n", nw, x, best, bytes / best,
             229376.0 / (bytes / best));
    }
  }
  return 0;
}
