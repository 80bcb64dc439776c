// mbarrier micro-benchmark for sm_100a: what one poll / arrive costs a thread, alone and while other warps of the CTA poll.
//   (1) test_wait on a completed phase, back to back (latency of a successful poll)
//   (2) the same while W other warps spin on an incomplete barrier with test_wait / try_wait / try_wait + nanosleep
//   (3) mbarrier.arrive latency under the same load
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench_mbar ubench_mbar.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ uint32_t try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}

// mode: 0 no pollers, 1 test_wait spin, 2 try_wait spin, 3 try_wait + nanosleep(100), 4 all 32 lanes try_wait spin
template <int MODE>
__global__ void k_mbar(int n_pollers, long long* out) {
  __shared__ __align__(8) unsigned long long bars[4];
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < 4; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[i])), "r"(1));
    stop = 0;
  }
  __syncthreads();
  const uint32_t done_bar = smem_u32(&bars[0]), pending_bar = smem_u32(&bars[1]), arr_bar = smem_u32(&bars[2]);
  if (tid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(done_bar) : "memory");   // phase 0 of bars[0] complete
  __syncthreads();
  if (warp == 0) {
    if (lane == 0) {
      for (int rep = 0; rep < 2; rep++) {
        long long t0 = clock64();
        uint32_t acc = 0;
        for (int i = 0; i < 64; i++) acc += test_wait(done_bar, 0);
        long long t1 = clock64();
        for (int i = 0; i < 64; i++) acc += try_wait(done_bar, 0);
        long long t2 = clock64();
        for (int i = 0; i < 64; i++) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(arr_bar) : "memory");
        long long t3 = clock64();
        if (blockIdx.x == 0) { out[0] = (t1 - t0) / 64; out[1] = (t2 - t1) / 64; out[2] = (t3 - t2) / 64; out[3] = acc; }
      }
      stop = 1;
    }
  } else if (warp <= n_pollers) {
    if (MODE == 4 || lane == 0) {
      while (!stop) {
        if (MODE == 1) test_wait(pending_bar, 0);
        if (MODE == 2 || MODE == 4) try_wait(pending_bar, 0);
        if (MODE == 3) { try_wait(pending_bar, 0); __nanosleep(100); }
      }
    }
  }
}

template <int MODE> static void run(int n_pollers, const char* what) {
  long long* d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
  k_mbar<MODE><<<148, 352>>>(n_pollers, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
  printf("%-44s pollers %2d: test_wait %4lld  try_wait %4lld  arrive %4lld cycles%s\n", what, n_pollers, h[0], h[1], h[2], e ? "  ERROR" : "");
  cudaFree(d);
}

int main() {
  run<0>(0, "idle");
  for (int w : {1, 4, 10}) {
    run<1>(w, "others: lane 0 test_wait spin");
    run<2>(w, "others: lane 0 try_wait spin");
    run<3>(w, "others: lane 0 try_wait + nanosleep(100)");
    run<4>(w, "others: 32 lanes try_wait spin");
  }
  return 0;
}
