// latency micro-benchmarks (single warp, dependent chains) for the FP64 path on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__global__ void k(double* out, long long* cyc, double seed) {
  __shared__ double sm[1024];
  int lane = threadIdx.x;
  for (int i = lane; i < 1024; i += blockDim.x) sm[i] = seed + i;
  __syncthreads();
  double x = seed + lane * 1e-6, y = 1.0000001, acc = 0.0;
  long long t0, t1;
  // DFMA chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = fma(x, y, 1e-9);
  t1 = clock64(); if (lane == 0 && blockIdx.x == 0 && threadIdx.x < 32) cyc[0] = t1 - t0; acc += x;
  // DADD chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = x + 1e-9;
  t1 = clock64(); if (lane == 0) cyc[1] = t1 - t0; acc += x;
  // DMUL chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = x * y;
  t1 = clock64(); if (lane == 0) cyc[2] = t1 - t0; acc += x;
  // shfl (64-bit) chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = __shfl_xor_sync(0xffffffffu, x, 1);
  t1 = clock64(); if (lane == 0) cyc[3] = t1 - t0; acc += x;
  // LDS.64 dependent (pointer chase through index)
  int idx = lane;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) { double v = sm[idx & 1023]; idx = (int)__double_as_longlong(v) & 1023; }
  t1 = clock64(); if (lane == 0) cyc[4] = t1 - t0; acc += idx;
  // MUFU.RSQ64H + refinement chain (rsqrt)
  x = 1.5 + lane * 1e-3;
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; i++) x = rsqrt(x) + 1.0;
  t1 = clock64(); if (lane == 0) cyc[5] = t1 - t0; acc += x;
  // F2I + I2F chain
  x = 3.7 + lane;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = (double)((int)x) + 0.5;
  t1 = clock64(); if (lane == 0) cyc[6] = t1 - t0; acc += x;
  // independent DFMA throughput, 8 chains, one warp
  double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; i++) { a0 = fma(a0, y, 1e-9); a1 = fma(a1, y, 1e-9); a2 = fma(a2, y, 1e-9); a3 = fma(a3, y, 1e-9); a4 = fma(a4, y, 1e-9); a5 = fma(a5, y, 1e-9); a6 = fma(a6, y, 1e-9); a7 = fma(a7, y, 1e-9); }
  t1 = clock64(); if (lane == 0) cyc[7] = t1 - t0; acc += a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  // __syncthreads cost with the whole block
  __syncthreads();
  t0 = clock64();
  for (int i = 0; i < 64; i++) __syncthreads();
  t1 = clock64(); if (threadIdx.x == 0) cyc[8] = t1 - t0;
  if (acc == 123.456) out[0] = acc;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8); cudaMallocManaged(&cyc, 16 * 8);
  for (int blk : {32, 256}) {
    k<<<1, blk>>>(out, cyc, 1.0); cudaDeviceSynchronize();
    k<<<1, blk>>>(out, cyc, 1.0); cudaDeviceSynchronize();
    printf("block %d threads: per-op cycles: DFMA %.1f DADD %.1f DMUL %.1f SHFL64 %.1f LDS.64chase %.1f rsqrt+add %.1f F2I+I2F+add %.1f  8xindepDFMA(per 8) %.1f  syncthreads %.1f\n", blk,
           cyc[0] / (double)N, cyc[1] / (double)N, cyc[2] / (double)N, cyc[3] / (double)N, cyc[4] / (double)N, cyc[5] / (double)N, cyc[6] / (double)N, cyc[7] / (double)N, cyc[8] / 64.0);
  }
  return 0;
}
