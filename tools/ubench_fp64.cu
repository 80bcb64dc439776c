// FP64 pipe micro-benchmark for sm_100a: DMMA (mma.sync f64) shapes vs DFMA, alone and mixed.
// Also verifies the f64 fragment layouts assumed by the chi-squared GEMM kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp64 ubench_fp64.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void mma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double (&c)[4], const double (&a)[2], double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

constexpr int NACC = 8;   // independent accumulator chains per warp

template <int SHAPE>  // 0: m8n8k4, 1: m16n8k4, 2: m16n8k8, 3: m16n8k16
__global__ void __launch_bounds__(512) k_dmma(double* out, int iters, double seed) {
  double a[8], b[4];
  for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x * 1e-9 + i;
  for (int i = 0; i < 4; i++) b[i] = seed * 0.5 + i;
  double c[NACC][4];
  for (int j = 0; j < NACC; j++) for (int i = 0; i < 4; i++) c[j][i] = 0.0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < NACC; j++) {
      if (SHAPE == 0) { double cc[2] = {c[j][0], c[j][1]}; mma884(cc, a[0], b[0]); c[j][0] = cc[0]; c[j][1] = cc[1]; }
      if (SHAPE == 1) { double aa[2] = {a[0], a[1]}; mma1684(c[j], aa, b[0]); }
      if (SHAPE == 2) { double aa[4] = {a[0], a[1], a[2], a[3]}; double bb[2] = {b[0], b[1]}; mma1688(c[j], aa, bb); }
      if (SHAPE == 3) { mma16816(c[j], a, b); }
    }
  }
  double s = 0;
  for (int j = 0; j < NACC; j++) for (int i = 0; i < 4; i++) s += c[j][i];
  if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(1024) k_dfma(double* out, int iters, double seed) {
  double x[16];
  for (int i = 0; i < 16; i++) x[i] = seed + i + threadIdx.x * 1e-9;
  double m = 1.0 + seed * 1e-9, q = seed * 1e-3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = fma(x[i], m, q);
  }
  double s = 0;
  for (int i = 0; i < 16; i++) s += x[i];
  if (s == 12345.678) out[0] = s;
}

// mixed: per iteration NACC m16n8k16 DMMA + NF DFMA
template <int NF>
__global__ void __launch_bounds__(512) k_mixed(double* out, int iters, double seed) {
  double a[8], b[4];
  for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x * 1e-9 + i;
  for (int i = 0; i < 4; i++) b[i] = seed * 0.5 + i;
  double c[NACC][4];
  for (int j = 0; j < NACC; j++) for (int i = 0; i < 4; i++) c[j][i] = 0.0;
  double x[16];
  for (int i = 0; i < 16; i++) x[i] = seed + i + threadIdx.x * 1e-9;
  double m = 1.0 + seed * 1e-9, q = seed * 1e-3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < NACC; j++) {
      { double cc[2] = {c[j][0], c[j][1]}; mma884(cc, a[0], b[0]); c[j][0] = cc[0]; c[j][1] = cc[1]; }
      { double cc[2] = {c[j][2], c[j][3]}; mma884(cc, a[1], b[1]); c[j][2] = cc[0]; c[j][3] = cc[1]; }
#pragma unroll
      for (int i = 0; i < NF / NACC; i++) x[(j * (NF / NACC) + i) % 16] = fma(x[(j * (NF / NACC) + i) % 16], m, q);
    }
  }
  double s = 0;
  for (int j = 0; j < NACC; j++) for (int i = 0; i < 4; i++) s += c[j][i];
  for (int i = 0; i < 16; i++) s += x[i];
  if (s == 12345.678) out[0] = s;
}

// ---- layout verification: C[16x8] = A[16xK] * B[Kx8], A row-major, B given as Bt[n][k]
template <int K>
__global__ void k_verify(const double* A, const double* Bt, double* C) {
  int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  double c[4] = {0, 0, 0, 0};
  if (K == 4) {
    double a[2] = {A[g * K + t], A[(g + 8) * K + t]};
    mma1684(c, a, Bt[g * K + t]);
  } else if (K == 8) {
    double a[4] = {A[g * K + t], A[(g + 8) * K + t], A[g * K + t + 4], A[(g + 8) * K + t + 4]};
    double b[2] = {Bt[g * K + t], Bt[g * K + t + 4]};
    mma1688(c, a, b);
  } else {
    double a[8], b[4];
    for (int i = 0; i < 8; i++) a[i] = A[(g + 8 * (i & 1)) * K + t + 4 * (i >> 1)];
    for (int i = 0; i < 4; i++) b[i] = Bt[g * K + t + 4 * i];
    mma16816(c, a, b);
  }
  C[g * 8 + 2 * t] = c[0]; C[g * 8 + 2 * t + 1] = c[1];
  C[(g + 8) * 8 + 2 * t] = c[2]; C[(g + 8) * 8 + 2 * t + 1] = c[3];
}

template <int K> static void verify() {
  double hA[16 * K], hB[8 * K], hC[128], ref[128];
  for (int i = 0; i < 16 * K; i++) hA[i] = sin(0.37 * i + 0.1);
  for (int i = 0; i < 8 * K; i++) hB[i] = cos(0.91 * i + 0.3);
  for (int m = 0; m < 16; m++) for (int n = 0; n < 8; n++) { double s = 0; for (int k = 0; k < K; k++) s += hA[m * K + k] * hB[n * K + k]; ref[m * 8 + n] = s; }
  double *dA, *dB, *dC;
  CK(cudaMalloc(&dA, sizeof hA)); CK(cudaMalloc(&dB, sizeof hB)); CK(cudaMalloc(&dC, sizeof hC));
  CK(cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice));
  k_verify<K><<<1, 32>>>(dA, dB, dC);
  CK(cudaMemcpy(hC, dC, sizeof hC, cudaMemcpyDeviceToHost));
  double md = 0; for (int i = 0; i < 128; i++) md = fmax(md, fabs(hC[i] - ref[i]));
  printf("verify m16n8k%d layout: max|diff| = %.3e %s\n", K, md, md < 1e-13 ? "OK" : "MISMATCH");
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
}

template <typename F> static float time_ms(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms); }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount; int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("device %s, %d SMs, max clock %d kHz\n", p.name, sms, clk);
  verify<4>(); verify<8>(); verify<16>();
  double* out; CK(cudaMalloc(&out, 8));
  const int iters = 20000;
  for (int warps : {4, 8, 16}) {
    int thr = warps * 32; int grid = sms;
    double nw = (double)grid * warps;
    float t;
    t = time_ms([&] { k_dmma<0><<<grid, thr>>>(out, iters, 1.0); });
    printf("warps/SM %2d  m8n8k4  : %8.2f TFLOP/s (%.1f MAC/clk/SM @max clk)\n", warps, nw * iters * NACC * 256 * 2 / t / 1e9, nw * iters * NACC * 256 / (t * 1e-3) / sms / (clk * 1e3));
    t = time_ms([&] { k_dmma<1><<<grid, thr>>>(out, iters, 1.0); });
    printf("warps/SM %2d  m16n8k4 : %8.2f TFLOP/s\n", warps, nw * iters * NACC * 512 * 2 / t / 1e9);
    t = time_ms([&] { k_dmma<2><<<grid, thr>>>(out, iters, 1.0); });
    printf("warps/SM %2d  m16n8k8 : %8.2f TFLOP/s\n", warps, nw * iters * NACC * 1024 * 2 / t / 1e9);
    t = time_ms([&] { k_dmma<3><<<grid, thr>>>(out, iters, 1.0); });
    printf("warps/SM %2d  m16n8k16: %8.2f TFLOP/s (%.1f MAC/clk/SM @max clk)\n", warps, nw * iters * NACC * 2048 * 2 / t / 1e9, nw * iters * NACC * 2048 / (t * 1e-3) / sms / (clk * 1e3));
    t = time_ms([&] { k_dfma<<<grid, thr>>>(out, iters, 1.0); });
    printf("warps/SM %2d  DFMA    : %8.2f TFLOP/s (%.1f FMA/clk/SM @max clk)\n", warps, nw * 32 * iters * 16 * 2 / t / 1e9, nw * 32 * iters * 16 / (t * 1e-3) / sms / (clk * 1e3));
    t = time_ms([&] { k_mixed<16><<<grid, thr>>>(out, iters / 4, 1.0); });
    { double fl = nw * (iters / 4) * (NACC * 512.0 + 32 * 16.0) * 2; printf("warps/SM %2d  mixed 16xDMMA884+16 DFMA: %8.2f TFLOP/s total (DMMA part %.2f)\n", warps, fl / t / 1e9, nw * (iters / 4) * NACC * 512.0 * 2 / t / 1e9); }
    t = time_ms([&] { k_mixed<64><<<grid, thr>>>(out, iters / 4, 1.0); });
    { double fl = nw * (iters / 4) * (NACC * 512.0 + 32 * 64.0) * 2; printf("warps/SM %2d  mixed 16xDMMA884+64 DFMA: %8.2f TFLOP/s total (DMMA part %.2f)\n", warps, fl / t / 1e9, nw * (iters / 4) * NACC * 512.0 * 2 / t / 1e9); }
  }
  // sustained DMMA for ~3 s to see the power-capped rate
  {
    int warps = 16, thr = warps * 32; double nw = (double)sms * warps;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    int reps = 40;
    for (int r = 0; r < reps; r++) k_dmma<3><<<sms, thr>>>(out, iters * 4, 1.0);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("sustained m16n8k16 x%d launches: %.1f ms, %.2f TFLOP/s\n", reps, ms, nw * iters * 4.0 * NACC * 2048 * 2 * reps / ms / 1e9);
  }
  return 0;
}
