// DMMA (mma.sync.m8n8k4.f64) issue rate on one SM of B200 against the vector DFMA rate, and whether the two share a pipe.
// One CTA, clock64; every warp runs `n` rounds of 8 independent accumulators.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
// mode 0: DMMA only; 1: DFMA only; 2: even warps DMMA, odd warps DFMA (same sub-partition pairs: warps w and w + 4 differ in kind)
__global__ void k(double *out, long long *cyc, int n, int mode) {
  const int w = threadIdx.x >> 5;
  double c[8][2], a[16];
  for (int q = 0; q < 8; ++q) { c[q][0] = q; c[q][1] = q + 0.5; }
  for (int q = 0; q < 16; ++q) a[q] = 1.0 + q + threadIdx.x * 1e-9;
  const double x = 1.0000001 + threadIdx.x * 1e-12, y = 0.9999999;
  const bool use_mma = mode == 0 || (mode == 2 && ((w >> 2) & 1) == 0);
  __syncthreads();
  const long long t0 = clock64();
  if (use_mma) {
    for (int i = 0; i < n; ++i)
#pragma unroll
      for (int q = 0; q < 8; ++q) dmma(c[q], x, y);
  } else {
    for (int i = 0; i < n; ++i)
#pragma unroll
      for (int q = 0; q < 16; ++q) a[q] = fma(a[q], y, 1e-9);
  }
  const long long t1 = clock64();
  double s = 0;
  for (int q = 0; q < 8; ++q) s += c[q][0] + c[q][1];
  for (int q = 0; q < 16; ++q) s += a[q];
  out[threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) cyc[w] = t1 - t0;
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8 * 32);
  const int n = 2000;
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%s, %d SMs, clock %d kHz\n", pr.name, pr.multiProcessorCount, clk);
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {1, 4, 8, 16, 32}) {
      if (mode == 2 && warps < 8) continue;
      long long h[32];
      for (int rep = 0; rep < 2; ++rep) { k<<<1, 32 * warps>>>(out, cyc, n, mode); cudaDeviceSynchronize(); }
      cudaMemcpy(h, cyc, 8 * 32, cudaMemcpyDeviceToHost);
      long long mx_m = 0, mx_f = 0; int nm = 0, nf = 0;
      for (int w = 0; w < warps; ++w) {
        const bool mm = mode == 0 || (mode == 2 && ((w >> 2) & 1) == 0);
        if (mm) { ++nm; if (h[w] > mx_m) mx_m = h[w]; } else { ++nf; if (h[w] > mx_f) mx_f = h[w]; }
      }
      const char *what = mode == 0 ? "DMMA only" : mode == 1 ? "DFMA only" : "DMMA + DFMA warps";
      printf("%-18s %2d warps:", what, warps);
      if (nm) printf("  DMMA %6.2f cycles per warp instruction, %7.1f FMA per clock per SM", (double)mx_m / (8.0 * n), 256.0 * 8.0 * n * nm / (double)mx_m);
      if (nf) printf("  DFMA %6.2f cycles per warp instruction, %7.1f FMA per clock per SM", (double)mx_f / (16.0 * n), 32.0 * 16.0 * n * nf / (double)mx_f);
      printf("\n");
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
