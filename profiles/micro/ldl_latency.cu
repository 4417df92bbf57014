// Cycles of one bordered LDL' elimination (M = 20, 43 rows, 64 threads), repeated in a loop (warm I-cache).
#include <cstdio>
#include <cuda_runtime.h>
constexpr int MMAX = 24;
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(count) : "memory"); }
template <int VARIANT>
__global__ void k(double *out, long long *cyc, int m, int nrows, int reps) {
  __shared__ double col[128], Yd[64 * 25], Z[3 * MMAX], dvec[MMAX];
  const int r = threadIdx.x;
  for (int e = r; e < 64 * 25; e += 64) { const int i = e / 25, j = e % 25; Yd[e] = i == j ? 30.0 + i : 1.0 / (1 + i + j); }
  __syncthreads();
  double a[MMAX];
  long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
    for (int j = 0; j < MMAX; ++j) a[j] = (j < m && r < nrows) ? Yd[r * 25 + j] : 0.0;
#pragma unroll
    for (int k = 0; k < MMAX; ++k) {
      if (k < m) {
        double *cb = col + 64 * (k & 1);
        cb[r] = a[k];
        if (VARIANT >= 1) { if (r >= m && r < m + 3) Z[(r - m) * MMAX + k] = a[k]; if (r == k) dvec[k] = a[k]; }
        named_sync(1, 64);
        const double rk = VARIANT == 3 ? 1.0 / cb[k] : __drcp_rn(cb[k]);
        if (r > k && r < nrows) {
          const double l = a[k] * rk;
          a[k] = l;
#pragma unroll
          for (int j = k + 1; j < MMAX; ++j)
            if (j < m) a[j] = fma(-l, cb[j], a[j]);
        }
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < MMAX; ++j) s += a[j];
  out[r] = s + Z[0] + dvec[0];
  if (r == 0) cyc[0] = (t1 - t0) / reps;
}
int main() {
  double *out; long long *cyc, h;
  cudaMalloc(&out, 1024); cudaMalloc(&cyc, 8);
  for (int v = 0; v < 2; ++v)
    for (int rep = 0; rep < 2; ++rep) {
      if (v == 0) k<0><<<1, 64>>>(out, cyc, 20, 43, 200); else k<1><<<1, 64>>>(out, cyc, 20, 43, 200);
      cudaDeviceSynchronize();
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      if (rep) printf("variant %d: %lld cycles per elimination = %.1f per column\n", v, h, h / 20.0);
    }
  // several CTAs of 4 x 64 threads each running eliminations concurrently on one SM would need more plumbing; cold run:
  k<1><<<1, 64>>>(out, cyc, 20, 43, 1);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("single cold pass: %lld cycles = %.1f per column\n", h, h / 20.0);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
