// DFMA issue rate on one SM of B200 as a function of how many warps run and where they sit (one CTA, clock64).
// Answers: is the FP64 pipe per SM sub-partition or shared, and what is its rate per warp instruction.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, int n, int mask_stride) {
  // active warps: warp w runs iff (w % mask_stride) == 0
  const int w = threadIdx.x >> 5;
  double a[16];
  for (int q = 0; q < 16; ++q) a[q] = 1.0 + q + threadIdx.x * 1e-9;
  const double y = 1.0000001;
  __syncthreads();
  long long t0 = clock64();
  if (w % mask_stride == 0)
    for (int i = 0; i < n; ++i)
#pragma unroll
      for (int q = 0; q < 16; ++q) a[q] = fma(a[q], y, 1e-9);
  long long t1 = clock64();
  double x = 0;
  for (int q = 0; q < 16; ++q) x += a[q];
  out[threadIdx.x] = x;
  if ((threadIdx.x & 31) == 0) cyc[w] = t1 - t0;
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8 * 32);
  const int n = 2000;
  struct { int warps, stride; const char *what; } cases[] = {
    {1, 1, "1 warp"}, {2, 1, "2 warps, 2 sub-partitions"}, {4, 1, "4 warps, 4 sub-partitions"},
    {8, 1, "8 warps, 2 per sub-partition"}, {16, 1, "16 warps, 4 per sub-partition"},
    {8, 4, "2 warps, same sub-partition"}, {16, 4, "4 warps, same sub-partition"}, {32, 4, "8 warps, same sub-partition"}};
  for (auto &c : cases) {
    long long h[32];
    for (int rep = 0; rep < 2; ++rep) { k<<<1, 32 * c.warps>>>(out, cyc, n, c.stride); cudaDeviceSynchronize(); }
    cudaMemcpy(h, cyc, 8 * 32, cudaMemcpyDeviceToHost);
    long long mx = 0; int active = 0;
    for (int w = 0; w < c.warps; ++w) if (w % c.stride == 0) { ++active; if (h[w] > mx) mx = h[w]; }
    printf("%-34s %7.2f cycles per DFMA per warp, %6.2f DFMA lanes per clock per SM\n", c.what, (double)mx / (16.0 * n),
           32.0 * 16.0 * n * active / (double)mx);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
