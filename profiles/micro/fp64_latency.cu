// Dependent-issue latencies that bound the per-chain Gibbs step on B200 (one warp, clock64).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, double x0, int n) {
  __shared__ double sm[64];
  double x = x0 + threadIdx.x * 1e-9, y = 1.0000001;
  long long t0, t1;
  int r = 0;
  // 1. dependent DFMA
  t0 = clock64();
  for (int i = 0; i < n; ++i) x = fma(x, y, 1e-9);
  t1 = clock64(); cyc[r++] = t1 - t0;
  // 2. dependent DADD
  t0 = clock64();
  for (int i = 0; i < n; ++i) x = x + y;
  t1 = clock64(); cyc[r++] = t1 - t0;
  // 3. dependent __drcp_rn
  x = 1.5 + threadIdx.x * 1e-9;
  t0 = clock64();
  for (int i = 0; i < n; ++i) x = __drcp_rn(x) + 1.0;
  t1 = clock64(); cyc[r++] = t1 - t0;
  // 4. dependent division
  t0 = clock64();
  for (int i = 0; i < n; ++i) x = 1.0 / x + 1.0;
  t1 = clock64(); cyc[r++] = t1 - t0;
  // 5. dependent sqrt
  t0 = clock64();
  for (int i = 0; i < n; ++i) x = sqrt(x) + 1.0;
  t1 = clock64(); cyc[r++] = t1 - t0;
  // 6. dependent log
  t0 = clock64();
  for (int i = 0; i < n; ++i) x = log(x) + 2.0;
  t1 = clock64(); cyc[r++] = t1 - t0;
  // 7. dependent exp
  t0 = clock64();
  for (int i = 0; i < n; ++i) x = exp(x * 1e-3) + 1.0;
  t1 = clock64(); cyc[r++] = t1 - t0;
  // 8. STS -> bar.sync(32 threads) -> LDS round trip
  t0 = clock64();
  for (int i = 0; i < n; ++i) { sm[threadIdx.x] = x; __syncwarp(); x = sm[(threadIdx.x + 1) & 31] + 1.0; __syncwarp(); }
  t1 = clock64(); cyc[r++] = t1 - t0;
  // 9. 64-bit shuffle + add (one butterfly step)
  t0 = clock64();
  for (int i = 0; i < n; ++i) x += __shfl_xor_sync(0xffffffffu, x, 1);
  t1 = clock64(); cyc[r++] = t1 - t0;
  // 10. dependent FFMA for scale
  float f = (float)x0;
  t0 = clock64();
  for (int i = 0; i < n; ++i) f = fmaf(f, 1.0000001f, 1e-9f);
  t1 = clock64(); cyc[r++] = t1 - t0;
  // 11. 16 independent DFMA chains (throughput per warp)
  double a[16];
  for (int q = 0; q < 16; ++q) a[q] = x0 + q;
  t0 = clock64();
  for (int i = 0; i < n; ++i)
#pragma unroll
    for (int q = 0; q < 16; ++q) a[q] = fma(a[q], y, 1e-9);
  t1 = clock64(); cyc[r++] = t1 - t0;
  for (int q = 0; q < 16; ++q) x += a[q];
  out[threadIdx.x] = x + f;
}
__global__ void k2(double *out, long long *cyc, int n) {   // bar.sync with 64 threads: STS -> bar -> LDS
  __shared__ double sm[64];
  double x = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    sm[threadIdx.x] = x;
    asm volatile("bar.sync 1, 64;" ::: "memory");
    x = sm[(threadIdx.x + 1) & 63] + 1.0;
    asm volatile("bar.sync 1, 64;" ::: "memory");
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = x;
}
int main() {
  double *out; long long *cyc, h[16];
  cudaMalloc(&out, 1024); cudaMalloc(&cyc, 128);
  const int n = 2000;
  const char *names[] = {"DFMA dep", "DADD dep", "__drcp_rn+add", "1/x+add", "sqrt+add", "log+add", "exp+mul+add",
                         "STS->syncwarp->LDS->syncwarp (+add)", "SHFL64+add", "FFMA dep", "16 indep DFMA (per 16)"};
  for (int rep = 0; rep < 2; ++rep) { k<<<1, 32>>>(out, cyc, 1.0, n); cudaDeviceSynchronize(); }
  cudaMemcpy(h, cyc, 88, cudaMemcpyDeviceToHost);
  for (int i = 0; i < 11; ++i) printf("%-40s %8.1f cycles\n", names[i], (double)h[i] / n);
  for (int rep = 0; rep < 2; ++rep) { k2<<<1, 64>>>(out, cyc, n); cudaDeviceSynchronize(); }
  cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-40s %8.1f cycles\n", "STS->bar.sync(64)->LDS->bar.sync(64) (+add)", (double)h[0] / n);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
