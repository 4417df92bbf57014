// Which part of a bordered-LDL' column step costs what (64 threads, M = 20, MMAX = 24; timing only).
#include <cstdio>
#include <cuda_runtime.h>
constexpr int MMAX = 24;
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(count) : "memory"); }
// V: 0 straight-line; 1 no reciprocal; 2 no barrier; 3 vector loads of the column; 4 rolled loop over smem rows (left-looking style cost model)
template <int V>
__global__ void k(double *out, long long *cyc, int m, int reps) {
  __shared__ __align__(16) double col[128], Yd[64 * 25], dvec[MMAX];
  const int r = threadIdx.x;
  for (int e = r; e < 64 * 25; e += 64) { const int i = e / 25, j = e % 25; Yd[e] = i == j ? 30.0 + i : 1.0 / (1 + i + j); }
  __syncthreads();
  double a[MMAX];
  long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
    for (int j = 0; j < MMAX; ++j) a[j] = Yd[r * 25 + j];
#pragma unroll
    for (int k = 0; k < MMAX; ++k) {
      if (k < m) {
        double *cb = col + 64 * (k & 1);
        cb[r] = a[k];
        if (r == k) dvec[k] = a[k];
        if (V != 2) named_sync(1, 64);
        const double l = V == 1 ? a[k] * 0.03 : a[k] * __drcp_rn(cb[k]);
        a[k] = l;
        if (V == 3) {
          constexpr int dummy = 0; (void)dummy;
          int j = k + 1;
          if (j & 1) { a[j] = fma(-l, cb[j], a[j]); ++j; }
#pragma unroll
          for (int jj = 0; jj < MMAX; jj += 2)
            if (jj >= k + 1 + ((k + 1) & 1)) { const double2 c2 = *reinterpret_cast<const double2 *>(cb + jj); a[jj] = fma(-l, c2.x, a[jj]); a[jj + 1] = fma(-l, c2.y, a[jj + 1]); }
        } else {
#pragma unroll
          for (int j = k + 1; j < MMAX; ++j) a[j] = fma(-l, cb[j], a[j]);
        }
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < MMAX; ++j) s += a[j];
  out[r] = s + dvec[0];
  if (r == 0) cyc[0] = (t1 - t0) / reps;
}
template <int V> void run(const char *name, double *out, long long *cyc) {
  long long h;
  for (int rep = 0; rep < 2; ++rep) { k<V><<<1, 64>>>(out, cyc, 20, 200); cudaDeviceSynchronize(); }
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s %6lld cycles per elimination = %6.1f per column\n", name, h, h / 20.0);
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 1024); cudaMalloc(&cyc, 8);
  run<0>("straight-line", out, cyc);
  run<1>("no reciprocal", out, cyc);
  run<2>("no barrier", out, cyc);
  run<3>("LDS.128 column loads", out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
