// What does a per-quarter-warp broadcast of (value, column) cost beside conflict-free 128-byte row reads from shared memory?
// Design question of the SEM lag formation (csrc/semsweep.cu): quarter warp = one output row, per nonzero one LDS.128 of
// the gathered X row (8 lanes x 16 B = one wavefront) + the nonzero's value (8 B) and table index (4 B) for all 8 lanes.
//   V0  row reads only (index from arithmetic)                    -- floor
//   V1  + LDS.64 value + LDS.32 index, one address per quarter    -- broadcast loads
//   V2  + values / indices preloaded one per lane, 3 SHFL (width 8) per nonzero
//   V3  + one LDS.128 {value, index, pad} per quarter
//   V4  + LDS.64 value only
//   V5  + LDS.32 index only (constant weight)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o smem_bcast smem_bcast.cu && ./smem_bcast
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ROWS = 256, NNZ = 4096, ITER = 2048;

template <int V>
__global__ void __launch_bounds__(256, 1) k(double *out, long long *cyc) {
  extern __shared__ __align__(128) unsigned char dyn[];
  double *T = reinterpret_cast<double *>(dyn);                 // [ROWS][16]
  double *va = T + ROWS * 16;                                  // [NNZ]
  int *ci = reinterpret_cast<int *>(va + NNZ);                 // [NNZ]
  struct Rec { double v; int c, pad; };
  Rec *rec = reinterpret_cast<Rec *>(ci + NNZ);                 // [NNZ] {va, ci}: one 16-byte load
  const int tid = threadIdx.x, lane = tid & 31, l = lane & 7, q = tid >> 3;
  for (int e = tid; e < ROWS * 16; e += 256) T[e] = 1.0 + e * 1e-6;
  for (int e = tid; e < NNZ; e += 256) {
    va[e] = 0.5 + e * 1e-5; ci[e] = (e * 37 + 11) & (ROWS - 1);
    rec[e].v = va[e]; rec[e].c = ci[e]; rec[e].pad = 0;
  }
  __syncthreads();
  double a0 = 0.0, a1 = 0.0;
  const long long t0 = clock64();
  int p = (q * 61) & (NNZ - 1);
#pragma unroll 1
  for (int it = 0; it < ITER; it += 8) {
    double vreg = 0.0; int creg = 0;
    if (V == 2) { vreg = va[(p + l) & (NNZ - 1)]; creg = ci[(p + l) & (NNZ - 1)]; }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int pp = (p + u) & (NNZ - 1);
      double v = 0.5; int c = (pp * 37 + 11) & (ROWS - 1);
      if (V == 1) { v = va[pp]; c = ci[pp]; }
      if (V == 2) { v = __shfl_sync(0xffffffffu, vreg, u, 8); c = __shfl_sync(0xffffffffu, creg, u, 8); }
      if (V == 3) { const int4 r = *reinterpret_cast<const int4 *>(rec + pp); v = __hiloint2double(r.y, r.x); c = r.z; }
      if (V == 4) { v = va[pp]; }
      if (V == 5) { c = ci[pp]; }
      const double2 x = *reinterpret_cast<const double2 *>(T + c * 16 + 2 * (l ^ (2 * (c & 3))));
      a0 = fma(v, x.x, a0); a1 = fma(v, x.y, a1);
    }
    p = (p + 8 * 32 + 8) & (NNZ - 1);
  }
  const long long t1 = clock64();
  out[blockIdx.x * 256 + tid] = a0 + a1;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int V>
void run(const char *name) {
  double *out; long long *cyc;
  cudaMalloc(&out, 148 * 256 * 8); cudaMalloc(&cyc, 148 * 8);
  const size_t smem = ROWS * 128 + NNZ * 8 + NNZ * 4 + NNZ * 16;
  cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<V><<<148, 256, smem>>>(out, cyc);
  k<V><<<148, 256, smem>>>(out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double m = 0; for (int i = 0; i < 148; ++i) m += h[i];
  m /= 148;
  // 8 warps x ITER row-read instructions (4 wavefronts each)
  printf("%-44s %s  %.0f cycles, %.1f cycles per step of 8 warps x 4 rows (row reads alone: 32 wavefronts)\n", name,
         cudaGetErrorString(e), m, m / ITER);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("V0 rows only");
  run<1>("V1 + LDS.64 value + LDS.32 index");
  run<2>("V2 + 3 SHFL per nonzero (preloaded)");
  run<3>("V3 + LDS.128 record");
  run<4>("V4 + LDS.64 value");
  run<5>("V5 + LDS.32 index");
  return 0;
}
