// TMA tile::gather4 (sm_100): four arbitrary rows of a 2-D tensor per instruction, landed as a 4 x W box in shared memory.
// Questions for the SEM sweep's foreign-row gathers (~150 rows of 160 B per tile and SM): which box shape does the tensor
// map need, and how many gather4 operations per clock can one / several issuing threads sustain?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_gather4 tma_gather4.cu   (driver entry point fetched at run time)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int W = 20, BATCH = 16, ROUNDS = 64, MAXW = 8;

// warps 0 .. nw-1: lane 0 issues BATCH gather4 copies, waits for them on its own mbarrier, ROUNDS times
__global__ void __launch_bounds__(256) k(const __grid_constant__ CUtensorMap tm, int nrows, int nw, double *out, long long *cyc, int *fail) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar[MAXW];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int w = 0; w < MAXW; ++w) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(s32(&bar[w])));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (wid < nw && lane == 0) {
    double *dst = reinterpret_cast<double *>(sm) + (size_t)wid * BATCH * 4 * W;
    uint32_t ph = 0, seed = 12345u + blockIdx.x * 977u + wid * 131u;
    const long long t0 = clock64();
    for (int r = 0; r < ROUNDS; ++r) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s32(&bar[wid])), "r"(BATCH * 4 * W * 8) : "memory");
#pragma unroll 4
      for (int b = 0; b < BATCH; ++b) {
        int rows[4];
        for (int q = 0; q < 4; ++q) { seed = seed * 1664525u + 1013904223u; rows[q] = (int)((seed >> 8) % (unsigned)nrows); }
        asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n"
                     ::"r"(s32(dst + (size_t)b * 4 * W)), "l"(&tm), "r"(0), "r"(rows[0]), "r"(rows[1]), "r"(rows[2]), "r"(rows[3]),
                       "r"(s32(&bar[wid])) : "memory");
      }
      uint32_t ok = 0;
      const long long w0 = clock64();
      while (!ok) {
        asm volatile("{\n .reg .pred q;\n mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n selp.u32 %0, 1, 0, q;\n}\n"
                     : "=r"(ok) : "r"(s32(&bar[wid])), "r"(ph) : "memory");
        if (!ok && clock64() - w0 > 200000000LL) { *fail = 1; return; }
      }
      ph ^= 1;
    }
    cyc[blockIdx.x * MAXW + wid] = clock64() - t0;
    out[blockIdx.x * MAXW + wid] = dst[3] + dst[4 * W + 1];
  }
}

int main() {
  const int nrows = 1000000;
  double *X;
  cudaMalloc(&X, (size_t)nrows * W * 8);
  {
    double *h = (double *)malloc((size_t)nrows * W * 8);
    for (size_t i = 0; i < (size_t)nrows * W; ++i) h[i] = (double)(i / W) + 0.001 * (i % W);
    cudaMemcpy(X, h, (size_t)nrows * W * 8, cudaMemcpyHostToDevice);
    free(h);
  }
  typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                          const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
  if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  double *out; long long *cyc; int *fail;
  cudaMalloc(&out, 148 * MAXW * 8); cudaMalloc(&cyc, 148 * MAXW * 8); cudaMalloc(&fail, 4);
  const size_t smem = (size_t)MAXW * BATCH * 4 * W * 8;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int boxrows : {1, 4}) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)nrows}, strides[1] = {(cuuint64_t)W * 8};
    cuuint32_t box[2] = {(cuuint32_t)W, (cuuint32_t)boxrows}, es[2] = {1, 1};
    CUresult r = ((Enc)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, X, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box {%d, %d}: encode %d\n", W, boxrows, (int)r);
    if (r != CUDA_SUCCESS) continue;
    for (int nw : {1, 2, 4, 8}) {
      cudaMemset(fail, 0, 4); cudaMemset(cyc, 0, 148 * MAXW * 8);
      k<<<148, 256, smem>>>(tm, nrows, nw, out, cyc, fail);
      cudaError_t e = cudaDeviceSynchronize();
      int hf = 0; long long hc[148 * MAXW]; double ho[4];
      cudaMemcpy(&hf, fail, 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
      cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < 148 * MAXW; ++i) if (hc[i] > mx) mx = hc[i];
      printf("  %d issuing warps per SM: %s%s  %.0f cycles per gather4 per warp, %.1f cycles per gather4 per SM, %.1f B/clk/SM (sample %.3f)\n",
             nw, cudaGetErrorString(e), hf ? " TIMEOUT" : "", (double)mx / (ROUNDS * BATCH), (double)mx / (ROUNDS * BATCH * nw),
             4.0 * W * 8 * ROUNDS * BATCH * nw / (double)mx, ho[0]);
      if (e != cudaSuccess || hf) break;
    }
  }
  return 0;
}
