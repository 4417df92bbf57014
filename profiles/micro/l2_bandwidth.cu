// L2 -> SM bandwidth of B200 for a working set that stays in L2 (the in-loop regime of the cfg4 sweep: 29 MB of operands read
// once per Gibbs iteration, L2-resident after the first).  Two access paths: 16-byte loads (ld.global.cg) and bulk copies into
// shared memory (cp.async.bulk + mbarrier, the path the ring sweep uses).  Also the same loops over 1 GB (HBM) for reference.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2_bandwidth l2_bandwidth.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512) ldg_kernel(const int4 *__restrict__ p, size_t n16, int reps, int *out) {
  int acc = 0;
  for (int r = 0; r < reps; ++r)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
      int4 v;
      asm volatile("ld.global.cg.v4.s32 {%0,%1,%2,%3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + i));
      acc += v.x ^ v.y ^ v.z ^ v.w;
    }
  if (acc == 0x12345678) *out = acc;
}

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one CTA per SM, one producer thread, S stages of CH bytes; nobody reads the data (pure transfer rate)
template <int S, int CH>
__global__ void __launch_bounds__(128) bulk_kernel(const char *__restrict__ p, size_t bytes, int reps) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) uint64_t full[S];
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(s32(&full[s])));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const size_t nch = bytes / CH;
  uint32_t ph[S];
  for (int s = 0; s < S; ++s) ph[s] = 0;
  int s = 0;
  long long issued = 0;
  for (int r = 0; r < reps; ++r)
    for (size_t c = blockIdx.x; c < nch; c += gridDim.x) {
      if (issued >= S) {      // wait for the copy that used this stage
        uint32_t ok = 0;
        while (!ok)
          asm volatile("{\n .reg .pred q;\n mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n selp.u32 %0, 1, 0, q;\n}\n"
                       : "=r"(ok) : "r"(s32(&full[s])), "r"(ph[s]) : "memory");
        ph[s] ^= 1;
      }
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s32(&full[s])), "r"(CH) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                   ::"r"(s32(sm + (size_t)s * CH)), "l"(p + c * CH), "r"(CH), "r"(s32(&full[s])) : "memory");
      ++issued;
      if (++s == S) s = 0;
    }
  for (int k = 0; k < S && k < issued; ++k) {     // drain
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n .reg .pred q;\n mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n selp.u32 %0, 1, 0, q;\n}\n"
                   : "=r"(ok) : "r"(s32(&full[s])), "r"(ph[s]) : "memory");
    ph[s] ^= 1;
    if (++s == S) s = 0;
  }
}

static float timed(void (*launch)(void *), void *ctx) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  launch(ctx);                      // warm-up (fills L2)
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  launch(ctx);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

struct Ctx { const char *p; size_t bytes; int reps, grid; int *out; };

int main() {
  char *buf; int *out;
  const size_t big = (size_t)1 << 30;
  cudaMalloc(&buf, big); cudaMemset(buf, 1, big); cudaMalloc(&out, 4);
  constexpr int S = 8, CH = 16384;
  cudaFuncSetAttribute(bulk_kernel<S, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, S * CH);
  for (size_t mb : {16, 29, 58, 100, 1024}) {
    const size_t bytes = mb == 1024 ? big : (mb << 20) / CH * CH;
    const int reps = mb == 1024 ? 4 : (int)(4096 / mb);
    Ctx c{buf, bytes, reps, 148, out};
    for (int cps : {1, 2, 4}) {
      c.grid = 148 * cps;
      const float ms = timed([](void *q) { Ctx *c = (Ctx *)q; ldg_kernel<<<c->grid, 512>>>((const int4 *)c->p, c->bytes / 16, c->reps, c->out); }, &c);
      printf("%5zu MB  ld.global.cg 16 B, %d CTAs x 512 per SM: %8.1f GB/s\n", mb, cps, (double)bytes * reps / ms / 1e6);
    }
    c.grid = 148;
    const float ms = timed([](void *q) { Ctx *c = (Ctx *)q; bulk_kernel<S, CH><<<c->grid, 128, S * CH>>>(c->p, c->bytes, c->reps); }, &c);
    printf("%5zu MB  cp.async.bulk 16 KB x 8 stages per SM:      %8.1f GB/s\n", mb, (double)bytes * reps / ms / 1e6);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
