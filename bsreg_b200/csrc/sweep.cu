// N-scale kernels of the Gibbs hot path (hand-written for sm_100a, FP64, no tensor cores):
//
//   K1  csr_spmm_kernel      W*B for dense row-major B        (R/15_sar.R:57, R/15_sem.R:63-64, R/12_slx.R:65)
//   K1+K2 sweep_gram_kernel  ONE pass over W, y, X per Gibbs iteration: forms the spatial lags
//                            Wy (and WX) on the fly and accumulates G = Z'Z, Z = [y|Wy|X|WX]
//                            (R/10_lm.R:36, R/15_sar.R:133-135, R/15_sem.R:34-35 and every
//                             N-scale quadratic form of R/10_lm.R:176, R/15_sar.R:110-114,
//                             R/15_sem.R:116-120 are polynomials in lambda of blocks of G)
//   K2  residual_ss_kernel   direct per-parameter-set ||(y - lam Wy) - (X - rho WX) beta||^2 with
//                            warp-shuffle + block reductions (R/10_lm.R:176,192)
//   K4  probes / dot         Rademacher probes and u'T_j(W)u dots for the trace moments
//
// All reductions are bit-reproducible: fixed-order trees inside a CTA, and 64-bit
// fixed-point integer atomics across CTAs (integer addition is associative, so the
// result does not depend on CTA arrival order).
#include "internal.cuh"
#include "rng.cuh"

namespace bsreg {

static int g_sweep_launches = 0;
int sweep_launch_count() { return g_sweep_launches; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------
// layout helpers
// ---------------------------------------------------------------------------
__global__ void transpose_kernel(const double *__restrict__ src, double *__restrict__ dst, int n, int m, int ld, int col0) {
  __shared__ double tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {   // src column-major: element (r, c) at c*n + r
    const int r = r0 + threadIdx.x, c = c0 + j;
    tile[j][threadIdx.x] = (r < n && c < m) ? src[(size_t)c * n + r] : 0.0;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < n && c < m) dst[(size_t)r * ld + col0 + c] = tile[threadIdx.x][j];
  }
}

void launch_transpose(const double *src, double *dst, int n, int m, int ld, int col0, cudaStream_t s) {
  dim3 grid((n + 31) / 32, (m + 31) / 32), block(32, 8);
  transpose_kernel<<<grid, block, 0, s>>>(src, dst, n, m, ld, col0);
}

// ---------------------------------------------------------------------------
// K1: out[i, col0 + c] = alpha * sum_p va[p] * B[ci[p], c] + beta * C[i, col0 + c]
// one thread per (row, column); consecutive threads read consecutive columns of a
// gathered row, so the gather is coalesced for m >= 4.
// ---------------------------------------------------------------------------
__global__ void csr_spmm_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci,
                                const double *__restrict__ va, const double *__restrict__ B, int m, int ldb,
                                double *__restrict__ out, int ldo, int col0, double alpha,
                                const double *__restrict__ C, double beta) {
  const size_t total = (size_t)n * m;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(idx / m), c = (int)(idx % m);
    double acc = 0.0;
    for (int p = rp[row], pe = rp[row + 1]; p < pe; ++p) acc = fma(va[p], B[(size_t)ci[p] * ldb + c], acc);
    double v = alpha * acc;
    if (C != nullptr) v += beta * C[(size_t)row * ldo + col0 + c];
    out[(size_t)row * ldo + col0 + c] = v;
  }
}

void launch_spmm(int n, const int *rp, const int *ci, const double *va, const double *B, int m, int ldb, double *out,
                 int ldo, int col0, double alpha, const double *C, double beta, cudaStream_t s) {
  const size_t total = (size_t)n * m;
  const int block = 256;
  const int grid = (int)((total + block - 1) / block < 148 * 16 ? (total + block - 1) / block : 148 * 16);
  csr_spmm_kernel<<<grid > 0 ? grid : 1, block, 0, s>>>(n, rp, ci, va, B, m, ldb, out, ldo, col0, alpha, C, beta);
}

// ---------------------------------------------------------------------------
// column sums of squares of Z (setup only; fixes the fixed-point scales)
// ---------------------------------------------------------------------------
__device__ __forceinline__ double z_entry(const DeviceData &dd, int row, int c) {
  if (c == 0) return dd.y[row];
  if (c == 1) return dd.Wy ? dd.Wy[row] : 0.0;
  if (c < 2 + dd.m) return dd.X[(size_t)row * dd.m + c - 2];
  return dd.WX[(size_t)row * dd.m + c - 2 - dd.m];
}

__global__ void colnorm_kernel(DeviceData dd, double *norms) {
  __shared__ double red[256];
  const int c = blockIdx.x;
  double acc = 0.0;
  for (int row = threadIdx.x; row < dd.n; row += blockDim.x) {
    const double v = z_entry(dd, row, c);
    acc = fma(v, v, acc);
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) norms[c] = red[0];
}

void launch_colnorms(const DeviceData &dd, double *norms, cudaStream_t s) {
  colnorm_kernel<<<dd.d, 256, 0, s>>>(dd, norms);
}

// ---------------------------------------------------------------------------
// K1 + K2 fused sweep (fast path, d <= 88).
//
// A CTA stages a tile of R rows of Z in shared memory (X and y streamed from HBM
// exactly once, Wy/WX gathered through W), then accumulates Z_tile' Z_tile in
// registers: the upper triangle of G is cut into 4x4 blocks, thread = (block,
// row group).  Lanes of a warp hold different blocks of the SAME row, so the
// shared-memory reads are broadcasts, and each thread issues 16 DFMA per 4
// LDS.128.  CTAs are persistent (grid = SMs x occupancy, static tile->CTA map).
// ---------------------------------------------------------------------------
template <int NB>
struct SweepCfg {
  static constexpr int B = NB * (NB + 1) / 2;
  static constexpr int RG = (B > 128) ? 1 : (256 + B / 2) / B;
  static constexpr int T = ((B * RG + 31) / 32) * 32;
  static constexpr int RPT = (64 + RG / 2) / RG > 0 ? (64 + RG / 2) / RG : 1;
  static constexpr int R = RG * RPT;
  static constexpr int DP = 4 * NB;
};

template <int NB>
__global__ void __launch_bounds__(SweepCfg<NB>::T) sweep_gram_kernel(DeviceData dd, int ntiles) {
  using Cfg = SweepCfg<NB>;
  constexpr int B = Cfg::B, RG = Cfg::RG, T = Cfg::T, RPT = Cfg::RPT, R = Cfg::R, DP = Cfg::DP;
  __shared__ __align__(16) double Z[R * DP];
  __shared__ int s_last;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int m = dd.m, d = dd.d, n = dd.n;
  const bool sem = dd.model == 2, has_w = dd.model != 0;

  // thread -> (block of G, row group)
  const int blk = tid % B, rg = tid / B;
  const bool active = rg < RG;
  int bi = 0, bj = blk;
  while (bj >= NB - bi) { bj -= NB - bi; ++bi; }
  bj += bi;                                   // bi <= bj < NB

  double acc[4][4];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[p][q] = 0.0;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int r0 = tile * R;
    // ---- phase A1: stream y and X rows (coalesced), zero the padding ----
    for (int e = tid; e < R * DP; e += T) {
      const int r = e / DP, c = e - r * DP;
      const int row = r0 + r;
      if (c == 1 || (sem && c >= 2 + m && c < 2 + 2 * m)) continue;   // lag columns are written below
      double v = 0.0;
      if (row < n) {
        if (c == 0) v = dd.y[row];
        else if (c < 2 + m) v = dd.X[(size_t)row * m + (c - 2)];
      }
      Z[e] = v;
    }
    // ---- phase A2: Wy, four lanes per row, fixed-order combine ----
    for (int rb = wid * 8; rb < R; rb += (T / 32) * 8) {
      const int r = rb + (lane >> 2), q = lane & 3;
      const int row = r0 + r;
      double s = 0.0;
      if (has_w && r < R && row < n) {
        for (int p = dd.rp[row] + q, pe = dd.rp[row + 1]; p < pe; p += 4) s = fma(dd.va[p], dd.y[dd.ci[p]], s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (q == 0 && r < R) Z[r * DP + 1] = s;
    }
    // ---- phase A3 (SEM): WX, one thread per (row, column), coalesced row gathers ----
    if (sem) {
      for (int e = tid; e < R * m; e += T) {
        const int r = e / m, c = e - r * m;
        const int row = r0 + r;
        double s = 0.0;
        if (row < n) {
          for (int p = dd.rp[row], pe = dd.rp[row + 1]; p < pe; ++p)
            s = fma(dd.va[p], dd.X[(size_t)dd.ci[p] * m + c], s);
        }
        Z[r * DP + 2 + m + c] = s;
      }
    }
    __syncthreads();
    // ---- phase B: rank-R update of the 4x4 register block ----
    if (active) {
#pragma unroll 4
      for (int i = 0; i < RPT; ++i) {
        const double *zr = Z + (rg + i * RG) * DP;
        const double2 a01 = *reinterpret_cast<const double2 *>(zr + 4 * bi);
        const double2 a23 = *reinterpret_cast<const double2 *>(zr + 4 * bi + 2);
        const double2 b01 = *reinterpret_cast<const double2 *>(zr + 4 * bj);
        const double2 b23 = *reinterpret_cast<const double2 *>(zr + 4 * bj + 2);
        const double a[4] = {a01.x, a01.y, a23.x, a23.y};
        const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[p][q] = fma(a[p], b[q], acc[p][q]);
      }
    }
    __syncthreads();
  }

  // ---- combine the row groups of this CTA in fixed order ----
  if (RG > 1) {
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        Z[tid] = active ? acc[p][q] : 0.0;
        __syncthreads();
        if (rg == 0) {
          double s = 0.0;
          for (int g = 0; g < RG; ++g) s += Z[blk + g * B];
          acc[p][q] = s;
        }
        __syncthreads();
      }
  }
  // ---- order-independent accumulation across CTAs: 64-bit fixed point ----
  if (rg == 0) {
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int gi = 4 * bi + p, gj = 4 * bj + q;
        if (gi < d && gj < d && gi <= gj) {
          const long long f = __double2ll_rn(acc[p][q] * dd.Ginv_scale[gi * d + gj]);
          atomicAdd(reinterpret_cast<unsigned long long *>(dd.Gfix + gi * d + gj), (unsigned long long)f);
        }
      }
  }
  // ---- last CTA converts to double, mirrors, and re-zeroes the accumulators ----
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(dd.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
    for (int e = tid; e < d * d; e += T) {
      const int i = e / d, j = e - i * d;
      if (i <= j) {
        const long long f = __ldcg(dd.Gfix + e);
        const double v = (double)f * dd.Gscale[e];
        dd.G[i * d + j] = v;
        dd.G[j * d + i] = v;
        dd.Gfix[e] = 0;
      }
    }
    if (tid == 0) *dd.ticket = 0u;
  }
}


// ---------------------------------------------------------------------------
// K1 + K2 fused sweep, main path for d <= 24 (LM / SAR up to K = 22): TMA-staged.
//
// One persistent, warp-specialised CTA per SM around an S-stage shared-memory ring:
//   producer warp : per tile of R rows, 1-D TMA bulk copies (cp.async.bulk, complete_tx on
//                   an mbarrier) of the K column segments of X (X is kept COLUMN-major and
//                   zero-padded to a whole number of tiles, so a tile column is one
//                   contiguous R*8-byte run), of y, and of the tile's CSR slice
//                   (row_ptr, col_idx, val).  Nothing N-scale passes through registers.
//   gather warps  : spatial lag of the tile, Wy[r] = sum_p val[p] * y[col[p]] (CSR slice
//                   from shared memory, y gathered from L2), written as column 1 of the tile.
//   math warps    : lane = row, warp = one 6x6 block of G = Z'Z (diagonal blocks 21 DFMA,
//                   off-diagonal 36 DFMA per row for 6 / 12 LDS.64); the column-major tile
//                   makes every LDS conflict-free.  Blocks are dealt to warps so that the
//                   four SM sub-partitions carry equal DFMA work.
// full[s] / lag[s] / empty[s] mbarriers order producer -> gather -> math -> producer.
// ---------------------------------------------------------------------------
struct BlockMap { unsigned char gi[32], gj[32]; };

template <int NG>
struct S7 {
  static constexpr int NBLK = NG * (NG + 1) / 2;
  static constexpr int WPB = (NBLK >= 8) ? 1 : (NBLK >= 5 ? 2 : 4);   // math warps per block of G
  static constexpr int MATH = NBLK * WPB;
  static constexpr int GATHER = 4;
  static constexpr int T = (MATH + GATHER + 1) * 32;                 // + producer warp
  static constexpr int RPT = 4;                                       // rows per lane per tile
  static constexpr int R = 32 * RPT;                                  // rows per tile
  static constexpr int NC = 6 * NG;                                   // tile columns incl. zero padding
  static constexpr int S = 4;                                         // ring stages
};

constexpr int S7_MAXNZ = 2560;   // nonzeros of one tile that fit the ring

__host__ __device__ inline size_t s7_stage_doubles(int R, int NC, int maxnz) {
  // Zt[NC][R] | VA[maxnz+4] | CI[maxnz+8] (ints) | RP[R+8] (ints); maxnz % 4 == 0 keeps 16-byte alignment
  return (size_t)R * NC + (maxnz + 4) + (maxnz + 8) / 2 + (R + 8) / 2;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int NG>
__global__ void __launch_bounds__(S7<NG>::T, 1)
sweep_tma_kernel(DeviceData dd, int ntiles, const int *__restrict__ tile_rp, int maxnz, BlockMap bm) {
  using C = S7<NG>;
  constexpr int T = C::T, R = C::R, NC = C::NC, S = C::S, RPT = C::RPT, WPB = C::WPB, MATH = C::MATH, GATHER = C::GATHER;
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) uint64_t full[S], lag[S], empty[S];
  __shared__ int s_last;
  const size_t stage_d = s7_stage_doubles(R, NC, maxnz);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int m = dd.m, d = dd.d;
  const bool has_w = dd.model != 0;

  auto stZ = [&](int s) { return smem + (size_t)s * stage_d; };
  auto stVA = [&](int s) { return stZ(s) + (size_t)R * NC; };
  auto stCI = [&](int s) { return reinterpret_cast<int *>(stVA(s) + maxnz + 4); };
  auto stRP = [&](int s) { return stCI(s) + maxnz + 8; };

  for (size_t e = tid; e < (size_t)S * stage_d; e += T) smem[e] = 0.0;   // padding columns stay zero
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&lag[s], GATHER); mbar_init(&empty[s], MATH); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic zero-fill before async-proxy writes

  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const size_t ldx = (size_t)ntiles * R;                            // padded column stride of Xc / length of ypad

  if (wid == MATH + GATHER) {
    // =========================== producer ===========================
    int p0 = 0, p1 = 0;
    if (has_w && my_tiles > 0) { p0 = tile_rp[blockIdx.x]; p1 = tile_rp[blockIdx.x + 1]; }
    for (int j = 0; j < my_tiles; ++j) {
      const int s = j % S, tile = blockIdx.x + j * gridDim.x, r0 = tile * R;
      int q0 = 0, q1 = 0;                                           // prefetch next tile's CSR range
      if (has_w && j + 1 < my_tiles) { q0 = tile_rp[tile + gridDim.x]; q1 = tile_rp[tile + gridDim.x + 1]; }
      mbar_wait(&empty[s], ((j / S) & 1) ^ 1);
      const int a0 = p0 & ~3, b0 = p0 & ~1;
      const uint32_t ci_bytes = has_w ? (uint32_t)((p1 - a0 + 3) / 4) * 16u : 0u;
      const uint32_t va_bytes = has_w ? (uint32_t)((p1 - b0 + 1) / 2) * 16u : 0u;
      const uint32_t rp_bytes = has_w ? (uint32_t)(R + 4) * 4u : 0u;
      if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)(m + 1) * R * 8u + ci_bytes + va_bytes + rp_bytes);
      __syncwarp();
      double *Z = stZ(s);
      for (int c = lane; c < m; c += 32) tma_bulk_g2s(Z + (size_t)(2 + c) * R, dd.Xc + (size_t)c * ldx + r0, R * 8u, &full[s]);
      if (lane == 31) tma_bulk_g2s(Z, dd.ypad + r0, R * 8u, &full[s]);
      if (has_w) {
        if (lane == 30) tma_bulk_g2s(stRP(s), dd.rp + r0, rp_bytes, &full[s]);
        if (lane == 29 && ci_bytes) tma_bulk_g2s(stCI(s), dd.ci + a0, ci_bytes, &full[s]);
        if (lane == 28 && va_bytes) tma_bulk_g2s(stVA(s), dd.va + b0, va_bytes, &full[s]);
      }
      p0 = q0; p1 = q1;
    }
  } else if (wid >= MATH) {
    // =========================== gather: Wy of the tile ===========================
    const int g = tid - MATH * 32;                                  // 0 .. 127
    for (int j = 0; j < my_tiles; ++j) {
      const int s = j % S, tile = blockIdx.x + j * gridDim.x, r0 = tile * R;
      mbar_wait(&full[s], (j / S) & 1);
      double *Z = stZ(s);
      if (has_w) {
        const int *RP = stRP(s), *CI = stCI(s);
        const double *VA = stVA(s);
        const int p0 = RP[0], a0 = p0 & ~3, b0 = p0 & ~1;
        for (int r = g; r < R; r += GATHER * 32) {
          double acc = 0.0;
          if (r0 + r < dd.n) {
            const int pb = RP[r], pe = RP[r + 1];
            int p = pb;
            for (; p + 4 <= pe; p += 4) {                           // four gathers in flight
              const double y0 = dd.y[CI[p - a0]], y1 = dd.y[CI[p + 1 - a0]], y2 = dd.y[CI[p + 2 - a0]], y3 = dd.y[CI[p + 3 - a0]];
              acc = fma(VA[p - b0], y0, acc); acc = fma(VA[p + 1 - b0], y1, acc);
              acc = fma(VA[p + 2 - b0], y2, acc); acc = fma(VA[p + 3 - b0], y3, acc);
            }
            for (; p < pe; ++p) acc = fma(VA[p - b0], dd.y[CI[p - a0]], acc);
          }
          Z[R + r] = acc;                                           // column 1 = Wy
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&lag[s]);
    }
  } else {
    // =========================== math: this warp's 6x6 block ===========================
    const int blk = wid % C::NBLK, wr = wid / C::NBLK;
    const int gi = bm.gi[blk], gj = bm.gj[blk];
    const bool diag = gi == gj;
    double acc[6][6];
#pragma unroll
    for (int p = 0; p < 6; ++p)
#pragma unroll
      for (int q = 0; q < 6; ++q) acc[p][q] = 0.0;
    for (int j = 0; j < my_tiles; ++j) {
      const int s = j % S;
      mbar_wait(&full[s], (j / S) & 1);
      mbar_wait(&lag[s], (j / S) & 1);
      const double *Z = stZ(s);
#pragma unroll
      for (int rr = 0; rr < RPT / WPB; ++rr) {
        const int r = lane + 32 * (wr + WPB * rr);
        double a[6], b[6];
#pragma unroll
        for (int p = 0; p < 6; ++p) a[p] = Z[(6 * gi + p) * R + r];
        if (diag) {
#pragma unroll
          for (int p = 0; p < 6; ++p)
#pragma unroll
            for (int q = p; q < 6; ++q) acc[p][q] = fma(a[p], a[q], acc[p][q]);
        } else {
#pragma unroll
          for (int p = 0; p < 6; ++p) b[p] = Z[(6 * gj + p) * R + r];
#pragma unroll
          for (int p = 0; p < 6; ++p)
#pragma unroll
            for (int q = 0; q < 6; ++q) acc[p][q] = fma(a[p], b[q], acc[p][q]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    // lanes hold disjoint rows: fixed-order shuffle tree, then order-independent fixed-point atomics
#pragma unroll
    for (int p = 0; p < 6; ++p)
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        if (diag && q < p) continue;
        const double v = warp_sum(acc[p][q]);
        const int ci_ = 6 * gi + p, cj_ = 6 * gj + q;
        if (lane == 0 && ci_ < d && cj_ < d) {
          const long long f = __double2ll_rn(v * dd.Ginv_scale[ci_ * d + cj_]);
          atomicAdd(reinterpret_cast<unsigned long long *>(dd.Gfix + ci_ * d + cj_), (unsigned long long)f);
        }
      }
  }
  // ---- last CTA converts to double, mirrors, and re-zeroes the accumulators ----
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(dd.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
    for (int e = tid; e < d * d; e += T) {
      const int i = e / d, j2 = e - i * d;
      if (i <= j2) {
        const long long f = __ldcg(dd.Gfix + e);
        const double v = (double)f * dd.Gscale[e];
        dd.G[i * d + j2] = v;
        dd.G[j2 * d + i] = v;
        dd.Gfix[e] = 0;
      }
    }
    if (tid == 0) *dd.ticket = 0u;
  }
}

// generic path (any d): one warp per entry of the upper triangle over materialised lags
__global__ void gram_generic_kernel(DeviceData dd) {
  const int d = dd.d;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int total = d * (d + 1) / 2;
  for (int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < total; e += warps) {
    int i = 0, j = e;
    while (j >= d - i) { j -= d - i; ++i; }
    j += i;
    double acc = 0.0;
    for (int row = lane; row < dd.n; row += 32)
      acc = fma(z_entry(dd, row, i), z_entry(dd, row, j), acc);
    acc = warp_sum(acc);
    if (lane == 0) { dd.G[i * d + j] = acc; dd.G[j * d + i] = acc; }
  }
}

bool sweep_fast_supported(const DeviceData &dd) { return dd.d <= 88; }

int sweep_tma_rows_per_tile(int d) { return d <= 24 ? S7<4>::R : 0; }
bool sweep_tma_supported(const DeviceData &dd) {
  return dd.d <= 24 && dd.model != 2 && dd.Xc != nullptr && (dd.model == 0 || dd.tile_rp != nullptr) &&
         dd.tile_maxnz <= S7_MAXNZ;
}

// deal the blocks of G to warps so that the four SM sub-partitions (warp % 4) carry equal DFMA work
template <int NG>
static BlockMap make_block_map() {
  using C = S7<NG>;
  BlockMap bm{};
  int gi[32], gj[32], nb = 0;
  for (int i = 0; i < NG; ++i) for (int j = i + 1; j < NG; ++j) { gi[nb] = i; gj[nb] = j; ++nb; }   // 36 FMAs per row
  for (int i = 0; i < NG; ++i) { gi[nb] = i; gj[nb] = i; ++nb; }                                       // 21 FMAs per row
  int load[4] = {0, 0, 0, 0}, used[32] = {0};
  for (int b = 0; b < nb; ++b) {
    int best = -1;
    for (int w = 0; w < C::NBLK; ++w) {
      if (used[w]) continue;
      if (best < 0 || load[w % 4] < load[best % 4]) best = w;
    }
    used[best] = 1;
    load[best % 4] += (gi[b] == gj[b]) ? 21 : 36;
    bm.gi[best] = (unsigned char)gi[b];
    bm.gj[best] = (unsigned char)gj[b];
  }
  return bm;
}

template <int NG>
static void launch_sweep_tma(const DeviceData &dd, int sm_count, cudaStream_t s) {
  using C = S7<NG>;
  static BlockMap bm = make_block_map<NG>();
  const size_t smem = (size_t)C::S * s7_stage_doubles(C::R, C::NC, dd.tile_maxnz) * sizeof(double);
  static size_t configured = 0;
  if (smem > configured) {
    cudaFuncSetAttribute(sweep_tma_kernel<NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  const int ntiles = (dd.n + C::R - 1) / C::R;
  const int grid = ntiles < sm_count ? ntiles : sm_count;
  sweep_tma_kernel<NG><<<grid, C::T, smem, s>>>(dd, ntiles, dd.tile_rp, dd.tile_maxnz, bm);
}

template <int NB>
static void launch_sweep_nb(const DeviceData &dd, int sm_count, cudaStream_t s) {
  using Cfg = SweepCfg<NB>;
  static int occ = 0;
  if (occ == 0) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sweep_gram_kernel<NB>, Cfg::T, 0);
    if (occ < 1) occ = 1;
    if (occ > 4) occ = 4;
  }
  const int ntiles = (dd.n + Cfg::R - 1) / Cfg::R;
  int grid = sm_count * occ;
  if (grid > ntiles) grid = ntiles;
  sweep_gram_kernel<NB><<<grid, Cfg::T, 0, s>>>(dd, ntiles);
}

void launch_sweep(const DeviceData &dd, int sm_count, cudaStream_t s) {
  ++g_sweep_launches;
  if (!sweep_fast_supported(dd)) {   // wide designs (d > 88): refresh the lags through W, then one warp per entry
    if (dd.model != 0) launch_spmm(dd.n, dd.rp, dd.ci, dd.va, dd.y, 1, 1, dd.Wy, 1, 0, 1.0, nullptr, 0.0, s);
    if (dd.model == 2) launch_spmm(dd.n, dd.rp, dd.ci, dd.va, dd.X, dd.m, dd.m, dd.WX, dd.m, 0, 1.0, nullptr, 0.0, s);
    gram_generic_kernel<<<296, 256, 0, s>>>(dd);
    return;
  }
  if (sweep_tma_supported(dd)) {
    if (dd.d <= 6) launch_sweep_tma<1>(dd, sm_count, s);
    else if (dd.d <= 12) launch_sweep_tma<2>(dd, sm_count, s);
    else if (dd.d <= 18) launch_sweep_tma<3>(dd, sm_count, s);
    else launch_sweep_tma<4>(dd, sm_count, s);
    return;
  }
  const int nb = (dd.d + 3) / 4;
  if (nb <= 2) launch_sweep_nb<2>(dd, sm_count, s);
  else if (nb <= 3) launch_sweep_nb<3>(dd, sm_count, s);
  else if (nb <= 4) launch_sweep_nb<4>(dd, sm_count, s);
  else if (nb <= 6) launch_sweep_nb<6>(dd, sm_count, s);
  else if (nb <= 8) launch_sweep_nb<8>(dd, sm_count, s);
  else if (nb <= 11) launch_sweep_nb<11>(dd, sm_count, s);
  else if (nb <= 16) launch_sweep_nb<16>(dd, sm_count, s);
  else launch_sweep_nb<22>(dd, sm_count, s);
}

// ---------------------------------------------------------------------------
// K2 direct form: one warp per row, lanes over regressors, shuffle reduction per
// row, block reduction per CTA, fixed-order combine over CTAs.
// ---------------------------------------------------------------------------
constexpr int RSS_EC = 4;          // evaluations per pass
constexpr int RSS_GRID = 592;      // CTAs (4 per SM)

__global__ void __launch_bounds__(256) residual_ss_kernel(DeviceData dd, int e0, int n_eval,
                                                          const double *__restrict__ beta,
                                                          const double *__restrict__ lam,
                                                          const double *__restrict__ rho, double *partial) {
  extern __shared__ double sb[];            // beta chunk [RSS_EC][m]
  __shared__ double red[8][RSS_EC];
  const int m = dd.m, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const double *__restrict__ Wy = dd.Wy;
  const double *__restrict__ WX = dd.model == 2 ? dd.WX : nullptr;
  const int ne = min(RSS_EC, n_eval - e0);
  for (int i = threadIdx.x; i < ne * m; i += blockDim.x) sb[i] = beta[(size_t)e0 * m + i];
  double l[RSS_EC], r[RSS_EC], acc[RSS_EC];
#pragma unroll
  for (int e = 0; e < RSS_EC; ++e) {
    l[e] = e < ne ? lam[e0 + e] : 0.0;
    r[e] = e < ne ? rho[e0 + e] : 0.0;
    acc[e] = 0.0;
  }
  __syncthreads();
  const int warps = gridDim.x * 8;
  for (int row = blockIdx.x * 8 + wid; row < dd.n; row += warps) {
    double part[RSS_EC];
#pragma unroll
    for (int e = 0; e < RSS_EC; ++e) part[e] = 0.0;
    for (int a = lane; a < m; a += 32) {
      const double x = dd.X[(size_t)row * m + a];
      const double wx = WX ? WX[(size_t)row * m + a] : 0.0;
#pragma unroll
      for (int e = 0; e < RSS_EC; ++e)
        if (e < ne) part[e] = fma(x - r[e] * wx, sb[e * m + a], part[e]);
    }
    const double yv = dd.y[row], wy = Wy ? Wy[row] : 0.0;
#pragma unroll
    for (int e = 0; e < RSS_EC; ++e) {
      const double res = (yv - l[e] * wy) - warp_sum(part[e]);
      acc[e] = fma(res, res, acc[e]);
    }
  }
  if (lane == 0)
#pragma unroll
    for (int e = 0; e < RSS_EC; ++e) red[wid][e] = acc[e];
  __syncthreads();
  if (threadIdx.x < RSS_EC) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    partial[(size_t)blockIdx.x * RSS_EC + threadIdx.x] = s;
  }
}

__global__ void combine_partials_kernel(const double *partial, int nparts, int stride, int count, double *out) {
  const int e = threadIdx.x;
  if (e < count) {
    double s = 0.0;
    for (int b = 0; b < nparts; ++b) s += partial[(size_t)b * stride + e];
    out[e] = s;
  }
}

void launch_residual_ss(const DeviceData &dd, int n_eval, const double *beta,
                        const double *lam, const double *rho, double *partial, double *out, cudaStream_t s) {
  for (int e0 = 0; e0 < n_eval; e0 += RSS_EC) {
    residual_ss_kernel<<<RSS_GRID, 256, RSS_EC * dd.m * sizeof(double), s>>>(dd, e0, n_eval, beta, lam, rho, partial);
    const int ne = n_eval - e0 < RSS_EC ? n_eval - e0 : RSS_EC;
    combine_partials_kernel<<<1, 32, 0, s>>>(partial, RSS_GRID, RSS_EC, ne, out + e0);
  }
}

// ---------------------------------------------------------------------------
// K4 helpers: Rademacher probes (same bits as oracle/core.py:rademacher_probes) and dots
// ---------------------------------------------------------------------------
__global__ void probes_kernel(double *U, int n, int n_probes, uint64_t seed) {
  const int nb = (n_probes + 127) / 128;
  const Philox ph(seed, 0xffffffffu);
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (size_t)n * nb; idx += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(idx / nb), b = (int)(idx % nb);
    const uint4 r = ph.raw((uint32_t)row, SLOT_PROBE, (uint32_t)b, 0xffffffffu);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    for (int k = 0; k < 4; ++k)
      for (int j = 0; j < 32; ++j) {
        const int p = 128 * b + 32 * k + j;
        if (p < n_probes) U[(size_t)row * n_probes + p] = 1.0 - 2.0 * (double)((w[k] >> j) & 1u);
      }
  }
}

void launch_probes(double *U, int n, int n_probes, uint64_t seed, cudaStream_t s) {
  probes_kernel<<<592, 256, 0, s>>>(U, n, n_probes, seed);
}

constexpr int DOT_GRID = 592;
__global__ void __launch_bounds__(256) dot_kernel(const double *__restrict__ a, const double *__restrict__ b, int64_t len,
                                                  double *partial) {
  __shared__ double red[8];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x)
    acc = fma(a[i], b[i], acc);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    partial[blockIdx.x] = s;
  }
}

void launch_dot(const double *a, const double *b, int64_t len, double *partial, double *out, cudaStream_t s) {
  dot_kernel<<<DOT_GRID, 256, 0, s>>>(a, b, len, partial);
  combine_partials_kernel<<<1, 32, 0, s>>>(partial, DOT_GRID, 1, 1, out);
}

}  // namespace bsreg
