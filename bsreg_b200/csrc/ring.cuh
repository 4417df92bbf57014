// Shared pieces of the TMA-fed ring sweep (sweep.cu: one launch per sweep; persist.cu: persistent Gibbs kernel).
#pragma once
#include <cstdlib>

#include "internal.cuh"

namespace bsreg {

struct BlockMap { unsigned char gi[32], gj[32]; };

// deal the blocks of G to math warps so that the four SM sub-partitions (warp % 4) carry equal DFMA work
template <int NG>
constexpr BlockMap make_block_map() {
  constexpr int NBLK = NG * (NG + 1) / 2;
  BlockMap bm{};
  int gi[32] = {}, gj[32] = {}, nb = 0;
  for (int i = 0; i < NG; ++i) for (int j = i + 1; j < NG; ++j) { gi[nb] = i; gj[nb] = j; ++nb; }   // 36 FMAs per row
  for (int i = 0; i < NG; ++i) { gi[nb] = i; gj[nb] = i; ++nb; }                                       // 21 FMAs per row
  int load[4] = {0, 0, 0, 0}, used[32] = {};
  for (int b = 0; b < nb; ++b) {
    int best = -1;
    for (int w = 0; w < NBLK; ++w) {
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
template <int NG> __constant__ BlockMap c_block_map = make_block_map<NG>();

template <int NG>
struct Ring {
  static constexpr int NBLK = NG * (NG + 1) / 2;
  static constexpr int WPB = (NBLK >= 8) ? 1 : 2;                    // math warps per block of G (row split)
  static constexpr int MATH = NBLK * WPB;
  static constexpr int MATHW = (MATH + 3) / 4 * 4;                    // roles are aligned to warpgroups (setmaxnreg)
  static constexpr int GATHER = 7;                                    // gather warps = tiles whose lag is in flight
  static constexpr int T = (MATHW + GATHER + 1) * 32;                // gather warps + the producer warp = 2 warpgroups
  // setmaxnreg moves registers inside the CTA's own allocation (T x the launch-bound count, 96 here):
  // MATHW * 32 * REG_MATH + 8 * 32 * REG_LIGHT must not exceed it
  static constexpr int REG_MATH = 120, REG_LIGHT = 56;
  static_assert(MATHW * 32 * REG_MATH + 8 * 32 * REG_LIGHT <= T * 96, "register pool");
  static constexpr int R = SWEEP_TILE_ROWS;                          // rows per tile
  static constexpr int NC = 6 * NG;                                   // tile columns incl. zero padding
  static constexpr int NRED = NBLK * 36;                              // entries reduced in the tail
  static constexpr size_t RED_BYTES = (size_t)NRED * (WPB * 32 + 1) * 8;
};
constexpr int RING_MAX_STAGES = 12;
constexpr size_t RING_SMEM_BUDGET = 200 * 1024;

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


// FP64 tensor-core form of the Gram update (NG = 4, i.e. 24 tile columns = 3 groups of 8): D (8 x 8) += A (8 x 4) B (4 x 8)
// with A[g][t] = Z[r0 + t][8 I + g], B[t][g] = Z[r0 + t][8 J + g] -- the same register serves as A fragment of column group I
// and as B fragment of column group I, so a step of 4 rows costs 3 LDS.64 and 6 mma (the upper-triangular block pairs)
// instead of 96 LDS.64 per row from ten warps.  Lane l holds D[l >> 2][2 (l & 3)], D[l >> 2][2 (l & 3) + 1].
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
constexpr int MMA_BLOCKS = 6, MMA_WARPS = 4;                          // block pairs (0,0) (0,1) (0,2) (1,1) (1,2) (2,2)
__host__ __device__ constexpr int mma_bi(int b) { return b < 3 ? 0 : (b < 5 ? 1 : 2); }
__host__ __device__ constexpr int mma_bj(int b) { return b < 3 ? b : (b < 5 ? b - 2 : 2); }
// one tile on the tensor cores: warp w of MMA_WARPS takes the 4-row steps w, w + 4, ... (row swizzle: tile_row_pos)
__device__ __forceinline__ void mma_tile(double (&acc)[MMA_BLOCKS][2], const double *Z, const int (&fo)[3], int wid) {
  constexpr int R = SWEEP_TILE_ROWS;
#pragma unroll
  for (int ks = 0; ks < R / (4 * MMA_WARPS); ++ks) {
    const int r0 = 4 * (wid + MMA_WARPS * ks);
    double f[3];
#pragma unroll
    for (int cg = 0; cg < 3; ++cg) f[cg] = Z[fo[cg] ^ r0];            // row r0 + t at position (r0 | t) ^ sw = (t ^ sw) ^ r0
    dmma884(acc[0], f[0], f[0]); dmma884(acc[1], f[0], f[1]); dmma884(acc[2], f[0], f[2]);
    dmma884(acc[3], f[1], f[1]); dmma884(acc[4], f[1], f[2]); dmma884(acc[5], f[2], f[2]);
  }
}
// offsets of a lane's fragment element in the three column groups
__device__ __forceinline__ void mma_frag_offsets(int (&fo)[3], int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int cg = 0; cg < 3; ++cg) fo[cg] = (8 * cg + g) * SWEEP_TILE_ROWS + (t ^ (((8 * cg + g) & 3) << 2));
}

// Wy of one staged tile, executed by one gather warp.  The tile's blob lists the DISTINCT columns its rows reference
// (UL, ascending ids of the permuted ordering) and the nonzeros index that list (CI), so a y value several rows share is
// fetched once: with observations in a locality-preserving order a 64-row kNN tile (640 nonzeros) has about 200 of them
// in about 70 sectors.  cp.async (LDGSTS) copies them straight into the stage's Yg area -- no destination registers, so
// every gather of the tile is in flight at once -- then lane = rows lane, lane + 32 sums each row in CSR order with FMAs
// (the order of K1), reading Yg through shared memory.  Blob layouts: tile_blob_bytes (internal.cuh).
__device__ __forceinline__ void ring_gather_tile(double *Z, const unsigned char *C, double *Yg,
                                                 const double *__restrict__ yg, int m, int lane) {
  constexpr int R = SWEEP_TILE_ROWS, GU = 10;
  const int *RP = reinterpret_cast<const int *>(C);
  const int nz = RP[R], nu = RP[R + 1], wt = RP[R + 2];
  const double *VA = reinterpret_cast<const double *>(C + (R + 4) * 4);
  const int *CI = reinterpret_cast<const int *>(VA + (wt > 0 ? wt * R : ((nz + 1) & ~1)));
  const int *UL = CI + (wt > 0 ? wt * R : ((nz + 3) & ~3));
  for (int t = lane; t < nu; t += 128) {
    int u[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) u[q] = t + 32 * q < nu ? UL[t + 32 * q] : 0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (t + 32 * q < nu)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_u32(Yg + t + 32 * q)), "l"(yg + u[q]) : "memory");
  }
  if (wt > 0) {
    // slot-major layout: slot u of rows lane, lane + 32 sits at [u][lane], [u][lane + 32] -- consecutive lanes, consecutive
    // addresses; rows shorter than wt end in (0.0, 0) pairs, which add an exact zero
    int cl[R / 32][GU];
#pragma unroll
    for (int hh = 0; hh < R / 32; ++hh)
#pragma unroll
      for (int u = 0; u < GU; ++u) cl[hh][u] = u < wt ? CI[u * R + lane + 32 * hh] : 0;
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int hh = 0; hh < R / 32; ++hh) {
      double a = 0.0;
#pragma unroll
      for (int u = 0; u < GU; ++u) if (u < wt) a = fma(VA[u * R + lane + 32 * hh], Yg[cl[hh][u]], a);
      for (int u = GU; u < wt; ++u) a = fma(VA[u * R + lane + 32 * hh], Yg[CI[u * R + lane + 32 * hh]], a);
      Z[(m + 1) * R + tile_row_pos(m + 1, lane + 32 * hh)] = a;
    }
  } else {
    int pb[R / 32], pe[R / 32];
#pragma unroll
    for (int hh = 0; hh < R / 32; ++hh) { pb[hh] = RP[lane + 32 * hh]; pe[hh] = RP[lane + 32 * hh + 1]; }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int hh = 0; hh < R / 32; ++hh) {
      double a = 0.0;
      for (int p = pb[hh]; p < pe[hh]; ++p) a = fma(VA[p], Yg[CI[p]], a);
      Z[(m + 1) * R + tile_row_pos(m + 1, lane + 32 * hh)] = a;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic writes before the stage is refilled
}

// stage = tile columns | CSR blob | gathered y values (one per distinct column of the tile)
inline int ring_yg_offset(int d, int max_blob) { return 6 * ((d + 5) / 6) * SWEEP_TILE_ROWS * 8 + ((max_blob + 15) & ~15); }
inline int sweep_ring_stage_bytes(int d, int max_blob, int max_nu) {
  if (d > 24) return 0;
  return ring_yg_offset(d, max_blob) + ((max_nu * 8 + 15) & ~15);
}
inline int ring_stages(int d, int max_blob, int max_nu) {
  const int sb = sweep_ring_stage_bytes(d, max_blob, max_nu);
  if (sb == 0) return 0;
  int S = (int)(RING_SMEM_BUDGET / sb);
  if (S > RING_MAX_STAGES) S = RING_MAX_STAGES;
  static const int cap = getenv("BSREG_RING_STAGES") ? atoi(getenv("BSREG_RING_STAGES")) : 0;   // tuning knob
  if (cap >= 3 && S > cap) S = cap;
  return S >= 3 ? S : 0;
}
inline bool ring_supported_impl(int d, int model, int max_blob, int max_nu) { return model != 2 && ring_stages(d, max_blob, max_nu) > 0; }

}  // namespace bsreg
