// K1 + K2 fused sweep for SEM (Gram matrix of Z = [y | Wy | X | WX], d = 2M + 2 <= 48): TMA-fed ring, lags formed in shared
// memory from de-duplicated row gathers, Gram update on the FP64 tensor cores.
//
// Replaces, per Gibbs iteration, `W %*% y`, `W %*% X` (R/15_sem.R:63-64) and the cross products of the filtered data
// (R/15_sem.R:28-36, :116-120): every one of them is a polynomial in lambda of blocks of G.
//
// Why a kernel of its own (round 1 ran SEM through sweep_gram_kernel: W.X gathered 160 B of X per nonzero straight from
// L2, 842 us per sweep at N = 1e6, 0.054 of the HBM roofline):
//   * the rows of X a tile needs are fetched ONCE per tile: the tile's own 64 rows arrive with the tile itself, the
//     FOREIGN rows it references (about 150 with the breadth-first ordering, instead of 640 nonzeros) are copied by
//     cp.async behind them into the same table, so a nonzero's column index addresses one table T[64 + nuf][ldx];
//   * W.X and W.y are formed from shared memory (lane = column, 8 rows per warp in flight);
//   * Z'Z runs as 8 x 8 x 4 FP64 mma blocks over column groups of 8 (mma.sync.m8n8k4.f64, DMMA in SASS).
//
// HBM operands (internal ordering of the observations, see api.cu):
//   Xr   [ntiles * 64][ldx]   X row-major, ldx = smallest value >= M with ldx = 4 (mod 8): 16-byte aligned rows whose
//                             fragment loads (4 columns x 4 rows per half warp) fall into 16 distinct banks
//   yp   [ntiles * 64]        y
//   blob per tile             as the ring sweep's (internal.cuh), column indices relative to T: < 64 = own row, else foreign
// One tile = three bulk copies (cp.async.bulk, complete_tx on an mbarrier) + nuf row gathers.
//
// Roles (512 threads, one CTA per SM): warps 0-3 mma, warps 4-11 gather + lag formation (one tile at a time, all eight),
// warp 12 producer.  full[s] / lag[s] / empty[s] order producer -> gather -> mma -> producer.
#include <cstdlib>

#include "internal.cuh"
#include "ring.cuh"

namespace bsreg {

constexpr int SEM_T = 512, SEM_MMA_WARPS = 4, SEM_GATHER_WARPS = 8, SEM_MAX_STAGES = 3, SEM_U_STAGES = 6;
constexpr size_t SEM_SMEM_BUDGET = 226 * 1024;   // dynamic shared memory of the CTA (227 KB less the static barriers)
constexpr int SEM_ZERO = 80;                 // doubles of the per-stage zero block the padding lanes of a fragment read
constexpr unsigned SEM_POLL_NS = 96;         // pause between two polls of an mbarrier
constexpr int SEM_AHEAD = 2;                 // tiles whose foreign rows are in flight ahead of the lag formation

struct SemGeom {
  int ntiles, S, stage_doubles;
  int ldx, nuf_cap, GA;                      // GA = columns of the [X | y | Wy] part padded to a multiple of 8
  int off_yT, off_WX, off_Wy, off_zero, off_blob;   // doubles from the start of a stage
  int off_u, u_ints;                         // ring of foreign-row lists behind the stages: offset (doubles), ints per list
  int lag_rows;                              // lag formation: 1 = quarter warp per row (16-byte loads), 0 = lane per column
  int dbg;                                   // profiling only (BSREG_SEM_DBG): 1 = no mma, 2 = no lag formation, 4 = no foreign-row gathers
};

#ifdef BSREG_TRACE
// clock64 stamps of tiles 40 .. 47 of CTA 0: slots 0-6 gather warp 0, 8-10 mma warp 0, 12-13 producer (profiles/trace_sem.py)
static __device__ long long g_sem_trace[8 * 16];
#define STRACE(j, k) do { if (blockIdx.x == 0 && lane == 0 && (j) >= 40 && (j) < 48) g_sem_trace[((j) - 40) * 16 + (k)] = clock64(); } while (0)
extern "C" int bsreg_debug_trace_sem(long long *out) { return (int)cudaMemcpyFromSymbol(out, g_sem_trace, sizeof(long long) * 128); }
#else
#define STRACE(j, k) do { } while (0)
#endif

__host__ __device__ inline int sem_ldx(int m) { int l = 4; while (l < m) l += 8; return l; }
__device__ __forceinline__ void named_sync_sem(int id, int count) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(count) : "memory"); }
// Polling an mbarrier is a shared-memory operation, and shared-memory bandwidth is what bounds this kernel: a waiter
// that is not served at once sleeps between polls instead of spinning (a tile takes microseconds).
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *b, uint32_t parity) {
  for (;;) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    if (ok) break;
    __nanosleep(SEM_POLL_NS);
  }
}
// one lane polls, the others look once the phase has completed (every thread still observes the barrier itself)
__device__ __forceinline__ void mbar_wait_warp(uint64_t *b, uint32_t parity, int lane) {
  if (lane == 0) mbar_wait_sleep(b, parity);
  __syncwarp();
  mbar_wait(b, parity);
}

// slot of the 8-column groups -> Gram column, -1 = padding
__host__ __device__ inline int sem_slot_gram(int slot, int m, int GA) {
  if (slot < m) return 2 + slot;             // X
  if (slot == m) return 0;                   // y
  if (slot == m + 1) return 1;               // Wy
  if (slot < GA) return -1;
  const int c = slot - GA;
  return c < m ? 2 + m + c : -1;             // WX
}

// Stage s (three of them): T = [own rows 64 x ldx (bulk copy) | foreign rows nuf x ldx (cp.async)], yT alike, WX, Wy, zeros, blob.
// The foreign part of a stage is written two tiles ahead of its use, while the mma warps may still be reading the own rows,
// WX and Wy of the stage's previous tile: the two regions have separate lifetimes.  The lists of foreign rows travel through
// a small ring of their own (six deep), so that a tile's gathers do not wait for the tile's bulk copies.
template <int NGR>
__global__ void __launch_bounds__(SEM_T, 1) sweep_sem_kernel(DeviceData dd, SemGeom ge) {
  constexpr int R = SWEEP_TILE_ROWS, NP = NGR * (NGR + 1) / 2;
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ __align__(8) uint64_t full[SEM_MAX_STAGES], lag[SEM_MAX_STAGES], empty[SEM_MAX_STAGES];
  __shared__ __align__(8) uint64_t ufull[SEM_U_STAGES], uempty[SEM_U_STAGES];
  double *sm = reinterpret_cast<double *>(dyn);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int m = dd.m, d = dd.d, S = ge.S, ldx = ge.ldx;

  if (blockIdx.x == 0 && dd.gzero >= 0) {     // the accumulator of the sweep after next
    long long *z = dd.Gfix + (size_t)dd.gzero * d * d;
    for (int e = tid; e < d * d; e += SEM_T) z[e] = 0;
  }
  for (int e = tid; e < S * SEM_ZERO; e += SEM_T) sm[(size_t)(e / SEM_ZERO) * ge.stage_doubles + ge.off_zero + e % SEM_ZERO] = 0.0;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&lag[s], 1); mbar_init(&empty[s], SEM_MMA_WARPS); }
    for (int s = 0; s < SEM_U_STAGES; ++s) { mbar_init(&ufull[s], 1); mbar_init(&uempty[s], SEM_GATHER_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");

  // tiles dealt round-robin: at any time the CTAs work on ~gridDim.x CONSECUTIVE tiles, whose foreign rows (within a few
  // thousand rows with the breadth-first ordering) are the own rows of tiles streamed a moment ago or about to be streamed
  // -- they are served by L2, so every row of X leaves HBM once per sweep
  const int nb = gridDim.x, sb = blockIdx.x;
  const int my_tiles = (ge.ntiles - sb + nb - 1) / nb;
  const uint32_t x_bytes = (uint32_t)R * ldx * 8u, y_bytes = (uint32_t)R * 8u, u_bytes = (uint32_t)ge.u_ints * 4u;
  int *uring = reinterpret_cast<int *>(sm + ge.off_u);
  constexpr int LD = SEM_MMA_WARPS + 1;
  double *red = sm;                                                   // tail: [NP * 64][MMA_WARPS + 1], over the ring

  if (wid == SEM_MMA_WARPS + SEM_GATHER_WARPS) {
    // =========================== producer of the tiles ===========================
    if (lane == 0 && my_tiles > 0) {
      int s = 0;
      uint32_t ph = 1;                                                // fresh barriers: the "previous" phase is complete
      long long o0 = dd.tile_off[sb];
      int o1 = dd.tile_len[sb];
      for (int j = 0; j < my_tiles; ++j) {
        const int tile = sb + j * nb;
        long long n0 = 0;
        int n1 = 0;
        if (j + 1 < my_tiles) { n0 = dd.tile_off[tile + nb]; n1 = dd.tile_len[tile + nb]; }
        STRACE(j, 12);
        mbar_wait_sleep(&empty[s], ph);
        STRACE(j, 13);
        double *st = sm + (size_t)s * ge.stage_doubles;
        mbar_expect_tx(&full[s], x_bytes + y_bytes + (uint32_t)o1);
        tma_bulk_g2s(st, dd.Xr + (size_t)tile * R * ldx, x_bytes, &full[s]);
        tma_bulk_g2s(st + ge.off_yT, dd.yp + (size_t)tile * R, y_bytes, &full[s]);
        tma_bulk_g2s(st + ge.off_blob, dd.csr_blob + o0, (uint32_t)o1, &full[s]);
        o0 = n0; o1 = n1;
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
  } else if (wid == SEM_MMA_WARPS + SEM_GATHER_WARPS + 1) {
    // =========================== producer of the foreign-row lists ===========================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;
      for (int j = 0; j < my_tiles; ++j) {
        const int tile = sb + j * nb;
        mbar_wait_sleep(&uempty[s], ph);
        mbar_expect_tx(&ufull[s], u_bytes);
        tma_bulk_g2s(uring + (size_t)s * ge.u_ints, dd.tile_ul + (size_t)tile * ge.u_ints, u_bytes, &ufull[s]);
        if (++s == SEM_U_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (wid >= SEM_MMA_WARPS && wid < SEM_MMA_WARPS + SEM_GATHER_WARPS) {
    // =========================== gather: foreign rows -> T, then Wy and WX of the tile ===========================
    constexpr int GT = SEM_GATHER_WARPS * 32;
    const int gt = tid - SEM_MMA_WARPS * 32, gw = gt >> 5;
    const int H = ldx >> 1;                                           // 16-byte chunks per row of X
    const bool isy = lane == m;
    // lane = column; lane m forms Wy; the lanes beyond repeat a column whose banks no active lane of their half warp touches
    const int col = lane < m ? lane : (lane - 16 >= 0 && lane - 16 < m ? lane - 16 : 0), mul = isy ? 1 : ldx;
    // every foreign row of a tile once: 16-byte copies straight into the table (no destination registers).  The load / store
    // unit spends ~13 cycles per cp.async warp instruction whatever the number of active lanes (clock64 trace: 150
    // instructions per tile = 2000 cycles of the gather warps), so the instructions are filled: 32 / H rows per instruction
    // (lane = (row, 16-byte chunk)), and the rows' y values 32 rows at a time (lane = row)
    const int rpi = 32 / H, gr = lane / H, q = lane - gr * H;         // rows per instruction, this lane's row and chunk
    const bool px = gr < rpi;
    const char *gx = reinterpret_cast<const char *>(dd.Xr) + 16 * q;
    const size_t row_bytes = (size_t)ldx * 8;
    int us = 0;
    uint32_t uph = 0;
    auto issue_gathers = [&](int jt) {
      if (jt < my_tiles) {
        const int s_ = jt % S;
        double *st = sm + (size_t)s_ * ge.stage_doubles;
        const uint32_t dx = smem_u32(st + (size_t)R * ldx) + 16 * q, dy = smem_u32(st + ge.off_yT + R);
        const int *UL = uring + (size_t)us * ge.u_ints;
        mbar_wait_warp(&ufull[us], uph, lane);
        const int nuf = UL[ge.u_ints - 1];
        const int nuf_ = (ge.dbg & 4) ? 0 : nuf;
        if (px)
          for (int t = rpi * gw + gr; t < nuf_; t += rpi * SEM_GATHER_WARPS)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dx + (uint32_t)(t * (int)row_bytes)), "l"(gx + (size_t)UL[t] * row_bytes) : "memory");
        for (int t = 32 * gw + lane; t < nuf_; t += 32 * SEM_GATHER_WARPS)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dy + 8u * t), "l"(dd.yp + UL[t]) : "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&uempty[us]);
        if (++us == SEM_U_STAGES) { us = 0; uph ^= 1; }
      }
      asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    for (int a = 0; a < SEM_AHEAD; ++a) issue_gathers(a);
    int s = 0;
    uint32_t ph = 0;
    for (int j = 0; j < my_tiles; ++j) {
      double *st = sm + (size_t)s * ge.stage_doubles;
      double *T = st, *yT = st + ge.off_yT, *WXs = st + ge.off_WX, *Wys = st + ge.off_Wy;
      const unsigned char *C = reinterpret_cast<const unsigned char *>(st + ge.off_blob);
      if (gw == 0) STRACE(j, 0);
      issue_gathers(j + SEM_AHEAD);
      if (gw == 0) STRACE(j, 1);
      asm volatile("cp.async.wait_group %0;\n" ::"n"(SEM_AHEAD) : "memory");
      if (gw == 0) STRACE(j, 2);
      mbar_wait_warp(&full[s], ph, lane);
      if (gw == 0) STRACE(j, 3);
      named_sync_sem(1, GT);
      if (gw == 0) STRACE(j, 4);
      const int *RP = reinterpret_cast<const int *>(C);
      const int nz = RP[R], wt = RP[R + 2];
      const double *VA = reinterpret_cast<const double *>(C + (R + 4) * 4);
      const int *CI = reinterpret_cast<const int *>(VA + (wt > 0 ? wt * R : ((nz + 1) & ~1)));
      if (ge.dbg & 2) {
      } else if (ge.lag_rows) {
        // quarter warp = one row of the tile, lane l of the quarter = 16-byte chunk l of a gathered row: a nonzero costs one
        // LDS.128 per quarter (8 lanes x 16 B = one conflict-free wavefront whatever the row's address), one 8-byte load
        // for the columns beyond the first 16 (lanes < tail) and for y (lane tail), and the broadcast of its value and table
        // index to the quarter's lanes (~1 wavefront each, profiles/r2_micro_smem_bcast.log) -- 2.2 wavefronts per nonzero
        // against 4.4 of the lane = column form below
        const int l = lane & 7, qtr = lane >> 3;
        const int tail = ldx - 16;                                      // doubles of a row beyond its first 128 bytes
        const bool a_on = 2 * l < ldx;
        const bool t64 = tail > 0 && tail < 8;                          // tail by 8-byte loads, y rides on lane `tail`
        const bool b_on = !t64 && tail > 0 && 2 * l < tail;             // long tail (ldx = 28): 16-byte loads, y by lane 7
        const int ylane = t64 ? tail : 7;
#pragma unroll 1
        for (int rd = 0; rd < R / (4 * SEM_GATHER_WARPS); ++rd) {
          const int i = rd * 4 * SEM_GATHER_WARPS + 4 * gw + qtr;
          double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0, tv = 0.0;
          auto body = [&](double v, int c) {
            const double *row = T + c * ldx;
            if (a_on) {
              const double2 x = *reinterpret_cast<const double2 *>(row + 2 * l);
              a0 = fma(v, x.x, a0); a1 = fma(v, x.y, a1);
            }
            if (b_on) {
              const double2 x = *reinterpret_cast<const double2 *>(row + 16 + 2 * l);
              b0 = fma(v, x.x, b0); b1 = fma(v, x.y, b1);
            }
            if (t64 ? l <= tail : l == 7) tv = fma(v, (t64 && l < tail) ? row[16 + l] : yT[c], tv);
          };
          if (wt > 0) {
#pragma unroll 5
            for (int u = 0; u < wt; ++u) body(VA[u * R + i], CI[u * R + i]);
          } else {
            const int p0 = RP[i], len = RP[i + 1] - p0;
            // (a vote per step, not redux.sync for the longest row: ncu could not prepare the kernel with REDUX in it, and
            // a shuffle reduction here doubled the time of the whole kernel, 257 -> 490 us)
            for (int u = 0; __any_sync(0xffffffffu, u < len); ++u) {
              const bool on = u < len;
              body(on ? VA[p0 + u] : 0.0, on ? CI[p0 + u] : 0);
            }
          }
          if (a_on) *reinterpret_cast<double2 *>(WXs + i * ldx + 2 * l) = make_double2(a0, a1);
          if (b_on) *reinterpret_cast<double2 *>(WXs + i * ldx + 16 + 2 * l) = make_double2(b0, b1);
          if (t64 && l < tail) WXs[i * ldx + 16 + l] = tv;
          if (l == ylane) Wys[i] = tv;
        }
      } else {
      const int r0 = 8 * gw;                                          // this warp's rows
      const double *src = (isy ? yT : T + col);
      double a8[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) a8[k] = 0.0;
      if (wt > 0) {
        // slot-major layout: the values / columns of slot u of rows r0 .. r0 + 7 are 64 / 32 contiguous bytes (broadcast loads)
#pragma unroll 2
        for (int u = 0; u < wt; ++u) {
          const double2 *vp = reinterpret_cast<const double2 *>(VA + u * R + r0);
          const int4 *cp = reinterpret_cast<const int4 *>(CI + u * R + r0);
          const double2 v01 = vp[0], v23 = vp[1], v45 = vp[2], v67 = vp[3];
          const int4 c03 = cp[0], c47 = cp[1];
          const double x0 = src[c03.x * mul], x1 = src[c03.y * mul], x2 = src[c03.z * mul], x3 = src[c03.w * mul];
          const double x4 = src[c47.x * mul], x5 = src[c47.y * mul], x6 = src[c47.z * mul], x7 = src[c47.w * mul];
          a8[0] = fma(v01.x, x0, a8[0]); a8[1] = fma(v01.y, x1, a8[1]); a8[2] = fma(v23.x, x2, a8[2]); a8[3] = fma(v23.y, x3, a8[3]);
          a8[4] = fma(v45.x, x4, a8[4]); a8[5] = fma(v45.y, x5, a8[5]); a8[6] = fma(v67.x, x6, a8[6]); a8[7] = fma(v67.y, x7, a8[7]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          double a = 0.0;
          for (int p = RP[r0 + k], pe = RP[r0 + k + 1]; p < pe; ++p) a = fma(VA[p], src[CI[p] * mul], a);
          a8[k] = a;
        }
      }
      if (lane < m) {
#pragma unroll
        for (int k = 0; k < 8; ++k) WXs[(r0 + k) * ldx + lane] = a8[k];
      } else if (isy) {
#pragma unroll
        for (int k = 0; k < 8; ++k) Wys[r0 + k] = a8[k];
      }
      }
      if (gw == 0) STRACE(j, 5);
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic accesses before the stage is refilled
      named_sync_sem(1, GT);
      if (gw == 0) STRACE(j, 6);
      if (gt == 0) mbar_arrive(&lag[s]);
      if (++s == S) { s = 0; ph ^= 1; }
    }
  } else if (wid < SEM_MMA_WARPS) {
    // =========================== mma: warp w takes the 4-row steps w, w + 4, w + 8, w + 12 of every tile ===========================
    const int g = lane >> 2, t = lane & 3;
    double acc[NP][2];
#pragma unroll
    for (int k = 0; k < NP; ++k) { acc[k][0] = 0.0; acc[k][1] = 0.0; }
    int off[NGR], str[NGR];
#pragma unroll
    for (int cg = 0; cg < NGR; ++cg) {
      const int slot = 8 * cg + g;
      int o = ge.off_zero + 4 * (g & 3) + t, sd = 1;                  // padding: zeros, one bank phase per column
      if (slot < m) { o = t * ldx + slot; sd = ldx; }
      else if (slot == m) { o = ge.off_yT + t; }
      else if (slot == m + 1) { o = ge.off_Wy + t; }
      else if (slot >= ge.GA && slot - ge.GA < m) { o = ge.off_WX + t * ldx + slot - ge.GA; sd = ldx; }
      off[cg] = o; str[cg] = sd;
    }
    int s = 0;
    uint32_t ph = 0;
    for (int j = 0; j < my_tiles; ++j) {
      if (wid == 0) STRACE(j, 8);
      mbar_wait_warp(&lag[s], ph, lane);
      if (wid == 0) STRACE(j, 9);
      const double *Z = sm + (size_t)s * ge.stage_doubles;
#pragma unroll
      for (int ks = 0; ks < ((ge.dbg & 1) ? 0 : R / (4 * SEM_MMA_WARPS)); ++ks) {
        const int r0 = 4 * (wid + SEM_MMA_WARPS * ks);
        double f[NGR];
#pragma unroll
        for (int cg = 0; cg < NGR; ++cg) f[cg] = Z[off[cg] + r0 * str[cg]];
        int p = 0;
#pragma unroll
        for (int I = 0; I < NGR; ++I)
#pragma unroll
          for (int J = I; J < NGR; ++J) { dmma884(acc[p], f[I], f[J]); ++p; }
      }
      if (wid == 0) STRACE(j, 10);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      if (++s == S) { s = 0; ph ^= 1; }
    }
    // ---- tail: once every mma warp has consumed its last stage the ring is dead: partial blocks -> shared memory ----
    named_sync_sem(2, SEM_MMA_WARPS * 32);
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      red[(k * 64 + g * 8 + 2 * t) * LD + wid] = acc[k][0];
      red[(k * 64 + g * 8 + 2 * t + 1) * LD + wid] = acc[k][1];
    }
  }
  // ---- one thread per entry of G adds the four partials in fixed order and issues ONE fixed-point atomic ----
  __syncthreads();
  long long *Gacc = dd.Gfix + (size_t)dd.gcur * d * d;
  for (int e = tid; e < NP * 64; e += SEM_T) {
    const int k = e >> 6, pp = (e >> 3) & 7, qq = e & 7;
    int I = 0, rem = k;
    while (rem >= NGR - I) { rem -= NGR - I; ++I; }
    const int J = I + rem;
    if (I == J && qq < pp) continue;                                  // lower triangle of a diagonal block
    const int ga = sem_slot_gram(8 * I + pp, m, ge.GA), gb = sem_slot_gram(8 * J + qq, m, ge.GA);
    if (ga < 0 || gb < 0) continue;
    const int gidx = min(ga, gb) * d + max(ga, gb);
    const double *v = red + e * LD;
    const double sum = (v[0] + v[1]) + (v[2] + v[3]);
    const long long f = __double2ll_rn(sum * dd.Ginv_scale[gidx]);
    atomicAdd(reinterpret_cast<unsigned long long *>(Gacc + gidx), (unsigned long long)f);
  }
}

// The lists of foreign rows as fixed-size records [ntiles][u_ints] (last int = count) for the list ring, cut out of the
// packed blobs; tile_len shrinks to the part of the blob the tile copy still needs (header, values, column slots).
__global__ void extract_ul_kernel(const unsigned char *__restrict__ blob, const long long *__restrict__ tile_off,
                                  int *__restrict__ tile_len, int *__restrict__ max_len, int *__restrict__ ul, int u_ints) {
  constexpr int R = SWEEP_TILE_ROWS;
  const int t = blockIdx.x;
  const int *hdr = reinterpret_cast<const int *>(blob + tile_off[t]);
  const int nz = hdr[R], nuf = hdr[R + 1], wt = hdr[R + 2];
  const int body = wt > 0 ? wt * R * 12 : ((nz + 1) & ~1) * 8 + ((nz + 3) & ~3) * 4;
  const int *UL = reinterpret_cast<const int *>(blob + tile_off[t] + (R + 4) * 4 + body);
  for (int i = threadIdx.x; i < u_ints - 1; i += blockDim.x) ul[(size_t)t * u_ints + i] = i < nuf ? UL[i] : 0;
  if (threadIdx.x == 0) {
    ul[(size_t)t * u_ints + u_ints - 1] = nuf;
    const int len = ((R + 4) * 4 + body + 15) & ~15;
    tile_len[t] = len;
    atomicMax(max_len, len);
  }
}

int sem_ul_ints(int max_nuf) { return (max_nuf + 1 + 3) & ~3; }

void launch_extract_ul(const unsigned char *blob, const long long *tile_off, int *tile_len, int *max_len, int *ul, int u_ints,
                       int n, cudaStream_t s) {
  const int ntiles = (n + SWEEP_TILE_ROWS - 1) / SWEEP_TILE_ROWS;
  extract_ul_kernel<<<ntiles, 128, 0, s>>>(blob, tile_off, tile_len, max_len, ul, u_ints);
}

// X rows and y in the internal ordering, zero padded to whole tiles and to ldx columns
__global__ void pack_rows_kernel(const double *__restrict__ X, const double *__restrict__ y, const int *__restrict__ perm,
                                 double *__restrict__ Xr, double *__restrict__ yp, int n, int m, int ldx, int nrows) {
  const size_t total = (size_t)nrows * ldx;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(idx / ldx), c = (int)(idx - (size_t)row * ldx);
    double v = 0.0;
    if (row < n && c < m) v = X[(size_t)(perm ? perm[row] : row) * m + c];
    Xr[idx] = v;
    if (c == 0) yp[row] = row < n ? y[perm ? perm[row] : row] : 0.0;
  }
}

void launch_pack_rows(const double *X, const double *y, const int *perm, double *Xr, double *yp, int n, int m, int ldx,
                      cudaStream_t s) {
  const int nrows = (n + SWEEP_TILE_ROWS - 1) / SWEEP_TILE_ROWS * SWEEP_TILE_ROWS;
  pack_rows_kernel<<<592, 256, 0, s>>>(X, y, perm, Xr, yp, n, m, ldx, nrows);
}

int sem_ring_ldx(int m) { return sem_ldx(m); }

static bool sem_geometry(int n, int m, int max_blob, int max_nuf, SemGeom &ge) {
  constexpr int R = SWEEP_TILE_ROWS;
  if (m < 1 || m > 22) return false;
  auto up16 = [](int v) { return (v + 15) & ~15; };                   // sub-arrays start on 128-byte boundaries (bank phase 0)
  ge.ntiles = (n + R - 1) / R;
  ge.ldx = sem_ldx(m);
  ge.nuf_cap = max_nuf;
  ge.GA = (m + 2 + 7) / 8 * 8;
  int o = up16((R + max_nuf) * ge.ldx);
  ge.off_yT = o; o += up16(R + max_nuf);
  ge.off_WX = o; o += up16(R * ge.ldx);
  ge.off_Wy = o + 4; o += up16(R + 4);                                // bank phase 4: beside y (phase 0) in a fragment load
  ge.off_zero = o; o += SEM_ZERO;
  ge.off_blob = o; o += up16((max_blob + 7) / 8);
  ge.stage_doubles = o;
  ge.u_ints = sem_ul_ints(max_nuf);
  {
    const char *e = getenv("BSREG_SEM_LAG");                          // A/B switch: 0 = the lane-per-column form
    ge.lag_rows = !(e && *e == '0');
    const char *g = getenv("BSREG_SEM_DBG");
    ge.dbg = g ? atoi(g) : 0;
  }
  const size_t u_bytes = (size_t)SEM_U_STAGES * ge.u_ints * 4;
  int S = (int)((SEM_SMEM_BUDGET - u_bytes) / ((size_t)o * 8));
  if (S > SEM_MAX_STAGES) S = SEM_MAX_STAGES;
  ge.S = S;
  ge.off_u = S * o;
  return S >= SEM_AHEAD + 1;                                          // the foreign rows of SEM_AHEAD tiles are in flight
}

bool sem_ring_supported(int n, int m, int max_blob, int max_nuf) {
  SemGeom ge;
  return sem_geometry(n, m, max_blob, max_nuf, ge);
}

template <int NGR>
static void launch_sem_ngr(const DeviceData &dd, const SemGeom &ge, int sm_count, cudaStream_t s) {
  constexpr size_t red_bytes = (size_t)(NGR * (NGR + 1) / 2) * 64 * (SEM_MMA_WARPS + 1) * 8;
  size_t smem = (size_t)ge.S * ge.stage_doubles * 8 + (size_t)SEM_U_STAGES * ge.u_ints * 4;
  if (smem < red_bytes) smem = red_bytes;
  if (smem_optin((const void *)sweep_sem_kernel<NGR>, smem) != cudaSuccess) return;
  const int grid = ge.ntiles < sm_count ? ge.ntiles : sm_count;
  sweep_sem_kernel<NGR><<<grid, SEM_T, smem, s>>>(dd, ge);
}

void launch_sweep_sem(const DeviceData &dd, int sm_count, cudaStream_t s) {
  SemGeom ge;
  if (!sem_geometry(dd.n, dd.m, dd.tile_max_blob, dd.tile_max_nu, ge)) return;
  const int ngr = (ge.GA + (dd.m + 7) / 8 * 8) / 8;
  if (ngr <= 2) launch_sem_ngr<2>(dd, ge, sm_count, s);
  else if (ngr <= 4) launch_sem_ngr<4>(dd, ge, sm_count, s);
  else launch_sem_ngr<6>(dd, ge, sm_count, s);
}

}  // namespace bsreg
