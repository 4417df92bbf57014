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
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

#include "internal.cuh"
#include "ring.cuh"
#include "rng.cuh"

namespace bsreg {

cudaError_t smem_optin(const void *kernel, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void *>, size_t> configured;
  if (bytes <= 48 * 1024) return cudaSuccess;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(mu);
  size_t &have = configured[{dev, kernel}];
  if (bytes > have) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    have = bytes;
  }
  return cudaSuccess;
}

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

// Zt[tile][c][r] for the ring sweep: columns X_1..X_m then y, zero padded to whole tiles; tile row r is observation
// perm[r] (perm == null: the caller's order); yp = y in the same order (the gather source of the ring)
__global__ void pack_tiles_kernel(const double *__restrict__ X, const double *__restrict__ y, const int *__restrict__ perm,
                                  double *__restrict__ Zt, double *__restrict__ yp, int n, int m, int ntiles) {
  constexpr int R = SWEEP_TILE_ROWS;
  const size_t total = (size_t)ntiles * (m + 1) * R;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(idx % R);
    const size_t tc = idx / R;
    const int c = (int)(tc % (m + 1)), tile = (int)(tc / (m + 1));
    const int row = tile * R + tile_row_pos(c, r);                   // position r of column c holds this row
    double v = 0.0;
    if (row < n) {
      const int src = perm ? perm[row] : row;
      v = c < m ? X[(size_t)src * m + c] : y[src];
      if (c == m && yp != nullptr) yp[row] = v;
    }
    Zt[idx] = v;
  }
}

void launch_pack_tiles(const double *X, const double *y, const int *perm, double *Zt, double *yp, int n, int m, cudaStream_t s) {
  const int ntiles = (n + SWEEP_TILE_ROWS - 1) / SWEEP_TILE_ROWS;
  pack_tiles_kernel<<<592, 256, 0, s>>>(X, y, perm, Zt, yp, n, m, ntiles);
}

// per-tile CSR blobs of the ring sweep (layout: tile_blob_bytes in ring.cuh).  One CTA per tile: rows perm[r0..r0+R) of the
// caller's CSR, columns relabelled by inv; the entries of a row keep their order (Wy is summed as K1 sums it).  The
// distinct columns of the tile are found by a bitonic sort in shared memory; tiles with more than PACK_CAP nonzeros
// are written without de-duplication (ul = the columns as they come).
constexpr int PACK_CAP = SWEEP_PACK_CAP, PACK_T = 256;
__global__ void __launch_bounds__(PACK_T) pack_csr_tiles_kernel(const int *__restrict__ rp, const int *__restrict__ ci,
                                                                const double *__restrict__ va, const int *__restrict__ perm,
                                                                const int *__restrict__ inv, const long long *__restrict__ tile_off,
                                                                int *__restrict__ tile_len, int *__restrict__ max_len_nu,
                                                                unsigned char *__restrict__ blob, int n, int split_local) {
  constexpr int R = SWEEP_TILE_ROWS;
  __shared__ int rpl[R + 1], src0[R];
  __shared__ int keys[PACK_CAP];
  __shared__ int s_nu, s_maxlen;
  const int t = blockIdx.x, r0 = t * R, tid = threadIdx.x;
  if (tid == 0) {
    int acc = 0, mx = 0;
    for (int i = 0; i < R; ++i) {
      rpl[i] = acc;
      const int row = r0 + i;
      if (row < n) {
        const int src = perm ? perm[row] : row;
        src0[i] = rp[src];
        acc += rp[src + 1] - rp[src];
        mx = max(mx, rp[src + 1] - rp[src]);
      } else {
        src0[i] = 0;
      }
    }
    rpl[R] = acc;
    s_maxlen = mx;
  }
  __syncthreads();
  const int nz = rpl[R], wt = ell_width(nz, s_maxlen);
  unsigned char *b = blob + tile_off[t];
  int *hdr = reinterpret_cast<int *>(b);
  double *bva = reinterpret_cast<double *>(b + (R + 4) * 4);
  int *bci = reinterpret_cast<int *>(bva + (wt > 0 ? wt * R : ((nz + 1) & ~1)));
  int *bul = bci + (wt > 0 ? wt * R : ((nz + 3) & ~3));
  // position of nonzero q of row i in the blob: slot-major when the tile is regular enough, CSR otherwise
  auto slot = [&](int i, int q) { return wt > 0 ? q * R + i : rpl[i] + q; };
  if (wt > 0) for (int e = tid; e < wt * R; e += PACK_T) { bva[e] = 0.0; bci[e] = 0x7fffffff; }   // padding of short rows
  __syncthreads();
  // values, and the relabelled columns (into bci for now)
  for (int i = tid / 32; i < R; i += PACK_T / 32) {
    const int len = rpl[i + 1] - rpl[i];
    for (int q = tid % 32; q < len; q += 32) {
      const int c = ci[src0[i] + q];
      bva[slot(i, q)] = va[src0[i] + q];
      bci[slot(i, q)] = inv ? inv[c] : c;
    }
  }
  if (wt == 0) for (int q = nz + tid; q < ((nz + 1) & ~1); q += PACK_T) bva[q] = 0.0;
  __syncthreads();
  const int nslot = wt > 0 ? wt * R : nz;                    // entries of bci to translate (padding included)
  int nu = nz;
  if (nz <= PACK_CAP) {
    int np2 = 1;
    while (np2 < nz) np2 <<= 1;
    // the columns of the real nonzeros (ELL padding carries 0x7fffffff and sorts to the end)
    if (wt > 0) {
      __shared__ int s_cnt;
      if (tid == 0) s_cnt = 0;
      __syncthreads();
      for (int e = tid; e < nslot; e += PACK_T) if (bci[e] != 0x7fffffff) keys[atomicAdd(&s_cnt, 1)] = bci[e];
      __syncthreads();
      for (int q = nz + tid; q < np2; q += PACK_T) keys[q] = 0x7fffffff;
    } else {
      for (int q = tid; q < np2; q += PACK_T) keys[q] = q < nz ? bci[q] : 0x7fffffff;
    }
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int q = tid; q < np2; q += PACK_T) {
          const int l = q ^ j;
          if (l > q) {
            const int a = keys[q], c = keys[l];
            const bool up = (q & k) == 0;
            if ((a > c) == up) { keys[q] = c; keys[l] = a; }
          }
        }
        __syncthreads();
      }
    // distinct keys, compacted in place by one warp (ballot scan), ascending
    if (tid < 32) {
      int base = 0;
      for (int q0 = 0; q0 < nz; q0 += 32) {
        const int q = q0 + tid;
        const int kq = q < nz ? keys[q] : 0;
        const bool first = q < nz && (q == 0 || keys[q - 1] != kq);
        const unsigned mask = __ballot_sync(0xffffffffu, first);
        __syncwarp();
        if (first) keys[base + __popc(mask & ((1u << tid) - 1))] = kq;   // base + rank <= q: never ahead of the reads
        base += __popc(mask);
        __syncwarp();
      }
      if (tid == 0) s_nu = base;
    }
    __syncthreads();
    nu = s_nu;
    // split_local (SEM ring sweep): the tile's own rows [r0, r0 + R) are addressed directly (index = row in the tile), the
    // foreign columns follow as R + position in the list of foreign columns; they form a contiguous run of the sorted keys
    int lo_own = 0, n_own = 0;
    if (split_local) {
      int lo = 0, hi = nu;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid] < r0) lo = mid + 1; else hi = mid; }
      lo_own = lo; hi = nu;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid] < r0 + R) lo = mid + 1; else hi = mid; }
      n_own = lo - lo_own;
    }
    for (int q = tid; q < nslot; q += PACK_T) {            // column -> position in the distinct list (padding -> 0)
      const int c = bci[q];
      int lo = 0, hi = nu - 1;
      if (c == 0x7fffffff) { lo = 0; }
      else {
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid] < c) lo = mid + 1; else hi = mid; }
        if (split_local) lo = (c >= r0 && c < r0 + R) ? c - r0 : R + (lo < lo_own ? lo : lo - n_own);
      }
      bci[q] = lo;
    }
    __syncthreads();
    if (split_local) {
      nu -= n_own;
      for (int q = tid; q < ((nu + 3) & ~3); q += PACK_T) bul[q] = q < nu ? keys[q < lo_own ? q : q + n_own] : 0;
    } else {
      for (int q = tid; q < ((nu + 3) & ~3); q += PACK_T) bul[q] = q < nu ? keys[q] : 0;
    }
  } else {
    // more nonzeros than the sort holds: always the CSR layout (64 * max_len <= 1.25 nz + 64 cannot hold with nz > PACK_CAP
    // unless the rows are long and regular; such tiles do not fit a ring stage anyway)
    for (int q = tid; q < ((nz + 3) & ~3); q += PACK_T) bul[q] = q < nz ? bci[q] : 0;
    __syncthreads();
    for (int q = tid; q < nz; q += PACK_T) bci[q] = split_local ? R + q : q;
  }
  if (wt == 0) for (int q = nz + tid; q < ((nz + 3) & ~3); q += PACK_T) bci[q] = 0;
  for (int i = tid; i < R + 4; i += PACK_T) hdr[i] = i <= R ? rpl[i] : (i == R + 1 ? nu : (i == R + 2 ? wt : 0));
  if (tid == 0) {
    const int len = tile_blob_bytes(nz, nu, wt);
    tile_len[t] = len;
    atomicMax(&max_len_nu[0], len);
    atomicMax(&max_len_nu[1], nu);
  }
}

void launch_pack_csr_tiles(const int *rp, const int *ci, const double *va, const int *perm, const int *inv,
                           const long long *tile_off, int *tile_len, int *max_len_nu, unsigned char *blob, int n,
                           bool split_local, cudaStream_t s) {
  const int ntiles = (n + SWEEP_TILE_ROWS - 1) / SWEEP_TILE_ROWS;
  pack_csr_tiles_kernel<<<ntiles, PACK_T, 0, s>>>(rp, ci, va, perm, inv, tile_off, tile_len, max_len_nu, blob, n,
                                                  split_local ? 1 : 0);
}

// W, X, y in the internal ordering (row r = caller's observation perm[r], columns relabelled by inv; the entries of a row
// keep their order): one warp per row
__global__ void permute_problem_kernel(const int *__restrict__ rp, const int *__restrict__ ci, const double *__restrict__ va,
                                       const double *__restrict__ X, const double *__restrict__ y, const int *__restrict__ perm,
                                       const int *__restrict__ inv, const int *__restrict__ rp_new, int *__restrict__ ci_new,
                                       double *__restrict__ va_new, double *__restrict__ X_new, double *__restrict__ y_new,
                                       int n, int m) {
  const int lane = threadIdx.x & 31, warps = (gridDim.x * blockDim.x) >> 5;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
    const int src = perm[r], p0 = rp[src], len = rp[src + 1] - p0, q0 = rp_new[r];
    for (int q = lane; q < len; q += 32) { ci_new[q0 + q] = inv[ci[p0 + q]]; va_new[q0 + q] = va[p0 + q]; }
    for (int c = lane; c < m; c += 32) X_new[(size_t)r * m + c] = X[(size_t)src * m + c];
    if (lane == 0) y_new[r] = y[src];
  }
}

void launch_permute_problem(const int *rp, const int *ci, const double *va, const double *X, const double *y, const int *perm,
                            const int *inv, const int *rp_new, int *ci_new, double *va_new, double *X_new, double *y_new,
                            int n, int m, cudaStream_t s) {
  permute_problem_kernel<<<148 * 8, 256, 0, s>>>(rp, ci, va, X, y, perm, inv, rp_new, ci_new, va_new, X_new, y_new, n, m);
}

// rowsum[i] = sum_p W_ip W_{col[p], i}: tr(W^2) = sum_i rowsum[i] (exact second Chebyshev moment)
__global__ void trace_w2_rows_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci,
                                     const double *__restrict__ va, double *__restrict__ rowsum) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int p = rp[i], pe = rp[i + 1]; p < pe; ++p) {
      const int j = ci[p];
      for (int q = rp[j], qe = rp[j + 1]; q < qe; ++q)     // no ordering of the column indices is assumed
        if (ci[q] == i) acc = fma(va[p], va[q], acc);
    }
    rowsum[i] = acc;
  }
}

void launch_trace_w2_rows(int n, const int *rp, const int *ci, const double *va, double *rowsum, cudaStream_t s) {
  trace_w2_rows_kernel<<<592, 256, 0, s>>>(n, rp, ci, va, rowsum);
}

// fixed-point accumulator -> symmetric double matrix (parity entry point bsreg_eval_gram)
__global__ void gram_to_double_kernel(DeviceData dd, double *G) {
  const int d = dd.d;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < d * d; e += gridDim.x * blockDim.x) G[e] = g_load(dd, e / d, e % d);
}

void launch_gram_to_double(const DeviceData &dd, double *G, cudaStream_t s) {
  gram_to_double_kernel<<<(dd.d * dd.d + 255) / 256, 256, 0, s>>>(dd, G);
}

// ---------------------------------------------------------------------------
// K1: out[i, col0 + c] = alpha * sum_p va[p] * B[ci[p], c] + beta * C[i, col0 + c]
// one thread per (row, column); consecutive threads read consecutive columns of a
// gathered row, so the gather is coalesced for m >= 4.
// ---------------------------------------------------------------------------
__global__ void combine_partials_kernel(const double *partial, int nparts, int stride, int count, double *out);

// DOT: also sum_{row, c} U[row][c] * out[row][c] (the Hutchinson dot u'T_j u of the trace moments, fused so that the
// product is not read back): per-CTA partials in fixed order, combined by combine_partials_kernel
template <bool DOT>
__global__ void __launch_bounds__(256) csr_spmm_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci,
                                                       const double *__restrict__ va, const double *__restrict__ B, int m, int ldb,
                                                       double *__restrict__ out, int ldo, int col0, double alpha,
                                                       const double *__restrict__ C, double beta,
                                                       const double *__restrict__ U, double *__restrict__ partial) {
  const size_t total = (size_t)n * m;
  double dot = 0.0;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(idx / m), c = (int)(idx % m);
    double acc = 0.0;
    for (int p = rp[row], pe = rp[row + 1]; p < pe; ++p) acc = fma(va[p], B[(size_t)ci[p] * ldb + c], acc);
    double v = alpha * acc;
    if (C != nullptr) v += beta * C[(size_t)row * ldo + col0 + c];
    out[(size_t)row * ldo + col0 + c] = v;
    if (DOT) dot = fma(U[(size_t)row * ldo + col0 + c], v, dot);
  }
  if (DOT) {
    __shared__ double red[8];
    dot = warp_sum(dot);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += red[w];
      partial[blockIdx.x] = t;
    }
  }
}

void launch_spmm(int n, const int *rp, const int *ci, const double *va, const double *B, int m, int ldb, double *out,
                 int ldo, int col0, double alpha, const double *C, double beta, cudaStream_t s) {
  const size_t total = (size_t)n * m;
  const int block = 256;
  const int grid = (int)((total + block - 1) / block < 148 * 16 ? (total + block - 1) / block : 148 * 16);
  csr_spmm_kernel<false><<<grid > 0 ? grid : 1, block, 0, s>>>(n, rp, ci, va, B, m, ldb, out, ldo, col0, alpha, C, beta,
                                                                nullptr, nullptr);
}

// out = alpha W B + beta C and *dot = sum(U .* out), U laid out like out; `partial` holds one double per CTA (<= 2368)
void launch_spmm_dot(int n, const int *rp, const int *ci, const double *va, const double *B, int m, double *out, double alpha,
                     const double *C, double beta, const double *U, double *partial, double *dot, cudaStream_t s) {
  const size_t total = (size_t)n * m;
  const int block = 256;
  int grid = (int)((total + block - 1) / block < 148 * 16 ? (total + block - 1) / block : 148 * 16);
  if (grid < 1) grid = 1;
  csr_spmm_kernel<true><<<grid, block, 0, s>>>(n, rp, ci, va, B, m, m, out, m, 0, alpha, C, beta, U, partial);
  if (dot != nullptr) combine_partials_kernel<<<1, 32, 0, s>>>(partial, grid, 1, 1, dot);   // null: the caller combines all orders at once
}

int spmm_dot_parts(int n, int m) {
  const size_t total = (size_t)n * m;
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  return grid < 1 ? 1 : grid;
}

// out[e] = sum of row e of partial [count][nparts] (one warp per entry, fixed order): all orders of the recurrence at once
__global__ void combine_rows_kernel(const double *partial, int nparts, int count, double *out) {
  const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (e >= count) return;
  double s0 = 0.0;
  for (int b = lane; b < nparts; b += 32) s0 += partial[(size_t)e * nparts + b];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s0 += __shfl_xor_sync(0xffffffffu, s0, o);
  if (lane == 0) out[e] = s0;
}

void launch_combine_rows(const double *partial, int nparts, int count, double *out, cudaStream_t s) {
  combine_rows_kernel<<<(count + 7) / 8, 256, 0, s>>>(partial, nparts, count, out);
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

// One warp per row reads the row's y, Wy, X (and WX) entries with coalesced loads; per-CTA column sums in shared memory
// (fixed order inside the CTA), partial[cta][d]; the host adds the COLNORM_CTAS partials in order (the norms only fix
// power-of-two scales, but a run-to-run difference in one of them would change the rounding of the fixed-point Gram).
constexpr int COLNORM_CTAS = 296, COLNORM_MAXD = 512;
__global__ void __launch_bounds__(256) colnorm_kernel(DeviceData dd, double *partial) {
  __shared__ double acc[8][COLNORM_MAXD];
  const int d = dd.d, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int c = lane; c < d; c += 32) acc[wid][c] = 0.0;
  __syncwarp();
  for (int row = blockIdx.x * 8 + wid; row < dd.n; row += gridDim.x * 8)
    for (int c = lane; c < d; c += 32) {
      const double v = z_entry(dd, row, c);
      acc[wid][c] = fma(v, v, acc[wid][c]);
    }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += 256) {
    double s0 = 0.0;
    for (int w = 0; w < 8; ++w) s0 += acc[w][c];
    partial[(size_t)blockIdx.x * d + c] = s0;
  }
}

int colnorm_parts() { return COLNORM_CTAS; }

// partial: colnorm_parts() * d doubles
void launch_colnorms(const DeviceData &dd, double *partial, cudaStream_t s) {
  colnorm_kernel<<<COLNORM_CTAS, 256, 0, s>>>(dd, partial);
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

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int m = dd.m, d = dd.d, n = dd.n;
  const bool sem = dd.model == 2, has_w = dd.model != 0;
  if (blockIdx.x == 0 && dd.gzero >= 0) {   // accumulator of the sweep after next
    long long *z = dd.Gfix + (size_t)dd.gzero * d * d;
    for (int e = tid; e < d * d; e += T) z[e] = 0;
  }

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
  // (no finalisation pass: consumers read the fixed-point buffer, internal.cuh: g_load)
  if (rg == 0) {
    long long *Gacc = dd.Gfix + (size_t)dd.gcur * d * d;
    double inv[4][4];                                  // scales first: the atomics below must not wait on these loads
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int gi = 4 * bi + p, gj = 4 * bj + q;
        inv[p][q] = (gi < d && gj < d && gi <= gj) ? dd.Ginv_scale[gi * d + gj] : 0.0;
      }
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int gi = 4 * bi + p, gj = 4 * bj + q;
        if (gi < d && gj < d && gi <= gj) {
          const long long f = __double2ll_rn(acc[p][q] * inv[p][q]);
          atomicAdd(reinterpret_cast<unsigned long long *>(Gacc + gi * d + gj), (unsigned long long)f);
        }
      }
  }
}


// ---------------------------------------------------------------------------
// K1 + K2 fused sweep, main path for d <= 24 (LM / SAR up to K = 22): TMA-fed ring.
//
// HBM layout (packed once at create, see api.cu): the rows are cut into tiles of R = 64;
//   Zt   [ntiles][m+1][R]  column-major inside a tile, columns = X_1..X_m, y; zero padded
//   blob per tile, 16-byte aligned: int rp_local[R+4] | double va[nz~2] | int ci[nz~4]
// so one tile is TWO contiguous bulk copies (cp.async.bulk, complete_tx on an mbarrier).
//
// One persistent, warp-specialised CTA per SM around an S-stage shared-memory ring:
//   producer warp   : one elected lane issues the two bulk copies of a tile into a free stage.
//   7 gather warps  : one tile each (7 tiles in flight): lane = 2 rows, Wy[r] = sum_p va[p] *
//                     y[col[p]] in CSR order with all gathers of the lane in flight at once (CSR
//                     slice from shared memory, y from L2), written as tile column m+1.
//   math warps      : lane = row, warp = one 6x6 block of G = Z'Z (diagonal blocks 21 DFMA,
//                     off-diagonal 36 DFMA per row for 6 / 12 conflict-free LDS.64).
// full[s] / lag[s] / empty[s] mbarriers order producer -> gather -> math -> producer.
// Tail: per-lane partial blocks go through shared memory, one thread per entry of G adds the
// 32 (64) lane partials in a fixed order and issues ONE 64-bit fixed-point atomic; there is
// no finalisation pass -- consumers read the fixed-point buffer (internal.cuh: g_load).
// ---------------------------------------------------------------------------
template <int NG>
__global__ void __launch_bounds__(Ring<NG>::T, 1)
sweep_ring_kernel(DeviceData dd, int ntiles, int S, int stage_bytes, int yg_off, int dbg) {
  using C = Ring<NG>;
  const BlockMap &bm = c_block_map<NG>;
  constexpr bool MMA = NG == 4;                                        // 24 tile columns: Gram update on the FP64 tensor cores
  constexpr int T = C::T, R = C::R, NC = C::NC, WPB = C::WPB, MATHW = C::MATHW, GATHER = C::GATHER;
  constexpr int MATH = MMA ? MMA_WARPS : C::MATH, NRED = MMA ? MMA_BLOCKS * 64 : C::NRED;
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ __align__(8) uint64_t full[RING_MAX_STAGES], lag[RING_MAX_STAGES], empty[RING_MAX_STAGES];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int m = dd.m, d = dd.d;
  const bool has_w = dd.model != 0;
  const double *__restrict__ yg = dd.yp;

  // the accumulator of the sweep after next: its last reader finished before this launch
  if (blockIdx.x == 0 && dd.gzero >= 0) {
    long long *z = dd.Gfix + (size_t)dd.gzero * d * d;
    for (int e = tid; e < d * d; e += T) z[e] = 0;
  }
  // tail bookkeeping, loaded early so that its latency is off the critical path
  int gidx = -1;
  double inv = 0.0;
  if (tid < NRED) {
    int ca, cb;                                                       // tile columns
    bool lower;
    if (MMA) {
      const int blk = tid >> 6, p = (tid >> 3) & 7, q = tid & 7;
      ca = 8 * mma_bi(blk) + p; cb = 8 * mma_bj(blk) + q;
      lower = mma_bi(blk) == mma_bj(blk) && q < p;
    } else {
      const int blk = tid / 36, e = tid - blk * 36, p = e / 6, q = e - p * 6;
      ca = 6 * bm.gi[blk] + p; cb = 6 * bm.gj[blk] + q;
      lower = bm.gi[blk] == bm.gj[blk] && q < p;
    }
    const int ncol = has_w ? m + 2 : m + 1;
    if (!lower && ca < ncol && cb < ncol) {
      const int ga = ca < m ? 2 + ca : ca - m, gb = cb < m ? 2 + cb : cb - m;   // G index: y = 0, Wy = 1, X_a = 2 + a
      gidx = min(ga, gb) * d + max(ga, gb);
      inv = dd.Ginv_scale[gidx];
    }
  }
  auto stZ = [&](int s) { return reinterpret_cast<double *>(ring + (size_t)s * stage_bytes); };
  auto stC = [&](int s) { return ring + (size_t)s * stage_bytes + (size_t)NC * R * 8; };
  // columns the copies never write (padding, and Wy when there is no W) stay zero
  {
    const int c0 = has_w ? m + 2 : m + 1, npad = (NC - c0) * R;
    for (int e = tid; e < S * npad; e += T) {
      const int s = e / npad, o = e - s * npad;
      stZ(s)[c0 * R + o] = 0.0;
    }
  }
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&lag[s], 1); mbar_init(&empty[s], MATH); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");

  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const uint32_t z_bytes = (uint32_t)(m + 1) * R * 8u;

  double *red = reinterpret_cast<double *>(ring);
  constexpr int LD = MMA ? MMA_WARPS + 1 : WPB * 32 + 1;

  if (wid >= MATHW) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(C::REG_LIGHT));
  if (wid == MATHW + GATHER) {
    // =========================== producer ===========================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;                                                // fresh barriers: the "previous" phase is complete
      long long o0 = 0;
      int o1 = 0;
      if (has_w && my_tiles > 0) { o0 = dd.tile_off[blockIdx.x]; o1 = dd.tile_len[blockIdx.x]; }
      for (int j = 0; j < my_tiles; ++j) {
        const int tile = blockIdx.x + j * gridDim.x;
        long long n0 = 0;                                             // next tile's blob, fetched early
        int n1 = 0;
        if (has_w && j + 1 < my_tiles) { n0 = dd.tile_off[tile + gridDim.x]; n1 = dd.tile_len[tile + gridDim.x]; }
        mbar_wait(&empty[s], ph);
        const uint32_t c_bytes = (uint32_t)o1;
        mbar_expect_tx(&full[s], z_bytes + c_bytes);
        tma_bulk_g2s(stZ(s), dd.Zt + (size_t)tile * (m + 1) * R, z_bytes, &full[s]);
        if (c_bytes) tma_bulk_g2s(stC(s), dd.csr_blob + o0, c_bytes, &full[s]);
        o0 = n0; o1 = n1;
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
    asm volatile("bar.sync 0;\n" ::: "memory");
  } else if (wid >= MATHW) {
    // =========================== gather: Wy of the tile ===========================
    // one warp per tile, GATHER tiles in flight; lane -> rows lane, lane + 32; the first GU
    // nonzeros of both rows are gathered with every load in flight, longer rows finish serially
    // a waiter may be at most one phase away from its barrier: never more gather warps than stages
    const int g = wid - MATHW, ng = S < GATHER ? S : GATHER;
    int s = g;
    uint32_t ph = 0;
    for (int j = g; j < (g < ng ? my_tiles : 0); j += ng) {
      mbar_wait(&full[s], ph);
      if (has_w && !(dbg & 4)) ring_gather_tile(stZ(s), stC(s), reinterpret_cast<double *>(stC(s) + yg_off), yg, m, lane);
      __syncwarp();
      if (lane == 0) mbar_arrive(&lag[s]);
      s += ng;
      if (s >= S) { s -= S; ph ^= 1; }
    }
    asm volatile("bar.sync 0;\n" ::: "memory");
  } else {
    // =========================== math: this warp's 6x6 block ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(C::REG_MATH));
    const int blk = wid % C::NBLK, wr = wid / C::NBLK;
    const int gi = bm.gi[blk], gj = bm.gj[blk];
    const bool diag = gi == gj;
    // offsets of this lane's row in the block's columns (row swizzle of the tile: tile_row_pos; bits 2-3 of the lane only)
    int oa[6], ob[6];
#pragma unroll
    for (int p = 0; p < 6; ++p) {
      oa[p] = (6 * gi + p) * R + tile_row_pos(6 * gi + p, lane);
      ob[p] = (6 * gj + p) * R + tile_row_pos(6 * gj + p, lane);
    }
    if constexpr (MMA) {
      double acc[MMA_BLOCKS][2];
#pragma unroll
      for (int q = 0; q < MMA_BLOCKS; ++q) { acc[q][0] = 0.0; acc[q][1] = 0.0; }
      int fo[3];
      mma_frag_offsets(fo, lane);
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < (wid < MATH ? my_tiles : 0); ++j) {
        mbar_wait(&lag[s], ph);
        if (!(dbg & 2)) mma_tile(acc, stZ(s), fo, wid);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == S) { s = 0; ph ^= 1; }
      }
      asm volatile("bar.sync 0;\n" ::: "memory");                     // every stage has been consumed
      if (wid < MATH) {
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int q = 0; q < MMA_BLOCKS; ++q) {
          red[(q * 64 + g * 8 + 2 * t) * LD + wid] = acc[q][0];
          red[(q * 64 + g * 8 + 2 * t + 1) * LD + wid] = acc[q][1];
        }
      }
    } else {
    double acc[6][6];
#pragma unroll
    for (int p = 0; p < 6; ++p)
#pragma unroll
      for (int q = 0; q < 6; ++q) acc[p][q] = 0.0;
    int s = 0;
    uint32_t ph = 0;
    for (int j = 0; j < (wid < MATH ? my_tiles : 0); ++j) {
      mbar_wait(&lag[s], ph);
      const double *Z = stZ(s);
      if (!(dbg & 2))
#pragma unroll
      for (int rr = 0; rr < (R / 32) / WPB; ++rr) {
        const int r32 = 32 * (wr + WPB * rr);
        double a[6], b[6];
#pragma unroll
        for (int p = 0; p < 6; ++p) a[p] = Z[oa[p] + r32];
        if (diag) {
#pragma unroll
          for (int p = 0; p < 6; ++p)
#pragma unroll
            for (int q = p; q < 6; ++q) acc[p][q] = fma(a[p], a[q], acc[p][q]);
        } else {
#pragma unroll
          for (int p = 0; p < 6; ++p) b[p] = Z[ob[p] + r32];
#pragma unroll
          for (int p = 0; p < 6; ++p)
#pragma unroll
            for (int q = 0; q < 6; ++q) acc[p][q] = fma(a[p], b[q], acc[p][q]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      if (++s == S) { s = 0; ph ^= 1; }
    }
    // ---- tail: lane partials -> shared memory -> one thread per entry -> one atomic ----
    asm volatile("bar.sync 0;\n" ::: "memory");                       // every stage has been consumed
    if (wid < MATH) {
#pragma unroll
      for (int p = 0; p < 6; ++p)
#pragma unroll
        for (int q = 0; q < 6; ++q) red[(blk * 36 + p * 6 + q) * LD + wr * 32 + lane] = acc[p][q];
    }
    }
  }
  __syncthreads();
  if (gidx >= 0) {
    const double *v = red + tid * LD;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;                    // fixed order
#pragma unroll 8
    for (int l = 0; l < LD - 1; l += 4) { s0 += v[l]; s1 += v[l + 1]; s2 += v[l + 2]; s3 += v[l + 3]; }
    const long long f = __double2ll_rn(((s0 + s1) + (s2 + s3)) * inv);
    atomicAdd(reinterpret_cast<unsigned long long *>(dd.Gfix + (size_t)dd.gcur * d * d + gidx), (unsigned long long)f);
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
    if (lane == 0) dd.Gfix[(size_t)dd.gcur * d * d + i * d + j] = __double2ll_rn(acc * dd.Ginv_scale[i * d + j]);
  }
}

bool sweep_fast_supported(const DeviceData &dd) { return dd.d <= 88; }

// ring geometry for a given design: stages that fit the shared-memory budget (0 = use the other paths)
bool sweep_ring_supported(int d, int model, int max_blob, int max_nu) { return ring_supported_impl(d, model, max_blob, max_nu); }

template <int NG>
static void launch_sweep_ring(const DeviceData &dd, int sm_count, cudaStream_t s) {
  using C = Ring<NG>;
  const int S = ring_stages(dd.d, dd.tile_max_blob, dd.tile_max_nu), sb = sweep_ring_stage_bytes(dd.d, dd.tile_max_blob, dd.tile_max_nu);
  size_t smem = (size_t)S * sb;
  if (smem < C::RED_BYTES) smem = C::RED_BYTES;
  if (smem_optin((const void *)sweep_ring_kernel<NG>, smem) != cudaSuccess) return;   // the error stays pending for cudaGetLastError
  const int ntiles = (dd.n + C::R - 1) / C::R;
  const int grid = ntiles < sm_count ? ntiles : sm_count;
  static const int dbg = getenv("BSREG_DBG") ? atoi(getenv("BSREG_DBG")) : 0;   // ablation switches (profiling only)
  sweep_ring_kernel<NG><<<grid, C::T, smem, s>>>(dd, ntiles, S, sb, ring_yg_offset(dd.d, dd.tile_max_blob) - C::NC * C::R * 8, dbg);
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
  DeviceData dv = dd;                              // Z'Z does not depend on the order of the observations
  if (dd.sw_rp != nullptr) { dv.rp = dd.sw_rp; dv.ci = dd.sw_ci; dv.va = dd.sw_va; dv.X = dd.sw_X; dv.y = dd.sw_y; }
  sweep_gram_kernel<NB><<<grid, Cfg::T, 0, s>>>(dv, ntiles);
}

void launch_sweep(const DeviceData &dd, int sm_count, cudaStream_t s) {
  ++g_sweep_launches;
  if (!sweep_fast_supported(dd)) {   // wide designs (d > 88): refresh the lags through W, then one warp per entry
    if (dd.model != 0) launch_spmm(dd.n, dd.rp, dd.ci, dd.va, dd.y, 1, 1, dd.Wy, 1, 0, 1.0, nullptr, 0.0, s);
    if (dd.model == 2) launch_spmm(dd.n, dd.rp, dd.ci, dd.va, dd.X, dd.m, dd.m, dd.WX, dd.m, 0, 1.0, nullptr, 0.0, s);
    gram_generic_kernel<<<296, 256, 0, s>>>(dd);
    return;
  }
  if (dd.Xr != nullptr) { launch_sweep_sem(dd, sm_count, s); return; }   // SEM ring sweep (packed at create)
  if (dd.Zt != nullptr) {                       // packed at create iff the ring covers this design
    if (dd.d <= 6) launch_sweep_ring<1>(dd, sm_count, s);
    else if (dd.d <= 12) launch_sweep_ring<2>(dd, sm_count, s);
    else if (dd.d <= 18) launch_sweep_ring<3>(dd, sm_count, s);
    else launch_sweep_ring<4>(dd, sm_count, s);
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

// out[e] = sum over the CTAs' partials, one warp per entry: lane l adds parts l, l + 32, ... in order, then a fixed
// shuffle tree -- the same order on every run
__global__ void combine_partials_kernel(const double *partial, int nparts, int stride, int count, double *out) {
  const int e = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (e < count) {
    double s = 0.0;
    for (int b = lane; b < nparts; b += 32) s += partial[(size_t)b * stride + e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[e] = s;
  }
}

void launch_residual_ss(const DeviceData &dd, int n_eval, const double *beta,
                        const double *lam, const double *rho, double *partial, double *out, cudaStream_t s) {
  for (int e0 = 0; e0 < n_eval; e0 += RSS_EC) {
    residual_ss_kernel<<<RSS_GRID, 256, RSS_EC * dd.m * sizeof(double), s>>>(dd, e0, n_eval, beta, lam, rho, partial);
    const int ne = n_eval - e0 < RSS_EC ? n_eval - e0 : RSS_EC;
    combine_partials_kernel<<<1, 32 * RSS_EC, 0, s>>>(partial, RSS_GRID, RSS_EC, ne, out + e0);
  }
}

// ---------------------------------------------------------------------------
// K4 helpers: Rademacher probes (same bits as oracle/core.py:rademacher_probes) and dots
// ---------------------------------------------------------------------------
// row `row` of U holds the probes of observation perm[row] (perm == null: of observation row); rows [n, n_rows) are zero
__global__ void probes_kernel(double *U, int n, int n_probes, uint64_t seed, const int *__restrict__ perm, int n_rows) {
  const int nb = (n_probes + 127) / 128;
  const Philox ph(seed, 0xffffffffu);
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (size_t)n_rows * nb; idx += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(idx / nb), b = (int)(idx % nb);
    if (row >= n) {
      for (int p = 128 * b; p < min(n_probes, 128 * b + 128); ++p) U[(size_t)row * n_probes + p] = 0.0;
      continue;
    }
    const uint4 r = ph.raw((uint32_t)(perm ? perm[row] : row), SLOT_PROBE, (uint32_t)b, 0xffffffffu);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    for (int k = 0; k < 4; ++k)
      for (int j = 0; j < 32; ++j) {
        const int p = 128 * b + 32 * k + j;
        if (p < n_probes) U[(size_t)row * n_probes + p] = 1.0 - 2.0 * (double)((w[k] >> j) & 1u);
      }
  }
}

void launch_probes(double *U, int n, int n_probes, uint64_t seed, cudaStream_t s, const int *perm, int n_rows) {
  probes_kernel<<<592, 256, 0, s>>>(U, n, n_probes, seed, perm, n_rows > n ? n_rows : n);
}

constexpr int DOT_GRID = 592;
__global__ void __launch_bounds__(256) dot_kernel(const double *__restrict__ a, const double *__restrict__ b, int64_t len,
                                                  double *partial) {
  __shared__ double red[8];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x)
    acc = b ? fma(a[i], b[i], acc) : acc + a[i];
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
