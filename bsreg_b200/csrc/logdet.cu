// Grid + spline log-determinant (`ldet_SAR / ldet_SEM = list(grid = TRUE)`), R/15_sar.R:79-87 = R/15_sem.R:91-99:
//
//   pars  <- i_seq(i_lambda)                                        (R/00_aux.R:105-107)
//   ldets <- determinant(diag(size) - x * get_W(), logarithm = TRUE)$modulus * reps     for x in pars
//   get_ldet <- splinefun(pars, ldets)                              (stats::splinefun, method "fmm")
//
// Here: the G grid matrices I - x_g W_sub are built dense on the device from the CSR of W (column-major, one after the
// other), factorised in place by one CTA each (LU with partial pivoting, the algorithm behind R's determinant()), and
// log|det| = sum log|u_kk|.  The FMM spline coefficients (Forsythe, Malcolm, Moler 1977, SPLINE) are formed on the host
// from the G values; the chains evaluate the piecewise cubic per proposal (SEVAL; chainlib.cuh: spline_ldet).
#include <cmath>
#include <vector>

#include "internal.cuh"

namespace bsreg {

// A_g = I - x_g * W[0:size, 0:size], column-major size x size, g = blockIdx.y
__global__ void build_grid_matrices_kernel(int size, const int *__restrict__ rp, const int *__restrict__ ci,
                                           const double *__restrict__ va, const double *__restrict__ xs, double *__restrict__ A) {
  const int g = blockIdx.y;
  const double x = xs[g];
  double *Ag = A + (size_t)g * size * size;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < size; i += gridDim.x * blockDim.x) {
    Ag[(size_t)i * size + i] += 1.0;                            // the buffer was zeroed; a diagonal entry of W adds below
    for (int p = rp[i], pe = rp[i + 1]; p < pe; ++p) {
      const int j = ci[p];
      if (j < size) Ag[(size_t)j * size + i] -= x * va[p];      // row i, column j (duplicates in a row would race: CSR is canonical)
    }
  }
}

// In-place LU with partial pivoting of one column-major matrix per CTA; out[g] = reps * sum_k log|u_kk|.
constexpr int LU_T = 256;
__global__ void __launch_bounds__(LU_T) lu_logdet_kernel(int n, double *__restrict__ A, double reps, double *__restrict__ out) {
  __shared__ double s_val[LU_T / 32];
  __shared__ int s_idx[LU_T / 32];
  __shared__ int s_piv;
  __shared__ double s_pivval;
  double *a = A + (size_t)blockIdx.x * n * n;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  double logdet = 0.0;
  for (int k = 0; k < n; ++k) {
    // pivot: largest |a_ik|, i >= k (ties: the smallest i, as LAPACK's idamax)
    double best = -1.0;
    int bi = k;
    for (int i = k + tid; i < n; i += LU_T) {
      const double v = fabs(a[(size_t)k * n + i]);
      if (v > best) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { s_val[wid] = best; s_idx[wid] = bi; }
    __syncthreads();
    if (tid == 0) {
      double b = s_val[0];
      int p = s_idx[0];
      for (int w = 1; w < LU_T / 32; ++w)
        if (s_val[w] > b || (s_val[w] == b && s_idx[w] < p)) { b = s_val[w]; p = s_idx[w]; }
      s_piv = p;
      s_pivval = a[(size_t)k * n + p];
    }
    __syncthreads();
    const int p = s_piv;
    const double piv = s_pivval;
    logdet += log(fabs(piv));                                     // every thread carries the same sum
    if (p != k)
      for (int j = tid; j < n; j += LU_T) {                       // swap rows k and p
        const double t = a[(size_t)j * n + k];
        a[(size_t)j * n + k] = a[(size_t)j * n + p];
        a[(size_t)j * n + p] = t;
      }
    __syncthreads();
    if (piv == 0.0) break;                                        // singular: log|det| = -Inf (R: modulus -Inf)
    const double rinv = 1.0 / piv;
    for (int i = k + 1 + tid; i < n; i += LU_T) a[(size_t)k * n + i] *= rinv;
    __syncthreads();
    // trailing update a_ij -= l_i a_kj: consecutive threads take consecutive rows of a column (coalesced)
    const int rem = n - k - 1;
    for (int e = tid; e < rem * rem; e += LU_T) {
      const int j = k + 1 + e / rem, i = k + 1 + e % rem;
      a[(size_t)j * n + i] = fma(-a[(size_t)k * n + i], a[(size_t)j * n + k], a[(size_t)j * n + i]);
    }
    __syncthreads();
  }
  if (tid == 0) out[blockIdx.x] = reps * logdet;
}

void launch_grid_logdets(int size, const int *rp, const int *ci, const double *va, int n_grid, const double *xs_dev,
                         double *A, double reps, double *out, cudaStream_t s) {
  cudaMemsetAsync(A, 0, (size_t)n_grid * size * size * 8, s);
  build_grid_matrices_kernel<<<dim3((size + 255) / 256, n_grid), 256, 0, s>>>(size, rp, ci, va, xs_dev, A);
  lu_logdet_kernel<<<n_grid, LU_T, 0, s>>>(size, A, reps, out);
}

// FMM's SPLINE on the host: coefficients (b, c, d) of s(u) = y_i + dx (b_i + dx (c_i + dx d_i)), dx = u - x_i
void fmm_spline_coefficients(int n, const double *x, const double *y, double *b, double *c, double *d) {
  for (int i = 0; i < n; ++i) { b[i] = 0.0; c[i] = 0.0; d[i] = 0.0; }
  if (n < 2) return;
  if (n < 3) {
    b[0] = b[1] = (y[1] - y[0]) / (x[1] - x[0]);
    return;
  }
  d[0] = x[1] - x[0];
  c[1] = (y[1] - y[0]) / d[0];
  for (int i = 1; i < n - 1; ++i) {
    d[i] = x[i + 1] - x[i];
    b[i] = 2.0 * (d[i - 1] + d[i]);
    c[i + 1] = (y[i + 1] - y[i]) / d[i];
    c[i] = c[i + 1] - c[i];
  }
  b[0] = -d[0];
  b[n - 1] = -d[n - 2];
  c[0] = c[n - 1] = 0.0;
  if (n > 3) {
    c[0] = c[2] / (x[3] - x[1]) - c[1] / (x[2] - x[0]);
    c[n - 1] = c[n - 2] / (x[n - 1] - x[n - 3]) - c[n - 3] / (x[n - 2] - x[n - 4]);
    c[0] = c[0] * d[0] * d[0] / (x[3] - x[0]);
    c[n - 1] = -c[n - 1] * d[n - 2] * d[n - 2] / (x[n - 1] - x[n - 4]);
  }
  for (int i = 1; i < n; ++i) {
    const double t = d[i - 1] / b[i - 1];
    b[i] = b[i] - t * d[i - 1];
    c[i] = c[i] - t * c[i - 1];
  }
  c[n - 1] = c[n - 1] / b[n - 1];
  for (int i = n - 2; i >= 0; --i) c[i] = (c[i] - d[i] * c[i + 1]) / b[i];
  b[n - 1] = (y[n - 1] - y[n - 2]) / d[n - 2] + d[n - 2] * (c[n - 2] + 2.0 * c[n - 1]);
  for (int i = 0; i < n - 1; ++i) {
    b[i] = (y[i + 1] - y[i]) / d[i] - d[i] * (c[i + 1] + 2.0 * c[i]);
    d[i] = (c[i + 1] - c[i]) / d[i];
    c[i] = 3.0 * c[i];
  }
  c[n - 1] = 3.0 * c[n - 1];
  d[n - 1] = d[n - 2];
}

}  // namespace bsreg
