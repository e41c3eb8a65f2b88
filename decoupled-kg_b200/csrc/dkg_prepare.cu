// Candidate-independent preparation kernels: everything GPyTorch's exact prediction strategy
// caches (training Cholesky, mean cache) plus the tables the reference recomputes per candidate
// although they do not depend on it (posterior means at the discretisation, the scalarised
// intercept table, K^-1 k(X_train, X_disc)).  Replaces the per-candidate posterior calls at
// discretekg.py:275-284 / :300.
#include "dkg_kernels.cuh"

namespace dkg {

// out[r, k] = x[r, k] / ls[k]   (GPyTorch divides inputs by the lengthscale)
__global__ void scale_rows_kernel(const double* __restrict__ x, int rows, int d,
                                  const double* __restrict__ ls, double* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * d) out[i] = x[i] / ls[i % d];
}

struct LsArg {
  double v[MAX_D];
};

__global__ void scale_rows_kernel_v(const double* __restrict__ x, int rows, int d, LsArg ls,
                                    double* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * d) out[i] = x[i] / ls.v[i % d];
}

int scale_rows(const double* x, int rows, int d, const double* ls_host, double* out,
               cudaStream_t st) {
  LsArg a;
  for (int k = 0; k < MAX_D; ++k) a.v[k] = k < d ? ls_host[k] : 1.0;
  int total = rows * d;
  if (total == 0) return DKG_OK;
  scale_rows_kernel_v<<<ceil_div(total, 256), 256, 0, st>>>(x, rows, d, a, out);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// K[t, u] = k(xs_t, xs_u) + noise * [t == u]
__global__ void kmat_train_kernel(const double* __restrict__ xs, int n, int d, int kind,
                                  double outputscale, double noise, double* __restrict__ K,
                                  int ld) {
  int u = blockIdx.x * blockDim.x + threadIdx.x;
  int t = blockIdx.y;
  if (u >= n) return;
  double sq = 0.0;
  for (int k = 0; k < d; ++k) {
    double df = xs[t * d + k] - xs[u * d + k];
    sq += df * df;
  }
  double v = stationary_from_sq(kind, outputscale, sq);
  if (t == u) v += noise;
  K[(size_t)t * ld + u] = v;
}

int kmat_train(const ObjState& o, int d, double jitter, double* K, int ld, cudaStream_t st) {
  dim3 grid(ceil_div(o.n, 128), o.n);
  kmat_train_kernel<<<grid, 128, 0, st>>>(o.xs, o.n, d, o.kernel, o.outputscale,
                                          o.noise + jitter, K, ld);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// ------------------------------------------------------------------------------------------
// In-place lower Cholesky of a dense n x n matrix held in global memory (L2 resident), one CTA.
// Right-looking, column at a time; the strict upper triangle is zeroed at the end.
// info[0] = 0 on success, j+1 if the pivot of column j was not positive.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1)
cholesky_kernel(double* __restrict__ A, int n, int* __restrict__ info) {
  __shared__ double s_piv;
  __shared__ int s_fail;
  const int tid = threadIdx.x;
  const int nt = blockDim.x;
  if (tid == 0) s_fail = 0;
  __syncthreads();
  for (int j = 0; j < n; ++j) {
    if (tid == 0) {
      double dj = A[(size_t)j * n + j];
      if (!(dj > 0.0)) {
        s_fail = j + 1;
        s_piv = 1.0;
      } else {
        s_piv = sqrt(dj);
        A[(size_t)j * n + j] = s_piv;
      }
    }
    __syncthreads();
    if (s_fail) break;
    const double piv = s_piv;
    for (int i = j + 1 + tid; i < n; i += nt) A[(size_t)i * n + j] /= piv;
    __syncthreads();
    // trailing update: A[i, k] -= A[i, j] * A[k, j] for j < k <= i.  Flattened over the
    // (n-j-1) x (n-j-1) square, lower part only; k fastest so that A[i, k] is coalesced.
    const int m = n - j - 1;
    const long long total = (long long)m * m;
    for (long long e = tid; e < total; e += nt) {
      int ii = (int)(e / m);
      int kk = (int)(e - (long long)ii * m);
      if (kk <= ii) {
        int i = j + 1 + ii, k = j + 1 + kk;
        A[(size_t)i * n + k] -= A[(size_t)i * n + j] * A[(size_t)k * n + j];
      }
    }
    __syncthreads();
  }
  if (tid == 0) info[0] = s_fail;
  __syncthreads();
  if (!s_fail) {
    for (long long e = tid; e < (long long)n * n; e += nt) {
      int i = (int)(e / n), k = (int)(e % n);
      if (k > i) A[e] = 0.0;
    }
  }
}

int cholesky_inplace(double* A, int n, int* info_dev, cudaStream_t st) {
  cholesky_kernel<<<1, 1024, 0, st>>>(A, n, info_dev);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

__global__ void transpose_kernel(const double* __restrict__ in, int rows, int cols, int ld_in,
                                 double* __restrict__ out, int ld_out) {
  __shared__ double tile[32][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  int r0 = blockIdx.y * 32;
  for (int k = threadIdx.y; k < 32; k += blockDim.y) {
    int r = r0 + k;
    tile[k][threadIdx.x] = (r < rows && c < cols) ? in[(size_t)r * ld_in + c] : 0.0;
  }
  __syncthreads();
  int oc = blockIdx.y * 32 + threadIdx.x;  // output column = input row
  for (int k = threadIdx.y; k < 32; k += blockDim.y) {
    int orow = blockIdx.x * 32 + k;  // output row = input column
    if (orow < cols && oc < rows) out[(size_t)orow * ld_out + oc] = tile[threadIdx.x][k];
  }
}

int transpose(const double* in, int rows, int cols, int ld_in, double* out, int ld_out,
              cudaStream_t st) {
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32));
  transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(in, rows, cols, ld_in, out, ld_out);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// ------------------------------------------------------------------------------------------
// Triangular solves with many right-hand sides, one thread per column (coalesced across
// columns; the triangular factor is read as warp-uniform broadcasts).
//   forward : L Y = R      (L lower, row-major n x n)
//   backward: L^T X = Y    (LT = L^T given row-major so that the inner loop reads a row)
// R is [n, ld] and is overwritten in place.
// ------------------------------------------------------------------------------------------
__global__ void trsm_forward_kernel(const double* __restrict__ L, int n, double* __restrict__ R,
                                    int ncols, int ld) {
  int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  for (int t = 0; t < n; ++t) {
    const double* Lrow = L + (size_t)t * n;
    double acc = R[(size_t)t * ld + col];
    for (int u = 0; u < t; ++u) acc -= Lrow[u] * R[(size_t)u * ld + col];
    R[(size_t)t * ld + col] = acc / Lrow[t];
  }
}

__global__ void trsm_backward_kernel(const double* __restrict__ LT, int n, double* __restrict__ R,
                                     int ncols, int ld) {
  int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  for (int t = n - 1; t >= 0; --t) {
    const double* Urow = LT + (size_t)t * n;  // Urow[u] = L[u, t]
    double acc = R[(size_t)t * ld + col];
    for (int u = n - 1; u > t; --u) acc -= Urow[u] * R[(size_t)u * ld + col];
    R[(size_t)t * ld + col] = acc / Urow[t];
  }
}

int cholesky_solve_inplace(const double* L, const double* LT, int n, double* R, int ncols, int ld,
                           cudaStream_t st) {
  if (ncols == 0) return DKG_OK;
  int threads = 64;
  trsm_forward_kernel<<<ceil_div(ncols, threads), threads, 0, st>>>(L, n, R, ncols, ld);
  DKG_LAUNCH_CHECK();
  trsm_backward_kernel<<<ceil_div(ncols, threads), threads, 0, st>>>(LT, n, R, ncols, ld);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// R[t, n] = k_i(xtr_t, xd_n) for t < n_train, n < N   (zero elsewhere; R is [n_pad, N_pad])
__global__ void kcross_kernel(const double* __restrict__ xs, int ntr, const double* __restrict__ xd_s,
                              int N, int d, int kind, double outputscale, double* __restrict__ R,
                              int ld) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int t = blockIdx.y;
  if (n >= N) return;
  double sq = 0.0;
  for (int k = 0; k < d; ++k) {
    double df = xs[t * d + k] - xd_s[(size_t)n * d + k];
    sq += df * df;
  }
  R[(size_t)t * ld + n] = stationary_from_sq(kind, outputscale, sq);
}

int kcross(const ObjState& o, const double* xd_s, int N, int d, double* R, int ld,
           cudaStream_t st) {
  dim3 grid(ceil_div(N, 128), o.n);
  kcross_kernel<<<grid, 128, 0, st>>>(o.xs, o.n, xd_s, N, d, o.kernel, o.outputscale, R, ld);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

__global__ void identity_kernel(double* __restrict__ A, int n, int ld) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  int r = blockIdx.y;
  if (c < n) A[(size_t)r * ld + c] = (r == c) ? 1.0 : 0.0;
}

int set_identity(double* A, int n, int ld, cudaStream_t st) {
  dim3 grid(ceil_div(n, 128), n);
  identity_kernel<<<grid, 128, 0, st>>>(A, n, ld);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// rhs[t] = y[t] - c
__global__ void resid_kernel(const double* __restrict__ y, int n, double c, double* __restrict__ out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = y[t] - c;
}

int residual(const double* y, int n, double c, double* out, cudaStream_t st) {
  resid_kernel<<<ceil_div(n, 128), 128, 0, st>>>(y, n, c, out);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// mu[n, m] = (c_m + sum_t k_m(xd_n, xtr_t) alpha_m[t]) * ystd_m + ymean_m ; one warp per point.
__global__ void mu_disc_kernel(const double* __restrict__ xd, int N, int d, const double* __restrict__ xs,
                               const double* __restrict__ alpha, int ntr, LsArg ls, int kind,
                               double outputscale, double mean_const, double y_mean, double y_std,
                               double* __restrict__ mu, int M, int m) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= N) return;
  double xq[MAX_D];
#pragma unroll
  for (int k = 0; k < MAX_D; ++k) xq[k] = k < d ? xd[(size_t)warp * d + k] / ls.v[k] : 0.0;
  double acc = 0.0;
  for (int t = lane; t < ntr; t += 32) {
    double sq = 0.0;
#pragma unroll
    for (int k = 0; k < MAX_D; ++k)
      if (k < d) {
        double df = xq[k] - xs[t * d + k];
        sq += df * df;
      }
    acc += stationary_from_sq(kind, outputscale, sq) * alpha[t];
  }
  acc = warp_sum(acc);
  if (lane == 0) mu[(size_t)warp * M + m] = (mean_const + acc) * y_std + y_mean;
}

int mu_disc(const double* xd, int N, int d, const ObjState& o, double* mu, int M, int m,
            cudaStream_t st) {
  LsArg a;
  for (int k = 0; k < MAX_D; ++k) a.v[k] = k < d ? o.ls[k] : 1.0;
  int threads = 256;
  long long total = (long long)N * 32;
  mu_disc_kernel<<<(int)((total + threads - 1) / threads), threads, 0, st>>>(
      xd, N, d, o.xs, o.alpha, o.n, a, o.kernel, o.outputscale, o.mean_const, o.y_mean, o.y_std, mu,
      M, m);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// A0[j, n] = sum_m W[j, m] * mu[n, m]  (products rounded, summed left to right -- matches
// torch.sum(weights * means, dim=-1) at discretekg.py:320 for the discretisation lines).
__global__ void a0_kernel(const double* __restrict__ mu, int N, int M, const double* __restrict__ W,
                          int S, double* __restrict__ A0, float* __restrict__ A0f, int ld) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y;
  if (n >= ld) return;
  double acc = 0.0;
  if (n < N) {
    acc = __dmul_rn(W[j * M + 0], mu[(size_t)n * M + 0]);
    for (int m = 1; m < M; ++m) acc = __dadd_rn(acc, __dmul_rn(W[j * M + m], mu[(size_t)n * M + m]));
  } else {
    acc = -INFINITY;  // padding lines can never win a max
  }
  A0[(size_t)j * ld + n] = acc;
  A0f[(size_t)j * ld + n] = (float)acc;  // the fp32 chord filter budgets for this rounding
}

// A0max[j], A0arg[j] (first index of the maximum); one CTA per scalarisation.
__global__ void a0max_kernel(const double* __restrict__ A0, int N, int ld, double* __restrict__ A0max,
                             int* __restrict__ A0arg) {
  __shared__ double sv[32];
  __shared__ int si[32];
  int j = blockIdx.x;
  double best = -INFINITY;
  int bi = 0x7fffffff;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    double v = A0[(size_t)j * ld + n];
    if (v > best || (v == best && n < bi)) {
      best = v;
      bi = n;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    sv[w] = best;
    si[w] = bi;
  }
  __syncthreads();
  if (w == 0) {
    int nw = blockDim.x >> 5;
    best = l < nw ? sv[l] : -INFINITY;
    bi = l < nw ? si[l] : 0x7fffffff;
    for (int o = 16; o > 0; o >>= 1) {
      double ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) {
        best = ov;
        bi = oi;
      }
    }
    if (l == 0) {
      A0max[j] = best;
      A0arg[j] = bi;
    }
  }
}

// A0tmax[t, j] = max of the float intercepts of scalarisation j over the lines [t * tile, (t + 1) * tile):
// the fp32 chord filter culls a whole warp tile of lines with one test against it
__global__ void a0_tilemax_kernel(const float* __restrict__ A0f, int ld, int tile, int ntiles, float* __restrict__ out,
                                  int* __restrict__ out_arg) {
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), j = blockIdx.y, lane = threadIdx.x & 31;
  if (t >= ntiles) return;
  float m = -INFINITY;
  int a = 0x7fffffff;  // first line of the tile that attains the maximum (the tile's "champion"); -1: padding only
  for (int n = t * tile + lane; n < min((t + 1) * tile, ld); n += 32) {
    const float v = A0f[(size_t)j * ld + n];
    if (v > m) { m = v; a = n; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oa = __shfl_xor_sync(0xffffffffu, a, o);
    if (om > m || (om == m && oa < a)) { m = om; a = oa; }
  }
  if (lane == 0) {
    out[(size_t)t * gridDim.y + j] = m;  // [tile][S]: a half-warp of the tile filter reads 16 neighbours
    if (out_arg != nullptr) out_arg[(size_t)t * gridDim.y + j] = a == 0x7fffffff ? -1 : a;
  }
}

int build_a0_tilemax(const float* A0f, int ld, int S, int tile, int ntiles, float* out, int* out_arg, cudaStream_t st) {
  dim3 grid(ceil_div(ntiles, 4), S);
  a0_tilemax_kernel<<<grid, 128, 0, st>>>(A0f, ld, tile, ntiles, out, out_arg);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// out[i, :] = in[perm[i], :]
__global__ void gather_rows_kernel(const double* __restrict__ in, const int* __restrict__ perm, int rows, int d,
                                   double* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * d) return;
  const int i = e / d, k = e - i * d;
  out[e] = in[(size_t)perm[i] * d + k];
}
int gather_rows(const double* in, const int* perm, int rows, int d, double* out, cudaStream_t st) {
  gather_rows_kernel<<<ceil_div(rows * d, 256), 256, 0, st>>>(in, perm, rows, d, out);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// undo the plan's internal line order for introspection: element (r, n) of src (row stride ld) belongs to
// original line perm[n]; by_rows: the permuted index is the ROW (src is [n_perm, cols])
__global__ void unpermute_kernel(const double* __restrict__ src, long long ld, long long rows, long long cols,
                                 const int* __restrict__ perm, int n_perm, int by_rows, double* __restrict__ out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * cols) return;
  const long long r = e / cols, c = e - r * cols;
  const double v = src[r * ld + c];
  if (by_rows) out[(r < n_perm ? (long long)perm[r] : r) * cols + c] = v;
  else out[r * cols + (c < n_perm ? (long long)perm[c] : c)] = v;
}
int unpermute(const double* src, long long ld, long long rows, long long cols, const int* perm, int n_perm, bool by_rows,
              double* out, cudaStream_t st) {
  if (rows * cols == 0) return DKG_OK;
  unpermute_kernel<<<(unsigned)((rows * cols + 255) / 256), 256, 0, st>>>(src, ld, rows, cols, perm, n_perm, by_rows ? 1 : 0, out);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

int build_a0(const double* mu, int N, int M, const double* W, int S, double* A0, float* A0f, int ld,
             double* A0max, int* A0arg, cudaStream_t st) {
  dim3 grid(ceil_div(ld, 128), S);
  a0_kernel<<<grid, 128, 0, st>>>(mu, N, M, W, S, A0, A0f, ld);
  DKG_LAUNCH_CHECK();
  a0max_kernel<<<S, 256, 0, st>>>(A0, N, ld, A0max, A0arg);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

}  // namespace dkg

// ==========================================================================================
// Fast path of the plan preparation (n <= CHOL_FAST_MAX): blocked Cholesky in one CTA with the
// panel in shared memory, explicit inverse of the triangular factor by block columns, and the
// two big solves  K^-1 R = L^-T (L^-1 R)  as DMMA GEMMs (dkg_gemm.cu) instead of per-column
// substitution (which was 5-9 ms per solve at n = 400, N = 16384).
// ==========================================================================================
namespace dkg {

constexpr int CB = 32;                 // block size
constexpr int CHOL_FAST_MAX = 800;     // panel (n x 33 doubles) must fit in shared memory

// Lower Cholesky, in place, row-major n x n with leading dimension ld (>= n).  One CTA.
__global__ void __launch_bounds__(1024, 1)
cholesky_blocked_kernel(double* __restrict__ A, int n, int ld, int* __restrict__ info) {
  extern __shared__ __align__(16) double cs[];
  double* Ld = cs;                 // [CB][CB + 1] diagonal block
  double* Lp = cs + CB * (CB + 1); // [n][CB + 1]  panel below the diagonal block
  __shared__ int s_fail;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_fail = 0;
  __syncthreads();
  for (int k0 = 0; k0 < n; k0 += CB) {
    const int kb = min(CB, n - k0);
    for (int e = tid; e < kb * kb; e += nt) {
      const int i = e / kb, j = e - i * kb;
      Ld[i * (CB + 1) + j] = (j <= i) ? A[(size_t)(k0 + i) * ld + k0 + j] : 0.0;
    }
    __syncthreads();
    if (warp == 0) {  // unblocked factorisation of the kb x kb block, lane = row
      for (int j = 0; j < kb; ++j) {
        const double djj = Ld[j * (CB + 1) + j];
        if (!(djj > 0.0)) {
          if (lane == 0) s_fail = k0 + j + 1;
          break;
        }
        const double piv = sqrt(djj);
        __syncwarp();
        if (lane == j) Ld[j * (CB + 1) + j] = piv;
        if (lane > j && lane < kb) Ld[lane * (CB + 1) + j] /= piv;
        __syncwarp();
        if (lane > j && lane < kb) {
          const double lij = Ld[lane * (CB + 1) + j];
          for (int c = j + 1; c <= lane; ++c) Ld[lane * (CB + 1) + c] -= lij * Ld[c * (CB + 1) + j];
        }
        __syncwarp();
      }
    }
    __syncthreads();
    if (s_fail) break;
    for (int e = tid; e < kb * kb; e += nt) {
      const int i = e / kb, j = e - i * kb;
      A[(size_t)(k0 + i) * ld + k0 + j] = Ld[i * (CB + 1) + j];  // upper part of the block -> 0
    }
    const int m = n - k0 - kb;  // rows below
    // panel: row r of A[k0+kb.., k0..k0+kb) times Ld^-T ; one thread per row
    for (int r = tid; r < m; r += nt) {
      double x[CB];
      const double* src = A + (size_t)(k0 + kb + r) * ld + k0;
#pragma unroll
      for (int j = 0; j < CB; ++j) x[j] = j < kb ? src[j] : 0.0;
#pragma unroll
      for (int j = 0; j < CB; ++j) {
        if (j < kb) {
          double acc = x[j];
#pragma unroll
          for (int c = 0; c < CB; ++c)
            if (c < j) acc -= x[c] * Ld[j * (CB + 1) + c];
          x[j] = acc / Ld[j * (CB + 1) + j];
        }
      }
      double* dst = A + (size_t)(k0 + kb + r) * ld + k0;
#pragma unroll
      for (int j = 0; j < CB; ++j)
        if (j < kb) {
          dst[j] = x[j];
          Lp[r * (CB + 1) + j] = x[j];
        }
    }
    __syncthreads();
    // trailing update of the lower triangle: A22[i, c] -= Lp[i, :] . Lp[c, :]   (c <= i)
    const int mt = (m + 3) / 4;  // 4 x 4 register tiles
    const int ntiles = mt * (mt + 1) / 2;
    for (int t = tid; t < ntiles; t += nt) {
      // t -> (ti, tc) with tc <= ti:  ti = floor((sqrt(8t+1)-1)/2)
      int ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
      while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
      while (ti * (ti + 1) / 2 > t) --ti;
      const int tc = t - ti * (ti + 1) / 2;
      double acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
      for (int j = 0; j < kb; ++j) {
        double li[4], lc[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int ri = ti * 4 + a, rc = tc * 4 + a;
          li[a] = ri < m ? Lp[ri * (CB + 1) + j] : 0.0;
          lc[a] = rc < m ? Lp[rc * (CB + 1) + j] : 0.0;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] += li[a] * lc[b];
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int ri = ti * 4 + a, rc = tc * 4 + b;
          if (ri < m && rc <= ri) A[(size_t)(k0 + kb + ri) * ld + k0 + kb + rc] -= acc[a][b];
        }
    }
    __syncthreads();
  }
  if (tid == 0) info[0] = s_fail;
  __syncthreads();
  if (!s_fail)
    for (long long e = tid; e < (long long)n * n; e += nt) {
      const int i = (int)(e / n), k = (int)(e % n);
      if (k > i) A[(size_t)i * ld + k] = 0.0;
    }
}

int cholesky_blocked(double* A, int n, int ld, int* info_dev, cudaStream_t st) {
  const size_t smem = sizeof(double) * ((size_t)CB * (CB + 1) + (size_t)n * (CB + 1));
  DKG_CUDA_OK(cudaFuncSetAttribute(cholesky_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)(sizeof(double) * ((size_t)CB * (CB + 1) + (size_t)CHOL_FAST_MAX * (CB + 1)))));
  cholesky_blocked_kernel<<<1, 1024, smem, st>>>(A, n, ld, info_dev);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// X = L^-1 (lower triangular), by block columns: CTA jb computes X[:, jb-block].
//   X[jb][jb] = L[jb][jb]^-1 ;  X[ib][jb] = -L[ib][ib]^-1 * sum_{kb=jb}^{ib-1} L[ib][kb] X[kb][jb]
// Every CTA inverts the diagonal blocks it needs itself (cheap, avoids a second kernel).
__global__ void __launch_bounds__(1024, 1)
tri_inverse_kernel(const double* __restrict__ L, int n, int ldl, double* __restrict__ X, int ldx) {
  extern __shared__ __align__(16) double ts[];
  const int nb = (n + CB - 1) / CB;
  const int jb = blockIdx.x;
  double* Xs = ts;                              // [nb][CB][CB+1] blocks X[kb][jb] computed so far
  double* Dinv = Xs + (size_t)nb * CB * (CB + 1);  // [CB][CB+1] inverse of the current diagonal block
  double* Sb = Dinv + CB * (CB + 1);            // [CB][CB+1] accumulated product
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = tid >> 5, cidx = tid & 31;      // one output element per thread: (r, cidx)
  for (int ib = jb; ib < nb; ++ib) {
    const int rb = min(CB, n - ib * CB);        // rows of this block
    // ---- Dinv = L[ib][ib]^-1 : warp w solves column w (forward substitution, lane = row) ----
    {
      // column `warp` of the inverse: solve L d = e_warp ; serial over rows, done by lane 0..: small
      if (warp < rb) {
        double dcol = 0.0;  // value of row `lane`
        for (int i = 0; i < rb; ++i) {
          // row i: d_i = (delta_{i,warp} - sum_{k<i} L[i][k] d_k) / L[i][i] ; lanes hold d_k
          double part = (lane < i) ? L[(size_t)(ib * CB + i) * ldl + ib * CB + lane] * dcol : 0.0;
          part = warp_sum(part);
          const double di = ((i == warp ? 1.0 : 0.0) - part) / L[(size_t)(ib * CB + i) * ldl + ib * CB + i];
          if (lane == i) dcol = di;
        }
        if (lane < rb) Dinv[lane * (CB + 1) + warp] = dcol;
      }
    }
    // ---- S = sum_{kb=jb}^{ib-1} L[ib][kb] X[kb][jb] ----
    double acc = 0.0;
    const int cb = min(CB, n - jb * CB);        // columns of the block column
    for (int kb = jb; kb < ib; ++kb) {
      const int kk = min(CB, n - kb * CB);
      if (r < rb && cidx < cb) {
        const double* lrow = L + (size_t)(ib * CB + r) * ldl + kb * CB;
        const double* xcol = Xs + (size_t)kb * CB * (CB + 1) + cidx;
        for (int q = 0; q < kk; ++q) acc += lrow[q] * xcol[q * (CB + 1)];
      }
    }
    __syncthreads();  // Dinv ready
    if (ib > jb) Sb[r * (CB + 1) + cidx] = acc;
    __syncthreads();
    double v = 0.0;
    if (r < rb && cidx < cb) {
      if (ib == jb) v = Dinv[r * (CB + 1) + cidx];
      else {
        for (int q = 0; q < rb; ++q) v -= Dinv[r * (CB + 1) + q] * Sb[q * (CB + 1) + cidx];
      }
      Xs[(size_t)ib * CB * (CB + 1) + r * (CB + 1) + cidx] = v;
      X[(size_t)(ib * CB + r) * ldx + jb * CB + cidx] = v;
    }
    __syncthreads();
  }
}

int tri_inverse(const double* L, int n, int ldl, double* X, int ldx, cudaStream_t st) {
  const int nb = ceil_div(n, CB);
  const size_t smem = sizeof(double) * ((size_t)nb + 2) * CB * (CB + 1);
  const int nbmax = ceil_div(CHOL_FAST_MAX, CB);
  DKG_CUDA_OK(cudaFuncSetAttribute(tri_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)(sizeof(double) * ((size_t)nbmax + 2) * CB * (CB + 1))));
  tri_inverse_kernel<<<nb, 1024, smem, st>>>(L, n, ldl, X, ldx);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// y = A x  (rows x cols, row-major, one warp per row)
__global__ void matvec_kernel(const double* __restrict__ A, int lda, int rows, int cols,
                              const double* __restrict__ x, double* __restrict__ y) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  double acc = 0.0;
  for (int k = lane; k < cols; k += 32) acc += A[(size_t)row * lda + k] * x[k];
  acc = warp_sum(acc);
  if (lane == 0) y[row] = acc;
}

int matvec(const double* A, int lda, int rows, int cols, const double* x, double* y, cudaStream_t st) {
  if (rows == 0) return DKG_OK;
  matvec_kernel<<<ceil_div(rows * 32, 256), 256, 0, st>>>(A, lda, rows, cols, x, y);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

int chol_fast_max() { return CHOL_FAST_MAX; }

// ==========================================================================================
// Incremental refresh: one training point appended to one objective (dkg_plan_append_point).
// With K' = [[K, k], [k^T, kappa]], v = K^-1 k and s = kappa - k^T v (the Schur complement):
//   K'^-1 = [[K^-1 + v v^T / s, -v / s], [-v^T / s, 1 / s]]
//   alpha' = [alpha - beta v; beta],                 beta = (y - c - k^T alpha) / s
//   B' = K'^-1 [Kxd; r] = [B - v w^T; w^T],          w = (r - Kxd^T v) / s,  r = k(x_new, X_disc)
// i.e. O(n^2) for the training side and O(n N) for the discretisation side, instead of the
// O(n^3 + n^2 N) of a fresh plan.
// ==========================================================================================
__global__ void kernel_row_kernel(const double* __restrict__ xq, const double* __restrict__ pts, int npts, int d,
                                  int kind, double outputscale, double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npts) return;
  double sq = 0.0;
  for (int k = 0; k < d; ++k) {
    const double df = xq[k] - pts[(size_t)t * d + k];
    sq += df * df;
  }
  out[t] = stationary_from_sq(kind, outputscale, sq);
}
int kernel_row(const double* xq_dev, const double* pts, int npts, int d, int kind, double outputscale, double* out,
               cudaStream_t st) {
  if (npts == 0) return DKG_OK;
  kernel_row_kernel<<<ceil_div(npts, 128), 128, 0, st>>>(xq_dev, pts, npts, d, kind, outputscale, out);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// y_out = (y_in or 0) + alpha * A x ; one warp per row
__global__ void matvec_axpy_kernel(const double* __restrict__ A, int lda, int rows, int cols, const double* __restrict__ x,
                                   const double* __restrict__ y_in, double alpha, double* __restrict__ y_out) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  double acc = 0.0;
  for (int k = lane; k < cols; k += 32) acc += A[(size_t)row * lda + k] * x[k];
  acc = warp_sum(acc);
  if (lane == 0) y_out[row] = (y_in != nullptr ? y_in[row] : 0.0) + alpha * acc;
}
int matvec_axpy(const double* A, int lda, int rows, int cols, const double* x, const double* y_in, double alpha,
                double* y_out, cudaStream_t st) {
  if (rows == 0) return DKG_OK;
  matvec_axpy_kernel<<<ceil_div(rows * 32, 256), 256, 0, st>>>(A, lda, rows, cols, x, y_in, alpha, y_out);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// scal[0] = s = kappa - kv . v ; scal[1] = beta = (yc - kv . alpha) / s ; scal[2] = 1 / s   (one CTA)
__global__ void append_scalars_kernel(const double* __restrict__ kv, const double* __restrict__ v,
                                      const double* __restrict__ alpha, int n, double kappa, double yc,
                                      double* __restrict__ scal) {
  __shared__ double s1[8], s2[8];
  double a = 0.0, b = 0.0;
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    a += kv[t] * v[t];
    b += kv[t] * alpha[t];
  }
  a = warp_sum(a);
  b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = a; s2[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { ta += s1[k]; tb += s2[k]; }
    const double s = kappa - ta;
    scal[0] = s;
    scal[1] = (yc - tb) / s;
    scal[2] = 1.0 / s;
  }
}
int append_scalars(const double* kv, const double* v, const double* alpha, int n, double kappa, double yc, double* scal,
                   cudaStream_t st) {
  append_scalars_kernel<<<1, 256, 0, st>>>(kv, v, alpha, n, kappa, yc, scal);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

__global__ void append_update_k_kernel(double* __restrict__ Kinv, double* __restrict__ Kmat, int ld, int n,
                                       const double* __restrict__ kv, const double* __restrict__ v,
                                       const double* __restrict__ scal, double kappa) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j > n) return;
  const double rs = scal[2];
  if (i < n && j < n) {
    Kinv[(size_t)i * ld + j] += v[i] * v[j] * rs;
  } else if (i == n && j == n) {
    Kinv[(size_t)n * ld + n] = rs;
    Kmat[(size_t)n * ld + n] = kappa;
  } else if (i == n) {
    Kinv[(size_t)n * ld + j] = -v[j] * rs;
    Kmat[(size_t)n * ld + j] = kv[j];
  } else {  // j == n
    Kinv[(size_t)i * ld + n] = -v[i] * rs;
    Kmat[(size_t)i * ld + n] = kv[i];
  }
}
int append_update_k(double* Kinv, double* Kmat, int ld, int n, const double* kv, const double* v, const double* scal,
                    double kappa, cudaStream_t st) {
  dim3 grid(ceil_div(n + 1, 128), n + 1);
  append_update_k_kernel<<<grid, 128, 0, st>>>(Kinv, Kmat, ld, n, kv, v, scal, kappa);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

__global__ void append_update_alpha_kernel(double* __restrict__ alpha, const double* __restrict__ v,
                                           const double* __restrict__ scal, int n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) alpha[t] -= scal[1] * v[t];
  else if (t == n) alpha[n] = scal[1];
}
int append_update_alpha(double* alpha, const double* v, const double* scal, int n, cudaStream_t st) {
  append_update_alpha_kernel<<<ceil_div(n + 1, 128), 128, 0, st>>>(alpha, v, scal, n);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// w[j] = (r[j] - sum_t v[t] Kxd[t, j]) / s, and the new row of Kxd
__global__ void append_w_kernel(double* __restrict__ Kxd, int ld, int n, int N, const double* __restrict__ v,
                                const double* __restrict__ rrow, const double* __restrict__ scal, double* __restrict__ w) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  double acc = 0.0;
  for (int t = 0; t < n; ++t) acc = fma(v[t], Kxd[(size_t)t * ld + j], acc);
  w[j] = (rrow[j] - acc) * scal[2];
  Kxd[(size_t)n * ld + j] = rrow[j];
}
int append_w(double* Kxd, int ld, int n, int N, const double* v, const double* rrow, const double* scal, double* w,
             cudaStream_t st) {
  append_w_kernel<<<ceil_div(N, 128), 128, 0, st>>>(Kxd, ld, n, N, v, rrow, scal, w);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// B[t, j] -= v[t] w[j] (t < n), B[n, j] = w[j];  BT likewise (transposed copy for the backward's row gathers)
__global__ void append_update_b_kernel(double* __restrict__ B, int ldb, double* __restrict__ BT, int ldbt, int n, int N,
                                       const double* __restrict__ v, const double* __restrict__ w) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
  if (j >= N) return;
  if (t < n) {
    const double nv = B[(size_t)t * ldb + j] - v[t] * w[j];
    B[(size_t)t * ldb + j] = nv;
    BT[(size_t)j * ldbt + t] = nv;
  } else {
    B[(size_t)n * ldb + j] = w[j];
    BT[(size_t)j * ldbt + n] = w[j];
  }
}
int append_update_b(double* B, int ldb, double* BT, int ldbt, int n, int N, const double* v, const double* w,
                    cudaStream_t st) {
  dim3 grid(ceil_div(N, 128), n + 1);
  append_update_b_kernel<<<grid, 128, 0, st>>>(B, ldb, BT, ldbt, n, N, v, w);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

}  // namespace dkg
