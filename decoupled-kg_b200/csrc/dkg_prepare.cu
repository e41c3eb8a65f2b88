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
                                  double outputscale, double noise, double* __restrict__ K) {
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
  K[(size_t)t * n + u] = v;
}

int kmat_train(const ObjState& o, int d, double jitter, double* K, cudaStream_t st) {
  dim3 grid(ceil_div(o.n, 128), o.n);
  kmat_train_kernel<<<grid, 128, 0, st>>>(o.xs, o.n, d, o.kernel, o.outputscale,
                                          o.noise + jitter, K);
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
                          int S, double* __restrict__ A0, int ld) {
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

int build_a0(const double* mu, int N, int M, const double* W, int S, double* A0, int ld,
             double* A0max, int* A0arg, cudaStream_t st) {
  dim3 grid(ceil_div(ld, 128), S);
  a0_kernel<<<grid, 128, 0, st>>>(mu, N, M, W, S, A0, ld);
  DKG_LAUNCH_CHECK();
  a0max_kernel<<<S, 256, 0, st>>>(A0, N, ld, A0max, A0arg);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

}  // namespace dkg
