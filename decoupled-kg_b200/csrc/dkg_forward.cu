// Candidate-side GP conditioning that precedes the big contraction:
//   xprep : k_m(x_c, X_train_m) for every objective, posterior means mu_m(x_c), the scalarised
//           intercept of the candidate's own line, and the scaled candidate coordinates
//           (replaces the per-candidate posterior calls at discretekg.py:275-284 for the row of
//           xnew itself).
//   var   : noisy predictive variance  k(x,x) - k_x^T K^-1 k_x + noise  (discretekg.py:302) and
//           the slope of the candidate's own line  Cov(x, x) / sd  (discretekg.py:313, entry 0).
#include "dkg_kernels.cuh"

namespace dkg {


__global__ void __launch_bounds__(128) xprep_kernel(XprepArgs p) {
  __shared__ double s_part[4];
  __shared__ double s_mu[MAX_M];
  const int c = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int m = 0; m < p.M; ++m) {
    double xm[MAX_D];
#pragma unroll
    for (int k = 0; k < MAX_D; ++k) xm[k] = k < p.d ? p.X[(size_t)c * p.d + k] / p.ls[m][k] : 0.0;
    if (m == p.target && threadIdx.x < p.d) p.Xs[(size_t)c * p.d + threadIdx.x] = xm[threadIdx.x];
    double acc = 0.0;
    for (int t = threadIdx.x; t < p.ntr[m]; t += blockDim.x) {
      double sq = 0.0;
#pragma unroll
      for (int k = 0; k < MAX_D; ++k)
        if (k < p.d) {
          double df = xm[k] - p.xs[m][(size_t)t * p.d + k];
          sq += df * df;
        }
      const double kv = stationary_from_sq(p.kind[m], p.outputscale[m], sq);
      if (m == p.target) p.KX[(size_t)c * p.n_pad + t] = kv;
      acc += kv * p.alpha[m][t];
    }
    acc = warp_sum(acc);
    if (lane == 0) s_part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = ((s_part[0] + s_part[1]) + (s_part[2] + s_part[3]));
      s_mu[m] = (p.mean_const[m] + tot) * p.y_std[m] + p.y_mean[m];
    }
    __syncthreads();
  }
  // intercept of the candidate's own line: sum_m W[j, m] mu_m(x)  (discretekg.py:320, row 0)
  for (int j = threadIdx.x; j < p.S; j += blockDim.x) {
    double a = __dmul_rn(p.W[j * p.M + 0], s_mu[0]);
    for (int m = 1; m < p.M; ++m) a = __dadd_rn(a, __dmul_rn(p.W[j * p.M + m], s_mu[m]));
    p.a_new[(size_t)c * p.S + j] = a;
  }
}

// posterior means only (one CTA per point)
__global__ void __launch_bounds__(128) mean_kernel(XprepArgs p, double* __restrict__ mu_out) {
  __shared__ double s_part[4];
  const int c = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int m = 0; m < p.M; ++m) {
    double xm[MAX_D];
#pragma unroll
    for (int k = 0; k < MAX_D; ++k) xm[k] = k < p.d ? p.X[(size_t)c * p.d + k] / p.ls[m][k] : 0.0;
    double acc = 0.0;
    for (int t = threadIdx.x; t < p.ntr[m]; t += blockDim.x) {
      double sq = 0.0;
#pragma unroll
      for (int k = 0; k < MAX_D; ++k)
        if (k < p.d) {
          double df = xm[k] - p.xs[m][(size_t)t * p.d + k];
          sq += df * df;
        }
      acc += stationary_from_sq(p.kind[m], p.outputscale[m], sq) * p.alpha[m][t];
    }
    acc = warp_sum(acc);
    if (lane == 0) s_part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = ((s_part[0] + s_part[1]) + (s_part[2] + s_part[3]));
      mu_out[(size_t)c * p.M + m] = (p.mean_const[m] + tot) * p.y_std[m] + p.y_mean[m];
    }
    __syncthreads();
  }
}

int launch_mean(const XprepArgs& p, double* mu_out, cudaStream_t st) {
  if (p.C == 0) return DKG_OK;
  mean_kernel<<<p.C, 128, 0, st>>>(p, mu_out);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

int launch_xprep(const XprepArgs& p, cudaStream_t st) {
  if (p.C == 0) return DKG_OK;
  xprep_kernel<<<p.C, 128, 0, st>>>(p);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// one warp per candidate
__global__ void var_kernel(const double* __restrict__ KX, int n_pad, const double* __restrict__ T,
                           int ldk, int ntr, int C, int kind, double outputscale, double noise,
                           double ystd2, double* __restrict__ var, double* __restrict__ sd,
                           double* __restrict__ zown) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  double acc = 0.0;
  for (int t = lane; t < ntr; t += 32) acc += KX[(size_t)c * n_pad + t] * T[(size_t)c * ldk + t];
  acc = warp_sum(acc);
  if (lane == 0) {
    const double kxx = stationary_from_sq(kind, outputscale, 0.0);
    const double var_lat = kxx - acc;                 // Cov(x, x), noise free   (:301, entry 0)
    const double v = (var_lat + noise) * ystd2;       // observation_noise=True  (:302)
    const double s = sqrt(v);
    var[c] = v;
    sd[c] = s;
    zown[c] = (var_lat * ystd2) / s;                  // znew_coefficients[0]    (:313)
  }
}

int launch_var(const double* KX, int n_pad, const double* T, int ldk, int ntr, int C, int kind,
               double outputscale, double noise, double ystd2, double* var, double* sd,
               double* zown, cudaStream_t st) {
  if (C == 0) return DKG_OK;
  const int threads = 256;
  var_kernel<<<ceil_div(C * 32, threads), threads, 0, st>>>(KX, n_pad, T, ldk, ntr, C, kind,
                                                            outputscale, noise, ystd2, var, sd, zown);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// Z[r, N] = zown[r] for the rows of one chunk (the candidate's own line lives in column N)
__global__ void place_own_kernel(const double* __restrict__ zown, int rows, double* __restrict__ Z,
                                 int ldz, int N) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows) Z[(size_t)r * ldz + N] = zown[r];
}

int launch_place_own(const double* zown, int rows, double* Z, int ldz, int N, cudaStream_t st) {
  if (rows == 0) return DKG_OK;
  place_own_kernel<<<ceil_div(rows, 128), 128, 0, st>>>(zown, rows, Z, ldz, N);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// ---- helpers of the coupled path ------------------------------------------------------------
struct LsArgF { double v[MAX_D]; };

__global__ void xscale_kernel(const double* __restrict__ X, int C, int d, LsArgF ls, double* __restrict__ Xs) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C * d) Xs[i] = X[i] / ls.v[i % d];
}
int launch_xprep_scaled(const double* X, int C, int d, const double* ls_host, double* Xs, cudaStream_t st) {
  if (C == 0) return DKG_OK;
  LsArgF a;
  for (int k = 0; k < MAX_D; ++k) a.v[k] = k < d ? ls_host[k] : 1.0;
  xscale_kernel<<<ceil_div(C * d, 256), 256, 0, st>>>(X, C, d, a, Xs);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

__global__ void fill_ones_kernel(double* __restrict__ p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 1.0;
}
int fill_ones(double* p, int n, cudaStream_t st) {
  if (n == 0) return DKG_OK;
  fill_ones_kernel<<<ceil_div(n, 256), 256, 0, st>>>(p, n);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// COV[c, N] = (k(x,x) - KX[c,:] . T[c,:]) * ystd2  (Cov_m(x_c, x_c), noise free); warp per candidate
__global__ void place_latent_var_kernel(const double* __restrict__ KX, int n_pad, const double* __restrict__ T,
                                        int ldk, int ntr, int C, int kind, double outputscale,
                                        double ystd2, double* __restrict__ COV, int ldz, int N) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  double acc = 0.0;
  for (int t = lane; t < ntr; t += 32) acc += KX[(size_t)c * n_pad + t] * T[(size_t)c * ldk + t];
  acc = warp_sum(acc);
  if (lane == 0) COV[(size_t)c * ldz + N] = (stationary_from_sq(kind, outputscale, 0.0) - acc) * ystd2;
}
int place_latent_var(const double* KX, int n_pad, const double* T, int ldk, int ntr, int C, int kind,
                     double outputscale, double ystd2, double* COV, int ldz, int N, cudaStream_t st) {
  if (C == 0) return DKG_OK;
  place_latent_var_kernel<<<ceil_div(C * 32, 256), 256, 0, st>>>(KX, n_pad, T, ldk, ntr, C, kind,
                                                                outputscale, ystd2, COV, ldz, N);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

}  // namespace dkg

// ==========================================================================================
// Backward-stable batched solve  T[c, :] = K^-1 k(X_train, x_c)  through the Cholesky factor
// (forward + back substitution, blocked by 32 rows), 32 candidates per CTA with their
// right-hand sides resident in shared memory.
//
// Why not T = KX @ Kinv with an explicit inverse (one GEMM)?  The explicit inverse has huge
// columns when K is ill conditioned and its rounding errors do not cancel in
// k_x^T Kinv k_x: on the c2 problem (cond 4.6e7) the predictive variance came out with a relative
// error of 2e-4, the slopes with 1e-6.  A backward-stable solve gives (K + E) w = k_x with
// |E| ~ eps |K|, and the quantities the path needs (k_x . w and Kxd[:, n] . w) then carry errors
// ~ eps |K| |w|^2 -- 1e-13 on the same problem, which is what LAPACK-based GPyTorch achieves.
// ==========================================================================================
namespace dkg {

constexpr int TS_NB = 32;       // block rows
constexpr int TS_COLS = 32;     // candidates per CTA
constexpr int TS_LDY = TS_COLS + 1;

__global__ void __launch_bounds__(1024, 1)
batched_cholesky_solve_kernel(const double* __restrict__ L, int n, const double* __restrict__ KX,
                              int n_pad, int C, double* __restrict__ T, int ldk) {
  extern __shared__ __align__(16) double ys[];  // [n][TS_LDY] right-hand sides / solution
  __shared__ double Ld[TS_NB][TS_NB + 1];      // current diagonal block of L
  const int tid = threadIdx.x;
  const int r = tid >> 5;        // row inside the block (0..31) == warp id
  const int c = tid & 31;        // candidate inside the tile == lane
  const int c0 = blockIdx.x * TS_COLS;
  const int nb = (n + TS_NB - 1) / TS_NB;
  // load: ys[t][c] = KX[c0 + c][t]   (each warp reads 32 consecutive t of one candidate)
  for (int e = tid; e < n * TS_COLS; e += blockDim.x) {
    const int cc = e / n, t = e - cc * n;
    ys[t * TS_LDY + cc] = (c0 + cc < C) ? KX[(size_t)(c0 + cc) * n_pad + t] : 0.0;
  }
  __syncthreads();
  // ---- forward: L y = b ----
  for (int kb = 0; kb < nb; ++kb) {
    const int rows = min(TS_NB, n - kb * TS_NB);
    const int row = kb * TS_NB + r;
    double acc = 0.0;
    if (r < rows) {
      const double* Lrow = L + (size_t)row * n;
      const int kmax = kb * TS_NB;
      for (int j = 0; j < kmax; ++j) acc += Lrow[j] * ys[j * TS_LDY + c];
    }
    if (r < rows) Ld[r][c] = (c < rows) ? L[(size_t)row * n + kb * TS_NB + c] : 0.0;
    __syncthreads();
    if (r < rows) ys[row * TS_LDY + c] -= acc;
    __syncthreads();
    // substitution inside the diagonal block: warp w handles candidate w, lane = row
    {
      const int cand = r, lr = c;  // reuse: warp id -> candidate, lane -> row
      double v = (lr < rows) ? ys[(kb * TS_NB + lr) * TS_LDY + cand] : 0.0;
      for (int i = 0; i < rows; ++i) {
        const double yi = __shfl_sync(0xffffffffu, v, i) / Ld[i][i];
        if (lr == i) v = yi;
        else if (lr > i && lr < rows) v -= Ld[lr][i] * yi;
      }
      if (lr < rows) ys[(kb * TS_NB + lr) * TS_LDY + cand] = v;
    }
    __syncthreads();
  }
  // ---- backward: L^T t = y ----
  for (int kb = nb - 1; kb >= 0; --kb) {
    const int rows = min(TS_NB, n - kb * TS_NB);
    const int row = kb * TS_NB + r;
    double acc = 0.0;
    if (r < rows) {
      for (int j = (kb + 1) * TS_NB; j < n; ++j) acc += L[(size_t)j * n + row] * ys[j * TS_LDY + c];
    }
    if (r < rows) Ld[r][c] = (c < rows) ? L[(size_t)row * n + kb * TS_NB + c] : 0.0;
    __syncthreads();
    if (r < rows) ys[row * TS_LDY + c] -= acc;
    __syncthreads();
    {
      const int cand = r, lr = c;
      double v = (lr < rows) ? ys[(kb * TS_NB + lr) * TS_LDY + cand] : 0.0;
      for (int i = rows - 1; i >= 0; --i) {
        const double ti = __shfl_sync(0xffffffffu, v, i) / Ld[i][i];
        if (lr == i) v = ti;
        else if (lr < i) v -= Ld[i][lr] * ti;  // (L^T)[lr][i] = L[i][lr]
      }
      if (lr < rows) ys[(kb * TS_NB + lr) * TS_LDY + cand] = v;
    }
    __syncthreads();
  }
  for (int e = tid; e < n * TS_COLS; e += blockDim.x) {
    const int cc = e / n, t = e - cc * n;
    if (c0 + cc < C) T[(size_t)(c0 + cc) * ldk + t] = ys[t * TS_LDY + cc];
  }
}

int batched_solve_max_n() { return 800; }

int launch_batched_cholesky_solve(const double* L, int n, const double* KX, int n_pad, int C,
                                  double* T, int ldk, cudaStream_t st) {
  if (C == 0) return DKG_OK;
  const size_t smem = sizeof(double) * (size_t)n * TS_LDY;
  DKG_CUDA_OK(cudaFuncSetAttribute(batched_cholesky_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)(sizeof(double) * (size_t)batched_solve_max_n() * TS_LDY)));
  batched_cholesky_solve_kernel<<<ceil_div(C, TS_COLS), 1024, smem, st>>>(L, n, KX, n_pad, C, T, ldk);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

}  // namespace dkg
