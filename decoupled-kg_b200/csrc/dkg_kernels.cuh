// Host-side launch interfaces shared by the translation units of libdkg_b200.
#pragma once

#include "dkg_plan.cuh"

namespace dkg {

// ---- dkg_prepare.cu --------------------------------------------------------------------------
int scale_rows(const double* x, int rows, int d, const double* ls_host, double* out, cudaStream_t st);
int kmat_train(const ObjState& o, int d, double jitter, double* K, int ld, cudaStream_t st);
int cholesky_inplace(double* A, int n, int* info_dev, cudaStream_t st);
int transpose(const double* in, int rows, int cols, int ld_in, double* out, int ld_out, cudaStream_t st);
int cholesky_solve_inplace(const double* L, const double* LT, int n, double* R, int ncols, int ld,
                           cudaStream_t st);
int kcross(const ObjState& o, const double* xd_s, int N, int d, double* R, int ld, cudaStream_t st);
int set_identity(double* A, int n, int ld, cudaStream_t st);
int residual(const double* y, int n, double c, double* out, cudaStream_t st);
int mu_disc(const double* xd, int N, int d, const ObjState& o, double* mu, int M, int m, cudaStream_t st);
int build_a0(const double* mu, int N, int M, const double* W, int S, double* A0, float* A0f, int ld,
             double* A0max, int* A0arg, cudaStream_t st);

int build_a0_tilemax(const float* A0f, int ld, int S, int tile, int ntiles, float* out, int* out_arg, cudaStream_t st);
int gather_rows(const double* in, const int* perm, int rows, int d, double* out, cudaStream_t st);
int unpermute(const double* src, long long ld, long long rows, long long cols, const int* perm, int n_perm, bool by_rows,
              double* out, cudaStream_t st);
int cholesky_blocked(double* A, int n, int ld, int* info_dev, cudaStream_t st);
int tri_inverse(const double* L, int n, int ldl, double* X, int ldx, cudaStream_t st);
int matvec(const double* A, int lda, int rows, int cols, const double* x, double* y, cudaStream_t st);
int chol_fast_max();
// incremental refresh (dkg_plan_append_point)
int kernel_row(const double* xq_dev, const double* pts, int npts, int d, int kind, double outputscale, double* out,
               cudaStream_t st);
int matvec_axpy(const double* A, int lda, int rows, int cols, const double* x, const double* y_in, double alpha,
                double* y_out, cudaStream_t st);
int append_scalars(const double* kv, const double* v, const double* alpha, int n, double kappa, double yc, double* scal,
                   cudaStream_t st);
int append_update_k(double* Kinv, double* Kmat, int ld, int n, const double* kv, const double* v, const double* scal,
                    double kappa, cudaStream_t st);
int append_update_alpha(double* alpha, const double* v, const double* scal, int n, cudaStream_t st);
int append_w(double* Kxd, int ld, int n, int N, const double* v, const double* rrow, const double* scal, double* w,
             cudaStream_t st);
int append_update_b(double* B, int ldb, double* BT, int ldbt, int n, int N, const double* v, const double* w,
                    cudaStream_t st);

// ---- dkg_gemm.cu -----------------------------------------------------------------------------
struct CovEpilogue {
  const double* xs;    // [rows, d] candidates / lengthscale_i
  const double* xd_s;  // [N_pad, d] discretisation / lengthscale_i
  const double* sd;    // [rows]     sqrt(noisy predictive variance), un-standardised
  double* Z;           // [rows, ldz]
  int ldz;
  int C;               // valid rows
  int N;               // valid columns
  int d;
  int kind;
  double outputscale;
  double ystd2;
};
int gemm_store(const double* A, int lda, const double* B, int ldb, int M_pad, int N_pad, int K,
               double* D, int ldd, cudaStream_t st);
int gemm_axpy(const double* A, int lda, const double* B, int ldb, int M_pad, int N_pad, int K,
              const double* Cin, int ldc, double alpha, double* D, int ldd, cudaStream_t st);
int gemm_cov(const double* KX, int lda, const double* B, int ldb, int M_pad, int N_pad, int K,
             const CovEpilogue& ep, cudaStream_t st);

// ---- dkg_ozaki.cu ----------------------------------------------------------------------------
constexpr int OZ_DEFAULT_DIGITS = 7;
constexpr int OZ_DEFAULT_DIAGONALS = 7;  // balanced digits: the dropped diagonals are zero-mean (see dkg_ozaki.cu)
constexpr int OZ_MAX_K = 4096;
int ozaki_kp(int K);
size_t ozaki_digit_bytes(int rows_pad, int K, int NS);
// rows per digit block of the B operand: 64 when the CTA-pair kernel serves this configuration, else 128
int ozaki_b_block_rows(int NS, int NG);
int ozaki_slice_rows(const double* X, int ld, int rows, int K, int block_rows, int NS, unsigned char* digits,
                     double* scale, cudaStream_t st);
int ozaki_cov(const unsigned char* a_digits, const double* sa, int M_pad, const unsigned char* b_digits,
              const double* sb, int N_pad, int K, int NS, int NG, bool b_nonneg, const CovEpilogue& ep,
              cudaStream_t st);
int ozaki_store(const unsigned char* a_digits, const double* sa, int M_pad, const unsigned char* b_digits,
                const double* sb, int N_pad, int K, int NS, int NG, bool b_nonneg, double* D, int ldd, int M,
                int N, cudaStream_t st);
int ozaki_store_axpy(const unsigned char* a_digits, const double* sa, int M_pad, const unsigned char* b_digits,
                     const double* sb, int N_pad, int K, int NS, int NG, const double* Cin, int ldc, double alpha,
                     double* D, int ldd, int M, int N, cudaStream_t st);
// b_nonneg: every entry of the B operand is >= 0 (true for kernel values), which lets all its digit planes be
// multiplied as UINT8 and pairs of them be merged into N = 256 instructions
int ozaki_mma_peak(int M_pad, int N_pad, int K, int NS, int NG, int reps, int mode, double* tops_out, double* ms_out,
                   cudaStream_t st);
int ozaki_any_negative(const double* X, int ld, int rows, int cols, int* flag_dev, cudaStream_t st);

// ---- dkg_forward.cu --------------------------------------------------------------------------
struct XprepArgs {
  const double* X;  // [C, d]
  int C, d, M, S, target;
  const double* xs[MAX_M];
  const double* alpha[MAX_M];
  int ntr[MAX_M];
  int kind[MAX_M];
  double outputscale[MAX_M], mean_const[MAX_M], y_mean[MAX_M], y_std[MAX_M];
  double ls[MAX_M][MAX_D];
  const double* W;  // [S, M]
  double* Xs;       // [C, d] candidates / lengthscale_target
  double* KX;       // [C_pad, n_pad]
  int n_pad;
  double* a_new;    // [C, S]
};
int launch_xprep(const XprepArgs& p, cudaStream_t st);
int launch_mean(const XprepArgs& p, double* mu_out, cudaStream_t st);
int launch_var(const double* KX, int n_pad, const double* T, int ldk, int ntr, int C, int kind,
               double outputscale, double noise, double ystd2, double* var, double* sd,
               double* zown, cudaStream_t st);
int launch_place_own(const double* zown, int rows, double* Z, int ldz, int N, cudaStream_t st);
int launch_batched_cholesky_solve(const double* L, int n, const double* KX, int n_pad, int C,
                                  double* T, int ldk, cudaStream_t st);
int batched_solve_max_n();
int launch_xprep_scaled(const double* X, int C, int d, const double* ls_host, double* Xs, cudaStream_t st);
int fill_ones(double* p, int n, cudaStream_t st);
int place_latent_var(const double* KX, int n_pad, const double* T, int ldk, int ntr, int C, int kind,
                     double outputscale, double ystd2, double* COV, int ldz, int N, cudaStream_t st);

}  // namespace dkg
