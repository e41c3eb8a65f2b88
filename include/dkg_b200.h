/*
 * dkg_b200.h -- C-ABI of libdkg_b200.so: the B200 (sm_100a) implementation of the discrete
 * knowledge-gradient hot path of quasirandom/decoupled-kg.
 *
 * Plain C types only (pointers, sizes, doubles); no torch / C++ types cross this boundary.
 * Reference = /root/reference (paths below are relative to it).  Each entry point names the
 * reference interface it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, a negative DKG_E* code on failure; the message of the
 *     last failure on the calling thread is available from dkg_last_error().
 *   - "dev" pointers are CUDA device pointers on the current device; "host" pointers are ordinary
 *     host memory.  `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - all matrices are row-major float64.
 *   - the library never frees or retains caller buffers; everything it allocates belongs to the
 *     plan and is released by dkg_plan_destroy().
 *   - kernels are enqueued on `stream` and the *_dev entry points do not synchronise the host
 *     (except for a first-use workspace allocation).
 */
#ifndef DKG_B200_H
#define DKG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DKG_API __attribute__((visibility("default")))
#else
#define DKG_API
#endif

#define DKG_ABI_VERSION 4

/* error codes */
#define DKG_OK 0
#define DKG_EINVAL (-1)   /* bad argument (shape, NULL, unsupported kernel id, ...)            */
#define DKG_ECUDA (-2)    /* a CUDA runtime call failed                                        */
#define DKG_ENOTPD (-3)   /* training covariance not positive definite even with 1e-6 jitter   */
#define DKG_ENOMEM (-4)
#define DKG_EEMPTY (-5)   /* zero lines given to the expected-max stage (ValueError upstream)  */
#define DKG_ECAPACITY (-7) /* dkg_plan_append_point: the plan's buffers have no room for another training point
                             of that objective (or the plan was built by the large-n path): build a new plan   */
#define DKG_ETRUNC (-6)   /* hull records lost (spill pool pinned too small by DKG_SPILL_BLOCKS): values are
                             exact, the gradient rows concerned are NaN (dkg_forward_host only; the device
                             entry point reports it through dkg_plan_stats()[6] and the NaN rows)          */

/* stationary kernels of the reference's model factory
 * (src/decoupledbo/modules/model/factory.py:116-135: ScaleKernel(MaternKernel(nu=2.5)|RBFKernel)) */
#define DKG_KERNEL_MATERN52 0
#define DKG_KERNEL_RBF 1

/* One objective of the ModelListGP (factory.py:63-88), in *model* space. */
typedef struct dkg_objective {
  const double* train_x_dev; /* [n, d]  unit-cube inputs (factory.py:65)                       */
  const double* train_y_dev; /* [n]     targets (standardised if an outcome transform is used) */
  int32_t n;                 /* training points of THIS objective (ragged across objectives,
                                src/decoupledbo/pipeline/nodes/bo_loop.py:403-405)             */
  int32_t kernel;            /* DKG_KERNEL_*                                                   */
  const double* lengthscale_host; /* [d] ARD lengthscales                                      */
  double outputscale;        /* ScaleKernel.outputscale                                        */
  double mean_const;         /* ConstantMean.constant                                          */
  double noise;              /* GaussianLikelihood.noise (variance)                            */
  double y_mean;             /* Standardize.means (0 if no outcome transform)                  */
  double y_std;              /* Standardize.stdvs (1 if no outcome transform)                  */
} dkg_objective;

typedef struct dkg_plan dkg_plan; /* opaque: candidate-independent state of one acquisition fn */

DKG_API int dkg_abi_version(void);
DKG_API const char* dkg_last_error(void);

/*
 * dkg_plan_create -- replaces DiscreteKnowledgeGradient.__init__
 *   (src/decoupledbo/modules/acquisition/discretekg.py:62-123) PLUS everything the reference
 *   recomputes for every candidate although it does not depend on the candidate: the training
 *   Cholesky and mean cache, the posterior means at the discretisation, the scalarised intercept
 *   table, and K^-1 k(X_train, X_disc) (discretekg.py:275-300 via GPyTorch's exact prediction).
 *
 *   objs[M]        GP state per objective
 *   d              input dimension
 *   x_disc_dev     [N, d] discretisation (discretekg.py:121)
 *   weights_host   [S, M] scalarisation weights (discretekg.py:122)
 *   target_ix      objective whose observation is fantasised (discretekg.py:123, decoupled path
 *                  discretekg.py:238-338), in [0, M); or -1 for the coupled evaluation in which all
 *                  objectives are observed together (target_output_ix=None, discretekg.py:162-235)
 *   flags          0, or DKG_PLAN_* bits
 */
#define DKG_PLAN_DEFAULT 0u
/* Reduced-precision mode: the covariance contraction (the O(C n N) term) keeps 4 base-256 digits per
 * operand (10 int8 digit products instead of 28), i.e. ~2^-31 of every row's scale -- float32-class
 * accuracy for that term, everything else stays float64.  Stated tolerance of the KG values in
 * this mode: |dKG| <= 1e-4 |KG| + 1e-7 max|intercept| (tests/test_gpu_fast_mode.py). */
#define DKG_PLAN_FAST32 1u
DKG_API int dkg_plan_create(const dkg_objective* objs, int32_t M, int32_t d, const double* x_disc_dev,
                    int32_t N, const double* weights_host, int32_t S, int32_t target_ix,
                    uint32_t flags, void* stream, dkg_plan** out_plan);

DKG_API void dkg_plan_destroy(dkg_plan* plan);

/*
 * dkg_plan_append_point -- incremental refresh of a plan when ONE observation is added to ONE objective
 *   with unchanged hyper-parameters: what the reference's BO loop does between iterations
 *   (src/decoupledbo/pipeline/nodes/bo_loop.py:403-405 appends the point, :457-465 rebuilds the model;
 *   with --fit-hyperparams=never/once the hyper-parameters stay fixed, :574-589).  Block-inverse update of
 *   (K + noise I)^-1, the mean cache, K^-1 k(X_train, X_disc) and the digit planes, then the posterior
 *   means at the discretisation and the intercept table: O(n^2 + n N) instead of the O(n^3 + n^2 N) of
 *   dkg_plan_create.  Works for decoupled and coupled plans and for any objective m
 *   (the fantasised one or not; objectives may have different numbers of points).
 *
 *   m        objective that received the observation, in [0, M)
 *   x_host   [d] the new input (unit cube, as train_x)
 *   y        the new target in the same space as train_y (standardised if an outcome transform is used)
 *   returns DKG_ECAPACITY when there is no room left (rows are padded to a multiple of 128, so a plan
 *   built with n points takes up to 128 - n % 128 appends) and DKG_ENOTPD when the extended covariance is
 *   not positive definite; the plan is unchanged in both cases.  Synchronises the stream.
 */
DKG_API int dkg_plan_append_point(dkg_plan* plan, int32_t m, const double* x_host, double y, void* stream);

/*
 * dkg_forward_dev -- replaces DiscreteKnowledgeGradient.forward (discretekg.py:131-159), i.e. the
 *   Python loop over candidates calling calculate_discrete_kg_conditioning_on_single_output
 *   (discretekg.py:238-338), and -- when dX_dev != NULL -- the autograd backward through it
 *   (driven by botorch gen_candidates_scipy from
 *   src/decoupledbo/modules/acquisition_optimisation_strategy.py:217-224).
 *
 *   X_dev   [C, d] candidates (the t-batch flattened, q == 1)
 *   kg_dev  [C]    out: KG value per candidate
 *   dX_dev  [C, d] out (optional): d KG[c] / d X[c, :]
 */
DKG_API int dkg_forward_dev(dkg_plan* plan, const double* X_dev, int32_t C, double* kg_dev,
                    double* dX_dev, void* stream);

/* Same with HOST buffers: H2D copy of X, kernels, D2H copy of kg (and dX), then a stream sync.
 * This is the call timed as the end-to-end ("e2e") number. */
DKG_API int dkg_forward_host(dkg_plan* plan, const double* X_host, int32_t C, double* kg_host,
                     double* dX_host, void* stream);

/*
 * dkg_posterior_mean_dev -- posterior means of all objectives at a batch of points,
 *   mu[c, m] = (c_m + k_m(x_c, X_train_m) . alpha_m) * y_std_m + y_mean_m,
 *   i.e. what model.posterior(X).mean returns (the intercept side of discretekg.py:300), exposed
 *   because the reference's metrics evaluate it for whole NSGA-II populations
 *   (src/decoupledbo/modules/pareto/sample.py:138-144 BoTorchModel.batch_fitness).
 *   X_dev [C, d] -> mu_dev [C, M]
 */
DKG_API int dkg_posterior_mean_dev(dkg_plan* plan, const double* X_dev, int32_t C, double* mu_dev,
                                   void* stream);

/*
 * dkg_expected_max_lines_dev -- replaces, for P independent sets of L lines each,
 *   calculate_epigraph_indices (discretekg.py:341-412) followed by
 *   calculate_expected_value_of_piecewise_linear_function (discretekg.py:415-452):
 *   E[max_n (a[p,n] + b[p,n] Z)], Z ~ N(0,1).
 *
 *   a_dev, b_dev   [P, L] intercepts / slopes
 *   emax_dev       [P]    out: expectation
 *   hull_count_dev [P]    out (optional): number of lines on the upper envelope
 *   hull_idx_dev   [P, hull_cap] out (optional): their indices, left to right
 *   hull_x_dev     [P, hull_cap] out (optional): the hull_count-1 intersections
 *   dE_da_dev, dE_db_dev [P, L] out (optional): gradient of the expectation (zero off the hull)
 *   L == 0 -> DKG_EEMPTY (the reference raises ValueError, discretekg.py:466-470)
 */
DKG_API int dkg_expected_max_lines_dev(const double* a_dev, const double* b_dev, int32_t P, int32_t L,
                               double* emax_dev, int32_t* hull_count_dev, int32_t* hull_idx_dev,
                               double* hull_x_dev, int32_t hull_cap, double* dE_da_dev,
                               double* dE_db_dev, void* stream);

/*
 * dkg_piecewise_expectation_dev -- replaces calculate_expected_value_of_piecewise_linear_function
 *   (discretekg.py:415-452) for P piecewise-linear functions of H pieces each with CALLER-GIVEN break
 *   points: E[f(Z)], Z ~ N(0,1), f = a[p,k] + b[p,k] z on (z[p,k-1], z[p,k]), z[p,-1] = -inf,
 *   z[p,H-1] = +inf.  Also the gradient autograd would produce (optional outputs):
 *   dE/da [P, H], dE/db [P, H], dE/dz [P, H-1].
 *
 *   a_dev, b_dev [P, H]; z_dev [P, H-1] (may be NULL when H == 1); e_dev [P]
 *   H == 0 -> DKG_EEMPTY (the reference raises ValueError, discretekg.py:466-470)
 */
DKG_API int dkg_piecewise_expectation_dev(const double* a_dev, const double* b_dev, const double* z_dev,
                                          int32_t P, int32_t H, double* e_dev, double* dE_da_dev,
                                          double* dE_db_dev, double* dE_dz_dev, void* stream);

/*
 * dkg_int8_matmul_dev -- D[M, N] = A[M, K] . Bt[N, K]^T in fp64 accuracy on the int8 tensor cores
 *   (tcgen05.mma.kind::i8 over exact base-256 digit planes; csrc/dkg_ozaki.cu).  This is the
 *   contraction engine behind the covariance rows of discretekg.py:301, exposed so that it can
 *   be checked on its own.  n_digits in [1, 7]; n_diagonals in [1, 2 n_digits - 1]
 *   (0, 0 selects the defaults 7 / 8).  K <= 4096.
 */
DKG_API int dkg_int8_matmul_dev(const double* A_dev, int32_t lda, const double* Bt_dev, int32_t ldb,
                                int32_t M, int32_t N, int32_t K, int32_t n_digits,
                                int32_t n_diagonals, double* D_dev, int32_t ldd, void* stream);

/*
 * dkg_int8_peak -- measured int8 tensor-core peak of this library's own MMA instruction stream: the
 *   contraction kernel of dkg_int8_matmul_dev launched `reps` times on an M x N x K problem with the
 *   operand copies and the accumulator drains switched off (results discarded).  Used by bench.py as the
 *   roofline denominator of the conditioning contraction; not part of the hot path.
 *   mode: 3 = MMA stream only (the peak); 1 = no operand copies; 2 = no accumulator drains; 0 = the full
 *   kernel on dummy digits (mode 0..2 decompose what keeps the real launch below the peak).
 *   tops_host: executed int8 TOP/s; ms_host (optional): average launch duration.  Synchronises.
 */
DKG_API int dkg_int8_peak(int32_t M, int32_t N, int32_t K, int32_t reps, int32_t mode, double* tops_host,
                          double* ms_host, void* stream);

/*
 * Introspection for parity tests (all copies are device-to-device on `stream`).
 * name is one of:
 *   "B"        [n_i, N]   K_i^-1 k_i(X_train, X_disc)            (target objective i)
 *   "Kinv"     [n_i, n_i] (K_i + noise I)^-1
 *   "alpha"    [sum n_m]  mean caches K_m^-1 (y_m - c_m), objectives concatenated
 *   "mu_disc"  [N, M]     posterior means at the discretisation (un-standardised)
 *   "A0"       [S, N]     scalarised intercept table sum_m W[j,m] mu_m(x_n)
 *   "A0max"    [S]        its row maxima
 *   "chol"     [n_i, n_i] lower Cholesky factor of K_i + noise I
 * After a forward call:
 *   "slopes"   [C, N+1]   cov_i(x_c, .)/sd_i(x_c); LAST column is the candidate's own line
 *   "a_new"    [C, S]     scalarised posterior mean at the candidate (intercept of its own line)
 *   "var"      [C]        noisy predictive variance at the candidate
 *   "kg_terms" [C, S]     per-scalarisation E[max] - max (discretekg.py:336)
 * Returns the number of doubles the tensor holds (>= 0) or a negative error; copies
 * min(count, capacity) doubles when out_dev != NULL.
 */
DKG_API int64_t dkg_plan_read(dkg_plan* plan, const char* name, double* out_dev, int64_t capacity,
                      void* stream);

/* counters: kernels launched by this library since load / since the last reset */
DKG_API int64_t dkg_launch_count(void);
DKG_API void dkg_launch_count_reset(void);

/* Optional per-kernel timing for bench.py's roofline line: when enabled, every launch of a
 * forward is bracketed by CUDA events on the launching stream.  dkg_profile_read synchronises the
 * device, sums the elapsed milliseconds and launch counts per category and clears the samples.
 * Categories: 0 xprep, 1 gemm_T (KX K^-1), 2 var, 3 gemm_cov (the conditioning contraction:
 * int8 tensor-core kernel, or the DMMA kernel), 4 place_own, 5 zstat (+ slope completion),
 * 6 filter, 7 hull, 8 overflow (cooperative path), 9 finalize (+ backward), 10 digits (base-256
 * digit planes of the T rows for the int8 contraction). */
#define DKG_PROFILE_CATEGORIES 11
DKG_API void dkg_profile_enable(int on);
DKG_API int dkg_profile_read(double* ms_host, int64_t* count_host, int32_t ncat);

/* per-plan statistics of the last forward (8 host ints): [0] candidates, [1] lines surviving the
 * chord filter (sum over the sets finished by the warp kernel), [2] (candidate, scalarisation) sets
 * handled by the cooperative overflow kernel, [3] total hull vertices, [4] sets taking the
 * |slope| < 1e-9 shortcut, [5] sets that needed the block-wide exact march, [6] sets whose hull
 * records did not all fit (more than 64 vertices AND the spill pool, DKG_SPILL_BLOCKS, exhausted:
 * their value is exact, the gradient rows of their candidates are NaN), [7] 1 when the covariance
 * contraction of this plan runs on the int8 tensor cores (0: fp64 DMMA kernel) */
DKG_API int dkg_plan_stats(dkg_plan* plan, int64_t* out8_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DKG_B200_H */
