// Internal layout of a dkg_plan: the candidate-independent state of one acquisition function
// plus the per-forward workspace.  See DESIGN.md "Data layout in HBM".
#pragma once

#include "dkg_common.cuh"

namespace dkg {

constexpr int GEMM_BM = 128;  // candidate rows per GEMM tile
constexpr int GEMM_BN = 128;  // discretisation columns per GEMM tile
constexpr int GEMM_BK = 16;   // training points per pipeline stage

struct ObjState {
  int n = 0;          // training points
  int n_pad = 0;      // rounded up to GEMM_BK (zero padded)
  int kernel = 0;
  double outputscale = 1, mean_const = 0, noise = 0, y_mean = 0, y_std = 1;
  double ls[MAX_D];
  int cap = 0;        // rows the n-dependent buffers below can hold (dkg_plan_append_point grows n up to it)
  double jitter = 0;  // Cholesky jitter this objective needed (0 normally)
  double* xs = nullptr;     // [cap, d] train inputs / lengthscale
  double* alpha = nullptr;  // [cap]    mean cache K^-1 (y - c), zero padded
  double* resid = nullptr;  // [cap]    y - c (kept for the mean-cache refresh of an append)
  // conditioning state (only for objectives whose observation is fantasised: the target in the
  // decoupled path, every objective in the coupled path)
  int ldk = 0;               // row stride of Kinv / T (n rounded up to GEMM_BN)
  double* chol = nullptr;    // [n, n]        lower Cholesky of K + noise I
  double* Kinv = nullptr;    // [n_pad, ldk]
  bool refine = true;        // T = KX K^-1 needs the refinement step (set at plan time from a bound on cond(K))
  double* Kmat = nullptr;    // [n_pad, ldk]  K + (noise + jitter) I, zero padded (refinement residual)
  double* Kxd = nullptr;     // [n_pad, N_pad]  k(X_train, X_disc) (raw cross-kernel, GEMM B operand)
  unsigned char* Kxd_dig = nullptr;  // [digits][N_pad][KP] base-256 digit planes of Kxd^T (int8 tensor-core path)
  double* Kxd_scale = nullptr;       // [N_pad] power-of-two scale of each discretisation point's column
  // digit planes of the rows of K^-1 and K (B operands of the int8 T = KX K^-1 products and their refinement step)
  unsigned char* Kinv_dig = nullptr;
  unsigned char* Kmat_dig = nullptr;
  double* Kinv_scale = nullptr;      // [ldk]
  double* Kmat_scale = nullptr;      // [ldk]
  double* B = nullptr;       // [n_pad, N_pad]  K^-1 k(X_train, X_disc), zero padded (backward only)
  double* BT = nullptr;      // [N, ldbt]
  int ldbt = 0;              // row stride of BT (>= n_pad; the capacity, so that appends stay in place)
  double* xd_s = nullptr;    // [N_pad, d]    discretisation / lengthscale
};

struct Workspace {
  int cap_C = 0;        // candidate capacity (multiple of GEMM_BM)
  int chunk_C = 0;      // candidates per slope-buffer chunk (multiple of GEMM_BM)
  double* X = nullptr;      // [cap_C, d]    staging for host-buffer calls
  double* kg = nullptr;     // [cap_C]
  double* dX = nullptr;     // [cap_C, d]
  double* KX = nullptr;     // [cap_C, ldk]      k_i(x_c, X_train)   (target objective)
  double* T = nullptr;      // [cap_C, ldk]      KX K^-1 (Kinv product + one refinement step)
  double* R = nullptr;      // [cap_C, ldk_max]  refinement residual KX - T K
  unsigned char* T_dig = nullptr;  // [digits][chunk_C][KP_max] digit planes of the T rows of one chunk
  double* T_scale = nullptr;       // [chunk_C]
  double* var = nullptr;    // [cap_C]           noisy predictive variance (un-standardised)
  double* sd = nullptr;     // [cap_C]           sqrt(var)
  double* zown = nullptr;   // [cap_C]           slope of the candidate's own line
  // coupled path only (per objective m, capacity cap_C each)
  double* KXm[MAX_M] = {};   // [cap_C, ldk_m]
  double* Tm[MAX_M] = {};    // [cap_C, ldk_m]
  double* varlat[MAX_M] = {};// [cap_C]  latent predictive variance (model space)
  double* COVm[MAX_M] = {};  // [chunk_C, ldz]  covariance rows Cov_m(x_c, .), column N = Cov_m(x_c, x_c)
  double* sdj = nullptr;     // [cap_C, S] sqrt of the scalarised noisy variance
  double* Zc = nullptr;      // [chunk_C * S, ldz]  per-(candidate, scalarisation) slope rows
  double* Xs = nullptr;     // [cap_C, d]        candidates / lengthscale_i
  double* a_new = nullptr;  // [cap_C, S]        intercept of the candidate's own line
  double* kg_terms = nullptr;  // [cap_C, S]
  double* Z = nullptr;      // [chunk_C, ldz]    slopes cov/sd; column N = candidate's own line
  double* zst = nullptr;    // [chunk_C, 2]      min / max of the slope row
  int* zarg = nullptr;      // [chunk_C, 2]      argmin, argmax of the slope row
  double* zpv = nullptr;    // [chunk_C, ZP_TILES, 2] per-tile min / max (tiled row-statistics pass)
  int* zpi = nullptr;       // [chunk_C, ZP_TILES, 2]
  void* chain = nullptr;    // [chunk_C, S] double4 chord-chain parameters
  void* chainv = nullptr;   // [chunk_C, S, 2] double4 end points of the two chords (second-level chain)
  void* chain5 = nullptr;   // [chunk_C, S, 2] double4 second-level chain
  void* chain32 = nullptr;  // [chunk_C, S, 2] float4: the chain rounded conservatively for the fp32 filter
  void* chain5f = nullptr;  // [chunk_C, S, 2] float4: float image of the second-level chain (tile-first filter)
  void* ztile = nullptr;    // [chunk_C, ztiles] float2: per-tile (min, max) of the float-rounded slope row
  int ztiles = 0;
  int* surv_cnt = nullptr;  // [chunk_C, S]      chord-filter survivors per (candidate, scal.)
  void* surv = nullptr;     // [chunk_C, S, SURV_CAP] SurvEntry (intercept, slope, index)
  unsigned long long* far = nullptr;  // [chunk_C, S, 2] farthest late survivors (chain seeds)
  int* ovf_sets = nullptr;  // [chunk_C * S] queue of sets for the cooperative kernel
  int* long_sets = nullptr; // [chunk_C * S + 1] queue of the sets hull_short_kernel left to hull_kernel; last entry = count
  int* ovf_count = nullptr; // [1]
  int* hull_cnt = nullptr;  // [chunk_C, S]
  int* hull_idx = nullptr;  // [chunk_C, S, HULL_CAP]
  double* hull_p = nullptr; // [chunk_C, S, HULL_CAP]  dE/da  (Phi differences)
  double* hull_q = nullptr; // [chunk_C, S, HULL_CAP]  dE/db  (-phi differences)
  int spill_blocks = 0;        // 32-record blocks of the hull-record spill pool (sets with > HULL_CAP vertices)
  int* spill_head = nullptr;   // [chunk_C, S]
  int* spill_next = nullptr;   // [spill_blocks]
  int* spill_idx = nullptr;    // [spill_blocks * 32]
  double* spill_p = nullptr;   // [spill_blocks * 32]
  double* spill_q = nullptr;   // [spill_blocks * 32]
  int* spill_used = nullptr;   // [1]
  int* amax_is_new = nullptr;  // [chunk_C, S]  1 if the max intercept is the candidate's own line
  long long* stats = nullptr;  // [8] device counters
  // host mirror of `stats`, copied asynchronously after every forward with gradients: the NEXT call
  // (or the host-buffer entry point, which synchronises anyway) sees whether the spill pool ran dry
  // and grows it -- without a host sync on the device-resident path
  long long* stats_pinned = nullptr;  // [8] pinned host memory
  cudaEvent_t stats_ev = nullptr;
  bool stats_pending = false;
  bool spill_checked = false;  // the first forward with gradients after a (re)allocation is checked synchronously
  // Small batches are launch-bound (~17 kernels of a few microseconds each): their launch sequence is
  // captured once per (C, with / without gradient) into a CUDA graph over the plan's own staging
  // buffers (X, kg, dX above) and replayed with one launch.
  struct GraphSlot { int C = 0; int grad = 0; cudaGraphExec_t exec = nullptr; unsigned long long used = 0; };
  static constexpr int GRAPH_SLOTS = 8;
  GraphSlot graphs[GRAPH_SLOTS];
  unsigned long long graph_clock = 0;
  cudaStream_t cap_stream = nullptr;  // capture stream (the caller's stream may be the legacy default stream)
  int last_C = 0;
};

constexpr int STATS_WORDS = 10;   // 8 counters (dkg_plan_stats) + the spill pool's fill count + padding
constexpr int FILTER_TILE = 128;  // lines per warp of the fp32 chord filter (4 per lane)
constexpr int SURV_CAP = 2048;  // survivors per (candidate, scalarisation) before the slow path
constexpr int HULL_CAP = 64;    // hull vertices recorded per (candidate, scalarisation)

}  // namespace dkg

struct dkg_plan {
  int M = 0, d = 0, N = 0, S = 0, target = 0;  // target < 0: coupled evaluation
  int N_pad = 0;  // N rounded up to GEMM_BN
  int ldz = 0;    // row stride of the slope buffer (>= N+1, multiple of 16)
  int ldk = 0;    // row stride of Kinv / T (n_i rounded up to GEMM_BN)
  int device = 0;
  int cov_digits = 7, cov_diagonals = 7;  // base-256 digit configuration of the int8 contraction (DKG_PLAN_FAST32: 4 / 4)
  dkg::ObjState obj[dkg::MAX_M];
  double W_host[dkg::MAX_S * dkg::MAX_M];
  double* W = nullptr;        // [S, M]
  double* W2 = nullptr;       // [S, M]  squared weights (coupled slopes)
  double* wt = nullptr;       // [S]  W[:, target]
  double* xd = nullptr;       // [N, d]      raw discretisation
  double* xd_s = nullptr;     // [N_pad, d]  discretisation / lengthscale_i (padding rows zero)
  double* chol = nullptr;     // [n_i, n_i]  lower Cholesky of K_i + noise I
  double* cholT = nullptr;    // [n_i, n_i]  its transpose (row access in the back substitution)
  double* Kinv = nullptr;     // [n_pad, ldk]
  double* B = nullptr;        // [n_pad, N_pad]  K_i^-1 k_i(X_train, X_disc), zero padded
  double* BT = nullptr;       // [N, n_pad]      transpose, for the backward gather
  double* alpha_all = nullptr;  // concatenated mean caches (exposed by dkg_plan_read)
  double* mu_disc = nullptr;  // [N, M]
  double* A0 = nullptr;       // [S, N_pad]   scalarised intercepts of the discretisation lines
  float* A0f = nullptr;       // [S, N_pad]   float copy of A0 (fp32 chord filter; padding = -inf)
  float* A0tmax = nullptr;    // [a0_tiles, S] max of A0f over tiles of FILTER_TILE consecutive lines (tile culling)
  int* A0targ = nullptr;      // [a0_tiles, S] the line attaining it (the tile's champion: sample of the second-level chain)
  int a0_tiles = 0;
  // Internal line order: the discretisation is re-ordered along a Morton curve at plan time, so that
  // FILTER_TILE consecutive lines are neighbours in input space -- similar posterior means and similar
  // covariance with any candidate -- and the chord filter can drop most tiles with one test.  Line n of
  // every internal table is line perm[n] of the caller's x_discretisation (nullptr: identity).
  int* perm = nullptr;        // [N]
  double* A0max = nullptr;    // [S]
  int* A0arg = nullptr;       // [S]
  double jitter = 0.0;        // Cholesky jitter that was needed (0 normally)
  dkg::Workspace ws;
};
