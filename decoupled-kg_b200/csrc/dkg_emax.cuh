// Expected maximum over fantasy outcomes: E[max_n (a_n + b_n Z)], Z ~ N(0,1), for many
// independent line sets.  Interfaces between dkg_emax.cu and dkg_api.cu.
#pragma once

#include "dkg_plan.cuh"

namespace dkg {

// A batch of line sets.  "row" c in [0, C), "scalarisation" j in [0, S): the set (c, j) has
// NL lines; line n has slope  b = wt[j] * Z[c*ldz + n]  and intercept
//   n <  NA : A[c*a_sc + j*a_sj + n]
//   n == NA : a_own[c*S + j]            (only when a_own != nullptr; then NL == NA + 1)
// In the KG path the intercept table is shared by all candidates (a_sc == 0) and line NA is the
// candidate's own line (the reference puts it at index 0, discretekg.py:277); the generic
// entry point (dkg_expected_max_lines_dev) uses S == 1, a_sc == L, wt == nullptr.
struct LineBatch {
  const double* Z = nullptr;
  int ldz = 0;
  const double* A = nullptr;
  const float* A32 = nullptr;      // optional float copy of A (same strides; shared table only): fp32 chord filter
  const float* A32tmax = nullptr;  // optional [a32_tiles, S] max of A32 over tiles of FILTER_TILE lines (tile culling)
  const int* A32targ = nullptr;    // optional [a32_tiles, S] the line attaining it (-1: the tile is padding only)
  int a32_tiles = 0;
  long long a_sc = 0;
  int a_sj = 0;
  const double* a_own = nullptr;
  const double* wt = nullptr;      // [S] or nullptr (== 1.0)
  const double* Amax = nullptr;    // max_n<NA A[..]: [S] when am_sc == 0, else [C, S]
  const int* Aarg = nullptr;
  long long am_sc = 0;
  int NA = 0, NL = 0, S = 0, C = 0;
  // row_mod > 0 (coupled path): every row is its own set (S == 1) and the shared intercept table /
  // maxima are indexed by (row % row_mod) -- rows are (candidate, scalarisation) pairs
  int row_mod = 0;
  // cov_M > 0 (coupled path): the slope rows are NOT materialised (they were 8.6 GB per 4096 candidates at c4);
  // the slope of line n in row (c, j) is formed on the fly from the M covariance rows of candidate c,
  //   z = (sum_m w2[j, m] cov_m[c, n]) / sd[row]          (line_slope() in dkg_emax.cu; Z is unused)
  int cov_M = 0;
  const double* cov[MAX_M] = {};   // [C, ldz] covariance rows Cov_m(x_c, .); column NA = Cov_m(x_c, x_c)
  const double* cov_w2 = nullptr;  // [row_mod, cov_M] squared scalarisation weights
  const double* cov_sd = nullptr;  // [rows] sqrt of the scalarised noisy variance
};

#ifdef __CUDACC__
// RN(s / sd) through the correctly rounded reciprocal and one residual step (Markstein); the SAME sequence
// everywhere a coupled slope is formed, so that every stage sees identical bits
__device__ __forceinline__ double coupled_quotient(double s, double sd, double rinv) {
  if (sd > 1e-290 && sd < 1e290) {
    const double q = s * rinv;
    return fma(fma(-q, sd, s), rinv, q);
  }
  return s / sd;
}
// (out of line: line_slope() is inlined at a dozen sites of the hull / overflow / finalize kernels, and two fp64
// divisions per site added 1.9 k of the 7.5 k instructions of hull_kernel -- instruction-cache misses on the
// decoupled path, which never executes them)
static __device__ __noinline__ double coupled_quotient_div(double s, double sd) {
  return coupled_quotient(s, sd, 1.0 / sd);
}
__device__ __forceinline__ double line_slope(const LineBatch& lb, int row, int n) {
  if (lb.cov_M == 0) return lb.Z[(size_t)row * lb.ldz + n];
  const int c = row / lb.row_mod, j = row - c * lb.row_mod;
  double s = 0.0;
  for (int m = 0; m < lb.cov_M; ++m) s = fma(lb.cov_w2[j * lb.cov_M + m], lb.cov[m][(size_t)c * lb.ldz + n], s);
  return coupled_quotient_div(s, lb.cov_sd[row]);
}
#endif

// one line that survived the streaming chord filter (intercept, raw slope coordinate, index)
struct SurvEntry {
  double a;
  double z;
  int idx;
  int pad;
};

struct EmaxScratch {
  double* zst = nullptr;        // [C, 2] min / max of the slope row
  int* zarg = nullptr;          // [C, 2] their first indices
  double* zpv = nullptr;        // [C, tiles, 2] per-tile min / max of the slope row (tiled statistics pass;
  int* zpi = nullptr;           //               optional) and their line indices
  double4* chain = nullptr;     // [C, S] chord-chain parameters of every set (zstat -> filter)
  double4* chainv = nullptr;    // [C, S, 2] end points (z0, a0, z1, a1) of the left / right chord in the raw
                                // slope coordinate (optional; second-level chain of the fp32 filter)
  double4* chain5 = nullptr;    // [C, S, 2] second-level chain (c_PU, m_PU, c_UT, m_UT), (c_TV, m_TV, c_VQ, m_VQ)
                                // built from the sampled phase of the fp32 filter (optional)
  float4* chain32 = nullptr;    // [C, S, 2] the same chain rounded conservatively to float, (m, m, c, c)
                                // per chord (optional; enables the fp32 filter)
  float4* chain5f = nullptr;    // [C, S, 2] float image of the second-level chain: (m_PU, m_UT, m_TV, m_VQ),
                                // (c_PU, c_UT, c_TV, c_VQ) (optional; tile-first filter)
  float2* ztile = nullptr;      // [C, ztiles] (min, max) of the float-rounded slopes of every FILTER_TILE
  int ztiles = 0;               // consecutive lines, written by the row-statistics pass (optional)
  int* surv_cnt = nullptr;      // [C, S]  lines that passed the filter (may exceed SURV_CAP)
  SurvEntry* surv = nullptr;    // [C, S, SURV_CAP]
  unsigned long long* far = nullptr;  // [C, S, 2] farthest late survivor above the left / right
                                      // chord, packed (float excess << 32 | line); 0 = none
  int* ovf_sets = nullptr;      // [C * S] queue of sets for the cooperative kernel
  int* long_sets = nullptr;     // optional [C * S] queue of the sets hull_short_kernel leaves to hull_kernel ...
  int* long_count = nullptr;    // ... and its length (reset by chain_kernel)
  int* ovf_count = nullptr;     // [1]
  int* spill_used = nullptr;    // [1] blocks claimed from the hull-record spill pool (optional)
  long long* stats = nullptr;   // [8] (optional)
  // surv_cnt, far, ovf_count and spill_used are zeroed by chain_kernel (emax_zstat*), i.e. before the
  // filter of the same batch runs
};

struct EmaxOut {
  double* terms = nullptr;     // [C, S]: E - max a (subtract_max) or E
  int subtract_max = 1;
  int* hull_cnt = nullptr;     // [C, S]           (optional)
  int* hull_idx = nullptr;     // [C, S, hull_cap] (optional)
  double* hull_p = nullptr;    // [C, S, hull_cap] dE/da (optional)
  double* hull_q = nullptr;    // [C, S, hull_cap] dE/db (optional)
  double* hull_x = nullptr;    // [C, S, hull_cap] intersections (optional)
  int hull_cap = 0;
  int* amax_is_own = nullptr;  // [C, S] (optional)
  double* kg = nullptr;        // [C] mean over scalarisations of terms (optional)
  long long* truncated = nullptr;  // [1] counts sets whose hull records could not all be stored (optional)
  // Spill chain for sets with more than hull_cap vertices (optional; KG path): records k >= hull_cap
  // live in 32-record blocks claimed from a pool as the march proceeds and chained per set, so the
  // backward sees EVERY vertex (the reference handles any hull size, discretekg.py:378-412).
  int* spill_head = nullptr;   // [C, S] first block of the set (-1: claim failed); valid iff hull_cnt > hull_cap
  int* spill_next = nullptr;   // [spill_blocks] next block of the same set (-1: none / claim failed)
  int* spill_idx = nullptr;    // [spill_blocks * 32]
  double* spill_p = nullptr;   // [spill_blocks * 32]
  double* spill_q = nullptr;   // [spill_blocks * 32]
  int* spill_used = nullptr;   // [1] blocks claimed so far (zeroed per forward)
  int spill_blocks = 0;
  double* dense_da = nullptr;  // [C*S, NL] dE/da scattered by line index (optional, pre-zeroed)
  double* dense_db = nullptr;  // [C*S, NL] dE/db (optional, pre-zeroed)
};

constexpr int SPILL_BLOCK = 32;  // records per spill block (one march batch)

#ifdef __CUDACC__
// Sequential reader of one set's hull records (k must not decrease between seek() calls).
struct HullReader {
  const EmaxOut* o;
  size_t set;
  int blk, bno;   // current spill block and its ordinal in the set's chain (bno < 0: chain not entered)
  size_t cur;
  bool spill;
  __device__ HullReader(const EmaxOut& out, size_t set_) : o(&out), set(set_), blk(-1), bno(-1), cur(0), spill(false) {}
  // false: the record was dropped because the spill pool was exhausted
  __device__ __forceinline__ bool seek(int k) {
    if (k < o->hull_cap) { cur = set * (size_t)o->hull_cap + k; spill = false; return true; }
    if (o->spill_head == nullptr) return false;
    const int b = (k - o->hull_cap) / SPILL_BLOCK;
    if (bno < 0) { blk = o->spill_head[set]; bno = 0; }
    while (bno < b && blk >= 0) { blk = o->spill_next[blk]; ++bno; }
    if (blk < 0) return false;
    cur = (size_t)blk * SPILL_BLOCK + (k - o->hull_cap) % SPILL_BLOCK;
    spill = true;
    return true;
  }
  __device__ __forceinline__ int idx() const { return spill ? o->spill_idx[cur] : o->hull_idx[cur]; }
  __device__ __forceinline__ double p() const { return spill ? o->spill_p[cur] : o->hull_p[cur]; }
  __device__ __forceinline__ double q() const { return spill ? o->spill_q[cur] : o->hull_q[cur]; }
};
#endif

constexpr int FIN_RMAX = 2560; // most hull records per candidate merged in shared memory (10 per scalarisation)
// slots of the distinct-line hash table (a power of two; distinct hull lines per candidate: tens at S = 16,
// many hundreds at S = 256)
__host__ __device__ inline int fin_hash_bits(int S) { return S <= 64 ? 10 : 12; }
__host__ __device__ inline int fin_rmax(int S) { const int r = 10 * S; return r < 256 ? 256 : r > FIN_RMAX ? FIN_RMAX : r; }

// Backward of the KG path (finalize kernel, one CTA per candidate).
struct BackwardArgs {
  double* dX = nullptr;          // [C, d]; nullptr -> forward only
  const double* X = nullptr;     // [C, d] raw candidates
  const double* T = nullptr;     // [C, ldk]  Kinv k_i(X_train, x_c)
  int ldk = 0;
  const double* BT = nullptr;    // [N, ldbt]
  int n_pad = 0;
  int ldbt = 0;
  const double* xd_s = nullptr;  // [N_pad, d]
  const double* var = nullptr;   // [C]
  const double* sd = nullptr;    // [C]
  const double* W = nullptr;     // [S, M]
  int M = 0, d = 0, target = 0;
  // per objective
  const double* xs[MAX_M];
  const double* alpha[MAX_M];
  int ntr[MAX_M];
  int kind[MAX_M];
  double outputscale[MAX_M];
  double y_std[MAX_M];
  double ls[MAX_M][MAX_D];
};

// Optional fused completion of the slope rows: when the covariance contraction ran on the int8
// tensor cores in product mode, Z holds P[c, n] = T[c, :] . k(X_train, x_n) and the row statistics
// pass turns it into the slopes  (k(x_c, x_n) - P) ystd^2 / sd[c]  (discretekg.py:301, :313) in
// place while it reads the row anyway.
struct CovFinish {
  const double* xs;    // [C, d] candidates / lengthscale
  const double* xd_s;  // [N_pad, d] discretisation / lengthscale
  const double* sd;    // [C]
  int d, kind, N;
  double outputscale, ystd2;
};
int emax_zstat(const LineBatch& lb, const EmaxScratch& sc, double* amax_out, int* aarg_out,
               cudaStream_t st, const CovFinish* fin = nullptr, bool* ztile_written = nullptr);
// row statistics from per-tile partials (sc.zpv / sc.zpi filled by the producer of the slope rows;
// ntiles == 0: sc.zst / sc.zarg are final already, only the chord chains are written),
// then the chord chains: replaces emax_zstat when the producer already saw every slope
int emax_zstat_from_partials(const LineBatch& lb, const EmaxScratch& sc, int ntiles, cudaStream_t st);
// ztile_valid: sc.ztile holds the per-tile slope ranges of this batch (written by emax_zstat)
int emax_filter(const LineBatch& lb, const EmaxScratch& sc, cudaStream_t st, bool ztile_valid = false);
// warp-per-set exact hull + closed-form expectation; sets it cannot finish go to the queue
// survivors_hint: average survivors per set seen by an earlier forward of the same plan (< 0: unknown); only
// steers which kernels are launched (8-lane groups for short sets or not) -- the results are identical
int emax_hull(const LineBatch& lb, const EmaxScratch& sc, const EmaxOut& out, cudaStream_t st,
              double survivors_hint = -1.0);
// CTA-per-set cooperative path for the queued sets (any input; always terminates)
int emax_overflow(const LineBatch& lb, const EmaxScratch& sc, const EmaxOut& out, cudaStream_t st);
// kg[c] = mean_j terms[c, j] and, if bw.dX, the fused envelope-theorem backward
int emax_finalize(const LineBatch& lb, const EmaxOut& out, const BackwardArgs& bw, cudaStream_t st);

// ---- small discretisations: the whole forward of one candidate in ONE CTA ----------------------
// At the sizes the reference's BO loop runs (11 x 11 grid, 10 restarts; bo_loop.py:123-131) the
// staged pipeline is ~17 dependent launches of a few microseconds each: ~100 us of pure kernel-chain
// latency per call whatever the batch (93 us device-side at C = 10 even when replayed as a CUDA graph).
// For N + 1 <= SMALL_MAX_LINES lines one CTA does everything up to the hull records: kernel rows and
// posterior means, T = K^-1 k_x (same refinement step), the variance, the covariance row (plain fp64
// dot products), and one warp per scalarisation running the reference's march over ALL lines: 36 us.
// (Measured limit: at N = 1024 the unfiltered marches make this path 2x SLOWER than the staged one --
// c2 at 512 candidates 0.57 vs 0.25 ms -- so it is kept to discretisations of a few hundred lines.)
constexpr int SMALL_MAX_LINES = 256;
constexpr int SMALL_MAX_TRAIN = 512;
struct SmallArgs {
  const double* X = nullptr;  // [C, d]
  int C = 0, d = 0, M = 0, target = 0;
  const double* xs[MAX_M] = {};     // [n_m, d] training inputs / lengthscale
  const double* alpha[MAX_M] = {};  // [n_m]
  int ntr[MAX_M] = {};
  int kind[MAX_M] = {};
  double outputscale[MAX_M] = {}, mean_const[MAX_M] = {}, y_mean[MAX_M] = {}, y_std[MAX_M] = {};
  double ls[MAX_M][MAX_D] = {};
  const double* W = nullptr;      // [S, M]
  const double* Kinv = nullptr;   // [n_pad, ldk]   target objective
  const double* Kmat = nullptr;   // [n_pad, ldk]
  int ldk = 0, refine = 0;
  const double* Kxd = nullptr;    // [n_pad, ldx]   k(X_train, X_disc)
  int ldx = 0;
  const double* xd_s = nullptr;   // [N_pad, d]
  double noise = 0.0;
  double* a_new = nullptr;        // [C, S] out
  double* T = nullptr;            // [C, ldk] out
  double* var = nullptr;          // [C] out
  double* sd = nullptr;           // [C] out
};
// writes lb.Z rows (non-const in practice), a_new, T, var, sd, the per-set terms and hull records
int emax_small_forward(const SmallArgs& a, const LineBatch& lb, const EmaxScratch& sc, const EmaxOut& out,
                       cudaStream_t st);

// E[f(Z)] for P piecewise-linear functions with H pieces each and ARBITRARY break points z [P, H-1]
// (calculate_expected_value_of_piecewise_linear_function, discretekg.py:415-452) and its gradient
int piecewise_expectation(const double* a, const double* b, const double* z, int P, int H, double* e,
                          double* de_da, double* de_db, double* de_dz, cudaStream_t st);

// Coupled evaluation (reference calculate_discrete_kg, discretekg.py:162-235): the slope of line n
// for scalarisation j is  sum_m W[j,m]^2 Cov_m(x, x_n) / sqrt(sum_m W[j,m]^2 var_m(x)).
struct CoupledArgs {
  int C = 0, S = 0, M = 0, d = 0, N = 0, ldz = 0;
  const double* W = nullptr;        // [S, M]
  const double* COV[MAX_M] = {};    // [C, ldz] covariance rows (un-standardised); column N = Cov_m(x,x)
  const double* varn[MAX_M] = {};   // [C] noisy predictive variance of objective m (un-standardised)
  double* sdj = nullptr;            // [C, S] out: sqrt of the scalarised noisy variance
  double* Zc = nullptr;             // [C * S, ldz] out: slope rows (nullptr: statistics only, rows not materialised)
  const double* W2 = nullptr;       // [S, M] squared weights (the table line_slope() reads)
  // optional: per-(row, tile) min / max of the finished slope rows (first index wins ties), so the row
  // statistics need no second pass over Zc; tile = CS_TILE_LINES consecutive entries of the N + 1
  double* zpv = nullptr;            // [C * S, tiles, 2]
  int* zpi = nullptr;               // [C * S, tiles, 2]
  // statistics-only form: the final row statistics are written here directly (zpv then only holds the
  // per-segment float ranges of the slope numerators)
  double* zst = nullptr;            // [C * S, 2]
  int* zarg = nullptr;              // [C * S, 2]
};
constexpr int CS_TILE_LINES = 256;  // one warp's segment (materialised slope rows)
constexpr int CST_SEG = 512;        // one warp's segment (statistics only: the covariance values stay in registers)
int coupled_slopes(const CoupledArgs& a, cudaStream_t st);
// segments per row of the partial statistics coupled_slopes() writes (input of emax_zstat_from_partials)
int coupled_stat_segments(int N, bool materialised);

struct CoupledBackward {
  double* dX = nullptr;             // [C, d]
  const double* X = nullptr;        // [C, d]
  const double* W = nullptr;        // [S, M]
  const double* COV[MAX_M] = {};    // [C, ldz] covariance rows (the slopes are formed on the fly, see LineBatch)
  const double* W2 = nullptr;       // [S, M] squared weights
  const double* sdj = nullptr;      // [C, S]
  int ldz = 0, M = 0, d = 0, S = 0, N = 0;
  const double* T[MAX_M] = {};      // [C, ldk_m]
  int ldk[MAX_M] = {};
  const double* BT[MAX_M] = {};     // [N, ldbt_m]
  int n_pad[MAX_M] = {};
  int ldbt[MAX_M] = {};
  const double* xd_s[MAX_M] = {};   // [N_pad, d]
  const double* xs[MAX_M] = {};
  const double* alpha[MAX_M] = {};
  int ntr[MAX_M] = {};
  int kind[MAX_M] = {};
  double outputscale[MAX_M] = {};
  double y_std[MAX_M] = {};
  double ls[MAX_M][MAX_D] = {};
};
// kg[c] = mean_j terms[c*S + j]; optional backward for the coupled path (CTA per candidate)
int emax_finalize_coupled(int C, int S, const EmaxOut& out, const CoupledBackward& bw, cudaStream_t st);

}  // namespace dkg
