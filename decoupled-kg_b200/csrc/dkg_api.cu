// C-ABI of libdkg_b200.so (see include/dkg_b200.h) and the orchestration of one forward pass.
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <algorithm>
#include <numeric>
#include <string>

#include "dkg_emax.cuh"
#include "dkg_kernels.cuh"

namespace dkg {

// ---- error / counters ------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- optional per-kernel timing (CUDA events on the launching stream) -------------------------
static bool g_prof_on = false;
struct ProfSample { int cat; cudaEvent_t e0, e1; };
static std::vector<ProfSample> g_prof;
struct ProfScope {
  cudaStream_t st; int idx = -1;
  ProfScope(int cat, cudaStream_t s) : st(s) {
    if (!g_prof_on) return;
    ProfSample ps; ps.cat = cat;
    cudaEventCreate(&ps.e0); cudaEventCreate(&ps.e1);
    cudaEventRecord(ps.e0, st);
    g_prof.push_back(ps); idx = (int)g_prof.size() - 1;
  }
  ~ProfScope() { if (idx >= 0) cudaEventRecord(g_prof[idx].e1, st); }
};

// ---- allocation helpers ----------------------------------------------------------------------
template <class T>
static int dev_alloc(T** p, size_t count, bool zero = true) {
  *p = nullptr;
  if (count == 0) count = 1;
  cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
  if (e != cudaSuccess) {
    set_error("cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
    return DKG_ENOMEM;
  }
  if (zero) {
    e = cudaMemset(*p, 0, count * sizeof(T));
    if (e != cudaSuccess) {
      set_error("cudaMemset failed: %s", cudaGetErrorString(e));
      return DKG_ECUDA;
    }
  }
  return DKG_OK;
}
template <class T>
static void dev_free(T*& p) {
  if (p) cudaFree(p);
  p = nullptr;
}

static void drop_graphs(Workspace& w) {
  for (auto& g : w.graphs) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    g = Workspace::GraphSlot();
  }
}

static void free_workspace(Workspace& w) {
  drop_graphs(w);
  if (w.cap_stream) { cudaStreamDestroy(w.cap_stream); w.cap_stream = nullptr; }
  dev_free(w.X); dev_free(w.kg); dev_free(w.dX); dev_free(w.KX); dev_free(w.T); dev_free(w.R); dev_free(w.T_dig); dev_free(w.T_scale); dev_free(w.var);
  dev_free(w.sd); dev_free(w.zown); dev_free(w.Xs); dev_free(w.a_new); dev_free(w.kg_terms);
  dev_free(w.Z); dev_free(w.zst); dev_free(w.zarg); dev_free(w.zpv); dev_free(w.zpi); dev_free(w.surv_cnt); { SurvEntry* t = (SurvEntry*)w.surv; dev_free(t); w.surv = nullptr; } { double4* t = (double4*)w.chain; dev_free(t); w.chain = nullptr; } { float4* t = (float4*)w.chain32; dev_free(t); w.chain32 = nullptr; } { double4* t = (double4*)w.chainv; dev_free(t); w.chainv = nullptr; } { double4* t = (double4*)w.chain5; dev_free(t); w.chain5 = nullptr; } { float4* t = (float4*)w.chain5f; dev_free(t); w.chain5f = nullptr; } { float2* t = (float2*)w.ztile; dev_free(t); w.ztile = nullptr; } dev_free(w.far); dev_free(w.ovf_sets); dev_free(w.long_sets); dev_free(w.ovf_count);
  dev_free(w.hull_cnt); dev_free(w.hull_idx); dev_free(w.hull_p); dev_free(w.hull_q);
  dev_free(w.spill_head); dev_free(w.spill_next); dev_free(w.spill_idx); dev_free(w.spill_p); dev_free(w.spill_q);
  w.spill_used = nullptr;  // (part of the stats allocation)
  dev_free(w.amax_is_new); dev_free(w.stats); dev_free(w.sdj); dev_free(w.Zc);
  for (int m = 0; m < MAX_M; ++m) { dev_free(w.KXm[m]); dev_free(w.Tm[m]); dev_free(w.varlat[m]); dev_free(w.COVm[m]); }
  if (w.stats_pinned) { cudaFreeHost(w.stats_pinned); w.stats_pinned = nullptr; }
  if (w.stats_ev) { cudaEventDestroy(w.stats_ev); w.stats_ev = nullptr; }
  w.stats_pending = false;
  w.spill_checked = false;
  w.cap_C = w.chunk_C = 0;
}

static int alloc_spill(Workspace& w, long long blocks) {
  drop_graphs(w);  // captured launches carry the pool's addresses and size
  dev_free(w.spill_next); dev_free(w.spill_idx); dev_free(w.spill_p); dev_free(w.spill_q);
  if (blocks < 1) blocks = 1;
  w.spill_blocks = (int)(blocks > 0x03ffffff ? 0x03ffffff : blocks);
  DKG_TRY(dev_alloc(&w.spill_next, (size_t)w.spill_blocks, false));
  DKG_TRY(dev_alloc(&w.spill_idx, (size_t)w.spill_blocks * SPILL_BLOCK, false));
  DKG_TRY(dev_alloc(&w.spill_p, (size_t)w.spill_blocks * SPILL_BLOCK, false));
  DKG_TRY(dev_alloc(&w.spill_q, (size_t)w.spill_blocks * SPILL_BLOCK, false));
  return DKG_OK;
}

// The last forward dropped hull records (spill pool exhausted): size the pool from what that forward
// saw -- its total hull-vertex count bounds the records that can spill -- with 50 % headroom.
// Returns 1 when the pool was grown (the caller may re-run), 0 if there was nothing to do.
static int grow_spill_if_needed(dkg_plan* p, const long long* h, int* grew) {
  *grew = 0;
  Workspace& w = p->ws;
  if (h[6] <= 0 || getenv("DKG_SPILL_BLOCKS") != nullptr) return DKG_OK;  // an explicit size is never overridden
  const long long sets = (long long)w.chunk_C * p->S;
  long long need = (h[3] / SPILL_BLOCK + sets) * 3 / 2;
  if (need <= w.spill_blocks) need = 2ll * w.spill_blocks;
  DKG_CUDA_OK(cudaDeviceSynchronize());
  DKG_TRY(alloc_spill(w, need));
  *grew = 1;
  return DKG_OK;
}

static int ensure_workspace(dkg_plan* p, int C) {
  Workspace& w = p->ws;
  if (C <= w.cap_C) return DKG_OK;
  // a (rare) growth: synchronise so nothing in flight still uses the old buffers
  DKG_CUDA_OK(cudaDeviceSynchronize());
  free_workspace(w);
  const int cap = round_up(C, GEMM_BM);
  const bool coupled = p->target < 0;
  // Candidates are processed in chunks that bound the scratch memory (slope rows + survivor
  // lists).  The budget for both is a quarter of the memory that is free now, at least 6 GB and at most 32 GB
  // (DKG_CHUNK_MB overrides): at c4 shapes 3.8 GB cover all 4096 candidates in one chunk either way; with many
  // scalarisations (c5: S = 256, 12.7 MB per candidate) a 6 GB budget meant chunks of 384 candidates -- launches of
  // 2.6 waves of tiles, the last chunk below the size the tile-first filter needs -- and 8.5 instead of 5.9 ms for
  // 2048 candidates.  B200 has 180 GB; the chunks are for whatever does not fit, not a habit.
  double chunk_mb = 6144.0;
  {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
      const double quarter = (double)free_b / 1048576.0 / 4.0;
      if (quarter > chunk_mb) chunk_mb = quarter < 32768.0 ? quarter : 32768.0;
    } else {
      cudaGetLastError();
    }
  }
  if (const char* e = getenv("DKG_CHUNK_MB")) chunk_mb = atof(e);
  double per_row = (double)p->ldz * sizeof(double) +
                   (double)p->S * (SURV_CAP * sizeof(SurvEntry) + HULL_CAP * 20.0 + 64.0);
  if (coupled) per_row += (double)p->ldz * sizeof(double) * p->M;  // M covariance rows; the S slope rows are not materialised
  long long rows = (long long)(chunk_mb * 1048576.0 / per_row);
  int chunk = (int)(rows / GEMM_BM) * GEMM_BM;
  if (chunk < GEMM_BM) chunk = GEMM_BM;
  if (chunk > cap) chunk = cap;
  if (chunk < cap) {  // equal chunks: the same number of launches, no small remainder chunk
    const int nch = ceil_div(cap, chunk);
    chunk = round_up(ceil_div(cap, nch), GEMM_BM);
  }
  const size_t S = p->S;
  DKG_TRY(dev_alloc(&w.X, (size_t)cap * p->d));
  DKG_TRY(dev_alloc(&w.kg, (size_t)cap));
  DKG_TRY(dev_alloc(&w.dX, (size_t)cap * p->d));
  int ldk_max = p->ldk;
  if (coupled) for (int m = 0; m < p->M; ++m) ldk_max = ldk_max > p->obj[m].ldk ? ldk_max : p->obj[m].ldk;
  DKG_TRY(dev_alloc(&w.KX, (size_t)cap * (coupled ? GEMM_BN : p->ldk)));
  DKG_TRY(dev_alloc(&w.T, (size_t)cap * (coupled ? GEMM_BN : p->ldk)));
  DKG_TRY(dev_alloc(&w.R, (size_t)cap * ldk_max));
  {
    int n_max = 1;
    for (int m = 0; m < p->M; ++m)
      if (coupled || m == p->target) {
        const int nc = p->obj[m].cap > p->obj[m].n ? p->obj[m].cap : p->obj[m].n;  // (room for appended points)
        n_max = n_max > nc ? n_max : nc;
      }
    // (sized for the whole batch, not for a chunk: the T solve slices all rows at once -- with many scalarisations
    // a chunk is a few hundred candidates and chunk-sized products were launches of a dozen tiles each)
    DKG_TRY(dev_alloc(&w.T_dig, ozaki_digit_bytes(cap, n_max, p->cov_digits)));
    DKG_TRY(dev_alloc(&w.T_scale, (size_t)cap + 128));  // + one row block: the CTA-pair kernel reads 256-row pairs
  }
  DKG_TRY(dev_alloc(&w.var, (size_t)cap));
  DKG_TRY(dev_alloc(&w.sd, (size_t)cap));
  DKG_TRY(dev_alloc(&w.zown, (size_t)cap));
  DKG_TRY(dev_alloc(&w.Xs, (size_t)cap * p->d));
  DKG_TRY(dev_alloc(&w.a_new, (size_t)cap * S));
  DKG_TRY(dev_alloc(&w.kg_terms, (size_t)cap * S));
  // the slope-row statistics are per ROW of the expected-max batch: one row per candidate in the
  // decoupled path, one per (candidate, scalarisation) in the coupled path
  const size_t rows_per_cand = coupled ? S : 1;
  DKG_TRY(dev_alloc(&w.Z, coupled ? (size_t)1 : (size_t)chunk * p->ldz));
  DKG_TRY(dev_alloc(&w.zst, (size_t)chunk * rows_per_cand * 2));
  DKG_TRY(dev_alloc(&w.zarg, (size_t)chunk * rows_per_cand * 2));
  {  // tiles of >= 1024 lines (64 KB of points at d <= 8); coupled: one row per (candidate, scalarisation)
    const size_t tiles = coupled ? (size_t)((p->N + 1) / CS_TILE_LINES + 2) : (size_t)(p->N / 1024 + 2);
    DKG_TRY(dev_alloc(&w.zpv, (size_t)chunk * rows_per_cand * tiles * 2));
    DKG_TRY(dev_alloc(&w.zpi, (size_t)chunk * rows_per_cand * tiles * 2));
  }
  if (coupled) {
    for (int m = 0; m < p->M; ++m) {
      DKG_TRY(dev_alloc(&w.KXm[m], (size_t)cap * p->obj[m].ldk));
      DKG_TRY(dev_alloc(&w.Tm[m], (size_t)cap * p->obj[m].ldk));
      DKG_TRY(dev_alloc(&w.varlat[m], (size_t)cap));
      DKG_TRY(dev_alloc(&w.COVm[m], (size_t)chunk * p->ldz));
    }
    DKG_TRY(dev_alloc(&w.sdj, (size_t)cap * S));
    if (getenv("DKG_COUPLED_ROWS") != nullptr) DKG_TRY(dev_alloc(&w.Zc, (size_t)chunk * S * p->ldz));  // (old materialised form, for comparison)
  }
  DKG_TRY(dev_alloc(&w.surv_cnt, (size_t)chunk * S));
  { SurvEntry* t = nullptr; DKG_TRY(dev_alloc(&t, (size_t)chunk * S * SURV_CAP, false)); w.surv = t; }
  DKG_TRY(dev_alloc(&w.far, (size_t)chunk * S * 2));
  { double4* t = nullptr; DKG_TRY(dev_alloc(&t, (size_t)chunk * S, false)); w.chain = t; }
  { float4* t = nullptr; DKG_TRY(dev_alloc(&t, (size_t)chunk * S * 2, false)); w.chain32 = t; }
  { double4* t = nullptr; DKG_TRY(dev_alloc(&t, (size_t)chunk * S * 2, false)); w.chainv = t; }
  { double4* t = nullptr; DKG_TRY(dev_alloc(&t, (size_t)chunk * S * 2, false)); w.chain5 = t; }
  if (!coupled) {
    { float4* t = nullptr; DKG_TRY(dev_alloc(&t, (size_t)chunk * S * 2, false)); w.chain5f = t; }
    w.ztiles = ceil_div(p->N, FILTER_TILE);
    { float2* t = nullptr; DKG_TRY(dev_alloc(&t, (size_t)chunk * w.ztiles, false)); w.ztile = t; }
  }
  DKG_TRY(dev_alloc(&w.ovf_sets, (size_t)chunk * S, false));
  DKG_TRY(dev_alloc(&w.long_sets, (size_t)chunk * S + 1));
  DKG_TRY(dev_alloc(&w.ovf_count, (size_t)1));
  DKG_TRY(dev_alloc(&w.hull_cnt, (size_t)chunk * S));
  DKG_TRY(dev_alloc(&w.hull_idx, (size_t)chunk * S * HULL_CAP, false));
  DKG_TRY(dev_alloc(&w.hull_p, (size_t)chunk * S * HULL_CAP, false));
  DKG_TRY(dev_alloc(&w.hull_q, (size_t)chunk * S * HULL_CAP, false));
  {  // spill pool for sets with more than HULL_CAP hull vertices: one block per 4 sets to start with;
     // it grows on demand (grow_spill_if_needed) unless DKG_SPILL_BLOCKS pins its size
    long long blocks = (long long)chunk * (long long)S / 4;
    if (blocks < 512) blocks = 512;
    if (const char* e = getenv("DKG_SPILL_BLOCKS")) blocks = atoll(e) >= 0 ? atoll(e) : blocks;
    DKG_TRY(dev_alloc(&w.spill_head, (size_t)chunk * S, false));
    DKG_TRY(alloc_spill(w, blocks));
    DKG_CUDA_OK(cudaHostAlloc((void**)&w.stats_pinned, sizeof(long long) * 8, cudaHostAllocDefault));
    memset(w.stats_pinned, 0, sizeof(long long) * 8);
    DKG_CUDA_OK(cudaEventCreateWithFlags(&w.stats_ev, cudaEventDisableTiming));
    w.stats_pending = false;
    w.spill_checked = false;
  }
  DKG_TRY(dev_alloc(&w.amax_is_new, (size_t)chunk * S));
  DKG_TRY(dev_alloc(&w.stats, (size_t)STATS_WORDS));  // 8 counters + the spill pool's fill count (one memset clears all)
  w.spill_used = reinterpret_cast<int*>(w.stats + 8);
  w.cap_C = cap;
  w.chunk_C = chunk;
  return DKG_OK;
}

static void destroy_plan(dkg_plan* p) {
  if (!p) return;
  cudaDeviceSynchronize();
  for (int m = 0; m < p->M; ++m) {
    ObjState& o = p->obj[m];
    dev_free(o.xs); dev_free(o.alpha); dev_free(o.resid); dev_free(o.chol); dev_free(o.Kinv); dev_free(o.Kmat); dev_free(o.Kxd_dig); dev_free(o.Kxd_scale); dev_free(o.Kinv_dig); dev_free(o.Kmat_dig); dev_free(o.Kinv_scale); dev_free(o.Kmat_scale); dev_free(o.B); dev_free(o.Kxd);
    dev_free(o.BT); dev_free(o.xd_s);
  }
  dev_free(p->W); dev_free(p->W2); dev_free(p->wt); dev_free(p->xd); dev_free(p->alpha_all);
  dev_free(p->mu_disc); dev_free(p->A0); dev_free(p->A0f); dev_free(p->A0max); dev_free(p->A0arg);
  dev_free(p->A0tmax); dev_free(p->A0targ); dev_free(p->perm);
  free_workspace(p->ws);
  delete p;
}

// The covariance contraction runs on the int8 tensor cores (dkg_ozaki.cu) unless DKG_COV_GEMM=dmma
// asks for the fp64 DMMA kernel (dkg_gemm.cu), which also serves n > OZ_MAX_K.
static bool use_int8_cov() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DKG_COV_GEMM");
    v = (e != nullptr && strcmp(e, "dmma") == 0) ? 0 : 1;
  }
  return v != 0;
}

// digit planes of the rows of K^-1 and K (both symmetric): B operands of T0 = KX K^-1, R = KX - T0 K, T = T0 + R K^-1
// on the int8 engine (solve_T).  Rows n .. ldk keep zero digits and a zero scale.
static int make_solve_digits(ObjState& o, const dkg_plan* p, cudaStream_t st) {
  if (!use_int8_cov() || o.n > OZ_MAX_K || o.Kinv == nullptr || o.Kmat == nullptr || (o.ldk % 128) != 0) return DKG_OK;
  const int k_cap = o.cap > o.n ? (o.cap < OZ_MAX_K ? o.cap : OZ_MAX_K) : o.n;
  const size_t bytes = ozaki_digit_bytes(o.ldk, k_cap, p->cov_digits);
  if (o.Kinv_dig == nullptr) DKG_TRY(dev_alloc(&o.Kinv_dig, bytes));
  if (o.Kmat_dig == nullptr) DKG_TRY(dev_alloc(&o.Kmat_dig, bytes));
  if (o.Kinv_scale == nullptr) DKG_TRY(dev_alloc(&o.Kinv_scale, (size_t)o.ldk + 128));
  if (o.Kmat_scale == nullptr) DKG_TRY(dev_alloc(&o.Kmat_scale, (size_t)o.ldk + 128));
  // (the k block count follows n: planes written for a smaller n are stale, so clear before re-slicing)
  DKG_CUDA_OK(cudaMemsetAsync(o.Kinv_dig, 0, bytes, st));
  DKG_CUDA_OK(cudaMemsetAsync(o.Kmat_dig, 0, bytes, st));
  DKG_TRY(ozaki_slice_rows(o.Kinv, o.ldk, o.n, o.n, 128, p->cov_digits, o.Kinv_dig, o.Kinv_scale, st));
  DKG_TRY(ozaki_slice_rows(o.Kmat, o.ldk, o.n, o.n, 128, p->cov_digits, o.Kmat_dig, o.Kmat_scale, st));
  return DKG_OK;
}

// digit planes of Kxd^T (one row per discretisation point) for the int8 path
static int make_kxd_digits(ObjState& o, const dkg_plan* p, cudaStream_t st) {
  DKG_TRY(make_solve_digits(o, p, st));
  if (!use_int8_cov() || o.n > OZ_MAX_K) return DKG_OK;
  double* KT = nullptr;
  DKG_TRY(dev_alloc(&KT, (size_t)p->N * o.n_pad, false));
  int rc = transpose(o.Kxd, o.n, p->N, p->N_pad, KT, o.n_pad, st);
  // (sized for the objective's row capacity, so that an appended training point only needs a re-slice)
  const int k_cap = o.cap > o.n ? (o.cap < OZ_MAX_K ? o.cap : OZ_MAX_K) : o.n;
  if (rc == DKG_OK && o.Kxd_dig == nullptr) rc = dev_alloc(&o.Kxd_dig, ozaki_digit_bytes(p->N_pad, k_cap, p->cov_digits));
  if (rc == DKG_OK && o.Kxd_scale == nullptr) rc = dev_alloc(&o.Kxd_scale, (size_t)p->N_pad);
  if (rc == DKG_OK)
    rc = ozaki_slice_rows(KT, o.n_pad, p->N, o.n, ozaki_b_block_rows(p->cov_digits, p->cov_diagonals), p->cov_digits,
                          o.Kxd_dig, o.Kxd_scale, st);
  cudaStreamSynchronize(st);
  dev_free(KT);
  return rc;
}

// Z rows of one chunk: (k(x_c, x_n) - T[c, :] . k(X_train, x_n)) ystd^2 / sd[c]
static int cov_rows(const ObjState& o, const dkg_plan* p, Workspace& w, const double* T, int cc, int cc_pad,
                    const CovEpilogue& ep, cudaStream_t st) {
  if (o.Kxd_dig != nullptr) {
    DKG_TRY(ozaki_slice_rows(T, o.ldk, cc, o.n, 128, p->cov_digits, w.T_dig, w.T_scale, st));
    return ozaki_cov(w.T_dig, w.T_scale, cc_pad, o.Kxd_dig, o.Kxd_scale, p->N_pad, o.n, p->cov_digits,
                     p->cov_diagonals, /*b_nonneg=*/true, ep, st);  // stationary kernel values are positive
  }
  return gemm_cov(T, o.ldk, o.Kxd, p->N_pad, cc_pad, p->N_pad, o.n_pad, ep, st);
}

// Cholesky of K_m + noise I with GPyTorch's jitter retries (psd_safe_cholesky: 1e-8 * 10^k,
// k < 3, in double) [recalled].  Leaves L in `Lbuf` (n x n).
static int factor_objective(const ObjState& o, int d, double* Lbuf, int* info_dev, double* jitter_out,
                            cudaStream_t st, bool blocked) {
  const double jitters[4] = {0.0, 1e-8, 1e-7, 1e-6};
  for (int k = 0; k < 4; ++k) {
    DKG_TRY(kmat_train(o, d, jitters[k], Lbuf, o.n, st));
    if (blocked) DKG_TRY(cholesky_blocked(Lbuf, o.n, o.n, info_dev, st));
    else DKG_TRY(cholesky_inplace(Lbuf, o.n, info_dev, st));
    int info = 0;
    DKG_CUDA_OK(cudaMemcpyAsync(&info, info_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    DKG_CUDA_OK(cudaStreamSynchronize(st));
    if (info == 0) {
      *jitter_out = jitters[k];
      return DKG_OK;
    }
  }
  set_error("training covariance is not positive definite even with 1e-6 jitter");
  return DKG_ENOTPD;
}

static bool needs_refinement(const ObjState& o, double jitter) {
  const double s2 = o.noise + jitter;
  if (!(s2 > 0.0)) return true;
  const double bound = (double)o.n * (o.outputscale + s2) * o.outputscale / (s2 * s2) * 2.220446049250313e-16;
  return !(bound < 1e-11);
}

// p->xd <- the discretisation in the plan's internal line order (see dkg_plan::perm): sorted by the
// Morton code of the points (bits interleaved across the d coordinates, as many bits per coordinate as
// fit into 63).  DKG_NO_REORDER=1 keeps the caller's order.
static int order_discretisation(dkg_plan* p, const double* x_disc_dev, cudaStream_t st) {
  const int N = p->N, d = p->d;
  if (N < 2 * FILTER_TILE || getenv("DKG_NO_REORDER") != nullptr) {
    DKG_CUDA_OK(cudaMemcpyAsync(p->xd, x_disc_dev, sizeof(double) * N * d, cudaMemcpyDeviceToDevice, st));
    return DKG_OK;
  }
  std::vector<double> h((size_t)N * d);
  DKG_CUDA_OK(cudaMemcpyAsync(h.data(), x_disc_dev, sizeof(double) * N * d, cudaMemcpyDeviceToHost, st));
  DKG_CUDA_OK(cudaStreamSynchronize(st));
  std::vector<double> lo(d, INFINITY), hi(d, -INFINITY);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < d; ++k) {
      const double v = h[(size_t)n * d + k];
      if (v < lo[k]) lo[k] = v;
      if (v > hi[k]) hi[k] = v;
    }
  const int bits = 63 / d < 16 ? 63 / d : 16;
  std::vector<unsigned long long> code(N, 0ull);
  for (int n = 0; n < N; ++n) {
    unsigned long long c = 0ull;
    for (int k = 0; k < d; ++k) {
      const double span = hi[k] - lo[k];
      double u = span > 0.0 ? (h[(size_t)n * d + k] - lo[k]) / span : 0.0;
      if (!(u >= 0.0)) u = 0.0;  // NaN coordinates sort first
      unsigned long long q = (unsigned long long)(u * (double)((1ull << bits) - 1) + 0.5);
      for (int b = 0; b < bits; ++b) c |= ((q >> b) & 1ull) << (b * d + k);
    }
    code[n] = c;
  }
  std::vector<int> perm(N);
  std::iota(perm.begin(), perm.end(), 0);
  std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return code[a] < code[b]; });
  DKG_TRY(dev_alloc(&p->perm, (size_t)N, false));
  DKG_CUDA_OK(cudaMemcpyAsync(p->perm, perm.data(), sizeof(int) * N, cudaMemcpyHostToDevice, st));
  DKG_TRY(gather_rows(x_disc_dev, p->perm, N, d, p->xd, st));
  DKG_CUDA_OK(cudaStreamSynchronize(st));  // `perm` is a stack-lifetime host buffer
  return DKG_OK;
}

static int build_plan(dkg_plan* p, const dkg_objective* objs, const double* x_disc_dev,
                      cudaStream_t st) {
  const int d = p->d, N = p->N, M = p->M, S = p->S, tgt = p->target;
  // --- per objective: scaled inputs, Cholesky, mean cache ---
  int n_max = 0, n_sum = 0;
  for (int m = 0; m < M; ++m) {
    n_max = n_max > objs[m].n ? n_max : objs[m].n;
    n_sum += objs[m].n;
  }
  double *Lbuf = nullptr, *LTbuf = nullptr;
  int* info_dev = nullptr;
  DKG_TRY(dev_alloc(&Lbuf, (size_t)n_max * n_max));
  DKG_TRY(dev_alloc(&LTbuf, (size_t)n_max * n_max));
  DKG_TRY(dev_alloc(&info_dev, 1));
  {
    int cap_sum = 0;
    for (int m = 0; m < M; ++m) cap_sum += round_up(objs[m].n, GEMM_BM);
    DKG_TRY(dev_alloc(&p->alpha_all, (size_t)(cap_sum > n_sum ? cap_sum : n_sum)));
  }
  DKG_TRY(dev_alloc(&p->mu_disc, (size_t)N * M));
  DKG_TRY(dev_alloc(&p->xd, (size_t)N * d));
  DKG_TRY(order_discretisation(p, x_disc_dev, st));

  int rc = DKG_OK;
  int off = 0;
  for (int m = 0; m < M && rc == DKG_OK; ++m) {
    ObjState& o = p->obj[m];
    const dkg_objective& s = objs[m];
    o.n = s.n;
    o.n_pad = round_up(s.n, GEMM_BK);
    o.kernel = s.kernel;
    o.outputscale = s.outputscale;
    o.mean_const = s.mean_const;
    o.noise = s.noise;
    o.y_mean = s.y_mean;
    o.y_std = s.y_std;
    for (int k = 0; k < MAX_D; ++k) o.ls[k] = k < d ? s.lengthscale_host[k] : 1.0;
    const bool fast = o.n <= chol_fast_max() && getenv("DKG_SLOW_PREPARE") == nullptr;
    o.cap = fast ? round_up(o.n, GEMM_BM) : o.n_pad;  // the fast path's 128-padded buffers leave room for appends
    if ((rc = dev_alloc(&o.xs, (size_t)o.cap * d)) != DKG_OK) break;
    if ((rc = dev_alloc(&o.alpha, (size_t)o.cap)) != DKG_OK) break;
    if ((rc = dev_alloc(&o.resid, (size_t)o.cap)) != DKG_OK) break;
    if ((rc = scale_rows(s.train_x_dev, o.n, d, o.ls, o.xs, st)) != DKG_OK) break;
    if ((rc = residual(s.train_y_dev, o.n, o.mean_const, o.resid, st)) != DKG_OK) break;
    const bool need_state = (m == tgt || tgt < 0);
    const int n = o.n;
    double jit = 0.0;
    if (fast) {
      // ---- fast path: blocked Cholesky, explicit L^-1, solves as DMMA GEMMs ----
      if ((rc = factor_objective(o, d, Lbuf, info_dev, &jit, st, /*blocked=*/true)) != DKG_OK) break;
      o.refine = needs_refinement(o, jit);
      const int np = round_up(n, GEMM_BM);  // 128-padded square buffers for the GEMM operands
      double *Linv = nullptr, *LinvT = nullptr, *tmpv = nullptr, *Ybuf = nullptr;
      auto cleanup = [&]() { cudaStreamSynchronize(st); dev_free(Linv); dev_free(LinvT); dev_free(tmpv); dev_free(Ybuf); };
      if ((rc = dev_alloc(&Linv, (size_t)np * np)) != DKG_OK) break;
      if ((rc = dev_alloc(&LinvT, (size_t)np * np)) != DKG_OK) { cleanup(); break; }
      if ((rc = dev_alloc(&tmpv, (size_t)np)) != DKG_OK) { cleanup(); break; }
      if ((rc = tri_inverse(Lbuf, n, n, Linv, np, st)) != DKG_OK) { cleanup(); break; }
      if ((rc = transpose(Linv, n, n, np, LinvT, np, st)) != DKG_OK) { cleanup(); break; }
      // mean cache: alpha = K^-1 (y - c) = L^-T (L^-1 r)
      if ((rc = residual(s.train_y_dev, n, o.mean_const, o.alpha, st)) != DKG_OK) { cleanup(); break; }
      if ((rc = matvec(Linv, np, n, n, o.alpha, tmpv, st)) != DKG_OK) { cleanup(); break; }
      if ((rc = matvec(LinvT, np, n, n, tmpv, o.alpha, st)) != DKG_OK) { cleanup(); break; }
      cudaMemcpyAsync(p->alpha_all + off, o.alpha, sizeof(double) * n, cudaMemcpyDeviceToDevice, st);
      off += n;
      if ((rc = mu_disc(p->xd, N, d, o, p->mu_disc, M, m, st)) != DKG_OK) { cleanup(); break; }
      o.jitter = jit;
      {
        // K and K^-1 of EVERY objective are kept (n^2 doubles each): the mean cache of any objective can
        // then be refreshed in O(n^2) when a training point is appended (dkg_plan_append_point)
        o.ldk = np;
        // Kinv = L^-T L^-1
        if ((rc = dev_alloc(&o.Kinv, (size_t)np * o.ldk)) != DKG_OK) { cleanup(); break; }
        if ((rc = gemm_store(LinvT, np, Linv, np, np, o.ldk, o.n_pad, o.Kinv, o.ldk, st)) != DKG_OK) { cleanup(); break; }
        if ((rc = dev_alloc(&o.Kmat, (size_t)np * o.ldk)) != DKG_OK) { cleanup(); break; }
        if ((rc = kmat_train(o, d, jit, o.Kmat, o.ldk, st)) != DKG_OK) { cleanup(); break; }
      }
      if (need_state) {
        if ((rc = dev_alloc(&o.chol, (size_t)n * n)) != DKG_OK) { cleanup(); break; }
        cudaMemcpyAsync(o.chol, Lbuf, sizeof(double) * n * n, cudaMemcpyDeviceToDevice, st);
        // B = K^-1 k(X_train, X_disc) = L^-T (L^-1 Kxd)
        if ((rc = dev_alloc(&o.xd_s, (size_t)p->N_pad * d)) != DKG_OK) { cleanup(); break; }
        if ((rc = scale_rows(p->xd, N, d, o.ls, o.xd_s, st)) != DKG_OK) { cleanup(); break; }
        if ((rc = dev_alloc(&o.B, (size_t)np * p->N_pad)) != DKG_OK) { cleanup(); break; }
        if ((rc = dev_alloc(&Ybuf, (size_t)np * p->N_pad)) != DKG_OK) { cleanup(); break; }
        if ((rc = dev_alloc(&o.Kxd, (size_t)np * p->N_pad)) != DKG_OK) { cleanup(); break; }
        if ((rc = kcross(o, o.xd_s, N, d, o.Kxd, p->N_pad, st)) != DKG_OK) { cleanup(); break; }
        if ((rc = make_kxd_digits(o, p, st)) != DKG_OK) { cleanup(); break; }
        if ((rc = gemm_store(Linv, np, o.Kxd, p->N_pad, np, p->N_pad, o.n_pad, Ybuf, p->N_pad, st)) != DKG_OK) { cleanup(); break; }
        if ((rc = gemm_store(LinvT, np, Ybuf, p->N_pad, np, p->N_pad, o.n_pad, o.B, p->N_pad, st)) != DKG_OK) { cleanup(); break; }
        // one refinement step (see solve_T): R = Kxd - K B;  B += Kinv R
        if ((rc = gemm_axpy(o.Kmat, o.ldk, o.B, p->N_pad, np, p->N_pad, o.n_pad, o.Kxd, p->N_pad, -1.0, Ybuf, p->N_pad, st)) != DKG_OK) { cleanup(); break; }
        if ((rc = gemm_axpy(o.Kinv, o.ldk, Ybuf, p->N_pad, np, p->N_pad, o.n_pad, o.B, p->N_pad, 1.0, o.B, p->N_pad, st)) != DKG_OK) { cleanup(); break; }
        o.ldbt = np;
        if ((rc = dev_alloc(&o.BT, (size_t)N * o.ldbt)) != DKG_OK) { cleanup(); break; }
        if ((rc = transpose(o.B, o.n_pad, N, p->N_pad, o.BT, o.ldbt, st)) != DKG_OK) { cleanup(); break; }
      }
      cleanup();
    } else {
      // ---- general path (large n): column-at-a-time Cholesky, per-column substitution ----
      if ((rc = factor_objective(o, d, Lbuf, info_dev, &jit, st, /*blocked=*/false)) != DKG_OK) break;
      o.refine = needs_refinement(o, jit);
      if ((rc = transpose(Lbuf, n, n, n, LTbuf, n, st)) != DKG_OK) break;
      if ((rc = residual(s.train_y_dev, n, o.mean_const, o.alpha, st)) != DKG_OK) break;
      if ((rc = cholesky_solve_inplace(Lbuf, LTbuf, n, o.alpha, 1, 1, st)) != DKG_OK) break;
      cudaMemcpyAsync(p->alpha_all + off, o.alpha, sizeof(double) * n, cudaMemcpyDeviceToDevice, st);
      off += n;
      if ((rc = mu_disc(p->xd, N, d, o, p->mu_disc, M, m, st)) != DKG_OK) break;
      if (need_state) {
        o.ldk = round_up(n, GEMM_BN);
        if ((rc = dev_alloc(&o.chol, (size_t)n * n)) != DKG_OK) break;
        cudaMemcpyAsync(o.chol, Lbuf, sizeof(double) * n * n, cudaMemcpyDeviceToDevice, st);
        if ((rc = dev_alloc(&o.Kinv, (size_t)o.n_pad * o.ldk)) != DKG_OK) break;
        if ((rc = set_identity(o.Kinv, n, o.ldk, st)) != DKG_OK) break;
        if ((rc = cholesky_solve_inplace(Lbuf, LTbuf, n, o.Kinv, n, o.ldk, st)) != DKG_OK) break;
        if ((rc = dev_alloc(&o.Kmat, (size_t)o.n_pad * o.ldk)) != DKG_OK) break;
        if ((rc = kmat_train(o, d, jit, o.Kmat, o.ldk, st)) != DKG_OK) break;
        if ((rc = dev_alloc(&o.xd_s, (size_t)p->N_pad * d)) != DKG_OK) break;
        if ((rc = scale_rows(p->xd, N, d, o.ls, o.xd_s, st)) != DKG_OK) break;
        if ((rc = dev_alloc(&o.B, (size_t)o.n_pad * p->N_pad)) != DKG_OK) break;
        if ((rc = dev_alloc(&o.Kxd, (size_t)o.n_pad * p->N_pad)) != DKG_OK) break;
        if ((rc = kcross(o, o.xd_s, N, d, o.Kxd, p->N_pad, st)) != DKG_OK) break;
        if ((rc = make_kxd_digits(o, p, st)) != DKG_OK) break;
        cudaMemcpyAsync(o.B, o.Kxd, sizeof(double) * (size_t)o.n_pad * p->N_pad, cudaMemcpyDeviceToDevice, st);
        if ((rc = cholesky_solve_inplace(Lbuf, LTbuf, n, o.B, N, p->N_pad, st)) != DKG_OK) break;
        o.ldbt = o.n_pad;
        if ((rc = dev_alloc(&o.BT, (size_t)N * o.ldbt)) != DKG_OK) break;
        if ((rc = transpose(o.B, o.n_pad, N, p->N_pad, o.BT, o.ldbt, st)) != DKG_OK) break;
      }
    }
    if (need_state) p->jitter = jit > p->jitter ? jit : p->jitter;
    if (m == tgt) {  // aliases used by the decoupled path
      p->ldk = o.ldk; p->chol = o.chol; p->Kinv = o.Kinv; p->B = o.B; p->BT = o.BT; p->xd_s = o.xd_s;
    }
  }
  if (rc == DKG_OK) {
    // scalarised intercept table
    rc = dev_alloc(&p->W, (size_t)S * M);
    if (rc == DKG_OK) rc = dev_alloc(&p->W2, (size_t)S * M);
    if (rc == DKG_OK) rc = dev_alloc(&p->wt, (size_t)S);
    if (rc == DKG_OK) rc = dev_alloc(&p->A0, (size_t)S * p->N_pad);
    if (rc == DKG_OK) rc = dev_alloc(&p->A0f, (size_t)S * p->N_pad);
    if (rc == DKG_OK) rc = dev_alloc(&p->A0max, (size_t)S);
    if (rc == DKG_OK) rc = dev_alloc(&p->A0arg, (size_t)S);
    if (rc == DKG_OK) {
      double wt_host[MAX_S];
      for (int j = 0; j < S; ++j) wt_host[j] = tgt >= 0 ? p->W_host[j * M + tgt] : 1.0;
      double w2_host[MAX_S * MAX_M];
      for (int e = 0; e < S * M; ++e) w2_host[e] = p->W_host[e] * p->W_host[e];
      cudaMemcpyAsync(p->W, p->W_host, sizeof(double) * S * M, cudaMemcpyHostToDevice, st);
      cudaMemcpyAsync(p->W2, w2_host, sizeof(double) * S * M, cudaMemcpyHostToDevice, st);
      cudaMemcpyAsync(p->wt, wt_host, sizeof(double) * S, cudaMemcpyHostToDevice, st);
      cudaStreamSynchronize(st);  // wt_host is a stack buffer
      rc = build_a0(p->mu_disc, N, M, p->W, S, p->A0, p->A0f, p->N_pad, p->A0max, p->A0arg, st);
      p->a0_tiles = ceil_div(p->N_pad, FILTER_TILE);
      if (rc == DKG_OK) rc = dev_alloc(&p->A0tmax, (size_t)S * p->a0_tiles);
      if (rc == DKG_OK) rc = dev_alloc(&p->A0targ, (size_t)S * p->a0_tiles);
      if (rc == DKG_OK) rc = build_a0_tilemax(p->A0f, p->N_pad, S, FILTER_TILE, p->a0_tiles, p->A0tmax, p->A0targ, st);
    }
  }
  cudaStreamSynchronize(st);
  cudaError_t e = cudaGetLastError();
  dev_free(Lbuf);
  dev_free(LTbuf);
  dev_free(info_dev);
  if (rc == DKG_OK && e != cudaSuccess) {
    set_error("CUDA error while building the plan: %s", cudaGetErrorString(e));
    rc = DKG_ECUDA;
  }
  return rc;
}

// T = KX K^-1 for a batch of cross-kernel rows (KX, T, R all with row stride o.ldk).
//
// A plain product with the explicit inverse loses kappa(K) * eps of *backward* accuracy: on the
// survey-c2 target objective (kappa 4.6e7) the quadratic form kx^T K^-1 kx came out with 2e-4
// relative error in the predictive variance.  One step of fixed-precision iterative refinement,
//     T0 = KX Kinv;   R = KX - T0 K;   T = T0 + R Kinv,
// restores the residual KX - T K to O(eps |T||K|), i.e. the accuracy of the two triangular
// solves the reference performs (gpytorch cholesky_solve), while staying on the DMMA GEMM.
// The step is skipped when it cannot matter: with K = K_f + s2 I (K_f PSD, s2 = noise + jitter) we have
// lambda_min >= s2 and lambda_max <= trace = n (os + s2), and the relative error of the predictive variance
// of the unrefined product is at most cond(K) eps os / s2 <= n (os + s2) os / s2^2 eps; below 1e-11 the
// refinement is below the other rounding errors (c4 objective 0, noise 1: 9e-14).
// DKG_T_SOLVE=trsm selects the batched substitution kernel, =kinv the unrefined product, =refine forces the step.
static int t_solve_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("DKG_T_SOLVE");
    mode = 0;
    if (e != nullptr && strcmp(e, "trsm") == 0) mode = 1;
    if (e != nullptr && strcmp(e, "kinv") == 0) mode = 2;
    if (e != nullptr && strcmp(e, "refine") == 0) mode = 3;  // refine even when the plan says it is not needed
  }
  return mode;
}

static bool t_solve_int8() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DKG_T_GEMM");  // =dmma: the fp64 DMMA GEMMs for every batch size
    v = (e != nullptr && strcmp(e, "dmma") == 0) ? 0 : 1;
  }
  return v != 0;
}

static int solve_T(const dkg_plan* p, Workspace& w, const ObjState& o, const double* KX, double* T, double* R, int C,
                   int C_pad, cudaStream_t st) {
  const int mode = t_solve_mode();
  if (mode == 1 && o.chol != nullptr && o.n <= batched_solve_max_n())
    return launch_batched_cholesky_solve(o.chol, o.n, KX, o.ldk, C, T, o.ldk, st);
  const bool refine = !(mode == 2 || (!o.refine && mode != 3));
  if (t_solve_int8() && o.Kinv_dig != nullptr && w.T_dig != nullptr && o.n >= 256) {
    // (a rule of the PLAN -- its number of training points --, never of the batch size: sub-batches and full
    // batches must give identical bits.  At n = 100 the DMMA GEMMs take ~17 us each and win: 52 vs 83 us per solve.)
    // The same exact int8 digit-plane products as the covariance contraction.  The fp64 DMMA
    // GEMM needs 75 us for a [4096, 416] x [416, 416] product (one under-filled wave of 128 x 64 tiles at 72 %
    // of the DMMA rate); here it is one wave of 128 x 128 tiles of ~25 us plus 20 us of digit slicing.
    const int NS = p->cov_digits, NG = p->cov_diagonals;
    for (int c0 = 0; c0 < C; c0 += w.cap_C) {  // (one pass: the digit buffer holds cap_C >= C rows)
      const int cc = (C - c0) < w.cap_C ? (C - c0) : w.cap_C;
      const int cc_pad = round_up(cc, GEMM_BM);
      const double* kx = KX + (size_t)c0 * o.ldk;
      double* t = T + (size_t)c0 * o.ldk;
      double* r = R + (size_t)c0 * o.ldk;
      DKG_TRY(ozaki_slice_rows(kx, o.ldk, cc, o.n, 128, NS, w.T_dig, w.T_scale, st));
      DKG_TRY(ozaki_store_axpy(w.T_dig, w.T_scale, cc_pad, o.Kinv_dig, o.Kinv_scale, o.ldk, o.n, NS, NG, nullptr, 0, 1.0,
                               t, o.ldk, cc, o.ldk, st));
      if (!refine) continue;
      DKG_TRY(ozaki_slice_rows(t, o.ldk, cc, o.n, 128, NS, w.T_dig, w.T_scale, st));
      DKG_TRY(ozaki_store_axpy(w.T_dig, w.T_scale, cc_pad, o.Kmat_dig, o.Kmat_scale, o.ldk, o.n, NS, NG, kx, o.ldk, -1.0,
                               r, o.ldk, cc, o.ldk, st));
      DKG_TRY(ozaki_slice_rows(r, o.ldk, cc, o.n, 128, NS, w.T_dig, w.T_scale, st));
      DKG_TRY(ozaki_store_axpy(w.T_dig, w.T_scale, cc_pad, o.Kinv_dig, o.Kinv_scale, o.ldk, o.n, NS, NG, t, o.ldk, 1.0,
                               t, o.ldk, cc, o.ldk, st));
    }
    return DKG_OK;
  }
  DKG_TRY(gemm_store(KX, o.ldk, o.Kinv, o.ldk, C_pad, o.ldk, o.n_pad, T, o.ldk, st));
  if (!refine) return DKG_OK;
  DKG_TRY(gemm_axpy(T, o.ldk, o.Kmat, o.ldk, C_pad, o.ldk, o.n_pad, KX, o.ldk, -1.0, R, o.ldk, st));
  DKG_TRY(gemm_axpy(R, o.ldk, o.Kinv, o.ldk, C_pad, o.ldk, o.n_pad, T, o.ldk, 1.0, T, o.ldk, st));
  return DKG_OK;
}

static int forward_coupled(dkg_plan* p, const double* X, int C, double* kg, double* dX, cudaStream_t st);

static int forward_once(dkg_plan* p, const double* X, int C, double* kg, double* dX, cudaStream_t st);
static int forward_coupled(dkg_plan* p, const double* X, int C, double* kg, double* dX, cudaStream_t st);

// ---- CUDA graphs for small batches -------------------------------------------------------------
static bool graphs_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DKG_GRAPHS");
    v = (e != nullptr && atoi(e) == 0) ? 0 : 1;
  }
  return v != 0;
}

static bool graph_eligible(const dkg_plan* p, int C) {
  // launch-bound sizes only: a few hundred microseconds of kernels at most
  static long long max_cn = -1;
  if (max_cn < 0) {
    const char* e = getenv("DKG_GRAPH_MAX_CN");  // (measurement switch: replay larger batches from a graph too)
    max_cn = e != nullptr ? atoll(e) : (1ll << 21);
  }
  return graphs_enabled() && !g_prof_on && C <= p->ws.chunk_C && (long long)C * p->N <= max_cn;
}

// Replays (capturing at first use) the launch sequence of one forward over the staging buffers
// w.X -> w.kg, w.dX.  Returns DKG_OK with *done = false when capture is not possible (the caller
// then launches the kernels directly).
static int forward_graph(dkg_plan* p, int C, bool grad, cudaStream_t st, bool* done) {
  Workspace& w = p->ws;
  *done = false;
  Workspace::GraphSlot* slot = nullptr;
  for (auto& g : w.graphs)
    if (g.exec != nullptr && g.C == C && g.grad == (grad ? 1 : 0)) slot = &g;
  if (slot == nullptr) {
    if (w.cap_stream == nullptr) DKG_CUDA_OK(cudaStreamCreateWithFlags(&w.cap_stream, cudaStreamNonBlocking));
    // everything the kernels of this forward were given before must have been consumed: the capture
    // itself runs nothing, but first-use attribute calls inside it are not stream ordered
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(w.cap_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      return DKG_OK;
    }
    const int rc = forward_once(p, w.X, C, w.kg, grad ? w.dX : nullptr, w.cap_stream);
    const cudaError_t ce = cudaStreamEndCapture(w.cap_stream, &graph);
    if (rc != DKG_OK || ce != cudaSuccess || graph == nullptr) {
      cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      return rc != DKG_OK ? rc : DKG_OK;
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess || exec == nullptr) {
      cudaGetLastError();
      return DKG_OK;
    }
    slot = &w.graphs[0];
    for (auto& g : w.graphs)
      if (g.exec == nullptr) { slot = &g; break; }
      else if (g.used < slot->used) slot = &g;
    if (slot->exec) cudaGraphExecDestroy(slot->exec);
    slot->C = C; slot->grad = grad ? 1 : 0; slot->exec = exec;
  }
  slot->used = ++w.graph_clock;
  w.last_C = C;
  DKG_CUDA_OK(cudaGraphLaunch(slot->exec, st));
  count_launch();
  *done = true;
  return DKG_OK;
}

// one forward on device buffers: through a graph over the staging buffers when the batch is small
static int enqueue_forward(dkg_plan* p, const double* X, int C, double* kg, double* dX, cudaStream_t st) {
  Workspace& w = p->ws;
  if (graph_eligible(p, C)) {
    const size_t xb = sizeof(double) * (size_t)C * p->d;
    if (X != w.X) DKG_CUDA_OK(cudaMemcpyAsync(w.X, X, xb, cudaMemcpyDeviceToDevice, st));
    bool done = false;
    DKG_TRY(forward_graph(p, C, dX != nullptr, st, &done));
    if (done) {
      if (kg != w.kg) DKG_CUDA_OK(cudaMemcpyAsync(kg, w.kg, sizeof(double) * (size_t)C, cudaMemcpyDeviceToDevice, st));
      if (dX != nullptr && dX != w.dX) DKG_CUDA_OK(cudaMemcpyAsync(dX, w.dX, xb, cudaMemcpyDeviceToDevice, st));
      return DKG_OK;
    }
    return forward_once(p, X != w.X ? w.X : X, C, kg, dX, st);
  }
  return forward_once(p, X, C, kg, dX, st);
}

// outcome of an EARLIER forward (copied asynchronously, looked at without a sync): grow the pool if
// that forward dropped records
static int spill_precheck(dkg_plan* p) {
  Workspace& w = p->ws;
  if (w.stats_pending && w.stats_ev != nullptr && cudaEventQuery(w.stats_ev) == cudaSuccess) {
    w.stats_pending = false;
    int grew = 0;
    DKG_TRY(grow_spill_if_needed(p, w.stats_pinned, &grew));
  }
  return DKG_OK;
}

// One forward (+ backward) on device buffers with the spill-pool bookkeeping around it: the first
// forward with gradients on a workspace is checked synchronously (and re-run with a larger pool if
// records were dropped), later ones leave their outcome for the next call's precheck -- no host sync.
static int forward_impl(dkg_plan* p, const double* X, int C, double* kg, double* dX, cudaStream_t st) {
  if (C == 0) return DKG_OK;
  DKG_TRY(spill_precheck(p));
  DKG_TRY(ensure_workspace(p, C));
  Workspace& w = p->ws;
  for (int attempt = 0;; ++attempt) {
    DKG_TRY(enqueue_forward(p, X, C, kg, dX, st));
    if (dX == nullptr) return DKG_OK;
    DKG_CUDA_OK(cudaMemcpyAsync(w.stats_pinned, w.stats, sizeof(long long) * 8, cudaMemcpyDeviceToHost, st));
    if (w.spill_checked) {
      DKG_CUDA_OK(cudaEventRecord(w.stats_ev, st));
      w.stats_pending = true;
      return DKG_OK;
    }
    DKG_CUDA_OK(cudaStreamSynchronize(st));
    w.spill_checked = true;
    int grew = 0;
    DKG_TRY(grow_spill_if_needed(p, w.stats_pinned, &grew));
    if (!grew || attempt >= 3) return DKG_OK;
  }
}

// Small discretisations (SmallArgs in dkg_emax.cuh): one CTA per candidate does the whole forward,
// then the usual finalize kernel -- two launches instead of ~17.  The rule depends on the PLAN only,
// so a sub-batch and the full batch of the same plan always take the same path (identical bits).
static bool small_path(const dkg_plan* p) {
  const char* e = getenv("DKG_SMALL");  // (read per forward so that tests can compare the two paths)
  if ((e != nullptr && atoi(e) == 0) || p->target < 0 || p->N + 1 > SMALL_MAX_LINES) return false;
  if (p->cov_digits != OZ_DEFAULT_DIGITS) return false;  // DKG_PLAN_FAST32 asks for the reduced-precision contraction
  for (int m = 0; m < p->M; ++m)
    if (p->obj[m].n > SMALL_MAX_TRAIN) return false;
  return p->obj[p->target].Kxd != nullptr;
}

static int forward_small(dkg_plan* p, const double* X, int C, double* kg, double* dX, cudaStream_t st) {
  Workspace& w = p->ws;
  const ObjState& ot = p->obj[p->target];
  const int d = p->d, N = p->N, S = p->S, M = p->M;
  w.last_C = C;
  DKG_CUDA_OK(cudaMemsetAsync(w.stats, 0, sizeof(long long) * STATS_WORDS, st));
  for (int c0 = 0; c0 < C; c0 += w.chunk_C) {
    const int cc = (C - c0) < w.chunk_C ? (C - c0) : w.chunk_C;
    SmallArgs a;
    a.X = X + (size_t)c0 * d; a.C = cc; a.d = d; a.M = M; a.target = p->target;
    for (int m = 0; m < M; ++m) {
      const ObjState& o = p->obj[m];
      a.xs[m] = o.xs; a.alpha[m] = o.alpha; a.ntr[m] = o.n; a.kind[m] = o.kernel;
      a.outputscale[m] = o.outputscale; a.mean_const[m] = o.mean_const; a.y_mean[m] = o.y_mean; a.y_std[m] = o.y_std;
      for (int k = 0; k < MAX_D; ++k) a.ls[m][k] = o.ls[k];
    }
    a.W = p->W; a.Kinv = ot.Kinv; a.Kmat = ot.Kmat; a.ldk = ot.ldk; a.refine = ot.refine ? 1 : 0;
    a.Kxd = ot.Kxd; a.ldx = p->N_pad; a.xd_s = p->xd_s; a.noise = ot.noise;
    a.a_new = w.a_new + (size_t)c0 * S; a.T = w.T + (size_t)c0 * p->ldk; a.var = w.var + c0; a.sd = w.sd + c0;
    LineBatch lb;
    lb.Z = w.Z; lb.ldz = p->ldz;
    lb.A = p->A0; lb.a_sc = 0; lb.a_sj = p->N_pad;
    lb.a_own = a.a_new;
    lb.wt = p->wt;
    lb.Amax = p->A0max; lb.Aarg = p->A0arg; lb.am_sc = 0;
    lb.NA = N; lb.NL = N + 1; lb.S = S; lb.C = cc;
    EmaxScratch sc;
    sc.stats = w.stats; sc.spill_used = w.spill_used;
    EmaxOut out;
    out.terms = w.kg_terms + (size_t)c0 * S;
    out.subtract_max = 1;
    out.hull_cnt = w.hull_cnt;
    out.hull_cap = HULL_CAP;
    out.amax_is_own = w.amax_is_new;
    out.kg = kg + c0;
    out.truncated = w.stats + 6;
    BackwardArgs bw{};
    if (dX != nullptr) {
      out.hull_idx = w.hull_idx; out.hull_p = w.hull_p; out.hull_q = w.hull_q;
      out.spill_head = w.spill_head; out.spill_next = w.spill_next; out.spill_idx = w.spill_idx;
      out.spill_p = w.spill_p; out.spill_q = w.spill_q; out.spill_used = w.spill_used;
      out.spill_blocks = w.spill_blocks;
      if (c0 > 0) DKG_CUDA_OK(cudaMemsetAsync(w.spill_used, 0, sizeof(int), st));  // the pool is per chunk
      bw.dX = dX + (size_t)c0 * d; bw.X = a.X; bw.T = a.T; bw.ldk = p->ldk; bw.BT = p->BT; bw.n_pad = ot.n_pad; bw.ldbt = ot.ldbt;
      bw.xd_s = p->xd_s; bw.var = a.var; bw.sd = a.sd; bw.W = p->W; bw.M = M; bw.d = d; bw.target = p->target;
      for (int m = 0; m < M; ++m) {
        const ObjState& o = p->obj[m];
        bw.xs[m] = o.xs; bw.alpha[m] = o.alpha; bw.ntr[m] = o.n; bw.kind[m] = o.kernel;
        bw.outputscale[m] = o.outputscale; bw.y_std[m] = o.y_std;
        for (int k = 0; k < MAX_D; ++k) bw.ls[m][k] = o.ls[k];
      }
    }
    { ProfScope ps(7, st); DKG_TRY(emax_small_forward(a, lb, sc, out, st)); }
    { ProfScope ps(9, st); DKG_TRY(emax_finalize(lb, out, bw, st)); }
  }
  return DKG_OK;
}

static int forward_once(dkg_plan* p, const double* X, int C, double* kg, double* dX, cudaStream_t st) {
  if (p->target < 0) return forward_coupled(p, X, C, kg, dX, st);
  if (small_path(p)) return forward_small(p, X, C, kg, dX, st);
  Workspace& w = p->ws;
  const ObjState& ot = p->obj[p->target];
  const int d = p->d, N = p->N, S = p->S, M = p->M;
  const int C_pad = round_up(C, GEMM_BM);
  w.last_C = C;
  DKG_CUDA_OK(cudaMemsetAsync(w.stats, 0, sizeof(long long) * STATS_WORDS, st));

  XprepArgs xa{};
  xa.X = X; xa.C = C; xa.d = d; xa.M = M; xa.S = S; xa.target = p->target;
  for (int m = 0; m < M; ++m) {
    const ObjState& o = p->obj[m];
    xa.xs[m] = o.xs; xa.alpha[m] = o.alpha; xa.ntr[m] = o.n; xa.kind[m] = o.kernel;
    xa.outputscale[m] = o.outputscale; xa.mean_const[m] = o.mean_const;
    xa.y_mean[m] = o.y_mean; xa.y_std[m] = o.y_std;
    for (int k = 0; k < MAX_D; ++k) xa.ls[m][k] = o.ls[k];
  }
  xa.W = p->W; xa.Xs = w.Xs; xa.KX = w.KX; xa.n_pad = p->ldk; xa.a_new = w.a_new;
  { ProfScope ps(0, st); DKG_TRY(launch_xprep(xa, st)); }

  // T = KX K^-1 (backward stable), then the predictive variance
  { ProfScope ps(1, st); DKG_TRY(solve_T(p, w, ot, w.KX, w.T, w.R, C, C_pad, st)); }
  const double ystd2 = ot.y_std * ot.y_std;
  { ProfScope ps(2, st);
    DKG_TRY(launch_var(w.KX, p->ldk, w.T, p->ldk, ot.n, C, ot.kernel, ot.outputscale, ot.noise,
                       ystd2, w.var, w.sd, w.zown, st)); }

  for (int c0 = 0; c0 < C; c0 += w.chunk_C) {
    const int cc = (C - c0) < w.chunk_C ? (C - c0) : w.chunk_C;
    const int cc_pad = round_up(cc, GEMM_BM);
    CovEpilogue ep{};
    ep.xs = w.Xs + (size_t)c0 * d;
    ep.xd_s = p->xd_s;
    ep.sd = w.sd + c0;
    ep.Z = w.Z;
    ep.ldz = p->ldz; ep.C = cc; ep.N = N; ep.d = d; ep.kind = ot.kernel;
    ep.outputscale = ot.outputscale; ep.ystd2 = ystd2;
    {
      // cov[c, n] = k(x_c, x_n) - T[c, :] . k(X_train, x_n)   with T = K^-1 k(X_train, x_c)
      if (ot.Kxd_dig != nullptr) {
        // int8 tensor cores in product mode: Z <- T . Kxd; the row-statistics pass below turns
        // the products into slopes while it reads them (CovFinish)
        { ProfScope pd(10, st);
          DKG_TRY(ozaki_slice_rows(w.T + (size_t)c0 * p->ldk, p->ldk, cc, ot.n, 128, p->cov_digits, w.T_dig,
                                   w.T_scale, st)); }
        ProfScope ps(3, st);
        DKG_TRY(ozaki_store(w.T_dig, w.T_scale, cc_pad, ot.Kxd_dig, ot.Kxd_scale, p->N_pad, ot.n, p->cov_digits,
                            p->cov_diagonals, /*b_nonneg=*/true, w.Z, p->ldz, cc, N, st));
      } else {
        ProfScope ps(3, st);
        DKG_TRY(cov_rows(ot, p, w, w.T + (size_t)c0 * p->ldk, cc, cc_pad, ep, st));
      } }
    { ProfScope ps(4, st); DKG_TRY(launch_place_own(w.zown + c0, cc, w.Z, p->ldz, N, st)); }

    LineBatch lb;
    lb.Z = w.Z; lb.ldz = p->ldz;
    lb.A = p->A0; lb.a_sc = 0; lb.a_sj = p->N_pad;
    lb.A32 = p->A0f;
    lb.A32tmax = p->A0tmax; lb.A32targ = p->A0targ; lb.a32_tiles = p->a0_tiles;
    lb.a_own = w.a_new + (size_t)c0 * S;
    lb.wt = p->wt;
    lb.Amax = p->A0max; lb.Aarg = p->A0arg; lb.am_sc = 0;
    lb.NA = N; lb.NL = N + 1; lb.S = S; lb.C = cc;
    EmaxScratch sc;
    sc.zst = w.zst; sc.zarg = w.zarg; sc.surv_cnt = w.surv_cnt; sc.surv = (SurvEntry*)w.surv;
    sc.ovf_sets = w.ovf_sets; sc.ovf_count = w.ovf_count; sc.far = w.far; sc.chain = (double4*)w.chain;
    sc.long_sets = w.long_sets; sc.long_count = w.long_sets + (size_t)w.chunk_C * S;
    sc.chain32 = (float4*)w.chain32;
    sc.chainv = (double4*)w.chainv;
    sc.zpv = w.zpv; sc.zpi = w.zpi;
    sc.chain5 = (double4*)w.chain5;
    sc.chain5f = (float4*)w.chain5f;
    sc.ztile = (float2*)w.ztile; sc.ztiles = w.ztiles;
    sc.stats = w.stats;
    sc.spill_used = w.spill_used;
    CovFinish fin{};
    fin.xs = ep.xs; fin.xd_s = ep.xd_s; fin.sd = ep.sd; fin.d = d; fin.kind = ep.kind; fin.N = N;
    fin.outputscale = ep.outputscale; fin.ystd2 = ep.ystd2;
    bool ztile_valid = false;
    { ProfScope ps(5, st);
      DKG_TRY(emax_zstat(lb, sc, nullptr, nullptr, st, ot.Kxd_dig != nullptr ? &fin : nullptr, &ztile_valid)); }
    { ProfScope ps(6, st); DKG_TRY(emax_filter(lb, sc, st, ztile_valid)); }
    EmaxOut out;
    out.terms = w.kg_terms + (size_t)c0 * S;
    out.subtract_max = 1;
    out.hull_cnt = w.hull_cnt;  // hull records are per chunk (consumed by finalize below)
    if (dX != nullptr) {        // only the backward reads the records
      out.hull_idx = w.hull_idx;
      out.hull_p = w.hull_p;
      out.hull_q = w.hull_q;
    }
    out.hull_x = nullptr;
    out.hull_cap = HULL_CAP;
    out.amax_is_own = w.amax_is_new;
    out.kg = kg + c0;
    out.truncated = w.stats + 6;
    if (dX != nullptr) {
      out.spill_head = w.spill_head; out.spill_next = w.spill_next; out.spill_idx = w.spill_idx;
      out.spill_p = w.spill_p; out.spill_q = w.spill_q; out.spill_used = w.spill_used;
      out.spill_blocks = w.spill_blocks;
    }
    BackwardArgs bw{};
    if (dX != nullptr) {
      bw.dX = dX + (size_t)c0 * d;
      bw.X = X + (size_t)c0 * d;
      bw.T = w.T + (size_t)c0 * p->ldk;
      bw.ldk = p->ldk;
      bw.BT = p->BT;
      bw.n_pad = ot.n_pad;
      bw.ldbt = ot.ldbt;
      bw.xd_s = p->xd_s;
      bw.var = w.var + c0;
      bw.sd = w.sd + c0;
      bw.W = p->W;
      bw.M = M; bw.d = d; bw.target = p->target;
      for (int m = 0; m < M; ++m) {
        const ObjState& o = p->obj[m];
        bw.xs[m] = o.xs; bw.alpha[m] = o.alpha; bw.ntr[m] = o.n; bw.kind[m] = o.kernel;
        bw.outputscale[m] = o.outputscale; bw.y_std[m] = o.y_std;
        for (int k = 0; k < MAX_D; ++k) bw.ls[m][k] = o.ls[k];
      }
    }
    // (survivors per set of the previous forward with gradients, read without synchronisation: a stale or
    // torn value only changes which hull kernels are launched, never a result)
    const double hint = (w.stats_pinned != nullptr && w.stats_pinned[0] >= 0 && w.spill_checked && w.last_C > 0)
                            ? (double)w.stats_pinned[1] / ((double)w.last_C * S) : -1.0;
    { ProfScope ps(7, st); DKG_TRY(emax_hull(lb, sc, out, st, hint)); }
    { ProfScope ps(8, st); DKG_TRY(emax_overflow(lb, sc, out, st)); }
    { ProfScope ps(9, st); DKG_TRY(emax_finalize(lb, out, bw, st)); }
  }
  return DKG_OK;
}

// Coupled evaluation (reference calculate_discrete_kg, discretekg.py:162-235).
static int forward_coupled(dkg_plan* p, const double* X, int C, double* kg, double* dX, cudaStream_t st) {
  Workspace& w = p->ws;
  const int d = p->d, N = p->N, S = p->S, M = p->M;
  const int C_pad = round_up(C, GEMM_BM);
  w.last_C = C;
  DKG_CUDA_OK(cudaMemsetAsync(w.stats, 0, sizeof(long long) * STATS_WORDS, st));
  // per objective: k_m(x, X_train), means, T_m = KX_m Kinv_m, latent variance
  for (int m = 0; m < M; ++m) {
    const ObjState& o = p->obj[m];
    XprepArgs xa{};
    xa.X = X; xa.C = C; xa.d = d; xa.M = M; xa.S = S; xa.target = m;
    for (int q = 0; q < M; ++q) {
      const ObjState& oq = p->obj[q];
      xa.xs[q] = oq.xs; xa.alpha[q] = oq.alpha; xa.ntr[q] = oq.n; xa.kind[q] = oq.kernel;
      xa.outputscale[q] = oq.outputscale; xa.mean_const[q] = oq.mean_const;
      xa.y_mean[q] = oq.y_mean; xa.y_std[q] = oq.y_std;
      for (int k = 0; k < MAX_D; ++k) xa.ls[q][k] = oq.ls[k];
    }
    xa.W = p->W; xa.Xs = w.Xs; xa.KX = w.KXm[m]; xa.n_pad = o.ldk; xa.a_new = w.a_new;
    { ProfScope ps(0, st); DKG_TRY(launch_xprep(xa, st)); }  // (a_new / means are recomputed identically each time)
    { ProfScope ps(1, st); DKG_TRY(solve_T(p, w, o, w.KXm[m], w.Tm[m], w.R, C, C_pad, st)); }
    // var = noisy variance (un-standardised) -> w.var reused per objective below via varlat
    DKG_TRY(launch_var(w.KXm[m], o.ldk, w.Tm[m], o.ldk, o.n, C, o.kernel, o.outputscale, o.noise,
                       o.y_std * o.y_std, w.varlat[m], w.sd, w.zown, st));
    // launch_var wrote: varlat[m] <- noisy variance, sd <- sqrt, zown <- Cov(x,x)/sd.  Recover the
    // covariance Cov_m(x, x) (un-standardised) into column N of COV_m per chunk below: zown * sd.
    // (kept simple: a tiny kernel multiplies them in place)
  }
  for (int c0 = 0; c0 < C; c0 += w.chunk_C) {
    const int cc = (C - c0) < w.chunk_C ? (C - c0) : w.chunk_C;
    const int cc_pad = round_up(cc, GEMM_BM);
    CoupledArgs ca;
    ca.C = cc; ca.S = S; ca.M = M; ca.d = d; ca.N = N; ca.ldz = p->ldz;
    ca.W = p->W; ca.W2 = p->W2; ca.sdj = w.sdj + (size_t)c0 * S; ca.Zc = w.Zc;  // Zc == nullptr: statistics only
    ca.zpv = w.zpv; ca.zpi = w.zpi; ca.zst = w.zst; ca.zarg = w.zarg;
    for (int m = 0; m < M; ++m) {
      const ObjState& o = p->obj[m];
      // covariance rows of objective m: same GEMM as the decoupled path with sd = 1
      DKG_TRY(launch_xprep_scaled(X + (size_t)c0 * d, cc, d, o.ls, w.Xs, st));
      DKG_TRY(fill_ones(w.sd, cc, st));
      CovEpilogue ep{};
      ep.xs = w.Xs; ep.xd_s = o.xd_s; ep.sd = w.sd; ep.Z = w.COVm[m];
      ep.ldz = p->ldz; ep.C = cc; ep.N = N; ep.d = d; ep.kind = o.kernel;
      ep.outputscale = o.outputscale; ep.ystd2 = o.y_std * o.y_std;
      { ProfScope ps(3, st); DKG_TRY(cov_rows(o, p, w, w.Tm[m] + (size_t)c0 * o.ldk, cc, cc_pad, ep, st)); }
      DKG_TRY(place_latent_var(w.KXm[m] + (size_t)c0 * o.ldk, o.ldk, w.Tm[m] + (size_t)c0 * o.ldk,
                               o.ldk, o.n, cc, o.kernel, o.outputscale, o.y_std * o.y_std,
                               w.COVm[m], p->ldz, N, st));
      ca.COV[m] = w.COVm[m];
      ca.varn[m] = w.varlat[m] + c0;
    }
    { ProfScope ps(4, st); DKG_TRY(coupled_slopes(ca, st)); }  // (profile category 4: slope assembly)

    LineBatch lb;
    lb.Z = w.Zc; lb.ldz = p->ldz;
    if (w.Zc == nullptr) {  // slopes formed on the fly from the covariance rows (line_slope)
      lb.cov_M = M; lb.cov_w2 = p->W2; lb.cov_sd = ca.sdj;
      for (int m = 0; m < M; ++m) lb.cov[m] = w.COVm[m];
    }
    lb.A = p->A0; lb.a_sc = 0; lb.a_sj = p->N_pad;
    lb.A32 = p->A0f;
    lb.A32tmax = p->A0tmax; lb.A32targ = p->A0targ; lb.a32_tiles = p->a0_tiles;
    lb.a_own = w.a_new + (size_t)c0 * S;
    lb.wt = nullptr;
    lb.Amax = p->A0max; lb.Aarg = p->A0arg; lb.am_sc = 0;
    lb.NA = N; lb.NL = N + 1; lb.S = 1; lb.C = cc * S; lb.row_mod = S;
    EmaxScratch sc;
    sc.zst = w.zst; sc.zarg = w.zarg; sc.surv_cnt = w.surv_cnt; sc.surv = (SurvEntry*)w.surv;
    sc.ovf_sets = w.ovf_sets; sc.ovf_count = w.ovf_count; sc.far = w.far; sc.chain = (double4*)w.chain;
    sc.long_sets = w.long_sets; sc.long_count = w.long_sets + (size_t)w.chunk_C * S;
    sc.chain32 = (float4*)w.chain32; sc.chainv = (double4*)w.chainv; sc.chain5 = (double4*)w.chain5;
    sc.stats = w.stats;
    sc.spill_used = w.spill_used;
    sc.zpv = w.zpv; sc.zpi = w.zpi;
    { ProfScope ps(5, st); DKG_TRY(emax_zstat_from_partials(lb, sc, coupled_stat_segments(N, w.Zc != nullptr), st)); }
    { ProfScope ps(6, st); DKG_TRY(emax_filter(lb, sc, st)); }
    EmaxOut out;
    out.terms = w.kg_terms + (size_t)c0 * S;
    out.subtract_max = 1;
    out.hull_cnt = w.hull_cnt;
    if (dX != nullptr) { out.hull_idx = w.hull_idx; out.hull_p = w.hull_p; out.hull_q = w.hull_q; }
    out.hull_x = nullptr; out.hull_cap = HULL_CAP; out.amax_is_own = w.amax_is_new;
    out.kg = kg + c0;
    out.truncated = w.stats + 6;
    if (dX != nullptr) {
      out.spill_head = w.spill_head; out.spill_next = w.spill_next; out.spill_idx = w.spill_idx;
      out.spill_p = w.spill_p; out.spill_q = w.spill_q; out.spill_used = w.spill_used;
      out.spill_blocks = w.spill_blocks;
    }
    { ProfScope ps(7, st); DKG_TRY(emax_hull(lb, sc, out, st)); }
    { ProfScope ps(8, st); DKG_TRY(emax_overflow(lb, sc, out, st)); }
    CoupledBackward bw;
    if (dX != nullptr) {
      bw.dX = dX + (size_t)c0 * d; bw.X = X + (size_t)c0 * d; bw.W = p->W; bw.W2 = p->W2;
      bw.sdj = w.sdj + (size_t)c0 * S; bw.ldz = p->ldz; bw.M = M; bw.d = d; bw.S = S; bw.N = N;
      for (int m = 0; m < M; ++m) {
        const ObjState& o = p->obj[m];
        bw.COV[m] = w.COVm[m];
        bw.T[m] = w.Tm[m] + (size_t)c0 * o.ldk; bw.ldk[m] = o.ldk; bw.BT[m] = o.BT; bw.n_pad[m] = o.n_pad; bw.ldbt[m] = o.ldbt;
        bw.xd_s[m] = o.xd_s; bw.xs[m] = o.xs; bw.alpha[m] = o.alpha; bw.ntr[m] = o.n; bw.kind[m] = o.kernel;
        bw.outputscale[m] = o.outputscale; bw.y_std[m] = o.y_std;
        for (int k = 0; k < MAX_D; ++k) bw.ls[m][k] = o.ls[k];
      }
    }
    { ProfScope ps(9, st); DKG_TRY(emax_finalize_coupled(cc, S, out, bw, st)); }
  }
  return DKG_OK;
}

}  // namespace dkg

using namespace dkg;

extern "C" {

int dkg_abi_version(void) { return DKG_ABI_VERSION; }
const char* dkg_last_error(void) { return g_err; }
int64_t dkg_launch_count(void) { return g_launches.load(); }
void dkg_launch_count_reset(void) { g_launches.store(0); }

int dkg_plan_create(const dkg_objective* objs, int32_t M, int32_t d, const double* x_disc_dev,
                    int32_t N, const double* weights_host, int32_t S, int32_t target_ix,
                    uint32_t flags, void* stream, dkg_plan** out_plan) {
  if (!out_plan) { set_error("out_plan is NULL"); return DKG_EINVAL; }
  *out_plan = nullptr;
  if (!objs || !x_disc_dev || !weights_host) { set_error("NULL argument"); return DKG_EINVAL; }
  if (M < 1 || M > MAX_M) { set_error("M=%d outside [1, %d]", M, MAX_M); return DKG_EINVAL; }
  if (d < 1 || d > MAX_D) { set_error("d=%d outside [1, %d]", d, MAX_D); return DKG_EINVAL; }
  if (S < 1 || S > MAX_S) { set_error("S=%d outside [1, %d]", S, MAX_S); return DKG_EINVAL; }
  if (N < 1) { set_error("the discretisation is empty"); return DKG_EINVAL; }
  if (target_ix < -1 || target_ix >= M) {
    set_error("target_ix=%d outside [-1, %d)  (-1 = coupled evaluation)", target_ix, M);
    return DKG_EINVAL;
  }
  for (int m = 0; m < M; ++m) {
    if (objs[m].n < 1 || !objs[m].train_x_dev || !objs[m].train_y_dev || !objs[m].lengthscale_host) {
      set_error("objective %d: empty or NULL training data", m);
      return DKG_EINVAL;
    }
    if (objs[m].kernel != DKG_KERNEL_MATERN52 && objs[m].kernel != DKG_KERNEL_RBF) {
      set_error("objective %d: unsupported kernel id %d", m, objs[m].kernel);
      return DKG_EINVAL;
    }
    if (!(objs[m].noise >= 0.0) || !(objs[m].outputscale > 0.0) || !(objs[m].y_std > 0.0)) {
      set_error("objective %d: noise/outputscale/y_std out of range", m);
      return DKG_EINVAL;
    }
    for (int k = 0; k < d; ++k)
      if (!(objs[m].lengthscale_host[k] > 0.0)) {
        set_error("objective %d: lengthscale[%d] must be positive", m, k);
        return DKG_EINVAL;
      }
  }
  dkg_plan* p = new (std::nothrow) dkg_plan();
  if (!p) { set_error("out of host memory"); return DKG_ENOMEM; }
  p->M = M; p->d = d; p->N = N; p->S = S; p->target = target_ix;
  static_assert(OZ_DEFAULT_DIGITS == 7 && OZ_DEFAULT_DIAGONALS == 7, "dkg_plan's default digit configuration");
  if (flags & DKG_PLAN_FAST32) { p->cov_digits = 4; p->cov_diagonals = 4; }
  else if (const char* e = getenv("DKG_OZ_DIAGONALS")) {  // measurement switch (8 = the round-1 cut-off)
    const int v = atoi(e);
    if (v >= 1 && v <= 2 * p->cov_digits - 1) p->cov_diagonals = v;
  }
  p->N_pad = round_up(N, GEMM_BN);
  p->ldz = round_up(N + 1, 16);
  cudaGetDevice(&p->device);
  memcpy(p->W_host, weights_host, sizeof(double) * S * M);
  int rc = build_plan(p, objs, x_disc_dev, (cudaStream_t)stream);
  if (rc != DKG_OK) {
    destroy_plan(p);
    return rc;
  }
  *out_plan = p;
  return DKG_OK;
}

void dkg_plan_destroy(dkg_plan* plan) { destroy_plan(plan); }

int dkg_plan_append_point(dkg_plan* plan, int32_t m, const double* x_host, double y, void* stream) {
  if (!plan || !x_host) { set_error("NULL argument"); return DKG_EINVAL; }
  if (m < 0 || m >= plan->M) { set_error("objective %d outside [0, %d)", m, plan->M); return DKG_EINVAL; }
  dkg_plan* p = plan;
  ObjState& o = p->obj[m];
  cudaStream_t st = (cudaStream_t)stream;
  const int n = o.n, d = p->d, N = p->N;
  const bool state = (m == p->target || p->target < 0);
  if (o.Kinv == nullptr || o.Kmat == nullptr || n + 1 > o.cap || n + 1 > o.ldk || (state && o.Kxd_dig != nullptr && n + 1 > OZ_MAX_K)) {
    set_error("no room for another training point of objective %d (n = %d, capacity %d): build a new plan", m, n, o.cap);
    return DKG_ECAPACITY;
  }
  DKG_CUDA_OK(cudaStreamSynchronize(st));  // nothing in flight may still read the tables that change
  double xq[MAX_D];
  for (int k = 0; k < d; ++k) xq[k] = x_host[k] / o.ls[k];
  double *scr = nullptr;  // [kv | v | tmp | scal(4)] + [rrow | w] for a state objective
  const size_t nv = (size_t)round_up(n + 1, 16);
  DKG_TRY(dev_alloc(&scr, 3 * nv + 4 + (state ? 2 * (size_t)p->N_pad : 0)));
  double *kv = scr, *v = scr + nv, *tmp = scr + 2 * nv, *scal = scr + 3 * nv, *rrow = scal + 4, *wv = rrow + p->N_pad;
  int rc = DKG_OK;
  auto A = [&](int r) { if (rc == DKG_OK) rc = r; };
  // the new row of the scaled inputs is written first: nothing reads beyond row n - 1 until n grows
  if (cudaMemcpyAsync(o.xs + (size_t)n * d, xq, sizeof(double) * d, cudaMemcpyHostToDevice, st) != cudaSuccess) rc = DKG_ECUDA;
  if (rc == DKG_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = DKG_ECUDA;  // xq is a stack buffer
  const double* xq_dev = o.xs + (size_t)n * d;
  const double kappa = o.outputscale + o.noise + o.jitter;  // k(x, x) of a stationary kernel + noise
  A(kernel_row(xq_dev, o.xs, n, d, o.kernel, o.outputscale, kv, st));
  // v = K^-1 k with one refinement step (as solve_T)
  A(matvec_axpy(o.Kinv, o.ldk, n, n, kv, nullptr, 1.0, v, st));
  A(matvec_axpy(o.Kmat, o.ldk, n, n, v, kv, -1.0, tmp, st));
  A(matvec_axpy(o.Kinv, o.ldk, n, n, tmp, v, 1.0, v, st));
  const double yc = y - o.mean_const;
  A(append_scalars(kv, v, o.alpha, n, kappa, yc, scal, st));
  double sh[4] = {0, 0, 0, 0};
  if (rc == DKG_OK) {
    if (cudaMemcpyAsync(sh, scal, sizeof(double) * 3, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) rc = DKG_ECUDA;
  }
  if (rc == DKG_OK && !(sh[0] > 0.0)) {
    set_error("the covariance extended by the new point is not positive definite (Schur complement %g)", sh[0]);
    rc = DKG_ENOTPD;
  }
  if (rc != DKG_OK) { cudaStreamSynchronize(st); dev_free(scr); return rc; }
  // ---- from here on the plan changes ----
  A(append_update_k(o.Kinv, o.Kmat, o.ldk, n, kv, v, scal, kappa, st));
  A(append_update_alpha(o.alpha, v, scal, n, st));
  if (rc == DKG_OK && cudaMemcpyAsync(o.resid + n, &yc, sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess) rc = DKG_ECUDA;
  if (rc == DKG_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = DKG_ECUDA;
  // mean cache: one refinement step on the extended system  alpha += K'^-1 (r - K' alpha)
  A(matvec_axpy(o.Kmat, o.ldk, n + 1, n + 1, o.alpha, o.resid, -1.0, tmp, st));
  A(matvec_axpy(o.Kinv, o.ldk, n + 1, n + 1, tmp, o.alpha, 1.0, o.alpha, st));
  if (state) {
    A(kernel_row(xq_dev, o.xd_s, N, d, o.kernel, o.outputscale, rrow, st));
    A(append_w(o.Kxd, p->N_pad, n, N, v, rrow, scal, wv, st));
    A(append_update_b(o.B, p->N_pad, o.BT, o.ldbt, n, N, v, wv, st));
  }
  o.n = n + 1;
  o.n_pad = round_up(o.n, GEMM_BK);
  // repeated block updates drift by ~eps cond(K) per step: the forward's refinement step (solve_T) makes
  // T = K^-1 k_x insensitive to that, so it is switched on for good once a plan has been extended
  o.refine = true;
  dev_free(o.chol);  // the Cholesky factor is not maintained (only the DKG_T_SOLVE=trsm experiment reads it)
  if (state) A(make_kxd_digits(o, p, st));
  A(mu_disc(p->xd, N, d, o, p->mu_disc, p->M, m, st));
  A(build_a0(p->mu_disc, N, p->M, p->W, p->S, p->A0, p->A0f, p->N_pad, p->A0max, p->A0arg, st));
  A(build_a0_tilemax(p->A0f, p->N_pad, p->S, FILTER_TILE, p->a0_tiles, p->A0tmax, p->A0targ, st));
  if (m == p->target) { p->chol = o.chol; }
  drop_graphs(p->ws);  // captured launches carry n
  cudaError_t e = cudaStreamSynchronize(st);
  dev_free(scr);
  if (rc == DKG_OK && e != cudaSuccess) { set_error("dkg_plan_append_point: %s", cudaGetErrorString(e)); rc = DKG_ECUDA; }
  return rc;
}

int dkg_forward_dev(dkg_plan* plan, const double* X_dev, int32_t C, double* kg_dev, double* dX_dev,
                    void* stream) {
  if (!plan || (C > 0 && (!X_dev || !kg_dev))) { set_error("NULL argument"); return DKG_EINVAL; }
  if (C < 0) { set_error("C=%d is negative", C); return DKG_EINVAL; }
  return forward_impl(plan, X_dev, C, kg_dev, dX_dev, (cudaStream_t)stream);
}

int dkg_forward_host(dkg_plan* plan, const double* X_host, int32_t C, double* kg_host,
                     double* dX_host, void* stream) {
  if (!plan || (C > 0 && (!X_host || !kg_host))) { set_error("NULL argument"); return DKG_EINVAL; }
  if (C < 0) { set_error("C=%d is negative", C); return DKG_EINVAL; }
  if (C == 0) return DKG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  DKG_TRY(spill_precheck(plan));
  DKG_TRY(ensure_workspace(plan, C));
  Workspace& w = plan->ws;
  const int d = plan->d;
  DKG_CUDA_OK(cudaMemcpyAsync(w.X, X_host, sizeof(double) * (size_t)C * d, cudaMemcpyHostToDevice, st));
  for (int attempt = 0;; ++attempt) {
    DKG_TRY(enqueue_forward(plan, w.X, C, w.kg, dX_host ? w.dX : nullptr, st));
    DKG_CUDA_OK(cudaMemcpyAsync(kg_host, w.kg, sizeof(double) * (size_t)C, cudaMemcpyDeviceToHost, st));
    if (dX_host) {
      DKG_CUDA_OK(cudaMemcpyAsync(dX_host, w.dX, sizeof(double) * (size_t)C * d, cudaMemcpyDeviceToHost, st));
      DKG_CUDA_OK(cudaMemcpyAsync(w.stats_pinned, w.stats, sizeof(long long) * 8, cudaMemcpyDeviceToHost, st));
    }
    DKG_CUDA_OK(cudaStreamSynchronize(st));
    if (!dX_host) break;
    // the one sync of this entry point also tells whether hull records were dropped: re-run with a
    // larger spill pool instead of handing back NaN gradient rows
    w.spill_checked = true;
    w.stats_pending = false;
    int grew = 0;
    DKG_TRY(grow_spill_if_needed(plan, w.stats_pinned, &grew));
    if (!grew || attempt >= 3) {
      if (w.stats_pinned[6] > 0) {
        set_error("%lld (candidate, scalarisation) sets have more upper-envelope vertices than the hull-record spill "
                  "pool holds; their gradient rows are NaN.  Raise or unset DKG_SPILL_BLOCKS (32 vertices per block).",
                  w.stats_pinned[6]);
        return DKG_ETRUNC;
      }
      break;
    }
  }
  return DKG_OK;
}

int dkg_posterior_mean_dev(dkg_plan* plan, const double* X_dev, int32_t C, double* mu_dev, void* stream) {
  if (!plan || (C > 0 && (!X_dev || !mu_dev))) { set_error("NULL argument"); return DKG_EINVAL; }
  if (C < 0) { set_error("C=%d is negative", C); return DKG_EINVAL; }
  XprepArgs xa{};
  xa.X = X_dev; xa.C = C; xa.d = plan->d; xa.M = plan->M; xa.S = plan->S; xa.target = 0;
  for (int m = 0; m < plan->M; ++m) {
    const ObjState& o = plan->obj[m];
    xa.xs[m] = o.xs; xa.alpha[m] = o.alpha; xa.ntr[m] = o.n; xa.kind[m] = o.kernel;
    xa.outputscale[m] = o.outputscale; xa.mean_const[m] = o.mean_const;
    xa.y_mean[m] = o.y_mean; xa.y_std[m] = o.y_std;
    for (int k = 0; k < MAX_D; ++k) xa.ls[m][k] = o.ls[k];
  }
  return launch_mean(xa, mu_dev, (cudaStream_t)stream);
}

int dkg_int8_matmul_dev(const double* A_dev, int32_t lda, const double* Bt_dev, int32_t ldb, int32_t M,
                        int32_t N, int32_t K, int32_t n_digits, int32_t n_diagonals, double* D_dev,
                        int32_t ldd, void* stream) {
  if (n_digits == 0) n_digits = OZ_DEFAULT_DIGITS;
  if (n_diagonals == 0) n_diagonals = OZ_DEFAULT_DIAGONALS < 2 * n_digits - 1 ? OZ_DEFAULT_DIAGONALS : 2 * n_digits - 1;
  if (!A_dev || !Bt_dev || !D_dev || M <= 0 || N <= 0 || K <= 0 || K > OZ_MAX_K || lda < K || ldb < K ||
      ldd < N || n_digits < 1 || n_digits > 7 || n_diagonals < 1 || n_diagonals > 2 * n_digits - 1) {
    set_error("dkg_int8_matmul_dev: invalid argument");
    return DKG_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int M_pad = round_up(M, GEMM_BM), N_pad = round_up(N, GEMM_BN);
  unsigned char *da = nullptr, *db = nullptr;
  double *sa = nullptr, *sb = nullptr;
  int rc = dev_alloc(&da, ozaki_digit_bytes(M_pad, K, n_digits));
  if (rc == DKG_OK) rc = dev_alloc(&db, ozaki_digit_bytes(N_pad, K, n_digits));
  if (rc == DKG_OK) rc = dev_alloc(&sa, (size_t)M_pad + 128);
  if (rc == DKG_OK) rc = dev_alloc(&sb, (size_t)N_pad);
  if (rc == DKG_OK) rc = ozaki_slice_rows(A_dev, lda, M, K, 128, n_digits, da, sa, st);
  if (rc == DKG_OK) rc = ozaki_slice_rows(Bt_dev, ldb, N, K, ozaki_b_block_rows(n_digits, n_diagonals), n_digits, db, sb, st);
  int* neg_dev = nullptr;
  int neg = 1;
  if (rc == DKG_OK) rc = dev_alloc(&neg_dev, 1);
  if (rc == DKG_OK) rc = ozaki_any_negative(Bt_dev, ldb, N, K, neg_dev, st);
  if (rc == DKG_OK) {
    cudaMemcpyAsync(&neg, neg_dev, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
  }
  if (rc == DKG_OK)
    rc = ozaki_store(da, sa, M_pad, db, sb, N_pad, K, n_digits, n_diagonals, /*b_nonneg=*/neg == 0, D_dev, ldd, M, N, st);
  cudaError_t e = cudaStreamSynchronize(st);
  if (rc == DKG_OK && e != cudaSuccess) {
    set_error("dkg_int8_matmul_dev: %s", cudaGetErrorString(e));
    rc = DKG_ECUDA;
  }
  dev_free(da); dev_free(db); dev_free(sa); dev_free(sb); dev_free(neg_dev);
  return rc;
}

int dkg_expected_max_lines_dev(const double* a_dev, const double* b_dev, int32_t P, int32_t L,
                               double* emax_dev, int32_t* hull_count_dev, int32_t* hull_idx_dev,
                               double* hull_x_dev, int32_t hull_cap, double* dE_da_dev,
                               double* dE_db_dev, void* stream) {
  if (L == 0) {
    set_error("Expected inputs to specify at least one line. Got intercepts.shape[-1]=0.");
    return DKG_EEMPTY;
  }
  if (P < 0 || L < 0 || hull_cap < 0) { set_error("negative size"); return DKG_EINVAL; }
  if (P == 0) return DKG_OK;
  if (!a_dev || !b_dev || !emax_dev) { set_error("NULL argument"); return DKG_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  double *zst = nullptr, *amax = nullptr;
  int *zarg = nullptr, *aarg = nullptr, *scnt = nullptr, *oset = nullptr, *ocnt = nullptr;
  SurvEntry* surv = nullptr;
  unsigned long long* far = nullptr;
  double4* chain = nullptr;
  int rc = DKG_OK;
  auto A = [&](int r) { if (rc == DKG_OK) rc = r; };
  A(dev_alloc(&zst, (size_t)P * 2));
  A(dev_alloc(&zarg, (size_t)P * 2));
  A(dev_alloc(&amax, (size_t)P));
  A(dev_alloc(&aarg, (size_t)P));
  A(dev_alloc(&scnt, (size_t)P));
  A(dev_alloc(&surv, (size_t)P * SURV_CAP, false));
  A(dev_alloc(&oset, (size_t)P, false));
  A(dev_alloc(&far, (size_t)P * 2));
  A(dev_alloc(&chain, (size_t)P, false));
  A(dev_alloc(&ocnt, (size_t)1));
  if (rc == DKG_OK) {
    LineBatch lb;
    lb.Z = b_dev; lb.ldz = L;
    lb.A = a_dev; lb.a_sc = L; lb.a_sj = 0;
    lb.a_own = nullptr; lb.wt = nullptr;
    lb.Amax = amax; lb.Aarg = aarg; lb.am_sc = 1;
    lb.NA = L; lb.NL = L; lb.S = 1; lb.C = P;
    EmaxScratch sc;
    sc.zst = zst; sc.zarg = zarg; sc.surv_cnt = scnt; sc.surv = surv;
    sc.ovf_sets = oset; sc.ovf_count = ocnt; sc.far = far; sc.chain = chain; sc.stats = nullptr;
    EmaxOut out;
    out.terms = emax_dev; out.subtract_max = 0;
    out.hull_cnt = hull_count_dev;
    out.hull_idx = hull_idx_dev; out.hull_x = hull_x_dev; out.hull_cap = hull_cap;
    out.dense_da = dE_da_dev; out.dense_db = dE_db_dev;
    if (dE_da_dev) cudaMemsetAsync(dE_da_dev, 0, sizeof(double) * (size_t)P * L, st);
    if (dE_db_dev) cudaMemsetAsync(dE_db_dev, 0, sizeof(double) * (size_t)P * L, st);
    A(emax_zstat(lb, sc, amax, aarg, st));
    A(emax_filter(lb, sc, st));
    A(emax_hull(lb, sc, out, st));
    A(emax_overflow(lb, sc, out, st));
  }
  cudaError_t e = cudaStreamSynchronize(st);
  if (rc == DKG_OK && e != cudaSuccess) {
    set_error("expected-max kernels failed: %s", cudaGetErrorString(e));
    rc = DKG_ECUDA;
  }
  dev_free(zst); dev_free(zarg); dev_free(amax); dev_free(aarg); dev_free(scnt); dev_free(surv);
  dev_free(oset); dev_free(ocnt); dev_free(far); dev_free(chain);
  return rc;
}

int dkg_int8_peak(int32_t M, int32_t N, int32_t K, int32_t reps, int32_t mode, double* tops_host, double* ms_host,
                  void* stream) {
  if (M <= 0 || N <= 0 || K <= 0 || K > OZ_MAX_K || reps < 1 || !tops_host) { set_error("dkg_int8_peak: invalid argument"); return DKG_EINVAL; }
  return ozaki_mma_peak(round_up(M, 256), round_up(N, GEMM_BN), K, OZ_DEFAULT_DIGITS, OZ_DEFAULT_DIAGONALS, reps, mode, tops_host,
                        ms_host, (cudaStream_t)stream);
}

int dkg_piecewise_expectation_dev(const double* a_dev, const double* b_dev, const double* z_dev, int32_t P,
                                  int32_t H, double* e_dev, double* dE_da_dev, double* dE_db_dev,
                                  double* dE_dz_dev, void* stream) {
  if (H == 0) {
    set_error("Expected inputs to specify at least one line. Got intercepts.shape[-1]=0.");
    return DKG_EEMPTY;
  }
  if (P < 0 || H < 0) { set_error("negative size"); return DKG_EINVAL; }
  if (P == 0) return DKG_OK;
  if (!a_dev || !b_dev || !e_dev || (H > 1 && !z_dev)) { set_error("NULL argument"); return DKG_EINVAL; }
  return piecewise_expectation(a_dev, b_dev, z_dev, P, H, e_dev, dE_da_dev, dE_db_dev, dE_dz_dev,
                               (cudaStream_t)stream);
}

int64_t dkg_plan_read(dkg_plan* plan, const char* name, double* out_dev, int64_t capacity,
                      void* stream) {
  if (!plan || !name) { set_error("NULL argument"); return DKG_EINVAL; }
  const std::string nm0(name);
  if (plan->target < 0 && (nm0 == "B" || nm0 == "Kinv" || nm0 == "chol" || nm0 == "slopes" || nm0 == "var")) {
    set_error("tensor '%s' is per target objective; not available for the coupled plan", name);
    return DKG_EINVAL;
  }
  const ObjState& ot = plan->obj[plan->target < 0 ? 0 : plan->target];
  const Workspace& w = plan->ws;
  const double* src = nullptr;
  int64_t rows = 0, cols = 0, ld = 0;
  int n_sum = 0;
  for (int m = 0; m < plan->M; ++m) n_sum += plan->obj[m].n;
  const std::string nm(name);
  if (nm == "B") { src = plan->B; rows = ot.n; cols = plan->N; ld = plan->N_pad; }
  else if (nm == "Kinv") { src = plan->Kinv; rows = ot.n; cols = ot.n; ld = plan->ldk; }
  else if (nm == "chol") { src = plan->chol; rows = ot.n; cols = ot.n; ld = ot.n; }
  else if (nm == "alpha") {
    int off = 0;  // (gathered at read time: appends change the objectives' lengths)
    for (int m = 0; m < plan->M; ++m) {
      if (out_dev) cudaMemcpyAsync(plan->alpha_all + off, plan->obj[m].alpha, sizeof(double) * plan->obj[m].n,
                                   cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
      off += plan->obj[m].n;
    }
    src = plan->alpha_all; rows = 1; cols = n_sum; ld = n_sum;
  }
  else if (nm == "mu_disc") { src = plan->mu_disc; rows = plan->N; cols = plan->M; ld = plan->M; }
  else if (nm == "A0") { src = plan->A0; rows = plan->S; cols = plan->N; ld = plan->N_pad; }
  else if (nm == "A0max") { src = plan->A0max; rows = 1; cols = plan->S; ld = plan->S; }
  else if (nm == "slopes") {
    if (w.last_C > w.chunk_C) { set_error("slopes are only retained for C <= chunk (%d)", w.chunk_C); return DKG_EINVAL; }
    src = w.Z; rows = w.last_C; cols = plan->N + 1; ld = plan->ldz;
  }
  else if (nm == "a_new") { src = w.a_new; rows = w.last_C; cols = plan->S; ld = plan->S; }
  else if (nm == "var") { src = w.var; rows = 1; cols = w.last_C; ld = w.last_C; }
  else if (nm == "kg_terms") { src = w.kg_terms; rows = w.last_C; cols = plan->S; ld = plan->S; }
  else { set_error("unknown tensor name '%s'", name); return DKG_EINVAL; }
  const int64_t count = rows * cols;
  if (out_dev && count > 0) {
    if (capacity < count) { set_error("capacity %lld < %lld", (long long)capacity, (long long)count); return DKG_EINVAL; }
    if (!src) { set_error("tensor '%s' not available yet", name); return DKG_EINVAL; }
    // tensors indexed by discretisation line are handed out in the CALLER's line order
    const bool by_cols = nm == "B" || nm == "A0" || nm == "slopes";
    const bool by_rows = nm == "mu_disc";
    if (plan->perm != nullptr && (by_cols || by_rows)) {
      int rc = unpermute(src, ld, rows, cols, plan->perm, plan->N, by_rows, out_dev, (cudaStream_t)stream);
      return rc == DKG_OK ? count : rc;
    }
    cudaError_t e = cudaMemcpy2DAsync(out_dev, cols * sizeof(double), src, ld * sizeof(double),
                                      cols * sizeof(double), rows, cudaMemcpyDeviceToDevice,
                                      (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("copy failed: %s", cudaGetErrorString(e)); return DKG_ECUDA; }
  }
  return count;
}

void dkg_profile_enable(int on) { g_prof_on = on != 0; }

int dkg_profile_read(double* ms_host, int64_t* count_host, int32_t ncat) {
  if (!ms_host || !count_host || ncat < 1) { set_error("bad argument"); return DKG_EINVAL; }
  DKG_CUDA_OK(cudaDeviceSynchronize());
  for (int k = 0; k < ncat; ++k) { ms_host[k] = 0.0; count_host[k] = 0; }
  for (auto& ps : g_prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ps.e0, ps.e1);
    if (ps.cat < ncat) { ms_host[ps.cat] += ms; count_host[ps.cat] += 1; }
    cudaEventDestroy(ps.e0); cudaEventDestroy(ps.e1);
  }
  g_prof.clear();
  return DKG_OK;
}

int dkg_plan_stats(dkg_plan* plan, int64_t* out5_host /* [8] */, void* stream) {
  if (!plan || !out5_host) { set_error("NULL argument"); return DKG_EINVAL; }
  long long h[8] = {0};
  if (plan->ws.stats) {
    DKG_CUDA_OK(cudaMemcpyAsync(h, plan->ws.stats, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    DKG_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
  }
  if (getenv("DKG_DEBUG_STATS")) fprintf(stderr, "[dkg] sets with a truncated survivor list: %lld\n", h[0]);
  out5_host[0] = plan->ws.last_C;
  for (int k = 1; k < 8; ++k) out5_host[k] = h[k];
  // [7]: 1 when the covariance contraction of this plan runs on the int8 tensor cores
  bool int8 = false;
  for (int m = 0; m < plan->M; ++m) int8 = int8 || plan->obj[m].Kxd_dig != nullptr;
  out5_host[7] = int8 ? 1 : 0;
  return DKG_OK;
}

}  // extern "C"
