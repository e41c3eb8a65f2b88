// Expected maximum over fantasy outcomes (discretekg.py:329-336 / :341-452) for
// [candidates x scalarisations] line sets of |X_disc|+1 lines each, plus the fused backward.
//
// The reference finds each upper envelope with a Python Jarvis march over ALL lines.  Here:
//   1. zstat   : min / max slope of each candidate's slope row (shared by all scalarisations,
//                because slopes are w_j * z with one z row per candidate, discretekg.py:321); in
//                the KG path the same pass turns the int8 products into slopes (tiled over lines
//                for large batches: zfinish_tiled_kernel + zreduce_kernel).  chain_kernel then
//                writes every set's chord chain (fp64, a conservatively rounded float image, the
//                chord end points).
//   2. filter  : one streaming, coalesced pass over the slope rows and the (L2-resident)
//                intercept table.  A line whose dual point (slope, intercept) lies on or below
//                the chain P -> T -> Q (P/Q = extreme-slope lines, T = max-intercept line; all
//                three are lines of the set) is inside the convex hull of the set, hence can
//                never be a strict vertex of the upper envelope and is dropped.  filter32_kernel
//                runs the test in float arithmetic (packed FFMA2) against the rounded chain and
//                re-tests the ~1 % that pass exactly (fp64) against a second-level chain found in
//                a sampled first launch (chain5_kernel); filter_kernel is the all-fp64 form for
//                per-row intercepts.  Survivors are appended (intercept, slope, index) to a
//                per-set list.
//   3. hull    : one warp per (candidate, scalarisation).  Long lists get up to five QuickHull-
//                style refinements (farthest survivor above each chord becomes a new chain vertex,
//                the list is re-filtered); then the warp runs the reference's march EXACTLY (same
//                ordering rule, strict-slope filter, division and tie-breaks) on what is left,
//                accumulating the closed-form expectation segment by segment and recording
//                dE/da, dE/db for the backward.
//   4. overflow: sets whose list overflowed (or stayed long) are queued and handled by a
//                cooperative CTA-per-set kernel: iterative chain refinement over ALL lines, then
//                the same warp march; if even that keeps too many lines (e.g. every line is a
//                hull vertex) a block-wide exact march over all lines finishes the job.
//   5. finalize: kg[c] = mean_j (E_j - max_n a_jn) and the envelope-theorem backward, sparse
//                over the recorded hull vertices (one CTA per candidate; distinct lines merged
//                through a shared-memory hash table).
#include "dkg_emax.cuh"

#include <cstdlib>
#include <cstring>

namespace dkg {

constexpr int E_THREADS = 256;
constexpr double SHORTCUT_TOL = 1e-9;  // discretekg.py:363
constexpr int STAGE_CAP = 128;         // lines a warp marches over from registers
constexpr int CHAIN_MAXV = 32;         // vertices of a warp's refinement chain (one lane each)
constexpr int HULL_LEVELS = 5;         // refinement passes of the warp kernel before a set is queued
#ifndef DKG_HULL_CTAS
#define DKG_HULL_CTAS 3
#endif
constexpr int HULL_CTAS = DKG_HULL_CTAS;  // resident CTAs per SM the warp hull kernel is compiled for
constexpr double EPS128 = 2.84217094304040074e-14;  // 128 * 2^-52

// ------------------------------------------------------------------------------------------
// block-wide (value, index) reductions with first-index tie-breaking
// ------------------------------------------------------------------------------------------
struct MinOp {
  __device__ static bool better(double v, int i, double bv, int bi) {
    return v < bv || (v == bv && i < bi);
  }
};
struct MaxOp {
  __device__ static bool better(double v, int i, double bv, int bi) {
    return v > bv || (v == bv && i < bi);
  }
};

template <class Op>
__device__ void block_arg_reduce(double& v, int& i, double* s_v, int* s_i) {
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (Op::better(ov, oi, v, i)) {
      v = ov;
      i = oi;
    }
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) {
    s_v[w] = v;
    s_i[w] = i;
  }
  __syncthreads();
  const int nw = blockDim.x >> 5;
  v = s_v[0];
  i = s_i[0];
  for (int k = 1; k < nw; ++k)
    if (Op::better(s_v[k], s_i[k], v, i)) {
      v = s_v[k];
      i = s_i[k];
    }
}

struct ChainMag { double amag, zmag; };  // magnitudes of the intercepts / slope coordinates on the chain
__device__ double4 chain_params(const LineBatch& lb, const EmaxScratch& sc, int c, int j, ChainMag* mag,
                                double4* verts = nullptr);
__device__ float4 chord32(double c, double m, const ChainMag& mag);

// one CTA per row: min / max (with first index) of the slope row; optionally the max intercept
template <int D>
__global__ void __launch_bounds__(E_THREADS, 5)
zstat_kernel(LineBatch lb, EmaxScratch sc, double* __restrict__ zst, int* __restrict__ zarg,
             double* __restrict__ amax_out, int* __restrict__ aarg_out, CovFinish fin) {
  __shared__ double s_v[E_THREADS / 32];
  __shared__ int s_i[E_THREADS / 32];
  const int c = blockIdx.x;
  double vmin = INFINITY, vmax = -INFINITY;
  int imin = 0x7fffffff, imax = 0x7fffffff;
  if (D > 0) {
    // finish the slopes in place (see CovFinish), then the statistics of the finished values
    double* zw = const_cast<double*>(lb.Z) + (size_t)c * lb.ldz;
    double xr[D > 0 ? D : 1];
#pragma unroll
    for (int k = 0; k < D; ++k) xr[k] = fin.xs[(size_t)c * D + k];
    const double rsd = fin.ystd2 / fin.sd[c];
    const int kind = fin.kind;
    const double os = fin.outputscale;
#pragma unroll 4
    for (int n = threadIdx.x; n < lb.NL; n += blockDim.x) {
      double v = zw[n];
      if (n < fin.N) {
        double sq = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          const double df = xr[k] - fin.xd_s[(size_t)n * D + k];
          sq = fma(df, df, sq);
        }
        v = (stationary_from_sq(kind, os, sq) - v) * rsd;
        zw[n] = v;
      }
      if (v < vmin) { vmin = v; imin = n; }
      if (v > vmax) { vmax = v; imax = n; }
    }
    __threadfence_block();
  } else {
    const double* z = lb.Z + (size_t)c * lb.ldz;
#pragma unroll 8
    for (int n = threadIdx.x; n < lb.NL; n += blockDim.x) {
      double v = z[n];
      if (v < vmin) { vmin = v; imin = n; }
      if (v > vmax) { vmax = v; imax = n; }
    }
  }
  block_arg_reduce<MinOp>(vmin, imin, s_v, s_i);
  block_arg_reduce<MaxOp>(vmax, imax, s_v, s_i);
  if (threadIdx.x == 0) {
    zst[c * 2 + 0] = vmin;
    zst[c * 2 + 1] = vmax;
    zarg[c * 2 + 0] = imin;
    zarg[c * 2 + 1] = imax;
  }
  if (amax_out != nullptr) {  // generic entry point: S == 1, intercepts differ per row
    const double* a = lb.A + (size_t)c * lb.a_sc;
    double am = -INFINITY;
    int ai = 0x7fffffff;
    for (int n = threadIdx.x; n < lb.NA; n += blockDim.x) {
      double v = a[n];
      if (v > am) { am = v; ai = n; }
    }
    block_arg_reduce<MaxOp>(am, ai, s_v, s_i);
    if (threadIdx.x == 0) {
      amax_out[c] = am;
      aarg_out[c] = ai;
    }
  }
}

// chord-chain parameters of every set (one thread per set), consumed by the filter kernels: the fp64
// chain, its conservatively rounded float image and the chord end points (second-level chain)
__global__ void __launch_bounds__(E_THREADS)
chain_kernel(LineBatch lb, EmaxScratch sc) {
  const size_t set = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (set >= (size_t)lb.C * lb.S) return;
  const int c = (int)(set / lb.S), j = (int)(set - (size_t)c * lb.S);
  // per-forward counters of the stages that follow (saves three memset launches per chunk)
  sc.surv_cnt[set] = 0;
  if (sc.far != nullptr) { sc.far[set * 2] = 0ull; sc.far[set * 2 + 1] = 0ull; }
  if (set == 0) {
    *sc.ovf_count = 0;
    if (sc.long_count != nullptr) *sc.long_count = 0;
    if (sc.spill_used != nullptr) *sc.spill_used = 0;
  }
  ChainMag mag;
  double4 verts[2];
  const double4 par = chain_params(lb, sc, c, j, &mag, verts);
  sc.chain[set] = par;
  if (sc.chainv != nullptr) {
    sc.chainv[set * 2 + 0] = verts[0];
    sc.chainv[set * 2 + 1] = verts[1];
  }
  if (sc.chain32 != nullptr) {
    sc.chain32[set * 2 + 0] = chord32(par.x, par.y, mag);
    sc.chain32[set * 2 + 1] = chord32(par.z, par.w, mag);
  }
}

// ---- tiled variant of the product-mode pass (large batches) ----------------------------------
// zstat_kernel reads the scaled discretisation point of every line (D doubles) through L2 once per
// ROW: at c4 that is 2.1 GB of L2->SM traffic next to 1.07 GB of slope rows, and the pass sat at
// ~75 % of the L2 throughput cap with HBM at 34 %.  Here a CTA keeps the points of one TILE of lines
// in shared memory (structure of arrays, conflict-free) and streams many rows through it, one warp
// per row segment: slopes are read and rewritten once, the points are read once per CTA, and the
// per-(row, tile) minima / maxima are combined by zreduce_kernel.
constexpr int ZT_BYTES = 64 * 1024;
// lines per CTA tile: whole filter tiles, 64 KB of scaled points
__host__ __device__ constexpr int zt_tile_lines(int D) { return (ZT_BYTES / (int)sizeof(double) / D) & ~(FILTER_TILE - 1); }

__device__ __forceinline__ float warp_min_f32(float v) {
  float r;
  asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float warp_max_f32(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}

// TRACK = false (the KG path, per-tile slope ranges requested): the pass keeps NO running (min, max, index) of
// the row -- that bookkeeping was ~30 of the ~120 instructions per slope of this issue-bound kernel (fp64
// DSETP.MIN/MAX with their NaN fix-ups, four 64-bit selects, two index selects; ncu source view, r02f).  The
// float (min, max) of every 128-line tile is written anyway (one conversion and two FMNMX per slope, one
// CREDUX pair per tile); rounding to float is monotone, so the row's exact minimum lies in a tile whose float
// minimum equals the smallest one of the row, and zreduce_tiles_kernel re-reads just those tiles.
template <int D, bool TRACK>
__global__ void __launch_bounds__(E_THREADS, 3)
zfinish_tiled_kernel(LineBatch lb, CovFinish fin, double* __restrict__ zpv, int* __restrict__ zpi,
                     int rows_per_cta, float2* __restrict__ ztile, int ztiles) {
  extern __shared__ __align__(16) unsigned char e_smem[];
  constexpr int TL = zt_tile_lines(D);
  double* s_xd = reinterpret_cast<double*>(e_smem);  // [D][TL]; lines past the discretisation hold zeros
  const int ntiles = gridDim.x, tile = blockIdx.x;
  const int n_lo = tile * TL;
  const int n_cnt = min(TL, fin.N - n_lo);
  for (int idx = threadIdx.x; idx < TL * D; idx += blockDim.x) {
    const int n = idx / D, k = idx - n * D;
    s_xd[k * TL + n] = n < n_cnt ? fin.xd_s[(size_t)n_lo * D + idx] : 0.0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int r_end = min(lb.C, (int)(blockIdx.y + 1) * rows_per_cta);
  const int kind = fin.kind;
  const double os = fin.outputscale;
  constexpr int U = 4;
  static_assert(32 * U == FILTER_TILE, "one trip of a warp = one filter tile");
  for (int r = blockIdx.y * rows_per_cta + warp; r < r_end; r += nwarp) {
    double xr[D];
#pragma unroll
    for (int k = 0; k < D; ++k) xr[k] = fin.xs[(size_t)r * D + k];
    const double rsd = fin.ystd2 / fin.sd[r];
    double* zw = const_cast<double*>(lb.Z) + (size_t)r * lb.ldz + n_lo;
    double vmin = INFINITY, vmax = -INFINITY;
    int imin = 0x7fffffff, imax = 0x7fffffff;
    double vn[U];  // the next trip's products are requested before this trip's kernel evaluations
#pragma unroll
    for (int u = 0; u < U; ++u) vn[u] = lane + 32 * u < n_cnt ? zw[lane + 32 * u] : 0.0;
    // (warp-uniform trip count: the tile statistics below use warp-wide reductions)
#pragma unroll 2
    for (int b0 = 0; b0 < n_cnt; b0 += 32 * U) {
      const int i0 = b0 + lane;
      const double* xd0 = s_xd + i0;
      double v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = vn[u];
      const int j0 = i0 + 32 * U;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = j0 + 32 * u;
        if (i < n_cnt) vn[u] = zw[i];
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        double sq = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          const double df = xr[k] - xd0[k * TL + 32 * u];  // (zero-filled past n_cnt: no bounds check)
          sq = fma(df, df, sq);
        }
        v[u] = (stationary_from_sq(kind, os, sq) - v[u]) * rsd;
      }
      float flo = INFINITY, fhi = -INFINITY;  // this trip's 128 consecutive lines = one filter tile
      if (b0 + 32 * U <= n_cnt) {  // (warp-uniform) whole trip inside the discretisation
#pragma unroll
        for (int u = 0; u < U; ++u) {
          zw[i0 + 32 * u] = v[u];
          if (TRACK) {
            if (v[u] < vmin) { vmin = v[u]; imin = n_lo + i0 + 32 * u; }
            if (v[u] > vmax) { vmax = v[u]; imax = n_lo + i0 + 32 * u; }
          }
          // the slope as the chord filter will see it (rounded to nearest float; rounding is monotone, so
          // the extremes of the rounded values bracket every rounded slope of the tile)
          const float f = __double2float_rn(v[u]);
          flo = fminf(flo, f);
          fhi = fmaxf(fhi, f);
        }
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = i0 + 32 * u;
          if (i < n_cnt) {
            zw[i] = v[u];
            if (TRACK) {
              if (v[u] < vmin) { vmin = v[u]; imin = n_lo + i; }
              if (v[u] > vmax) { vmax = v[u]; imax = n_lo + i; }
            }
            const float f = __double2float_rn(v[u]);
            flo = fminf(flo, f);
            fhi = fmaxf(fhi, f);
          }
        }
      }
      if (ztile != nullptr) {
        flo = warp_min_f32(flo);
        fhi = warp_max_f32(fhi);
        if (lane == 0) ztile[(size_t)r * ztiles + (n_lo + b0) / FILTER_TILE] = make_float2(flo, fhi);
      }
    }
    if (TRACK) {
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, vmin, o);
        const int oi = __shfl_xor_sync(0xffffffffu, imin, o);
        if (MinOp::better(ov, oi, vmin, imin)) { vmin = ov; imin = oi; }
        const double pv = __shfl_xor_sync(0xffffffffu, vmax, o);
        const int pi = __shfl_xor_sync(0xffffffffu, imax, o);
        if (MaxOp::better(pv, pi, vmax, imax)) { vmax = pv; imax = pi; }
      }
      if (lane == 0) {
        const size_t q = ((size_t)r * ntiles + tile) * 2;
        zpv[q] = vmin; zpv[q + 1] = vmax;
        zpi[q] = imin; zpi[q + 1] = imax;
      }
    }
  }
}

// one thread per row: combine the tile partials (increasing line order, first index wins ties) and
// the entries past the discretisation lines (the candidate's own line, final already)
__global__ void zreduce_kernel(LineBatch lb, int N, int ntiles, const double* __restrict__ zpv,
                               const int* __restrict__ zpi, double* __restrict__ zst, int* __restrict__ zarg) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= lb.C) return;
  double vmin = INFINITY, vmax = -INFINITY;
  int imin = 0x7fffffff, imax = 0x7fffffff;
  for (int t = 0; t < ntiles; ++t) {
    const size_t q = ((size_t)r * ntiles + t) * 2;
    if (MinOp::better(zpv[q], zpi[q], vmin, imin)) { vmin = zpv[q]; imin = zpi[q]; }
    if (MaxOp::better(zpv[q + 1], zpi[q + 1], vmax, imax)) { vmax = zpv[q + 1]; imax = zpi[q + 1]; }
  }
  for (int n = N; n < lb.NL; ++n) {
    const double v = lb.Z[(size_t)r * lb.ldz + n];
    if (v < vmin) { vmin = v; imin = n; }
    if (v > vmax) { vmax = v; imax = n; }
  }
  zst[r * 2 + 0] = vmin; zst[r * 2 + 1] = vmax;
  zarg[r * 2 + 0] = imin; zarg[r * 2 + 1] = imax;
}

// one warp per row: exact (min, max, first indices) of the finished slope row from the per-tile float ranges.
// float(min of the row) is the smallest float tile minimum, and every tile that attains it may hold the exact
// minimum: those tiles (usually one) are re-read in increasing line order, first index winning ties.
__global__ void __launch_bounds__(E_THREADS)
zreduce_tiles_kernel(LineBatch lb, int N, const float2* __restrict__ ztile, int ztiles, double* __restrict__ zst,
                     int* __restrict__ zarg) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (E_THREADS / 32) + (threadIdx.x >> 5);
  if (r >= lb.C) return;
  const float2* zt = ztile + (size_t)r * ztiles;
  float glo = INFINITY, ghi = -INFINITY;
  for (int t = lane; t < ztiles; t += 32) {
    const float2 f = zt[t];
    glo = fminf(glo, f.x);
    ghi = fmaxf(ghi, f.y);
  }
  glo = warp_min_f32(glo);
  ghi = warp_max_f32(ghi);
  const double* z = lb.Z + (size_t)r * lb.ldz;
  double vmin = INFINITY, vmax = -INFINITY;
  int imin = 0x7fffffff, imax = 0x7fffffff;
  for (int t0 = 0; t0 < ztiles; t0 += 32) {
    const int t = t0 + lane;
    const float2 f = t < ztiles ? zt[t] : make_float2(INFINITY, -INFINITY);
    unsigned hit = __ballot_sync(0xffffffffu, f.x == glo || f.y == ghi);
    while (hit) {
      const int tt = t0 + __ffs(hit) - 1;
      hit &= hit - 1u;
      for (int u = 0; u < FILTER_TILE / 32; ++u) {
        const int n = tt * FILTER_TILE + lane + 32 * u;
        if (n < N) {
          const double v = z[n];
          if (v < vmin) { vmin = v; imin = n; }
          if (v > vmax) { vmax = v; imax = n; }
        }
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, vmin, o);
    const int oi = __shfl_xor_sync(0xffffffffu, imin, o);
    if (MinOp::better(ov, oi, vmin, imin)) { vmin = ov; imin = oi; }
    const double pv = __shfl_xor_sync(0xffffffffu, vmax, o);
    const int pi = __shfl_xor_sync(0xffffffffu, imax, o);
    if (MaxOp::better(pv, pi, vmax, imax)) { vmax = pv; imax = pi; }
  }
  if (lane == 0) {
    for (int n = N; n < lb.NL; ++n) {  // the candidate's own line (final already)
      const double v = z[n];
      if (v < vmin) { vmin = v; imin = n; }
      if (v > vmax) { vmax = v; imax = n; }
    }
    zst[r * 2 + 0] = vmin; zst[r * 2 + 1] = vmax;
    zarg[r * 2 + 0] = imin; zarg[r * 2 + 1] = imax;
  }
}

template <int D>
static int launch_zfinish_tiled(const LineBatch& lb, const EmaxScratch& sc, const CovFinish& f, cudaStream_t st) {
  constexpr int tile_lines = zt_tile_lines(D);
  const int ntiles = ceil_div(f.N, tile_lines);
  int gy = ceil_div(3 * 148, ntiles);
  if (gy > ceil_div(lb.C, E_THREADS / 32)) gy = ceil_div(lb.C, E_THREADS / 32);
  const int rows_per_cta = ceil_div(lb.C, gy);
  gy = ceil_div(lb.C, rows_per_cta);
  const size_t smem = (size_t)tile_lines * D * sizeof(double);
  // (DKG_ZSTAT_TRACK=1: the round-2 form with running row statistics in the pass itself, for comparison)
  const bool track = sc.ztile == nullptr || getenv("DKG_ZSTAT_TRACK") != nullptr;
  if (track) {
    DKG_CUDA_OK(cudaFuncSetAttribute(zfinish_tiled_kernel<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    zfinish_tiled_kernel<D, true><<<dim3(ntiles, gy), E_THREADS, smem, st>>>(lb, f, sc.zpv, sc.zpi, rows_per_cta, sc.ztile, sc.ztiles);
    DKG_LAUNCH_CHECK();
    zreduce_kernel<<<ceil_div(lb.C, 128), 128, 0, st>>>(lb, f.N, ntiles, sc.zpv, sc.zpi, sc.zst, sc.zarg);
    DKG_LAUNCH_CHECK();
  } else {
    DKG_CUDA_OK(cudaFuncSetAttribute(zfinish_tiled_kernel<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    zfinish_tiled_kernel<D, false><<<dim3(ntiles, gy), E_THREADS, smem, st>>>(lb, f, sc.zpv, sc.zpi, rows_per_cta, sc.ztile, sc.ztiles);
    DKG_LAUNCH_CHECK();
    zreduce_tiles_kernel<<<ceil_div(lb.C, E_THREADS / 32), E_THREADS, 0, st>>>(lb, f.N, sc.ztile, sc.ztiles, sc.zst, sc.zarg);
    DKG_LAUNCH_CHECK();
  }
  return DKG_OK;
}

int emax_zstat(const LineBatch& lb, const EmaxScratch& sc, double* amax_out, int* aarg_out,
               cudaStream_t st, const CovFinish* fin, bool* ztile_written) {
  if (ztile_written != nullptr) *ztile_written = false;
  if (lb.C == 0) return DKG_OK;
  CovFinish f{};
  if (fin != nullptr) f = *fin;
  if (fin != nullptr && f.d >= 1 && amax_out == nullptr && sc.zpv != nullptr && getenv("DKG_ZSTAT_ROWS") == nullptr &&
      (long long)lb.C * f.N >= (1ll << 22)) {
    int rc = DKG_OK;
    switch (f.d) {
      case 1: rc = launch_zfinish_tiled<1>(lb, sc, f, st); break;
      case 2: rc = launch_zfinish_tiled<2>(lb, sc, f, st); break;
      case 3: rc = launch_zfinish_tiled<3>(lb, sc, f, st); break;
      case 4: rc = launch_zfinish_tiled<4>(lb, sc, f, st); break;
      case 5: rc = launch_zfinish_tiled<5>(lb, sc, f, st); break;
      case 6: rc = launch_zfinish_tiled<6>(lb, sc, f, st); break;
      case 7: rc = launch_zfinish_tiled<7>(lb, sc, f, st); break;
      default: rc = launch_zfinish_tiled<8>(lb, sc, f, st); break;
    }
    if (rc != DKG_OK) return rc;
    if (ztile_written != nullptr) *ztile_written = sc.ztile != nullptr;
    const long long sets = (long long)lb.C * lb.S;
    chain_kernel<<<(unsigned)((sets + E_THREADS - 1) / E_THREADS), E_THREADS, 0, st>>>(lb, sc);
    DKG_LAUNCH_CHECK();
    return DKG_OK;
  }
#define DKG_ZSTAT(DD) zstat_kernel<DD><<<lb.C, E_THREADS, 0, st>>>(lb, sc, sc.zst, sc.zarg, amax_out, aarg_out, f)
  switch (fin != nullptr ? fin->d : 0) {
    case 0: DKG_ZSTAT(0); break;
    case 1: DKG_ZSTAT(1); break;
    case 2: DKG_ZSTAT(2); break;
    case 3: DKG_ZSTAT(3); break;
    case 4: DKG_ZSTAT(4); break;
    case 5: DKG_ZSTAT(5); break;
    case 6: DKG_ZSTAT(6); break;
    case 7: DKG_ZSTAT(7); break;
    default: DKG_ZSTAT(8); break;
  }
#undef DKG_ZSTAT
  DKG_LAUNCH_CHECK();
  const long long sets = (long long)lb.C * lb.S;
  chain_kernel<<<(unsigned)((sets + E_THREADS - 1) / E_THREADS), E_THREADS, 0, st>>>(lb, sc);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

int emax_zstat_from_partials(const LineBatch& lb, const EmaxScratch& sc, int ntiles, cudaStream_t st) {
  if (lb.C == 0) return DKG_OK;
  if (ntiles > 0) {
    zreduce_kernel<<<ceil_div(lb.C, 128), 128, 0, st>>>(lb, lb.NL, ntiles, sc.zpv, sc.zpi, sc.zst, sc.zarg);
    DKG_LAUNCH_CHECK();
  }
  const long long sets = (long long)lb.C * lb.S;
  chain_kernel<<<(unsigned)((sets + E_THREADS - 1) / E_THREADS), E_THREADS, 0, st>>>(lb, sc);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// ------------------------------------------------------------------------------------------
// per-set facts shared by all stages
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ size_t a_base(const LineBatch& lb, int c, int j) {
  const int jt = lb.row_mod ? (c % lb.row_mod) : j;
  return (size_t)c * lb.a_sc + (size_t)jt * lb.a_sj;
}
__device__ __forceinline__ size_t am_index(const LineBatch& lb, int c, int j) {
  const int jt = lb.row_mod ? (c % lb.row_mod) : j;
  return (size_t)c * lb.am_sc + jt;
}
__device__ __forceinline__ double line_intercept(const LineBatch& lb, int c, int j, int n) {
  return n < lb.NA ? lb.A[a_base(lb, c, j) + n] : lb.a_own[(size_t)c * lb.S + j];
}

struct SetInfo {
  double w;       // slope weight of this scalarisation
  double amax;    // max_n a_n  (torch.max(intercepts), discretekg.py:336)
  int iT;         // a line attaining it (the candidate's own line wins ties: reference index 0)
  int own_is_max;
  int iP, iQ;     // first indices of the min / max raw slope coordinate
  double zmin, zmax;
  bool shortcut;  // all |slopes| < 1e-9 (discretekg.py:363)
};

__device__ __forceinline__ SetInfo set_info(const LineBatch& lb, const EmaxScratch& sc, int c, int j) {
  SetInfo s;
  s.w = lb.wt ? lb.wt[j] : 1.0;
  s.zmin = sc.zst[c * 2 + 0];
  s.zmax = sc.zst[c * 2 + 1];
  s.iP = sc.zarg[c * 2 + 0];
  s.iQ = sc.zarg[c * 2 + 1];
  s.amax = lb.Amax[am_index(lb, c, j)];
  s.iT = lb.Aarg[am_index(lb, c, j)];
  s.own_is_max = 0;
  if (lb.a_own != nullptr) {
    const double ao = lb.a_own[(size_t)c * lb.S + j];
    if (ao >= s.amax) { s.amax = ao; s.iT = lb.NA; s.own_is_max = 1; }
  }
  s.shortcut = fabs(__dmul_rn(s.w, s.zmin)) < SHORTCUT_TOL && fabs(__dmul_rn(s.w, s.zmax)) < SHORTCUT_TOL;
  return s;
}

// ------------------------------------------------------------------------------------------
// chord filter (streaming)
// ------------------------------------------------------------------------------------------
// Chain parameters for set (c, j): a line (z, a) survives iff
//   a > min(par.x + par.y * z, par.z + par.w * z)
// (left chord P->T and right chord T->Q in the dual plane; the chain is concave because T has
// the maximum intercept).  A slack of 128 eps of every magnitude entering c + m*z keeps all
// lines within rounding of the chain: extra survivors cost time, never accuracy.
__device__ double4 chain_params(const LineBatch& lb, const EmaxScratch& sc, int c, int j, ChainMag* mag,
                                double4* verts) {
  const double inf = INFINITY;
  const SetInfo s = set_info(lb, sc, c, j);
  if (mag != nullptr) mag->amag = mag->zmag = 0.0;
  if (verts != nullptr) verts[0] = verts[1] = make_double4(0.0, 0.0, 0.0, 0.0);
  if (s.shortcut) return make_double4(inf, 0.0, inf, 0.0);
  const double sgn = s.w < 0.0 ? -1.0 : 1.0;  // effective slope coordinate z' = sgn * z
  const int iP = sgn > 0 ? s.iP : s.iQ;
  const int iQ = sgn > 0 ? s.iQ : s.iP;
  const double zP = sgn > 0 ? s.zmin : -s.zmax;
  const double zQ = sgn > 0 ? s.zmax : -s.zmin;
  const double aT = s.amax;
  const double zT = sgn * line_slope(lb, c, s.iT);
  const double aP = line_intercept(lb, c, j, iP);
  const double aQ = line_intercept(lb, c, j, iQ);
  if (mag != nullptr) {
    mag->amag = fmax(fabs(aT), fmax(fabs(aP), fabs(aQ)));
    mag->zmag = fmax(fabs(s.zmin), fabs(s.zmax));
  }
  if (verts != nullptr) {  // chord end points in the RAW slope coordinate: (z0, a0, z1, a1) per side
    verts[0] = make_double4(sgn * zP, aP, sgn * zT, aT);
    verts[1] = make_double4(sgn * zT, aT, sgn * zQ, aQ);
  }
  double c1 = inf, m1 = 0.0, c2 = inf, m2 = 0.0;
  if (zT > zP) {
    m1 = (aT - aP) / (zT - zP);
    const double slack = EPS128 * (fabs(aT) + fabs(aP) + fabs(m1) * fmax(fabs(zP), fabs(zT)));
    c1 = aP - m1 * zP - slack;
  }
  if (zQ > zT) {
    m2 = (aQ - aT) / (zQ - zT);
    const double slack = EPS128 * (fabs(aT) + fabs(aQ) + fabs(m2) * fmax(fabs(zQ), fabs(zT)));
    c2 = aT - m2 * zT - slack;
  }
  return make_double4(c1, m1 * sgn, c2, m2 * sgn);
}

// Float image (m, m, c, c) of one chord for the fp32 filter, which keeps a line iff
//   float(a) > fmaf(m32, float(z), c32)      (either chord).
// Every line that passes the fp64 test  a > c + m z  must pass this one (extra survivors only cost
// time).  Error budget, all relative to  mag = |c| + |m| zmag + amag  (|z| <= zmag on the row; a
// true survivor has |a| <= amag because the concave chain runs between members of the set):
// a, z, m rounded to nearest float (2^-24 each), c32 rounded DOWN, the float fma (2^-24 |t|):
// together < 4.1 * 2^-24 mag.  c32 is lowered by 2^-21 mag = 8 * 2^-24 mag.  Sets whose magnitudes
// leave the range where those float bounds hold get (0, 0, -inf, -inf): every line survives and
// the exact overflow path takes the set.
__device__ float4 chord32(double c, double m, const ChainMag& mag) {
  if (c == INFINITY) return make_float4(0.f, 0.f, INFINITY, INFINITY);  // no chord / shortcut set
  const double g = fabs(c) + fabs(m) * mag.zmag + mag.amag;
  const bool ok = g > 1e-30 && g < 1e30 && fabs(m) < 1e30 && mag.zmag < 1e30 &&
                  (mag.zmag + fabs(m)) <= g * 1e37;  // float denormal losses stay below the slack
  if (!ok) return make_float4(0.f, 0.f, -INFINITY, -INFINITY);
  const float cf = __double2float_rd(c - 4.76837158203125e-07 * g);  // 2^-21
  const float mf = __double2float_rn(m);
  return make_float4(mf, mf, cf, cf);
}

// Each CTA owns G candidates x (E_THREADS * R) lines; each thread keeps the slope coordinates of
// its R lines for the G candidates in registers (G*R doubles), then walks the scalarisations:
// per scalarisation it loads the chain parameters of the G sets once (shared-memory broadcast)
// and the R intercepts once (coalesced, shared by the G candidates), and runs G*R tests of
// 2 DFMA + DMNMX + DSETP each.  Survivors are rare (~1%) and appended with an atomic slot claim.
__device__ __forceinline__ unsigned long long pack_excess(double e, int n) {
  // e > 0: float bits are order preserving; only used to CHOOSE a seed -- any line above the
  // chord is a valid chain vertex, so the float rounding is harmless
  return ((unsigned long long)__float_as_uint((float)e) << 32) | (unsigned)n;
}



// Survivors are first collected in a CTA-local shared-memory pool (shared atomics are ~30 cycles;
// a global atomicAdd whose return value is needed stalls the warp for a full L2 round trip, and
// with ~5 survivors per warp per scalarisation that latency dominated the kernel: ncu,
// profiles/r01_ncu_summary.md).  At the end the CTA claims one contiguous range per set with a
// single global atomic and copies its entries out in parallel.
constexpr int POOL_CAP = 2048;

// store one survivor; returns its packed excess above its chord (0 if none) and the side
__device__ __forceinline__ unsigned long long append_survivor(const LineBatch& lb, const EmaxScratch& sc,
                                                              const double4& par, size_t set, int c,
                                                              int j, int n, int pos, int* side_out) {
  const double av = lb.A[a_base(lb, c, j) + n];
  const double zv = line_slope(lb, c, n);
  if (pos < SURV_CAP) {
    SurvEntry e;
    e.a = av; e.z = zv; e.idx = n; e.pad = 0;
    sc.surv[set * SURV_CAP + pos] = e;
  }
  // the farthest survivor above each chord seeds the QuickHull refinement in the hull /
  // overflow kernels (any survivor is a valid chain vertex)
  const double t1 = fma(par.y, zv, par.x), t2 = fma(par.w, zv, par.z);
  const int side = t1 <= t2 ? 0 : 1;
  const double ex = av - (side == 0 ? t1 : t2);
  *side_out = side;
  return ex > 0.0 ? pack_excess(ex, n) : 0ull;
}

template <int G, int R, bool SHARED_A>
__global__ void __launch_bounds__(E_THREADS, (G * R > 16) ? 2 : 3)
filter_kernel(LineBatch lb, EmaxScratch sc) {
  extern __shared__ __align__(16) unsigned char e_smem[];
  const int S = lb.S;
  double4* s_par = reinterpret_cast<double4*>(e_smem);     // [S][G]
  int* s_cnt = reinterpret_cast<int*>(s_par + G * S);      // [S*G] survivors pooled per set
  int* s_base = s_cnt + G * S;                             // [S*G] claimed global range start
  int* s_fill = s_base + G * S;                            // [S*G]
  int2* s_pool = reinterpret_cast<int2*>(s_fill + G * S);  // [POOL_CAP] (line, local set)
  unsigned long long* s_far = reinterpret_cast<unsigned long long*>(s_pool + POOL_CAP);  // [S*G][2]
  __shared__ int s_pool_n;
  const int c0 = blockIdx.y * G;
  for (int e = threadIdx.x; e < G * S; e += blockDim.x) {
    const int j = e / G, g = e - j * G;
    const int c = c0 + g;
    s_par[e] = (c < lb.C) ? sc.chain[(size_t)c * S + j] : make_double4(INFINITY, 0.0, INFINITY, 0.0);
    s_cnt[e] = 0;
    s_fill[e] = 0;
    s_far[2 * e] = 0ull;
    s_far[2 * e + 1] = 0ull;
  }
  if (threadIdx.x == 0) s_pool_n = 0;
  // lines of this thread; out-of-range slots alias the last line (a duplicate survivor is
  // harmless to the exact march, and it keeps the loop free of bounds predicates)
  int nc[R];
#pragma unroll
  for (int r = 0; r < R; ++r)
    nc[r] = min(blockIdx.x * (E_THREADS * R) + r * E_THREADS + (int)threadIdx.x, lb.NA - 1);
  double z[G][R];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int c = min(c0 + g, lb.C - 1);
#pragma unroll
    for (int r = 0; r < R; ++r) z[g][r] = line_slope(lb, c, nc[r]);
  }
  double a_nx[R];  // intercepts of the NEXT scalarisation (software prefetch, SHARED_A only)
  if (SHARED_A) {
#pragma unroll
    for (int r = 0; r < R; ++r) a_nx[r] = lb.A[nc[r]];
  }
  __syncthreads();

  for (int j = 0; j < S; ++j) {
    double4 p[G];
#pragma unroll
    for (int g = 0; g < G; ++g) p[g] = s_par[j * G + g];
    double a[R];
    if (SHARED_A) {
      const double* nxt = lb.A + (size_t)min(j + 1, S - 1) * lb.a_sj;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        a[r] = a_nx[r];
        a_nx[r] = nxt[nc[r]];
      }
    }
    // branch-free tests -> bit mask (bit g*R + r).  a > min(l1, l2)  <=>  a > l1 or a > l2.
    unsigned mask = 0u;
#pragma unroll
    for (int g = 0; g < G; ++g) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        double av;
        if (SHARED_A) av = a[r];
        else av = lb.A[a_base(lb, min(c0 + g, lb.C - 1), j) + nc[r]];
        const double t1 = fma(p[g].y, z[g][r], p[g].x);
        const double t2 = fma(p[g].w, z[g][r], p[g].z);
        if ((av > t1) | (av > t2)) mask |= 1u << (g * R + r);
      }
    }
    while (mask) {  // rare: ~1 % of the tests
      const int bit = __ffs(mask) - 1;
      mask &= mask - 1u;
      const int g = bit / R, r = bit - g * R;
      int n = nc[0];
#pragma unroll
      for (int rr = 1; rr < R; ++rr)
        if (rr == r) n = nc[rr];
      const int setl = j * G + g;
      const int pp = atomicAdd(&s_pool_n, 1);
      if (pp < POOL_CAP) {
        s_pool[pp] = make_int2(n, setl);
        atomicAdd(&s_cnt[setl], 1);
      } else {  // pool full (very dense survivors): claim a global slot directly
        const int c = c0 + g;
        const size_t set = (size_t)c * S + j;
        int side;
        const unsigned long long key =
            append_survivor(lb, sc, s_par[setl], set, c, j, n, atomicAdd(&sc.surv_cnt[set], 1), &side);
        if (key) atomicMax(&s_far[2 * setl + side], key);
      }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < G * S; e += blockDim.x) {
    const int k = s_cnt[e];
    if (k > 0) {
      const int j = e / G, g = e - j * G;
      s_base[e] = atomicAdd(&sc.surv_cnt[(size_t)(c0 + g) * S + j], k);
    }
  }
  __syncthreads();
  const int np = min(s_pool_n, POOL_CAP);
  for (int e = threadIdx.x; e < np; e += blockDim.x) {
    const int2 it = s_pool[e];
    const int setl = it.y;
    const int j = setl / G, g = setl - j * G;
    const int c = c0 + g;
    const int pos = s_base[setl] + atomicAdd(&s_fill[setl], 1);
    int side;
    const unsigned long long key =
        append_survivor(lb, sc, s_par[setl], (size_t)c * S + j, c, j, it.x, pos, &side);
    if (key) atomicMax(&s_far[2 * setl + side], key);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * G * S; e += blockDim.x) {
    const unsigned long long key = s_far[e];
    if (key) {
      const int setl = e >> 1;
      const int j = setl / G, g = setl - j * G;
      atomicMax(&sc.far[((size_t)(c0 + g) * S + j) * 2 + (e & 1)], key);
    }
  }
}

// ---- fp32 variant (KG path: one intercept table shared by all candidates) --------------------
// The same test in float arithmetic against the conservatively rounded chain (chord32): 64-bit
// compares and FMAs run at half rate on the fp64 pipe, the float test issues as one packed FFMA2
// per chord for two lines plus two FSETP and one predicated add per line.  Survivors are stored
// from the fp64 tables as before, so the hull stage sees exact values; false positives (lines
// within ~5e-7 of the chain) only lengthen the lists.  Each thread owns 4 CONSECUTIVE lines
// (one 16-byte load of the float intercepts per scalarisation, two of the slope row per candidate).
typedef unsigned long long u64;

template <unsigned BIT>
__device__ __forceinline__ unsigned pair_test32(unsigned mask, u64 mm1, u64 cc1, u64 mm2, u64 cc2, u64 zz,
                                                float a0, float a1) {
  asm("{\n"
      ".reg .b64 t1, t2;\n .reg .f32 t1l, t1h, t2l, t2h;\n .reg .pred p, q;\n"
      "fma.rn.f32x2 t1, %1, %5, %2;\n"
      "fma.rn.f32x2 t2, %3, %5, %4;\n"
      "mov.b64 {t1l, t1h}, t1;\n mov.b64 {t2l, t2h}, t2;\n"
      "setp.gt.f32 p, %6, t1l;\n setp.gt.or.f32 p, %6, t2l, p;\n @p add.u32 %0, %0, %8;\n"
      "setp.gt.f32 q, %7, t1h;\n setp.gt.or.f32 q, %7, t2h, q;\n @q add.u32 %0, %0, %9;\n"
      "}"
      : "+r"(mask)
      : "l"(mm1), "l"(cc1), "l"(mm2), "l"(cc2), "l"(zz), "f"(a0), "f"(a1), "n"(BIT), "n"(BIT << 1));
  return mask;
}

__device__ __forceinline__ u64 pack_f32x2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}

// line through two members of a set as (c, m): t(z) = c + m z, lowered by the same 128-ulp slack as
// the 3-point chain (the end points themselves pass the test  a > t)
__device__ __forceinline__ double2 chord_through(double z0, double a0, double z1, double a1) {
  const double m = (a1 - a0) / (z1 - z0);
  const double slack = EPS128 * (fabs(a0) + fabs(a1) + fabs(m) * fmax(fabs(z0), fabs(z1)));
  return make_double2(a0 - m * z0 - slack, m);
}

// Survivors are parked in a per-WARP shared-memory pool (position = ballot prefix, no atomics, no
// CTA barrier) and written out by the warp itself: one global slot claim per survivor, all lanes in
// flight at once, so the L2 round trip is paid per 32 survivors, not per survivor.
constexpr int F32_THREADS = 128;
constexpr int WPOOL = 256;
// scalarisations per CTA (blockIdx.z walks the batches): all of them up to 64 (20 KB of chain
// parameters, 7 CTAs per SM), batches of 32 beyond
static inline int f32_jb(int S) { return S <= 64 ? S : 32; }

// The filter runs in two launches.  Phase 1 covers a SAMPLE of the lines (a few line blocks spread
// over the row) and stores its survivors as they are; the per-set, per-side farthest sampled line
// above the 3-point chord (sc.far) then becomes a vertex U / V of a SECOND-LEVEL chain
// (chain5_kernel: P -> U -> T and T -> V -> Q; any member of the set is a valid vertex, the sample
// only decides how tight the chain is).  Phase 2 covers the remaining lines and re-tests every
// parked line, exactly (fp64), against that chain before it claims a slot: on smooth GP posteriors
// this removes ~70 % of the stored survivors (and the hull-stage work that follows them).
template <int G, bool REFINE>
__device__ __forceinline__ void flush_warp_pool(const LineBatch& lb, const EmaxScratch& sc, const int2* pool,
                                                int cnt, int c0, int j_lo, unsigned long long* s_far) {
  __syncwarp();
  const int S = lb.S;
  for (int e = threadIdx.x & 31; e < cnt; e += 32) {
    const int2 it = pool[e];
    const int setl = it.y;
    const int j = setl / G, g = setl - j * G;
    const int c = c0 + g;
    const size_t set = (size_t)c * S + j;
    const int n = it.x;
    const double av = lb.A[a_base(lb, c, j) + n];
    const double zv = line_slope(lb, c, n);
    const double4 par = sc.chain[set];
    const double t1 = fma(par.y, zv, par.x), t2 = fma(par.w, zv, par.z);
    const int side = t1 <= t2 ? 0 : 1;
    if (REFINE) {
      const double4 q = sc.chain5[set * 2 + side];
      if (!((av > fma(q.y, zv, q.x)) | (av > fma(q.w, zv, q.z)))) continue;
    }
    const int pos = atomicAdd(&sc.surv_cnt[set], 1);
    if (pos < SURV_CAP) {
      SurvEntry en;
      en.a = av; en.z = zv; en.idx = n; en.pad = 0;
      sc.surv[set * SURV_CAP + pos] = en;
    }
    // the farthest survivor above each 3-point chord: second-level chain vertex (phase 1) and seed
    // of the QuickHull refinement in the hull stage
    const double ex = av - (side == 0 ? t1 : t2);
    if (ex > 0.0) atomicMax(&s_far[2 * (setl - j_lo * G) + side], pack_excess(ex, n));
  }
  __syncwarp();
}

// second-level chain of every set from the phase-1 maxima: chain5[set] = (c_PU, m_PU, c_UT, m_UT),
// (c_TV, m_TV, c_VQ, m_VQ); a missing vertex repeats the 3-point chord
__global__ void __launch_bounds__(E_THREADS)
chain5_kernel(LineBatch lb, EmaxScratch sc) {
  const size_t set = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (set >= (size_t)lb.C * lb.S) return;
  const int c = (int)(set / lb.S), j = (int)(set - (size_t)c * lb.S);
  const double4 par = sc.chain[set];
  float mf[4], cf[4];
  ChainMag mag;
  {  // magnitudes for the conservative float rounding (chord32): as in chain_params
    const double4 v0 = sc.chainv[set * 2 + 0], v1 = sc.chainv[set * 2 + 1];
    mag.amag = fmax(fabs(v0.w), fmax(fabs(v0.y), fabs(v1.w)));
    mag.zmag = fmax(fabs(sc.zst[c * 2 + 0]), fabs(sc.zst[c * 2 + 1]));
  }
#pragma unroll
  for (int side = 0; side < 2; ++side) {
    const double pc = side == 0 ? par.x : par.z, pm = side == 0 ? par.y : par.w;
    double4 q = make_double4(pc, pm, pc, pm);
    const unsigned long long key = sc.far[set * 2 + side];
    if (key != 0ull && pc != INFINITY) {
      const int nu = (int)(key & 0xffffffffull);
      const double au = lb.A[a_base(lb, c, j) + nu];
      const double zu = line_slope(lb, c, nu);
      const double4 v = sc.chainv[set * 2 + side];  // (z0, a0, z1, a1), raw slope coordinate
      if ((zu - v.x) * (v.z - zu) > 0.0) {          // U strictly between the chord's end points
        const double2 c0u = chord_through(v.x, v.y, zu, au), cu1 = chord_through(zu, au, v.z, v.w);
        q = make_double4(c0u.x, c0u.y, cu1.x, cu1.y);
      }
    }
    sc.chain5[set * 2 + side] = q;
    const float4 f0 = chord32(q.x, q.y, mag), f1 = chord32(q.z, q.w, mag);
    mf[2 * side] = f0.x; cf[2 * side] = f0.z;
    mf[2 * side + 1] = f1.x; cf[2 * side + 1] = f1.z;
  }
  if (sc.chain5f != nullptr) {
    sc.chain5f[set * 2 + 0] = make_float4(mf[0], mf[1], mf[2], mf[3]);
    sc.chain5f[set * 2 + 1] = make_float4(cf[0], cf[1], cf[2], cf[3]);
  }
}

// ---- tile-first fp32 filter (KG path, N >= 1024) ------------------------------------------------
// The plan orders the discretisation along a Morton curve, so FILTER_TILE consecutive lines are
// neighbours in input space: their slopes span a short interval [zlo, zhi] (written per (row, tile)
// by the row-statistics pass) and their intercepts are bounded by the tile maximum of the shared
// table (plan time).  The per-line test keeps a line iff  a > fma(m_k, z, c_k)  for one of the chords
// k of the (second-level) chain; fma is monotone in z, so
//     amax_tile <= min(fma(m_k, zlo, c_k), fma(m_k, zhi, c_k))   for EVERY chord k
// proves that no line of the tile passes: the (tile, set) pair is dropped with one test and the
// survivor lists are exactly those of the per-line filter.  At c4 81 % / 92 % of the (tile, set) pairs
// and 64 % / 85 % of the (tile, row) pairs -- whose slopes are then not even read -- go this way.
//
// 1. probe32_kernel: 16 consecutive lines out of every 256 (a spatially spread sample in Morton order) against
//    the 3-point chain; records the farthest line above each chord (sc.far) -> chain5_kernel builds the
//    second-level chain P-U-T-V-Q (fp64) and its conservatively rounded float image.
// 2. tilefilter_kernel: one warp per (row, 16 scalarisations, range of tile pairs); lane (half, j)
//    decides tile 2 tp + half for scalarisation j; surviving (tile, set) pairs are tested line by line
//    (4 consecutive lines per lane) against the 4 float chords, parked in a per-warp pool and written out
//    after the exact fp64 re-test against the same chain.
// the sample: PROBE_CHUNK consecutive lines (one 128-byte segment of the slope row) out of every
// PROBE_PERIOD -- a spatially spread sample in Morton order that costs 1/16 of a pass over the rows
// (single lines at a stride of 16 pull every 128-byte segment of the rows through DRAM: measured)
constexpr int PROBE_CHUNK = 16;
constexpr int PROBE_PERIOD = 256;
constexpr int PROBE_THREADS = 128;
constexpr int PROBE_G = 4;

__global__ void __launch_bounds__(PROBE_THREADS)
probe32_kernel(LineBatch lb, EmaxScratch sc, int JB) {
  extern __shared__ __align__(16) unsigned char e_smem[];
  const int S = lb.S;
  const int j_lo = blockIdx.z * JB, j_hi = min(S, j_lo + JB), SB = j_hi - j_lo;
  float4* s_p32 = reinterpret_cast<float4*>(e_smem);  // [SB][G][2] (m, m, c, c) per chord
  unsigned long long* s_far = reinterpret_cast<unsigned long long*>(s_p32 + 2 * PROBE_G * JB);  // [SB * G][2]
  const int c0 = blockIdx.y * PROBE_G;
  for (int e = threadIdx.x; e < PROBE_G * SB; e += blockDim.x) {
    const int j = j_lo + e / PROBE_G, g = e % PROBE_G;
    const int c = c0 + g;
    const bool in = c < lb.C;
    const float4 none = make_float4(0.f, 0.f, INFINITY, INFINITY);
    s_p32[2 * e] = in ? sc.chain32[((size_t)c * S + j) * 2] : none;
    s_p32[2 * e + 1] = in ? sc.chain32[((size_t)c * S + j) * 2 + 1] : none;
    s_far[2 * e] = 0ull;
    s_far[2 * e + 1] = 0ull;
  }
  __syncthreads();
  const int sample = blockIdx.x * PROBE_THREADS + (int)threadIdx.x;
  const int n = (sample / PROBE_CHUNK) * PROBE_PERIOD + sample % PROBE_CHUNK;
  if (n < lb.NA) {
    float z[PROBE_G];
#pragma unroll
    for (int g = 0; g < PROBE_G; ++g) z[g] = __double2float_rn(lb.Z[(size_t)min(c0 + g, lb.C - 1) * lb.ldz + n]);
    float a_nx = lb.A32[(size_t)j_lo * lb.a_sj + n];
    for (int j = j_lo; j < j_hi; ++j) {
      const float a = a_nx;
      if (j + 1 < j_hi) a_nx = lb.A32[(size_t)(j + 1) * lb.a_sj + n];
#pragma unroll
      for (int g = 0; g < PROBE_G; ++g) {
        const float4 q1 = s_p32[2 * ((j - j_lo) * PROBE_G + g)], q2 = s_p32[2 * ((j - j_lo) * PROBE_G + g) + 1];
        if ((a > fmaf(q1.x, z[g], q1.z)) | (a > fmaf(q2.x, z[g], q2.z))) {  // rare (~1 %)
          const int c = c0 + g;
          if (c < lb.C) {
            const size_t set = (size_t)c * S + j;
            const double av = lb.A[a_base(lb, c, j) + n];
            const double zv = lb.Z[(size_t)c * lb.ldz + n];
            const double4 par = sc.chain[set];
            const double t1 = fma(par.y, zv, par.x), t2 = fma(par.w, zv, par.z);
            const int side = t1 <= t2 ? 0 : 1;
            const double ex = av - (side == 0 ? t1 : t2);
            if (ex > 0.0) atomicMax(&s_far[2 * ((j - j_lo) * PROBE_G + g) + side], pack_excess(ex, n));
          }
        }
      }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * PROBE_G * SB; e += blockDim.x) {
    const unsigned long long key = s_far[e];
    if (key) {
      const int setl = e >> 1;
      const int j = j_lo + setl / PROBE_G, g = setl % PROBE_G;
      atomicMax(&sc.far[((size_t)(c0 + g) * S + j) * 2 + (e & 1)], key);
    }
  }
}

// Champion probe (default when the plan carries the tile champions).  The sample above covers 1/16 of the lines
// wherever they happen to lie; the vertices U / V that make the second-level chain tight are HIGH lines.  The
// plan knows, per (128-line tile, scalarisation), the line with the largest intercept of the tile (the intercept
// table is candidate-independent), and within a Morton tile the slopes vary little, so that line is (nearly) the
// tile's farthest line above any chord.  Probing just those N / 128 champions per set -- one gathered slope each,
// the 16 champions of a tile lie within 1 KB of the slope row -- finds better vertices with an eighth of the
// tests: survivors per set 40 -> 20 / 8.5 -> 5.6 on the two c4 objectives (CPU simulation before the kernel was
// written; measured: see DESIGN.md), which is what the hull stage pays for.
constexpr int PC_THREADS = 256;

__global__ void __launch_bounds__(PC_THREADS)
probe_champ_kernel(LineBatch lb, EmaxScratch sc) {
  extern __shared__ __align__(16) unsigned char e_smem[];
  unsigned long long* s_far = reinterpret_cast<unsigned long long*>(e_smem);  // [S][2]
  const int S = lb.S, c = blockIdx.x;
  for (int e = threadIdx.x; e < 2 * S; e += blockDim.x) s_far[e] = 0ull;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int half = lane >> 4, jl = lane & 15;
  const int ntiles = lb.a32_tiles;
  const double* zrow = lb.Z + (size_t)c * lb.ldz;
  for (int jb = 0; jb < S; jb += 16) {
    const int j = jb + jl;
    const bool j_ok = j < S;
    float4 q1 = make_float4(0.f, 0.f, INFINITY, INFINITY), q2 = q1;  // (m, m, c, c) of the chords P-T, T-Q
    if (j_ok) {
      q1 = sc.chain32[((size_t)c * S + j) * 2];
      q2 = sc.chain32[((size_t)c * S + j) * 2 + 1];
    }
    for (int t0 = 2 * warp; t0 < ntiles; t0 += 2 * nwarp) {
      const int tile = t0 + half;
      if (!j_ok || tile >= ntiles) continue;
      const int n = lb.A32targ[(size_t)tile * S + j];
      if (n < 0 || n >= lb.NA) continue;
      const float a = lb.A32tmax[(size_t)tile * S + j];
      const double zv = zrow[n];
      const float z = __double2float_rn(zv);
      if ((a > fmaf(q1.x, z, q1.z)) | (a > fmaf(q2.x, z, q2.z))) {
        const double av = lb.A[a_base(lb, c, j) + n];
        const double4 par = sc.chain[(size_t)c * S + j];
        const double t1 = fma(par.y, zv, par.x), t2 = fma(par.w, zv, par.z);
        const int side = t1 <= t2 ? 0 : 1;
        const double ex = av - (side == 0 ? t1 : t2);
        if (ex > 0.0) atomicMax(&s_far[2 * j + side], pack_excess(ex, n));
      }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * S; e += blockDim.x) {
    const unsigned long long key = s_far[e];
    if (key) atomicMax(&sc.far[(size_t)c * S * 2 + e], key);
  }
}

constexpr int TF_THREADS = 256;
constexpr int TF_POOL = 128;  // parked (line, scalarisation) entries per warp

// write out this warp's parked lines of row c after the exact (fp64) re-test against the second-level chain
__device__ __forceinline__ void tf_flush(const LineBatch& lb, const EmaxScratch& sc, const int2* pool, int cnt, int c) {
  __syncwarp();
  const int S = lb.S;
  for (int e = threadIdx.x & 31; e < cnt; e += 32) {
    const int2 it = pool[e];
    const int n = it.x, j = it.y;
    const size_t set = (size_t)c * S + j;
    const double av = lb.A[a_base(lb, c, j) + n];
    const double zv = lb.Z[(size_t)c * lb.ldz + n];
    const double4 par = sc.chain[set];
    const double t1 = fma(par.y, zv, par.x), t2 = fma(par.w, zv, par.z);
    const int side = t1 <= t2 ? 0 : 1;
    const double4 q = sc.chain5[set * 2 + side];
    if (!((av > fma(q.y, zv, q.x)) | (av > fma(q.w, zv, q.z)))) continue;
    const int pos = atomicAdd(&sc.surv_cnt[set], 1);
    if (pos < SURV_CAP) {
      SurvEntry en;
      en.a = av; en.z = zv; en.idx = n; en.pad = 0;
      sc.surv[set * SURV_CAP + pos] = en;
    }
  }
  __syncwarp();
}

#ifndef DKG_TF_CTAS
#define DKG_TF_CTAS 4
#endif
__global__ void __launch_bounds__(TF_THREADS, DKG_TF_CTAS)
tilefilter_kernel(LineBatch lb, EmaxScratch sc, int pairs_per_warp, int njb, int use_ztile) {
  __shared__ int2 s_pool[TF_THREADS / 32][TF_POOL];
  __shared__ ulonglong2 s_ch[TF_THREADS / 32][16][4];  // per warp, per scalarisation: (m_k, m_k | c_k, c_k) of chord k
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int S = lb.S;
  const int ntiles = (lb.NA + FILTER_TILE - 1) / FILTER_TILE, ntp = (ntiles + 1) / 2;
  const int chunks = (ntp + pairs_per_warp - 1) / pairs_per_warp;
  // work item = (row, scalarisation batch, chunk of tile pairs); chunk fastest so that neighbouring
  // warps share a row's chain parameters and slope row in L1 / L2
  // (the launcher keeps the item count below 2^31: 32-bit index arithmetic)
  const unsigned item = blockIdx.x * (unsigned)(TF_THREADS / 32) + (unsigned)warp;
  if (item >= (unsigned)lb.C * (unsigned)njb * (unsigned)chunks) return;
  const int chunk = (int)(item % (unsigned)chunks);
  const unsigned item_cj = item / (unsigned)chunks;
  const int jb = (int)(item_cj % (unsigned)njb);
  const int c = (int)(item_cj / (unsigned)njb);
  const int half = lane >> 4, jl = lane & 15;
  const int j_mine = jb * 16 + jl;
  const bool j_ok = j_mine < S;
  // chain of (c, j_mine): slopes m[4], intercepts c[4] of the chords P-U, U-T, T-V, V-Q (float, conservative)
  float4 mm = make_float4(0.f, 0.f, 0.f, 0.f), cc = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
  if (j_ok) {
    mm = sc.chain5f[((size_t)c * S + j_mine) * 2 + 0];
    cc = sc.chain5f[((size_t)c * S + j_mine) * 2 + 1];
  }
  const double* zrow = lb.Z + (size_t)c * lb.ldz;
  int2* pool = s_pool[warp];
  int wcnt = 0;
  const unsigned lt = (1u << lane) - 1u;
  const int tp_lo = chunk * pairs_per_warp, tp_hi = min(ntp, tp_lo + pairs_per_warp);
  // tile statistics of the NEXT pair are requested before this pair is worked on (L2 latency)
  float2 zr_nx = make_float2(-INFINITY, INFINITY);  // (no tile statistics: nothing can be culled)
  float am_nx = INFINITY;
  {
    const int tile = 2 * tp_lo + half;
    if (j_ok && tile < ntiles) {
      if (use_ztile) zr_nx = sc.ztile[(size_t)c * sc.ztiles + tile];
      am_nx = lb.A32tmax[(size_t)tile * S + j_mine];
    }
  }
  if (half == 0) {  // packed copies for the per-line tests (one broadcast LDS.128 per chord)
    s_ch[warp][jl][0] = make_ulonglong2(pack_f32x2(mm.x, mm.x), pack_f32x2(cc.x, cc.x));
    s_ch[warp][jl][1] = make_ulonglong2(pack_f32x2(mm.y, mm.y), pack_f32x2(cc.y, cc.y));
    s_ch[warp][jl][2] = make_ulonglong2(pack_f32x2(mm.z, mm.z), pack_f32x2(cc.z, cc.z));
    s_ch[warp][jl][3] = make_ulonglong2(pack_f32x2(mm.w, mm.w), pack_f32x2(cc.w, cc.w));
  }
  __syncwarp();
  for (int tp = tp_lo; tp < tp_hi; ++tp) {
    const int tile = 2 * tp + half;
    const float2 zr = zr_nx;
    const float am = am_nx;
    {
      const int tn = tile + 2;
      if (tp + 1 < tp_hi && j_ok && tn < ntiles) {
        if (use_ztile) zr_nx = sc.ztile[(size_t)c * sc.ztiles + tn];
        am_nx = lb.A32tmax[(size_t)tn * S + j_mine];
      }
    }
    // side_live bit 0: a line of the tile may pass one of the LEFT chords (P-U, U-T); bit 1: one of the right
    // chords (T-V, V-Q).  A chord that culls the whole tile cannot pass any single line of it (fma is monotone
    // in z), so the per-line phase below skips the sides that are dead for the tile: same survivors.
    unsigned side_live = 0u;
    if (j_ok && tile < ntiles) {
      const float lo = zr.x, hi = zr.y;
      const bool dead_l = (am <= fminf(fmaf(mm.x, lo, cc.x), fmaf(mm.x, hi, cc.x))) & (am <= fminf(fmaf(mm.y, lo, cc.y), fmaf(mm.y, hi, cc.y)));
      const bool dead_r = (am <= fminf(fmaf(mm.z, lo, cc.z), fmaf(mm.z, hi, cc.z))) & (am <= fminf(fmaf(mm.w, lo, cc.w), fmaf(mm.w, hi, cc.w)));
      side_live = (dead_l ? 0u : 1u) | (dead_r ? 0u : 2u);
    }
    const unsigned live = __ballot_sync(0xffffffffu, side_live != 0u);  // bit (half * 16 + jl): this (tile, set) pair needs the lines
    if (live == 0u) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      unsigned bits = (live >> (16 * h)) & 0xffffu;
      if (bits == 0u) continue;
      const int t = 2 * tp + h;
      const int n0 = t * FILTER_TILE + lane * 4;  // 4 consecutive lines per lane
      u64 z01 = 0ull, z23 = 0ull;
      const bool in = n0 < lb.NA;  // lines past NA carry -inf in the float intercept table: never kept
      if (in) {
        const double2 v0 = *reinterpret_cast<const double2*>(zrow + n0);
        const double2 v1 = *reinterpret_cast<const double2*>(zrow + n0 + 2);
        z01 = pack_f32x2(__double2float_rn(v0.x), __double2float_rn(v0.y));
        z23 = pack_f32x2(__double2float_rn(v1.x), __double2float_rn(v1.y));
      }
      // (lanes past the last line read line 0 of the tile and are masked out below: no predicate on the loads)
      const float* arow = lb.A32 + (size_t)jb * 16 * lb.a_sj + (in ? n0 : t * FILTER_TILE);
      const unsigned inmask = in ? 0xfu : 0u;
      const ulonglong2* chw = &s_ch[warp][0][0];
      const size_t a_sj = (size_t)lb.a_sj;
      float4 a_nx = *reinterpret_cast<const float4*>(arow + (size_t)(__ffs(bits) - 1) * a_sj);
      while (bits) {
        const int jj = __ffs(bits) - 1;
        bits &= bits - 1u;
        const int j = jb * 16 + jj;
        const float4 a = a_nx;
        if (bits) a_nx = *reinterpret_cast<const float4*>(arow + (size_t)(__ffs(bits) - 1) * a_sj);
        const unsigned sides = __shfl_sync(0xffffffffu, side_live, 16 * h + jj);  // (warp-uniform)
        unsigned mask = 0u;
        // a line is kept iff a > fma(m_k, z, c_k) for one of the four chords: two pairs of chords per
        // pair of lines, each as in filter32_kernel (one packed FFMA2 per chord)
        if (sides & 1u) {
          const ulonglong2 q0 = chw[jj * 4 + 0], q1 = chw[jj * 4 + 1];
          mask = pair_test32<1u>(mask, q0.x, q0.y, q1.x, q1.y, z01, a.x, a.y);
          mask = pair_test32<4u>(mask, q0.x, q0.y, q1.x, q1.y, z23, a.z, a.w);
        }
        if (sides & 2u) {
          const ulonglong2 q2 = chw[jj * 4 + 2], q3 = chw[jj * 4 + 3];
          unsigned mask2 = 0u;
          mask2 = pair_test32<1u>(mask2, q2.x, q2.y, q3.x, q3.y, z01, a.x, a.y);
          mask2 = pair_test32<4u>(mask2, q2.x, q2.y, q3.x, q3.y, z23, a.z, a.w);
          mask |= mask2;
        }
        mask &= inmask;
        for (;;) {  // survivors are rare
          const unsigned vote = __ballot_sync(0xffffffffu, mask != 0u);
          if (vote == 0u) break;
          if (wcnt + 32 > TF_POOL) {
            tf_flush(lb, sc, pool, wcnt, c);
            wcnt = 0;
          }
          if (mask) {
            const int bit = __ffs(mask) - 1;
            mask &= mask - 1u;
            pool[wcnt + __popc(vote & lt)] = make_int2(n0 + bit, j);
          }
          wcnt += __popc(vote);
        }
      }
    }
  }
  tf_flush(lb, sc, pool, wcnt, c);
}

static int launch_tilefilter(const LineBatch& lb, const EmaxScratch& sc, cudaStream_t st, bool ztile_valid) {
  const long long sets = (long long)lb.C * lb.S;
  if (lb.A32targ != nullptr && getenv("DKG_PROBE_SAMPLE") == nullptr) {
    // probe: the tile champions against the 3-point chain -> farthest champion above each chord
    probe_champ_kernel<<<lb.C, PC_THREADS, (size_t)2 * lb.S * sizeof(unsigned long long), st>>>(lb, sc);
    DKG_LAUNCH_CHECK();
  } else {  // probe: 16 lines out of every 256 against the 3-point chain -> farthest line above each chord
    const int nsamp = ceil_div(lb.NA, PROBE_PERIOD) * PROBE_CHUNK;
    const int JB = lb.S <= 64 ? lb.S : 32;
    const size_t smem = (size_t)PROBE_G * JB * (2 * sizeof(float4) + 2 * sizeof(unsigned long long));
    dim3 grid(ceil_div(nsamp, PROBE_THREADS), ceil_div(lb.C, PROBE_G), ceil_div(lb.S, JB));
    probe32_kernel<<<grid, PROBE_THREADS, smem, st>>>(lb, sc, JB);
    DKG_LAUNCH_CHECK();
  }
  chain5_kernel<<<(unsigned)((sets + E_THREADS - 1) / E_THREADS), E_THREADS, 0, st>>>(lb, sc);
  DKG_LAUNCH_CHECK();
  const int ntiles = ceil_div(lb.NA, FILTER_TILE), ntp = (ntiles + 1) / 2;
  const int njb = ceil_div(lb.S, 16);
  // enough warps to fill the machine several times over, but as many tile pairs per warp as that allows
  // (the chain parameters are loaded once per warp)
  // (4 pairs per warp measured best at c4: 0.62 ms vs 0.72 at 16 and 1.08 at 64 -- short items balance the
  // very uneven work per row and keep more loads in flight)
  long long ppw = (long long)lb.C * njb * ntp / (148ll * 64);
  if (ppw > 3) ppw = 3;  // (re-measured with 4 CTAs per SM and the champion probe: 0.452 / 0.288 ms at 3, 0.458 / 0.301 at 2, 0.478 / 0.305 at 4)
  if (const char* e = getenv("DKG_TF_PPW")) ppw = atoi(e);
  if (ppw < 1) ppw = 1;
  if (ppw > ntp) ppw = ntp;
  while ((long long)lb.C * njb * ceil_div(ntp, (int)ppw) >= (1ll << 31) && ppw < ntp) ppw *= 2;  // 32-bit item index in the kernel
  const int chunks = ceil_div(ntp, (int)ppw);
  const long long items = (long long)lb.C * njb * chunks;
  const int wpb = TF_THREADS / 32;
  tilefilter_kernel<<<(unsigned)((items + wpb - 1) / wpb), TF_THREADS, 0, st>>>(lb, sc, (int)ppw, njb, ztile_valid ? 1 : 0);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// blk_first / blk_step / blk_skip map blockIdx.x to a line block: phase 1 visits blocks
// i * blk_step (blk_skip == 0), phase 2 the others (blk_skip == number of sampled blocks)
__device__ __forceinline__ int line_block(int i, int step, int nsb, bool phase2) {
  if (!phase2) return i * step;
  const int body = (step - 1) * nsb;  // non-sampled blocks inside the sampled stretch
  if (i < body) return (i / (step - 1)) * step + 1 + i % (step - 1);
  return nsb * step + (i - body);
}

// ROWS: the coupled path -- every row is its own set (S == 1) and row r reads the intercept row
// r % row_mod of the shared table, so the G rows of a CTA carry G different intercept rows.
template <int G, bool REFINE, bool ROWS>
__global__ void __launch_bounds__(F32_THREADS, 7)
filter32_kernel(LineBatch lb, EmaxScratch sc, int blk_step, int nsb, int F32_JB) {
  static_assert(G == 4, "bit layout below assumes 4 candidates x 4 lines");
  extern __shared__ __align__(16) unsigned char e_smem[];
  const int S = lb.S;
  // blockIdx.z selects a batch of F32_JB scalarisations: shared memory (and with it the number of
  // resident CTAs) does not grow with S; the slope rows are re-read once per batch (mostly from L2)
  const int j_lo = blockIdx.z * F32_JB, j_hi = min(S, j_lo + F32_JB), SB = j_hi - j_lo;
  ulonglong2* s_p32 = reinterpret_cast<ulonglong2*>(e_smem);  // [SB][G][2] (m,m | c,c) per chord
  unsigned long long* s_far = reinterpret_cast<unsigned long long*>(s_p32 + 2 * G * F32_JB);  // [SB*G][2]
  int2* pool = reinterpret_cast<int2*>(s_far + 2 * G * F32_JB) + (threadIdx.x >> 5) * WPOOL;  // this warp's
  const int c0 = blockIdx.y * G;
  const ulonglong2* g_p32 = reinterpret_cast<const ulonglong2*>(sc.chain32);
  const ulonglong2 none = make_ulonglong2(0ull, 0x7f8000007f800000ull);  // (0, 0, +inf, +inf)
  for (int e = threadIdx.x; e < G * SB; e += blockDim.x) {
    const int j = j_lo + e / G, g = e % G;
    const int c = c0 + g;
    const bool in = c < lb.C;
    s_p32[2 * e] = in ? g_p32[((size_t)c * S + j) * 2] : none;
    s_p32[2 * e + 1] = in ? g_p32[((size_t)c * S + j) * 2 + 1] : none;
    s_far[2 * e] = 0ull;
    s_far[2 * e + 1] = 0ull;
  }
  const int n0 = (line_block(blockIdx.x, blk_step, nsb, REFINE) * F32_THREADS + (int)threadIdx.x) * 4;
  const bool live = n0 < lb.NA;  // lines n0..n0+3; beyond NA the float table holds -inf (never kept)
  u64 zp[G][2];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    double2 v0 = make_double2(0.0, 0.0), v1 = v0;
    if (live) {
      const int row = min(c0 + g, lb.C - 1);
      if (ROWS && lb.cov_M > 0) {
        // coupled rows are not materialised: 4 slopes from the M covariance rows of this row's candidate
        const int cc = row / lb.row_mod, jj = row - cc * lb.row_mod;
        double s4[4] = {0.0, 0.0, 0.0, 0.0};
        for (int m = 0; m < lb.cov_M; ++m) {
          const double w2 = lb.cov_w2[jj * lb.cov_M + m];
          const double* cr = lb.cov[m] + (size_t)cc * lb.ldz + n0;
          const double2 c0v = *reinterpret_cast<const double2*>(cr), c1v = *reinterpret_cast<const double2*>(cr + 2);
          s4[0] = fma(w2, c0v.x, s4[0]); s4[1] = fma(w2, c0v.y, s4[1]);
          s4[2] = fma(w2, c1v.x, s4[2]); s4[3] = fma(w2, c1v.y, s4[3]);
        }
        const double sd = lb.cov_sd[row], rinv = 1.0 / sd;
        v0 = make_double2(coupled_quotient(s4[0], sd, rinv), coupled_quotient(s4[1], sd, rinv));
        v1 = make_double2(coupled_quotient(s4[2], sd, rinv), coupled_quotient(s4[3], sd, rinv));
      } else {
        const double* zr = lb.Z + (size_t)row * lb.ldz + n0;
        v0 = *reinterpret_cast<const double2*>(zr);
        v1 = *reinterpret_cast<const double2*>(zr + 2);
      }
    }
    zp[g][0] = pack_f32x2(__double2float_rn(v0.x), __double2float_rn(v0.y));
    zp[g][1] = pack_f32x2(__double2float_rn(v1.x), __double2float_rn(v1.y));
  }
  const float ninf = -INFINITY;
  float4 a_nx = make_float4(ninf, ninf, ninf, ninf);
  const float* ap = lb.A32 + (size_t)j_lo * lb.a_sj + n0;
  if (live && !ROWS) a_nx = *reinterpret_cast<const float4*>(ap);
  const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
  int wcnt = 0;  // entries in this warp's pool (warp-uniform)
  __syncthreads();

  for (int j = j_lo; j < j_hi; ++j) {
    float4 ag[G];
    if (ROWS) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        ag[g] = make_float4(ninf, ninf, ninf, ninf);
        if (live) ag[g] = *reinterpret_cast<const float4*>(lb.A32 + (size_t)(min(c0 + g, lb.C - 1) % lb.row_mod) * lb.a_sj + n0);
      }
    } else {
      ag[0] = ag[1] = ag[2] = ag[3] = a_nx;
      ap += lb.a_sj;
      if (live && j + 1 < j_hi) a_nx = *reinterpret_cast<const float4*>(ap);
    }
    const ulonglong2* pj = s_p32 + (size_t)(j - j_lo) * (2 * G);
    unsigned mask = 0u;
    {
      const ulonglong2 q1 = pj[0], q2 = pj[1];
      mask = pair_test32<1u << 0>(mask, q1.x, q1.y, q2.x, q2.y, zp[0][0], ag[0].x, ag[0].y);
      mask = pair_test32<1u << 2>(mask, q1.x, q1.y, q2.x, q2.y, zp[0][1], ag[0].z, ag[0].w);
    }
    {
      const ulonglong2 q1 = pj[2], q2 = pj[3];
      mask = pair_test32<1u << 4>(mask, q1.x, q1.y, q2.x, q2.y, zp[1][0], ag[1].x, ag[1].y);
      mask = pair_test32<1u << 6>(mask, q1.x, q1.y, q2.x, q2.y, zp[1][1], ag[1].z, ag[1].w);
    }
    {
      const ulonglong2 q1 = pj[4], q2 = pj[5];
      mask = pair_test32<1u << 8>(mask, q1.x, q1.y, q2.x, q2.y, zp[2][0], ag[2].x, ag[2].y);
      mask = pair_test32<1u << 10>(mask, q1.x, q1.y, q2.x, q2.y, zp[2][1], ag[2].z, ag[2].w);
    }
    {
      const ulonglong2 q1 = pj[6], q2 = pj[7];
      mask = pair_test32<1u << 12>(mask, q1.x, q1.y, q2.x, q2.y, zp[3][0], ag[3].x, ag[3].y);
      mask = pair_test32<1u << 14>(mask, q1.x, q1.y, q2.x, q2.y, zp[3][1], ag[3].z, ag[3].w);
    }
    for (;;) {  // survivors are rare (~1 % of the tests)
      const unsigned vote = __ballot_sync(0xffffffffu, mask != 0u);
      if (vote == 0u) break;
      if (wcnt + 32 > WPOOL) {
        flush_warp_pool<G, REFINE>(lb, sc, pool, wcnt, c0, j_lo, s_far);
        wcnt = 0;
      }
      if (mask) {
        const int bit = __ffs(mask) - 1;
        mask &= mask - 1u;
        pool[wcnt + __popc(vote & lt)] = make_int2(n0 + (bit & 3), j * G + (bit >> 2));
      }
      wcnt += __popc(vote);
    }
  }
  flush_warp_pool<G, REFINE>(lb, sc, pool, wcnt, c0, j_lo, s_far);
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * G * SB; e += blockDim.x) {
    const unsigned long long key = s_far[e];
    if (key) {
      const int setl = e >> 1;
      const int j = j_lo + setl / G, g = setl % G;
      atomicMax(&sc.far[((size_t)(c0 + g) * S + j) * 2 + (e & 1)], key);
    }
  }
}

static int launch_filter32(const LineBatch& lb, const EmaxScratch& sc, cudaStream_t st) {
  constexpr int G = 4;
  const int nblk = ceil_div(lb.NA, F32_THREADS * 4);
  const unsigned gy = ceil_div(lb.C, G);
  const int F32_JB = f32_jb(lb.S);
  const size_t smem = (size_t)G * F32_JB * (2 * sizeof(ulonglong2) + 2 * sizeof(unsigned long long)) +
                      (F32_THREADS / 32) * WPOOL * sizeof(int2);
  const unsigned gz = ceil_div(lb.S, F32_JB);
  const bool rows = lb.row_mod > 0;
  if (smem > 47 * 1024) {
    DKG_CUDA_OK(cudaFuncSetAttribute(filter32_kernel<G, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DKG_CUDA_OK(cudaFuncSetAttribute(filter32_kernel<G, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DKG_CUDA_OK(cudaFuncSetAttribute(filter32_kernel<G, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DKG_CUDA_OK(cudaFuncSetAttribute(filter32_kernel<G, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  // two sampled blocks (1024 lines) out of >= 8; smaller rows are filtered in one launch
  const bool two_phase = sc.chain5 != nullptr && sc.chainv != nullptr && nblk >= 8;
  int sample_blocks = 2;
  if (const char* e = getenv("DKG_FILTER_NSB")) sample_blocks = atoi(e) >= 1 && atoi(e) <= nblk / 4 ? atoi(e) : 2;
  const int nsb = two_phase ? sample_blocks : nblk, step = two_phase ? nblk / sample_blocks : 1;
  if (rows) filter32_kernel<G, false, true><<<dim3(nsb, gy, gz), F32_THREADS, smem, st>>>(lb, sc, step, nsb, F32_JB);
  else filter32_kernel<G, false, false><<<dim3(nsb, gy, gz), F32_THREADS, smem, st>>>(lb, sc, step, nsb, F32_JB);
  DKG_LAUNCH_CHECK();
  if (two_phase) {
    const long long sets = (long long)lb.C * lb.S;
    chain5_kernel<<<(unsigned)((sets + E_THREADS - 1) / E_THREADS), E_THREADS, 0, st>>>(lb, sc);
    DKG_LAUNCH_CHECK();
    if (rows) filter32_kernel<G, true, true><<<dim3(nblk - nsb, gy, gz), F32_THREADS, smem, st>>>(lb, sc, step, nsb, F32_JB);
    else filter32_kernel<G, true, false><<<dim3(nblk - nsb, gy, gz), F32_THREADS, smem, st>>>(lb, sc, step, nsb, F32_JB);
    DKG_LAUNCH_CHECK();
  }
  return DKG_OK;
}

// ---- coupled path, slope rows not materialised --------------------------------------------------
// The rows of one candidate (one per scalarisation) are all linear combinations of the same M covariance
// rows, so a CTA takes FC_CG candidates x 512 lines, every thread keeps the covariance values of its 4
// consecutive lines in registers (FC_CG x M x 4 doubles) and walks the scalarisations two at a time: the
// slope is M fp64 FMAs and one multiplication by 1 / sd (the float image of s * RN(1 / sd) differs from the
// float image of the correctly rounded quotient by at most 2^-52 |z| beyond its own 2^-24 |z| rounding -- far
// inside the 3.9 * 2^-24 margin chord32() leaves), the float intercepts of a scalarisation are loaded once
// for both candidates, and the test is the packed float test of filter32_kernel.  Survivors are re-formed
// exactly (line_slope) when they are written out, so the hull stage sees the same bits as every other stage.
constexpr int FC_CG = 2;  // candidates per CTA
constexpr int FC_JG = 2;  // scalarisations per step
constexpr int FC_NB = 4;  // line blocks (512 lines each) per CTA

template <int MT, bool REFINE>
__global__ void __launch_bounds__(F32_THREADS, 4)
filter32_cov_kernel(LineBatch lb, EmaxScratch sc, int blk_step, int nsb, int nblk_phase) {
  extern __shared__ __align__(16) unsigned char e_smem[];
  const int SR = lb.row_mod, M = lb.cov_M;
  const int ncand = lb.C / SR;
  const int cbase = blockIdx.y * FC_CG;
  const int row0 = cbase * SR;
  const int nset = FC_CG * SR;                                                            // local sets: cl * SR + j
  ulonglong2* s_p32 = reinterpret_cast<ulonglong2*>(e_smem);                              // [nset][2]
  unsigned long long* s_far = reinterpret_cast<unsigned long long*>(s_p32 + 2 * nset);    // [nset][2]
  double* s_ri = reinterpret_cast<double*>(s_far + 2 * nset);                             // [nset] 1 / sd
  double* s_w2 = s_ri + nset;                                                             // [SR][MT]
  int2* pool = reinterpret_cast<int2*>(s_w2 + SR * MT) + (threadIdx.x >> 5) * WPOOL;      // this warp's
  const ulonglong2* g_p32 = reinterpret_cast<const ulonglong2*>(sc.chain32);
  const ulonglong2 none = make_ulonglong2(0ull, 0x7f8000007f800000ull);  // (0, 0, +inf, +inf): nothing passes
  for (int e = threadIdx.x; e < nset; e += blockDim.x) {
    const bool in = cbase + e / SR < ncand;
    s_p32[2 * e] = in ? g_p32[((size_t)row0 + e) * 2] : none;
    s_p32[2 * e + 1] = in ? g_p32[((size_t)row0 + e) * 2 + 1] : none;
    s_far[2 * e] = 0ull;
    s_far[2 * e + 1] = 0ull;
    s_ri[e] = in ? 1.0 / lb.cov_sd[row0 + e] : 0.0;
  }
  for (int e = threadIdx.x; e < SR * MT; e += blockDim.x) {
    const int j = e / MT, m = e - j * MT;
    s_w2[e] = m < M ? lb.cov_w2[j * M + m] : 0.0;
  }
  const float ninf = -INFINITY;
  const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
  int wcnt = 0;  // entries in this warp's pool (warp-uniform)
  __syncthreads();
  // FC_NB line blocks per CTA: the chain parameters / reciprocals / weights above are loaded once for all of
  // them (with one block per CTA that set-up and its barrier were a fifth of the kernel's instructions)
  for (int bi = blockIdx.x * FC_NB; bi < min(nblk_phase, (int)(blockIdx.x + 1) * FC_NB); ++bi) {
  const int n0 = (line_block(bi, blk_step, nsb, REFINE) * F32_THREADS + (int)threadIdx.x) * 4;
  const bool live = n0 < lb.NA;  // lines n0..n0+3; beyond NA the float table holds -inf (never kept)
  double cv[FC_CG][MT][4];
#pragma unroll
  for (int cl = 0; cl < FC_CG; ++cl)
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      double2 v0 = make_double2(0.0, 0.0), v1 = v0;
      if (live && m < M) {
        const double* cr = lb.cov[m] + (size_t)min(cbase + cl, ncand - 1) * lb.ldz + n0;
        v0 = *reinterpret_cast<const double2*>(cr);
        v1 = *reinterpret_cast<const double2*>(cr + 2);
      }
      cv[cl][m][0] = v0.x; cv[cl][m][1] = v0.y; cv[cl][m][2] = v1.x; cv[cl][m][3] = v1.y;
    }
  for (int j0 = 0; j0 < SR; j0 += FC_JG) {
    unsigned mask = 0u;
#pragma unroll
    for (int jl = 0; jl < FC_JG; ++jl) {
      const int j = min(j0 + jl, SR - 1);
      float4 ag = make_float4(ninf, ninf, ninf, ninf);
      if (live && j0 + jl < SR) ag = *reinterpret_cast<const float4*>(lb.A32 + (size_t)j * lb.a_sj + n0);
      double w2[MT];
#pragma unroll
      for (int m = 0; m < MT; ++m) w2[m] = s_w2[j * MT + m];
#pragma unroll
      for (int cl = 0; cl < FC_CG; ++cl) {
        const int setl = cl * SR + j;
        const double ri = s_ri[setl];
        float zf[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          double sacc = 0.0;
#pragma unroll
          for (int m = 0; m < MT; ++m) sacc = fma(w2[m], cv[cl][m][u], sacc);  // (w2 = cv = 0 beyond M)
          zf[u] = __double2float_rn(sacc * ri);
        }
        const ulonglong2 q1 = s_p32[2 * setl], q2 = s_p32[2 * setl + 1];
        constexpr unsigned B0 = 1u;  // bit layout: (jl * FC_CG + cl) * 4 + line
        if (jl == 0 && cl == 0) {
          mask = pair_test32<B0 << 0>(mask, q1.x, q1.y, q2.x, q2.y, pack_f32x2(zf[0], zf[1]), ag.x, ag.y);
          mask = pair_test32<B0 << 2>(mask, q1.x, q1.y, q2.x, q2.y, pack_f32x2(zf[2], zf[3]), ag.z, ag.w);
        } else if (jl == 0 && cl == 1) {
          mask = pair_test32<B0 << 4>(mask, q1.x, q1.y, q2.x, q2.y, pack_f32x2(zf[0], zf[1]), ag.x, ag.y);
          mask = pair_test32<B0 << 6>(mask, q1.x, q1.y, q2.x, q2.y, pack_f32x2(zf[2], zf[3]), ag.z, ag.w);
        } else if (jl == 1 && cl == 0) {
          mask = pair_test32<B0 << 8>(mask, q1.x, q1.y, q2.x, q2.y, pack_f32x2(zf[0], zf[1]), ag.x, ag.y);
          mask = pair_test32<B0 << 10>(mask, q1.x, q1.y, q2.x, q2.y, pack_f32x2(zf[2], zf[3]), ag.z, ag.w);
        } else {
          mask = pair_test32<B0 << 12>(mask, q1.x, q1.y, q2.x, q2.y, pack_f32x2(zf[0], zf[1]), ag.x, ag.y);
          mask = pair_test32<B0 << 14>(mask, q1.x, q1.y, q2.x, q2.y, pack_f32x2(zf[2], zf[3]), ag.z, ag.w);
        }
      }
    }
    for (;;) {  // survivors are rare (~1 % of the tests)
      const unsigned vote = __ballot_sync(0xffffffffu, mask != 0u);
      if (vote == 0u) break;
      if (wcnt + 32 > WPOOL) {
        flush_warp_pool<(1 << 20), REFINE>(lb, sc, pool, wcnt, row0, 0, s_far);
        wcnt = 0;
      }
      if (mask) {
        const int bit = __ffs(mask) - 1;
        mask &= mask - 1u;
        const int slot = bit >> 2;  // jl * FC_CG + cl
        pool[wcnt + __popc(vote & lt)] = make_int2(n0 + (bit & 3), (slot % FC_CG) * SR + j0 + slot / FC_CG);
      }
      wcnt += __popc(vote);
    }
  }
  }  // line blocks
  flush_warp_pool<(1 << 20), REFINE>(lb, sc, pool, wcnt, row0, 0, s_far);
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * nset; e += blockDim.x) {
    const unsigned long long key = s_far[e];
    if (key) atomicMax(&sc.far[((size_t)row0 + (e >> 1)) * 2 + (e & 1)], key);
  }
}

template <int MT>
static int launch_filter32_cov(const LineBatch& lb, const EmaxScratch& sc, cudaStream_t st) {
  static_assert(FC_CG == 2 && FC_JG == 2, "bit layout of filter32_cov_kernel");
  const int SR = lb.row_mod, ncand = lb.C / SR;
  const int nblk = ceil_div(lb.NA, F32_THREADS * 4);
  const unsigned gy = ceil_div(ncand, FC_CG);
  const size_t smem = (size_t)FC_CG * SR * (2 * sizeof(ulonglong2) + 2 * sizeof(unsigned long long) + sizeof(double)) +
                      (size_t)SR * MT * sizeof(double) + (F32_THREADS / 32) * WPOOL * sizeof(int2);
  if (smem > 47 * 1024) {
    DKG_CUDA_OK(cudaFuncSetAttribute(filter32_cov_kernel<MT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DKG_CUDA_OK(cudaFuncSetAttribute(filter32_cov_kernel<MT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  // (two phases as in launch_filter32: sampled line blocks, second-level chain, the rest)
  const bool two_phase = sc.chain5 != nullptr && sc.chainv != nullptr && nblk >= 8;
  const int nsb = two_phase ? 2 : nblk, step = two_phase ? nblk / 2 : 1;
  filter32_cov_kernel<MT, false><<<dim3(ceil_div(nsb, FC_NB), gy), F32_THREADS, smem, st>>>(lb, sc, step, nsb, nsb);
  DKG_LAUNCH_CHECK();
  if (two_phase) {
    chain5_kernel<<<(unsigned)((lb.C + E_THREADS - 1) / E_THREADS), E_THREADS, 0, st>>>(lb, sc);
    DKG_LAUNCH_CHECK();
    filter32_cov_kernel<MT, true><<<dim3(ceil_div(nblk - nsb, FC_NB), gy), F32_THREADS, smem, st>>>(lb, sc, step, nsb, nblk - nsb);
    DKG_LAUNCH_CHECK();
  }
  return DKG_OK;
}

template <int G, int R>
static int launch_filter(const LineBatch& lb, const EmaxScratch& sc, cudaStream_t st) {
  dim3 grid(ceil_div(lb.NA, E_THREADS * R), ceil_div(lb.C, G));
  const size_t smem = (size_t)G * lb.S * (sizeof(double4) + 3 * sizeof(int) + 2 * sizeof(unsigned long long)) +
                      POOL_CAP * sizeof(int2);
  if (smem > 47 * 1024) {  // many scalarisations: opt in to large dynamic shared memory
    DKG_CUDA_OK(cudaFuncSetAttribute(filter_kernel<G, R, true>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DKG_CUDA_OK(cudaFuncSetAttribute(filter_kernel<G, R, false>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (lb.a_sc == 0 && lb.row_mod == 0)
    filter_kernel<G, R, true><<<grid, E_THREADS, smem, st>>>(lb, sc);
  else
    filter_kernel<G, R, false><<<grid, E_THREADS, smem, st>>>(lb, sc);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

int emax_filter(const LineBatch& lb, const EmaxScratch& sc, cudaStream_t st, bool ztile_valid) {
  if (lb.C == 0 || lb.NA == 0) return DKG_OK;
  {
    // KG path on a large discretisation: the tile-first filter (DKG_FILTER=f64 / f32 / line32 select the older kernels)
    const char* fe0 = getenv("DKG_FILTER");
    if (fe0 == nullptr && ztile_valid && lb.A32 != nullptr && lb.A32tmax != nullptr && sc.chain32 != nullptr &&
        sc.chain5 != nullptr && sc.chain5f != nullptr && sc.chainv != nullptr && lb.a_sc == 0 && lb.row_mod == 0 &&
        lb.NA >= 8 * FILTER_TILE && (lb.ldz & 1) == 0 && (lb.a_sj & 3) == 0 && lb.ldz >= ((lb.NA + 3) & ~3) &&
        lb.a_sj >= ((lb.NA + 3) & ~3))
      return launch_tilefilter(lb, sc, st, ztile_valid);
  }
  // big batches: 4 lines per thread (fewer parameter loads per test); small ones: more CTAs
  const long long ctas4 = (long long)ceil_div(lb.C, 4) * ceil_div(lb.NA, E_THREADS * 4);
  // KG path with a shared intercept table: the float test (DKG_FILTER=f64 keeps the fp64 kernel)
  // (DKG_FILTER=f32 uses it for small batches too: tests)
  const char* fe = getenv("DKG_FILTER");  // read per launch so tests can compare the two kernels
  const int mode = fe == nullptr ? 0 : strcmp(fe, "f64") == 0 ? 1 : strcmp(fe, "f32") == 0 ? 2 : 0;
  if (mode != 1 && lb.A32 != nullptr && sc.chain32 != nullptr && lb.a_sc == 0 && (lb.row_mod == 0 || lb.S == 1) &&
      (lb.ldz & 1) == 0 && (lb.a_sj & 3) == 0 && lb.ldz >= ((lb.NA + 3) & ~3) && lb.a_sj >= ((lb.NA + 3) & ~3) &&
      (ctas4 >= 148 || mode == 2)) {
    if (lb.cov_M > 0 && lb.cov_M <= 4 && lb.row_mod > 0 && lb.C % lb.row_mod == 0 && getenv("DKG_FILTER_COV") == nullptr)
      return lb.cov_M <= 2 ? launch_filter32_cov<2>(lb, sc, st) : launch_filter32_cov<4>(lb, sc, st);
    return launch_filter32(lb, sc, st);
  }
  if (getenv("DKG_FILTER_R8") && ctas4 >= 6 * 148) return launch_filter<4, 8>(lb, sc, st);
  if (ctas4 >= 3 * 148) return launch_filter<4, 4>(lb, sc, st);
  return launch_filter<4, 1>(lb, sc, st);
}

// ------------------------------------------------------------------------------------------
// exact march (restates discretekg.py:370-410 on whatever lines `fetch` exposes)
// ------------------------------------------------------------------------------------------
struct Line {
  double a, b;
  int idx;  // internal index (own line == NA); -1 = empty
  int ref;  // position in the reference's ordering of the inputs (own line first)
};

__device__ __forceinline__ Line empty_line() {
  Line L;
  L.a = 0.0; L.b = 0.0; L.idx = -1; L.ref = 0x7fffffff;
  return L;
}

// ordering of the reference's sorted array: slope ascending, then intercept descending; among
// identical lines the input position decides (the reference's first sort is not stable there).
__device__ __forceinline__ bool sorted_before(const Line& p, const Line& q) {
  if (p.b != q.b) return p.b < q.b;
  if (p.a != q.a) return p.a > q.a;
  return p.ref < q.ref;
}

__device__ __forceinline__ int ref_index(const LineBatch& lb, int idx) {
  return (lb.a_own != nullptr) ? (idx == lb.NA ? 0 : idx + 1) : idx;
}

__device__ __forceinline__ Line make_line(const LineBatch& lb, double w, double a, double z, int idx) {
  Line L;
  L.a = a;
  L.b = __dmul_rn(w, z);  // slopes = weights[..., i] * znew_coefficients (:321)
  L.idx = idx;
  L.ref = ref_index(lb, idx);
  return L;
}

__device__ __forceinline__ Line gather_line(const LineBatch& lb, int c, int j, double w, int idx) {
  return make_line(lb, w, line_intercept(lb, c, j, idx), line_slope(lb, c, idx), idx);
}

__device__ __forceinline__ Line shfl_line(const Line& L, int o) {
  Line r;
  r.a = __shfl_xor_sync(0xffffffffu, L.a, o);
  r.b = __shfl_xor_sync(0xffffffffu, L.b, o);
  r.idx = __shfl_xor_sync(0xffffffffu, L.idx, o);
  r.ref = __shfl_xor_sync(0xffffffffu, L.ref, o);
  return r;
}

// candidate for the next hull vertex: line + its intersection with the current line
struct Next {
  Line L;
  double x;
};
__device__ __forceinline__ void consider_next(Next& best, const Line& cur, const Line& L) {
  if (L.idx >= 0 && L.b > cur.b) {  // strictly different slope (:388); sorted => larger
    const double x = -(cur.a - L.a) / (cur.b - L.b);  // (:395)
    if (best.L.idx < 0 || x < best.x || (x == best.x && sorted_before(L, best.L))) {
      best.L = L;
      best.x = x;
    }
  }
}
__device__ __forceinline__ void merge_next(Next& best, const Next& o) {
  if (o.L.idx >= 0 &&
      (best.L.idx < 0 || o.x < best.x || (o.x == best.x && sorted_before(o.L, best.L))))
    best = o;
}

// per-set output sink: the first hull_cap records inline, the rest in chained spill blocks
static_assert(HULL_CAP % SPILL_BLOCK == 0, "a march batch must map to one spill block");
struct Recorder {
  const EmaxOut* out;
  size_t set;
  int NL;
  int blk = -1;          // current spill block
  bool dropped = false;  // the pool ran dry for this set

  __device__ __forceinline__ void dense(const Line& L, double p, double q) const {
    const EmaxOut& o = *out;
    if (o.dense_da) o.dense_da[set * (size_t)NL + L.idx] = p;
    if (o.dense_db) o.dense_db[set * (size_t)NL + L.idx] = q;
  }
  __device__ __forceinline__ void store_inline(int k, const Line& L, double p, double q, double x, bool last) const {
    const EmaxOut& o = *out;
    const size_t r = set * (size_t)o.hull_cap + k;
    if (o.hull_idx) o.hull_idx[r] = L.idx;
    if (o.hull_p) o.hull_p[r] = p;
    if (o.hull_q) o.hull_q[r] = q;
    if (o.hull_x && !last) o.hull_x[r] = x;
  }
  // claim the block holding records k .. (called by ONE thread); links it behind the current one
  __device__ __forceinline__ int claim(int k) {
    const EmaxOut& o = *out;
    int nb = atomicAdd(o.spill_used, 1);
    if (nb >= o.spill_blocks) nb = -1;
    else o.spill_next[nb] = -1;
    if (k == o.hull_cap) o.spill_head[set] = nb;
    else o.spill_next[blk] = nb;
    if (nb < 0 && o.truncated != nullptr) atomicAdd((unsigned long long*)o.truncated, 1ull);
    return nb;
  }
  __device__ __forceinline__ void store_spill(int k, const Line& L, double p, double q) const {
    const EmaxOut& o = *out;
    const size_t r = (size_t)blk * SPILL_BLOCK + (k - o.hull_cap) % SPILL_BLOCK;
    o.spill_idx[r] = L.idx;
    o.spill_p[r] = p;
    o.spill_q[r] = q;
  }
  // one record, from a single thread that calls this for k = 0, 1, 2, ... in order
  __device__ void single(int k, const Line& L, double p, double q, double x, bool last) {
    const EmaxOut& o = *out;
    if (k < o.hull_cap) store_inline(k, L, p, q, x, last);
    else if (o.spill_head != nullptr && !dropped) {
      if ((k - o.hull_cap) % SPILL_BLOCK == 0) { blk = claim(k); dropped = blk < 0; }
      if (!dropped) store_spill(k, L, p, q);
    }
    dense(L, p, q);
  }
  // warp-wide: lane l holds record k0 + l of a march batch (l < cnt); k0 is a multiple of 32
  __device__ void batch(int k0, int cnt, const Line& L, double p, double q, double x, bool last_batch) {
    const EmaxOut& o = *out;
    const int lane = threadIdx.x & 31;
    const bool act = lane < cnt;
    const int k = k0 + lane;
    if (act && k < o.hull_cap) store_inline(k, L, p, q, x, last_batch && lane == cnt - 1);
    if (o.spill_head != nullptr && k0 >= o.hull_cap && !dropped) {
      int nb = 0;
      if (lane == 0) nb = claim(k0);
      blk = __shfl_sync(0xffffffffu, nb, 0);
      dropped = blk < 0;
      if (act && !dropped) store_spill(k, L, p, q);
    }
    if (act) dense(L, p, q);
  }
};

struct HullResult {
  double E;
  int h;
};

// One warp marches over `total` lines exposed by fetch(k); the first 32*LANE_LINES are cached
// in registers.  Accumulates sum_k a_k dPhi_k - b_k dphi_k (:449-451) in hull order.
constexpr int LANE_LINES = STAGE_CAP / 32;

template <class Fetch>
__device__ HullResult warp_march(int total, Fetch fetch, Recorder& rec) {
  const int lane = threadIdx.x & 31;
  Line cache[LANE_LINES];
#pragma unroll
  for (int r = 0; r < LANE_LINES; ++r) {
    const int k = lane + 32 * r;
    cache[r] = (k < total) ? fetch(k) : empty_line();
  }
  // first line of the sorted order: minimum slope, maximum intercept among ties (:371-374)
  Line cur = empty_line();
#pragma unroll
  for (int r = 0; r < LANE_LINES; ++r)
    if (cache[r].idx >= 0 && (cur.idx < 0 || sorted_before(cache[r], cur))) cur = cache[r];
  for (int k = lane + 32 * LANE_LINES; k < total; k += 32) {
    const Line L = fetch(k);
    if (cur.idx < 0 || sorted_before(L, cur)) cur = L;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const Line oth = shfl_line(cur, o);
    if (oth.idx >= 0 && (cur.idx < 0 || sorted_before(oth, cur))) cur = oth;
  }

  // The march only needs the intersections; Phi / phi of the breakpoints are evaluated afterwards
  // for up to 32 vertices at once (lane k owns vertex k of the batch), so the erf / exp latency is
  // paid once per batch instead of once per vertex.  The expectation is still summed in hull order.
  double E = 0.0, carry_cdf = 0.0, carry_pdf = 0.0;
  int h = 0;
  Line mine = empty_line();   // vertex owned by this lane in the current batch
  double mine_x = INFINITY;   // its right breakpoint (+inf for the last vertex)
  while (true) {
    // next vertex: among lines with a strictly different (larger) slope, the one whose intersection
    // with the current line comes first (:388-396); ties -> earliest in the sorted order
    Next best;
    best.L = empty_line();
    best.x = INFINITY;
#pragma unroll
    for (int r = 0; r < LANE_LINES; ++r) consider_next(best, cur, cache[r]);
    for (int k = lane + 32 * LANE_LINES; k < total; k += 32) consider_next(best, cur, fetch(k));
    const unsigned have = __ballot_sync(0xffffffffu, best.L.idx >= 0);
    const bool last = have == 0u;
    Line nxt = empty_line();
    double nx = INFINITY;
    if (!last) {
      // Smallest intersection of the warp: rounding to float is monotone, so the exact minimum is held by a
      // lane whose ROUNDED value equals the smallest rounded value -- one CREDUX instead of five rounds of 64-bit
      // shuffles and NaN-aware fp64 minima.  Usually one lane qualifies; several (equal or float-equal
      // intersections) are merged exactly: smaller x first, then the earliest line in the sorted order.
      const double myx = best.L.idx >= 0 ? best.x : INFINITY;
      const float fx = __double2float_rn(myx);
      const float fm = warp_min_f32(fx);
      const unsigned tied = __ballot_sync(0xffffffffu, best.L.idx >= 0 && fx == fm);
      const int src = __ffs(tied) - 1;
      if (__popc(tied) > 1) {
        Next b2 = best;
        if (!((tied >> lane) & 1u)) b2.L = empty_line();
        for (int o = 16; o > 0; o >>= 1) {
          Next oth;
          oth.L = shfl_line(b2.L, o);
          oth.x = __shfl_xor_sync(0xffffffffu, b2.x, o);
          merge_next(b2, oth);
        }
        nxt = b2.L;
        nx = b2.x;
      } else {
        nxt.a = __shfl_sync(0xffffffffu, best.L.a, src);
        nxt.b = __shfl_sync(0xffffffffu, best.L.b, src);
        nxt.idx = __shfl_sync(0xffffffffu, best.L.idx, src);
        nxt.ref = __shfl_sync(0xffffffffu, best.L.ref, src);
        nx = __shfl_sync(0xffffffffu, myx, src);
      }
    }
    const int slot = h & 31;
    if (lane == slot) { mine = cur; mine_x = nx; }
    ++h;
    if (slot == 31 || last) {
      const int cnt = slot + 1;
      const bool act = lane < cnt;
      const double cdf = act ? std_normal_cdf(mine_x) : 0.0;  // Phi(+inf) = 1, phi(+inf) = 0
      const double pdf = act ? std_normal_pdf(mine_x) : 0.0;
      double lcdf = __shfl_up_sync(0xffffffffu, cdf, 1);
      double lpdf = __shfl_up_sync(0xffffffffu, pdf, 1);
      if (lane == 0) { lcdf = carry_cdf; lpdf = carry_pdf; }
      const double dP = cdf - lcdf, dp = pdf - lpdf;
      // intercepts * (cdf[1:] - cdf[:-1]) - slopes * (pdf[1:] - pdf[:-1])   (:449-451)
      const double term = act ? __dsub_rn(__dmul_rn(mine.a, dP), __dmul_rn(mine.b, dp)) : 0.0;
      rec.batch(h - cnt, cnt, mine, dP, -dp, mine_x, last);
      for (int k = 0; k < cnt; ++k) E += __shfl_sync(0xffffffffu, term, k);
      carry_cdf = __shfl_sync(0xffffffffu, cdf, cnt - 1);
      carry_pdf = __shfl_sync(0xffffffffu, pdf, cnt - 1);
    }
    if (last) break;
    cur = nxt;
  }
  HullResult res;
  res.E = E;
  res.h = h;
  return res;
}

__device__ __forceinline__ void finish_set(const LineBatch& lb, const EmaxOut& out, size_t set,
                                           const SetInfo& s, double E, int h) {
  if (out.hull_cnt) out.hull_cnt[set] = h;
  // more hull vertices than record slots and no spill chain: E is exact, but a backward would miss
  // vertices (with a chain the Recorder counts the sets whose blocks could not be claimed)
  if (out.truncated != nullptr && out.hull_p != nullptr && out.spill_head == nullptr && h > out.hull_cap)
    atomicAdd((unsigned long long*)out.truncated, 1ull);
  out.terms[set] = out.subtract_max ? (E - s.amax) : E;  // kg[j] = E - max(intercepts) (:336)
}

// chord through two dual points; excess(L) > 0 <=> L lies above the chord
struct Chord {
  double b0, a0, m, slack;
  __device__ void set(const Line& l, const Line& r) {
    b0 = l.b;
    a0 = l.a;
    m = (r.a - l.a) / (r.b - l.b);
    slack = EPS128 * (fabs(l.a) + fabs(r.a) + fabs(m) * fmax(fabs(l.b), fabs(r.b)));
  }
  __device__ double excess(const Line& L) const { return L.a - fma(m, L.b - b0, a0); }
};

__device__ __forceinline__ bool higher_slope(const Line& p, const Line& q) {  // p better as "Q"
  if (p.b != q.b) return p.b > q.b;
  if (p.a != q.a) return p.a > q.a;
  return p.ref < q.ref;
}
__device__ __forceinline__ bool higher_intercept(const Line& p, const Line& q) {
  if (p.a != q.a) return p.a > q.a;
  return p.ref < q.ref;
}

// ------------------------------------------------------------------------------------------
// hull kernel: one warp per set
// ------------------------------------------------------------------------------------------
// hull_short_kernel<G>: G lanes per set (8 or 16), 4 lines per lane in registers -> sets of up to 32 / 64 lines

// one warp, one set (the warp's slices of the staging arrays are passed in)
__device__ __forceinline__ void hull_one_set(const LineBatch& lb, const EmaxScratch& sc, const EmaxOut& out, size_t set, int skip_short,
                                             double* sa_w, double* sb_w, int* si_w, double* scm_w, double* scs_w,
                                             unsigned long long* sfk_w) {
  const int lane = threadIdx.x & 31;
  const int c = (int)(set / lb.S);
  const int j = (int)(set - (size_t)c * lb.S);
  const SetInfo s = set_info(lb, sc, c, j);
  if (out.amax_is_own != nullptr && lane == 0) out.amax_is_own[set] = s.own_is_max;
  Recorder rec{&out, set, lb.NL};

  if (s.shortcut) {
    // all |slopes| < 1e-9: the reference returns argmax(intercepts) only (:363-367)
    if (lane == 0) {
      Line L;
      L.idx = s.iT; L.a = s.amax; L.b = 0; L.ref = 0;
      rec.single(0, L, 1.0, 0.0, 0.0, true);
      finish_set(lb, out, set, s, s.amax, 1);
      if (sc.stats) atomicAdd((unsigned long long*)&sc.stats[4], 1ull);
    }
    return;
  }
  const int cnt = sc.surv_cnt[set];
  if (cnt > SURV_CAP) {  // list truncated: the cooperative kernel redoes this set from all lines
    if (lane == 0) {
      sc.ovf_sets[atomicAdd(sc.ovf_count, 1)] = (int)set;
      if (sc.stats) atomicAdd((unsigned long long*)&sc.stats[0], 1ull);  // (debug: DKG_DEBUG_STATS)
    }
    return;
  }
  const SurvEntry* list = sc.surv + set * SURV_CAP;
  int seeds[4];
  int nseed = 0;
  seeds[nseed++] = s.iP;
  seeds[nseed++] = s.iQ;
  seeds[nseed++] = lb.Aarg[am_index(lb, c, j)];
  if (lb.a_own != nullptr) seeds[nseed++] = lb.NA;
  const int total = cnt + nseed;
  if (skip_short && total <= skip_short) return;  // hull_short_kernel finishes these, several to a warp
  const double w = s.w;
  auto fetch_global = [&](int k) -> Line {
    if (k < cnt) {
      const SurvEntry e = list[k];
      return make_line(lb, w, e.a, e.z, e.idx);
    }
    return gather_line(lb, c, j, w, seeds[k - cnt]);
  };

  HullResult r;
  int staged = 0;
  if (total <= STAGE_CAP) {
    // short list: stage it as it is (one march instantiation for both branches keeps this kernel's
    // code, and its instruction-cache footprint, a third smaller)
    for (int k = lane; k < total; k += 32) {
      const Line L = fetch_global(k);
      sa_w[k] = L.a; sb_w[k] = L.b; si_w[k] = L.idx;
    }
    staged = total;
  } else {
    // ---- QuickHull-style refinement over the list, compacting into shared memory ----
    // chain in the (b, a) plane: P (min slope), T (max intercept), Q (max slope) plus the
    // farthest late survivors F1 / F2 recorded by the filter (any set member is a valid vertex).
    // The chain vertices (strictly increasing slope) are the head of the staging arrays; every
    // pass classifies the list against the current chain, optimistically compacts the lines that
    // stay into the tail and records the farthest line above every chord; if the result does
    // not fit, those lines become vertices (the chain stays concave) and the pass is repeated.
    const bool posw = !(w < 0.0);
    const Line P = gather_line(lb, c, j, w, posw ? s.iP : s.iQ);
    const Line Q = gather_line(lb, c, j, w, posw ? s.iQ : s.iP);
    const Line T = gather_line(lb, c, j, w, s.iT);
    const unsigned long long f1 = sc.far[set * 2 + 0], f2 = sc.far[set * 2 + 1];
    Line F1 = f1 ? gather_line(lb, c, j, w, (int)(f1 & 0xffffffffull)) : empty_line();
    Line F2 = f2 ? gather_line(lb, c, j, w, (int)(f2 & 0xffffffffull)) : empty_line();
    const bool hasL = T.b > P.b, hasR = Q.b > T.b;
    if (!(hasL && F1.idx >= 0 && F1.b > P.b && F1.b < T.b)) F1 = empty_line();
    if (!(hasR && F2.idx >= 0 && F2.b > T.b && F2.b < Q.b)) F2 = empty_line();
    if (F1.idx >= 0) { Chord ch; ch.set(P, T); if (!(ch.excess(F1) > 0.0)) F1 = empty_line(); }  // concave
    if (F2.idx >= 0) { Chord ch; ch.set(T, Q); if (!(ch.excess(F2) > 0.0)) F2 = empty_line(); }
    double* vb = sb_w;
    double* va = sa_w;
    int* vi = si_w;
    int nv = 0;
    if (lane == 0) {
      const Line vs[5] = {hasL ? P : empty_line(), F1, T, F2, hasR ? Q : empty_line()};
      for (int v = 0; v < 5; ++v)
        if (vs[v].idx >= 0) { va[nv] = vs[v].a; vb[nv] = vs[v].b; vi[nv] = vs[v].idx; ++nv; }
    }
    nv = __shfl_sync(0xffffffffu, nv, 0);
    bool fits = false;
    for (int level = 0; level < HULL_LEVELS; ++level) {
      __syncwarp();
      if (lane < nv - 1) {  // chord lane: vertex lane -> vertex lane + 1
        const double m = (va[lane + 1] - va[lane]) / (vb[lane + 1] - vb[lane]);
        scm_w[lane] = m;
        scs_w[lane] = EPS128 * (fabs(va[lane]) + fabs(va[lane + 1]) + fabs(m) * fmax(fabs(vb[lane]), fabs(vb[lane + 1])));
      }
      if (lane < CHAIN_MAXV) sfk_w[lane] = 0ull;
      __syncwarp();
      staged = nv;
      for (int k0 = 0; k0 < total; k0 += 32) {
        const int k = k0 + lane;
        bool keep = false;
        Line L = empty_line();
        if (k < total) {
          L = fetch_global(k);
          // a line strictly inside a chord's slope range survives only above that chord; a line
          // at a vertex slope is kept (the march discards it if it is dominated)
          keep = true;
          int q = 0;  // largest vertex with vb[q] <= L.b (P / Q are the extreme slopes of the set)
          for (int step = CHAIN_MAXV / 2; step > 0; step >>= 1)
            if (q + step < nv && vb[q + step] <= L.b) q += step;
          if (q < nv - 1 && L.b > vb[q]) {
            const double ex = L.a - fma(scm_w[q], L.b - vb[q], va[q]);
            keep = ex > -scs_w[q];
            if (ex > 0.0) atomicMax(&sfk_w[q], pack_excess(ex, k));
          }
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        const int pos = staged + __popc(m & ((1u << lane) - 1u));
        if (keep && pos < STAGE_CAP) { sa_w[pos] = L.a; sb_w[pos] = L.b; si_w[pos] = L.idx; }
        staged += __popc(m);
      }
      if (staged <= STAGE_CAP) { fits = true; break; }
      // the farthest line above every chord becomes a vertex
      __syncwarp();
      const unsigned long long key = lane < nv - 1 ? sfk_w[lane] : 0ull;
      const unsigned ins = __ballot_sync(0xffffffffu, key != 0ull);
      if (ins == 0u || nv + __popc(ins) > CHAIN_MAXV || level + 1 == HULL_LEVELS) break;
      const Line F = key ? fetch_global((int)(key & 0xffffffffull)) : empty_line();
      const double oa = lane < nv ? va[lane] : 0.0, ob = lane < nv ? vb[lane] : 0.0;
      const int oi = lane < nv ? vi[lane] : -1;
      const int before = __popc(ins & ((1u << lane) - 1u));  // inserts from chords 0 .. lane-1
      __syncwarp();
      if (lane < nv) { va[lane + before] = oa; vb[lane + before] = ob; vi[lane + before] = oi; }
      if (key) { va[lane + before + 1] = F.a; vb[lane + before + 1] = F.b; vi[lane + before + 1] = F.idx; }
      nv += __popc(ins);
    }
    if (!fits) {
      if (lane == 0) sc.ovf_sets[atomicAdd(sc.ovf_count, 1)] = (int)set;
      return;
    }
  }
  __syncwarp();
  auto fetch_smem = [&](int k) -> Line {
    Line L;
    L.a = sa_w[k]; L.b = sb_w[k]; L.idx = si_w[k];
    L.ref = ref_index(lb, L.idx);
    return L;
  };
  r = warp_march(staged, fetch_smem, rec);
  if (lane == 0) {
    finish_set(lb, out, set, s, r.E, r.h);
    if (sc.stats) {
      atomicAdd((unsigned long long*)&sc.stats[1], (unsigned long long)cnt);
      atomicAdd((unsigned long long*)&sc.stats[3], (unsigned long long)r.h);
    }
  }
}

// One warp per set, or -- when hull_short_kernel ran first and queued the sets it left (sc.long_sets) -- a fixed
// grid of warps walking that queue: a warp per set just to find out that the set was short cost 60 us per launch.
__global__ void __launch_bounds__(E_THREADS, HULL_CTAS)
hull_kernel(LineBatch lb, EmaxScratch sc, EmaxOut out, int skip_short, int from_queue) {
  __shared__ double s_a[E_THREADS / 32][STAGE_CAP];
  __shared__ double s_b[E_THREADS / 32][STAGE_CAP];
  __shared__ int s_i[E_THREADS / 32][STAGE_CAP];
  __shared__ double s_cm[E_THREADS / 32][CHAIN_MAXV];  // chord slopes / slacks of the refinement chain
  __shared__ double s_cs[E_THREADS / 32][CHAIN_MAXV];
  __shared__ unsigned long long s_fk[E_THREADS / 32][CHAIN_MAXV];  // farthest line above every chord
  const int warp = threadIdx.x >> 5;
  if (from_queue) {
    const int n = *sc.long_count;
    for (int qi = blockIdx.x * (E_THREADS / 32) + warp; qi < n; qi += gridDim.x * (E_THREADS / 32)) {
      hull_one_set(lb, sc, out, (size_t)sc.long_sets[qi], 0, s_a[warp], s_b[warp], s_i[warp], s_cm[warp], s_cs[warp], s_fk[warp]);
      __syncwarp();
    }
    return;
  }
  const long long set_ll = (long long)blockIdx.x * (E_THREADS / 32) + warp;
  if (set_ll >= (long long)lb.C * lb.S) return;
  hull_one_set(lb, sc, out, (size_t)set_ll, skip_short, s_a[warp], s_b[warp], s_i[warp], s_cm[warp], s_cs[warp], s_fk[warp]);
}

// Short sets (survivors + seeds <= 32 lines: most sets of a well-filtered batch): a warp per set leaves
// three quarters of the lanes without a line.  Here 8 lanes share a set (4 lines per lane in registers),
// four sets per warp advance in lock-step through the same exact march; the Phi / phi evaluations of a
// set's vertices are batched (8 per group) and deferred to one flush at the end wherever possible.
template <int HS_G>
__global__ void __launch_bounds__(E_THREADS, 3)
hull_short_kernel(LineBatch lb, EmaxScratch sc, EmaxOut out) {
  constexpr int HS_LINES = 4 * HS_G;  // lines a group holds in registers (4 per lane)
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, gl = lane & (HS_G - 1);
  const long long nsets = (long long)lb.C * lb.S;
  const long long set_ll = ((long long)blockIdx.x * (E_THREADS / 32) + (threadIdx.x >> 5)) * (32 / HS_G) + (lane / HS_G);
  const bool in_range = set_ll < nsets;
  const size_t set = in_range ? (size_t)set_ll : 0;
  const int c = (int)(set / lb.S);
  const int j = (int)(set - (size_t)c * lb.S);
  const SetInfo s = set_info(lb, sc, c, j);
  const int cnt = sc.surv_cnt[set];
  int seeds[4];
  int nseed = 0;
  seeds[nseed++] = s.iP;
  seeds[nseed++] = s.iQ;
  seeds[nseed++] = lb.Aarg[am_index(lb, c, j)];
  if (lb.a_own != nullptr) seeds[nseed++] = lb.NA;
  const int total = cnt + nseed;
  bool active = in_range && !s.shortcut && cnt <= SURV_CAP && total <= HS_LINES;  // (group-uniform)
  if (in_range && gl == 0) {
    if (active) {
      if (out.amax_is_own != nullptr) out.amax_is_own[set] = s.own_is_max;
    } else if (sc.long_sets != nullptr) {
      sc.long_sets[atomicAdd(sc.long_count, 1)] = (int)set;  // left to hull_kernel (long list, shortcut set, truncated list)
    }
  }
  const double w = s.w;
  const SurvEntry* list = sc.surv + set * SURV_CAP;
  Line cache[HS_LINES / HS_G];
#pragma unroll
  for (int r = 0; r < HS_LINES / HS_G; ++r) {
    const int k = gl + HS_G * r;
    cache[r] = empty_line();
    if (active && k < total) {
      if (k < cnt) {
        const SurvEntry e = list[k];
        cache[r] = make_line(lb, w, e.a, e.z, e.idx);
      } else {
        cache[r] = gather_line(lb, c, j, w, seeds[k - cnt]);
      }
    }
  }
  // first line of the sorted order (:371-374)
  Line cur = empty_line();
#pragma unroll
  for (int r = 0; r < HS_LINES / HS_G; ++r)
    if (cache[r].idx >= 0 && (cur.idx < 0 || sorted_before(cache[r], cur))) cur = cache[r];
#pragma unroll
  for (int o = HS_G / 2; o > 0; o >>= 1) {
    const Line oth = shfl_line(cur, o);
    if (oth.idx >= 0 && (cur.idx < 0 || sorted_before(oth, cur))) cur = oth;
  }
  Recorder rec{&out, set, lb.NL};
  double E = 0.0, carry_cdf = 0.0, carry_pdf = 0.0;
  int h = 0, hbase = 0;      // vertices found / already flushed
  Line mine = empty_line();  // vertex hbase + gl of this group's current batch
  double mine_x = INFINITY;
  bool finished = false;     // the march of this group has ended; its last batch waits for the final flush
  int fin_cnt = 0;

  // evaluates Phi / phi for the batches of the groups with `go` (cnt_b vertices each) and folds them in
  auto flush = [&](bool go, int cnt_b, bool is_last) {
    const bool act = go && gl < cnt_b;
    const double cdf = act ? std_normal_cdf(mine_x) : 0.0;
    const double pdf = act ? std_normal_pdf(mine_x) : 0.0;
    double lcdf = __shfl_up_sync(full, cdf, 1, HS_G);
    double lpdf = __shfl_up_sync(full, pdf, 1, HS_G);
    if (gl == 0) { lcdf = carry_cdf; lpdf = carry_pdf; }
    const double dP = cdf - lcdf, dp = pdf - lpdf;
    const double term = act ? __dsub_rn(__dmul_rn(mine.a, dP), __dmul_rn(mine.b, dp)) : 0.0;
    if (act) {
      rec.store_inline(hbase + gl, mine, dP, -dp, mine_x, is_last && gl == cnt_b - 1);
      rec.dense(mine, dP, -dp);
    }
#pragma unroll
    for (int k = 0; k < HS_G; ++k) {
      const double t = __shfl_sync(full, term, k, HS_G);
      if (go && k < cnt_b) E += t;  // in hull order, as a sequential sum
    }
    const int last_lane = go ? cnt_b - 1 : 0;
    const double cc = __shfl_sync(full, cdf, last_lane, HS_G), pp = __shfl_sync(full, pdf, last_lane, HS_G);
    if (go) {
      carry_cdf = cc;
      carry_pdf = pp;
      hbase += cnt_b;
      mine = empty_line();
      mine_x = INFINITY;
    }
  };

  while (__any_sync(full, active)) {
    Next best;
    best.L = empty_line();
    best.x = INFINITY;
    if (active) {
#pragma unroll
      for (int r = 0; r < HS_LINES / HS_G; ++r) consider_next(best, cur, cache[r]);
    }
#pragma unroll
    for (int o = HS_G / 2; o > 0; o >>= 1) {  // earliest intersection; ties -> earliest line in the sorted order
      Next oth;
      oth.L = shfl_line(best.L, o);
      oth.x = __shfl_xor_sync(full, best.x, o);
      merge_next(best, oth);
    }
    const bool last = best.L.idx < 0;
    const int slot = h - hbase;
    if (active && gl == slot) { mine = cur; mine_x = last ? INFINITY : best.x; }
    if (active) ++h;
    const bool batch_full = active && !last && slot == HS_G - 1;
    if (__any_sync(full, batch_full)) flush(batch_full, HS_G, false);
    if (active && last) {
      active = false;
      finished = true;
      fin_cnt = slot + 1;
    } else if (active) {
      cur = best.L;
    }
  }
  if (__any_sync(full, finished)) {
    flush(finished, fin_cnt, true);
    if (finished && gl == 0) {
      finish_set(lb, out, set, s, E, h);
      if (sc.stats) {
        atomicAdd((unsigned long long*)&sc.stats[1], (unsigned long long)cnt);
        atomicAdd((unsigned long long*)&sc.stats[3], (unsigned long long)h);
      }
    }
  }
}

int emax_hull(const LineBatch& lb, const EmaxScratch& sc, const EmaxOut& out, cudaStream_t st, double survivors_hint) {
  const long long sets = (long long)lb.C * lb.S;
  if (sets == 0) return DKG_OK;
  const int wpb = E_THREADS / 32;
  const char* e = getenv("DKG_HULL_SHORT");
  // hull records beyond the inline capacity go through the warp-wide recorder of hull_kernel only.
  // Measured at c4: ~9 survivors per set: 8 lanes per set take the hull stage from 0.227 to 0.13 ms; ~29 per set (most
  // sets hold 33 .. 64 lines with their seeds): 16 lanes per set.  The sets a short kernel leaves are queued for hull_kernel.
  static const double short_max = getenv("DKG_HULL_SHORT_MAX") != nullptr ? atof(getenv("DKG_HULL_SHORT_MAX")) : 48.0;
  static const double short8_max = getenv("DKG_HULL_SHORT8_MAX") != nullptr ? atof(getenv("DKG_HULL_SHORT8_MAX")) : 16.0;
  const int G = (survivors_hint < 0.0 || survivors_hint <= short8_max) ? 8 : 16;
  const bool use_short = !(e != nullptr && atoi(e) == 0) && (survivors_hint < 0.0 || survivors_hint <= short_max) &&
                         (out.hull_cap >= 4 * G || (out.hull_idx == nullptr && out.hull_x == nullptr));
  if (use_short) {
    const long long per_cta = (long long)wpb * (32 / G);
    if (G == 8) hull_short_kernel<8><<<(unsigned)((sets + per_cta - 1) / per_cta), E_THREADS, 0, st>>>(lb, sc, out);
    else hull_short_kernel<16><<<(unsigned)((sets + per_cta - 1) / per_cta), E_THREADS, 0, st>>>(lb, sc, out);
    DKG_LAUNCH_CHECK();
  }
  if (use_short && sc.long_sets != nullptr && sc.long_count != nullptr) {
    long long ctas = (sets + wpb - 1) / wpb;
    if (ctas > 148 * HULL_CTAS * 4) ctas = 148 * HULL_CTAS * 4;
    hull_kernel<<<(unsigned)ctas, E_THREADS, 0, st>>>(lb, sc, out, 0, 1);
  } else {
    hull_kernel<<<(unsigned)((sets + wpb - 1) / wpb), E_THREADS, 0, st>>>(lb, sc, out, use_short ? 4 * G : 0, 0);
  }
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// ------------------------------------------------------------------------------------------
// overflow kernel: one CTA per queued set, all lines, always terminates
// ------------------------------------------------------------------------------------------
constexpr int OVF_MAXV = 130;  // chain vertices
constexpr int OVF_ROUNDS = 6;
// one set per CTA and usually only a handful of queued sets: the time of this stage is the latency of the slowest
// set, so the CTAs are wide (the block-wide scans and marches scale with the thread count)
constexpr int OVF_THREADS = 512;
constexpr int OVF_CTAS_PER_SM = 2;

// index k of the chain vertex with the largest v_b[k] <= b  (requires b >= v_b[0])
__device__ __forceinline__ int chain_locate(const double* v_b, int nv, double b) {
  int lo = 0, hi = nv - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (v_b[mid] <= b) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(OVF_THREADS, OVF_CTAS_PER_SM)
overflow_kernel(LineBatch lb, EmaxScratch sc, EmaxOut out) {
  __shared__ double v_b[OVF_MAXV], v_a[OVF_MAXV], v_m[OVF_MAXV], v_slack[OVF_MAXV];
  __shared__ int v_idx[OVF_MAXV];
  __shared__ unsigned long long v_best[OVF_MAXV];
  __shared__ double n_b[OVF_MAXV], n_a[OVF_MAXV], g_b[OVF_MAXV], g_a[OVF_MAXV];
  __shared__ int n_i[OVF_MAXV], g_i[OVF_MAXV];
  __shared__ int s_nv, s_grew, s_list;
  __shared__ double r_a[OVF_THREADS / 32], r_b[OVF_THREADS / 32], r_x[OVF_THREADS / 32];
  __shared__ int r_idx[OVF_THREADS / 32], r_ref[OVF_THREADS / 32];
  __shared__ double c_a, c_b;
  __shared__ int c_idx, c_ref;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nq = *sc.ovf_count;
  for (int qi = blockIdx.x; qi < nq; qi += gridDim.x) {
    const size_t set = (size_t)sc.ovf_sets[qi];
    const int c = (int)(set / lb.S);
    const int j = (int)(set - (size_t)c * lb.S);
    const SetInfo s = set_info(lb, sc, c, j);
    const double w = s.w;
    Recorder rec{&out, set, lb.NL};
    __syncthreads();  // the previous queue entry is done with the shared state

    // ---- initial chain P, T, Q in the (b, a) plane (strictly increasing b) ----
    if (tid == 0) {
      const bool pos = !(w < 0.0);
      const Line P = gather_line(lb, c, j, w, pos ? s.iP : s.iQ);
      const Line Q = gather_line(lb, c, j, w, pos ? s.iQ : s.iP);
      const Line T = gather_line(lb, c, j, w, s.iT);
      int nv = 0;
      const unsigned long long f1 = sc.far[set * 2 + 0], f2 = sc.far[set * 2 + 1];
      const Line F1 = f1 ? gather_line(lb, c, j, w, (int)(f1 & 0xffffffffull)) : empty_line();
      const Line F2 = f2 ? gather_line(lb, c, j, w, (int)(f2 & 0xffffffffull)) : empty_line();
      const Line vs[5] = {P, F1, T, F2, Q};
      for (int v = 0; v < 5; ++v) {
        const Line& L = vs[v];
        if (L.idx < 0) continue;
        if ((v == 1 && !(L.b > P.b && L.b < T.b)) || (v == 3 && !(L.b > T.b && L.b < Q.b))) continue;
        if (nv > 0 && L.b <= v_b[nv - 1]) {  // same slope: keep the higher line
          if (L.b == v_b[nv - 1] && L.a > v_a[nv - 1]) { v_a[nv - 1] = L.a; v_idx[nv - 1] = L.idx; }
          continue;
        }
        v_b[nv] = L.b; v_a[nv] = L.a; v_idx[nv] = L.idx; ++nv;
      }
      // drop an inserted seed that is not above the chord of its neighbours (chain must be concave)
      for (int k = 1; k + 1 < nv;) {
        const double m = (v_a[k + 1] - v_a[k - 1]) / (v_b[k + 1] - v_b[k - 1]);
        const bool is_t = v_idx[k] == T.idx;
        if (!is_t && !(v_a[k] > fma(m, v_b[k] - v_b[k - 1], v_a[k - 1]))) {
          for (int q2 = k; q2 + 1 < nv; ++q2) { v_b[q2] = v_b[q2 + 1]; v_a[q2] = v_a[q2 + 1]; v_idx[q2] = v_idx[q2 + 1]; }
          --nv;
        } else ++k;
      }
      s_nv = nv;
    }
    __syncthreads();

    // Every round classifies all lines against the current chain, optimistically compacting the
    // survivors into this set's (global) list; if they fit, the list is final.  Otherwise the
    // farthest line above every chord becomes a new chain vertex and the pass is repeated.
    SurvEntry* list = sc.surv + set * SURV_CAP;
    bool small_enough = false;
    for (int round = 0; round <= OVF_ROUNDS; ++round) {
      const int nv = s_nv;
      for (int k = tid; k < nv; k += blockDim.x) {  // chord k joins vertices k and k+1
        v_best[k] = 0ull;
        if (k + 1 < nv) {
          const double m = (v_a[k + 1] - v_a[k]) / (v_b[k + 1] - v_b[k]);
          v_m[k] = m;
          v_slack[k] = EPS128 * (fabs(v_a[k]) + fabs(v_a[k + 1]) + fabs(m) * fmax(fabs(v_b[k]), fabs(v_b[k + 1])));
        }
      }
      if (tid == 0) s_list = 0;
      __syncthreads();
      for (int n0 = tid; n0 < lb.NL; n0 += 4 * blockDim.x) {
        // four lines per trip: all eight loads are issued before any of them is consumed
        double zq[4], aq[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int n = n0 + u * blockDim.x;
          const bool ok = n < lb.NL;
          zq[u] = ok ? line_slope(lb, c, n) : 0.0;
          aq[u] = ok ? line_intercept(lb, c, j, n) : -INFINITY;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int n = n0 + u * blockDim.x;
          if (n >= lb.NL) break;
          const double zn = zq[u];
          const Line L = make_line(lb, w, aq[u], zn, n);
          bool keep;
          if (L.b < v_b[0]) keep = true;  // outside the chain: cannot be dropped
          else {
            const int k = chain_locate(v_b, nv, L.b);
            if (L.b == v_b[k]) keep = L.a > v_a[k];  // same slope as a vertex: only if higher
            else if (k == nv - 1) keep = true;       // beyond the last vertex
            else {
              const double e = L.a - fma(v_m[k], L.b - v_b[k], v_a[k]);
              keep = e > -v_slack[k];
              if (e > 0.0) atomicMax(&v_best[k], pack_excess(e, n));
            }
          }
          if (keep) {
            const int pos = atomicAdd(&s_list, 1);
            if (pos < SURV_CAP) {
              SurvEntry e;
              e.a = L.a; e.z = zn; e.idx = n; e.pad = 0;
              list[pos] = e;
            }
          }
        }
      }
      __syncthreads();
      if (s_list <= SURV_CAP) { small_enough = true; break; }
      if (round == OVF_ROUNDS) break;
      // insert the farthest line of every chord: the candidates are gathered in parallel (two dependent global
      // loads each -- one thread doing that for up to 64 chords was most of a late round), thread 0 merges
      for (int k = tid; k + 1 < nv; k += blockDim.x) {
        g_i[k] = -1;
        if (v_best[k] != 0ull) {
          const Line L = gather_line(lb, c, j, w, (int)(v_best[k] & 0xffffffffull));
          if (L.b > v_b[k] && L.b < v_b[k + 1]) { g_b[k] = L.b; g_a[k] = L.a; g_i[k] = L.idx; }
        }
      }
      __syncthreads();
      if (tid == 0) {
        int m = 0;
        bool grew = false;
        for (int k = 0; k < nv; ++k) {
          n_b[m] = v_b[k]; n_a[m] = v_a[k]; n_i[m] = v_idx[k]; ++m;
          if (k + 1 < nv && g_i[k] >= 0 && m + (nv - k) < OVF_MAXV) {
            n_b[m] = g_b[k]; n_a[m] = g_a[k]; n_i[m] = g_i[k]; ++m;
            grew = true;
          }
        }
        for (int k = 0; k < m; ++k) { v_b[k] = n_b[k]; v_a[k] = n_a[k]; v_idx[k] = n_i[k]; }
        s_nv = m;
        s_grew = grew ? 1 : 0;
      }
      __syncthreads();
      if (!s_grew) break;  // no chord has anything above it, yet too many lines: all near-hull
    }

    // ---- exact block-wide march over `total` lines exposed by fetch(k): every thread scans its share for the next
    // vertex, warps and then the CTA reduce (earliest intersection, ties by the reference's sorted order) ----
    auto block_march = [&](int total, auto fetch, bool all_lines) {
      Line cur = empty_line();
      for (int k = tid; k < total; k += blockDim.x) {
        const Line L = fetch(k);
        if (cur.idx < 0 || sorted_before(L, cur)) cur = L;
      }
      for (int o = 16; o > 0; o >>= 1) {
        const Line oth = shfl_line(cur, o);
        if (oth.idx >= 0 && (cur.idx < 0 || sorted_before(oth, cur))) cur = oth;
      }
      if (lane == 0) { r_a[warp] = cur.a; r_b[warp] = cur.b; r_idx[warp] = cur.idx; r_ref[warp] = cur.ref; }
      __syncthreads();
      if (tid == 0) {
        Line best = empty_line();
        for (int k = 0; k < OVF_THREADS / 32; ++k) {
          Line L;
          L.a = r_a[k]; L.b = r_b[k]; L.idx = r_idx[k]; L.ref = r_ref[k];
          if (L.idx >= 0 && (best.idx < 0 || sorted_before(L, best))) best = L;
        }
        c_a = best.a; c_b = best.b; c_idx = best.idx; c_ref = best.ref;
      }
      __syncthreads();
      double E = 0.0, prev_cdf = 0.0, prev_pdf = 0.0;
      int h = 0;
      while (true) {
        cur.a = c_a; cur.b = c_b; cur.idx = c_idx; cur.ref = c_ref;
        Next best;
        best.L = empty_line();
        best.x = INFINITY;
        for (int k = tid; k < total; k += blockDim.x) consider_next(best, cur, fetch(k));
        for (int o = 16; o > 0; o >>= 1) {
          Next oth;
          oth.L = shfl_line(best.L, o);
          oth.x = __shfl_xor_sync(0xffffffffu, best.x, o);
          merge_next(best, oth);
        }
        if (lane == 0) {
          r_a[warp] = best.L.a; r_b[warp] = best.L.b; r_idx[warp] = best.L.idx;
          r_ref[warp] = best.L.ref; r_x[warp] = best.x;
        }
        __syncthreads();
        Next nx;
        nx.L = empty_line();
        nx.x = INFINITY;
        for (int k = 0; k < OVF_THREADS / 32; ++k) {
          Next o;
          o.L.a = r_a[k]; o.L.b = r_b[k]; o.L.idx = r_idx[k]; o.L.ref = r_ref[k]; o.x = r_x[k];
          merge_next(nx, o);
        }
        const bool last = nx.L.idx < 0;
        const double cdf = last ? 1.0 : std_normal_cdf(nx.x);
        const double pdf = last ? 0.0 : std_normal_pdf(nx.x);
        const double dP = cdf - prev_cdf, dp = pdf - prev_pdf;
        E += __dsub_rn(__dmul_rn(cur.a, dP), __dmul_rn(cur.b, dp));
        __syncthreads();  // everyone has consumed c_* and r_* of this step
        if (tid == 0) {
          rec.single(h, cur, dP, -dp, nx.x, last);
          c_a = nx.L.a; c_b = nx.L.b; c_idx = nx.L.idx; c_ref = nx.L.ref;
        }
        ++h;
        prev_cdf = cdf;
        prev_pdf = pdf;
        __syncthreads();
        if (last) break;
      }
      if (tid == 0) {
        finish_set(lb, out, set, s, E, h);
        if (sc.stats) {
          atomicAdd((unsigned long long*)&sc.stats[2], 1ull);
          if (all_lines) atomicAdd((unsigned long long*)&sc.stats[5], 1ull);
          atomicAdd((unsigned long long*)&sc.stats[3], (unsigned long long)h);
        }
      }
    };
    if (small_enough) {
      // ---- the list is complete: march over it plus the chain vertices.  (A single warp marching over a list of
      // up to SURV_CAP lines in global memory took ~100 us per set: 60 dependent trips per vertex.) ----
      const int nv = s_nv;
      __syncthreads();
      const int cnt = min(s_list, SURV_CAP);
      block_march(cnt + nv, [&](int k) -> Line {
        if (k < cnt) {
          const SurvEntry e = list[k];
          return make_line(lb, w, e.a, e.z, e.idx);
        }
        Line L;
        L.a = v_a[k - cnt]; L.b = v_b[k - cnt]; L.idx = v_idx[k - cnt];
        L.ref = ref_index(lb, L.idx);
        return L;
      }, false);
    } else {
      // ---- every line (e.g. all of them are hull vertices) ----
      block_march(lb.NL, [&](int n) -> Line { return gather_line(lb, c, j, w, n); }, true);
    }
  }
}

int emax_overflow(const LineBatch& lb, const EmaxScratch& sc, const EmaxOut& out, cudaStream_t st) {
  if (lb.C == 0) return DKG_OK;
  long long sets = (long long)lb.C * lb.S;
  int grid = (int)(sets < OVF_CTAS_PER_SM * 148 ? sets : OVF_CTAS_PER_SM * 148);
  overflow_kernel<<<grid, OVF_THREADS, 0, st>>>(lb, sc, out);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// ------------------------------------------------------------------------------------------
// finalize: kg[c] and the fused backward (envelope theorem); CTA per candidate
//   dKG/da_jn = (p_jn - [n == argmax a_j]) / S ; dKG/db_jn = q_jn / S on hull lines only.
// ------------------------------------------------------------------------------------------
// (templated on the input dimension: the coordinate loops below are written over MAX_D with a
// `k < d` guard; with d a compile-time constant the dead iterations disappear, which halves the
// code - and the instruction-cache misses - at d = 4)
template <int D>
__global__ void __launch_bounds__(E_THREADS, 3)
finalize_kernel(LineBatch lb, EmaxOut out, BackwardArgs bw, int rec_batch) {
  extern __shared__ __align__(16) unsigned char e_smem[];
  const int c = blockIdx.x;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int S = lb.S;

  // kg[c] = mean_j terms[c, j] (:338): the S loads are issued in parallel, the sum is taken in
  // index order by one thread (deterministic, same association as a sequential sum)
  double* s_term = reinterpret_cast<double*>(e_smem);  // [S]
  for (int j = threadIdx.x; j < S; j += blockDim.x) s_term[j] = out.terms[(size_t)c * S + j];
  __syncthreads();
  if (threadIdx.x == 0) {
    double acc = 0.0;
    for (int j = 0; j < S; ++j) acc += s_term[j];
    out.kg[c] = acc / (double)S;
  }
  if (bw.dX == nullptr) return;

  double* s_r = s_term + S;                         // [n_pad]  sum_n Gz[n] B[:, n]
  double* s_ga = s_r + bw.n_pad;                    // [S]      dKG/d a_own[j]
  double* s_sc = s_ga + S;                          // [0] Gsum, [1] GzOwn, [2..2+MAX_D) gkd, then Gm[m]
  double* s_red = s_sc + 2 + MAX_D + MAX_M;         // [nwarps * MAX_D]
  const int tgt = bw.target;
  constexpr int d = D;
  const double invS = 1.0 / (double)S;
  const double* zrow = lb.Z + (size_t)c * lb.ldz;

  // ---- gather the hull records of this candidate and merge duplicates: the same few lines are
  // hull vertices for most scalarisations, so the B^T row gathers and kernel-gradient evaluations
  // below run once per DISTINCT line (order = first occurrence, hence deterministic) ----
  const int RM = fin_rmax(S);
  int* s_ridx = reinterpret_cast<int*>(s_red + (E_THREADS / 32) * MAX_D);  // [RM] record line index
  int* s_uidx = s_ridx + RM;                               // [RM] distinct line index
  int* s_flag = s_uidx + RM;                               // [RM] first occurrence?
  int* s_roff = s_flag + RM;                               // [S + 1] record offsets per set
  double* s_rcz = reinterpret_cast<double*>(s_roff + ((S + 2) & ~1));  // [RM] w_j q / S
  double* s_ucz = s_rcz + RM;                              // [RM] merged coefficient
  int* s_hkey = reinterpret_cast<int*>(s_ucz + RM);        // [HN] line index held by a slot (-1: free)
  const int HB = fin_hash_bits(S), HN = 1 << HB;
  int* s_hfirst = s_hkey + HN;                             // [HN] first record of that line
  __shared__ int s_nrec, s_nuniq, s_hovf, s_broken;
  for (int j = threadIdx.x; j < S; j += blockDim.x)
    s_roff[j + 1] = out.hull_cnt[(size_t)c * S + j];  // counts first, scanned below
  __syncthreads();
  if (threadIdx.x == 0) {
    int off = 0;
    for (int j = 0; j < S; ++j) {
      const int h = s_roff[j + 1];
      s_roff[j] = off;
      off += h;
    }
    s_roff[S] = off;
    s_nrec = off;
    s_nuniq = 0;
    s_hovf = 0;
    s_broken = 0;
  }
  for (int h = threadIdx.x; h < HN; h += blockDim.x) { s_hkey[h] = -1; s_hfirst[h] = 0x7fffffff; }
  __syncthreads();
  const int nrec = s_nrec;
  // The records are merged in BATCHES of up to RM (what the shared-memory tables hold): a candidate with more
  // records than that (S = 256 on a noisy objective: thousands) used to fall back to loops in which every thread
  // walked every record of every scalarisation through the spill-chain reader for every training point -- ~3 ms
  // per candidate, 13 of the 21 ms of a c5 S = 256 launch.  r_t and the scalar terms simply accumulate over the
  // batches (one batch -- every S <= 64 case -- adds to zero: the same bits as before).
  bool merged = true;
  for (int t = threadIdx.x; t < bw.n_pad; t += blockDim.x) s_r[t] = 0.0;
  for (int j = threadIdx.x; j < S; j += blockDim.x) s_ga[j] = 0.0;  // p / S of the set's own-line record, filled while staging
  double gsum = 0.0, gzown = 0.0, gkd[MAX_D], gm[MAX_M];  // (warp 0)
#pragma unroll
  for (int k = 0; k < MAX_D; ++k) gkd[k] = 0.0;
#pragma unroll
  for (int m = 0; m < MAX_M; ++m) gm[m] = 0.0;
  double xs_t[MAX_D];
#pragma unroll
  for (int k = 0; k < MAX_D; ++k)
    xs_t[k] = k < d ? bw.X[(size_t)c * d + k] / bw.ls[tgt][k] : 0.0;
  auto slope_terms = [&](int idx, double cz) {
    gsum += cz * zrow[idx];
    if (idx == lb.NA) {
      gzown += cz;
    } else {
      double sq = 0.0;
#pragma unroll
      for (int q = 0; q < MAX_D; ++q)
        if (q < d) {
          double df = xs_t[q] - bw.xd_s[(size_t)idx * d + q];
          sq += df * df;
        }
      const double gc = stationary_grad_coeff(bw.kind[tgt], bw.outputscale[tgt], sq);
#pragma unroll
      for (int q = 0; q < MAX_D; ++q)
        if (q < d)
          gkd[q] += cz * gc * (xs_t[q] - bw.xd_s[(size_t)idx * d + q]) / bw.ls[tgt][q];
    }
  };
  const int RB = rec_batch > 0 && rec_batch < RM ? rec_batch : RM;  // (a smaller batch only as a test hook: DKG_FIN_BATCH)
  for (int b0 = 0; b0 < nrec && merged; b0 += RB) {
    const int nb = min(RB, nrec - b0);  // records b0 .. b0 + nb of the candidate (set-major order)
    if (b0 > 0) {
      __syncthreads();  // the previous batch's tables are done with
      for (int h = threadIdx.x; h < HN; h += blockDim.x) { s_hkey[h] = -1; s_hfirst[h] = 0x7fffffff; }
      if (threadIdx.x == 0) s_nuniq = 0;
      __syncthreads();
    }
    for (int j = warp; j < S; j += nwarps) {
      if (s_roff[j + 1] <= b0 || s_roff[j] >= b0 + nb) continue;  // no record of this set in the batch
      const size_t set = (size_t)c * S + j;
      const double wj = bw.W[j * bw.M + tgt];
      const int h = s_roff[j + 1] - s_roff[j];
      HullReader rd(out, set);
      for (int k = lane; k < h; k += 32) {
        const int e = s_roff[j] + k - b0;
        if (e < 0 || e >= nb) continue;
        if (!rd.seek(k)) { s_broken = 1; s_ridx[e] = lb.NA; s_rcz[e] = 0.0; continue; }
        const int idx = rd.idx();
        s_ridx[e] = idx;
        s_rcz[e] = wj * rd.q() * invS;
        if (idx == lb.NA) s_ga[j] = rd.p() * invS;  // (a line is a hull vertex of its set at most once)
      }
    }
    __syncthreads();
    // first occurrence of every line through a small hash table: slot claim by compare-and-swap,
    // first record by atomicMin (order independent, hence deterministic); a quadratic scan here cost
    // 17 % of the kernel at S = 16 and made S = 256 unusable
    for (int e = threadIdx.x; e < nb; e += blockDim.x) {
      const int idx = s_ridx[e];
      unsigned h = ((unsigned)idx * 2654435761u) >> (32 - HB);
      int probes = 0;
      for (;; h = (h + 1) & (HN - 1)) {
        const int prev = atomicCAS(&s_hkey[h], -1, idx);
        if (prev == -1 || prev == idx) break;
        if (++probes >= HN) { s_hovf = 1; break; }
      }
      atomicMin(&s_hfirst[h], e);
      s_flag[e] = (int)h;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nb; e += blockDim.x) s_flag[e] = s_hfirst[s_flag[e]] == e ? 1 : 0;
    __syncthreads();
    if (s_hovf) { merged = false; break; }  // more distinct lines than slots (cannot happen while HN > RM): unmerged loops
    for (int e = threadIdx.x; e < nb; e += blockDim.x) {
      if (!s_flag[e]) continue;
      const int idx = s_ridx[e];
      int rank = 0;  // distinct lines that first occur before e
      for (int f = 0; f < e; ++f) rank += s_flag[f];
      double acc = 0.0;
      for (int f = e; f < nb; ++f)
        if (s_ridx[f] == idx) acc += s_rcz[f];
      s_uidx[rank] = idx;
      s_ucz[rank] = acc;
      atomicAdd(&s_nuniq, 1);
    }
    __syncthreads();
    const int nuniq = s_nuniq;
    // r_t over this thread's training points; every thread walks the same list in order
    for (int t = threadIdx.x; t < bw.n_pad; t += blockDim.x) {
      double acc = 0.0;
      for (int u = 0; u < nuniq; ++u) {
        const int idx = s_uidx[u];
        if (idx < lb.NA) acc += s_ucz[u] * bw.BT[(size_t)idx * bw.ldbt + t];
      }
      s_r[t] += acc;
    }
    if (warp == 0)
      for (int u = lane; u < nuniq; u += 32) slope_terms(s_uidx[u], s_ucz[u]);
  }
  if (!merged) {
    __syncthreads();
    for (int t = threadIdx.x; t < bw.n_pad; t += blockDim.x) {
      double acc = 0.0;
      for (int j = 0; j < S; ++j) {
        const size_t set = (size_t)c * S + j;
        const int h = out.hull_cnt[set];
        const double wj = bw.W[j * bw.M + tgt];
        HullReader rd(out, set);
        for (int k = 0; k < h; ++k) {
          if (!rd.seek(k)) { s_broken = 1; break; }
          const int idx = rd.idx();
          if (idx < lb.NA) {
            const double cz = wj * rd.q() * invS;
            acc += cz * bw.BT[(size_t)idx * bw.ldbt + t];
          }
        }
      }
      s_r[t] = acc;
    }
    gsum = gzown = 0.0;
#pragma unroll
    for (int k = 0; k < MAX_D; ++k) gkd[k] = 0.0;
  }
  // scalars: warp 0
  if (warp == 0) {
    for (int j = lane; j < S; j += 32) {  // lane l owns scalarisations l, l+32, ...
      const size_t set = (size_t)c * S + j;
      double ga = out.amax_is_own[set] ? -invS : 0.0;
      if (merged) {
        ga += s_ga[j];  // (staged with the records: no dependent walk over the set's records on the critical path)
      } else {
        const int h = out.hull_cnt[set];
        const double wj = bw.W[j * bw.M + tgt];
        HullReader rd(out, set);
        for (int k = 0; k < h; ++k) {
          if (!rd.seek(k)) { s_broken = 1; break; }
          const int idx = rd.idx();
          if (idx == lb.NA) ga += rd.p() * invS;
          slope_terms(idx, wj * rd.q() * invS);
        }
      }
      s_ga[j] = ga;
#pragma unroll
      for (int m = 0; m < MAX_M; ++m)
        if (m < bw.M) gm[m] += ga * bw.W[j * bw.M + m];
    }
    gsum = warp_sum(gsum);
    gzown = warp_sum(gzown);
#pragma unroll
    for (int k = 0; k < MAX_D; ++k) gkd[k] = warp_sum(gkd[k]);
#pragma unroll
    for (int m = 0; m < MAX_M; ++m) gm[m] = warp_sum(gm[m]);
    if (lane == 0) {
      s_sc[0] = gsum;
      s_sc[1] = gzown;
      for (int k = 0; k < MAX_D; ++k) s_sc[2 + k] = gkd[k];
      for (int m = 0; m < MAX_M; ++m) s_sc[2 + MAX_D + m] = gm[m];
    }
  }
  __syncthreads();

  const double var = bw.var[c], sd = bw.sd[c];
  const double s2 = bw.y_std[tgt] * bw.y_std[tgt];
  const double cT = -2.0 * s2 * (s_sc[1] / sd - s_sc[0] / (2.0 * var));
  const double cr = -s2 / sd;
  double grad[MAX_D];
#pragma unroll
  for (int k = 0; k < MAX_D; ++k) grad[k] = 0.0;
  for (int m = 0; m < bw.M; ++m) {
    const double cm = s_sc[2 + MAX_D + m] * bw.y_std[m];
    if (m != tgt && cm == 0.0) continue;
    double xm[MAX_D];
#pragma unroll
    for (int k = 0; k < MAX_D; ++k) xm[k] = k < d ? bw.X[(size_t)c * d + k] / bw.ls[m][k] : 0.0;
    for (int t = threadIdx.x; t < bw.ntr[m]; t += blockDim.x) {
      double u = cm * bw.alpha[m][t];
      if (m == tgt) u += cr * s_r[t] + cT * bw.T[(size_t)c * bw.ldk + t];
      double sq = 0.0;
#pragma unroll
      for (int k = 0; k < MAX_D; ++k)
        if (k < d) {
          double df = xm[k] - bw.xs[m][(size_t)t * d + k];
          sq += df * df;
        }
      const double gc = u * stationary_grad_coeff(bw.kind[m], bw.outputscale[m], sq);
#pragma unroll
      for (int k = 0; k < MAX_D; ++k)
        if (k < d) grad[k] += gc * (xm[k] - bw.xs[m][(size_t)t * d + k]) / bw.ls[m][k];
    }
  }
#pragma unroll
  for (int k = 0; k < MAX_D; ++k) grad[k] = warp_sum(grad[k]);
  if (lane == 0)
    for (int k = 0; k < MAX_D; ++k) s_red[warp * MAX_D + k] = grad[k];
  __syncthreads();
  if (threadIdx.x < d) {
    double acc = 0.0;
    for (int wv = 0; wv < nwarps; ++wv) acc += s_red[wv * MAX_D + threadIdx.x];
    acc += (s2 / sd) * s_sc[2 + threadIdx.x];
    // hull records lost to an exhausted spill pool: fail loudly (NaN) instead of a partial gradient
    bw.dX[(size_t)c * d + threadIdx.x] = s_broken ? __longlong_as_double(0x7ff8000000000000ll) : acc;
  }
}

int emax_finalize(const LineBatch& lb, const EmaxOut& out, const BackwardArgs& bw, cudaStream_t st) {
  if (lb.C == 0 || out.kg == nullptr) return DKG_OK;
  size_t smem = sizeof(double) * lb.S;
  if (bw.dX != nullptr)
    smem = sizeof(double) * ((size_t)bw.n_pad + 2 * lb.S + 2 + MAX_D + MAX_M + (E_THREADS / 32) * MAX_D) +
           sizeof(int) * (3 * fin_rmax(lb.S) + ((lb.S + 2) & ~1) + (2 << fin_hash_bits(lb.S))) + sizeof(double) * 2 * fin_rmax(lb.S);
  const char* fte = getenv("DKG_FIN_THREADS");
  // 128-thread CTAs (6 per SM) hide the short barrier-separated phases best; with many scalarisations
  // the hull records exceed the merge capacity and the per-record loops want the wider CTA
  const int fin_default = lb.S > 32 ? 256 : 128;
  const char* fbe = getenv("DKG_FIN_BATCH");  // test hook: hull records merged per batch (default: the table capacity)
  const int rec_batch = fbe != nullptr ? atoi(fbe) : 0;
  const int fin_threads = fte != nullptr && (atoi(fte) == 256 || atoi(fte) == 128 || atoi(fte) == 64) ? atoi(fte) : fin_default;
#define DKG_FINALIZE(DD)                                                                                       \
  do {                                                                                                         \
    if (smem > 47 * 1024)                                                                                      \
      DKG_CUDA_OK(cudaFuncSetAttribute(finalize_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    finalize_kernel<DD><<<lb.C, fin_threads, smem, st>>>(lb, out, bw, rec_batch);                              \
  } while (0)
  switch (bw.dX != nullptr ? bw.d : 1) {
    case 1: DKG_FINALIZE(1); break;
    case 2: DKG_FINALIZE(2); break;
    case 3: DKG_FINALIZE(3); break;
    case 4: DKG_FINALIZE(4); break;
    case 5: DKG_FINALIZE(5); break;
    case 6: DKG_FINALIZE(6); break;
    case 7: DKG_FINALIZE(7); break;
    default: DKG_FINALIZE(8); break;
  }
#undef DKG_FINALIZE
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// ------------------------------------------------------------------------------------------
// small discretisations: one CTA per candidate (see SmallArgs in dkg_emax.cuh)
// ------------------------------------------------------------------------------------------
constexpr int SM_THREADS = 256;

__device__ __forceinline__ double block_sum(double v, double* s_red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) s_red[w] = v;
  __syncthreads();
  double tot = 0.0;
  for (int k = 0; k < SM_THREADS / 32; ++k) tot += s_red[k];  // fixed order: deterministic
  return tot;
}

template <int D>
__global__ void __launch_bounds__(SM_THREADS)
small_kg_kernel(SmallArgs a, LineBatch lb, EmaxScratch sc, EmaxOut out) {
  extern __shared__ __align__(16) unsigned char e_smem[];
  const int c = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = lb.S, N = lb.NA, tgt = a.target;
  const int n = a.ntr[tgt];
  const int n_al = (n + 1) & ~1;
  double* zs = reinterpret_cast<double*>(e_smem);  // [NL] slope row
  double* kx = zs + ((lb.NL + 1) & ~1);            // [n]  k(x, X_train)
  double* t0 = kx + n_al;                          // [n]
  double* rr = t0 + n_al;                          // [n]
  double* tt = rr + n_al;                          // [n]  T
  double* aw = tt + n_al + (size_t)warp * ((lb.NL + 1) & ~1);  // [NL] this warp's intercept row (march stage)
  __shared__ double s_red[SM_THREADS / 32], s_mu[MAX_M], s_scal[4];
  __shared__ double s_zmin[SM_THREADS / 32], s_zmax[SM_THREADS / 32];

  // ---- kernel rows and posterior means of every objective (xprep_kernel) ----
  for (int m = 0; m < a.M; ++m) {
    double xm[D];
#pragma unroll
    for (int k = 0; k < D; ++k) xm[k] = a.X[(size_t)c * D + k] / a.ls[m][k];
    double acc = 0.0;
    for (int t = tid; t < a.ntr[m]; t += SM_THREADS) {
      double sq = 0.0;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const double df = xm[k] - a.xs[m][(size_t)t * D + k];
        sq += df * df;
      }
      const double kv = stationary_from_sq(a.kind[m], a.outputscale[m], sq);
      if (m == tgt) kx[t] = kv;
      acc += kv * a.alpha[m][t];
    }
    const double tot = block_sum(acc, s_red);
    if (tid == 0) s_mu[m] = (a.mean_const[m] + tot) * a.y_std[m] + a.y_mean[m];
  }
  __syncthreads();
  for (int j = tid; j < S; j += SM_THREADS) {  // own-line intercepts (discretekg.py:320, row 0)
    double v = __dmul_rn(a.W[j * a.M + 0], s_mu[0]);
    for (int m = 1; m < a.M; ++m) v = __dadd_rn(v, __dmul_rn(a.W[j * a.M + m], s_mu[m]));
    a.a_new[(size_t)c * S + j] = v;
  }
  // ---- T = K^-1 k_x with one refinement step (solve_T): T0 = Kinv kx; R = kx - K T0; T = T0 + Kinv R ----
  // (thread t walks column t of the symmetric matrices: coalesced across the threads.  A warp-per-row
  // variant with shuffle reductions measured SLOWER at these sizes: 48 vs 36 us per forward at n = 60)
  for (int t = tid; t < n; t += SM_THREADS) {
    double acc = 0.0;
    for (int s2 = 0; s2 < n; ++s2) acc = fma(kx[s2], a.Kinv[(size_t)s2 * a.ldk + t], acc);
    t0[t] = acc;
  }
  __syncthreads();
  if (a.refine) {
    for (int t = tid; t < n; t += SM_THREADS) {
      double acc = 0.0;
      for (int s2 = 0; s2 < n; ++s2) acc = fma(t0[s2], a.Kmat[(size_t)s2 * a.ldk + t], acc);
      rr[t] = kx[t] - acc;
    }
    __syncthreads();
    for (int t = tid; t < n; t += SM_THREADS) {
      double acc = 0.0;
      for (int s2 = 0; s2 < n; ++s2) acc = fma(rr[s2], a.Kinv[(size_t)s2 * a.ldk + t], acc);
      tt[t] = t0[t] + acc;
    }
  } else {
    for (int t = tid; t < n; t += SM_THREADS) tt[t] = t0[t];
  }
  __syncthreads();
  for (int t = tid; t < n; t += SM_THREADS) a.T[(size_t)c * a.ldk + t] = tt[t];
  // ---- noisy predictive variance and the own line's slope (var_kernel) ----
  {
    double acc = 0.0;
    for (int t = tid; t < n; t += SM_THREADS) acc += kx[t] * tt[t];
    const double dot = block_sum(acc, s_red);
    if (tid == 0) {
      const double ystd2 = a.y_std[tgt] * a.y_std[tgt];
      const double kxx = stationary_from_sq(a.kind[tgt], a.outputscale[tgt], 0.0);
      const double var_lat = kxx - dot;
      const double v = (var_lat + a.noise) * ystd2;
      const double sdv = sqrt(v);
      a.var[c] = v;
      a.sd[c] = sdv;
      s_scal[0] = ystd2 / sdv;
      s_scal[1] = (var_lat * ystd2) / sdv;  // znew_coefficients[0] (:313)
    }
    __syncthreads();
  }
  // ---- covariance row over the predictive sd: z_n = (k(x, x_n) - T . k(X_train, x_n)) ystd^2 / sd ----
  {
    const double rsd = s_scal[0];
    double xt[D];
#pragma unroll
    for (int k = 0; k < D; ++k) xt[k] = a.X[(size_t)c * D + k] / a.ls[tgt][k];
    double* zg = const_cast<double*>(lb.Z) + (size_t)c * lb.ldz;
    double vmin = INFINITY, vmax = -INFINITY;
    for (int nn = tid; nn < N; nn += SM_THREADS) {
      double sq = 0.0;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const double df = xt[k] - a.xd_s[(size_t)nn * D + k];
        sq = fma(df, df, sq);
      }
      double dot = 0.0;
      for (int t = 0; t < n; ++t) dot = fma(tt[t], a.Kxd[(size_t)t * a.ldx + nn], dot);
      const double z = (stationary_from_sq(a.kind[tgt], a.outputscale[tgt], sq) - dot) * rsd;
      zs[nn] = z;
      zg[nn] = z;
      vmin = fmin(vmin, z);
      vmax = fmax(vmax, z);
    }
    if (tid == 0) {
      zs[N] = s_scal[1];
      zg[N] = s_scal[1];
      vmin = fmin(vmin, s_scal[1]);
      vmax = fmax(vmax, s_scal[1]);
    }
    vmin = warp_min(vmin);
    vmax = warp_max(vmax);
    if (lane == 0) { s_zmin[warp] = vmin; s_zmax[warp] = vmax; }
    __syncthreads();
  }
  double zmin = s_zmin[0], zmax = s_zmax[0];
  for (int k = 1; k < SM_THREADS / 32; ++k) { zmin = fmin(zmin, s_zmin[k]); zmax = fmax(zmax, s_zmax[k]); }

  // ---- one warp per scalarisation: the reference's march over ALL lines (no filter needed) ----
  for (int j = warp; j < S; j += SM_THREADS / 32) {
    const size_t set = (size_t)c * S + j;
    const double w = lb.wt ? lb.wt[j] : 1.0;
    SetInfo s;
    s.w = w;
    s.zmin = zmin; s.zmax = zmax; s.iP = s.iQ = 0;
    s.amax = lb.Amax[j];
    s.iT = lb.Aarg[j];
    s.own_is_max = 0;
    const double ao = a.a_new[(size_t)c * S + j];
    if (ao >= s.amax) { s.amax = ao; s.iT = N; s.own_is_max = 1; }
    s.shortcut = fabs(__dmul_rn(w, zmin)) < SHORTCUT_TOL && fabs(__dmul_rn(w, zmax)) < SHORTCUT_TOL;
    if (out.amax_is_own != nullptr && lane == 0) out.amax_is_own[set] = s.own_is_max;
    Recorder rec{&out, set, lb.NL};
    if (s.shortcut) {
      if (lane == 0) {
        Line L;
        L.idx = s.iT; L.a = s.amax; L.b = 0; L.ref = 0;
        rec.single(0, L, 1.0, 0.0, 0.0, true);
        finish_set(lb, out, set, s, s.amax, 1);
        if (sc.stats) atomicAdd((unsigned long long*)&sc.stats[4], 1ull);
      }
      continue;
    }
    // the set's intercepts go to shared memory first: the march re-reads every line once per hull
    // vertex, and from L2 each of those reads is a dependent ~700-cycle access
    const double* arow = lb.A + (size_t)j * lb.a_sj;
    __syncwarp();
    for (int k = lane; k < N; k += 32) aw[k] = arow[k];
    if (lane == 0) aw[N] = ao;
    __syncwarp();
    auto fetch = [&](int k) -> Line {
      return make_line(lb, w, aw[k], zs[k], k);
    };
    const HullResult r = warp_march(lb.NL, fetch, rec);
    if (lane == 0) {
      finish_set(lb, out, set, s, r.E, r.h);
      if (sc.stats) atomicAdd((unsigned long long*)&sc.stats[3], (unsigned long long)r.h);
    }
  }
}

int emax_small_forward(const SmallArgs& a, const LineBatch& lb, const EmaxScratch& sc, const EmaxOut& out,
                       cudaStream_t st) {
  if (a.C == 0) return DKG_OK;
  const int n_al = (a.ntr[a.target] + 1) & ~1;
  const size_t smem = sizeof(double) * ((size_t)((lb.NL + 1) & ~1) * (1 + SM_THREADS / 32) + 4 * (size_t)n_al);
#define DKG_SMALL(DD)                                                                                              \
  do {                                                                                                             \
    if (smem > 47 * 1024)                                                                                          \
      DKG_CUDA_OK(cudaFuncSetAttribute(small_kg_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    small_kg_kernel<DD><<<a.C, SM_THREADS, smem, st>>>(a, lb, sc, out);                                            \
  } while (0)
  switch (a.d) {
    case 1: DKG_SMALL(1); break;
    case 2: DKG_SMALL(2); break;
    case 3: DKG_SMALL(3); break;
    case 4: DKG_SMALL(4); break;
    case 5: DKG_SMALL(5); break;
    case 6: DKG_SMALL(6); break;
    case 7: DKG_SMALL(7); break;
    default: DKG_SMALL(8); break;
  }
#undef DKG_SMALL
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// ------------------------------------------------------------------------------------------
// stand-alone expectation of a piecewise-linear function of Z ~ N(0, 1) with caller-given break
// points (discretekg.py:415-452): E = sum_k a_k (Phi_{k+1} - Phi_k) - b_k (phi_{k+1} - phi_k) with
// z_0 = -inf, z_H = +inf, summed in piece order; and its gradient
//   dE/da_k = Phi_{k+1} - Phi_k,  dE/db_k = -(phi_{k+1} - phi_k),
//   dE/dz_k = phi(z_k) ((a_{k-1} - a_k) + z_k (b_{k-1} - b_k))      (phi'(z) = -z phi(z)).
// One warp per function: the lanes evaluate Phi / phi of 32 break points at a time, lane 0 adds the
// terms in order (same association as a sequential sum).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(E_THREADS)
piecewise_kernel(const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ z,
                 int P, int H, double* __restrict__ e, double* __restrict__ de_da, double* __restrict__ de_db,
                 double* __restrict__ de_dz) {
  const int lane = threadIdx.x & 31;
  const long long p = (long long)blockIdx.x * (E_THREADS / 32) + (threadIdx.x >> 5);
  if (p >= P) return;
  const double* ap = a + (size_t)p * H;
  const double* bp = b + (size_t)p * H;
  const double* zp = z + (size_t)p * (H - 1);
  double E = 0.0, carry_cdf = 0.0, carry_pdf = 0.0;
  for (int k0 = 0; k0 < H; k0 += 32) {
    const int k = k0 + lane;
    const bool act = k < H;
    const double zr = (act && k < H - 1) ? zp[k] : INFINITY;  // right break point of piece k
    const double cdf = act ? std_normal_cdf(zr) : 0.0;
    const double pdf = act ? std_normal_pdf(zr) : 0.0;
    double lcdf = __shfl_up_sync(0xffffffffu, cdf, 1);
    double lpdf = __shfl_up_sync(0xffffffffu, pdf, 1);
    if (lane == 0) { lcdf = carry_cdf; lpdf = carry_pdf; }
    const double dP = cdf - lcdf, dp = pdf - lpdf;
    const double ak = act ? ap[k] : 0.0, bk = act ? bp[k] : 0.0;
    const double term = act ? __dsub_rn(__dmul_rn(ak, dP), __dmul_rn(bk, dp)) : 0.0;
    if (act) {
      if (de_da) de_da[(size_t)p * H + k] = dP;
      if (de_db) de_db[(size_t)p * H + k] = -dp;
      if (de_dz && k < H - 1) {
        const double an = ap[k + 1], bn = bp[k + 1];
        de_dz[(size_t)p * (H - 1) + k] = pdf * ((ak - an) + zr * (bk - bn));
      }
    }
    const int cnt = min(32, H - k0);
    for (int q = 0; q < cnt; ++q) E += __shfl_sync(0xffffffffu, term, q);
    carry_cdf = __shfl_sync(0xffffffffu, cdf, cnt - 1);
    carry_pdf = __shfl_sync(0xffffffffu, pdf, cnt - 1);
  }
  if (lane == 0) e[p] = E;
}

int piecewise_expectation(const double* a, const double* b, const double* z, int P, int H, double* e,
                          double* de_da, double* de_db, double* de_dz, cudaStream_t st) {
  const int wpb = E_THREADS / 32;
  piecewise_kernel<<<(unsigned)(((long long)P + wpb - 1) / wpb), E_THREADS, 0, st>>>(a, b, z, P, H, e, de_da, de_db, de_dz);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

}  // namespace dkg
