// Expected maximum over fantasy outcomes (discretekg.py:329-336 / :341-452) for
// [candidates x scalarisations] line sets of |X_disc|+1 lines each, plus the fused backward.
//
// The reference finds each upper envelope with a Python Jarvis march over ALL lines.  Here:
//   1. zstat  : min / max slope of each candidate's slope row (shared by all scalarisations,
//               because slopes are w_j * z with one z row per candidate, discretekg.py:321).
//   2. filter : one streaming, coalesced pass over the slope rows and the (L2-resident)
//               intercept table.  A line whose dual point (slope, intercept) lies on or below the
//               chain P -> T -> Q (P/Q = extreme-slope lines, T = max-intercept line; all three
//               are lines of the set) is inside the convex hull of the set, hence can never be a
//               strict vertex of the upper envelope and is dropped.  On GP-shaped inputs this
//               leaves ~1% of the lines.
//   3. hull   : one warp per (candidate, scalarisation) runs the reference's march EXACTLY
//               (same ordering rule, same strict-slope filter, same division, same tie-breaks)
//               on the survivors, accumulates the closed-form expectation segment by segment
//               and records dE/da, dE/db for the backward.
//   4. backward (same kernel, CTA per candidate): envelope-theorem gradient, sparse over the
//               recorded hull vertices.
#include "dkg_emax.cuh"

namespace dkg {

constexpr int E_THREADS = 256;
constexpr double SHORTCUT_TOL = 1e-9;  // discretekg.py:363

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

// one CTA per row: min / max (with first index) of the slope row; optionally the max intercept
__global__ void __launch_bounds__(E_THREADS)
zstat_kernel(LineBatch lb, double* __restrict__ zst, int* __restrict__ zarg,
             double* __restrict__ amax_out, int* __restrict__ aarg_out) {
  __shared__ double s_v[E_THREADS / 32];
  __shared__ int s_i[E_THREADS / 32];
  const int c = blockIdx.x;
  const double* z = lb.Z + (size_t)c * lb.ldz;
  double vmin = INFINITY, vmax = -INFINITY;
  int imin = 0x7fffffff, imax = 0x7fffffff;
  for (int n = threadIdx.x; n < lb.NL; n += blockDim.x) {
    double v = z[n];
    if (v < vmin) { vmin = v; imin = n; }
    if (v > vmax) { vmax = v; imax = n; }
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

int emax_zstat(const LineBatch& lb, const EmaxScratch& sc, double* amax_out, int* aarg_out,
               cudaStream_t st) {
  if (lb.C == 0) return DKG_OK;
  zstat_kernel<<<lb.C, E_THREADS, 0, st>>>(lb, sc.zst, sc.zarg, amax_out, aarg_out);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// ------------------------------------------------------------------------------------------
// chord filter
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double line_intercept(const LineBatch& lb, int c, int j, int n) {
  return n < lb.NA ? lb.A[(size_t)c * lb.a_sc + (size_t)j * lb.a_sj + n]
                   : lb.a_own[(size_t)c * lb.S + j];
}

// Chain parameters for set (c, j): a line (z, a) survives iff
//   a > min(par.x + par.y * z, par.z + par.w * z)
// (left chord P->T and right chord T->Q in the dual plane; the chain is concave because T has
// the maximum intercept).  A small slack keeps everything within rounding of the chain.
__device__ double4 chain_params(const LineBatch& lb, const EmaxScratch& sc, int c, int j) {
  const double inf = INFINITY;
  const double w = lb.wt ? lb.wt[j] : 1.0;
  const double zmin = sc.zst[c * 2 + 0], zmax = sc.zst[c * 2 + 1];
  if (fabs(__dmul_rn(w, zmin)) < SHORTCUT_TOL && fabs(__dmul_rn(w, zmax)) < SHORTCUT_TOL)
    return make_double4(inf, 0.0, inf, 0.0);  // reference shortcut: the hull is argmax a only
  const double sgn = w < 0.0 ? -1.0 : 1.0;   // effective slope coordinate z' = sgn * z
  const int iP = sgn > 0 ? sc.zarg[c * 2 + 0] : sc.zarg[c * 2 + 1];
  const int iQ = sgn > 0 ? sc.zarg[c * 2 + 1] : sc.zarg[c * 2 + 0];
  const double zP = sgn > 0 ? zmin : -zmax;
  const double zQ = sgn > 0 ? zmax : -zmin;
  double aT = lb.Amax[(size_t)c * lb.am_sc + j];
  int iT = lb.Aarg[(size_t)c * lb.am_sc + j];
  if (lb.a_own != nullptr) {
    double ao = lb.a_own[(size_t)c * lb.S + j];
    if (ao > aT) { aT = ao; iT = lb.NA; }
  }
  const double zT = sgn * lb.Z[(size_t)c * lb.ldz + iT];
  const double aP = line_intercept(lb, c, j, iP);
  const double aQ = line_intercept(lb, c, j, iQ);
  // slack = 128 eps of every magnitude that enters  c + m * z  (keeps all lines that are within
  // rounding of the chain; extra survivors only cost time, never accuracy)
  const double eps128 = 2.84217094304040074e-14;
  double c1 = inf, m1 = 0.0, c2 = inf, m2 = 0.0;
  if (zT > zP) {
    m1 = (aT - aP) / (zT - zP);
    const double slack = eps128 * (fabs(aT) + fabs(aP) + fabs(m1) * fmax(fabs(zP), fabs(zT)));
    c1 = aP - m1 * zP - slack;
  }
  if (zQ > zT) {
    m2 = (aQ - aT) / (zQ - zT);
    const double slack = eps128 * (fabs(aT) + fabs(aQ) + fabs(m2) * fmax(fabs(zQ), fabs(zT)));
    c2 = aT - m2 * zT - slack;
  }
  return make_double4(c1, m1 * sgn, c2, m2 * sgn);
}

template <int G, bool SHARED_A>
__global__ void __launch_bounds__(E_THREADS)
filter_kernel(LineBatch lb, EmaxScratch sc, int slice_len) {
  extern __shared__ __align__(16) unsigned char e_smem[];
  double4* s_par = reinterpret_cast<double4*>(e_smem);  // [G][S]
  const int c0 = blockIdx.y * G;
  const int S = lb.S;
  for (int e = threadIdx.x; e < G * S; e += blockDim.x) {
    int g = e / S, j = e - g * S;
    int c = c0 + g;
    s_par[e] = (c < lb.C) ? chain_params(lb, sc, c, j) : make_double4(INFINITY, 0.0, INFINITY, 0.0);
  }
  __syncthreads();

  const int lo = blockIdx.x * slice_len;
  const int hi = min(lo + slice_len, lb.NA);
  for (int n = lo + threadIdx.x; n < hi; n += blockDim.x) {
    double z[G];
#pragma unroll
    for (int g = 0; g < G; ++g)
      z[g] = (c0 + g < lb.C) ? lb.Z[(size_t)(c0 + g) * lb.ldz + n] : 0.0;
    for (int j = 0; j < S; ++j) {
      double a_sh = 0.0;
      if (SHARED_A) a_sh = lb.A[(size_t)j * lb.a_sj + n];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        double a = a_sh;
        if (!SHARED_A)
          a = (c0 + g < lb.C) ? lb.A[(size_t)(c0 + g) * lb.a_sc + (size_t)j * lb.a_sj + n]
                              : -INFINITY;
        const double4 p = s_par[g * S + j];
        const double thr = fmin(fma(p.y, z[g], p.x), fma(p.w, z[g], p.z));
        if (a > thr) {
          const size_t set = (size_t)(c0 + g) * S + j;
          int pos = atomicAdd(&sc.surv_cnt[set], 1);
          if (pos < SURV_CAP) sc.surv_idx[set * SURV_CAP + pos] = n;
        }
      }
    }
  }
}

int emax_filter(const LineBatch& lb, const EmaxScratch& sc, cudaStream_t st) {
  if (lb.C == 0 || lb.NA == 0) return DKG_OK;
  constexpr int G = 4;
  // slices sized so the grid has a few waves of CTAs on 148 SMs
  int rows = ceil_div(lb.C, G);
  int slice = 2048;
  while (slice > 256 && (long long)rows * ceil_div(lb.NA, slice) < 2 * 148) slice >>= 1;
  dim3 grid(ceil_div(lb.NA, slice), rows);
  size_t smem = (size_t)G * lb.S * sizeof(double4);
  if (lb.a_sc == 0)
    filter_kernel<G, true><<<grid, E_THREADS, smem, st>>>(lb, sc, slice);
  else
    filter_kernel<G, false><<<grid, E_THREADS, smem, st>>>(lb, sc, slice);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// ------------------------------------------------------------------------------------------
// exact warp-level march (restates discretekg.py:370-410 on the surviving lines)
// ------------------------------------------------------------------------------------------
struct Line {
  double a, b;
  int idx;  // internal index (own line == NA)
  int ref;  // position in the reference's ordering of the inputs (own line first)
};

// ordering of the reference's sorted array: slope ascending, then intercept descending; among
// identical lines the input position decides (the reference's first sort is not stable there).
__device__ __forceinline__ bool sorted_before(const Line& p, const Line& q) {
  if (p.b != q.b) return p.b < q.b;
  if (p.a != q.a) return p.a > q.a;
  return p.ref < q.ref;
}

struct SetView {
  const LineBatch* lb;
  const int* surv;  // survivor list of this set (or nullptr when marching over all lines)
  int cnt;          // survivors used
  int seeds[4];
  int nseed;
  int total;        // number of lines visible to the march
  int c, j;
  double w;
};

__device__ __forceinline__ Line fetch_line(const SetView& v, int k) {
  const LineBatch& lb = *v.lb;
  int idx;
  if (v.surv == nullptr) idx = k;
  else idx = k < v.cnt ? v.surv[k] : v.seeds[k - v.cnt];
  Line L;
  L.idx = idx;
  L.ref = (lb.a_own != nullptr) ? (idx == lb.NA ? 0 : idx + 1) : idx;
  L.b = __dmul_rn(v.w, lb.Z[(size_t)v.c * lb.ldz + idx]);  // slopes = w_ji * znew (:321)
  L.a = line_intercept(lb, v.c, v.j, idx);
  return L;
}

constexpr int LANE_LINES = 4;  // lines cached in registers per lane (128 per warp)

// One warp.  Returns E[max]; writes hull records through `rec` callbacks in lane 0.
struct HullResult {
  double E;
  int h;
};

template <class Rec>
__device__ HullResult warp_march(const SetView& v, Rec rec) {
  const int lane = threadIdx.x & 31;
  Line cache[LANE_LINES];
#pragma unroll
  for (int r = 0; r < LANE_LINES; ++r) {
    int k = lane + 32 * r;
    if (k < v.total) cache[r] = fetch_line(v, k);
    else { cache[r].a = 0; cache[r].b = 0; cache[r].idx = -1; cache[r].ref = 0x7fffffff; }
  }

  // first line of the sorted order: minimum slope, maximum intercept among ties (:371-374)
  Line cur;
  cur.idx = -1; cur.a = 0; cur.b = 0; cur.ref = 0x7fffffff;
  {
#pragma unroll
    for (int r = 0; r < LANE_LINES; ++r)
      if (cache[r].idx >= 0 && (cur.idx < 0 || sorted_before(cache[r], cur))) cur = cache[r];
    for (int k = lane + 32 * LANE_LINES; k < v.total; k += 32) {
      Line L = fetch_line(v, k);
      if (cur.idx < 0 || sorted_before(L, cur)) cur = L;
    }
    for (int o = 16; o > 0; o >>= 1) {
      Line oth;
      oth.a = __shfl_xor_sync(0xffffffffu, cur.a, o);
      oth.b = __shfl_xor_sync(0xffffffffu, cur.b, o);
      oth.idx = __shfl_xor_sync(0xffffffffu, cur.idx, o);
      oth.ref = __shfl_xor_sync(0xffffffffu, cur.ref, o);
      if (oth.idx >= 0 && (cur.idx < 0 || sorted_before(oth, cur))) cur = oth;
    }
  }

  double E = 0.0, prev_cdf = 0.0, prev_pdf = 0.0;
  int h = 0;
  while (true) {
    // next vertex: among lines with a strictly different (i.e. larger) slope, the one whose
    // intersection with the current line comes first (:388-396); ties -> earliest in sorted order
    Line best;
    best.idx = -1; best.a = 0; best.b = 0; best.ref = 0x7fffffff;
    double bx = INFINITY;
    auto consider = [&](const Line& L) {
      if (L.b > cur.b) {
        double x = -(cur.a - L.a) / (cur.b - L.b);
        if (best.idx < 0 || x < bx || (x == bx && sorted_before(L, best))) {
          best = L;
          bx = x;
        }
      }
    };
#pragma unroll
    for (int r = 0; r < LANE_LINES; ++r)
      if (cache[r].idx >= 0) consider(cache[r]);
    for (int k = lane + 32 * LANE_LINES; k < v.total; k += 32) consider(fetch_line(v, k));
    for (int o = 16; o > 0; o >>= 1) {
      Line oth;
      oth.a = __shfl_xor_sync(0xffffffffu, best.a, o);
      oth.b = __shfl_xor_sync(0xffffffffu, best.b, o);
      oth.idx = __shfl_xor_sync(0xffffffffu, best.idx, o);
      oth.ref = __shfl_xor_sync(0xffffffffu, best.ref, o);
      double ox = __shfl_xor_sync(0xffffffffu, bx, o);
      if (oth.idx >= 0 && (best.idx < 0 || ox < bx || (ox == bx && sorted_before(oth, best)))) {
        best = oth;
        bx = ox;
      }
    }
    const bool last = best.idx < 0;
    const double cdf = last ? 1.0 : std_normal_cdf(bx);
    const double pdf = last ? 0.0 : std_normal_pdf(bx);
    const double dP = cdf - prev_cdf;
    const double dp = pdf - prev_pdf;
    // intercepts * (cdf[1:] - cdf[:-1]) - slopes * (pdf[1:] - pdf[:-1])   (:449-451)
    E += __dsub_rn(__dmul_rn(cur.a, dP), __dmul_rn(cur.b, dp));
    if (lane == 0) rec(h, cur, dP, -dp, bx, last);
    ++h;
    if (last) break;
    cur = best;
    prev_cdf = cdf;
    prev_pdf = pdf;
  }
  HullResult res;
  res.E = E;
  res.h = h;
  return res;
}

// ------------------------------------------------------------------------------------------
// hull + expectation + (optional) backward: one CTA per candidate row
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(E_THREADS)
hull_kernel(LineBatch lb, EmaxScratch sc, EmaxOut out, BackwardArgs bw) {
  extern __shared__ __align__(16) unsigned char e_smem[];
  const int c = blockIdx.x;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int S = lb.S;

  long long st_surv = 0, st_slow = 0, st_hull = 0, st_short = 0;

  for (int j = warp; j < S; j += nwarps) {
    const size_t set = (size_t)c * S + j;
    const double w = lb.wt ? lb.wt[j] : 1.0;
    const double zmin = sc.zst[c * 2 + 0], zmax = sc.zst[c * 2 + 1];
    double amax = lb.Amax[(size_t)c * lb.am_sc + j];
    int iT = lb.Aarg[(size_t)c * lb.am_sc + j];
    int own_is_max = 0;
    if (lb.a_own != nullptr) {
      double ao = lb.a_own[set];
      // torch.max over [own line, discretisation...] returns the first maximum: the own line
      // (reference index 0) wins ties.
      if (ao >= amax) { amax = ao; iT = lb.NA; own_is_max = 1; }
    }
    if (out.amax_is_own != nullptr && lane == 0) out.amax_is_own[set] = own_is_max;

    double E;
    int h;
    const size_t rbase = set * (size_t)out.hull_cap;
    auto rec = [&](int k, const Line& L, double p, double q, double x, bool last) {
      if (k < out.hull_cap) {
        if (out.hull_idx) out.hull_idx[rbase + k] = L.idx;
        if (out.hull_p) out.hull_p[rbase + k] = p;
        if (out.hull_q) out.hull_q[rbase + k] = q;
        if (out.hull_x && !last) out.hull_x[rbase + k] = x;
      }
      if (out.dense_da) out.dense_da[set * (size_t)lb.NL + L.idx] = p;
      if (out.dense_db) out.dense_db[set * (size_t)lb.NL + L.idx] = q;
    };

    const bool shortcut =
        fabs(__dmul_rn(w, zmin)) < SHORTCUT_TOL && fabs(__dmul_rn(w, zmax)) < SHORTCUT_TOL;
    if (shortcut) {
      // all |slopes| < 1e-9: the reference returns argmax(intercepts) only (:363-367)
      E = amax;
      h = 1;
      if (lane == 0) {
        Line L;
        L.idx = iT; L.a = amax; L.b = 0; L.ref = 0;
        rec(0, L, 1.0, 0.0, 0.0, true);
      }
      ++st_short;
    } else {
      SetView v;
      v.lb = &lb;
      v.c = c;
      v.j = j;
      v.w = w;
      const int cnt = sc.surv_cnt[set];
      if (cnt <= SURV_CAP) {
        v.surv = sc.surv_idx + set * SURV_CAP;
        v.cnt = cnt;
        v.nseed = 0;
        v.seeds[v.nseed++] = sc.zarg[c * 2 + 0];
        v.seeds[v.nseed++] = sc.zarg[c * 2 + 1];
        v.seeds[v.nseed++] = lb.Aarg[(size_t)c * lb.am_sc + j];
        if (lb.a_own != nullptr) v.seeds[v.nseed++] = lb.NA;
        v.total = cnt + v.nseed;
        st_surv += cnt;
      } else {  // filter kept too many lines: march over everything (slow, exact)
        v.surv = nullptr;
        v.cnt = 0;
        v.nseed = 0;
        v.total = lb.NL;
        ++st_slow;
      }
      HullResult r = warp_march(v, rec);
      E = r.E;
      h = r.h;
    }
    st_hull += h;
    if (lane == 0) {
      if (out.hull_cnt) out.hull_cnt[set] = h;
      out.terms[set] = out.subtract_max ? (E - amax) : E;  // kg[j] = E - max(intercepts) (:336)
    }
  }

  if (sc.stats != nullptr && lane == 0) {
    if (st_surv) atomicAdd((unsigned long long*)&sc.stats[1], (unsigned long long)st_surv);
    if (st_slow) atomicAdd((unsigned long long*)&sc.stats[2], (unsigned long long)st_slow);
    if (st_hull) atomicAdd((unsigned long long*)&sc.stats[3], (unsigned long long)st_hull);
    if (st_short) atomicAdd((unsigned long long*)&sc.stats[4], (unsigned long long)st_short);
  }
  if (out.kg == nullptr) return;
  __syncthreads();

  if (threadIdx.x == 0) {
    double acc = 0.0;
    for (int j = 0; j < S; ++j) acc += out.terms[(size_t)c * S + j];
    out.kg[c] = acc / (double)S;  // kg.mean() (:338)
  }
  if (bw.dX == nullptr) return;

  // ---------------- fused backward (envelope theorem) ----------------
  // dKG/da_jn = (p_jn - [n == argmax a_j]) / S ; dKG/db_jn = q_jn / S on hull lines only.
  double* s_r = reinterpret_cast<double*>(e_smem);  // [n_pad]  sum_n Gz[n] B[:, n]
  double* s_ga = s_r + bw.n_pad;                    // [S]      dKG/d a_own[j]
  double* s_sc = s_ga + S;                          // scalars: [0] Gsum, [1] GzOwn, [2..2+d) gkd,
                                                    //          [2+MAX_D .. ) Gm[m]
  double* s_red = s_sc + 2 + MAX_D + MAX_M;         // [nwarps * MAX_D] reduction scratch
  const int tgt = bw.target;
  const int d = bw.d;
  const double invS = 1.0 / (double)S;
  const double* zrow = lb.Z + (size_t)c * lb.ldz;
  const int hcap = out.hull_cap;

  // r_t over this thread's training points; every thread walks the same record list in order
  for (int t = threadIdx.x; t < bw.n_pad; t += blockDim.x) {
    double acc = 0.0;
    for (int j = 0; j < S; ++j) {
      const size_t set = (size_t)c * S + j;
      const int h = min(out.hull_cnt[set], hcap);
      const double wj = bw.W[j * bw.M + tgt];
      for (int k = 0; k < h; ++k) {
        const int idx = out.hull_idx[set * hcap + k];
        if (idx < lb.NA) {
          const double cz = wj * out.hull_q[set * hcap + k] * invS;
          acc += cz * bw.BT[(size_t)idx * bw.n_pad + t];
        }
      }
    }
    s_r[t] = acc;
  }
  // scalars: warp 0, lane l owns scalarisations l, l+32, ...
  if (warp == 0) {
    double gsum = 0.0, gzown = 0.0, gkd[MAX_D], gm[MAX_M];
#pragma unroll
    for (int k = 0; k < MAX_D; ++k) gkd[k] = 0.0;
#pragma unroll
    for (int m = 0; m < MAX_M; ++m) gm[m] = 0.0;
    double xs_t[MAX_D];
#pragma unroll
    for (int k = 0; k < MAX_D; ++k)
      xs_t[k] = k < d ? bw.X[(size_t)c * d + k] / bw.ls[tgt][k] : 0.0;
    for (int j = lane; j < S; j += 32) {
      const size_t set = (size_t)c * S + j;
      const int h = min(out.hull_cnt[set], hcap);
      const double wj = bw.W[j * bw.M + tgt];
      double ga = out.amax_is_own[set] ? -invS : 0.0;
      for (int k = 0; k < h; ++k) {
        const int idx = out.hull_idx[set * hcap + k];
        const double cz = wj * out.hull_q[set * hcap + k] * invS;
        gsum += cz * zrow[idx];
        if (idx == lb.NA) {
          gzown += cz;
          ga += out.hull_p[set * hcap + k] * invS;
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
    bw.dX[(size_t)c * d + threadIdx.x] = acc;
  }
}

int emax_hull(const LineBatch& lb, const EmaxScratch& sc, const EmaxOut& out,
              const BackwardArgs& bw, cudaStream_t st) {
  if (lb.C == 0) return DKG_OK;
  size_t smem = 0;
  if (bw.dX != nullptr)
    smem = sizeof(double) * ((size_t)bw.n_pad + lb.S + 2 + MAX_D + MAX_M + (E_THREADS / 32) * MAX_D);
  hull_kernel<<<lb.C, E_THREADS, smem, st>>>(lb, sc, out, bw);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

}  // namespace dkg
