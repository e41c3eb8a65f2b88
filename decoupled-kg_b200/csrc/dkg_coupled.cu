// Coupled evaluation (target_output_ix=None): reference calculate_discrete_kg
// (discretekg.py:162-235).  One joint posterior over all objectives, scalarised per weight vector
// with ScalarizedPosteriorTransform: for independent outputs (ModelListGP) the scalarised
// covariance is sum_m w_m^2 Cov_m and the noisy variance sum_m w_m^2 (var_m + noise_m) [BoTorch,
// recalled; pinned by the reference's coupled goldens, tests/test_reference_goldens.py].
//
// The conditioning GEMM runs once per objective (same kernel as the decoupled path, sd = 1), then
//   coupled_slopes : Zc[(c,j), n] = sum_m W[j,m]^2 Cov_m[c,n] / sqrt(sum_m W[j,m]^2 var_m[c])
// after which every (candidate, scalarisation) pair is an independent line set for the generic
// expected-max kernels (slope ORDER now differs per scalarisation), and
//   finalize_coupled: kg[c] = mean_j and the envelope-theorem backward through all M objectives.
#include "dkg_emax.cuh"

#include <cstdlib>

namespace dkg {

constexpr int CP_THREADS = 256;

__global__ void __launch_bounds__(CP_THREADS)
coupled_sd_kernel(CoupledArgs a) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.C * a.S) return;
  const int c = e / a.S, j = e - c * a.S;
  double v = 0.0;
  for (int m = 0; m < a.M; ++m) {
    const double w = a.W[j * a.M + m];
    v += (w * w) * a.varn[m][c];
  }
  a.sdj[e] = sqrt(v);  // xnew_variance.sqrt()  (:223)
}

__global__ void __launch_bounds__(CP_THREADS)
coupled_slopes_kernel(CoupledArgs a) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (n > a.N) return;  // column N is the candidate's own line
  double cov[MAX_M];
#pragma unroll
  for (int m = 0; m < MAX_M; ++m) cov[m] = m < a.M ? a.COV[m][(size_t)c * a.ldz + n] : 0.0;
  for (int j = 0; j < a.S; ++j) {
    double s = 0.0;
#pragma unroll
    for (int m = 0; m < MAX_M; ++m)
      if (m < a.M) {
        const double w = a.W[j * a.M + m];
        s += (w * w) * cov[m];
      }
    if (a.Zc != nullptr) a.Zc[((size_t)c * a.S + j) * a.ldz + n] = s / a.sdj[(size_t)c * a.S + j];
  }
}

// Same slope rows, plus the row statistics while the values are in registers.  Every WARP owns one
// segment of CS_TILE_LINES lines of one candidate: it keeps the M covariance segments in shared
// memory, walks the scalarisations, writes each slope once and reduces min / max (first index wins
// ties) with shuffles only -- no CTA barrier after the load.  The separate statistics pass re-read
// all of Zc (8.6 GB at c4: 2.6 ms).
// The quotient s / sd uses the correctly rounded reciprocal of the row's sd and one residual step,
//   q = s r;  q' = q + (s - q sd) r      (Markstein: q' = RN(s / sd) for r = RN(1 / sd)),
// 3 fp64 instructions instead of the ~25 of a full division, of which this kernel would need 1.07e9.
__global__ void __launch_bounds__(CP_THREADS)
coupled_slopes_stats_kernel(CoupledArgs a, int nseg) {
  extern __shared__ __align__(16) unsigned char c_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int seg = blockIdx.x * nwarp + warp, c = blockIdx.y;
  if (seg >= nseg) return;
  double* s_cov = reinterpret_cast<double*>(c_smem) + (size_t)warp * a.M * CS_TILE_LINES;  // [M][CS_TILE_LINES]
  const int n_lo = seg * CS_TILE_LINES;
  const int n_cnt = min(CS_TILE_LINES, a.N + 1 - n_lo);  // column N is the candidate's own line
  for (int m = 0; m < a.M; ++m)
    for (int i = lane; i < n_cnt; i += 32) s_cov[m * CS_TILE_LINES + i] = a.COV[m][(size_t)c * a.ldz + n_lo + i];
  __syncwarp();
  for (int j = 0; j < a.S; ++j) {
    double w2[MAX_M];
#pragma unroll
    for (int m = 0; m < MAX_M; ++m) w2[m] = m < a.M ? a.W2[j * a.M + m] : 0.0;
    const size_t row = (size_t)c * a.S + j;
    const double sd = a.sdj[row];
    const double rinv = 1.0 / sd;
    double* zrow = a.Zc != nullptr ? a.Zc + row * a.ldz + n_lo : nullptr;  // nullptr: statistics only
    double vmin = INFINITY, vmax = -INFINITY;
    int imin = 0x7fffffff, imax = 0x7fffffff;
#pragma unroll 4
    for (int i = lane; i < n_cnt; i += 32) {
      double s = 0.0;  // (same fma order and quotient as line_slope(): identical bits in every stage)
#pragma unroll
      for (int m = 0; m < MAX_M; ++m)
        if (m < a.M) s = fma(w2[m], s_cov[m * CS_TILE_LINES + i], s);
      const double z = coupled_quotient(s, sd, rinv);
      if (zrow != nullptr) zrow[i] = z;
      if (z < vmin) { vmin = z; imin = n_lo + i; }
      if (z > vmax) { vmax = z; imax = n_lo + i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, vmin, o);
      const int oi = __shfl_xor_sync(0xffffffffu, imin, o);
      if (ov < vmin || (ov == vmin && oi < imin)) { vmin = ov; imin = oi; }
      const double pv = __shfl_xor_sync(0xffffffffu, vmax, o);
      const int pi = __shfl_xor_sync(0xffffffffu, imax, o);
      if (pv > vmax || (pv == vmax && pi < imax)) { vmax = pv; imax = pi; }
    }
    if (lane == 0) {
      const size_t q = (row * nseg + seg) * 2;
      a.zpv[q] = vmin; a.zpv[q + 1] = vmax;
      a.zpi[q] = imin; a.zpi[q + 1] = imax;
    }
  }
}

// Statistics only (the slope rows are not materialised).  Every warp owns CST_SEG lines of one candidate and
// keeps the M covariance segments in REGISTERS (CST_SEG / 32 lines per lane); per scalarisation the lane-local
// min / max of the numerator  s = sum_m w2[j, m] cov_m[n]  costs M fp64 FMAs, two compares and the selects --
// the quotient is monotone in s (correctly rounded, sd > 0), so it is taken once per (segment, scalarisation) on
// the reduced value instead of once per line.  The 32 lane-local results of 16 scalarisations x {min, max} are
// transposed through shared memory: lane k reduces quantity k over the 32 lanes (one pass over 32 padded rows)
// instead of 16 x 2 five-step shuffle butterflies, which were half of the instructions of the fused
// assembly + statistics kernel above.
constexpr int CST_WARPS = 4;
constexpr int CST_PITCH = 33;  // padded row of the transpose buffer (conflict-free column reads)

template <int MT>
__global__ void __launch_bounds__(CST_WARPS * 32)
coupled_stats_kernel(CoupledArgs a, int nseg) {
  extern __shared__ __align__(16) unsigned char c_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int seg = blockIdx.x * CST_WARPS + warp, c = blockIdx.y;
  if (seg >= nseg) return;  // (no CTA barrier below: warps are independent)
  constexpr int E = CST_SEG / 32;
  double* s_v = reinterpret_cast<double*>(c_smem) + (size_t)warp * (32 * CST_PITCH);
  int* s_i = reinterpret_cast<int*>(reinterpret_cast<double*>(c_smem) + (size_t)CST_WARPS * 32 * CST_PITCH) +
             (size_t)warp * (32 * CST_PITCH);
  const int n_lo = seg * CST_SEG + lane;
  const double qnan = __longlong_as_double(0x7ff8000000000000ll);
  double cov[MT][E];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int n = n_lo + e * 32;  // column N is the candidate's own line; beyond it: NaN (never selected)
      cov[m][e] = (m < a.M && n <= a.N) ? a.COV[m][(size_t)c * a.ldz + n] : (m == 0 ? qnan : 0.0);
    }
  for (int jb = 0; jb < a.S; jb += 16) {
#pragma unroll 1
    for (int jj = 0; jj < 16; ++jj) {
      const int j = jb + jj;
      double vmin = INFINITY, vmax = -INFINITY;
      int emin = 0x7fffffff, emax = 0x7fffffff;
      if (j < a.S) {
        double w2[MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) w2[m] = m < a.M ? a.W2[j * a.M + m] : 0.0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          double s = 0.0;  // (same fma order as line_slope(): identical bits in every stage)
#pragma unroll
          for (int m = 0; m < MT; ++m)
            if (m < a.M) s = fma(w2[m], cov[m][e], s);
          if (s < vmin) { vmin = s; emin = e; }
          if (s > vmax) { vmax = s; emax = e; }
        }
      }
      s_v[jj * CST_PITCH + lane] = vmin;
      s_i[jj * CST_PITCH + lane] = emin == 0x7fffffff ? emin : n_lo + emin * 32;
      s_v[(16 + jj) * CST_PITCH + lane] = vmax;
      s_i[(16 + jj) * CST_PITCH + lane] = emax == 0x7fffffff ? emax : n_lo + emax * 32;
    }
    __syncwarp();
    {  // lane k < 16: min of scalarisation jb + k; lane 16 + k: its max (first index wins ties)
      const bool is_max = lane >= 16;
      const double* rv = s_v + lane * CST_PITCH;
      const int* ri = s_i + lane * CST_PITCH;
      double v = rv[0];
      int i = ri[0];
#pragma unroll 8
      for (int l = 1; l < 32; ++l) {
        const double ov = rv[l];
        const int oi = ri[l];
        const bool better = (is_max ? ov > v : ov < v) || (ov == v && oi < i);
        if (better) { v = ov; i = oi; }
      }
      const int j = jb + (lane & 15);
      if (j < a.S) {
        const size_t row = (size_t)c * a.S + j;
        const double sd = a.sdj[row];
        const size_t q = (row * nseg + seg) * 2 + (is_max ? 1 : 0);
        a.zpv[q] = coupled_quotient(v, sd, 1.0 / sd);
        a.zpi[q] = i;
      }
    }
    __syncwarp();
  }
}

// ---- statistics only, coarse ranges first (default) ---------------------------------------------
// The exact (min, max, index) bookkeeping above costs ~8 of the ~12 instructions per (line, scalarisation) of
// an issue-bound kernel.  The HIGH WORD of a double (sign, exponent, 20 mantissa bits), mapped from
// sign-magnitude to two's complement, is a monotone (non-strict) integer image of the value, so the exact minimum
// of a row lies in a segment whose smallest key equals the smallest key of the row: pass 1 only keeps the key
// range of the numerators per (row, segment) -- M fp64 FMAs and four integer ALU instructions per (line,
// scalarisation), one REDUX pair per (segment, scalarisation); a float conversion instead of the key runs on the
// quarter-rate conversion pipe and was the limiter -- and pass 2 (one warp per row) re-forms the numerators of
// just the segments that attain the row's extreme keys, exactly, with the first index winning ties.
__device__ __forceinline__ int order_key(double s) {
  const int hi = __double2hiint(s);
  return hi ^ ((hi >> 31) & 0x7fffffff);
}

template <int MT>
__global__ void __launch_bounds__(CST_WARPS * 32)
coupled_range_kernel(CoupledArgs a, int nseg, int2* __restrict__ zrange) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int seg = blockIdx.x * CST_WARPS + warp, c = blockIdx.y;
  if (seg >= nseg) return;
  constexpr int E = CST_SEG / 32;
  const int n_lo = seg * CST_SEG + lane;
  const bool whole = (seg + 1) * CST_SEG <= a.N + 1;  // (warp-uniform) every line of the segment exists
  double cov[MT][E];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int n = n_lo + e * 32;  // column N is the candidate's own line
      cov[m][e] = (m < a.M && n <= a.N) ? a.COV[m][(size_t)c * a.ldz + n] : 0.0;
    }
#pragma unroll 2
  for (int j = 0; j < a.S; ++j) {
    double w2[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) w2[m] = m < a.M ? a.W2[j * a.M + m] : 0.0;
    int klo = 0x7fffffff, khi = (int)0x80000000;
    if (whole) {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        double s = 0.0;  // (same fma order as line_slope(): identical bits in every stage)
#pragma unroll
        for (int m = 0; m < MT; ++m) s = fma(w2[m], cov[m][e], s);  // (w2 = cov = 0 beyond M)
        const int k = order_key(s);
        klo = min(klo, k);
        khi = max(khi, k);
      }
    } else {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        double s = 0.0;
#pragma unroll
        for (int m = 0; m < MT; ++m) s = fma(w2[m], cov[m][e], s);  // (w2 = cov = 0 beyond M)
        const int k = order_key(s);
        if (n_lo + e * 32 <= a.N) {
          klo = min(klo, k);
          khi = max(khi, k);
        }
      }
    }
    klo = __reduce_min_sync(0xffffffffu, klo);
    khi = __reduce_max_sync(0xffffffffu, khi);
    if (lane == 0) zrange[((size_t)c * a.S + j) * nseg + seg] = make_int2(klo, khi);
  }
}

// one warp per row (candidate, scalarisation): exact extremes of the numerator from the segments that attain
// the extreme keys, then the quotient (monotone in the numerator) once per extreme
__global__ void __launch_bounds__(CP_THREADS)
coupled_zreduce_kernel(CoupledArgs a, int nseg, const int2* __restrict__ zrange) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (CP_THREADS / 32) + (threadIdx.x >> 5);
  if (row >= a.C * a.S) return;
  const int c = row / a.S, j = row - c * a.S;
  const int2* zr = zrange + (size_t)row * nseg;
  int glo = 0x7fffffff, ghi = (int)0x80000000;
  for (int t = lane; t < nseg; t += 32) {
    const int2 f = zr[t];
    glo = min(glo, f.x);
    ghi = max(ghi, f.y);
  }
  glo = __reduce_min_sync(0xffffffffu, glo);
  ghi = __reduce_max_sync(0xffffffffu, ghi);
  double w2[MAX_M];
#pragma unroll
  for (int m = 0; m < MAX_M; ++m) w2[m] = m < a.M ? a.W2[j * a.M + m] : 0.0;
  double vmin = INFINITY, vmax = -INFINITY;
  int imin = 0x7fffffff, imax = 0x7fffffff;
  for (int t0 = 0; t0 < nseg; t0 += 32) {
    const int t = t0 + lane;
    const int2 f = t < nseg ? zr[t] : make_int2(0x7fffffff, (int)0x80000000);
    unsigned hit = __ballot_sync(0xffffffffu, t < nseg && (f.x == glo || f.y == ghi));
    while (hit) {
      const int seg = t0 + __ffs(hit) - 1;
      hit &= hit - 1u;
      for (int e = 0; e < CST_SEG / 32; ++e) {
        const int n = seg * CST_SEG + lane + 32 * e;
        if (n <= a.N) {
          double s = 0.0;
#pragma unroll
          for (int m = 0; m < MAX_M; ++m)
            if (m < a.M) s = fma(w2[m], a.COV[m][(size_t)c * a.ldz + n], s);
          if (s < vmin) { vmin = s; imin = n; }
          if (s > vmax) { vmax = s; imax = n; }
        }
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, vmin, o);
    const int oi = __shfl_xor_sync(0xffffffffu, imin, o);
    if (ov < vmin || (ov == vmin && oi < imin)) { vmin = ov; imin = oi; }
    const double pv = __shfl_xor_sync(0xffffffffu, vmax, o);
    const int pi = __shfl_xor_sync(0xffffffffu, imax, o);
    if (pv > vmax || (pv == vmax && pi < imax)) { vmax = pv; imax = pi; }
  }
  if (lane == 0) {
    const double sd = a.sdj[row], rinv = 1.0 / sd;
    a.zst[row * 2 + 0] = coupled_quotient(vmin, sd, rinv);
    a.zst[row * 2 + 1] = coupled_quotient(vmax, sd, rinv);
    a.zarg[row * 2 + 0] = imin;
    a.zarg[row * 2 + 1] = imax;
  }
}

template <int MT>
static int launch_coupled_ranges(const CoupledArgs& a, cudaStream_t st) {
  const int nseg = ceil_div(a.N + 1, CST_SEG);
  int2* zrange = reinterpret_cast<int2*>(a.zpv);  // [C * S, nseg] int2 <= the [C * S, tiles, 2] doubles reserved
  coupled_range_kernel<MT><<<dim3(ceil_div(nseg, CST_WARPS), a.C), CST_WARPS * 32, 0, st>>>(a, nseg, zrange);
  DKG_LAUNCH_CHECK();
  coupled_zreduce_kernel<<<ceil_div(a.C * a.S, CP_THREADS / 32), CP_THREADS, 0, st>>>(a, nseg, zrange);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// segments of per-(row, segment) partials the caller has to reduce (emax_zstat_from_partials); 0: the row
// statistics were written directly
int coupled_stat_segments(int N, bool materialised) {
  if (!materialised && getenv("DKG_COUPLED_STATS_EXACT") == nullptr) return 0;
  return ceil_div(N + 1, materialised ? CS_TILE_LINES : CST_SEG);
}

template <int MT>
static int launch_coupled_stats(const CoupledArgs& a, cudaStream_t st) {
  if (a.zst != nullptr && a.zarg != nullptr && getenv("DKG_COUPLED_STATS_EXACT") == nullptr) return launch_coupled_ranges<MT>(a, st);
  const int nseg = ceil_div(a.N + 1, CST_SEG);
  const size_t smem = (size_t)CST_WARPS * 32 * CST_PITCH * (sizeof(double) + sizeof(int));
  DKG_CUDA_OK(cudaFuncSetAttribute(coupled_stats_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  coupled_stats_kernel<MT><<<dim3(ceil_div(nseg, CST_WARPS), a.C), CST_WARPS * 32, smem, st>>>(a, nseg);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

int coupled_slopes(const CoupledArgs& a, cudaStream_t st) {
  if (a.C == 0) return DKG_OK;
  coupled_sd_kernel<<<ceil_div(a.C * a.S, CP_THREADS), CP_THREADS, 0, st>>>(a);
  DKG_LAUNCH_CHECK();
  if (a.Zc == nullptr && a.zpv != nullptr && a.zpi != nullptr) {
    if (a.M <= 2) return launch_coupled_stats<2>(a, st);
    if (a.M <= 4) return launch_coupled_stats<4>(a, st);
    return launch_coupled_stats<MAX_M>(a, st);
  }
  if (a.zpv != nullptr && a.zpi != nullptr) {
    const int nseg = ceil_div(a.N + 1, CS_TILE_LINES);
    const int nwarp = CP_THREADS / 32;
    const size_t smem = sizeof(double) * (size_t)nwarp * a.M * CS_TILE_LINES;
    if (smem > 47 * 1024)
      DKG_CUDA_OK(cudaFuncSetAttribute(coupled_slopes_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    coupled_slopes_stats_kernel<<<dim3(ceil_div(nseg, nwarp), a.C), CP_THREADS, smem, st>>>(a, nseg);
    DKG_LAUNCH_CHECK();
    return DKG_OK;
  }
  dim3 grid(ceil_div(a.N + 1, CP_THREADS), a.C);
  coupled_slopes_kernel<<<grid, CP_THREADS, 0, st>>>(a);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// ------------------------------------------------------------------------------------------
// finalize (coupled): CTA per candidate.  With b_jn = sum_m om_jm cov_m[n], om_jm = W_jm^2 / sd_j,
// var_j = sum_m W_jm^2 var_m:
//   dKG = sum_j Ga_j d a_own_j + sum_m [ sum_n Gz_m[n] d cov_m[n] + Cv_m d var_m ]
//   Gz_m[n] = sum_j om_jm q_jn / S,   Cv_m = -sum_j W_jm^2 / (2 var_j) sum_n (q_jn / S) b_jn
// The hull records of the candidate (all scalarisations) are staged in shared memory once and their DISTINCT
// lines found through the same hash table as in finalize_kernel: the same few lines are hull vertices of
// most scalarisations, so per objective the B_m^T row gathers and the kernel-gradient evaluations run once
// per distinct line with the merged coefficient Gz_m[n].  (Before: every thread walked every record of every
// scalarisation through the spill-chain reader for every training point -- 2.05 of the 7.5 ms of a c4 step.)
// Candidates with more records than the staging capacity are handled in batches of that capacity.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CP_THREADS)
finalize_coupled_kernel(int S, EmaxOut out, CoupledBackward bw, int rec_batch) {
  extern __shared__ __align__(16) unsigned char c_smem[];
  const int c = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  double* s_term = reinterpret_cast<double*>(c_smem);  // [S]
  for (int j = threadIdx.x; j < S; j += blockDim.x) s_term[j] = out.terms[(size_t)c * S + j];
  __syncthreads();
  if (threadIdx.x == 0) {
    double acc = 0.0;
    for (int j = 0; j < S; ++j) acc += s_term[j];
    out.kg[c] = acc / (double)S;  // kg.mean() (:235)
  }
  if (bw.dX == nullptr) return;

  const int d = bw.d, M = bw.M, NA = bw.N;
  __shared__ int s_broken, s_nrec, s_nuniq, s_hovf;
  const double invS = 1.0 / (double)S;
  int n_pad_max = 0;
  for (int m = 0; m < M; ++m) n_pad_max = max(n_pad_max, bw.n_pad[m]);
  constexpr int SCW = 2 + MAX_D;                 // per objective: [0] gzown, [1] Cv, [2..) gkd
  double* s_r = s_term + S;                      // [M][n_pad_max]  r_m, accumulated over the record batches
  double* s_sc = s_r + (size_t)M * n_pad_max;    // [M][SCW], then gm[MAX_M]
  double* s_gm = s_sc + MAX_M * SCW;             // [MAX_M]
  double* s_red = s_gm + MAX_M;                  // [nwarps * MAX_D]   (sized for CP_THREADS / 32 warps)
  const int RM = fin_rmax(S);
  const int HB = fin_hash_bits(S), HN = 1 << HB;
  double* s_rq = s_red + (CP_THREADS / 32) * MAX_D;  // [RM] q / S of a record
  double* s_ucz = s_rq + RM;                     // [RM] merged coefficient Gz_m of a distinct line
  double* s_om = s_ucz + RM;                     // [S]  om_jm of the current objective
  double* s_qb = s_om + S;                       // [S]  sum_n (q_jn / S) b_jn
  int* s_ridx = reinterpret_cast<int*>(s_qb + S);  // [RM] record line index
  int* s_rj = s_ridx + RM;                       // [RM] record scalarisation
  int* s_flag = s_rj + RM;                       // [RM] hash slot, then first-occurrence flag
  int* s_uidx = s_flag + RM;                     // [RM] distinct line index
  int* s_ue = s_uidx + RM;                       // [RM] its first record
  int* s_roff = s_ue + RM;                       // [S + 1] record offsets per set
  int* s_hkey = s_roff + ((S + 2) & ~1);         // [HN]
  int* s_hfirst = s_hkey + HN;                   // [HN]
  double grad[MAX_D];
#pragma unroll
  for (int k = 0; k < MAX_D; ++k) grad[k] = 0.0;

  for (int j = threadIdx.x; j < S; j += blockDim.x) { s_roff[j + 1] = out.hull_cnt[(size_t)c * S + j]; s_qb[j] = 0.0; }
  for (int h = threadIdx.x; h < HN; h += blockDim.x) { s_hkey[h] = -1; s_hfirst[h] = 0x7fffffff; }
  for (int e = threadIdx.x; e < M * n_pad_max; e += blockDim.x) s_r[e] = 0.0;
  for (int e = threadIdx.x; e < MAX_M * SCW; e += blockDim.x) s_sc[e] = 0.0;
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
  __syncthreads();
  const int nrec = s_nrec;

  // gm[m] = sum_j Ga_j W[j, m]  (own-line intercept path), computed once by warp 0
  if (warp == 0) {
    double gm[MAX_M];
#pragma unroll
    for (int m = 0; m < MAX_M; ++m) gm[m] = 0.0;
    for (int j = lane; j < S; j += 32) {
      const size_t set = (size_t)c * S + j;
      const int h = out.hull_cnt[set];
      double ga = out.amax_is_own[set] ? -invS : 0.0;
      HullReader rd(out, set);
      for (int k = 0; k < h; ++k) {
        if (!rd.seek(k)) { s_broken = 1; break; }
        if (rd.idx() == NA) ga += rd.p() * invS;
      }
#pragma unroll
      for (int m = 0; m < MAX_M; ++m)
        if (m < M) gm[m] += ga * bw.W[j * M + m];
    }
#pragma unroll
    for (int m = 0; m < MAX_M; ++m) gm[m] = warp_sum(gm[m]);
    if (lane == 0)
      for (int m = 0; m < MAX_M; ++m) s_gm[m] = gm[m];
  }

  // ---- the hull records in batches of what the shared-memory tables hold (one batch unless S is large): distinct
  // lines of the batch through the hash table, then per objective the merged coefficients, the B_m^T row gathers
  // into r_m and the prior-covariance terms; everything accumulates over the batches ----
  const int RB = rec_batch > 0 && rec_batch < RM ? rec_batch : RM;  // (a smaller batch only as a test hook: DKG_FIN_BATCH)
  for (int b0 = 0; b0 < nrec; b0 += RB) {
    const int nb = min(RB, nrec - b0);
    __syncthreads();  // the previous batch's tables are done with (first batch: s_gm / the clears above are visible)
    if (b0 > 0) {
      for (int h = threadIdx.x; h < HN; h += blockDim.x) { s_hkey[h] = -1; s_hfirst[h] = 0x7fffffff; }
      if (threadIdx.x == 0) s_nuniq = 0;
      __syncthreads();
    }
    for (int j = warp; j < S; j += nwarps) {
      if (s_roff[j + 1] <= b0 || s_roff[j] >= b0 + nb) continue;
      const size_t set = (size_t)c * S + j;
      const int h = s_roff[j + 1] - s_roff[j];
      HullReader rd(out, set);
      for (int k = lane; k < h; k += 32) {
        const int e = s_roff[j] + k - b0;
        if (e < 0 || e >= nb) continue;
        s_rj[e] = j;
        if (!rd.seek(k)) { s_broken = 1; s_ridx[e] = NA; s_rq[e] = 0.0; continue; }
        s_ridx[e] = rd.idx();
        s_rq[e] = rd.q() * invS;
      }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nb; e += blockDim.x) {
      const int idx = s_ridx[e];
      unsigned h = ((unsigned)idx * 2654435761u) >> (32 - HB);
      int probes = 0;
      for (;; h = (h + 1) & (HN - 1)) {
        const int prev = atomicCAS(&s_hkey[h], -1, idx);
        if (prev == -1 || prev == idx) break;
        if (++probes >= HN) { s_hovf = 1; break; }  // (cannot happen: HN > RM)
      }
      atomicMin(&s_hfirst[h], e);
      s_flag[e] = (int)h;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nb; e += blockDim.x) s_flag[e] = s_hfirst[s_flag[e]] == e ? 1 : 0;
    __syncthreads();
    if (s_hovf) s_broken = 1;  // fail loudly rather than silently drop lines
    for (int e = threadIdx.x; e < nb; e += blockDim.x) {
      if (!s_flag[e]) continue;
      int rank = 0;  // distinct lines that first occur before e (order = first occurrence: deterministic)
      for (int f = 0; f < e; ++f) rank += s_flag[f];
      s_uidx[rank] = s_ridx[e];
      s_ue[rank] = e;
      atomicAdd(&s_nuniq, 1);
    }
    // qb_j = sum_n (q_jn / S) b_jn: the same for every objective; this batch's records of set j
    for (int j = threadIdx.x; j < S; j += blockDim.x) {
      const int e_lo = max(s_roff[j], b0) - b0, e_hi = min(s_roff[j + 1], b0 + nb) - b0;
      if (e_lo >= e_hi) continue;
      const size_t set = (size_t)c * S + j;
      const double sd = bw.sdj[set], rinv = 1.0 / sd;
      double qb = 0.0;
      for (int e = e_lo; e < e_hi; ++e) {
        double sacc = 0.0;  // b_jn, formed as everywhere else (line_slope)
        for (int mm = 0; mm < M; ++mm) sacc = fma(bw.W2[j * M + mm], bw.COV[mm][(size_t)c * bw.ldz + s_ridx[e]], sacc);
        qb += s_rq[e] * coupled_quotient(sacc, sd, rinv);
      }
      s_qb[j] += qb;
    }
    __syncthreads();
    const int nuniq = s_nuniq;
    for (int m = 0; m < M; ++m) {
      for (int j = threadIdx.x; j < S; j += blockDim.x) {
        const double w = bw.W[j * M + m];
        s_om[j] = (w * w) / bw.sdj[(size_t)c * S + j];
      }
      __syncthreads();
      for (int u = threadIdx.x; u < nuniq; u += blockDim.x) {
        const int idx = s_uidx[u];
        double acc = 0.0;
        for (int f = s_ue[u]; f < nb; ++f)
          if (s_ridx[f] == idx) acc += s_om[s_rj[f]] * s_rq[f];
        s_ucz[u] = acc;
      }
      __syncthreads();
      // r_m[t] += sum over the distinct discretisation lines of Gz_m[n] * B_m^T[n, t]
      double* rm = s_r + (size_t)m * n_pad_max;
      for (int t = threadIdx.x; t < bw.n_pad[m]; t += blockDim.x) {
        double acc = 0.0;
        for (int u = 0; u < nuniq; ++u) {
          const int idx = s_uidx[u];
          if (idx < NA) acc += s_ucz[u] * bw.BT[m][(size_t)idx * bw.ldbt[m] + t];
        }
        rm[t] += acc;
      }
      if (warp == 0) {  // prior-covariance path of the distinct lines with their coefficient Gz_m[n]
        double gzown = 0.0, gkd[MAX_D];
#pragma unroll
        for (int k = 0; k < MAX_D; ++k) gkd[k] = 0.0;
        double xm[MAX_D];
#pragma unroll
        for (int k = 0; k < MAX_D; ++k) xm[k] = k < d ? bw.X[(size_t)c * d + k] / bw.ls[m][k] : 0.0;
        for (int u = lane; u < nuniq; u += 32) {
          const int idx = s_uidx[u];
          const double cz = s_ucz[u];
          if (idx == NA) {
            gzown += cz;
          } else {
            double sq = 0.0;
#pragma unroll
            for (int q = 0; q < MAX_D; ++q)
              if (q < d) {
                const double df = xm[q] - bw.xd_s[m][(size_t)idx * d + q];
                sq += df * df;
              }
            const double gc = stationary_grad_coeff(bw.kind[m], bw.outputscale[m], sq);
#pragma unroll
            for (int q = 0; q < MAX_D; ++q)
              if (q < d) gkd[q] += cz * gc * (xm[q] - bw.xd_s[m][(size_t)idx * d + q]) / bw.ls[m][q];
          }
        }
        gzown = warp_sum(gzown);
#pragma unroll
        for (int k = 0; k < MAX_D; ++k) gkd[k] = warp_sum(gkd[k]);
        if (lane == 0) {
          s_sc[m * SCW + 0] += gzown;
          for (int k = 0; k < MAX_D; ++k) s_sc[m * SCW + 2 + k] += gkd[k];
        }
      }
      __syncthreads();  // s_om / s_ucz are reused by the next objective
    }
  }
  __syncthreads();

  for (int m = 0; m < M; ++m) {
    const double s2 = bw.y_std[m] * bw.y_std[m];
    if (warp == 0) {  // Cv_m from the complete qb_j
      double cv = 0.0;
      for (int j = lane; j < S; j += 32) {
        const double w = bw.W[j * M + m];
        const double sd = bw.sdj[(size_t)c * S + j];
        cv -= (w * w) / (2.0 * sd * sd) * s_qb[j];
      }
      cv = warp_sum(cv);
      if (lane == 0) s_sc[m * SCW + 1] = cv;
    }
    __syncthreads();
    const double* rm = s_r + (size_t)m * n_pad_max;
    const double cT = -2.0 * s2 * (s_sc[m * SCW + 0] + s_sc[m * SCW + 1]);
    const double cm = s_gm[m] * bw.y_std[m];
    double xm[MAX_D];
#pragma unroll
    for (int k = 0; k < MAX_D; ++k) xm[k] = k < d ? bw.X[(size_t)c * d + k] / bw.ls[m][k] : 0.0;
    for (int t = threadIdx.x; t < bw.ntr[m]; t += blockDim.x) {
      const double u = cm * bw.alpha[m][t] - s2 * rm[t] + cT * bw.T[m][(size_t)c * bw.ldk[m] + t];
      double sq = 0.0;
#pragma unroll
      for (int k = 0; k < MAX_D; ++k)
        if (k < d) {
          const double df = xm[k] - bw.xs[m][(size_t)t * d + k];
          sq += df * df;
        }
      const double gc = u * stationary_grad_coeff(bw.kind[m], bw.outputscale[m], sq);
#pragma unroll
      for (int k = 0; k < MAX_D; ++k)
        if (k < d) grad[k] += gc * (xm[k] - bw.xs[m][(size_t)t * d + k]) / bw.ls[m][k];
    }
    if (threadIdx.x == 0)
      for (int k = 0; k < MAX_D; ++k) grad[k] += s2 * s_sc[m * SCW + 2 + k];  // prior-covariance term, once
  }
#pragma unroll
  for (int k = 0; k < MAX_D; ++k) grad[k] = warp_sum(grad[k]);
  if (lane == 0)
    for (int k = 0; k < MAX_D; ++k) s_red[warp * MAX_D + k] = grad[k];
  __syncthreads();
  if (threadIdx.x < d) {
    double acc = 0.0;
    for (int wv = 0; wv < nwarps; ++wv) acc += s_red[wv * MAX_D + threadIdx.x];
    // hull records lost to an exhausted spill pool: fail loudly (NaN) instead of a partial gradient
    bw.dX[(size_t)c * d + threadIdx.x] = s_broken ? __longlong_as_double(0x7ff8000000000000ll) : acc;
  }
}

int emax_finalize_coupled(int C, int S, const EmaxOut& out, const CoupledBackward& bw, cudaStream_t st) {
  if (C == 0 || out.kg == nullptr) return DKG_OK;
  int n_pad_max = 0;
  for (int m = 0; m < bw.M; ++m) n_pad_max = n_pad_max > bw.n_pad[m] ? n_pad_max : bw.n_pad[m];
  size_t smem = sizeof(double) * S;
  if (bw.dX != nullptr)
    smem = sizeof(double) * ((size_t)S + (size_t)bw.M * n_pad_max + (size_t)MAX_M * (2 + MAX_D) + MAX_M +
                             (CP_THREADS / 32) * MAX_D + 2 * fin_rmax(S) + 2 * S) +
           sizeof(int) * ((size_t)5 * fin_rmax(S) + ((S + 2) & ~1) + (2 << fin_hash_bits(S)));
  if (smem > 47 * 1024)
    DKG_CUDA_OK(cudaFuncSetAttribute(finalize_coupled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const char* fbe = getenv("DKG_FIN_BATCH");  // test hook: hull records merged per batch (default: the table capacity)
  finalize_coupled_kernel<<<C, CP_THREADS / 2, smem, st>>>(S, out, bw, fbe != nullptr ? atoi(fbe) : 0);  // short barrier-separated phases: more CTAs per SM
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

}  // namespace dkg
