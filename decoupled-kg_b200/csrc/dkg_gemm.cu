// fp64 tensor-core (DMMA) GEMM for the GP-conditioning contraction.
//
//   cov-mode : Z[c, n] = ( k_i(x_c, xd_n) - sum_t KX[c, t] * B[t, n] ) * ystd^2 / sd[c]
//              i.e. the fantasy-conditioned cross-covariance row of discretekg.py:301 divided by
//              the predictive standard deviation (discretekg.py:313), for a whole tile of
//              candidates at once and WITHOUT ever forming the (N+1)^2 covariance.
//   store-mode: D = A @ B   (used for T = KX @ Kinv, the variance quadratic form)
//
// tcgen05 has no f64 kind, so the fp64 tensor path on sm_100a is warp-level
// mma.sync.m8n8k4.f64 (SASS: DMMA.8x8x4; measured peak 37.1 TFLOP/s on B200, equal to the
// cuBLAS DGEMM rate).  CTA tile 128 x 64 x 16, 8 warps (4 x 2), warp tile 32 x 32 (4 x 4 DMMA
// fragments, 32 fp64 accumulators per thread) so that TWO CTAs fit per SM: while one CTA runs
// its kernel-evaluation epilogue the other keeps the DMMA pipe busy.  3-stage cp.async pipeline;
// shared-memory rows padded (+4 doubles) so every fragment load is bank-conflict free.  The
// epilogue first parks the accumulators in shared memory, then runs as a ROLLED loop (4 rows = 8
// kernel evaluations in flight per thread) with coalesced 512-byte row stores: fully unrolling it
// over the fragments made ~150 KB of straight-line code and the SM stalled on instruction fetch
// (ncu: stall_no_instruction 3.1 per issue, profiles/r01_ncu_summary.md).
#include "dkg_kernels.cuh"

namespace dkg {

constexpr int G_THREADS = 256;
constexpr int G_STAGES = 3;
constexpr int G_BN = 64;             // columns per CTA tile (GEMM_BN = 128 is the padding unit)
constexpr int LDA_S = GEMM_BK + 4;   // 20 doubles: rows land on distinct bank groups
constexpr int LDB_S = G_BN + 4;      // 68 doubles
constexpr int LDC_S = G_BN + 8;      // 72 doubles: accumulator staging, conflict-free double2 stores
constexpr int A_STAGE = GEMM_BM * LDA_S;
constexpr int B_STAGE = GEMM_BK * LDB_S;
constexpr size_t GEMM_SMEM_PIPE = (size_t)G_STAGES * (A_STAGE + B_STAGE) * sizeof(double);
constexpr size_t GEMM_SMEM_EPI = (size_t)(GEMM_BM * LDC_S + GEMM_BM * MAX_D + GEMM_BM) * sizeof(double);
constexpr size_t GEMM_SMEM = GEMM_SMEM_PIPE > GEMM_SMEM_EPI ? GEMM_SMEM_PIPE : GEMM_SMEM_EPI;
static_assert(GEMM_BN % G_BN == 0, "padding unit must be a multiple of the CTA tile width");

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier: used for the B
// operand, whose tile rows are 512 contiguous bytes.  (The A rows of a K-slice are only 128 bytes,
// and dense / swizzled cp.async.bulk.tensor tiles cannot be read by fp64 DMMA fragments without
// bank conflicts, so A stays on cp.async with padded rows: see DESIGN.md.)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MBAR_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MBAR_DONE_%=;\n"
      "bra MBAR_WAIT_%=;\n"
      "MBAR_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

template <bool COV>
__global__ void __launch_bounds__(G_THREADS, 2)
dmma_gemm_kernel(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                 int K, double* __restrict__ D, int ldd, CovEpilogue ep) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + G_STAGES * A_STAGE;
  __shared__ __align__(8) unsigned long long b_full[G_STAGES];  // "B tile of this stage has landed"

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int wm = (warp & 3) * 32;
  const int wn = (warp >> 2) * 32;
  const int g = lane >> 2;  // fragment row (A, C) / column (B)
  const int q = lane & 3;   // fragment k index (A, B) / column pair (C)

  const int m_base = blockIdx.y * GEMM_BM;
  const int n_base = blockIdx.x * G_BN;

  const double* Ag = A + (size_t)m_base * lda;
  const double* Bg = B + n_base;

  auto load_stage = [&](int stage, int k0) {
    double* as = As + stage * A_STAGE;
    double* bs = Bs + stage * B_STAGE;
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // A: 128 rows x 16 doubles = 1024 16-byte chunks
      int c = tid + i * G_THREADS;
      int row = c >> 3, ch = c & 7;
      cp_async16(as + row * LDA_S + ch * 2, Ag + (size_t)row * lda + k0 + ch * 2);
    }
    if (warp == 0) {  // B: 16 rows x 512 bytes, one TMA bulk copy per row into the padded tile
      if (lane == 0) mbar_arrive_expect_tx(&b_full[stage], GEMM_BK * G_BN * (unsigned)sizeof(double));
      __syncwarp();
      if (lane < GEMM_BK)
        bulk_copy_g2s(bs + lane * LDB_S, Bg + (size_t)(k0 + lane) * ldb, G_BN * (unsigned)sizeof(double),
                      &b_full[stage]);
    }
  };

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < G_STAGES; ++s) mbar_init(&b_full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  const int KT = K / GEMM_BK;
#pragma unroll
  for (int s = 0; s < G_STAGES - 1; ++s) {
    if (s < KT) load_stage(s, s * GEMM_BK);
    cp_async_commit();
  }

  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<G_STAGES - 2>();                                  // this thread's A chunks
    mbar_wait(&b_full[kt % G_STAGES], (kt / G_STAGES) & 1);         // the stage's B rows (TMA)
    __syncthreads();                                                // everyone's A chunks
    {
      int nk = kt + G_STAGES - 1;
      if (nk < KT) load_stage(nk % G_STAGES, nk * GEMM_BK);
      cp_async_commit();
    }
    const double* as = As + (kt % G_STAGES) * A_STAGE;
    const double* bs = Bs + (kt % G_STAGES) * B_STAGE;
#pragma unroll
    for (int kk = 0; kk < GEMM_BK / 4; ++kk) {
      double af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = as[(wm + i * 8 + g) * LDA_S + kk * 4 + q];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = bs[(kk * 4 + q) * LDB_S + wn + j * 8 + g];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();

  if (!COV) {
    // store mode: D = acc, or D = Cin + alpha * acc when an addend is given (ep.xs doubles as the
    // addend pointer with leading dimension ep.ldz and ep.ystd2 as alpha)
    const double* Cin = ep.xs;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int row = m_base + wm + i * 8 + g;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int col = n_base + wn + j * 8 + q * 2;
        double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
        if (Cin != nullptr) {
          const double2 cin = *reinterpret_cast<const double2*>(Cin + (size_t)row * ep.ldz + col);
          v.x = fma(ep.ystd2, v.x, cin.x);
          v.y = fma(ep.ystd2, v.y, cin.y);
        }
        *reinterpret_cast<double2*>(D + (size_t)row * ldd + col) = v;
      }
    }
    return;
  } else {
    // ---- epilogue: accumulators -> shared memory (the pipeline buffers are free now), then every
    // warp finishes whole rows: lane l owns columns 2l, 2l+1, so each row is one coalesced 512-byte
    // store and the kernel evaluations of 4 rows (8 outputs) are in flight per thread ----
    double* s_acc = smem;                               // [128][LDC_S]
    double* s_xr = smem + GEMM_BM * LDC_S;              // [128][MAX_D] candidates / lengthscale
    double* s_sd = s_xr + GEMM_BM * MAX_D;              // [128] 1 / sd
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<double2*>(s_acc + (wm + i * 8 + g) * LDC_S + wn + j * 8 + q * 2) =
            make_double2(acc[i][j][0], acc[i][j][1]);
    const int d = ep.d;
    for (int e = tid; e < GEMM_BM * MAX_D; e += G_THREADS) {
      int r = e / MAX_D, k = e - r * MAX_D;
      s_xr[e] = (k < d && m_base + r < ep.C) ? ep.xs[(size_t)(m_base + r) * d + k] : 0.0;
    }
    for (int e = tid; e < GEMM_BM; e += G_THREADS)
      s_sd[e] = (m_base + e < ep.C) ? 1.0 / ep.sd[m_base + e] : 1.0;
    // this thread's two discretisation points (scaled coordinates), fixed for all rows
    const int col = n_base + lane * 2;
    double xc0[MAX_D], xc1[MAX_D];
#pragma unroll
    for (int k = 0; k < MAX_D; ++k) {
      xc0[k] = (k < d) ? ep.xd_s[(size_t)col * d + k] : 0.0;
      xc1[k] = (k < d) ? ep.xd_s[(size_t)(col + 1) * d + k] : 0.0;
    }
    __syncthreads();
    const int kind = ep.kind;
    const double os = ep.outputscale, ystd2 = ep.ystd2;
#pragma unroll 1
    for (int rg = 0; rg < GEMM_BM / 8 / 4; ++rg) {  // 4 rows per iteration, rows warp + 8 * (..)
      double z0[4], z1[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int lr = warp + 8 * (rg * 4 + u);
        double sq0 = 0.0, sq1 = 0.0;
#pragma unroll
        for (int k = 0; k < MAX_D; ++k)
          if (k < d) {
            const double x = s_xr[lr * MAX_D + k];
            const double d0 = x - xc0[k], d1 = x - xc1[k];
            sq0 += d0 * d0;
            sq1 += d1 * d1;
          }
        const double2 a2 = *reinterpret_cast<const double2*>(s_acc + lr * LDC_S + lane * 2);
        // cov / sd (discretekg.py:313) as cov * (1 / sd): one rounding more than the division
        // (<= 1 ulp), ~12 fp64 instructions less per output
        const double rsd = s_sd[lr] * ystd2;
        z0[u] = (stationary_from_sq(kind, os, sq0) - a2.x) * rsd;
        z1[u] = (stationary_from_sq(kind, os, sq1) - a2.y) * rsd;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int row = m_base + warp + 8 * (rg * 4 + u);
        if (row < ep.C) {
          double* dst = ep.Z + (size_t)row * ep.ldz + col;
          if (col + 1 < ep.N) *reinterpret_cast<double2*>(dst) = make_double2(z0[u], z1[u]);
          else if (col < ep.N) dst[0] = z0[u];
        }
      }
    }
  }
}

static int ensure_gemm_attr() {
  // per device (the attribute lives in the context of the current device); cheap enough to
  // repeat, and correct when one process drives several GPUs
  static int done_for_device[64] = {0};
  int dev = 0;
  DKG_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !done_for_device[dev]) {
    DKG_CUDA_OK(cudaFuncSetAttribute(dmma_gemm_kernel<true>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
    DKG_CUDA_OK(cudaFuncSetAttribute(dmma_gemm_kernel<false>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
    if (dev >= 0 && dev < 64) done_for_device[dev] = 1;
  }
  return DKG_OK;
}

// D[M_pad, N_pad] = A[M_pad, K] @ B[K, N_pad]; all dims multiples of the tile sizes.
int gemm_store(const double* A, int lda, const double* B, int ldb, int M_pad, int N_pad, int K,
               double* D, int ldd, cudaStream_t st) {
  return gemm_axpy(A, lda, B, ldb, M_pad, N_pad, K, nullptr, 0, 1.0, D, ldd, st);
}

// D = Cin + alpha * (A @ B)   (Cin == nullptr: D = A @ B).  Cin / D may alias.
int gemm_axpy(const double* A, int lda, const double* B, int ldb, int M_pad, int N_pad, int K,
              const double* Cin, int ldc, double alpha, double* D, int ldd, cudaStream_t st) {
  DKG_TRY(ensure_gemm_attr());
  dim3 grid(N_pad / G_BN, M_pad / GEMM_BM);
  CovEpilogue ep{};
  ep.xs = Cin;
  ep.ldz = ldc;
  ep.ystd2 = alpha;
  dmma_gemm_kernel<false><<<grid, G_THREADS, GEMM_SMEM, st>>>(A, lda, B, ldb, K, D, ldd, ep);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

int gemm_cov(const double* KX, int lda, const double* B, int ldb, int M_pad, int N_pad, int K,
             const CovEpilogue& ep, cudaStream_t st) {
  DKG_TRY(ensure_gemm_attr());
  dim3 grid(N_pad / G_BN, M_pad / GEMM_BM);
  dmma_gemm_kernel<true><<<grid, G_THREADS, GEMM_SMEM, st>>>(KX, lda, B, ldb, K, nullptr, 0, ep);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

}  // namespace dkg
