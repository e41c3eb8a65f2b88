// fp64-accurate GP-conditioning contraction on the 5th-generation tensor cores (tcgen05, int8).
//
//   cov-mode  : Z[c, n] = ( k_i(x_c, xd_n) - sum_t T[c, t] * Kxd[t, n] ) * ystd^2 / sd[c]
//               (the fantasy-conditioned cross-covariance row of discretekg.py:301 over the
//               predictive standard deviation, discretekg.py:313)
//   store-mode: D = A @ B^T   (test hook)
//
// tcgen05.mma has no f64 kind, and the fp64 DMMA pipe of B200 tops out at 37 TFLOP/s.  The int8
// kind runs at ~4.5 POP/s, so the product is evaluated with the Ozaki scheme: every row of T and
// every column of Kxd is scaled by a power of two into [-127, 127], rounded to its last digit
// (round-to-nearest: the representation error is unbiased, <= 2^-56 of full scale) and cut into NS
// BALANCED base-256 digits,
//     x = d_0 + d_1 256^-1 + d_2 256^-2 + ...      every d_s in [-128, 127]  (signed int8),
// which is exact (integer arithmetic on the rounded 55-bit fixed-point value).  Then
//     sum_t x_t y_t = sum_g 256^-g  sum_{i+j=g} sum_t d_i(t) e_j(t),
// and every inner sum is an integer GEMM whose int32 accumulator cannot overflow
// (7 pairs * K <= 4096 * 128^2 < 2^31).  Diagonals g < NG are kept.  Balanced digits have (almost)
// zero mean, so the dropped diagonals are sums of zero-mean terms: with NS = 7, NG = 7 the largest
// dropped one (g = 7, six digit pairs) contributes sqrt(6 K) * 74^2 * 2^-56 ~ 4e-12 in units where
// |row|max, |col|max are in [64, 128), i.e. ~1e-15 * sqrt(K / 400) relative to |row|max * |col|max --
// the rounding error of an fp64 dot product of that length.  (Round 1 used unsigned digits d_s in
// [0, 255] for s >= 1: their mean of 127.5 makes every dropped diagonal a BIASED sum ~ K * 127.5^2,
// which forced NG = 8: 34 digit products instead of 28.)
//
// Kernel (ozaki_kernel): persistent, one CTA per SM (grid = #SMs), 128 x 128 output tile, 640 threads.
//   warp 0      : producer - per k block (K = 32 bytes) two LINEAR cp.async.bulk copies (the digits
//                 0..nd-1 of one (row block, k block) are one contiguous run of ready-made UMMA operand
//                 images in global memory: 128 rows x 32 B, canonical NO-SWIZZLE core-matrix order),
//                 into a 2-stage ring of 64 KB stages, mbarrier complete_tx.
//   warp 1      : MMA issuer - one lane chosen with elect.sync issues tcgen05.mma.cta_group::1.kind::i8
//                 (M128 N128 K32, or N256 when two adjacent B digit planes are merged).  The loop is
//                 k-outer with one TMEM accumulator slot (128 columns) per DIAGONAL: pass 0 folds the
//                 diagonals 6..3 into the four slots (all 512 TMEM columns), pass 1 the diagonals 2..0,
//                 so every digit block is fetched once per pass.  tcgen05.commit frees the stage and
//                 publishes each slot.  The first and the last k block of a pass are issued slot by
//                 slot in the order the epilogue drains them, so the drain of one pass overlaps the
//                 MMAs at the seam of two passes instead of stalling the tensor pipe.
//   warps 4..19 : epilogue (setmaxnreg 104) - tcgen05.ld the int32 diagonal sums (32 lanes x 32
//                 columns per warp), convert exactly to fp64 and fold acc += 256^-g * S_g into 32 fp64
//                 registers per thread; after the last pass apply the two power-of-two scales and
//                 write the 32 x 32 block through an 8-column shared-memory transpose.
//   (warps 2, 3 idle; setmaxnreg 56 for warps 0..3.)
// Every digit plane is signed (a_format = b_format = INT8), so two adjacent B planes can always share
// one N = 256 instruction.
// ozaki_pair_kernel is the cta_group::2 variant (cluster of 2, M = 256); measured no faster, not default.
#include <cstdlib>

#include "dkg_kernels.cuh"

namespace dkg {

namespace {

constexpr int OZ_BM = 128;
constexpr int OZ_BN = 128;
constexpr int OZ_KB = 32;             // bytes of K per k block = one MMA (K32); no-swizzle core-matrix images
constexpr int OZ_BLK_BYTES = OZ_BM * OZ_KB;  // one digit block (A or B) of a stage: 4 KB
constexpr int OZ_MAX_DIGITS = 7;
constexpr int OZ_STAGES = 2;
constexpr int OZ_STAGE_CAP = 65536;   // bytes per stage: 1 k block of 7+7 digit blocks, or 2 of 4+4
constexpr int OZ_MAX_KPS = 2;         // k blocks per stage
constexpr int OZ_ACC = 4;             // TMEM accumulator slots (diagonals) per pass: 4 x 128 = all 512 columns
constexpr int OZ_MAX_ENT = OZ_ACC * OZ_MAX_DIGITS;  // MMAs per k block of one pass
constexpr int OZ_THREADS = 640;       // warp 0 bulk copies, warp 1 MMA, (2, 3 idle), warps 4..19 epilogue
constexpr int OZ_EPI_WARPS = 16;
constexpr int OZ_EPI_LD = 9;
constexpr int OZ_EPI_STAGE = 32 * OZ_EPI_LD;  // doubles per warp (32 rows x 8 columns, padded)
constexpr int OZ_ROWDATA = 32 * (MAX_D + 1);  // doubles per warp (covariance mode)
constexpr size_t OZ_SMEM = 1024 + (size_t)OZ_STAGES * OZ_STAGE_CAP +
                           (size_t)OZ_EPI_WARPS * (OZ_EPI_STAGE + OZ_ROWDATA) * sizeof(double) + 256 +
                           (size_t)4 * OZ_MAX_ENT * 16;  // barriers, slot table, MMA programs of up to 4 passes
constexpr int OZ_TMEM_COLS = 512;
static_assert(OZ_SMEM <= 227 * 1024, "shared memory budget");
static_assert(2 * OZ_MAX_DIGITS * OZ_BLK_BYTES <= OZ_STAGE_CAP, "a stage must hold one k block of all digits");

struct OzakiArgs {
  int NS, NG;
  int KP;             // digits per row (bytes), multiple of 32
  const unsigned char* a_digits;  // [row block][k block][digit][4 KB block image]
  const unsigned char* b_digits;
  int m_tiles, n_tiles;
  const double* sa;   // [rows] power-of-two row scale of A
  const double* sb;   // [cols] power-of-two row scale of B
  int cov;            // 1: covariance epilogue, 0: plain store
  CovEpilogue ep;
  double* D;          // store mode: D = Cin + alpha * A B^T (Cin == nullptr: D = alpha * A B^T)
  int ldd;
  const double* Cin;
  int ldc;
  double alpha;
  int M, N;           // valid rows / cols (store mode)
  int max_kps;        // k blocks per stage (tuning knob)
  int slot_wait;      // wait per accumulator slot (1) or for all slots before the first MMA (0)
  int b_unsigned;     // (name kept from round 1) 1: issue adjacent B digit planes as merged N = 256 MMAs
  int dbg;            // measurement only: 1 = no operand copies (MMAs run on stale smem), 2 = no TMEM drain
};

__device__ __forceinline__ unsigned s_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void bar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void bar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "W_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra D_%=;\n"
      "bra W_%=;\n"
      "D_%=:\n"
      "}\n" ::"r"(s_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(s_u32(dst)),
               "l"(src), "r"(bytes), "r"(s_u32(bar))
               : "memory");
}
// one lane of a converged warp (ptxas knows the predicate of elect.sync selects a single lane and
// issues the tcgen05 instructions under it without a per-lane serialisation loop)
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_i8(unsigned d_tmem, unsigned long long adesc, unsigned long long bdesc, unsigned idesc,
                                        unsigned accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major operand block of 128 rows x 32 bytes in the canonical NO-SWIZZLE layout: 8 x 16-byte core
// matrices (128 contiguous bytes), ordered [row group][k chunk]: K-adjacent core matrices are
// 128 B apart (LBO), row-group-adjacent ones 256 B (SBO).  The digit planes are stored in global
// memory in exactly this image, so a block arrives with one linear bulk copy.
__device__ __forceinline__ unsigned long long umma_desc(unsigned smem_addr) {
  unsigned long long d = 0;
  d |= (unsigned long long)((smem_addr & 0x3FFFF) >> 4);
  d |= (unsigned long long)(128 >> 4) << 16;
  d |= (unsigned long long)(256 >> 4) << 32;
  d |= (unsigned long long)1 << 46;
  return d;
}
// c_format S32 (2) | a_format | b_format (1 = signed, 0 = unsigned) | K-major both | N >> 3 | M >> 4
__device__ __forceinline__ unsigned umma_idesc(int a_signed, int b_signed, int n = OZ_BN) {
  return (2u << 4) | ((unsigned)a_signed << 7) | ((unsigned)b_signed << 10) | ((unsigned)(n >> 3) << 17) |
         ((unsigned)(OZ_BM >> 4) << 24);
}

#define OZ_TMEM_LD32(r, taddr)                                                                                         \
  asm volatile(                                                                                                        \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                        \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                        \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"                      \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),    \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),         \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),        \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                      \
      : "r"(taddr))

// exact int32 -> fp64 without the (slow, XU-pipe) I2F.F64: 2^52 + 2^31 + v has the integer in its
// low mantissa word; one DADD removes the offset
__device__ __forceinline__ double i32_to_f64(int v) {
  return __hiloint2double(0x43300000, v ^ 0x80000000) - 4503601774854144.0;
}

// All MMAs of one k block of pass P, unrolled at compile time.  Accumulator slot a holds the
// diagonal g_hi - a.  On the first k block of a pass the issuing thread waits per slot until the
// epilogue has drained it (so the drain of the previous pass overlaps the first MMAs), on the last
// one it publishes each slot as soon as its MMAs are queued.
template <int NS, int NG, int P, bool FIRST, bool LAST, bool MERGE>
__device__ __forceinline__ void issue_pass_k(unsigned tmem_base, unsigned long long base, unsigned long long* tfull,
                                             unsigned long long* tempty, unsigned par) {
  constexpr int g_hi = NG - 1 - P * OZ_ACC;
  constexpr int g_lo = g_hi - OZ_ACC + 1 > 0 ? g_hi - OZ_ACC + 1 : 0;
  constexpr int nd = (g_hi < NS - 1 ? g_hi : NS - 1) + 1;  // digits 0 .. nd-1 of both operands are staged
  // accumulator slot a holds the diagonal g_lo + a
  if (FIRST || LAST) {
    // Seam of two passes.  The epilogue drains the slots in the order a = OZ_ACC-1 .. 0 and needs
    // ~800 cycles per slot; waiting for all four before the first MMA (and publishing all four after
    // the last) left the tensor pipe idle for a full drain per pass (~15 % of the kernel).  Here
    // the last k block finishes slot by slot in drain order and publishes each slot at once, and
    // the first k block of the next pass restarts each slot as soon as IT has been drained.
#pragma unroll
    for (int a = OZ_ACC - 1; a >= 0; --a) {
      const int g = g_lo + a;
      if (FIRST) {
        bar_wait(&tempty[a], par ^ 1);
        tc_fence_after();
      }
      if (g <= g_hi) {
        const int ilo = g - NS + 1 > 0 ? g - NS + 1 : 0, ihi = g < NS - 1 ? g : NS - 1;
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          if (i >= ilo && i <= ihi) {
            const int j = g - i;
            umma_i8(tmem_base + (unsigned)(a * OZ_BN), base + (unsigned long long)((i * OZ_BLK_BYTES) >> 4),
                    base + (unsigned long long)(((nd + j) * OZ_BLK_BYTES) >> 4), umma_idesc(1, 1),
                    (i > ilo || !FIRST) ? 1u : 0u);
          }
        }
      }
      if (LAST) umma_commit(&tfull[a]);
    }
    return;
  }
  if (MERGE) {
    // All digit planes have the same format, so for a fixed A digit i the products with two
    // consecutive B digits j, j+1 (diagonals i+j, i+j+1 = adjacent slots; the two digit blocks are
    // adjacent in the stage) go out as ONE M128 N256 K32 instruction: 16 instead of 28 MMAs per
    // k block.  The tensor pipe needs ~98 cycles per N = 128 instruction where the arithmetic
    // takes 64 (measured with operand copies and TMEM drains switched off), i.e. there is a
    // fixed cost per instruction that the wider shape amortises.
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      const int jlo = g_lo - i > 0 ? g_lo - i : 0, jhi = g_hi - i < NS - 1 ? g_hi - i : NS - 1;
#pragma unroll
      for (int t = 0; t < (NS + 1) / 2; ++t) {
        const int j = jlo + 2 * t;
        if (j <= jhi) {
          const bool wide = j + 1 <= jhi;
          umma_i8(tmem_base + (unsigned)((i + j - g_lo) * OZ_BN), base + (unsigned long long)((i * OZ_BLK_BYTES) >> 4),
                  base + (unsigned long long)(((nd + j) * OZ_BLK_BYTES) >> 4), umma_idesc(1, 1, wide ? 2 * OZ_BN : OZ_BN),
                  1u);
        }
      }
    }
  } else {
    // one N = 128 instruction per digit pair, round-robin over the slots
#pragma unroll
    for (int t = 0; t < NS; ++t) {
#pragma unroll
      for (int a = 0; a < OZ_ACC; ++a) {
        const int g = g_lo + a;
        if (g <= g_hi) {
          const int ilo = g - NS + 1 > 0 ? g - NS + 1 : 0, ihi = g < NS - 1 ? g : NS - 1;
          const int i = ilo + t;
          if (i <= ihi) {
            const int j = g - i;
            umma_i8(tmem_base + (unsigned)(a * OZ_BN), base + (unsigned long long)((i * OZ_BLK_BYTES) >> 4),
                    base + (unsigned long long)(((nd + j) * OZ_BLK_BYTES) >> 4), umma_idesc(1, 1),
                    1u);
          }
        }
      }
    }
  }
}

// (the k block position is a template parameter so that the steady-state instantiation is a
// straight run of MMAs whose descriptors stay in uniform registers)
template <int NS, int NG, int P, bool MERGE>
__device__ __forceinline__ void issue_pass_m(unsigned tmem_base, unsigned long long base, bool first_k, bool last_k,
                                             unsigned long long* tfull, unsigned long long* tempty, unsigned par) {
  if (first_k) {
    if (last_k) issue_pass_k<NS, NG, P, true, true, MERGE>(tmem_base, base, tfull, tempty, par);
    else issue_pass_k<NS, NG, P, true, false, MERGE>(tmem_base, base, tfull, tempty, par);
  } else {
    if (last_k) issue_pass_k<NS, NG, P, false, true, MERGE>(tmem_base, base, tfull, tempty, par);
    else issue_pass_k<NS, NG, P, false, false, MERGE>(tmem_base, base, tfull, tempty, par);
  }
}
template <int NS, int NG, int P>
__device__ __forceinline__ void issue_pass(unsigned tmem_base, unsigned long long base, bool first_k, bool last_k,
                                           unsigned long long* tfull, unsigned long long* tempty, unsigned par,
                                           bool merge) {
  if (merge) issue_pass_m<NS, NG, P, true>(tmem_base, base, first_k, last_k, tfull, tempty, par);
  else issue_pass_m<NS, NG, P, false>(tmem_base, base, first_k, last_k, tfull, tempty, par);
}

// Final epilogue of one warp's 32 x 32 block in covariance mode: apply the row / column scales,
// evaluate the kernel term and write Z.  The products go through a shared-memory transpose 8
// columns at a time so that lanes map to (row mod 4, column): four 64-byte row segments per store.
template <int D>
__device__ __forceinline__ void cov_tail(const double (&acc)[32], double sa_r, const OzakiArgs& args, double* my_stage,
                                         double* my_rows, int row0, int col0, int lane) {
  const CovEpilogue& ep = args.ep;
  const int hr = lane >> 3, hc = lane & 7;
  for (int e = lane; e < 32 * D; e += 32) {
    const int r = e / D, k = e - r * D;
    my_rows[r * (MAX_D + 1) + k] = (row0 + r < ep.C) ? ep.xs[(size_t)(row0 + r) * D + k] : 0.0;
  }
  my_rows[lane * (MAX_D + 1) + MAX_D] = (row0 + lane < ep.C) ? ep.ystd2 / ep.sd[row0 + lane] : 0.0;
  const int kind = ep.kind;
  const double os = ep.outputscale;
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 8; ++c) my_stage[lane * OZ_EPI_LD + c] = acc[h * 8 + c] * sa_r;
    __syncwarp();
    const int col = col0 + h * 8 + hc;
    const double sb_c = args.sb[col];
    double xc[D];
#pragma unroll
    for (int k = 0; k < D; ++k) xc[k] = ep.xd_s[(size_t)col * D + k];
    double* zrow = ep.Z + (size_t)(row0 + hr) * ep.ldz + col;
    const bool col_ok = col < ep.N;
#pragma unroll 4
    for (int it = 0; it < 8; ++it) {
      const int r = 4 * it + hr;
      double sq = 0.0;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const double df = my_rows[r * (MAX_D + 1) + k] - xc[k];
        sq = fma(df, df, sq);
      }
      const double v = my_stage[r * OZ_EPI_LD + hc] * sb_c;
      const double z = (stationary_from_sq(kind, os, sq) - v) * my_rows[r * (MAX_D + 1) + MAX_D];
      if (col_ok && row0 + r < ep.C) zrow[(size_t)(4 * it) * ep.ldz] = z;
    }
  }
  __syncwarp();
}

// product mode: scale and store the warp's 32 x 32 block through the 8-column transpose
__device__ __forceinline__ void store_tail(const double (&acc)[32], double sa_r, const OzakiArgs& args, double* my_stage,
                                           int row0, int col0, int lane) {
  const int hr = lane >> 3, hc = lane & 7;
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 8; ++c) my_stage[lane * OZ_EPI_LD + c] = acc[h * 8 + c] * sa_r;
    __syncwarp();
    const int col = col0 + h * 8 + hc;
    const double sb_c = args.sb[col];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int r = 4 * it + hr;
      if (row0 + r < args.M && col < args.N) {
        double v = my_stage[r * OZ_EPI_LD + hc] * sb_c;
        if (args.Cin != nullptr) v = fma(args.alpha, v, args.Cin[(size_t)(row0 + r) * args.ldc + col]);
        args.D[(size_t)(row0 + r) * args.ldd + col] = v;
      }
    }
  }
  __syncwarp();
}

__global__ void __launch_bounds__(OZ_THREADS, 1)
ozaki_kernel(const OzakiArgs args) {
  extern __shared__ unsigned char oz_smem_raw[];
  // (offset arithmetic on the __shared__ array itself, so that the compiler keeps the shared address
  // space and emits LDS / STS instead of generic loads)
  unsigned char* smem = oz_smem_raw + ((1024u - (s_u32(oz_smem_raw) & 1023u)) & 1023u);
  const int NS = args.NS, NG = args.NG;
  unsigned char* s_stage = smem;
  double* s_epi = reinterpret_cast<double*>(smem + (size_t)OZ_STAGES * OZ_STAGE_CAP);
  double* s_rowdata = s_epi + OZ_EPI_WARPS * OZ_EPI_STAGE;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(s_rowdata + OZ_EPI_WARPS * OZ_ROWDATA);
  unsigned long long* full = bars;               // [OZ_STAGES]
  unsigned long long* empty = bars + OZ_STAGES;  // [OZ_STAGES]
  unsigned long long* tfull = bars + 2 * OZ_STAGES;  // [OZ_ACC]
  unsigned long long* tempty = tfull + OZ_ACC;       // [OZ_ACC]
  unsigned* s_tmem = reinterpret_cast<unsigned*>(tempty + OZ_ACC);
  int* s_slotend = reinterpret_cast<int*>(s_tmem + 2);  // [4 passes][OZ_ACC] end of each slot's program entries
  uint4* s_prog = reinterpret_cast<uint4*>(bars + 32);  // [4][OZ_MAX_ENT]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = args.KP / OZ_KB;
  const int n_tiles_total = args.m_tiles * args.n_tiles;
  const int n_pass = (NG + OZ_ACC - 1) / OZ_ACC;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < OZ_STAGES; ++s) {
      bar_init(&full[s], 1);
      bar_init(&empty[s], 1);
    }
    for (int a = 0; a < OZ_ACC; ++a) {
      bar_init(&tfull[a], 1);
      bar_init(&tempty[a], OZ_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 2 && lane == 0) {
    // generic digit configurations: the MMA "program" of every pass (descriptor offsets, instruction
    // descriptor, accumulate flag), grouped by accumulator slot
    for (int p = 0; p < n_pass; ++p) {
      const int g_hi = NG - 1 - p * OZ_ACC;
      const int g_lo = g_hi - OZ_ACC + 1 > 0 ? g_hi - OZ_ACC + 1 : 0;
      const int nd = (g_hi < NS - 1 ? g_hi : NS - 1) + 1;
      int ne = 0;
      for (int a = 0; a < OZ_ACC; ++a) {
        const int g = g_lo + a;
        if (g <= g_hi) {
          const int ilo = g - NS + 1 > 0 ? g - NS + 1 : 0, ihi = g < NS - 1 ? g : NS - 1;
          for (int i = ilo; i <= ihi; ++i, ++ne) {
            const int j = g - i;
            s_prog[p * OZ_MAX_ENT + ne] = make_uint4((unsigned)(i * OZ_BLK_BYTES) >> 4, (unsigned)((nd + j) * OZ_BLK_BYTES) >> 4,
                                                     umma_idesc(1, 1), i > ilo ? 1u : 0u);
          }
        }
        s_slotend[p * OZ_ACC + a] = ne;
      }
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(s_u32(s_tmem)), "n"(OZ_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = *s_tmem;

  // Pass p folds the diagonals g_hi(p) .. g_lo(p) (at most OZ_ACC of them, one TMEM accumulator
  // slot each), lowest weights first.  It needs the digits 0 .. nd-1 of both operands; a stage
  // holds as many k blocks of them as fit in OZ_STAGE_CAP (two in the second pass of the default
  // configuration, which keeps the short stages of that pass ahead of the copy latency).
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;\n");
    if (warp == 0 && lane == 0 && !(args.dbg & 1)) {
      // ===== bulk-copy producer =====
      int s = 0;
      unsigned ph = 0;
      for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x) {
        const int m_blk = tile % args.m_tiles, n_blk = tile / args.m_tiles;
        for (int p = 0; p < n_pass; ++p) {
          const int g_hi = NG - 1 - p * OZ_ACC;
          const int nd = (g_hi < NS - 1 ? g_hi : NS - 1) + 1;
          const unsigned kbytes = (unsigned)(nd * OZ_BLK_BYTES);  // per operand per k block
          const int kps = (OZ_STAGE_CAP / (2 * (int)kbytes)) < args.max_kps ? (OZ_STAGE_CAP / (2 * (int)kbytes)) : args.max_kps;
          for (int kb = 0; kb < KB; kb += kps) {
            const int nk = KB - kb < kps ? KB - kb : kps;
            bar_wait(&empty[s], ph ^ 1);
            bar_expect_tx(&full[s], 2u * kbytes * nk);
            unsigned char* st = s_stage + (size_t)s * OZ_STAGE_CAP;
            for (int kk = 0; kk < nk; ++kk) {
              // digits 0 .. nd-1 of one (row block, k block) are one contiguous run in global memory
              bulk_g2s(st + (size_t)kk * 2 * kbytes, args.a_digits + ((size_t)(m_blk * KB + kb + kk) * NS) * OZ_BLK_BYTES,
                       kbytes, &full[s]);
              bulk_g2s(st + (size_t)kk * 2 * kbytes + kbytes,
                       args.b_digits + ((size_t)(n_blk * KB + kb + kk) * NS) * OZ_BLK_BYTES, kbytes, &full[s]);
            }
            if (++s == OZ_STAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    } else if (warp == 1) {
      // ===== MMA issuer (the whole warp runs the loops, one elected lane issues) =====
      int s = 0;
      unsigned ph = 0;
      unsigned pcount = 0;
      const bool std_cfg = NS == OZ_DEFAULT_DIGITS && NG == OZ_DEFAULT_DIAGONALS;
      const bool fast_cfg = NS == 4 && NG == 4;
      for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x) {
        for (int p = 0; p < n_pass; ++p, ++pcount) {
          const unsigned par = pcount & 1;
          const int g_hi = NG - 1 - p * OZ_ACC;
          const int nd = (g_hi < NS - 1 ? g_hi : NS - 1) + 1;
          const unsigned kbytes = (unsigned)(nd * OZ_BLK_BYTES);
          const int kps = (OZ_STAGE_CAP / (2 * (int)kbytes)) < args.max_kps ? (OZ_STAGE_CAP / (2 * (int)kbytes)) : args.max_kps;
          const uint4* prog = s_prog + p * OZ_MAX_ENT;
          const int* slotend = s_slotend + p * OZ_ACC;
          for (int kb = 0; kb < KB; kb += kps) {
            const int nk = KB - kb < kps ? KB - kb : kps;
            if (!(args.dbg & 1)) bar_wait(&full[s], ph);
            tc_fence_after();
            const unsigned st_addr = s_u32(s_stage + (size_t)s * OZ_STAGE_CAP);
            if (elect_one()) {
            for (int kk = 0; kk < nk; ++kk) {
              const unsigned long long base = umma_desc(st_addr + kk * 2 * kbytes);
              const bool first_k = kb + kk == 0, last_k = kb + kk == KB - 1;
              if (std_cfg) {
                // default digit configuration: offsets and instruction descriptors are immediates
                // (~10 uniform-datapath instructions per MMA; through the table the issuing thread
                // needed ~23 and the tensor pipe waited on it)
                if (p == 0)
                  issue_pass<OZ_DEFAULT_DIGITS, OZ_DEFAULT_DIAGONALS, 0>(tmem_base, base, first_k, last_k, tfull, tempty, par, args.b_unsigned != 0);
                else
                  issue_pass<OZ_DEFAULT_DIGITS, OZ_DEFAULT_DIAGONALS, 1>(tmem_base, base, first_k, last_k, tfull, tempty, par, args.b_unsigned != 0);
              } else if (fast_cfg) {
                // reduced-precision configuration (DKG_PLAN_FAST32): 4 digits, 4 diagonals, one pass
                issue_pass<4, 4, 0>(tmem_base, base, first_k, last_k, tfull, tempty, par, args.b_unsigned != 0);
              } else {
                int e = 0;
                for (int a = 0; a < OZ_ACC; ++a) {
                  if (first_k) {
                    bar_wait(&tempty[a], par ^ 1);
                    tc_fence_after();
                  }
                  for (; e < slotend[a]; ++e) {
                    const uint4 en = prog[e];
                    umma_i8(tmem_base + (unsigned)(a * OZ_BN), base + en.x, base + en.y, en.z, (first_k ? 0u : 1u) | en.w);
                  }
                  if (last_k) umma_commit(&tfull[a]);
                }
              }
            }
            if (!(args.dbg & 1)) umma_commit(&empty[s]);
            }
            __syncwarp();
            if (++s == OZ_STAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;\n");
    // ===== epilogue warps =====
    const int ew = warp - 4;
    const int q = warp & 3;    // TMEM lane quarter this warp may read
    const int cg = ew >> 2;    // which 32 of the 128 accumulator columns
    double* my_stage = s_epi + ew * OZ_EPI_STAGE;
    double* my_rows = s_rowdata + ew * OZ_ROWDATA;
    unsigned pcount = 0;
    for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x) {
      const int m_blk = tile % args.m_tiles, n_blk = tile / args.m_tiles;
      double acc[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) acc[c] = 0.0;
      for (int p = 0; p < n_pass; ++p, ++pcount) {
        const unsigned par = pcount & 1;
        const int g_hi = NG - 1 - p * OZ_ACC;
        const int g_lo = g_hi - OZ_ACC + 1 > 0 ? g_hi - OZ_ACC + 1 : 0;
        double w = 1.0;
        for (int g = 0; g < g_hi; ++g) w *= 0.00390625;  // 256^-g_hi
        const unsigned taddr = tmem_base + ((unsigned)(q * 32) << 16) + cg * 32;
        for (int a = OZ_ACC - 1; a >= 0; --a) {  // slot a = diagonal g_lo + a; smallest weights first
          bar_wait(&tfull[a], par);
          tc_fence_after();
          if (g_lo + a <= g_hi) {
            if (!(args.dbg & 2)) {
              int r[32];
              OZ_TMEM_LD32(r, taddr + (unsigned)(a * OZ_BN));
              asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
              for (int c = 0; c < 32; ++c) acc[c] = fma(i32_to_f64(r[c]), w, acc[c]);
            }
            w *= 256.0;
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) bar_arrive(&tempty[a]);
        }
      }
      // ---- final epilogue ----
      const int row0 = m_blk * OZ_BM + q * 32;
      const int col0 = n_blk * OZ_BN + cg * 32;
      const double sa_r = args.sa[row0 + lane];
      if (args.cov) {
        switch (args.ep.d) {
          case 1: cov_tail<1>(acc, sa_r, args, my_stage, my_rows, row0, col0, lane); break;
          case 2: cov_tail<2>(acc, sa_r, args, my_stage, my_rows, row0, col0, lane); break;
          case 3: cov_tail<3>(acc, sa_r, args, my_stage, my_rows, row0, col0, lane); break;
          case 4: cov_tail<4>(acc, sa_r, args, my_stage, my_rows, row0, col0, lane); break;
          case 5: cov_tail<5>(acc, sa_r, args, my_stage, my_rows, row0, col0, lane); break;
          case 6: cov_tail<6>(acc, sa_r, args, my_stage, my_rows, row0, col0, lane); break;
          case 7: cov_tail<7>(acc, sa_r, args, my_stage, my_rows, row0, col0, lane); break;
          default: cov_tail<8>(acc, sa_r, args, my_stage, my_rows, row0, col0, lane); break;
        }
      } else {
        store_tail(acc, sa_r, args, my_stage, row0, col0, lane);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(OZ_TMEM_COLS));
  }
}

// ------------------------------------------------------------------------------------------
// CTA-pair variant (cluster of 2, tcgen05.mma.cta_group::2, M = 256 across the two SMs).
// With N = 128 a single-CTA MMA reads 8 KB of operands per 64 cycles - all of the 128 B/clk of
// shared-memory bandwidth, which it shares with the bulk copies and the epilogue.  In pair mode
// each SM feeds its own 128 rows of A and only HALF of B (64 rows, 2 KB per digit): 96 B/clk.
//   rank 0 (leader): issues every MMA; its tcgen05.commit multicasts to both CTAs' barriers.
//   rank 1         : warp 1 relays "my stage has landed" to the leader (remote mbarrier arrive),
//                    since a linear bulk copy can only signal a barrier of its own CTA.
//   both           : producer (own A rows + own B half), 16 epilogue warps draining the local
//                    TMEM; slot releases (tempty) are remote arrivals on the leader's barriers.
// Default digit configuration only (7 digits, 8 diagonals: two passes of four slots).
// ------------------------------------------------------------------------------------------
constexpr int OZP_STAGES = 3;
constexpr int OZP_STAGE_CAP = 49152;  // one k block of 7 x (4 KB + 2 KB), or two of 4 x (4 KB + 2 KB)
constexpr int OZP_HALF_BYTES = OZ_BLK_BYTES / 2;
constexpr size_t OZP_SMEM = 1024 + (size_t)OZP_STAGES * OZP_STAGE_CAP +
                            (size_t)OZ_EPI_WARPS * (OZ_EPI_STAGE + OZ_ROWDATA) * sizeof(double) + 256;
static_assert(OZP_SMEM <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned map_to_rank(unsigned smem_addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void bar_arrive_cluster(unsigned cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void bar_wait_cluster(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "W_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra D_%=;\n"
      "bra W_%=;\n"
      "D_%=:\n"
      "}\n" ::"r"(s_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(unsigned long long* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(s_u32(bar)),
      "h"((unsigned short)3)
      : "memory");
}
__device__ __forceinline__ void umma_i8_pair(unsigned d_tmem, unsigned long long adesc, unsigned long long bdesc,
                                             unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ unsigned umma_idesc_pair(int a_signed, int b_signed) {
  return (2u << 4) | ((unsigned)a_signed << 7) | ((unsigned)b_signed << 10) | ((unsigned)(OZ_BN >> 3) << 17) |
         ((unsigned)(256 >> 4) << 24);
}

template <int NS, int NG, int P, bool FIRST, bool LAST>
__device__ __forceinline__ void issue_pass_pair_k(unsigned tmem_base, unsigned long long base, unsigned long long* tfull,
                                                  unsigned long long* tempty, unsigned par) {
  constexpr int g_hi = NG - 1 - P * OZ_ACC;
  constexpr int g_lo = g_hi - OZ_ACC + 1 > 0 ? g_hi - OZ_ACC + 1 : 0;
  constexpr int nd = (g_hi < NS - 1 ? g_hi : NS - 1) + 1;
#pragma unroll
  for (int a = 0; a < OZ_ACC; ++a) {
    const int g = g_hi - a;
    if (FIRST) {
      bar_wait_cluster(&tempty[a], par ^ 1);
      tc_fence_after();
    }
    if (g >= g_lo) {
      const int ilo = g - NS + 1 > 0 ? g - NS + 1 : 0, ihi = g < NS - 1 ? g : NS - 1;
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        if (i >= ilo && i <= ihi) {
          const int j = g - i;
          umma_i8_pair(tmem_base + (unsigned)(a * OZ_BN), base + (unsigned long long)((i * OZ_BLK_BYTES) >> 4),
                       base + (unsigned long long)((nd * OZ_BLK_BYTES + j * OZP_HALF_BYTES) >> 4),
                       umma_idesc_pair(1, 1), (i > ilo || !FIRST) ? 1u : 0u);
        }
      }
    }
    if (LAST) umma_commit_pair(&tfull[a]);
  }
}
template <int NS, int NG, int P>
__device__ __forceinline__ void issue_pass_pair(unsigned tmem_base, unsigned long long base, bool first_k, bool last_k,
                                                unsigned long long* tfull, unsigned long long* tempty, unsigned par) {
  if (first_k) {
    if (last_k) issue_pass_pair_k<NS, NG, P, true, true>(tmem_base, base, tfull, tempty, par);
    else issue_pass_pair_k<NS, NG, P, true, false>(tmem_base, base, tfull, tempty, par);
  } else {
    if (last_k) issue_pass_pair_k<NS, NG, P, false, true>(tmem_base, base, tfull, tempty, par);
    else issue_pass_pair_k<NS, NG, P, false, false>(tmem_base, base, tfull, tempty, par);
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(OZ_THREADS, 1)
ozaki_pair_kernel(const OzakiArgs args) {
  constexpr int NS = OZ_DEFAULT_DIGITS, NG = OZ_DEFAULT_DIAGONALS;
  constexpr int n_pass = (NG + OZ_ACC - 1) / OZ_ACC;
  extern __shared__ unsigned char oz_smem_raw[];
  unsigned char* smem = oz_smem_raw + ((1024u - (s_u32(oz_smem_raw) & 1023u)) & 1023u);
  unsigned char* s_stage = smem;
  double* s_epi = reinterpret_cast<double*>(smem + (size_t)OZP_STAGES * OZP_STAGE_CAP);
  double* s_rowdata = s_epi + OZ_EPI_WARPS * OZ_EPI_STAGE;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(s_rowdata + OZ_EPI_WARPS * OZ_ROWDATA);
  unsigned long long* full = bars;                    // [OZP_STAGES] own stage landed
  unsigned long long* pfull = full + OZP_STAGES;      // [OZP_STAGES] (leader) peer's stage landed
  unsigned long long* empty = pfull + OZP_STAGES;     // [OZP_STAGES] stage consumed (multicast commit)
  unsigned long long* tfull = empty + OZP_STAGES;     // [OZ_ACC] slot ready (multicast commit)
  unsigned long long* tempty = tfull + OZ_ACC;        // [OZ_ACC] (leader) slot drained by both CTAs
  unsigned* s_tmem = reinterpret_cast<unsigned*>(tempty + OZ_ACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned rank = cluster_ctarank();
  const int KB = args.KP / OZ_KB;
  const int m_pairs = (args.m_tiles + 1) / 2;
  const int n_tiles_total = m_pairs * args.n_tiles;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < OZP_STAGES; ++s) {
      bar_init(&full[s], 1);
      bar_init(&pfull[s], 1);
      bar_init(&empty[s], 1);
    }
    for (int a = 0; a < OZ_ACC; ++a) {
      bar_init(&tfull[a], 1);
      bar_init(&tempty[a], 2 * OZ_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(s_u32(s_tmem)), "n"(OZ_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const unsigned tmem_base = *s_tmem;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;\n");
    if (warp == 0 && lane == 0) {
      // ===== producer: own 128 rows of A, own 64-row half of B =====
      int s = 0;
      unsigned ph = 0;
      for (int tile = cluster_id; tile < n_tiles_total; tile += n_clusters) {
        const int m_blk = (tile % m_pairs) * 2 + (int)rank, n_half = (tile / m_pairs) * 2 + (int)rank;
        for (int p = 0; p < n_pass; ++p) {
          const int g_hi = NG - 1 - p * OZ_ACC;
          const int nd = (g_hi < NS - 1 ? g_hi : NS - 1) + 1;
          const unsigned abytes = (unsigned)(nd * OZ_BLK_BYTES), bbytes = (unsigned)(nd * OZP_HALF_BYTES);
          const int kps = OZP_STAGE_CAP / (int)(abytes + bbytes) < OZ_MAX_KPS ? OZP_STAGE_CAP / (int)(abytes + bbytes) : OZ_MAX_KPS;
          for (int kb = 0; kb < KB; kb += kps) {
            const int nk = KB - kb < kps ? KB - kb : kps;
            bar_wait_cluster(&empty[s], ph ^ 1);
            bar_expect_tx(&full[s], (abytes + bbytes) * nk);
            unsigned char* st = s_stage + (size_t)s * OZP_STAGE_CAP;
            for (int kk = 0; kk < nk; ++kk) {
              bulk_g2s(st + (size_t)kk * (abytes + bbytes), args.a_digits + ((size_t)(m_blk * KB + kb + kk) * NS) * OZ_BLK_BYTES,
                       abytes, &full[s]);
              bulk_g2s(st + (size_t)kk * (abytes + bbytes) + abytes,
                       args.b_digits + ((size_t)(n_half * KB + kb + kk) * NS) * OZP_HALF_BYTES, bbytes, &full[s]);
            }
            if (++s == OZP_STAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    } else if (warp == 1) {
      int s = 0;
      unsigned ph = 0;
      unsigned pcount = 0;
      if (rank == 0) {
        // ===== leader: MMA issuer for the pair =====
        for (int tile = cluster_id; tile < n_tiles_total; tile += n_clusters) {
          for (int p = 0; p < n_pass; ++p, ++pcount) {
            const unsigned par = pcount & 1;
            const int g_hi = NG - 1 - p * OZ_ACC;
            const int nd = (g_hi < NS - 1 ? g_hi : NS - 1) + 1;
            const unsigned kbytes = (unsigned)(nd * (OZ_BLK_BYTES + OZP_HALF_BYTES));
            const int kps = OZP_STAGE_CAP / (int)kbytes < OZ_MAX_KPS ? OZP_STAGE_CAP / (int)kbytes : OZ_MAX_KPS;
            for (int kb = 0; kb < KB; kb += kps) {
              const int nk = KB - kb < kps ? KB - kb : kps;
              bar_wait(&full[s], ph);
              bar_wait_cluster(&pfull[s], ph);
              tc_fence_after();
              if (elect_one()) {
                const unsigned st_addr = s_u32(s_stage + (size_t)s * OZP_STAGE_CAP);
                for (int kk = 0; kk < nk; ++kk) {
                  const unsigned long long base = umma_desc(st_addr + kk * kbytes);
                  const bool first_k = kb + kk == 0, last_k = kb + kk == KB - 1;
                  if (p == 0) issue_pass_pair<NS, NG, 0>(tmem_base, base, first_k, last_k, tfull, tempty, par);
                  else issue_pass_pair<NS, NG, 1>(tmem_base, base, first_k, last_k, tfull, tempty, par);
                }
                umma_commit_pair(&empty[s]);
              }
              __syncwarp();
              if (++s == OZP_STAGES) { s = 0; ph ^= 1; }
            }
          }
        }
      } else if (lane == 0) {
        // ===== peer: tell the leader when this CTA's stage has landed =====
        for (int tile = cluster_id; tile < n_tiles_total; tile += n_clusters) {
          for (int p = 0; p < n_pass; ++p) {
            const int g_hi = NG - 1 - p * OZ_ACC;
            const int nd = (g_hi < NS - 1 ? g_hi : NS - 1) + 1;
            const unsigned kbytes = (unsigned)(nd * (OZ_BLK_BYTES + OZP_HALF_BYTES));
            const int kps = OZP_STAGE_CAP / (int)kbytes < OZ_MAX_KPS ? OZP_STAGE_CAP / (int)kbytes : OZ_MAX_KPS;
            for (int kb = 0; kb < KB; kb += kps) {
              bar_wait(&full[s], ph);
              bar_arrive_cluster(map_to_rank(s_u32(&pfull[s]), 0));
              if (++s == OZP_STAGES) { s = 0; ph ^= 1; }
            }
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;\n");
    // ===== epilogue warps (both CTAs drain their own TMEM) =====
    const int ew = warp - 4;
    const int q = warp & 3;
    const int cg = ew >> 2;
    double* my_stage = s_epi + ew * OZ_EPI_STAGE;
    double* my_rows = s_rowdata + ew * OZ_ROWDATA;
    unsigned tempty_leader[OZ_ACC];
#pragma unroll
    for (int a = 0; a < OZ_ACC; ++a) tempty_leader[a] = map_to_rank(s_u32(&tempty[a]), 0);
    unsigned pcount = 0;
    for (int tile = cluster_id; tile < n_tiles_total; tile += n_clusters) {
      const int m_blk = (tile % m_pairs) * 2 + (int)rank, n_blk = tile / m_pairs;
      double acc[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) acc[c] = 0.0;
      for (int p = 0; p < n_pass; ++p, ++pcount) {
        const unsigned par = pcount & 1;
        const int g_hi = NG - 1 - p * OZ_ACC;
        const int g_lo = g_hi - OZ_ACC + 1 > 0 ? g_hi - OZ_ACC + 1 : 0;
        double w = 1.0;
        for (int g = 0; g < g_hi; ++g) w *= 0.00390625;
        const unsigned taddr = tmem_base + ((unsigned)(q * 32) << 16) + cg * 32;
#pragma unroll
        for (int a = 0; a < OZ_ACC; ++a) {
          bar_wait_cluster(&tfull[a], par);
          tc_fence_after();
          if (g_hi - a >= g_lo) {
            int r[32];
            OZ_TMEM_LD32(r, taddr + (unsigned)(a * OZ_BN));
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
            for (int c = 0; c < 32; ++c) acc[c] = fma(i32_to_f64(r[c]), w, acc[c]);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) bar_arrive_cluster(tempty_leader[a]);
          w *= 256.0;
        }
      }
      const int row0 = m_blk * OZ_BM + q * 32;
      const int col0 = n_blk * OZ_BN + cg * 32;
      const double sa_r = args.sa[row0 + lane];
      if (args.cov) {
        switch (args.ep.d) {
          case 1: cov_tail<1>(acc, sa_r, args, my_stage, my_rows, row0, col0, lane); break;
          case 2: cov_tail<2>(acc, sa_r, args, my_stage, my_rows, row0, col0, lane); break;
          case 3: cov_tail<3>(acc, sa_r, args, my_stage, my_rows, row0, col0, lane); break;
          case 4: cov_tail<4>(acc, sa_r, args, my_stage, my_rows, row0, col0, lane); break;
          case 5: cov_tail<5>(acc, sa_r, args, my_stage, my_rows, row0, col0, lane); break;
          case 6: cov_tail<6>(acc, sa_r, args, my_stage, my_rows, row0, col0, lane); break;
          case 7: cov_tail<7>(acc, sa_r, args, my_stage, my_rows, row0, col0, lane); break;
          default: cov_tail<8>(acc, sa_r, args, my_stage, my_rows, row0, col0, lane); break;
        }
      } else {
        store_tail(acc, sa_r, args, my_stage, row0, col0, lane);
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(OZ_TMEM_COLS));
  }
}

// Power-of-two scale of every row into [-127, 127], then NS exact balanced base-256 digits, written as UMMA
// operand images.  A CTA takes 8 consecutive rows (one row group of the core-matrix layout): warp w first finds
// the maximum of row w; then every warp converts 16-column chunks of all 8 rows at once -- lane (row, quarter)
// handles 4 consecutive columns and stores ONE 32-bit word per digit, so a warp writes the 128 contiguous bytes
// of a core matrix per digit (the first version stored single bytes, 32 useful bytes per store instruction, and
// took 21 us for the 4096 x 416 rows of a c4 batch -- five times per forward once T = KX K^-1 runs on int8 too).
constexpr int SL_ROWS = 8;
__global__ void __launch_bounds__(SL_ROWS * 32)
slice_rows_kernel(const double* __restrict__ X, int ld, int rows, int K, int KP, int NS, int block_rows,
                  unsigned char* __restrict__ out, double* __restrict__ scale) {
  __shared__ double s_mul[SL_ROWS];  // 2^(8 (NS - 1) - e7) of the CTA's rows (0: row beyond the matrix / not finite)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r0 = blockIdx.x * SL_ROWS;
  {
    const int r = r0 + warp;
    double m = 0.0;
    if (r < rows) {
      const double* x = X + (size_t)r * ld;
      for (int k = lane; k < K; k += 32) m = fmax(m, fabs(x[k]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    int e = 0;
    if (m > 0.0 && m < 1e300) frexp(m, &e);  // m = f 2^e, f in [0.5, 1)
    int e7 = e - 7;                           // |x| 2^-e7 < 128
    // balanced digits d_s in [-128, 127] represent exactly the fixed-point values in [-127, 127] (the carries
    // of the lower digits can raise the leading digit by one): rows whose maximum lies above 127 give up one bit
    if (scalbn(m, -e7) > 127.0) e7 += 1;
    if (lane == 0) {
      if (r < rows) scale[r] = scalbn(1.0, e7);
      s_mul[warp] = (r < rows && m < 1e300) ? scalbn(1.0, 8 * (NS - 1) - e7) : 0.0;
    }
  }
  __syncthreads();
  // block image: [row block][k block][digit][row group][k chunk (2)][row in group (8)][16 B]
  // with block_rows (128, or 64 for the B halves of the CTA-pair kernel) rows x 32 bytes per block
  const int rl = lane >> 2, kq = lane & 3;  // row of the group, 4-column quarter of the 16-column chunk
  const int r = r0 + rl;
  if (r >= rows) return;
  const double mul = s_mul[rl];
  const double* x = X + (size_t)r * ld;
  const int rb = r / block_rows, ri = r - rb * block_rows;
  for (int kc = warp; kc < (KP >> 4); kc += SL_ROWS) {
    const int k0 = kc * 16 + kq * 4;
    // round to the last digit first (|V| <= 127 * 256^(NS-1) < 2^55; multiplying by a power of two is exact), then
    // peel balanced digits off the low end: d = ((V + 128) mod 256) - 128, V <- (V - d) / 256; what is left after
    // NS - 1 steps is the leading digit
    long long V[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) V[u] = (k0 + u < K) ? __double2ll_rn(x[k0 + u] * mul) : 0ll;
    const size_t blk0 = ((size_t)rb * (KP >> 5) + (k0 >> 5)) * NS;
    const int in_blk = (((ri >> 3) * 2 + ((k0 & 31) >> 4)) * 8 + (ri & 7)) * 16 + (k0 & 15);
    for (int s = NS - 1; s >= 0; --s) {
      unsigned word = 0u;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int dg;
        if (s > 0) {
          dg = (int)((V[u] + 128) & 255) - 128;
          V[u] = (V[u] - dg) >> 8;
        } else {
          dg = (int)V[u];
        }
        word |= ((unsigned)dg & 0xffu) << (8 * u);  // two's complement bytes, little endian = increasing column
      }
      *reinterpret_cast<unsigned*>(out + (blk0 + s) * (size_t)(block_rows * OZ_KB) + in_blk) = word;
    }
  }
}

int ensure_ozaki_attr(int* n_sm) {
  static int done_for_device[64] = {0};
  static int sms[64] = {0};
  int dev = 0;
  DKG_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) dev = 0;
  if (!done_for_device[dev]) {
    DKG_CUDA_OK(cudaFuncSetAttribute(ozaki_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OZ_SMEM));
    DKG_CUDA_OK(cudaFuncSetAttribute(ozaki_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OZP_SMEM));
    DKG_CUDA_OK(cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev));
    done_for_device[dev] = 1;
  }
  *n_sm = sms[dev];
  return DKG_OK;
}

}  // namespace

// DKG_OZ_PAIR=1 selects the CTA-pair kernel for the default digit configuration.  Measured at c4 it
// is no faster than the single-CTA kernel (0.751 vs 0.747 ms), i.e. the shared-memory operand
// bandwidth is not what holds the tensor pipe at ~60 %, so the simpler kernel stays the default.
static bool pair_mode(int NS, int NG) {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DKG_OZ_PAIR");
    v = (e != nullptr && atoi(e) != 0) ? 1 : 0;
  }
  return v != 0 && NS == OZ_DEFAULT_DIGITS && NG == OZ_DEFAULT_DIAGONALS;
}

int ozaki_b_block_rows(int NS, int NG) {
  if (NG == 0) NG = OZ_DEFAULT_DIAGONALS;
  return pair_mode(NS, NG) ? 64 : 128;
}

__global__ void any_negative_kernel(const double* __restrict__ X, int ld, int rows, int cols, int* __restrict__ flag) {
  const long long total = (long long)rows * cols;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e / cols), c = (int)(e - (long long)r * cols);
    if (X[(size_t)r * ld + c] < 0.0) *flag = 1;
  }
}

// flag_dev[0] = 1 if any entry of X[rows, cols] is negative (flag_dev must be zero on entry)
int ozaki_any_negative(const double* X, int ld, int rows, int cols, int* flag_dev, cudaStream_t st) {
  if (rows == 0 || cols == 0) return DKG_OK;
  any_negative_kernel<<<296, 256, 0, st>>>(X, ld, rows, cols, flag_dev);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

int ozaki_kp(int K) { return round_up(K, 32); }

size_t ozaki_digit_bytes(int rows_pad, int K, int NS) { return (size_t)NS * round_up(rows_pad, 256) * ozaki_kp(K); }

// block-image digit planes (see slice_rows_kernel) and scale[rows] of the row-major matrix X[rows, K]
int ozaki_slice_rows(const double* X, int ld, int rows, int K, int block_rows, int NS, unsigned char* digits,
                     double* scale, cudaStream_t st) {
  if (rows == 0) return DKG_OK;
  slice_rows_kernel<<<ceil_div(rows, SL_ROWS), SL_ROWS * 32, 0, st>>>(X, ld, rows, K, ozaki_kp(K), NS, block_rows, digits, scale);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

static int ozaki_launch(const unsigned char* a_digits, const double* sa, int M_pad, const unsigned char* b_digits,
                        const double* sb, int N_pad, int K, int NS, int NG, OzakiArgs& args, cudaStream_t st) {
  int n_sm = 0;
  DKG_TRY(ensure_ozaki_attr(&n_sm));
  const int KP = ozaki_kp(K);
  args.NS = NS;
  args.NG = NG;
  args.KP = KP;
  args.a_digits = a_digits;
  args.b_digits = b_digits;
  args.m_tiles = M_pad / OZ_BM;
  args.n_tiles = N_pad / OZ_BN;
  args.sa = sa;
  args.sb = sb;
  {
    // tuning / measurement switches, read once per process
    static int kps = 0, slot_wait = 1, dbg = 0, merge = 1;
    if (kps == 0) {
      const char* e = getenv("DKG_OZ_KPS");
      int v = e != nullptr ? atoi(e) : OZ_MAX_KPS;
      e = getenv("DKG_OZ_SLOTWAIT");
      slot_wait = e != nullptr ? atoi(e) : 1;
      e = getenv("DKG_OZ_DBG");
      dbg = e != nullptr ? atoi(e) : 0;
      e = getenv("DKG_OZ_MERGE");
      merge = (e != nullptr && atoi(e) == 0) ? 0 : 1;
      kps = v < 1 ? 1 : v > OZ_MAX_KPS ? OZ_MAX_KPS : v;
    }
    args.max_kps = kps;
    args.slot_wait = slot_wait;
    args.dbg |= dbg;
    if (!merge) args.b_unsigned = 0;
  }
  if (pair_mode(NS, NG)) {
    const int tiles = ((args.m_tiles + 1) / 2) * args.n_tiles;
    // how many CTA pairs can be co-resident (SMs pair up within a GPC; an odd SM is left over)
    static int max_clusters[64] = {0};
    int dev = 0;
    DKG_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) dev = 0;
    if (max_clusters[dev] == 0) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(n_sm & ~1);
      cfg.blockDim = dim3(OZ_THREADS);
      cfg.dynamicSmemBytes = OZP_SMEM;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, ozaki_pair_kernel, &cfg) != cudaSuccess || n < 1) {
        cudaGetLastError();
        n = n_sm / 2;
      }
      max_clusters[dev] = n;
    }
    const int clusters = tiles < max_clusters[dev] ? tiles : max_clusters[dev];
    ozaki_pair_kernel<<<2 * clusters, OZ_THREADS, OZP_SMEM, st>>>(args);
    DKG_LAUNCH_CHECK();
    return DKG_OK;
  }
  const int tiles = args.m_tiles * args.n_tiles;
  const int grid = tiles < n_sm ? tiles : n_sm;
  ozaki_kernel<<<grid, OZ_THREADS, OZ_SMEM, st>>>(args);
  DKG_LAUNCH_CHECK();
  return DKG_OK;
}

// Z = (k(x_c, xd_n) - A B^T) ystd^2 / sd from digit planes (see the file header)
int ozaki_cov(const unsigned char* a_digits, const double* sa, int M_pad, const unsigned char* b_digits,
              const double* sb, int N_pad, int K, int NS, int NG, bool b_nonneg, const CovEpilogue& ep,
              cudaStream_t st) {
  OzakiArgs args{};
  (void)b_nonneg;
  args.b_unsigned = 1;  // balanced digits: every plane is signed, adjacent B planes can always be merged (DKG_OZ_MERGE=0 disables)
  args.cov = 1;
  args.ep = ep;
  return ozaki_launch(a_digits, sa, M_pad, b_digits, sb, N_pad, K, NS, NG, args, st);
}

// D[M, N] = A B^T from digit planes (test hook)
int ozaki_store(const unsigned char* a_digits, const double* sa, int M_pad, const unsigned char* b_digits,
                const double* sb, int N_pad, int K, int NS, int NG, bool b_nonneg, double* D, int ldd, int M,
                int N, cudaStream_t st) {
  OzakiArgs args{};
  (void)b_nonneg;
  args.b_unsigned = 1;  // balanced digits: every plane is signed, adjacent B planes can always be merged (DKG_OZ_MERGE=0 disables)
  args.cov = 0;
  args.D = D;
  args.ldd = ldd;
  args.M = M;
  args.N = N;
  return ozaki_launch(a_digits, sa, M_pad, b_digits, sb, N_pad, K, NS, NG, args, st);
}

// D[M, N] = Cin + alpha * A B^T from digit planes (the T = KX K^-1 products and their refinement step)
int ozaki_store_axpy(const unsigned char* a_digits, const double* sa, int M_pad, const unsigned char* b_digits,
                     const double* sb, int N_pad, int K, int NS, int NG, const double* Cin, int ldc, double alpha,
                     double* D, int ldd, int M, int N, cudaStream_t st) {
  OzakiArgs args{};
  args.b_unsigned = 1;
  args.cov = 0;
  args.D = D;
  args.ldd = ldd;
  args.Cin = Cin;
  args.ldc = ldc;
  args.alpha = alpha;
  args.M = M;
  args.N = N;
  return ozaki_launch(a_digits, sa, M_pad, b_digits, sb, N_pad, K, NS, NG, args, st);
}

// Measured int8 tensor peak of THIS kernel's instruction stream: the same launch with the operand
// copies and the TMEM drains switched off (MMAs run on whatever the stages hold; results discarded),
// i.e. the rate at which the tensor pipe retires M128 N128/256 K32 kind::i8 instructions under the
// box's power / clock conditions.  Returns executed int8 TOP/s over `reps` launches (CUDA events).
int ozaki_mma_peak(int M_pad, int N_pad, int K, int NS, int NG, int reps, int mode, double* tops_out, double* ms_out,
                   cudaStream_t st) {
  const int KP = ozaki_kp(K);
  unsigned char* dig = nullptr;
  double* sc = nullptr;
  const size_t bytes = ozaki_digit_bytes(M_pad > N_pad ? M_pad : N_pad, K, NS);
  DKG_CUDA_OK(cudaMalloc((void**)&dig, bytes));
  DKG_CUDA_OK(cudaMalloc((void**)&sc, sizeof(double) * (size_t)((M_pad > N_pad ? M_pad : N_pad) + 256)));
  DKG_CUDA_OK(cudaMemsetAsync(dig, 1, bytes, st));
  DKG_CUDA_OK(cudaMemsetAsync(sc, 0, sizeof(double) * (size_t)((M_pad > N_pad ? M_pad : N_pad) + 256), st));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int rc = DKG_OK;
  for (int it = 0; it < reps + 2 && rc == DKG_OK; ++it) {
    if (it == 2) cudaEventRecord(e0, st);
    OzakiArgs args{};
    args.b_unsigned = 1;
    args.cov = 0;
    args.D = nullptr;
    args.M = 0;  // store_tail writes nothing
    args.N = 0;
    args.ldd = 0;
    args.dbg = mode & 3;  // bit 0: no operand copies, bit 1: no accumulator drains (3 = MMA stream only)
    rc = ozaki_launch(dig, sc, M_pad, dig, sc, N_pad, K, NS, NG, args, st);
  }
  cudaEventRecord(e1, st);
  cudaError_t e = cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(dig);
  cudaFree(sc);
  if (rc != DKG_OK) return rc;
  if (e != cudaSuccess) { set_error("ozaki_mma_peak: %s", cudaGetErrorString(e)); return DKG_ECUDA; }
  int pairs = 0;
  for (int g = 0; g < NG; ++g) pairs += (g < NS - 1 ? g : NS - 1) - (g - NS + 1 > 0 ? g - NS + 1 : 0) + 1;
  const double ops = 2.0 * (double)M_pad * (double)N_pad * (double)KP * pairs * reps;
  if (ms_out) *ms_out = ms / reps;
  if (tops_out) *tops_out = ops / (ms * 1e-3) / 1e12;
  return DKG_OK;
}

}  // namespace dkg
