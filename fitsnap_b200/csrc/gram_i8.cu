// K2+K3+K4 on the 5th-generation tensor cores: exact-integer Gram through int8 tcgen05.mma.
//
// tcgen05.mma has no f64 kind, and the fit needs a Gram that is good to ~1e-15 (cond(G) = cond(Aw)^2),
// so the fp64 contraction  G = [Aw|bw]^T [Aw|bw]  is recast as EXACT integer arithmetic that the int8
// tensor cores can do (Ozaki scheme II: Chinese remainder theorem instead of mantissa slices):
//
//   per slab of <= 2^18 rows
//   1. i8_colmax_kernel   m_c = max_r |fl(w_r a_rc)|  (augmented column c = k holds w_r b_r).  One HBM pass; the maxima of
//                         ALL slabs come from one launch in front of the slab loop (blockIdx.z = slab).
//   2. i8_convert_kernel  q_rc = rint(fl(w_r a_rc) * 2^e_c), e_c = BETA-1-ilogb(m_c), an integer of at most
//                         BETA <= 53 bits: the fp64 value itself when the column maximum has that exponent,
//                         an absolute truncation at m_c 2^-BETA otherwise (what an fp64 dot product keeps of
//                         it anyway).  For 16 pairwise coprime moduli p_t <= 256 it writes the symmetric
//                         residues q mod p_t as int8 planes  R_t[c][r]  (row index contiguous: both MMA
//                         operands become K-major).  q is split into its 7 low bytes + sign and two
//                         dp4a instructions per modulus fold them with the byte weights 256^i mod p_t; the
//                         sum is then reduced with one fp32 fma and one integer multiply-add on the float's bit
//                         pattern (see convert_one_modulus) -- no division, no integer<->float conversion.
//                         No shared memory: a warp owns 4 columns x 128 rows.  (The same code can run as
//                         converter warps inside the tensor-core kernel -- FSB_I8_CONVERTERS, off by default:
//                         measured slower under the power cap)
//   3. i8_gemm_kernel     for every modulus and every 128 x 256 tile of the lower triangle:
//                         C_t = R_t^T R_t  with tcgen05.mma.kind::i8 (M128 N256 K32, int32 accumulators in
//                         TMEM, operands TMA-loaded with 128-byte swizzle into a 4-stage mbarrier ring; one
//                         elected thread issues the MMAs, one the TMA loads, four warps drain TMEM; persistent,
//                         one CTA per SM, two accumulators so that draining overlaps the next unit).  A unit
//                         covers <= 2^16 rows, so |C_t| <= 2^16 * 128^2 = 2^30 never wraps.  Accumulators are
//                         added into an int64 table with integer atomics (exact, order-independent).
//   4. i8_crt_kernel      reduces the table mod p_t, rebuilds the exact integer G'_ij = sum_r q_ri q_rj
//                         (|G'| <= 2^18 2^106 < P/2 = 2^124.4) by Garner's mixed-radix algorithm in 128-bit
//                         integers, rounds ONCE to fp64, scales by 2^-(e_i+e_j) and adds the slab into gaug.
//
// The result is the correctly rounded Gram of the quantised slab: independent of summation order, tile
// shape (and, per slab, of the row order), at least as accurate normwise as the DMMA path
// (tests/test_gpu_gram_int8.py).  Measured: DESIGN.md section 3.1a, profiles/r01_i8_gram_k1000.txt, r02_i8_k1000.txt.
// References: solvers/svd.py:35-53, solvers/ridge.py:28-43 (the products this replaces).
#include "fsb_common.cuh"
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <utility>

namespace {

constexpr int NMOD = 16;
constexpr int BM = 128;                 // tile rows   (columns of A, operand "A" of the MMA)
constexpr int BN = 256;                 // tile columns (columns of A, operand "B" of the MMA)
constexpr int BK = 128;                 // rows of the design matrix per pipeline stage (bytes of K per operand row)
constexpr int UK = 32;                  // K of one kind::i8 MMA
constexpr int NST = 4;
constexpr int A_BYTES = BM * BK;        // 16 KB
constexpr int B_BYTES = BN * BK;        // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int GEMM_THREADS = 192;       // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
constexpr int TMEM_COLS = 512;       // two 128 x 256 int32 accumulators
constexpr int64_t SLAB_ROWS = 262144;   // 2^18: with BETA = 53, 2^18 * 2^106 < P/2
constexpr int64_t UNIT_ROWS_MAX = 65536;
constexpr int CV_ROWS = 128, CV_COLS = 32, CV_THREADS = 256;
// Residue planes: byte (t, c, r) of a slab lives at  (t * kpc + c) * ldr + r  (ldr = slab height).  Modulus-major ON
// PURPOSE: a column-major variant with constant strides (the 16 stores of a column as immediate offsets) made the
// conversion 7 % faster and the tensor-core kernel 3.7x SLOWER -- the 128 lines of a TMA box were 4 MB apart, one 2 MB
// page each, and the units of one (chunk, modulus) touched >1000 pages instead of ~130 (gpurun_out/s4_launches_i8.csv).
#ifndef FSB_NCW
#define FSB_NCW 16
#endif
constexpr int NCW = FSB_NCW;                 // converter warps riding in the GEMM CTA (two groups of CV_THREADS / 32)
constexpr int GEMM_THREADS_CV = GEMM_THREADS + 32 * NCW;

struct I8Tables {
  int mod[NMOD];
  int half[NMOD];
  unsigned magic[NMOD];     // floor(2^32 / p) + 1: exact floor(u / p) for u < 2^23
  int off[NMOD];            // p * 2048 + half: makes the dp4a sum positive, folds the symmetric shift
  int wlo[NMOD], whi[NMOD]; // packed int8 weights 256^i mod p (i = 0..3 | 4..6, sign weight -2^56 mod p)
  // fp32 reduction of the dp4a sum S (|S| < 2^18), moduli <= 253: the dp4a accumulator starts at the BIT PATTERN
  // of the float T + 2048 p with T = 2^23 + 256 j = p s (j = -2^15 mod p), so its result, read as a float, is
  // uf = T + 2048 p + S exactly.  qm = fma(uf, 1/p, 1.5 2^23) rounds uf/p to an integer, q = qm - (1.5 2^23 + s)
  // = 2048 + rint(S/p), and fma(-q, p, uf) = 2^23 + 256 j + (S - p rint(S/p)): its low byte is the symmetric
  // residue in two's complement.  Three fp32 operations, no integer<->float conversion instruction.
  int finit[NMOD];
  float fsub[NMOD], finv[NMOD], fmod[NMOD];
  double dinv[NMOD];
  int garner_w[NMOD][NMOD]; // (p_0 ... p_{j-1}) mod p_k
  int garner_inv[NMOD];     // (p_0 ... p_{k-1})^-1 mod p_k
  unsigned long long p_lo, p_hi;        // P = prod p_t
  unsigned long long half_lo, half_hi;  // floor(P / 2)
};

constexpr int kMods[NMOD] = {256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197, 193};

constexpr int sym_mod(long long x, int p) {
  long long r = x % p;
  if (r < 0) r += p;
  return (int)(r > p / 2 ? r - p : r);
}
constexpr int pow_mod(int base, int e, int p) {
  long long r = 1, b = base % p;
  for (int i = 0; i < e; ++i) r = (r * b) % p;
  return (int)r;
}
constexpr int inv_mod(int a, int p) {
  a %= p;
  for (int x = 1; x < p; ++x)
    if ((a * x) % p == 1) return x;
  return 0;
}
constexpr int pack4(int a, int b, int c, int d) {
  return (int)(((unsigned)(a & 255)) | ((unsigned)(b & 255) << 8) | ((unsigned)(c & 255) << 16) |
               ((unsigned)(d & 255) << 24));
}
constexpr I8Tables make_tables() {
  I8Tables t{};
  unsigned __int128 P = 1;
  for (int i = 0; i < NMOD; ++i) {
    const int p = kMods[i];
    t.mod[i] = p;
    t.half[i] = p / 2;
    t.magic[i] = (unsigned)((1ull << 32) / (unsigned)p + 1ull);
    t.off[i] = p * 2048 + p / 2;
    int w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < 7; ++j) w[j] = sym_mod(pow_mod(256, j, p), p);
    w[7] = sym_mod(-(long long)pow_mod(256, 7, p), p);
    t.wlo[i] = pack4(w[0], w[1], w[2], w[3]);
    t.whi[i] = pack4(w[4], w[5], w[6], w[7]);
    {
      int j = (int)((((-32768ll) % p) + p) % p);
      if (j == 0) j = p;
      const int T = (1 << 23) + 256 * j;          // multiple of p (p odd)
      t.finit[i] = 0x4B000000 + 256 * j + 2048 * p;
      t.fsub[i] = (float)(12582912 + (p % 2 ? T / p : 0));
      t.finv[i] = 1.0f / (float)p;
      t.fmod[i] = (float)p;
      t.dinv[i] = 1.0 / (double)p;
    }
    int prod = 1;
    for (int j = 0; j < NMOD; ++j) {
      t.garner_w[i][j] = j < i ? prod : 0;
      if (j < i) prod = (prod * (kMods[j] % p)) % p;
    }
    t.garner_inv[i] = i == 0 ? 1 : inv_mod(prod, p);
    P *= (unsigned)p;
  }
  t.p_lo = (unsigned long long)P;
  t.p_hi = (unsigned long long)(P >> 64);
  const unsigned __int128 H = P >> 1;
  t.half_lo = (unsigned long long)H;
  t.half_hi = (unsigned long long)(H >> 64);
  return t;
}

__constant__ I8Tables c_tab = make_tables();
constexpr I8Tables k_tab = make_tables();      // the same values as compile-time constants (immediates in the converter)

// ------------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded spin: a protocol error becomes a launch failure (trap) instead of a hung GPU
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned done = 0;
  for (unsigned spin = 0; spin < (1u << 22); ++spin) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc,
                                          unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(unsigned taddr, unsigned (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// shared-memory matrix descriptor: K-major operand, 128-byte swizzle, 8-row groups 1024 bytes apart
__device__ __forceinline__ unsigned long long umma_desc_sw128(unsigned smem_addr) {
  unsigned long long d = 0;
  d |= (unsigned long long)((smem_addr & 0x3FFFFu) >> 4);   // start address
  d |= (unsigned long long)(1024u >> 4) << 32;              // stride byte offset (between 8-row groups)
  d |= 1ull << 46;                                          // descriptor version (sm_100)
  d |= 2ull << 61;                                          // SWIZZLE_128B
  return d;
}
// instruction descriptor of kind::i8: D = s32, A = B = signed 8 bit, both K-major, dense
__device__ __forceinline__ unsigned umma_idesc_i8(int m, int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(m >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------ 1
// column maxima of |w a| over the slab; bit patterns of non-negative doubles order like integers
__global__ void __launch_bounds__(256) i8_colmax_kernel(const double* __restrict__ A, int64_t lda,
                                                        const double* __restrict__ b,
                                                        const double* __restrict__ weff, int64_t nrows_total, int k,
                                                        int64_t slab_rows, int64_t rows_per_cta,
                                                        unsigned long long* __restrict__ colmax_all, int colmax_ld,
                                                        int* __restrict__ nonfinite) {
  // blockIdx.z = slab: all slabs of a Gram in ONE launch, maxima per slab in colmax_all[slab][colmax_ld]
  const int64_t slab0 = (int64_t)blockIdx.z * slab_rows;
  int64_t nrows = nrows_total - slab0;
  if (nrows > slab_rows) nrows = slab_rows;
  A += slab0 * lda;
  b += slab0;
  weff += slab0;
  unsigned long long* colmax = colmax_all + (size_t)blockIdx.z * colmax_ld;
  const int c = blockIdx.x * 256 + threadIdx.x;
  const int ka = k + 1;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
  int64_t r1 = r0 + rows_per_cta;
  if (r1 > nrows) r1 = nrows;
  if (c >= ka) return;
  const double* src = (c < k) ? A + c : b;
  const int64_t stride = (c < k) ? lda : 1;
  double m = 0.0;
  bool bad = false;
  int64_t r = r0;
  for (; r + 8 <= r1; r += 8) {
    double v[8], wv[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      v[q] = __ldg(src + (r + q) * stride);
      wv[q] = __ldg(weff + r + q);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const double t = fabs(v[q] * wv[q]);
      bad |= !(t <= 1.7976931348623157e308);
      m = fmax(m, t);
    }
  }
  for (; r < r1; ++r) {
    const double t = fabs(__ldg(src + r * stride) * __ldg(weff + r));
    bad |= !(t <= 1.7976931348623157e308);
    m = fmax(m, t);
  }
  if (bad) atomicAdd(nonfinite, 1);
  if (m > 0.0 && !bad) atomicMax(colmax + c, (unsigned long long)__double_as_longlong(m));
}

__device__ __forceinline__ int scale_exponent(unsigned long long max_bits, int beta) {
  const double m = __longlong_as_double((long long)max_bits);
  if (!(m > 0.0)) return 0;
  int e = beta - 1 - ilogb(m);
  if (e > 1000) e = 1000;
  if (e < -1000) e = -1000;
  return e;
}

// residues of four consecutive rows (one column) modulo p_T, packed into one word of plane T
template <int T>
__device__ __forceinline__ void convert_one_modulus(const unsigned (&lo)[4], const unsigned (&hi)[4], unsigned* dst,
                                                    size_t plane_words) {
  constexpr int P = kMods[T];
  // per-modulus constants as immediates: inside the converter loop of i8_gemm_kernel constant-bank operands would be
  // hoisted into ~80 registers
  constexpr int WLO = k_tab.wlo[T], WHI = k_tab.whi[T], OFF = k_tab.off[T], FINIT = k_tab.finit[T];
  constexpr unsigned MAGIC = k_tab.magic[T];
  constexpr float FINV = k_tab.finv[T];
  unsigned res[4];   // residue of row q in the low byte
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if constexpr (P == 256) {          // two's complement low byte
      res[q] = lo[q];
    } else if constexpr (P > 253) {    // 255: no slack for a sloppy quotient, exact integer reduction
      const unsigned u = (unsigned)dp4a_us(hi[q], WHI, dp4a_us(lo[q], WLO, OFF));
      const unsigned qq = __umulhi(u, MAGIC);
      res[q] = (unsigned)((int)(u - qq * (unsigned)P) - P / 2);
    } else {
      // u = bit pattern of the float uf = T + 2048 p + S (see I8Tables).  qm = 1.5 2^23 + n with n = rint(uf / p) =
      // T/p + 2048 + rint(S / p); its bit pattern is 0x4B400000 + n.  Only the LOW BYTE of the result is kept, and
      // T, 2048 p and 0x4B400000 P are multiples of 256: (u - bits(qm) P) mod 256 = (S - P rint(S / p)) mod 256, the
      // symmetric residue in two's complement -- one integer multiply-add instead of a float subtract and a fused
      // multiply-add (same n, hence the same residue bytes as the three-float-operation form).
      const int u = dp4a_us(hi[q], WHI, dp4a_us(lo[q], WLO, FINIT));
      const float qm = __fmaf_rn(__int_as_float(u), FINV, 12582912.0f);
      res[q] = (unsigned)(u - __float_as_int(qm) * P);
    }
  }
  const unsigned t01 = __byte_perm(res[0], res[1], 0x0040);
  const unsigned t23 = __byte_perm(res[2], res[3], 0x0040);
  dst[(size_t)T * plane_words] = __byte_perm(t01, t23, 0x5410);
}
template <int... T>
__device__ __forceinline__ void convert_all_moduli(const unsigned (&lo)[4], const unsigned (&hi)[4], unsigned* dst,
                                                   size_t plane_words, std::integer_sequence<int, T...>) {
  (convert_one_modulus<T>(lo, hi, dst, plane_words), ...);
}

// ------------------------------------------------------------------------------------------------ 2
// residue planes: planes[(t * kpc + c) * ldr + r] = (rint(w_r a_rc 2^e_c)) mod p_t, symmetric, int8
// A warp owns 4 columns x 128 rows: lane l holds rows 4l..4l+3 of each, so every store instruction writes one
// full 128-byte line of one plane (no shared-memory transpose) and every load fetches whole 32-byte sectors.
// Eight warps side by side cover 32 columns = 256 contiguous bytes of every row.
struct ConvArgs {
  const double* A;
  int64_t lda;
  const double* b;
  const double* weff;
  int64_t nrows;                       // rows of this slab
  int k;
  const unsigned long long* colmax;    // column maxima of this slab
  int beta;
  unsigned* planes;
  int64_t ldr;                         // rows a plane column holds (slab height, multiple of CV_ROWS)
  int ncolt;                           // column tiles (kpc / CV_COLS)
  int ntile;                           // ncolt * row blocks; 0 = nothing to convert
};

__device__ __forceinline__ void convert_tile(const ConvArgs& cv, int c0, int64_t r0) {
  const int k = cv.k, ka = cv.k + 1;
  const size_t col_words = (size_t)cv.ldr / 4, plane_words = (size_t)cv.ncolt * CV_COLS * col_words;
  unsigned* colbase = cv.planes + (size_t)c0 * col_words + (size_t)r0 / 4;
  if (c0 >= ka) {                                        // padding columns of the last tile: zero residues
    for (int j = 0; j < 4; ++j)
      for (int t = 0; t < NMOD; ++t) colbase[(size_t)t * plane_words + (size_t)j * col_words] = 0u;
    return;
  }
  const double* __restrict__ A = cv.A;
  const int64_t lda = cv.lda, nrows = cv.nrows;
  double v[4][4], wv[4];
  const bool vec = ((lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int64_t r = r0 + q;
    const int64_t rc = r < nrows ? r : nrows - 1;         // rows past the end: clamped address, weight 0
    wv[q] = r < nrows ? __ldg(cv.weff + rc) : 0.0;
    const double* row = A + rc * lda;
    if (vec && c0 + 3 < k) {     // whole 32-byte sector of this row in two 16-byte loads
      const double2 x0 = __ldg(reinterpret_cast<const double2*>(row + c0));
      const double2 x1 = __ldg(reinterpret_cast<const double2*>(row + c0 + 2));
      v[q][0] = x0.x; v[q][1] = x0.y; v[q][2] = x1.x; v[q][3] = x1.y;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + j;
        v[q][j] = (c < k) ? __ldg(row + c) : ((c == k) ? __ldg(cv.b + rc) : 0.0);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c0 + j;
    double scale = 0.0;
    if (c < ka)   // 2^e, e in [-1000, 1000]: assemble the exponent field directly
      scale = __longlong_as_double((long long)(scale_exponent(__ldg(cv.colmax + c), cv.beta) + 1023) << 52);
    unsigned lo[4], hi[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long long iv = __double2ll_rn((v[q][j] * wv[q]) * scale);   // fl(w*a) first: the reference's aw (svd.py:44)
      lo[q] = (unsigned)iv;
      hi[q] = ((unsigned)((unsigned long long)iv >> 32) & 0x00FFFFFFu) | (iv < 0 ? 0x01000000u : 0u);
    }
    convert_all_moduli(lo, hi, colbase + (size_t)j * col_words, plane_words, std::make_integer_sequence<int, NMOD>{});
  }
}

// stand-alone conversion (first slab of a Gram; the following slabs are converted inside i8_gemm_kernel)
__global__ void __launch_bounds__(CV_THREADS, 3) i8_convert_kernel(ConvArgs cv) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  convert_tile(cv, blockIdx.x * CV_COLS + warp * 4, (int64_t)blockIdx.y * CV_ROWS + lane * 4);
}

// ------------------------------------------------------------------------------------------------ 3
struct GemmArgs {
  int n_i;                 // 128-column blocks of the augmented matrix
  int ntile;               // (I, JJ) tiles of the lower triangle
  int ka;
  int kp;                  // leading dimension of the int64 table (n_i * 128)
  int nunits;              // chunks x NMOD x ntile
  int64_t slab_rows;
  int64_t unit_rows;
  long long* table;        // [NMOD][kp (column j)][kp (row i)]
};

struct Unit {
  int t, col_i, col_j, n_mma, nks;
  int64_t row0;
  bool wide;
};

// unit index -> (row chunk, modulus, tile); tile fastest: the tiles of one (chunk, modulus) run at the same
// time on neighbouring SMs and share the plane rows in L2
__device__ __forceinline__ Unit decode_unit(const GemmArgs& p, int unit) {
  Unit u;
  const int tile = unit % p.ntile;
  unit /= p.ntile;
  u.t = unit % NMOD;
  const int64_t chunk = unit / NMOD;
  int ti = 0, tj = 0, rem = tile;
  for (ti = 0; ti < p.n_i; ++ti) {
    const int cnt = ti / 2 + 1;
    if (rem < cnt) { tj = rem; break; }
    rem -= cnt;
  }
  u.col_i = ti * BM;
  u.col_j = tj * BN;
  // second 128-column half of JJ: it must exist AND reach the lower triangle (for an even block row the last tile's
  // second half lies entirely above the diagonal: N = 128 there saves 10 % of the MMAs and operand loads at 8 block rows)
  u.wide = ((u.col_j + BM) < p.n_i * BM) && (u.col_j + BM <= u.col_i);
  u.n_mma = u.wide ? BN : BM;
  u.row0 = chunk * p.unit_rows;
  int64_t row1 = u.row0 + p.unit_rows;
  if (row1 > p.slab_rows) row1 = p.slab_rows;
  u.nks = (int)((row1 - u.row0 + BK - 1) / BK);
  return u;
}

// Persistent: one CTA per SM walks units blockIdx.x, blockIdx.x + gridDim.x, ...  (all CTAs of the grid are
// resident from the start, so the block scheduler can place the conversion kernel of the next slab beside them).
// Two accumulators in TMEM (2 x 256 columns): the epilogue of unit n drains one while the MMAs of unit n+1
// fill the other.
//
// Warps 6.. (present when the launch has GEMM_THREADS_CV threads) are CONVERTERS: while the tensor cores contract the
// planes of slab s, they turn slab s+1 of the design matrix into the OTHER plane buffer (convert_tile, the code of
// i8_convert_kernel).  The MMA issuer, the TMA producer and the epilogue warps leave nearly every issue slot of the
// SM unused, and the conversion needs no shared memory: both halves of the Gram pipeline run at the same time
// inside one launch, the kernel boundary is the only synchronisation between them.  Column tile fastest, tiles
// strided over (CTA, converter group): at any moment the whole GPU reads one neighbourhood of rows of A.
__global__ void __launch_bounds__(GEMM_THREADS_CV, 1) i8_gemm_kernel(const __grid_constant__ CUtensorMap tmap, GemmArgs p,
                                                                    ConvArgs cv) {
  extern __shared__ __align__(1024) unsigned char gm_sm[];
  __shared__ __align__(8) unsigned long long s_full[NST];
  __shared__ __align__(8) unsigned long long s_empty[NST];
  __shared__ __align__(8) unsigned long long s_acc_full[2];
  __shared__ __align__(8) unsigned long long s_acc_empty[2];
  __shared__ unsigned s_tmem;

  // 1024-byte aligned stage buffers (128-byte swizzle atoms are 1024 bytes)
  const unsigned sm_base = (smem_u32(gm_sm) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) {
      mbar_init(smem_u32(&s_full[i]), 1);
      mbar_init(smem_u32(&s_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&s_acc_full[i]), 1);
      mbar_init(smem_u32(&s_acc_empty[i]), 4);   // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                 "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem = s_tmem;

  if (warp == 0) {
    if (lane == 0) {
      unsigned it = 0;                            // pipeline stage counter, runs on across units
      for (int unit = blockIdx.x; unit < p.nunits; unit += gridDim.x) {
        const Unit u = decode_unit(p, unit);
        const unsigned bytes = (unsigned)(A_BYTES + u.n_mma * BK);
        for (int ks = 0; ks < u.nks; ++ks, ++it) {
          const unsigned slot = it % NST, n = it / NST;
          if (it >= NST) mbar_wait(smem_u32(&s_empty[slot]), (n - 1) & 1);
          const unsigned full = smem_u32(&s_full[slot]);
          const unsigned dst = sm_base + slot * STAGE_BYTES;
          const int r = (int)(u.row0 + (int64_t)ks * BK);
          mbar_expect_tx(full, bytes);
          tma_load_3d(dst, &tmap, r, u.col_i, u.t, full);
          tma_load_3d(dst + A_BYTES, &tmap, r, u.col_j, u.t, full);
          if (u.wide) tma_load_3d(dst + A_BYTES + BM * BK, &tmap, r, u.col_j + BM, u.t, full);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      unsigned it = 0, nu = 0;
      for (int unit = blockIdx.x; unit < p.nunits; unit += gridDim.x, ++nu) {
        const Unit u = decode_unit(p, unit);
        const unsigned idesc = umma_idesc_i8(BM, u.n_mma);
        const unsigned ab = nu & 1;               // accumulator buffer
        if (nu >= 2) mbar_wait(smem_u32(&s_acc_empty[ab]), ((nu >> 1) - 1) & 1);   // drained by the epilogue
        tc_fence_after();
        const unsigned acc = tmem + ab * BN;
        for (int ks = 0; ks < u.nks; ++ks, ++it) {
          const unsigned slot = it % NST, n = it / NST;
          mbar_wait(smem_u32(&s_full[slot]), n & 1);
          tc_fence_after();
          const unsigned a_addr = sm_base + slot * STAGE_BYTES;
          const unsigned long long adesc = umma_desc_sw128(a_addr);
          const unsigned long long bdesc = umma_desc_sw128(a_addr + A_BYTES);
#pragma unroll
          for (int k4 = 0; k4 < BK / UK; ++k4)   // +32 bytes of K inside the swizzle atom = +2 in the address field
            tc_mma_i8(acc, adesc + (unsigned long long)(2 * k4), bdesc + (unsigned long long)(2 * k4), idesc,
                      (unsigned)((ks | k4) != 0));
          tc_commit(smem_u32(&s_empty[slot]));   // frees the stage when these MMAs have read it
        }
        tc_commit(smem_u32(&s_acc_full[ab]));    // accumulator complete
      }
    }
  } else if (warp >= GEMM_THREADS / 32) {
    const int cw = warp - GEMM_THREADS / 32;
    const int grp = cw >> 3, w8 = cw & 7;
    const int ngrp = (int)(blockDim.x - GEMM_THREADS) / CV_THREADS;
    for (int tile = (int)blockIdx.x * ngrp + grp; tile < cv.ntile; tile += (int)gridDim.x * ngrp) {
      const int colt = tile % cv.ncolt, rowb = tile / cv.ncolt;
      convert_tile(cv, colt * CV_COLS + w8 * 4, (int64_t)rowb * CV_ROWS + lane * 4);
    }
  } else {
    // epilogue: warp w may touch TMEM lanes [32 (w % 4), +32)
    const int quad = warp & 3;
    unsigned nu = 0;
    for (int unit = blockIdx.x; unit < p.nunits; unit += gridDim.x, ++nu) {
      const Unit u = decode_unit(p, unit);
      const unsigned ab = nu & 1;
      const int grow = u.col_i + quad * 32 + lane;   // row i of the Gram (column index of A)
      mbar_wait(smem_u32(&s_acc_full[ab]), (nu >> 1) & 1);
      tc_fence_after();
      long long* tab = p.table + (size_t)u.t * p.kp * p.kp;
      for (int c0 = 0; c0 < u.n_mma; c0 += 32) {
        if (u.col_j + c0 > u.col_i + BM - 1) break;    // the rest of the tile lies above the diagonal
        unsigned v[32];
        tc_ld32(tmem + ((unsigned)(quad * 32) << 16) + ab * BN + (unsigned)c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int gcol = u.col_j + c0 + j;
          if (grow < p.ka && gcol <= grow && u.nks > 0)
            atomicAdd(reinterpret_cast<unsigned long long*>(tab + (size_t)gcol * p.kp + grow),
                      (unsigned long long)(long long)(int)v[j]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        unsigned long long st;
        asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(st) : "r"(smem_u32(&s_acc_empty[ab])) : "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ 4
// exact reconstruction of the slab's integer Gram and accumulation into gaug (lower triangle + mirror)
__global__ void __launch_bounds__(128) i8_crt_kernel(long long* __restrict__ table, int kp, int ka,
                                                     const unsigned long long* __restrict__ colmax, int beta,
                                                     const int* __restrict__ nonfinite, int first,
                                                     double* __restrict__ gaug) {
  const int j = blockIdx.y;
  const int i = j + blockIdx.x * 128 + threadIdx.x;
  if (i >= ka) return;
  int res[NMOD];
  long long tv[NMOD];
#pragma unroll
  for (int t = 0; t < NMOD; ++t) tv[t] = __ldcg(table + ((size_t)t * kp + j) * kp + i);   // 16 loads in flight per thread
#pragma unroll
  for (int t = 0; t < NMOD; ++t) {
    long long* cell = table + ((size_t)t * kp + j) * kp + i;
    const long long v = tv[t];
    *cell = 0;                                   // ready for the next slab
    // |v| <= 4 units x 2^30: quotient estimate in fp64, then one correction step either way
    const int p = c_tab.mod[t];
    int r = (int)(v - (long long)floor((double)v * c_tab.dinv[t]) * p);
    r = r < 0 ? r + p : r;
    r = r >= p ? r - p : r;
    res[t] = r;
  }
  // Garner: x = d_0 + d_1 p_0 + d_2 p_0 p_1 + ...   (all operands < 2^21: floor(u/p) = umulhi(u, magic))
  int dig[NMOD];
  dig[0] = res[0];
#pragma unroll
  for (int kk = 1; kk < NMOD; ++kk) {
    const unsigned p = (unsigned)c_tab.mod[kk];
    unsigned acc = 0;
#pragma unroll
    for (int jj = 0; jj < kk; ++jj) acc += (unsigned)(dig[jj] * c_tab.garner_w[kk][jj]);
    acc -= __umulhi(acc, c_tab.magic[kk]) * p;                 // acc mod p
    unsigned d = (unsigned)res[kk] + p - acc;                  // in (0, 2p)
    d = d >= p ? d - p : d;
    d *= (unsigned)c_tab.garner_inv[kk];
    dig[kk] = (int)(d - __umulhi(d, c_tab.magic[kk]) * p);
  }
  unsigned __int128 x = (unsigned)dig[NMOD - 1];
#pragma unroll
  for (int kk = NMOD - 2; kk >= 0; --kk) x = x * (unsigned)c_tab.mod[kk] + (unsigned)dig[kk];
  const unsigned __int128 P = ((unsigned __int128)c_tab.p_hi << 64) | c_tab.p_lo;
  const unsigned __int128 H = ((unsigned __int128)c_tab.half_hi << 64) | c_tab.half_lo;
  double sign = 1.0;
  if (x > H) { x = P - x; sign = -1.0; }
  // one rounding: normalise to 64 significant bits, keep a sticky bit, let the int64 -> double conversion round
  const unsigned long long xh = (unsigned long long)(x >> 64), xl = (unsigned long long)x;
  double val;
  if (xh == 0) {
    val = (double)xl;
  } else {
    const int sh = 64 - __clzll((long long)xh);            // 1..64 bits live in the high word
    unsigned long long top = (sh == 64) ? xh : ((xh << (64 - sh)) | (xl >> sh));
    const unsigned long long lost = (sh == 64) ? xl : (xl << (64 - sh));
    if (lost) top |= 1ull;
    val = ldexp((double)top, sh);
  }
  const int e = scale_exponent(colmax[i], beta) + scale_exponent(colmax[j], beta);
  val = sign * ldexp(val, -e);
  if (*nonfinite) val = __longlong_as_double(0x7ff8000000000000ll);
  if (first) {
    gaug[(size_t)i * ka + j] = val;
    gaug[(size_t)j * ka + i] = val;
  } else {
    const double s = gaug[(size_t)i * ka + j] + val;
    gaug[(size_t)i * ka + j] = s;
    gaug[(size_t)j * ka + i] = s;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

int converter_threads() {             // development switch: FSB_I8_CONVERTERS = 0 (default) | 256 | 512
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FSB_I8_CONVERTERS");
    v = e ? atoi(e) : 0;
    if (v < 0) v = 0;
    v = (v / CV_THREADS) * CV_THREADS;
    if (v > 32 * NCW) v = 32 * NCW;
  }
  return v;
}

struct I8Plan {
  int ka, n_i, ntile, kp, kpc, ka_pad, nbuf;
  int64_t slab_rows, nslab;
  size_t off_colmax, off_flag, off_table, off_planes, plane_bytes, total;
};

I8Plan plan_i8(int64_t n_rows, int k) {
  I8Plan pl;
  pl.ka = k + 1;
  pl.n_i = (pl.ka + BM - 1) / BM;
  pl.ntile = 0;
  for (int i = 0; i < pl.n_i; ++i) pl.ntile += i / 2 + 1;
  pl.kp = pl.n_i * BM;
  pl.kpc = (int)fsb_round_up(pl.ka, CV_COLS);
  pl.ka_pad = (int)fsb_round_up(pl.ka, 32);
  const int64_t n = n_rows > 0 ? n_rows : 1;
  pl.nslab = fsb_ceil_div(n, SLAB_ROWS);
  pl.slab_rows = fsb_round_up(fsb_ceil_div(n, pl.nslab), CV_ROWS);
  pl.nbuf = (pl.nslab > 1 && converter_threads() > 0) ? 2 : 1;   // a second plane buffer only when slab s+1 is converted beside the contraction of slab s
  pl.off_colmax = 0;
  pl.off_flag = align256((size_t)pl.nslab * pl.ka_pad * sizeof(unsigned long long));
  pl.off_table = pl.off_flag + 256;
  size_t off = pl.off_table + align256((size_t)NMOD * pl.kp * pl.kp * sizeof(long long));
  pl.off_planes = (off + 1023) & ~(size_t)1023;
  pl.plane_bytes = ((size_t)NMOD * pl.kpc * (size_t)pl.slab_rows + 1023) & ~(size_t)1023;
  pl.total = pl.off_planes + pl.nbuf * pl.plane_bytes + 1024;   // + slack to align the base
  return pl;
}

}  // namespace

bool fsb_gram_i8_available() { return get_encode() != nullptr; }

size_t fsb_gram_i8_ws_bytes(int64_t n_rows, int k) { return plan_i8(n_rows, k).total; }

// `weff` already carries the test mask (weight 0).  gaug is fully overwritten.  Everything runs on `s`.
// Per Gram: column maxima of every slab (one launch), conversion of slab 0, then per slab ONE launch that contracts
// slab s on the tensor cores and converts slab s+1 beside it (converter warps of i8_gemm_kernel), and the CRT.
int fsb_launch_gram_i8(const fsb_context* h, const double* A, int64_t lda, const double* b, const double* weff,
                       int64_t n_rows, int k, double* gaug, void* ws, size_t ws_bytes, cudaStream_t s) {
  const I8Plan pl = plan_i8(n_rows, k);
  if (ws_bytes < pl.total) return FSB_ERR_WORKSPACE_TOO_SMALL;
  if (!get_encode()) return FSB_ERR_UNSUPPORTED;
  char* base = (char*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  unsigned long long* colmax = (unsigned long long*)(base + pl.off_colmax);
  int* flag = (int*)(base + pl.off_flag);
  long long* table = (long long*)(base + pl.off_table);
  char* planes = base + pl.off_planes;
  const int ka = pl.ka;
  const int beta = 53;

  FSB_CUDA_TRY(cudaMemsetAsync(table, 0, (size_t)NMOD * pl.kp * pl.kp * sizeof(long long), s));
  FSB_CUDA_TRY(cudaMemsetAsync(flag, 0, sizeof(int), s));
  if (n_rows <= 0) {
    FSB_CUDA_TRY(cudaMemsetAsync(gaug, 0, (size_t)ka * ka * sizeof(double), s));
    return FSB_OK;
  }
  const size_t gemm_smem = (size_t)NST * STAGE_BYTES + 1024;
  FSB_CUDA_TRY(cudaFuncSetAttribute(i8_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem));
  const int64_t nslab = pl.nslab;
  FSB_CUDA_TRY(cudaMemsetAsync(colmax, 0, (size_t)nslab * pl.ka_pad * sizeof(unsigned long long), s));
  {
    const int64_t nr = n_rows < pl.slab_rows ? n_rows : pl.slab_rows;
    // ~32 CTAs per SM in total: with exactly one resident wave + a few CTAs (1200 on 1184 slots at 5 slabs) the
    // stragglers ran alone and the launch took 2.5 ms instead of 1.6 (profiles/r02_launches_bench_default.csv)
    int64_t rb = fsb_ceil_div(nr, fsb_ceil_div((int64_t)h->sm_count * 32, fsb_ceil_div(ka, 256) * nslab));
    rb = fsb_round_up(rb < 64 ? 64 : rb, 8);
    dim3 grid((unsigned)fsb_ceil_div(ka, 256), (unsigned)fsb_ceil_div(nr, rb), (unsigned)nslab);
    i8_colmax_kernel<<<grid, 256, 0, s>>>(A, lda, b, weff, n_rows, k, pl.slab_rows, rb, colmax, pl.ka_pad, flag);
    FSB_LAUNCH_CHECK("i8_colmax_kernel");
  }
  auto conv_args = [&](int64_t i) {
    ConvArgs cv;
    const int64_t r0 = i * pl.slab_rows;
    cv.A = A + r0 * lda; cv.lda = lda; cv.b = b + r0; cv.weff = weff + r0;
    cv.nrows = (n_rows - r0) < pl.slab_rows ? (n_rows - r0) : pl.slab_rows;
    cv.k = k; cv.colmax = colmax + (size_t)i * pl.ka_pad; cv.beta = beta;
    cv.planes = (unsigned*)(planes + (size_t)(i % pl.nbuf) * pl.plane_bytes);
    cv.ldr = pl.slab_rows;
    cv.ncolt = pl.kpc / CV_COLS;
    cv.ntile = cv.ncolt * (int)fsb_ceil_div(cv.nrows, CV_ROWS);
    return cv;
  };
  const int cv_threads = converter_threads();
  for (int64_t i = 0; i < nslab; ++i) {
    const ConvArgs cur = conv_args(i);
    const int64_t nr = cur.nrows;
    if (i == 0 || cv_threads == 0) {
      dim3 grid((unsigned)cur.ncolt, (unsigned)fsb_ceil_div(nr, CV_ROWS));
      i8_convert_kernel<<<grid, CV_THREADS, 0, s>>>(cur);
      FSB_LAUNCH_CHECK("i8_convert_kernel");
    }
    {
      CUtensorMap tmap;
      const cuuint64_t gdim[3] = {(cuuint64_t)nr, (cuuint64_t)pl.kpc, (cuuint64_t)NMOD};
      const cuuint64_t gstr[2] = {(cuuint64_t)pl.slab_rows, (cuuint64_t)pl.slab_rows * (cuuint64_t)pl.kpc};
      const cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)BM, 1};
      const cuuint32_t estr[3] = {1, 1, 1};
      CUresult cr = get_encode()(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)cur.planes, gdim, gstr, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (cr != CUDA_SUCCESS) return FSB_ERR_UNSUPPORTED;
      GemmArgs ga;
      ga.n_i = pl.n_i; ga.ntile = pl.ntile; ga.ka = ka; ga.kp = pl.kp; ga.slab_rows = nr; ga.table = table;
      // enough units for ~8 per SM, each between 4096 and 65536 rows
      int64_t want_chunks = fsb_ceil_div((int64_t)h->sm_count * 8, (int64_t)pl.ntile * NMOD);
      int64_t ur = fsb_round_up(fsb_ceil_div(nr, want_chunks), BK);
      if (ur < 4096) ur = 4096;
      if (ur > UNIT_ROWS_MAX) ur = UNIT_ROWS_MAX;
      ga.unit_rows = ur;
      const int64_t nchunk = fsb_ceil_div(nr, ur);
      ga.nunits = (int)(nchunk * NMOD * pl.ntile);
      const int grid = ga.nunits < h->sm_count ? ga.nunits : h->sm_count;
      ConvArgs next;
      next.ntile = 0;
      int threads = GEMM_THREADS;
      if (i + 1 < nslab && cv_threads > 0) {
        next = conv_args(i + 1);
        threads = GEMM_THREADS + cv_threads;
      } else {
        next = cur;
        next.ntile = 0;
      }
      i8_gemm_kernel<<<(unsigned)grid, (unsigned)threads, gemm_smem, s>>>(tmap, ga, next);
      FSB_LAUNCH_CHECK("i8_gemm_kernel");
    }
    {
      dim3 grid((unsigned)fsb_ceil_div(ka, 128), (unsigned)ka);
      i8_crt_kernel<<<grid, 128, 0, s>>>(table, pl.kp, ka, cur.colmax, beta, flag, i == 0 ? 1 : 0, gaug);
      FSB_LAUNCH_CHECK("i8_crt_kernel");
    }
  }
  return FSB_OK;
}
