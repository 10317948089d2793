// K2+K3+K4, wide matrices (k + 1 > 104): pre-weight pass + TMA-fed DMMA Gram.
//
// What the measurements on B200 said: the fp64 pipe is shared between DMMA and
// DMUL, so ANY multiply issued while the DMMAs run is expensive (consumer-side weighting: -14 %;
// producer-side weighting makes the producers the critical path), cp.async (LDGSTS) is slow
// (~46 cycles per warp instruction per SM), and register-staged producers cannot keep enough bytes
// in flight next to 512 consumer threads holding the accumulators.  So:
//   pass 1  preweight_kernel   Waug[r, :] = [ w_r * A[r, :k] | w_r * b_r | 0 ... ]   (HBM-bound,
//           one read of A, one write; rows 128-byte aligned, leading dimension ldw = ceil16(k+1))
//           -- fl(w*a), the same rounding as the reference's `aw` (svd.py:44); test rows have w = 0.
//   pass 2  gram_tma_kernel    ONE elected thread per CTA issues cp.async.bulk.tensor (TMA, SASS
//           UTMALDG) box loads [16 rows x 132 columns] of the two column ranges of the super-tile
//           straight into the padded (pitch 132, conflict-free) shared-memory layout of a 6-deep
//           ring; completion is tracked by mbarriers (expect_tx / complete_tx); rows and columns
//           past the edge are zero-filled by the TMA unit.  16 consumer warps: LDS + DMMA only.
// Pass 1 costs one extra read+write of A (2.5 ms per 1e6 x 1000) against ~40 ms of DMMA work, and the
// ~8x re-reads of the operand columns by the 36 super-tiles hit L2.
#include "fsb_common.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace {

constexpr int T_RCH = 16;
constexpr int T_NSTAGE = 6;
constexpr int T_RANGE = T_RCH * FSB_GLDS;          // doubles
constexpr int T_STAGE = 2 * T_RANGE;               // I range, J range
constexpr int T_CONSUMERS = FSB_GTHREADS;          // 512
constexpr int T_THREADS = T_CONSUMERS + 32;        // + one producer warp

struct TmaArgs {
  int64_t n_rows;
  int k;
  int ntile;
  int64_t rows_per_chunk;
  double* partial;
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "FSB_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra FSB_DONE_%=;\n\t"
      "bra FSB_WAIT_%=;\n\t"
      "FSB_DONE_%=:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

__device__ __forceinline__ void tile_coords(int tile, int& ti, int& tj) {
  int t = (int)((sqrtf(8.0f * (float)tile + 1.0f) - 1.0f) * 0.5f);
  while ((t + 1) * (t + 2) / 2 <= tile) ++t;
  while (t * (t + 1) / 2 > tile) --t;
  ti = t;
  tj = tile - t * (t + 1) / 2;
}

template <int MI, int NJ, bool OND>
__device__ __forceinline__ void compute_stage(double (&acc)[4][4][2], const double* __restrict__ fI,
                                              const double* __restrict__ fJ) {
#pragma unroll
  for (int ks = 0; ks < T_RCH / 4; ++ks) {
    double af[MI], bf[NJ];
#pragma unroll
    for (int i = 0; i < MI; ++i) af[i] = fI[ks * 4 * FSB_GLDS + i * 8];
    if (OND) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) bf[j] = af[j];
    } else {
#pragma unroll
      for (int j = 0; j < NJ; ++j) bf[j] = fJ[ks * 4 * FSB_GLDS + j * 8];
    }
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (!OND || j <= i) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
  }
}

__global__ void __launch_bounds__(T_THREADS, 1) gram_tma_kernel(const __grid_constant__ CUtensorMap tmap, TmaArgs p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);       // [T_NSTAGE][ I: 16 x 132 | J: 16 x 132 ]
  __shared__ __align__(8) unsigned long long s_full[T_NSTAGE];
  __shared__ __align__(8) unsigned long long s_empty[T_NSTAGE];
  __shared__ unsigned char s_map[16];

  const int tile = blockIdx.x % p.ntile;
  const int64_t chunk = blockIdx.x / p.ntile;
  int ti, tj;
  tile_coords(tile, ti, tj);
  const bool diag = (ti == tj);
  const int ka = p.k + 1;
  const int colI0 = ti * FSB_GT, colJ0 = tj * FSB_GT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  int nbI = (ka - colI0 + 7) >> 3; nbI = nbI > 16 ? 16 : nbI;
  int nbJ = (ka - colJ0 + 7) >> 3; nbJ = nbJ > 16 ? 16 : nbJ;

  if (tid == 0) {
    for (int s = 0; s < T_NSTAGE; ++s) {
      mbar_init(smem_u32(&s_full[s]), 1);                     // the producer's expect_tx arrival
      mbar_init(smem_u32(&s_empty[s]), T_CONSUMERS / 32);     // one arrival per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // deal the 16 warp tiles to the consumer warps (LPT greedy over the four schedulers, warp % 4)
    int cost[16], order[16], load[4] = {0, 0, 0, 0}, cnt[4] = {0, 0, 0, 0};
    for (int t = 0; t < 16; ++t) {
      const int r = t >> 2, c = t & 3;
      int m = nbI - 4 * r; m = m > 4 ? 4 : (m < 0 ? 0 : m);
      int n = nbJ - 4 * c; n = n > 4 ? 4 : (n < 0 ? 0 : n);
      cost[t] = (diag && c > r) ? 0 : ((diag && c == r) ? m * (m + 1) / 2 : m * n);
      order[t] = t;
    }
    for (int a = 1; a < 16; ++a) {
      const int o = order[a];
      int q = a - 1;
      while (q >= 0 && cost[order[q]] < cost[o]) { order[q + 1] = order[q]; --q; }
      order[q + 1] = o;
    }
    for (int a = 0; a < 16; ++a) {
      int best = -1;
      for (int q = 0; q < 4; ++q)
        if (cnt[q] < 4 && (best < 0 || load[q] < load[best])) best = q;
      s_map[cnt[best] * 4 + best] = (unsigned char)order[a];
      load[best] += cost[order[a]];
      cnt[best] += 1;
    }
  }
  __syncthreads();

  const int64_t row_begin = chunk * p.rows_per_chunk;       // multiple of T_RCH: stages never straddle chunks
  int64_t row_end = row_begin + p.rows_per_chunk;
  if (row_end > p.n_rows) row_end = p.n_rows;
  const int nsteps = row_end > row_begin ? (int)((row_end - row_begin + T_RCH - 1) / T_RCH) : 0;

  if (warp == T_CONSUMERS / 32) {
    // ------------------------------------ producer: TMA ------------------------------------------
    if (lane == 0) {
      const unsigned bytes = (diag ? 1u : 2u) * (unsigned)(T_RANGE * sizeof(double));
      for (int s = 0; s < nsteps; ++s) {
        const int slot = s % T_NSTAGE;
        if (s >= T_NSTAGE) mbar_wait(smem_u32(&s_empty[slot]), (unsigned)(((s / T_NSTAGE) - 1) & 1));
        const unsigned full = smem_u32(&s_full[slot]);
        mbar_expect_tx(full, bytes);
        const int row = (int)(row_begin + (int64_t)s * T_RCH);
        double* st = smem + (size_t)slot * T_STAGE;
        tma_load_2d(smem_u32(st), &tmap, colI0, row, full);
        if (!diag) tma_load_2d(smem_u32(st + T_RANGE), &tmap, colJ0, row, full);
      }
    }
    return;
  }

  // ------------------------------------ consumers: LDS -> DMMA -----------------------------------
  const int wt = s_map[warp];
  const int wr = wt >> 2, wc = wt & 3;
  int mi = nbI - 4 * wr; mi = mi > 4 ? 4 : mi;
  int nj = nbJ - 4 * wc; nj = nj > 4 ? 4 : nj;
  const bool active = (mi > 0) && (nj > 0) && (!diag || wc <= wr);
  const bool on_diag = diag && (wc == wr);
  const int shape = !active ? -1 : (on_diag ? 16 + (mi - 1) : (mi - 1) * 4 + (nj - 1));

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int frag_off = (lane & 3) * FSB_GLDS + (lane >> 2);
  for (int s = 0; s < nsteps; ++s) {
    const int slot = s % T_NSTAGE;
    mbar_wait(smem_u32(&s_full[slot]), (unsigned)((s / T_NSTAGE) & 1));
    if (active) {
      const double* st = smem + (size_t)slot * T_STAGE;
      const double* fI = st + frag_off + wr * 32;
      const double* fJ = (diag ? st : st + T_RANGE) + frag_off + wc * 32;
      switch (shape) {
#define FSB_CASE(MI, NJ) case (MI - 1) * 4 + (NJ - 1): compute_stage<MI, NJ, false>(acc, fI, fJ); break;
        FSB_CASE(4, 4) FSB_CASE(4, 3) FSB_CASE(4, 2) FSB_CASE(4, 1)
        FSB_CASE(3, 4) FSB_CASE(3, 3) FSB_CASE(3, 2) FSB_CASE(3, 1)
        FSB_CASE(2, 4) FSB_CASE(2, 3) FSB_CASE(2, 2) FSB_CASE(2, 1)
        FSB_CASE(1, 4) FSB_CASE(1, 3) FSB_CASE(1, 2) FSB_CASE(1, 1)
#undef FSB_CASE
        case 16: compute_stage<1, 1, true>(acc, fI, fJ); break;
        case 17: compute_stage<2, 2, true>(acc, fI, fJ); break;
        case 18: compute_stage<3, 3, true>(acc, fI, fJ); break;
        case 19: compute_stage<4, 4, true>(acc, fI, fJ); break;
        default: break;
      }
    }
    __syncwarp();
    if (lane == 0 && s + T_NSTAGE < nsteps) mbar_arrive(smem_u32(&s_empty[slot]));
  }

  if (active) {
    double* out = p.partial + ((size_t)chunk * p.ntile + tile) * (size_t)(FSB_GT * FSB_GT);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (i < mi && j < nj && (!on_diag || j <= i)) {
          const int row = wr * 32 + i * 8 + (lane >> 2);
          const int col = wc * 32 + j * 8 + 2 * (lane & 3);
          *reinterpret_cast<double2*>(out + row * FSB_GT + col) = make_double2(acc[i][j][0], acc[i][j][1]);
        }
  }
}

// pass 1: Waug[r, c] = w_r * A[r, c] (c < k), w_r * b_r (c == k), 0 (k < c < ldw); one warp per row
__global__ void __launch_bounds__(256) preweight_kernel(const double* __restrict__ A, int64_t lda,
                                                        const double* __restrict__ b,
                                                        const double* __restrict__ weff, int64_t n_rows, int k,
                                                        double* __restrict__ waug, int64_t ldw) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r0 = warp_global * 2; r0 < n_rows; r0 += nwarps * 2) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int64_t r = r0 + q;
      if (r >= n_rows) break;
      const double wv = __ldg(weff + r);
      const double* src = A + r * lda;
      double* dst = waug + r * ldw;
      for (int c0 = 0; c0 < ldw; c0 += 128) {
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + lane + 32 * u;
          v[u] = (c < k) ? __ldg(src + c) : ((c == k) ? __ldg(b + r) : 0.0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + lane + 32 * u;
          if (c < ldw) dst[c] = (c <= k) ? v[u] * wv : 0.0;     // fl(w*a): the reference's aw (svd.py:44)
        }
      }
    }
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

}  // namespace

int64_t fsb_gram_tma_ldw(int k) { return fsb_round_up((int64_t)k + 1, 16); }

bool fsb_gram_tma_available() { return get_encode() != nullptr && !getenv("FSB_GRAM_NO_TMA"); }

// `waug` must hold n_rows * ldw doubles, 128-byte aligned.
int fsb_launch_gram_tma(const fsb_context* h, const double* A, int64_t lda, const double* b, const double* weff,
                        int64_t n_rows, int k, int ntile, int nchunk, int64_t rows_per_chunk, double* partial,
                        double* waug, cudaStream_t s) {
  const int64_t ldw = fsb_gram_tma_ldw(k);
  if (n_rows > 0) {
    int64_t ctas = fsb_ceil_div(fsb_ceil_div(n_rows, 2), 8);
    const int64_t cap = (int64_t)h->sm_count * 8;
    if (ctas > cap) ctas = cap;
    preweight_kernel<<<(unsigned)ctas, 256, 0, s>>>(A, lda, b, weff, n_rows, k, waug, ldw);
    FSB_LAUNCH_CHECK("preweight_kernel");
  }
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {(cuuint64_t)ldw, (cuuint64_t)(n_rows > 0 ? n_rows : 1)};
  const cuuint64_t gstr[1] = {(cuuint64_t)(ldw * sizeof(double))};
  const cuuint32_t box[2] = {(cuuint32_t)FSB_GLDS, (cuuint32_t)T_RCH};
  const cuuint32_t estr[2] = {1, 1};
  CUresult cr = get_encode()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)waug, gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return FSB_ERR_UNSUPPORTED;
  TmaArgs a;
  a.n_rows = n_rows; a.k = k; a.ntile = ntile; a.rows_per_chunk = rows_per_chunk; a.partial = partial;
  const size_t smem = (size_t)T_NSTAGE * T_STAGE * sizeof(double);
  FSB_CUDA_TRY(cudaFuncSetAttribute(gram_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gram_tma_kernel<<<(unsigned)(nchunk * ntile), T_THREADS, smem, s>>>(tmap, a);
  FSB_LAUNCH_CHECK("gram_tma_kernel");
  return FSB_OK;
}
