// K2+K3+K4: fused training-mask + row weighting + augmented Gram  [aw|bw]^T [aw|bw]  in fp64.
//
// Replaces the reference prologue `aw = w[:,None]*A[training]; bw = w*b[training]`
// (fitsnap3lib/solvers/svd.py:35-46, ridge.py:28-39) and the contraction `aw.T @ aw`,
// `aw.T @ bw` (svd.py:50-51, ridge.py:42-43, examples/library/transpose_trick/example.py:233-234)
// without ever materialising aw.
//
// Design (B200, sm_100a):
//   * tcgen05 has no fp64 MMA kind, so the fp64-exact contraction runs on the fp64 tensor
//     sub-pipe that does exist on sm_100a: mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4; measured peak
//     on this pool 37.1 TFLOP/s, tools/ubench/fp64_rates.cu).
//   * The augmented matrix has ka = k+1 columns (column k is bw), so G, c = aw^T bw and
//     bw^T bw come out of ONE symmetric product.  Only lower-triangular 128x128 super-tiles
//     are computed; the final reduction mirrors them.
//   * grid = (#lower-tri super-tiles) x (#row chunks); blockIdx is tile-fastest so the CTAs
//     that share a row chunk run together and re-reads of A hit the 126 MB L2.
//   * warp-specialised CTA of 640 threads: 4 PRODUCER warps stage 16-row slices of the two column
//     ranges of the super-tile (plain coalesced LDG -> weight/mask/b-column/zero padding -> STS)
//     into a 6-deep shared-memory ring, 16 CONSUMER warps do nothing but LDS + DMMA.  FULL[s] /
//     EMPTY[s] named barriers (bar.sync / bar.arrive) hand the stages over.  cp.async (LDGSTS)
//     was measured at ~46 cycles per warp instruction per SM on this part and, issued by the
//     compute warps, serialised with the DMMAs; TMA needs 16-byte row pitches, which real FitSNAP
//     widths (31, 69, 110 ...) with lda = k do not give.  No alignment requirement here.
//   * weights (0 for test rows) are applied by the producers, once per element: fl(w*a), the
//     same rounding as the reference's aw; the consumers issue no fp64 op besides DMMA.
//   * smem row pitch 132 doubles (== 4 mod 16) makes every fragment load conflict-free.
//   * the 16 warp tiles (32x32 each) of a partial / diagonal super-tile carry unequal DMMA
//     counts: they are dealt to warps by a longest-processing-time greedy so the four DMMA
//     sub-pipes of the SM (warp % 4) stay balanced; the inner loop is specialised on the tile
//     shape so no predicate sits between DMMAs.
//   * split-K partials go to a workspace and are summed in a fixed order (deterministic).
#include "fsb_common.cuh"
#include <stdlib.h>

namespace {

struct GramArgs {
  const double* A;
  int64_t lda;
  const double* b;
  const double* w;      // effective weights: 0 for rows excluded from training
  int64_t n_rows;
  int k;
  int ntile;
  int64_t rows_per_chunk;
  double* partial;
};

constexpr int RCH = 16;                          // rows per stage
constexpr int NSTAGE = 6;                        // ring depth: 6 x 33.8 KB
constexpr int RANGE_DOUBLES = RCH * FSB_GLDS;    // one column range of one stage
constexpr int STAGE_DOUBLES = 2 * RANGE_DOUBLES; // I range, J range
constexpr int N_CONSUMERS = FSB_GTHREADS;        // 512: warps 0-15
constexpr int N_PRODUCERS = 128;                 // warps 16-19
constexpr int N_THREADS = N_CONSUMERS + N_PRODUCERS;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

__device__ __forceinline__ void tile_coords(int tile, int& ti, int& tj) {
  int t = (int)((sqrtf(8.0f * (float)tile + 1.0f) - 1.0f) * 0.5f);
  while ((t + 1) * (t + 2) / 2 <= tile) ++t;
  while (t * (t + 1) / 2 > tile) --t;
  ti = t;
  tj = tile - t * (t + 1) / 2;
}

// DMMA block of one stage for one warp tile of MI x NJ 8x8 blocks (OND: tile sits on the
// diagonal of a diagonal super-tile: only blocks j <= i, and the B fragments are the A fragments).
template <int MI, int NJ, bool OND>
__device__ __forceinline__ void compute_stage(double (&acc)[4][4][2], const double* __restrict__ fI,
                                              const double* __restrict__ fJ) {
#pragma unroll
  for (int ks = 0; ks < RCH / 4; ++ks) {
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

__global__ void __launch_bounds__(N_THREADS, 1) gram_dmma_kernel(GramArgs p) {
  extern __shared__ double smem[];  // [NSTAGE][ I: RCH x GLDS | J: RCH x GLDS ]

  const int tile = blockIdx.x % p.ntile;
  const int64_t chunk = blockIdx.x / p.ntile;
  int ti, tj;
  tile_coords(tile, ti, tj);
  const bool diag = (ti == tj);
  const int k = p.k;
  const int ka = k + 1;
  const int colI0 = ti * FSB_GT, colJ0 = tj * FSB_GT;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;

  int nbI = (ka - colI0 + 7) >> 3; nbI = nbI > 16 ? 16 : nbI;
  int nbJ = (ka - colJ0 + 7) >> 3; nbJ = nbJ > 16 ? 16 : nbJ;

  // Deal the 16 warp tiles to the consumer warps (LPT greedy over the four schedulers, warp % 4).
  __shared__ unsigned char s_map[16];
  if (tid == 0) {
    int cost[16], order[16], load[4] = {0, 0, 0, 0}, cnt[4] = {0, 0, 0, 0};
    for (int t = 0; t < 16; ++t) {
      const int r = t >> 2, c = t & 3;
      int m = nbI - 4 * r; m = m > 4 ? 4 : (m < 0 ? 0 : m);
      int n = nbJ - 4 * c; n = n > 4 ? 4 : (n < 0 ? 0 : n);
      cost[t] = (diag && c > r) ? 0 : ((diag && c == r) ? m * (m + 1) / 2 : m * n);
      order[t] = t;
    }
    for (int a = 1; a < 16; ++a) {  // insertion sort, descending cost
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

  const int64_t row_begin = chunk * p.rows_per_chunk;
  int64_t row_end = row_begin + p.rows_per_chunk;
  if (row_end > p.n_rows) row_end = p.n_rows;
  const int nsteps = row_end > row_begin ? (int)((row_end - row_begin + RCH - 1) / RCH) : 0;

  if (warp >= N_CONSUMERS / 32) {
    // ------------------------------ producers: LDG -> weight -> STS ------------------------------
    // thread c owns column c of the I range and column c of the J range; the source of a column is
    // a per-thread constant: a column of A, the b vector (column k), or nothing (zero padding).
    const int c = tid - N_CONSUMERS;
    const int gcI = colI0 + c, gcJ = colJ0 + c;
    const double* srcI; int64_t strI; bool useI;
    const double* srcJ; int64_t strJ; bool useJ;
    if (gcI < k) { srcI = p.A + gcI; strI = p.lda; useI = true; }
    else if (gcI == k) { srcI = p.b; strI = 1; useI = true; }
    else { srcI = p.w; strI = 0; useI = false; }
    if (diag) { srcJ = p.w; strJ = 0; useJ = false; }
    else if (gcJ < k) { srcJ = p.A + gcJ; strJ = p.lda; useJ = true; }
    else if (gcJ == k) { srcJ = p.b; strJ = 1; useJ = true; }
    else { srcJ = p.w; strJ = 0; useJ = false; }
    const int64_t last_row = p.n_rows > 0 ? p.n_rows - 1 : 0;
    for (int s = 0; s < nsteps; ++s) {
      const int slot = s % NSTAGE;
      const int64_t r0 = row_begin + (int64_t)s * RCH;
      const int64_t rw = r0 + (lane & (RCH - 1));       // row weights: lanes 0..15 hold the stage's w
      const double wl = (rw < row_end) ? __ldg(p.w + rw) : 0.0;
      double vI[RCH], vJ[RCH];
#pragma unroll
      for (int i = 0; i < RCH; ++i) {
        const int64_t r = r0 + i;
        const int64_t rc = r < row_end ? r : last_row;   // clamped; rows past the end get weight 0
        vI[i] = __ldg(srcI + rc * strI);
        if (!diag) vJ[i] = __ldg(srcJ + rc * strJ);
      }
      if (s >= NSTAGE) bar_sync(1 + NSTAGE + slot, N_THREADS);      // EMPTY[slot]
      double* stI = smem + (size_t)slot * STAGE_DOUBLES + c;
#pragma unroll
      for (int i = 0; i < RCH; ++i) {
        const double wi = __shfl_sync(0xffffffffu, wl, i);
        stI[i * FSB_GLDS] = useI ? vI[i] * wi : 0.0;                  // fl(w*a): the reference's aw (svd.py:44)
        if (!diag) stI[RANGE_DOUBLES + i * FSB_GLDS] = useJ ? vJ[i] * wi : 0.0;
      }
      __threadfence_block();
      bar_arrive(1 + slot, N_THREADS);                               // FULL[slot]
    }
    return;
  }

  // -------------------------------- consumers: LDS -> DMMA ---------------------------------------
  const int wt = s_map[warp];
  const int wr = wt >> 2, wc = wt & 3;
  int mi = nbI - 4 * wr; mi = mi > 4 ? 4 : mi;
  int nj = nbJ - 4 * wc; nj = nj > 4 ? 4 : nj;
  const bool active = (mi > 0) && (nj > 0) && (!diag || wc <= wr);
  const bool on_diag = diag && (wc == wr);
  // shape code for the specialised inner loops (warp-uniform)
  const int shape = !active ? -1 : (on_diag ? 16 + (mi - 1) : (mi - 1) * 4 + (nj - 1));

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int frag_off = (lane & 3) * FSB_GLDS + (lane >> 2);
  for (int s = 0; s < nsteps; ++s) {
    const int slot = s % NSTAGE;
    bar_sync(1 + slot, N_THREADS);                                   // FULL[slot]
    if (active) {
      const double* st = smem + (size_t)slot * STAGE_DOUBLES;
      const double* fI = st + frag_off + wr * 32;
      const double* fJ = (diag ? st : st + RANGE_DOUBLES) + frag_off + wc * 32;
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
    if (s + NSTAGE < nsteps) bar_arrive(1 + NSTAGE + slot, N_THREADS);   // EMPTY[slot]
  }

  // split-K partial: [chunk][tile][128][128]; only entries that the reduction reads need be valid
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

// effective weights: w for training rows, 0 for test rows (pt.fitsnap_dict['Testing'], svd.py:35-40)
__global__ void mask_weights_kernel(const double* __restrict__ w, const uint8_t* __restrict__ testing,
                                    int64_t n, double* __restrict__ weff) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) weff[i] = testing[i] ? 0.0 : w[i];
}

// Deterministic split-K reduction + symmetric mirror into the (k+1)x(k+1) output.  One block
// owns 32 columns of one row; its 8 warps each add every 8th chunk in index order and the 8 partial
// sums are combined in a fixed order, so the result does not depend on scheduling.
__global__ void __launch_bounds__(256) gram_reduce_kernel(const double* __restrict__ partial, int nchunk,
                                                          int ntile, int ka, double* __restrict__ gaug) {
  __shared__ double sh[8][33];
  const int i = blockIdx.y;
  const int jx = threadIdx.x & 31, gy = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + jx;
  const bool live = (i < ka) && (j <= i);
  double s = 0.0;
  if (live) {
    const int ti = i / FSB_GT, tj = j / FSB_GT;
    const int tile = ti * (ti + 1) / 2 + tj;
    const size_t off = (size_t)tile * (FSB_GT * FSB_GT) + (size_t)(i % FSB_GT) * FSB_GT + (j % FSB_GT);
    const size_t stride = (size_t)ntile * (FSB_GT * FSB_GT);
    for (int c = gy; c < nchunk; c += 8) s += partial[off + (size_t)c * stride];
  }
  sh[gy][jx] = s;
  __syncthreads();
  if (gy == 0 && live) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += sh[q][jx];
    gaug[(size_t)i * ka + j] = t;
    gaug[(size_t)j * ka + i] = t;
  }
}

struct GramPlan {
  int ntile;
  int nchunk;
  int64_t rows_per_chunk;
};

GramPlan plan_gram(const fsb_context* h, int64_t n_rows, int k) {
  GramPlan pl;
  const int ka = k + 1;
  const int nt = (ka + FSB_GT - 1) / FSB_GT;
  pl.ntile = nt * (nt + 1) / 2;
  // one resident CTA per SM (512 threads x <=128 regs); several waves when tiles are heterogeneous
  int64_t want = pl.ntile == 1 ? h->sm_count : fsb_ceil_div(8 * (int64_t)h->sm_count, pl.ntile);
  int64_t max_chunks = fsb_ceil_div(n_rows > 0 ? n_rows : 1, 4 * RCH);
  if (want > max_chunks) want = max_chunks;
  if (want < 1) want = 1;
  pl.rows_per_chunk = fsb_round_up(fsb_ceil_div(n_rows > 0 ? n_rows : 1, want), RCH);
  pl.nchunk = (int)fsb_ceil_div(n_rows > 0 ? n_rows : 1, pl.rows_per_chunk);
  return pl;
}

}  // namespace

// narrow matrices (<= 13 blocks of 8 augmented columns) take the row-split kernel of gram_small.cu
int fsb_gram_small_ctas(const fsb_context* h, int64_t n_rows);
int fsb_gram_small_groups();
int fsb_launch_gram_small(const fsb_context* h, const double* A, int64_t lda, const double* b, const double* weff,
                          int64_t n_rows, int k, double* partial, cudaStream_t s);
static bool use_small(int k) { return (k + 1 + 7) / 8 <= 13; }

// wide matrices: pre-weight pass + TMA-fed kernel (gram_tma.cu); the LDG-staged kernel above is the
// fallback when the driver entry point for tensor maps is unavailable or FSB_GRAM_NO_TMA is set
int64_t fsb_gram_tma_ldw(int k);
bool fsb_gram_tma_available();
int fsb_launch_gram_tma(const fsb_context* h, const double* A, int64_t lda, const double* b, const double* weff,
                        int64_t n_rows, int k, int ntile, int nchunk, int64_t rows_per_chunk, double* partial,
                        double* waug, cudaStream_t s);
static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
static size_t waug_bytes(int64_t n_rows, int k) {
  return (k + 1 <= FSB_GT) ? 0 : align256((size_t)(n_rows > 0 ? n_rows : 1) * (size_t)fsb_gram_tma_ldw(k) * sizeof(double));
}

// exact-integer Gram on the int8 tcgen05 tensor cores (gram_i8.cu)
bool fsb_gram_i8_available();
size_t fsb_gram_i8_ws_bytes(int64_t n_rows, int k);
int fsb_launch_gram_i8(const fsb_context* h, const double* A, int64_t lda, const double* b, const double* weff,
                       int64_t n_rows, int k, double* gaug, void* ws, size_t ws_bytes, cudaStream_t s);

// FSB_GRAM_AUTO: the int8 path pays a conversion pass (~100 integer operations per matrix element) and
// wins once the contraction dominates it: wide matrices with enough rows to fill the machine.
int fsb_gram_path_for(const fsb_context* h, int64_t n_rows, int k) {
  if (h->gram_path == FSB_GRAM_INT8) return fsb_gram_i8_available() ? FSB_GRAM_INT8 : FSB_GRAM_FP64;
  if (h->gram_path == FSB_GRAM_FP64) return FSB_GRAM_FP64;
  if (k + 1 >= FSB_I8_AUTO_MIN_COLS && n_rows >= FSB_I8_AUTO_MIN_ROWS && fsb_gram_i8_available())
    return FSB_GRAM_INT8;
  return FSB_GRAM_FP64;
}

static GramPlan effective_plan(const fsb_context* h, int64_t n_rows, int k) {
  GramPlan pl = plan_gram(h, n_rows, k);
  if (use_small(k)) {
    pl.ntile = 1;
    pl.nchunk = fsb_gram_small_ctas(h, n_rows) * fsb_gram_small_groups();
  }
  return pl;
}

static size_t partial_bytes(const GramPlan& pl) {
  return (size_t)pl.nchunk * pl.ntile * FSB_GT * FSB_GT * sizeof(double);
}

// fused K1 + K2..K4 for narrow matrices (gram_small.cu: scatter_gram_kernel) + the deterministic split-K reduction
bool fsb_scatter_gram_supported(const ScatterArgs& sc, int64_t total);
int fsb_launch_scatter_gram_small(const fsb_context* h, const ScatterArgs& sc, const uint8_t* testing, int64_t total,
                                  int store_a, double* partial, void* desc, cudaStream_t s);
size_t fsb_scatter_gram_desc_bytes(int64_t total);

int fsb_launch_scatter_gram(const fsb_context* h, const ScatterArgs& sc, const uint8_t* testing, int64_t total,
                            int store_a, double* gaug, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (!fsb_scatter_gram_supported(sc, total)) return FSB_ERR_UNSUPPORTED;
  const bool bzero = sc.flags & FSB_BZEROFLAG;
  const int k = sc.ncoeff * sc.numtypes + (bzero ? 0 : sc.numtypes);
  if (fsb_gram_path_for(h, total, k) != FSB_GRAM_FP64 || !use_small(k)) return FSB_ERR_UNSUPPORTED;
  GramPlan pl = effective_plan(h, total, k);
  // the per-row descriptors live where the unfused path keeps its masked weights / pre-weighted copy
  const size_t off_desc = align256(partial_bytes(pl));
  if (ws_bytes < off_desc + fsb_scatter_gram_desc_bytes(total)) return FSB_ERR_WORKSPACE_TOO_SMALL;
  int st = fsb_launch_scatter_gram_small(h, sc, testing, total, store_a, (double*)ws, (char*)ws + off_desc, s);
  if (st != FSB_OK) return st;
  const int ka = k + 1;
  dim3 rgrid((unsigned)fsb_ceil_div(ka, 32), (unsigned)ka);
  gram_reduce_kernel<<<rgrid, 256, 0, s>>>((const double*)ws, pl.nchunk, 1, ka, gaug);
  FSB_LAUNCH_CHECK("gram_reduce_kernel");
  return FSB_OK;
}

size_t fsb_gram_ws_bytes(const fsb_context* h, int64_t n_rows, int k) {
  if (fsb_gram_path_for(h, n_rows, k) == FSB_GRAM_INT8)
    return align256((size_t)(n_rows > 0 ? n_rows : 1) * sizeof(double)) + fsb_gram_i8_ws_bytes(n_rows, k);
  GramPlan pl = effective_plan(h, n_rows, k);
  // split-K partials + masked weight vector (test mask given) + pre-weighted copy (wide matrices); narrow matrices:
  // room for the 24-byte row descriptors of the fused scatter + Gram kernel instead
  const size_t tail = align256((size_t)(n_rows > 0 ? n_rows : 1) * sizeof(double)) + waug_bytes(n_rows, k);
  const size_t desc = use_small(k) ? align256(fsb_scatter_gram_desc_bytes(n_rows)) : 0;
  return align256(partial_bytes(pl)) + (tail > desc ? tail : desc);
}

int fsb_launch_gram(const fsb_context* h, const double* A, int64_t lda, const double* b, const double* w,
                    const uint8_t* testing, int64_t n_rows, int k, double* gaug, void* ws, size_t ws_bytes,
                    cudaStream_t s) {
  if (fsb_gram_path_for(h, n_rows, k) == FSB_GRAM_INT8) {
    const size_t wbytes = align256((size_t)(n_rows > 0 ? n_rows : 1) * sizeof(double));
    if (ws_bytes < wbytes + fsb_gram_i8_ws_bytes(n_rows, k)) return FSB_ERR_WORKSPACE_TOO_SMALL;
    const double* weff8 = w;
    if (testing && n_rows > 0) {
      mask_weights_kernel<<<(unsigned)fsb_ceil_div(n_rows, 256), 256, 0, s>>>(w, testing, n_rows, (double*)ws);
      FSB_LAUNCH_CHECK("mask_weights_kernel");
      weff8 = (const double*)ws;
    }
    return fsb_launch_gram_i8(h, A, lda, b, weff8, n_rows, k, gaug, (char*)ws + wbytes, ws_bytes - wbytes, s);
  }
  GramPlan pl = effective_plan(h, n_rows, k);
  const size_t off_w = align256(partial_bytes(pl));
  const size_t off_waug = off_w + align256((size_t)(n_rows > 0 ? n_rows : 1) * sizeof(double));
  const size_t need = off_waug + waug_bytes(n_rows, k);
  if (ws_bytes < need) return FSB_ERR_WORKSPACE_TOO_SMALL;
  const double* weff = w;
  if (testing && n_rows > 0) {
    double* wbuf = (double*)((char*)ws + off_w);
    mask_weights_kernel<<<(unsigned)fsb_ceil_div(n_rows, 256), 256, 0, s>>>(w, testing, n_rows, wbuf);
    FSB_LAUNCH_CHECK("mask_weights_kernel");
    weff = wbuf;
  }
  const int ka = k + 1;
  dim3 rgrid((unsigned)fsb_ceil_div(ka, 32), (unsigned)ka);
  if (use_small(k) && !getenv("FSB_GRAM_FORCE_TILED")) {
    int st = fsb_launch_gram_small(h, A, lda, b, weff, n_rows, k, (double*)ws, s);
    if (st != FSB_OK) return st;
    gram_reduce_kernel<<<rgrid, 256, 0, s>>>((const double*)ws, pl.nchunk, 1, ka, gaug);
    FSB_LAUNCH_CHECK("gram_reduce_kernel");
    return FSB_OK;
  }
  pl = plan_gram(h, n_rows, k);
  // one super-tile (k + 1 <= 128): every column is read once, the pre-weight pass would double the
  // traffic for nothing -> LDG-staged kernel; three tiles and more: pre-weight + TMA
  if (pl.ntile >= 3 && fsb_gram_tma_available() && (reinterpret_cast<uintptr_t>(ws) & 127) == 0) {
    int st = fsb_launch_gram_tma(h, A, lda, b, weff, n_rows, k, pl.ntile, pl.nchunk, pl.rows_per_chunk,
                                 (double*)ws, (double*)((char*)ws + off_waug), s);
    if (st != FSB_OK) return st;
    gram_reduce_kernel<<<rgrid, 256, 0, s>>>((const double*)ws, pl.nchunk, pl.ntile, ka, gaug);
    FSB_LAUNCH_CHECK("gram_reduce_kernel");
    return FSB_OK;
  }
  GramArgs a;
  a.A = A; a.lda = lda; a.b = b; a.w = weff; a.n_rows = n_rows; a.k = k;
  a.ntile = pl.ntile; a.rows_per_chunk = pl.rows_per_chunk; a.partial = (double*)ws;
  const size_t smem = (size_t)NSTAGE * STAGE_DOUBLES * sizeof(double);
  FSB_CUDA_TRY(cudaFuncSetAttribute(gram_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gram_dmma_kernel<<<(unsigned)(pl.nchunk * pl.ntile), N_THREADS, smem, s>>>(a);
  FSB_LAUNCH_CHECK("gram_dmma_kernel");
  gram_reduce_kernel<<<rgrid, 256, 0, s>>>((const double*)ws, pl.nchunk, pl.ntile, ka, gaug);
  FSB_LAUNCH_CHECK("gram_reduce_kernel");
  return FSB_OK;
}
