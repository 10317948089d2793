// K2+K3+K4: fused training-mask + row weighting + augmented Gram  [aw|bw]^T [aw|bw]  in fp64.
//
// Replaces the reference prologue `aw = w[:,None]*A[training]; bw = w*b[training]`
// (fitsnap3lib/solvers/svd.py:35-46, ridge.py:28-39) and the contraction `aw.T @ aw`,
// `aw.T @ bw` (svd.py:50-51, ridge.py:42-43, examples/library/transpose_trick/example.py:233-234)
// without ever materialising aw.
//
// Design (B200, sm_100a):
//   * tcgen05 has no fp64 MMA kind, so the fp64-exact contraction runs on the fp64 tensor
//     path that does exist on sm_100a: mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).
//   * The augmented matrix has ka = k+1 columns (column k is bw), so G, c = aw^T bw and
//     bw^T bw come out of ONE symmetric product.  Only lower-triangular 128x128 super-tiles
//     are computed; the final reduction mirrors them.
//   * grid = (#lower-tri super-tiles) x (#row chunks); blockIdx is tile-fastest so the CTAs
//     that share a row chunk run together and re-reads of A hit the 126 MB L2.
//   * each CTA streams its row chunk through a double-buffered shared-memory stage of 16 rows;
//     weighting / masking / the b column are applied on the way in (registers), so HBM sees
//     A exactly once per super-tile column range and nothing else.
//   * smem row pitch 132 doubles (== 4 mod 16) makes every DMMA fragment load conflict-free.
//   * split-K partials go to a workspace and are summed in a fixed order (deterministic).
#include "fsb_common.cuh"

namespace {

struct GramArgs {
  const double* A;
  int64_t lda;
  const double* b;
  const double* w;
  const uint8_t* testing;
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

__device__ __forceinline__ void tile_coords(int tile, int& ti, int& tj) {
  int t = (int)((sqrtf(8.0f * (float)tile + 1.0f) - 1.0f) * 0.5f);
  while ((t + 1) * (t + 2) / 2 <= tile) ++t;
  while (t * (t + 1) / 2 > tile) --t;
  ti = t;
  tj = tile - t * (t + 1) / 2;
}

constexpr int RANGE_DOUBLES = FSB_GRCH * FSB_GLDS;  // one column range of one stage
constexpr int ROWS_PER_THREAD = FSB_GRCH / (FSB_GTHREADS / FSB_GT);  // 16 / 4 = 4

__global__ void __launch_bounds__(FSB_GTHREADS, 1) gram_dmma_kernel(GramArgs p) {
  extern __shared__ double smem[];  // [2 stages][2 ranges][GRCH][GLDS]

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
  const int wr = warp >> 2, wc = warp & 3;  // 4x4 grid of 32x32 warp tiles

  int nbI = (ka - colI0 + 7) >> 3; nbI = nbI > 16 ? 16 : nbI;
  int nbJ = (ka - colJ0 + 7) >> 3; nbJ = nbJ > 16 ? 16 : nbJ;
  int mi = nbI - 4 * wr; mi = mi > 4 ? 4 : mi;
  int nj = nbJ - 4 * wc; nj = nj > 4 ? 4 : nj;
  const bool active = (mi > 0) && (nj > 0) && (!diag || wc <= wr);
  const bool on_diag = diag && (wc == wr);

  const int64_t row_begin = chunk * p.rows_per_chunk;
  int64_t row_end = row_begin + p.rows_per_chunk;
  if (row_end > p.n_rows) row_end = p.n_rows;
  const int nsteps = row_end > row_begin ? (int)((row_end - row_begin + FSB_GRCH - 1) / FSB_GRCH) : 0;

  // staging map: thread -> one column of the super-tile, 4 rows of the stage
  const int scol = tid & (FSB_GT - 1);
  const int srow = tid >> 7;  // 0..3
  const int gcI = colI0 + scol, gcJ = colJ0 + scol;

  double pfI[ROWS_PER_THREAD], pfJ[ROWS_PER_THREAD];

  auto load_stage = [&](int step) {
    const int64_t r0 = row_begin + (int64_t)step * FSB_GRCH + srow;
#pragma unroll
    for (int i = 0; i < ROWS_PER_THREAD; ++i) {
      const int64_t r = r0 + 4 * i;
      double vI = 0.0, vJ = 0.0;
      if (r < row_end) {
        const bool keep = p.testing ? (p.testing[r] == 0) : true;
        if (keep) {
          const double wv = __ldg(p.w + r);
          if (gcI < k) vI = __ldg(p.A + r * p.lda + gcI) * wv;
          else if (gcI == k) vI = __ldg(p.b + r) * wv;
          if (!diag) {
            if (gcJ < k) vJ = __ldg(p.A + r * p.lda + gcJ) * wv;
            else if (gcJ == k) vJ = __ldg(p.b + r) * wv;
          }
        }
      }
      pfI[i] = vI;
      pfJ[i] = vJ;
    }
  };
  auto store_stage = [&](int buf) {
    double* sI = smem + buf * (2 * RANGE_DOUBLES);
    double* sJ = sI + RANGE_DOUBLES;
#pragma unroll
    for (int i = 0; i < ROWS_PER_THREAD; ++i) {
      sI[(srow + 4 * i) * FSB_GLDS + scol] = pfI[i];
      if (!diag) sJ[(srow + 4 * i) * FSB_GLDS + scol] = pfJ[i];
    }
  };

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  if (nsteps > 0) {
    load_stage(0);
    store_stage(0);
  }
  __syncthreads();

  const int frag_off = (lane & 3) * FSB_GLDS + (lane >> 2);
  for (int s = 0; s < nsteps; ++s) {
    const bool more = (s + 1 < nsteps);
    if (more) load_stage(s + 1);
    if (active) {
      const double* sI = smem + (s & 1) * (2 * RANGE_DOUBLES);
      const double* sJ = diag ? sI : sI + RANGE_DOUBLES;
      const double* fI = sI + frag_off + wr * 32;
      const double* fJ = sJ + frag_off + wc * 32;
#pragma unroll
      for (int ks = 0; ks < FSB_GRCH / 4; ++ks) {
        double af[4], bf[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          af[i] = fI[ks * 4 * FSB_GLDS + i * 8];
          bf[i] = fJ[ks * 4 * FSB_GLDS + i * 8];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (i < mi && j < nj && (!on_diag || j <= i)) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
    }
    if (more) store_stage((s + 1) & 1);
    __syncthreads();
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

// Deterministic split-K reduction + symmetric mirror into the (k+1)x(k+1) output.
__global__ void gram_reduce_kernel(const double* __restrict__ partial, int nchunk, int ntile, int ka,
                                   double* __restrict__ gaug) {
  const int i = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ka || j > i) return;
  const int ti = i / FSB_GT, tj = j / FSB_GT;
  const int tile = ti * (ti + 1) / 2 + tj;
  const size_t off = (size_t)tile * (FSB_GT * FSB_GT) + (size_t)(i % FSB_GT) * FSB_GT + (j % FSB_GT);
  const size_t stride = (size_t)ntile * (FSB_GT * FSB_GT);
  double s = 0.0;
  for (int c = 0; c < nchunk; ++c) s += partial[off + (size_t)c * stride];
  gaug[(size_t)i * ka + j] = s;
  gaug[(size_t)j * ka + i] = s;
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
  int64_t want = pl.ntile == 1 ? h->sm_count : fsb_ceil_div(4 * (int64_t)h->sm_count, pl.ntile);
  int64_t max_chunks = fsb_ceil_div(n_rows > 0 ? n_rows : 1, 4 * FSB_GRCH);
  if (want > max_chunks) want = max_chunks;
  if (want < 1) want = 1;
  pl.rows_per_chunk = fsb_round_up(fsb_ceil_div(n_rows > 0 ? n_rows : 1, want), FSB_GRCH);
  pl.nchunk = (int)fsb_ceil_div(n_rows > 0 ? n_rows : 1, pl.rows_per_chunk);
  return pl;
}

}  // namespace

size_t fsb_gram_ws_bytes(const fsb_context* h, int64_t n_rows, int k) {
  GramPlan pl = plan_gram(h, n_rows, k);
  return (size_t)pl.nchunk * pl.ntile * FSB_GT * FSB_GT * sizeof(double);
}

int fsb_launch_gram(const fsb_context* h, const double* A, int64_t lda, const double* b, const double* w,
                    const uint8_t* testing, int64_t n_rows, int k, double* gaug, void* ws, size_t ws_bytes,
                    cudaStream_t s) {
  GramPlan pl = plan_gram(h, n_rows, k);
  const size_t need = (size_t)pl.nchunk * pl.ntile * FSB_GT * FSB_GT * sizeof(double);
  if (ws_bytes < need) return FSB_ERR_WORKSPACE_TOO_SMALL;
  GramArgs a;
  a.A = A; a.lda = lda; a.b = b; a.w = w; a.testing = testing; a.n_rows = n_rows; a.k = k;
  a.ntile = pl.ntile; a.rows_per_chunk = pl.rows_per_chunk; a.partial = (double*)ws;
  const size_t smem = (size_t)2 * 2 * RANGE_DOUBLES * sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    FSB_CUDA_TRY(cudaFuncSetAttribute(gram_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  gram_dmma_kernel<<<(unsigned)(pl.nchunk * pl.ntile), FSB_GTHREADS, smem, s>>>(a);
  FSB_LAUNCH_CHECK("gram_dmma_kernel");
  const int ka = k + 1;
  dim3 rgrid((unsigned)fsb_ceil_div(ka, 128), (unsigned)ka);
  gram_reduce_kernel<<<rgrid, 128, 0, s>>>((const double*)ws, pl.nchunk, pl.ntile, ka, gaug);
  FSB_LAUNCH_CHECK("gram_reduce_kernel");
  return FSB_OK;
}
