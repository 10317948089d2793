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
//   * each CTA streams its row chunk through a 3-stage cp.async (LDGSTS) ring of 32-row stages:
//     global -> shared memory with no register staging, zero-fill for rows/columns past the
//     edge, one __syncthreads per stage.  lda is arbitrary (K = 31, 69, 110 ... are real
//     FitSNAP widths), so the copies are 8-byte; TMA needs 16-byte pitches and is not used here.
//   * row weights (0 for test rows) ride along in the stage and are applied when a DMMA
//     fragment is read from shared memory: fl(w*a), the same rounding as the reference's aw.
//     The multiplies run on the plain fp64 pipe, which is separate from the DMMA sub-pipe.
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

constexpr int RCH = FSB_GRCH;                    // rows per stage
constexpr int NSTAGE = 3;
constexpr int RANGE_DOUBLES = RCH * FSB_GLDS;    // one column range of one stage
constexpr int STAGE_DOUBLES = 2 * RANGE_DOUBLES + RCH;   // I range, J range, weights

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// 8-byte asynchronous global->shared copy; src_bytes = 0 zero-fills without touching memory.
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc, bool valid) {
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int nbytes = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(saddr), "l"(gsrc), "r"(nbytes) : "memory");
}
// 16-byte variant (both addresses 16-byte aligned)
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc, int nbytes) {
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(gsrc), "r"(nbytes) : "memory");
}
// plain 16-byte copy (no src-size operand): the common case of a full stage
__device__ __forceinline__ void cp_async16_full(double* smem_dst, const double* gsrc) {
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8_full(double* smem_dst, const double* gsrc) {
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(saddr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

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
                                              const double* __restrict__ fJ, const double* __restrict__ fW) {
#pragma unroll 2
  for (int ks = 0; ks < RCH / 4; ++ks) {
    const double wv = fW[ks * 4];
    double af[MI], bf[NJ];
#pragma unroll
    for (int i = 0; i < MI; ++i) af[i] = fI[ks * 4 * FSB_GLDS + i * 8] * wv;
    if (OND) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) bf[j] = af[j];
    } else {
#pragma unroll
      for (int j = 0; j < NJ; ++j) bf[j] = fJ[ks * 4 * FSB_GLDS + j * 8] * wv;
    }
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (!OND || j <= i) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
  }
}

template <bool VEC16>
__global__ void __launch_bounds__(FSB_GTHREADS, 1) gram_dmma_kernel(GramArgs p) {
  extern __shared__ double smem[];  // [NSTAGE][ I: RCH x GLDS | J: RCH x GLDS | w: RCH ]

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

  // Deal the 16 warp tiles to warps (LPT greedy over the four schedulers, warp % 4).
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
  const int wt = s_map[warp];
  const int wr = wt >> 2, wc = wt & 3;
  int mi = nbI - 4 * wr; mi = mi > 4 ? 4 : mi;
  int nj = nbJ - 4 * wc; nj = nj > 4 ? 4 : nj;
  const bool active = (mi > 0) && (nj > 0) && (!diag || wc <= wr);
  const bool on_diag = diag && (wc == wr);
  // shape code for the specialised inner loops (warp-uniform)
  const int shape = !active ? -1 : (on_diag ? 16 + (mi - 1) : (mi - 1) * 4 + (nj - 1));

  const int64_t row_begin = chunk * p.rows_per_chunk;
  int64_t row_end = row_begin + p.rows_per_chunk;
  if (row_end > p.n_rows) row_end = p.n_rows;
  const int nsteps = row_end > row_begin ? (int)((row_end - row_begin + RCH - 1) / RCH) : 0;

  // ---- copy plan ------------------------------------------------------------------------------
  // A "slot" is what one thread copies per stage row: 8 bytes of one column (scalar plan) or 16
  // bytes of a column pair (VEC16: lda even and A 16-byte aligned).  The source of a slot is a
  // per-thread constant: a column of A, the b vector (column k), or nothing (zero-fill).
  constexpr int SLOTS_PER_ROW = VEC16 ? FSB_GT / 2 : FSB_GT;
  constexpr int ROW_GROUPS = FSB_GTHREADS / SLOTS_PER_ROW;      // 8 (VEC16) or 4
  constexpr int COPIES = RCH / ROW_GROUPS;                      // 4 (VEC16) or 8 per range per stage
  const int slot = tid % SLOTS_PER_ROW;
  const int srow = tid / SLOTS_PER_ROW;
  const int scol = VEC16 ? 2 * slot : slot;
  const int64_t last_row = p.n_rows > 0 ? p.n_rows - 1 : 0;

  struct Src { const double* ptr; int64_t stride; int nbytes; };
  auto scalar_src = [&](int gc) {
    Src r;
    if (gc < k) { r.ptr = p.A + gc; r.stride = p.lda; r.nbytes = 8; }
    else if (gc == k) { r.ptr = p.b; r.stride = 1; r.nbytes = 8; }
    else { r.ptr = p.w; r.stride = 0; r.nbytes = 0; }
    return r;
  };
  // per range: mode 0 = one 16-byte copy from A, 1 = two scalar slots (pair straddles column k),
  // 2 = scalar plan (one 8-byte slot)
  Src sI0, sI1, sJ0, sJ1;
  int modeI, modeJ;
  {
    const int gc = colI0 + scol;
    if (VEC16) {
      if (gc + 1 < k) { modeI = 0; sI0.ptr = p.A + gc; sI0.stride = p.lda; sI0.nbytes = 16; sI1 = sI0; }
      else if (gc > k) { modeI = 0; sI0.ptr = p.w; sI0.stride = 0; sI0.nbytes = 0; sI1 = sI0; }
      else { modeI = 1; sI0 = scalar_src(gc); sI1 = scalar_src(gc + 1); }
    } else { modeI = 2; sI0 = scalar_src(gc); sI1 = sI0; }
    const int gj = colJ0 + scol;
    if (VEC16) {
      if (gj + 1 < k) { modeJ = 0; sJ0.ptr = p.A + gj; sJ0.stride = p.lda; sJ0.nbytes = 16; sJ1 = sJ0; }
      else if (gj > k) { modeJ = 0; sJ0.ptr = p.w; sJ0.stride = 0; sJ0.nbytes = 0; sJ1 = sJ0; }
      else { modeJ = 1; sJ0 = scalar_src(gj); sJ1 = scalar_src(gj + 1); }
    } else { modeJ = 2; sJ0 = scalar_src(gj); sJ1 = sJ0; }
  }

  auto copy_slot = [&](double* dst, const Src& s0, const Src& s1, int mode, int64_t rc, bool in) {
    if (mode == 0) cp_async16(dst, s0.ptr + rc * s0.stride, in ? s0.nbytes : 0);
    else {
      cp_async8(dst, s0.ptr + rc * s0.stride, in && s0.nbytes);
      if (mode == 1) cp_async8(dst + 1, s1.ptr + rc * s1.stride, in && s1.nbytes);
    }
  };

  // Running source pointers: the rows a thread copies form one arithmetic progression across
  // stages (srow, srow+RG, ..., then the same rows of the next stage), so every copy costs one
  // 64-bit add instead of a 64-bit multiply-add plus clamps.  Stages are issued in order.
  const double* runI = sI0.ptr + (row_begin + srow) * sI0.stride;
  const double* runJ = sJ0.ptr + (row_begin + srow) * sJ0.stride;
  const int64_t stepI = (int64_t)ROW_GROUPS * sI0.stride;
  const int64_t stepJ = (int64_t)ROW_GROUPS * sJ0.stride;
  const double* runW = p.w + row_begin + (tid < RCH ? tid : 0);

  auto issue_stage = [&](int step) {
    if (step < nsteps) {
      double* st = smem + (size_t)(step % NSTAGE) * STAGE_DOUBLES;
      const int64_t r0 = row_begin + (int64_t)step * RCH;
      const bool full = (r0 + RCH <= row_end);       // every stage but the last of a chunk
      if (full) {
#pragma unroll
        for (int i = 0; i < COPIES; ++i) {
          double* dI = st + (srow + ROW_GROUPS * i) * FSB_GLDS + scol;
          if (modeI == 0) { if (sI0.nbytes) cp_async16_full(dI, runI); else cp_async16(dI, runI, 0); }
          else if (modeI == 2) { if (sI0.nbytes) cp_async8_full(dI, runI); else cp_async8(dI, runI, false); }
          else copy_slot(dI, sI0, sI1, 1, r0 + srow + ROW_GROUPS * i, true);
          runI += stepI;
          if (!diag) {
            double* dJ = dI + RANGE_DOUBLES;
            if (modeJ == 0) { if (sJ0.nbytes) cp_async16_full(dJ, runJ); else cp_async16(dJ, runJ, 0); }
            else if (modeJ == 2) { if (sJ0.nbytes) cp_async8_full(dJ, runJ); else cp_async8(dJ, runJ, false); }
            else copy_slot(dJ, sJ0, sJ1, 1, r0 + srow + ROW_GROUPS * i, true);
            runJ += stepJ;
          }
        }
        if (tid < RCH) cp_async8(st + 2 * RANGE_DOUBLES + tid, runW, true);
        runW += RCH;
      } else {   // ragged last stage of the chunk: per-row predicate, clamped addresses
#pragma unroll
        for (int i = 0; i < COPIES; ++i) {
          const int lr = srow + ROW_GROUPS * i;
          const int64_t r = r0 + lr;
          const bool in = r < row_end;
          const int64_t rc = in ? r : last_row;
          copy_slot(st + lr * FSB_GLDS + scol, sI0, sI1, modeI, rc, in);
          if (!diag) copy_slot(st + RANGE_DOUBLES + lr * FSB_GLDS + scol, sJ0, sJ1, modeJ, rc, in);
        }
        if (tid < RCH) {
          const int64_t r = r0 + tid;
          const bool in = r < row_end;
          cp_async8(st + 2 * RANGE_DOUBLES + tid, p.w + (in ? r : last_row), in);
        }
      }
    }
    cp_async_commit();   // always commit (possibly empty) so the group accounting stays uniform
  };

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < NSTAGE - 1; ++s) issue_stage(s);

  const int frag_off = (lane & 3) * FSB_GLDS + (lane >> 2);
  for (int s = 0; s < nsteps; ++s) {
    cp_async_wait<NSTAGE - 2>();   // this thread's copies of stage s have landed
    __syncthreads();               // everyone's have, and everyone is done reading stage s-1
    issue_stage(s + NSTAGE - 1);   // refill the buffer stage s-1 used
    if (active) {
      const double* st = smem + (size_t)(s % NSTAGE) * STAGE_DOUBLES;
      const double* fI = st + frag_off + wr * 32;
      const double* fJ = (diag ? st : st + RANGE_DOUBLES) + frag_off + wc * 32;
      const double* fW = st + 2 * RANGE_DOUBLES + (lane & 3);
      switch (shape) {
#define FSB_CASE(MI, NJ) case (MI - 1) * 4 + (NJ - 1): compute_stage<MI, NJ, false>(acc, fI, fJ, fW); break;
        FSB_CASE(4, 4) FSB_CASE(4, 3) FSB_CASE(4, 2) FSB_CASE(4, 1)
        FSB_CASE(3, 4) FSB_CASE(3, 3) FSB_CASE(3, 2) FSB_CASE(3, 1)
        FSB_CASE(2, 4) FSB_CASE(2, 3) FSB_CASE(2, 2) FSB_CASE(2, 1)
        FSB_CASE(1, 4) FSB_CASE(1, 3) FSB_CASE(1, 2) FSB_CASE(1, 1)
#undef FSB_CASE
        case 16: compute_stage<1, 1, true>(acc, fI, fJ, fW); break;
        case 17: compute_stage<2, 2, true>(acc, fI, fJ, fW); break;
        case 18: compute_stage<3, 3, true>(acc, fI, fJ, fW); break;
        case 19: compute_stage<4, 4, true>(acc, fI, fJ, fW); break;
        default: break;
      }
    }
  }
  cp_async_wait<0>();

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
  int64_t max_chunks = fsb_ceil_div(n_rows > 0 ? n_rows : 1, 4 * FSB_GRCH);
  if (want > max_chunks) want = max_chunks;
  if (want < 1) want = 1;
  pl.rows_per_chunk = fsb_round_up(fsb_ceil_div(n_rows > 0 ? n_rows : 1, want), FSB_GRCH);
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

size_t fsb_gram_ws_bytes(const fsb_context* h, int64_t n_rows, int k) {
  GramPlan pl = effective_plan(h, n_rows, k);
  // split-K partials + room for the masked weight vector (used only when a test mask is given)
  return partial_bytes(pl) + (size_t)(n_rows > 0 ? n_rows : 1) * sizeof(double);
}

int fsb_launch_gram(const fsb_context* h, const double* A, int64_t lda, const double* b, const double* w,
                    const uint8_t* testing, int64_t n_rows, int k, double* gaug, void* ws, size_t ws_bytes,
                    cudaStream_t s) {
  GramPlan pl = effective_plan(h, n_rows, k);
  const size_t need = partial_bytes(pl) + (size_t)(n_rows > 0 ? n_rows : 1) * sizeof(double);
  if (ws_bytes < need) return FSB_ERR_WORKSPACE_TOO_SMALL;
  const double* weff = w;
  if (testing && n_rows > 0) {
    double* wbuf = (double*)((char*)ws + partial_bytes(pl));
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
  GramArgs a;
  a.A = A; a.lda = lda; a.b = b; a.w = weff; a.n_rows = n_rows; a.k = k;
  a.ntile = pl.ntile; a.rows_per_chunk = pl.rows_per_chunk; a.partial = (double*)ws;
  const size_t smem = (size_t)NSTAGE * STAGE_DOUBLES * sizeof(double);
  const bool vec16 = (lda % 2 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  if (vec16) {
    FSB_CUDA_TRY(cudaFuncSetAttribute(gram_dmma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gram_dmma_kernel<true><<<(unsigned)(pl.nchunk * pl.ntile), FSB_GTHREADS, smem, s>>>(a);
  } else {
    FSB_CUDA_TRY(cudaFuncSetAttribute(gram_dmma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gram_dmma_kernel<false><<<(unsigned)(pl.nchunk * pl.ntile), FSB_GTHREADS, smem, s>>>(a);
  }
  FSB_LAUNCH_CHECK("gram_dmma_kernel");
  gram_reduce_kernel<<<rgrid, 256, 0, s>>>((const double*)ws, pl.nchunk, pl.ntile, ka, gaug);
  FSB_LAUNCH_CHECK("gram_reduce_kernel");
  return FSB_OK;
}
