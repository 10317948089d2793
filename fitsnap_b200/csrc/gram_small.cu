// K2+K3+K4 for narrow design matrices (k + 1 <= 104, i.e. <= 13 DMMA blocks of 8 columns):
// the whole lower triangle of the augmented Gram lives in the registers of a PAIR of warps.
//
// Why a second kernel: with one 128x128 super-tile (gram.cu) only 10 of its 16 warp tiles exist at
// k = 100 and the busiest scheduler carries 26 of the 91 DMMAs of a k-step (ncu: DMMA sub-pipe 49 %
// active).  Here the ROWS are split instead of the output: a CTA has 4 warp pairs; pair g takes
// k-steps g, g+4, ... of every stage, and the two warps of a pair split the 91 (i, j <= i) blocks
// 45 / 46 by block row.  Every warp issues the same number of DMMAs per k-step, both warps of a pair
// sit on the same scheduler (w, w+4), so the four DMMA sub-pipes carry identical work, all 8 warps
// are busy, and a warp reuses 13 fragments for 46 DMMAs.  The 4 partial Grams of a CTA go to the
// split-K workspace as 4 "chunks"; gram_reduce_kernel (gram.cu) sums them in a fixed order.
//
// Staging is the same 8/16-byte cp.async ring as gram.cu (64-row stages, pitch 8*NB+4 doubles,
// == 4 or 12 mod 16: conflict-free fragment loads), weights applied as fl(w*a) at fragment load.
#include "fsb_common.cuh"

namespace {

struct SmallArgs {
  const double* A;
  int64_t lda;
  const double* b;
  const double* w;      // effective weights (0 for excluded rows)
  int64_t n_rows;
  int k;
  int64_t rows_per_cta;
  double* partial;      // [cta * 4 + group][128][128]
};

constexpr int S_THREADS = 256;
constexpr int S_RCH = 64;       // rows per stage
constexpr int S_NSTAGE = 3;
constexpr int S_GROUPS = 4;     // warp pairs

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc, bool valid) {
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int nbytes = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(saddr), "l"(gsrc), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc, int nbytes) {
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(gsrc), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// block rows [0, split) go to warp 0 of a pair, [split, NB) to warp 1: split balances i(i+1)/2
__host__ __device__ constexpr int split_row(int nb) {
  int total = nb * (nb + 1) / 2, r = 0;
  while ((r + 1) * (r + 2) / 2 <= total / 2) ++r;
  // r(r+1)/2 <= total/2 < (r+1)(r+2)/2 ; pick the closer one
  return ((r + 1) * (r + 2) / 2 - total / 2) < (total / 2 - r * (r + 1) / 2) ? r + 1 : r;
}

template <int R0, int R1, int NACC>
__device__ __forceinline__ void part_kstep(double (&acc)[NACC][2], const double* __restrict__ frow, double wv) {
  // fragments of the block columns this part touches: 0 .. R1-1
  double f[R1 > 0 ? R1 : 1];
#pragma unroll
  for (int c = 0; c < R1; ++c) f[c] = frow[c * 8] * wv;
#pragma unroll
  for (int i = R0; i < R1; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      constexpr int base = R0 * (R0 + 1) / 2;
      dmma884(acc[i * (i + 1) / 2 + j - base][0], acc[i * (i + 1) / 2 + j - base][1], f[i], f[j]);
    }
}

template <int R0, int R1, typename IssueFn>
__device__ __forceinline__ void run_part(const SmallArgs& p, double* smem, int pitch, int group, int lane,
                                         int nsteps, int64_t cta, IssueFn& issue_stage) {
  constexpr int NACC = (R1 * (R1 + 1) - R0 * (R0 + 1)) / 2 > 0 ? (R1 * (R1 + 1) - R0 * (R0 + 1)) / 2 : 1;
  double acc[NACC][2];
#pragma unroll
  for (int q = 0; q < NACC; ++q) acc[q][0] = acc[q][1] = 0.0;
  const int stage_doubles = S_RCH * pitch + S_RCH;
  const int frag_off = (lane & 3) * pitch + (lane >> 2);
  for (int s = 0; s < nsteps; ++s) {
    cp_async_wait<S_NSTAGE - 2>();
    __syncthreads();
    issue_stage(s + S_NSTAGE - 1);
    const double* st = smem + (size_t)(s % S_NSTAGE) * stage_doubles;
    const double* sw = st + S_RCH * pitch;
    if (R1 > R0) {
#pragma unroll
      for (int q = 0; q < S_RCH / 4 / S_GROUPS; ++q) {
        const int ks = group + S_GROUPS * q;          // this pair's k-steps of the stage
        part_kstep<R0, R1, NACC>(acc, st + ks * 4 * pitch + frag_off, sw[ks * 4 + (lane & 3)]);
      }
    }
  }
  cp_async_wait<0>();
  // partial Gram of this pair -> workspace chunk (cta*4 + group), tile 0, 128x128 layout
  double* out = p.partial + ((size_t)cta * S_GROUPS + group) * (size_t)(FSB_GT * FSB_GT);
#pragma unroll
  for (int i = R0; i < R1; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      constexpr int base = R0 * (R0 + 1) / 2;
      const int row = i * 8 + (lane >> 2), col = j * 8 + 2 * (lane & 3);
      *reinterpret_cast<double2*>(out + row * FSB_GT + col) =
          make_double2(acc[i * (i + 1) / 2 + j - base][0], acc[i * (i + 1) / 2 + j - base][1]);
    }
}

template <int NB, bool VEC16>
__global__ void __launch_bounds__(S_THREADS, 1) gram_rowsplit_kernel(SmallArgs p) {
  extern __shared__ double smem[];   // [S_NSTAGE][ S_RCH x pitch | w: S_RCH ]
  constexpr int PITCH = 8 * NB + 4;
  constexpr int SPLIT = split_row(NB);
  const int k = p.k;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int group = warp & 3, part = warp >> 2;     // warps (g, g+4) form pair g: same scheduler
  const int64_t cta = blockIdx.x;
  const int64_t row_begin = cta * p.rows_per_cta;
  int64_t row_end = row_begin + p.rows_per_cta;
  if (row_end > p.n_rows) row_end = p.n_rows;
  const int nsteps = row_end > row_begin ? (int)((row_end - row_begin + S_RCH - 1) / S_RCH) : 0;
  const int64_t last_row = p.n_rows > 0 ? p.n_rows - 1 : 0;

  // ---- copy plan (see gram.cu): slot = one column (8 B) or a column pair (16 B) of one stage row
  constexpr int SLOTS_PER_ROW = VEC16 ? 64 : 128;
  constexpr int ROW_GROUPS = S_THREADS / SLOTS_PER_ROW;        // 4 (VEC16) or 2
  constexpr int COPIES = S_RCH / ROW_GROUPS;                   // 16 or 32 per thread per stage
  const int slot = tid % SLOTS_PER_ROW, srow = tid / SLOTS_PER_ROW;
  const int scol = VEC16 ? 2 * slot : slot;
  const bool col_live = scol < 8 * NB;                         // columns past the padded width: nothing to do
  struct Src { const double* ptr; int64_t stride; int nbytes; };
  auto scalar_src = [&](int gc) {
    Src r;
    if (gc < k) { r.ptr = p.A + gc; r.stride = p.lda; r.nbytes = 8; }
    else if (gc == k) { r.ptr = p.b; r.stride = 1; r.nbytes = 8; }
    else { r.ptr = p.w; r.stride = 0; r.nbytes = 0; }
    return r;
  };
  Src s0, s1;
  int mode;   // 0: one 16-byte copy, 1: pair straddling column k (two scalar slots), 2: scalar plan
  if (VEC16) {
    if (scol + 1 < k) { mode = 0; s0.ptr = p.A + scol; s0.stride = p.lda; s0.nbytes = 16; s1 = s0; }
    else if (scol > k) { mode = 0; s0.ptr = p.w; s0.stride = 0; s0.nbytes = 0; s1 = s0; }
    else { mode = 1; s0 = scalar_src(scol); s1 = scalar_src(scol + 1); }
  } else { mode = 2; s0 = scalar_src(scol); s1 = s0; }
  const double* run = s0.ptr + (row_begin + srow) * s0.stride;
  const int64_t step = (int64_t)ROW_GROUPS * s0.stride;
  const double* runW = p.w + row_begin + (tid < S_RCH ? tid : 0);
  constexpr int STAGE_DOUBLES = S_RCH * PITCH + S_RCH;

  auto issue_stage = [&](int stp) {
    if (stp < nsteps) {
      double* st = smem + (size_t)(stp % S_NSTAGE) * STAGE_DOUBLES;
      const int64_t r0 = row_begin + (int64_t)stp * S_RCH;
      const bool full = (r0 + S_RCH <= row_end);
      if (col_live) {
#pragma unroll 4
        for (int i = 0; i < COPIES; ++i) {
          const int lr = srow + ROW_GROUPS * i;
          double* d = st + lr * PITCH + scol;
          if (full && mode != 1) {
            if (mode == 0) cp_async16(d, run, s0.nbytes);
            else cp_async8(d, run, s0.nbytes != 0);
          } else {
            const int64_t r = r0 + lr;
            const bool in = r < row_end;
            const int64_t rc = in ? r : last_row;
            if (mode == 0) cp_async16(d, s0.ptr + rc * s0.stride, in ? s0.nbytes : 0);
            else {
              cp_async8(d, s0.ptr + rc * s0.stride, in && s0.nbytes);
              if (mode == 1) cp_async8(d + 1, s1.ptr + rc * s1.stride, in && s1.nbytes);
            }
          }
          run += step;
        }
      }
      if (tid < S_RCH) {
        const int64_t r = r0 + tid;
        const bool in = r < row_end;
        cp_async8(st + S_RCH * PITCH + tid, in ? runW : p.w + last_row, in);
      }
      runW += S_RCH;
    }
    cp_async_commit();
  };

#pragma unroll
  for (int s = 0; s < S_NSTAGE - 1; ++s) issue_stage(s);

  if (part == 0) run_part<0, SPLIT>(p, smem, PITCH, group, lane, nsteps, cta, issue_stage);
  else run_part<SPLIT, NB>(p, smem, PITCH, group, lane, nsteps, cta, issue_stage);
}

template <int NB>
int launch_nb(const SmallArgs& a, int ncta, bool vec16, cudaStream_t s) {
  constexpr int PITCH = 8 * NB + 4;
  const size_t smem = (size_t)S_NSTAGE * (S_RCH * PITCH + S_RCH) * sizeof(double);
  if (vec16) {
    FSB_CUDA_TRY(cudaFuncSetAttribute(gram_rowsplit_kernel<NB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gram_rowsplit_kernel<NB, true><<<ncta, S_THREADS, smem, s>>>(a);
  } else {
    FSB_CUDA_TRY(cudaFuncSetAttribute(gram_rowsplit_kernel<NB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gram_rowsplit_kernel<NB, false><<<ncta, S_THREADS, smem, s>>>(a);
  }
  FSB_LAUNCH_CHECK("gram_rowsplit_kernel");
  return FSB_OK;
}

}  // namespace

// number of row chunks (CTAs) the row-split kernel uses; each contributes S_GROUPS partials
int fsb_gram_small_ctas(const fsb_context* h, int64_t n_rows) {
  int64_t want = h->sm_count;
  int64_t maxc = fsb_ceil_div(n_rows > 0 ? n_rows : 1, 2 * S_RCH);
  if (want > maxc) want = maxc;
  return (int)(want < 1 ? 1 : want);
}

int fsb_gram_small_groups() { return S_GROUPS; }

int fsb_launch_gram_small(const fsb_context* h, const double* A, int64_t lda, const double* b, const double* weff,
                          int64_t n_rows, int k, double* partial, cudaStream_t s) {
  const int nb = (k + 1 + 7) / 8;
  SmallArgs a;
  a.A = A; a.lda = lda; a.b = b; a.w = weff; a.n_rows = n_rows; a.k = k; a.partial = partial;
  const int want = fsb_gram_small_ctas(h, n_rows);
  a.rows_per_cta = fsb_round_up(fsb_ceil_div(n_rows > 0 ? n_rows : 1, want), S_RCH);
  const int ncta = (int)fsb_ceil_div(n_rows > 0 ? n_rows : 1, a.rows_per_cta);
  // every (cta < want, group) slot of the workspace must be written: launch `want` CTAs, the
  // surplus ones see an empty row range and write zeros
  (void)ncta;
  const bool vec16 = (lda % 2 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  switch (nb) {
#define FSB_NB(N) case N: return launch_nb<N>(a, want, vec16, s);
    FSB_NB(1) FSB_NB(2) FSB_NB(3) FSB_NB(4) FSB_NB(5) FSB_NB(6) FSB_NB(7)
    FSB_NB(8) FSB_NB(9) FSB_NB(10) FSB_NB(11) FSB_NB(12) FSB_NB(13)
#undef FSB_NB
    default: return FSB_ERR_UNSUPPORTED;
  }
}
