// HBM-bound streaming kernels of the linear-fit path:
//   K1  scatter   raw LAMMPS descriptor blocks -> rows of A, b, w
//                 (fitsnap3lib/calculators/lammps_snap.py:391-556, lammps_pace.py:369-509)
//   K7  residual  g = aw^T (bw - aw x)         (refinement pass; cf. solvers/ridge.py:60)
//       predict   y = A x                      (solvers/solver.py:377)
// One warp owns one row: lanes stride across the columns, so every global access is a
// coalesced 256-byte request; several rows per warp iteration keep enough loads in flight to
// cover HBM latency for narrow rows.  No shared-memory staging: there is no reuse.
#include "fsb_common.cuh"
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

namespace {

constexpr double VIRIAL_UNIT = 1.6021765e6;  // lammps_snap.py:526

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__constant__ int c_prefetch = 0;   // L2 prefetch ahead of the demand loads: measured no gain on B200, off by default
__device__ __forceinline__ bool prefetch_enabled() { return c_prefetch != 0; }

// ------------------------------------------------------------------------------ K7
template <int NPL, int RPI, bool WITH_G, int MINB>
__global__ void __launch_bounds__(256, MINB) rowpass_kernel(const double* __restrict__ A, int64_t lda,
                                                      const double* __restrict__ b,
                                                      const double* __restrict__ w,
                                                      const uint8_t* __restrict__ testing, int64_t n_rows,
                                                      int k, const double* __restrict__ x,
                                                      double* __restrict__ out, int64_t rows_per_cta) {
  // WITH_G: out = per-CTA partial g [gridDim.x][k];  else: out = y [n_rows]
  extern __shared__ double sg[];  // k doubles (WITH_G only)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  double xr[NPL], gacc[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = lane + 32 * i;
    xr[i] = (c < k) ? x[c] : 0.0;
    gacc[i] = 0.0;
  }
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta;
  int64_t r_end = r_begin + rows_per_cta;
  if (r_end > n_rows) r_end = n_rows;

  const bool do_pf = prefetch_enabled();
  for (int64_t r0 = r_begin + (int64_t)warp * RPI; r0 < r_end; r0 += (int64_t)nwarp * RPI) {
    if (do_pf) {
      // pull the rows this warp will need two iterations from now into L2 (one line per lane)
      const int64_t rp = r0 + 2 * (int64_t)nwarp * RPI;
      if (rp < r_end) {
        const int64_t nr = (r_end - rp) < RPI ? (r_end - rp) : RPI;
        const char* base = reinterpret_cast<const char*>(A + rp * lda);
        const int64_t nbytes = ((nr - 1) * lda + k) * (int64_t)sizeof(double);
        for (int64_t off = (int64_t)lane * 128; off < nbytes; off += 32 * 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
      }
    }
    double a[RPI][NPL];
    double wv[RPI], bv[RPI];
    unsigned tb[RPI];
    // all loads are unconditional (row index clamped) so none waits on another; rows past the end
    // and test rows get weight 0 afterwards (their finite values then contribute exactly 0)
#pragma unroll
    for (int q = 0; q < RPI; ++q) {
      const int64_t r = r0 + q;
      const int64_t rc = r < r_end ? r : r_end - 1;
      wv[q] = 0.0; bv[q] = 0.0; tb[q] = 0u;
      if (WITH_G) {
        wv[q] = __ldg(w + rc);
        bv[q] = __ldg(b + rc);
        if (testing) tb[q] = (unsigned)__ldg(testing + rc);
      }
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        const int c = lane + 32 * i;
        a[q][i] = (c < k) ? __ldg(A + rc * lda + c) : 0.0;
      }
    }
    if (WITH_G) {
#pragma unroll
      for (int q = 0; q < RPI; ++q)
        if (r0 + q >= r_end || tb[q] != 0u) wv[q] = 0.0;
    }
#pragma unroll
    for (int q = 0; q < RPI; ++q) {
      if (WITH_G) {
        double dot = 0.0;
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
          a[q][i] *= wv[q];              // aw = w * a, rounded as the reference does (svd.py:44)
          dot += a[q][i] * xr[i];
        }
        dot = warp_sum(dot);
        const double res = wv[q] * bv[q] - dot;   // bw - aw x
#pragma unroll
        for (int i = 0; i < NPL; ++i) gacc[i] += a[q][i] * res;
      } else {
        double dot = 0.0;
#pragma unroll
        for (int i = 0; i < NPL; ++i) dot += a[q][i] * xr[i];
        dot = warp_sum(dot);
        if (lane == 0 && r0 + q < r_end) out[r0 + q] = dot;
      }
    }
  }

  if (WITH_G) {
    // fixed-order reduction over the CTA's warps, then one partial vector per CTA
    for (int c = threadIdx.x; c < k; c += blockDim.x) sg[c] = 0.0;
    __syncthreads();
    for (int wv_ = 0; wv_ < nwarp; ++wv_) {
      if (warp == wv_) {
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
          const int c = lane + 32 * i;
          if (c < k) sg[c] += gacc[i];
        }
      }
      __syncthreads();
    }
    for (int c = threadIdx.x; c < k; c += blockDim.x) out[(size_t)blockIdx.x * k + c] = sg[c];
  }
}

// ------------------------------------------------------------------------------ K7, bulk-copy staged
// The register-staged kernel above is latency-bound for narrow rows (k ~ 100: 25 % occupancy, 4.7 TB/s):
// its loads in flight are bounded by registers.  Here one producer thread streams tiles of BK_ROWS whole rows
// -- contiguous in memory when lda == k -- into a shared-memory ring with cp.async.bulk (TMA 1-D, SASS
// UBLKCP), completion on mbarriers; ~150-200 KB per SM are in flight at no register cost.  8 consumer warps
// take 8 rows each per tile (lanes across columns, LDS), b / w / test flags of the NEXT tile are prefetched
// into registers while the current one is processed.  Same arithmetic and the same fixed-order reduction as
// rowpass_kernel<.., true>.
constexpr int BK_ROWS = 64;
constexpr int BK_CONSUMERS = 256;
constexpr int BK_NW = BK_CONSUMERS / 32;          // consumer warps
constexpr int BK_RPW = BK_ROWS / BK_NW;            // rows of a tile per warp
constexpr int BK_THREADS = BK_CONSUMERS + 32;

__device__ __forceinline__ unsigned bk_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bk_mbar_wait(unsigned bar, unsigned parity) {
  unsigned done = 0;
  for (unsigned spin = 0; spin < (1u << 24); ++spin) {
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

// Consumer layout (round 2): EIGHT lanes per row, four rows per warp instruction.  Lane (rs = lane / 8, cl = lane % 8)
// owns the column pairs 2 cl + 16 i of row rs of the current group of four rows: one 16-byte shared-memory load per
// pair, the dot product needs three shuffle levels (8 lanes) instead of five, every lane slot carries a real column
// (100 columns on 7 x 16 slots instead of 4 x 32), and the four row groups keep separate column sums that are combined
// once at the end.  79 -> ~20 warp instructions per row: the consumers were the limit of this kernel (4.8 TB/s with
// 42 % issue utilisation and two warps per scheduler, profiles/r02_rowpass_c2.txt), not the bulk copies.
template <int NI>
__global__ void __launch_bounds__(BK_THREADS, 1) rowpass_bulk_kernel(const double* __restrict__ A,
                                                                     const double* __restrict__ b,
                                                                     const double* __restrict__ w,
                                                                     const uint8_t* __restrict__ testing,
                                                                     int64_t n_rows, int k,
                                                                     const double* __restrict__ x,
                                                                     double* __restrict__ out, int64_t rows_per_cta,
                                                                     int nstage) {
  extern __shared__ __align__(128) unsigned char bk_raw[];
  __shared__ __align__(8) unsigned long long s_full[8];
  __shared__ __align__(8) unsigned long long s_empty[8];
  double* ring = reinterpret_cast<double*>(bk_raw);              // nstage x (BK_ROWS x k)
  const size_t stage_doubles = (size_t)BK_ROWS * k;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta;
  int64_t r_end = r_begin + rows_per_cta;
  if (r_end > n_rows) r_end = n_rows;
  const int ntile = r_end > r_begin ? (int)((r_end - r_begin + BK_ROWS - 1) / BK_ROWS) : 0;

  if (tid == 0) {
    for (int i = 0; i < nstage; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bk_smem_u32(&s_full[i])), "r"(1) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bk_smem_u32(&s_empty[i])), "r"(BK_NW) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == BK_CONSUMERS / 32) {
    // ---- producer
    if (lane == 0) {
      for (int t = 0; t < ntile; ++t) {
        const int slot = t % nstage, n = t / nstage;
        if (t >= nstage) bk_mbar_wait(bk_smem_u32(&s_empty[slot]), (unsigned)((n - 1) & 1));
        const int64_t r0 = r_begin + (int64_t)t * BK_ROWS;
        const int nr = (int)((r_end - r0) < BK_ROWS ? (r_end - r0) : BK_ROWS);
        const unsigned bytes = (unsigned)((size_t)nr * k * sizeof(double));
        const unsigned full = bk_smem_u32(&s_full[slot]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(bk_smem_u32(ring + slot * stage_doubles)), "l"(A + r0 * k), "r"(bytes), "r"(full)
                     : "memory");
      }
    }
    return;
  }

  // ---- consumers
  const int rs = lane >> 3, cl = lane & 7;
  double x0[NI], x1[NI], g0[NI], g1[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int c = 2 * cl + 16 * i;                   // k is even: a pair is inside or outside as a whole
    x0[i] = (c < k) ? x[c] : 0.0;
    x1[i] = (c < k) ? x[c + 1] : 0.0;
    g0[i] = 0.0; g1[i] = 0.0;
  }
  // lane q < BK_RPW keeps (w, b, test flag) of row BK_RPW * warp + q of the tile, loaded TWO tiles ahead and not
  // touched until that tile is current: the first version selected w = 0 for test rows right behind the loads, which
  // made every fetch wait for its own DRAM round trip (33 % of the kernel's stall samples on that one FSEL,
  // profiles/r02b_rowpass_c2.txt).  Rows past the end keep w = 0.
  auto fetch = [&](int t, double& wv, double& bv, unsigned& tv) {
    wv = 0.0; bv = 0.0; tv = 0u;
    const int64_t r = r_begin + (int64_t)t * BK_ROWS + warp * BK_RPW + (lane & (BK_RPW - 1));
    if (t < ntile && r < r_end) {
      bv = __ldg(b + r);
      wv = __ldg(w + r);
      if (testing) tv = (unsigned)__ldg(testing + r);
    }
  };
  double w_cur, b_cur, w_nxt, b_nxt, w_nx2, b_nx2;
  unsigned t_cur, t_nxt, t_nx2;
  fetch(0, w_cur, b_cur, t_cur);
  fetch(1, w_nxt, b_nxt, t_nxt);
  const unsigned kbytes = (unsigned)k * 8u;
  for (int t = 0; t < ntile; ++t) {
    fetch(t + 2, w_nx2, b_nx2, t_nx2);
    const int slot = t % nstage, n = t / nstage;
    bk_mbar_wait(bk_smem_u32(&s_full[slot]), (unsigned)(n & 1));
    const int64_t r0 = r_begin + (int64_t)t * BK_ROWS;
    const int nr = (int)((r_end - r0) < BK_ROWS ? (r_end - r0) : BK_ROWS);     // rows of this tile that exist
    const unsigned tile = bk_smem_u32(ring + slot * stage_doubles);
    const double w_use = t_cur ? 0.0 : w_cur;        // test rows: weight 0 (selected now, two tiles after the loads)
#pragma unroll
    for (int it = 0; it < BK_RPW / 4; ++it) {
      const int q = it * 4 + rs;                      // row of this warp's share of the tile
      int row = warp * BK_RPW + q;
      const double wv = __shfl_sync(0xffffffffu, w_use, q);     // 0 for a test row and for rows past the end
      const double bv = __shfl_sync(0xffffffffu, b_cur, q);
      row = row < nr ? row : nr - 1;                  // ragged last tile: an existing row, weight 0
      const unsigned raddr = tile + (unsigned)row * kbytes + (unsigned)cl * 16u;
      double a0[NI], a1[NI];
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        double v0 = 0.0, v1 = 0.0;
        if (2 * cl + 16 * i < k)
          asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(raddr + (unsigned)i * 128u));
        a0[i] = v0 * wv;                              // aw = w * a, rounded as the reference does (svd.py:44)
        a1[i] = v1 * wv;
      }
      double d0 = 0.0, d1 = 0.0;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        d0 += a0[i] * x0[i];
        d1 += a1[i] * x1[i];
      }
      double dot = d0 + d1;
      dot += __shfl_xor_sync(0xffffffffu, dot, 1);
      dot += __shfl_xor_sync(0xffffffffu, dot, 2);
      dot += __shfl_xor_sync(0xffffffffu, dot, 4);
      const double res = wv * bv - dot;               // bw - aw x
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        g0[i] += a0[i] * res;
        g1[i] += a1[i] * res;
      }
    }
    __syncwarp();
    if (lane == 0)
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bk_smem_u32(&s_empty[slot])) : "memory");
    w_cur = w_nxt; b_cur = b_nxt; t_cur = t_nxt;
    w_nxt = w_nx2; b_nxt = b_nx2; t_nxt = t_nx2;
  }

  // the four row groups of the warp, in a fixed order; then the consumer warps, in a fixed order (the ring is free
  // now: all tiles consumed by this warp, other warps may still read theirs -> a separate region past the ring)
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    g0[i] += __shfl_xor_sync(0xffffffffu, g0[i], 8);
    g1[i] += __shfl_xor_sync(0xffffffffu, g1[i], 8);
    g0[i] += __shfl_xor_sync(0xffffffffu, g0[i], 16);
    g1[i] += __shfl_xor_sync(0xffffffffu, g1[i], 16);
  }
  double* sg = ring + (size_t)nstage * stage_doubles;      // BK_NW x k doubles
  if (rs == 0) {
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int c = 2 * cl + 16 * i;
      if (c < k) { sg[warp * k + c] = g0[i]; sg[warp * k + c + 1] = g1[i]; }
    }
  }
  asm volatile("bar.sync 1, %0;" ::"n"(BK_CONSUMERS) : "memory");
  for (int c = tid; c < k; c += BK_CONSUMERS) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < BK_NW; ++q) t += sg[q * k + c];
    out[(size_t)blockIdx.x * k + c] = t;
  }
}

// Deterministic column sums of the per-CTA partial vectors: 32 columns x 32 part-groups per block,
// each group adds its parts in index order, the 32 group sums are combined in a fixed order.
__global__ void __launch_bounds__(1024) colsum_reduce_kernel(const double* __restrict__ partial, int nparts, int k,
                                                             double* __restrict__ g) {
  __shared__ double sh[32][33];
  const int cx = threadIdx.x & 31, gy = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  double s = 0.0;
  if (c < k)
    for (int p = gy; p < nparts; p += 32) s += partial[(size_t)p * k + c];
  sh[gy][cx] = s;
  __syncthreads();
  if (gy == 0 && c < k) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < 32; ++q) t += sh[q][cx];
    g[c] = t;
  }
}

// ------------------------------------------------------------------------------ error analysis
// Per-(group, train/test, row type) error sums of the linear error analysis
// (fitsnap3lib/solvers/solver.py:108-133, 368-429) in ONE pass over A: pred = a . x, res = b - pred.
// stats[g][0..9] = n, sum|res|, sum res^2, sum t, sum t^2, n(w != 0), sum|w res|, sum (w res)^2, sum w t, sum (w t)^2.
// Each warp walks a CONTIGUOUS row range and keeps the sums of the current group in registers (rows
// of one configuration / row type are adjacent), flushing to a shared-memory table only when the
// group id changes; the table goes to global memory with one atomicAdd per non-zero entry per CTA.
// (Atomic accumulation order is not fixed: metrics agree to rounding, not bit for bit.)
constexpr int GS_NSTAT = 10;

template <int NPL>
__global__ void __launch_bounds__(256) group_stats_kernel(const double* __restrict__ A, int64_t lda,
                                                          const double* __restrict__ b,
                                                          const double* __restrict__ w,
                                                          const int32_t* __restrict__ gid, int64_t n_rows, int k,
                                                          const double* __restrict__ x, int n_groups,
                                                          double* __restrict__ stats, int64_t rows_per_warp) {
  extern __shared__ double tab[];   // n_groups x GS_NSTAT
  for (int i = threadIdx.x; i < n_groups * GS_NSTAT; i += blockDim.x) tab[i] = 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  double xr[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = lane + 32 * i;
    xr[i] = (c < k) ? x[c] : 0.0;
  }
  const int64_t r_begin = warp_global * rows_per_warp;
  int64_t r_end = r_begin + rows_per_warp;
  if (r_end > n_rows) r_end = n_rows;
  double acc[GS_NSTAT];
#pragma unroll
  for (int q = 0; q < GS_NSTAT; ++q) acc[q] = 0.0;
  int cur = -1;
  auto flush = [&]() {
    if (cur >= 0 && lane == 0) {
#pragma unroll
      for (int q = 0; q < GS_NSTAT; ++q)
        if (acc[q] != 0.0) atomicAdd(&tab[cur * GS_NSTAT + q], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < GS_NSTAT; ++q) acc[q] = 0.0;
  };
  constexpr int RPI = (NPL <= 4) ? 4 : (NPL <= 8 ? 2 : 1);
  for (int64_t r0 = r_begin; r0 < r_end; r0 += RPI) {
    double a[RPI][NPL];
#pragma unroll
    for (int q = 0; q < RPI; ++q) {
      const int64_t rc = (r0 + q < r_end) ? r0 + q : r_end - 1;
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        const int c = lane + 32 * i;
        a[q][i] = (c < k) ? __ldg(A + rc * lda + c) : 0.0;
      }
    }
#pragma unroll
    for (int q = 0; q < RPI; ++q) {
      if (r0 + q >= r_end) break;
      double dot = 0.0;
#pragma unroll
      for (int i = 0; i < NPL; ++i) dot += a[q][i] * xr[i];
      dot = warp_sum(dot);
      const int64_t r = r0 + q;
      const int g = gid[r];
      if (g != cur) { flush(); cur = g; }
      const double t = b[r], wv = w[r];
      const double res = t - dot, wres = wv * res, wt = wv * t;
      acc[0] += 1.0; acc[1] += fabs(res); acc[2] += res * res; acc[3] += t; acc[4] += t * t;
      acc[5] += (wv != 0.0) ? 1.0 : 0.0; acc[6] += fabs(wres); acc[7] += wres * wres; acc[8] += wt; acc[9] += wt * wt;
    }
  }
  flush();
  __syncthreads();
  for (int i = threadIdx.x; i < n_groups * GS_NSTAT; i += blockDim.x)
    if (tab[i] != 0.0) atomicAdd(&stats[i], tab[i]);
}

struct RowPlan {
  int ncta;
  int64_t rows_per_cta;
};

RowPlan plan_rows(const fsb_context* h, int64_t n_rows) {
  RowPlan pl;
  int64_t want = (int64_t)h->sm_count * 8;
  int64_t maxc = fsb_ceil_div(n_rows > 0 ? n_rows : 1, 64);
  if (want > maxc) want = maxc;
  if (want < 1) want = 1;
  pl.rows_per_cta = fsb_ceil_div(n_rows > 0 ? n_rows : 1, want);
  pl.ncta = (int)fsb_ceil_div(n_rows > 0 ? n_rows : 1, pl.rows_per_cta);
  return pl;
}

// bulk-copy staged residual pass: whole rows must be contiguous (lda == k) and 16-byte tileable
bool rowpass_bulk_ok(const fsb_context* h, const double* A, int64_t lda, int64_t n_rows, int k) {
  static int off = -1;
  if (off < 0) off = getenv("FSB_ROWPASS_NO_BULK") ? 1 : 0;
  return !off && lda == k && (k & 1) == 0 && k <= 128 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 &&
         n_rows >= (int64_t)h->sm_count * BK_ROWS * 4;
}

int launch_rowpass_bulk(const fsb_context* h, const double* A, const double* b, const double* w,
                        const uint8_t* testing, int64_t n_rows, int k, const double* x, double* out, int* nparts,
                        cudaStream_t s) {
  const size_t stage = (size_t)BK_ROWS * k * sizeof(double);
  const size_t tail = (size_t)BK_NW * k * sizeof(double);
  int nstage = (int)((h->smem_optin - 2048 - tail) / stage);
  if (nstage > 6) nstage = 6;
  if (nstage < 2) return FSB_ERR_UNSUPPORTED;
  const size_t smem = (size_t)nstage * stage + tail;
  const int grid = h->sm_count;
  const int64_t rows_per_cta = fsb_round_up(fsb_ceil_div(n_rows, grid), BK_ROWS);
  const int ni = (k + 15) / 16;
#define FSB_BULK(NI)                                                                                           \
  do {                                                                                                         \
    FSB_CUDA_TRY(cudaFuncSetAttribute(rowpass_bulk_kernel<NI>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                      (int)smem));                                                             \
    rowpass_bulk_kernel<NI><<<grid, BK_THREADS, smem, s>>>(A, b, w, testing, n_rows, k, x, out, rows_per_cta,  \
                                                           nstage);                                            \
  } while (0)
  switch (ni) {
    case 1: FSB_BULK(1); break;
    case 2: FSB_BULK(2); break;
    case 3: FSB_BULK(3); break;
    case 4: FSB_BULK(4); break;
    case 5: FSB_BULK(5); break;
    case 6: FSB_BULK(6); break;
    case 7: FSB_BULK(7); break;
    case 8: FSB_BULK(8); break;
    default: return FSB_ERR_UNSUPPORTED;   // rowpass_bulk_ok admits k <= 128 only
  }
#undef FSB_BULK
  FSB_LAUNCH_CHECK("rowpass_bulk_kernel");
  *nparts = grid;
  return FSB_OK;
}

template <bool WITH_G>
int launch_rowpass(const fsb_context* h, const double* A, int64_t lda, const double* b, const double* w,
                   const uint8_t* testing, int64_t n_rows, int k, const double* x, double* out,
                   cudaStream_t s) {
  RowPlan pl = plan_rows(h, n_rows);
  const size_t smem = WITH_G ? (size_t)k * sizeof(double) : 0;
  const int npl = (k + 31) / 32;
  static int pf_set = 0;
  if (!pf_set) {
    const char* e = getenv("FSB_PREFETCH");
    const int v = e ? 1 : 0;
    cudaMemcpyToSymbol(c_prefetch, &v, sizeof(int));
    pf_set = 1;
  }
  // tuning knob (read once): FSB_ROWPASS_MINB = 1|2|3 resident-CTA target of the narrow-row variants
  static int minb = -1;
  if (minb < 0) {
    const char* e = getenv("FSB_ROWPASS_MINB");
    minb = e ? atoi(e) : 2;
    if (minb < 1 || minb > 3) minb = 2;
  }
#define FSB_ROWPASS_M(NPL, RPI, MINB)                                                                        \
  rowpass_kernel<NPL, RPI, WITH_G, MINB><<<pl.ncta, 256, smem, s>>>(A, lda, b, w, testing, n_rows, k, x, out, \
                                                                    pl.rows_per_cta)
#define FSB_ROWPASS(NPL, RPI)                                   \
  do {                                                          \
    if (minb == 3) FSB_ROWPASS_M(NPL, RPI, 3);                  \
    else if (minb == 2) FSB_ROWPASS_M(NPL, RPI, 2);             \
    else FSB_ROWPASS_M(NPL, RPI, 1);                            \
  } while (0)
#define FSB_ROWPASS1(NPL, RPI) FSB_ROWPASS_M(NPL, RPI, 1)
  if (npl <= 1) FSB_ROWPASS(1, 8);
  else if (npl <= 2) FSB_ROWPASS(2, 4);
  else if (npl <= 4) { if (getenv("FSB_ROWPASS_RPI8")) FSB_ROWPASS_M(4, 8, 1); else FSB_ROWPASS(4, 4); }
  else if (npl <= 8) FSB_ROWPASS(8, 2);
  else if (npl <= 16) FSB_ROWPASS1(16, 2);   // two rows per warp iteration: 8 KB in flight per warp, as at npl = 32
  else if (npl <= 32) FSB_ROWPASS1(32, 1);
  else if (npl <= 64) FSB_ROWPASS1(64, 1);
  else return FSB_ERR_UNSUPPORTED;   // k > 2048
#undef FSB_ROWPASS1
#undef FSB_ROWPASS_M
#undef FSB_ROWPASS
  FSB_LAUNCH_CHECK("rowpass_kernel");
  return FSB_OK;
}

// ------------------------------------------------------------------------------ K1
using fsb_dev::scrub;

// A CTA walks tiles of SC_TILE consecutive output rows.  Phase 1: one thread per row resolves the
// row's metadata (configuration, row family, source row, divisor) with independent global loads and
// writes b and w (coalesced); the results go to shared memory.  Phase 2: each warp streams rows,
// SC_RPW rows x SC_CU column chunks = 16 independent 8-byte loads per lane in flight, with no global
// dependency chain in front of them.  The row-independent column map (raw column -> output column)
// and the blank2J prefactors also live in shared memory: per element the work is two LDS, one LDG,
// at most one divide, one multiply, one STG.
constexpr int SC_TILE = 128;
constexpr int SC_RPW = 4;
constexpr int SC_CU = 4;

__global__ void __launch_bounds__(256, 3) scatter_kernel(ScatterArgs p, int64_t total) {
  extern __shared__ unsigned char sc_smem[];
  __shared__ int64_t m_src[SC_TILE];     // raw row index
  __shared__ double m_div[SC_TILE];      // N (energy rows) or V (virial rows)
  __shared__ int m_cfg[SC_TILE];
  __shared__ int m_kind[SC_TILE];        // 0 energy, 1 force, 2 virial
  const bool use_e = p.flags & FSB_ROWS_ENERGY, use_f = p.flags & FSB_ROWS_FORCE,
             bzero = p.flags & FSB_BZEROFLAG;
  const bool do_scrub = p.flags & FSB_SCRUB_NONFINITE;
  const int kraw = p.ncoeff * p.numtypes;
  const int k = bzero ? kraw : kraw + p.numtypes;
  const int seg = p.ncoeff + 1;
  const int64_t ldr = kraw + 1;
  double* s_b2j = reinterpret_cast<double*>(sc_smem);
  int* s_src = reinterpret_cast<int*>(s_b2j + k);   // >= 0: raw column; < 0: lead column of type (-v-1)
  for (int oc = threadIdx.x; oc < k; oc += blockDim.x) {
    s_b2j[oc] = p.blank2j[oc];
    int v = oc;
    if (!bzero) {
      const int t = oc / seg, q = oc - t * seg;
      v = (q == 0) ? -(t + 1) : t * p.ncoeff + q - 1;
    }
    s_src[oc] = v;
  }

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int64_t row0 = p.out_row_off[0];
  const int64_t ntiles = (total + SC_TILE - 1) / SC_TILE;
  bool bad = false;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();   // previous tile fully consumed (and the column map is visible)
    const int64_t t0 = tile * SC_TILE;
    const int nrow = (int)((total - t0) < SC_TILE ? (total - t0) : SC_TILE);
    if (threadIdx.x < nrow) {
      const int64_t rr = t0 + threadIdx.x;
      const int64_t row = row0 + rr;
      int c;
      if (p.row_cfg) {
        c = p.row_cfg[rr];
      } else {  // binary search: largest c with out_row_off[c] <= row
        int lo = 0, hi = p.ncfg - 1;
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (p.out_row_off[mid] <= row) lo = mid; else hi = mid - 1;
        }
        c = lo;
      }
      int64_t local = row - p.out_row_off[c];
      const int n = p.natoms[c];
      int kind;
      int64_t sub;
      if (use_e && local == 0) { kind = 0; sub = 0; }
      else {
        if (use_e) local -= 1;
        if (use_f && local < 3 * (int64_t)n) { kind = 1; sub = local; }
        else { if (use_f) local -= 3 * (int64_t)n; kind = 2; sub = local; }
      }
      const int64_t rraw = p.raw_row_off[c] + (kind == 0 ? 0 : (kind == 1 ? 1 + sub : 1 + 3 * (int64_t)n + sub));
      m_src[threadIdx.x] = rraw;
      m_cfg[threadIdx.x] = c;
      m_kind[threadIdx.x] = kind;
      m_div[threadIdx.x] = (kind == 0) ? (double)n : (kind == 2 ? p.volume[c] : 1.0);
      // b and w of this row
      const double ref = scrub(__ldg(p.raw + rraw * ldr + kraw), do_scrub, bad);
      double bv, wv;
      if (kind == 0) {
        bv = (p.energy[c] - ref) / (double)n;             // lammps_snap.py:473
        wv = p.eweight[c];
      } else if (kind == 1) {
        // atoms before configuration c = (raw rows before it - 7 c) / 3: an exact multiple of 3, so the
        // quotient is one multiply by the inverse of 3 modulo 2^64 (a 64-bit division costs ~100 instructions)
        const int64_t atom0 = (int64_t)((uint64_t)(p.raw_row_off[c] - p.raw_row_off[0] - 7 * (int64_t)c) *
                                        0xAAAAAAAAAAAAAAABull);
        bv = p.forces[3 * atom0 + sub] - ref;             // :506-507
        wv = p.fweight[c];
      } else {
        const int vi[6] = {0, 1, 2, 1, 0, 0}, vj[6] = {0, 1, 2, 2, 2, 1};
        bv = p.stress[(size_t)c * 9 + vi[sub] * 3 + vj[sub]] - ref;   // :540-541
        wv = p.vweight[c];
      }
      p.b[row] = bv;
      p.w[row] = wv;
    }
    __syncthreads();

    for (int lr0 = warp * SC_RPW; lr0 < nrow; lr0 += nwarp * SC_RPW) {
      const double* src[SC_RPW];
      double* dst[SC_RPW];
      double dv[SC_RPW];
      int kind[SC_RPW];
#pragma unroll
      for (int q = 0; q < SC_RPW; ++q) {
        const int lr = (lr0 + q < nrow) ? lr0 + q : lr0;   // duplicates the first row past the end (not stored)
        src[q] = p.raw + m_src[lr] * ldr;
        dst[q] = p.A + (row0 + t0 + lr) * p.lda;
        dv[q] = m_div[lr];
        kind[q] = (lr0 + q < nrow) ? m_kind[lr] : -1;
      }
      for (int oc0 = 0; oc0 < k; oc0 += 32 * SC_CU) {
        double v[SC_RPW][SC_CU], pref[SC_CU];
        int srcc[SC_CU];
#pragma unroll
        for (int u = 0; u < SC_CU; ++u) {
          const int oc = oc0 + lane + 32 * u;
          const bool live = oc < k;
          srcc[u] = live ? s_src[oc] : -1;
          pref[u] = live ? s_b2j[oc] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < SC_RPW; ++q)
#pragma unroll
          for (int u = 0; u < SC_CU; ++u) v[q][u] = (srcc[u] >= 0) ? __ldg(src[q] + srcc[u]) : 0.0;
        // non-finite detection on the raw bits (exponent all ones), scrubbed only when asked to
        unsigned nf = 0;
#pragma unroll
        for (int q = 0; q < SC_RPW; ++q)
#pragma unroll
          for (int u = 0; u < SC_CU; ++u)
            nf |= ((unsigned)__double2hiint(v[q][u]) & 0x7ff00000u) == 0x7ff00000u ? 1u : 0u;
        if (nf) {
          bad = true;
          if (do_scrub) {
#pragma unroll
            for (int q = 0; q < SC_RPW; ++q)
#pragma unroll
              for (int u = 0; u < SC_CU; ++u) v[q][u] = scrub(v[q][u], true, bad);
          }
        }
#pragma unroll
        for (int q = 0; q < SC_RPW; ++q) {
          double* d = dst[q] + oc0 + lane;
          if (kind[q] == 1) {                                   // force rows: A = R * blank2J   (:493-502)
#pragma unroll
            for (int u = 0; u < SC_CU; ++u)
              if (oc0 + lane + 32 * u < k) d[32 * u] = v[q][u] * pref[u];   // lead columns hold v = 0
          } else if (kind[q] == 2) {                            // virial rows: (1.6021765e6*R)/V * blank2J (:526-536)
#pragma unroll
            for (int u = 0; u < SC_CU; ++u)
              if (oc0 + lane + 32 * u < k)
                d[32 * u] = ((srcc[u] >= 0) ? (VIRIAL_UNIT * v[q][u]) / dv[q] : 0.0) * pref[u];
          } else if (kind[q] == 0) {                            // energy row: R/N (+ type fractions) * blank2J (:435-467)
#pragma unroll
            for (int u = 0; u < SC_CU; ++u)
              if (oc0 + lane + 32 * u < k) {
                const double val = (srcc[u] >= 0)
                                       ? v[q][u] / dv[q]
                                       : p.type_fraction[(size_t)m_cfg[lr0 + q] * p.numtypes + (-srcc[u] - 1)];
                d[32 * u] = val * pref[u];
              }
          }
        }
      }
    }
  }
  if (p.nonfinite && __any_sync(0xffffffffu, bad) && lane == 0) atomicAdd(p.nonfinite, 1);
}


// ------------------------------------------------------------------------------ K1, bulk-copy staged
// Fast path of the scatter for the common layout: energy + force + virial rows all assembled (output row i
// comes from raw row i), narrow rows (k <= 160), contiguous A (lda == k), row -> configuration map given.
// scatter_kernel above issues ~56 instructions per matrix element (address arithmetic, predicates, per-kind
// branches around global loads and stores) and is half issue-bound, half latency-bound (4.8 TB/s).  Here the
// global traffic is moved by the TMA unit: one producer thread streams tiles of SB_ROWS raw rows (contiguous)
// into a shared-memory ring (cp.async.bulk, mbarrier complete_tx), 8 consumer warps transform them from shared
// memory into a shared-memory output tile (same operations in the same order as scatter_kernel: bit-identical
// A, b, w), and one thread writes the tile back with cp.async.bulk.global.shared::cta (bulk_group).  Row
// metadata (configuration, row family, divisor, truth, weight) of the NEXT tile is prefetched into registers.
// Tiles that cannot be bulk-copied (the ragged last tile, 16-byte misalignment) are moved with plain loads
// and stores by the consumers; the arithmetic is shared.
constexpr int SB_ROWS = 32;
constexpr int SB_CONSUMERS = 256;
constexpr int SB_META = 6;                              // metadata warps (one per ring slot, <= nstage active)
constexpr int SB_THREADS = SB_CONSUMERS + 32 + 32 * SB_META;
constexpr int SB_RPW = SB_ROWS / (SB_CONSUMERS / 32);   // 4 rows of a tile per consumer warp
constexpr int SB_MAXK = 160;
constexpr int SB_NCH = SB_MAXK / 32;                    // column chunks of 32 held in registers per lane

struct SbRowMeta {      // one per row of a tile, written by the producer warp
  double div, wv, truth;
  int cfg, kind;        // kind: 0 energy, 1 force, 2 virial, -1 past the end
};

__global__ void __launch_bounds__(SB_THREADS, 1) scatter_bulk_kernel(ScatterArgs p, int64_t total,
                                                                     int64_t tiles_per_cta, int nstage) {
  extern __shared__ __align__(128) unsigned char sb_raw[];
  __shared__ __align__(8) unsigned long long s_full[8];
  __shared__ __align__(8) unsigned long long s_empty[8];
  __shared__ SbRowMeta s_meta[8][SB_ROWS];
  const bool bzero = p.flags & FSB_BZEROFLAG;
  const bool do_scrub = p.flags & FSB_SCRUB_NONFINITE;
  const int kraw = p.ncoeff * p.numtypes;
  const int k = bzero ? kraw : kraw + p.numtypes;
  const int seg = p.ncoeff + 1;
  const int ldr = kraw + 1;
  const size_t raw_doubles = (size_t)SB_ROWS * ldr;
  double* ring = reinterpret_cast<double*>(sb_raw);                 // nstage raw tiles
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  const int64_t ntiles = (total + SB_ROWS - 1) / SB_ROWS;
  const int64_t t_begin = (int64_t)blockIdx.x * tiles_per_cta;
  int64_t t_end = t_begin + tiles_per_cta;
  if (t_end > ntiles) t_end = ntiles;
  const int64_t row0 = p.out_row_off[0];
  const int64_t rraw0 = p.raw_row_off[0];
  // a tile goes through the TMA unit when it is full and its source is 16-byte aligned
  auto tile_is_bulk = [&](int64_t t) {
    const int64_t i0 = t * SB_ROWS;
    const bool full = (total - i0) >= SB_ROWS;
    const uintptr_t src = reinterpret_cast<uintptr_t>(p.raw + (rraw0 + i0) * ldr);
    return full && (src & 15) == 0;
  };

  if (tid == 0) {
    for (int i = 0; i < nstage; ++i) {
      // full: the expect_tx arrival of the bulk copy (or a plain arrival) + the arrival after the metadata
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bk_smem_u32(&s_full[i])), "r"(2) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bk_smem_u32(&s_empty[i])), "r"(SB_CONSUMERS / 32)
                   : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == SB_CONSUMERS / 32) {
    // ---- copy producer: one thread streams the raw tiles (non-bulk tiles are filled by the consumers; the slot /
    //      phase bookkeeping stays in step)
    if (lane == 0) {
      for (int64_t t = t_begin; t < t_end; ++t) {
        const int it = (int)(t - t_begin);
        const int slot = it % nstage, n = it / nstage;
        if (it >= nstage) bk_mbar_wait(bk_smem_u32(&s_empty[slot]), (unsigned)((n - 1) & 1));
        const unsigned full = bk_smem_u32(&s_full[slot]);
        if (tile_is_bulk(t)) {
          const unsigned bytes = (unsigned)(raw_doubles * sizeof(double));
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(bytes) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(bk_smem_u32(ring + slot * raw_doubles)), "l"(p.raw + (rraw0 + t * SB_ROWS) * ldr),
                         "r"(bytes), "r"(full) : "memory");
        } else {
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full) : "memory");   // slot handed over empty
        }
      }
    }
    return;
  }
  if (warp > SB_CONSUMERS / 32) {
    // ---- metadata producers: one warp per ring slot, lane r resolves row r of the tile.  Resolving a
    //      row is a chain of dependent global loads (row -> configuration -> offsets -> truth, ~2 us): done here,
    //      several tiles in flight, it never stalls the consumers.  Published in shared memory with the second
    //      arrival on the tile's barrier.
    const int mw = warp - SB_CONSUMERS / 32 - 1;
    const int64_t n_force = total - 7 * (int64_t)p.ncfg > 1 ? total - 7 * (int64_t)p.ncfg : 1;   // 3 * atoms
    // warp m owns ring slot m: the tiles of one slot are then resolved strictly in order, which the parity-based
    // mbarrier wait needs (a waiter two phases ahead of the barrier would see a matching parity and run early)
    for (int64_t t = t_begin + mw; mw < nstage && t < t_end; t += nstage) {
      const int it = (int)(t - t_begin);
      const int slot = it % nstage, n = it / nstage;
      if (it >= nstage) bk_mbar_wait(bk_smem_u32(&s_empty[slot]), (unsigned)((n - 1) & 1));
      SbRowMeta m;
      m.cfg = 0; m.kind = -1; m.div = 1.0; m.wv = 0.0; m.truth = 0.0;
      const int64_t i = t * SB_ROWS + lane;
      if (i < total) {
        // two levels of loads instead of four: everything that depends only on the configuration index is
        // requested at once (rows map 1:1, so the force component of row i is forces[i - 7 c - 1] whatever
        // the row family turns out to be; the index is clamped for the rows that will not use it)
        const int c = __ldg(p.row_cfg + i);
        int64_t fi = i - 7 * (int64_t)c - 1;
        fi = fi < 0 ? 0 : (fi > n_force - 1 ? n_force - 1 : fi);
        const int nat = __ldg(p.natoms + c);
        const int64_t off_c = __ldg(p.out_row_off + c);
        const double ew = __ldg(p.eweight + c), fw = __ldg(p.fweight + c), vw = __ldg(p.vweight + c);
        const double en = __ldg(p.energy + c), vol = __ldg(p.volume + c), fo = __ldg(p.forces + fi);
        const int64_t local = row0 + i - off_c;
        m.cfg = c;
        if (local == 0) {
          m.kind = 0; m.div = (double)nat; m.wv = ew; m.truth = en;
        } else if (local < 1 + 3 * (int64_t)nat) {
          m.kind = 1; m.wv = fw; m.truth = fo;
        } else {
          const int sub = (int)(local - 1 - 3 * (int64_t)nat);
          const int vi[6] = {0, 1, 2, 1, 0, 0}, vj[6] = {0, 1, 2, 2, 2, 1};
          m.kind = 2; m.div = vol; m.wv = vw;
          m.truth = __ldg(p.stress + (size_t)c * 9 + vi[sub] * 3 + vj[sub]);
        }
      }
      s_meta[slot][lane] = m;
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bk_smem_u32(&s_full[slot])) : "memory");
    }
    return;
  }

  // ---- consumers: the column map of this lane is loop-invariant, kept in registers
  bool bad = false;
  int srcc[SB_NCH];       // >= 0: raw column; < 0: lead column of type (-v-1); INT_MIN: column past k
  double pref[SB_NCH];
#pragma unroll
  for (int u = 0; u < SB_NCH; ++u) {
    const int oc = lane + 32 * u;
    srcc[u] = INT_MIN;
    pref[u] = 0.0;
    if (oc < k) {
      int v = oc;
      if (!bzero) {
        const int t = oc / seg, q = oc - t * seg;
        v = (q == 0) ? -(t + 1) : t * p.ncoeff + q - 1;
      }
      srcc[u] = v;
      pref[u] = __ldg(p.blank2j + oc);
    }
  }
  for (int64_t t = t_begin; t < t_end; ++t) {
    const int it = (int)(t - t_begin);
    const int slot = it % nstage, n = it / nstage;
    const int64_t i0 = t * SB_ROWS;
    const int nr = (int)((total - i0) < SB_ROWS ? (total - i0) : SB_ROWS);
    const bool bulk = tile_is_bulk(t);
    double* rin = ring + slot * raw_doubles;
    bk_mbar_wait(bk_smem_u32(&s_full[slot]), (unsigned)(n & 1));
    if (!bulk) {   // plain loads into the slot (ragged last tile / misaligned views)
      const double* src = p.raw + (rraw0 + i0) * ldr;
      for (int e = tid; e < nr * ldr; e += SB_CONSUMERS) rin[e] = __ldg(src + e);
    }
    if (!bulk) asm volatile("bar.sync 1, %0;" ::"n"(SB_CONSUMERS) : "memory");   // block-uniform

#pragma unroll
    for (int q = 0; q < SB_RPW; ++q) {
      const int lr = warp * SB_RPW + q;
      const SbRowMeta m = s_meta[slot][lr];                     // broadcast read
      if (m.kind < 0) continue;                                 // past the end (warp-uniform)
      const double* rrow = rin + (size_t)lr * ldr;
      double* orow = p.A + (row0 + i0 + lr) * (int64_t)k;     // written straight from registers (coalesced)
      double v[SB_NCH];
#pragma unroll
      for (int u = 0; u < SB_NCH; ++u) v[u] = (srcc[u] >= 0) ? rrow[srcc[u]] : 0.0;
      unsigned nf = 0;
#pragma unroll
      for (int u = 0; u < SB_NCH; ++u) nf |= ((unsigned)__double2hiint(v[u]) & 0x7ff00000u) == 0x7ff00000u ? 1u : 0u;
      if (nf) {
        bad = true;
        if (do_scrub) {
#pragma unroll
          for (int u = 0; u < SB_NCH; ++u) v[u] = scrub(v[u], true, bad);
        }
      }
      if (m.kind == 1) {                                        // force rows: A = R * blank2J (lammps_snap.py:493-502)
#pragma unroll
        for (int u = 0; u < SB_NCH; ++u)
          if (srcc[u] != INT_MIN) orow[lane + 32 * u] = v[u] * pref[u];
      } else if (m.kind == 2) {                                 // virial rows: (1.6021765e6*R)/V * blank2J (:526-536)
#pragma unroll
        for (int u = 0; u < SB_NCH; ++u)
          if (srcc[u] != INT_MIN) orow[lane + 32 * u] = ((srcc[u] >= 0) ? (VIRIAL_UNIT * v[u]) / m.div : 0.0) * pref[u];
      } else {                                                  // energy row: R/N (+ type fractions) * blank2J (:435-467)
#pragma unroll
        for (int u = 0; u < SB_NCH; ++u)
          if (srcc[u] != INT_MIN) {
            const double val = (srcc[u] >= 0) ? v[u] / m.div
                                              : p.type_fraction[(size_t)m.cfg * p.numtypes + (-srcc[u] - 1)];
            orow[lane + 32 * u] = val * pref[u];
          }
      }
    }
    // b and w of the warp's rows (reference column = last raw column)
    if (lane < SB_RPW) {
      const int lr = warp * SB_RPW + lane;
      const SbRowMeta m = s_meta[slot][lr];
      if (m.kind >= 0) {
        const double ref = scrub(rin[(size_t)lr * ldr + kraw], do_scrub, bad);
        p.b[row0 + i0 + lr] = (m.kind == 0) ? (m.truth - ref) / m.div : m.truth - ref;   // :473, :506-507, :540-541
        p.w[row0 + i0 + lr] = m.wv;
      }
    }
    __syncwarp();
    if (lane == 0)   // raw slot and its metadata consumed
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bk_smem_u32(&s_empty[slot])) : "memory");
  }
  if (p.nonfinite && __any_sync(0xffffffffu, bad) && lane == 0) atomicAdd(p.nonfinite, 1);
}


// row -> configuration map of a batch: row_cfg[i] = largest c with out_row_off[c] <= out_row_off[0] + i
__global__ void __launch_bounds__(256) row_map_kernel(const int64_t* __restrict__ out_row_off, int ncfg,
                                                      int32_t* __restrict__ row_cfg, int64_t n_rows) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  const int64_t row = __ldg(out_row_off) + i;
  int lo = 0, hi = ncfg - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(out_row_off + mid) <= row) lo = mid; else hi = mid - 1;
  }
  row_cfg[i] = lo;
}

}  // namespace

int fsb_launch_row_map(const int64_t* out_row_off, int ncfg, int32_t* row_cfg, int64_t n_rows, cudaStream_t s) {
  if (n_rows == 0) return FSB_OK;
  row_map_kernel<<<(unsigned)fsb_ceil_div(n_rows, 256), 256, 0, s>>>(out_row_off, ncfg, row_cfg, n_rows);
  FSB_LAUNCH_CHECK("row_map_kernel");
  return FSB_OK;
}

size_t fsb_residual_ws_bytes(const fsb_context* h, int64_t n_rows, int k) {
  RowPlan pl = plan_rows(h, n_rows);
  return (size_t)pl.ncta * k * sizeof(double);
}

int fsb_launch_residual(const fsb_context* h, const double* A, int64_t lda, const double* b, const double* w,
                        const uint8_t* testing, int64_t n_rows, int k, const double* x, double* g, void* ws,
                        size_t ws_bytes, cudaStream_t s) {
  RowPlan pl = plan_rows(h, n_rows);
  if (ws_bytes < (size_t)pl.ncta * k * sizeof(double)) return FSB_ERR_WORKSPACE_TOO_SMALL;
  int nparts = pl.ncta;
  int st;
  if (rowpass_bulk_ok(h, A, lda, n_rows, k) && h->sm_count <= pl.ncta)
    st = launch_rowpass_bulk(h, A, b, w, testing, n_rows, k, x, (double*)ws, &nparts, s);
  else
    st = launch_rowpass<true>(h, A, lda, b, w, testing, n_rows, k, x, (double*)ws, s);
  if (st != FSB_OK) return st;
  colsum_reduce_kernel<<<(unsigned)fsb_ceil_div(k, 32), 1024, 0, s>>>((const double*)ws, nparts, k, g);
  FSB_LAUNCH_CHECK("colsum_reduce_kernel");
  return FSB_OK;
}

int fsb_launch_group_stats(const fsb_context* h, const double* A, int64_t lda, const double* b, const double* w,
                           const int32_t* gid, int64_t n_rows, int k, const double* x, int n_groups, double* stats,
                           cudaStream_t s) {
  if (n_rows == 0) return FSB_OK;
  const size_t smem = (size_t)n_groups * GS_NSTAT * sizeof(double);
  if (smem > h->smem_optin) return FSB_ERR_UNSUPPORTED;
  int64_t nwarps = (int64_t)h->sm_count * 8 * 4;
  const int64_t maxw = fsb_ceil_div(n_rows, 32);
  if (nwarps > maxw) nwarps = maxw;
  if (nwarps < 8) nwarps = 8;
  nwarps = fsb_round_up(nwarps, 8);
  const int64_t rows_per_warp = fsb_ceil_div(n_rows, nwarps);
  const int ncta = (int)(nwarps / 8);
  const int npl = (k + 31) / 32;
#define FSB_GS(NPL)                                                                                          \
  do {                                                                                                       \
    FSB_CUDA_TRY(cudaFuncSetAttribute(group_stats_kernel<NPL>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                      (int)smem));                                                           \
    group_stats_kernel<NPL><<<ncta, 256, smem, s>>>(A, lda, b, w, gid, n_rows, k, x, n_groups, stats,        \
                                                    rows_per_warp);                                          \
  } while (0)
  if (npl <= 1) FSB_GS(1);
  else if (npl <= 2) FSB_GS(2);
  else if (npl <= 4) FSB_GS(4);
  else if (npl <= 8) FSB_GS(8);
  else if (npl <= 16) FSB_GS(16);
  else if (npl <= 32) FSB_GS(32);
  else if (npl <= 64) FSB_GS(64);
  else return FSB_ERR_UNSUPPORTED;
#undef FSB_GS
  FSB_LAUNCH_CHECK("group_stats_kernel");
  return FSB_OK;
}

int fsb_launch_predict(const fsb_context* h, const double* A, int64_t lda, int64_t n_rows, int k,
                       const double* x, double* y, cudaStream_t s) {
  if (n_rows == 0) return FSB_OK;
  return launch_rowpass<false>(h, A, lda, nullptr, nullptr, nullptr, n_rows, k, x, y, s);
}

int fsb_launch_scatter(const fsb_context* h, const double* raw, const int64_t* raw_row_off,
                       const int64_t* out_row_off, const int32_t* natoms, const double* volume,
                       const double* energy, const double* forces, const double* stress,
                       const double* eweight, const double* fweight, const double* vweight,
                       const double* type_fraction, const double* blank2j, int ncfg, int numtypes,
                       int ncoeff, int flags, double* A, int64_t lda, double* b, double* w,
                       int32_t* nonfinite, const int32_t* row_cfg, int64_t n_rows_hint, cudaStream_t s) {
  ScatterArgs a;
  a.raw = raw; a.raw_row_off = raw_row_off; a.out_row_off = out_row_off; a.natoms = natoms;
  a.volume = volume; a.energy = energy; a.forces = forces; a.stress = stress;
  a.eweight = eweight; a.fweight = fweight; a.vweight = vweight; a.type_fraction = type_fraction;
  a.blank2j = blank2j; a.ncfg = ncfg; a.numtypes = numtypes; a.ncoeff = ncoeff; a.flags = flags;
  a.A = A; a.lda = lda; a.b = b; a.w = w; a.nonfinite = nonfinite; a.row_cfg = row_cfg;
  const bool bzero = flags & FSB_BZEROFLAG;
  const int k = ncoeff * numtypes + (bzero ? 0 : numtypes);
  {
    // bulk-copy staged fast path (see scatter_bulk_kernel)
    static int off = -1;
    if (off < 0) off = getenv("FSB_SCATTER_NO_BULK") ? 1 : 0;
    const int all_rows = FSB_ROWS_ENERGY | FSB_ROWS_FORCE | FSB_ROWS_STRESS;
    if (!off && (flags & all_rows) == all_rows && row_cfg && lda == k && k <= SB_MAXK &&
        n_rows_hint >= (int64_t)h->sm_count * SB_ROWS * 4) {
      const int ldr = ncoeff * numtypes + 1;
      const size_t raw_b = (size_t)SB_ROWS * ldr * sizeof(double);
      const size_t fixed = 64;
      int nstage = (int)((h->smem_optin - 2048 - fixed) / raw_b);
      if (nstage > 6) nstage = 6;
      if (nstage >= 2) {
        const size_t smem_b = (size_t)nstage * raw_b + fixed;
        const int64_t ntiles = fsb_ceil_div(n_rows_hint, SB_ROWS);
        const int grid = h->sm_count;
        const int64_t tiles_per_cta = fsb_ceil_div(ntiles, grid);
        FSB_CUDA_TRY(cudaFuncSetAttribute(scatter_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem_b));
        scatter_bulk_kernel<<<grid, SB_THREADS, smem_b, s>>>(a, n_rows_hint, tiles_per_cta, nstage);
        FSB_LAUNCH_CHECK("scatter_bulk_kernel");
        return FSB_OK;
      }
    }
  }
  const size_t smem = (size_t)k * (sizeof(double) + sizeof(int));
  // tiles of SC_TILE rows, grid-stride; 3 CTAs of 8 warps resident per SM
  int64_t ctas = fsb_ceil_div(n_rows_hint, SC_TILE);
  const int64_t cap = (int64_t)h->sm_count * 3;
  if (ctas > cap) ctas = cap;
  if (ctas < 1) ctas = 1;
  scatter_kernel<<<(unsigned)ctas, 256, smem, s>>>(a, n_rows_hint);
  FSB_LAUNCH_CHECK("scatter_kernel");
  return FSB_OK;
}
