// K6: equilibrated Cholesky factorisation of (G + alpha I) and triangular solves, fp64.
//
// Replaces, on the reduced (k x k) system, what the reference delegates to LAPACK:
//   scipy.linalg.lstsq(aw, bw, 1.0e-13)            fitsnap3lib/solvers/svd.py:54
//   sklearn Ridge(alpha, fit_intercept=False).fit  fitsnap3lib/solvers/ridge.py:49-57
//   inv(XtX + alpha I) @ Xty                       fitsnap3lib/lib/ridge_solver/regressor.py:10-16
// The accuracy lost by squaring the condition number is recovered by iterative refinement
// with a residual streamed from A itself (stream_ops.cu, fsb_residual); see DESIGN.md.
//
// Layout of the caller-owned `factor` buffer (doubles), kp = round_up(k, 64):
//   d[kp]      power-of-two column scales (0 => coefficient pinned to 0)
//   flag[kp]   1.0 => column dropped during factorisation (pivot below tolerance)
//   L[kp*kp]   row-major, pitch kp; lower triangle holds the Cholesky factor of S = D (G+alpha I) D
//   Linv[kp*32] (k <= 128) inverses of the 32x32 diagonal blocks of L: the small solve is then four
//              block steps of tiny mat-vecs instead of k dependent scalar steps
//   Linv[kp2*kp2] (k > 128) the explicit inverse of L (lower triangle, pitch kp2 = 64 * next power of two of
//              kp/64), built once per factorisation by block doubling (batched 64x64-tile GEMMs); every solve /
//              refinement step is then two GEMVs spread over the whole GPU instead of a single-CTA substitution
//
// The factorisation is a right-looking blocked Cholesky with 64-wide panels, three small
// kernels per panel (diag potrf / panel trsm / trailing syrk).  k <= 1024 means <= 48
// launches of latency-bound work (k^3/3 = 0.36 GFLOP at k = 1024): negligible next to the
// Gram pass, so it is written for robustness, not peak.
#include "fsb_common.cuh"
#include <float.h>
#include <stdlib.h>

namespace {

constexpr int NB = FSB_NB;
constexpr int NBP = NB + 1;  // padded smem pitch

struct FactorView {
  double* d;
  double* flag;
  double* L;
  double* Linv;   // k <= 128: inverses of the 32x32 diagonal blocks of L, [kp/32][32][32]; else L^-1, pitch kp2
  int kp;
  int kp2;        // pitch of the explicit inverse (k > 128)
};

__host__ __device__ inline FactorView view_factor(void* buf, int k) {
  FactorView v;
  v.kp = ((k + NB - 1) / NB) * NB;
  if (v.kp == 0) v.kp = NB;
  v.d = (double*)buf;
  v.flag = v.d + v.kp;
  v.L = v.flag + v.kp;
  v.Linv = v.L + (size_t)v.kp * v.kp;
  int np2 = 1;
  while (np2 * NB < v.kp) np2 *= 2;
  v.kp2 = np2 * NB;
  return v;
}

__device__ __forceinline__ double pow2_scale(double g) {
  // d = 2^-floor(e/2) with g = m 2^e, m in [0.5,1)  =>  g d^2 in [0.5, 2)
  int e;
  frexp(g, &e);
  int h = (e >= 0) ? (e / 2) : -((-e + 1) / 2);
  return ldexp(1.0, -h);
}

// S = D (G + alpha I) D on the lower triangle; padding rows/cols get the identity.
__global__ void equilibrate_kernel(const double* __restrict__ gaug, int k, double alpha, FactorView f,
                                   int32_t* info) {
  const int i = blockIdx.x;
  const int ka = k + 1;
  const int kp = f.kp;
  double gii = 0.0, di = 0.0;
  if (i < k) {
    gii = gaug[(size_t)i * ka + i] + alpha;
    di = (gii > 0.0 && gii < DBL_MAX) ? pow2_scale(gii) : 0.0;
  }
  if (threadIdx.x == 0) {
    f.d[i] = di;
    f.flag[i] = 0.0;
    if (i < k && di == 0.0) atomicAdd(&info[FSB_INFO_NUM_PINNED], 1);
  }
  for (int j = threadIdx.x; j <= i; j += blockDim.x) {
    double v;
    if (i >= k || di == 0.0) {
      v = (i == j) ? 1.0 : 0.0;
    } else {
      const double gjj = gaug[(size_t)j * ka + j] + alpha;
      const double dj = (gjj > 0.0 && gjj < DBL_MAX) ? pow2_scale(gjj) : 0.0;
      if (dj == 0.0) v = 0.0;
      else v = (i == j) ? di * gii * di : di * gaug[(size_t)i * ka + j] * dj;
    }
    f.L[(size_t)i * kp + j] = v;
  }
  // clear the strict upper part of this row so the buffer is fully defined
  for (int j = i + 1 + threadIdx.x; j < kp; j += blockDim.x) f.L[(size_t)i * kp + j] = 0.0;
}

// Cholesky of the 64x64 diagonal block of panel p, in shared memory.
__global__ void __launch_bounds__(256) potrf_diag_kernel(FactorView f, int p, double tol, int32_t* info) {
  __shared__ double s[NB][NBP];
  __shared__ int dropped[NB];
  const int kp = f.kp;
  double* blk = f.L + (size_t)(p * NB) * kp + p * NB;
  for (int idx = threadIdx.x; idx < NB * NB; idx += blockDim.x) {
    const int r = idx / NB, c = idx % NB;
    s[r][c] = (c <= r) ? blk[(size_t)r * kp + c] : 0.0;
  }
  __syncthreads();
  if (threadIdx.x < NB) dropped[threadIdx.x] = 0;
  __syncthreads();
  // diag(S) lies in [0.5, 2) after equilibration, so `tol` is an absolute pivot threshold.
  for (int j = 0; j < NB; ++j) {
    if (threadIdx.x == 0) {
      const double piv = s[j][j];
      if (!(piv > tol)) {
        dropped[j] = 1;
        s[j][j] = 1.0;
      } else {
        s[j][j] = sqrt(piv);
      }
    }
    __syncthreads();
    const double ljj = s[j][j];
    const int dj = dropped[j];
    for (int r = j + 1 + threadIdx.x; r < NB; r += blockDim.x) s[r][j] = dj ? 0.0 : s[r][j] / ljj;
    __syncthreads();
    if (!dj) {
      // trailing update of the lower triangle: (r, c) with j < c <= r
      const int m = NB - 1 - j;
      for (int idx = threadIdx.x; idx < m * m; idx += blockDim.x) {
        const int r = j + 1 + idx / m, c = j + 1 + idx % m;
        if (c <= r) s[r][c] -= s[r][j] * s[c][j];
      }
    }
    __syncthreads();
  }
  for (int idx = threadIdx.x; idx < NB * NB; idx += blockDim.x) {
    const int r = idx / NB, c = idx % NB;
    if (c <= r) blk[(size_t)r * kp + c] = s[r][c];
  }
  if (threadIdx.x < NB && dropped[threadIdx.x]) {
    const int col = p * NB + threadIdx.x;
    f.flag[col] = 1.0;
    f.d[col] = 0.0;
    atomicAdd(&info[FSB_INFO_NUM_DEFICIENT], 1);
    atomicExch(&info[FSB_INFO_STATUS], 1);
    atomicMin(&info[FSB_INFO_FIRST_BAD_COLUMN], col);
  }
}

// Cholesky of the 64x64 diagonal block of panel p AND its inverse, one CTA of 16 x 16 threads, compact loops.
// (The first version -- potrf_diag_kernel above, right-looking with three block barriers per column, followed by
// trsm_panel_kernel / trtri_diag_kernel whose fully unrolled 64-step substitutions run ONCE per launch -- spent
// 64 + 42 us per panel at k = 1000 (profiles/r02_launches_bench_default.csv): barrier latency in the first,
// instruction-cache misses on ~6000 straight-line instructions in the other two.  A left-looking version with one
// barrier per column but a 16-slot dot product per thread and step took 76 us: ~220 instructions per step.)
// Both phases are right-looking in "unscaled" form, so that a step is: read the pivot, ONE reciprocal, a rank-1 update
// of at most 4 x 4 elements per thread, ONE block barrier -- no column / row scaling inside the loops:
//   factor : column j of S keeps the values it has when pivot p_j is reached (= L[r][j] sqrt(p_j));
//            S[r][c] -= S[r][j] S[c][j] / p_j for j < c <= r.  ip[j] = 1 / p_j, or 0 for a dropped column.
//   inverse: Y = I; row r of Y is final (up to the factor 1 / L[r][r]) when step r is reached;
//            Y[r2][c] -= (S[r2][r] / p_r) Y[r][c] for r2 > r, c <= r   (L[r2][r] / L[r][r] = S[r2][r] / p_r).
//   after the loops: L[r][c] = S[r][c] / sqrt(p_c), X[r][c] = Y[r][c] / sqrt(p_r).
// X goes to the diagonal block of f.Linv: the panel solve becomes a 64x64x64 product (trsm_gemm_kernel) and
// trtri_diag_kernel is not needed any more.
__global__ void __launch_bounds__(256) potrf_inv_kernel(FactorView f, int p, double tol, int32_t* info) {
  extern __shared__ double pi_sm[];             // s[NB][NBP] | y[NB][NBP]: 66.6 KB, dynamic (above the static limit)
  double (*s)[NBP] = reinterpret_cast<double (*)[NBP]>(pi_sm);
  double (*y)[NBP] = reinterpret_cast<double (*)[NBP]>(pi_sm + NB * NBP);
  __shared__ double pivv[NB], ipv[NB];          // pivot (0: dropped), reciprocal pivot (0: dropped)
  const int kp = f.kp, tid = threadIdx.x;
  double* blk = f.L + (size_t)(p * NB) * kp + p * NB;
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int r = idx / NB, c = idx % NB;
    s[r][c] = (c <= r) ? blk[(size_t)r * kp + c] : 0.0;
    y[r][c] = (c == r) ? 1.0 : 0.0;
  }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;
  // diag(S) lies in [0.5, 2) after equilibration, so `tol` is an absolute pivot threshold.
  for (int j = 0; j < NB; ++j) {
    const double pj = s[j][j];
    const bool drop = !(pj > tol);
    const double inv = drop ? 0.0 : 1.0 / pj;
    if (tid == 0) { pivv[j] = drop ? 0.0 : pj; ipv[j] = inv; }
    if (!drop) {
      double cj[4];
#pragma unroll
      for (int b2 = 0; b2 < 4; ++b2) cj[b2] = s[tx + 16 * b2][j];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int r = ty + 16 * a;
        if (r > j) {
          const double fr = s[r][j] * inv;
#pragma unroll
          for (int b2 = 0; b2 < 4; ++b2) {
            const int c = tx + 16 * b2;
            if (c > j && c <= r) s[r][c] -= fr * cj[b2];
          }
        }
      }
    }
    __syncthreads();
  }
  for (int r = 0; r < NB; ++r) {
    const double ipr = ipv[r];
    if (ipr != 0.0) {
      double yr[4];
#pragma unroll
      for (int b2 = 0; b2 < 4; ++b2) yr[b2] = y[r][tx + 16 * b2];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int r2 = ty + 16 * a;
        if (r2 > r) {
          const double fr = s[r2][r] * ipr;
#pragma unroll
          for (int b2 = 0; b2 < 4; ++b2) {
            const int c = tx + 16 * b2;
            if (c <= r) y[r2][c] -= fr * yr[b2];
          }
        }
      }
    }
    __syncthreads();
  }
  double* out = f.Linv + (size_t)(p * NB) * f.kp2 + p * NB;
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int r = idx / NB, c = idx % NB;
    const double pc = pivv[c], pr = pivv[r];
    if (c <= r) {
      double v;
      if (pc == 0.0) v = (r == c) ? 1.0 : 0.0;                 // dropped column: e_c
      else v = (r == c) ? sqrt(pc) : s[r][c] * rsqrt(pc);
      blk[(size_t)r * kp + c] = v;
    }
    out[(size_t)r * f.kp2 + c] = (c <= r) ? y[r][c] * (pr == 0.0 ? 1.0 : rsqrt(pr)) : 0.0;
  }
  if (tid < NB && pivv[tid] == 0.0) {
    const int col = p * NB + tid;
    f.flag[col] = 1.0;
    f.d[col] = 0.0;
    atomicAdd(&info[FSB_INFO_NUM_DEFICIENT], 1);
    atomicExch(&info[FSB_INFO_STATUS], 1);
    atomicMin(&info[FSB_INFO_FIRST_BAD_COLUMN], col);
  }
}

// Panel solve as a product: L[bi][p] <- A[bi][p] * Linv_pp^T (columns of dropped pivots forced to 0), in place.
__global__ void __launch_bounds__(256) trsm_gemm_kernel(FactorView f, int p) {
  constexpr int MH = 32;
  __shared__ double la[NB][MH + 1];
  __shared__ double lb[NB][MH + 1];
  const int kp = f.kp;
  const int bi = p + 1 + blockIdx.x;
  double* pa = f.L + (size_t)(bi * NB) * kp + p * NB;
  const double* pb = f.Linv + (size_t)(p * NB) * f.kp2 + p * NB;
  const int tr = (threadIdx.x / 16) * 4, tc = (threadIdx.x % 16) * 4;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  constexpr int EPT = NB * MH / 256;          // next half prefetched into registers, as in syrk_update_kernel
  double ra[EPT], rb[EPT];
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    const int idx = threadIdx.x + 256 * e, r = idx / MH, c = idx % MH;
    ra[e] = pa[(size_t)r * kp + c];
    rb[e] = pb[(size_t)r * f.kp2 + c];
  }
  for (int mh = 0; mh < NB; mh += MH) {
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const int idx = threadIdx.x + 256 * e, r = idx / MH, c = idx % MH;
      la[r][c] = ra[e];
      lb[r][c] = rb[e];
    }
    __syncthreads();
    if (mh + MH < NB) {
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        const int idx = threadIdx.x + 256 * e, r = idx / MH, c = idx % MH;
        ra[e] = pa[(size_t)r * kp + mh + MH + c];
        rb[e] = pb[(size_t)r * f.kp2 + mh + MH + c];
      }
    }
    for (int m = 0; m < MH; ++m) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = la[tr + i][m]; b[i] = lb[tc + i][m]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const bool dead = f.flag[p * NB + tc + j] != 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) pa[(size_t)(tr + i) * kp + tc + j] = dead ? 0.0 : acc[i][j];
  }
}

// Panel solve: rows of block-row bi (> p) against L_pp^T.  One thread per row, row in registers.
__global__ void __launch_bounds__(NB) trsm_panel_kernel(FactorView f, int p) {
  __shared__ double lpp[NB][NBP];
  __shared__ double fl[NB];
  const int kp = f.kp;
  const int bi = p + 1 + blockIdx.x;
  const double* dblk = f.L + (size_t)(p * NB) * kp + p * NB;
  for (int idx = threadIdx.x; idx < NB * NB; idx += blockDim.x) {
    const int r = idx / NB, c = idx % NB;
    lpp[r][c] = __ldg(dblk + (size_t)r * kp + c);
  }
  if (threadIdx.x < NB) fl[threadIdx.x] = f.flag[p * NB + threadIdx.x];
  __syncthreads();
  double* row = f.L + (size_t)(bi * NB + threadIdx.x) * kp + p * NB;
  double x[NB];
#pragma unroll
  for (int c = 0; c < NB; ++c) x[c] = row[c];
#pragma unroll
  for (int c = 0; c < NB; ++c) {
    const double xc = (fl[c] != 0.0) ? 0.0 : x[c] / lpp[c][c];
    x[c] = xc;
#pragma unroll
    for (int c2 = c + 1; c2 < NB; ++c2) x[c2] -= xc * lpp[c2][c];
  }
#pragma unroll
  for (int c = 0; c < NB; ++c) row[c] = x[c];
}

// Trailing update: C[bi][bj] -= L[bi][p] L[bj][p]^T for p < bj <= bi.
__global__ void __launch_bounds__(256) syrk_update_kernel(FactorView f, int p) {
  constexpr int MH = 32;  // inner dimension staged in halves to stay under 48 KB static smem
  __shared__ double la[NB][MH + 1];
  __shared__ double lb[NB][MH + 1];
  const int kp = f.kp;
  // decode lower-triangular tile index
  int t = blockIdx.x;
  int ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  while (ti * (ti + 1) / 2 > t) --ti;
  const int tj = t - ti * (ti + 1) / 2;
  const int bi = p + 1 + ti, bj = p + 1 + tj;
  const double* pa = f.L + (size_t)(bi * NB) * kp + p * NB;
  const double* pb = f.L + (size_t)(bj * NB) * kp + p * NB;
  const int tr = (threadIdx.x / 16) * 4, tc = (threadIdx.x % 16) * 4;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  // the next 32-column half is loaded into registers while the current one is multiplied (these kernels are a chain
  // of dependent launches of ~20 us each: the global-load latency of a chunk was not overlapped with anything)
  constexpr int EPT = NB * MH / 256;
  double ra[EPT], rb[EPT];
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    const int idx = threadIdx.x + 256 * e, r = idx / MH, c = idx % MH;
    ra[e] = __ldg(pa + (size_t)r * kp + c);
    rb[e] = __ldg(pb + (size_t)r * kp + c);
  }
  for (int mh = 0; mh < NB; mh += MH) {
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const int idx = threadIdx.x + 256 * e, r = idx / MH, c = idx % MH;
      la[r][c] = ra[e];
      lb[r][c] = rb[e];
    }
    __syncthreads();
    if (mh + MH < NB) {
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        const int idx = threadIdx.x + 256 * e, r = idx / MH, c = idx % MH;
        ra[e] = __ldg(pa + (size_t)r * kp + mh + MH + c);
        rb[e] = __ldg(pb + (size_t)r * kp + mh + MH + c);
      }
    }
    for (int m = 0; m < MH; ++m) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = la[tr + i][m]; b[i] = lb[tc + i][m]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
  double* pc = f.L + (size_t)(bi * NB) * kp + bj * NB;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = tr + i, c = tc + j;
      if (bi != bj || c <= r) pc[(size_t)r * kp + c] -= acc[i][j];
    }
}

// x_out = x_in + D L^-T L^-1 D (rhs - alpha x_in); single CTA, y kept in shared memory.
__global__ void __launch_bounds__(256) trsv_kernel(FactorView f, int k, const double* __restrict__ rhs,
                                                   int64_t rhs_stride, double alpha,
                                                   const double* __restrict__ x_in, double* __restrict__ x_out) {
  extern __shared__ double sm[];
  const int kp = f.kp;
  double* y = sm;                 // kp
  double* blk = sm + kp;          // NB x NBP
  __shared__ double s_flag[NB];
  const int np = kp / NB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;

  for (int i = tid; i < kp; i += blockDim.x) {
    double r = 0.0;
    if (i < k) {
      const double xi = x_in ? x_in[i] : 0.0;
      r = f.d[i] * (rhs[(size_t)i * rhs_stride] - alpha * xi);
    }
    y[i] = r;
  }
  __syncthreads();

  // ---- forward: L y = r
  for (int p = 0; p < np; ++p) {
    const double* dblk = f.L + (size_t)(p * NB) * kp + p * NB;
    for (int idx = tid; idx < NB * NB; idx += blockDim.x) {
      const int r = idx / NB, c = idx % NB;
      blk[r * NBP + c] = __ldg(dblk + (size_t)r * kp + c);
    }
    if (tid < NB) s_flag[tid] = f.flag[p * NB + tid];
    __syncthreads();
    if (warp == 0) {
      double* yp = y + p * NB;
      for (int c = 0; c < NB; ++c) {
        double part = 0.0;
        for (int m = lane; m < c; m += 32) part += blk[c * NBP + m] * yp[m];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) {
          const double v = (yp[c] - part) / blk[c * NBP + c];
          yp[c] = (s_flag[c] != 0.0) ? 0.0 : v;
        }
        __syncwarp();
      }
    }
    __syncthreads();
    // update the remaining right-hand side: warp per row, lanes across the 64 panel columns
    const double y0 = y[p * NB + lane], y1 = y[p * NB + 32 + lane];
    for (int i = (p + 1) * NB + warp; i < kp; i += nwarp) {
      const double* lrow = f.L + (size_t)i * kp + p * NB;
      double part = __ldg(lrow + lane) * y0 + __ldg(lrow + 32 + lane) * y1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (lane == 0) y[i] -= part;
    }
    __syncthreads();
  }

  // ---- backward: L^T z = y
  for (int p = np - 1; p >= 0; --p) {
    const double* dblk = f.L + (size_t)(p * NB) * kp + p * NB;
    for (int idx = tid; idx < NB * NB; idx += blockDim.x) {
      const int r = idx / NB, c = idx % NB;
      blk[r * NBP + c] = __ldg(dblk + (size_t)r * kp + c);
    }
    __syncthreads();
    if (warp == 0) {
      double* yp = y + p * NB;
      for (int c = NB - 1; c >= 0; --c) {
        double part = 0.0;
        for (int m = c + 1 + lane; m < NB; m += 32) part += blk[m * NBP + c] * yp[m];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) yp[c] = (yp[c] - part) / blk[c * NBP + c];
        __syncwarp();
      }
    }
    __syncthreads();
    // y[j] -= sum_r L[p*NB + r][j] z[r] for j < p*NB: thread per column (coalesced over j)
    for (int j = tid; j < p * NB; j += blockDim.x) {
      double part = 0.0;
      const double* lcol = f.L + (size_t)(p * NB) * kp + j;
#pragma unroll 8
      for (int r = 0; r < NB; ++r) part += __ldg(lcol + (size_t)r * kp) * y[p * NB + r];
      y[j] -= part;
    }
    __syncthreads();
  }

  for (int i = tid; i < k; i += blockDim.x) {
    const double xi = x_in ? x_in[i] : 0.0;
    x_out[i] = xi + f.d[i] * y[i];
  }
}

// ---------------------------------------------------------------------------------------------
// Small systems (k <= 128, e.g. BASELINE config 2: k = 100): the whole factorisation runs in ONE CTA
// with S resident in shared memory (pitch k|1 -> conflict-free column walks), one barrier per
// column: the trailing update uses the unscaled column and 1/pivot, columns are scaled at the end.
constexpr int SMALL_K = 128;

__global__ void __launch_bounds__(512) small_factor_kernel(const double* __restrict__ gaug, int k, double alpha,
                                                           FactorView f, double tol, int32_t* info) {
  extern __shared__ double sm[];
  const int P = k | 1;
  double* S = sm;              // k x P
  double* piv = sm + (size_t)k * P;   // k   (pivot, or 0 when the column was dropped)
  double* dsc = piv + k;       // k   scales
  const int ka = k + 1;
  const int kp = f.kp;
  const int tid = threadIdx.x, nt = blockDim.x;

  for (int i = tid; i < k; i += nt) {
    const double g = gaug[(size_t)i * ka + i] + alpha;
    dsc[i] = (g > 0.0 && g < DBL_MAX) ? pow2_scale(g) : 0.0;
  }
  __syncthreads();
#pragma unroll 4
  for (int idx = tid; idx < k * ka; idx += nt) {      // flat over the first k rows of gaug: coalesced, unrolled
    const int i = idx / ka, j = idx - i * ka;
    const double g = gaug[idx];
    if (j > i || j >= k) continue;
    const double di = dsc[i], dj = dsc[j];
    double v;
    if (di == 0.0 || dj == 0.0) v = (i == j) ? 1.0 : 0.0;
    else v = di * (g + (i == j ? alpha : 0.0)) * dj;
    S[i * P + j] = v;
  }
  __syncthreads();

  // Blocked left-looking Cholesky, panels of PB columns, in "unscaled column" form: column j of S keeps the values
  // it has when its pivot p_j is reached (= L[r][j] sqrt(p_j)); every later use multiplies by 1 / p_j instead
  // (S[r][c] -= S[r][j] S[c][j] / p_j).  No column-scaling pass inside the loop, hence ONE block barrier per column
  // (round 1 scaled the column in place: three barriers and a sqrt + division chain per column, ~1 us each); the
  // columns are scaled once, in parallel, after the loop.  ip[j] = 1 / p_j, or 0 for a dropped column (its
  // contributions vanish).  (1) a new panel is updated with all finished columns at once (dot products over m < p0,
  // no barrier inside), (2) its PB columns are eliminated one by one inside the panel.
  constexpr int PB = 16;
  double* ip = dsc + k;                  // k   reciprocal pivots (shared-memory region sized by the launcher)
  for (int p0 = 0; p0 < k; p0 += PB) {
    const int pw = (k - p0) < PB ? (k - p0) : PB;
    if (p0 > 0) {
      const int nel = (k - p0) * PB;
      for (int idx = tid; idx < nel; idx += nt) {
        const int r = p0 + (idx >> 4), c = p0 + (idx & (PB - 1));
        if (c < p0 + pw && c <= r) {
          const double* sr = S + r * P;
          const double* sc = S + c * P;
          double acc0 = 0.0, acc1 = 0.0;
          int m = 0;
          for (; m + 1 < p0; m += 2) {
            acc0 += sr[m] * sc[m] * ip[m];
            acc1 += sr[m + 1] * sc[m + 1] * ip[m + 1];
          }
          if (m < p0) acc0 += sr[m] * sc[m] * ip[m];
          S[r * P + c] -= acc0 + acc1;
        }
      }
      __syncthreads();
    }
    for (int j = p0; j < p0 + pw; ++j) {
      const double pj = S[j * P + j];
      const bool drop = !(pj > tol);
      const double inv = drop ? 0.0 : 1.0 / pj;
      if (tid == 0) { piv[j] = drop ? 0.0 : pj; ip[j] = inv; }
      const int nc = p0 + pw - (j + 1);      // remaining panel columns
      if (nc > 0 && !drop) {
        const int nel = (k - (j + 1)) * PB;
        for (int idx = tid; idx < nel; idx += nt) {
          const int r = j + 1 + (idx >> 4), c = j + 1 + (idx & (PB - 1));
          if (c < p0 + pw && c <= r) S[r * P + c] -= S[r * P + j] * (S[c * P + j] * inv);
        }
      }
      __syncthreads();
    }
  }
  // scale: L[r][j] = S[r][j] / sqrt(p_j); dropped columns become (1 on the diagonal, 0 below)
  for (int idx = tid; idx < k * k; idx += nt) {
    const int r = idx / k, c = idx - r * k;
    if (c > r) continue;
    const double pc = piv[c];
    double v;
    if (pc == 0.0) v = (r == c) ? 1.0 : 0.0;
    else v = (r == c) ? sqrt(pc) : S[r * P + c] * rsqrt(pc);
    S[r * P + c] = v;
  }
  __syncthreads();

  // scale the columns, publish the factor in the common layout (pitch kp, identity padding)
  if (tid < FSB_INFO_LEN) info[tid] = (tid == FSB_INFO_FIRST_BAD_COLUMN) ? k : 0;
  __syncthreads();
  for (int idx = tid; idx < kp * kp; idx += nt) {
    const int i = idx / kp, j = idx - i * kp;
    double v = 0.0;
    if (i < k && j <= i) {
      v = S[i * P + j];      // already the scaled factor; dropped columns hold (1 on the diagonal, 0 below)
    } else if (i == j) {
      v = 1.0;
    }
    f.L[(size_t)i * kp + j] = v;
  }
  // inverses of the 32x32 diagonal blocks (warp b inverts block b; lane c owns column c of the inverse
  // and runs a forward substitution on e_c with the block's rows broadcast from shared memory)
  {
    const int warp = tid >> 5, lane = tid & 31;
    const int nb32 = kp / 32;
    if (warp < nb32) {
      const int o = warp * 32;
      double x[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int gi = o + i;
        double acc = (i == lane) ? 1.0 : 0.0;
#pragma unroll
        for (int m = 0; m < 32; ++m)
          if (m < i) acc -= ((gi < k) ? S[gi * P + o + m] : 0.0) * x[m];
        const double lii = (gi < k) ? S[gi * P + gi] : 1.0;
        x[i] = (i >= lane) ? acc / lii : 0.0;
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) f.Linv[(size_t)warp * 1024 + i * 32 + lane] = x[i];
    }
  }
  for (int i = tid; i < kp; i += nt) {
    const bool real = i < k;
    const bool dropped = real && piv[i] == 0.0 && dsc[i] != 0.0;
    const bool pinned = real && dsc[i] == 0.0;
    f.d[i] = (real && !dropped) ? dsc[i] : 0.0;
    f.flag[i] = dropped ? 1.0 : 0.0;
    if (pinned) atomicAdd(&info[FSB_INFO_NUM_PINNED], 1);
    if (dropped) {
      atomicAdd(&info[FSB_INFO_NUM_DEFICIENT], 1);
      atomicExch(&info[FSB_INFO_STATUS], 1);
      atomicMin(&info[FSB_INFO_FIRST_BAD_COLUMN], i);
    }
  }
}

// x_out = x_in + D L^-T L^-1 D (rhs - alpha x_in) for k <= 128, blocked by 32: with the inverses of
// the diagonal blocks precomputed by small_factor_kernel a substitution is 4 block steps, each two
// tiny mat-vecs done by the whole CTA (4 threads per row, shuffle-combined), instead of k dependent
// scalar steps.  128 threads: thread t -> row (t >> 2) of the current block, quarter (t & 3).
// Launched with SOLVE_THREADS threads: all of them stage the factor (the load phase was most of this kernel when 128
// threads fetched 131 KB with 8 scalar loads in flight each), the first SMALL_K then run the block substitution on a
// named barrier of their own and the rest leave.
constexpr int SOLVE_THREADS = 512;
__global__ void __launch_bounds__(SOLVE_THREADS) small_solve_kernel(FactorView f, int k, const double* __restrict__ rhs,
                                                              int64_t rhs_stride, double alpha,
                                                              const double* __restrict__ x_in,
                                                              double* __restrict__ x_out) {
  extern __shared__ double sm[];
  const int kp = f.kp;                   // 64 or 128
  const int P = kp | 1;
  double* L = sm;                        // kp x P (lower triangle used)
  double* Li = L + (size_t)kp * P;       // [kp/32][32][33]
  double* y = Li + (size_t)(kp / 32) * 32 * 33;   // kp
  double* tb = y + kp;                   // 32 (block right-hand side)
  const int tid = threadIdx.x, nt = blockDim.x;
  {
    // lower-triangular 32 x 32 tiles only, 16-byte loads (kp is a multiple of 64: rows are 16-byte aligned)
    const int nb = kp / 32;
    const int ntile = nb * (nb + 1) / 2;
    const double2* L2 = reinterpret_cast<const double2*>(f.L);
#pragma unroll 4
    for (int idx = tid; idx < ntile * 512; idx += nt) {
      const int t = idx >> 9, e = idx & 511;
      int tr = 0, rem = t;
      while (rem > tr) { rem -= tr + 1; ++tr; }          // tile (tr, rem), rem <= tr
      const int r = tr * 32 + (e >> 4), c = rem * 32 + 2 * (e & 15);
      const double2 v = __ldg(L2 + ((size_t)r * kp + c) / 2);
      if (c <= r) L[r * P + c] = v.x;
      if (c + 1 <= r) L[r * P + c + 1] = v.y;
    }
    const int nli = (kp / 32) * 1024;
#pragma unroll 4
    for (int idx = tid; idx < nli; idx += nt) {
      const int b = idx >> 10, r = (idx >> 5) & 31, c = idx & 31;
      Li[(b * 32 + r) * 33 + c] = __ldg(f.Linv + idx);
    }
  }
  for (int i = tid; i < kp; i += nt) {
    double v = 0.0;
    if (i < k) v = f.d[i] * (rhs[(size_t)i * rhs_stride] - alpha * (x_in ? x_in[i] : 0.0));
    y[i] = v;
  }
  __syncthreads();
  if (tid >= SMALL_K) return;
#define FSB_SOLVE_SYNC() asm volatile("bar.sync 1, %0;" ::"n"(SMALL_K) : "memory")
  const int nb32 = kp / 32;
  const int row = tid >> 2, q = tid & 3;     // 32 rows x 4 quarters
  // ---- forward: L y = r
  for (int b = 0; b < nb32; ++b) {
    const int gi = b * 32 + row;
    double part = 0.0;
    const int ncol = b * 32;                 // columns already solved
    for (int c = q; c < ncol; c += 4) part += L[gi * P + c] * y[c];
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    if (q == 0) tb[row] = y[gi] - part;
    FSB_SOLVE_SYNC();
    double acc = 0.0;
    for (int m = q; m <= row; m += 4) acc += Li[(b * 32 + row) * 33 + m] * tb[m];
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    FSB_SOLVE_SYNC();                         // everyone is done reading tb / y of this block
    if (q == 0) y[gi] = (gi < k && f.flag[gi] != 0.0) ? 0.0 : acc;
    FSB_SOLVE_SYNC();
  }
  // ---- backward: L^T z = y
  for (int b = nb32 - 1; b >= 0; --b) {
    const int gi = b * 32 + row;
    double part = 0.0;
    for (int c = (b + 1) * 32 + q; c < kp; c += 4) part += L[c * P + gi] * y[c];   // (L^T)[gi][c] = L[c][gi]
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    if (q == 0) tb[row] = y[gi] - part;
    FSB_SOLVE_SYNC();
    double acc = 0.0;
    for (int m = row + q; m < 32; m += 4) acc += Li[(b * 32 + m) * 33 + row] * tb[m];   // (Linv^T)[row][m]
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    FSB_SOLVE_SYNC();
    if (q == 0) y[gi] = acc;
    FSB_SOLVE_SYNC();
  }
  for (int i = tid; i < k; i += SMALL_K) x_out[i] = (x_in ? x_in[i] : 0.0) + f.d[i] * y[i];
#undef FSB_SOLVE_SYNC
}

// ---------------------------------------------------------------------------------------------
// Large systems (k > 128): explicit inverse of the Cholesky factor by block doubling.
//   level 0   Linv_pp = L_pp^-1 for every 64x64 diagonal block (one CTA each, thread per column)
//   level l   blocks of size s = 64 2^l are paired:  Linv21 = -Linv22 (L21 Linv11)
//             as two batched tile GEMMs; the intermediate T = Linv22 L21 lives in the (unused) mirror
//             block of the upper triangle.  Zero blocks of the triangular operands are skipped.
// Dropped columns (flag) need no special care here: their column of L is e_c, and the solve masks y_c.
__global__ void __launch_bounds__(NB) trtri_diag_kernel(FactorView f) {
  __shared__ double l[NB][NBP];
  const int p = blockIdx.x, kp = f.kp;
  const double* blk = f.L + (size_t)(p * NB) * kp + p * NB;
  for (int idx = threadIdx.x; idx < NB * NB; idx += blockDim.x) {
    const int r = idx / NB, c = idx % NB;
    l[r][c] = __ldg(blk + (size_t)r * kp + c);
  }
  __syncthreads();
  // column j of the inverse: forward substitution on e_j, kept in registers
  const int j = threadIdx.x;
  double x[NB];
#pragma unroll
  for (int r = 0; r < NB; ++r) x[r] = (r == j) ? 1.0 : 0.0;
#pragma unroll
  for (int r = 0; r < NB; ++r) {
    const double xr = (r >= j) ? x[r] / l[r][r] : 0.0;
    x[r] = xr;
#pragma unroll
    for (int r2 = r + 1; r2 < NB; ++r2) x[r2] -= xr * l[r2][r];
  }
  double* out = f.Linv + (size_t)(p * NB) * f.kp2 + p * NB;
#pragma unroll
  for (int r = 0; r < NB; ++r) out[(size_t)r * f.kp2 + j] = x[r];
}

// one 64x64 tile of  C = sign * A B  with A (64 x 64 nkb) and B (64 nkb x 64) given by pointers + pitches;
// rows/columns of A or B at or beyond `a_lim` / `b_lim` (global indices kept by the caller) read as zero
struct TileGemm {
  const double* a; int64_t lda;
  const double* b; int64_t ldb;
  double* c; int64_t ldc;
  int kb0, kb1;     // 64-wide k blocks [kb0, kb1)
  double sign;
  bool a_zero;      // the A tile lies outside the stored matrix: the product is zero
};

__device__ __forceinline__ void tile_gemm(const TileGemm& g) {
  constexpr int MH = 16;
  __shared__ double sa[NB][MH + 1];
  __shared__ double sb[MH][NB + 1];
  const int tr = (threadIdx.x / 16) * 4, tc = (threadIdx.x % 16) * 4;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  if (!g.a_zero) {
    constexpr int EPT = NB * MH / 256;        // the next 16-wide chunk is loaded while the current one is multiplied
    double ra[EPT], rb[EPT];
    const int mbeg = g.kb0 * NB, mend = g.kb1 * NB;
    if (mbeg < mend) {
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        const int idx = threadIdx.x + 256 * e;
        ra[e] = g.a[(size_t)(idx / MH) * g.lda + mbeg + idx % MH];
        rb[e] = g.b[(size_t)(mbeg + idx / NB) * g.ldb + idx % NB];
      }
    }
    for (int m0 = mbeg; m0 < mend; m0 += MH) {
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        const int idx = threadIdx.x + 256 * e;
        sa[idx / MH][idx % MH] = ra[e];
        sb[idx / NB][idx % NB] = rb[e];
      }
      __syncthreads();
      if (m0 + MH < mend) {
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          const int idx = threadIdx.x + 256 * e;
          ra[e] = g.a[(size_t)(idx / MH) * g.lda + m0 + MH + idx % MH];
          rb[e] = g.b[(size_t)(m0 + MH + idx / NB) * g.ldb + idx % NB];
        }
      }
#pragma unroll
      for (int m = 0; m < MH; ++m) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { a[i] = sa[tr + i][m]; b[i] = sb[m][tc + i]; }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) g.c[(size_t)(tr + i) * g.ldc + tc + j] = g.sign * acc[i][j];
}

// step 1 of a level: T = Linv22 * L21 (Linv22 lower triangular: k blocks 0..ti), T stored in the mirror block
// step 2:            Linv21 = -T * Linv11 (Linv11 lower triangular: k blocks tj..nb-1)
template <int STEP>
__global__ void __launch_bounds__(256) trtri_level_kernel(FactorView f, int nb) {
  // grid: x = tile (ti * nb + tj) inside the s x s block, y = pair
  const int ti = blockIdx.x / nb, tj = blockIdx.x % nb;
  const int base = blockIdx.y * 2 * nb;               // first 64-block of the pair
  const int np = f.kp / NB;
  const int64_t p2 = f.kp2;
  TileGemm g;
  const int r1 = base + nb;                           // block row of the "2" half
  if (STEP == 1) {
    g.a = f.Linv + (size_t)((r1 + ti) * NB) * p2 + (size_t)r1 * NB;     g.lda = p2;
    g.b = f.L + (size_t)(r1 * NB) * f.kp + (size_t)(base + tj) * NB;    g.ldb = f.kp;
    g.c = f.Linv + (size_t)((base + ti) * NB) * p2 + (size_t)(r1 + tj) * NB;   g.ldc = p2;   // mirror block
    g.kb0 = 0; g.kb1 = ti + 1;
    g.sign = 1.0;
    g.a_zero = (r1 + ti >= np);                       // rows of L beyond the matrix: L21 = 0 there
    if (r1 + g.kb1 > np) g.kb1 = np - r1 > 0 ? np - r1 : 0;   // k blocks of L21 that exist
    if (g.kb1 <= g.kb0) g.a_zero = true;
  } else {
    g.a = f.Linv + (size_t)((base + ti) * NB) * p2 + (size_t)r1 * NB;   g.lda = p2;           // T
    g.b = f.Linv + (size_t)(base * NB) * p2 + (size_t)(base + tj) * NB; g.ldb = p2;
    g.c = f.Linv + (size_t)((r1 + ti) * NB) * p2 + (size_t)(base + tj) * NB;   g.ldc = p2;
    g.kb0 = tj; g.kb1 = nb;
    g.sign = -1.0;
    g.a_zero = (r1 + ti >= np);
  }
  tile_gemm(g);
}

// y = mask(Linv * D (rhs - alpha x_in)): warp per row
__global__ void __launch_bounds__(256) inv_forward_kernel(FactorView f, int k, const double* __restrict__ rhs,
                                                          int64_t rhs_stride, double alpha,
                                                          const double* __restrict__ x_in, double* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= f.kp) return;
  const double* lr = f.Linv + (size_t)row * f.kp2;
  double part = 0.0;
  for (int c = lane; c <= row && c < k; c += 32) {
    const double xi = x_in ? x_in[c] : 0.0;
    part += lr[c] * (f.d[c] * (rhs[(size_t)c * rhs_stride] - alpha * xi));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) y[row] = (row < k && f.flag[row] == 0.0) ? part : 0.0;
}

// x_out = x_in + D Linv^T y: 32 columns per CTA, 8 warps split the rows, fixed-order combine
__global__ void __launch_bounds__(256) inv_backward_kernel(FactorView f, int k, const double* __restrict__ y,
                                                           const double* __restrict__ x_in,
                                                           double* __restrict__ x_out) {
  __shared__ double part[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;
  double acc = 0.0;
  if (col < k)
    for (int r = blockIdx.x * 32 + warp; r < k; r += 8)
      if (r >= col) acc += f.Linv[(size_t)r * f.kp2 + col] * y[r];
  part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && col < k) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += part[q][lane];
    const double xi = x_in ? x_in[col] : 0.0;
    x_out[col] = xi + f.d[col] * t;
  }
}

__global__ void init_info_kernel(int32_t* info, int k) {
  if (threadIdx.x < FSB_INFO_LEN) info[threadIdx.x] = (threadIdx.x == FSB_INFO_FIRST_BAD_COLUMN) ? k : 0;
}

}  // namespace

size_t fsb_factor_bytes_impl(int k) {
  int kp = ((k + NB - 1) / NB) * NB;
  if (kp == 0) kp = NB;
  if (k <= SMALL_K) return ((size_t)2 * kp + (size_t)kp * kp + (size_t)kp * 32) * sizeof(double);
  int np2 = 1;
  while (np2 * NB < kp) np2 *= 2;
  const size_t kp2 = (size_t)np2 * NB;
  return ((size_t)2 * kp + (size_t)kp * kp + kp2 * kp2 + (size_t)kp) * sizeof(double);   // d, flag, L, L^-1, y
}

int fsb_launch_factor(const fsb_context* h, const double* gaug, int k, double alpha, void* factor,
                      size_t factor_bytes, int32_t* info, cudaStream_t s) {
  (void)h;
  if (factor_bytes < fsb_factor_bytes_impl(k)) return FSB_ERR_WORKSPACE_TOO_SMALL;
  FactorView f = view_factor(factor, k);
  if (k <= SMALL_K) {
    const int P = k | 1;
    const size_t smem = ((size_t)k * P + 3 * (size_t)k) * sizeof(double);
    FSB_CUDA_TRY(cudaFuncSetAttribute(small_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    small_factor_kernel<<<1, 512, smem, s>>>(gaug, k, alpha, f, 64.0 * (double)f.kp * DBL_EPSILON, info);
    FSB_LAUNCH_CHECK("small_factor_kernel");
    return FSB_OK;
  }
  init_info_kernel<<<1, 32, 0, s>>>(info, k);
  FSB_LAUNCH_CHECK("init_info_kernel");
  equilibrate_kernel<<<f.kp, 128, 0, s>>>(gaug, k, alpha, f, info);
  FSB_LAUNCH_CHECK("equilibrate_kernel");
  const int np = f.kp / NB;
  const double tol = 64.0 * (double)f.kp * DBL_EPSILON;
  const bool v1 = getenv("FSB_FACTOR_PANEL_V1") != nullptr;    // first-generation panel kernels, kept for cross-checks
  const size_t potrf_smem = (size_t)2 * NB * NBP * sizeof(double);
  if (!v1) {
    FSB_CUDA_TRY(cudaFuncSetAttribute(potrf_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)potrf_smem));
    FSB_CUDA_TRY(cudaMemsetAsync(f.Linv, 0, (size_t)f.kp2 * f.kp2 * sizeof(double), s));
  }
  for (int p = 0; p < np; ++p) {
    const int rem = np - p - 1;
    if (v1) {
      potrf_diag_kernel<<<1, 256, 0, s>>>(f, p, tol, info);
      FSB_LAUNCH_CHECK("potrf_diag_kernel");
      if (rem > 0) {
        trsm_panel_kernel<<<rem, NB, 0, s>>>(f, p);
        FSB_LAUNCH_CHECK("trsm_panel_kernel");
      }
    } else {
      potrf_inv_kernel<<<1, 256, potrf_smem, s>>>(f, p, tol, info);
      FSB_LAUNCH_CHECK("potrf_inv_kernel");
      if (rem > 0) {
        trsm_gemm_kernel<<<rem, 256, 0, s>>>(f, p);
        FSB_LAUNCH_CHECK("trsm_gemm_kernel");
      }
    }
    if (rem > 0) {
      syrk_update_kernel<<<rem * (rem + 1) / 2, 256, 0, s>>>(f, p);
      FSB_LAUNCH_CHECK("syrk_update_kernel");
    }
  }
  // explicit inverse of L by block doubling (the diagonal blocks are already inverted)
  if (v1) {
    FSB_CUDA_TRY(cudaMemsetAsync(f.Linv, 0, (size_t)f.kp2 * f.kp2 * sizeof(double), s));
    trtri_diag_kernel<<<np, NB, 0, s>>>(f);
    FSB_LAUNCH_CHECK("trtri_diag_kernel");
  }
  for (int nb = 1; nb * NB < f.kp2; nb *= 2) {
    const int pairs = f.kp2 / (2 * nb * NB);
    dim3 grid((unsigned)(nb * nb), (unsigned)pairs);
    trtri_level_kernel<1><<<grid, 256, 0, s>>>(f, nb);
    FSB_LAUNCH_CHECK("trtri_level_kernel<1>");
    trtri_level_kernel<2><<<grid, 256, 0, s>>>(f, nb);
    FSB_LAUNCH_CHECK("trtri_level_kernel<2>");
  }
  return FSB_OK;
}

int fsb_launch_factor_solve(const fsb_context* h, const void* factor, int k, const double* rhs,
                            int64_t rhs_stride, double alpha, const double* x_in, double* x_out,
                            cudaStream_t s) {
  FactorView f = view_factor(const_cast<void*>(factor), k);
  if (k <= SMALL_K) {
    const int kp_s = f.kp, P = kp_s | 1;
    const size_t smem_s = ((size_t)kp_s * P + (size_t)(kp_s / 32) * 32 * 33 + kp_s + 32) * sizeof(double);
    FSB_CUDA_TRY(cudaFuncSetAttribute(small_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
    small_solve_kernel<<<1, SOLVE_THREADS, smem_s, s>>>(f, k, rhs, rhs_stride, alpha, x_in, x_out);
    FSB_LAUNCH_CHECK("small_solve_kernel");
    return FSB_OK;
  }
  if (getenv("FSB_SOLVE_TRSV")) {   // the single-CTA substitution, kept for cross-checking the inverse
    const size_t smem = ((size_t)f.kp + NB * NBP) * sizeof(double);
    if (smem > h->smem_optin) return FSB_ERR_UNSUPPORTED;
    FSB_CUDA_TRY(cudaFuncSetAttribute(trsv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    trsv_kernel<<<1, 256, smem, s>>>(f, k, rhs, rhs_stride, alpha, x_in, x_out);
    FSB_LAUNCH_CHECK("trsv_kernel");
    return FSB_OK;
  }
  double* y = f.Linv + (size_t)f.kp2 * f.kp2;
  inv_forward_kernel<<<(unsigned)fsb_ceil_div(f.kp, 8), 256, 0, s>>>(f, k, rhs, rhs_stride, alpha, x_in, y);
  FSB_LAUNCH_CHECK("inv_forward_kernel");
  inv_backward_kernel<<<(unsigned)fsb_ceil_div(k, 32), 256, 0, s>>>(f, k, y, x_in, x_out);
  FSB_LAUNCH_CHECK("inv_backward_kernel");
  return FSB_OK;
}
