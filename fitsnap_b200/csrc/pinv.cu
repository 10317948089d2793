// Minimum-norm fallback for rank-deficient systems: pseudo-inverse of the Gram through a symmetric
// eigendecomposition (parallel cyclic Jacobi), applied as x += G^+ rhs.
//
// scipy.linalg.lstsq(aw, bw, 1e-13) (fitsnap3lib/solvers/svd.py:54, LAPACK gelsd) returns the
// MINIMUM-NORM least-squares solution when columns of aw are linearly dependent (duplicated
// descriptors; all-zero columns are already pinned by the Cholesky path).  Cholesky of the Gram breaks
// down there (fsb_factor reports it in info[] and drops the column: a basic solution).  This path
// reproduces the reference's answer instead:  G = V diag(lambda) V^T  (G is NOT equilibrated: the norm
// being minimised is |x|_2, as in gelsd),  G^+ = V diag(1/lambda_i for lambda_i > rcond*lambda_max) V^T,
// x0 = G^+ c, followed by the same refinement against A as the regular path (x += G^+ aw^T(bw - aw x)),
// which keeps x in the row space.  Singular values are resolved at Gram precision: the cut-off acts
// on eigenvalues of G (sigma^2), default rcond = k * eps, not on sigma/sigma_max = 1e-13.
//
// One CTA, matrix in global memory (L2-resident): rare path, written for robustness.  Round-robin
// tournament ordering: k/2 disjoint rotations per round, k-1 rounds per sweep.
#include "fsb_common.cuh"
#include <float.h>

namespace {

__global__ void __launch_bounds__(1024) jacobi_eig_kernel(const double* __restrict__ gaug, int k, double* __restrict__ Aw,
                                                          double* __restrict__ V, double* __restrict__ lam,
                                                          int max_sweeps, int32_t* info) {
  extern __shared__ double sm[];     // c[m/2], s[m/2], reduction scratch[32]
  const int tid = threadIdx.x, nt = blockDim.x;
  const int ka = k + 1;
  const int m = (k + 1) & ~1;        // players (even)
  const int half = m / 2;
  double* cs_c = sm;
  double* cs_s = sm + half;
  double* red = sm + 2 * half;
  __shared__ int s_p[1024], s_q[1024];
  __shared__ double s_off, s_diag;

  for (int idx = tid; idx < k * k; idx += nt) {
    const int i = idx / k, j = idx - i * k;
    Aw[idx] = gaug[(size_t)i * ka + j];
    V[idx] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  int sweeps = 0;
  for (int sw = 0; sw < max_sweeps; ++sw) {
    // convergence: off-diagonal Frobenius norm against the diagonal
    double off = 0.0, dg = 0.0;
    for (int idx = tid; idx < k * k; idx += nt) {
      const int i = idx / k, j = idx - i * k;
      const double v = Aw[idx];
      if (i == j) dg += v * v; else off += v * v;
    }
    for (int o = 16; o > 0; o >>= 1) { off += __shfl_xor_sync(0xffffffffu, off, o); dg += __shfl_xor_sync(0xffffffffu, dg, o); }
    if ((tid & 31) == 0) { red[tid >> 5] = off; red[32 + (tid >> 5)] = dg; }
    __syncthreads();
    if (tid == 0) {
      double a = 0.0, b = 0.0;
      for (int wv = 0; wv < nt / 32; ++wv) { a += red[wv]; b += red[32 + wv]; }
      s_off = a; s_diag = b;
    }
    __syncthreads();
    if (s_off <= 1e-30 * s_diag || s_off == 0.0) break;
    ++sweeps;
    for (int r = 0; r < m - 1; ++r) {
      // pairs of this round + their rotations (from the matrix as it stands before the round)
      if (tid < half) {
        int p, q;
        if (tid == 0) { p = m - 1; q = r; }
        else { p = (r + tid) % (m - 1); q = (r - tid + (m - 1)) % (m - 1); }
        if (p > q) { const int t = p; p = q; q = t; }
        double c = 1.0, s = 0.0;
        if (q < k) {
          const double apq = Aw[(size_t)p * k + q];
          if (apq != 0.0) {
            const double app = Aw[(size_t)p * k + p], aqq = Aw[(size_t)q * k + q];
            const double tau = (aqq - app) / (2.0 * apq);
            const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + t * t);
            s = t * c;
          }
        }
        s_p[tid] = p; s_q[tid] = q; cs_c[tid] = c; cs_s[tid] = s;
      }
      __syncthreads();
      // rows: A <- J^T A
      for (int idx = tid; idx < half * k; idx += nt) {
        const int t = idx / k, j = idx - t * k;
        const int p = s_p[t], q = s_q[t];
        if (q < k) {
          const double c = cs_c[t], s = cs_s[t];
          const double ap = Aw[(size_t)p * k + j], aq = Aw[(size_t)q * k + j];
          Aw[(size_t)p * k + j] = c * ap - s * aq;
          Aw[(size_t)q * k + j] = s * ap + c * aq;
        }
      }
      __syncthreads();
      // columns: A <- A J, V <- V J
      for (int idx = tid; idx < half * k; idx += nt) {
        const int t = idx / k, i = idx - t * k;
        const int p = s_p[t], q = s_q[t];
        if (q < k) {
          const double c = cs_c[t], s = cs_s[t];
          const double ap = Aw[(size_t)i * k + p], aq = Aw[(size_t)i * k + q];
          Aw[(size_t)i * k + p] = c * ap - s * aq;
          Aw[(size_t)i * k + q] = s * ap + c * aq;
          const double vp = V[(size_t)i * k + p], vq = V[(size_t)i * k + q];
          V[(size_t)i * k + p] = c * vp - s * vq;
          V[(size_t)i * k + q] = s * vp + c * vq;
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < k; i += nt) lam[i] = Aw[(size_t)i * k + i];
  if (tid == 0) info[1] = sweeps;
}

// P = V diag(1/(lambda_i + shift) | lambda_i > cut) V^T ;  cut = rcond * max lambda ; info[0] = numerical rank
// shift = 0: the pseudo-inverse (minimum-norm least squares); shift = alpha > 0: the ridge inverse restricted to the
// numerical range of G (the exact ridge solution has no component in the null space of G: there aw^T bw vanishes)
__global__ void __launch_bounds__(256) pinv_build_kernel(const double* __restrict__ V, const double* __restrict__ lam,
                                                         int k, double rcond, double shift, double* __restrict__ P,
                                                         int32_t* info) {
  extern __shared__ double inv[];   // k
  __shared__ double s_max;
  if (threadIdx.x == 0) {
    double mx = 0.0;
    for (int i = 0; i < k; ++i) mx = fmax(mx, lam[i]);
    s_max = mx;
  }
  __syncthreads();
  const double cut = rcond * s_max;
  int rank = 0;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const bool keep = lam[i] > cut && lam[i] > 0.0;
    inv[i] = keep ? 1.0 / (lam[i] + shift) : 0.0;
  }
  __syncthreads();
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    for (int i = 0; i < k; ++i) rank += (inv[i] != 0.0);
    info[0] = rank;
  }
  const int i = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k || j >= k) return;
  double acc = 0.0;
  for (int mth = 0; mth < k; ++mth) acc += V[(size_t)i * k + mth] * inv[mth] * V[(size_t)j * k + mth];
  P[(size_t)i * k + j] = acc;
}

// x_out = (x_in ? x_in : 0) + P rhs   (warp per row)
__global__ void __launch_bounds__(256) pinv_apply_kernel(const double* __restrict__ P, int k, const double* __restrict__ rhs,
                                                         int64_t rhs_stride, const double* __restrict__ x_in,
                                                         double* __restrict__ x_out) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= k) return;
  double acc = 0.0;
  for (int c = lane; c < k; c += 32) acc += P[(size_t)row * k + c] * rhs[(size_t)c * rhs_stride];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) x_out[row] = (x_in ? x_in[row] : 0.0) + acc;
}

}  // namespace

size_t fsb_pinv_bytes_impl(int k) { return ((size_t)3 * k * k + k) * sizeof(double); }   // P | work | V | lambda

int fsb_launch_pinv_factor(const fsb_context* h, const double* gaug, int k, double rcond, double shift, void* buf,
                           size_t bytes, int32_t* info, cudaStream_t s) {
  if (bytes < fsb_pinv_bytes_impl(k)) return FSB_ERR_WORKSPACE_TOO_SMALL;
  double* P = (double*)buf;
  double* work = P + (size_t)k * k;
  double* V = work + (size_t)k * k;
  double* lam = V + (size_t)k * k;
  const int m = (k + 1) & ~1;
  const size_t smem = ((size_t)m + 64) * sizeof(double);
  if (m / 2 > 1024) return FSB_ERR_UNSUPPORTED;
  (void)h;
  jacobi_eig_kernel<<<1, 1024, smem, s>>>(gaug, k, work, V, lam, 30, info);
  FSB_LAUNCH_CHECK("jacobi_eig_kernel");
  dim3 grid((unsigned)fsb_ceil_div(k, 256), (unsigned)k);
  pinv_build_kernel<<<grid, 256, (size_t)k * sizeof(double), s>>>(V, lam, k, rcond, shift, P, info);
  FSB_LAUNCH_CHECK("pinv_build_kernel");
  return FSB_OK;
}

int fsb_launch_pinv_apply(const void* buf, int k, const double* rhs, int64_t rhs_stride, const double* x_in,
                          double* x_out, cudaStream_t s) {
  pinv_apply_kernel<<<(unsigned)fsb_ceil_div((int64_t)k * 32, 256), 256, 0, s>>>((const double*)buf, k, rhs, rhs_stride,
                                                                                 x_in, x_out);
  FSB_LAUNCH_CHECK("pinv_apply_kernel");
  return FSB_OK;
}
