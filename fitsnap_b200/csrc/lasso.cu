// LASSO on the reduced problem: covariance-form cyclic coordinate descent on (G, c).
//
// Replaces sklearn.linear_model.Lasso(alpha, fit_intercept=False, max_iter).fit(aw, bw) as called at
// fitsnap3lib/solvers/lasso.py:25-29.  sklearn minimises  1/(2 n) |bw - aw x|^2 + alpha |x|_1
// (n = number of training rows); with G = aw^T aw, c = aw^T bw the coordinate update is
//     x_j <- S(c_j - sum_{m != j} G_jm x_m, n alpha) / G_jj,   S = soft threshold,
// which needs only the k x k Gram produced by gram.cu (one pass over A, one all-reduce), so the
// iteration never touches the design matrix again.  One CTA; the gradient-like vector
// r = c - G x is kept in shared memory and updated with row j of G (coalesced) after each change.
#include "fsb_common.cuh"

namespace {

__global__ void __launch_bounds__(1024) lasso_cd_kernel(const double* __restrict__ gaug, int k, double n_train,
                                                        double alpha, int max_iter, double tol,
                                                        double* __restrict__ x_out, int32_t* info) {
  extern __shared__ double sm[];
  double* x = sm;        // k
  double* r = sm + k;    // k
  __shared__ double s_delta, s_dmax, s_xmax;
  const int ka = k + 1;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < k; i += nt) { x[i] = 0.0; r[i] = gaug[(size_t)i * ka + k]; }
  __syncthreads();
  const double thr = n_train * alpha;
  int sweeps = 0, converged = 0;
  for (int it = 0; it < max_iter; ++it) {
    if (tid == 0) { s_dmax = 0.0; s_xmax = 0.0; }
    __syncthreads();
    for (int j = 0; j < k; ++j) {
      if (tid == 0) {
        const double gjj = gaug[(size_t)j * ka + j];
        const double xo = x[j];
        double xn = 0.0;
        if (gjj > 0.0) {
          const double rho = r[j] + gjj * xo;
          const double mag = fabs(rho) - thr;
          xn = mag > 0.0 ? copysign(mag, rho) / gjj : 0.0;
        }
        x[j] = xn;
        s_delta = xn - xo;
        s_dmax = fmax(s_dmax, fabs(xn - xo));
        s_xmax = fmax(s_xmax, fabs(xn));
      }
      __syncthreads();
      const double d = s_delta;
      if (d != 0.0) {
        const double* grow = gaug + (size_t)j * ka;   // row j == column j (symmetric), contiguous
        for (int i = tid; i < k; i += nt) r[i] -= d * grow[i];
      }
      __syncthreads();
    }
    ++sweeps;
    const double dmax = s_dmax, xmax = s_xmax;   // written before the last barrier of the sweep
    __syncthreads();                              // everyone has read them before thread 0 resets them
    if (xmax == 0.0 || dmax <= tol * xmax) { converged = 1; break; }
  }
  for (int i = tid; i < k; i += nt) x_out[i] = x[i];
  if (tid == 0) { info[0] = converged ? 0 : 1; info[1] = sweeps; }
}

}  // namespace

int fsb_launch_lasso(const fsb_context* h, const double* gaug, int k, double n_train, double alpha, int max_iter,
                     double tol, double* x_out, int32_t* info, cudaStream_t s) {
  const size_t smem = (size_t)2 * k * sizeof(double);
  if (smem > h->smem_optin) return FSB_ERR_UNSUPPORTED;
  FSB_CUDA_TRY(cudaFuncSetAttribute(lasso_cd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = k >= 1024 ? 1024 : (k >= 256 ? 256 : 128);
  lasso_cd_kernel<<<1, threads, smem, s>>>(gaug, k, n_train, alpha, max_iter, tol, x_out, info);
  FSB_LAUNCH_CHECK("lasso_cd_kernel");
  return FSB_OK;
}
