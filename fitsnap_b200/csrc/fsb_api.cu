// C-ABI entry points of libfitsnap_b200.so (declared in include/fitsnap_b200.h).
// Argument validation + dispatch to the kernel launchers; never throws, never syncs.
#include "fsb_common.cuh"
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include <new>

// launchers (gram.cu, solve.cu, stream_ops.cu)
size_t fsb_factor_bytes_impl(int k);
int fsb_launch_factor(const fsb_context* h, const double* gaug, int k, double alpha, void* factor,
                      size_t factor_bytes, int32_t* info, cudaStream_t s);
int fsb_launch_factor_solve(const fsb_context* h, const void* factor, int k, const double* rhs,
                            int64_t rhs_stride, double alpha, const double* x_in, double* x_out,
                            cudaStream_t s);
size_t fsb_residual_ws_bytes(const fsb_context* h, int64_t n_rows, int k);
int fsb_launch_residual(const fsb_context* h, const double* A, int64_t lda, const double* b, const double* w,
                        const uint8_t* testing, int64_t n_rows, int k, const double* x, double* g, void* ws,
                        size_t ws_bytes, cudaStream_t s);
int fsb_launch_predict(const fsb_context* h, const double* A, int64_t lda, int64_t n_rows, int k,
                       const double* x, double* y, cudaStream_t s);
int fsb_launch_scatter(const fsb_context* h, const double* raw, const int64_t* raw_row_off,
                       const int64_t* out_row_off, const int32_t* natoms, const double* volume,
                       const double* energy, const double* forces, const double* stress,
                       const double* eweight, const double* fweight, const double* vweight,
                       const double* type_fraction, const double* blank2j, int ncfg, int numtypes,
                       int ncoeff, int flags, double* A, int64_t lda, double* b, double* w,
                       int32_t* nonfinite, const int32_t* row_cfg, int64_t n_rows_hint, cudaStream_t s);

int fsb_launch_scatter_gram(const fsb_context* h, const ScatterArgs& sc, const uint8_t* testing, int64_t total,
                            int store_a, double* gaug, void* ws, size_t ws_bytes, cudaStream_t s);
int fsb_launch_row_map(const int64_t* out_row_off, int ncfg, int32_t* row_cfg, int64_t n_rows, cudaStream_t s);
bool fsb_gram_i8_available();
int fsb_gram_path_for(const fsb_context* h, int64_t n_rows, int k);

int fsb_launch_lasso(const fsb_context* h, const double* gaug, int k, double n_train, double alpha, int max_iter,
                     double tol, double* x_out, int32_t* info, cudaStream_t s);

int fsb_launch_group_stats(const fsb_context* h, const double* A, int64_t lda, const double* b, const double* w,
                           const int32_t* gid, int64_t n_rows, int k, const double* x, int n_groups, double* stats,
                           cudaStream_t s);

size_t fsb_pinv_bytes_impl(int k);
int fsb_launch_pinv_factor(const fsb_context* h, const double* gaug, int k, double rcond, double shift, void* buf,
                           size_t bytes, int32_t* info, cudaStream_t s);
int fsb_launch_pinv_apply(const void* buf, int k, const double* rhs, int64_t rhs_stride, const double* x_in,
                          double* x_out, cudaStream_t s);

static thread_local char g_cuda_err[512] = "";

void fsb_note_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
}

void fsb_note_text(const char* where, const char* detail) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", where, detail ? detail : "");
}

unsigned long long g_fsb_launches = 0;

static const int FSB_MAX_K = 2047;  // rowpass kernels hold ceil(k/32) <= 64 values per lane

extern "C" {

int fsb_version(void) { return 100; }

const char* fsb_status_string(int status) {
  switch (status) {
    case FSB_OK: return "ok";
    case FSB_ERR_INVALID_ARGUMENT: return "invalid argument";
    case FSB_ERR_CUDA: return "CUDA error";
    case FSB_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
    case FSB_ERR_UNSUPPORTED: return "unsupported size or configuration";
    case FSB_ERR_NO_DEVICE: return "no usable CUDA device";
    default: return "unknown status";
  }
}

const char* fsb_last_cuda_error(void) { return g_cuda_err; }

int fsb_create(fsb_handle_t* out, int device) {
  if (!out) return FSB_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    if (e != cudaSuccess) fsb_note_cuda_error(e, "cudaGetDeviceCount");
    return FSB_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= count) return FSB_ERR_INVALID_ARGUMENT;
  FSB_CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  FSB_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "device %d is sm_%d%d; this library is built for sm_100a only",
             device, prop.major, prop.minor);
    return FSB_ERR_UNSUPPORTED;
  }
  fsb_context* h = new (std::nothrow) fsb_context;
  if (!h) return FSB_ERR_INVALID_ARGUMENT;
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->smem_optin = prop.sharedMemPerBlockOptin;
  h->gram_path = FSB_GRAM_AUTO;
  if (const char* e = getenv("FSB_GRAM_PATH")) {
    if (!strcmp(e, "fp64")) h->gram_path = FSB_GRAM_FP64;
    else if (!strcmp(e, "int8")) h->gram_path = FSB_GRAM_INT8;
  }
  *out = h;
  return FSB_OK;
}

int fsb_destroy(fsb_handle_t h) {
  delete h;
  return FSB_OK;
}

int fsb_launch_count(fsb_handle_t h, uint64_t* count) {
  if (!h || !count) return FSB_ERR_INVALID_ARGUMENT;
  *count = (uint64_t)g_fsb_launches;
  return FSB_OK;
}

int fsb_sm_count(fsb_handle_t h, int* sm_count) {
  if (!h || !sm_count) return FSB_ERR_INVALID_ARGUMENT;
  *sm_count = h->sm_count;
  return FSB_OK;
}

int fsb_scatter(fsb_handle_t h, const double* raw, const int64_t* raw_row_off, const int64_t* out_row_off,
                const int32_t* natoms, const double* volume, const double* energy, const double* forces,
                const double* stress, const double* eweight, const double* fweight, const double* vweight,
                const double* type_fraction, const double* blank2j, int32_t ncfg, int32_t numtypes,
                int32_t ncoeff, int32_t flags, double* A, int64_t lda, double* b, double* w,
                int64_t n_rows_out, const int32_t* row_cfg, int32_t* nonfinite, void* stream) {
  if (!h || ncfg < 0 || numtypes < 1 || ncoeff < 1 || n_rows_out < 0) return FSB_ERR_INVALID_ARGUMENT;
  if (ncfg == 0 || n_rows_out == 0) return FSB_OK;
  if (!raw || !raw_row_off || !out_row_off || !natoms || !blank2j || !A || !b || !w)
    return FSB_ERR_INVALID_ARGUMENT;
  const bool bzero = flags & FSB_BZEROFLAG;
  const int k = ncoeff * numtypes + (bzero ? 0 : numtypes);
  if (lda < k) return FSB_ERR_INVALID_ARGUMENT;
  if ((flags & FSB_ROWS_ENERGY) && (!energy || !eweight || (!bzero && !type_fraction)))
    return FSB_ERR_INVALID_ARGUMENT;
  if ((flags & FSB_ROWS_FORCE) && (!forces || !fweight)) return FSB_ERR_INVALID_ARGUMENT;
  if ((flags & FSB_ROWS_STRESS) && (!stress || !vweight || !volume)) return FSB_ERR_INVALID_ARGUMENT;
  return fsb_launch_scatter(h, raw, raw_row_off, out_row_off, natoms, volume, energy, forces, stress, eweight,
                            fweight, vweight, type_fraction, blank2j, ncfg, numtypes, ncoeff, flags, A, lda, b,
                            w, nonfinite, row_cfg, n_rows_out, (cudaStream_t)stream);
}

int fsb_row_map(fsb_handle_t h, const int64_t* out_row_off, int32_t ncfg, int32_t* row_cfg, int64_t n_rows_out,
                void* stream) {
  if (!h || ncfg < 0 || n_rows_out < 0) return FSB_ERR_INVALID_ARGUMENT;
  if (ncfg == 0 || n_rows_out == 0) return FSB_OK;
  if (!out_row_off || !row_cfg) return FSB_ERR_INVALID_ARGUMENT;
  return fsb_launch_row_map(out_row_off, ncfg, row_cfg, n_rows_out, (cudaStream_t)stream);
}

int fsb_scatter_gram(fsb_handle_t h, const double* raw, const int64_t* raw_row_off, const int64_t* out_row_off,
                     const int32_t* natoms, const double* volume, const double* energy, const double* forces,
                     const double* stress, const double* eweight, const double* fweight, const double* vweight,
                     const double* type_fraction, const double* blank2j, int32_t ncfg, int32_t numtypes,
                     int32_t ncoeff, int32_t flags, double* A, int64_t lda, double* b, double* w, int64_t n_rows_out,
                     const int32_t* row_cfg, int32_t* nonfinite, const uint8_t* testing, double* gaug, void* workspace,
                     size_t workspace_bytes, void* stream) {
  if (!h || ncfg < 1 || numtypes < 1 || ncoeff < 1 || n_rows_out < 1 || !gaug || !workspace)
    return FSB_ERR_INVALID_ARGUMENT;
  if (!raw || !raw_row_off || !out_row_off || !natoms || !blank2j || !b || !w || !row_cfg)
    return FSB_ERR_INVALID_ARGUMENT;
  const bool bzero = flags & FSB_BZEROFLAG;
  const int k = ncoeff * numtypes + (bzero ? 0 : numtypes);
  if (A && lda < k) return FSB_ERR_INVALID_ARGUMENT;
  if (!energy || !eweight || (!bzero && !type_fraction) || !forces || !fweight || !stress || !vweight || !volume)
    return FSB_ERR_INVALID_ARGUMENT;
  ScatterArgs a;
  a.raw = raw; a.raw_row_off = raw_row_off; a.out_row_off = out_row_off; a.natoms = natoms;
  a.volume = volume; a.energy = energy; a.forces = forces; a.stress = stress;
  a.eweight = eweight; a.fweight = fweight; a.vweight = vweight; a.type_fraction = type_fraction;
  a.blank2j = blank2j; a.ncfg = ncfg; a.numtypes = numtypes; a.ncoeff = ncoeff; a.flags = flags;
  a.A = A; a.lda = A ? lda : k; a.b = b; a.w = w; a.nonfinite = nonfinite; a.row_cfg = row_cfg;
  return fsb_launch_scatter_gram(h, a, testing, n_rows_out, A ? 1 : 0, gaug, workspace, workspace_bytes,
                                 (cudaStream_t)stream);
}

int fsb_set_gram_path(fsb_handle_t h, int32_t path) {
  if (!h || path < FSB_GRAM_AUTO || path > FSB_GRAM_INT8) return FSB_ERR_INVALID_ARGUMENT;
  if (path == FSB_GRAM_INT8 && !fsb_gram_i8_available()) return FSB_ERR_UNSUPPORTED;
  h->gram_path = path;
  return FSB_OK;
}

int fsb_get_gram_path(fsb_handle_t h, int64_t n_rows, int32_t k, int32_t* path) {
  if (!h || !path || k < 1) return FSB_ERR_INVALID_ARGUMENT;
  *path = fsb_gram_path_for(h, n_rows, k);
  return FSB_OK;
}

size_t fsb_gram_workspace_bytes(fsb_handle_t h, int64_t n_rows, int32_t k) {
  if (!h || k < 1 || n_rows < 0) return 0;
  return fsb_gram_ws_bytes(h, n_rows, k);
}

int fsb_gram(fsb_handle_t h, const double* A, int64_t lda, const double* b, const double* w,
             const uint8_t* testing, int64_t n_rows, int32_t k, double* gaug, void* workspace,
             size_t workspace_bytes, void* stream) {
  if (!h || k < 1 || k > FSB_MAX_K || n_rows < 0 || lda < k || !gaug || !workspace)
    return FSB_ERR_INVALID_ARGUMENT;
  if (n_rows > 0 && (!A || !b || !w)) return FSB_ERR_INVALID_ARGUMENT;
  return fsb_launch_gram(h, A, lda, b, w, testing, n_rows, k, gaug, workspace, workspace_bytes,
                         (cudaStream_t)stream);
}

size_t fsb_factor_bytes(fsb_handle_t h, int32_t k) {
  if (!h || k < 1) return 0;
  return fsb_factor_bytes_impl(k);
}

int fsb_factor(fsb_handle_t h, const double* gaug, int32_t k, double alpha, void* factor,
               size_t factor_bytes, int32_t* info, void* stream) {
  if (!h || !gaug || k < 1 || k > FSB_MAX_K || !factor || !info || !(alpha >= 0.0))
    return FSB_ERR_INVALID_ARGUMENT;
  return fsb_launch_factor(h, gaug, k, alpha, factor, factor_bytes, info, (cudaStream_t)stream);
}

int fsb_factor_solve(fsb_handle_t h, const void* factor, int32_t k, const double* rhs, int64_t rhs_stride,
                     double alpha, const double* x_in, double* x_out, void* stream) {
  if (!h || !factor || k < 1 || k > FSB_MAX_K || !rhs || rhs_stride < 1 || !x_out)
    return FSB_ERR_INVALID_ARGUMENT;
  return fsb_launch_factor_solve(h, factor, k, rhs, rhs_stride, alpha, x_in, x_out, (cudaStream_t)stream);
}

size_t fsb_pinv_bytes(fsb_handle_t h, int32_t k) {
  if (!h || k < 1) return 0;
  return fsb_pinv_bytes_impl(k);
}

int fsb_pinv_factor(fsb_handle_t h, const double* gaug, int32_t k, double rcond, void* pinv, size_t pinv_bytes,
                    int32_t* info, void* stream) {
  if (!h || !gaug || k < 1 || k > FSB_MAX_K || !pinv || !info || !(rcond >= 0.0)) return FSB_ERR_INVALID_ARGUMENT;
  return fsb_launch_pinv_factor(h, gaug, k, rcond, 0.0, pinv, pinv_bytes, info, (cudaStream_t)stream);
}

int fsb_pinv_factor_shifted(fsb_handle_t h, const double* gaug, int32_t k, double rcond, double alpha, void* pinv,
                            size_t pinv_bytes, int32_t* info, void* stream) {
  if (!h || !gaug || k < 1 || k > FSB_MAX_K || !pinv || !info || !(rcond >= 0.0) || !(alpha >= 0.0))
    return FSB_ERR_INVALID_ARGUMENT;
  return fsb_launch_pinv_factor(h, gaug, k, rcond, alpha, pinv, pinv_bytes, info, (cudaStream_t)stream);
}

int fsb_pinv_apply(fsb_handle_t h, const void* pinv, int32_t k, const double* rhs, int64_t rhs_stride,
                   const double* x_in, double* x_out, void* stream) {
  if (!h || !pinv || k < 1 || k > FSB_MAX_K || !rhs || rhs_stride < 1 || !x_out) return FSB_ERR_INVALID_ARGUMENT;
  return fsb_launch_pinv_apply(pinv, k, rhs, rhs_stride, x_in, x_out, (cudaStream_t)stream);
}

int fsb_lasso(fsb_handle_t h, const double* gaug, int32_t k, int64_t n_train, double alpha, int32_t max_iter,
              double tol, double* x_out, int32_t* info, void* stream) {
  if (!h || !gaug || k < 1 || k > FSB_MAX_K || n_train < 0 || !(alpha >= 0.0) || max_iter < 1 || !x_out || !info)
    return FSB_ERR_INVALID_ARGUMENT;
  return fsb_launch_lasso(h, gaug, k, (double)n_train, alpha, max_iter, tol, x_out, info, (cudaStream_t)stream);
}

size_t fsb_residual_workspace_bytes(fsb_handle_t h, int64_t n_rows, int32_t k) {
  if (!h || k < 1 || n_rows < 0) return 0;
  return fsb_residual_ws_bytes(h, n_rows, k);
}

int fsb_residual(fsb_handle_t h, const double* A, int64_t lda, const double* b, const double* w,
                 const uint8_t* testing, int64_t n_rows, int32_t k, const double* x, double* g,
                 void* workspace, size_t workspace_bytes, void* stream) {
  if (!h || k < 1 || k > FSB_MAX_K || n_rows < 0 || lda < k || !x || !g || !workspace)
    return FSB_ERR_INVALID_ARGUMENT;
  if (n_rows > 0 && (!A || !b || !w)) return FSB_ERR_INVALID_ARGUMENT;
  return fsb_launch_residual(h, A, lda, b, w, testing, n_rows, k, x, g, workspace, workspace_bytes,
                             (cudaStream_t)stream);
}

int fsb_group_stats(fsb_handle_t h, const double* A, int64_t lda, const double* b, const double* w,
                    const int32_t* group_id, int64_t n_rows, int32_t k, const double* x, int32_t n_groups,
                    double* stats, void* stream) {
  if (!h || k < 1 || k > FSB_MAX_K || n_rows < 0 || lda < k || !x || !stats || n_groups < 1)
    return FSB_ERR_INVALID_ARGUMENT;
  if (n_rows > 0 && (!A || !b || !w || !group_id)) return FSB_ERR_INVALID_ARGUMENT;
  return fsb_launch_group_stats(h, A, lda, b, w, group_id, n_rows, k, x, n_groups, stats, (cudaStream_t)stream);
}

int fsb_predict(fsb_handle_t h, const double* A, int64_t lda, int64_t n_rows, int32_t k, const double* x,
                double* y, void* stream) {
  if (!h || k < 1 || k > FSB_MAX_K || n_rows < 0 || lda < k || !x) return FSB_ERR_INVALID_ARGUMENT;
  if (n_rows > 0 && (!A || !y)) return FSB_ERR_INVALID_ARGUMENT;
  return fsb_launch_predict(h, A, lda, n_rows, k, x, y, (cudaStream_t)stream);
}

}  // extern "C"
