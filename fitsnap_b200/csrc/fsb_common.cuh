// Shared declarations of the fitsnap_b200 native library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/fitsnap_b200.h"

struct fsb_context {
  int device;
  int sm_count;
  size_t smem_optin;   // max dynamic shared memory per block (opt-in)
  int gram_path;       // FSB_GRAM_AUTO / FSB_GRAM_FP64 / FSB_GRAM_INT8 (fsb_set_gram_path)
};

// thread-local last CUDA error text (fsb_last_cuda_error)
void fsb_note_cuda_error(cudaError_t e, const char* where);
void fsb_note_text(const char* where, const char* detail);
// kernels launched by this library in this process (fsb_launch_count): bumped once per successful launch
extern unsigned long long g_fsb_launches;

#define FSB_CUDA_TRY(expr)                                   \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) {                                 \
      fsb_note_cuda_error(_e, #expr);                        \
      return FSB_ERR_CUDA;                                   \
    }                                                        \
  } while (0)

#define FSB_LAUNCH_CHECK(where)                              \
  do {                                                       \
    cudaError_t _e = cudaGetLastError();                     \
    if (_e != cudaSuccess) {                                 \
      fsb_note_cuda_error(_e, where);                        \
      return FSB_ERR_CUDA;                                   \
    }                                                        \
    ++g_fsb_launches;                                        \
  } while (0)

static inline int64_t fsb_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t fsb_round_up(int64_t a, int64_t b) { return fsb_ceil_div(a, b) * b; }

// ---- geometry shared by host planners and kernels --------------------------------------
// Gram: output is tiled in GT x GT super-tiles of the augmented (k+1)x(k+1) matrix; only
// lower-triangular super-tiles are computed.
constexpr int FSB_GT = 128;          // super-tile edge (columns of A)
constexpr int FSB_GRCH = 32;         // rows of A per shared-memory pipeline stage
constexpr int FSB_GLDS = FSB_GT + 4; // smem row pitch in doubles: == 4 (mod 16) -> conflict-free DMMA fragment loads
constexpr int FSB_GTHREADS = 512;    // 16 warps, each owning a 32x32 sub-tile (4x4 DMMA 8x8 blocks)

// FSB_GRAM_AUTO switches to the int8 tcgen05 Gram from this shape on (gram.cu: fsb_gram_path_for)
// (measured crossover against the DMMA path: ~250 columns; 1.7x at 480, 2.3x at 1000 columns)
constexpr int FSB_I8_AUTO_MIN_COLS = 384;
constexpr int64_t FSB_I8_AUTO_MIN_ROWS = 65536;   // 30000 x 1000 measured 0.9x, 1e6 x 1000 2.3x

// Cholesky panel width
constexpr int FSB_NB = 64;

// launchers implemented in the individual .cu files
int fsb_launch_gram(const fsb_context* h, const double* A, int64_t lda, const double* b, const double* w,
                    const uint8_t* testing, int64_t n_rows, int k, double* gaug, void* ws, size_t ws_bytes,
                    cudaStream_t s);
size_t fsb_gram_ws_bytes(const fsb_context* h, int64_t n_rows, int k);

// ---- K1 arguments (stream_ops.cu: scatter kernels; gram_small.cu: fused scatter + Gram) -----------------
struct ScatterArgs {
  const double* raw;
  const int64_t* raw_row_off;
  const int64_t* out_row_off;
  const int32_t* natoms;
  const double* volume;
  const double* energy;
  const double* forces;
  const double* stress;
  const double* eweight;
  const double* fweight;
  const double* vweight;
  const double* type_fraction;
  const double* blank2j;
  int ncfg, numtypes, ncoeff, flags;
  double* A;
  int64_t lda;
  double* b;
  double* w;
  int32_t* nonfinite;
  const int32_t* row_cfg;   // optional: configuration index of every output row
};

constexpr double FSB_VIRIAL_UNIT = 1.6021765e6;  // lammps_snap.py:526

#ifdef __CUDACC__
namespace fsb_dev {
__device__ __forceinline__ double scrub(double v, bool do_scrub, bool& bad) {
  if (!isfinite(v)) {
    bad = true;
    if (do_scrub) {  // numpy.nan_to_num defaults (lammps_pace.py:401)
      if (isnan(v)) return 0.0;
      return v > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
    }
  }
  return v;
}
}  // namespace fsb_dev
#endif
