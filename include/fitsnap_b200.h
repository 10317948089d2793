/*
 * fitsnap_b200 -- C-ABI of the B200-native FitSNAP linear-fit hot path.
 *
 * Drop-in boundary for: per-configuration descriptor-row assembly into the design
 * matrix A (+ truth b, weights w) and the weighted least-squares / ridge solve for the
 * coefficient vector.  The reference has no FFI for this path (it is numpy/scipy/sklearn
 * called from Python); each entry point below cites the reference Python interface it
 * replaces (paths relative to the FitSNAP tree).  INTEGRATION.md shows the ctypes
 * binding a FitSNAP maintainer would add.
 *
 * Conventions
 *   - every function returns an fsb_status (0 = OK); nothing aborts or throws;
 *   - all data pointers are DEVICE pointers (fp64 unless stated), owned by the caller;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered and
 *     asynchronous, no entry point synchronises the host;
 *   - workspaces are caller-allocated; query the size with the *_workspace_bytes call;
 *   - A is row-major with leading dimension lda >= k (doubles); rows are
 *     configurations' energy/force/virial rows exactly as FitSNAP's
 *     pt.shared_arrays['a'].array (calculator.py:287).
 */
#ifndef FITSNAP_B200_H
#define FITSNAP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fsb_context* fsb_handle_t;

typedef enum {
  FSB_OK = 0,
  FSB_ERR_INVALID_ARGUMENT = 1,
  FSB_ERR_CUDA = 2,            /* a CUDA runtime call / launch failed; see fsb_last_cuda_error */
  FSB_ERR_WORKSPACE_TOO_SMALL = 3,
  FSB_ERR_UNSUPPORTED = 4,
  FSB_ERR_NO_DEVICE = 5
} fsb_status;

/* flags of fsb_scatter: which row families are assembled ([CALCULATOR] energy/force/stress,
 * io/sections/calculator_sections/calculator.py:13-47) and [BISPECTRUM]/[ACE] bzeroflag. */
enum {
  FSB_ROWS_ENERGY = 1,
  FSB_ROWS_FORCE = 2,
  FSB_ROWS_STRESS = 4,
  FSB_BZEROFLAG = 8,
  FSB_SCRUB_NONFINITE = 16   /* numpy.nan_to_num on the raw block (lammps_pace.py:399-403) */
};

/* info[] written by fsb_factor (device int32[8]) */
enum {
  FSB_INFO_STATUS = 0,        /* 0 = positive definite; 1 = at least one pivot broke down */
  FSB_INFO_FIRST_BAD_COLUMN = 1,
  FSB_INFO_NUM_PINNED = 2,    /* all-zero columns of A pinned to coefficient 0 (lstsq min-norm) */
  FSB_INFO_NUM_DEFICIENT = 3, /* columns dropped because their pivot fell below tolerance */
  FSB_INFO_LEN = 8
};

int fsb_version(void);
const char* fsb_status_string(int status);
/* text of the last CUDA error seen by this library on the calling thread ("" if none) */
const char* fsb_last_cuda_error(void);

/* Context: binds a CUDA device, caches its SM count.  No reference counterpart
 * (the reference's ParallelTools, parallel_tools.py:157, plays the runtime role). */
int fsb_create(fsb_handle_t* out, int device);
int fsb_destroy(fsb_handle_t h);
int fsb_sm_count(fsb_handle_t h, int* sm_count);
/* kernels this library has launched in the calling process so far (a real counter, bumped at every launch site);
 * bench.py's `gpu_launches` is a difference of two readings. */
int fsb_launch_count(fsb_handle_t h, uint64_t* count);

/* ---- K1: row build + scale + scatter ------------------------------------------------
 * Replaces LammpsSnap._collect_lammps (calculators/lammps_snap.py:391-556) and
 * LammpsPace._collect_lammps (calculators/lammps_pace.py:369-509), batched over ncfg
 * configurations whose raw LAMMPS compute blocks are concatenated in `raw`.
 *   raw          sum_c (1+3N_c+6) rows x (kraw+1) cols, row-major, kraw = ncoeff*numtypes;
 *                last column = reference-potential energy/force/virial
 *   raw_row_off  int64[ncfg+1] first raw row of each configuration
 *   out_row_off  int64[ncfg+1] first output row of each configuration
 *   natoms       int32[ncfg]
 *   volume       [ncfg]  lmp.get_thermo("vol")            (lammps_snap.py:406)
 *   energy       [ncfg]  data["Energy"]
 *   forces       [3*sum N_c] data["Forces"].ravel() concatenated
 *   stress       [ncfg*9] data["Stress"] 3x3 row-major; rows use entries
 *                (0,0)(1,1)(2,2)(1,2)(0,2)(0,1)           (lammps_snap.py:541)
 *   eweight/fweight/vweight [ncfg]                         (scrapers/scrape.py:323-353)
 *   type_fraction [ncfg*numtypes] fraction of atoms per type (lammps_snap.py:459-462),
 *                only read when FSB_BZEROFLAG is clear
 *   blank2j      [k] column prefactor, k = kraw + (bzeroflag ? 0 : numtypes)
 *   A,b,w        outputs; rows [out_row_off[0], out_row_off[ncfg]) are written
 *   n_rows_out   out_row_off[ncfg] - out_row_off[0] (host copy; sizes the grid)
 *   row_cfg      int32[n_rows_out] configuration index of every output row, or NULL (the kernel
 *                then binary-searches out_row_off per row; slower)
 *   nonfinite    device int32 counter (or NULL), incremented when a NaN/Inf is met in the
 *                raw values that were read -- the host raises the reference's ValueError
 *                (lammps_snap.py:426-428) from it.  Must be zeroed by the caller.
 * The raw blocks must be packed back to back: raw_row_off[c+1]-raw_row_off[c] = 1+3N_c+6.
 */
int fsb_scatter(fsb_handle_t h, const double* raw, const int64_t* raw_row_off,
                const int64_t* out_row_off, const int32_t* natoms, const double* volume,
                const double* energy, const double* forces, const double* stress,
                const double* eweight, const double* fweight, const double* vweight,
                const double* type_fraction, const double* blank2j, int32_t ncfg,
                int32_t numtypes, int32_t ncoeff, int32_t flags, double* A, int64_t lda,
                double* b, double* w, int64_t n_rows_out, const int32_t* row_cfg, int32_t* nonfinite,
                void* stream);

/* row -> configuration map of a batch (the `row_cfg` argument of fsb_scatter / fsb_scatter_gram, which selects
 * their TMA-staged fast paths): row_cfg[i] = configuration that owns output row out_row_off[0] + i. */
int fsb_row_map(fsb_handle_t h, const int64_t* out_row_off, int32_t ncfg, int32_t* row_cfg, int64_t n_rows_out,
                void* stream);

/* ---- K1 + K2..K4 fused: row build + scatter + Gram in ONE pass over the raw blocks --------------------------
 * The per-batch accumulation of examples/library/transpose_trick/example.py:226-246
 *     a, b, w = calculator.process_single(configuration, i);  aw, bw = w[:, None] * a, w * b
 *     c += aw.T @ aw;  d += aw.T @ bw
 * for a whole batch of configurations: arguments of fsb_scatter + the test mask, output of fsb_gram.  The rows are
 * weighted and contracted straight from the registers that assemble them -- the design matrix is not read back.
 * A may be NULL (streaming mode: A is never materialised; b and w always are).  Results are bit-identical to
 * fsb_scatter followed by fsb_gram.  Covers the common layout only -- all three row families (FSB_ROWS_ENERGY |
 * FORCE | STRESS), row_cfg given, k + 1 <= 104, fp64 Gram path -- and returns FSB_ERR_UNSUPPORTED otherwise (call
 * fsb_scatter + fsb_gram then).  Workspace: fsb_gram_workspace_bytes(h, n_rows_out, k). */
int fsb_scatter_gram(fsb_handle_t h, const double* raw, const int64_t* raw_row_off, const int64_t* out_row_off,
                     const int32_t* natoms, const double* volume, const double* energy, const double* forces,
                     const double* stress, const double* eweight, const double* fweight, const double* vweight,
                     const double* type_fraction, const double* blank2j, int32_t ncfg, int32_t numtypes,
                     int32_t ncoeff, int32_t flags, double* A, int64_t lda, double* b, double* w, int64_t n_rows_out,
                     const int32_t* row_cfg, int32_t* nonfinite, const uint8_t* testing, double* gaug, void* workspace,
                     size_t workspace_bytes, void* stream);

/* ---- K2+K3+K4: fused mask + row weighting + Gram ------------------------------------
 * Replaces the prologue and contraction of SVD/RIDGE/LASSO.perform_fit
 * (solvers/svd.py:35-53, solvers/ridge.py:28-43, solvers/lasso.py:19-24) and the
 * per-configuration accumulation of examples/library/transpose_trick/example.py:226-246.
 *   testing  uint8[n_rows] or NULL; 1 = row excluded (pt.fitsnap_dict['Testing'])
 *   gaug     out, (k+1)x(k+1) row-major symmetric:
 *              [ aw^T aw   aw^T bw ]
 *              [ bw^T aw   bw^T bw ]      aw = w[:,None]*A[train], bw = w*b[train]
 *
 * Two arithmetic paths produce gaug (fsb_set_gram_path; default FSB_GRAM_AUTO picks by shape):
 *   FSB_GRAM_FP64  fp64 tensor-core MMA (DMMA), split-K with a fixed reduction order;
 *   FSB_GRAM_INT8  exact integer Gram on the int8 tcgen05 tensor cores: the weighted rows are
 *                  quantised per column to 53-bit integers, contracted modulo 16 coprime moduli
 *                  (int32 accumulators in tensor memory) and rebuilt by the Chinese remainder
 *                  theorem -- one rounding per entry, independent of summation order.
 * Call fsb_set_gram_path BEFORE fsb_gram_workspace_bytes: the workspace size depends on it.
 */
enum { FSB_GRAM_AUTO = 0, FSB_GRAM_FP64 = 1, FSB_GRAM_INT8 = 2 };
int fsb_set_gram_path(fsb_handle_t h, int32_t path);
/* the path fsb_gram will take for this shape under the current setting (FSB_GRAM_FP64 or _INT8) */
int fsb_get_gram_path(fsb_handle_t h, int64_t n_rows, int32_t k, int32_t* path);
size_t fsb_gram_workspace_bytes(fsb_handle_t h, int64_t n_rows, int32_t k);
int fsb_gram(fsb_handle_t h, const double* A, int64_t lda, const double* b, const double* w,
             const uint8_t* testing, int64_t n_rows, int32_t k, double* gaug,
             void* workspace, size_t workspace_bytes, void* stream);

/* ---- K5: the collective of the row-sharded fit -------------------------------------------
 * Replaces comm.Allreduce([c, MPI.DOUBLE], [c_all, MPI.DOUBLE]) / ([d, ...]) of
 * examples/library/transpose_trick/example.py:241-242 (and the node-shared-array reductions of
 * parallel_tools.py): in-place SUM of `count` doubles over the ranks of a communicator, on the caller's stream.
 * One process per GPU; every rank makes the same sequence of calls with the same counts.
 *
 * Communicator set-up (host side, once):
 *   rank 0: fsb_comm_unique_id(id)  ->  ship the 128 bytes to every rank (MPI_Bcast, a torch.distributed store ...)
 *   all   : fsb_comm_init(h, id, world, rank, &comm)            -- ncclCommInitRank (libnccl.so.2 is dlopen'ed)
 *      or : fsb_comm_adopt(h, existing_ncclComm_t, world, rank, &comm)
 *   optional peer windows for small messages (ranks of ONE box, NVLink / NVSwitch):
 *   all   : fsb_comm_peer_export(comm, handle)  ->  all-gather the handles (fsb_comm_peer_handle_bytes() each)
 *   all   : fsb_comm_peer_attach(comm, handles_in_rank_order, world)   -- cudaIpcOpenMemHandle
 *           (if it fails on ANY rank, every rank must call fsb_comm_peer_disable)
 * fsb_allreduce then sends messages of at most fsb_comm_peer_max_bytes() through ONE kernel that stages the vector
 * in this rank's window, publishes an epoch flag, waits for every rank's flag and adds all windows in rank order over
 * peer loads (bit-identical sums on every rank; CUDA-graph capturable; no communicator stream), and larger messages
 * through ncclAllReduce(ncclDouble, ncclSum).  Ranks must enter a peer-window call within FSB_PEER_TIMEOUT_S
 * (default 120 s) of each other, otherwise the waiting kernel traps instead of hanging the device.
 * fsb_comm_info: out4 = {peer windows ready, peer-window calls, NCCL calls, world}.
 */
typedef struct fsb_comm* fsb_comm_t;
int fsb_comm_unique_id(void* id, size_t id_bytes);   /* id_bytes >= 128 */
int fsb_comm_init(fsb_handle_t h, const void* id, int world, int rank, fsb_comm_t* out);
int fsb_comm_adopt(fsb_handle_t h, void* nccl_comm, int world, int rank, fsb_comm_t* out);
size_t fsb_comm_peer_handle_bytes(void);
size_t fsb_comm_peer_max_bytes(void);
int fsb_comm_peer_export(fsb_comm_t c, void* handle, size_t handle_bytes);
int fsb_comm_peer_attach(fsb_comm_t c, const void* handles, int n);
int fsb_comm_peer_disable(fsb_comm_t c);
int fsb_comm_info(fsb_comm_t c, int64_t* out4);
int fsb_allreduce(fsb_handle_t h, fsb_comm_t c, double* buf, int64_t count, void* stream);
int fsb_comm_destroy(fsb_comm_t c);

/* ---- K6: factor + solve -------------------------------------------------------------
 * Replaces scipy.linalg.lstsq(aw,bw,1e-13) as called at solvers/svd.py:54 (alpha = 0) and
 * sklearn Ridge(alpha, fit_intercept=False).fit / Local_Ridge.fit as called at
 * solvers/ridge.py:49-57, lib/ridge_solver/regressor.py:10-16 (alpha > 0), on the
 * reduced system.  fsb_factor equilibrates S = D (G + alpha I) D, D = diag^-1/2, pins
 * all-zero columns, and Cholesky-factors S into `factor`; fsb_factor_solve applies
 *   x_out = (x_in ? x_in : 0) + D S^-1 D (rhs - alpha * (x_in ? x_in : 0)).
 * With rhs = aw^T bw, x_in = NULL this is the normal-equation solution; with
 * rhs = aw^T (bw - aw x) from fsb_residual it is one step of iterative refinement.
 */
size_t fsb_factor_bytes(fsb_handle_t h, int32_t k);
int fsb_factor(fsb_handle_t h, const double* gaug, int32_t k, double alpha, void* factor,
               size_t factor_bytes, int32_t* info, void* stream);
int fsb_factor_solve(fsb_handle_t h, const void* factor, int32_t k, const double* rhs,
                     int64_t rhs_stride, double alpha, const double* x_in, double* x_out,
                     void* stream);

/* ---- minimum-norm fallback (rank-deficient systems) -----------------------------------
 * scipy.linalg.lstsq(aw, bw, 1e-13) (solvers/svd.py:54, gelsd) returns the MINIMUM-NORM solution
 * when columns are linearly dependent; fsb_factor reports such a system in info[] (status 1).
 * fsb_pinv_factor builds G^+ = V diag(1/lambda_i | lambda_i > rcond * lambda_max) V^T from a Jacobi
 * eigendecomposition of the (un-equilibrated) Gram in `gaug`; fsb_pinv_apply does
 *   x_out = (x_in ? x_in : 0) + G^+ rhs,
 * i.e. the initial solve (rhs = aw^T bw) and each refinement step (rhs = fsb_residual output).
 * info[0] = numerical rank, info[1] = Jacobi sweeps (device int32[2]).  rcond acts on eigenvalues
 * of G (squared singular values); k * 2.2e-16 is the resolution of a Gram formed in fp64.
 */
size_t fsb_pinv_bytes(fsb_handle_t h, int32_t k);
int fsb_pinv_factor(fsb_handle_t h, const double* gaug, int32_t k, double rcond, void* pinv,
                    size_t pinv_bytes, int32_t* info, void* stream);
/* ridge on a numerically rank-deficient system (sklearn's Ridge falls back from Cholesky to an SVD solve there,
 * solvers/ridge.py:49-57): P = V diag(1/(lambda_i + alpha) | lambda_i > rcond * lambda_max) V^T -- the exact ridge
 * solution has no component in the null space of G, so those directions are dropped rather than divided by alpha.
 * Refinement: x_out = x_in + P (fsb_residual output - alpha * x_in). */
int fsb_pinv_factor_shifted(fsb_handle_t h, const double* gaug, int32_t k, double rcond, double alpha, void* pinv,
                            size_t pinv_bytes, int32_t* info, void* stream);
int fsb_pinv_apply(fsb_handle_t h, const void* pinv, int32_t k, const double* rhs, int64_t rhs_stride,
                   const double* x_in, double* x_out, void* stream);

/* ---- LASSO on the reduced problem ----------------------------------------------------
 * Replaces sklearn Lasso(alpha, fit_intercept=False, max_iter).fit(aw, bw) as called at
 * solvers/lasso.py:25-29: argmin 1/(2 n_train) |bw - aw x|^2 + alpha |x|_1, by cyclic coordinate
 * descent on (G, c) taken from `gaug` (fsb_gram output).  Stops when the largest coordinate
 * change of a sweep is <= tol * max|x| or after max_iter sweeps.
 * info[0] = 0 converged / 1 hit max_iter, info[1] = sweeps done (device int32[2]).
 */
int fsb_lasso(fsb_handle_t h, const double* gaug, int32_t k, int64_t n_train, double alpha,
              int32_t max_iter, double tol, double* x_out, int32_t* info, void* stream);

/* ---- K7: residual / prediction pass --------------------------------------------------
 * fsb_residual: g = aw^T (bw - aw x) in one streaming pass over A (refinement residual;
 *   the reference computes `aw @ coef - bw` at solvers/ridge.py:60).
 * fsb_predict:  y = A x for ALL rows (solvers/solver.py:377 `a @ self.fit`).
 */
size_t fsb_residual_workspace_bytes(fsb_handle_t h, int64_t n_rows, int32_t k);
int fsb_residual(fsb_handle_t h, const double* A, int64_t lda, const double* b,
                 const double* w, const uint8_t* testing, int64_t n_rows, int32_t k,
                 const double* x, double* g, void* workspace, size_t workspace_bytes,
                 void* stream);
int fsb_predict(fsb_handle_t h, const double* A, int64_t lda, int64_t n_rows, int32_t k,
                const double* x, double* y, void* stream);

/* ---- error analysis -------------------------------------------------------------------
 * One pass over A giving, per group id, the sums from which Solver.error_analysis builds its
 * MAE / RMSE / R^2 table (solvers/solver.py:108-133, 368-429): with pred = a . x, res = b - pred,
 *   stats[g*10 + 0..9] += n, sum|res|, sum res^2, sum t, sum t^2,
 *                         n(w != 0), sum|w res|, sum (w res)^2, sum w t, sum (w t)^2.
 * group_id: int32[n_rows] in [0, n_groups); the caller encodes (Groups, Testing, Row_Type) into it.
 * `stats` (n_groups*10 doubles) must be zeroed by the caller; accumulation uses atomics.
 */
int fsb_group_stats(fsb_handle_t h, const double* A, int64_t lda, const double* b, const double* w,
                    const int32_t* group_id, int64_t n_rows, int32_t k, const double* x,
                    int32_t n_groups, double* stats, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FITSNAP_B200_H */
