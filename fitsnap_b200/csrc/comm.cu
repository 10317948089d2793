// K5: the one collective of the linear-fit path -- sum of the packed Gram [(k+1)^2 doubles] and of the k-vector of
// every refinement round over the row shards (one per GPU).  Mirrors `comm.Allreduce([c, MPI.DOUBLE], ...)` of
// examples/library/transpose_trick/example.py:241-242 and the node-shared-array reductions of
// fitsnap3lib/parallel_tools.py.
//
// Two transports behind one call (fsb_allreduce):
//   * peer window (small messages, <= FSB_PEER_MAX_BYTES): every rank owns a device window that all other ranks of
//     the box map through CUDA IPC (NVLink / NVSwitch peer memory).  ONE kernel per call: stage my vector into my
//     window, publish an epoch flag (release.sys), wait for the flag of every rank (acquire.sys), read all windows
//     with plain loads over NVLink and add them IN RANK ORDER -- every rank computes bit-identical sums, so the
//     replicated solve stays replicated.  No host involvement, no communicator stream: the kernel is an ordinary
//     node of the caller's stream and is CUDA-graph capturable.  The messages of this path are 80 KB (k = 100) or
//     less: latency-bound, which is where a one-shot exchange beats a ring (NCCL ~37 us per call measured in round 1).
//   * NCCL (large messages, or no peer window): ncclAllReduce(ncclDouble, ncclSum) on the caller's stream.  libnccl is
//     resolved at run time with dlopen -- the library links neither NCCL nor torch; inside a torch process the
//     already-loaded libnccl.so.2 is picked up.
#include "fsb_common.cuh"
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>

namespace {

// ---- NCCL through dlopen ---------------------------------------------------------------------
struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
typedef int (*FnGetUniqueId)(NcclUniqueId*);
typedef int (*FnCommInitRank)(NcclComm*, int, NcclUniqueId, int);
typedef int (*FnCommDestroy)(NcclComm);
typedef int (*FnAllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef const char* (*FnGetErrorString)(int);
constexpr int kNcclFloat64 = 8;   // ncclDataType_t ncclDouble
constexpr int kNcclSum = 0;       // ncclRedOp_t ncclSum

struct NcclApi {
  void* lib = nullptr;
  FnGetUniqueId get_unique_id = nullptr;
  FnCommInitRank comm_init_rank = nullptr;
  FnCommDestroy comm_destroy = nullptr;
  FnAllReduce all_reduce = nullptr;
  FnGetErrorString error_string = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {getenv("FSB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n) continue;
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (api.lib) {
      api.get_unique_id = (FnGetUniqueId)dlsym(api.lib, "ncclGetUniqueId");
      api.comm_init_rank = (FnCommInitRank)dlsym(api.lib, "ncclCommInitRank");
      api.comm_destroy = (FnCommDestroy)dlsym(api.lib, "ncclCommDestroy");
      api.all_reduce = (FnAllReduce)dlsym(api.lib, "ncclAllReduce");
      api.error_string = (FnGetErrorString)dlsym(api.lib, "ncclGetErrorString");
      api.ok = api.get_unique_id && api.comm_init_rank && api.comm_destroy && api.all_reduce;
    }
  }
  return api;
}

void note(const char* what, const char* detail) { fsb_note_text(what, detail); }

// ---- peer window -----------------------------------------------------------------------------
constexpr int PEER_MAX_RANKS = 16;
constexpr size_t PEER_CTL_BYTES = 128;
constexpr size_t PEER_MAX_BYTES = 1u << 20;                       // per staging buffer (131072 doubles)
constexpr size_t PEER_WINDOW_BYTES = PEER_CTL_BYTES + 2 * PEER_MAX_BYTES;

struct PeerCtl {
  unsigned long long flag;     // epoch whose data is complete in this rank's window (written release.sys)
  unsigned long long epoch;    // epoch of the last finished call (read by the next launch, device-side state so that
                               // a captured graph replays correctly)
  unsigned int arrive;         // CTAs of this rank that have staged their part
  unsigned int depart;         // CTAs of this rank that have finished reading
};

struct PeerArgs {
  char* win[PEER_MAX_RANKS];
  int world, rank;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_peer(const double* p) {   // never served from a stale L1 line
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) peer_allreduce_kernel(PeerArgs p, double* __restrict__ buf, int64_t count,
                                                             long long timeout_cycles) {
  PeerCtl* my = reinterpret_cast<PeerCtl*>(p.win[p.rank]);
  const unsigned long long e = *reinterpret_cast<volatile unsigned long long*>(&my->epoch) + 1ull;
  const size_t stage_off = PEER_CTL_BYTES + (size_t)(e & 1ull) * PEER_MAX_BYTES;
  double* stage = reinterpret_cast<double*>(p.win[p.rank] + stage_off);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;

  // 1. my contribution -> my window; the last CTA to finish publishes the epoch
  for (int64_t i = tid; i < count; i += nthr) stage[i] = buf[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned old = atomicAdd(&my->arrive, 1u);
    if (old == gridDim.x - 1) {
      my->arrive = 0u;
      __threadfence_system();
      st_release_sys(&my->flag, e);
    }
  }
  // 2. wait until every rank (this one included) has published epoch e.  Safe to reuse the buffer of epoch e - 2:
  //    a rank publishes e - 1 only after it has finished reading e - 2, and this rank has seen every e - 1 flag.
  if (threadIdx.x < p.world) {
    const unsigned long long* f = &reinterpret_cast<const PeerCtl*>(p.win[threadIdx.x])->flag;
    const long long t0 = clock64();
    while (ld_acquire_sys(f) < e) {
      if (clock64() - t0 > timeout_cycles) __trap();   // a rank never arrived: fail the launch instead of hanging the GPU
    }
  }
  __syncthreads();
  // 3. sum in rank order (identical on every rank)
  for (int64_t i = tid; i < count; i += nthr) {
    double s = 0.0;
    for (int r = 0; r < p.world; ++r) s += ld_peer(reinterpret_cast<const double*>(p.win[r] + stage_off) + i);
    buf[i] = s;
  }
  // 4. the last CTA to leave advances the epoch for the next launch
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned old = atomicAdd(&my->depart, 1u);
    if (old == gridDim.x - 1) {
      my->depart = 0u;
      *reinterpret_cast<volatile unsigned long long*>(&my->epoch) = e;
      __threadfence();
    }
  }
}

}  // namespace

struct fsb_comm {
  int world = 1, rank = 0, device = 0;
  NcclComm nccl_comm = nullptr;
  bool owns_nccl = false;
  char* window = nullptr;                 // this rank's peer window (cudaMalloc)
  char* peer[PEER_MAX_RANKS] = {nullptr}; // mapped windows, rank order (own entry = window)
  bool peer_ready = false;
  long long timeout_cycles = 0;
  long long peer_calls = 0, nccl_calls = 0;
};

extern "C" {

int fsb_comm_unique_id(void* id, size_t id_bytes) {
  if (!id || id_bytes < sizeof(NcclUniqueId)) return FSB_ERR_INVALID_ARGUMENT;
  NcclApi& api = nccl();
  if (!api.ok) {
    note("fsb_comm_unique_id", "libnccl.so.2 could not be loaded (set FSB_NCCL_LIB)");
    return FSB_ERR_UNSUPPORTED;
  }
  NcclUniqueId uid;
  const int r = api.get_unique_id(&uid);
  if (r != 0) {
    note("ncclGetUniqueId", api.error_string ? api.error_string(r) : "error");
    return FSB_ERR_CUDA;
  }
  memcpy(id, &uid, sizeof(uid));
  return FSB_OK;
}

static fsb_comm* new_comm(fsb_handle_t h, int world, int rank) {
  fsb_comm* c = new (std::nothrow) fsb_comm;
  if (!c) return nullptr;
  c->world = world;
  c->rank = rank;
  c->device = h->device;
  // flag-wait timeout of the peer kernel: FSB_PEER_TIMEOUT_S seconds at ~2 GHz (default 120 s)
  double secs = 120.0;
  if (const char* e = getenv("FSB_PEER_TIMEOUT_S")) secs = atof(e) > 0 ? atof(e) : secs;
  c->timeout_cycles = (long long)(secs * 2.0e9);
  return c;
}

int fsb_comm_init(fsb_handle_t h, const void* id, int world, int rank, fsb_comm_t* out) {
  if (!h || !out || world < 1 || rank < 0 || rank >= world) return FSB_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  FSB_CUDA_TRY(cudaSetDevice(h->device));
  fsb_comm* c = new_comm(h, world, rank);
  if (!c) return FSB_ERR_INVALID_ARGUMENT;
  if (world > 1) {
    NcclApi& api = nccl();
    if (!id || !api.ok) {
      delete c;
      note("fsb_comm_init", "libnccl.so.2 could not be loaded (set FSB_NCCL_LIB) or no unique id given");
      return FSB_ERR_UNSUPPORTED;
    }
    NcclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    const int r = api.comm_init_rank(&c->nccl_comm, world, uid, rank);
    if (r != 0) {
      note("ncclCommInitRank", api.error_string ? api.error_string(r) : "error");
      delete c;
      return FSB_ERR_CUDA;
    }
    c->owns_nccl = true;
  }
  *out = c;
  return FSB_OK;
}

int fsb_comm_adopt(fsb_handle_t h, void* nccl_comm, int world, int rank, fsb_comm_t* out) {
  if (!h || !out || world < 1 || rank < 0 || rank >= world || (world > 1 && !nccl_comm))
    return FSB_ERR_INVALID_ARGUMENT;
  if (world > 1 && !nccl().ok) return FSB_ERR_UNSUPPORTED;
  fsb_comm* c = new_comm(h, world, rank);
  if (!c) return FSB_ERR_INVALID_ARGUMENT;
  c->nccl_comm = nccl_comm;
  c->owns_nccl = false;
  *out = c;
  return FSB_OK;
}

size_t fsb_comm_peer_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }
size_t fsb_comm_peer_max_bytes(void) { return PEER_MAX_BYTES; }

int fsb_comm_peer_export(fsb_comm_t c, void* handle, size_t handle_bytes) {
  if (!c || !handle || handle_bytes < sizeof(cudaIpcMemHandle_t)) return FSB_ERR_INVALID_ARGUMENT;
  if (c->world > PEER_MAX_RANKS) return FSB_ERR_UNSUPPORTED;
  FSB_CUDA_TRY(cudaSetDevice(c->device));
  if (!c->window) {
    void* p = nullptr;
    FSB_CUDA_TRY(cudaMalloc(&p, PEER_WINDOW_BYTES));
    FSB_CUDA_TRY(cudaMemset(p, 0, PEER_WINDOW_BYTES));
    FSB_CUDA_TRY(cudaDeviceSynchronize());
    c->window = (char*)p;
  }
  cudaIpcMemHandle_t hd;
  FSB_CUDA_TRY(cudaIpcGetMemHandle(&hd, c->window));
  memcpy(handle, &hd, sizeof(hd));
  return FSB_OK;
}

int fsb_comm_peer_attach(fsb_comm_t c, const void* handles, int n) {
  if (!c || !handles || n != c->world || !c->window) return FSB_ERR_INVALID_ARGUMENT;
  if (c->world > PEER_MAX_RANKS) return FSB_ERR_UNSUPPORTED;
  FSB_CUDA_TRY(cudaSetDevice(c->device));
  const char* hp = (const char*)handles;
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) {
      c->peer[r] = c->window;
      continue;
    }
    cudaIpcMemHandle_t hd;
    memcpy(&hd, hp + (size_t)r * sizeof(hd), sizeof(hd));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      fsb_note_cuda_error(e, "cudaIpcOpenMemHandle");
      for (int q = 0; q < r; ++q)
        if (q != c->rank && c->peer[q]) {
          cudaIpcCloseMemHandle(c->peer[q]);
          c->peer[q] = nullptr;
        }
      cudaGetLastError();
      return FSB_ERR_CUDA;
    }
    c->peer[r] = (char*)p;
  }
  c->peer_ready = true;
  return FSB_OK;
}

int fsb_comm_peer_disable(fsb_comm_t c) {
  if (!c) return FSB_ERR_INVALID_ARGUMENT;
  c->peer_ready = false;
  return FSB_OK;
}

int fsb_comm_info(fsb_comm_t c, int64_t* out4) {
  if (!c || !out4) return FSB_ERR_INVALID_ARGUMENT;
  out4[0] = c->peer_ready ? 1 : 0;
  out4[1] = c->peer_calls;
  out4[2] = c->nccl_calls;
  out4[3] = c->world;
  return FSB_OK;
}

int fsb_allreduce(fsb_handle_t h, fsb_comm_t c, double* buf, int64_t count, void* stream) {
  if (!h || !c || count < 0 || (count > 0 && !buf)) return FSB_ERR_INVALID_ARGUMENT;
  if (c->world == 1 || count == 0) return FSB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (c->peer_ready && (size_t)count * sizeof(double) <= PEER_MAX_BYTES) {
    PeerArgs pa;
    for (int r = 0; r < PEER_MAX_RANKS; ++r) pa.win[r] = r < c->world ? c->peer[r] : nullptr;
    pa.world = c->world;
    pa.rank = c->rank;
    int64_t grid = fsb_ceil_div(count, 256);
    if (grid > 64) grid = 64;
    if (grid < 1) grid = 1;
    peer_allreduce_kernel<<<(unsigned)grid, 256, 0, s>>>(pa, buf, count, c->timeout_cycles);
    FSB_LAUNCH_CHECK("peer_allreduce_kernel");
    c->peer_calls++;
    return FSB_OK;
  }
  if (!c->nccl_comm) return FSB_ERR_UNSUPPORTED;
  NcclApi& api = nccl();
  const int r = api.all_reduce(buf, buf, (size_t)count, kNcclFloat64, kNcclSum, c->nccl_comm, s);
  if (r != 0) {
    note("ncclAllReduce", api.error_string ? api.error_string(r) : "error");
    return FSB_ERR_CUDA;
  }
  c->nccl_calls++;
  return FSB_OK;
}

int fsb_comm_destroy(fsb_comm_t c) {
  if (!c) return FSB_OK;
  cudaSetDevice(c->device);
  for (int r = 0; r < c->world && r < PEER_MAX_RANKS; ++r)
    if (r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
  if (c->window) cudaFree(c->window);
  if (c->owns_nccl && c->nccl_comm && nccl().ok) nccl().comm_destroy(c->nccl_comm);
  cudaGetLastError();
  delete c;
  return FSB_OK;
}

}  // extern "C"
