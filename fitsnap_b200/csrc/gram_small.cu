// K2+K3+K4 for narrow design matrices (k + 1 <= 104, i.e. NB <= 13 DMMA blocks of 8 columns):
// warp-specialised, the whole lower triangle of the augmented Gram lives in registers.
//
// Why a second kernel, and why it looks like this (all measured on B200):
//  * With one 128x128 super-tile (gram.cu) only 10 of 16 warp tiles exist at k = 100 and the busiest
//    scheduler carries 26 of the 91 DMMAs of a k-step.  Here the ROWS are split instead of the
//    output: two consumer groups of 4 warps alternate the k-steps of a stage; inside a group the
//    91 (i, j <= i) blocks are split by block row into 4 parts of 21..25 DMMAs, and the second
//    group uses the parts in reverse order, so every scheduler (warp % 4) carries 45-46 DMMAs per
//    pair of k-steps and each warp reuses <= 13 fragments for its DMMAs.
//  * cp.async (LDGSTS) staging turned out to be the bottleneck: ~46 cycles per warp-instruction per
//    SM (3.1 TB/s with 16-byte copies, 1.3 TB/s with 8-byte ones), and because every warp issued
//    its copies after the stage barrier, copy time (0.27 ms) and DMMA time (0.38 ms) simply added
//    up.  Plain coalesced LDG runs at 4.5-5.8 TB/s on the same data, so staging is done by four
//    PRODUCER warps: LDG -> (weight, mask, b column, zero padding) -> STS into a ring of stages,
//    while the eight CONSUMER warps do nothing but LDS + DMMA.  Producers apply fl(w*a) once per
//    element, so the consumers' DMULs (which share the fp64 pipe with DMMA) are gone as well.
//  * ring synchronisation uses named barriers (bar.sync / bar.arrive): FULL[s] producers -> consumers,
//    EMPTY[s] consumers -> producers.  No alignment requirement on lda or k (31, 69, 110 are real).
//  * smem row pitch 8*NB+4 doubles (== 4 or 12 mod 16): conflict-free fragment loads.
//  * each consumer group writes its partial Gram as one split-K "chunk"; gram_reduce_kernel
//    (gram.cu) sums the chunks in a fixed order (deterministic).
#include "fsb_common.cuh"
#include <limits.h>
#include <stdlib.h>

namespace {

struct SmallArgs {
  const double* A;
  int64_t lda;
  const double* b;
  const double* w;      // effective weights (0 for excluded rows)
  int64_t n_rows;
  int k;
  int64_t rows_per_cta;
  double* partial;      // [cta * S_GROUPS + group][128][128]
  int debug;            // tuning: bit0 = L2 prefetch in the flattened producer
};

constexpr int S_CONSUMERS = 256;   // warps 0-7
constexpr int S_PRODUCERS = 128;   // warps 8-11
constexpr int S_THREADS = S_CONSUMERS + S_PRODUCERS;
constexpr int S_RCH = 32;          // rows per stage
constexpr int S_GROUPS = 2;        // consumer groups (4 warps each)
constexpr int S_MAXSTAGE = 7;      // 2 * stages named barriers + barrier 0 <= 16

__host__ __device__ constexpr int ring_depth(int nb) {
  int stage_bytes = S_RCH * (8 * nb + 4) * 8;
  int n = 200 * 1024 / stage_bytes;
  return n > S_MAXSTAGE ? S_MAXSTAGE : (n < 2 ? 2 : n);
}

// block rows [part_begin(nb,p), part_begin(nb,p+1)) belong to part p (4 parts, balanced by i(i+1)/2)
__host__ __device__ constexpr int part_begin(int nb, int p) {
  if (p <= 0) return 0;
  if (p >= 4) return nb;
  const int total = nb * (nb + 1) / 2;
  int r = 0;
  // first r whose cumulative count r(r+1)/2 reaches (about) total * p / 4
  while (r < nb && 4 * (r * (r + 1) / 2) + 2 * (r + 1) < total * p) ++r;
  return r;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// `wring` == nullptr: the ring holds weighted rows (gram_rowsplit_kernel: the producers multiplied).  Otherwise the
// ring holds UNWEIGHTED rows and wring[slot][row] their weights: the consumer scales its fragments, fl(w * a) exactly
// as a producer would (svd.py:44), right before the DMMAs that use them.  The fp64 pipe is shared by DMUL and DMMA: a
// producer warp's DMUL queues behind the consumers' DMMAs (~130 cycles each, measured: the producers of the fused
// kernel spent most of a stage waiting for 32 multiplies), whereas 13 DMULs in front of 21-25 DMMAs of the same
// warp cost ~7 % of the pipe.
template <int R0, int R1>
__device__ __forceinline__ void consume(double* partial, const double* smem, int pitch, int nstage, int group,
                                        int lane, int nsteps, int64_t cta, const double* wring = nullptr) {
  constexpr int NACC = (R1 * (R1 + 1) - R0 * (R0 + 1)) / 2 > 0 ? (R1 * (R1 + 1) - R0 * (R0 + 1)) / 2 : 1;
  constexpr int base = R0 * (R0 + 1) / 2;
  double acc[NACC][2];
#pragma unroll
  for (int q = 0; q < NACC; ++q) acc[q][0] = acc[q][1] = 0.0;
  const int stage_doubles = S_RCH * pitch;
  const int frag_off = (lane & 3) * pitch + (lane >> 2);
  for (int s = 0; s < nsteps; ++s) {
    const int slot = s % nstage;
    bar_sync(1 + slot, S_THREADS);                       // FULL[slot]: the producers' stores are visible
    const double* st = smem + (size_t)slot * stage_doubles + frag_off;
    // fragments of k-step q+1 are loaded before the DMMAs of k-step q are issued, so the LDS latency
    // of one warp never coincides with an empty DMMA queue on its scheduler
    constexpr int Q = S_RCH / 4 / S_GROUPS;
    double f[2][R1 > 0 ? R1 : 1];
#pragma unroll
    for (int c = 0; c < R1; ++c) f[0][c] = st[group * 4 * pitch + c * 8];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      if (q + 1 < Q) {
        const double* frow = st + (group + S_GROUPS * (q + 1)) * 4 * pitch;   // this group's next k-step
#pragma unroll
        for (int c = 0; c < R1; ++c) f[(q + 1) & 1][c] = frow[c * 8];
      }
      if (wring) {
        const double wq = wring[slot * S_RCH + (group + S_GROUPS * q) * 4 + (lane & 3)];
#pragma unroll
        for (int c = 0; c < R1; ++c) f[q & 1][c] *= wq;
      }
#pragma unroll
      for (int i = R0; i < R1; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j)
          dmma884(acc[i * (i + 1) / 2 + j - base][0], acc[i * (i + 1) / 2 + j - base][1], f[q & 1][i], f[q & 1][j]);
    }
    if (s + nstage < nsteps) bar_arrive(1 + S_MAXSTAGE + slot, S_THREADS);   // EMPTY[slot]
  }
  double* out = partial + ((size_t)cta * S_GROUPS + group) * (size_t)(FSB_GT * FSB_GT);
#pragma unroll
  for (int i = R0; i < R1; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      const int row = i * 8 + (lane >> 2), col = j * 8 + 2 * (lane & 3);
      *reinterpret_cast<double2*>(out + row * FSB_GT + col) =
          make_double2(acc[i * (i + 1) / 2 + j - base][0], acc[i * (i + 1) / 2 + j - base][1]);
    }
}

template <int NB>
__global__ void __launch_bounds__(S_THREADS, 1) gram_rowsplit_kernel(SmallArgs p) {
  extern __shared__ double smem[];   // [NSTAGE][S_RCH x PITCH]
  constexpr int PITCH = 8 * NB + 4;
  constexpr int NSTAGE = ring_depth(NB);
  constexpr int KP = 8 * NB;
  const int k = p.k;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t cta = blockIdx.x;
  const int64_t row_begin = cta * p.rows_per_cta;
  int64_t row_end = row_begin + p.rows_per_cta;
  if (row_end > p.n_rows) row_end = p.n_rows;
  const int nsteps = row_end > row_begin ? (int)((row_end - row_begin + S_RCH - 1) / S_RCH) : 0;

  if (warp >= S_CONSUMERS / 32) {
    if (KP >= 64) {
    // ------------------------------ producers: LDG -> weight -> STS ------------------------------
      const int c = tid - S_CONSUMERS;            // column of the augmented matrix owned by this thread
      const double* src; int64_t stride; bool use;
      if (c < k) { src = p.A + c; stride = p.lda; use = true; }
      else if (c == k) { src = p.b; stride = 1; use = true; }
      else { src = p.w; stride = 0; use = false; }
      const int64_t last_row = p.n_rows > 0 ? p.n_rows - 1 : 0;
      // Two register buffers: the loads of stage s+1 are in flight while stage s is weighted and
      // stored, so a producer thread always has 32-64 independent 8-byte loads outstanding.
      double va[S_RCH], vb[S_RCH];
      double wa = 0.0, wb = 0.0;
      auto load_stage = [&](int s, double (&v)[S_RCH], double& wl) {
        const int64_t r0 = row_begin + (int64_t)s * S_RCH;
        const int64_t rw = r0 + lane;                      // row weights: one coalesced load per warp
        wl = (rw < row_end) ? __ldg(p.w + rw) : 0.0;
        if (c < KP) {
#pragma unroll
          for (int i = 0; i < S_RCH; ++i) {
            const int64_t r = r0 + i;
            const int64_t rc = r < row_end ? r : last_row;   // clamped; rows past the end get weight 0
            v[i] = __ldg(src + rc * stride);
          }
        }
      };
      auto store_stage = [&](int s, const double (&v)[S_RCH], double wl) {
        const int slot = s % NSTAGE;
        if (s >= NSTAGE) bar_sync(1 + S_MAXSTAGE + slot, S_THREADS);      // EMPTY[slot]: consumers are done with it
        double* st = smem + (size_t)slot * (S_RCH * PITCH) + c;
#pragma unroll
        for (int i = 0; i < S_RCH; ++i) {
          const double wi = __shfl_sync(0xffffffffu, wl, i);
          if (c < KP) st[i * PITCH] = use ? v[i] * wi : 0.0;     // fl(w*a): the reference's aw (svd.py:44)
        }
        __threadfence_block();                                             // stores ordered before the arrival
        bar_arrive(1 + slot, S_THREADS);                                   // FULL[slot]
      };
      if (nsteps > 0) load_stage(0, va, wa);
      for (int s = 0; s < nsteps; s += 2) {
        if (s + 1 < nsteps) load_stage(s + 1, vb, wb);
        store_stage(s, va, wa);
        if (s + 1 < nsteps) {
          if (s + 2 < nsteps) load_stage(s + 2, va, wa);
          store_stage(s + 1, vb, wb);
        }
      }

      return;
    }
    // ------------------------------ producers: LDG -> weight -> STS ------------------------------
    // The stage (S_RCH rows x KP columns) is flattened over the 128 producer threads so that every
    // thread owns EPT elements whatever KP is; consecutive threads read consecutive addresses.
    constexpr int NEL = S_RCH * KP;
    constexpr int EPT = (NEL + S_PRODUCERS - 1) / S_PRODUCERS;
    const int tp = tid - S_CONSUMERS;
    const int64_t last_row = p.n_rows > 0 ? p.n_rows - 1 : 0;
    double va[EPT], vb[EPT];
    double wa = 0.0, wb = 0.0;
    auto load_stage = [&](int s, double (&v)[EPT], double& wl) {
      const int64_t r0 = row_begin + (int64_t)s * S_RCH;
      const int64_t rw = r0 + lane;                      // row weights: one coalesced load per warp
      wl = (rw < row_end) ? __ldg(p.w + rw) : 0.0;
#pragma unroll
      for (int i = 0; i < EPT; ++i) {
        const int e = tp + S_PRODUCERS * i;
        const int row = e / KP, col = e - row * KP;        // KP is a compile-time constant
        const int64_t r = r0 + row;
        const int64_t rc = r < row_end ? r : last_row;     // clamped; rows past the end get weight 0
        const double* src = (col < k) ? p.A + rc * p.lda + col : ((col == k) ? p.b + rc : p.w);
        v[i] = (e < NEL) ? __ldg(src) : 0.0;
      }
    };
    auto store_stage = [&](int s, const double (&v)[EPT], double wl) {
      const int slot = s % NSTAGE;
      if (s >= NSTAGE) bar_sync(1 + S_MAXSTAGE + slot, S_THREADS);      // EMPTY[slot]: consumers are done with it
      double* st = smem + (size_t)slot * (S_RCH * PITCH);
#pragma unroll
      for (int i = 0; i < EPT; ++i) {
        const int e = tp + S_PRODUCERS * i;
        const int row = e / KP, col = e - row * KP;
        const double wi = __shfl_sync(0xffffffffu, wl, row & 31);       // row < 32 whenever e < NEL
        if (e < NEL) st[row * PITCH + col] = (col <= k) ? v[i] * wi : 0.0;   // fl(w*a): the reference's aw (svd.py:44)
      }
      __threadfence_block();                                             // stores ordered before the arrival
      bar_arrive(1 + slot, S_THREADS);                                   // FULL[slot]
    };
    // DRAM latency is taken out of the demand loads by prefetching the rows of stage s+3 into L2
    // (one prefetch per 128-byte line, spread over the producer threads; no registers tied up).
    auto prefetch_stage = [&](int s) {
      if (s >= nsteps || !(p.debug & 1)) return;
      const int64_t r0 = row_begin + (int64_t)s * S_RCH;
      const int64_t nr = (row_end - r0) < S_RCH ? (row_end - r0) : S_RCH;
      const char* base = reinterpret_cast<const char*>(p.A + r0 * p.lda);
      const int64_t nbytes = ((nr - 1) * p.lda + k) * (int64_t)sizeof(double);
      for (int64_t off = (int64_t)tp * 128; off < nbytes; off += (int64_t)S_PRODUCERS * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
    };
    constexpr int PF = 3;
    for (int s = 1; s <= PF; ++s) prefetch_stage(s);
    if (nsteps > 0) load_stage(0, va, wa);
    for (int s = 0; s < nsteps; s += 2) {
      prefetch_stage(s + 1 + PF);
      if (s + 1 < nsteps) load_stage(s + 1, vb, wb);
      store_stage(s, va, wa);
      if (s + 1 < nsteps) {
        prefetch_stage(s + 2 + PF);
        if (s + 2 < nsteps) load_stage(s + 2, va, wa);
        store_stage(s + 1, vb, wb);
      }
    }
    return;
  }

  // -------------------------------- consumers: LDS -> DMMA ---------------------------------------
  const int group = warp >> 2;
  const int part = (group == 0) ? (warp & 3) : 3 - (warp & 3);   // reversed in group 1: schedulers balanced
  constexpr int B1 = part_begin(NB, 1), B2 = part_begin(NB, 2), B3 = part_begin(NB, 3);
  switch (part) {
    case 0: consume<0, B1>(p.partial, smem, PITCH, NSTAGE, group, lane, nsteps, cta); break;
    case 1: consume<B1, B2>(p.partial, smem, PITCH, NSTAGE, group, lane, nsteps, cta); break;
    case 2: consume<B2, B3>(p.partial, smem, PITCH, NSTAGE, group, lane, nsteps, cta); break;
    default: consume<B3, NB>(p.partial, smem, PITCH, NSTAGE, group, lane, nsteps, cta); break;
  }
}

// ------------------------------------------------------------------------------------------------
// K1 + K2..K4 FUSED: raw LAMMPS blocks -> rows of A, b, w (written once, for the refinement passes; or not at all)
// AND, from the same registers, the weighted rows that feed the DMMA consumers.  The design matrix is never read
// back for the Gram: the scatter's HBM traffic (16k + 24 bytes per row) hides behind the fp64 tensor work, which is
// the longer of the two (this is the per-batch `C += aw^T aw; d += aw^T bw` of
// examples/library/transpose_trick/example.py:226-246, with `_collect_lammps` (lammps_snap.py:391-556) fused in).
//
// Same CTA shape as gram_rowsplit_kernel: 8 consumer warps (identical code: the Gram is bit-identical to
// scatter -> gram_rowsplit on the same rows) + 4 producer warps.  Producer thread c owns column c of the augmented
// matrix [A | b]: per half stage it loads its raw column for 16 rows (coalesced across the warp: raw columns are
// contiguous), applies the row's transform (/N, *1.6021765e6/V, type fraction, blank2J -- the operations of
// scatter_kernel in the same order, bit-identical A, b, w), stores A (coalesced) and the weighted value into the
// ring.  Row metadata (row -> configuration -> kind / divisor / weight / truth) is a chain of dependent global
// loads: lane r of every producer warp resolves row r of a stage, three stages ahead of its use, one level of the
// chain per stage, and the store loop broadcasts it with shuffles.
// Requirements: energy + force + virial rows all assembled (raw row i <-> output row i), row_cfg given,
// k + 1 <= 104.  Anything else takes the two separate kernels.
struct FusedArgs {
  ScatterArgs sc;
  const uint8_t* testing;   // optional test mask of the output rows of this call (1 = excluded from the Gram)
  int64_t total;            // rows of this call
  int64_t rows_per_cta;
  double* partial;
  const struct FusedDesc* desc;   // per-row descriptors (row_resolve_body)
  int store_a;              // 0: streaming mode, A is not materialised (b and w still are)
  int spec_from_a;          // 1: the energy / virial rows of A were written by special_rows_body before this launch
};

struct FusedDesc {           // one per output row, written by row_resolve_body
  double wg;                 // weight seen by the Gram: w, or 0 for a test row
  double div;                // N (energy row), V (virial row), 1 (force row)
  int kind;                  // 0 energy, 1 force, 2 virial
  int cfg;
};

// Row metadata is a chain of dependent global loads (row -> configuration -> offsets / weights / truth -> stress
// component): resolved here for every row at once -- b and w are final after this kernel, the fused kernel below reads
// one 24-byte descriptor per row.  ~1e6 rows: tens of microseconds, 2 % of the fused kernel's traffic.
__device__ __forceinline__ void row_resolve_body(const ScatterArgs& a, const uint8_t* __restrict__ testing, int64_t total,
                                                 FusedDesc* __restrict__ desc, int64_t block) {
  const int64_t i = block * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const bool do_scrub = a.flags & FSB_SCRUB_NONFINITE;
  const int kraw = a.ncoeff * a.numtypes;
  const int64_t ldr = kraw + 1;
  const int64_t row0 = __ldg(a.out_row_off), rraw0 = __ldg(a.raw_row_off);
  const int64_t n_force = total - 7 * (int64_t)a.ncfg > 1 ? total - 7 * (int64_t)a.ncfg : 1;   // 3 * atoms
  const int c = __ldg(a.row_cfg + i);
  double ref = __ldg(a.raw + (rraw0 + i) * ldr + kraw);       // reference-potential column of this row
  int64_t fi = i - 7 * (int64_t)c - 1;                        // rows map 1:1 (see scatter_bulk_kernel)
  fi = fi < 0 ? 0 : (fi > n_force - 1 ? n_force - 1 : fi);
  const int nat = __ldg(a.natoms + c);
  const int64_t off_c = __ldg(a.out_row_off + c);
  const double ew = __ldg(a.eweight + c), fw = __ldg(a.fweight + c), vw = __ldg(a.vweight + c);
  const double en = __ldg(a.energy + c), vol = __ldg(a.volume + c), fo = __ldg(a.forces + fi);
  const int64_t local = row0 + i - off_c;
  FusedDesc d;
  double truth, wv;
  d.cfg = c;
  if (local == 0) {
    d.kind = 0; d.div = (double)nat; wv = ew; truth = en;
  } else if (local < 1 + 3 * (int64_t)nat) {
    d.kind = 1; d.div = 1.0; wv = fw; truth = fo;
  } else {
    const int sub = (int)(local - 1 - 3 * (int64_t)nat);
    const int vi[6] = {0, 1, 2, 1, 0, 0}, vj[6] = {0, 1, 2, 2, 2, 1};
    d.kind = 2; d.div = vol; wv = vw;
    truth = __ldg(a.stress + (size_t)c * 9 + vi[sub] * 3 + vj[sub]);
  }
  bool bad = false;
  ref = fsb_dev::scrub(ref, do_scrub, bad);
  if (bad && a.nonfinite) atomicAdd(a.nonfinite, 1);
  a.b[row0 + i] = (d.kind == 0) ? (truth - ref) / d.div : truth - ref;     // lammps_snap.py:473, :506-507, :540-541
  a.w[row0 + i] = wv;
  d.wg = (testing && __ldg(testing + i)) ? 0.0 : wv;
  desc[i] = d;
}

// Energy and virial rows of A (7 of the rows of a configuration), computed BEFORE the fused kernel: one CTA per
// configuration, the arithmetic of scatter_kernel (same operations, same order: bit-identical values).
// Why not inside scatter_gram_kernel: an fp64 division is ~10 dependent fp64 operations, and every fp64 operation of
// a producer warp queues behind the consumers' DMMAs on the shared fp64 pipe.  Measured on 1e6 x 100 (7 % special
// rows): those rows took ~46 % of the producers' busy time, the producers -- not the DMMA pipe -- set the pace of the
// kernel (consumers 27 % of their time at the FULL barrier), and the kernel time grew LINEARLY with the number of fp64
// operations the producers issue per stage (0.61 ms at ~27, 0.70 ms at ~40, 0.83 ms at ~96: interleaving independent
// division chains did not help, the pipe serialises them).  With this kernel the producers issue no fp64 operation
// at all for these rows: they copy the finished values from A (L2-resident, written microseconds earlier) into the
// DMMA ring.
__device__ __forceinline__ void special_rows_body(const ScatterArgs& a, int cfg_block) {
  // one CTA per configuration: its 7 special rows x k columns are dealt to the 256 threads (<= 3 elements each for
  // k <= 104), the per-configuration scalars are loaded once and every thread's raw loads are issued together -- with
  // one 128-thread CTA per ROW the kernel was a chain of dependent loads over 70 000 tiny CTAs (60 us at 1e4
  // configurations, a tenth of the fused phase)
  const int cfg = cfg_block;
  const bool bzero = a.flags & FSB_BZEROFLAG, do_scrub = a.flags & FSB_SCRUB_NONFINITE;
  const int kraw = a.ncoeff * a.numtypes;
  const int k = bzero ? kraw : kraw + a.numtypes;
  const int seg = a.ncoeff + 1;
  const int n = __ldg(a.natoms + cfg);
  const int64_t o0 = __ldg(a.out_row_off + cfg), r0 = __ldg(a.raw_row_off + cfg);
  const double vol = __ldg(a.volume + cfg), dn = (double)n;
  constexpr int EPT = 3;                                   // elements per thread: 7 * 104 <= 3 * 256
  int sub[EPT], col[EPT], srcc[EPT];
  double x[EPT];
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    const int idx = threadIdx.x + 256 * e;
    sub[e] = idx / k;                                      // 0: energy row, 1..6: virial rows; >= 7: nothing
    col[e] = idx - sub[e] * k;
    int v = col[e];
    if (!bzero) {
      const int t = col[e] / seg, q = col[e] - t * seg;
      v = (q == 0) ? -(t + 1) : t * a.ncoeff + q - 1;
    }
    srcc[e] = v;
    const int64_t local = (sub[e] == 0) ? 0 : 3 * (int64_t)n + sub[e];   // row inside the configuration (raw and output)
    x[e] = (sub[e] < 7 && v >= 0) ? __ldg(a.raw + (r0 + local) * (int64_t)(kraw + 1) + v) : 0.0;
  }
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    if (sub[e] >= 7) continue;
    double xv = x[e];
    if (do_scrub) { bool dummy = false; xv = fsb_dev::scrub(xv, true, dummy); }
    double val;
    if (sub[e] == 0) val = (srcc[e] >= 0) ? xv / dn : __ldg(a.type_fraction + (size_t)cfg * a.numtypes + (-srcc[e] - 1));
    else val = (srcc[e] >= 0) ? (FSB_VIRIAL_UNIT * xv) / vol : 0.0;                 // lammps_snap.py:526-536
    const int64_t local = (sub[e] == 0) ? 0 : 3 * (int64_t)n + sub[e];
    a.A[(o0 + local) * a.lda + col[e]] = val * __ldg(a.blank2j + col[e]);
  }
}

// value of A for an energy / virial row (or any row when non-finite raw values are being scrubbed): out of line, the two
// fp64 divisions exist once in the kernel
__device__ __noinline__ double fused_special_row(double x, int kind, double div, double tf, double pref, bool loads_raw,
                                                 bool do_scrub) {
  bool dummy = false;
  if (do_scrub) x = fsb_dev::scrub(x, true, dummy);
  if (kind == 1) return (loads_raw ? x : 0.0) * pref;                                   // lammps_snap.py:493-502
  if (kind == 2) return (loads_raw ? (FSB_VIRIAL_UNIT * x) / div : 0.0) * pref;         // :526-536
  return (loads_raw ? x / div : tf) * pref;                                             // :435-467
}

// The two preparation steps of the fused kernel in ONE launch (they are independent: blocks [0, nspecial) finish the
// energy / virial rows of one configuration each, the others resolve 256 row descriptors each), so that the ~20 us of
// the first hide behind the ~28 us of the second instead of preceding them.
__global__ void __launch_bounds__(256) fused_prep_kernel(ScatterArgs a, const uint8_t* __restrict__ testing, int64_t total,
                                                         FusedDesc* __restrict__ desc, int nspecial) {
  if ((int64_t)blockIdx.x < (int64_t)nspecial) special_rows_body(a, (int)blockIdx.x);
  else row_resolve_body(a, testing, total, desc, (int64_t)blockIdx.x - nspecial);
}

__device__ __forceinline__ unsigned fz_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fz_mbar_wait(unsigned bar, unsigned parity) {
  unsigned done = 0;
  for (unsigned spin = 0; spin < (1u << 26); ++spin) {
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
  __trap();      // protocol error: fail the launch instead of hanging the device
}

__device__ __forceinline__ double fz_lds(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void fz_sts(unsigned addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}

constexpr int FZ_RAW_STAGES = 3;     // raw tiles in flight per SM (TMA bulk copies)
constexpr int FZ_PBAR = 15;          // named barrier of the 128 producer threads

template <int NB>
__global__ void __launch_bounds__(S_THREADS, 1) scatter_gram_kernel(FusedArgs p, int nstage_d) {
  extern __shared__ __align__(128) double smem[];   // [nstage_d][S_RCH x PITCH] DMMA ring | [FZ_RAW_STAGES][S_RCH x ldr] raw ring
  __shared__ __align__(8) unsigned long long raw_full[FZ_RAW_STAGES];
  __shared__ __align__(8) unsigned long long raw_empty[FZ_RAW_STAGES];
  __shared__ __align__(16) double wring[S_MAXSTAGE * S_RCH];          // Gram weights of the rows of every ring stage
  constexpr int PITCH = 8 * NB + 4;
  constexpr int KP = 8 * NB;
  const ScatterArgs& a = p.sc;
  const bool bzero = a.flags & FSB_BZEROFLAG, do_scrub = a.flags & FSB_SCRUB_NONFINITE;
  const int kraw = a.ncoeff * a.numtypes;
  const int k = bzero ? kraw : kraw + a.numtypes;
  const int seg = a.ncoeff + 1;
  const int ldr = kraw + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t cta = blockIdx.x;
  const int64_t row_begin = cta * p.rows_per_cta;
  int64_t row_end = row_begin + p.rows_per_cta;
  if (row_end > p.total) row_end = p.total;
  const int nsteps = row_end > row_begin ? (int)((row_end - row_begin + S_RCH - 1) / S_RCH) : 0;
  double* raw_ring = smem + (size_t)nstage_d * (S_RCH * PITCH);
  const int raw_stage = S_RCH * ldr;                       // doubles per raw tile

  if (tid == S_CONSUMERS) {
    for (int i = 0; i < FZ_RAW_STAGES; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fz_smem_u32(&raw_full[i])), "r"(1) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fz_smem_u32(&raw_empty[i])), "r"(S_PRODUCERS / 32)
                   : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= S_CONSUMERS / 32) {
    const int c = tid - S_CONSUMERS;          // column of A owned by this thread (the b column belongs to warp 0's lanes)
    int srcc = INT_MIN;                       // >= 0: raw column; < 0: lead column of type (-v-1); INT_MIN: not a column of A
    double pref = 0.0;
    if (c < k) {
      int v = c;
      if (!bzero) {
        const int t = c / seg, q = c - t * seg;
        v = (q == 0) ? -(t + 1) : t * a.ncoeff + q - 1;
      }
      srcc = v;
      pref = __ldg(a.blank2j + c);
    }
    const bool unit_pref = pref == 1.0;       // blank2J is a 0/1 mask in practice: x * 1.0 == x, and the fp64 pipe
                                              // (shared with the consumers' DMMAs) is spared one multiply per element
    const int64_t row0 = a.out_row_off[0], rraw0 = a.raw_row_off[0];
    const double* raw0 = a.raw + rraw0 * (int64_t)ldr;          // raw row of output row index 0 of this call
    const bool loads_raw = srcc >= 0;
    const bool acol = c < k, do_store = acol && p.store_a;
    const bool bwarp = warp == S_CONSUMERS / 32;
    const bool ring_thread = c < KP && c != k;
    const bool warp_unit_pref = __all_sync(0xffffffffu, unit_pref || !acol);     // warp-uniform
    bool bad = false;

    // The TMA unit moves the raw tiles (one elected thread, cp.async.bulk + mbarrier complete_tx, FZ_RAW_STAGES tiles in
    // flight): a producer thread holds no load registers and computes no global load addresses.  A stage (32 rows) is
    // transformed from shared memory in two passes:
    //   pass 1, one basic block: every row is treated as a force row (93 % of them are: A = R * blank2J);
    //   pass 2, warp-uniform bit tests: energy / virial rows (divisions, type fractions; ~7 % of the rows) are redone
    //           by an out-of-line routine and overwrite what pass 1 stored.
    // Lane r of every producer warp holds the descriptor of row r of the stage (loaded one stage ahead); the b column
    // of the ring is written by lane r of producer warp 0.
    auto tile_is_bulk = [&](int t) {
      const int64_t i0 = row_begin + (int64_t)t * S_RCH;
      return (row_end - i0) >= S_RCH && (reinterpret_cast<uintptr_t>(raw0 + i0 * ldr) & 15) == 0;
    };
    auto issue_tile = [&](int t) {            // elected thread only
      if (t >= nsteps) return;
      const int slot = t % FZ_RAW_STAGES, n = t / FZ_RAW_STAGES;
      if (t >= FZ_RAW_STAGES) fz_mbar_wait(fz_smem_u32(&raw_empty[slot]), (unsigned)((n - 1) & 1));
      const unsigned full = fz_smem_u32(&raw_full[slot]);
      if (tile_is_bulk(t)) {
        const unsigned bytes = (unsigned)(raw_stage * sizeof(double));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(fz_smem_u32(raw_ring + (size_t)slot * raw_stage)),
                       "l"(raw0 + (row_begin + (int64_t)t * S_RCH) * ldr), "r"(bytes), "r"(full) : "memory");
      } else {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full) : "memory");   // filled by the producers
      }
    };
    struct StageDesc { double wg, div, bval; int kind, cfg; };
    auto load_desc = [&](int s, StageDesc& d) {
      d.wg = 0.0; d.div = 1.0; d.bval = 0.0; d.kind = -1; d.cfg = 0;
      const int64_t i = row_begin + (int64_t)s * S_RCH + lane;
      if (s < nsteps && i < row_end) {
        const FusedDesc* g = p.desc + i;
        d.wg = __ldg(&g->wg); d.div = __ldg(&g->div); d.kind = __ldg(&g->kind); d.cfg = __ldg(&g->cfg);
        if (bwarp) d.bval = __ldg(a.b + row0 + i);
      }
    };

    if (c == 0)
      for (int t = 0; t < FZ_RAW_STAGES - 1; ++t) issue_tile(t);
    StageDesc dA, dB;
    load_desc(0, dA);
    for (int s = 0; s < nsteps; ++s) {
      if (c == 0) issue_tile(s + FZ_RAW_STAGES - 1);
      load_desc(s + 1, dB);
      const unsigned valid = __ballot_sync(0xffffffffu, dA.kind >= 0);
      const unsigned special = __ballot_sync(0xffffffffu, dA.kind == 0 || dA.kind == 2);
      const unsigned spec_mask = p.spec_from_a ? special : 0u;       // rows of A that are final already: not stored here
      const int rslot = s % FZ_RAW_STAGES;
      double* rtile = raw_ring + (size_t)rslot * raw_stage;
      fz_mbar_wait(fz_smem_u32(&raw_full[rslot]), (unsigned)((s / FZ_RAW_STAGES) & 1));
      if (!tile_is_bulk(s)) {                 // ragged last tile / misaligned view: plain loads, all producers
        const int64_t i0 = row_begin + (int64_t)s * S_RCH;
        const int nr = (int)((row_end - i0) < S_RCH ? (row_end - i0) : S_RCH);
        const double* src = raw0 + i0 * ldr;
        for (int e = c; e < nr * ldr; e += S_PRODUCERS) rtile[e] = __ldg(src + e);
        bar_sync(FZ_PBAR, S_PRODUCERS);
      }
      if (s >= nstage_d) bar_sync(1 + S_MAXSTAGE + (s % nstage_d), S_THREADS);   // EMPTY[slot] of the DMMA ring
      double* st = smem + (size_t)(s % nstage_d) * (S_RCH * PITCH) + c;
      double* arow = a.A + (row0 + row_begin + (int64_t)s * S_RCH) * a.lda + c;
      const double* rs = rtile + (loads_raw ? srcc : 0);
      unsigned nf = 0u;
      if (valid == 0xffffffffu) {
        // fast path (every stage but a ragged last one): no per-row predicates.  Non-finite detection is one add and
        // one or per row: (hi & 0x7ff00000) + 0x00100000 has bit 31 set iff the exponent field is all ones.
        // Explicit 32-bit shared-memory addresses (no generic-pointer arithmetic); no fp64 instruction at all when
        // blank2J is a mask of ones: the rows go into the ring unweighted, the consumers apply the weights.
        unsigned acc = 0u;
        unsigned r_addr = fz_smem_u32(rs);
        const unsigned s_addr = fz_smem_u32(st);
        const unsigned ldr8 = (unsigned)ldr * 8u;
        // batches of 8 rows: the 8 shared-memory loads are issued back to back, then consumed (one warp per scheduler:
        // nothing else hides the LDS latency)
        if (warp_unit_pref) {
#pragma unroll
          for (int j0 = 0; j0 < S_RCH; j0 += 8) {
            double x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              x[u] = 0.0;
              if (loads_raw) x[u] = fz_lds(r_addr + (unsigned)u * ldr8);
            }
            r_addr += 8u * ldr8;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              acc |= ((unsigned)__double2hiint(x[u]) & 0x7ff00000u) + 0x00100000u;
              if (do_store && !((spec_mask >> (j0 + u)) & 1u)) arow[0] = x[u];   // blank2J == 1: A = R
              arow += a.lda;
              if (ring_thread) fz_sts(s_addr + (unsigned)((j0 + u) * PITCH * 8), x[u]);      // padding columns: x = 0
            }
          }
        } else {
#pragma unroll
          for (int j0 = 0; j0 < S_RCH; j0 += 8) {
            double x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              x[u] = 0.0;
              if (loads_raw) x[u] = fz_lds(r_addr + (unsigned)u * ldr8);
            }
            r_addr += 8u * ldr8;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              acc |= ((unsigned)__double2hiint(x[u]) & 0x7ff00000u) + 0x00100000u;
              const double val = x[u] * pref;                        // lead columns hold x = 0
              if (do_store && !((spec_mask >> (j0 + u)) & 1u)) arow[0] = val;
              arow += a.lda;
              if (ring_thread) fz_sts(s_addr + (unsigned)((j0 + u) * PITCH * 8), val);
            }
          }
        }
        nf = (acc >> 31) ? 0xffffffffu : 0u;                       // which row does not matter below
      } else {
#pragma unroll 1
        for (int j = 0; j < S_RCH; ++j) {
          const bool ok = (valid >> j) & 1u;
          const double x = (loads_raw && ok) ? rs[j * ldr] : 0.0;
          nf |= ((unsigned)__double2hiint(x) & 0x7ff00000u) == 0x7ff00000u ? (1u << j) : 0u;
          const double val = x * pref;
          if (do_store && ok && !((spec_mask >> j) & 1u)) arow[0] = val;
          arow += a.lda;
          if (ring_thread) st[j * PITCH] = (acol && ok) ? val : 0.0;
        }
      }
      nf = (loads_raw && acol) ? (nf & valid) : 0u;
      bad |= nf != 0u;
      unsigned redo = special;
      if (do_scrub && __any_sync(0xffffffffu, nf != 0u)) redo = valid;           // numpy.nan_to_num: rare
      if (redo) {
        arow = a.A + (row0 + row_begin + (int64_t)s * S_RCH) * a.lda + c;
        unsigned copy = redo & spec_mask;                        // finished by special_rows_body: A -> ring
        redo &= ~spec_mask;
        while (copy) {                                           // warp-uniform; 8 independent loads per batch
          int jj[8];
          double xv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            jj[u] = copy ? __ffs((int)copy) - 1 : -1;
            copy &= copy - 1u;                                   // 0 stays 0
            xv[u] = 0.0;
            if (acol && jj[u] >= 0) xv[u] = arow[(int64_t)jj[u] * a.lda];
          }
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (acol && jj[u] >= 0) st[jj[u] * PITCH] = xv[u];
        }
        while (redo) {                                           // warp-uniform: only the rows that need it
          const int j = __ffs((int)redo) - 1;
          redo &= redo - 1u;
          const int kind = __shfl_sync(0xffffffffu, dA.kind, j);
          const double div = __shfl_sync(0xffffffffu, dA.div, j);
          const int cfg = __shfl_sync(0xffffffffu, dA.cfg, j);
          if (acol) {
            const double x = loads_raw ? rs[j * ldr] : 0.0;
            const double tf = (kind == 0 && !loads_raw) ? __ldg(a.type_fraction + (size_t)cfg * a.numtypes + (-srcc - 1)) : 0.0;
            const double val = fused_special_row(x, kind, div, tf, pref, loads_raw, do_scrub);
            if (do_store) arow[(int64_t)j * a.lda] = val;
            st[j * PITCH] = val;
          }
        }
      }
      if (bwarp) {         // the b column of the ring and the weights of the stage's rows (0: test row / past the end)
        smem[(size_t)(s % nstage_d) * (S_RCH * PITCH) + lane * PITCH + k] = dA.bval;
        wring[(s % nstage_d) * S_RCH + lane] = dA.wg;
      }
      __syncwarp();
      if (lane == 0)       // this warp is done reading the raw tile
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fz_smem_u32(&raw_empty[rslot])) : "memory");
      __threadfence_block();
      bar_arrive(1 + (s % nstage_d), S_THREADS);                               // FULL[slot] of the DMMA ring
      if (p.spec_from_a) {
        // the finished energy / virial rows of the NEXT stage: pull them towards this SM now (they were written by
        // special_rows_body and have usually left L2 again), so that the copy loop above does not wait on DRAM
        unsigned nxt = __ballot_sync(0xffffffffu, dB.kind == 0 || dB.kind == 2);
        const double* nrow = a.A + (row0 + row_begin + (int64_t)(s + 1) * S_RCH) * a.lda + c;
        while (nxt) {
          const int j = __ffs((int)nxt) - 1;
          nxt &= nxt - 1u;
          if (acol) asm volatile("prefetch.global.L1 [%0];" ::"l"(nrow + (int64_t)j * a.lda));
        }
      }
      dA = dB;
    }
    if (a.nonfinite && __any_sync(0xffffffffu, bad) && lane == 0) atomicAdd(a.nonfinite, 1);
    return;
  }

  const int group = warp >> 2;
  const int part = (group == 0) ? (warp & 3) : 3 - (warp & 3);
  constexpr int B1 = part_begin(NB, 1), B2 = part_begin(NB, 2), B3 = part_begin(NB, 3);
  switch (part) {
    case 0: consume<0, B1>(p.partial, smem, PITCH, nstage_d, group, lane, nsteps, cta, wring); break;
    case 1: consume<B1, B2>(p.partial, smem, PITCH, nstage_d, group, lane, nsteps, cta, wring); break;
    case 2: consume<B2, B3>(p.partial, smem, PITCH, nstage_d, group, lane, nsteps, cta, wring); break;
    default: consume<B3, NB>(p.partial, smem, PITCH, nstage_d, group, lane, nsteps, cta, wring); break;
  }
}

template <int NB>
int launch_fused_nb(const FusedArgs& a, int ncta, size_t smem_optin, cudaStream_t s) {
  constexpr int PITCH = 8 * NB + 4;
  const int ldr = a.sc.ncoeff * a.sc.numtypes + 1;
  const size_t stage_d = (size_t)S_RCH * PITCH * sizeof(double);
  const size_t raw_b = (size_t)FZ_RAW_STAGES * S_RCH * ldr * sizeof(double);
  if (smem_optin < raw_b + 2 * stage_d + 2048) return FSB_ERR_UNSUPPORTED;
  int nstage_d = (int)((smem_optin - 2048 - raw_b) / stage_d);
  if (nstage_d > S_MAXSTAGE) nstage_d = S_MAXSTAGE;
  const size_t smem = (size_t)nstage_d * stage_d + raw_b;
  FSB_CUDA_TRY(cudaFuncSetAttribute(scatter_gram_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  scatter_gram_kernel<NB><<<ncta, S_THREADS, smem, s>>>(a, nstage_d);
  FSB_LAUNCH_CHECK("scatter_gram_kernel");
  return FSB_OK;
}

template <int NB>
int launch_nb(const SmallArgs& a, int ncta, cudaStream_t s) {
  constexpr int PITCH = 8 * NB + 4;
  const size_t smem = (size_t)ring_depth(NB) * (S_RCH * PITCH) * sizeof(double);
  FSB_CUDA_TRY(cudaFuncSetAttribute(gram_rowsplit_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gram_rowsplit_kernel<NB><<<ncta, S_THREADS, smem, s>>>(a);
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
  { const char* e = getenv("FSB_GRAM_DEBUG"); a.debug = e ? atoi(e) : 0; }
  const int want = fsb_gram_small_ctas(h, n_rows);
  a.rows_per_cta = fsb_round_up(fsb_ceil_div(n_rows > 0 ? n_rows : 1, want), S_RCH);
  // every (cta < want, group) slot of the workspace must be written: launch `want` CTAs; surplus
  // ones see an empty row range and write zeros
  switch (nb) {
#define FSB_NB(N) case N: return launch_nb<N>(a, want, s);
    FSB_NB(1) FSB_NB(2) FSB_NB(3) FSB_NB(4) FSB_NB(5) FSB_NB(6) FSB_NB(7)
    FSB_NB(8) FSB_NB(9) FSB_NB(10) FSB_NB(11) FSB_NB(12) FSB_NB(13)
#undef FSB_NB
    default: return FSB_ERR_UNSUPPORTED;
  }
}

// fused scatter + Gram (see scatter_gram_kernel); the caller reduces the split-K partials like the unfused path
bool fsb_scatter_gram_supported(const ScatterArgs& sc, int64_t total) {
  const int all_rows = FSB_ROWS_ENERGY | FSB_ROWS_FORCE | FSB_ROWS_STRESS;
  const bool bzero = sc.flags & FSB_BZEROFLAG;
  const int k = sc.ncoeff * sc.numtypes + (bzero ? 0 : sc.numtypes);
  static int off = -1;
  if (off < 0) off = getenv("FSB_NO_FUSED_SCATTER_GRAM") ? 1 : 0;
  return !off && (sc.flags & all_rows) == all_rows && sc.row_cfg != nullptr && (k + 1 + 7) / 8 <= 13 && total > 0;
}

size_t fsb_scatter_gram_desc_bytes(int64_t total) { return (size_t)(total > 0 ? total : 1) * sizeof(FusedDesc); }

int fsb_launch_scatter_gram_small(const fsb_context* h, const ScatterArgs& sc, const uint8_t* testing, int64_t total,
                                  int store_a, double* partial, void* desc, cudaStream_t s) {
  const bool bzero = sc.flags & FSB_BZEROFLAG;
  const int k = sc.ncoeff * sc.numtypes + (bzero ? 0 : sc.numtypes);
  const int nb = (k + 1 + 7) / 8;
  FusedArgs a;
  a.sc = sc; a.testing = testing; a.total = total; a.partial = partial; a.store_a = store_a;
  a.desc = (const FusedDesc*)desc;
  static int no_pre = -1;          // development switch: the divisions of the special rows inside the fused kernel
  if (no_pre < 0) no_pre = getenv("FSB_FUSED_NO_PRESPECIAL") ? 1 : 0;
  a.spec_from_a = (store_a && sc.A && sc.ncfg > 0 && !no_pre) ? 1 : 0;
  const int nspecial = a.spec_from_a ? sc.ncfg : 0;
  fused_prep_kernel<<<(unsigned)(nspecial + fsb_ceil_div(total, 256)), 256, 0, s>>>(sc, testing, total, (FusedDesc*)desc,
                                                                                   nspecial);
  FSB_LAUNCH_CHECK("fused_prep_kernel");
  const int want = fsb_gram_small_ctas(h, total);
  a.rows_per_cta = fsb_round_up(fsb_ceil_div(total > 0 ? total : 1, want), S_RCH);
  switch (nb) {
#define FSB_NB(N) case N: return launch_fused_nb<N>(a, want, h->smem_optin, s);
    FSB_NB(1) FSB_NB(2) FSB_NB(3) FSB_NB(4) FSB_NB(5) FSB_NB(6) FSB_NB(7)
    FSB_NB(8) FSB_NB(9) FSB_NB(10) FSB_NB(11) FSB_NB(12) FSB_NB(13)
#undef FSB_NB
    default: return FSB_ERR_UNSUPPORTED;
  }
}
