// Microbenchmark behind DESIGN.md finding 5: how fast can one B200 stream a buffer out of DRAM
//   (a) with plain LDG, U independent 8-byte loads per thread, at the occupancy the launch allows,
//   (b) with cp.async.bulk (TMA 1-D) tiles into a shared-memory ring of S stages, one producer thread per CTA,
// for several tile sizes / ring depths / CTAs per SM.  Every kernel sums what it reads (one fp64 add per
// element) so that nothing is optimised away.  Build:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/stream_rates tools/ubench/stream_rates.cu
// Run on the GPU box:  tools/ubench/stream_rates [GiB]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

template <int U>
__global__ void __launch_bounds__(256) ldg_sum(const double* __restrict__ p, size_t n, double* out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0.0;
  for (; i + (U - 1) * stride < n; i += U * stride) {
    double v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldg(p + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u];
  }
  for (; i < n; i += stride) acc += __ldg(p + i);
  if (acc == 12345.678) out[0] = acc;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned done = 0;
  for (unsigned spin = 0; spin < (1u << 26); ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}

// one producer thread + 8 consumer warps; tile_bytes per stage, nstage stages
__global__ void __launch_bounds__(288) bulk_sum(const double* __restrict__ p, size_t n, int tile_doubles, int nstage,
                                                double* out) {
  extern __shared__ __align__(128) unsigned char raw[];
  __shared__ __align__(8) unsigned long long full[16], empty[16];
  double* ring = reinterpret_cast<double*>(raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  const size_t ntile = n / tile_doubles;
  if (tid == 0) {
    for (int i = 0; i < nstage; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[i])), "r"(1) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[i])), "r"(8) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 8) {
    if (tid == 256) {
      int it = 0;
      for (size_t t = blockIdx.x; t < ntile; t += gridDim.x, ++it) {
        const int slot = it % nstage, k = it / nstage;
        if (it >= nstage) mbar_wait(smem_u32(&empty[slot]), (unsigned)((k - 1) & 1));
        const unsigned bytes = (unsigned)(tile_doubles * sizeof(double));
        const unsigned bar = smem_u32(&full[slot]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(ring + (size_t)slot * tile_doubles)), "l"(p + t * tile_doubles), "r"(bytes), "r"(bar)
                     : "memory");
      }
    }
    return;
  }
  double acc = 0.0;
  int it = 0;
  for (size_t t = blockIdx.x; t < ntile; t += gridDim.x, ++it) {
    const int slot = it % nstage, k = it / nstage;
    mbar_wait(smem_u32(&full[slot]), (unsigned)(k & 1));
    const double* tile = ring + (size_t)slot * tile_doubles;
    for (int e = tid; e < tile_doubles; e += 256) acc += tile[e];
    __syncwarp();
    if ((tid & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[slot])) : "memory");
  }
  if (acc == 12345.678) out[0] = acc;
}

template <typename F>
static float time_ms(F launch) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    best = ms < best ? ms : best;
  }
  return best;
}

int main(int argc, char** argv) {
  const double gib = argc > 1 ? atof(argv[1]) : 2.0;
  const size_t n = (size_t)(gib * (1ull << 30)) / sizeof(double);
  double *p, *out;
  CK(cudaMalloc(&p, n * sizeof(double)));
  CK(cudaMalloc(&out, 64));
  CK(cudaMemset(p, 0, n * sizeof(double)));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("%s, %d SMs, buffer %.1f GiB\n", prop.name, sms, gib);
  const double gb = n * sizeof(double) / 1e9;
  for (int cps : {2, 4, 8}) {
    float t1 = time_ms([&] { ldg_sum<1><<<sms * cps, 256>>>(p, n, out); });
    float t4 = time_ms([&] { ldg_sum<4><<<sms * cps, 256>>>(p, n, out); });
    float t8 = time_ms([&] { ldg_sum<8><<<sms * cps, 256>>>(p, n, out); });
    printf("LDG  %d CTAs/SM x 256 thr: U=1 %.0f GB/s  U=4 %.0f GB/s  U=8 %.0f GB/s\n", cps, gb / t1 * 1e3, gb / t4 * 1e3,
           gb / t8 * 1e3);
  }
  CK(cudaFuncSetAttribute(bulk_sum, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  for (int tile_kb : {8, 16, 32, 64}) {
    for (int nstage : {2, 3, 6}) {
      for (int cps : {1, 2}) {
        const size_t smem = (size_t)tile_kb * 1024 * nstage;
        if (smem * cps > 220 * 1024 || nstage > 16) continue;
        const int td = tile_kb * 1024 / 8;
        float t = time_ms([&] { bulk_sum<<<sms * cps, 288, smem>>>(p, n, td, nstage, out); });
        printf("TMA  tile %2d KB x %d stages, %d CTA/SM (%3zu KB in flight per SM): %.0f GB/s\n", tile_kb, nstage, cps,
               smem * cps / 1024, gb / t * 1e3);
      }
    }
  }
  CK(cudaGetLastError());
  return 0;
}
