// Microbenchmark of the Gram inner loop in isolation: fragments from shared memory (pitch 132),
// optional row-weight multiply, 4x4 DMMA blocks per warp, 16 warps per CTA, one CTA per SM.
// Variants: 0 = LDS + DMUL + DMMA (as in the kernel)   1 = LDS + DMMA (no weighting)
//           2 = software-pipelined fragments (next k-step loaded before the DMMAs of this one)
//           3 = variant 0 with a __syncthreads every 8 k-steps (stage boundary)
#include <cstdio>
#include <cuda_runtime.h>
constexpr int GLDS = 132, RCH = 32;
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int VAR>
__global__ void __launch_bounds__(512, 1) inner(double* out, int iters) {
  extern __shared__ double sm[];
  double* sI = sm; double* sJ = sm + RCH * GLDS; double* sW = sJ + RCH * GLDS;
  for (int i = threadIdx.x; i < 2 * RCH * GLDS + RCH; i += blockDim.x) sm[i] = 1.0 + 1e-3 * (i % 97);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wr = warp >> 2, wc = warp & 3;
  const double* fI = sI + (lane & 3) * GLDS + (lane >> 2) + wr * 32;
  const double* fJ = sJ + (lane & 3) * GLDS + (lane >> 2) + wc * 32;
  const double* fW = sW + (lane & 3);
  double acc[4][4][2];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
    if (VAR == 2) {
      double af[4], bf[4], an[4], bn[4];
      { const double wv = fW[0];
#pragma unroll
        for (int i = 0; i < 4; ++i) { af[i] = fI[i * 8] * wv; bf[i] = fJ[i * 8] * wv; } }
#pragma unroll
      for (int ks = 0; ks < RCH / 4; ++ks) {
        if (ks + 1 < RCH / 4) {
          const double wv = fW[(ks + 1) * 4];
#pragma unroll
          for (int i = 0; i < 4; ++i) { an[i] = fI[(ks + 1) * 4 * GLDS + i * 8] * wv; bn[i] = fJ[(ks + 1) * 4 * GLDS + i * 8] * wv; }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
#pragma unroll
        for (int i = 0; i < 4; ++i) { af[i] = an[i]; bf[i] = bn[i]; }
      }
    } else {
#pragma unroll 2
      for (int ks = 0; ks < RCH / 4; ++ks) {
        const double wv = fW[ks * 4];
        double af[4], bf[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          af[i] = fI[ks * 4 * GLDS + i * 8]; bf[i] = fJ[ks * 4 * GLDS + i * 8];
          if (VAR != 1) { af[i] *= wv; bf[i] *= wv; }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
      if (VAR == 3) __syncthreads();
    }
  }
  double s = 0;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j][0] + acc[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int VAR>
void run(int sms, double* out) {
  const int iters = 2000; const size_t smem = (2 * RCH * GLDS + RCH) * sizeof(double);
  cudaFuncSetAttribute(inner<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  inner<VAR><<<sms, 512, smem>>>(out, iters); cudaDeviceSynchronize();
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  cudaEventRecord(s); inner<VAR><<<sms, 512, smem>>>(out, iters); cudaEventRecord(e); cudaEventSynchronize(e);
  float ms; cudaEventElapsedTime(&ms, s, e);
  double flops = (double)sms * 16 * iters * 8 * 16 * 512.0;
  printf("variant %d: %.3f ms  %.2f TFLOP/s (%s)\n", VAR, ms, flops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 512);
  run<0>(sms, out); run<1>(sms, out); run<2>(sms, out); run<3>(sms, out);
  return 0;
}
