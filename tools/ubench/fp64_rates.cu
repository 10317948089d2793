// Microbenchmark: peak issue rate of DMMA.8x8x4 vs DFMA on this GPU (register-resident loops).
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void dmma_loop(double* out, int iters) {
  double c[ILP][2];
  double a = threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-4;
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void dfma_loop(double* out, int iters) {
  double c[ILP];
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename K>
float run(K kern, int blocks, int threads, double* out, int iters) {
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  kern<<<blocks, threads>>>(out, iters); cudaDeviceSynchronize();
  cudaEventRecord(s); kern<<<blocks, threads>>>(out, iters); cudaEventRecord(e); cudaEventSynchronize(e);
  float ms; cudaEventElapsedTime(&ms, s, e); return ms;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount; double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  printf("%s SMs=%d clock=%d kHz\n", p.name, sms, p.clockRate);
  const int iters = 20000;
  for (int warps : {1, 2, 4, 8, 16, 32}) {
    int threads = warps * 32 > 1024 ? 1024 : warps * 32; int bps = warps * 32 / threads;
    float ms = run(dmma_loop<8>, sms * bps, threads, out, iters);
    double flops = (double)sms * warps * iters * 8 * 512.0;
    printf("DMMA ilp8 warps/SM=%2d : %.3f ms  %.2f TFLOP/s\n", warps, ms, flops / ms / 1e9);
  }
  for (int warps : {4, 8, 16, 32}) {
    int threads = warps * 32 > 1024 ? 1024 : warps * 32;
    float ms = run(dmma_loop<16>, sms, threads, out, iters);
    double flops = (double)sms * warps * iters * 16 * 512.0;
    printf("DMMA ilp16 warps/SM=%2d : %.3f ms  %.2f TFLOP/s\n", warps, ms, flops / ms / 1e9);
  }
  for (int warps : {4, 8, 16, 32}) {
    int threads = warps * 32;
    float ms = run(dfma_loop<16>, sms, threads, out, iters);
    double flops = (double)sms * warps * 32.0 * iters * 16 * 2.0;
    printf("DFMA ilp16 warps/SM=%2d : %.3f ms  %.2f TFLOP/s\n", warps, ms, flops / ms / 1e9);
  }
  return 0;
}
