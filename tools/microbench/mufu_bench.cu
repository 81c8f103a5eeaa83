// Microbenchmark: MUFU (XU pipe) throughput on sm_100a for the two approximations the dense-Adam replay uses
// (sqrt.approx.ftz.f32, rcp.approx.ftz.f32), and for the replay step itself (mfb_rowops.cuh row_replay inner loop,
// restated).  Prints lane-ops per clock per SM.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a mufu_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float sqrt_approx(float x) { float r; asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_approx(float x) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rsqrt_approx(float x) { float r; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

template <int OP, int ILP>
__global__ void k_mufu(float *out, int iters, float seed) {
  float x[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) x[j] = seed + threadIdx.x * 1e-3f + j;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      if (OP == 0) x[j] = rcp_approx(x[j]) + 1.0f;
      if (OP == 1) x[j] = sqrt_approx(x[j]) + 1.0f;
      if (OP == 2) x[j] = rsqrt_approx(x[j]) + 1.0f;
      if (OP == 3) x[j] = rcp_approx(sqrt_approx(x[j]) + 1.0f) + 1.0f;   // the replay's pair
    }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < ILP; ++j) s += x[j];
  if (s == 12345.678f) out[0] = s;
}

template <int OP, int ILP>
void run(const char *name, int mufu_per_iter) {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  float *out;
  cudaMalloc(&out, 4);
  const int iters = 20000, threads = 1024, blocks = sms * 2;
  k_mufu<OP, ILP><<<blocks, threads>>>(out, 100, 1.5f);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k_mufu<OP, ILP><<<blocks, threads>>>(out, iters, 1.5f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  const double ops = (double)blocks * threads * iters * ILP * mufu_per_iter;
  printf("%-28s %8.3f ms  %7.2f G lane-ops/s  = %5.2f lane-ops/clk/SM at %d MHz nominal\n", name, ms, ops / ms / 1e6,
         ops / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000);
  cudaFree(out);
}

int main() {
  run<0, 8>("rcp.approx.ftz", 1);
  run<1, 8>("sqrt.approx.ftz", 1);
  run<2, 8>("rsqrt.approx.ftz", 1);
  run<3, 8>("rcp(sqrt(x)+1) pair", 2);
  return 0;
}
