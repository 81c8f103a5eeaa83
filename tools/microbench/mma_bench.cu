// Microbenchmark: issue rate of tcgen05.mma kind::f16 (fp16 x fp16 -> fp32 in TMEM) with cta_group::1 on sm_100a.
// One CTA per SM, operands are whatever shared memory holds (zeros), one thread issues back-to-back MMAs.
// Prints cycles per instruction and the dense TFLOP/s that issue rate corresponds to.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mma_bench mma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int N>
__global__ void __launch_bounds__(128, 1) k_mma(int iters, long long *cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t *smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t adesc = desc_sw128(smem_u32(smem)), bdesc = desc_sw128(smem_u32(smem + 32 * 1024));
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t acc = 1u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + (uint32_t)((i & 1) * 256)),
            "l"(adesc + (uint64_t)(2 * k)), "l"(bdesc + (uint64_t)(2 * k)), "r"(idesc), "r"(acc)
            : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile(
        "{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(
            smem_u32(&bar))
        : "memory");
    const long long t1 = clock64();
    if (blockIdx.x == 0) cycles[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

template <int N>
void run() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long *cyc;
  cudaMalloc(&cyc, 8);
  const int smem = 66 * 1024, iters = 20000;
  cudaFuncSetAttribute(k_mma<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_mma<N><<<sms, 128, smem>>>(100, cyc);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k_mma<N><<<sms, 128, smem>>>(iters, cyc);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  long long h = 0;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double n_mma = (double)iters * 4;
  const double flops = n_mma * 2.0 * 128 * N * 16 * sms;
  printf("M=128 N=%d K=16 cta_group::1: %.1f cycles per tcgen05.mma (clock64), %.3f ms, %.0f TFLOP/s over %d SMs (%s)\n", N,
         (double)h / n_mma, ms, flops / (ms * 1e-3) / 1e12, sms, cudaGetErrorString(cudaGetLastError()));
  cudaFree(cyc);
}

int main() {
  run<128>();
  run<256>();
  run<64>();
  return 0;
}
