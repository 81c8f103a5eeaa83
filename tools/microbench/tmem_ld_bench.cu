// Microbenchmark: read bandwidth of tensor memory (tcgen05.ld) on sm_100a, per SM.
// One CTA per SM allocates all 512 TMEM columns; W warps (warp w reads the lane quarter w % 4, as the hardware
// requires) issue back-to-back tcgen05.ld.32x32b.xN over the columns and consume the registers with a cheap
// xor so the loads cannot be dropped.  Prints bytes per clock per SM for every (W, N, loads per wait).
// Optionally the tensor pipe is kept busy by another warp issuing tcgen05.mma (MMA=1) to see whether the
// accumulator writes and the read-out share a port.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tmem_ld_bench tmem_ld_bench.cu
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

#define LD_X32(r, taddr)                                                                                             \
  asm volatile(                                                                                                      \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                      \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "   \
      "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                                                         \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),  \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),       \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),      \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                    \
      : "r"(taddr)                                                                                                   \
      : "memory")

#define LD_X16(r, taddr)                                                                                             \
  asm volatile(                                                                                                      \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                                      \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"                               \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),  \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                     \
      : "r"(taddr)                                                                                                   \
      : "memory")

// 16x256b.x8: 16 lanes x 8 x 256 bits = the same 32 registers per thread, other access shape
#define LD_16x256_X8(r, taddr)                                                                                       \
  asm volatile(                                                                                                      \
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "                                                                      \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "   \
      "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                                                         \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),  \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),       \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),      \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                    \
      : "r"(taddr)                                                                                                   \
      : "memory")

// SHAPE: 0 = 32x32b.x32 (4 KB per warp instruction), 1 = 32x32b.x16 (2 KB), 2 = 16x256b.x8 (4 KB)
// PER_WAIT: loads issued before each tcgen05.wait::ld
template <int SHAPE, int PER_WAIT, bool MMA>
__global__ void __launch_bounds__(544, 1) k_ld(int iters, int ld_warps, long long *cycles, uint32_t *sink) {
  extern __shared__ uint8_t raw[];
  uint8_t *smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ int stop;
  const int warp = threadIdx.x >> 5;
  if (MMA)
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0u;
  if (threadIdx.x == 0) {
    stop = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  const int lw = warp - 1;   // warp 0 = TMEM owner / MMA issuer; loaders from warp 1 on
  if (MMA && threadIdx.x == 0) {
    // keep the tensor pipe saturated while the others read (results are garbage, timing is what matters)
    const uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t adesc = desc_sw128(smem_u32(smem)), bdesc = desc_sw128(smem_u32(smem + 32 * 1024));
    int n = 0;
    while (*((volatile int *)&stop) < ld_warps) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + (uint32_t)((n & 3) * 128)),
            "l"(adesc + (uint64_t)(2 * (k & 3))), "l"(bdesc + (uint64_t)(2 * (k & 3))), "r"(idesc), "r"(1u)
            : "memory");
      }
      ++n;
      if ((n & 7) == 0) {   // bound the queue: wait for what was issued
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        const uint32_t ph = (uint32_t)((n >> 3) - 1) & 1u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tW1:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D1;\n\tbra W1;\n\tD1:\n\t}" ::"r"(
                smem_u32(&bar)), "r"(ph)
            : "memory");
      }
    }
    if (blockIdx.x == 0) cycles[1] = n * 8;
  }
  if (lw >= 0 && lw < ld_warps) {
    const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32)) << 16;
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      uint32_t r[PER_WAIT][32];
#pragma unroll
      for (int j = 0; j < PER_WAIT; ++j) {
        const uint32_t col = (uint32_t)(((i * PER_WAIT + j) * 32 + lw * 64) & 511);
        if (SHAPE == 0) LD_X32(r[j], tmem + lane_addr + col);
        else if (SHAPE == 1) {
          LD_X16(r[j], tmem + lane_addr + (col & ~15u));
#pragma unroll
          for (int c = 16; c < 32; ++c) r[j][c] = 0;
        } else LD_16x256_X8(r[j], tmem + lane_addr + (col & ~63u));
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < PER_WAIT; ++j)
#pragma unroll
        for (int c = 0; c < 32; c += 8) acc ^= r[j][c];
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const long long t1 = clock64();
    if (blockIdx.x == 0 && lw == 0 && (threadIdx.x & 31) == 0) cycles[0] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    if ((threadIdx.x & 31) == 0) atomicAdd(&stop, 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

template <int SHAPE, int PER_WAIT, bool MMA>
void run(int ld_warps) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long *cyc;
  uint32_t *sink;
  cudaMalloc(&cyc, 16);
  cudaMalloc(&sink, 4);
  cudaMemset(cyc, 0, 16);
  const int smem = 66 * 1024, iters = 20000 / PER_WAIT;
  cudaFuncSetAttribute(k_ld<SHAPE, PER_WAIT, MMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_ld<SHAPE, PER_WAIT, MMA><<<sms, 544, smem>>>(100, ld_warps, cyc, sink);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k_ld<SHAPE, PER_WAIT, MMA><<<sms, 544, smem>>>(iters, ld_warps, cyc, sink);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  long long h[2] = {0, 0};
  cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
  const double bytes_per_ld = (SHAPE == 1) ? 2048.0 : 4096.0;
  const double bytes = (double)iters * PER_WAIT * bytes_per_ld * ld_warps;
  const char *shape = SHAPE == 0 ? "32x32b.x32" : (SHAPE == 1 ? "32x32b.x16" : "16x256b.x8");
  printf("%s  warps %2d  loads/wait %d  mma %d: %7.1f B/clk/SM (clock64)  %6.1f cycles per warp-load  %.3f ms", shape, ld_warps,
         PER_WAIT, (int)MMA, bytes / (double)h[0], (double)h[0] / ((double)iters * PER_WAIT), ms);
  if (MMA) printf("  mma: %.1f cycles each", (double)h[0] / (double)h[1]);
  printf("  (%s)\n", cudaGetErrorString(cudaGetLastError()));
  cudaFree(cyc);
  cudaFree(sink);
}

int main() {
  for (int w : {1, 4, 8, 16}) run<0, 1, false>(w);
  for (int w : {4, 8, 16}) run<0, 2, false>(w);
  for (int w : {4, 16}) run<1, 2, false>(w);
  for (int w : {4, 16}) run<2, 2, false>(w);
  for (int w : {4, 8, 16}) run<0, 2, true>(w);
  return 0;
}
