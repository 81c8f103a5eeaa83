// Microbenchmark: the MMA stream of k_tc_gemm in isolation -- 256 resident "user" rows (two 128-row A blocks, K = 128 in
// two 64-element swizzle atoms), a ring of two 128-row "item" tiles (B), four 128-column accumulators (2 buffers x 2
// blocks), 16 tcgen05.mma M128 N128 K16 + 2 commits per tile, no TMA, no epilogue.  Variants add what the real kernel has
// around that stream, one thing at a time, to see what takes the MMA from 64 to ~100-130 cycles:
//   V0 plain stream          V1 + 16 warps parked on an mbarrier (try_wait spin, like the idle epilogue)
//   V2 + 16 warps reading the accumulators (tcgen05.ld 32x32b.x32, one 64-column strip per warp per tile)
//   V3 + 16 warps streaming 16-byte shared-memory stores/loads (epilogue scratch traffic)
//   V4 N = 256 per MMA (one 256-row B tile per instruction, same flops per tile: 8 MMAs of 128 cycles)
//   V5 two commits per tile    V6 the kernel's accumulator handshake (16 warps wait tfull, fence, arrive tempty; the issuer
//   waits tempty)    V7 = V6 + the producer's ring handshake (empty -> full) with back-off waits, still no TMA
//   V8 = V7 + the epilogue warps also issue one 8-byte global load and one 16-byte cp.async per tile
// The issuing thread is chosen with elect.sync: under a plain `lane == 0` branch the compiler wraps EVERY tcgen05.mma in an
// ELECT / BRA.U.ANY loop (173 of them in this file) and the issuing thread, not the tensor pipe, sets the pace.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mma_pattern_bench mma_pattern_bench.cu
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
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(1u)
      : "memory");
}
__device__ __forceinline__ void commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WD;\n\tbra WL;\n\tWD:\n\t}" ::"r"(
          smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity) {
  for (;;) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (done) break;
    __nanosleep(64);
  }
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

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
constexpr int THREADS = 64 + 16 * 32;

template <int V>
__global__ void __launch_bounds__(THREADS, 1) k_pat(int tiles, long long *cycles, uint32_t *sink, uint32_t seed) {
  extern __shared__ uint8_t raw[];
  uint8_t *smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t *sU = smem;                      // [2 katoms][256 rows][128 B]
  uint8_t *sV = smem + 64 * 1024;          // [2 stages][2 katoms][128 rows][128 B]   (V4: one 256-row tile = both stages)
  uint8_t *scr = smem + 128 * 1024;        // 16 warps x 4.5 KB scratch
  __shared__ uint64_t tfull[2], park, tempty[2], full[2], empty[2];
  __shared__ __align__(16) float bias_s[16 * 64];
  __shared__ uint32_t slot;
  __shared__ int done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // pseudo-random fp16 bit patterns with small exponents (finite values)
  for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) {
    uint32_t x = (uint32_t)i * 2654435761u + seed;
    x ^= x >> 15;
    reinterpret_cast<uint32_t *>(smem)[i] = (x & 0x83FF83FFu) | 0x30003000u;
  }
  if (threadIdx.x == 0) {
    done = 0;
    for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(tfull + b)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&park)));
    for (int b = 0; b < 2; ++b) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 16;" ::"r"(smem_u32(tempty + b)));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(full + b)));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(empty + b)));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (warp == 1 && elect_one()) {
    const uint32_t idesc128 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc256 = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const long long t0 = clock64();
    for (int i = 0; i < tiles; ++i) {
      const int s = i & 1, b = i & 1;
      if (V >= 6) {
        if (i >= 2) mbar_wait_backoff(tempty + b, (uint32_t)((i - 2) >> 1) & 1u);
        if (V >= 7) mbar_wait_backoff(full + s, (uint32_t)(i >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      } else if (i >= 2) mbar_wait(tfull + b, (uint32_t)((i - 2) >> 1) & 1u);   // the buffer's previous MMAs have retired
      if (V == 4) {
        // 128 users x 256 items per instruction: tile i covers user block (i & 1) and the whole 256-row B
        const uint32_t d = tmem + (uint32_t)(b * 256);
        for (int ka = 0; ka < 2; ++ka) {
          const uint64_t ud = desc_sw128(smem_u32(sU + ka * 256 * 128 + (i & 1) * 128 * 128));
          const uint64_t vd = desc_sw128(smem_u32(sV + ka * 256 * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) mma(d, ud + (uint64_t)(2 * k), vd + (uint64_t)(2 * k), idesc256);
        }
      } else {
#pragma unroll
        for (int ub = 0; ub < 2; ++ub) {
          const uint32_t d = tmem + (uint32_t)(b * 256 + ub * 128);
          for (int ka = 0; ka < 2; ++ka) {
            const uint64_t ud = desc_sw128(smem_u32(sU + ka * 256 * 128 + ub * 128 * 128));
            const uint64_t vd = desc_sw128(smem_u32(sV + s * 32 * 1024 + ka * 128 * 128));
#pragma unroll
            for (int k = 0; k < 4; ++k) mma(d, ud + (uint64_t)(2 * k), vd + (uint64_t)(2 * k), idesc128);
          }
        }
      }
      if (V == 5 || V >= 7) commit(empty + s);
      commit(tfull + b);
    }
    mbar_wait(tfull + ((tiles - 1) & 1), (uint32_t)((tiles - 1) >> 1) & 1u);
    const long long t1 = clock64();
    if (blockIdx.x == 0) cycles[0] = t1 - t0;
    *((volatile int *)&done) = 1;
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&park)) : "memory");
  } else if (warp == 0 && lane == 0 && V >= 7) {
    for (int i = 0; i < tiles; ++i) {
      const int s = i & 1;
      mbar_wait_backoff(empty + s, ((uint32_t)(i >> 1) & 1u) ^ 1u);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(full + s)) : "memory");
    }
  } else if (warp >= 2) {
    if (V >= 6) {
      const float *g = reinterpret_cast<const float *>(sink);
      float accf = 0.f;
      for (int i = 0; i < tiles; ++i) {
        const int b = i & 1;
        if (V >= 8) {
          if (lane < 16)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(bias_s + (warp - 2) * 64 + lane * 4)), "l"(g + 64 + ((i * 64 + lane * 4) & 1023)) : "memory");
          asm volatile("cp.async.commit_group;" ::: "memory");
          accf += __ldg(g + 2048 + ((i * 512 + threadIdx.x) & 4095));
        }
        mbar_wait(tfull + b, (uint32_t)(i >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (V >= 8) asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(tempty + b)) : "memory");
      }
      if (accf == 123.456f) sink[1] = 1;
    } else if (V == 1) {
      mbar_wait(&park, 0);
    } else if (V == 2) {
      const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32)) << 16;
      const int e = (warp - 2) >> 2;
      uint32_t acc = 0;
      int n = 0;
      while (*((volatile int *)&done) == 0) {
        uint32_t r[32];
        LD_X32(r, tmem + lane_addr + (uint32_t)(((n & 1) * 256) + e * 64 + ((n >> 1) & 1) * 32));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 32; ++c) acc += r[c] >> 7;
        ++n;
        __nanosleep(200);   // ~ the real epilogue's pace: two strips per ~2000-cycle tile
      }
      if (acc == 0x1234567u) sink[0] = acc;
      if (blockIdx.x == 0 && threadIdx.x == 64) cycles[1] = n;
    } else if (V == 3) {
      uint4 *row = reinterpret_cast<uint4 *>(scr + (warp - 2) * 4608 + lane * 144);
      uint4 v = make_uint4(lane, warp, 3, 4);
      int n = 0;
      while (*((volatile int *)&done) == 0) {
#pragma unroll
        for (int c = 0; c < 8; ++c) row[c] = v;
        __syncwarp();
        v.x += row[(n & 7)].y;
        ++n;
      }
      if (v.x == 0x1234567u) sink[0] = v.x;
      if (blockIdx.x == 0 && threadIdx.x == 64) cycles[1] = n;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

template <int V>
void run(const char *what) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long *cyc;
  uint32_t *sink;
  cudaMalloc(&cyc, 16);
  cudaMalloc(&sink, 65536);
  cudaMemset(sink, 0, 65536);
  cudaMemset(cyc, 0, 16);
  const int smem = 1024 + 128 * 1024 + 16 * 4608, tiles = 4000;
  cudaFuncSetAttribute(k_pat<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_pat<V><<<sms, THREADS, smem>>>(50, cyc, sink, 1u);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k_pat<V><<<sms, THREADS, smem>>>(tiles, cyc, sink, 7u);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  long long h[2] = {0, 0};
  cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
  const double flops = (double)tiles * 2.0 * 256 * 128 * 128 * sms;
  printf("V%d %-52s %7.1f cycles per 128x128x16 of MMA work (clock64)  %.3f ms  %.0f TFLOP/s  side-loop iterations per tile %.2f (%s)\n",
         V, what, (double)h[0] / ((double)tiles * 16), ms, flops / (ms * 1e-3) / 1e12, (double)h[1] / tiles,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(cyc);
  cudaFree(sink);
}

int main() {
  run<0>("plain MMA stream (2 A blocks, 2 B stages, 4 accum)");
  run<1>("+ 16 warps parked on an mbarrier");
  run<2>("+ 16 warps reading accumulators (tcgen05.ld)");
  run<3>("+ 16 warps streaming smem scratch stores");
  run<4>("N = 256 per MMA");
  run<5>("two commits per tile");
  run<6>("accumulator handshake with 16 warps (tfull -> tempty)");
  run<7>("+ ring handshake with a producer thread (empty -> full)");
  run<8>("+ per-tile global load and cp.async in the 16 warps");
  return 0;
}
