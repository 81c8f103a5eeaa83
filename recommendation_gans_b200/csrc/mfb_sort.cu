// Stable LSD radix sort of (uint32 key, uint32 value) pairs -- the "planner" of the training
// step.  The reference reduces duplicate rows of a batch inside torch's embedding_dense_backward
// (index_add in batch order, implicit.py:361); here the rows of many steps are sorted at once by
// (step, table, row) so each unique row's contributions form one contiguous, batch-ordered
// segment and can be reduced deterministically without atomics.
#include "mfb_internal.cuh"

namespace {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int ITEMS_PER_WARP_ITER = 32;
constexpr int ITERS = 16;                                  // keys per thread
constexpr int TILE = SORT_THREADS * ITERS;                 // keys per block
constexpr int WARP_SPAN = ITEMS_PER_WARP_ITER * ITERS;     // contiguous keys owned by a warp

__global__ void __launch_bounds__(SORT_THREADS) k_radix_hist(const uint32_t *__restrict__ keys, long long n,
                                                             int shift, uint32_t *__restrict__ hist, int nblk) {
  __shared__ uint32_t h[RADIX];
  for (int i = threadIdx.x; i < RADIX; i += SORT_THREADS) h[i] = 0;
  __syncthreads();
  long long base = (long long)blockIdx.x * TILE;
  for (int it = 0; it < ITERS; ++it) {
    long long i = base + (long long)it * SORT_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & (RADIX - 1)], 1u);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < RADIX; d += SORT_THREADS) hist[(long long)d * nblk + blockIdx.x] = h[d];
}

// Per-digit exclusive scan across blocks: CTA d scans hist[d*nblk .. (d+1)*nblk) in place and writes
// the digit's total; the scatter kernel turns the 256 totals into digit bases itself.
__global__ void __launch_bounds__(1024) k_radix_scan(uint32_t *__restrict__ hist, int nblk,
                                                     uint32_t *__restrict__ totals) {
  __shared__ uint32_t warp_tot[32], warp_excl[32];
  __shared__ uint32_t carry_s, chunk_total;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  uint32_t *h = hist + (long long)blockIdx.x * nblk;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int start = 0; start < nblk; start += 1024) {
    const int i = start + tid;
    const uint32_t v = (i < nblk) ? h[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[wid] = x;
    __syncthreads();
    if (wid == 0) {  // scan the 32 warp totals
      const uint32_t w = warp_tot[lane];
      uint32_t ws = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, ws, o);
        if (lane >= o) ws += y;
      }
      warp_excl[lane] = ws - w;
      if (lane == 31) chunk_total = ws;
    }
    __syncthreads();
    if (i < nblk) h[i] = carry_s + warp_excl[wid] + x - v;
    __syncthreads();
    if (tid == 0) carry_s += chunk_total;
    __syncthreads();
  }
  if (tid == 0) totals[blockIdx.x] = carry_s;
}

// Each warp owns a contiguous span of the block's tile and walks it in order, so ranks are
// stable: rank = (global digit base for this block) + (same-digit keys of earlier warps)
//             + (same-digit keys earlier in this warp's span).
__global__ void __launch_bounds__(SORT_THREADS) k_radix_scatter(const uint32_t *__restrict__ keys_in,
                                                                const uint32_t *__restrict__ vals_in,
                                                                uint32_t *__restrict__ keys_out,
                                                                uint32_t *__restrict__ vals_out, long long n, int shift,
                                                                const uint32_t *__restrict__ hist, int nblk,
                                                                const uint32_t *__restrict__ totals) {
  __shared__ uint32_t cnt[SORT_WARPS][RADIX];
  __shared__ uint32_t dbase[RADIX];
  __shared__ uint32_t wtot[SORT_WARPS];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int i = tid; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&cnt[0][0])[i] = 0;
  {  // exclusive scan of the 256 digit totals (SORT_THREADS == RADIX)
    const uint32_t v = totals[tid];
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wtot[wid] = x;
    __syncthreads();
    uint32_t before = 0;
    for (int w = 0; w < wid; ++w) before += wtot[w];
    dbase[tid] = before + x - v;
  }
  __syncthreads();
  const long long wbase = (long long)blockIdx.x * TILE + (long long)wid * WARP_SPAN;
  // pass 1: per-warp digit counts
  for (int it = 0; it < ITERS; ++it) {
    long long i = wbase + it * 32 + lane;
    uint32_t d = (i < n) ? ((keys_in[i] >> shift) & (RADIX - 1)) : (uint32_t)RADIX;  // RADIX = invalid
    unsigned peers = __match_any_sync(0xffffffffu, d);
    if (d < RADIX && lane == (__ffs(peers) - 1)) cnt[wid][d] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // exclusive prefix over warps, seeded with the block's global base for each digit
  for (int d = tid; d < RADIX; d += SORT_THREADS) {
    uint32_t run = dbase[d] + hist[(long long)d * nblk + blockIdx.x];
    for (int w = 0; w < SORT_WARPS; ++w) {
      uint32_t c = cnt[w][d];
      cnt[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
  // pass 2: scatter in the same order
  for (int it = 0; it < ITERS; ++it) {
    long long i = wbase + it * 32 + lane;
    bool valid = i < n;
    uint32_t key = valid ? keys_in[i] : 0u;
    uint32_t val = valid ? vals_in[i] : 0u;
    uint32_t d = valid ? ((key >> shift) & (RADIX - 1)) : (uint32_t)RADIX;
    unsigned peers = __match_any_sync(0xffffffffu, d);
    int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (valid && lane == leader) base = cnt[wid][d];
    base = __shfl_sync(0xffffffffu, base, leader);
    uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    if (valid) {
      keys_out[base + rank] = key;
      vals_out[base + rank] = val;
    }
    __syncwarp();
    if (valid && lane == leader) cnt[wid][d] = base + __popc(peers);
    __syncwarp();
  }
}

}  // namespace

int mfb_radix_sort_pairs(uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b, int64_t n,
                         int nbits, DevBuf &hist, uint32_t **out_keys, uint32_t **out_vals, cudaStream_t st) {
  *out_keys = keys_a;
  *out_vals = vals_a;
  if (n <= 1 || nbits <= 0) return MFB_OK;
  if (n >= (1ll << 32)) {
    mfb_set_error("radix sort: too many keys (%lld)", (long long)n);
    return MFB_ERR_INVALID;
  }
  int nblk = (int)((n + TILE - 1) / TILE);
  long long total = (long long)RADIX * nblk + RADIX;
  MFB_CHECK(hist.reserve((size_t)total * sizeof(uint32_t)));
  uint32_t *totals = hist.as<uint32_t>() + (long long)RADIX * nblk;
  static_assert(SORT_THREADS == RADIX, "scatter scans the digit totals with one thread per digit");
  uint32_t *kin = keys_a, *vin = vals_a, *kout = keys_b, *vout = vals_b;
  for (int shift = 0; shift < nbits; shift += RADIX_BITS) {
    k_radix_hist<<<nblk, SORT_THREADS, 0, st>>>(kin, n, shift, hist.as<uint32_t>(), nblk);
    MFB_KERNEL_CHECK();
    k_radix_scan<<<RADIX, 1024, 0, st>>>(hist.as<uint32_t>(), nblk, totals);
    MFB_KERNEL_CHECK();
    k_radix_scatter<<<nblk, SORT_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift, hist.as<uint32_t>(), nblk, totals);
    MFB_KERNEL_CHECK();
    uint32_t *t = kin;
    kin = kout;
    kout = t;
    t = vin;
    vin = vout;
    vout = t;
  }
  *out_keys = kin;
  *out_vals = vin;
  return MFB_OK;
}
