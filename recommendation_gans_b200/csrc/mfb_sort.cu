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
// keys per thread: 16 for bulk sorts (fewer blocks, smaller histograms); 4 for small inputs, where a kernel's latency
// is the serial work of one block (a short call's first plan sits on the training step's critical path)
constexpr int ITERS_BULK = 16, ITERS_SMALL = 4;
constexpr long long SMALL_SORT_KEYS = 1ll << 20;

template <int ITERS>
__global__ void __launch_bounds__(SORT_THREADS) k_radix_hist(const uint32_t *__restrict__ keys, long long n,
                                                             int shift, uint32_t *__restrict__ hist, int nblk) {
  constexpr int TILE = SORT_THREADS * ITERS;                 // keys per block
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
// The pairs are first placed into shared memory in digit order (block-local rank) and then streamed out: consecutive
// threads write consecutive addresses inside each digit's run, instead of 32 scattered 4-byte stores per warp
// instruction (which cost a 32-byte sector each).
template <int ITERS>
__global__ void __launch_bounds__(SORT_THREADS) k_radix_scatter(const uint32_t *__restrict__ keys_in,
                                                                const uint32_t *__restrict__ vals_in,
                                                                uint32_t *__restrict__ keys_out,
                                                                uint32_t *__restrict__ vals_out, long long n, int shift,
                                                                const uint32_t *__restrict__ hist, int nblk,
                                                                const uint32_t *__restrict__ totals) {
  constexpr int TILE = SORT_THREADS * ITERS;                 // keys per block
  constexpr int WARP_SPAN = ITEMS_PER_WARP_ITER * ITERS;     // contiguous keys owned by a warp
  __shared__ uint32_t cnt[SORT_WARPS][RADIX];
  __shared__ uint32_t gbase[RADIX];     // global position of this block's first key of digit d
  __shared__ uint32_t lstart[RADIX];    // block-local position of the first key of digit d
  __shared__ uint32_t wtot[SORT_WARPS];
  __shared__ uint32_t sk[TILE], sv[TILE];
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
    gbase[tid] = before + x - v + hist[(long long)tid * nblk + blockIdx.x];
  }
  __syncthreads();
  const long long wbase = (long long)blockIdx.x * TILE + (long long)wid * WARP_SPAN;
  // pass 1: per-warp digit counts
  uint32_t kreg[ITERS], vreg[ITERS];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    long long i = wbase + it * 32 + lane;
    const bool valid = i < n;
    kreg[it] = valid ? keys_in[i] : 0u;
    vreg[it] = valid ? vals_in[i] : 0u;
  }
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    long long i = wbase + it * 32 + lane;
    uint32_t d = (i < n) ? ((kreg[it] >> shift) & (RADIX - 1)) : (uint32_t)RADIX;  // RADIX = invalid
    unsigned peers = __match_any_sync(0xffffffffu, d);
    if (d < RADIX && lane == (__ffs(peers) - 1)) cnt[wid][d] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // per digit: exclusive prefix over warps (block-local), the digit's block total, then a scan of the totals over digits
  {
    const int d = tid;
    uint32_t run = 0;
    for (int w = 0; w < SORT_WARPS; ++w) {
      uint32_t c = cnt[w][d];
      cnt[w][d] = run;
      run += c;
    }
    uint32_t x = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    __syncthreads();            // wtot is reused
    if (lane == 31) wtot[wid] = x;
    __syncthreads();
    uint32_t before = 0;
    for (int w = 0; w < wid; ++w) before += wtot[w];
    lstart[d] = before + x - run;
  }
  __syncthreads();
  // pass 2: place every pair at its block-local sorted position
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    long long i = wbase + it * 32 + lane;
    const bool valid = i < n;
    const uint32_t key = kreg[it];
    uint32_t d = valid ? ((key >> shift) & (RADIX - 1)) : (uint32_t)RADIX;
    unsigned peers = __match_any_sync(0xffffffffu, d);
    int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (valid && lane == leader) base = cnt[wid][d];
    base = __shfl_sync(0xffffffffu, base, leader);
    uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    if (valid) {
      const uint32_t at = lstart[d] + base + rank;
      sk[at] = key;
      sv[at] = vreg[it];
    }
    __syncwarp();
    if (valid && lane == leader) cnt[wid][d] = base + __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // stream out: position i of the block-local order goes to gbase[digit] + (i - lstart[digit])
  const long long left = n - (long long)blockIdx.x * TILE;
  const int count = left < TILE ? (int)left : TILE;
  for (int i = tid; i < count; i += SORT_THREADS) {
    const uint32_t key = sk[i];
    const uint32_t d = (key >> shift) & (RADIX - 1);
    const uint32_t at = gbase[d] + ((uint32_t)i - lstart[d]);
    keys_out[at] = key;
    vals_out[at] = sv[i];
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
  const bool small = n < SMALL_SORT_KEYS;
  const int tile = SORT_THREADS * (small ? ITERS_SMALL : ITERS_BULK);
  int nblk = (int)((n + tile - 1) / tile);
  long long total = (long long)RADIX * nblk + RADIX;
  MFB_CHECK(hist.reserve((size_t)total * sizeof(uint32_t)));
  uint32_t *totals = hist.as<uint32_t>() + (long long)RADIX * nblk;
  static_assert(SORT_THREADS == RADIX, "scatter scans the digit totals with one thread per digit");
  uint32_t *kin = keys_a, *vin = vals_a, *kout = keys_b, *vout = vals_b;
  for (int shift = 0; shift < nbits; shift += RADIX_BITS) {
    if (small) k_radix_hist<ITERS_SMALL><<<nblk, SORT_THREADS, 0, st>>>(kin, n, shift, hist.as<uint32_t>(), nblk);
    else k_radix_hist<ITERS_BULK><<<nblk, SORT_THREADS, 0, st>>>(kin, n, shift, hist.as<uint32_t>(), nblk);
    MFB_KERNEL_CHECK();
    k_radix_scan<<<RADIX, 1024, 0, st>>>(hist.as<uint32_t>(), nblk, totals);
    MFB_KERNEL_CHECK();
    if (small)
      k_radix_scatter<ITERS_SMALL><<<nblk, SORT_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift, hist.as<uint32_t>(), nblk,
                                                                 totals);
    else
      k_radix_scatter<ITERS_BULK><<<nblk, SORT_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift, hist.as<uint32_t>(), nblk,
                                                                totals);
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
