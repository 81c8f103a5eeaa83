// MT19937 index streams on the device (integer work, bit-exact with CPython `random` and numpy
// `RandomState`): replaces the host-side negative draws of the reference
//   implicit.py:352,370      random.choices(neg_examples, k = num_neg * batch)
//   spotlight/sampling.py:33 random_state.randint(0, num_items, shape, dtype=int64)
// The generator state travels as (624 words, position) exactly as random.getstate()[1], so the
// host RNG objects stay in sync with what the reference would have consumed.
#include <mutex>

#include "mfb_internal.cuh"
#include "mfb_mt.cuh"

namespace {

__global__ void __launch_bounds__(MT_THREADS) k_mt_generate(uint32_t *state_io, unsigned long long nwords,
                                                            uint32_t *out) {
  mt_generate_body(state_io, state_io, nwords, out, false);
}

// random.random(): a = w0 >> 5, b = w1 >> 6, x = (a*2^26 + b) / 2^53; index = floor(x * n).
__global__ void k_choices(const uint32_t *__restrict__ words, long long k, long long pop_len,
                          const long long *__restrict__ pop_u, const long long *__restrict__ pop_i,
                          long long *__restrict__ out_u, long long *__restrict__ out_i,
                          long long *__restrict__ out_idx) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= k) return;
  uint32_t a = words[2 * s] >> 5;
  uint32_t b = words[2 * s + 1] >> 6;
  double x = __dmul_rn(__dadd_rn(__dmul_rn((double)a, 67108864.0), (double)b), 1.0 / 9007199254740992.0);
  long long idx = (long long)floor(__dmul_rn(x, (double)pop_len));
  if (out_idx) out_idx[s] = idx;
  if (out_u) {
    out_u[s] = pop_u[idx];
    out_i[s] = pop_i[idx];
  }
}

// numpy legacy masked rejection: accept (w & mask) when <= rng, in stream order.
// Single CTA; writes the first `count` accepted values, the number of stream words consumed to
// obtain them (result[0]) and the total number accepted among nwords (result[1]).
__global__ void __launch_bounds__(1024) k_masked_compact(const uint32_t *__restrict__ words, long long nwords,
                                                         uint32_t mask, uint32_t rng, long long count,
                                                         long long *__restrict__ out, long long *__restrict__ result) {
  __shared__ int warp_sums[32];
  __shared__ long long base_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) {
    base_s = 0;
    result[0] = -1;
  }
  __syncthreads();
  for (long long start = 0; start < nwords; start += 1024) {
    long long i = start + tid;
    uint32_t v = 0;
    int ok = 0;
    if (i < nwords) {
      v = words[i] & mask;
      ok = (v <= rng);
    }
    unsigned bal = __ballot_sync(0xffffffffu, ok);
    int in_warp = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) warp_sums[wid] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < 32; ++w) {
      int c = warp_sums[w];
      if (w < wid) before += c;
      total += c;
    }
    long long rank = base_s + before + in_warp;
    if (ok && rank < count) {
      out[rank] = (long long)v;
      if (rank == count - 1) result[0] = i + 1;
    }
    __syncthreads();
    if (tid == 0) base_s += total;
    __syncthreads();
    if (base_s >= count) break;
  }
  if (tid == 0) result[1] = base_s;
}

DevBuf g_state, g_words, g_result, g_jump;  // library-global scratch for the model-less RNG entry points

// The model-less entry points share the scratch above: one caller at a time (ranks running as threads of one process
// call these concurrently, and ctypes releases the GIL), and the scratch follows the calling thread's device.
std::recursive_mutex g_rng_mutex;
int g_rng_device = -1;
void release_rng_scratch();
struct RngScratchGuard {
  std::lock_guard<std::recursive_mutex> lock;
  RngScratchGuard() : lock(g_rng_mutex) {
    int dev = -1;
    if (cudaGetDevice(&dev) == cudaSuccess && dev != g_rng_device) {
      if (g_rng_device >= 0) release_rng_scratch();
      g_rng_device = dev;
    }
  }
};

int upload_state(const uint32_t *h_state, cudaStream_t st) {
  MFB_CHECK(g_state.reserve(625 * sizeof(uint32_t)));
  if (h_state[624] > 624) {
    mfb_set_error("MT19937 position %u out of range", h_state[624]);
    return MFB_ERR_INVALID;
  }
  MFB_CUDA(cudaMemcpyAsync(g_state.ptr, h_state, 625 * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  return MFB_OK;
}

}  // namespace

// Device-resident stream: advances d_state (625 words) by nwords outputs, no host round trip.
int mfb_mt_generate_async(uint32_t *d_state, int64_t nwords, uint32_t *d_words, cudaStream_t st) {
  if (nwords <= 0) return MFB_OK;
  mfb_count_library_launch(1);
  k_mt_generate<<<1, MT_THREADS, 0, st>>>(d_state, (unsigned long long)nwords, d_words);
  MFB_KERNEL_CHECK();
  return MFB_OK;
}

// random.choices index mapping of 2k words -> k (user,item) pairs gathered from the population
int mfb_choices_async(const uint32_t *d_words, int64_t k, int64_t pop_len, const int64_t *d_pop_users,
                      const int64_t *d_pop_items, int64_t *d_out_users, int64_t *d_out_items, cudaStream_t st) {
  if (k <= 0) return MFB_OK;
  mfb_count_library_launch(1);
  k_choices<<<(unsigned)((k + 255) / 256), 256, 0, st>>>(d_words, k, pop_len, (const long long *)d_pop_users,
                                                         (const long long *)d_pop_items, (long long *)d_out_users,
                                                         (long long *)d_out_items, nullptr);
  MFB_KERNEL_CHECK();
  return MFB_OK;
}

// Generates nwords outputs into d_words (may be nullptr: advance only) and updates h_state.
int mfb_mt_generate(uint32_t *h_state, int64_t nwords, uint32_t *d_words, cudaStream_t st) {
  if (nwords < 0) return MFB_ERR_INVALID;
  MFB_CHECK(upload_state(h_state, st));
  if (nwords >= 8 * 65536 && d_words != nullptr) {
    // long draws (the row-sharded path draws the negatives of hundreds of steps at once): up to 64 shares by jump-ahead
    int64_t share = (nwords + 63) / 64;
    share = share < 65536 ? 65536 : ((share + 65535) / 65536) * 65536;   // few distinct sizes -> few polynomial sets
    MFB_CHECK(mfb_mt_generate_parallel(g_state.as<uint32_t>(), nwords, d_words, share, &g_jump, st));
  } else if (nwords > 0) {
    mfb_count_library_launch(1);
    k_mt_generate<<<1, MT_THREADS, 0, st>>>(g_state.as<uint32_t>(), (unsigned long long)nwords, d_words);
    MFB_KERNEL_CHECK();
  }
  MFB_CUDA(cudaMemcpyAsync(h_state, g_state.ptr, 625 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  MFB_CUDA(cudaStreamSynchronize(st));
  return MFB_OK;
}

extern "C" int mfb_mt_words(uint32_t *h_state, int64_t nwords, uint32_t *d_out, mfb_stream stream) {
  RngScratchGuard guard;
  if (!h_state || (nwords > 0 && !d_out)) return MFB_ERR_INVALID;
  return mfb_mt_generate(h_state, nwords, d_out, (cudaStream_t)stream);
}

static int choices_impl(uint32_t *h_state, const int64_t *d_pop_users, const int64_t *d_pop_items, int64_t pop_len,
                        int64_t k, int64_t *d_out_users, int64_t *d_out_items, int64_t *d_out_idx, cudaStream_t st) {
  if (!h_state || pop_len <= 0 || k < 0) {
    mfb_set_error("mt_choices: bad arguments (pop_len=%lld, k=%lld)", (long long)pop_len, (long long)k);
    return MFB_ERR_INVALID;
  }
  if (k == 0) return MFB_OK;
  MFB_CHECK(g_words.reserve((size_t)(2 * k) * sizeof(uint32_t)));
  MFB_CHECK(mfb_mt_generate(h_state, 2 * k, g_words.as<uint32_t>(), st));
  int threads = 256;
  long long blocks = (k + threads - 1) / threads;
  mfb_count_library_launch(1);
  k_choices<<<(unsigned)blocks, threads, 0, st>>>(g_words.as<uint32_t>(), k, pop_len, (const long long *)d_pop_users,
                                                  (const long long *)d_pop_items, (long long *)d_out_users,
                                                  (long long *)d_out_items, (long long *)d_out_idx);
  MFB_KERNEL_CHECK();
  return MFB_OK;
}

extern "C" int mfb_mt_choices_pairs(uint32_t *h_state, const int64_t *d_pop_users, const int64_t *d_pop_items,
                                    int64_t pop_len, int64_t k, int64_t *d_out_users, int64_t *d_out_items,
                                    mfb_stream stream) {
  RngScratchGuard guard;
  if (!d_pop_users || !d_pop_items || !d_out_users || !d_out_items) return MFB_ERR_INVALID;
  return choices_impl(h_state, d_pop_users, d_pop_items, pop_len, k, d_out_users, d_out_items, nullptr,
                      (cudaStream_t)stream);
}

extern "C" int mfb_mt_choices_indices(uint32_t *h_state, int64_t pop_len, int64_t k, int64_t *d_out,
                                      mfb_stream stream) {
  RngScratchGuard guard;
  if (!d_out) return MFB_ERR_INVALID;
  return choices_impl(h_state, nullptr, nullptr, pop_len, k, nullptr, nullptr, d_out, (cudaStream_t)stream);
}

extern "C" int mfb_mt_sample_items(uint32_t *h_state, int64_t num_items, int64_t count, int64_t *d_out,
                                   mfb_stream stream) {
  RngScratchGuard guard;
  cudaStream_t st = (cudaStream_t)stream;
  if (!h_state || num_items <= 0 || count < 0 || (count > 0 && !d_out) || num_items > 0xFFFFFFFFll) {
    mfb_set_error("mt_sample_items: bad arguments (num_items=%lld, count=%lld)", (long long)num_items,
                  (long long)count);
    return MFB_ERR_INVALID;
  }
  if (count == 0) return MFB_OK;
  const uint32_t rng = (uint32_t)(num_items - 1);
  if (rng == 0) {  // numpy returns zeros without touching the stream
    MFB_CUDA(cudaMemsetAsync(d_out, 0, (size_t)count * sizeof(int64_t), st));
    return MFB_OK;
  }
  uint32_t mask = rng;
  mask |= mask >> 1;
  mask |= mask >> 2;
  mask |= mask >> 4;
  mask |= mask >> 8;
  mask |= mask >> 16;
  uint32_t saved[625];
  memcpy(saved, h_state, sizeof(saved));
  double accept = ((double)rng + 1.0) / ((double)mask + 1.0);
  int64_t est = (int64_t)((double)count / accept * 1.05) + 4096;
  MFB_CHECK(g_result.reserve(2 * sizeof(long long)));
  for (int attempt = 0; attempt < 8; ++attempt) {
    uint32_t tmp[625];
    memcpy(tmp, saved, sizeof(tmp));
    MFB_CHECK(g_words.reserve((size_t)est * sizeof(uint32_t)));
    MFB_CHECK(mfb_mt_generate(tmp, est, g_words.as<uint32_t>(), st));
    mfb_count_library_launch(1);
    k_masked_compact<<<1, 1024, 0, st>>>(g_words.as<uint32_t>(), est, mask, rng, count, (long long *)d_out,
                                         g_result.as<long long>());
    MFB_KERNEL_CHECK();
    long long res[2];
    MFB_CUDA(cudaMemcpyAsync(res, g_result.ptr, sizeof(res), cudaMemcpyDeviceToHost, st));
    MFB_CUDA(cudaStreamSynchronize(st));
    if (res[0] > 0) {
      // advance the caller's state by exactly the number of words the reference would consume
      memcpy(h_state, saved, sizeof(saved));
      return mfb_mt_generate(h_state, res[0], nullptr, st);
    }
    est *= 2;
  }
  mfb_set_error("mt_sample_items: rejection sampling did not converge");
  return MFB_ERR_INVALID;
}

// ---------------------------------------------------------------------------------------------------------
// Offline negative-pair generator: spotlight/sampling.py:46-70 (get_negative_samples) with the rank-shift
// resampling of sampling.py:37-44 (negsamp_vectorized_bsearch_preverif), bit-exact with the reference's
// consumption of numpy's legacy global MT19937 stream:
//   users = np.random.choice(num_users, n); items = np.random.choice(num_items, n)      (two randint streams)
//   for i in order: if has_key(users[i], items[i]):                                       (CSR value == 1)
//       raw = np.random.randint(0, num_items - len(row(users[i])), 1)                     (masked rejection, 32-bit)
//       items[i] = raw + searchsorted(row - arange(len(row)), raw, 'right')               (raw-th item outside row)
// The membership test and the rank shift are data-parallel; only the order in which the re-draws consume the stream
// is sequential, and that part touches two row pointers and a few stream words per re-draw.
// ---------------------------------------------------------------------------------------------------------
namespace {

__global__ void k_neg_flag(const long long *__restrict__ users, const long long *__restrict__ items, long long n,
                           const long long *__restrict__ key_indptr, const int *__restrict__ key_indices,
                           int *__restrict__ flag) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long u = users[i];
  const int it = (int)items[i];
  long long lo = key_indptr[u], end = key_indptr[u + 1], hi = end;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (key_indices[mid] < it) lo = mid + 1; else hi = mid;
  }
  flag[i] = (lo < end && key_indices[lo] == it) ? 1 : 0;
}

// ordered compaction of the flagged sample indices (single CTA, stream order preserved)
__global__ void __launch_bounds__(1024) k_neg_compact(const int *__restrict__ flag, long long n,
                                                      long long *__restrict__ list, long long *__restrict__ count) {
  __shared__ int warp_sums[32];
  __shared__ long long base_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) base_s = 0;
  __syncthreads();
  for (long long start = 0; start < n; start += 1024) {
    const long long i = start + tid;
    const int ok = (i < n) ? flag[i] : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    const int in_warp = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) warp_sums[wid] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < 32; ++w) {
      const int c = warp_sums[w];
      if (w < wid) before += c;
      total += c;
    }
    if (ok) list[base_s + before + in_warp] = i;
    __syncthreads();
    if (tid == 0) base_s += total;
    __syncthreads();
  }
  if (tid == 0) *count = base_s;
}

// the sequential part: re-draw c takes stream words until (word & mask_c) <= rng_c, rng_c = free items of its user - 1
// result[0] = words consumed (or -1: the buffer ran out), result[1] = 1 if some user has no free item
__global__ void k_neg_redraw_serial(const long long *__restrict__ list, long long cnt,
                                    const long long *__restrict__ users, const long long *__restrict__ row_indptr,
                                    long long num_items, const uint32_t *__restrict__ words, long long nwords,
                                    uint32_t *__restrict__ raw, long long *__restrict__ result) {
  long long w = 0;
  result[1] = 0;
  for (long long c = 0; c < cnt; ++c) {
    const long long u = users[list[c]];
    const long long free_items = num_items - (row_indptr[u + 1] - row_indptr[u]);
    if (free_items <= 0) {   // numpy: randint(0, 0) raises ValueError("high <= 0")
      result[0] = w;
      result[1] = 1;
      return;
    }
    const uint32_t rng = (uint32_t)(free_items - 1);
    if (rng == 0) {          // numpy returns low without touching the stream
      raw[c] = 0u;
      continue;
    }
    uint32_t mask = rng;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    uint32_t v;
    do {
      if (w >= nwords) {
        result[0] = -1;
        return;
      }
      v = words[w++] & mask;
    } while (v > rng);
    raw[c] = v;
  }
  result[0] = w;
}

// items[list[c]] = the raw[c]-th item (0-based) outside the user's sorted row: raw + #{j : row[j] - j <= raw}
__global__ void k_neg_rank_shift(const long long *__restrict__ list, long long cnt, const long long *__restrict__ users,
                                 const long long *__restrict__ row_indptr, const int *__restrict__ row_indices,
                                 const uint32_t *__restrict__ raw, long long *__restrict__ items) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cnt) return;
  const long long i = list[c], u = users[i];
  const long long r0 = row_indptr[u], len = row_indptr[u + 1] - r0;
  const long long x = (long long)raw[c];
  long long lo = 0, hi = len;                       // searchsorted(row - arange, x, side='right')
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if ((long long)row_indices[r0 + mid] - mid <= x) lo = mid + 1; else hi = mid;
  }
  items[i] = x + lo;
}

DevBuf g_nflag, g_nlist, g_nraw;
void release_rng_scratch() {
  for (DevBuf *b : {&g_state, &g_words, &g_result, &g_jump, &g_nflag, &g_nlist, &g_nraw}) b->release();
}

}  // namespace

extern "C" int mfb_negative_pairs(uint32_t *h_state, int64_t num_users, int64_t num_items, int64_t num_samples,
                                  const int64_t *d_key_indptr, const int32_t *d_key_indices,
                                  const int64_t *d_row_indptr, const int32_t *d_row_indices, int64_t *d_out_users,
                                  int64_t *d_out_items, int64_t *h_n_redrawn, mfb_stream stream) {
  RngScratchGuard guard;
  cudaStream_t st = (cudaStream_t)stream;
  if (!h_state || num_users <= 0 || num_items <= 0 || num_samples < 0 || !d_key_indptr || !d_row_indptr ||
      (num_samples > 0 && (!d_out_users || !d_out_items))) {
    mfb_set_error("negative_pairs: bad arguments");
    return MFB_ERR_INVALID;
  }
  if (h_n_redrawn) *h_n_redrawn = 0;
  if (num_samples == 0) return MFB_OK;
  MFB_CHECK(mfb_mt_sample_items(h_state, num_users, num_samples, d_out_users, stream));
  MFB_CHECK(mfb_mt_sample_items(h_state, num_items, num_samples, d_out_items, stream));
  MFB_CHECK(g_nflag.reserve((size_t)num_samples * sizeof(int)));
  MFB_CHECK(g_nlist.reserve((size_t)num_samples * sizeof(long long)));
  MFB_CHECK(g_result.reserve(2 * sizeof(long long)));
  k_neg_flag<<<(unsigned)((num_samples + 255) / 256), 256, 0, st>>>((const long long *)d_out_users,
                                                                    (const long long *)d_out_items, num_samples,
                                                                    (const long long *)d_key_indptr, d_key_indices,
                                                                    g_nflag.as<int>());
  k_neg_compact<<<1, 1024, 0, st>>>(g_nflag.as<int>(), num_samples, g_nlist.as<long long>(), g_result.as<long long>());
  MFB_KERNEL_CHECK();
  mfb_count_library_launch(2);
  long long cnt = 0;
  MFB_CUDA(cudaMemcpyAsync(&cnt, g_result.ptr, sizeof(cnt), cudaMemcpyDeviceToHost, st));
  MFB_CUDA(cudaStreamSynchronize(st));
  if (h_n_redrawn) *h_n_redrawn = cnt;
  if (cnt == 0) return MFB_OK;
  MFB_CHECK(g_nraw.reserve((size_t)cnt * sizeof(uint32_t)));
  uint32_t saved[625];
  memcpy(saved, h_state, sizeof(saved));
  int64_t est = 4 * cnt + 1024;
  for (int attempt = 0; attempt < 8; ++attempt) {
    uint32_t tmp[625];
    memcpy(tmp, saved, sizeof(tmp));
    MFB_CHECK(g_words.reserve((size_t)est * sizeof(uint32_t)));
    MFB_CHECK(mfb_mt_generate(tmp, est, g_words.as<uint32_t>(), st));
    k_neg_redraw_serial<<<1, 1, 0, st>>>(g_nlist.as<long long>(), cnt, (const long long *)d_out_users,
                                         (const long long *)d_row_indptr, num_items, g_words.as<uint32_t>(), est,
                                         g_nraw.as<uint32_t>(), g_result.as<long long>());
    MFB_KERNEL_CHECK();
    mfb_count_library_launch(2);
    long long res[2];
    MFB_CUDA(cudaMemcpyAsync(res, g_result.ptr, sizeof(res), cudaMemcpyDeviceToHost, st));
    MFB_CUDA(cudaStreamSynchronize(st));
    if (res[1]) {
      mfb_set_error("high <= 0");   // numpy's ValueError from randint(0, 0): a user interacted with every item
      return MFB_ERR_RANGE;
    }
    if (res[0] >= 0) {
      k_neg_rank_shift<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(
          g_nlist.as<long long>(), cnt, (const long long *)d_out_users, (const long long *)d_row_indptr, d_row_indices,
          g_nraw.as<uint32_t>(), (long long *)d_out_items);
      MFB_KERNEL_CHECK();
      mfb_count_library_launch(1);
      memcpy(h_state, saved, sizeof(saved));
      if (res[0] > 0) MFB_CHECK(mfb_mt_generate(h_state, res[0], nullptr, st));   // advance by the words consumed
      MFB_CUDA(cudaStreamSynchronize(st));
      return MFB_OK;
    }
    est *= 2;
  }
  mfb_set_error("negative_pairs: rejection sampling did not converge");
  return MFB_ERR_INVALID;
}
