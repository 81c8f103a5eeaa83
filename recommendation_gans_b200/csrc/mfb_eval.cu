// Full-catalog evaluation: user x item scoring with the train-interaction mask and top-k fused,
// so the score matrix never reaches HBM.  Replaces the per-user loop of
// spotlight/evaluation.py:155-180 (model.predict(user) -> mask -> argsort -> set intersections).
//
// This file holds the exact-fp32 CUDA-core path: scores are sequential-FMA dot products
// (d = 0..D-1) + user bias + item bias, ranked on the pre-sigmoid value with ties -> lower item
// id.  It is the reference ordering the tensor-core path (mfb_eval_tc.cu) is
// re-scored against, and the path used for shapes the tensor-core kernel does not cover.
#include <math.h>

#include "mfb_internal.cuh"

namespace {

constexpr int EV_THREADS = 256;
constexpr int EV_ITEMS = 128;   // items per tile
constexpr int EV_USERS = 16;    // users per CTA (2 groups of 8, one group per half of the CTA)
constexpr int EV_UPT = 8;       // users per thread
constexpr float MASKED_SCORE = -3.402823466e38f;  // ranks after every real score, before empty slots

// Warp-resident sorted top-k list of up to 32 * KL entries: rank r lives in register r / 32 of lane r % 32.
// Candidates arrive in ascending item id, so a new entry goes after every entry with val >= nv (ties -> lower id first).
template <int KL>
struct TopK {
  float val[KL];
  int id[KL];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int j = 0; j < KL; ++j) {
      val[j] = -INFINITY;
      id[j] = 0x7fffffff;
    }
  }
  // value at rank r (warp-uniform r)
  __device__ __forceinline__ float value_at(int r) const {
    float v = val[0];
#pragma unroll
    for (int j = 1; j < KL; ++j) v = (r >> 5) == j ? val[j] : v;
    return __shfl_sync(0xffffffffu, v, r & 31);
  }
  __device__ __forceinline__ void insert(float nv, int nid, int lane) {
    int pos = 0;
#pragma unroll
    for (int j = 0; j < KL; ++j) pos += __popc(__ballot_sync(0xffffffffu, val[j] >= nv));
#pragma unroll
    for (int j = KL - 1; j >= 0; --j) {     // high registers first: register j-1 still holds its old entries
      float up_v = __shfl_up_sync(0xffffffffu, val[j], 1);
      int up_i = __shfl_up_sync(0xffffffffu, id[j], 1);
      if (j > 0) {
        const float pv = __shfl_sync(0xffffffffu, val[j - 1], 31);
        const int pi = __shfl_sync(0xffffffffu, id[j - 1], 31);
        if (lane == 0) {
          up_v = pv;
          up_i = pi;
        }
      }
      const int r = j * 32 + lane;
      if (r > pos) {
        val[j] = up_v;
        id[j] = up_i;
      } else if (r == pos) {
        val[j] = nv;
        id[j] = nid;
      }
    }
  }
};

// A candidate is admissible when it ranks strictly after the cut-off (cut_val, cut_id) of the previous pass
// (k > 32 * KL is served in passes of 32 * KL ranks); cut_id < 0: no cut-off.
__device__ __forceinline__ bool after_cut(float sc, int it, float cut_val, int cut_id) {
  return cut_id < 0 || sc < cut_val || (sc == cut_val && it > cut_id);
}

// One tile of EV_USERS listed users [u0, u0 + EV_USERS) against the whole catalog.
// dynamic smem: Vs[EV_ITEMS][D+1] | Us[EV_USERS][D] | S[EV_USERS][EV_ITEMS]
// out_pos (may be null): row of the output arrays that listed user ui writes (the tensor-core path's re-done users go
// straight to their rows of the caller's result).
template <int KL>
__device__ __forceinline__ void topk_exact_tile(const long long *__restrict__ user_ids, int n_users, int u0,
                                                const TableView &users, const TableView &items, int D,
                                                const long long *__restrict__ indptr, const int *__restrict__ indices,
                                                int k, int out_stride, int *__restrict__ out_ids,
                                                float *__restrict__ out_scores, const float *__restrict__ cut_val_in,
                                                const int *__restrict__ cut_id_in, float *__restrict__ cut_val_out,
                                                int *__restrict__ cut_id_out, const int *__restrict__ out_pos) {
  extern __shared__ float smem[];
  const int ldv = D + 1;
  float *Vs = smem;
  float *Us = Vs + EV_ITEMS * ldv;
  float *S = Us + EV_USERS * D;
  __shared__ long long s_uid[EV_USERS];
  __shared__ float s_ub[EV_USERS];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int I = items.rows;

  if (tid < EV_USERS) {
    int ui = u0 + tid;
    long long uid = (ui < n_users) ? user_ids[ui] : -1;
    s_uid[tid] = uid;
    s_ub[tid] = (uid >= 0) ? users.bp[uid] : 0.f;
  }
  __syncthreads();
  // user rows transposed ([d][user]) so a thread reads its 8 users at one d with two LDS.128
  for (int e = tid; e < EV_USERS * D; e += EV_THREADS) {
    int u = e / D, d = e - u * D;
    long long uid = s_uid[u];
    Us[d * EV_USERS + u] = (uid >= 0) ? users.p[uid * D + d] : 0.f;
  }

  // each warp owns two users of the tile for masking / top-k
  TopK<KL> top[2];
  long long cur[2], end[2];
  float cutv[2];
  int cuti[2];
#pragma unroll
  for (int w = 0; w < 2; ++w) {
    top[w].init();
    const int ui = u0 + wid * 2 + w;
    cutv[w] = 0.f;
    cuti[w] = -1;
    if (cut_id_in != nullptr && ui < n_users) {
      cutv[w] = cut_val_in[ui];
      cuti[w] = cut_id_in[ui];
    }
    long long uid = s_uid[wid * 2 + w];
    cur[w] = end[w] = 0;
    if (uid >= 0 && indptr != nullptr) {
      cur[w] = indptr[uid];
      end[w] = indptr[uid + 1];
    }
  }

  const int item_l = tid & (EV_ITEMS - 1);
  const int ugrp = tid / EV_ITEMS;  // 0 or 1 -> users ugrp*8 .. ugrp*8+7

  for (int i0 = 0; i0 < I; i0 += EV_ITEMS) {
    __syncthreads();  // previous tile fully consumed (also covers the Us fill on the first pass)
    // stage the item tile: coalesced global reads, padded rows in smem
    for (int e = tid; e < EV_ITEMS * D; e += EV_THREADS) {
      int r = e / D, d = e - r * D;
      int it = i0 + r;
      Vs[r * ldv + d] = (it < I) ? items.p[(long long)it * D + d] : 0.f;
    }
    __syncthreads();
    {
      float acc[EV_UPT];
#pragma unroll
      for (int u = 0; u < EV_UPT; ++u) acc[u] = 0.f;
      const float *vrow = Vs + item_l * ldv;
      const float *ubase = Us + ugrp * EV_UPT;
      for (int d = 0; d < D; ++d) {
        const float v = vrow[d];
        const float4 ua = *reinterpret_cast<const float4 *>(ubase + d * EV_USERS);
        const float4 ub = *reinterpret_cast<const float4 *>(ubase + d * EV_USERS + 4);
        acc[0] = fmaf(ua.x, v, acc[0]);
        acc[1] = fmaf(ua.y, v, acc[1]);
        acc[2] = fmaf(ua.z, v, acc[2]);
        acc[3] = fmaf(ua.w, v, acc[3]);
        acc[4] = fmaf(ub.x, v, acc[4]);
        acc[5] = fmaf(ub.y, v, acc[5]);
        acc[6] = fmaf(ub.z, v, acc[6]);
        acc[7] = fmaf(ub.w, v, acc[7]);
      }
      const int it = i0 + item_l;
      const float ib = (it < I) ? items.bp[it] : 0.f;
#pragma unroll
      for (int u = 0; u < EV_UPT; ++u) {
        int ul = ugrp * EV_UPT + u;
        S[ul * EV_ITEMS + item_l] = (acc[u] + s_ub[ul]) + ib;
      }
    }
    __syncthreads();
    // mask + top-k: warp wid handles users 2*wid, 2*wid+1
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const int ul = wid * 2 + w;
      if (s_uid[ul] < 0) continue;  // warp-uniform
      float *srow = S + ul * EV_ITEMS;
      // train items inside [i0, i0+EV_ITEMS): CSR indices are sorted, cursor only moves forward
      const int hi = i0 + EV_ITEMS;
      while (cur[w] < end[w]) {
        long long p = cur[w] + lane;
        int idx = (p < end[w]) ? indices[p] : 0x7fffffff;
        bool in = idx < hi;
        if (in && idx >= i0) srow[idx - i0] = MASKED_SCORE;
        unsigned nin = __popc(__ballot_sync(0xffffffffu, in));
        cur[w] += nin;
        if (nin < 32) break;
      }
      __syncwarp();
#pragma unroll
      for (int c = 0; c < EV_ITEMS / 32; ++c) {
        const int il = c * 32 + lane;
        const int it = i0 + il;
        float sc = srow[il];
        float thr = top[w].value_at(k - 1);
        unsigned cand = __ballot_sync(0xffffffffu, (it < I) && (sc > thr) && after_cut(sc, it, cutv[w], cuti[w]));
        while (cand) {
          int src = __ffs(cand) - 1;
          cand &= cand - 1;
          float nv = __shfl_sync(0xffffffffu, sc, src);
          int nid = i0 + c * 32 + src;
          thr = top[w].value_at(k - 1);
          if (nv > thr) top[w].insert(nv, nid, lane);
        }
      }
    }
  }
#pragma unroll
  for (int w = 0; w < 2; ++w) {
    const int ui = u0 + wid * 2 + w;
    if (ui >= n_users) continue;
    const long long orow = out_pos ? out_pos[ui] : ui;
#pragma unroll
    for (int j = 0; j < KL; ++j) {
      const int r = j * 32 + lane;
      if (r < k) {
        out_ids[orow * out_stride + r] = top[w].id[j];
        if (out_scores) {
          float z = top[w].val[j];
          out_scores[orow * out_stride + r] = (z == MASKED_SCORE) ? 0.f : 1.0f / (1.0f + expf(-z));
        }
        if (cut_id_out != nullptr && r == k - 1) {    // where the next pass resumes
          cut_val_out[ui] = top[w].val[j];
          cut_id_out[ui] = top[w].id[j];
        }
      }
    }
  }
}

// grid: ceil(n_users / EV_USERS) blocks, or fewer: a block walks the user tiles blockIdx.x, blockIdx.x + gridDim.x, ...
// n_dev (may be null): device-resident count that caps n_users -- the tensor-core path enqueues the re-do of its
// uncertified users without knowing on the host how many there are (usually none: every block leaves at once).
template <int KL>
__global__ void __launch_bounds__(EV_THREADS) k_topk_exact(const long long *__restrict__ user_ids, int n_users,
                                                           TableView users, TableView items, int D,
                                                           const long long *__restrict__ indptr,
                                                           const int *__restrict__ indices, int k, int out_stride,
                                                           int *__restrict__ out_ids, float *__restrict__ out_scores,
                                                           const float *__restrict__ cut_val_in,
                                                           const int *__restrict__ cut_id_in,
                                                           float *__restrict__ cut_val_out,
                                                           int *__restrict__ cut_id_out, const int *__restrict__ n_dev,
                                                           const int *__restrict__ out_pos) {
  if (n_dev != nullptr) n_users = min(n_users, *n_dev);
  for (int u0 = blockIdx.x * EV_USERS; u0 < n_users; u0 += gridDim.x * EV_USERS) {
    __syncthreads();   // the previous tile's shared-memory state is no longer read
    topk_exact_tile<KL>(user_ids, n_users, u0, users, items, D, indptr, indices, k, out_stride, out_ids, out_scores,
                        cut_val_in, cut_id_in, cut_val_out, cut_id_out, out_pos);
  }
}

// Top-k of the rows of a dense score matrix [n_rows, n_items] (scores produced by any model, e.g. an MLP / NeuMF
// `representation=`): descending score, ties -> lower item id, the row user's train items last.  One warp per row.
template <int KL>
__global__ void __launch_bounds__(EV_THREADS) k_topk_dense(const float *__restrict__ scores, int n_rows, int n_items,
                                                           const long long *__restrict__ user_ids,
                                                           const long long *__restrict__ indptr,
                                                           const int *__restrict__ indices, int k, int out_stride,
                                                           int *__restrict__ out_ids, float *__restrict__ out_scores,
                                                           const float *__restrict__ cut_val_in,
                                                           const int *__restrict__ cut_id_in,
                                                           float *__restrict__ cut_val_out,
                                                           int *__restrict__ cut_id_out) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (EV_THREADS / 32) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  TopK<KL> top;
  top.init();
  float cutv = 0.f;
  int cuti = -1;
  if (cut_id_in != nullptr) {
    cutv = cut_val_in[row];
    cuti = cut_id_in[row];
  }
  long long cur = 0, end = 0;
  if (indptr != nullptr) {
    const long long uid = user_ids ? user_ids[row] : row;
    cur = indptr[uid];
    end = indptr[uid + 1];
  }
  const float *srow = scores + (long long)row * n_items;
  for (int i0 = 0; i0 < n_items; i0 += 32) {
    const int it = i0 + lane;
    float sc = (it < n_items) ? srow[it] : -INFINITY;
    // train items inside [i0, i0+32): the CSR row is sorted, the cursor only moves forward
    unsigned masked = 0u;
    while (cur < end) {
      const long long p = cur + lane;
      const int idx = (p < end) ? indices[p] : 0x7fffffff;
      const bool in = idx < i0 + 32;
      masked |= __reduce_or_sync(0xffffffffu, (in && idx >= i0) ? (1u << (idx - i0)) : 0u);
      const unsigned nin = __popc(__ballot_sync(0xffffffffu, in));
      cur += nin;
      if (nin < 32) break;
    }
    if ((masked >> lane) & 1u) sc = MASKED_SCORE;
    float thr = top.value_at(k - 1);
    unsigned cand = __ballot_sync(0xffffffffu, (it < n_items) && (sc > thr) && after_cut(sc, it, cutv, cuti));
    while (cand) {
      const int src = __ffs(cand) - 1;
      cand &= cand - 1;
      const float nv = __shfl_sync(0xffffffffu, sc, src);
      thr = top.value_at(k - 1);
      if (nv > thr) top.insert(nv, i0 + src, lane);
    }
  }
#pragma unroll
  for (int j = 0; j < KL; ++j) {
    const int r = j * 32 + lane;
    if (r < k) {
      out_ids[(long long)row * out_stride + r] = top.id[j];
      if (out_scores) out_scores[(long long)row * out_stride + r] = top.val[j];
      if (cut_id_out != nullptr && r == k - 1) {
        cut_val_out[row] = top.val[j];
        cut_id_out[row] = top.id[j];
      }
    }
  }
}

// hits[u, j] = |topk[u, :ks[j]] ∩ test_row(user_ids[u])|  (evaluation.py:108-113); one thread per user
__global__ void k_topk_hits(const int *__restrict__ topk, const long long *__restrict__ user_ids, int n_users, int k,
                            const long long *__restrict__ indptr, const int *__restrict__ indices, int k0, int k1,
                            int k2, int k3, int nk, int *__restrict__ hits, int *__restrict__ ntargets) {
  int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_users) return;
  const int ks[4] = {k0, k1, k2, k3};
  long long uid = user_ids[u];
  long long lo = indptr[uid], hi = indptr[uid + 1];
  ntargets[u] = (int)(hi - lo);
  int h = 0, next = 0;
  for (int r = 0; r < k; ++r) {
    int id = topk[(long long)u * k + r];
    long long a = lo, b = hi;
    while (a < b) {
      long long mid = (a + b) >> 1;
      if (indices[mid] < id) a = mid + 1; else b = mid;
    }
    if (a < hi && indices[a] == id) ++h;
    while (next < nk && ks[next] == r + 1) hits[(long long)u * nk + next++] = h;
  }
}

}  // namespace

constexpr int EV_PASS_RANKS = 256;   // ranks per pass of the warp-resident list (8 registers x 32 lanes)

// launch one pass of KERNEL<KL> with KL chosen from the pass width
#define MFB_TOPK_DISPATCH(KERNEL, kpass, GRID, SMEM, ...)                                   \
  do {                                                                                      \
    if ((kpass) <= 32) KERNEL<1><<<GRID, EV_THREADS, SMEM, st>>>(__VA_ARGS__);              \
    else if ((kpass) <= 64) KERNEL<2><<<GRID, EV_THREADS, SMEM, st>>>(__VA_ARGS__);         \
    else if ((kpass) <= 128) KERNEL<4><<<GRID, EV_THREADS, SMEM, st>>>(__VA_ARGS__);        \
    else KERNEL<8><<<GRID, EV_THREADS, SMEM, st>>>(__VA_ARGS__);                            \
  } while (0)

// d_n_dev / d_out_pos: see k_topk_exact (both null for an ordinary call)
static int topk_exact_impl(mfb_model *m, const int64_t *d_user_ids, int64_t n_users, const int64_t *d_train_indptr,
                           const int32_t *d_train_indices, int32_t k, int32_t *d_out_ids, float *d_out_scores,
                           cudaStream_t st, const int *d_n_dev, const int *d_out_pos) {
  const int D = m->desc.dim;
  size_t smem = ((size_t)EV_ITEMS * (D + 1) + (size_t)EV_USERS * D + (size_t)EV_USERS * EV_ITEMS) * sizeof(float);
  if (smem > 220 * 1024) {
    mfb_set_error("topk: embedding_dim %d needs %zu B of shared memory", D, smem);
    return MFB_ERR_UNSUPPORTED;
  }
  static size_t smem_set = 0;   // (one device per process: the attributes are set when the size first grows)
  if (smem > smem_set) {
    MFB_CUDA(cudaFuncSetAttribute(k_topk_exact<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MFB_CUDA(cudaFuncSetAttribute(k_topk_exact<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MFB_CUDA(cudaFuncSetAttribute(k_topk_exact<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MFB_CUDA(cudaFuncSetAttribute(k_topk_exact<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  int grid = (int)((n_users + EV_USERS - 1) / EV_USERS);
  if (d_n_dev != nullptr && grid > 2 * m->num_sms) grid = 2 * m->num_sms;   // count unknown here: the blocks stride
  // k beyond one pass of the warp-resident list: further passes resume after the last (score, id) of the previous one
  float *cut_val = nullptr;
  int *cut_id = nullptr;
  if (k > EV_PASS_RANKS) {
    MFB_CHECK(m->eval.cut.reserve((size_t)n_users * (sizeof(float) + sizeof(int))));
    cut_val = m->eval.cut.as<float>();
    cut_id = reinterpret_cast<int *>(cut_val + n_users);
  }
  for (int done = 0; done < k; done += EV_PASS_RANKS) {
    const int kpass = (k - done < EV_PASS_RANKS) ? (k - done) : EV_PASS_RANKS;
    int tk = m->prof.begin(PK_TOPK, st);
    MFB_TOPK_DISPATCH(k_topk_exact, kpass, grid, smem, (const long long *)d_user_ids, (int)n_users, m->users, m->items, D,
                      (const long long *)d_train_indptr, d_train_indices, kpass, (int)k, d_out_ids + done,
                      d_out_scores ? d_out_scores + done : nullptr, done ? cut_val : nullptr, done ? cut_id : nullptr,
                      cut_val, cut_id, d_n_dev, d_out_pos);
    m->prof.end(tk, st);
    MFB_KERNEL_CHECK();
  }
  return MFB_OK;
}

static int topk_impl(mfb_model *m, const int64_t *d_user_ids, int64_t n_users, const int64_t *d_train_indptr,
                     const int32_t *d_train_indices, int32_t k, int32_t *d_out_ids, float *d_out_scores,
                     mfb_stream stream, uint64_t plan_key) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!m || !d_user_ids || !d_out_ids || n_users < 0) return MFB_ERR_INVALID;
  if (k <= 0 || k > m->items.rows) {
    mfb_set_error("topk: k=%d outside 1..num_items (%d)", k, m->items.rows);
    return MFB_ERR_UNSUPPORTED;
  }
  if ((d_train_indptr == nullptr) != (d_train_indices == nullptr)) return MFB_ERR_INVALID;
  if (n_users == 0) return MFB_OK;
  MFB_CHECK(mfb_flush(m, stream));
  m->last_topk_redo = 0;
  // large problems go through the tensor-core path (fp16 candidates + exact fp32 re-score: same ids as the
  // exact kernel); small ones -- and k beyond MFB_MAX_TOPK -- take the exact kernel
  if (mfb_tc_supported(m, k) && n_users >= 64 && n_users < (1ll << 30))
    return mfb_topk_tc(m, d_user_ids, n_users, d_train_indptr, d_train_indices, k, d_out_ids, d_out_scores, st,
                       topk_exact_impl, &m->last_topk_redo, plan_key);
  return topk_exact_impl(m, d_user_ids, n_users, d_train_indptr, d_train_indices, k, d_out_ids, d_out_scores, st, nullptr,
                         nullptr);
}

extern "C" int mfb_topk(mfb_model *m, const int64_t *d_user_ids, int64_t n_users, const int64_t *d_train_indptr,
                        const int32_t *d_train_indices, int32_t k, int32_t *d_out_ids, float *d_out_scores,
                        mfb_stream stream) {
  return topk_impl(m, d_user_ids, n_users, d_train_indptr, d_train_indices, k, d_out_ids, d_out_scores, stream, 0);
}

extern "C" int mfb_topk_keyed(mfb_model *m, const int64_t *d_user_ids, int64_t n_users, const int64_t *d_train_indptr,
                              const int32_t *d_train_indices, int32_t k, int32_t *d_out_ids, float *d_out_scores,
                              uint64_t plan_key, mfb_stream stream) {
  return topk_impl(m, d_user_ids, n_users, d_train_indptr, d_train_indices, k, d_out_ids, d_out_scores, stream,
                   plan_key);
}

// Top-k of dense score rows (scores from any model): see k_topk_dense.  d_user_ids maps a row to the user whose train
// row masks it (null: row r is user r); d_cut_scratch: n_rows * 8 bytes of device scratch, needed only when k > 256.
extern "C" int mfb_topk_scores(const float *d_scores, int64_t n_rows, int64_t n_items, const int64_t *d_user_ids,
                               const int64_t *d_train_indptr, const int32_t *d_train_indices, int32_t k,
                               int32_t *d_out_ids, float *d_out_scores, void *d_cut_scratch, mfb_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!d_scores || !d_out_ids || n_rows < 0 || n_items <= 0 || n_items >= (1ll << 31) || n_rows >= (1ll << 31))
    return MFB_ERR_INVALID;
  if (k <= 0 || k > n_items) {
    mfb_set_error("topk_scores: k=%d outside 1..num_items (%lld)", k, (long long)n_items);
    return MFB_ERR_UNSUPPORTED;
  }
  if ((d_train_indptr == nullptr) != (d_train_indices == nullptr)) return MFB_ERR_INVALID;
  if (k > EV_PASS_RANKS && d_cut_scratch == nullptr) {
    mfb_set_error("topk_scores: k=%d > %d needs the cut-off scratch", k, EV_PASS_RANKS);
    return MFB_ERR_INVALID;
  }
  if (n_rows == 0) return MFB_OK;
  float *cut_val = reinterpret_cast<float *>(d_cut_scratch);
  int *cut_id = cut_val ? reinterpret_cast<int *>(cut_val + n_rows) : nullptr;
  const int grid = (int)((n_rows + EV_THREADS / 32 - 1) / (EV_THREADS / 32));
  for (int done = 0; done < k; done += EV_PASS_RANKS) {
    const int kpass = (k - done < EV_PASS_RANKS) ? (k - done) : EV_PASS_RANKS;
    mfb_count_library_launch(1);
    MFB_TOPK_DISPATCH(k_topk_dense, kpass, grid, 0, d_scores, (int)n_rows, (int)n_items, (const long long *)d_user_ids,
                      (const long long *)d_train_indptr, d_train_indices, kpass, (int)k, d_out_ids + done,
                      d_out_scores ? d_out_scores + done : nullptr, done ? cut_val : nullptr, done ? cut_id : nullptr,
                      cut_val, cut_id);
    MFB_KERNEL_CHECK();
  }
  return MFB_OK;
}

// Number of users the last mfb_topk call re-did with the exact kernel (tensor-core path only).
extern "C" int mfb_topk_last_redo(const mfb_model *mc) {
  mfb_model *m = const_cast<mfb_model *>(mc);
  if (!m) return -1;
  if (m->last_topk_redo < 0 && m->eval.redo_cnt_dev != nullptr) {   // the tensor-core pass left the count on the device
    int n = 0;
    if (cudaStreamSynchronize(m->eval.redo_stream) != cudaSuccess ||
        cudaMemcpy(&n, m->eval.redo_cnt_dev, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess)
      return -1;
    m->last_topk_redo = n;
  }
  return m->last_topk_redo;
}

extern "C" int mfb_debug_tc_stats(mfb_model *m, int64_t n_users, int64_t *h_out, mfb_stream stream) {
  if (!m || !h_out || n_users <= 0) return MFB_ERR_INVALID;
  return mfb_tc_stats(m, (int)n_users, (long long *)h_out, (cudaStream_t)stream);
}

// Test hook: raw tensor-core (fp16 x fp16 -> fp32, + item bias) scores of the listed users, item-major
// [num_items][ceil(n_users/256)*256].
extern "C" int mfb_debug_tc_scores(mfb_model *m, const int64_t *d_user_ids, int64_t n_users, float *d_out,
                                   mfb_stream stream) {
  if (!m || !d_user_ids || !d_out || n_users <= 0) return MFB_ERR_INVALID;
  if (m->desc.dim < 1 || m->desc.dim > 128) return MFB_ERR_UNSUPPORTED;
  MFB_CHECK(mfb_flush(m, stream));
  return mfb_tc_dump_scores(m, d_user_ids, (int)n_users, d_out, (cudaStream_t)stream);
}

extern "C" int mfb_topk_hits(const int32_t *d_topk_ids, const int64_t *d_user_ids, int64_t n_users, int32_t k,
                             const int64_t *d_test_indptr, const int32_t *d_test_indices, const int32_t *h_ks,
                             int32_t nk, int32_t *d_hits, int32_t *d_ntargets, mfb_stream stream) {
  if (!d_topk_ids || !d_user_ids || !d_test_indptr || !d_test_indices || !h_ks || !d_hits || !d_ntargets)
    return MFB_ERR_INVALID;
  if (nk < 1 || nk > 4) {
    mfb_set_error("topk_hits: 1..4 cut-offs per call (got %d)", nk);
    return MFB_ERR_UNSUPPORTED;
  }
  int ks[4] = {0, 0, 0, 0};
  for (int j = 0; j < nk; ++j) {
    ks[j] = h_ks[j];
    if (ks[j] < 1 || ks[j] > k || (j > 0 && ks[j] <= ks[j - 1])) {
      mfb_set_error("topk_hits: cut-offs must be ascending within 1..k");
      return MFB_ERR_INVALID;
    }
  }
  if (n_users == 0) return MFB_OK;
  mfb_count_library_launch(1);
  k_topk_hits<<<(unsigned)((n_users + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      d_topk_ids, (const long long *)d_user_ids, (int)n_users, k, (const long long *)d_test_indptr, d_test_indices,
      ks[0], ks[1], ks[2], ks[3], nk, d_hits, d_ntargets);
  MFB_KERNEL_CHECK();
  return MFB_OK;
}
