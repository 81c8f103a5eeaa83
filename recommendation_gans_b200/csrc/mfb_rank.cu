// Full ranks of the test items (mean reciprocal rank, spotlight/evaluation.py:13-60).
//
// The reference computes, per user with test items, predictions = -model.predict(user), sets the user's train items to
// FLOAT_MAX, ranks the whole vector with scipy.stats.rankdata (average ranks for ties) and reads the ranks of the test
// items.  Here a CTA owns a user: every warp keeps the user's row in registers and walks the catalog, computing each
// item's probability exactly as predict does (same per-lane FMA chain, warp-shuffle sum, bias order, sigmoid) and
// comparing it with the user's test-item probabilities -- no score vector and no sort:
//   rank(t) = #{free items j : p_j > p_t} + (#{free items j : p_j == p_t} + 1) / 2          (t itself counts in ==)
// and a test item that is also a train item ties with all train items after every free item.
#include "mfb_internal.cuh"
#include "mfb_rowops.cuh"

namespace {

constexpr int RK_WARPS = 8;
constexpr int RK_THREADS = RK_WARPS * 32;
constexpr int RK_TB = 64;   // test items per pass over the catalog (two per lane)

__device__ __forceinline__ bool in_sorted(const int *__restrict__ a, long long lo, long long end, int x) {
  long long hi = end;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    const int v = a[mid];
    if (v < x) lo = mid + 1; else hi = mid;
  }
  return lo < end && a[lo] == x;
}

template <int VEC, int NIT>
__global__ void __launch_bounds__(RK_THREADS) k_rank_test_items(const long long *__restrict__ user_ids, int n_users,
                                                                TableView users, TableView items, int D,
                                                                const long long *__restrict__ test_indptr,
                                                                const int *__restrict__ test_indices,
                                                                const long long *__restrict__ train_indptr,
                                                                const int *__restrict__ train_indices,
                                                                float *__restrict__ out_rank) {
  __shared__ float tp[RK_TB];
  __shared__ int t_train[RK_TB];
  __shared__ unsigned int c_gt[RK_TB], c_eq[RK_TB];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int ui = blockIdx.x;
  if (ui >= n_users) return;
  const long long u = user_ids[ui];
  const int I = items.rows;
  Frag<VEC, NIT> fu;
  frag_load<VEC, NIT>(fu, users.p + u * D, D, lane);
  const float bu = users.bp[u];
  const long long t0 = test_indptr[u], t1 = test_indptr[u + 1];
  long long r0 = 0, r1 = 0;
  if (train_indptr != nullptr) {
    r0 = train_indptr[u];
    r1 = train_indptr[u + 1];
  }
  const int n_train = (int)(r1 - r0);
  auto prob = [&](int i) {
    Frag<VEC, NIT> fi;
    frag_load<VEC, NIT>(fi, items.p + (long long)i * D, D, lane);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < NIT * VEC; ++k) acc = fmaf(fu.x[k], fi.x[k], acc);
    acc = warp_sum(acc);
    return sigmoidf_acc((acc + bu) + items.bp[i]);
  };
  for (long long tb = t0; tb < t1; tb += RK_TB) {
    const int nt = (int)((t1 - tb < RK_TB) ? (t1 - tb) : RK_TB);
    __syncthreads();
    for (int t = wid; t < nt; t += RK_WARPS) {          // probabilities of this batch of test items
      const int item = test_indices[tb + t];
      const float p = prob(item);
      if (lane == 0) {
        tp[t] = p;
        t_train[t] = (n_train > 0 && in_sorted(train_indices, r0, r1, item)) ? 1 : 0;
        c_gt[t] = 0u;
        c_eq[t] = 0u;
      }
    }
    __syncthreads();
    const float p_a = lane < nt ? tp[lane] : 2.0f, p_b = lane + 32 < nt ? tp[lane + 32] : 2.0f;   // 2 > any probability
    unsigned int gt_a = 0, eq_a = 0, gt_b = 0, eq_b = 0;
    long long cur = r0;                                                  // cursor into the sorted train row
    for (int i = wid; i < I; i += RK_WARPS) {                            // ascending items: the cursor only advances
      const float p = prob(i);                                           // same value in every lane
      while (cur < r1 && train_indices[cur] < i) ++cur;
      if (cur < r1 && train_indices[cur] == i) continue;                 // train items rank after every free item
      gt_a += p > p_a;
      eq_a += p == p_a;
      gt_b += p > p_b;
      eq_b += p == p_b;
    }
    if (lane < nt) {
      atomicAdd(&c_gt[lane], gt_a);
      atomicAdd(&c_eq[lane], eq_a);
    }
    if (lane + 32 < nt) {
      atomicAdd(&c_gt[lane + 32], gt_b);
      atomicAdd(&c_eq[lane + 32], eq_b);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nt; t += RK_THREADS) {
      float rank;
      if (t_train[t]) rank = (float)(I - n_train) + 0.5f * (float)(n_train + 1);
      else rank = (float)c_gt[t] + 0.5f * (float)(c_eq[t] + 1u);
      out_rank[tb + t] = rank;
    }
  }
}

}  // namespace

extern "C" int mfb_rank_test_items(mfb_model *m, const int64_t *d_user_ids, int64_t n_users,
                                   const int64_t *d_test_indptr, const int32_t *d_test_indices,
                                   const int64_t *d_train_indptr, const int32_t *d_train_indices, float *d_out_rank,
                                   mfb_stream stream) {
  if (!m || !d_user_ids || !d_test_indptr || !d_test_indices || !d_out_rank || n_users < 0) return MFB_ERR_INVALID;
  if ((d_train_indptr == nullptr) != (d_train_indices == nullptr)) return MFB_ERR_INVALID;
  if (n_users == 0) return MFB_OK;
  MFB_CHECK(mfb_flush(m, stream));
  Shape sh;
  MFB_CHECK(pick_shape(m->desc.dim, &sh));
  cudaStream_t st = (cudaStream_t)stream;
  int tk = m->prof.begin(PK_TOPK, st);
#define CALL(V, N)                                                                                                  \
  k_rank_test_items<V, N><<<(unsigned)n_users, RK_THREADS, 0, st>>>(                                                \
      (const long long *)d_user_ids, (int)n_users, m->users, m->items, m->desc.dim, (const long long *)d_test_indptr, \
      d_test_indices, (const long long *)d_train_indptr, d_train_indices, d_out_rank)
  MFB_DISPATCH_SHAPE(sh, CALL);
#undef CALL
  m->prof.end(tk, st);
  MFB_KERNEL_CHECK();
  return MFB_OK;
}
