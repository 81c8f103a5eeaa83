// Fused implicit-MF training step for sm_100a.
//
// Replaces ImplicitFactorizationModel.run_train_iteration (implicit.py:347-364): BilinearNet
// forward on the positive batch and on the sampled negative pairs (representations.py:80-91),
// the ranking loss (spotlight/losses.py:20-172), autograd's dense embedding backward, and the
// dense torch optimiser step (spotlight/optimizers.py -> torch.optim.Adam / SGD).
//
// Semantics are the reference's DENSE optimiser (every row of every table steps every
// iteration); traffic is row-sparse: each row carries `last` = the optimiser step it
// is current for, and missed steps (gradient = weight_decay * p) are replayed elementwise in
// registers when the row is next gathered (k_catchup) or at mfb_flush.
//
// Per step:  k_catchup (unique rows -> state at t-1)  ->  k_forward (gather, dot, sigmoid; row
// snapshots)  ->  k_loss (loss value, dLoss/dz per slot)  ->  k_update (per unique row: batch-
// ordered segment reduction of the slot gradients, optimiser step t).  Unique rows / segments
// come from the planner's stable radix sort (mfb_sort.cu), so the reduction order is the batch
// order used by torch's index_add and no atomics are involved.
#include <math.h>

#include "mfb_internal.cuh"

namespace {

constexpr int WARPS_PER_BLOCK = 8;
constexpr int BLOCK_THREADS = WARPS_PER_BLOCK * 32;

// ---------------------------------------------------------------------------------------
// per-lane row fragments: element index e = (it*32 + lane)*VEC + k
// ---------------------------------------------------------------------------------------
template <int VEC, int NIT>
struct Frag {
  float x[NIT * VEC];
};

template <int VEC, int NIT>
__device__ __forceinline__ void frag_load(Frag<VEC, NIT> &f, const float *__restrict__ row, int D, int lane) {
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int e = (it * 32 + lane) * VEC;
    if constexpr (VEC == 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < D) v = *reinterpret_cast<const float4 *>(row + e);
      f.x[it * 4 + 0] = v.x;
      f.x[it * 4 + 1] = v.y;
      f.x[it * 4 + 2] = v.z;
      f.x[it * 4 + 3] = v.w;
    } else {
      f.x[it] = (e < D) ? row[e] : 0.f;
    }
  }
}

template <int VEC, int NIT>
__device__ __forceinline__ void frag_store(const Frag<VEC, NIT> &f, float *__restrict__ row, int D, int lane) {
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int e = (it * 32 + lane) * VEC;
    if (e < D) {
      if constexpr (VEC == 4) {
        *reinterpret_cast<float4 *>(row + e) =
            make_float4(f.x[it * 4 + 0], f.x[it * 4 + 1], f.x[it * 4 + 2], f.x[it * 4 + 3]);
      } else {
        row[e] = f.x[it];
      }
    }
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float sigmoidf_acc(float z) { return 1.0f / (1.0f + expf(-z)); }

__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// One dense torch.optim.Adam update of a single element (torch/optim/adam.py _single_tensor_adam,
// op order as on the CPU path): grad += wd*p; exp_avg.lerp_(grad, 1-b1);
// exp_avg_sq = exp_avg_sq*b2 + (1-b2)*grad*grad; denom = sqrt(v)/bc2_sqrt + eps;
// p += (-step_size*m)/denom.
template <bool FAST>
__device__ __forceinline__ void adam_elem(float &p, float &m, float &v, float g, float neg_ss, float bc2,
                                          const OptView &o) {
  g = fmaf(o.wd, p, g);
  float base = o.lerp_small ? m : g;
  m = fmaf(o.lerp_coeff, __fsub_rn(g, m), base);
  v = __fadd_rn(__fmul_rn(v, o.beta2), __fmul_rn(__fmul_rn(o.one_minus_beta2, g), g));
  if (FAST) {
    float denom = __fadd_rn(__fdividef(sqrt_approx(v), bc2), o.eps);
    p = __fadd_rn(p, __fdividef(__fmul_rn(neg_ss, m), denom));
  } else {
    float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2), o.eps);
    p = __fadd_rn(p, __fdiv_rn(__fmul_rn(neg_ss, m), denom));
  }
}

// torch.optim.SGD, momentum 0: p <- p - lr*(g + wd*p)
__device__ __forceinline__ void sgd_elem(float &p, float g, const OptView &o) {
  g = fmaf(o.wd, p, g);
  p = fmaf(-o.lr, g, p);
}

// Full optimiser state of one table row held by a warp.
template <int VEC, int NIT>
struct RowState {
  Frag<VEC, NIT> p, m, v;
  float bp, bm, bv;  // bias (replicated in all lanes)
};

template <int VEC, int NIT>
__device__ __forceinline__ void row_load(RowState<VEC, NIT> &r, const TableView &T, long long row, int D, int lane,
                                         bool adam) {
  frag_load<VEC, NIT>(r.p, T.p + row * D, D, lane);
  r.bp = T.bp[row];
  if (adam) {
    frag_load<VEC, NIT>(r.m, T.m + row * D, D, lane);
    frag_load<VEC, NIT>(r.v, T.v + row * D, D, lane);
    r.bm = T.bm[row];
    r.bv = T.bv[row];
  }
}

template <int VEC, int NIT>
__device__ __forceinline__ void row_store(const RowState<VEC, NIT> &r, const TableView &T, long long row, int D,
                                          int lane, bool adam) {
  frag_store<VEC, NIT>(r.p, T.p + row * D, D, lane);
  if (adam) {
    frag_store<VEC, NIT>(r.m, T.m + row * D, D, lane);
    frag_store<VEC, NIT>(r.v, T.v + row * D, D, lane);
  }
  if (lane == 0) {
    T.bp[row] = r.bp;
    if (adam) {
      T.bm[row] = r.bm;
      T.bv[row] = r.bv;
    }
  }
}

// Replays the dense zero-gradient updates of optimiser steps (from, to] on a row.
template <int VEC, int NIT, bool FAST>
__device__ __forceinline__ void row_replay(RowState<VEC, NIT> &r, int from, int to, const OptView &o) {
  if (o.kind == MFB_OPT_ADAM) {
    for (int s = from + 1; s <= to; ++s) {
      float neg_ss = -__ldg(o.step_size + s);
      float bc2 = __ldg(o.bc2_sqrt + s);
#pragma unroll
      for (int k = 0; k < NIT * VEC; ++k) adam_elem<FAST>(r.p.x[k], r.m.x[k], r.v.x[k], 0.f, neg_ss, bc2, o);
      adam_elem<FAST>(r.bp, r.bm, r.bv, 0.f, neg_ss, bc2, o);
    }
  } else {
    if (o.wd == 0.f) return;  // p - lr*(0 + 0*p) == p
    for (int s = from + 1; s <= to; ++s) {
#pragma unroll
      for (int k = 0; k < NIT * VEC; ++k) sgd_elem(r.p.x[k], 0.f, o);
      sgd_elem(r.bp, 0.f, o);
    }
  }
}

// ---------------------------------------------------------------------------------------
// planner: pack slot ids and sort keys for a chunk of steps
// ---------------------------------------------------------------------------------------
// Step s (chunk-local) has b_s positives followed by m negatives; slots of step s start at
// s*Lfull (only the final step of an epoch can be partial and it is last in its chunk).
// Key = step << (rb+1) | table << rb | row; keys of a step are laid out [users of all slots]
// [items of all slots] so a stable sort leaves each (step,table,row) group in slot order.
__global__ void k_pack(const long long *__restrict__ pos_u, const long long *__restrict__ pos_i, long long n_pos,
                       const long long *__restrict__ neg_u, const long long *__restrict__ neg_i, int batch, int m_neg,
                       long long step0, int nsteps, int rb, int num_users, int num_items, int *__restrict__ slot_u,
                       int *__restrict__ slot_i, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals,
                       int *__restrict__ err_flag) {
  const int Lfull = batch + m_neg;
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)nsteps * Lfull) return;
  int s = (int)(gid / Lfull);
  int j = (int)(gid % Lfull);
  long long gstep = step0 + s;
  long long first = gstep * batch;
  int b = (int)((n_pos - first < batch) ? (n_pos - first) : batch);
  int L = b + m_neg;
  if (j >= L) return;
  long long u, i;
  if (j < b) {
    u = pos_u[first + j];
    i = pos_i[first + j];
  } else {
    long long q = gstep * m_neg + (j - b);
    u = neg_u[q];
    i = neg_i[q];
  }
  if (u < 0 || u >= num_users || i < 0 || i >= num_items) {
    atomicExch(err_flag, 1);
    u = 0;
    i = 0;
  }
  long long so = (long long)s * Lfull;
  slot_u[so + j] = (int)u;
  slot_i[so + j] = (int)i;
  long long ko = 2 * so;
  keys[ko + j] = ((uint32_t)s << (rb + 1)) | (uint32_t)u;
  vals[ko + j] = (uint32_t)j;
  keys[ko + L + j] = ((uint32_t)s << (rb + 1)) | (1u << rb) | (uint32_t)i;
  vals[ko + L + j] = (uint32_t)j;
}

// ---------------------------------------------------------------------------------------
// k_catchup: one warp per sorted position; segment heads bring their row to step t-1
// ---------------------------------------------------------------------------------------
template <int VEC, int NIT, bool FAST>
__global__ void __launch_bounds__(BLOCK_THREADS) k_catchup(const uint32_t *__restrict__ skeys, int n, int rb,
                                                           TableView users, TableView items, OptView opt, int D,
                                                           int t) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (q >= n) return;
  const uint32_t key = skeys[q];
  if (q > 0 && skeys[q - 1] == key) return;
  const long long row = key & ((1u << rb) - 1u);
  const TableView &T = ((key >> rb) & 1u) ? items : users;
  const int last = T.last[row];
  if (last >= t - 1) return;
  const bool adam = opt.kind == MFB_OPT_ADAM;
  RowState<VEC, NIT> r;
  row_load<VEC, NIT>(r, T, row, D, lane, adam);
  row_replay<VEC, NIT, FAST>(r, last, t - 1, opt);
  row_store<VEC, NIT>(r, T, row, D, lane, adam);
  if (lane == 0) T.last[row] = t - 1;
}

// ---------------------------------------------------------------------------------------
// k_forward: one warp per slot.  pred[j] = sigmoid(<U[u],V[i]> + bu + bi); optionally snapshots
// both rows (the values every gradient of this step must be computed from).
// ---------------------------------------------------------------------------------------
template <int VEC, int NIT>
__global__ void __launch_bounds__(BLOCK_THREADS) k_forward(const int *__restrict__ slot_u,
                                                           const int *__restrict__ slot_i, int L, TableView users,
                                                           TableView items, int D, float *__restrict__ snap_u,
                                                           float *__restrict__ snap_i, float *__restrict__ pred) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (j >= L) return;
  const long long u = slot_u[j], i = slot_i[j];
  Frag<VEC, NIT> fu, fi;
  frag_load<VEC, NIT>(fu, users.p + u * D, D, lane);
  frag_load<VEC, NIT>(fi, items.p + i * D, D, lane);
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < NIT * VEC; ++k) acc = fmaf(fu.x[k], fi.x[k], acc);
  acc = warp_sum(acc);
  if (snap_u != nullptr) {
    frag_store<VEC, NIT>(fu, snap_u + (long long)j * D, D, lane);
    frag_store<VEC, NIT>(fi, snap_i + (long long)j * D, D, lane);
  }
  if (lane == 0) {
    float z = (acc + users.bp[u]) + items.bp[i];
    pred[j] = sigmoidf_acc(z);
  }
}

// predict(user) against every item: one warp per item, user row shared
template <int VEC, int NIT>
__global__ void __launch_bounds__(BLOCK_THREADS) k_forward_user(long long u, int num_items, TableView users,
                                                                TableView items, int D, float *__restrict__ pred) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (i >= num_items) return;
  Frag<VEC, NIT> fu, fi;
  frag_load<VEC, NIT>(fu, users.p + u * D, D, lane);
  frag_load<VEC, NIT>(fi, items.p + (long long)i * D, D, lane);
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < NIT * VEC; ++k) acc = fmaf(fu.x[k], fi.x[k], acc);
  acc = warp_sum(acc);
  if (lane == 0) pred[i] = sigmoidf_acc((acc + users.bp[u]) + items.bp[i]);
}

__global__ void k_pack_pairs(const long long *__restrict__ users, const long long *__restrict__ items, long long n,
                             int num_users, int num_items, int *__restrict__ slot_u, int *__restrict__ slot_i,
                             int *__restrict__ err_flag) {
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  long long u = users[g], i = items[g];
  if (u < 0 || u >= num_users || i < 0 || i >= num_items) {
    atomicExch(err_flag, 1);
    u = 0;
    i = 0;
  }
  slot_u[g] = (int)u;
  slot_i[g] = (int)i;
}

// ---------------------------------------------------------------------------------------
// k_loss: one CTA.  Loss value + dLoss/dpred (or dLoss/dz when to_logit) for 1-D pos[b], neg[m]
// probabilities; formulas are torch's backward of spotlight/losses.py.
// ---------------------------------------------------------------------------------------
constexpr int LOSS_THREADS = 1024;

__device__ double block_sum(double v, double *sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  double tot = 0.0;
  for (int w = 0; w < LOSS_THREADS / 32; ++w) tot += sh[w];  // fixed order -> deterministic
  return tot;
}

__global__ void __launch_bounds__(LOSS_THREADS) k_loss(int kind, const float *__restrict__ pos, int b,
                                                       const float *__restrict__ neg, int m,
                                                       float *__restrict__ loss_out, float *__restrict__ dpos,
                                                       float *__restrict__ dneg, int to_logit) {
  __shared__ double sh[LOSS_THREADS / 32];
  __shared__ float sh_max[LOSS_THREADS / 32];
  __shared__ int sh_arg[LOSS_THREADS / 32];
  const int tid = threadIdx.x;
  const float inv_b = 1.0f / (float)b;
  const bool grad = dpos != nullptr;
  if (kind == MFB_LOSS_POINTWISE) {
    double sp = 0.0, sn = 0.0;
    for (int j = tid; j < b; j += LOSS_THREADS) {
      float x = pos[j];
      sp += (double)(-fmaxf(logf(x), -100.0f));
      if (grad) {
        float d = __fdiv_rn(__fdiv_rn(x - 1.0f, fmaxf((1.0f - x) * x, 1e-12f)), (float)b);
        dpos[j] = to_logit ? (d * (1.0f - x)) * x : d;
      }
    }
    for (int j = tid; j < m; j += LOSS_THREADS) {
      float x = neg[j];
      sn += (double)(-fmaxf(logf(1.0f - x), -100.0f));
      if (grad) {
        float d = __fdiv_rn(__fdiv_rn(x, fmaxf((1.0f - x) * x, 1e-12f)), (float)m);
        dneg[j] = to_logit ? (d * (1.0f - x)) * x : d;
      }
    }
    sp = block_sum(sp, sh);
    sn = block_sum(sn, sh);
    if (tid == 0) {
      float l = (float)(sp / (double)b);
      if (m > 0) l = l + (float)(sn / (double)m);
      *loss_out = l;
    }
  } else if (kind == MFB_LOSS_HINGE || kind == MFB_LOSS_BPR) {
    double s = 0.0;
    for (int j = tid; j < b; j += LOSS_THREADS) {
      float xp = pos[j], xn = neg[j];
      float dp, dn;
      if (kind == MFB_LOSS_HINGE) {
        float d = (xn - xp) + 1.0f;
        s += (double)fmaxf(d, 0.0f);
        float a = (d >= 0.0f) ? inv_b : 0.0f;
        dp = -a;
        dn = a;
      } else {
        float sg = sigmoidf_acc(xp - xn);
        s += (double)(1.0f - sg);
        float g = ((-inv_b) * (1.0f - sg)) * sg;
        dp = g;
        dn = -g;
      }
      if (grad) {
        dpos[j] = to_logit ? (dp * (1.0f - xp)) * xp : dp;
        dneg[j] = to_logit ? (dn * (1.0f - xn)) * xn : dn;
      }
    }
    s = block_sum(s, sh);
    if (tid == 0) *loss_out = (float)(s / (double)b);
  } else {  // adaptive hinge, 1-D negatives: hinge against the first global maximum
    float best = -INFINITY;
    int arg = 0x7fffffff;
    for (int j = tid; j < m; j += LOSS_THREADS) {
      float x = neg[j];
      if (x > best) {  // strided ascending j: first occurrence kept per thread
        best = x;
        arg = j;
      }
    }
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ob = __shfl_xor_sync(0xffffffffu, best, o);
      int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ob > best || (ob == best && oa < arg)) {
        best = ob;
        arg = oa;
      }
    }
    if (lane == 0) {
      sh_max[wid] = best;
      sh_arg[wid] = arg;
    }
    __syncthreads();
    best = sh_max[0];
    arg = sh_arg[0];
    for (int w = 1; w < LOSS_THREADS / 32; ++w) {
      if (sh_max[w] > best || (sh_max[w] == best && sh_arg[w] < arg)) {
        best = sh_max[w];
        arg = sh_arg[w];
      }
    }
    double s = 0.0, cnt = 0.0;
    for (int j = tid; j < b; j += LOSS_THREADS) {
      float xp = pos[j];
      float d = (best - xp) + 1.0f;
      s += (double)fmaxf(d, 0.0f);
      bool act = d >= 0.0f;
      cnt += act ? 1.0 : 0.0;
      if (grad) {
        float dp = act ? -inv_b : 0.0f;
        dpos[j] = to_logit ? (dp * (1.0f - xp)) * xp : dp;
      }
    }
    if (grad)
      for (int j = tid; j < m; j += LOSS_THREADS) dneg[j] = 0.0f;
    s = block_sum(s, sh);
    cnt = block_sum(cnt, sh);
    if (tid == 0) {
      *loss_out = (float)(s / (double)b);
      if (grad) {
        float dn = (float)cnt * inv_b;
        dneg[arg] = to_logit ? (dn * (1.0f - best)) * best : dn;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// k_update: one warp per sorted position; segment heads reduce the slot gradients of their row
// in slot order and apply optimiser step t.
//   d/dU[u] += dz_j * V_old[i_j]   d/dbu[u] += dz_j   (and symmetrically for items)
// ---------------------------------------------------------------------------------------
template <int VEC, int NIT, bool FAST>
__global__ void __launch_bounds__(BLOCK_THREADS) k_update(const uint32_t *__restrict__ skeys,
                                                          const uint32_t *__restrict__ svals, int n, int rb,
                                                          TableView users, TableView items, OptView opt, int D,
                                                          const float *__restrict__ snap_u,
                                                          const float *__restrict__ snap_i,
                                                          const float *__restrict__ dz, int t) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (q >= n) return;
  const uint32_t key = skeys[q];
  if (q > 0 && skeys[q - 1] == key) return;
  const long long row = key & ((1u << rb) - 1u);
  const bool is_item = (key >> rb) & 1u;
  const TableView &T = is_item ? items : users;
  const float *__restrict__ other = is_item ? snap_u : snap_i;
  const bool adam = opt.kind == MFB_OPT_ADAM;

  Frag<VEC, NIT> g;
#pragma unroll
  for (int k = 0; k < NIT * VEC; ++k) g.x[k] = 0.f;
  float gb = 0.f;
  for (int q2 = q; q2 < n && skeys[q2] == key; ++q2) {
    const int j = (int)svals[q2];
    const float d = dz[j];
    if (d != 0.f) {
      Frag<VEC, NIT> o;
      frag_load<VEC, NIT>(o, other + (long long)j * D, D, lane);
#pragma unroll
      for (int k = 0; k < NIT * VEC; ++k) g.x[k] = __fadd_rn(g.x[k], __fmul_rn(d, o.x[k]));
      gb = __fadd_rn(gb, d);
    }
  }
  RowState<VEC, NIT> r;
  row_load<VEC, NIT>(r, T, row, D, lane, adam);
  if (adam) {
    const float neg_ss = -__ldg(opt.step_size + t);
    const float bc2 = __ldg(opt.bc2_sqrt + t);
#pragma unroll
    for (int k = 0; k < NIT * VEC; ++k) adam_elem<FAST>(r.p.x[k], r.m.x[k], r.v.x[k], g.x[k], neg_ss, bc2, opt);
    adam_elem<FAST>(r.bp, r.bm, r.bv, gb, neg_ss, bc2, opt);
  } else {
#pragma unroll
    for (int k = 0; k < NIT * VEC; ++k) sgd_elem(r.p.x[k], g.x[k], opt);
    sgd_elem(r.bp, gb, opt);
  }
  row_store<VEC, NIT>(r, T, row, D, lane, adam);
  if (lane == 0) T.last[row] = t;
}

// k_flush: one warp per row of one table; replay (last, t]
template <int VEC, int NIT, bool FAST>
__global__ void __launch_bounds__(BLOCK_THREADS) k_flush(TableView T, OptView opt, int D, int t) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= T.rows) return;
  const int last = T.last[row];
  if (last >= t) return;
  const bool adam = opt.kind == MFB_OPT_ADAM;
  RowState<VEC, NIT> r;
  row_load<VEC, NIT>(r, T, row, D, lane, adam);
  row_replay<VEC, NIT, FAST>(r, last, t, opt);
  row_store<VEC, NIT>(r, T, row, D, lane, adam);
  if (lane == 0) T.last[row] = t;
}

// ---------------------------------------------------------------------------------------
// dispatch on (vector width, iterations per lane, fast math)
// ---------------------------------------------------------------------------------------
struct Shape {
  int vec, nit;
};

int pick_shape(int D, Shape *s) {
  if (D <= 0 || D > 512) {
    mfb_set_error("embedding_dim %d unsupported (1..512)", D);
    return MFB_ERR_UNSUPPORTED;
  }
  if (D % 4 == 0) {
    s->vec = 4;
    int per = (D + 127) / 128;
    s->nit = per <= 1 ? 1 : (per <= 2 ? 2 : 4);
  } else {
    s->vec = 1;
    int per = (D + 31) / 32;
    s->nit = per <= 2 ? 2 : (per <= 4 ? 4 : (per <= 8 ? 8 : 16));
  }
  return MFB_OK;
}

#define MFB_DISPATCH_SHAPE(SH, CALL)                       \
  do {                                                     \
    if ((SH).vec == 4) {                                   \
      if ((SH).nit == 1) { CALL(4, 1); }                   \
      else if ((SH).nit == 2) { CALL(4, 2); }              \
      else { CALL(4, 4); }                                 \
    } else {                                               \
      if ((SH).nit == 2) { CALL(1, 2); }                   \
      else if ((SH).nit == 4) { CALL(1, 4); }              \
      else if ((SH).nit == 8) { CALL(1, 8); }              \
      else { CALL(1, 16); }                                \
    }                                                      \
  } while (0)

inline int grid_for_warps(long long warps) { return (int)((warps + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK); }

int bits_for(uint32_t maxval) {
  int b = 1;
  while (b < 32 && (maxval >> b) != 0) ++b;
  return b;
}

int check_err_flag(int *d_flag, cudaStream_t st, const char *what) {
  int h = 0;
  MFB_CUDA(cudaMemcpyAsync(&h, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  MFB_CUDA(cudaStreamSynchronize(st));
  if (h) {
    mfb_set_error("%s: id out of range for the model's tables", what);
    return MFB_ERR_RANGE;
  }
  return MFB_OK;
}

int launch_flush(mfb_model *m, cudaStream_t st) {
  if (m->flushed_step == m->step) return MFB_OK;
  Shape sh;
  MFB_CHECK(pick_shape(m->desc.dim, &sh));
  MFB_CHECK(mfb_ensure_scalars(m, m->step));
  const int D = m->desc.dim, t = (int)m->step;
  const bool fast = m->desc.fast_math != 0;
#define CALL(V, N)                                                                                      \
  if (fast) {                                                                                           \
    k_flush<V, N, true><<<grid_for_warps(m->users.rows), BLOCK_THREADS, 0, st>>>(m->users, m->opt, D, t); \
    k_flush<V, N, true><<<grid_for_warps(m->items.rows), BLOCK_THREADS, 0, st>>>(m->items, m->opt, D, t); \
  } else {                                                                                              \
    k_flush<V, N, false><<<grid_for_warps(m->users.rows), BLOCK_THREADS, 0, st>>>(m->users, m->opt, D, t); \
    k_flush<V, N, false><<<grid_for_warps(m->items.rows), BLOCK_THREADS, 0, st>>>(m->items, m->opt, D, t); \
  }
  MFB_DISPATCH_SHAPE(sh, CALL);
#undef CALL
  MFB_KERNEL_CHECK();
  m->flushed_step = m->step;
  return MFB_OK;
}

int validate_loss_shape(int loss, int64_t n_pos, int batch, int n_neg) {
  if (loss < MFB_LOSS_POINTWISE || loss > MFB_LOSS_ADAPTIVE_HINGE) {
    mfb_set_error("unknown loss %d", loss);
    return MFB_ERR_INVALID;
  }
  if (n_pos <= 0 || batch <= 0 || n_neg < 0) {
    mfb_set_error("bad batching (n_pos=%lld batch=%d n_neg=%d)", (long long)n_pos, batch, n_neg);
    return MFB_ERR_INVALID;
  }
  if (n_neg == 0 && loss != MFB_LOSS_POINTWISE) {
    mfb_set_error("loss %d needs negative predictions", loss);
    return MFB_ERR_SHAPE;
  }
  if (loss == MFB_LOSS_HINGE || loss == MFB_LOSS_BPR) {
    // elementwise neg - pos on 1-D tensors of lengths n_neg*batch and b: torch broadcasting
    // (losses.py:88,121) only accepts equal lengths here
    if (n_neg != 1 || n_pos % batch != 0) {
      mfb_set_error("hinge/bpr: len(negatives)=%d*%d must equal len(positives) in every batch (n_pos=%lld)", n_neg,
                    batch, (long long)n_pos);
      return MFB_ERR_SHAPE;
    }
  }
  return MFB_OK;
}

}  // namespace

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" int mfb_flush(mfb_model *m, mfb_stream stream) {
  if (!m) return MFB_ERR_INVALID;
  return launch_flush(m, (cudaStream_t)stream);
}

extern "C" int mfb_loss_forward_backward(int loss, const float *d_pos, int64_t n_pos, const float *d_neg,
                                         int64_t n_neg, float *d_loss, float *d_dpos, float *d_dneg,
                                         mfb_stream stream) {
  if (!d_pos || !d_loss || n_pos <= 0 || n_neg < 0 || (n_neg > 0 && !d_neg)) return MFB_ERR_INVALID;
  if (loss < MFB_LOSS_POINTWISE || loss > MFB_LOSS_ADAPTIVE_HINGE) return MFB_ERR_INVALID;
  if (n_neg == 0 && loss != MFB_LOSS_POINTWISE) return MFB_ERR_SHAPE;
  if ((loss == MFB_LOSS_HINGE || loss == MFB_LOSS_BPR) && n_neg != n_pos) {
    mfb_set_error("hinge/bpr: len(neg)=%lld != len(pos)=%lld", (long long)n_neg, (long long)n_pos);
    return MFB_ERR_SHAPE;
  }
  if ((d_dpos == nullptr) != (d_dneg == nullptr) && n_neg > 0) return MFB_ERR_INVALID;
  k_loss<<<1, LOSS_THREADS, 0, (cudaStream_t)stream>>>(loss, d_pos, (int)n_pos, d_neg, (int)n_neg, d_loss, d_dpos,
                                                       d_dneg, 0);
  MFB_KERNEL_CHECK();
  return MFB_OK;
}

extern "C" int mfb_predict_pairs(mfb_model *m, const int64_t *d_users, const int64_t *d_items, int64_t count,
                                 float *d_out, mfb_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!m || !d_users || !d_items || !d_out || count < 0) return MFB_ERR_INVALID;
  if (count == 0) return MFB_OK;
  MFB_CHECK(launch_flush(m, st));
  Shape sh;
  MFB_CHECK(pick_shape(m->desc.dim, &sh));
  MFB_CHECK(m->ws_ids.reserve((size_t)count * 2 * sizeof(int) + 16));
  MFB_CHECK(m->ws_scalars.reserve(64));
  int *slot_u = m->ws_ids.as<int>();
  int *slot_i = slot_u + count;
  int *flag = m->ws_scalars.as<int>();
  MFB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
  k_pack_pairs<<<(unsigned)((count + 255) / 256), 256, 0, st>>>((const long long *)d_users, (const long long *)d_items,
                                                               count, m->users.rows, m->items.rows, slot_u, slot_i,
                                                               flag);
  MFB_KERNEL_CHECK();
  MFB_CHECK(check_err_flag(flag, st, "predict"));
  const int D = m->desc.dim;
  // chunk so the per-launch slot count stays an int
  for (int64_t off = 0; off < count; off += (1 << 24)) {
    int L = (int)((count - off < (1 << 24)) ? (count - off) : (1 << 24));
#define CALL(V, N)                                                                                               \
  k_forward<V, N><<<grid_for_warps(L), BLOCK_THREADS, 0, st>>>(slot_u + off, slot_i + off, L, m->users, m->items, D, \
                                                                nullptr, nullptr, d_out + off)
    MFB_DISPATCH_SHAPE(sh, CALL);
#undef CALL
    MFB_KERNEL_CHECK();
  }
  return MFB_OK;
}

extern "C" int mfb_predict_user(mfb_model *m, int64_t user, float *d_out, mfb_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!m || !d_out) return MFB_ERR_INVALID;
  if (user < 0 || user >= m->users.rows) {
    mfb_set_error("predict: user id %lld out of range", (long long)user);
    return MFB_ERR_RANGE;
  }
  MFB_CHECK(launch_flush(m, st));
  Shape sh;
  MFB_CHECK(pick_shape(m->desc.dim, &sh));
  const int D = m->desc.dim, I = m->items.rows;
#define CALL(V, N) \
  k_forward_user<V, N><<<grid_for_warps(I), BLOCK_THREADS, 0, st>>>((long long)user, I, m->users, m->items, D, d_out)
  MFB_DISPATCH_SHAPE(sh, CALL);
#undef CALL
  MFB_KERNEL_CHECK();
  return MFB_OK;
}

static int run_steps(mfb_model *m, int loss, const int64_t *d_pos_users, const int64_t *d_pos_items, int64_t n_pos,
                     int32_t batch, int32_t n_neg, const int64_t *d_neg_users, const int64_t *d_neg_items,
                     float *d_step_losses, cudaStream_t st, bool train) {
  if (!m || !d_pos_users || !d_pos_items || !d_step_losses) return MFB_ERR_INVALID;
  MFB_CHECK(validate_loss_shape(loss, n_pos, batch, n_neg));
  if (n_neg > 0 && (!d_neg_users || !d_neg_items)) return MFB_ERR_INVALID;
  Shape sh;
  MFB_CHECK(pick_shape(m->desc.dim, &sh));
  const int D = m->desc.dim;
  const int64_t m_neg64 = (int64_t)n_neg * batch;
  if (m_neg64 + batch > (1 << 24)) {
    mfb_set_error("batch*(1+n_neg) too large");
    return MFB_ERR_UNSUPPORTED;
  }
  const int m_neg = (int)m_neg64;
  const int Lfull = batch + m_neg;
  const int64_t nsteps = (n_pos + batch - 1) / batch;
  const uint32_t maxrow = (uint32_t)((m->users.rows > m->items.rows ? m->users.rows : m->items.rows) - 1);
  const int rb = bits_for(maxrow);
  int sb_max = 31 - rb;  // bits left for the chunk-local step
  if (sb_max < 0) {
    mfb_set_error("tables too large for 32-bit planner keys");
    return MFB_ERR_UNSUPPORTED;
  }
  int chunk = 1 << (sb_max > 6 ? 6 : sb_max);
  while (chunk > 1 && (int64_t)chunk * Lfull > (1ll << 25)) chunk >>= 1;  // bound workspace
  if (!train) MFB_CHECK(launch_flush(m, st));
  if (train) MFB_CHECK(mfb_ensure_scalars(m, m->step + nsteps));

  const size_t slots = (size_t)chunk * Lfull;
  MFB_CHECK(m->ws_slots.reserve(slots * 2 * sizeof(int)));
  MFB_CHECK(m->ws_pred.reserve((size_t)Lfull * sizeof(float)));
  MFB_CHECK(m->ws_scalars.reserve(64));
  int *slot_u = m->ws_slots.as<int>();
  int *slot_i = slot_u + slots;
  int *flag = m->ws_scalars.as<int>();
  MFB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
  if (train) {
    MFB_CHECK(m->ws_keys_a.reserve(slots * 2 * sizeof(uint32_t)));
    MFB_CHECK(m->ws_keys_b.reserve(slots * 2 * sizeof(uint32_t)));
    MFB_CHECK(m->ws_vals_a.reserve(slots * 2 * sizeof(uint32_t)));
    MFB_CHECK(m->ws_vals_b.reserve(slots * 2 * sizeof(uint32_t)));
    MFB_CHECK(m->ws_rows.reserve((size_t)Lfull * D * 2 * sizeof(float)));
    MFB_CHECK(m->ws_dz.reserve((size_t)Lfull * sizeof(float)));
  } else {
    // keys are still written by k_pack; give it somewhere to put them
    MFB_CHECK(m->ws_keys_a.reserve(slots * 2 * sizeof(uint32_t)));
    MFB_CHECK(m->ws_vals_a.reserve(slots * 2 * sizeof(uint32_t)));
  }
  float *snap_u = train ? m->ws_rows.as<float>() : nullptr;
  float *snap_i = train ? snap_u + (size_t)Lfull * D : nullptr;
  float *pred = m->ws_pred.as<float>();
  float *dz = train ? m->ws_dz.as<float>() : nullptr;
  const bool fast = m->desc.fast_math != 0;

  for (int64_t s0 = 0; s0 < nsteps; s0 += chunk) {
    const int ns = (int)((nsteps - s0 < chunk) ? (nsteps - s0) : chunk);
    const long long total = (long long)ns * Lfull;
    k_pack<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        (const long long *)d_pos_users, (const long long *)d_pos_items, n_pos, (const long long *)d_neg_users,
        (const long long *)d_neg_items, batch, m_neg, s0, ns, rb, m->users.rows, m->items.rows, slot_u, slot_i,
        m->ws_keys_a.as<uint32_t>(), m->ws_vals_a.as<uint32_t>(), flag);
    MFB_KERNEL_CHECK();
    // number of keys actually present: all steps full except possibly the last of the epoch
    const int64_t last_first = (s0 + ns - 1) * (int64_t)batch;
    const int b_last = (int)((n_pos - last_first < batch) ? (n_pos - last_first) : batch);
    const int64_t nkeys = 2 * ((int64_t)(ns - 1) * Lfull + (b_last + m_neg));
    uint32_t *skeys = nullptr, *svals = nullptr;
    if (train) {
      int nbits = rb + 1 + bits_for((uint32_t)(ns > 1 ? ns - 1 : 1));
      if (ns == 1) nbits = rb + 1;
      MFB_CHECK(mfb_radix_sort_pairs(m->ws_keys_a.as<uint32_t>(), m->ws_vals_a.as<uint32_t>(),
                                     m->ws_keys_b.as<uint32_t>(), m->ws_vals_b.as<uint32_t>(), nkeys, nbits,
                                     m->ws_hist, &skeys, &svals, st));
    }
    for (int s = 0; s < ns; ++s) {
      const int64_t gstep = s0 + s;
      const int b = (s == ns - 1) ? b_last : batch;
      const int L = b + m_neg;
      const int *su = slot_u + (size_t)s * Lfull;
      const int *si = slot_i + (size_t)s * Lfull;
      if (train) {
        const int t = (int)(m->step + 1);
        const uint32_t *k = skeys + 2 * (size_t)s * Lfull;
        const uint32_t *v = svals + 2 * (size_t)s * Lfull;
        const int nk = 2 * L;
#define CALL(V, N)                                                                                                 \
  if (fast)                                                                                                        \
    k_catchup<V, N, true><<<grid_for_warps(nk), BLOCK_THREADS, 0, st>>>(k, nk, rb, m->users, m->items, m->opt, D, t); \
  else                                                                                                             \
    k_catchup<V, N, false><<<grid_for_warps(nk), BLOCK_THREADS, 0, st>>>(k, nk, rb, m->users, m->items, m->opt, D, t)
        MFB_DISPATCH_SHAPE(sh, CALL);
#undef CALL
#define CALL(V, N) \
  k_forward<V, N><<<grid_for_warps(L), BLOCK_THREADS, 0, st>>>(su, si, L, m->users, m->items, D, snap_u, snap_i, pred)
        MFB_DISPATCH_SHAPE(sh, CALL);
#undef CALL
        k_loss<<<1, LOSS_THREADS, 0, st>>>(loss, pred, b, pred + b, m_neg, d_step_losses + gstep, dz, dz + b, 1);
#define CALL(V, N)                                                                                                   \
  if (fast)                                                                                                          \
    k_update<V, N, true><<<grid_for_warps(nk), BLOCK_THREADS, 0, st>>>(k, v, nk, rb, m->users, m->items, m->opt, D,  \
                                                                        snap_u, snap_i, dz, t);                      \
  else                                                                                                               \
    k_update<V, N, false><<<grid_for_warps(nk), BLOCK_THREADS, 0, st>>>(k, v, nk, rb, m->users, m->items, m->opt, D, \
                                                                         snap_u, snap_i, dz, t)
        MFB_DISPATCH_SHAPE(sh, CALL);
#undef CALL
        MFB_KERNEL_CHECK();
        m->step += 1;
      } else {
#define CALL(V, N) \
  k_forward<V, N><<<grid_for_warps(L), BLOCK_THREADS, 0, st>>>(su, si, L, m->users, m->items, D, nullptr, nullptr, pred)
        MFB_DISPATCH_SHAPE(sh, CALL);
#undef CALL
        k_loss<<<1, LOSS_THREADS, 0, st>>>(loss, pred, b, pred + b, m_neg, d_step_losses + gstep, nullptr, nullptr, 0);
        MFB_KERNEL_CHECK();
      }
    }
  }
  // ids were clamped on the device; report a bad id once, after the queue drains
  MFB_CHECK(check_err_flag(flag, st, train ? "train" : "loss"));
  return MFB_OK;
}

extern "C" int mfb_train_steps(mfb_model *m, int loss, const int64_t *d_pos_users, const int64_t *d_pos_items,
                               int64_t n_pos, int32_t batch, int32_t n_neg, const int64_t *d_neg_users,
                               const int64_t *d_neg_items, float *d_step_losses, mfb_stream stream) {
  return run_steps(m, loss, d_pos_users, d_pos_items, n_pos, batch, n_neg, d_neg_users, d_neg_items, d_step_losses,
                   (cudaStream_t)stream, true);
}

extern "C" int mfb_loss_steps(mfb_model *m, int loss, const int64_t *d_pos_users, const int64_t *d_pos_items,
                              int64_t n_pos, int32_t batch, int32_t n_neg, const int64_t *d_neg_users,
                              const int64_t *d_neg_items, float *d_step_losses, mfb_stream stream) {
  return run_steps(m, loss, d_pos_users, d_pos_items, n_pos, batch, n_neg, d_neg_users, d_neg_items, d_step_losses,
                   (cudaStream_t)stream, false);
}

extern "C" int mfb_train_epoch_host(mfb_model *m, int loss, const int64_t *h_pos_users, const int64_t *h_pos_items,
                                    int64_t n_pos, int32_t batch, int32_t n_neg, uint32_t *h_state,
                                    const int64_t *d_pop_users, const int64_t *d_pop_items, int64_t pop_len,
                                    float *h_step_losses, mfb_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!m || !h_pos_users || !h_pos_items || !h_step_losses) return MFB_ERR_INVALID;
  MFB_CHECK(validate_loss_shape(loss, n_pos, batch, n_neg));
  const int64_t nsteps = (n_pos + batch - 1) / batch;
  const int64_t k = nsteps * (int64_t)n_neg * batch;
  MFB_CHECK(m->ws_ids.reserve((size_t)n_pos * 2 * sizeof(int64_t)));
  MFB_CHECK(m->ws_losses.reserve((size_t)nsteps * sizeof(float)));
  int64_t *d_u = m->ws_ids.as<int64_t>();
  int64_t *d_i = d_u + n_pos;
  MFB_CUDA(cudaMemcpyAsync(d_u, h_pos_users, (size_t)n_pos * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  MFB_CUDA(cudaMemcpyAsync(d_i, h_pos_items, (size_t)n_pos * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  int64_t *d_nu = nullptr, *d_ni = nullptr;
  if (k > 0) {
    if (!h_state || !d_pop_users || !d_pop_items || pop_len <= 0) return MFB_ERR_INVALID;
    MFB_CHECK(m->ws_neg_u.reserve((size_t)k * sizeof(int64_t)));
    MFB_CHECK(m->ws_neg_i.reserve((size_t)k * sizeof(int64_t)));
    d_nu = m->ws_neg_u.as<int64_t>();
    d_ni = m->ws_neg_i.as<int64_t>();
    MFB_CHECK(mfb_mt_choices_pairs(h_state, d_pop_users, d_pop_items, pop_len, k, d_nu, d_ni, stream));
  }
  MFB_CHECK(run_steps(m, loss, d_u, d_i, n_pos, batch, n_neg, d_nu, d_ni, m->ws_losses.as<float>(), st, true));
  MFB_CUDA(cudaMemcpyAsync(h_step_losses, m->ws_losses.ptr, (size_t)nsteps * sizeof(float), cudaMemcpyDeviceToHost, st));
  MFB_CUDA(cudaStreamSynchronize(st));
  return MFB_OK;
}
