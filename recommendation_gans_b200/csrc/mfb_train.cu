// Fused implicit-MF training step for sm_100a.
//
// Replaces ImplicitFactorizationModel.run_train_iteration (implicit.py:347-364): BilinearNet
// forward on the positive batch and on the sampled negative pairs (representations.py:80-91),
// the ranking loss (spotlight/losses.py:20-172), autograd's dense embedding backward, and the
// dense torch optimiser step (spotlight/optimizers.py -> torch.optim.Adam / SGD).
//
// Semantics are the reference's DENSE optimiser (every row of every table steps every
// iteration); traffic is row-sparse.  Each row carries `last` = the optimiser step it
// is current for; the zero-gradient steps it misses (gradient = weight_decay * p) are replayed
// elementwise in registers:
//   * eagerly, right after the row is updated at step t: the planner knows the row's next use
//     (t + gap) inside the chunk, so k_update replays it to t + gap - 1 while it still holds
//     the state -- no extra pass over the row;
//   * lazily, for a row's first use in a chunk (or a gap longer than the eager bound): the
//     planner lists those rows per step, and dedicated blocks of k_update(s) bring the rows of
//     step s+1's list up to date while step s updates its own (disjoint) rows;
//   * at mfb_flush, for everything else.
//
// Pipeline
//   planner stream  sample negatives (MT19937) -> pack ids -> stable radix sort by (step,table,row)
//                   -> segment table -> re-sort by (table,row) -> next-use gaps + lazy lists -> one
//                   16-byte record per sorted position.  Chunk c+1 is planned while chunk c trains.
//   main stream     per step: k_forward (gather, dot, sigmoid, row snapshots, adaptive-hinge max)
//                   -> k_update (per unique row: batch-ordered segment reduction of slot gradients,
//                   optimiser step t, eager replay; + lazy catch-up blocks for step s+1).
//                   per chunk: k_loss_steps (all loss values in one launch).
// Unique rows / segments come from a stable sort, so the reduction order is the batch order used by
// torch's index_add; no floating-point atomics anywhere.
#include <math.h>

#include "mfb_internal.cuh"
#include "mfb_rowops.cuh"

namespace {

constexpr int WARPS_PER_BLOCK = 8;
constexpr int BLOCK_THREADS = WARPS_PER_BLOCK * 32;

// Programmatic dependent launch: the per-step kernels are launched back to back on one stream with
// programmatic stream serialization, so a kernel's CTAs start (and load their planner records) while the
// previous kernel drains; everything that depends on the previous kernel's output comes after pdl_wait().
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------------------
// planner: pack slot ids and sort keys for a chunk of steps
// ---------------------------------------------------------------------------------------
// Step s (chunk-local) has b_s positives followed by m negatives; slots of step s start at
// s*Lfull (only the final step of an epoch can be partial and it is last in its chunk).
// Key = step << (rb+1) | table << rb | row; keys of a step are laid out [users of all slots]
// [items of all slots] so a stable sort leaves each (step,table,row) group in slot order.
__global__ void k_pack(const long long *__restrict__ pos_u, const long long *__restrict__ pos_i, long long n_pos,
                       const long long *__restrict__ neg_u, const long long *__restrict__ neg_i, int batch, int m_neg,
                       long long step0, long long neg_step0, int nsteps, int rb, int num_users, int num_items,
                       int *__restrict__ slot_u,
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
    long long q = (gstep - neg_step0) * m_neg + (j - b);
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

// When neg_begin >= 0 (adaptive hinge), the maximum over slots j >= neg_begin is reduced per block
// and folded into *gmax with one (usually skipped) 64-bit atomicMax per block.
// A warp owns FWD_SPW consecutive slots and issues the row loads of all of them before the first use: the kernel is
// bound by dependent-load latency (ids -> rows), so loads in flight per warp are what counts.
#ifndef MFB_FWD_SPW
#define MFB_FWD_SPW 2
#endif
constexpr int FWD_SPW = MFB_FWD_SPW;
#ifndef MFB_FWD_PREFETCH
#define MFB_FWD_PREFETCH 0   // measured on B200 (cfg3): update 35.4 -> 34.7 us, forward 13.3 -> 16.0 us per step: a net loss
#endif

template <int VEC, int NIT>
__global__ void __launch_bounds__(BLOCK_THREADS) k_forward(const int *__restrict__ slot_u,
                                                           const int *__restrict__ slot_i, int L, TableView users,
                                                           TableView items, int D, float *__restrict__ snap_u,
                                                           float *__restrict__ snap_i, float *__restrict__ pred,
                                                           int neg_begin, unsigned long long *__restrict__ gmax) {
  __shared__ unsigned long long wmax[WARPS_PER_BLOCK];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int j0 = (blockIdx.x * WARPS_PER_BLOCK + wid) * FWD_SPW;
  unsigned long long mine = 0ull;
  pdl_launch_dependents();
  long long u[FWD_SPW], i[FWD_SPW];
#pragma unroll
  for (int s = 0; s < FWD_SPW; ++s) {
    const bool ok = j0 + s < L;
    u[s] = ok ? slot_u[j0 + s] : 0;       // planner output: independent of the previous kernel
    i[s] = ok ? slot_i[j0 + s] : 0;
  }
  pdl_wait();            // tables are written by the previous step's update
  if (j0 < L) {
    Frag<VEC, NIT> fu[FWD_SPW], fi[FWD_SPW];
    float bu[FWD_SPW], bi[FWD_SPW];
#pragma unroll
    for (int s = 0; s < FWD_SPW; ++s) {   // (slots past L re-read row 0: harmless, never stored)
      frag_load<VEC, NIT>(fu[s], users.p + u[s] * D, D, lane);
      frag_load<VEC, NIT>(fi[s], items.p + i[s] * D, D, lane);
      bu[s] = users.bp[u[s]];
      bi[s] = items.bp[i[s]];
    }
    if (MFB_FWD_PREFETCH && snap_u != nullptr && users.m != nullptr) {
      // training step: the update kernel that follows reads the optimiser moments of exactly these rows; asking L2 for
      // them now takes the DRAM round trip out of its dependent-load chain
#pragma unroll
      for (int s = 0; s < FWD_SPW; ++s) {
        const int e = lane * 32;                      // one 128-byte line per lane
        if (e < D) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(users.m + u[s] * D + e));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(users.v + u[s] * D + e));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(items.m + i[s] * D + e));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(items.v + i[s] * D + e));
        }
      }
    }
#pragma unroll
    for (int s = 0; s < FWD_SPW; ++s) {
      const int j = j0 + s;
      if (j >= L) break;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < NIT * VEC; ++k) acc = fmaf(fu[s].x[k], fi[s].x[k], acc);
      acc = warp_sum(acc);
      if (snap_u != nullptr) {
        frag_store<VEC, NIT>(fu[s], snap_u + (long long)j * D, D, lane);
        frag_store<VEC, NIT>(fi[s], snap_i + (long long)j * D, D, lane);
      }
      const float z = (acc + bu[s]) + bi[s];
      const float y = sigmoidf_acc(z);
      if (lane == 0) pred[j] = y;
      if (neg_begin >= 0 && j >= neg_begin) {
        const unsigned long long cand = pack_max(y, j - neg_begin);
        mine = cand > mine ? cand : mine;
      }
    }
  }
  if (neg_begin >= 0) {  // uniform across the grid
    if (lane == 0) wmax[wid] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long best = wmax[0];
#pragma unroll
      for (int w = 1; w < WARPS_PER_BLOCK; ++w) best = wmax[w] > best ? wmax[w] : best;
      if (best > *reinterpret_cast<volatile unsigned long long *>(gmax)) atomicMax(gmax, best);
    }
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

__device__ void loss_block(int kind, const float *__restrict__ pos, int b, const float *__restrict__ neg, int m,
                           float *__restrict__ loss_out, float *__restrict__ dpos, float *__restrict__ dneg,
                           int to_logit) {
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


__global__ void __launch_bounds__(LOSS_THREADS) k_loss(int kind, const float *__restrict__ pos, int b,
                                                       const float *__restrict__ neg, int m,
                                                       float *__restrict__ loss_out, float *__restrict__ dpos,
                                                       float *__restrict__ dneg, int to_logit) {
  loss_block(kind, pos, b, neg, m, loss_out, dpos, dneg, to_logit);
}

// ---------------------------------------------------------------------------------------
// k_loss_ex: the function-level forms the model driver never produces (spotlight/losses.py): a `mask` over the b
// pairs (loss*mask summed, divided by mask.sum(), losses.py:51-55,91-95,124-128) and 2-D negatives [n_rows, b]: adaptive
// hinge takes the per-positive maximum over dim 0 (first maximal row, losses.py:170), hinge / bpr broadcast.  neg is row-major [n_rows, b];
// n_rows == 0 means 1-D negatives of length m (adaptive hinge: the global maximum, as k_loss).  One CTA.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LOSS_THREADS) k_loss_ex(int kind, const float *__restrict__ pos, int b,
                                                          const float *__restrict__ neg, int n_rows, int m,
                                                          const float *__restrict__ mask, float *__restrict__ loss_out,
                                                          float *__restrict__ dpos, float *__restrict__ dneg) {
  __shared__ double sh[LOSS_THREADS / 32];
  __shared__ float sh_max[LOSS_THREADS / 32];
  __shared__ int sh_arg[LOSS_THREADS / 32];
  const int tid = threadIdx.x;
  const bool grad = dpos != nullptr;
  double wsum = 0.0;
  for (int j = tid; j < b; j += LOSS_THREADS) wsum += mask ? (double)mask[j] : 1.0;
  wsum = block_sum(wsum, sh);
  const float W = (float)wsum;                        // mask.sum() (or b): fp32 like the reference's tensor
  if (kind == MFB_LOSS_POINTWISE) {
    // scalar BCE sum, then (loss*mask).sum()/mask.sum(): the mask cancels unless it is all zero (-> nan, as torch)
    double sp = 0.0, sn = 0.0;
    for (int j = tid; j < b; j += LOSS_THREADS) sp += (double)(-fmaxf(logf(pos[j]), -100.0f));
    for (int j = tid; j < m; j += LOSS_THREADS) sn += (double)(-fmaxf(logf(1.0f - neg[j]), -100.0f));
    sp = block_sum(sp, sh);
    sn = block_sum(sn, sh);
    float l = (float)(sp / (double)b);
    if (m > 0) l = l + (float)(sn / (double)m);
    const float scale = mask ? __fdiv_rn(W, W) : 1.0f;            // d/dl of (l*mask).sum()/mask.sum()
    if (tid == 0) *loss_out = mask ? __fdiv_rn(l * W, W) : l;
    if (grad) {
      for (int j = tid; j < b; j += LOSS_THREADS) {
        const float x = pos[j];
        dpos[j] = scale * __fdiv_rn(__fdiv_rn(x - 1.0f, fmaxf((1.0f - x) * x, 1e-12f)), (float)b);
      }
      for (int j = tid; j < m; j += LOSS_THREADS) {
        const float x = neg[j];
        dneg[j] = scale * __fdiv_rn(__fdiv_rn(x, fmaxf((1.0f - x) * x, 1e-12f)), (float)m);
      }
    }
    return;
  }
  // pairwise losses over the b positives; the negative of pair j is neg[j] (1-D), the column maximum (2-D) or the
  // global maximum (adaptive hinge on 1-D negatives)
  float gbest = -INFINITY;
  int garg = 0x7fffffff;
  const bool adaptive = kind == MFB_LOSS_ADAPTIVE_HINGE;
  if (adaptive && n_rows == 0) {
    for (int j = tid; j < m; j += LOSS_THREADS) {
      const float x = neg[j];
      if (x > gbest) { gbest = x; garg = j; }
    }
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, gbest, o);
      const int oa = __shfl_xor_sync(0xffffffffu, garg, o);
      if (ob > gbest || (ob == gbest && oa < garg)) { gbest = ob; garg = oa; }
    }
    if (lane == 0) { sh_max[wid] = gbest; sh_arg[wid] = garg; }
    __syncthreads();
    gbest = sh_max[0];
    garg = sh_arg[0];
    for (int w = 1; w < LOSS_THREADS / 32; ++w)
      if (sh_max[w] > gbest || (sh_max[w] == gbest && sh_arg[w] < garg)) { gbest = sh_max[w]; garg = sh_arg[w]; }
    if (grad)
      for (int j = tid; j < m; j += LOSS_THREADS) dneg[j] = 0.0f;
    __syncthreads();
  } else if (adaptive && grad) {
    for (long long j = tid; j < (long long)n_rows * b; j += LOSS_THREADS) dneg[j] = 0.0f;
    __syncthreads();
  }
  if (!adaptive && n_rows > 0) {
    // hinge / bpr broadcast over the rows of [n_rows, b] negatives: loss[r, j] against pos[j]; the reference takes
    // loss.mean() over n_rows*b entries, or with a mask (loss*mask[b]).sum() / mask.sum() (the mask's own sum)
    const double denom = mask ? wsum : (double)n_rows * (double)b;
    const float Wn = (float)denom;
    double s2 = 0.0;
    for (int j = tid; j < b; j += LOSS_THREADS) {
      const float xp = pos[j], w = mask ? mask[j] : 1.0f;
      const float scale = __fdiv_rn(w, Wn);
      float dp_acc = 0.f;
      for (int r = 0; r < n_rows; ++r) {
        const float xn = neg[(long long)r * b + j];
        float dp, dn;
        if (kind == MFB_LOSS_BPR) {
          const float sg = sigmoidf_acc(xp - xn);
          s2 += (double)((1.0f - sg) * w);
          const float g = ((-scale) * (1.0f - sg)) * sg;
          dp = g;
          dn = -g;
        } else {
          const float d = (xn - xp) + 1.0f;
          s2 += (double)(fmaxf(d, 0.0f) * w);
          const float a = (d >= 0.0f) ? scale : 0.0f;
          dp = -a;
          dn = a;
        }
        dp_acc += dp;
        if (grad) dneg[(long long)r * b + j] = dn;
      }
      if (grad) dpos[j] = dp_acc;
    }
    s2 = block_sum(s2, sh);
    if (tid == 0) *loss_out = (float)(s2 / denom);
    return;
  }
  double s = 0.0, gsum = 0.0;
  for (int j = tid; j < b; j += LOSS_THREADS) {
    const float xp = pos[j], w = mask ? mask[j] : 1.0f;
    float xn;
    int arg = j;
    if (!adaptive) {
      xn = neg[j];
    } else if (n_rows == 0) {
      xn = gbest;
    } else {
      xn = neg[j];
      arg = j;
      for (int r = 1; r < n_rows; ++r) {
        const float x = neg[(long long)r * b + j];
        if (x > xn) { xn = x; arg = r * b + j; }
      }
    }
    float dp, dn;
    if (kind == MFB_LOSS_BPR) {
      const float sg = sigmoidf_acc(xp - xn);
      s += (double)((1.0f - sg) * w);
      const float g = ((-__fdiv_rn(w, W)) * (1.0f - sg)) * sg;
      dp = g;
      dn = -g;
    } else {
      const float d = (xn - xp) + 1.0f;
      s += (double)(fmaxf(d, 0.0f) * w);
      const float a = (d >= 0.0f) ? __fdiv_rn(w, W) : 0.0f;
      dp = -a;
      dn = a;
    }
    if (grad) {
      dpos[j] = dp;
      if (adaptive && n_rows == 0) gsum += (double)dn;     // every pair feeds the one global maximum
      else dneg[arg] = dn;
    }
  }
  s = block_sum(s, sh);
  gsum = block_sum(gsum, sh);
  if (tid == 0) {
    *loss_out = (float)(s / wsum);
    if (grad && adaptive && n_rows == 0) dneg[garg] = (float)gsum;
  }
}

// Loss values of a whole chunk of steps in one launch (block s <-> step step0+s): the per-step
// predictions are kept until the chunk ends, so loss.item() costs nothing on the step's critical path.
__global__ void __launch_bounds__(LOSS_THREADS) k_loss_steps(int kind, const float *__restrict__ pred_chunk,
                                                             int Lfull, int batch, int m_neg, long long n_pos,
                                                             long long step0, float *__restrict__ losses) {
  const long long gstep = step0 + blockIdx.x;
  const long long first = gstep * batch;
  const int b = (int)((n_pos - first < batch) ? (n_pos - first) : batch);
  const float *pred = pred_chunk + (long long)blockIdx.x * Lfull;
  loss_block(kind, pred, b, pred + b, m_neg, losses + gstep, nullptr, nullptr, 0);
}

// dLoss/dz of slot j, recomputed where it is consumed (k_update) from the step's predictions and the
// adaptive-hinge maximum; same formulas as loss_block (torch backward of spotlight/losses.py).
// Hinge-type terms on probabilities are always active: neg - pos + 1 >= 0 for values in [0, 1].
template <int KIND>
__device__ __forceinline__ float slot_dz(int j, int b, int m, const float *pred, float gmax, int jstar) {
  const float x = pred[j];
  float d;
  if (KIND == MFB_LOSS_POINTWISE) {
    const float den = fmaxf((1.0f - x) * x, 1e-12f);
    d = (j < b) ? __fdiv_rn(__fdiv_rn(x - 1.0f, den), (float)b) : __fdiv_rn(__fdiv_rn(x, den), (float)m);
  } else if (KIND == MFB_LOSS_HINGE) {
    const float inv_b = 1.0f / (float)b;
    const float xp = (j < b) ? x : pred[j - b], xn = (j < b) ? pred[j + b] : x;
    const float a = (((xn - xp) + 1.0f) >= 0.0f) ? inv_b : 0.0f;
    d = (j < b) ? -a : a;
  } else if (KIND == MFB_LOSS_BPR) {
    const float inv_b = 1.0f / (float)b;
    const float xp = (j < b) ? x : pred[j - b], xn = (j < b) ? pred[j + b] : x;
    const float sg = sigmoidf_acc(xp - xn);
    const float g = ((-inv_b) * (1.0f - sg)) * sg;
    d = (j < b) ? g : -g;
  } else {
    const float inv_b = 1.0f / (float)b;
    if (j < b) {
      d = (((gmax - x) + 1.0f) >= 0.0f) ? -inv_b : 0.0f;
    } else {
      d = (j - b == jstar) ? (float)b * inv_b : 0.0f;  // sum over the b active positives of 1/b
    }
  }
  return (d * (1.0f - x)) * x;  // sigmoid backward: grad * (1 - y) * y
}

// ---------------------------------------------------------------------------------------
// planner kernels (model-independent integer work, bulk per chunk)
// ---------------------------------------------------------------------------------------
// Keys for the second, row-major sort: (table,row) only, payload = step-major position.
__global__ void k_rowkeys(const uint32_t *__restrict__ skeys, long long n, int rb, uint32_t *__restrict__ rkeys,
                          uint32_t *__restrict__ rvals) {
  long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  rkeys[q] = skeys[q] & ((2u << rb) - 1u);  // table bit + row bits
  rvals[q] = (uint32_t)q;
}

// The stable (table,row) sort lists every row's uses in step order.  For the head of each use:
//   seg_gap[head] = steps until the row's next use inside the chunk (0: none)  -> eager replay
//   and the use is appended to its step's lazy list when nothing will have replayed the row up to it:
//   first use in the chunk, or the previous use is more than eager_max steps back.
// List order is arbitrary (atomic append) but every entry is an independent row, so results do not depend on it.
__global__ void k_next_use(const uint32_t *__restrict__ rkeys, const uint32_t *__restrict__ rvals, long long n, int rb,
                           const uint32_t *__restrict__ skeys, const uint32_t *__restrict__ seg_first,
                           const uint32_t *__restrict__ seg_len, uint32_t *__restrict__ seg_gap, int eager_max,
                           uint32_t *__restrict__ lazy_rows, int *__restrict__ lazy_cnt, int lazy_cap) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t q = rvals[i];
  if (seg_first[q] != q) return;                        // only heads
  const uint32_t rk = rkeys[i];
  const int step = (int)(skeys[q] >> (rb + 1));
  const long long i2 = i + seg_len[q];                  // the segment's positions are adjacent here too
  uint32_t gap = 0;
  if (i2 < n && rkeys[i2] == rk) gap = (skeys[rvals[i2]] >> (rb + 1)) - (uint32_t)step;
  seg_gap[q] = gap;
  bool lazy = true;
  if (i > 0 && rkeys[i - 1] == rk) {                    // previous use: its head decides whether it replays to us
    const uint32_t pq = seg_first[rvals[i - 1]];
    const int pgap = step - (int)(skeys[pq] >> (rb + 1));
    lazy = pgap > 1 && pgap > eager_max;                // pgap == 1: the previous step's own update covers it
  }
  if (lazy) {
    const int at = atomicAdd(lazy_cnt + step, 1);
    lazy_rows[(long long)step * lazy_cap + at] = rk;
  }
}

// ---------------------------------------------------------------------------------------
// k_update: one warp per sorted position of the step (+ catch-up blocks, see below).
// A row's slot gradients form one segment
//   d/dU[u] += dz_j * V_old[i_j]   d/dbu[u] += dz_j   (and symmetrically for items)
// which is reduced in slot order.  Segments longer than a window (popular items) are cut at
// absolute window boundaries (positions that are multiples of UPD_WIN): each piece is summed by
// its own warp into a partial, and the last piece to finish (atomic ticket) adds the partials in
// position order and applies optimiser step t.  Piece boundaries depend only on the sorted ids,
// so the summation order -- and the result -- is identical from run to run.
//
// Blocks [0, cu_blocks) do not update: they walk the lazy list of the NEXT step and replay those
// rows (disjoint from this step's rows) up to optimiser step t, so the next forward finds every
// row current.  With n == 0 the kernel is a pure catch-up pass (first step of a chunk).
// ---------------------------------------------------------------------------------------
constexpr int UPD_WIN = 32;
// Measured on B200 (cfg3): 4-warp blocks capped at ~51 registers beat 8-warp/64-register blocks by ~12%
// (a block retires only when its slowest warp -- the longest eager replay -- is done).
#ifndef MFB_UPD_WARPS
#define MFB_UPD_WARPS 4
#endif
#ifndef MFB_UPD_MINB
#define MFB_UPD_MINB 10
#endif
#ifndef MFB_UPD_INFLIGHT
#define MFB_UPD_INFLIGHT 2   // measured: 2 slots in flight (48 regs, 16 B spill) beats 4 (96 B spill) by 8%
#endif
#ifndef MFB_UPD_INTERLEAVE
#define MFB_UPD_INTERLEAVE 1
#endif
constexpr int UPD_WARPS = MFB_UPD_WARPS;   // warps per k_update block

struct UpdArgs {
  const PosInfo *info;        // chunk-global
  const uint32_t *svals;      // chunk-global: slot id per sorted position
  long long base;             // first sorted position of this step
  int n;                      // sorted positions of this step (2 * slots)
  int rb, D;
  TableView users, items;
  OptView opt;
  const float *snap_u, *snap_i, *pred;
  int b, m_neg;
  const unsigned long long *gmax_cell;
  float *partial;
  int pstride;
  int *tickets;
  int t;                      // optimiser step applied by this launch
  int eager_max;
  // lazy catch-up role
  int cu_blocks;
  const uint32_t *lazy_rows;
  const int *lazy_cnt;
  int cu_target;              // optimiser step the listed rows are brought to
};

// One row of the lazy list: replay its zero-gradient steps up to cu_target.
template <int VEC, int NIT, bool FAST>
__device__ __forceinline__ void catchup_row(const UpdArgs &a, uint32_t rk, int lane) {
  const int D = a.D;
  const int adam = opt_state_bits(a.opt.kind);
  const long long row = rk & ((1u << a.rb) - 1u);
  const TableView &T = ((rk >> a.rb) & 1u) ? a.items : a.users;
  const int last = T.last[row];
  if (last >= a.cu_target) return;
  RowState<VEC, NIT> r;
  row_load<VEC, NIT>(r, T, row, D, lane, adam);
  row_replay<VEC, NIT, FAST>(r, last, a.cu_target, a.opt);
  row_store<VEC, NIT>(r, T, row, D, lane, adam);
  if (lane == 0) T.last[row] = a.cu_target;
}

// The update role for sorted position ql of the step (one warp): `raw` = its PosInfo record, j0 = its slot.
template <int VEC, int NIT, bool FAST, int KIND>
__device__ __forceinline__ void update_position(const UpdArgs &a, int ql, const uint4 raw, int j0, int lane) {
  const int D = a.D;
  const int adam = opt_state_bits(a.opt.kind);   // which per-row optimiser state exists
  const long long q = a.base + ql;
  const uint32_t key = raw.x;
  const long long first = raw.y;
  const bool head = first == q;
  if (!head && (ql % UPD_WIN) != 0) return;
  const int fl = (int)(first - a.base);
  const int seg_end = fl + (int)raw.z;                                 // local, exclusive
  const int win_end = (ql / UPD_WIN + 1) * UPD_WIN;
  const int run_end = seg_end < win_end ? seg_end : win_end;
  const bool whole = head && run_end == seg_end;
  const long long row = key & ((1u << a.rb) - 1u);
  const bool is_item = (key >> a.rb) & 1u;
  const TableView &T = is_item ? a.items : a.users;
  const float *__restrict__ other = is_item ? a.snap_u : a.snap_i;
  const float *__restrict__ pred = a.pred;
  const int b = a.b, m_neg = a.m_neg;

  // the row's optimiser state does not depend on the gradient: get it in flight first.  (The adaptive-hinge cell is
  // one address for the whole grid; the compiler moves it to a uniform register and the warp waits for it there --
  // ncu put 12 % of the stall samples on that wait when it preceded the row loads.)
  RowState<VEC, NIT> r;
  if (whole) row_load<VEC, NIT>(r, T, row, D, lane, adam);
  unsigned long long gcell = 0ull;
  if (KIND == MFB_LOSS_ADAPTIVE_HINGE) gcell = __ldcg(a.gmax_cell + (lane & 0));   // per-lane load: no uniform-register wait
  const float gmax = unpack_max_val(gcell);
  const int jstar = (KIND == MFB_LOSS_ADAPTIVE_HINGE) ? unpack_max_idx(gcell) : -1;

  Frag<VEC, NIT> g;
#pragma unroll
  for (int k = 0; k < NIT * VEC; ++k) g.x[k] = 0.f;
  float gb = 0.f;
  {  // first slot of the run (most rows have exactly one)
    // adaptive hinge: only the first maximal negative carries gradient (known without its prediction)
    const bool use = !(KIND == MFB_LOSS_ADAPTIVE_HINGE && j0 >= b && j0 - b != jstar);
    if (use) {
      Frag<VEC, NIT> o;
      frag_load<VEC, NIT>(o, other + (long long)j0 * D, D, lane);
      const float d = slot_dz<KIND>(j0, b, m_neg, pred, gmax, jstar);
#pragma unroll
      for (int k = 0; k < NIT * VEC; ++k) g.x[k] = __fmul_rn(d, o.x[k]);   // 0 + d*o
      gb = d;
    }
  }
  for (int p0 = ql + 1; p0 < run_end; p0 += MFB_UPD_INFLIGHT) {  // remaining slots, several in flight, added in slot order
    int jj[MFB_UPD_INFLIGHT];
    bool use[MFB_UPD_INFLIGHT];
    float dd[MFB_UPD_INFLIGHT];
    Frag<VEC, NIT> o[MFB_UPD_INFLIGHT];
#pragma unroll
    for (int u = 0; u < MFB_UPD_INFLIGHT; ++u) {
      jj[u] = (p0 + u < run_end) ? (int)a.svals[a.base + p0 + u] : 0;
      use[u] = (p0 + u < run_end) && !(KIND == MFB_LOSS_ADAPTIVE_HINGE && jj[u] >= b && jj[u] - b != jstar);
    }
#pragma unroll
    for (int u = 0; u < MFB_UPD_INFLIGHT; ++u) {
      dd[u] = 0.f;
      if (use[u]) {
        frag_load<VEC, NIT>(o[u], other + (long long)jj[u] * D, D, lane);
        dd[u] = slot_dz<KIND>(jj[u], b, m_neg, pred, gmax, jstar);
      }
    }
#pragma unroll
    for (int u = 0; u < MFB_UPD_INFLIGHT; ++u) {
      if (use[u]) {
#pragma unroll
        for (int k = 0; k < NIT * VEC; ++k) g.x[k] = __fadd_rn(g.x[k], __fmul_rn(dd[u], o[u].x[k]));
        gb = __fadd_rn(gb, dd[u]);
      }
    }
  }

  if (!whole) {
    // piece of a long segment: publish the partial, take a ticket
    const int first_win = fl / UPD_WIN, last_win = (seg_end - 1) / UPD_WIN;
    const int npieces = last_win - first_win + 1;
    const int nwin = (a.n + UPD_WIN - 1) / UPD_WIN;
    // partial slots: [0, nwin) pieces that start at a window boundary, [nwin, 2*nwin) pieces that start at a head
    const int pslot = head ? (nwin + ql / UPD_WIN) : (ql / UPD_WIN);
    float *pp = a.partial + (long long)pslot * a.pstride;
    frag_store<VEC, NIT>(g, pp, D, lane);
    if (lane == 0) pp[D] = gb;
    __threadfence();
    __syncwarp();
    int old = 0;
    if (lane == 0) old = atomicAdd(a.tickets + fl, 1);
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old != npieces - 1) return;
    if (lane == 0) a.tickets[fl] = 0;  // self-cleaning for the next step
    __threadfence();
    // ordered sum: the head piece, then the window-aligned pieces in position order
    const float *hp = a.partial + (long long)(nwin + first_win) * a.pstride;
    frag_load_cg<VEC, NIT>(g, hp, D, lane);
    gb = __ldcg(hp + D);
    for (int w0 = first_win + 1; w0 <= last_win; w0 += 4) {   // four partial rows in flight, added in position order
      Frag<VEC, NIT> o[4];
      float ob[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (w0 + u <= last_win) {
          const float *wp = a.partial + (long long)(w0 + u) * a.pstride;
          frag_load_cg<VEC, NIT>(o[u], wp, D, lane);
          ob[u] = __ldcg(wp + D);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (w0 + u <= last_win) {
#pragma unroll
          for (int k = 0; k < NIT * VEC; ++k) g.x[k] = __fadd_rn(g.x[k], o[u].x[k]);
          gb = __fadd_rn(gb, ob[u]);
        }
      }
    }
    row_load<VEC, NIT>(r, T, row, D, lane, adam);
  }
  apply_step<VEC, NIT, FAST>(r, g, gb, a.opt, a.t);
  // eager replay: the row is next used `gap` steps from now, so the zero-gradient dense updates of the
  // steps in between are applied right here (no extra pass over the row later)
  int upto = a.t;
  const int gap = (int)raw.w;
  if (gap > 1 && gap <= a.eager_max) {   // (every position of a segment carries the segment's gap)
    upto = a.t + gap - 1;
    row_replay<VEC, NIT, FAST>(r, a.t, upto, a.opt);
  }
  row_store<VEC, NIT>(r, T, row, D, lane, adam);
  if (lane == 0) T.last[row] = upto;
}

template <int VEC, int NIT, bool FAST, int KIND>
__global__ void __launch_bounds__(UPD_WARPS * 32, MFB_UPD_MINB) k_update(const UpdArgs a) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

  pdl_launch_dependents();
  if ((int)blockIdx.x < a.cu_blocks) {
    // ---- lazy catch-up role ------------------------------------------------------------
    const int cnt = *a.lazy_cnt;
    pdl_wait();
    const int stride = a.cu_blocks * UPD_WARPS;
    for (int i = blockIdx.x * UPD_WARPS + wid; i < cnt; i += stride) catchup_row<VEC, NIT, FAST>(a, a.lazy_rows[i], lane);
    return;
  }

  // ---- update role -----------------------------------------------------------------------
  // Position within the step.  The sorted positions of a step are [user rows | item rows] (n/2 each).  User rows carry
  // long eager replays (MUFU-bound), item rows are touched almost every step (load/store-bound): even blocks walk the
  // user half, odd blocks the item half, so both kinds are resident on every SM at the same time instead of one
  // after the other.
  int ql;
  if (MFB_UPD_INTERLEAVE) {
    const int ub = (int)blockIdx.x - a.cu_blocks;
    const int half = a.n >> 1;
    const int in_half = (ub >> 1) * UPD_WARPS + wid;
    if (in_half >= half) return;
    ql = (ub & 1) * half + in_half;
  } else {
    ql = ((int)blockIdx.x - a.cu_blocks) * UPD_WARPS + wid;
    if (ql >= a.n) return;
  }
  const long long q = a.base + ql;
  // independent loads first: this kernel is bound by dependent-load latency, not bandwidth
  const uint4 raw = *reinterpret_cast<const uint4 *>(a.info + q);
  const int j0 = (int)a.svals[q];
  pdl_wait();   // predictions, snapshots and the adaptive-hinge maximum come from this step's forward kernel
  update_position<VEC, NIT, FAST, KIND>(a, ql, raw, j0, lane);
}

// ---------------------------------------------------------------------------------------
// k_flush: one warp per row of one table; replay (last, t]
template <int VEC, int NIT, bool FAST>
__global__ void __launch_bounds__(BLOCK_THREADS) k_flush(TableView T, OptView opt, int D, int t) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= T.rows) return;
  const int last = T.last[row];
  if (last >= t) return;
  const int adam = opt_state_bits(opt.kind);
  RowState<VEC, NIT> r;
  row_load<VEC, NIT>(r, T, row, D, lane, adam);
  row_replay<VEC, NIT, FAST>(r, last, t, opt);
  row_store<VEC, NIT>(r, T, row, D, lane, adam);
  if (lane == 0) T.last[row] = t;
}

// ---------------------------------------------------------------------------------------
// dispatch on (vector width, iterations per lane, fast math, loss kind)
// ---------------------------------------------------------------------------------------
inline int grid_for_warps(long long warps) { return (int)((warps + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK); }

// launch with programmatic stream serialization (see pdl_wait)
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), int grid, int block, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
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
  int tk = m->prof.begin(PK_FLUSH, st, 2);
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
  m->prof.end(tk, st);
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
  mfb_count_library_launch(1);
  k_loss<<<1, LOSS_THREADS, 0, (cudaStream_t)stream>>>(loss, d_pos, (int)n_pos, d_neg, (int)n_neg, d_loss, d_dpos,
                                                       d_dneg, 0);
  MFB_KERNEL_CHECK();
  return MFB_OK;
}

extern "C" int mfb_loss_forward_backward_ex(int loss, const float *d_pos, int64_t n_pos, const float *d_neg,
                                            int64_t neg_rows, int64_t neg_cols, const float *d_mask, float *d_loss,
                                            float *d_dpos, float *d_dneg, mfb_stream stream) {
  if (!d_pos || !d_loss || n_pos <= 0 || neg_rows < 0 || neg_cols < 0 || (neg_cols > 0 && !d_neg)) return MFB_ERR_INVALID;
  if (loss < MFB_LOSS_POINTWISE || loss > MFB_LOSS_ADAPTIVE_HINGE) return MFB_ERR_INVALID;
  if (neg_cols == 0 && loss != MFB_LOSS_POINTWISE) return MFB_ERR_SHAPE;
  if (neg_rows > 0 && loss == MFB_LOSS_POINTWISE) {
    mfb_set_error("2-D negatives: pointwise_loss flattens its inputs, pass them 1-D");
    return MFB_ERR_UNSUPPORTED;
  }
  const bool pairwise = loss == MFB_LOSS_HINGE || loss == MFB_LOSS_BPR || neg_rows > 0;
  if (pairwise && neg_cols != n_pos) {
    mfb_set_error("pairwise loss: negatives have %lld columns, positives %lld", (long long)neg_cols, (long long)n_pos);
    return MFB_ERR_SHAPE;
  }
  if ((d_dpos == nullptr) != (d_dneg == nullptr) && neg_cols > 0) return MFB_ERR_INVALID;
  mfb_count_library_launch(1);
  k_loss_ex<<<1, LOSS_THREADS, 0, (cudaStream_t)stream>>>(loss, d_pos, (int)n_pos, d_neg, (int)neg_rows, (int)neg_cols,
                                                          d_mask, d_loss, d_dpos, d_dneg);
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
  int tk = m->prof.begin(PK_PREDICT, st, 2);
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
  k_forward<V, N><<<grid_for_warps((L + FWD_SPW - 1) / FWD_SPW), BLOCK_THREADS, 0, st>>>(slot_u + off, slot_i + off, L, m->users, m->items, D, \
                                                                nullptr, nullptr, d_out + off, -1, nullptr)
    MFB_DISPATCH_SHAPE(sh, CALL);
#undef CALL
    MFB_KERNEL_CHECK();
  }
  m->prof.end(tk, st);
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
  int tk = m->prof.begin(PK_PREDICT, st);
#define CALL(V, N) \
  k_forward_user<V, N><<<grid_for_warps(I), BLOCK_THREADS, 0, st>>>((long long)user, I, m->users, m->items, D, d_out)
  MFB_DISPATCH_SHAPE(sh, CALL);
#undef CALL
  m->prof.end(tk, st);
  MFB_KERNEL_CHECK();
  return MFB_OK;
}

// ---------------------------------------------------------------------------------------
// host pipeline
// ---------------------------------------------------------------------------------------
struct StepGeom {
  int D, batch, m_neg, Lfull, rb, chunk, loss;
  int64_t n_pos, nsteps;
  bool train, fast, adaptive;
  int eager_max = 0;   // longest next-use distance replayed eagerly inside k_update
  int cu_blocks = 0;   // catch-up blocks prepended to every k_update launch
  // negatives: explicit per-epoch arrays, or drawn per chunk from the model's MT19937 stream
  const int64_t *pop_users = nullptr, *pop_items = nullptr;
  int64_t pop_len = 0;
};

static int ensure_streams(mfb_model *m) {
  if (m->st_plan) return MFB_OK;
  // The planner's and the generator's kernels are small and run while the step kernels fill every SM: highest stream
  // priority lets their blocks take the next free slots instead of queueing behind a whole k_update grid.
  int prio_lo = 0, prio_hi = 0;
  MFB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  MFB_CUDA(cudaStreamCreateWithPriority(&m->st_plan, cudaStreamNonBlocking, prio_hi));
  MFB_CUDA(cudaStreamCreateWithPriority(&m->st_rng, cudaStreamNonBlocking, prio_hi));
  auto mk = [](cudaEvent_t *e) { return cudaEventCreateWithFlags(e, cudaEventDisableTiming); };
  for (auto &e : m->ev_plan) MFB_CUDA(mk(&e));
  for (auto &e : m->ev_done) MFB_CUDA(mk(&e));
  for (auto &e : m->ev_rng) MFB_CUDA(mk(&e));
  MFB_CUDA(mk(&m->ev_join));
  MFB_CUDA(mk(&m->ev_seed));
  int dev = 0;
  MFB_CUDA(cudaGetDevice(&dev));
  MFB_CUDA(cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, dev));
  return MFB_OK;
}

static void chunk_extent(const StepGeom &g, int64_t s0, int cap, int *ns, int *b_last, int64_t *nkeys) {
  *ns = (int)((g.nsteps - s0 < cap) ? (g.nsteps - s0) : cap);
  const int64_t last_first = (s0 + *ns - 1) * (int64_t)g.batch;
  *b_last = (int)((g.n_pos - last_first < g.batch) ? (g.n_pos - last_first) : g.batch);
  // all steps are full except possibly the last of the epoch (which is last in its chunk)
  *nkeys = 2 * ((int64_t)(*ns - 1) * g.Lfull + (*b_last + g.m_neg));
}

// MT19937 words of a chunk's negative draws, on the RNG stream, ahead of the planner.  The generator is one sequential
// stream; with 16 384 or more words per step it is cut into shares by jump-ahead and generated by many CTAs
// (mfb_mt_jump.cu), otherwise by one CTA.
static int draw_chunk_words(mfb_model *m, PlanBuf &pb, const StepGeom &g, int64_t s0, int cap, cudaStream_t sr) {
  if (!(g.pop_len > 0 && g.m_neg > 0)) return MFB_OK;
  int ns, b_last;
  int64_t nkeys;
  chunk_extent(g, s0, cap, &ns, &b_last, &nkeys);
  const int64_t k = (int64_t)ns * g.m_neg;
  MFB_CHECK(pb.words.reserve((size_t)(2 * (int64_t)g.chunk * g.m_neg) * sizeof(uint32_t)));
  int tk = m->prof.begin(PK_SAMPLE, sr, 0);
  MFB_CHECK(mfb_mt_generate_parallel(m->rng_state.as<uint32_t>(), 2 * k, pb.words.as<uint32_t>(), 2 * (int64_t)g.m_neg,
                                     &m->rng_jump, sr));
  m->prof.end(tk, sr);
  return MFB_OK;
}

static int plan_chunk(mfb_model *m, PlanBuf &pb, const StepGeom &g, const int64_t *d_pos_users,
                      const int64_t *d_pos_items, const int64_t *d_neg_users, const int64_t *d_neg_items, int64_t s0,
                      int cap, int *flag, cudaStream_t sp) {
  int ns, b_last;
  int64_t nkeys;
  chunk_extent(g, s0, cap, &ns, &b_last, &nkeys);
  const size_t slots = (size_t)g.chunk * g.Lfull;
  MFB_CHECK(pb.slots.reserve(slots * 2 * sizeof(int)));
  MFB_CHECK(pb.pred.reserve(slots * sizeof(float)));
  MFB_CHECK(pb.gmax.reserve((size_t)g.chunk * sizeof(unsigned long long)));
  MFB_CHECK(pb.keys_a.reserve(slots * 2 * sizeof(uint32_t)));
  MFB_CHECK(pb.vals_a.reserve(slots * 2 * sizeof(uint32_t)));
  if (g.train) {
    MFB_CHECK(pb.keys_b.reserve(slots * 2 * sizeof(uint32_t)));
    MFB_CHECK(pb.vals_b.reserve(slots * 2 * sizeof(uint32_t)));
    MFB_CHECK(pb.keys_c.reserve(slots * 2 * sizeof(uint32_t)));
    MFB_CHECK(pb.vals_c.reserve(slots * 2 * sizeof(uint32_t)));
    MFB_CHECK(pb.seg.reserve(slots * 2 * 3 * sizeof(uint32_t)));
    MFB_CHECK(pb.info.reserve(slots * 2 * sizeof(PosInfo)));
    MFB_CHECK(pb.lazy_rows.reserve(slots * 2 * sizeof(uint32_t)));
    MFB_CHECK(pb.lazy_cnt.reserve((size_t)g.chunk * sizeof(int)));
  }
  int *slot_u = pb.slots.as<int>();
  int *slot_i = slot_u + slots;
  const long long total = (long long)ns * g.Lfull;
  int tk;
  long long neg_step0 = 0;
  if (g.pop_len > 0 && g.m_neg > 0) {
    // random.choices(neg_examples, k = n_neg*batch) for each step of the chunk, in step order
    // (the words were drawn by draw_chunk_words on the RNG stream; the caller made this stream wait for them)
    const int64_t k = (int64_t)ns * g.m_neg;
    MFB_CHECK(pb.neg_u.reserve((size_t)((int64_t)g.chunk * g.m_neg) * sizeof(int64_t)));
    MFB_CHECK(pb.neg_i.reserve((size_t)((int64_t)g.chunk * g.m_neg) * sizeof(int64_t)));
    tk = m->prof.begin(PK_SAMPLE, sp, 0);
    MFB_CHECK(mfb_choices_async(pb.words.as<uint32_t>(), k, g.pop_len, g.pop_users, g.pop_items,
                                pb.neg_u.as<int64_t>(), pb.neg_i.as<int64_t>(), sp));
    m->prof.end(tk, sp);
    d_neg_users = pb.neg_u.as<int64_t>();
    d_neg_items = pb.neg_i.as<int64_t>();
    neg_step0 = s0;
  }
  tk = m->prof.begin(PK_PACK, sp);
  k_pack<<<(unsigned)((total + 255) / 256), 256, 0, sp>>>(
      (const long long *)d_pos_users, (const long long *)d_pos_items, g.n_pos, (const long long *)d_neg_users,
      (const long long *)d_neg_items, g.batch, g.m_neg, s0, neg_step0, ns, g.rb, m->users.rows, m->items.rows,
      slot_u, slot_i, pb.keys_a.as<uint32_t>(), pb.vals_a.as<uint32_t>(), flag);
  m->prof.end(tk, sp);
  MFB_KERNEL_CHECK();
  if (g.adaptive) MFB_CUDA(cudaMemsetAsync(pb.gmax.ptr, 0, (size_t)ns * sizeof(unsigned long long), sp));
  if (g.train) {
    const unsigned gk = (unsigned)((nkeys + 255) / 256);
    const int step_bits = ns > 1 ? bits_for((uint32_t)(ns - 1)) : 0;
    const int npass = (g.rb + 1 + step_bits + 7) / 8 + (g.rb + 1 + 7) / 8;
    tk = m->prof.begin(PK_SORT, sp, 3 * npass + 4);
    MFB_CHECK(mfb_radix_sort_pairs(pb.keys_a.as<uint32_t>(), pb.vals_a.as<uint32_t>(), pb.keys_b.as<uint32_t>(),
                                   pb.vals_b.as<uint32_t>(), nkeys, g.rb + 1 + step_bits, m->ws_hist, &pb.skeys,
                                   &pb.svals, sp));
    uint32_t *seg_first = pb.seg.as<uint32_t>();
    uint32_t *seg_len = seg_first + slots * 2;
    uint32_t *seg_gap = seg_len + slots * 2;
    k_segments<<<gk, 256, 0, sp>>>(pb.skeys, nkeys, seg_first, seg_len);
    MFB_KERNEL_CHECK();
    // re-sort by (table,row): the free ping-pong pair holds the input, keys_c/vals_c the other side
    uint32_t *fk = (pb.skeys == pb.keys_a.as<uint32_t>()) ? pb.keys_b.as<uint32_t>() : pb.keys_a.as<uint32_t>();
    uint32_t *fv = (pb.svals == pb.vals_a.as<uint32_t>()) ? pb.vals_b.as<uint32_t>() : pb.vals_a.as<uint32_t>();
    k_rowkeys<<<gk, 256, 0, sp>>>(pb.skeys, nkeys, g.rb, fk, fv);
    MFB_KERNEL_CHECK();
    uint32_t *rk = nullptr, *rv = nullptr;
    MFB_CHECK(mfb_radix_sort_pairs(fk, fv, pb.keys_c.as<uint32_t>(), pb.vals_c.as<uint32_t>(), nkeys, g.rb + 1,
                                   m->ws_hist, &rk, &rv, sp));
    MFB_CUDA(cudaMemsetAsync(pb.lazy_cnt.ptr, 0, (size_t)ns * sizeof(int), sp));
    k_next_use<<<gk, 256, 0, sp>>>(rk, rv, nkeys, g.rb, pb.skeys, seg_first, seg_len, seg_gap, g.eager_max,
                                   pb.lazy_rows.as<uint32_t>(), pb.lazy_cnt.as<int>(), 2 * g.Lfull);
    MFB_KERNEL_CHECK();
    k_posinfo<<<gk, 256, 0, sp>>>(pb.skeys, nkeys, seg_first, seg_len, seg_gap, pb.info.as<PosInfo>());
    m->prof.end(tk, sp);
    MFB_KERNEL_CHECK();
  }
  return MFB_OK;
}

template <int V, int N, bool FAST>
static void launch_update_kind(const UpdArgs &a, int loss, int grid, cudaStream_t st) {
  switch (loss) {
    case MFB_LOSS_POINTWISE:
      launch_pdl(k_update<V, N, FAST, MFB_LOSS_POINTWISE>, grid, UPD_WARPS * 32, st, a);
      break;
    case MFB_LOSS_BPR:
      launch_pdl(k_update<V, N, FAST, MFB_LOSS_BPR>, grid, UPD_WARPS * 32, st, a);
      break;
    case MFB_LOSS_HINGE:
      launch_pdl(k_update<V, N, FAST, MFB_LOSS_HINGE>, grid, UPD_WARPS * 32, st, a);
      break;
    default:
      launch_pdl(k_update<V, N, FAST, MFB_LOSS_ADAPTIVE_HINGE>, grid, UPD_WARPS * 32, st, a);
      break;
  }
}

static int launch_update(mfb_model *m, const Shape &sh, const StepGeom &g, const UpdArgs &a, int cls,
                         cudaStream_t st) {
  const int grid = a.cu_blocks + (MFB_UPD_INTERLEAVE ? 2 * ((a.n / 2 + UPD_WARPS - 1) / UPD_WARPS)
                                                     : (a.n + UPD_WARPS - 1) / UPD_WARPS);
  if (grid == 0) return MFB_OK;
  int tk = m->prof.begin(cls, st);
#define CALL(V, N)                                        \
  if (g.fast)                                             \
    launch_update_kind<V, N, true>(a, g.loss, grid, st);  \
  else                                                    \
    launch_update_kind<V, N, false>(a, g.loss, grid, st)
  MFB_DISPATCH_SHAPE(sh, CALL);
#undef CALL
  m->prof.end(tk, st);
  MFB_KERNEL_CHECK();
  return MFB_OK;
}

static int exec_chunk(mfb_model *m, PlanBuf &pb, const Shape &sh, const StepGeom &g, int64_t s0, int cap,
                      float *d_step_losses, cudaStream_t st) {
  int ns, b_last;
  int64_t nkeys;
  chunk_extent(g, s0, cap, &ns, &b_last, &nkeys);
  const int D = g.D, Lfull = g.Lfull, m_neg = g.m_neg;
  const size_t slots = (size_t)g.chunk * Lfull;
  const int *slot_u = pb.slots.as<int>();
  const int *slot_i = slot_u + slots;
  float *pred_chunk = pb.pred.as<float>();
  unsigned long long *gmax = pb.gmax.as<unsigned long long>();
  float *snap_u = g.train ? m->ws_rows.as<float>() : nullptr;
  float *snap_i = g.train ? snap_u + (size_t)Lfull * D : nullptr;
  int tk;

  UpdArgs a;
  if (g.train) {
    a.info = pb.info.as<PosInfo>();
    a.svals = pb.svals;
    a.rb = g.rb;
    a.D = D;
    a.users = m->users;
    a.items = m->items;
    a.opt = m->opt;
    a.snap_u = snap_u;
    a.snap_i = snap_i;
    a.m_neg = m_neg;
    a.partial = m->ws_partial.as<float>();
    a.pstride = ((D + 1 + 31) / 32) * 32;  // partial rows own whole 128-byte lines
    a.tickets = m->ws_tickets.as<int>();
    a.eager_max = g.eager_max;
    // first step of the chunk: nothing has replayed its listed rows yet -> pure catch-up pass to t-1
    a.base = 0;
    a.n = 0;
    a.pred = nullptr;
    a.b = 0;
    a.gmax_cell = nullptr;
    a.t = (int)(m->step + 1);
    a.cu_blocks = 4 * m->num_sms * (WARPS_PER_BLOCK / UPD_WARPS);
    a.lazy_rows = pb.lazy_rows.as<uint32_t>();
    a.lazy_cnt = pb.lazy_cnt.as<int>();
    a.cu_target = (int)m->step;
  }
  if (g.train) MFB_CHECK(launch_update(m, sh, g, a, PK_CATCHUP, st));
  for (int s = 0; s < ns; ++s) {
    const int b = (s == ns - 1) ? b_last : g.batch;
    const int L = b + m_neg;
    const int *su = slot_u + (size_t)s * Lfull;
    const int *si = slot_i + (size_t)s * Lfull;
    float *pred = pred_chunk + (size_t)s * Lfull;
    const int neg_begin = g.adaptive ? b : -1;
    tk = m->prof.begin(PK_FORWARD, st);
#define CALL(V, N)                                                                                             \
  launch_pdl(k_forward<V, N>, grid_for_warps((L + FWD_SPW - 1) / FWD_SPW), BLOCK_THREADS, st, su, si, L, m->users, m->items, D, snap_u, snap_i, \
             pred, neg_begin, gmax + s)
    MFB_DISPATCH_SHAPE(sh, CALL);
#undef CALL
    m->prof.end(tk, st);
    MFB_KERNEL_CHECK();
    if (g.train) {
      const int t = (int)(m->step + 1);
      a.base = 2 * (long long)s * Lfull;
      a.n = 2 * L;
      a.pred = pred;
      a.b = b;
      a.gmax_cell = gmax + s;
      a.t = t;
      // rows first needed by step s+1 (and not replayed by anyone) are brought to step t by extra blocks
      if (s + 1 < ns) {
        a.cu_blocks = g.cu_blocks * (WARPS_PER_BLOCK / UPD_WARPS);
        a.lazy_rows = pb.lazy_rows.as<uint32_t>() + (size_t)(s + 1) * 2 * Lfull;
        a.lazy_cnt = pb.lazy_cnt.as<int>() + (s + 1);
        a.cu_target = t;
      } else {
        a.cu_blocks = 0;
      }
      MFB_CHECK(launch_update(m, sh, g, a, PK_UPDATE, st));
      m->step += 1;
    }
  }
  // loss values of the whole chunk in one launch (the predictions of every step were kept)
  tk = m->prof.begin(PK_LOSS, st);
  k_loss_steps<<<ns, LOSS_THREADS, 0, st>>>(g.loss, pred_chunk, Lfull, g.batch, m_neg, g.n_pos, s0, d_step_losses);
  m->prof.end(tk, st);
  MFB_KERNEL_CHECK();
  return MFB_OK;
}

static int run_steps(mfb_model *m, int loss, const int64_t *d_pos_users, const int64_t *d_pos_items, int64_t n_pos,
                     int32_t batch, int32_t n_neg, const int64_t *d_neg_users, const int64_t *d_neg_items,
                     float *d_step_losses, cudaStream_t st, bool train, const int64_t *d_pop_users = nullptr,
                     const int64_t *d_pop_items = nullptr, int64_t pop_len = 0, cudaEvent_t ev_rng_ready = nullptr,
                     bool defer_check = false) {
  if (!m || !d_pos_users || !d_pos_items || !d_step_losses) return MFB_ERR_INVALID;
  MFB_CHECK(validate_loss_shape(loss, n_pos, batch, n_neg));
  const bool from_stream = pop_len > 0;
  if (n_neg > 0 && !from_stream && (!d_neg_users || !d_neg_items)) return MFB_ERR_INVALID;
  if (n_neg > 0 && from_stream) {
    if (!d_pop_users || !d_pop_items) return MFB_ERR_INVALID;
    if (!m->rng_seeded) {
      mfb_set_error("negative sampler not seeded (mfb_model_rng_seed)");
      return MFB_ERR_INVALID;
    }
  }
  Shape sh;
  MFB_CHECK(pick_shape(m->desc.dim, &sh));
  MFB_CHECK(ensure_streams(m));
  StepGeom g;
  g.D = m->desc.dim;
  const int64_t m_neg64 = (int64_t)n_neg * batch;
  if (m_neg64 + batch > (1 << 24)) {
    mfb_set_error("batch*(1+n_neg) too large");
    return MFB_ERR_UNSUPPORTED;
  }
  g.batch = batch;
  g.m_neg = (int)m_neg64;
  g.Lfull = batch + g.m_neg;
  g.n_pos = n_pos;
  g.nsteps = (n_pos + batch - 1) / batch;
  g.loss = loss;
  g.train = train;
  g.fast = m->desc.fast_math != 0;
  g.adaptive = loss == MFB_LOSS_ADAPTIVE_HINGE;
  g.eager_max = train ? m->tune_eager_max : 0;
  g.cu_blocks = m->tune_cu_blocks_per_sm * m->num_sms;
  if (n_neg > 0 && from_stream) {
    g.pop_users = d_pop_users;
    g.pop_items = d_pop_items;
    g.pop_len = pop_len;
  }
  const uint32_t maxrow = (uint32_t)((m->users.rows > m->items.rows ? m->users.rows : m->items.rows) - 1);
  g.rb = bits_for(maxrow);
  const int sb_max = 31 - g.rb;  // bits left for the chunk-local step
  if (sb_max < 0) {
    mfb_set_error("tables too large for 32-bit planner keys");
    return MFB_ERR_UNSUPPORTED;
  }
  const int chunk_bits = m->tune_chunk_bits;
  g.chunk = 1 << (sb_max > chunk_bits ? chunk_bits : sb_max);
  while (g.chunk > 1 && (int64_t)g.chunk * g.Lfull > (1ll << 25)) g.chunk >>= 1;  // bound workspace
  if (!train) MFB_CHECK(launch_flush(m, st));
  if (train) MFB_CHECK(mfb_ensure_scalars(m, m->step + g.nsteps + 1));

  MFB_CHECK(m->ws_scalars.reserve(64));
  int *flag = m->ws_scalars.as<int>();
  MFB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
  if (train) {
    const int nwin_full = (2 * g.Lfull + UPD_WIN - 1) / UPD_WIN;
    const int pstride = ((g.D + 1 + 31) / 32) * 32;
    MFB_CHECK(m->ws_rows.reserve((size_t)g.Lfull * g.D * 2 * sizeof(float)));
    MFB_CHECK(m->ws_partial.reserve((size_t)2 * nwin_full * pstride * sizeof(float)));
    const size_t before = m->ws_tickets.cap;
    MFB_CHECK(m->ws_tickets.reserve((size_t)2 * g.Lfull * sizeof(int)));
    if (m->ws_tickets.cap != before) MFB_CUDA(cudaMemsetAsync(m->ws_tickets.ptr, 0, m->ws_tickets.cap, st));
  }
  // Chunk schedule: the first chunk's plan sits on the critical path (nothing to overlap it with), so chunks start
  // small and grow -- every later plan then hides behind the execution of the chunk before it (a plan costs a fixed
  // ~30 launches + ~14 us per step, a step executes in ~40 us: sizes may roughly triple from chunk to chunk).
  std::vector<std::pair<int64_t, int>> sched;
  {
    int cap = m->tune_chunk_ramp > 0 ? m->tune_chunk_ramp : g.chunk;
    if (cap > g.chunk) cap = g.chunk;
    int rep = 0;
    for (int64_t s0 = 0; s0 < g.nsteps;) {
      const int ns = (int)((g.nsteps - s0 < cap) ? (g.nsteps - s0) : cap);
      sched.push_back({s0, cap});
      s0 += ns;
      if (++rep >= 2 && cap < g.chunk) {   // two chunks of the first size, then doubling
        cap = cap * 2 > g.chunk ? g.chunk : cap * 2;
      }
    }
  }
  // the planner and RNG streams start after everything already queued on the caller's stream (the ids may have
  // been produced there, the generator state is uploaded there) ...
  cudaStream_t sp = m->st_plan, sr = m->st_rng;
  MFB_CUDA(cudaEventRecord(m->ev_join, st));
  MFB_CUDA(cudaStreamWaitEvent(sp, m->ev_join, 0));
  // (the generator needs only the state, not the ids: a caller that uploads ids after the state names the earlier point)
  MFB_CUDA(cudaStreamWaitEvent(sr, ev_rng_ready ? ev_rng_ready : m->ev_join, 0));
  const bool draws = g.pop_len > 0 && g.m_neg > 0;
  const int64_t nchunks = (int64_t)sched.size();
  auto plan = [&](int64_t c) -> int {
    PlanBuf &pb = m->plan[c & 1];
    if (draws) {
      // plan buffer c&1 was last read by the planner of chunk c-2 (ev_plan) -- its words may be overwritten now
      if (c >= 2) MFB_CUDA(cudaStreamWaitEvent(sr, m->ev_plan[c & 1], 0));
      MFB_CHECK(draw_chunk_words(m, pb, g, sched[c].first, sched[c].second, sr));
      MFB_CUDA(cudaEventRecord(m->ev_rng[c & 1], sr));
    }
    // ... re-uses a plan buffer only after the chunk that trained from it has finished
    if (c >= 2) MFB_CUDA(cudaStreamWaitEvent(sp, m->ev_done[c & 1], 0));
    if (draws) MFB_CUDA(cudaStreamWaitEvent(sp, m->ev_rng[c & 1], 0));
    MFB_CHECK(plan_chunk(m, pb, g, d_pos_users, d_pos_items, d_neg_users, d_neg_items, sched[c].first, sched[c].second,
                         flag, sp));
    MFB_CUDA(cudaEventRecord(m->ev_plan[c & 1], sp));
    return MFB_OK;
  };
  MFB_CHECK(plan(0));
  for (int64_t c = 0; c < nchunks; ++c) {
    const int cur = (int)(c & 1);
    if (c + 1 < nchunks) MFB_CHECK(plan(c + 1));
    MFB_CUDA(cudaStreamWaitEvent(st, m->ev_plan[cur], 0));
    MFB_CHECK(exec_chunk(m, m->plan[cur], sh, g, sched[c].first, sched[c].second, d_step_losses, st));
    MFB_CUDA(cudaEventRecord(m->ev_done[cur], st));
  }
  // ids were clamped on the device; report a bad id once, after the queue drains
  // (defer_check: the caller reads the flag -- m->ws_scalars -- together with its own results, one synchronisation)
  if (!defer_check) MFB_CHECK(check_err_flag(flag, st, train ? "train" : "loss"));
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

extern "C" int mfb_train_epoch(mfb_model *m, int loss, const int64_t *d_pos_users, const int64_t *d_pos_items,
                               int64_t n_pos, int32_t batch, int32_t n_neg, const int64_t *d_pop_users,
                               const int64_t *d_pop_items, int64_t pop_len, float *d_step_losses, mfb_stream stream) {
  if (n_neg > 0 && pop_len <= 0) return MFB_ERR_INVALID;
  return run_steps(m, loss, d_pos_users, d_pos_items, n_pos, batch, n_neg, nullptr, nullptr, d_step_losses,
                   (cudaStream_t)stream, true, d_pop_users, d_pop_items, n_neg > 0 ? pop_len : 0);
}

extern "C" int mfb_loss_epoch(mfb_model *m, int loss, const int64_t *d_pos_users, const int64_t *d_pos_items,
                              int64_t n_pos, int32_t batch, int32_t n_neg, const int64_t *d_pop_users,
                              const int64_t *d_pop_items, int64_t pop_len, float *d_step_losses, mfb_stream stream) {
  if (n_neg > 0 && pop_len <= 0) return MFB_ERR_INVALID;
  return run_steps(m, loss, d_pos_users, d_pos_items, n_pos, batch, n_neg, nullptr, nullptr, d_step_losses,
                   (cudaStream_t)stream, false, d_pop_users, d_pop_items, n_neg > 0 ? pop_len : 0);
}

extern "C" int mfb_train_epoch_host(mfb_model *m, int loss, const int64_t *h_pos_users, const int64_t *h_pos_items,
                                    int64_t n_pos, int32_t batch, int32_t n_neg, uint32_t *h_state,
                                    const int64_t *d_pop_users, const int64_t *d_pop_items, int64_t pop_len,
                                    float *h_step_losses, mfb_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!m || !h_pos_users || !h_pos_items || !h_step_losses) return MFB_ERR_INVALID;
  MFB_CHECK(validate_loss_shape(loss, n_pos, batch, n_neg));
  const int64_t nsteps = (n_pos + batch - 1) / batch;
  MFB_CHECK(m->ws_ids.reserve((size_t)n_pos * 2 * sizeof(int64_t)));
  MFB_CHECK(m->ws_losses.reserve((size_t)nsteps * sizeof(float)));
  int64_t *d_u = m->ws_ids.as<int64_t>();
  int64_t *d_i = d_u + n_pos;
  if (n_neg > 0) {
    if (!h_state || !d_pop_users || !d_pop_items || pop_len <= 0) return MFB_ERR_INVALID;
    MFB_CHECK(mfb_model_rng_seed(m, h_state, stream));
  }
  // the word generator starts as soon as the state is up; the id upload (pinned memory: asynchronous) runs beside it
  MFB_CHECK(ensure_streams(m));
  MFB_CUDA(cudaEventRecord(m->ev_seed, st));
  MFB_CUDA(cudaMemcpyAsync(d_u, h_pos_users, (size_t)n_pos * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  MFB_CUDA(cudaMemcpyAsync(d_i, h_pos_items, (size_t)n_pos * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  MFB_CHECK(run_steps(m, loss, d_u, d_i, n_pos, batch, n_neg, nullptr, nullptr, m->ws_losses.as<float>(), st, true,
                      d_pop_users, d_pop_items, n_neg > 0 ? pop_len : 0, m->ev_seed, true));
  // losses, the generator state and the id-range flag come back together: one synchronisation.  (The last generator
  // kernel precedes the last chunk's plan, which the last chunk's steps on `st` waited for.)
  int h_flag = 0;
  MFB_CUDA(cudaMemcpyAsync(h_step_losses, m->ws_losses.ptr, (size_t)nsteps * sizeof(float), cudaMemcpyDeviceToHost, st));
  MFB_CUDA(cudaMemcpyAsync(&h_flag, m->ws_scalars.ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
  if (n_neg > 0)
    MFB_CUDA(cudaMemcpyAsync(h_state, m->rng_state.ptr, 625 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  MFB_CUDA(cudaStreamSynchronize(st));
  if (h_flag) {
    mfb_set_error("train: id out of range for the model's tables");
    return MFB_ERR_RANGE;
  }
  return MFB_OK;
}
