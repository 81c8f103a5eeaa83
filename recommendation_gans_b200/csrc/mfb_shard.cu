// Row-sharded implicit-MF training step (BASELINE cfg5: 10M users x 2M items x 128).
//
// The reference has one process and one set of tables (implicit.py:163-199).  At catalog sizes whose
// tables + Adam state are split over G GPUs, rank r OWNS the rows {g : g mod G == r} of all four tables
// (local index g / G) together with their moments and `last` counters, and COMPUTES a contiguous slice of
// every minibatch's slots.  One step is
//   owner    k_shard_catchup  rows requested this step are replayed to optimiser step t-1 (dense-optimiser
//                             semantics, see mfb_train.cu) -- unique rows only
//            k_shard_gather   requested rows (+ bias) -> send buffer, grouped by requesting rank
//   exchange all-to-all of rows over NVLink            (host: torch.distributed / NCCL)
//   compute  k_shard_forward  dot + biases + sigmoid per local slot; local adaptive-hinge maximum
//   exchange all-reduce(MAX) of the packed (probability, index) cell   (adaptive hinge only)
//   compute  k_shard_backward dLoss/dlogit per slot -> gradient rows, same layout as the received rows
//            k_shard_loss     this rank's partial loss sums (double, fixed order)
//   exchange all-to-all of gradient rows back to the owners
//   owner    k_shard_update   per unique row: ordered segment reduction of its gradient rows, optimiser
//                             step t (run_train_iteration, implicit.py:347-364, for that row)
//
// Every rank knows every id of the step (positives are replicated, negatives come from the same MT19937
// stream on every rank), so no id exchange is needed: the planner derives both directions' layouts and the
// per-pair row counts from the ids alone -- a stable radix sort of (step, owner, computing rank) keys for
// the exchange layout, and a stable sort of this rank's served rows by (step, table, row) for the
// deterministic, atomics-free gradient reduction.
#include <algorithm>

#include "mfb_internal.cuh"
#include "mfb_rowops.cuh"

namespace {

constexpr int SH_WARPS = 8;
constexpr int SH_THREADS = SH_WARPS * 32;
constexpr int SH_WIN = 32;        // long segments are cut at absolute multiples of SH_WIN (as UPD_WIN)
constexpr int SH_UPD_WARPS = 4;

// contiguous block partition of n units over G ranks (blocks differ by at most one; sharding.shard_range)
__host__ __device__ inline int part_lo(int r, int n, int G) {
  const int base = n / G, extra = n % G;
  return r * base + (r < extra ? r : extra);
}
__host__ __device__ inline int part_owner(int j, int n, int G) {
  const int base = n / G, extra = n % G;
  const int thr = extra * (base + 1);
  return j < thr ? j / (base + 1) : extra + (j - thr) / base;
}

struct Geom {
  long long n_pos, step0;
  int batch, m_neg, ns, Lfull, G, rank, gb, GP, rb, Lloc_cap;
};

constexpr int MAX_PEERS = 16;
struct Peers {
  char *base[MAX_PEERS];
};
struct XLayout {
  size_t recv_off, grecv_off, cells_off, flags_off, err_off, total;
};
enum { FLAG_ROWS = 0, FLAG_MAX = 1, FLAG_GRADS = 2 };

__device__ __forceinline__ int step_batch(const Geom &g, int s) {
  const long long first = (g.step0 + s) * g.batch;
  return (int)((g.n_pos - first < g.batch) ? (g.n_pos - first) : g.batch);
}

// ---- planner ---------------------------------------------------------------------------------
// entry (s, table, slot j): key = s | owner | computing rank, value = table<<31 | j.  Within a step the
// entries are laid out [user entries in slot order][item entries in slot order].
__global__ void k_shard_pack(const long long *__restrict__ pos_u, const long long *__restrict__ pos_i,
                             const long long *__restrict__ neg_u, const long long *__restrict__ neg_i, Geom g,
                             long long g_users, long long g_items, int *__restrict__ ids_u, int *__restrict__ ids_i,
                             uint32_t *__restrict__ keys, uint32_t *__restrict__ vals, int *__restrict__ err_flag) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)g.ns * g.Lfull) return;
  const int s = (int)(gid / g.Lfull), j = (int)(gid % g.Lfull);
  const int b = step_batch(g, s), L = b + g.m_neg;
  if (j >= L) return;
  long long u, i;
  int c;
  if (j < b) {
    const long long at = (g.step0 + s) * g.batch + j;
    u = pos_u[at];
    i = pos_i[at];
    c = part_owner(j, b, g.G);
  } else {
    const long long at = (long long)s * g.m_neg + (j - b);
    u = neg_u[at];
    i = neg_i[at];
    c = part_owner(j - b, g.m_neg, g.G);
  }
  if (u < 0 || u >= g_users || i < 0 || i >= g_items) {
    atomicExch(err_flag, 1);
    u = 0;
    i = 0;
  }
  const long long so = (long long)s * g.Lfull, ko = 2 * so;
  ids_u[so + j] = (int)u;
  ids_i[so + j] = (int)i;
  const uint32_t sk = (uint32_t)s << (2 * g.gb);
  keys[ko + j] = sk | ((uint32_t)(u % g.G) << g.gb) | (uint32_t)c;
  vals[ko + j] = (uint32_t)j;
  keys[ko + L + j] = sk | ((uint32_t)(i % g.G) << g.gb) | (uint32_t)c;
  vals[ko + L + j] = 0x80000000u | (uint32_t)j;
}

// start[k] = first sorted position whose key is >= k, k in [0, ns*GP*GP]
__global__ void k_shard_starts(const uint32_t *__restrict__ skeys, long long n, int nkeys, uint32_t *__restrict__ start) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > nkeys) return;
  long long lo = 0, hi = n;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (skeys[mid] < (uint32_t)k) lo = mid + 1; else hi = mid;
  }
  start[k] = (uint32_t)lo;
}

// per step: rows this rank serves (prefix over steps -> own_off) and, for the rows it receives, the offset of
// every owner's block in its receive buffer (rbase).  One block.
__global__ void k_shard_offsets(const uint32_t *__restrict__ start, Geom g, uint32_t *__restrict__ own_off,
                                uint32_t *__restrict__ rbase) {
  const int GP = g.GP, r = g.rank;
  for (int s = threadIdx.x; s < g.ns; s += blockDim.x) {
    uint32_t acc = 0;
    for (int o = 0; o < GP; ++o) {
      rbase[s * GP + o] = acc;
      acc += start[(s * GP + o) * GP + r + 1] - start[(s * GP + o) * GP + r];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t acc = 0;
    for (int s = 0; s < g.ns; ++s) {
      own_off[s] = acc;
      acc += start[(s * GP + r + 1) * GP] - start[(s * GP + r) * GP];
    }
    own_off[g.ns] = acc;
  }
}

__global__ void k_shard_layout(const uint32_t *__restrict__ skeys, const uint32_t *__restrict__ svals, long long n,
                               Geom g, const int *__restrict__ ids_u, const int *__restrict__ ids_i,
                               const uint32_t *__restrict__ start, const uint32_t *__restrict__ own_off,
                               const uint32_t *__restrict__ rbase, uint32_t *__restrict__ ent,
                               uint32_t *__restrict__ keys2, uint32_t *__restrict__ vals2, int *__restrict__ rpos,
                               uint32_t *__restrict__ sdst, uint32_t *__restrict__ gdst) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const uint32_t key = skeys[q], val = svals[q];
  const int GP = g.GP, r = g.rank;
  const int s = (int)(key >> (2 * g.gb)), o = (int)((key >> g.gb) & (uint32_t)(GP - 1)), c = (int)(key & (uint32_t)(GP - 1));
  const int t = (int)(val >> 31), j = (int)(val & 0x7fffffffu);
  const int id = (t ? ids_i : ids_u)[(long long)s * g.Lfull + j];
  if (o == r) {
    const uint32_t p = (uint32_t)q - start[(s * GP + r) * GP];
    const uint32_t at = own_off[s] + p;
    const uint32_t lrow = (uint32_t)(id / g.G);
    ent[at] = ((uint32_t)t << 31) | lrow;
    keys2[at] = ((uint32_t)s << (g.rb + 1)) | ((uint32_t)t << g.rb) | lrow;
    vals2[at] = p;
    // direct exchange: where this row lands in computing rank c's receive buffer (blocks of owners 0..r-1 first)
    uint32_t rb_c = 0;
    for (int o2 = 0; o2 < r; ++o2) rb_c += start[(s * GP + o2) * GP + c + 1] - start[(s * GP + o2) * GP + c];
    sdst[at] = ((uint32_t)c << 26) | (rb_c + ((uint32_t)q - start[(s * GP + r) * GP + c]));
  }
  if (c == r) {
    const int b = step_batch(g, s);
    const int b_lo = part_lo(r, b, g.G), b_loc = part_lo(r + 1, b, g.G) - b_lo;
    const int jl = (j < b) ? (j - b_lo) : (b_loc + (j - b - part_lo(r, g.m_neg, g.G)));
    const long long at = ((long long)t * g.ns + s) * g.Lloc_cap + jl;
    rpos[at] = (int)(rbase[s * GP + o] + ((uint32_t)q - start[(s * GP + o) * GP + r]));
    // direct exchange: the row's gradient goes to owner o, at the row's position in o's serve order
    gdst[at] = ((uint32_t)o << 26) | ((uint32_t)q - start[(s * GP + o) * GP]);
  }
}

// ---- owner: catch-up + gather ------------------------------------------------------------------
template <int VEC, int NIT, bool FAST>
__global__ void __launch_bounds__(SH_THREADS) k_shard_catchup(const PosInfo *__restrict__ info, long long base, int n,
                                                              int rb, TableView users, TableView items, OptView opt,
                                                              int D, int target) {
  const int lane = threadIdx.x & 31;
  const int ql = blockIdx.x * SH_WARPS + (threadIdx.x >> 5);
  if (ql >= n) return;
  const long long q = base + ql;
  const uint4 raw = *reinterpret_cast<const uint4 *>(info + q);
  if ((long long)raw.y != q) return;  // unique rows only
  const long long row = raw.x & ((1u << rb) - 1u);
  const TableView &T = ((raw.x >> rb) & 1u) ? items : users;
  const int last = T.last[row];
  if (last >= target) return;
  const int adam = opt_state_bits(opt.kind);
  RowState<VEC, NIT> r;
  row_load<VEC, NIT>(r, T, row, D, lane, adam);
  row_replay<VEC, NIT, FAST>(r, last, target, opt);
  row_store<VEC, NIT>(r, T, row, D, lane, adam);
  if (lane == 0) T.last[row] = target;
}

template <int VEC, int NIT>
__global__ void __launch_bounds__(SH_THREADS) k_shard_gather(const uint32_t *__restrict__ ent, int n, TableView users,
                                                             TableView items, int D, int Dp, int stride,
                                                             float *__restrict__ send) {
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * SH_WARPS + (threadIdx.x >> 5);
  if (p >= n) return;
  const uint32_t e = ent[p];
  const TableView &T = (e >> 31) ? items : users;
  const long long row = e & 0x7fffffffu;
  Frag<VEC, NIT> f;
  frag_load<VEC, NIT>(f, T.p + row * D, D, lane);
  float *dst = send + (long long)p * stride;
  frag_store<VEC, NIT>(f, dst, D, lane);
  if (lane == 0) dst[Dp] = T.bp[row];
}

// ---- compute rank: forward / backward / loss -----------------------------------------------------
template <int VEC, int NIT>
__global__ void __launch_bounds__(SH_THREADS) k_shard_forward(const int *__restrict__ rpos_u,
                                                              const int *__restrict__ rpos_i, int Lloc, int b_loc,
                                                              int neg_lo, const float *__restrict__ recv, int D, int Dp,
                                                              int stride, float *__restrict__ pred, int adaptive,
                                                              unsigned long long *__restrict__ gmax) {
  __shared__ unsigned long long wmax[SH_WARPS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int jl = blockIdx.x * SH_WARPS + wid;
  unsigned long long mine = 0ull;
  if (jl < Lloc) {
    const float *ru = recv + (long long)rpos_u[jl] * stride, *ri = recv + (long long)rpos_i[jl] * stride;
    Frag<VEC, NIT> fu, fi;
    frag_load_cg<VEC, NIT>(fu, ru, D, lane);   // exchange buffers are written by peers: L2-only loads
    frag_load_cg<VEC, NIT>(fi, ri, D, lane);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < NIT * VEC; ++k) acc = fmaf(fu.x[k], fi.x[k], acc);
    acc = warp_sum(acc);
    const float y = sigmoidf_acc((acc + __ldcg(ru + Dp)) + __ldcg(ri + Dp));
    if (lane == 0) pred[jl] = y;
    if (adaptive && jl >= b_loc) mine = pack_max(y, neg_lo + (jl - b_loc));
  }
  if (adaptive) {
    if (lane == 0) wmax[wid] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long best = wmax[0];
#pragma unroll
      for (int w = 1; w < SH_WARPS; ++w) best = wmax[w] > best ? wmax[w] : best;
      if (best > *reinterpret_cast<volatile unsigned long long *>(gmax)) atomicMax(gmax, best);
    }
  }
}

// dLoss/dlogit of local slot jl (formulas of mfb_train.cu slot_dz, with the GLOBAL batch sizes as normalisers)
template <int KIND>
__device__ __forceinline__ float shard_dz(int jl, int b_loc, int b_glob, int m_glob, int neg_lo,
                                          const float *__restrict__ pred, float gmax, int jstar) {
  const float x = pred[jl];
  float d;
  if (KIND == MFB_LOSS_POINTWISE) {
    const float den = fmaxf((1.0f - x) * x, 1e-12f);
    d = (jl < b_loc) ? __fdiv_rn(__fdiv_rn(x - 1.0f, den), (float)b_glob) : __fdiv_rn(__fdiv_rn(x, den), (float)m_glob);
  } else if (KIND == MFB_LOSS_HINGE) {
    const float inv_b = 1.0f / (float)b_glob;
    const float xp = (jl < b_loc) ? x : pred[jl - b_loc], xn = (jl < b_loc) ? pred[jl + b_loc] : x;
    const float a = (((xn - xp) + 1.0f) >= 0.0f) ? inv_b : 0.0f;
    d = (jl < b_loc) ? -a : a;
  } else if (KIND == MFB_LOSS_BPR) {
    const float inv_b = 1.0f / (float)b_glob;
    const float xp = (jl < b_loc) ? x : pred[jl - b_loc], xn = (jl < b_loc) ? pred[jl + b_loc] : x;
    const float sg = sigmoidf_acc(xp - xn);
    const float gg = ((-inv_b) * (1.0f - sg)) * sg;
    d = (jl < b_loc) ? gg : -gg;
  } else {
    const float inv_b = 1.0f / (float)b_glob;
    if (jl < b_loc) d = (((gmax - x) + 1.0f) >= 0.0f) ? -inv_b : 0.0f;
    else d = (neg_lo + (jl - b_loc) == jstar) ? (float)b_glob * inv_b : 0.0f;
  }
  return (d * (1.0f - x)) * x;
}

template <int VEC, int NIT, int KIND>
__global__ void __launch_bounds__(SH_THREADS) k_shard_backward(const int *__restrict__ rpos_u,
                                                               const int *__restrict__ rpos_i, int Lloc, int b_loc,
                                                               int b_glob, int m_glob, int neg_lo,
                                                               const float *__restrict__ recv, int D, int Dp, int stride,
                                                               const float *__restrict__ pred,
                                                               const unsigned long long *__restrict__ gmax_cell,
                                                               float *__restrict__ gsend,
                                                               const uint32_t *__restrict__ gdst_u,
                                                               const uint32_t *__restrict__ gdst_i,
                                                               char *const *__restrict__ peers, size_t grecv_off) {
  const int lane = threadIdx.x & 31;
  const int jl = blockIdx.x * SH_WARPS + (threadIdx.x >> 5);
  if (jl >= Lloc) return;
  unsigned long long gcell = 0ull;
  if (KIND == MFB_LOSS_ADAPTIVE_HINGE) gcell = *gmax_cell;
  const float gmax = unpack_max_val(gcell);
  const int jstar = (KIND == MFB_LOSS_ADAPTIVE_HINGE) ? unpack_max_idx(gcell) : -1;
  const long long pu = rpos_u[jl], pi = rpos_i[jl];
  float *gu, *gi;
  if (gdst_u != nullptr) {   // direct exchange: straight into the owners' gradient buffers
    const uint32_t du = gdst_u[jl], di = gdst_i[jl];
    gu = reinterpret_cast<float *>(peers[du >> 26] + grecv_off) + (long long)(du & 0x3ffffffu) * stride;
    gi = reinterpret_cast<float *>(peers[di >> 26] + grecv_off) + (long long)(di & 0x3ffffffu) * stride;
  } else {
    gu = gsend + pu * stride;
    gi = gsend + pi * stride;
  }
  Frag<VEC, NIT> fu, fi;
  // adaptive hinge: only the first maximal negative carries gradient; the others send zero rows
  const bool zero = KIND == MFB_LOSS_ADAPTIVE_HINGE && jl >= b_loc && (neg_lo + (jl - b_loc)) != jstar;
  float d = 0.f;
  if (zero) {
#pragma unroll
    for (int k = 0; k < NIT * VEC; ++k) fu.x[k] = fi.x[k] = 0.f;
  } else {
    frag_load_cg<VEC, NIT>(fu, recv + pu * stride, D, lane);
    frag_load_cg<VEC, NIT>(fi, recv + pi * stride, D, lane);
    d = shard_dz<KIND>(jl, b_loc, b_glob, m_glob, neg_lo, pred, gmax, jstar);
#pragma unroll
    for (int k = 0; k < NIT * VEC; ++k) {
      const float a = fu.x[k];
      fu.x[k] = __fmul_rn(d, fi.x[k]);   // d/dU[u] contribution: dz * V[i]
      fi.x[k] = __fmul_rn(d, a);         // d/dV[i] contribution: dz * U[u]
    }
  }
  // tail of an exchanged row: [Dp] bias gradient, [Dp+1] 1 = the row carries a gradient, 0 = all-zero row whose values
  // were not written (the owner then skips it; a row that only receives such entries is not touched at all and
  // catches up lazily -- identical under the dense-optimiser semantics, and half of the gradient traffic of
  // adaptive hinge, where only one negative of the whole batch has a gradient)
  if (!zero) {
    frag_store<VEC, NIT>(fu, gu, D, lane);
    frag_store<VEC, NIT>(fi, gi, D, lane);
  }
  if (lane == 0) {
    const float2 tail = make_float2(d, zero ? 0.f : 1.f);
    *reinterpret_cast<float2 *>(gu + Dp) = tail;
    *reinterpret_cast<float2 *>(gi + Dp) = tail;
  }
}


// ---- direct exchange over peer memory (NVLink): no collective call on the step's critical path -----------
// Every rank owns one exchange buffer (xbuf) that its peers map (CUDA IPC across processes): [receive rows]
// [gradient rows][adaptive-hinge cells, one per peer][flags: 3 kinds x one 64-bit sequence number per peer].
// Producers store straight into the consumer's buffer; a one-thread-per-peer signal kernel publishes the step's
// sequence number after the producing kernel (stream order + release.sys), and a one-warp wait kernel holds the
// consumer's stream until all peers have signalled (acquire.sys).  Buffers are reused every step: a peer can only
// overwrite a region after this rank signalled the phase that read it last.

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

template <int VEC, int NIT>
__global__ void __launch_bounds__(SH_THREADS) k_shard_gather_direct(const uint32_t *__restrict__ ent,
                                                                    const uint32_t *__restrict__ sdst, int n,
                                                                    TableView users, TableView items, int D, int Dp,
                                                                    int stride, char *const *__restrict__ peers,
                                                                    size_t recv_off) {
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * SH_WARPS + (threadIdx.x >> 5);
  if (p >= n) return;
  const uint32_t e = ent[p], d = sdst[p];
  const TableView &T = (e >> 31) ? items : users;
  const long long row = e & 0x7fffffffu;
  Frag<VEC, NIT> f;
  frag_load<VEC, NIT>(f, T.p + row * D, D, lane);
  float *dst = reinterpret_cast<float *>(peers[d >> 26] + recv_off) + (long long)(d & 0x3ffffffu) * stride;
  frag_store<VEC, NIT>(f, dst, D, lane);
  if (lane == 0) dst[Dp] = T.bp[row];
}

// One thread per peer c.  Publishes this rank's sequence number for `kind` in peer c's flag array (after copying
// the local adaptive-hinge cell into c's cell array when local_cell != nullptr) -- everything this rank stored into
// c's buffer earlier in the stream is ordered before it (kernel boundary + fence + release.sys) -- and then spins
// until peer c's own flag for `kind` arrives here (bounded: a dead peer must end in an error, not in a hung GPU).
// With out_cell, thread 0 finally folds the peers' cells into *out_cell.
__global__ void k_shard_signal_wait(char *const *__restrict__ peers, char *__restrict__ xbuf, int G, int me,
                                    size_t flags_off, int kind, unsigned long long seq, size_t cells_off,
                                    const unsigned long long *__restrict__ local_cell,
                                    unsigned long long *__restrict__ out_cell, int *__restrict__ err,
                                    long long timeout_clocks, int do_signal, int do_wait) {
  const int c = threadIdx.x;
  if (c < G && do_signal) {
    char *pb = peers[c];
    if (local_cell != nullptr) reinterpret_cast<unsigned long long *>(pb + cells_off)[me] = *local_cell;
    __threadfence_system();
    st_release_sys(reinterpret_cast<unsigned long long *>(pb + flags_off) + kind * MAX_PEERS + me, seq);
  }
  if (c < G && do_wait) {
    const unsigned long long *f = reinterpret_cast<const unsigned long long *>(xbuf + flags_off) + kind * MAX_PEERS + c;
    const long long t0 = clock64();
    while (ld_acquire_sys(f) < seq) {
      if (clock64() - t0 > timeout_clocks) {
        atomicExch(err, 1 + kind);
        break;
      }
      __nanosleep(32);
    }
  }
  __syncthreads();
  if (out_cell != nullptr && do_wait && threadIdx.x == 0) {
    const unsigned long long *cells = reinterpret_cast<const unsigned long long *>(xbuf + cells_off);
    unsigned long long best = 0ull;
    for (int k = 0; k < G; ++k) {
      const unsigned long long v = ld_acquire_sys(cells + k);
      best = v > best ? v : best;
    }
    *out_cell = best;
  }
}

constexpr int SHL_THREADS = 1024;
__device__ double shl_block_sum(double v, double *sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  double tot = 0.0;
  for (int w = 0; w < SHL_THREADS / 32; ++w) tot += sh[w];
  return tot;
}

// this rank's partial sums of the loss (spotlight/losses.py); the host adds the ranks' partials in rank order and
// divides by the global batch sizes.  out[0]: positive / pairwise sum, out[1]: negative sum (pointwise only).
__global__ void __launch_bounds__(SHL_THREADS) k_shard_loss(int kind, const float *__restrict__ pred, int b_loc,
                                                            int m_loc, const unsigned long long *__restrict__ gmax_cell,
                                                            double *__restrict__ out) {
  __shared__ double sh[SHL_THREADS / 32];
  const int tid = threadIdx.x;
  double s0 = 0.0, s1 = 0.0;
  if (kind == MFB_LOSS_POINTWISE) {
    for (int j = tid; j < b_loc; j += SHL_THREADS) s0 += (double)(-fmaxf(logf(pred[j]), -100.0f));
    for (int j = tid; j < m_loc; j += SHL_THREADS) s1 += (double)(-fmaxf(logf(1.0f - pred[b_loc + j]), -100.0f));
  } else if (kind == MFB_LOSS_HINGE) {
    for (int j = tid; j < b_loc; j += SHL_THREADS) s0 += (double)fmaxf((pred[b_loc + j] - pred[j]) + 1.0f, 0.0f);
  } else if (kind == MFB_LOSS_BPR) {
    for (int j = tid; j < b_loc; j += SHL_THREADS) s0 += (double)(1.0f - sigmoidf_acc(pred[j] - pred[b_loc + j]));
  } else {
    const float gmax = unpack_max_val(*gmax_cell);
    for (int j = tid; j < b_loc; j += SHL_THREADS) s0 += (double)fmaxf((gmax - pred[j]) + 1.0f, 0.0f);
  }
  s0 = shl_block_sum(s0, sh);
  s1 = shl_block_sum(s1, sh);
  if (tid == 0) {
    out[0] = s0;
    out[1] = s1;
  }
}

// ---- owner: ordered segment reduction of the received gradient rows + optimiser step t ---------------
struct ShUpdArgs {
  const PosInfo *info;
  const uint32_t *svals;   // sorted position -> position in the step's gradient buffer
  long long base;
  int n, rb, D, Dp, stride;
  TableView users, items;
  OptView opt;
  const float *grecv;
  float *partial;
  int pstride;
  int *tickets;
  int t;
};

template <int VEC, int NIT, bool FAST>
__global__ void __launch_bounds__(SH_UPD_WARPS * 32) k_shard_update(const ShUpdArgs a) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int D = a.D;
  const int adam = opt_state_bits(a.opt.kind);   // which per-row optimiser state exists
  const int ql = (int)blockIdx.x * SH_UPD_WARPS + wid;
  if (ql >= a.n) return;
  const long long q = a.base + ql;
  const uint4 raw = *reinterpret_cast<const uint4 *>(a.info + q);
  const long long first = raw.y;
  const bool head = first == q;
  if (!head && (ql % SH_WIN) != 0) return;
  const int fl = (int)(first - a.base);
  const int seg_end = fl + (int)raw.z;
  const int win_end = (ql / SH_WIN + 1) * SH_WIN;
  const int run_end = seg_end < win_end ? seg_end : win_end;
  const bool whole = head && run_end == seg_end;
  const long long row = raw.x & ((1u << a.rb) - 1u);
  const TableView &T = ((raw.x >> a.rb) & 1u) ? a.items : a.users;

  // tails first: [Dp] bias gradient, [Dp+1] != 0 when the row's values were written (see k_shard_backward)
  const float *src0 = a.grecv + (long long)a.svals[q] * a.stride;
  const float2 tail0 = __ldcg(reinterpret_cast<const float2 *>(src0 + a.Dp));
  if (whole && run_end == ql + 1 && tail0.y == 0.f) return;   // only a zero-gradient entry: the row is not touched now

  RowState<VEC, NIT> r;
  if (whole) row_load<VEC, NIT>(r, T, row, D, lane, adam);

  Frag<VEC, NIT> g;
  float gb = tail0.x;
  bool any = tail0.y != 0.f;
  if (any) {
    frag_load_cg<VEC, NIT>(g, src0, D, lane);
  } else {
#pragma unroll
    for (int k = 0; k < NIT * VEC; ++k) g.x[k] = 0.f;
  }
  for (int p0 = ql + 1; p0 < run_end; p0 += 2) {   // remaining rows of the run, two in flight, added in order
    const bool has2 = p0 + 1 < run_end;
    const float *s0 = a.grecv + (long long)a.svals[a.base + p0] * a.stride;
    const float *s1 = a.grecv + (long long)a.svals[a.base + (has2 ? p0 + 1 : p0)] * a.stride;
    const float2 t0 = __ldcg(reinterpret_cast<const float2 *>(s0 + a.Dp));
    const float2 t1 = __ldcg(reinterpret_cast<const float2 *>(s1 + a.Dp));
    const bool u0 = t0.y != 0.f, u1 = has2 && t1.y != 0.f;
    Frag<VEC, NIT> o0, o1;
    if (u0) frag_load_cg<VEC, NIT>(o0, s0, D, lane);
    if (u1) frag_load_cg<VEC, NIT>(o1, s1, D, lane);
    if (u0) {
#pragma unroll
      for (int k = 0; k < NIT * VEC; ++k) g.x[k] = __fadd_rn(g.x[k], o0.x[k]);
      gb = __fadd_rn(gb, t0.x);
      any = true;
    }
    if (u1) {
#pragma unroll
      for (int k = 0; k < NIT * VEC; ++k) g.x[k] = __fadd_rn(g.x[k], o1.x[k]);
      gb = __fadd_rn(gb, t1.x);
      any = true;
    }
  }
  if (whole && !any) return;   // every entry of the row was a zero-gradient one: nothing to apply, `last` stays

  if (!whole) {
    const int first_win = fl / SH_WIN, last_win = (seg_end - 1) / SH_WIN;
    const int npieces = last_win - first_win + 1;
    const int nwin = (a.n + SH_WIN - 1) / SH_WIN;
    const int pslot = head ? (nwin + ql / SH_WIN) : (ql / SH_WIN);
    float *pp = a.partial + (long long)pslot * a.pstride;
    frag_store<VEC, NIT>(g, pp, D, lane);
    if (lane == 0) pp[a.Dp] = gb;
    __threadfence();
    __syncwarp();
    int old = 0;
    if (lane == 0) old = atomicAdd(a.tickets + fl, 1);
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old != npieces - 1) return;
    if (lane == 0) a.tickets[fl] = 0;
    __threadfence();
    const float *hp = a.partial + (long long)(nwin + first_win) * a.pstride;
    frag_load_cg<VEC, NIT>(g, hp, D, lane);
    gb = __ldcg(hp + a.Dp);
    for (int w = first_win + 1; w <= last_win; ++w) {
      const float *wp = a.partial + (long long)w * a.pstride;
      Frag<VEC, NIT> o;
      frag_load_cg<VEC, NIT>(o, wp, D, lane);
#pragma unroll
      for (int k = 0; k < NIT * VEC; ++k) g.x[k] = __fadd_rn(g.x[k], o.x[k]);
      gb = __fadd_rn(gb, __ldcg(wp + a.Dp));
    }
    row_load<VEC, NIT>(r, T, row, D, lane, adam);
  }
  apply_step<VEC, NIT, FAST>(r, g, gb, a.opt, a.t);
  row_store<VEC, NIT>(r, T, row, D, lane, adam);
  if (lane == 0) T.last[row] = a.t;
}

inline int grid_warps(long long warps, int per) { return (int)((warps + per - 1) / per); }

}  // namespace

struct PlanSet {
  Geom g = {};
  bool planned = false;
  std::vector<uint32_t> h_own_off;
  uint32_t *skeys2 = nullptr, *svals2 = nullptr;
  DevBuf ent, k2a, k2b, v2a, v2b, segf, segl, info, rpos, pred, sdst, gdst;
  cudaEvent_t ev_planned = nullptr;   // recorded on the planning stream when the set is complete
  cudaEvent_t ev_used = nullptr;      // recorded on the step stream after the last kernel that reads the set
  bool wait_planned = false, was_used = false;
};

struct mfb_shard {
  mfb_model *m = nullptr;
  int rank = 0, world = 1, gb = 1, GP = 2;
  long long g_users = 0, g_items = 0;
  Shape shape = {4, 1};
  int Dp = 0, stride = 0;
  // Planner output is double-buffered: chunk c+1 is planned (on the caller's planning stream) into the set that chunk
  // c-1 used while chunk c executes from the other one.  `active` = the set the step functions read.
  PlanSet sets[2];
  int active = 0;
  std::vector<uint32_t> h_start;
  DevBuf ids_u, ids_i, keys_a, keys_b, vals_a, vals_b, hist, start, own_off, rbase, err, partial, tickets, gcell;
  int64_t launches = 0;
  // direct exchange over peer memory
  void *xbuf = nullptr;
  XLayout xl = {};
  Peers peers = {};
  DevBuf peers_dev;
  bool peers_set = false;
  unsigned long long seq = 0;
  int x_batch = 0, x_m_neg = 0;
};

extern "C" int mfb_shard_create(mfb_model *local, int32_t rank, int32_t world, int64_t global_users,
                                int64_t global_items, mfb_shard **out) {
  if (!local || !out || world < 1 || world > 64 || rank < 0 || rank >= world) {
    mfb_set_error("shard_create: bad arguments (rank %d of %d)", rank, world);
    return MFB_ERR_INVALID;
  }
  const long long lu = (global_users - rank + world - 1) / world, li = (global_items - rank + world - 1) / world;
  if (local->desc.num_users != lu || local->desc.num_items != li) {
    mfb_set_error("shard_create: rank %d of %d owns %lld user rows and %lld item rows of %lld x %lld, the local model has %d x %d",
                  rank, world, lu, li, (long long)global_users, (long long)global_items, local->desc.num_users,
                  local->desc.num_items);
    return MFB_ERR_INVALID;
  }
  mfb_shard *sh = new mfb_shard();
  sh->m = local;
  sh->rank = rank;
  sh->world = world;
  sh->gb = bits_for((uint32_t)(world - 1));
  sh->GP = 1 << sh->gb;
  sh->g_users = global_users;
  sh->g_items = global_items;
  int rc = pick_shape(local->desc.dim, &sh->shape);
  if (rc != MFB_OK) {
    delete sh;
    return rc;
  }
  sh->Dp = (local->desc.dim + 3) / 4 * 4;
  sh->stride = sh->Dp + 4;
  for (PlanSet &ps : sh->sets) {
    cudaEventCreateWithFlags(&ps.ev_planned, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ps.ev_used, cudaEventDisableTiming);
  }
  *out = sh;
  return MFB_OK;
}

extern "C" int mfb_shard_destroy(mfb_shard *sh) {
  if (!sh) return MFB_OK;
  cudaDeviceSynchronize();
  DevBuf *bufs[] = {&sh->ids_u, &sh->ids_i, &sh->keys_a, &sh->keys_b, &sh->vals_a, &sh->vals_b, &sh->hist, &sh->start,
                    &sh->own_off, &sh->rbase, &sh->err, &sh->partial, &sh->tickets, &sh->gcell, &sh->peers_dev};
  for (DevBuf *b : bufs) b->release();
  for (PlanSet &ps : sh->sets) {
    DevBuf *pbufs[] = {&ps.ent, &ps.k2a, &ps.k2b, &ps.v2a, &ps.v2b, &ps.segf, &ps.segl, &ps.info, &ps.rpos, &ps.pred,
                       &ps.sdst, &ps.gdst};
    for (DevBuf *b : pbufs) b->release();
    if (ps.ev_planned) cudaEventDestroy(ps.ev_planned);
    if (ps.ev_used) cudaEventDestroy(ps.ev_used);
  }
  if (sh->xbuf) cudaFree(sh->xbuf);
  delete sh;
  return MFB_OK;
}

extern "C" int32_t mfb_shard_row_stride(const mfb_shard *sh) { return sh ? sh->stride : -1; }
extern "C" int64_t mfb_shard_launches(const mfb_shard *sh) { return sh ? sh->launches : -1; }

extern "C" int mfb_shard_plan(mfb_shard *sh, const int64_t *d_pos_users, const int64_t *d_pos_items, int64_t n_pos,
                              int32_t batch, int32_t n_neg, const int64_t *d_neg_users, const int64_t *d_neg_items,
                              int64_t step0, int32_t nsteps, int64_t *h_counts, mfb_stream stream) {
  if (!sh || !d_pos_users || !d_pos_items || !h_counts || n_pos <= 0 || batch <= 0 || n_neg < 0 || nsteps <= 0 ||
      step0 < 0 || (n_neg > 0 && (!d_neg_users || !d_neg_items))) {
    mfb_set_error("shard_plan: bad arguments");
    return MFB_ERR_INVALID;
  }
  const int64_t total_steps = (n_pos + batch - 1) / batch;
  if (step0 + nsteps > total_steps) {
    mfb_set_error("shard_plan: steps [%lld, %lld) exceed the epoch's %lld", (long long)step0,
                  (long long)(step0 + nsteps), (long long)total_steps);
    return MFB_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  PlanSet &ps = sh->sets[1 - sh->active];   // the set the previous-but-one chunk used; the active one may be executing
  if (ps.was_used) MFB_CUDA(cudaStreamWaitEvent(st, ps.ev_used, 0));
  ps.planned = false;
  Geom &g = ps.g;
  g.n_pos = n_pos;
  g.step0 = step0;
  g.batch = batch;
  g.m_neg = n_neg * batch;
  g.ns = nsteps;
  g.Lfull = batch + g.m_neg;
  g.G = sh->world;
  g.rank = sh->rank;
  g.gb = sh->gb;
  g.GP = sh->GP;
  const int max_rows = std::max(sh->m->desc.num_users, sh->m->desc.num_items);
  g.rb = bits_for((uint32_t)(max_rows > 1 ? max_rows - 1 : 1));
  g.Lloc_cap = (batch + g.G - 1) / g.G + (g.m_neg + g.G - 1) / g.G;
  const int sb = bits_for((uint32_t)(nsteps > 1 ? nsteps - 1 : 1));
  if (sb + g.rb + 1 > 32 || sb + 2 * g.gb > 32 || (long long)g.Lfull >= (1ll << 31)) {
    mfb_set_error("shard_plan: %d steps x %d local rows do not fit the 32-bit sort keys; plan fewer steps per chunk",
                  nsteps, max_rows);
    return MFB_ERR_UNSUPPORTED;
  }
  // valid entries: all steps full except possibly the epoch's last one
  const int64_t last_first = (step0 + nsteps - 1) * (int64_t)batch;
  const int b_last = (int)std::min<int64_t>(batch, n_pos - last_first);
  const int64_t n_e = 2 * ((int64_t)(nsteps - 1) * g.Lfull + b_last + g.m_neg);
  const int64_t cap_e = 2 * (int64_t)nsteps * g.Lfull;
  MFB_CHECK(sh->ids_u.reserve((size_t)cap_e / 2 * sizeof(int)));
  MFB_CHECK(sh->ids_i.reserve((size_t)cap_e / 2 * sizeof(int)));
  MFB_CHECK(sh->keys_a.reserve((size_t)cap_e * 4));
  MFB_CHECK(sh->keys_b.reserve((size_t)cap_e * 4));
  MFB_CHECK(sh->vals_a.reserve((size_t)cap_e * 4));
  MFB_CHECK(sh->vals_b.reserve((size_t)cap_e * 4));
  MFB_CHECK(sh->err.reserve(sizeof(int)));
  const int nkeys = nsteps * g.GP * g.GP;
  MFB_CHECK(sh->start.reserve((size_t)(nkeys + 1) * 4));
  MFB_CHECK(sh->own_off.reserve((size_t)(nsteps + 1) * 4));
  MFB_CHECK(sh->rbase.reserve((size_t)nsteps * g.GP * 4));
  MFB_CHECK(ps.rpos.reserve((size_t)2 * nsteps * g.Lloc_cap * sizeof(int)));
  MFB_CHECK(ps.pred.reserve((size_t)nsteps * g.Lloc_cap * sizeof(float)));
  MFB_CHECK(ps.gdst.reserve((size_t)2 * nsteps * g.Lloc_cap * 4));
  if (2ll * g.Lfull >= (1ll << 26) || g.G > 64) {
    mfb_set_error("shard_plan: batch too large for the packed exchange positions");
    return MFB_ERR_UNSUPPORTED;
  }
  MFB_CUDA(cudaMemsetAsync(sh->err.ptr, 0, sizeof(int), st));
  {
    const long long nthreads = (long long)nsteps * g.Lfull;
    k_shard_pack<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(
        (const long long *)d_pos_users, (const long long *)d_pos_items, (const long long *)d_neg_users,
        (const long long *)d_neg_items, g, sh->g_users, sh->g_items, sh->ids_u.as<int>(), sh->ids_i.as<int>(),
        sh->keys_a.as<uint32_t>(), sh->vals_a.as<uint32_t>(), sh->err.as<int>());
    MFB_KERNEL_CHECK();
  }
  uint32_t *sk = nullptr, *sv = nullptr;
  MFB_CHECK(mfb_radix_sort_pairs(sh->keys_a.as<uint32_t>(), sh->vals_a.as<uint32_t>(), sh->keys_b.as<uint32_t>(),
                                 sh->vals_b.as<uint32_t>(), n_e, sb + 2 * g.gb, sh->hist, &sk, &sv, st));
  k_shard_starts<<<(nkeys + 1 + 255) / 256, 256, 0, st>>>(sk, n_e, nkeys, sh->start.as<uint32_t>());
  MFB_KERNEL_CHECK();
  k_shard_offsets<<<1, 256, 0, st>>>(sh->start.as<uint32_t>(), g, sh->own_off.as<uint32_t>(), sh->rbase.as<uint32_t>());
  MFB_KERNEL_CHECK();
  sh->h_start.resize((size_t)nkeys + 1);
  ps.h_own_off.resize((size_t)nsteps + 1);
  int h_err = 0;
  MFB_CUDA(cudaMemcpyAsync(sh->h_start.data(), sh->start.ptr, (size_t)(nkeys + 1) * 4, cudaMemcpyDeviceToHost, st));
  MFB_CUDA(cudaMemcpyAsync(ps.h_own_off.data(), sh->own_off.ptr, (size_t)(nsteps + 1) * 4, cudaMemcpyDeviceToHost, st));
  MFB_CUDA(cudaMemcpyAsync(&h_err, sh->err.ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
  MFB_CUDA(cudaStreamSynchronize(st));
  if (h_err) {
    mfb_set_error("shard_plan: id out of range for the global tables (%lld users, %lld items)", sh->g_users,
                  sh->g_items);
    return MFB_ERR_RANGE;
  }
  for (int s = 0; s < nsteps; ++s)
    for (int o = 0; o < g.G; ++o)
      for (int c = 0; c < g.G; ++c) {
        const int k = (s * g.GP + o) * g.GP + c;
        h_counts[((int64_t)s * g.G + o) * g.G + c] = (int64_t)sh->h_start[k + 1] - (int64_t)sh->h_start[k];
      }
  const int64_t n_own = ps.h_own_off[nsteps];
  const size_t own_cap = (size_t)std::max<int64_t>(n_own, 1);
  MFB_CHECK(ps.ent.reserve(own_cap * 4));
  MFB_CHECK(ps.k2a.reserve(own_cap * 4));
  MFB_CHECK(ps.k2b.reserve(own_cap * 4));
  MFB_CHECK(ps.v2a.reserve(own_cap * 4));
  MFB_CHECK(ps.v2b.reserve(own_cap * 4));
  MFB_CHECK(ps.segf.reserve(own_cap * 4));
  MFB_CHECK(ps.segl.reserve(own_cap * 4));
  MFB_CHECK(ps.info.reserve(own_cap * sizeof(PosInfo)));
  MFB_CHECK(ps.sdst.reserve(own_cap * 4));
  k_shard_layout<<<(unsigned)((n_e + 255) / 256), 256, 0, st>>>(
      sk, sv, n_e, g, sh->ids_u.as<int>(), sh->ids_i.as<int>(), sh->start.as<uint32_t>(), sh->own_off.as<uint32_t>(),
      sh->rbase.as<uint32_t>(), ps.ent.as<uint32_t>(), ps.k2a.as<uint32_t>(), ps.v2a.as<uint32_t>(),
      ps.rpos.as<int>(), ps.sdst.as<uint32_t>(), ps.gdst.as<uint32_t>());
  MFB_KERNEL_CHECK();
  ps.skeys2 = ps.k2a.as<uint32_t>();
  ps.svals2 = ps.v2a.as<uint32_t>();
  int sort2_kernels = 0;
  if (n_own > 0) {
    MFB_CHECK(mfb_radix_sort_pairs(ps.k2a.as<uint32_t>(), ps.v2a.as<uint32_t>(), ps.k2b.as<uint32_t>(),
                                   ps.v2b.as<uint32_t>(), n_own, sb + g.rb + 1, sh->hist, &ps.skeys2, &ps.svals2, st));
    const unsigned grid = (unsigned)((n_own + 255) / 256);
    k_segments<<<grid, 256, 0, st>>>(ps.skeys2, n_own, ps.segf.as<uint32_t>(), ps.segl.as<uint32_t>());
    MFB_KERNEL_CHECK();
    k_posinfo<<<grid, 256, 0, st>>>(ps.skeys2, n_own, ps.segf.as<uint32_t>(), ps.segl.as<uint32_t>(), nullptr,
                                    ps.info.as<PosInfo>());
    MFB_KERNEL_CHECK();
    sort2_kernels = 3 * ((sb + g.rb + 1 + 7) / 8) + 2;
  }
  // gradient-reduction scratch for the largest step of the chunk
  int64_t max_serve = 0;
  for (int s = 0; s < nsteps; ++s) max_serve = std::max<int64_t>(max_serve, ps.h_own_off[s + 1] - ps.h_own_off[s]);
  const int64_t nwin = (max_serve + SH_WIN - 1) / SH_WIN;
  MFB_CHECK(sh->partial.reserve((size_t)std::max<int64_t>(2 * nwin, 1) * sh->stride * sizeof(float)));
  const size_t old_cap = sh->tickets.cap;
  MFB_CHECK(sh->tickets.reserve((size_t)std::max<int64_t>(max_serve, 1) * sizeof(int)));
  if (sh->tickets.cap != old_cap) MFB_CUDA(cudaMemsetAsync(sh->tickets.ptr, 0, sh->tickets.cap, st));
  MFB_CHECK(mfb_ensure_scalars(sh->m, sh->m->step + nsteps + 1));
  sh->launches += 4 + 3 * ((sb + 2 * g.gb + 7) / 8) + sort2_kernels;
  MFB_CUDA(cudaEventRecord(ps.ev_planned, st));
  ps.planned = true;
  ps.wait_planned = true;     // the first step call makes its stream wait for the planning stream
  ps.was_used = false;
  sh->active = 1 - sh->active;
  return MFB_OK;
}

namespace {
struct StepView {
  int b, b_lo, b_loc, m_lo, m_loc, Lloc;
  long long base;
  int n_serve;
};
int step_view(const mfb_shard *sh, int s, StepView *v) {
  if (!sh) return MFB_ERR_INVALID;
  const PlanSet &ps = sh->sets[sh->active];
  if (!ps.planned || s < 0 || s >= ps.g.ns) {
    mfb_set_error("shard step %d: no plan covers it (call mfb_shard_plan first)", s);
    return MFB_ERR_INVALID;
  }
  const Geom &g = ps.g;
  const long long first = (g.step0 + s) * g.batch;
  v->b = (int)std::min<long long>(g.batch, g.n_pos - first);
  v->b_lo = part_lo(g.rank, v->b, g.G);
  v->b_loc = part_lo(g.rank + 1, v->b, g.G) - v->b_lo;
  v->m_lo = part_lo(g.rank, g.m_neg, g.G);
  v->m_loc = part_lo(g.rank + 1, g.m_neg, g.G) - v->m_lo;
  v->Lloc = v->b_loc + v->m_loc;
  v->base = ps.h_own_off[s];
  v->n_serve = (int)(ps.h_own_off[s + 1] - ps.h_own_off[s]);
  return MFB_OK;
}
// first use of a freshly planned set on the step stream: order it after the planner's last kernel
int sync_plan(mfb_shard *sh, cudaStream_t st) {
  PlanSet &ps = sh->sets[sh->active];
  if (ps.wait_planned) {
    MFB_CUDA(cudaStreamWaitEvent(st, ps.ev_planned, 0));
    ps.wait_planned = false;
  }
  return MFB_OK;
}
}  // namespace

extern "C" int mfb_shard_gather(mfb_shard *sh, int32_t s, float *d_send, mfb_stream stream) {
  StepView v;
  MFB_CHECK(step_view(sh, s, &v));
  cudaStream_t st = (cudaStream_t)stream;
  MFB_CHECK(sync_plan(sh, st));
  PlanSet &ps = sh->sets[sh->active];
  if (v.n_serve == 0) return MFB_OK;
  if (!d_send) return MFB_ERR_INVALID;
  mfb_model *m = sh->m;
  const int D = m->desc.dim;
  const int target = (int)m->step;   // the rows must be current for the step BEFORE the one being computed
  const bool fast = m->desc.fast_math != 0;
  const int grid = grid_warps(v.n_serve, SH_WARPS);
  {
#define CALL(V, N)                                                                                                   \
  if (fast)                                                                                                          \
    k_shard_catchup<V, N, true><<<grid, SH_THREADS, 0, st>>>(ps.info.as<PosInfo>(), v.base, v.n_serve, ps.g.rb,    \
                                                             m->users, m->items, m->opt, D, target);                 \
  else                                                                                                               \
    k_shard_catchup<V, N, false><<<grid, SH_THREADS, 0, st>>>(ps.info.as<PosInfo>(), v.base, v.n_serve, ps.g.rb,   \
                                                              m->users, m->items, m->opt, D, target);
    MFB_DISPATCH_SHAPE(sh->shape, CALL);
#undef CALL
    MFB_KERNEL_CHECK();
  }
#define CALL(V, N)                                                                                                 \
  k_shard_gather<V, N><<<grid, SH_THREADS, 0, st>>>(ps.ent.as<uint32_t>() + v.base, v.n_serve, m->users, m->items, D, \
                                                    sh->Dp, sh->stride, d_send);
  MFB_DISPATCH_SHAPE(sh->shape, CALL);
#undef CALL
  MFB_KERNEL_CHECK();
  sh->launches += 2;
  return MFB_OK;
}

extern "C" int mfb_shard_forward(mfb_shard *sh, int loss, int32_t s, const float *d_recv, int64_t *d_gmax_cell,
                                 mfb_stream stream) {
  StepView v;
  MFB_CHECK(step_view(sh, s, &v));
  if (loss < MFB_LOSS_POINTWISE || loss > MFB_LOSS_ADAPTIVE_HINGE) return MFB_ERR_INVALID;
  if ((loss == MFB_LOSS_HINGE || loss == MFB_LOSS_BPR) && sh->sets[sh->active].g.m_neg != v.b) {
    mfb_set_error("%s loss needs as many negatives as positives (got %d and %d)", loss == MFB_LOSS_HINGE ? "hinge" : "bpr",
                  sh->sets[sh->active].g.m_neg, v.b);
    return MFB_ERR_SHAPE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  MFB_CHECK(sync_plan(sh, st));
  PlanSet &ps = sh->sets[sh->active];
  const int adaptive = loss == MFB_LOSS_ADAPTIVE_HINGE;
  if (adaptive) {
    if (!d_gmax_cell) return MFB_ERR_INVALID;
    MFB_CUDA(cudaMemsetAsync(d_gmax_cell, 0, sizeof(int64_t), st));
  }
  if (v.Lloc == 0) return MFB_OK;
  if (!d_recv) return MFB_ERR_INVALID;
  const Geom &g = ps.g;
  const int *rpu = ps.rpos.as<int>() + (long long)s * g.Lloc_cap;
  const int *rpi = ps.rpos.as<int>() + ((long long)g.ns + s) * g.Lloc_cap;
  float *pred = ps.pred.as<float>() + (long long)s * g.Lloc_cap;
  const int D = sh->m->desc.dim;
#define CALL(V, N)                                                                                               \
  k_shard_forward<V, N><<<grid_warps(v.Lloc, SH_WARPS), SH_THREADS, 0, st>>>(                                    \
      rpu, rpi, v.Lloc, v.b_loc, v.m_lo, d_recv, D, sh->Dp, sh->stride, pred, adaptive, (unsigned long long *)d_gmax_cell);
  MFB_DISPATCH_SHAPE(sh->shape, CALL);
#undef CALL
  MFB_KERNEL_CHECK();
  sh->launches += 1;
  return MFB_OK;
}

static int shard_backward_impl(mfb_shard *sh, int loss, int32_t s, const float *d_recv, const int64_t *d_gmax_cell,
                               float *d_gsend, bool direct, double *d_loss_partial, cudaStream_t st) {
  StepView v;
  MFB_CHECK(step_view(sh, s, &v));
  if (loss < MFB_LOSS_POINTWISE || loss > MFB_LOSS_ADAPTIVE_HINGE || !d_loss_partial) return MFB_ERR_INVALID;
  if (loss == MFB_LOSS_ADAPTIVE_HINGE && !d_gmax_cell) return MFB_ERR_INVALID;
  MFB_CHECK(sync_plan(sh, st));
  PlanSet &ps = sh->sets[sh->active];
  const Geom &g = ps.g;
  const int *rpu = ps.rpos.as<int>() + (long long)s * g.Lloc_cap;
  const int *rpi = ps.rpos.as<int>() + ((long long)g.ns + s) * g.Lloc_cap;
  const float *pred = ps.pred.as<float>() + (long long)s * g.Lloc_cap;
  const int D = sh->m->desc.dim;
  const unsigned long long *cell = (const unsigned long long *)d_gmax_cell;
  const uint32_t *gdu = direct ? ps.gdst.as<uint32_t>() + (long long)s * g.Lloc_cap : nullptr;
  const uint32_t *gdi = direct ? ps.gdst.as<uint32_t>() + ((long long)g.ns + s) * g.Lloc_cap : nullptr;
  if (v.Lloc > 0) {
    if (!d_recv || (!d_gsend && !direct)) return MFB_ERR_INVALID;
    const int grid = grid_warps(v.Lloc, SH_WARPS);
#define CALLK(V, N, K)                                                                                          \
  k_shard_backward<V, N, K><<<grid, SH_THREADS, 0, st>>>(rpu, rpi, v.Lloc, v.b_loc, v.b, g.m_neg, v.m_lo, d_recv, D, \
                                                         sh->Dp, sh->stride, pred, cell, d_gsend, gdu, gdi,  \
                                                         sh->peers_dev.as<char *>(), sh->xl.grecv_off)
#define CALL(V, N)                                                          \
  switch (loss) {                                                           \
    case MFB_LOSS_POINTWISE: CALLK(V, N, MFB_LOSS_POINTWISE); break;        \
    case MFB_LOSS_BPR: CALLK(V, N, MFB_LOSS_BPR); break;                    \
    case MFB_LOSS_HINGE: CALLK(V, N, MFB_LOSS_HINGE); break;                \
    default: CALLK(V, N, MFB_LOSS_ADAPTIVE_HINGE); break;                   \
  }
    MFB_DISPATCH_SHAPE(sh->shape, CALL);
#undef CALL
#undef CALLK
    MFB_KERNEL_CHECK();
    sh->launches += 1;
  }
  k_shard_loss<<<1, SHL_THREADS, 0, st>>>(loss, pred, v.b_loc, v.m_loc, cell, d_loss_partial);
  MFB_KERNEL_CHECK();
  sh->launches += 1;
  return MFB_OK;
}

extern "C" int mfb_shard_backward(mfb_shard *sh, int loss, int32_t s, const float *d_recv, const int64_t *d_gmax_cell,
                                  float *d_gsend, double *d_loss_partial, mfb_stream stream) {
  return shard_backward_impl(sh, loss, s, d_recv, d_gmax_cell, d_gsend, false, d_loss_partial, (cudaStream_t)stream);
}

extern "C" int mfb_shard_update(mfb_shard *sh, int32_t s, const float *d_grecv, mfb_stream stream) {
  StepView v;
  MFB_CHECK(step_view(sh, s, &v));
  cudaStream_t st = (cudaStream_t)stream;
  MFB_CHECK(sync_plan(sh, st));
  PlanSet &ps = sh->sets[sh->active];
  mfb_model *m = sh->m;
  const int t = (int)m->step + 1;
  MFB_CHECK(mfb_ensure_scalars(m, t + 1));
  if (v.n_serve > 0) {
    if (!d_grecv) return MFB_ERR_INVALID;
    ShUpdArgs a;
    a.info = ps.info.as<PosInfo>();
    a.svals = ps.svals2;
    a.base = v.base;
    a.n = v.n_serve;
    a.rb = ps.g.rb;
    a.D = m->desc.dim;
    a.Dp = sh->Dp;
    a.stride = sh->stride;
    a.users = m->users;
    a.items = m->items;
    a.opt = m->opt;
    a.grecv = d_grecv;
    a.partial = sh->partial.as<float>();
    a.pstride = sh->stride;
    a.tickets = sh->tickets.as<int>();
    a.t = t;
    const bool fast = m->desc.fast_math != 0;
    const int grid = grid_warps(v.n_serve, SH_UPD_WARPS);
#define CALL(V, N)                                                            \
  if (fast) k_shard_update<V, N, true><<<grid, SH_UPD_WARPS * 32, 0, st>>>(a); \
  else k_shard_update<V, N, false><<<grid, SH_UPD_WARPS * 32, 0, st>>>(a);
    MFB_DISPATCH_SHAPE(sh->shape, CALL);
#undef CALL
    MFB_KERNEL_CHECK();
    sh->launches += 1;
  }
  m->step = t;   // the dense optimiser stepped every row; rows not served here catch up lazily
  MFB_CUDA(cudaEventRecord(ps.ev_used, st));   // the planner may recycle this set once the step has run
  ps.was_used = true;
  return MFB_OK;
}

// ---- direct exchange over peer memory ---------------------------------------------------------------------
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

extern "C" int mfb_shard_xbuf_alloc(mfb_shard *sh, int32_t batch, int32_t n_neg, void **d_ptr, int64_t *bytes) {
  if (!sh || batch <= 0 || n_neg < 0 || !d_ptr || !bytes) return MFB_ERR_INVALID;
  if (sh->world > MAX_PEERS) {
    mfb_set_error("direct exchange supports up to %d ranks", MAX_PEERS);
    return MFB_ERR_UNSUPPORTED;
  }
  const int G = sh->world;
  const long long m_neg = (long long)n_neg * batch, Lfull = batch + m_neg;
  const long long Lloc_cap = (batch + G - 1) / G + (m_neg + G - 1) / G;
  XLayout x;
  x.recv_off = 0;
  x.grecv_off = align_up((size_t)(2 * Lloc_cap) * sh->stride * sizeof(float), 256);
  x.cells_off = x.grecv_off + align_up((size_t)(2 * Lfull) * sh->stride * sizeof(float), 256);
  x.flags_off = x.cells_off + align_up(MAX_PEERS * 8, 256);
  x.err_off = x.flags_off + align_up(3 * MAX_PEERS * 8, 256);
  x.total = x.err_off + 256;
  if (sh->xbuf) {
    MFB_CUDA(cudaDeviceSynchronize());
    cudaFree(sh->xbuf);
    sh->xbuf = nullptr;
  }
  cudaError_t e = cudaMalloc(&sh->xbuf, x.total);
  if (e != cudaSuccess) {
    mfb_set_error("cudaMalloc(%zu) for the exchange buffer failed: %s", x.total, cudaGetErrorString(e));
    return MFB_ERR_NOMEM;
  }
  MFB_CUDA(cudaMemset((char *)sh->xbuf + x.cells_off, 0, x.total - x.cells_off));
  MFB_CHECK(sh->gcell.reserve(2 * sizeof(unsigned long long)));
  MFB_CUDA(cudaMemset(sh->gcell.ptr, 0, 2 * sizeof(unsigned long long)));
  MFB_CUDA(cudaDeviceSynchronize());
  sh->xl = x;
  sh->peers_set = false;
  sh->seq = 0;
  sh->x_batch = batch;
  sh->x_m_neg = (int)m_neg;
  *d_ptr = sh->xbuf;
  *bytes = (int64_t)x.total;
  return MFB_OK;
}

extern "C" int mfb_shard_xbuf_set_peers(mfb_shard *sh, void *const *peer_ptrs) {
  if (!sh || !sh->xbuf || !peer_ptrs) return MFB_ERR_INVALID;
  for (int c = 0; c < sh->world; ++c) {
    if (!peer_ptrs[c]) return MFB_ERR_INVALID;
    sh->peers.base[c] = (char *)peer_ptrs[c];
  }
  if (sh->peers.base[sh->rank] != (char *)sh->xbuf) {
    mfb_set_error("xbuf_set_peers: entry %d must be this rank's own buffer", sh->rank);
    return MFB_ERR_INVALID;
  }
  MFB_CHECK(sh->peers_dev.reserve(MAX_PEERS * sizeof(char *)));
  MFB_CUDA(cudaMemcpy(sh->peers_dev.ptr, sh->peers.base, MAX_PEERS * sizeof(char *), cudaMemcpyHostToDevice));
  sh->peers_set = true;
  return MFB_OK;
}

extern "C" int mfb_ipc_export(void *d_ptr, void *h_handle64) {
  if (!d_ptr || !h_handle64) return MFB_ERR_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  MFB_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)h_handle64, d_ptr));
  return MFB_OK;
}

extern "C" int mfb_ipc_open(const void *h_handle64, void **d_ptr) {
  if (!d_ptr || !h_handle64) return MFB_ERR_INVALID;
  cudaIpcMemHandle_t h;
  memcpy(&h, h_handle64, sizeof(h));
  MFB_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return MFB_OK;
}

extern "C" int mfb_ipc_close(void *d_ptr) {
  if (!d_ptr) return MFB_OK;
  MFB_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return MFB_OK;
}

// One step of the direct exchange is four phases; between two phases lies one signal/wait pair:
//   0 owner     catch-up, rows straight into the computing ranks' receive buffers        | ROWS
//   1 compute   forward (+ local adaptive-hinge maximum)                                 | MAX (adaptive hinge only)
//   2 compute   gradient rows straight into the owners' buffers, partial loss sums       | GRADS
//   3 owner     ordered reduction + optimiser step
// fused = true (one process per GPU): the signal and the wait of a pair are ONE kernel at the head of the next phase.
// fused = false: each phase ends with its signal and the next begins with the wait -- a host that drives several
// ranks of ONE device phase by phase on one stream (the 1-GPU tests) then never launches a wait before the matching
// signals, so nothing ever spins.
static int shard_phase(mfb_shard *sh, int loss, int s, int phase, double *d_loss_partial, cudaStream_t st, bool fused) {
  PlanSet &ps = sh->sets[sh->active];
  mfb_model *m = sh->m;
  const int G = sh->world, D = m->desc.dim;
  const bool fast = m->desc.fast_math != 0, adaptive = loss == MFB_LOSS_ADAPTIVE_HINGE;
  const XLayout &x = sh->xl;
  char *xb = (char *)sh->xbuf;
  float *recv = (float *)(xb + x.recv_off), *grecv = (float *)(xb + x.grecv_off);
  int *err = (int *)(xb + x.err_off);
  unsigned long long *local_cell = sh->gcell.as<unsigned long long>(), *global_cell = local_cell + 1;
  char *const *pd = sh->peers_dev.as<char *>();
  const long long timeout = 6000000000ll;   // ~3 s of SM clocks: a dead peer ends in an error, not in a hung GPU
  StepView v;
  MFB_CHECK(step_view(sh, s, &v));
  auto flag = [&](int kind, const unsigned long long *cell_in, unsigned long long *cell_out, int do_signal, int do_wait) {
    k_shard_signal_wait<<<1, 32, 0, st>>>(pd, xb, G, sh->rank, x.flags_off, kind, sh->seq, x.cells_off, cell_in, cell_out,
                                          err, timeout, do_signal, do_wait);
    sh->launches += 1;
  };
  if (phase == 0) {
    ++sh->seq;
    if (v.n_serve > 0) {
      const int grid = grid_warps(v.n_serve, SH_WARPS);
      const int target = (int)m->step;
#define CALL(V, N)                                                                                                   \
  if (fast)                                                                                                          \
    k_shard_catchup<V, N, true><<<grid, SH_THREADS, 0, st>>>(ps.info.as<PosInfo>(), v.base, v.n_serve, ps.g.rb,    \
                                                             m->users, m->items, m->opt, D, target);                 \
  else                                                                                                               \
    k_shard_catchup<V, N, false><<<grid, SH_THREADS, 0, st>>>(ps.info.as<PosInfo>(), v.base, v.n_serve, ps.g.rb,   \
                                                              m->users, m->items, m->opt, D, target);                \
  k_shard_gather_direct<V, N><<<grid, SH_THREADS, 0, st>>>(ps.ent.as<uint32_t>() + v.base,                          \
                                                           ps.sdst.as<uint32_t>() + v.base, v.n_serve, m->users,    \
                                                           m->items, D, sh->Dp, sh->stride, pd, x.recv_off);
      MFB_DISPATCH_SHAPE(sh->shape, CALL);
#undef CALL
      MFB_KERNEL_CHECK();
      sh->launches += 2;
    }
    if (!fused) flag(FLAG_ROWS, nullptr, nullptr, 1, 0);
  } else if (phase == 1) {
    flag(FLAG_ROWS, nullptr, nullptr, fused ? 1 : 0, 1);
    MFB_KERNEL_CHECK();
    MFB_CHECK(mfb_shard_forward(sh, loss, s, recv, (int64_t *)local_cell, (mfb_stream)st));
    if (adaptive && !fused) flag(FLAG_MAX, local_cell, nullptr, 1, 0);
  } else if (phase == 2) {
    if (adaptive) flag(FLAG_MAX, local_cell, global_cell, fused ? 1 : 0, 1);
    MFB_KERNEL_CHECK();
    MFB_CHECK(shard_backward_impl(sh, loss, s, recv, (const int64_t *)global_cell, nullptr, true, d_loss_partial, st));
    if (!fused) flag(FLAG_GRADS, nullptr, nullptr, 1, 0);
  } else {
    flag(FLAG_GRADS, nullptr, nullptr, fused ? 1 : 0, 1);
    MFB_KERNEL_CHECK();
    MFB_CHECK(mfb_shard_update(sh, s, grecv, (mfb_stream)st));
  }
  MFB_KERNEL_CHECK();
  return MFB_OK;
}

static int shard_direct_ready(mfb_shard *sh, const double *d_loss_partial, int s_begin, int s_end) {
  if (!sh) return MFB_ERR_INVALID;
  PlanSet &ps = sh->sets[sh->active];
  if (!ps.planned || !sh->peers_set || !d_loss_partial || s_begin < 0 || s_end > ps.g.ns || s_begin > s_end) {
    mfb_set_error("direct exchange: needs a plan, an exchange buffer with peers, and a step range inside the plan");
    return MFB_ERR_INVALID;
  }
  if (ps.g.batch != sh->x_batch || ps.g.m_neg != sh->x_m_neg) {
    mfb_set_error("direct exchange: the buffer was sized for batch %d / %d negatives, the plan has %d / %d",
                  sh->x_batch, sh->x_m_neg, ps.g.batch, ps.g.m_neg);
    return MFB_ERR_INVALID;
  }
  return MFB_OK;
}

// Steps [s_begin, s_end) of the planned chunk with the direct exchange: everything is enqueued on `stream`, no host
// synchronisation and no collective call.  d_loss_partial: 2 doubles per step (as mfb_shard_backward).
extern "C" int mfb_shard_run_steps(mfb_shard *sh, int loss, int32_t s_begin, int32_t s_end, double *d_loss_partial,
                                   mfb_stream stream) {
  MFB_CHECK(shard_direct_ready(sh, d_loss_partial, s_begin, s_end));
  cudaStream_t st = (cudaStream_t)stream;
  MFB_CHECK(sync_plan(sh, st));
  for (int s = s_begin; s < s_end; ++s)
    for (int phase = 0; phase < 4; ++phase)
      MFB_CHECK(shard_phase(sh, loss, s, phase, d_loss_partial + 2 * (s - s_begin), st, true));
  return MFB_OK;
}

// One phase (0..3, see above) of step s with unfused flags, for a host that drives several ranks of one device in
// lockstep: call phase p for every rank before phase p+1 for any.  d_loss_partial: the step's 2 doubles.
extern "C" int mfb_shard_run_phase(mfb_shard *sh, int loss, int32_t s, int32_t phase, double *d_loss_partial,
                                   mfb_stream stream) {
  MFB_CHECK(shard_direct_ready(sh, d_loss_partial, s, s + 1));
  if (phase < 0 || phase > 3) return MFB_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  MFB_CHECK(sync_plan(sh, st));
  return shard_phase(sh, loss, s, phase, d_loss_partial, st, false);
}

// Synchronises the stream and reports whether a wait on a peer timed out since the last check.
extern "C" int mfb_shard_direct_check(mfb_shard *sh, mfb_stream stream) {
  if (!sh || !sh->xbuf) return MFB_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  int h = 0;
  MFB_CUDA(cudaMemcpyAsync(&h, (char *)sh->xbuf + sh->xl.err_off, sizeof(int), cudaMemcpyDeviceToHost, st));
  MFB_CUDA(cudaStreamSynchronize(st));
  if (h) {
    MFB_CUDA(cudaMemsetAsync((char *)sh->xbuf + sh->xl.err_off, 0, sizeof(int), st));
    mfb_set_error("direct exchange: timed out waiting for a peer (%s)",
                  h == 1 ? "rows" : (h == 2 ? "adaptive-hinge maximum" : "gradient rows"));
    return MFB_ERR_CUDA;
  }
  return MFB_OK;
}
