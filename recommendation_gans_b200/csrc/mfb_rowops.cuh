// Per-row device helpers shared by the replicated-table step (mfb_train.cu) and the row-sharded step
// (mfb_shard.cu): warp-held row fragments, torch.optim.Adam / SGD element updates (op order of
// torch/optim/adam.py _single_tensor_adam), dense zero-gradient replay, shape dispatch.
#pragma once
#include <math.h>

#include "mfb_internal.cuh"

namespace {

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

// L2-only loads (ld.global.cg): for scratch written by other SMs earlier in the same launch
template <int VEC, int NIT>
__device__ __forceinline__ void frag_load_cg(Frag<VEC, NIT> &f, const float *__restrict__ row, int D, int lane) {
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int e = (it * 32 + lane) * VEC;
    if constexpr (VEC == 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < D) v = __ldcg(reinterpret_cast<const float4 *>(row + e));
      f.x[it * 4 + 0] = v.x;
      f.x[it * 4 + 1] = v.y;
      f.x[it * 4 + 2] = v.z;
      f.x[it * 4 + 3] = v.w;
    } else {
      f.x[it] = (e < D) ? __ldcg(row + e) : 0.f;
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

// MUFU approximations with flush-to-zero: no denormal fix-up code around them (a second-moment
// below 1.2e-38 means |grad| < 1e-19, i.e. denom == eps either way).
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
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
    // bc2 carries 1/bc2_sqrt in fast mode (see row_replay/apply_step): one multiply, no division
    float denom = fmaf(sqrt_approx(v), bc2, o.eps);
    p = fmaf(__fmul_rn(neg_ss, m), rcp_approx(denom), p);
  } else {
    float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2), o.eps);
    p = __fadd_rn(p, __fdiv_rn(__fmul_rn(neg_ss, m), denom));
  }
}

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2, two IEEE fp32 operations per instruction) ----
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// Loop-invariant packed constants of the fast-mode Adam update
struct AdamPack {
  f32x2 wd, coeff, beta2, omb2, eps, minus1;
  int lerp_small;
};
__device__ __forceinline__ AdamPack make_adam_pack(const OptView &o) {
  AdamPack k;
  k.wd = pack2(o.wd, o.wd);
  k.coeff = pack2(o.lerp_coeff, o.lerp_coeff);
  k.beta2 = pack2(o.beta2, o.beta2);
  k.omb2 = pack2(o.one_minus_beta2, o.one_minus_beta2);
  k.eps = pack2(o.eps, o.eps);
  k.minus1 = pack2(-1.0f, -1.0f);
  k.lerp_small = o.lerp_small;
  return k;
}

// adam_elem<true> on two elements at once: same operations, same roundings, half the FP instructions
__device__ __forceinline__ void adam_pair_fast(float &p0, float &p1, float &m0, float &m1, float &v0, float &v1,
                                               float g0, float g1, f32x2 neg_ss, f32x2 inv_bc2, const AdamPack &k) {
  f32x2 p = pack2(p0, p1), m = pack2(m0, m1), v = pack2(v0, v1), g = pack2(g0, g1);
  g = fma2(k.wd, p, g);
  const f32x2 d = fma2(m, k.minus1, g);                   // g - m (exact: same as a subtraction)
  m = fma2(k.coeff, d, k.lerp_small ? m : g);
  v = add2(mul2(v, k.beta2), mul2(mul2(k.omb2, g), g));
  float s0, s1;
  unpack2(v, s0, s1);
  const f32x2 denom = fma2(pack2(sqrt_approx(s0), sqrt_approx(s1)), inv_bc2, k.eps);
  float d0, d1;
  unpack2(denom, d0, d1);
  p = fma2(mul2(neg_ss, m), pack2(rcp_approx(d0), rcp_approx(d1)), p);
  unpack2(p, p0, p1);
  unpack2(m, m0, m1);
  unpack2(v, v0, v1);
}

// torch.optim.SGD, momentum 0: p <- p - lr*(g + wd*p)
__device__ __forceinline__ void sgd_elem(float &p, float g, const OptView &o) {
  g = fmaf(o.wd, p, g);
  p = fmaf(-o.lr, g, p);
}

// which per-row state an optimiser keeps: bit 0 = exp_avg (Adam), bit 1 = exp_avg_sq (Adam) / square_avg (RMSprop)
__device__ __forceinline__ int opt_state_bits(int kind) {
  return kind == MFB_OPT_ADAM ? 3 : (kind == MFB_OPT_RMSPROP ? 2 : 0);
}

// One dense torch.optim.RMSprop update of a single element (torch/optim/rmsprop.py _single_tensor_rmsprop, momentum 0,
// not centered): grad += wd*p; square_avg = square_avg*alpha + (1-alpha)*grad*grad; avg = sqrt(square_avg) + eps;
// p += (-lr*grad)/avg.  (o.beta2 carries alpha.)
__device__ __forceinline__ void rms_elem(float &p, float &v, float g, const OptView &o) {
  g = fmaf(o.wd, p, g);
  v = __fadd_rn(__fmul_rn(v, o.beta2), __fmul_rn(__fmul_rn(o.one_minus_beta2, g), g));
  const float avg = __fadd_rn(__fsqrt_rn(v), o.eps);
  p = __fadd_rn(p, __fdiv_rn(__fmul_rn(-o.lr, g), avg));
}

// Full optimiser state of one table row held by a warp.
template <int VEC, int NIT>
struct RowState {
  Frag<VEC, NIT> p, m, v;
  float bp, bm, bv;  // bias (replicated in all lanes)
};

template <int VEC, int NIT>
__device__ __forceinline__ void row_load(RowState<VEC, NIT> &r, const TableView &T, long long row, int D, int lane,
                                         int state) {   // state bits: opt_state_bits()
  frag_load<VEC, NIT>(r.p, T.p + row * D, D, lane);
  r.bp = T.bp[row];
  if (state & 1) {
    frag_load<VEC, NIT>(r.m, T.m + row * D, D, lane);
    r.bm = T.bm[row];
  }
  if (state & 2) {
    frag_load<VEC, NIT>(r.v, T.v + row * D, D, lane);
    r.bv = T.bv[row];
  }
}

template <int VEC, int NIT>
__device__ __forceinline__ void row_store(const RowState<VEC, NIT> &r, const TableView &T, long long row, int D,
                                          int lane, int state) {
  frag_store<VEC, NIT>(r.p, T.p + row * D, D, lane);
  if (state & 1) frag_store<VEC, NIT>(r.m, T.m + row * D, D, lane);
  if (state & 2) frag_store<VEC, NIT>(r.v, T.v + row * D, D, lane);
  if (lane == 0) {
    T.bp[row] = r.bp;
    if (state & 1) T.bm[row] = r.bm;
    if (state & 2) T.bv[row] = r.bv;
  }
}

// Replays the dense zero-gradient updates of optimiser steps (from, to] on a row.
template <int VEC, int NIT, bool FAST>
__device__ __forceinline__ void row_replay(RowState<VEC, NIT> &r, int from, int to, const OptView &o) {
  if (o.kind == MFB_OPT_ADAM) {
    if constexpr (FAST && VEC == 4) {
      // Zero-gradient steps in fast mode: g = wd*p is folded into the moment updates
      //   m' = (1-w)*m + (w*wd)*p          v' = beta2*v + ((1-beta2)*wd^2 * p) * p
      // (8 packed FP instructions per element pair instead of 10; a few ulps from the IEEE op order, like the MUFU
      // sqrt/rcp this mode already uses -- the 1e-5 parity tests cover it); 8% faster at cfg5 (long replays).
      // ncu on the cfg5 catch-up kernel: XU (MUFU) pipe 81% busy, FMA pipe 42%, issue 47%.  Trading the MUFU.RCP for a
      // reciprocal carried from step to step and refined by two Newton updates (4 packed FMAs per pair) was measured
      // twice on B200 -- with a residual check + MUFU fallback: 30% slower; branch-free from optimiser step 64 on: 12%
      // slower (results identical to 1e-11) -- so the packed FMA forms are no cheaper than the MUFU they replace.
      const float w = o.lerp_small ? o.lerp_coeff : o.lerp_coeff + 1.0f;      // 1 - beta1
      const float c1 = 1.0f - w, cw = w * o.wd, k3 = (o.one_minus_beta2 * o.wd) * o.wd;
      const f32x2 c1_2 = pack2(c1, c1), cw_2 = pack2(cw, cw), k3_2 = pack2(k3, k3), b2_2 = pack2(o.beta2, o.beta2),
                  eps_2 = pack2(o.eps, o.eps);
      constexpr int NP = NIT * VEC / 2;
      f32x2 P[NP], M[NP], V[NP];
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        P[k] = pack2(r.p.x[2 * k], r.p.x[2 * k + 1]);
        M[k] = pack2(r.m.x[2 * k], r.m.x[2 * k + 1]);
        V[k] = pack2(r.v.x[2 * k], r.v.x[2 * k + 1]);
      }
      for (int s = from + 1; s <= to; ++s) {
        const float neg_ss = -__ldg(o.step_size + s);
        const float ibc = __ldg(o.inv_bc2_sqrt + s);
        const f32x2 neg_ss2 = pack2(neg_ss, neg_ss), ibc2 = pack2(ibc, ibc);
#pragma unroll
        for (int k = 0; k < NP; ++k) {
          M[k] = fma2(c1_2, M[k], mul2(cw_2, P[k]));
          V[k] = fma2(mul2(k3_2, P[k]), P[k], mul2(V[k], b2_2));
          float s0, s1;
          unpack2(V[k], s0, s1);
          const f32x2 denom = fma2(pack2(sqrt_approx(s0), sqrt_approx(s1)), ibc2, eps_2);
          float d0, d1;
          unpack2(denom, d0, d1);
          P[k] = fma2(mul2(neg_ss2, M[k]), pack2(rcp_approx(d0), rcp_approx(d1)), P[k]);
        }
        r.bm = fmaf(c1, r.bm, cw * r.bp);
        r.bv = fmaf(k3 * r.bp, r.bp, r.bv * o.beta2);
        r.bp = fmaf(neg_ss * r.bm, rcp_approx(fmaf(sqrt_approx(r.bv), ibc, o.eps)), r.bp);
      }
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        unpack2(P[k], r.p.x[2 * k], r.p.x[2 * k + 1]);
        unpack2(M[k], r.m.x[2 * k], r.m.x[2 * k + 1]);
        unpack2(V[k], r.v.x[2 * k], r.v.x[2 * k + 1]);
      }
      return;
    }
    for (int s = from + 1; s <= to; ++s) {
      float neg_ss = -__ldg(o.step_size + s);
      float bc2 = FAST ? __ldg(o.inv_bc2_sqrt + s) : __ldg(o.bc2_sqrt + s);
#pragma unroll
      for (int k = 0; k < NIT * VEC; ++k) adam_elem<FAST>(r.p.x[k], r.m.x[k], r.v.x[k], 0.f, neg_ss, bc2, o);
      adam_elem<FAST>(r.bp, r.bm, r.bv, 0.f, neg_ss, bc2, o);
    }
  } else if (o.kind == MFB_OPT_RMSPROP) {
    for (int s = from + 1; s <= to; ++s) {   // (wd == 0: p stays, square_avg still decays every step)
#pragma unroll
      for (int k = 0; k < NIT * VEC; ++k) rms_elem(r.p.x[k], r.v.x[k], 0.f, o);
      rms_elem(r.bp, r.bv, 0.f, o);
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


// (probability, index) packed so that an unsigned 64-bit max picks the largest probability and,
// among equals, the smallest index -- torch.max(neg, 0) returns the first maximal element.
__device__ __forceinline__ unsigned long long pack_max(float prob, int idx) {
  return ((unsigned long long)__float_as_uint(prob) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)idx);
}
__device__ __forceinline__ float unpack_max_val(unsigned long long v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ int unpack_max_idx(unsigned long long v) {
  return (int)(0xFFFFFFFFu - (uint32_t)(v & 0xFFFFFFFFull));
}

template <int VEC, int NIT, bool FAST>
__device__ __forceinline__ void apply_step(RowState<VEC, NIT> &r, const Frag<VEC, NIT> &g, float gb,
                                           const OptView &opt, int t) {
  if (opt.kind == MFB_OPT_ADAM) {
    const float neg_ss = -__ldg(opt.step_size + t);
    const float bc2 = FAST ? __ldg(opt.inv_bc2_sqrt + t) : __ldg(opt.bc2_sqrt + t);
    if constexpr (FAST && VEC == 4) {
      const AdamPack kp = make_adam_pack(opt);
      const f32x2 neg_ss2 = pack2(neg_ss, neg_ss), ibc2 = pack2(bc2, bc2);
#pragma unroll
      for (int k = 0; k < NIT * VEC; k += 2)
        adam_pair_fast(r.p.x[k], r.p.x[k + 1], r.m.x[k], r.m.x[k + 1], r.v.x[k], r.v.x[k + 1], g.x[k], g.x[k + 1], neg_ss2,
                       ibc2, kp);
      adam_elem<true>(r.bp, r.bm, r.bv, gb, neg_ss, bc2, opt);
      return;
    }
#pragma unroll
    for (int k = 0; k < NIT * VEC; ++k) adam_elem<FAST>(r.p.x[k], r.m.x[k], r.v.x[k], g.x[k], neg_ss, bc2, opt);
    adam_elem<FAST>(r.bp, r.bm, r.bv, gb, neg_ss, bc2, opt);
  } else if (opt.kind == MFB_OPT_RMSPROP) {
#pragma unroll
    for (int k = 0; k < NIT * VEC; ++k) rms_elem(r.p.x[k], r.v.x[k], g.x[k], opt);
    rms_elem(r.bp, r.bv, gb, opt);
  } else {
#pragma unroll
    for (int k = 0; k < NIT * VEC; ++k) sgd_elem(r.p.x[k], g.x[k], opt);
    sgd_elem(r.bp, gb, opt);
  }
}

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
    s->nit = (D <= 128) ? 4 : 16;
  }
  return MFB_OK;
}

#define MFB_DISPATCH_SHAPE(SH, CALL)            \
  do {                                          \
    if ((SH).vec == 4) {                        \
      if ((SH).nit == 1) { CALL(4, 1); }        \
      else if ((SH).nit == 2) { CALL(4, 2); }   \
      else { CALL(4, 4); }                      \
    } else {                                    \
      if ((SH).nit == 4) { CALL(1, 4); }        \
      else { CALL(1, 16); }                     \
    }                                           \
  } while (0)

// After the sort, equal keys (same step, table, row) are adjacent: seg_first[q] = first position of
// q's segment, seg_len[head] = its length.
__global__ void k_segments(const uint32_t *__restrict__ skeys, long long n, uint32_t *__restrict__ seg_first,
                           uint32_t *__restrict__ seg_len) {
  long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const uint32_t key = skeys[q];
  if (q == 0 || skeys[q - 1] != key) {  // head: upper bound of key in (q, n)
    long long lo = q + 1, hi = n;
    if (lo < n && skeys[lo] == key) {    // (most segments have length 1 and skip the search)
      while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (skeys[mid] <= key) lo = mid + 1; else hi = mid;
      }
    }
    seg_first[q] = (uint32_t)q;
    seg_len[q] = (uint32_t)(lo - q);
  } else {                               // inside a segment: lower bound of key in [0, q)
    long long lo = 0, hi = q - 1;        // skeys[q-1] == key, so the answer is <= q-1
    if (q >= 2 && skeys[q - 2] != key) {
      lo = q - 1;
    } else {
      while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (skeys[mid] < key) lo = mid + 1; else hi = mid;
      }
    }
    seg_first[q] = (uint32_t)lo;
  }
}

// One 16-byte record per sorted position, so k_update needs a single load for its bookkeeping.
struct __align__(16) PosInfo {
  uint32_t key;    // step | table | row
  uint32_t first;  // chunk-global position of the segment head
  uint32_t len;    // segment length
  uint32_t gap;    // steps to the row's next use in the chunk (0: none)
};

__global__ void k_posinfo(const uint32_t *__restrict__ skeys, long long n, const uint32_t *__restrict__ seg_first,
                          const uint32_t *__restrict__ seg_len, const uint32_t *__restrict__ seg_gap,
                          PosInfo *__restrict__ info) {
  long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const uint32_t first = seg_first[q];
  PosInfo p;
  p.key = skeys[q];
  p.first = first;
  p.len = seg_len[first];
  p.gap = seg_gap ? seg_gap[first] : 0u;
  info[q] = p;
}

int bits_for(uint32_t maxval) {
  int b = 1;
  while (b < 32 && (maxval >> b) != 0) ++b;
  return b;
}

}  // namespace
