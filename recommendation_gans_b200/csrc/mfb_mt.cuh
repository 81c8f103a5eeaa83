// MT19937 constants and word functions shared by the generator (mfb_mt19937.cu) and the jump-ahead (mfb_mt_jump.cu).
#pragma once
#include <stdint.h>

namespace {

constexpr int MT_N = 624;
constexpr int MT_M = 397;
constexpr int MT_THREADS = 640;   // one thread per state word (624) in the generator kernel

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9D2C5680u;
  y ^= (y << 15) & 0xEFC60000u;
  y ^= y >> 18;
  return y;
}

__device__ __forceinline__ uint32_t mt_mix(uint32_t cur, uint32_t nxt, uint32_t far) {
  uint32_t y = (cur & 0x80000000u) | (nxt & 0x7FFFFFFFu);
  return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
}

// One CTA walks the stream sequentially, one thread per state word.  new[k] needs old[k], old[k+1] and
// new-or-old[(k+397) % 624]; the "new" operands are themselves mixes of old words (at most two levels deep, plus
// new[0] for k = 623), so every word of the next state is computed from the OLD state alone: one data-parallel phase
// and one barrier per 624-word regeneration instead of three dependent phases.
// state_in: 624 words + position (625 uint32); state_out (may be null): where the advanced state goes.
// out may be nullptr (advance only); raw: emit the state words themselves (no tempering) -- the jump-ahead's input.
// pos_override >= 0 replaces the stored position (624: start with a regeneration).
__device__ __forceinline__ void mt_generate_body(const uint32_t *state_in, uint32_t *state_out,
                                                 unsigned long long nwords, uint32_t *out, bool raw,
                                                 int pos_override = -1) {
  __shared__ uint32_t buf[2][MT_N];
  const int tid = threadIdx.x;
  constexpr int G = MT_N - MT_M;   // 227
  int cur = 0;
  for (int i = tid; i < MT_N; i += MT_THREADS) buf[0][i] = state_in[i];
  int pos = pos_override >= 0 ? pos_override : (int)state_in[MT_N];
  __syncthreads();
  unsigned long long emitted = 0;
  while (emitted < nwords) {
    if (pos >= MT_N) {
      const uint32_t *o = buf[cur];
      uint32_t *n = buf[cur ^ 1];
      for (int k = tid; k < MT_N; k += MT_THREADS) {
        uint32_t far;
        if (k < G) {
          far = o[k + MT_M];
        } else if (k < 2 * G) {
          far = mt_mix(o[k - G], o[k - G + 1], o[k - G + MT_M]);                 // new[k-227]
        } else {
          const uint32_t inner = mt_mix(o[k - 2 * G], o[k - 2 * G + 1], o[k - 2 * G + MT_M]);   // new[k-454]
          far = mt_mix(o[k - G], o[k - G + 1], inner);                             // new[k-227]
        }
        const uint32_t nxt = (k == MT_N - 1) ? mt_mix(o[0], o[1], o[MT_M]) : o[k + 1];   // k = 623 wraps to new[0]
        n[k] = mt_mix(o[k], nxt, far);
      }
      __syncthreads();   // the only barrier per regeneration: buffers alternate, so the words read above are not
      cur ^= 1;          // overwritten before the NEXT barrier
      pos = 0;
    }
    unsigned long long left = nwords - emitted;
    int take = (left < (unsigned long long)(MT_N - pos)) ? (int)left : (MT_N - pos);
    if (out != nullptr) {
      for (int i = tid; i < take; i += MT_THREADS) {
        const uint32_t w = buf[cur][pos + i];
        out[emitted + i] = raw ? w : mt_temper(w);
      }
    }
    emitted += take;
    pos += take;
  }
  __syncthreads();
  if (state_out != nullptr) {
    for (int i = tid; i < MT_N; i += MT_THREADS) state_out[i] = buf[cur][i];
    if (tid == 0) state_out[MT_N] = (uint32_t)pos;
  }
}

}  // namespace
