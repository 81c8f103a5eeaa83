// Internal declarations shared by the mfb200 translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/mfb200.h"

void mfb_set_error(const char *fmt, ...);
void mfb_count_library_launch(int n);  // kernels launched by model-less entry points

#define MFB_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t err__ = (call);                                                             \
    if (err__ != cudaSuccess) {                                                             \
      mfb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(err__)); \
      return MFB_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

#define MFB_CHECK(expr)            \
  do {                             \
    int rc__ = (expr);             \
    if (rc__ != MFB_OK) return rc__; \
  } while (0)

#define MFB_KERNEL_CHECK() MFB_CUDA(cudaGetLastError())

// A growable device buffer owned by the library (scratch only; parameters belong to torch).
struct DevBuf {
  void *ptr = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return MFB_OK;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e != cudaSuccess) {
      mfb_set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
      return MFB_ERR_NOMEM;
    }
    cap = want;
    return MFB_OK;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
  }
  template <typename T>
  T *as() const {
    return reinterpret_cast<T *>(ptr);
  }
};

// Device-visible view of the model (passed by value to kernels).
struct TableView {
  float *p;      // [rows, dim]   parameters
  float *m;      // Adam exp_avg      (nullptr for SGD)
  float *v;      // Adam exp_avg_sq   (nullptr for SGD)
  float *bp;     // [rows]        bias parameter
  float *bm;
  float *bv;
  int32_t *last;  // [rows] optimiser step this row is current for
  int32_t rows;
};

struct OptView {
  int32_t kind;  // mfb_optimizer
  int32_t fast;  // approximate sqrt/div in replay
  float lr, beta1, beta2, eps, wd;
  float lerp_coeff;           // ATen lerp: w<0.5 ? w : w-1
  int32_t lerp_small;         // w < 0.5
  float one_minus_beta2;
  const float *step_size;     // [t] -> lr/(1-beta1^t)   (Python double -> fp32), index 1..cap
  const float *bc2_sqrt;      // [t] -> sqrt(1-beta2^t)
  const float *inv_bc2_sqrt;  // [t] -> 1/sqrt(1-beta2^t)  (fast-math replay)
};

enum ProfClass {
  PK_SAMPLE = 0, PK_PACK, PK_SORT, PK_CATCHUP, PK_FORWARD, PK_LOSS, PK_UPDATE, PK_FLUSH, PK_PREDICT, PK_TOPK,
  PK_STEP, PK_OTHER
};

struct Profiler {
  bool on = false;
  int64_t launches = 0;  // always counted
  struct Rec { int cls; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
  // usage: auto t = prof.begin(cls, st, nkernels); launch...; prof.end(t, st);
  int begin(int cls, cudaStream_t st, int nkernels = 1) {
    launches += nkernels;
    if (!on) return -1;
    Rec r; r.cls = cls; r.a = get(); r.b = get();
    cudaEventRecord(r.a, st);
    recs.push_back(r);
    return (int)recs.size() - 1;
  }
  void end(int tok, cudaStream_t st) {
    if (tok >= 0) cudaEventRecord(recs[tok].b, st);
  }
};

// Planner output for one chunk of steps (double-buffered: chunk c+1 is planned on its own stream
// while chunk c trains).
struct PlanBuf {
  DevBuf slots, keys_a, keys_b, vals_a, vals_b, keys_c, vals_c, seg, info, lazy_rows, lazy_cnt, pred, gmax, words, neg_u,
      neg_i;
  uint32_t *skeys = nullptr, *svals = nullptr;
};

// Workspaces of the tensor-core evaluation path (mfb_eval_tc.cu)
struct EvalBuf {
  DevBuf ub, vb, unorm, vnorm, gmax, thr, cand, cnt, redo, mcnt, mptr, mpairs, cut, xk;
  // user ranges of the last tensor-core pass: first user, users, candidate sub-lists per user (2 parities x item-tile splits)
  int n_ranges = 0, range_first[2] = {0, 0}, range_users[2] = {0, 0}, range_nsub[2] = {2, 2};
  int n_users_pad = 0;   // padded user count of the last tensor-core pass (layout of the per-user scratch arrays)
  // train-mask bitmap currently held in mpairs / mptr: caller-chosen key of the (user list, train CSR) it was built
  // from (0 = none) and the geometry it was built for
  uint64_t mask_key = 0;
  int mask_users = 0, mask_tiles = 0, mask_step = 0, mask_pad = 0;
  const void *mask_bits_ptr = nullptr, *mask_dirty_ptr = nullptr;
  // users re-done by the exact kernel in the last tensor-core pass: counted on the device, read on demand
  const int *redo_cnt_dev = nullptr;
  cudaStream_t redo_stream = nullptr;
};

struct mfb_model {
  mfb_model_desc desc;
  Profiler prof;
  EvalBuf eval;
  int tune_tc = 1, tune_tc_sample_step = 4;   // MFB_TC=0 forces the exact-fp32 evaluation kernel
  int tune_tc_xk = 1;                         // MFB_TC_XK=0: bias pre-store + per-score subtraction instead of the extra K = 16 MMA step
  int tune_tc_tile_radius = 1;                // MFB_TC_TILE_RADIUS=0: per-item error radius in the re-score (a gather)
  int tune_tc_tail_split = 0;                 // MFB_TC_TAIL_SPLIT=1: item tiles of the partial last wave of user blocks split over 2-4 CTAs (measured: -0.7 %)
  int tune_tc_fused_thr = 1;                  // MFB_TC_FUSED_THR=0: group maxima through HBM + selection kernel
  int last_topk_redo = 0;                     // users re-done by the exact kernel in the last mfb_topk call (-1: still on the device)
  PlanBuf plan[2];
  cudaStream_t st_plan = nullptr;   // planner stream
  cudaStream_t st_rng = nullptr;    // MT19937 word generation (sequential, one CTA) runs ahead of the planner here
  cudaEvent_t ev_plan[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_rng[2] = {nullptr, nullptr};
  cudaEvent_t ev_join = nullptr, ev_seed = nullptr;
  int num_sms = 148;
  DevBuf rng_state;            // device-resident MT19937 state (624 words + position) of the negative sampler
  DevBuf rng_jump;             // jump-ahead scratch of the multi-CTA generator (raw words + share states)
  bool rng_seeded = false;
  // tuning knobs (environment overrides read at model creation: MFB_EAGER_MAX, MFB_CHUNK_BITS, MFB_CU_BLOCKS)
  int tune_eager_max = 64, tune_chunk_bits = 6, tune_cu_blocks_per_sm = 2;
  int tune_chunk_ramp = 0;     // steps in the first planner chunk of a call (MFB_CHUNK_RAMP; 0 = no ramp)
  TableView users, items;
  OptView opt;
  int64_t step = 0;            // optimiser steps applied so far
  int64_t flushed_step = 0;    // all rows are current for this step
  // per-step scalar tables
  std::vector<float> h_step_size, h_bc2_sqrt, h_inv_bc2_sqrt;
  DevBuf d_step_size, d_bc2_sqrt, d_inv_bc2_sqrt;
  int64_t scalars_cap = 0;
  DevBuf last_users, last_items;
  // training workspaces
  DevBuf ws_slots, ws_keys_a, ws_keys_b, ws_vals_a, ws_vals_b, ws_hist, ws_rows, ws_pred, ws_dz, ws_scalars;
  DevBuf ws_ids, ws_neg_u, ws_neg_i, ws_words, ws_losses, ws_seg, ws_partial, ws_tickets;
};

int mfb_ensure_scalars(mfb_model *m, int64_t upto);

// ---- tensor-core evaluation (mfb_eval_tc.cu)
bool mfb_tc_supported(const mfb_model *m, int k);
int mfb_topk_tc(mfb_model *m, const int64_t *d_user_ids, int64_t n_users, const int64_t *d_train_indptr,
                const int32_t *d_train_indices, int32_t k, int32_t *d_out_ids, float *d_out_scores, cudaStream_t st,
                int (*exact_topk)(mfb_model *, const int64_t *, int64_t, const int64_t *, const int32_t *, int32_t,
                                  int32_t *, float *, cudaStream_t, const int *, const int *),
                int *h_n_redo, uint64_t plan_key);
int mfb_tc_stats(mfb_model *m, int n_users, long long *h_out, cudaStream_t st);
int mfb_tc_dump_scores(mfb_model *m, const int64_t *d_user_ids, int n_users, float *d_out, cudaStream_t st);

// ---- MT19937 (mfb_mt19937.cu)
int mfb_mt_generate(uint32_t *h_state, int64_t nwords, uint32_t *d_words, cudaStream_t st);
int mfb_mt_generate_async(uint32_t *d_state, int64_t nwords, uint32_t *d_words, cudaStream_t st);
// the same stream cut into shares that many CTAs generate at once (mfb_mt_jump.cu); falls back to the call above
int mfb_mt_generate_parallel(uint32_t *d_state, int64_t nwords, uint32_t *d_words, int64_t words_per_step,
                             DevBuf *scratch, cudaStream_t st);
int mfb_choices_async(const uint32_t *d_words, int64_t k, int64_t pop_len, const int64_t *d_pop_users,
                      const int64_t *d_pop_items, int64_t *d_out_users, int64_t *d_out_items, cudaStream_t st);

// ---- radix sort (mfb_sort.cu): stable LSD sort of (key,val) pairs on bits [0,nbits)
// keys_a/vals_a hold the input; result pointer returned in *out_keys/*out_vals (a or b).
int mfb_radix_sort_pairs(uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b, int64_t n,
                         int nbits, DevBuf &hist, uint32_t **out_keys, uint32_t **out_vals, cudaStream_t st);
