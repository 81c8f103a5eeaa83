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
};

struct mfb_model {
  mfb_model_desc desc;
  TableView users, items;
  OptView opt;
  int64_t step = 0;            // optimiser steps applied so far
  int64_t flushed_step = 0;    // all rows are current for this step
  // per-step scalar tables
  std::vector<float> h_step_size, h_bc2_sqrt;
  DevBuf d_step_size, d_bc2_sqrt;
  int64_t scalars_cap = 0;
  DevBuf last_users, last_items;
  // training workspaces
  DevBuf ws_slots, ws_keys_a, ws_keys_b, ws_vals_a, ws_vals_b, ws_hist, ws_rows, ws_pred, ws_dz, ws_scalars;
  DevBuf ws_ids, ws_neg_u, ws_neg_i, ws_words, ws_losses;
};

int mfb_ensure_scalars(mfb_model *m, int64_t upto);

// ---- MT19937 (mfb_mt19937.cu)
int mfb_mt_generate(uint32_t *h_state, int64_t nwords, uint32_t *d_words, cudaStream_t st);

// ---- radix sort (mfb_sort.cu): stable LSD sort of (key,val) pairs on bits [0,nbits)
// keys_a/vals_a hold the input; result pointer returned in *out_keys/*out_vals (a or b).
int mfb_radix_sort_pairs(uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b, int64_t n,
                         int nbits, DevBuf &hist, uint32_t **out_keys, uint32_t **out_vals, cudaStream_t st);
