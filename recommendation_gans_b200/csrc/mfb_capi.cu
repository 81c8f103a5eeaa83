// Model handle, error reporting and optimiser scalar tables of the mfb200 C ABI.
// mfb_model_create replaces ImplicitFactorizationModel._initialize (implicit.py:163-199): it binds
// the BilinearNet tables (torch-owned storage) and the optimiser hyper-parameters.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include "mfb_internal.cuh"

static thread_local char g_err[1024] = "";

void mfb_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int64_t g_library_launches = 0;
void mfb_count_library_launch(int n) { g_library_launches += n; }
extern "C" int64_t mfb_library_launches(void) { return g_library_launches; }

extern "C" const char *mfb_last_error(void) { return g_err; }
extern "C" int mfb_version(void) { return MFB200_VERSION; }

// step_size[t] = lr / (1 - beta1^t), bc2_sqrt[t] = sqrt(1 - beta2^t): computed in double exactly as
// torch/optim/adam.py does in Python floats, then rounded to fp32 when they meet fp32 tensors.
int mfb_ensure_scalars(mfb_model *m, int64_t upto) {
  if (m->desc.optimizer != MFB_OPT_ADAM) return MFB_OK;
  if (upto < m->scalars_cap) return MFB_OK;
  int64_t cap = m->scalars_cap > 0 ? m->scalars_cap : 4096;
  while (cap <= upto) cap *= 2;
  if (cap >= (1ll << 31)) {
    mfb_set_error("optimiser step count too large");
    return MFB_ERR_UNSUPPORTED;
  }
  m->h_step_size.resize(cap);
  m->h_bc2_sqrt.resize(cap);
  m->h_inv_bc2_sqrt.resize(cap);
  for (int64_t t = (m->scalars_cap > 0 ? m->scalars_cap : 0); t < cap; ++t) {
    if (t == 0) {
      m->h_step_size[0] = 0.f;
      m->h_bc2_sqrt[0] = 1.f;
      m->h_inv_bc2_sqrt[0] = 1.f;
      continue;
    }
    double bc1 = 1.0 - pow(m->desc.beta1, (double)t);
    double bc2 = 1.0 - pow(m->desc.beta2, (double)t);
    m->h_step_size[t] = (float)(m->desc.lr / bc1);
    m->h_bc2_sqrt[t] = (float)sqrt(bc2);
    m->h_inv_bc2_sqrt[t] = (float)(1.0 / sqrt(bc2));
  }
  // the old arrays may still be read by queued kernels: drain before replacing them
  MFB_CUDA(cudaDeviceSynchronize());
  MFB_CHECK(m->d_step_size.reserve((size_t)cap * sizeof(float)));
  MFB_CHECK(m->d_bc2_sqrt.reserve((size_t)cap * sizeof(float)));
  MFB_CHECK(m->d_inv_bc2_sqrt.reserve((size_t)cap * sizeof(float)));
  MFB_CUDA(cudaMemcpy(m->d_inv_bc2_sqrt.ptr, m->h_inv_bc2_sqrt.data(), (size_t)cap * sizeof(float),
                      cudaMemcpyHostToDevice));
  MFB_CUDA(cudaMemcpy(m->d_step_size.ptr, m->h_step_size.data(), (size_t)cap * sizeof(float), cudaMemcpyHostToDevice));
  MFB_CUDA(cudaMemcpy(m->d_bc2_sqrt.ptr, m->h_bc2_sqrt.data(), (size_t)cap * sizeof(float), cudaMemcpyHostToDevice));
  m->opt.step_size = m->d_step_size.as<float>();
  m->opt.bc2_sqrt = m->d_bc2_sqrt.as<float>();
  m->opt.inv_bc2_sqrt = m->d_inv_bc2_sqrt.as<float>();
  m->scalars_cap = cap;
  return MFB_OK;
}

extern "C" int mfb_model_create(const mfb_model_desc *d, mfb_model **out) {
  if (!d || !out) return MFB_ERR_INVALID;
  *out = nullptr;
  if (d->num_users <= 0 || d->num_items <= 0 || d->dim <= 0 || d->dim > 512) {
    mfb_set_error("model_create: bad shape users=%d items=%d dim=%d (dim must be 1..512)", d->num_users,
                  d->num_items, d->dim);
    return MFB_ERR_INVALID;
  }
  if (!d->d_user_emb || !d->d_item_emb || !d->d_user_bias || !d->d_item_bias) {
    mfb_set_error("model_create: null parameter table");
    return MFB_ERR_INVALID;
  }
  if (d->optimizer != MFB_OPT_SGD && d->optimizer != MFB_OPT_ADAM && d->optimizer != MFB_OPT_RMSPROP) {
    mfb_set_error("model_create: optimizer %d unsupported (SGD without momentum, Adam, RMSprop)", d->optimizer);
    return MFB_ERR_UNSUPPORTED;
  }
  if (d->optimizer == MFB_OPT_RMSPROP &&
      (!d->d_user_emb_v || !d->d_item_emb_v || !d->d_user_bias_v || !d->d_item_bias_v)) {
    mfb_set_error("model_create: RMSprop needs square_avg buffers (the *_v fields) for all four tables");
    return MFB_ERR_INVALID;
  }
  if (d->optimizer == MFB_OPT_ADAM &&
      (!d->d_user_emb_m || !d->d_user_emb_v || !d->d_item_emb_m || !d->d_item_emb_v || !d->d_user_bias_m ||
       !d->d_user_bias_v || !d->d_item_bias_m || !d->d_item_bias_v)) {
    mfb_set_error("model_create: Adam needs exp_avg / exp_avg_sq buffers for all four tables");
    return MFB_ERR_INVALID;
  }
  if (d->dim % 4 == 0) {
    const uintptr_t align = (uintptr_t)d->d_user_emb | (uintptr_t)d->d_item_emb | (uintptr_t)d->d_user_emb_m |
                            (uintptr_t)d->d_user_emb_v | (uintptr_t)d->d_item_emb_m | (uintptr_t)d->d_item_emb_v;
    if (align & 15) {
      mfb_set_error("model_create: tables must be 16-byte aligned");
      return MFB_ERR_INVALID;
    }
  }
  mfb_model *m = new mfb_model();
  m->desc = *d;
  if (const char *e = getenv("MFB_EAGER_MAX")) m->tune_eager_max = atoi(e);
  if (const char *e = getenv("MFB_CHUNK_BITS")) m->tune_chunk_bits = atoi(e) < 0 ? 0 : (atoi(e) > 12 ? 12 : atoi(e));
  if (const char *e = getenv("MFB_TC")) m->tune_tc = atoi(e);
  if (const char *e = getenv("MFB_TC_SAMPLE_STEP")) m->tune_tc_sample_step = atoi(e) < 1 ? 1 : atoi(e);
  if (const char *e = getenv("MFB_TC_FUSED_THR")) m->tune_tc_fused_thr = atoi(e);
  if (const char *e = getenv("MFB_TC_XK")) m->tune_tc_xk = atoi(e);
  if (const char *e = getenv("MFB_TC_TAIL_SPLIT")) m->tune_tc_tail_split = atoi(e);
  if (const char *e = getenv("MFB_TC_TILE_RADIUS")) m->tune_tc_tile_radius = atoi(e);
  if (const char *e = getenv("MFB_CU_BLOCKS")) m->tune_cu_blocks_per_sm = atoi(e) < 1 ? 1 : atoi(e);
  if (const char *e = getenv("MFB_CHUNK_RAMP")) m->tune_chunk_ramp = atoi(e) < 0 ? 0 : atoi(e);
  auto bind = [&](TableView &T, int rows, float *p, float *pm, float *pv, float *b, float *bm, float *bv) {
    T.p = p; T.m = pm; T.v = pv; T.bp = b; T.bm = bm; T.bv = bv; T.rows = rows; T.last = nullptr;
  };
  bind(m->users, d->num_users, d->d_user_emb, d->d_user_emb_m, d->d_user_emb_v, d->d_user_bias, d->d_user_bias_m,
       d->d_user_bias_v);
  bind(m->items, d->num_items, d->d_item_emb, d->d_item_emb_m, d->d_item_emb_v, d->d_item_bias, d->d_item_bias_m,
       d->d_item_bias_v);
  int rc = m->last_users.reserve((size_t)d->num_users * sizeof(int32_t));
  if (rc == MFB_OK) rc = m->last_items.reserve((size_t)d->num_items * sizeof(int32_t));
  if (rc != MFB_OK) {
    delete m;
    return rc;
  }
  cudaMemset(m->last_users.ptr, 0, (size_t)d->num_users * sizeof(int32_t));
  cudaMemset(m->last_items.ptr, 0, (size_t)d->num_items * sizeof(int32_t));
  m->users.last = m->last_users.as<int32_t>();
  m->items.last = m->last_items.as<int32_t>();
  OptView &o = m->opt;
  o.kind = d->optimizer;
  o.fast = d->fast_math;
  o.lr = (float)d->lr;
  o.beta1 = (float)d->beta1;
  o.beta2 = (float)d->beta2;
  o.eps = (float)d->eps;
  o.wd = (float)d->weight_decay;
  const float w = (float)(1.0 - d->beta1);  // exp_avg.lerp_(grad, 1 - beta1): weight is a Python double -> fp32
  o.lerp_small = fabsf(w) < 0.5f;
  o.lerp_coeff = o.lerp_small ? w : (w - 1.0f);
  o.one_minus_beta2 = (float)(1.0 - d->beta2);
  o.step_size = nullptr;
  o.bc2_sqrt = nullptr;
  o.inv_bc2_sqrt = nullptr;
  rc = mfb_ensure_scalars(m, 4095);
  if (rc != MFB_OK) {
    delete m;
    return rc;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    mfb_set_error("model_create: %s", cudaGetErrorString(e));
    delete m;
    return MFB_ERR_CUDA;
  }
  *out = m;
  return MFB_OK;
}

extern "C" int mfb_model_destroy(mfb_model *m) {
  if (!m) return MFB_OK;
  cudaDeviceSynchronize();
  DevBuf *bufs[] = {&m->d_step_size, &m->d_bc2_sqrt, &m->d_inv_bc2_sqrt, &m->last_users, &m->last_items, &m->ws_slots, &m->ws_keys_a,
                    &m->ws_keys_b, &m->ws_vals_a, &m->ws_vals_b, &m->ws_hist, &m->ws_rows, &m->ws_pred, &m->ws_dz,
                    &m->ws_scalars, &m->ws_ids, &m->ws_neg_u, &m->ws_neg_i, &m->ws_words, &m->ws_losses,
                    &m->ws_seg, &m->ws_partial, &m->ws_tickets};
  for (DevBuf *b : bufs) b->release();
  for (PlanBuf &pb : m->plan) {
    DevBuf *pbufs[] = {&pb.slots, &pb.keys_a, &pb.keys_b, &pb.vals_a, &pb.vals_b, &pb.seg, &pb.pred, &pb.gmax,
                       &pb.words, &pb.neg_u, &pb.neg_i, &pb.keys_c, &pb.vals_c, &pb.info, &pb.lazy_rows, &pb.lazy_cnt};
    for (DevBuf *b : pbufs) b->release();
  }
  m->rng_state.release();
  m->rng_jump.release();
  {
    DevBuf *ebufs[] = {&m->eval.ub, &m->eval.vb, &m->eval.unorm, &m->eval.vnorm, &m->eval.gmax, &m->eval.thr,
                       &m->eval.cand, &m->eval.cnt, &m->eval.redo, &m->eval.mcnt, &m->eval.mptr, &m->eval.mpairs,
                       &m->eval.cut, &m->eval.xk};
    for (DevBuf *b : ebufs) b->release();
  }
  if (m->st_plan) cudaStreamDestroy(m->st_plan);
  if (m->st_rng) cudaStreamDestroy(m->st_rng);
  for (cudaEvent_t e : m->ev_rng) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : m->ev_plan) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : m->ev_done) if (e) cudaEventDestroy(e);
  if (m->ev_join) cudaEventDestroy(m->ev_join);
  if (m->ev_seed) cudaEventDestroy(m->ev_seed);
  for (auto &r : m->prof.recs) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  for (auto e : m->prof.pool) cudaEventDestroy(e);
  delete m;
  return MFB_OK;
}

extern "C" int64_t mfb_model_step(const mfb_model *m) { return m ? m->step : -1; }

extern "C" int mfb_model_set_step(mfb_model *m, int64_t t) {
  if (!m || t < 0) return MFB_ERR_INVALID;
  if (m->flushed_step != m->step) {
    mfb_set_error("set_step: flush pending updates first");
    return MFB_ERR_INVALID;
  }
  // all rows are current: re-base the per-row counters
  cudaDeviceSynchronize();
  if (t != m->step) {
    // rows are all at m->step == flushed_step; make them current for t
    std::vector<int32_t> fill_u((size_t)m->users.rows, (int32_t)t), fill_i((size_t)m->items.rows, (int32_t)t);
    MFB_CUDA(cudaMemcpy(m->users.last, fill_u.data(), fill_u.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    MFB_CUDA(cudaMemcpy(m->items.last, fill_i.data(), fill_i.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  m->step = t;
  m->flushed_step = t;
  return mfb_ensure_scalars(m, t + 1);
}

static const char *kProfNames[MFB_PROFILE_CLASSES] = {"sample", "pack", "sort", "catchup", "forward", "loss",
                                                      "update", "flush", "predict", "topk", "step", "other"};

extern "C" const char *mfb_profile_name(int cls) {
  return (cls >= 0 && cls < MFB_PROFILE_CLASSES) ? kProfNames[cls] : "";
}

extern "C" int mfb_profile_enable(mfb_model *m, int on) {
  if (!m) return MFB_ERR_INVALID;
  m->prof.on = on != 0;
  return MFB_OK;
}

extern "C" int64_t mfb_model_launches(const mfb_model *m) { return m ? m->prof.launches : -1; }

extern "C" int mfb_profile_read(mfb_model *m, double *h_ms, int64_t *h_launches) {
  if (!m || !h_ms || !h_launches) return MFB_ERR_INVALID;
  MFB_CUDA(cudaDeviceSynchronize());
  for (int c = 0; c < MFB_PROFILE_CLASSES; ++c) {
    h_ms[c] = 0.0;
    h_launches[c] = 0;
  }
  for (auto &r : m->prof.recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      h_ms[r.cls] += ms;
      h_launches[r.cls] += 1;
    }
    m->prof.pool.push_back(r.a);
    m->prof.pool.push_back(r.b);
  }
  m->prof.recs.clear();
  return MFB_OK;
}

// ---- negative-sampler stream bound to the model -------------------------------------------------
extern "C" int mfb_model_rng_seed(mfb_model *m, const uint32_t *h_state, mfb_stream stream) {
  if (!m || !h_state) return MFB_ERR_INVALID;
  if (h_state[624] > 624) {
    mfb_set_error("MT19937 position %u out of range", h_state[624]);
    return MFB_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  MFB_CHECK(m->rng_state.reserve(625 * sizeof(uint32_t)));
  // earlier draws may still be queued on the planner stream
  if (m->st_plan) MFB_CUDA(cudaStreamSynchronize(m->st_plan));
  if (m->st_rng) MFB_CUDA(cudaStreamSynchronize(m->st_rng));
  MFB_CUDA(cudaMemcpyAsync(m->rng_state.ptr, h_state, 625 * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  MFB_CUDA(cudaStreamSynchronize(st));
  m->rng_seeded = true;
  return MFB_OK;
}

extern "C" int mfb_model_rng_state(mfb_model *m, uint32_t *h_state, mfb_stream stream) {
  if (!m || !h_state) return MFB_ERR_INVALID;
  if (!m->rng_seeded) {
    mfb_set_error("model rng was never seeded");
    return MFB_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (m->st_plan) MFB_CUDA(cudaStreamSynchronize(m->st_plan));
  if (m->st_rng) MFB_CUDA(cudaStreamSynchronize(m->st_rng));
  MFB_CUDA(cudaMemcpyAsync(h_state, m->rng_state.ptr, 625 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  MFB_CUDA(cudaStreamSynchronize(st));
  return MFB_OK;
}
