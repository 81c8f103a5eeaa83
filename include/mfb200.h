/*
 * mfb200 -- C ABI of the B200-native implicit matrix-factorisation hot path.
 *
 * Drop-in boundary for the reference's (Stamatios-Korres/recommendation_Gans) implicit-MF
 * fit / predict / evaluate path.  The reference has no FFI of its own (pure Python on torch);
 * each entry point below replaces the torch/numpy/CPython work done at the cited reference
 * lines, and is bound from Python with ctypes (recommendation_gans_b200/_native.py; the
 * reference-side stub is shown in INTEGRATION.md).
 *
 * Conventions
 *   - Every function returns MFB_OK (0) or a negative mfb_status; mfb_last_error() gives text.
 *   - Pointers named d_* are DEVICE pointers (borrowed for the call; torch owns parameter and
 *     optimiser-state storage), h_* are HOST pointers.  No torch types cross this boundary.
 *   - Ids are int64 (the reference keeps all ids as torch.int64, implicit.py:264-268).
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued on it.  Calls that return
 *     host data synchronise that stream before returning.
 *   - Not re-entrant per model handle; no internal threads.
 */
#ifndef MFB200_H
#define MFB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFB200_VERSION 100

typedef struct mfb_model mfb_model; /* opaque */
typedef void *mfb_stream;           /* cudaStream_t */

typedef enum {
  MFB_OK = 0,
  MFB_ERR_INVALID = -1,     /* bad argument */
  MFB_ERR_CUDA = -2,        /* CUDA runtime error (text in mfb_last_error) */
  MFB_ERR_RANGE = -3,       /* id >= table size  (reference: ValueError, implicit.py:222-236) */
  MFB_ERR_SHAPE = -4,       /* hinge/bpr with len(neg) != len(pos) (reference: torch broadcast RuntimeError) */
  MFB_ERR_UNSUPPORTED = -5, /* feature outside the hot path */
  MFB_ERR_NOMEM = -6
} mfb_status;

/* spotlight/losses.py:20,59,99,133 */
typedef enum {
  MFB_LOSS_POINTWISE = 0,
  MFB_LOSS_BPR = 1,
  MFB_LOSS_HINGE = 2,
  MFB_LOSS_ADAPTIVE_HINGE = 3
} mfb_loss;

/* spotlight/optimizers.py:4-22 -> torch.optim.SGD (momentum 0) / torch.optim.Adam / torch.optim.RMSprop (momentum 0,
 * not centered), all dense.  RMSprop: beta2 carries alpha, the *_v buffers are square_avg, the *_m buffers are unused. */
typedef enum { MFB_OPT_SGD = 0, MFB_OPT_ADAM = 1, MFB_OPT_RMSPROP = 2 } mfb_optimizer;

/* BilinearNet parameters (spotlight/factorization/representations.py:47-60) and the
 * optimiser configuration (implicit.py:182-192).  All tables row-major fp32 on the device. */
typedef struct {
  int32_t num_users;
  int32_t num_items;
  int32_t dim;
  int32_t optimizer; /* mfb_optimizer */
  double lr;
  double beta1;
  double beta2;
  double eps;
  double weight_decay;
  float *d_user_emb;  /* [num_users, dim] */
  float *d_item_emb;  /* [num_items, dim] */
  float *d_user_bias; /* [num_users, 1]   */
  float *d_item_bias; /* [num_items, 1]   */
  /* Adam moments, same shapes as the parameters (NULL for SGD) */
  float *d_user_emb_m, *d_user_emb_v;
  float *d_item_emb_m, *d_item_emb_v;
  float *d_user_bias_m, *d_user_bias_v;
  float *d_item_bias_m, *d_item_bias_v;
  int32_t fast_math; /* 0: IEEE sqrt/div in the optimiser replay (parity mode); 1: MUFU approximations */
  int32_t reserved;
} mfb_model_desc;

int mfb_version(void);
const char *mfb_last_error(void);

/* ---- model handle ------------------------------------------------------------------ */
/* Replaces ImplicitFactorizationModel._initialize (implicit.py:163-199): binds the four
 * tables and the optimiser.  `opt_step` starts at 0 (torch Adam state['step']). */
int mfb_model_create(const mfb_model_desc *desc, mfb_model **out);
int mfb_model_destroy(mfb_model *m);
int64_t mfb_model_step(const mfb_model *m);      /* optimiser steps taken so far */
int mfb_model_set_step(mfb_model *m, int64_t t); /* e.g. when resuming from torch optimiser state */

/* ---- MT19937 index streams (integer work, bit-exact) -------------------------------- */
/* State layout = CPython random.getstate()[1] / numpy RandomState.get_state()[1:3]:
 * 624 words followed by the position (h_state[624] in 0..624).  Updated in place. */

/* implicit.py:352,370  random.choices(neg_examples, k): writes the k sampled (user,item)
 * pairs, gathered from the population arrays, consuming exactly 2k words of the stream. */
int mfb_mt_choices_pairs(uint32_t *h_state, const int64_t *d_pop_users, const int64_t *d_pop_items,
                         int64_t pop_len, int64_t k, int64_t *d_out_users, int64_t *d_out_items,
                         mfb_stream stream);
/* same stream, indices only (random.choices(range(pop_len), k)) */
int mfb_mt_choices_indices(uint32_t *h_state, int64_t pop_len, int64_t k, int64_t *d_out, mfb_stream stream);
/* spotlight/sampling.py:33  random_state.randint(0, num_items, count, dtype=int64):
 * masked rejection on 32-bit draws; consumes a data-dependent number of words. */
int mfb_mt_sample_items(uint32_t *h_state, int64_t num_items, int64_t count, int64_t *d_out, mfb_stream stream);
/* raw tempered 32-bit outputs (test hook) */
int mfb_mt_words(uint32_t *h_state, int64_t nwords, uint32_t *d_out, mfb_stream stream);
/* The same words generated by many CTAs: the stream is cut into shares of whole steps (words_per_step words each
 * step) whose starting states come from a GF(2) jump-ahead of the one sequential stream random.choices draws from
 * (implicit.py:352,370); mfb_train_epoch uses this path when a step consumes 16384 words or more.  Short streams fall
 * back to the one-CTA generator.  Same words, same state handed back as mfb_mt_words. */
int mfb_mt_words_parallel(uint32_t *h_state, int64_t nwords, int64_t words_per_step, uint32_t *d_out,
                          mfb_stream stream);
/* host only: the jump polynomial x^jump mod phi (phi = characteristic polynomial of MT19937's one-word transition) as
 * 312 little-endian 64-bit words; word n + jump of the stream is the XOR of the words n + i over its set bits i. */
int mfb_mt_jump_poly(int64_t jump, uint64_t *h_out);

/* spotlight/sampling.py:46-70 get_negative_samples (with the rank-shift resampling of :37-44): num_samples uniform
 * (user, item) pairs; a pair that is a known interaction (key CSR: the entries whose stored value == 1, as
 * Interactions.has_key, interactions.py:159-160) has its item re-drawn uniformly outside the user's row (row CSR: every
 * stored entry).  Consumes numpy's legacy MT19937 stream exactly as the reference (np.random.choice x2, then one
 * np.random.randint per re-draw in sample order); h_state is advanced in place.  CSR indices sorted within rows. */
int mfb_negative_pairs(uint32_t *h_state, int64_t num_users, int64_t num_items, int64_t num_samples,
                       const int64_t *d_key_indptr, const int32_t *d_key_indices, const int64_t *d_row_indptr,
                       const int32_t *d_row_indices, int64_t *d_out_users, int64_t *d_out_items, int64_t *h_n_redrawn,
                       mfb_stream stream);

/* ---- forward / predict ---------------------------------------------------------------- */
/* BilinearNet.forward (representations.py:62-91) on m (user,item) pairs:
 * out[j] = sigmoid(<U[u_j],V[i_j]> + bu[u_j] + bi[i_j]).  Also ImplicitFactorizationModel.predict
 * in its pairwise form (implicit.py:381-415).  Brings lazily-updated rows up to date first. */
int mfb_predict_pairs(mfb_model *m, const int64_t *d_users, const int64_t *d_items, int64_t count,
                      float *d_out, mfb_stream stream);
/* predict(user_id) for all items (implicit.py:410-415, _components.py:8-25) */
int mfb_predict_user(mfb_model *m, int64_t user, float *d_out, mfb_stream stream);

/* ---- losses on probability vectors (spotlight/losses.py:20-172, 1-D tensors) ---------- */
/* d_loss: 1 float.  d_dpos/d_dneg may be NULL (forward only); otherwise dLoss/dpred. */
int mfb_loss_forward_backward(int loss, const float *d_pos, int64_t n_pos, const float *d_neg, int64_t n_neg,
                              float *d_loss, float *d_dpos, float *d_dneg, mfb_stream stream);

/* Function-level forms of the same losses that the model driver never produces: `d_mask` (NULL or n_pos floats;
 * loss*mask summed and divided by mask.sum(), losses.py:51-55,91-95,124-128) and 2-D negatives
 * [neg_rows, neg_cols == n_pos] row-major: adaptive hinge takes the per-positive maximum over dim 0 (first maximal row,
 * losses.py:170); hinge and bpr broadcast the positives over the rows (mean over neg_rows*n_pos entries).
 * neg_rows == 0: 1-D negatives of length neg_cols. */
int mfb_loss_forward_backward_ex(int loss, const float *d_pos, int64_t n_pos, const float *d_neg, int64_t neg_rows,
                                 int64_t neg_cols, const float *d_mask, float *d_loss, float *d_dpos, float *d_dneg,
                                 mfb_stream stream);

/* ---- training ------------------------------------------------------------------------- */
/* Fused replacement of the inner loop of ImplicitFactorizationModel.fit
 * (implicit.py:290-298 -> run_train_iteration, implicit.py:347-364) over
 * nsteps = ceil(n_pos / batch) consecutive minibatches:
 *   step s uses positives [s*batch, min((s+1)*batch, n_pos)) and the n_neg*batch negative
 *   pairs d_neg_*[s*n_neg*batch ...] (k uses the full batch even on the last partial batch).
 * Forward, loss, backward and the optimiser update have DENSE-optimiser semantics (every row
 * of every table steps every iteration) implemented with row-sparse traffic: rows not in a
 * batch are brought up to date lazily (mfb_flush, or when next gathered).
 * d_step_losses[s] receives the batch loss of step s (what loss.item() returns). */
int mfb_train_steps(mfb_model *m, int loss, const int64_t *d_pos_users, const int64_t *d_pos_items,
                    int64_t n_pos, int32_t batch, int32_t n_neg, const int64_t *d_neg_users,
                    const int64_t *d_neg_items, float *d_step_losses, mfb_stream stream);
/* run_val_iteration (implicit.py:366-379) over consecutive minibatches: same batching, no update. */
int mfb_loss_steps(mfb_model *m, int loss, const int64_t *d_pos_users, const int64_t *d_pos_items,
                   int64_t n_pos, int32_t batch, int32_t n_neg, const int64_t *d_neg_users,
                   const int64_t *d_neg_items, float *d_step_losses, mfb_stream stream);
/* Apply all pending dense-optimiser updates to every row (before any external read of the
 * tables: predict, evaluation, best-model copy, checkpoint; implicit.py:321-324,338-343). */
int mfb_flush(mfb_model *m, mfb_stream stream);

/* Negative-sampler stream bound to the model: the MT19937 state (random.getstate()[1] layout) lives on
 * the device between calls, so whole epochs run without a host round trip per draw. */
int mfb_model_rng_seed(mfb_model *m, const uint32_t *h_state, mfb_stream stream);
int mfb_model_rng_state(mfb_model *m, uint32_t *h_state, mfb_stream stream); /* synchronises */
/* One epoch of implicit.py:289-298 with the negatives drawn on the device, chunk by chunk, from the
 * model's stream: step s uses random.choices(population, k = n_neg*batch) exactly as implicit.py:352
 * (2 words per sample, consumed in step order).  Sampling and planning of chunk c+1 overlap the
 * training of chunk c.  mfb_loss_epoch is the validation pass (implicit.py:308-320; it consumes the
 * same stream, implicit.py:370). */
int mfb_train_epoch(mfb_model *m, int loss, const int64_t *d_pos_users, const int64_t *d_pos_items, int64_t n_pos,
                    int32_t batch, int32_t n_neg, const int64_t *d_pop_users, const int64_t *d_pop_items,
                    int64_t pop_len, float *d_step_losses, mfb_stream stream);
int mfb_loss_epoch(mfb_model *m, int loss, const int64_t *d_pos_users, const int64_t *d_pos_items, int64_t n_pos,
                   int32_t batch, int32_t n_neg, const int64_t *d_pop_users, const int64_t *d_pop_items,
                   int64_t pop_len, float *d_step_losses, mfb_stream stream);

/* Host-buffer entry (end-to-end path): positives in pageable/pinned HOST memory, negatives drawn
 * on the device from the MT19937 stream in h_state (random.choices semantics over the device
 * population arrays); per-step losses are returned in h_step_losses.  H2D of the ids and D2H of
 * the losses happen inside the call.  Equivalent to one training epoch of implicit.py:289-298. */
int mfb_train_epoch_host(mfb_model *m, int loss, const int64_t *h_pos_users, const int64_t *h_pos_items,
                         int64_t n_pos, int32_t batch, int32_t n_neg, uint32_t *h_state,
                         const int64_t *d_pop_users, const int64_t *d_pop_items, int64_t pop_len,
                         float *h_step_losses, mfb_stream stream);

/* ---- full-catalog evaluation ------------------------------------------------------------ */
/* Scores every listed user against all items and keeps the top-k item ids (any 1 <= k <= num_items),
 * best first, ties -> lower item id, ranking on the pre-sigmoid score; items in the user's
 * train row (CSR, sorted indices; may be NULL) rank last (evaluation.py:160-169).
 * Replaces the per-user predict + argsort loop of precision_recall_score, which accepts any k
 * (evaluation.py:144-150).  k <= MFB_MAX_TOPK with embedding_dim <= 128 takes the tensor-core path
 * (bit-identical ids); larger k the exact fp32 kernel, in passes of 256 ranks. */
#define MFB_MAX_TOPK 32
int mfb_topk(mfb_model *m, const int64_t *d_user_ids, int64_t n_users, const int64_t *d_train_indptr,
             const int32_t *d_train_indices, int32_t k, int32_t *d_out_ids, float *d_out_scores,
             mfb_stream stream);
/* mfb_topk for a caller that ranks the SAME (user list, train CSR) repeatedly -- model.test() runs three ranking passes
 * over one test set (implicit.py:428-460), a validation loop one per epoch.  plan_key != 0 names that pair (the caller
 * guarantees: equal key => equal d_user_ids contents, equal CSR contents); the model-independent preprocessing of the
 * train mask (its per-tile bit images for the tensor-core epilogue) is then built once and reused while the key
 * matches.  plan_key == 0 is mfb_topk.  Results never depend on the key. */
int mfb_topk_keyed(mfb_model *m, const int64_t *d_user_ids, int64_t n_users, const int64_t *d_train_indptr,
                   const int32_t *d_train_indices, int32_t k, int32_t *d_out_ids, float *d_out_scores,
                   uint64_t plan_key, mfb_stream stream);
/* The same ranking over rows of a dense score matrix [n_rows, n_items] that some other model produced -- the
 * `representation=` escape hatch of ImplicitFactorizationModel (implicit.py:169-180: MLP, NeuMF, ...), whose scores
 * come from a torch module instead of the bilinear kernel.  Descending score, ties -> lower item id, train items of
 * the row's user (d_user_ids[row]; NULL: user = row) last.  d_out_scores (may be NULL) receives the given scores
 * (-FLT_MAX for train items, as evaluation.py:163 sets them).  d_cut_scratch: n_rows * 8 bytes of device scratch,
 * required only when k > 256. */
int mfb_topk_scores(const float *d_scores, int64_t n_rows, int64_t n_items, const int64_t *d_user_ids,
                    const int64_t *d_train_indptr, const int32_t *d_train_indices, int32_t k, int32_t *d_out_ids,
                    float *d_out_scores, void *d_cut_scratch, mfb_stream stream);
/* Users that the last mfb_topk call had to redo with the exact-fp32 kernel (tensor-core path bookkeeping). */
int mfb_topk_last_redo(const mfb_model *m);
/* Test hook: raw tensor-core scores (fp16 inputs, fp32 accumulate, + item bias), item-major
 * [num_items][ceil(n_users/256)*256]; embedding_dim <= 128. */
/* Debug: candidate-list statistics of the last tensor-core mfb_topk call, h_out[5]:
 * {users, listed items total, max per user, users over capacity, items re-scored exactly}. */
int mfb_debug_tc_stats(mfb_model *m, int64_t n_users, int64_t *h_out, mfb_stream stream);
int mfb_debug_tc_scores(mfb_model *m, const int64_t *d_user_ids, int64_t n_users, float *d_out, mfb_stream stream);

/* _get_precision_recall (evaluation.py:108-113): hits[u*nk + j] = |topk[u,:ks[j]] ∩ test_row(u)|,
 * ntargets[u] = len(test_row(u)).  ks ascending, ks[nk-1] <= k. */
int mfb_topk_hits(const int32_t *d_topk_ids, const int64_t *d_user_ids, int64_t n_users, int32_t k,
                  const int64_t *d_test_indptr, const int32_t *d_test_indices, const int32_t *h_ks, int32_t nk,
                  int32_t *d_hits, int32_t *d_ntargets, mfb_stream stream);

/* ---- row-sharded training (scaled catalog: tables split over the GPUs of one box) ------------- */
/* No counterpart exists in the reference (one process, one set of tables, implicit.py:163-199); the
 * arithmetic of a step is run_train_iteration's (implicit.py:347-364), distributed as follows.  Rank r of
 * `world` OWNS the rows {g : g mod world == r} of the four tables (local index g / world) -- `local` is an
 * mfb_model bound to those local tables, moments included -- and COMPUTES a contiguous block of every
 * minibatch's positives and of its negatives (blocks as sharding.shard_range).  All ranks pass the SAME ids
 * (positives replicated, negatives drawn from the same MT19937 stream), so ids are never exchanged.
 * Exchange buffers hold rows of mfb_shard_row_stride() floats: dim values (padded to a multiple of 4), the
 * bias, 3 pad floats.  One step s of a planned chunk:
 *   mfb_shard_gather   -> all-to-all (rows to the computing ranks; counts from mfb_shard_plan)
 *   mfb_shard_forward  -> [adaptive hinge: all-reduce MAX of the int64 cell]
 *   mfb_shard_backward -> all-to-all back (gradient rows to the owners, mirrored counts)
 *   mfb_shard_update.
 * The host side (recommendation_gans_b200/sharded.py) runs the collectives with torch.distributed (NCCL). */
typedef struct mfb_shard mfb_shard;
int mfb_shard_create(mfb_model *local, int32_t rank, int32_t world, int64_t global_users, int64_t global_items,
                     mfb_shard **out);
int mfb_shard_destroy(mfb_shard *sh);
int32_t mfb_shard_row_stride(const mfb_shard *sh);
int64_t mfb_shard_launches(const mfb_shard *sh);
/* Plans steps [step0, step0+nsteps) of the epoch over the n_pos positives (minibatches as mfb_train_steps).
 * d_neg_*: the n_neg*batch negative pairs of each planned step, step-major, starting at step0's.
 * h_counts[(s*world + o)*world + c] = rows owner o sends to computing rank c in step s (and receives back as
 * gradients).  Synchronises the stream once. */
int mfb_shard_plan(mfb_shard *sh, const int64_t *d_pos_users, const int64_t *d_pos_items, int64_t n_pos,
                   int32_t batch, int32_t n_neg, const int64_t *d_neg_users, const int64_t *d_neg_items,
                   int64_t step0, int32_t nsteps, int64_t *h_counts, mfb_stream stream);
/* Owner: brings the step's requested rows up to date (dense-optimiser replay) and packs them into d_send,
 * grouped by computing rank in rank order (sum_c counts[s][rank][c] rows). */
int mfb_shard_gather(mfb_shard *sh, int32_t s, float *d_send, mfb_stream stream);
/* Computing rank: d_recv holds the rows received from owner 0, 1, ... in that order.  BilinearNet.forward
 * (representations.py:80-91) on this rank's slots; for MFB_LOSS_ADAPTIVE_HINGE *d_gmax_cell receives the
 * packed (probability, first index) maximum of this rank's negatives -- all-reduce it with MAX as int64. */
int mfb_shard_forward(mfb_shard *sh, int loss, int32_t s, const float *d_recv, int64_t *d_gmax_cell,
                      mfb_stream stream);
/* Computing rank: gradient rows (dLoss/drow contributions per slot, bias gradient in the bias position) into
 * d_gsend, same layout as d_recv; d_loss_partial[2] receives this rank's partial loss sums (positive/pairwise
 * sum, pointwise negative sum) -- the loss of the step is sum_ranks(p[0])/b (+ sum_ranks(p[1])/m). */
int mfb_shard_backward(mfb_shard *sh, int loss, int32_t s, const float *d_recv, const int64_t *d_gmax_cell,
                       float *d_gsend, double *d_loss_partial, mfb_stream stream);
/* Owner: d_grecv holds the gradient rows in d_send's layout.  Ordered segment reduction per unique row and one
 * dense-optimiser step (torch Adam / SGD) of the local model; advances its step counter. */
int mfb_shard_update(mfb_shard *sh, int32_t s, const float *d_grecv, mfb_stream stream);

/* Direct exchange over peer memory (NVLink): the collectives above are replaced by stores into the peers' exchange
 * buffers from inside the gather / forward / backward kernels plus flag kernels (release/acquire at system scope),
 * so a step needs no collective call and no host synchronisation.
 *   mfb_shard_xbuf_alloc     allocates this rank's exchange buffer for minibatches of `batch` positives and
 *                            n_neg*batch negatives (cudaMalloc: exportable with CUDA IPC)
 *   mfb_ipc_export/open/close  64-byte CUDA IPC handle of a device pointer / mapping of a peer's handle
 *   mfb_shard_xbuf_set_peers peer_ptrs[world]: every rank's exchange buffer as mapped in THIS process
 *                            (entry `rank` = the pointer mfb_shard_xbuf_alloc returned); all ranks must have
 *                            allocated before any rank runs steps
 *   mfb_shard_run_steps      steps [s_begin, s_end) of the planned chunk, fully asynchronous;
 *                            d_loss_partial: 2 doubles per step as mfb_shard_backward
 *   mfb_shard_direct_check   synchronises the stream; error if a wait on a peer timed out (~3 s) */
int mfb_shard_xbuf_alloc(mfb_shard *sh, int32_t batch, int32_t n_neg, void **d_ptr, int64_t *bytes);
int mfb_shard_xbuf_set_peers(mfb_shard *sh, void *const *peer_ptrs);
int mfb_ipc_export(void *d_ptr, void *h_handle64);
int mfb_ipc_open(const void *h_handle64, void **d_ptr);
int mfb_ipc_close(void *d_ptr);
int mfb_shard_run_steps(mfb_shard *sh, int loss, int32_t s_begin, int32_t s_end, double *d_loss_partial,
                        mfb_stream stream);
/* One phase of step s (0 owner gather | 1 forward | 2 backward | 3 owner update) with the signal at the end of a phase
 * and the wait at the head of the next as separate launches: for a host that drives several ranks of ONE device in
 * lockstep on one stream (phase p for every rank before phase p+1 for any), so that no wait is ever launched before
 * its signals.  d_loss_partial: the step's 2 doubles. */
int mfb_shard_run_phase(mfb_shard *sh, int loss, int32_t s, int32_t phase, double *d_loss_partial, mfb_stream stream);
int mfb_shard_direct_check(mfb_shard *sh, mfb_stream stream);

/* mrr_score (evaluation.py:13-60): average ranks (scipy.stats.rankdata of -predict(user), train items forced last) of
 * every listed user's test items, written at the items' positions in the test CSR: d_out_rank[p] for
 * p in [test_indptr[u], test_indptr[u+1]).  Probabilities are computed exactly as mfb_predict_user does. */
int mfb_rank_test_items(mfb_model *m, const int64_t *d_user_ids, int64_t n_users, const int64_t *d_test_indptr,
                        const int32_t *d_test_indices, const int64_t *d_train_indptr, const int32_t *d_train_indices,
                        float *d_out_rank, mfb_stream stream);

/* ---- in-situ kernel timing (measurement only) ---------------------------------------------- */
/* When enabled, every kernel launched for this model is bracketed by CUDA events on the launch
 * stream.  mfb_profile_read synchronises, then reports per kernel class the summed device time
 * (ms) and the launch count since the last reset, and resets.  Classes: see mfb_profile_name. */
#define MFB_PROFILE_CLASSES 12
int mfb_profile_enable(mfb_model *m, int on);
int mfb_profile_read(mfb_model *m, double *h_ms, int64_t *h_launches);
const char *mfb_profile_name(int cls);
/* total kernels launched by the library for this model since creation (always counted) */
int64_t mfb_model_launches(const mfb_model *m);
/* kernels launched by the model-less entry points (MT19937 streams, loss functions, hit counting) */
int64_t mfb_library_launches(void);

#ifdef __cplusplus
}
#endif
#endif /* MFB200_H */
