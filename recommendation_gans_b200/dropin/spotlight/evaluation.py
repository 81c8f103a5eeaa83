"""Ranking evaluation (reference: spotlight/evaluation.py:108-190).

`precision_recall_score` keeps the reference's signature and return value (two scalars: the mean
over evaluated users and over all requested k) but replaces the per-user
predict -> mask -> argsort -> set-intersection loop with two CUDA kernels: fused scoring + train mask
+ top-k (mfb_topk; no score matrix is materialised) and a hit counter (mfb_topk_hits).
Ranking is on the pre-sigmoid score with ties -> lower item id (the reference's own tie order is
whatever numpy's unstable argsort produces).
"""
import itertools

import numpy as np
import torch

FLOAT_MAX = np.finfo(np.float32).max
_TOKENS = itertools.count(1)      # names of uploaded CSR contents (plan keys of mfb_topk_keyed)


class _ModuleRanker(object):
    """Ranking backend for a model whose scores come from an arbitrary torch module (`representation=` MLP / NeuMF,
    implicit.py:169-180): the module scores blocks of users against the whole catalogue and the CUDA ranking kernel
    (mfb_topk_scores) keeps the top-k with the train mask -- the score matrix lives only one block at a time."""

    def __init__(self, model):
        self.net = model._net
        self.device = next(self.net.parameters()).device
        self.num_users, self.num_items = int(model._num_users), int(model._num_items)

    def topk(self, user_ids, k, train_indptr=None, train_indices=None, with_scores=False, plan_key=0):
        from recommendation_gans_b200.engine import topk_scores_device
        user_ids = torch.as_tensor(user_ids, dtype=torch.int64, device=self.device).reshape(-1)
        n, I = user_ids.numel(), self.num_items
        ids = torch.empty((n, int(k)), dtype=torch.int32, device=self.device)
        scores = torch.empty((n, int(k)), dtype=torch.float32, device=self.device) if with_scores else None
        block = max(1, min(4096, (1 << 24) // I))
        items = torch.arange(I, device=self.device, dtype=torch.int64)
        was_training = self.net.training
        self.net.train(False)
        with torch.no_grad():
            for lo in range(0, n, block):
                u = user_ids[lo:lo + block]
                s = self.net(u.repeat_interleave(I), items.repeat(u.numel())).reshape(u.numel(), I)
                out = topk_scores_device(s, k, u, train_indptr, train_indices, with_scores=with_scores)
                if with_scores:
                    ids[lo:lo + block], scores[lo:lo + block] = out
                else:
                    ids[lo:lo + block] = out
        self.net.train(was_training)
        return (ids, scores) if with_scores else ids

    def topk_hits(self, topk_ids, user_ids, test_indptr, test_indices, ks):
        from recommendation_gans_b200.engine import topk_hits_device
        return topk_hits_device(topk_ids, user_ids, test_indptr, test_indices, ks)

    def rank_test_items(self, *args, **kwargs):
        raise NotImplementedError('mrr_score: average ranks are implemented for the BilinearNet representation only')


def _native_engine(model):
    """The ranking backend of a fitted model: the fused engine for BilinearNet, the module ranker for anything else."""
    net = getattr(model, '_net', None)
    if net is None:
        raise ValueError('evaluation: the model has no network (call fit or set_users first)')
    if not hasattr(net, '_engine'):
        return _ModuleRanker(model)
    eng = getattr(model, '_engine_', None)
    return eng if eng is not None else net._engine()


def _host_csr(interactions):
    """The CSR form of an Interactions object (or a scipy matrix).  Interactions builds it once at construction
    (interactions.py:115, `csr_matrix`); that cached matrix is used when it still describes the id arrays."""
    cached = getattr(interactions, 'csr_matrix', None)
    if cached is not None and hasattr(interactions, 'user_ids') and cached.shape == (interactions.num_users,
                                                                                    interactions.num_items):
        return cached
    return interactions.tocsr()


def _csr_to_device(csr, device):
    """(indptr int64, indices int32) of a CSR matrix on `device`; the copy is kept on the matrix object so the
    several metric passes of model.test() (implicit.py:428-460) upload a set of interactions once."""
    csr = csr.tocsr()
    cache = getattr(csr, '_mfb_device', None)
    if cache is not None and cache[0] == str(device) and cache[1] == csr.nnz:
        return cache[2], cache[3]
    if not csr.has_canonical_format:
        csr.sum_duplicates()
    if not csr.has_sorted_indices:
        csr.sort_indices()
    indptr = torch.from_numpy(np.ascontiguousarray(csr.indptr, dtype=np.int64)).to(device)
    indices = torch.from_numpy(np.ascontiguousarray(csr.indices, dtype=np.int32)).to(device)
    try:
        csr._mfb_device = (str(device), csr.nnz, indptr, indices, next(_TOKENS))
    except AttributeError:
        pass
    return indptr, indices


def _plan_key(test_csr, train_csr):
    """Key of the (evaluated users, train mask) pair for mfb_topk_keyed: both are functions of the uploaded CSR
    contents, which `_csr_to_device` names with a token; 0 (no reuse) when either is not cached."""
    a, b = getattr(test_csr, '_mfb_device', None), getattr(train_csr, '_mfb_device', None)
    if a is None or b is None or len(a) < 5 or len(b) < 5:
        return 0
    return ((b[4] & 0xFFFFFFFF) << 32) | (a[4] & 0xFFFFFFFF)


def _check_against_model(eng, test_csr, train_csr=None):
    """What model.predict's _check_input (implicit.py:215-236) enforces inside the reference's per-user loop: ids beyond
    the model's tables raise ValueError instead of reading out of bounds."""
    if test_csr.shape[0] > eng.num_users and np.diff(test_csr.indptr)[eng.num_users:].any():
        raise ValueError('Maximum user id greater than number of users in model.')
    for name, csr in (('test', test_csr), ('train', train_csr)):
        if csr is None:
            continue
        top = getattr(csr, '_mfb_max_item', None)          # one pass over the indices per matrix object
        if top is None or top[0] != csr.nnz:
            top = (csr.nnz, int(csr.indices.max()) if csr.nnz else -1)
            try:
                csr._mfb_max_item = top
            except AttributeError:
                pass
        if top[1] >= eng.num_items:
            raise ValueError('Maximum item id greater than number of items in model.')
    if train_csr is not None and train_csr.shape[0] < min(test_csr.shape[0], eng.num_users):
        # the reference indexes train[user_id] for every evaluated user (evaluation.py:162)
        raise IndexError('row index (%d) out of range' % train_csr.shape[0])


def _eval_users(eng, test_csr):
    """Users with at least one test interaction (evaluation.py:157), as ids into the model's tables."""
    row_len = np.diff(test_csr.indptr)
    return np.nonzero(row_len[:eng.num_users])[0].astype(np.int64)


def _get_precision_recall(predictions, targets, k):
    """Host mirror of evaluation.py:108-113 (kept for API parity; the GPU path uses mfb_topk_hits)."""
    predictions = predictions[:k]
    num_hit = len(set(predictions).intersection(set(targets)))
    return float(num_hit) / k, float(num_hit) / len(targets)


def _hits_at(eng, topk, d_users, t_indptr, t_indices, cutoffs):
    """hits[u, j] = |top-cutoffs[j] of user u  intersected with its test items| and the users' target counts.  A cut-off
    beyond the catalogue size sees the whole catalogue (the reference slices predictions[:k] of a num_items-long
    ranking, evaluation.py:110); the kernel takes up to 4 ascending cut-offs per call."""
    width = topk.shape[1]
    clamped = np.minimum(np.asarray(cutoffs, dtype=np.int64), width)
    uniq = np.unique(clamped)
    hits = np.zeros((topk.shape[0], len(uniq)), dtype=np.int64)
    ntargets = None
    for c0 in range(0, len(uniq), 4):
        chunk = uniq[c0:c0 + 4]
        h, nt = eng.topk_hits(topk, d_users, t_indptr, t_indices, chunk)
        hits[:, c0:c0 + len(chunk)] = h.cpu().numpy()
        ntargets = nt.cpu().numpy()
    col = {int(kk): j for j, kk in enumerate(uniq)}
    return hits[:, [col[int(c)] for c in clamped]], ntargets


def topk_for_users(model, user_ids, k, train=None):
    """Top-k item ids (best first) for each listed user; train items rank last when `train` is given."""
    eng = _native_engine(model)
    indptr = indices = None
    if train is not None:
        train_csr = _host_csr(train)
        user_arr = np.asarray(user_ids, dtype=np.int64).reshape(-1)
        if len(user_arr) and (int(user_arr.max()) >= min(eng.num_users, train_csr.shape[0]) or int(user_arr.min()) < 0):
            raise ValueError('Maximum user id greater than number of users in model.')
        if train_csr.nnz and int(train_csr.indices.max()) >= eng.num_items:
            raise ValueError('Maximum item id greater than number of items in model.')
        indptr, indices = _csr_to_device(train_csr, eng.device)
    return eng.topk(np.asarray(user_ids, dtype=np.int64), int(k), indptr, indices)


def precision_recall_score(model, test, train=None, k=10):
    eng = _native_engine(model)
    test_csr = _host_csr(test)
    train_csr = _host_csr(train) if train is not None else None
    _check_against_model(eng, test_csr, train_csr)
    ks = np.array([k]) if np.isscalar(k) else np.asarray(k)
    ks_sorted = np.unique(ks)
    kmax = int(ks_sorted[-1])
    user_ids = _eval_users(eng, test_csr)                    # users with >= 1 test item (evaluation.py:157)
    cold_start_users = 0
    if train_csr is not None:
        cold_start_users = int((np.diff(train_csr.indptr)[user_ids] == 0).sum())
    if len(user_ids) == 0:
        print("Cold start users: ", cold_start_users)
        return np.mean(np.array([])), np.mean(np.array([]))
    t_indptr, t_indices = _csr_to_device(test_csr, eng.device)
    m_indptr = m_indices = None
    if train_csr is not None:
        m_indptr, m_indices = _csr_to_device(train_csr, eng.device)
    d_users = torch.from_numpy(user_ids).to(eng.device)
    topk = eng.topk(d_users, min(kmax, eng.num_items), m_indptr, m_indices,
                    plan_key=_plan_key(test_csr, train_csr) if train_csr is not None else 0)
    hits, ntargets = _hits_at(eng, topk, d_users, t_indptr, t_indices, ks)
    precision = hits.astype(np.float64) / ks.astype(np.float64)[None, :]
    recall = hits.astype(np.float64) / ntargets.astype(np.float64)[:, None]
    print("Cold start users: ", cold_start_users)
    return np.mean(precision.squeeze()), np.mean(recall.squeeze())


def precision_recall_score_sharded(model, test, train=None, k=10):
    """Multi-GPU variant (one process per GPU, torch.distributed initialised, the same model replicated on
    every rank): the users with test items are split into contiguous blocks, each rank scores its block on its
    own GPU and the per-k sums are all-reduced.  Returns the same two scalars as precision_recall_score."""
    import torch.distributed as dist
    from recommendation_gans_b200.sharding import allreduce_precision_recall, shard_range
    eng = _native_engine(model)
    test_csr = _host_csr(test)
    train_csr = _host_csr(train) if train is not None else None
    _check_against_model(eng, test_csr, train_csr)
    ks = np.array([k]) if np.isscalar(k) else np.asarray(k)
    ks_sorted = np.unique(ks)
    all_users = _eval_users(eng, test_csr)
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(len(all_users), rank, world)
    user_ids = all_users[lo:hi]
    t_indptr, t_indices = _csr_to_device(test_csr, eng.device)
    m_indptr = m_indices = None
    if train_csr is not None:
        m_indptr, m_indices = _csr_to_device(train_csr, eng.device)
    hits = np.zeros((len(user_ids), len(ks)), dtype=np.int64)
    ntargets = np.zeros(len(user_ids), dtype=np.int64)
    if len(user_ids):
        d_users = torch.from_numpy(user_ids).to(eng.device)
        key = _plan_key(test_csr, train_csr) if train_csr is not None else 0
        topk = eng.topk(d_users, min(int(ks_sorted[-1]), eng.num_items), m_indptr, m_indices,
                        plan_key=(key ^ ((rank + 1) << 20) ^ (world << 28)) if key else 0)
        hits, ntargets = _hits_at(eng, topk, d_users, t_indptr, t_indices, ks)
    prec, rec, _ = allreduce_precision_recall(hits, ntargets, ks, dist=dist if world > 1 else None,
                                              device=eng.device if (world > 1 and dist.get_backend() == 'nccl') else None)
    return float(np.mean(prec)), float(np.mean(rec))


def rmse_score(net, user_ids, item_ids):
    """Sum of squared (1 - prediction) over the batch (evaluation.py:187-190; logged as "BCE")."""
    predictions = net(user_ids, item_ids)
    diff = 1.0 - predictions.detach().cpu().numpy()
    return np.sum(diff ** 2)


def _all_rank_hits(eng, topk, d_users, t_indptr, t_indices, k):
    """cum[u, r] = number of test items of user u among its first r+1 recommendations."""
    cum = np.zeros((topk.shape[0], k), dtype=np.int64)
    ntargets = None
    for c0 in range(0, k, 4):
        chunk = np.arange(c0 + 1, min(c0 + 4, k) + 1)
        h, nt = eng.topk_hits(topk, d_users, t_indptr, t_indices, chunk)
        cum[:, c0:c0 + len(chunk)] = h.cpu().numpy()
        ntargets = nt.cpu().numpy()
    return cum, ntargets


def map_at_k(model, test, k=5):
    """Mean average precision@k over users with test items (evaluation.py:278-353): no train mask,
    apk normalised by min(len(targets), k); a user whose only target is item 0 scores 0.0, as in the
    reference's `if not actual.any()` check (evaluation.py:308-309)."""
    eng = _native_engine(model)
    test_csr = _host_csr(test)
    _check_against_model(eng, test_csr)
    user_ids = _eval_users(eng, test_csr)
    if len(user_ids) == 0:
        return np.mean(np.array([]))
    t_indptr, t_indices = _csr_to_device(test_csr, eng.device)
    d_users = torch.from_numpy(user_ids).to(eng.device)
    kk = min(int(k), eng.num_items)
    topk = eng.topk(d_users, kk)
    cum, ntargets = _all_rank_hits(eng, topk, d_users, t_indptr, t_indices, kk)
    flags = np.diff(np.concatenate([np.zeros((len(user_ids), 1), dtype=np.int64), cum], axis=1), axis=1)
    score = (flags * cum / np.arange(1, kk + 1)[None, :]).sum(axis=1)
    apk_ = score / np.minimum(ntargets, k)
    csr = test_csr
    all_zero_targets = np.array([not csr.indices[csr.indptr[u]:csr.indptr[u + 1]].any() for u in user_ids])
    apk_[all_zero_targets] = 0.0
    return np.mean(apk_)


def mrr_score(model, test, train=None):
    """Mean reciprocal rank per user with test items (evaluation.py:13-60): the rank of every test item in
    -model.predict(user) with the user's train items forced last, average ranks for ties (scipy.stats.rankdata).  One
    kernel compares every item's probability with the user's test-item probabilities; no score vector, no sort."""
    eng = _native_engine(model)
    test_csr = _host_csr(test)
    _check_against_model(eng, test_csr)
    user_ids = _eval_users(eng, test_csr)
    if len(user_ids) == 0:
        return np.array([])
    t_indptr, t_indices = _csr_to_device(test_csr, eng.device)
    r_indptr = r_indices = None
    if train is not None:
        train_csr = _host_csr(train)
        _check_against_model(eng, test_csr, train_csr)
        r_indptr, r_indices = _csr_to_device(train_csr, eng.device)
    ranks = eng.rank_test_items(torch.from_numpy(user_ids).to(eng.device), t_indptr, t_indices, r_indptr,
                                r_indices).cpu().numpy().astype(np.float64)
    indptr = t_indptr.cpu().numpy()
    return np.array([(1.0 / ranks[indptr[u]:indptr[u + 1]]).mean() for u in user_ids])


def hit_ratio(model, test, k=10):
    """Fraction of users with test items whose target is among the top-k recommendations (evaluation.py:192-213; no
    train mask).  The reference evaluates `target in predictions[:k]` on numpy arrays: with one test item per user
    (leave-one-out, its intended use) that is membership; with exactly k targets numpy compares position by position;
    any other count makes numpy raise ValueError -- reproduced here."""
    eng = _native_engine(model)
    test_csr = _host_csr(test)
    _check_against_model(eng, test_csr)
    user_ids = _eval_users(eng, test_csr)
    if len(user_ids) == 0:
        raise ZeroDivisionError('division by zero')            # num_hits / num_users with no test users
    kk = min(int(k), eng.num_items)
    counts = np.diff(test_csr.indptr)[user_ids]
    bad = (counts != 1) & (counts != kk)
    if bad.any():
        raise ValueError('operands could not be broadcast together with shapes (%d,) (%d,) '
                         % (kk, counts[np.argmax(bad)]))
    d_users = torch.from_numpy(user_ids).to(eng.device)
    topk = eng.topk(d_users, kk)
    t_indptr, t_indices = _csr_to_device(test_csr, eng.device)
    hits, _ = eng.topk_hits(topk, d_users, t_indptr, t_indices, np.array([kk]))
    hit = hits.cpu().numpy().reshape(-1) > 0
    multi = np.nonzero(counts == kk)[0] if kk != 1 else np.array([], dtype=np.int64)
    if len(multi):                                              # k targets: numpy's elementwise comparison
        top_h = topk.cpu().numpy()
        for r in multi:
            u = user_ids[r]
            hit[r] = bool((top_h[r] == test_csr.indices[test_csr.indptr[u]:test_csr.indptr[u + 1]]).any())
    return hit.sum() / len(user_ids)


def evaluate_popItems(item_popularity, test, k=10):
    """Most-popular-items baseline (evaluation.py:215-243).  Host-side by nature: no model involved."""
    test_csr = test.tocsr()
    pop_top = item_popularity.values.argsort()[::-1][:k]
    ks = np.array([k]) if np.isscalar(k) else np.asarray(k)
    precision, recall = [], []
    for u in np.nonzero(np.diff(test_csr.indptr))[0]:
        targets = test_csr.indices[test_csr.indptr[u]:test_csr.indptr[u + 1]]
        p, r = zip(*[_get_precision_recall(pop_top, targets, x) for x in ks])
        precision.append(p)
        recall.append(r)
    return np.mean(precision), np.mean(recall),


def evaluate_random(item_popularity, test, k=10):
    """Random-recommendation baseline (evaluation.py:245-276); draws from the global numpy RNG."""
    all_items = test.num_items
    test_csr = test.tocsr()
    ks = np.array([k]) if np.isscalar(k) else np.asarray(k)
    precision, recall = [], []
    for u in np.nonzero(np.diff(test_csr.indptr))[0]:
        targets = test_csr.indices[test_csr.indptr[u]:test_csr.indptr[u + 1]]
        predictions = np.random.choice(all_items, len(targets))
        p, r = zip(*[_get_precision_recall(predictions, targets, x) for x in ks])
        precision.append(p)
        recall.append(r)
    return np.mean(np.array(precision).squeeze()), np.mean(np.array(recall).squeeze())
