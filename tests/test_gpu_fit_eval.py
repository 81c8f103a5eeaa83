"""Drop-in API level: ImplicitFactorizationModel.fit / predict / precision_recall_score on the GPU vs
golden results of the reference's own fit (same seeds, same inputs), and top-k id exactness of the
evaluation kernel vs the tie-defined oracle ranking."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import mf_oracle as O
from tests.gpu_helpers import make_net, tables_of, rel_err

pytestmark = pytest.mark.gpu


def _interactions(u, i, U, I):
    from spotlight.interactions import Interactions
    return Interactions(u.astype(np.int32), i.astype(np.int32), num_users=U, num_items=I)


@pytest.mark.parametrize('name', ['fit_pointwise', 'fit_bpr'])
def test_fit_predict_evaluate_match_reference(golden_dir, name, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)                               # the model creates experiments_results/ in cwd
    from implicit import ImplicitFactorizationModel
    from spotlight.evaluation import precision_recall_score
    import spotlight.optimizers as optimizers
    g = np.load(os.path.join(golden_dir, name + '.npz'))
    U, I, D, B, n_neg, n_epochs = [int(x) for x in g['meta']]
    lr, l2 = [float(x) for x in g['hyper']]
    a, b = [int(x) for x in g['split']]
    users, items = g['users'], g['items']
    train, valid = _interactions(users[:a], items[:a], U, I), _interactions(users[a:b], items[a:b], U, I)
    test = _interactions(users[b:], items[b:], U, I)
    net = make_net([g['init%d' % i] for i in range(4)])
    model = ImplicitFactorizationModel(
        loss=str(g['loss']), embedding_dim=D, n_iter=n_epochs, batch_size=B, l2=l2, learning_rate=lr,
        optimizer_func=optimizers.adam_optimizer, representation=net, random_state=np.random.RandomState(0),
        neg_examples=[tuple(p) for p in g['neg_pairs'].tolist()], num_negative_samples=n_neg,
        experiment_name='gpu_' + name, use_cuda=True)
    random.seed(int(g['py_seed']))
    model.fit(train, valid, verbose=False)
    # per-epoch summary.csv written like the reference's
    import csv
    with open(os.path.join(model.experiment_logs, 'summary.csv')) as f:
        rows = list(csv.reader(f))
    assert rows[0] == list(g['summary_header'])
    summary = np.array([[float(x) for x in r] for r in rows[1:]])
    np.testing.assert_allclose(summary[:, :2], g['summary'][:, :2], rtol=1e-5)
    assert model.best_epoch == int(g['best_epoch'])
    # python's global random stream ends where the reference leaves it
    assert tuple(int(x) for x in g['py_random_after']) == random.getstate()[1]
    for i, t in enumerate(tables_of(model._net)):
        assert rel_err(t, g['final%d' % i]) < 1e-5, i
    ck = torch.load(os.path.join(model.experiment_saved_models, 'best_model'))
    assert sorted(ck['network']) == ['item_biases.weight', 'item_embeddings.weight', 'user_biases.weight',
                                     'user_embeddings.weight']
    np.testing.assert_allclose(model.predict(3), g['predict_user3'], rtol=1e-5)
    np.testing.assert_allclose(model.predict(g['predict_pairs_users'], g['predict_pairs_items']),
                               g['predict_pairs'], rtol=1e-5)
    with pytest.raises(ValueError):
        model.predict(U)
    for k in (5, 10, 20):
        p, r = precision_recall_score(model, test, train=train, k=k)
        np.testing.assert_allclose([p, r], g['pr_masked_k%d' % k], atol=2e-3)   # tie order is the only freedom
        p, r = precision_recall_score(model, test, k=k)
        np.testing.assert_allclose([p, r], g['pr_nomask_k%d' % k], atol=2e-3)
    p, r = precision_recall_score(model, test, train=train, k=np.array([5, 10, 20]))
    np.testing.assert_allclose([p, r], g['pr_masked_karray'], atol=2e-3)
    res = model.test(test, None, k=5, rmse_flag=True, precision_recall=False, map_recall=True)
    assert set(res) >= {'k', 'bce', 'map'} and os.path.exists(os.path.join(model.experiment_logs, 'test_summary.json'))


def _dyadic_tables(rs, U, I, D):
    """Embeddings on a coarse dyadic grid: every dot product is exact in fp32 (and bf16/tf32) in any
    summation order, so top-k ids are well defined bit-for-bit."""
    ue = rs.randint(-8, 9, (U, D)).astype(np.float32) / 16.0
    ie = rs.randint(-8, 9, (I, D)).astype(np.float32) / 16.0
    ub = rs.randint(-4, 5, (U, 1)).astype(np.float32) / 8.0
    ib = rs.randint(-4, 5, (I, 1)).astype(np.float32) / 8.0
    return ue, ie, ub, ib


@pytest.mark.parametrize('U,I,D,k', [(70, 333, 32, 20), (33, 1000, 128, 10), (5, 40, 16, 32), (130, 257, 50, 5)])
def test_topk_ids_bit_exact_on_dyadic_grid(U, I, D, k):
    from recommendation_gans_b200.engine import MFEngine
    rs = np.random.RandomState(U + I)
    tabs = _dyadic_tables(rs, U, I, D)
    net = make_net(tabs)
    eng = MFEngine(net)
    tu, ti = rs.randint(0, U, 12 * U), rs.randint(0, I, 12 * U)
    tu[tu == 2] = 3                                             # user 2: cold start (no train row)
    train = O.csr_from_pairs(tu, ti, U, I)
    train.sort_indices()
    indptr = torch.from_numpy(train.indptr.astype(np.int64)).cuda()
    indices = torch.from_numpy(train.indices.astype(np.int32)).cuda()
    user_ids = np.arange(U, dtype=np.int64)[::-1].copy()       # arbitrary order
    model = O.OracleMF(*[torch.from_numpy(t) for t in tabs])
    for masked in (True, False):
        got = eng.topk(user_ids, k, indptr if masked else None, indices if masked else None).cpu().numpy()
        for row, u in enumerate(user_ids):
            rated = train[u].indices if masked else None
            exp = O.topk_stable(model.logits(int(u)), rated, k)
            assert (got[row] == exp).all(), (u, masked, got[row], exp)


def test_topk_more_masked_than_free_items():
    """k larger than the number of unmasked items: masked items follow, by ascending id."""
    from recommendation_gans_b200.engine import MFEngine
    rs = np.random.RandomState(3)
    U, I, D, k = 4, 24, 8, 20
    tabs = _dyadic_tables(rs, U, I, D)
    eng = MFEngine(make_net(tabs))
    tu = np.repeat(np.arange(U), 15)
    ti = np.concatenate([rs.permutation(I)[:15] for _ in range(U)])
    train = O.csr_from_pairs(tu, ti, U, I)
    train.sort_indices()
    got = eng.topk(np.arange(U), k, torch.from_numpy(train.indptr.astype(np.int64)).cuda(),
                   torch.from_numpy(train.indices.astype(np.int32)).cuda()).cpu().numpy()
    model = O.OracleMF(*[torch.from_numpy(t) for t in tabs])
    for u in range(U):
        assert (got[u] == O.topk_stable(model.logits(u), train[u].indices, k)).all()


def test_topk_random_fp32_respects_gaps():
    """Random fp32 embeddings: ids must agree with the oracle wherever the oracle's neighbouring scores
    differ by more than the fp32 summation-order error bound."""
    from recommendation_gans_b200.engine import MFEngine
    rs = np.random.RandomState(5)
    U, I, D, k = 64, 3000, 64, 20
    tabs = (rs.normal(0, 0.3, (U, D)).astype(np.float32), rs.normal(0, 0.3, (I, D)).astype(np.float32),
            rs.normal(0, 0.1, (U, 1)).astype(np.float32), rs.normal(0, 0.1, (I, 1)).astype(np.float32))
    eng = MFEngine(make_net(tabs))
    got, scores = eng.topk(np.arange(U), k, with_scores=True)
    got, scores = got.cpu().numpy(), scores.cpu().numpy()
    model = O.OracleMF(*[torch.from_numpy(t) for t in tabs])
    for u in range(U):
        z = model.logits(u).astype(np.float64)
        order = np.argsort(-z, kind='stable')
        bound = 1e-5
        for r in range(k):
            gap_ok = (z[order[r]] - z[order[r + 1]] > bound) and (r == 0 or z[order[r - 1]] - z[order[r]] > bound)
            if gap_ok:
                assert got[u, r] == order[r]
        np.testing.assert_allclose(scores[u], 1 / (1 + np.exp(-z[got[u]])), rtol=1e-5)


# ------------------------------------------------------------------------------------------------
# tensor-core evaluation path (TMA + tcgen05 GEMM -> candidates -> exact fp32 re-score)
# ------------------------------------------------------------------------------------------------
def _random_tables(rs, U, I, D, scale=0.3):
    return (rs.normal(0, scale, (U, D)).astype(np.float32), rs.normal(0, scale, (I, D)).astype(np.float32),
            rs.normal(0, 0.1, (U, 1)).astype(np.float32), rs.normal(0, 0.1, (I, 1)).astype(np.float32))


@pytest.mark.parametrize('U,I,D', [(300, 1500, 128), (70, 1100, 64), (513, 2049, 128), (90, 1300, 32), (257, 1100, 50)])
def test_tc_raw_scores_match_fp16_matmul(U, I, D):
    """The tcgen05 GEMM (descriptors, swizzle, TMEM layout, the extra K = 16 step that carries the bias) against torch on
    fp16-rounded inputs."""
    from recommendation_gans_b200.engine import MFEngine
    rs = np.random.RandomState(U)
    tabs = _random_tables(rs, U, I, D)
    eng = MFEngine(make_net(tabs))
    users = np.arange(U, dtype=np.int64)[::-1].copy()
    got = eng.debug_tc_scores(users).cpu().numpy()                       # [I, U]
    ub = torch.from_numpy(tabs[0][users]).cuda().half().float()
    vb = torch.from_numpy(tabs[1]).cuda().half().float()
    ref = (vb.double() @ ub.double().T).float().cpu().numpy() + tabs[3]   # + item bias (fp32)
    np.testing.assert_allclose(got, ref, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize('bias_scale', [1.0, 50.0, 3000.0])
def test_tc_bias_through_the_extra_mma_step_stays_inside_its_error_budget(bias_scale):
    """The item bias enters the accumulator as three fp16 pieces times (1, 2^-6, 2^-12) in an extra K = 16 MMA step,
    and the dot product is then accumulated ON TOP of it.  The error model (k_tc_xk_users) budgets 2^-16 * |b| for what
    that costs; here biases far larger than the dot products make the effect visible: the raw tensor-core scores stay
    within 2^-16 * |b|max + (the fp16-matmul tolerance of the test above) of the float64 value."""
    from recommendation_gans_b200.engine import MFEngine
    U, I, D = 300, 1500, 128
    rs = np.random.RandomState(11)
    tabs = list(_random_tables(rs, U, I, D, scale=0.05))
    tabs[3] = (rs.normal(0, bias_scale, (I, 1))).astype(np.float32)
    eng = MFEngine(make_net(tabs))
    users = np.arange(U, dtype=np.int64)
    got = eng.debug_tc_scores(users).cpu().numpy().astype(np.float64)    # [I, U]
    ub = torch.from_numpy(tabs[0][users]).cuda().half().double()
    vb = torch.from_numpy(tabs[1]).cuda().half().double()
    ref = (vb @ ub.T).cpu().numpy() + tabs[3].astype(np.float64)
    err = np.abs(got - ref)
    bmax = float(np.abs(tabs[3]).max())
    budget = 2.0 ** -16 * bmax + 2e-5
    assert err.max() <= budget, (err.max(), budget)
    # and the budget is not vacuous: the error is within a factor of ~50 of it for the large biases
    if bias_scale >= 50.0:
        assert err.max() >= 2.0 ** -24 * bmax * 0.5, (err.max(), bmax)


@pytest.mark.parametrize('U,I,D,k,scale,skew', [(600, 12000, 128, 20, 0.3, False), (300, 9000, 64, 5, 0.05, False),
                                                (1000, 20000, 128, 10, 1.0, False), (900, 16000, 128, 20, 0.01, True),
                                                (500, 9000, 32, 10, 0.2, False), (400, 8000, 50, 20, 0.1, True),
                                                (300, 7000, 100, 5, 0.3, False)])
@pytest.mark.parametrize('split', [None, '1', '3'], ids=['auto_split', 'split1', 'split3'])
def test_tc_topk_equals_exact_kernel(U, I, D, k, scale, skew, split, monkeypatch):
    """Tensor-core path returns bit-identical ids to the exact fp32 kernel (same exact re-score definition).
    split: item tiles of a user block divided over grid.y CTAs with their own candidate sub-lists (auto = what a small
    user shard gets; 1 = the full-size configuration).
    skew=True: heavy-tailed item norms and a shared popular direction, the shape a trained model has, with popularity
    correlated with the item id -- the item layout and per-item error radii must keep the candidate lists short."""
    from recommendation_gans_b200.engine import MFEngine
    rs = np.random.RandomState(I)
    tabs = list(_random_tables(rs, U, I, D, scale))
    if skew:
        pop = (1.0 + np.arange(I)) ** -0.6                          # popularity falls with the item id (as in ML-20M)
        common = rs.normal(0, 1, D).astype(np.float32)
        common /= np.linalg.norm(common)
        tabs[1] = (tabs[1] * (1 + 40 * pop[:, None]) + 8 * scale * np.sqrt(D) * pop[:, None] * common).astype(np.float32)
        tabs[0] = (tabs[0] + 2 * scale * np.sqrt(D) * rs.rand(U, 1).astype(np.float32) * common).astype(np.float32)
        tabs[3] = (tabs[3] + 4 * scale * pop[:, None]).astype(np.float32)
    tu, ti = rs.randint(0, U, 12 * U), rs.randint(0, I, 12 * U)
    tu[tu == 7] = 8                                                # a cold-start user
    train = O.csr_from_pairs(tu, ti, U, I)
    train.sort_indices()
    indptr = torch.from_numpy(train.indptr.astype(np.int64)).cuda()
    indices = torch.from_numpy(train.indices.astype(np.int32)).cuda()
    users = rs.permutation(U).astype(np.int64)
    if split is not None:
        monkeypatch.setenv('MFB_TC_SPLIT', split)
    monkeypatch.setenv('MFB_TC', '0')
    exact = MFEngine(make_net(tabs))
    monkeypatch.setenv('MFB_TC', '1')
    tc = MFEngine(make_net(tabs))
    for masked in (True, False):
        args = (indptr, indices) if masked else (None, None)
        ids_e, sc_e = exact.topk(users, k, *args, with_scores=True)
        ids_t, sc_t = tc.topk(users, k, *args, with_scores=True)
        assert (ids_e.cpu().numpy() == ids_t.cpu().numpy()).all()
        np.testing.assert_array_equal(sc_e.cpu().numpy(), sc_t.cpu().numpy())
        assert tc.topk_last_redo < U // 20, tc.topk_last_redo       # the fast path did the work
    assert exact.topk_last_redo == 0


def test_hit_ratio_matches_reference_loop():
    """evaluation.hit_ratio (evaluation.py:192-213) on a leave-one-out test set: membership of the single target in the
    top-k of the stable ranking (dyadic tables: every score exact, so the ranking is well defined)."""
    from spotlight.evaluation import hit_ratio
    from oracle import mf_oracle as O
    rs = np.random.RandomState(3)
    U, I, D, k = 300, 500, 16, 10
    tabs = _dyadic_tables(rs, U, I, D)
    net = make_net(tabs)

    class _Model(object):          # the only thing evaluation needs from a fitted model
        _net = net
        _num_items = I
    users = np.arange(0, U, 2)     # every second user has one test item
    targets = rs.randint(0, I, len(users))
    test = _interactions(users, targets, U, I)
    logits = tabs[0] @ tabs[1].T + tabs[2] + tabs[3].T
    expect = np.mean([targets[j] in O.topk_stable(logits[u], np.array([], dtype=np.int64), k)
                      for j, u in enumerate(users)])
    assert hit_ratio(_Model(), test, k=k) == pytest.approx(expect, abs=1e-12)
    two = _interactions(np.repeat(users[:5], 2), rs.randint(0, I, 10), U, I)      # 2 targets, k=10: numpy raises
    with pytest.raises(ValueError):
        hit_ratio(_Model(), two, k=k)


@pytest.mark.parametrize('with_train', [False, True], ids=['nomask', 'trainmask'])
def test_mrr_score_matches_reference_loop(with_train):
    """evaluation.mrr_score (evaluation.py:13-60) vs the reference's loop restated on the oracle's predictions:
    -predict(user), train items set to FLOAT_MAX, scipy.stats.rankdata, mean reciprocal rank of the test items."""
    import scipy.stats as st
    from spotlight.evaluation import mrr_score
    from oracle import mf_oracle as O
    rs = np.random.RandomState(8)
    U, I, D = 120, 700, 32
    tabs = _random_tables(rs, U, I, D, 0.5)
    tabs[1][5] = tabs[1][9]                                   # two items with identical rows and biases: an exact tie
    tabs[3][5] = tabs[3][9]
    net = make_net(tabs)

    class _Model(object):
        _net = net
        _num_items = I
    tu, ti = rs.randint(0, U, 600), rs.randint(0, I, 600)
    tu[tu == 3] = 4                                           # a user without test items
    ti[:40] = rs.choice([5, 9], 40)                           # the tied items are test items of several users
    test = _interactions(tu, ti, U, I)
    ru, ri = rs.randint(0, U, 3000), rs.randint(0, I, 3000)
    ru[:30], ri[:30] = tu[:30], ti[:30]                       # some test items are also train items
    train = _interactions(ru, ri, U, I) if with_train else None
    oracle = O.OracleMF(*[torch.from_numpy(t) for t in tabs])
    test_csr = test.tocsr()
    train_csr = train.tocsr() if with_train else None
    expect = []
    for u in range(U):
        row = test_csr[u].indices
        if not len(row):
            continue
        pred = -oracle.predict(u)
        if with_train:
            pred[train_csr[u].indices] = np.finfo(np.float32).max
        expect.append((1.0 / st.rankdata(pred)[row]).mean())
    got = mrr_score(_Model(), test, train=train)
    assert got.shape == (len(expect),)
    np.testing.assert_allclose(got, expect, rtol=1e-4)


@pytest.mark.parametrize('name', ['fit_pointwise', 'fit_bpr'])
def test_map_at_k_and_bce_match_reference(golden_dir, name, tmp_path, monkeypatch):
    """map_at_k (evaluation.py:334-353) and the "BCE" figure of model.test (rmse_score, evaluation.py:187-190,
    implicit.py:428-437) on the reference's fitted tables vs values frozen from the reference's own functions
    (tests/golden/metrics.npz, oracle/make_golden_metrics.py)."""
    monkeypatch.chdir(tmp_path)
    from implicit import ImplicitFactorizationModel
    from spotlight.evaluation import map_at_k
    g = np.load(os.path.join(golden_dir, name + '.npz'))
    gm = np.load(os.path.join(golden_dir, 'metrics.npz'))
    U, I, D, B, n_neg, n_epochs = [int(x) for x in g['meta']]
    b = int(g['split'][1])
    test = _interactions(g['users'][b:], g['items'][b:], U, I)
    net = make_net([g['final%d' % i] for i in range(4)])
    model = ImplicitFactorizationModel(embedding_dim=D, representation=net, batch_size=B, use_cuda=True,
                                       experiment_name='gpu_metrics_' + name)
    model.set_users(U, I)
    for k in (1, 5, 10):
        # the only freedom is the order inside groups of tied sigmoid outputs
        assert map_at_k(model, test, k=k) == pytest.approx(float(gm['%s_map_k%d' % (name, k)]), abs=2e-3)
    res = model.test(test, None, k=5, rmse_flag=True, precision_recall=False, map_recall=True)
    assert res['bce'] == pytest.approx(float(gm['%s_bce' % name]), rel=1e-5)
    assert res['map'] == pytest.approx(float(gm['%s_map_k5' % name]), abs=2e-3)


def test_keyed_topk_reuses_the_mask_images_without_changing_results():
    """mfb_topk_keyed: the second call with the same key skips the train-mask preprocessing; ids and scores equal the
    unkeyed call; a new key (other users, other CSR) rebuilds."""
    from recommendation_gans_b200.engine import MFEngine
    rs = np.random.RandomState(21)
    U, I, D, k = 700, 3000, 128, 20
    tabs = _random_tables(rs, U, I, D, 0.3)
    eng = MFEngine(make_net(tabs))

    def csr(seed):
        r = np.random.RandomState(seed)
        c = O.csr_from_pairs(r.randint(0, U, 30 * U), r.randint(0, I, 30 * U), U, I)
        c.sum_duplicates()
        c.sort_indices()
        return (torch.from_numpy(c.indptr.astype(np.int64)).cuda(), torch.from_numpy(c.indices.astype(np.int32)).cuda())
    a, b = csr(1), csr(2)
    users_a, users_b = rs.permutation(U).astype(np.int64), np.arange(0, U, 2, dtype=np.int64)
    ref_a = eng.topk(users_a, k, *a, with_scores=True)
    ref_b = eng.topk(users_b, k, *b, with_scores=True)
    for rep in range(2):
        got = eng.topk(users_a, k, *a, with_scores=True, plan_key=11)
        assert (got[0] == ref_a[0]).all() and (got[1] == ref_a[1]).all()
    got = eng.topk(users_b, k, *b, with_scores=True, plan_key=12)
    assert (got[0] == ref_b[0]).all() and (got[1] == ref_b[1]).all()
    got = eng.topk(users_a, k, *a, with_scores=True, plan_key=11)          # back to the first pair: rebuilt for key 11
    assert (got[0] == ref_a[0]).all() and (got[1] == ref_a[1]).all()


def test_tc_topk_certificates_hold_for_sparse_and_huge_rows():
    """Adversarial inputs for the fp16 error model: one-hot-like rows (no averaging over the row: the rounding error of a
    single product is the whole error), values in the fp16 subnormal range, and values beyond the fp16 range (the call
    must fall back to the exact kernel).  Ids and scores must equal the exact kernel's in every case."""
    from recommendation_gans_b200.engine import MFEngine
    rs = np.random.RandomState(31)
    U, I, D, k = 300, 4000, 128, 20
    cases = {}
    ue, ie = np.zeros((U, D), np.float32), np.zeros((I, D), np.float32)
    ue[np.arange(U), rs.randint(0, 4, U)] = rs.uniform(0.5, 2.0, U).astype(np.float32)
    ie[np.arange(I), rs.randint(0, 4, I)] = (1.0 + rs.randint(0, 2000, I) * 2.0 ** -11).astype(np.float32)   # dense near-ties
    cases['one_hot'] = (ue, ie)
    cases['subnormal'] = (rs.normal(0, 2e-5, (U, D)).astype(np.float32), rs.normal(0, 2e-5, (I, D)).astype(np.float32))
    big_u = rs.normal(0, 0.3, (U, D)).astype(np.float32)
    big_u[5, 7] = 1.0e5
    cases['beyond_fp16'] = (big_u, rs.normal(0, 0.3, (I, D)).astype(np.float32))
    import os
    for name, (a, b) in cases.items():
        tabs = (a, b, np.zeros((U, 1), np.float32), rs.normal(0, 1e-6, (I, 1)).astype(np.float32))
        old = os.environ.get('MFB_TC')
        os.environ['MFB_TC'] = '0'
        exact = MFEngine(make_net(tabs))
        os.environ['MFB_TC'] = '1'
        tc = MFEngine(make_net(tabs))
        if old is None:
            os.environ.pop('MFB_TC')
        else:
            os.environ['MFB_TC'] = old
        users = np.arange(U, dtype=np.int64)
        ids_e, sc_e = exact.topk(users, k, with_scores=True)
        ids_t, sc_t = tc.topk(users, k, with_scores=True)
        assert (ids_e == ids_t).all(), name
        assert (sc_e == sc_t).all(), name
        if name == 'beyond_fp16':
            assert tc.topk_last_redo == U
