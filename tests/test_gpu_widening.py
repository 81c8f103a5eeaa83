"""Rows of SURVEY section 8(f) that widen the path: k beyond 32, the `representation=` escape hatch (MLP / NeuMF scored by
a torch module, negatives and ranking still on the CUDA path), and differentiable BilinearNet.forward for
hand-written training loops."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import mf_oracle as O
from oracle import mt19937_ref as R
from tests.gpu_helpers import make_net

pytestmark = pytest.mark.gpu


def _interactions(u, i, U, I):
    from spotlight.interactions import Interactions
    return Interactions(u.astype(np.int32), i.astype(np.int32), num_users=U, num_items=I)


def _dyadic_tables(rs, U, I, D):
    return (rs.randint(-8, 9, (U, D)).astype(np.float32) / 16.0, rs.randint(-8, 9, (I, D)).astype(np.float32) / 16.0,
            rs.randint(-4, 5, (U, 1)).astype(np.float32) / 8.0, rs.randint(-4, 5, (I, 1)).astype(np.float32) / 8.0)


@pytest.mark.parametrize('U,I,D,k', [(40, 700, 32, 33), (33, 900, 128, 100), (20, 1500, 64, 256), (9, 1200, 16, 700),
                                     (6, 300, 50, 300)])
def test_topk_beyond_32_matches_stable_ranking(U, I, D, k):
    """precision_recall_score accepts any k (evaluation.py:144-150): k > 32 runs the exact kernel with a
    multi-register warp list (<= 256 ranks per pass) and resumes after the previous pass's last (score, id)."""
    from recommendation_gans_b200.engine import MFEngine
    rs = np.random.RandomState(U + k)
    tabs = _dyadic_tables(rs, U, I, D)
    eng = MFEngine(make_net(tabs))
    tu, ti = rs.randint(0, U, 40 * U), rs.randint(0, I, 40 * U)
    train = O.csr_from_pairs(tu, ti, U, I)
    train.sum_duplicates()
    train.sort_indices()
    indptr = torch.from_numpy(train.indptr.astype(np.int64)).cuda()
    indices = torch.from_numpy(train.indices.astype(np.int32)).cuda()
    model = O.OracleMF(*[torch.from_numpy(t) for t in tabs])
    for masked in (True, False):
        got = eng.topk(np.arange(U), k, indptr if masked else None, indices if masked else None).cpu().numpy()
        for u in range(U):
            exp = O.topk_stable(model.logits(u), train[u].indices if masked else None, k)
            assert (got[u] == exp).all(), (u, masked)


def test_precision_recall_with_large_and_oversized_k():
    """k = 50 and a k beyond the catalogue size (the reference slices predictions[:k] of a num_items-long ranking and
    still divides by k)."""
    from spotlight.evaluation import precision_recall_score
    rs = np.random.RandomState(11)
    U, I, D = 60, 90, 16
    tabs = _dyadic_tables(rs, U, I, D)
    net = make_net(tabs)

    class _Model(object):
        _net = net
        _num_users, _num_items = U, I
    test = _interactions(rs.randint(0, U, 400), rs.randint(0, I, 400), U, I)
    train = _interactions(rs.randint(0, U, 900), rs.randint(0, I, 900), U, I)
    oracle = O.OracleMF(*[torch.from_numpy(t) for t in tabs])
    for k in (50, 200, np.array([5, 50, 200])):
        p, r = precision_recall_score(_Model(), test, train=train, k=k)
        ep, er, _, _ = O.precision_recall_score(oracle, test.tocsr(), train.tocsr(), k=k)
        assert p == pytest.approx(ep, abs=1e-12) and r == pytest.approx(er, abs=1e-12)


@pytest.mark.parametrize('k', [5, 40, 300])
def test_dense_score_topk_matches_stable_ranking(k):
    """mfb_topk_scores: rows of an arbitrary dense score matrix, ties -> lower id, train items last."""
    from recommendation_gans_b200.engine import topk_scores_device
    rs = np.random.RandomState(k)
    n, I = 37, 1000
    scores = (rs.randint(0, 64, (n, I)) / 64.0).astype(np.float32)        # many exact ties
    users = rs.permutation(50)[:n].astype(np.int64)
    train = O.csr_from_pairs(rs.randint(0, 50, 2000), rs.randint(0, I, 2000), 50, I)
    train.sum_duplicates()
    train.sort_indices()
    ids, sc = topk_scores_device(torch.from_numpy(scores).cuda(), k, users,
                                 torch.from_numpy(train.indptr.astype(np.int64)).cuda(),
                                 torch.from_numpy(train.indices.astype(np.int32)).cuda(), with_scores=True)
    ids, sc = ids.cpu().numpy(), sc.cpu().numpy()
    for r in range(n):
        exp = O.topk_stable(scores[r], train[users[r]].indices, k)
        assert (ids[r] == exp).all(), r
        rated = set(train[users[r]].indices.tolist())
        for j in range(k):
            assert sc[r, j] == (np.float32(-3.402823466e38) if ids[r, j] in rated else scores[r, ids[r, j]])
    ids2 = topk_scores_device(torch.from_numpy(scores).cuda(), k).cpu().numpy()       # no mask, row = user
    for r in range(n):
        assert (ids2[r] == O.topk_stable(scores[r], None, k)).all()


@pytest.mark.parametrize('kind', ['neuMF', 'mlp'])
def test_representation_escape_hatch_trains_and_evaluates(kind, tmp_path, monkeypatch):
    """ImplicitFactorizationModel(representation=NeuMF/MLP) (ncf_spotlight.py / neuMF_spotlight.py usage,
    implicit.py:169-180): negatives come from Python's `random` stream exactly as the reference would consume it, the
    loss kernels train the module through autograd, predict works, and precision_recall_score ranks the module's own
    scores with the CUDA kernel (checked against a stable host ranking of model.predict)."""
    monkeypatch.chdir(tmp_path)
    from implicit import ImplicitFactorizationModel
    from spotlight.dnn_models.mlp import MLP
    from spotlight.dnn_models.neuMF import NeuMF
    from spotlight.evaluation import precision_recall_score, topk_for_users
    import spotlight.optimizers as optimizers
    rs = np.random.RandomState(4)
    U, I, B, n_neg, epochs = 80, 120, 64, 2, 3
    n_all = 3000
    users = rs.randint(0, U, n_all)
    items = ((users * 7 + rs.randint(0, 5, n_all)) % I).astype(np.int64)          # learnable structure
    a, b = int(0.81 * n_all), int(0.9 * n_all)
    train, valid, test = (_interactions(users[:a], items[:a], U, I), _interactions(users[a:b], items[a:b], U, I),
                          _interactions(users[b:], items[b:], U, I))
    neg_pairs = [(int(u), int(i)) for u, i in zip(rs.randint(0, U, a), rs.randint(0, I, a))]
    torch.manual_seed(0)
    net = (NeuMF(mlp_layers=[32, 16, 8], num_users=U, num_items=I, mf_embedding_dim=8, mlp_embedding_dim=16)
           if kind == 'neuMF' else MLP(layers=[32, 16, 8], num_users=U, num_items=I, embedding_dim=16))
    model = ImplicitFactorizationModel(loss='pointwise', n_iter=epochs, batch_size=B, l2=1e-6, learning_rate=5e-3,
                                       optimizer_func=optimizers.adam_optimizer, representation=net, use_cuda=True,
                                       random_state=np.random.RandomState(0), neg_examples=neg_pairs,
                                       num_negative_samples=n_neg, experiment_name='gpu_' + kind)
    random.seed(9)
    model.fit(train, valid, verbose=False)
    # the `random` stream ends where the reference's per-minibatch random.choices calls would leave it
    gen = R.MT19937.from_python_seed(9)
    steps = epochs * (-(-a // B) + -(-(b - a) // B))
    R.choices_indices(gen, len(neg_pairs), steps * n_neg * B)
    assert tuple(int(x) for x in gen.mt) + (gen.pos,) == random.getstate()[1]
    import csv
    with open(os.path.join(model.experiment_logs, 'summary.csv')) as f:
        rows = list(csv.reader(f))
    losses = np.array([[float(x) for x in r] for r in rows[1:]])
    assert losses.shape[0] == epochs and losses[-1, 0] < losses[0, 0]               # it learns
    pred = model.predict(3)
    assert pred.shape == (I,) and np.isfinite(pred).all()
    with pytest.raises(ValueError):
        model.predict(U)
    k = 10
    got = topk_for_users(model, np.arange(U), k, train=train).cpu().numpy()
    train_csr = train.tocsr()
    for u in range(0, U, 7):
        assert (got[u] == O.topk_stable(model.predict(u), train_csr[u].indices, k)).all()
    p, r = precision_recall_score(model, test, train=train, k=k)
    test_csr = test.tocsr()
    exp_p = np.mean([len(set(O.topk_stable(model.predict(u), train_csr[u].indices, k).tolist())
                         & set(test_csr[u].indices.tolist())) / k
                     for u in range(U) if len(test_csr[u].indices)])
    assert p == pytest.approx(exp_p, abs=1e-12)
    res = model.test(test, None, k=5, rmse_flag=True, precision_recall=False, map_recall=True)
    assert set(res) >= {'k', 'bce', 'map'}


def test_bilinear_forward_is_differentiable_like_the_reference():
    """A hand-written loop -- loss(net(u, i), net(u', i')).backward() -- gets the dense table gradients autograd
    produces through the reference's nn.Embedding expressions (checked against the oracle's torch CPU graph)."""
    import spotlight.losses as L
    rs = np.random.RandomState(2)
    U, I, D, b = 50, 70, 16, 40
    tabs = [t.numpy() for t in O.init_tables(U, I, D, torch_seed=1)]
    tabs[2] = rs.normal(0, 0.1, (U, 1)).astype(np.float32)
    tabs[3] = rs.normal(0, 0.1, (I, 1)).astype(np.float32)
    net = make_net(tabs)
    pu, pi = rs.randint(0, U, b), rs.randint(0, I, b)
    nu, ni = rs.randint(0, U, 3 * b), rs.randint(0, I, 3 * b)
    pu[:5] = pu[5:10]                                                   # duplicate rows in the batch
    loss = L.pointwise_loss(net(torch.from_numpy(pu).cuda(), torch.from_numpy(pi).cuda()),
                            net(torch.from_numpy(nu).cuda(), torch.from_numpy(ni).cuda()))
    loss.backward()
    ref = [torch.from_numpy(t).clone().requires_grad_(True) for t in tabs]
    ref_loss = O.pointwise_loss(O.bilinear_forward(*ref, torch.from_numpy(pu), torch.from_numpy(pi)),
                                O.bilinear_forward(*ref, torch.from_numpy(nu), torch.from_numpy(ni)))
    ref_loss.backward()
    assert loss.item() == pytest.approx(ref_loss.item(), rel=1e-5)
    params = (net.user_embeddings.weight, net.item_embeddings.weight, net.user_biases.weight, net.item_biases.weight)
    for p, r in zip(params, ref):
        np.testing.assert_allclose(p.grad.cpu().numpy(), r.grad.numpy(), rtol=2e-5, atol=1e-9)
