"""The oracle restatement (oracle/mf_oracle.py) against golden vectors produced by running the
reference itself (oracle/make_golden.py), and the lazy row catch-up specification
(tests/lazy_model.py) against both.  CPU only."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import mf_oracle as O
from oracle import mt19937_ref as R
from tests.lazy_model import LazyMF, loss_and_dpred

torch.set_num_threads(1)
STEP_FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', 'steps_*.npz')))


def _t(x):
    return torch.from_numpy(np.asarray(x))


def test_forward_and_losses(golden_dir):
    g = np.load(os.path.join(golden_dir, 'forward_losses.npz'))
    tabs = [_t(g['tables%d' % i]) for i in range(4)]
    pred = O.bilinear_forward(*tabs, _t(g['users']), _t(g['items'])).numpy()
    np.testing.assert_allclose(pred, g['pred'], rtol=1e-6, atol=1e-7)
    for tag in ('same', 'five'):
        for name in ('pointwise', 'bpr', 'hinge', 'adaptive_hinge'):
            key = '%s_%s' % (name, tag)
            if 'loss_' + key not in g:
                continue
            pos = _t(g['pos']).clone().requires_grad_(True)
            neg = _t(g['neg_' + tag]).clone().requires_grad_(True)
            val = O.LOSS_FUNCTIONS[name](pos, neg)
            val.backward()
            np.testing.assert_allclose(val.item(), g['loss_' + key], rtol=1e-6)
            np.testing.assert_allclose(pos.grad.numpy(), g['dpos_' + key], rtol=1e-6, atol=1e-9)
            np.testing.assert_allclose(neg.grad.numpy(), g['dneg_' + key], rtol=1e-6, atol=1e-9)
            # the closed-form gradients the CUDA kernels implement
            lv, dpos, dneg = loss_and_dpred(name, g['pos'], g['neg_' + tag])
            np.testing.assert_allclose(lv, g['loss_' + key], rtol=2e-6)
            np.testing.assert_allclose(dpos, g['dpos_' + key], rtol=2e-6, atol=1e-9)
            np.testing.assert_allclose(dneg, g['dneg_' + key], rtol=2e-6, atol=1e-9)


@pytest.mark.parametrize('path', STEP_FILES, ids=[os.path.basename(p)[6:-4] for p in STEP_FILES])
def test_steps_oracle_and_lazy_model(path):
    g = np.load(path)
    U, I, D, B, n_neg = [int(x) for x in g['meta']]
    lr, l2 = [float(x) for x in g['hyper']]
    loss, opt = str(g['loss']), str(g['optimizer'])
    init = [g['init%d' % i] for i in range(4)]
    model = O.OracleMF(*[_t(t) for t in init], loss=loss, optimizer=opt, lr=lr, l2=l2,
                       batch_size=B, num_negative_samples=n_neg, neg_pairs=g['neg_pairs'])
    kind = {'pointwise': 'pointwise', 'hinge': 'hinge'}.get(loss, 'adaptive_hinge')
    lazy = LazyMF(init, opt, lr, l2)
    gen = R.MT19937.from_python_seed(int(g['py_seed']))
    users, items = g['users'], g['items']
    n_steps = len(g['step_losses'])
    losses, lazy_losses = [], []
    for s in range(n_steps):
        pu, pi = users[s * B:(s + 1) * B], items[s * B:(s + 1) * B]
        nu, ni = model.draw_negatives(gen)
        assert (g['neg_pairs'][g['neg_idx'][s]] == np.stack([nu.numpy(), ni.numpy()], 1)).all()
        losses.append(model.train_step(_t(pu), _t(pi), nu, ni).item())
        lazy_losses.append(lazy.train_step(kind, pu, pi, nu.numpy(), ni.numpy()))
    np.testing.assert_allclose(losses, g['step_losses'], rtol=1e-6)
    np.testing.assert_allclose(lazy_losses, g['step_losses'], rtol=1e-5)
    for i, t in enumerate(model.numpy_tables()):
        np.testing.assert_allclose(t, g['final%d' % i], rtol=1e-6, atol=1e-9)
    lazy.flush()
    for i, t in enumerate(lazy.tables()):
        ref = g['final%d' % i]
        err = np.abs(t - ref).max() / np.abs(ref).max()
        # Bias rows that receive equal-and-opposite hinge gradients at init (pos and neg both at
        # p~0.5) have g ~ 0 up to rounding; Adam's m/sqrt(v) normalisation turns that 1-ulp noise into
        # O(lr) differences.  Those rows are bounded in units of lr instead.
        ok = err < 1e-5 or (i >= 2 and np.abs(t - ref).max() < 2e-2 * lr)
        assert ok, (i, err)
    # validation iterations continue the same stream (implicit.py:366-379)
    val = []
    for s in range(len(g['val_losses'])):
        nu, ni = model.draw_negatives(gen)
        val.append(model.val_step(_t(users[s * B:(s + 1) * B]), _t(items[s * B:(s + 1) * B]), nu, ni).item())
    np.testing.assert_allclose(val, g['val_losses'], rtol=1e-6)


@pytest.mark.parametrize('name', ['fit_pointwise', 'fit_bpr'])
def test_fit_predict_evaluate(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + '.npz'))
    U, I, D, B, n_neg, n_epochs = [int(x) for x in g['meta']]
    lr, l2 = [float(x) for x in g['hyper']]
    a, b = [int(x) for x in g['split']]
    users, items = g['users'], g['items']
    rs = np.random.RandomState(0)
    rs.randint(-10 ** 8, 10 ** 8)                       # implicit.py:146 consumes one draw
    perm = np.arange(a)
    rs.shuffle(perm)                                     # torch_utils.py:50-51
    model = O.OracleMF(*[_t(g['init%d' % i]) for i in range(4)], loss=str(g['loss']), optimizer='adam',
                       lr=lr, l2=l2, batch_size=B, num_negative_samples=n_neg, neg_pairs=g['neg_pairs'])
    gen = R.MT19937.from_python_seed(int(g['py_seed']))
    out = O.fit_epochs(model, users[:a][perm], items[:a][perm], users[a:b], items[a:b], n_epochs, gen)
    np.testing.assert_allclose(out['train'], g['summary'][:, 0], rtol=1e-6)
    np.testing.assert_allclose(out['val'], g['summary'][:, 1], rtol=1e-6)
    assert int(g['best_epoch']) == int(np.argmin(out['val']))   # best == last here, so final tables compare
    assert int(g['best_epoch']) == n_epochs - 1
    for i, t in enumerate(model.numpy_tables()):
        np.testing.assert_allclose(t, g['final%d' % i], rtol=1e-6, atol=1e-8)
    assert tuple(g['py_random_after'][:624]) == tuple(gen.mt.tolist()) and g['py_random_after'][624] == gen.pos
    np.testing.assert_allclose(model.predict(3), g['predict_user3'], rtol=1e-6)
    np.testing.assert_allclose(model.predict(g['predict_pairs_users'], g['predict_pairs_items']),
                               g['predict_pairs'], rtol=1e-6)
    train = O.csr_from_pairs(users[:a], items[:a], U, I)
    test = O.csr_from_pairs(users[b:], items[b:], U, I)
    for k in (5, 10, 20):
        p, r, cold, _ = O.precision_recall_score(model, test, train, k, ranking='reference')
        np.testing.assert_allclose([p, r], g['pr_masked_k%d' % k], rtol=1e-9)
        p, r, _, _ = O.precision_recall_score(model, test, None, k, ranking='reference')
        np.testing.assert_allclose([p, r], g['pr_nomask_k%d' % k], rtol=1e-9)
        # tie-defined oracle (stable, logits): same metric up to tie effects on this trained model
        p2, r2, _, _ = O.precision_recall_score(model, test, train, k, ranking='stable_logit')
        np.testing.assert_allclose([p2, r2], g['pr_masked_k%d' % k], atol=2e-3)
    p, r, _, _ = O.precision_recall_score(model, test, train, np.array([5, 10, 20]), ranking='reference')
    np.testing.assert_allclose([p, r], g['pr_masked_karray'], rtol=1e-9)


def test_losses_with_mask_and_2d_negatives_match_reference(golden_dir):
    """spotlight/losses.py `mask=` and adaptive_hinge_loss on [n, b] negatives (tests/golden/losses_ex.npz, frozen
    from the reference by oracle/make_golden_losses_ex.py)."""
    import torch
    from oracle import mf_oracle as O
    g = np.load(os.path.join(golden_dir, 'losses_ex.npz'))
    for case in g['cases']:
        name, negkey, use_mask = str(case).split('|')
        pos = torch.from_numpy(g['pos'].copy()).requires_grad_(True)
        neg = torch.from_numpy(g[negkey].copy()).requires_grad_(True)
        mask = torch.from_numpy(g['mask']) if use_mask == '1' else None
        val = O.LOSS_FUNCTIONS[name](pos, neg, mask=mask)
        val.backward()
        tag = '%s_%s_%s' % (name, negkey, 'mask' if use_mask == '1' else 'nomask')
        np.testing.assert_allclose(val.detach().numpy(), g['loss_' + tag], rtol=1e-6)
        np.testing.assert_allclose(pos.grad.numpy(), g['dpos_' + tag], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(neg.grad.numpy(), g['dneg_' + tag], rtol=1e-6, atol=1e-9)


def test_negative_pair_generator_matches_reference(golden_dir):
    """oracle.get_negative_samples vs the reference's get_negative_samples (tests/golden/neg_samples.npz, frozen by
    oracle/make_golden_negsamples.py): pairs bit-exact and numpy's generator left in the same state."""
    g = np.load(os.path.join(golden_dir, 'neg_samples.npz'))
    for name in g['cases']:
        name = str(name)
        U, I, N, seed = [int(x) for x in g[name + '_meta']]
        rs = np.random.RandomState(seed)
        pairs = O.get_negative_samples(g[name + '_users'], g[name + '_items'], U, I, N, rs)
        assert (pairs == g[name + '_pairs']).all()
        st = rs.get_state()
        assert (st[1] == g[name + '_state_key']).all() and st[2] == int(g[name + '_state_pos'])
        import collections
        seen = collections.Counter(zip(g[name + '_users'].tolist(), g[name + '_items'].tolist()))
        once = {p for p, c in seen.items() if c == 1}    # a repeated pair is stored as 2: has_key is False for it
        assert (1, 3) not in once and not [p for p in map(tuple, pairs.tolist()) if p in once]


@pytest.mark.parametrize('name', ['fit_pointwise', 'fit_bpr'])
def test_map_at_k_and_rmse_match_reference(golden_dir, name):
    """oracle map_at_k / rmse_score vs values frozen from the reference's own functions (oracle/make_golden_metrics.py,
    evaluation.py:187-190,278-353) on the fitted tables of the fit goldens."""
    g = np.load(os.path.join(golden_dir, name + '.npz'))
    gm = np.load(os.path.join(golden_dir, 'metrics.npz'))
    U, I, D, B, n_neg, n_epochs = [int(x) for x in g['meta']]
    b = int(g['split'][1])
    users, items = g['users'], g['items']
    model = O.OracleMF(*[_t(g['final%d' % i]) for i in range(4)])
    test_csr = O.csr_from_pairs(users[b:], items[b:], U, I)
    for k in (1, 5, 10):
        assert O.map_at_k(model, test_csr, k=k, ranking='reference') == pytest.approx(float(gm['%s_map_k%d' % (name, k)]),
                                                                                       abs=1e-12)
        # the tie-defined ranking (what the CUDA path implements) differs only through tie order
        assert O.map_at_k(model, test_csr, k=k) == pytest.approx(float(gm['%s_map_k%d' % (name, k)]), abs=2e-3)
    parts = [O.rmse_score(model, users[b:][s:s + B], items[b:][s:s + B]) for s in range(0, len(users) - b, B)]
    np.testing.assert_allclose(parts, gm['%s_rmse_parts' % name], rtol=1e-6)
