"""Fused CUDA training step vs golden vectors produced by the reference (and vs the oracle where
the golden file stops): per-step losses, tables after N steps, validation losses.
Tolerance (north star): 1e-5 relative on fp32 losses and tables."""
import glob
import os

import numpy as np
import pytest
import torch

from tests.gpu_helpers import make_engine, tables_of, rel_err

pytestmark = pytest.mark.gpu
STEP_FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', 'steps_*.npz')))


def _run(path, fast_math=False, chunked=True):
    g = np.load(path)
    U, I, D, B, n_neg = [int(x) for x in g['meta']]
    lr, l2 = [float(x) for x in g['hyper']]
    loss, opt = str(g['loss']), str(g['optimizer'])
    kind = {'pointwise': 'pointwise', 'hinge': 'hinge'}.get(loss, 'adaptive_hinge')   # implicit.py:194-199
    net, _, eng = make_engine([g['init%d' % i] for i in range(4)], opt, lr, l2, fast_math)
    n_steps = len(g['step_losses'])
    negs = g['neg_pairs'][g['neg_idx']]                     # [draws, n_neg*B, 2]
    nu = torch.from_numpy(negs[:n_steps, :, 0].reshape(-1).copy()).cuda()
    ni = torch.from_numpy(negs[:n_steps, :, 1].reshape(-1).copy()).cuda()
    users, items = torch.from_numpy(g['users']).cuda(), torch.from_numpy(g['items']).cuda()
    if chunked:
        losses = eng.train_steps(kind, users, items, B, n_neg, nu, ni).cpu().numpy()
    else:   # one call per step: exercises the single-step path and the partial last batch
        out = []
        for s in range(n_steps):
            k = n_neg * B
            out.append(eng.train_steps(kind, users[s * B:(s + 1) * B], items[s * B:(s + 1) * B], B, n_neg,
                                       nu[s * k:(s + 1) * k], ni[s * k:(s + 1) * k]).cpu().numpy()[0])
        losses = np.array(out)
    assert eng.step == n_steps
    n_val = len(g['val_losses'])
    vu = torch.from_numpy(negs[n_steps:, :, 0].reshape(-1).copy()).cuda()
    vi = torch.from_numpy(negs[n_steps:, :, 1].reshape(-1).copy()).cuda()
    val = eng.loss_steps(kind, users[:n_val * B], items[:n_val * B], B, n_neg, vu, vi).cpu().numpy()
    eng.flush()
    torch.cuda.synchronize()
    return g, lr, losses, val, tables_of(net)


@pytest.mark.parametrize('path', STEP_FILES, ids=[os.path.basename(p)[6:-4] for p in STEP_FILES])
@pytest.mark.parametrize('chunked', [True, False], ids=['bulk', 'stepwise'])
def test_steps_match_reference(path, chunked):
    g, lr, losses, val, tables = _run(path, chunked=chunked)
    np.testing.assert_allclose(losses, g['step_losses'], rtol=1e-5)
    np.testing.assert_allclose(val, g['val_losses'], rtol=1e-5)
    for i, t in enumerate(tables):
        ref = g['final%d' % i]
        err = rel_err(t, ref)
        # bias rows with cancelling hinge gradients are ill-conditioned under Adam: bounded in lr units
        assert err < 1e-5 or (i >= 2 and np.abs(t - ref).max() < 2e-2 * lr), (i, err)


@pytest.mark.parametrize('name', ['pointwise_adam', 'bpr_adam'])
def test_fast_math_replay_stays_within_tolerance(golden_dir, name):
    g, lr, losses, val, tables = _run(os.path.join(golden_dir, 'steps_%s.npz' % name), fast_math=True)
    np.testing.assert_allclose(losses, g['step_losses'], rtol=1e-5)
    for i, t in enumerate(tables):
        assert rel_err(t, g['final%d' % i]) < 1e-5, i


def test_out_of_range_ids_raise(golden_dir):
    g = np.load(os.path.join(golden_dir, 'steps_pointwise_sgd.npz'))
    U, I, D, B, n_neg = [int(x) for x in g['meta']]
    _, _, eng = make_engine([g['init%d' % i] for i in range(4)], 'sgd', 1e-2, 0.0)
    users = torch.from_numpy(g['users'][:B].copy()).cuda()
    items = torch.from_numpy(g['items'][:B].copy()).cuda()
    users[3] = U                                              # one past the end
    with pytest.raises(ValueError):
        eng.train_steps('pointwise', users, items, B, 0)
    with pytest.raises(ValueError):
        eng.predict_pairs(users, items)
    with pytest.raises(RuntimeError):                          # hinge with len(neg) != len(pos)
        eng.train_steps('hinge', users[:B - 1] * 0, items[:B - 1], B, 1, users, items)


def test_losses_match_reference(golden_dir):
    import spotlight.losses as L
    g = np.load(os.path.join(golden_dir, 'forward_losses.npz'))
    for tag in ('same', 'five'):
        for name in ('pointwise', 'bpr', 'hinge', 'adaptive_hinge'):
            key = '%s_%s' % (name, tag)
            if 'loss_' + key not in g:
                with pytest.raises(RuntimeError):
                    getattr(L, name + '_loss')(torch.from_numpy(g['pos']).cuda(), torch.from_numpy(g['neg_' + tag]).cuda())
                continue
            pos = torch.from_numpy(g['pos']).cuda().requires_grad_(True)
            neg = torch.from_numpy(g['neg_' + tag]).cuda().requires_grad_(True)
            val = getattr(L, name + '_loss')(pos, neg)
            val.backward()
            np.testing.assert_allclose(val.item(), g['loss_' + key], rtol=2e-6)
            np.testing.assert_allclose(pos.grad.cpu().numpy(), g['dpos_' + key], rtol=2e-6, atol=1e-9)
            np.testing.assert_allclose(neg.grad.cpu().numpy(), g['dneg_' + key], rtol=2e-6, atol=1e-9)


def test_losses_with_mask_and_2d_negatives_match_reference(golden_dir):
    """`mask=` and adaptive_hinge_loss on [n, b] negatives (losses.py:51-55,91-95,124-128,170) vs vectors frozen from
    the reference by oracle/make_golden_losses_ex.py; also checked against the oracle on fresh random inputs."""
    import spotlight.losses as L
    from oracle import mf_oracle as O
    g = np.load(os.path.join(golden_dir, 'losses_ex.npz'))
    for case in g['cases']:
        name, negkey, use_mask = str(case).split('|')
        pos = torch.from_numpy(g['pos'].copy()).cuda().requires_grad_(True)
        neg = torch.from_numpy(g[negkey].copy()).cuda().requires_grad_(True)
        mask = torch.from_numpy(g['mask']).cuda() if use_mask == '1' else None
        val = getattr(L, name + '_loss')(pos, neg, mask=mask)
        val.backward()
        tag = '%s_%s_%s' % (name, negkey, 'mask' if use_mask == '1' else 'nomask')
        np.testing.assert_allclose(val.item(), g['loss_' + tag], rtol=2e-6)
        np.testing.assert_allclose(pos.grad.cpu().numpy(), g['dpos_' + tag], rtol=2e-6, atol=1e-9)
        np.testing.assert_allclose(neg.grad.cpu().numpy(), g['dneg_' + tag], rtol=2e-6, atol=1e-9)
    rs = np.random.RandomState(5)
    for b, n in ((1, 1), (33, 7), (4097, 3)):                 # sizes around the block size; a single pair
        pos_h = rs.uniform(0.01, 0.99, b).astype(np.float32)
        neg_h = rs.uniform(0.01, 0.99, (n, b)).astype(np.float32)
        mask_h = (rs.uniform(0, 1, b) < 0.5).astype(np.float32)
        mask_h[0] = 1.0
        pc, nc = torch.from_numpy(pos_h).requires_grad_(True), torch.from_numpy(neg_h).requires_grad_(True)
        ref = O.adaptive_hinge_loss(pc, nc, mask=torch.from_numpy(mask_h))
        ref.backward()
        pg, ng = torch.from_numpy(pos_h).cuda().requires_grad_(True), torch.from_numpy(neg_h).cuda().requires_grad_(True)
        val = L.adaptive_hinge_loss(pg, ng, mask=torch.from_numpy(mask_h).cuda())
        val.backward()
        np.testing.assert_allclose(val.item(), ref.item(), rtol=2e-6)
        np.testing.assert_allclose(pg.grad.cpu().numpy(), pc.grad.numpy(), rtol=2e-6, atol=1e-9)
        np.testing.assert_allclose(ng.grad.cpu().numpy(), nc.grad.numpy(), rtol=2e-6, atol=1e-9)
    with pytest.raises(NotImplementedError):                   # pointwise flattens: 2-D negatives are outside the path
        L.pointwise_loss(torch.rand(4).cuda(), torch.rand(2, 4).cuda())
    with pytest.raises(RuntimeError):                          # column count must match the positives
        L.hinge_loss(torch.rand(4).cuda(), torch.rand(2, 5).cuda())


def test_forward_matches_reference(golden_dir):
    from tests.gpu_helpers import make_net
    g = np.load(os.path.join(golden_dir, 'forward_losses.npz'))
    net = make_net([g['tables%d' % i] for i in range(4)])
    pred = net(torch.from_numpy(g["users"]).cuda(), torch.from_numpy(g["items"]).cuda()).detach().cpu().numpy()
    np.testing.assert_allclose(pred, g['pred'], rtol=1e-5, atol=1e-7)
    with pytest.raises(IndexError):                            # the reference breaks on a single pair
        net(torch.tensor([1]).cuda(), torch.tensor([2]).cuda())


@pytest.mark.parametrize('loss,opt,n_neg', [('pointwise', 'adam', 2), ('adaptive_hinge', 'adam', 1),
                                            ('bpr', 'adam', 1), ('hinge', 'sgd', 1)])
@pytest.mark.parametrize('fast_math', [False, True], ids=['ieee', 'fast'])
def test_many_chunks_match_oracle(loss, opt, n_neg, fast_math):
    """Several planner chunks (look-ahead catch-up, double-buffered planning) and Zipf-skewed items
    (multi-window segment reduction) against the oracle run on the same inputs."""
    from oracle import mf_oracle as O
    torch.set_num_threads(1)
    rs = np.random.RandomState(17)
    U, I, D, B = 500, 300, 16, 32
    n_steps = 150
    n_pos = n_steps * B - (0 if loss in ('bpr', 'hinge') else 5)     # partial last batch where the loss allows it
    p = 1.0 / np.arange(1, I + 1) ** 1.05
    users = rs.randint(0, U, n_pos)
    items = rs.choice(I, n_pos, p=p / p.sum())
    neg = np.stack([rs.randint(0, U, n_steps * n_neg * B), rs.randint(0, I, n_steps * n_neg * B)], 1)
    tabs = [t.numpy() for t in O.init_tables(U, I, D, torch_seed=3)]
    lr, l2 = (1e-3, 1e-5) if opt == 'adam' else (5e-2, 1e-4)
    oracle = O.OracleMF(*[torch.from_numpy(t) for t in tabs], optimizer=opt, lr=lr, l2=l2, batch_size=B,
                        num_negative_samples=n_neg, loss_fn=O.LOSS_FUNCTIONS[loss])
    ref_losses = []
    for s in range(n_steps):
        k = n_neg * B
        ref_losses.append(oracle.train_step(torch.from_numpy(users[s * B:(s + 1) * B]),
                                            torch.from_numpy(items[s * B:(s + 1) * B]),
                                            torch.from_numpy(neg[s * k:(s + 1) * k, 0].copy()),
                                            torch.from_numpy(neg[s * k:(s + 1) * k, 1].copy())).item())
    net, _, eng = make_engine(tabs, opt, lr, l2, fast_math)
    losses = eng.train_steps(loss, users, items, B, n_neg, neg[:, 0].copy(), neg[:, 1].copy()).cpu().numpy()
    eng.flush()
    np.testing.assert_allclose(losses, ref_losses, rtol=1e-5)
    for i, (got, exp) in enumerate(zip(tables_of(net), oracle.numpy_tables())):
        err = rel_err(got, exp)
        assert err < 1e-5 or (i >= 2 and np.abs(got - exp).max() < 2e-2 * lr), (i, err)
