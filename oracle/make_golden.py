"""Generate tests/golden/*.npz by running the REAL reference (imported from /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python oracle/make_golden.py
The reference has no tests / golden vectors of its own, so these files are the
pin for the oracle (`oracle/mf_oracle.py`, `oracle/mt19937_ref.py`) and for the CUDA path.
Every array below is an input to or an output of unmodified reference code; nothing here
comes from this repository's implementation.
"""
import os
import random
import sys
import tempfile
import types

import numpy as np
import torch

REF = os.environ.get('REF_PATH', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')
# GOLDEN_ONLY_STEPS=name1,name2 regenerates only those step-level files (steps_<name>.npz)
ONLY_STEPS = set(filter(None, os.environ.get('GOLDEN_ONLY_STEPS', '').split(',')))


def import_reference():
    sys.path.insert(0, REF)
    sys.modules.setdefault('h5py', types.ModuleType('h5py'))  # only datasets/movielens.py imports it
    os.chdir(tempfile.mkdtemp(prefix='refrun_'))              # implicit.py:97-112 mkdirs in cwd
    import implicit as ref_implicit
    from spotlight import sampling, losses, evaluation, optimizers
    from spotlight.interactions import Interactions
    from spotlight.factorization.representations import BilinearNet
    return ref_implicit, sampling, losses, evaluation, optimizers, Interactions, BilinearNet


def synth(rs, num_users, num_items, n, zipf=False):
    users = rs.randint(0, num_users, n).astype(np.int64)
    if zipf:
        p = 1.0 / np.arange(1, num_items + 1) ** 1.05
        items = rs.choice(num_items, n, p=p / p.sum()).astype(np.int64)
    else:
        items = rs.randint(0, num_items, n).astype(np.int64)
    return users, items


def tables_of(net):
    sd = net.state_dict()
    return [sd[k].detach().cpu().numpy().copy() for k in
            ('user_embeddings.weight', 'item_embeddings.weight', 'user_biases.weight', 'item_biases.weight')]


def main():
    ref_implicit, sampling, losses, evaluation, optimizers, Interactions, BilinearNet = import_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    if not ONLY_STEPS:
        make_streams_forward_losses(sampling, losses, BilinearNet)
    make_steps(ref_implicit, optimizers, Interactions, BilinearNet)
    if not ONLY_STEPS:
        make_fits(ref_implicit, evaluation, optimizers, Interactions, BilinearNet)


def make_streams_forward_losses(sampling, losses, BilinearNet):

    # ---------------------------------------------------------------- A. index streams
    rng = {}
    for seed, n, cnt in ((0, 1682, 4096), (1, 3706, 4096), (2, 26744, 8192), (3, 2000000, 4096),
                         (4, 1, 16), (5, 2, 64), (6, 1025, 2000)):
        rng['sample_items_s%d_n%d' % (seed, n)] = sampling.sample_items(
            None, None, n, (cnt,), np.random.RandomState(seed))
    for seed, L, k in ((0, 80000, 5120), (5, 81000, 3000), (2 ** 70 + 17, 16200213, 16384), (9, 7, 50)):
        random.seed(seed)
        pop = range(L)
        rng['choices_s%d_L%d' % (seed % 1000, L)] = np.array(random.choices(pop, k=k), dtype=np.int64)
        rng['choices_s%d_L%d_second' % (seed % 1000, L)] = np.array(random.choices(pop, k=k // 2 + 1), dtype=np.int64)
    rng['seed_big'] = np.array([2 ** 70 + 17], dtype=object).astype(str)
    np.savez_compressed(os.path.join(OUT, 'rng_streams.npz'), **rng)

    # ---------------------------------------------------------------- B/C. forward + losses
    torch.manual_seed(0)
    net = BilinearNet(50, 40, 24)
    net.user_biases.weight.data.normal_(0, 0.3)
    net.item_biases.weight.data.normal_(0, 0.3)
    net.user_embeddings.weight.data.mul_(24 * 0.4)
    rs = np.random.RandomState(11)
    fu, fi = synth(rs, 50, 40, 300)
    pred = net(torch.from_numpy(fu), torch.from_numpy(fi)).detach().numpy()
    fl = {'tables%d' % i: t for i, t in enumerate(tables_of(net))}
    fl.update(users=fu, items=fi, pred=pred)
    pos = torch.from_numpy(rs.uniform(0.01, 0.99, 96).astype(np.float32)).requires_grad_(True)
    pos.data[0] = 1.0
    pos.data[1] = 1e-30
    negs = {'same': 96, 'five': 480}
    for tag, m in negs.items():
        neg = torch.from_numpy(rs.uniform(0.01, 0.99, m).astype(np.float32)).requires_grad_(True)
        if tag == 'same':
            neg.data[5] = neg.data[17] = 0.995  # tie for the max -> first index gets the gradient
            neg.data[2] = 1.0
        fl['neg_' + tag] = neg.detach().numpy().copy()
        for name in ('pointwise', 'bpr', 'hinge', 'adaptive_hinge'):
            if tag == 'five' and name in ('bpr', 'hinge'):
                continue  # broadcasting error in the reference (shapes differ)
            fn = getattr(losses, name + '_loss')
            pos.grad = None
            neg.grad = None
            val = fn(pos, neg)
            val.backward()
            fl['loss_%s_%s' % (name, tag)] = val.detach().numpy()
            fl['dpos_%s_%s' % (name, tag)] = pos.grad.numpy().copy()
            fl['dneg_%s_%s' % (name, tag)] = neg.grad.numpy().copy()
    fl['pos'] = pos.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, 'forward_losses.npz'), **fl)



def make_steps(ref_implicit, optimizers, Interactions, BilinearNet):
    # ---------------------------------------------------------------- D. step-level
    U, I, D, B = 120, 90, 16, 64
    cases = [
        # name, loss, optimizer, n_neg, l2, lr, n_pos, zipf
        ('pointwise_adam', 'pointwise', 'adam', 5, 1e-5, 1e-3, 64 * 12 + 40, True),
        ('pointwise_adam_wd0', 'pointwise', 'adam', 2, 0.0, 1e-2, 64 * 10 + 7, False),
        ('pointwise_sgd', 'pointwise', 'sgd', 3, 1e-4, 5e-2, 64 * 10 + 33, True),
        ('bpr_adam', 'bpr', 'adam', 1, 1e-5, 1e-3, 64 * 12 + 40, True),          # -> adaptive hinge (F2)
        ('adaptive_adam', 'adaptive_hinge', 'adam', 3, 1e-5, 1e-3, 64 * 12 + 9, False),
        ('hinge_adam', 'hinge', 'adam', 1, 1e-5, 1e-3, 64 * 12, True),           # full batches only
        ('hinge_sgd', 'hinge', 'sgd', 1, 0.0, 5e-2, 64 * 8, False),
        ('pointwise_rms', 'pointwise', 'rms', 3, 1e-5, 1e-3, 64 * 11 + 21, True),   # torch.optim.RMSprop (--optim rms)
        ('bpr_rms_wd0', 'bpr', 'rms', 1, 0.0, 1e-3, 64 * 10 + 5, False),
    ]
    for name, loss, opt, n_neg, l2, lr, n_pos, zipf in cases:
        if ONLY_STEPS and name not in ONLY_STEPS:
            continue
        rs = np.random.RandomState(100 + len(name))
        users, items = synth(rs, U, I, n_pos, zipf)
        neg_pairs = np.stack(synth(rs, U, I, n_pos), 1)
        torch.manual_seed(0)
        net = BilinearNet(U, I, D, sparse=False)
        init = tables_of(net)
        model = ref_implicit.ImplicitFactorizationModel(
            loss=loss, embedding_dim=D, n_iter=1, batch_size=B, l2=l2, learning_rate=lr,
            optimizer_func=getattr(optimizers, opt + '_optimizer'), representation=net,
            random_state=np.random.RandomState(0), neg_examples=[tuple(p) for p in neg_pairs.tolist()],
            num_negative_samples=n_neg, experiment_name='golden_' + name)
        train = Interactions(users.astype(np.int32), items.astype(np.int32), num_users=U, num_items=I)
        model._initialize(train)
        random.seed(7)
        tu, ti = torch.from_numpy(users), torch.from_numpy(items)
        step_losses = []
        for s in range(0, n_pos, B):
            step_losses.append(model.run_train_iteration(tu[s:s + B], ti[s:s + B]).item())
        mid = tables_of(model._net)
        val_losses = []
        for s in range(0, 3 * B, B):
            val_losses.append(model.run_val_iteration(tu[s:s + B], ti[s:s + B]).item())
        random.seed(7)
        n_draws = len(step_losses) + len(val_losses)
        neg_idx = np.array([random.choices(range(len(neg_pairs)), k=n_neg * B) for _ in range(n_draws)])
        d = dict(users=users, items=items, neg_pairs=neg_pairs, neg_idx=neg_idx,
                 step_losses=np.array(step_losses, dtype=np.float64),
                 val_losses=np.array(val_losses, dtype=np.float64),
                 meta=np.array([U, I, D, B, n_neg]), hyper=np.array([lr, l2]),
                 loss=np.array(loss), optimizer=np.array(opt), py_seed=np.array(7))
        for i in range(4):
            d['init%d' % i] = init[i]
            d['final%d' % i] = mid[i]
        np.savez_compressed(os.path.join(OUT, 'steps_%s.npz' % name), **d)
        print(name, step_losses[0], step_losses[-1])



def make_fits(ref_implicit, evaluation, optimizers, Interactions, BilinearNet):
    # ---------------------------------------------------------------- E. fit + predict + evaluate
    U, I, D, B = 150, 110, 16, 128
    for name, loss, n_neg in (('fit_pointwise', 'pointwise', 4), ('fit_bpr', 'bpr', 1)):
        rs = np.random.RandomState(42)
        n_all = 6000
        users, items = synth(rs, U, I, n_all, zipf=True)
        # make it learnable: users prefer items congruent to them
        items = np.where(rs.rand(n_all) < 0.7, (users * 7 + rs.randint(0, 6, n_all)) % I, items).astype(np.int64)
        a, b = int(0.81 * n_all), int(0.9 * n_all)
        mk = lambda u, i: Interactions(u.astype(np.int32), i.astype(np.int32), num_users=U, num_items=I)
        train, valid, test = mk(users[:a], items[:a]), mk(users[a:b], items[a:b]), mk(users[b:], items[b:])
        neg_pairs = np.stack(synth(rs, U, I, a), 1)
        torch.manual_seed(0)
        net = BilinearNet(U, I, D, sparse=False)
        init = tables_of(net)
        model = ref_implicit.ImplicitFactorizationModel(
            loss=loss, embedding_dim=D, n_iter=4, batch_size=B, l2=1e-5, learning_rate=2e-2,
            optimizer_func=optimizers.adam_optimizer, representation=net,
            random_state=np.random.RandomState(0), neg_examples=[tuple(p) for p in neg_pairs.tolist()],
            num_negative_samples=n_neg, experiment_name='golden_' + name)
        random.seed(3)
        model.fit(train, valid, verbose=False)
        final = tables_of(model._net)
        import csv
        with open(os.path.join(model.experiment_logs, 'summary.csv')) as f:
            rows = list(csv.reader(f))
        summary = np.array([[float(x) for x in r] for r in rows[1:]])
        d = dict(users=users, items=items, split=np.array([a, b]), neg_pairs=neg_pairs,
                 meta=np.array([U, I, D, B, n_neg, 4]), hyper=np.array([2e-2, 1e-5]), loss=np.array(loss),
                 py_seed=np.array(3), summary=summary, summary_header=np.array(rows[0]),
                 best_epoch=np.array(model.best_epoch), py_random_after=np.array(random.getstate()[1], dtype=np.int64))
        for i in range(4):
            d['init%d' % i] = init[i]
            d['final%d' % i] = final[i]
        d['predict_user3'] = model.predict(3)
        pu, pi = synth(np.random.RandomState(5), U, I, 500)
        d['predict_pairs_users'], d['predict_pairs_items'] = pu, pi
        d['predict_pairs'] = model.predict(pu, pi)
        for k in (5, 10, 20):
            p, r = evaluation.precision_recall_score(model, test, train=train, k=k)
            d['pr_masked_k%d' % k] = np.array([p, r])
            p, r = evaluation.precision_recall_score(model, test, k=k)
            d['pr_nomask_k%d' % k] = np.array([p, r])
        p, r = evaluation.precision_recall_score(model, test, train=train, k=np.array([5, 10, 20]))
        d['pr_masked_karray'] = np.array([p, r])
        np.savez_compressed(os.path.join(OUT, '%s.npz' % name), **d)
        print(name, summary[:, :2].tolist(), d['pr_masked_k10'])


if __name__ == '__main__':
    main()
