"""Config-level parity runs (BASELINE.json configs cfg1-cfg5) of the CUDA path against the oracle.

Helper module of tests/test_gpu_configs.py (and of tools/config_parity.py, which prints the same
diagnostics as JSON lines).  Every run returns a dict of measured deviations; the asserts live in
the test file.  Inputs are synthetic MovieLens-shaped ids: users uniform, items uniform
or Zipf(1.05); negative (user, item) pairs uniform; tables from `torch.manual_seed(0)` through the
reference's initialisers; Adam(0.5, 0.999), lr 1e-3, l2 1e-5 (mf_spotlight.py defaults).
"""
import os

import numpy as np
import torch

from oracle import mf_oracle as O

# BASELINE.json configs (SURVEY section 8 shorthand).  `loss` is what ImplicitFactorizationModel would be
# given; implicit.py:194-199 wires 'bpr' to adaptive_hinge_loss.
CONFIGS = {
    'cfg1': dict(U=943, I=1682, D=32, B=256, n_neg=1, loss='bpr', steps=317),            # one ML-100K epoch
    'cfg2': dict(U=6040, I=3706, D=64, B=1024, n_neg=5, loss='pointwise', steps=200),
    'cfg3': dict(U=138493, I=26744, D=128, B=8192, n_neg=1, loss='adaptive_hinge', steps=200),
    # cfg5 down-scaled to what the dense-optimiser oracle can step on the host (same D, same global batch)
    'cfg5s': dict(U=200000, I=40000, D=128, B=65536, n_neg=1, loss='bpr', steps=24),
}
LR, L2 = 1e-3, 1e-5


def loss_kind(name):
    return {'pointwise': 'pointwise', 'hinge': 'hinge'}.get(name, 'adaptive_hinge')      # implicit.py:194-199


def synth(cfg, zipf, seed=0):
    c = CONFIGS[cfg]
    rs = np.random.RandomState(seed)
    n = c['steps'] * c['B']
    users = rs.randint(0, c['U'], n).astype(np.int64)
    if zipf:
        p = 1.0 / np.arange(1, c['I'] + 1) ** 1.05
        items = rs.choice(c['I'], n, p=p / p.sum()).astype(np.int64)
    else:
        items = rs.randint(0, c['I'], n).astype(np.int64)
    m = c['steps'] * c['n_neg'] * c['B']
    neg_u = rs.randint(0, c['U'], m).astype(np.int64)
    neg_i = rs.randint(0, c['I'], m).astype(np.int64)
    return users, items, neg_u, neg_i


_ORACLE_CACHE = {}


def oracle_run(cfg, zipf, dtype=torch.float32):
    """The reference's step (oracle restatement: same torch CPU ops, dense torch.optim.Adam) over the config's
    steps.  Cached per (cfg, zipf, dtype): both arithmetic modes of the CUDA path are compared with the same run.
    dtype=float64 runs the identical algorithm in double precision: the yardstick for the fp32 reference's own
    rounding envelope (see `table_report`)."""
    key = (cfg, bool(zipf), dtype)
    if key in _ORACLE_CACHE:
        return _ORACLE_CACHE[key]
    c = CONFIGS[cfg]
    users, items, neg_u, neg_i = synth(cfg, zipf)
    init = [t.numpy().copy() for t in O.init_tables(c['U'], c['I'], c['D'], torch_seed=0)]
    model = O.OracleMF(*[torch.from_numpy(t).to(dtype) for t in init], loss=c['loss'], optimizer='adam', lr=LR, l2=L2,
                       batch_size=c['B'], num_negative_samples=c['n_neg'])
    if dtype == torch.float64:
        # the yardstick run only has to be (much) more accurate than fp32: torch's fused CPU Adam applies the same
        # update rule 2.4x faster (its own rounding sits at 1e-16); the fp32 run keeps the reference's default path
        model.opt = torch.optim.Adam(model.tables, lr=LR, betas=(0.5, 0.999), eps=1e-8, weight_decay=L2, fused=True)
    B, k = c['B'], c['n_neg'] * c['B']
    tu, ti = torch.from_numpy(users), torch.from_numpy(items)
    nu, ni = torch.from_numpy(neg_u), torch.from_numpy(neg_i)
    losses = [model.train_step(tu[s * B:(s + 1) * B], ti[s * B:(s + 1) * B], nu[s * k:(s + 1) * k],
                               ni[s * k:(s + 1) * k]).item() for s in range(c['steps'])]
    out = dict(init=init, losses=np.array(losses, dtype=np.float64), tables=model.numpy_tables(),
               inputs=(users, items, neg_u, neg_i))
    _ORACLE_CACHE[key] = out
    return out


def _rel(a, b):
    diff = np.abs(a.astype(np.float64) - b.astype(np.float64))
    scale = max(float(np.abs(b).max()), 1e-30)
    return diff, scale


def table_report(got, cfg, zipf, lr):
    """Deviation of the four tables from the fp32 reference run: `rel` is the max-norm relative error the 1e-5 bar
    is stated on, `hatch` counts the elements outside 1e-5 * max|ref|, `abs_lr` is the largest deviation in units
    of the learning rate.

    Where a table misses 1e-5 the report adds the rounding envelope of the reference itself: the same algorithm run
    in float64 (`rel64` = our deviation from it, `env` = the fp32 reference's deviation from it, `env_hatch` = how
    many of ITS elements are outside 1e-5).  Long duplicate segments (Zipf ids: one item row receives ~1000 gradient
    rows per step) make fp32 summation order visible after Adam's normalisation: the fp32 reference moves by
    5e-5 / 1e-4 when the batch is merely permuted (tools/fp32_envelope.py, profiles/r02_fp32_envelope.txt), so
    no implementation with its own summation order can be closer to one fp32 run than that run is to the exact
    arithmetic."""
    ref = oracle_run(cfg, zipf)['tables']
    rep = []
    for i, (t, r) in enumerate(zip(got, ref)):
        diff, scale = _rel(t, r)
        rep.append(dict(rel=float(diff.max() / scale), hatch=int((diff > 1e-5 * scale).sum()), size=int(diff.size),
                        abs_lr=float(diff.max() / lr)))
    if any(t['rel'] >= 1e-5 for t in rep):
        ref64 = oracle_run(cfg, zipf, torch.float64)['tables']
        for t, g, r32, r64 in zip(rep, got, ref, ref64):
            d_ours, scale = _rel(g, r64)
            d_ref, _ = _rel(r32, r64)
            t.update(rel64=float(d_ours.max() / scale), env=float(d_ref.max() / scale),
                     env_hatch=int((d_ref > 1e-5 * scale).sum()), hatch64=int((d_ours > 1e-5 * scale).sum()))
    return rep


def longest_segment(items, neg_i, B, n_neg):
    """Largest number of slots of one step that address the same item row (the ordered segment reduction
    splits a segment into windows of 32 and combines them through tickets)."""
    best = 0
    k = n_neg * B
    for s in range(0, min(len(items) // B, 8)):
        ids = np.concatenate([items[s * B:(s + 1) * B], neg_i[s * k:(s + 1) * k]])
        best = max(best, int(np.bincount(ids).max()))
    return best


def native_run(cfg, zipf, fast_math):
    from tests.gpu_helpers import make_engine, tables_of
    c = CONFIGS[cfg]
    ref = oracle_run(cfg, zipf)
    users, items, neg_u, neg_i = ref['inputs']
    net, _, eng = make_engine(ref['init'], 'adam', LR, L2, fast_math)
    losses = eng.train_steps(loss_kind(c['loss']), users, items, c['B'], c['n_neg'], neg_u, neg_i).cpu().numpy()
    eng.flush()
    torch.cuda.synchronize()
    tables = tables_of(net)
    loss_rel = float(np.max(np.abs(losses - ref['losses']) / np.abs(ref['losses'])))
    return dict(cfg=cfg, items='zipf' if zipf else 'uniform', fast_math=bool(fast_math), steps=c['steps'],
                loss_rel=loss_rel, tables=table_report(tables, cfg, zipf, LR),
                longest_segment=longest_segment(items, neg_i, c['B'], c['n_neg'])), net, eng


def sharded_run(cfg, zipf, world, fast_math, direct=True, chunk_steps=8):
    """cfg5 (down-scaled): G virtual ranks on one GPU, rows g % G == r per rank."""
    from recommendation_gans_b200 import sharded
    c = CONFIGS[cfg]
    ref = oracle_run(cfg, zipf)
    users, items, neg_u, neg_i = ref['inputs']

    def make(rank, comm):
        be = sharded.CudaShardBackend(rank, world, c['U'], c['I'], c['D'],
                                      local_tables=sharded.slice_tables(ref['init'], rank, world), optimizer='adam',
                                      lr=LR, l2=L2, fast_math=fast_math)
        return sharded.ShardedMF(be, comm, chunk_steps=chunk_steps, direct=direct)

    def work(shard):
        losses = shard.train_steps(loss_kind(c['loss']), users, items, c['B'], c['n_neg'], neg_u, neg_i)
        tables = shard.local_tables()
        torch.cuda.synchronize()
        shard.close()
        return losses, tables

    results = sharded.run_local_ranks(world, make, work)
    tables = sharded.assemble_tables([r[1] for r in results], world)
    loss_rel = max(float(np.max(np.abs(np.asarray(l) - ref['losses']) / np.abs(ref['losses']))) for l, _ in results)
    return dict(cfg=cfg, items='zipf' if zipf else 'uniform', fast_math=bool(fast_math), world=world,
                direct=bool(direct), steps=c['steps'], loss_rel=loss_rel,
                tables=table_report(tables, cfg, zipf, LR))


def train_csr(cfg, seed=1, per_user=117):
    """Synthetic train interactions for the cfg4 mask: ~117 per user (ML-20M), Zipf items."""
    import scipy.sparse as sp
    c = CONFIGS[cfg]
    rs = np.random.RandomState(seed)
    n = per_user * c['U']
    p = 1.0 / np.arange(1, c['I'] + 1) ** 1.05
    tu = rs.randint(0, c['U'], n)
    ti = rs.choice(c['I'], n, p=p / p.sum())
    csr = sp.coo_matrix((np.ones(n), (tu, ti)), shape=(c['U'], c['I'])).tocsr()
    csr.sum_duplicates()
    csr.sort_indices()
    return csr


def topk_report(net, cfg, k=20, sample=500, seed=2):
    """cfg4 on a TRAINED table: (1) tensor-core path == exact fp32 kernel for every user (ids and scores);
    (2) ids vs the stable-ranking oracle computed in float64 from the same tables for `sample` users, compared at
    every rank whose float64 score is separated from both neighbours by more than the fp32 summation bound."""
    from recommendation_gans_b200.engine import MFEngine
    from tests.gpu_helpers import tables_of
    c = CONFIGS[cfg]
    U, I, D = c['U'], c['I'], c['D']
    csr = train_csr(cfg)
    indptr = torch.from_numpy(csr.indptr.astype(np.int64)).cuda()
    indices = torch.from_numpy(csr.indices.astype(np.int32)).cuda()
    users = np.arange(U, dtype=np.int64)
    old = os.environ.get('MFB_TC')
    try:
        os.environ['MFB_TC'] = '0'
        exact = MFEngine(net)
        os.environ['MFB_TC'] = '1'
        tc = MFEngine(net)
    finally:
        if old is None:
            os.environ.pop('MFB_TC', None)
        else:
            os.environ['MFB_TC'] = old
    out = dict(cfg=cfg, users=U, k=k)
    for masked in (True, False):
        args = (indptr, indices) if masked else (None, None)
        ids_e, sc_e = exact.topk(users, k, *args, with_scores=True)
        ids_t, sc_t = tc.topk(users, k, *args, with_scores=True)
        tag = 'masked' if masked else 'nomask'
        out['tc_vs_exact_id_mismatch_' + tag] = int((ids_e != ids_t).sum().item())
        out['tc_vs_exact_score_mismatch_' + tag] = int((sc_e != sc_t).sum().item())
        out['tc_redo_' + tag] = tc.topk_last_redo
        if masked:
            got = ids_t.cpu().numpy()
    ue, ie, ub, ib = [t.astype(np.float64) for t in tables_of(net)]
    rs = np.random.RandomState(seed)
    picked = rs.choice(U, sample, replace=False)
    checked = wrong = 0
    for u in picked:
        z = ie @ ue[u] + ub[u, 0] + ib[:, 0]
        bound = 2.0 * D * 2.0 ** -24 * float((np.abs(ie) @ np.abs(ue[u])).max() + np.abs(ub[u, 0]) + np.abs(ib).max())
        rated = csr.indices[csr.indptr[u]:csr.indptr[u + 1]]
        key = -z
        key[rated] = np.inf
        order = np.argsort(key, kind='stable')[:k + 1]
        zs = z[order]
        for r in range(k):
            if zs[r] - zs[r + 1] > bound and (r == 0 or zs[r - 1] - zs[r] > bound):
                checked += 1
                wrong += int(got[u, r] != order[r])
    out.update(oracle_users=int(sample), oracle_ranks_checked=checked, oracle_ranks_wrong=wrong)
    return out
