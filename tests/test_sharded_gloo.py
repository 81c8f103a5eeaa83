"""N>1 host logic of the row-sharded training step on CPU: two gloo ranks run the real orchestration
(`ShardedMF.train_steps`) and the real transport (`DistComm`: all_to_all_single with the planned split sizes,
all_reduce MAX/SUM) over the numpy specification backend; the assembled tables and per-step losses must equal the
reference's single-process run frozen in tests/golden/steps_*.npz (1e-5 relative)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from recommendation_gans_b200 import sharded
from tests.shard_spec_backend import SpecShardBackend

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def _case(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        g = {k: z[k] for k in z.files}
    U, I, D, B, n_neg = [int(x) for x in g['meta']]
    n_steps = len(g['step_losses'])
    negs = g['neg_pairs'][g['neg_idx']]
    return g, dict(U=U, I=I, D=D, B=B, n_neg=n_neg, lr=float(g['hyper'][0]), l2=float(g['hyper'][1]),
                   loss=str(g['loss']), opt=str(g['optimizer']),
                   nu=negs[:n_steps, :, 0].reshape(-1).copy(), ni=negs[:n_steps, :, 1].reshape(-1).copy())


def _run_rank(rank, world, comm, name, chunk):
    g, c = _case(name)
    init = [g['init%d' % i] for i in range(4)]
    be = SpecShardBackend(rank, world, c['U'], c['I'], c['D'], sharded.slice_tables(init, rank, world), c['opt'],
                          c['lr'], c['l2'])
    shard = sharded.ShardedMF(be, comm, chunk_steps=chunk)
    losses = shard.train_steps(c['loss'], g['users'], g['items'], c['B'], c['n_neg'], c['nu'], c['ni'])
    return losses, shard.local_tables()


def _worker(rank, world, port, name, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    losses, tables = _run_rank(rank, world, sharded.DistComm(), name, 4)
    np.savez(os.path.join(out_dir, 'r%d.npz' % rank), losses=losses, t0=tables[0], t1=tables[1], t2=tables[2],
             t3=tables[3])
    dist.destroy_process_group()


def _check(name, results, world):
    g, c = _case(name)
    for losses, _ in results:
        np.testing.assert_allclose(losses, g['step_losses'], rtol=1e-5)
    tables = sharded.assemble_tables([r[1] for r in results], world)
    for i, t in enumerate(tables):
        ref = g['final%d' % i].reshape(t.shape)
        err = np.abs(t - ref).max() / max(np.abs(ref).max(), 1e-30)
        assert err < 1e-5 or (i >= 2 and np.abs(t - ref).max() < 2e-2 * c['lr']), (i, err)


@pytest.mark.parametrize('name', ['steps_bpr_adam.npz', 'steps_pointwise_adam.npz', 'steps_hinge_sgd.npz',
                                  'steps_pointwise_rms.npz'])
def test_two_gloo_ranks_match_reference(name, tmp_path):
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, name, str(tmp_path)), nprocs=2, join=True)
    res = []
    for r in range(2):
        z = np.load(tmp_path / ('r%d.npz' % r))
        res.append((z['losses'], [z['t%d' % k] for k in range(4)]))
    np.testing.assert_array_equal(res[0][0], res[1][0])      # every rank reports the same losses
    _check(name, res, 2)


@pytest.mark.parametrize('world', [1, 3])
def test_virtual_ranks_match_reference(world):
    """LocalComm (threads) transport with the same orchestration, incl. a world size that does not divide the batch."""
    name = 'steps_adaptive_adam.npz'
    res = sharded.run_local_ranks(world, lambda rank, comm: (rank, comm),
                                  lambda rc: _run_rank(rc[0], world, rc[1], name, 64))
    _check(name, res, world)


def test_assemble_inverts_slice():
    rs = np.random.RandomState(0)
    full = [rs.rand(11, 4), rs.rand(7, 4), rs.rand(11, 1), rs.rand(7, 1)]
    for world in (1, 2, 3, 8):
        parts = [sharded.slice_tables(full, r, world) for r in range(world)]
        assert [p[0].shape[0] for p in parts] == [sharded.local_rows(11, r, world) for r in range(world)]
        for a, b in zip(sharded.assemble_tables(parts, world), full):
            np.testing.assert_array_equal(a, b)


def test_more_ranks_than_rows_is_rejected_before_any_cuda_work():
    with pytest.raises(ValueError):
        sharded.CudaShardBackend.__init__(object.__new__(sharded.CudaShardBackend), 3, 4, 3, 100, 8)
