"""N>1 host logic on CPU: two gloo ranks shard the evaluated users and all-reduce the hit sums; the result
must equal the single-process value (no GPU needed)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from recommendation_gans_b200.sharding import allreduce_precision_recall, shard_range


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, hits, ntargets, ks, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = shard_range(len(ntargets), rank, world)
    p, r, n = allreduce_precision_recall(hits[lo:hi], ntargets[lo:hi], ks, dist=dist)
    np.save(os.path.join(out_dir, 'r%d.npy' % rank), np.concatenate([p, r, [n]]))
    dist.destroy_process_group()


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 138493):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_allreduce_matches_single_process(tmp_path):
    rs = np.random.RandomState(0)
    n, ks = 1001, [5, 10, 20]
    ntargets = rs.randint(1, 30, n)
    hits = np.minimum(np.sort(rs.randint(0, 6, (n, 3)), axis=1), ntargets[:, None])
    p1, r1, n1 = allreduce_precision_recall(hits, ntargets, ks)
    np.testing.assert_allclose(p1, (hits / np.array(ks)[None, :]).mean(0))
    np.testing.assert_allclose(r1, (hits / ntargets[:, None]).mean(0))
    mp.spawn(_worker, args=(2, _free_port(), hits, ntargets, ks, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / 'r0.npy'), np.load(tmp_path / 'r1.npy')
    np.testing.assert_array_equal(a, b)                       # every rank holds the same reduced metrics
    np.testing.assert_allclose(a[:3], p1, rtol=1e-12)
    np.testing.assert_allclose(a[3:6], r1, rtol=1e-12)
    assert int(a[6]) == n == n1
