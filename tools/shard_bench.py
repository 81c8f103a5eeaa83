#!/usr/bin/env python
"""cfg5 (BASELINE.json configs[4]): synthetic scaled catalog, tables row-sharded over the GPUs of one box.

    python tools/shard_bench.py [--users 10000000 --items 2000000 --dim 128 --batch 65536 --steps 200 --warmup 20]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/shard_bench.py ...

Strong scaling: the global batch is fixed, every rank computes batch/N positives (+ negatives) per step and owns
1/N of the rows.  Prints one JSON object (rank 0): interactions/s (device-timed, max over ranks), per-phase
kernel time, exchange volume.  `bench.py` imports `run()` for its `sharded_train` object."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def run(users=10_000_000, items=2_000_000, dim=128, batch=65536, n_neg=1, steps=192, warmup=64, loss='bpr',
        zipf=False, chunk_steps=32, fast_math=True, pop_len=4_000_000, comm=None, phase_times=False, direct=True,
        per_chunk=False):
    import recommendation_gans_b200  # noqa: F401
    from recommendation_gans_b200 import sharded
    import torch.distributed as dist
    if comm is None:
        comm = sharded.DistComm() if dist.is_initialized() else sharded.LocalGroup(1).comm(0)
    rank, world = comm.rank, comm.world
    dev = torch.device('cuda', torch.cuda.current_device())
    be = sharded.CudaShardBackend(rank, world, users, items, dim, optimizer='adam', lr=1e-3, l2=1e-5,
                                  fast_math=fast_math, device=dev, seed=0)
    shard = sharded.ShardedMF(be, comm, chunk_steps=chunk_steps, direct=direct)
    rs = np.random.RandomState(0)                       # same ids on every rank
    n_pos = (steps + warmup) * batch

    def ids(n, hi):
        if zipf:
            p = 1.0 / np.arange(1, hi + 1) ** 1.05
            return rs.choice(hi, n, p=p / p.sum()).astype(np.int64)
        return rs.randint(0, hi, n).astype(np.int64)
    pos_u, pos_i = be.ids(rs.randint(0, users, n_pos)), be.ids(ids(n_pos, items))
    pop_u, pop_i = be.ids(rs.randint(0, users, pop_len)), be.ids(ids(pop_len, items))
    import random
    random.seed(0)
    state = np.array(random.getstate()[1], dtype=np.uint32)
    m = n_neg * batch
    neg_u, neg_i = be.draw_negatives(state, pop_u, pop_i, (steps + warmup) * m)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if dist.is_initialized():
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up (also sizes every workspace and the exchange buffers)
    shard.train_steps(loss, pos_u, pos_i, batch, n_neg, neg_u[:warmup * m], neg_i[:warmup * m], step0=0, nsteps=warmup)
    barrier()
    l0 = be.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    chunk_ms, marks, losses = [], [e0], []
    call_steps = chunk_steps if per_chunk else steps   # per_chunk: one call per chunk puts every chunk's device time on
    for c0 in range(0, steps, call_steps):             # record, but planning then no longer overlaps the previous chunk
        ns = min(call_steps, steps - c0)
        lo = (warmup + c0) * m
        losses.append(shard.train_steps(loss, pos_u, pos_i, batch, n_neg, neg_u[lo:lo + ns * m], neg_i[lo:lo + ns * m],
                                        step0=warmup + c0, nsteps=ns))
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        marks.append(ev)
    losses = np.concatenate(losses)
    e1.record()
    barrier()
    chunk_ms = [marks[i].elapsed_time(marks[i + 1]) for i in range(len(marks) - 1)]
    wall = time.perf_counter() - t0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist.is_initialized():
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    stride = be.stride
    rows_per_step = 2 * (batch + m)
    out = {
        'metric': 'train interactions/s (row-sharded scaled catalog, BPR)', 'unit': 'interactions/s',
        'value': steps * batch / (ms * 1e-3), 'ms_per_step': ms / steps, 'wall_ms_per_step': wall * 1e3 / steps,
        'n_gpus': world, 'steps': steps, 'warmup': warmup, 'scaling': 'strong',
        'config': {'workload': 'synthetic scaled catalog %dx%d dim %d, loss=%s->adaptive_hinge, global batch %d, '
                               'n_neg %d, Adam(0.5,0.999) lr 1e-3 l2 1e-5, rows sharded by id %% %d'
                               % (users, items, dim, loss, batch, n_neg, world),
                   'items': 'zipf(1.05)' if zipf else 'uniform', 'chunk_steps': chunk_steps,
                   'transport': type(comm).__name__,
                   'exchange': 'peer-memory stores + flag kernels (NVLink), no collective per step' if direct else
                               'all_to_all_single + all_reduce per step'},
        'exchange_bytes_per_step_per_gpu_each_way': rows_per_step / world * stride * 4 * (world - 1) / world,
        'algorithmic_bytes_per_step': (6 * 2 * (1 + n_neg) * (dim + 1) * 4 + 16) * batch,
        'gpu_launches': be.launches - l0, 'final_loss': float(losses[-1]),
        'chunk_ms': [round(x, 2) for x in chunk_ms] if per_chunk else None,
    }
    out['hbm_gbs_algorithmic_total'] = out['algorithmic_bytes_per_step'] / (out['ms_per_step'] * 1e-3) / 1e9
    if phase_times and not direct:
        out['phase_us_per_step'] = phase_breakdown(shard, loss, pos_u, pos_i, batch, n_neg, neg_u, neg_i, warmup, m)
    shard.close()
    return out


def phase_breakdown(shard, loss, pos_u, pos_i, batch, n_neg, neg_u, neg_i, warmup, m, nsteps=16):
    """Device time of each phase (CUDA events around every call of a few extra steps; serialises the stream)."""
    be, comm = shard.backend, shard.comm
    names = ['gather', 'a2a_rows', 'forward', 'allreduce', 'backward', 'a2a_grads', 'update']
    acc = {k: 0.0 for k in names}
    evs = []

    def timed(name, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((name, a, b))

    class TimedBackend(object):
        def __getattr__(self, k):
            return getattr(be, k)

        def gather(self, *a):
            timed('gather', lambda: be.gather(*a))

        def forward(self, *a):
            timed('forward', lambda: be.forward(*a))

        def backward(self, *a):
            timed('backward', lambda: be.backward(*a))

        def update(self, *a):
            timed('update', lambda: be.update(*a))

    class TimedComm(object):
        rank, world = comm.rank, comm.world
        n = 0

        def all_to_all(self, *a):
            self.n += 1
            timed('a2a_rows' if self.n % 2 else 'a2a_grads', lambda: comm.all_to_all(*a))

        def all_reduce_max(self, t):
            timed('allreduce', lambda: comm.all_reduce_max(t))

        def all_reduce_sum(self, t):
            comm.all_reduce_sum(t)

    shard.backend, shard.comm = TimedBackend(), TimedComm()
    try:
        shard.train_steps(loss, pos_u, pos_i, batch, n_neg, neg_u[:nsteps * m], neg_i[:nsteps * m], step0=0,
                          nsteps=nsteps)
        torch.cuda.synchronize()
    finally:
        shard.backend, shard.comm = be, comm
    for name, a, b in evs:
        acc[name] += a.elapsed_time(b) * 1e3
    return {k: v / nsteps for k, v in acc.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--users', type=int, default=10_000_000)
    ap.add_argument('--items', type=int, default=2_000_000)
    ap.add_argument('--dim', type=int, default=128)
    ap.add_argument('--batch', type=int, default=65536)
    ap.add_argument('--n-neg', type=int, default=1)
    ap.add_argument('--steps', type=int, default=192)
    ap.add_argument('--warmup', type=int, default=64)
    ap.add_argument('--chunk-steps', type=int, default=32)
    ap.add_argument('--zipf', action='store_true')
    ap.add_argument('--phases', action='store_true')
    ap.add_argument('--per-chunk', action='store_true', help='one call per chunk: per-chunk device times (no plan overlap)')
    ap.add_argument('--collectives', action='store_true', help='use all_to_all_single/all_reduce instead of peer memory')
    args = ap.parse_args()
    import torch.distributed as dist
    if 'RANK' in os.environ and int(os.environ.get('WORLD_SIZE', '1')) > 1:
        local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    out = run(args.users, args.items, args.dim, args.batch, args.n_neg, args.steps, args.warmup, zipf=args.zipf,
              chunk_steps=args.chunk_steps, phase_times=args.phases, direct=not args.collectives, per_chunk=args.per_chunk)
    if not dist.is_initialized() or dist.get_rank() == 0:
        print(json.dumps(out))
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
