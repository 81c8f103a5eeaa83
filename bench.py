#!/usr/bin/env python
"""Benchmark of the implicit-MF hot path (BASELINE.json metric: train interactions/s on the
ML-20M-shape BPR MF config + top-k eval users/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

One JSON line on stdout (rank 0).  A "step" is one training minibatch (B = 8192 positive
interactions + their sampled negative pairs): negative draw (device MT19937, random.choices
semantics), forward, loss, backward, dense-semantics Adam -- everything
ImplicitFactorizationModel.run_train_iteration does (implicit.py:347-364).

  value     device-timed throughput, ids already resident in HBM (CUDA events around K steps);
            the timed region is repeated and the MEDIAN is reported (min/max under `timing`)
  e2e       same K steps through the host-buffer entry point (pinned HOST ids -> H2D -> steps ->
            D2H of the per-step losses), wall clock, median
  roofline  HBM roofline of the dominant kernel (k_update); roofline_step: of the whole step
  eval      cfg4: full-catalog top-k (k=20, train mask), users sharded over the ranks; `e2e` is
            spotlight.evaluation.precision_recall_score on HOST Interactions (CSR upload, top-k,
            hits, D2H inside the clock); `cpu_baseline` is the reference's per-user loop
            (evaluation.py:155-180) on a user sample
  zipf      the same train/eval measurements with Zipf(1.05) item ids
  cpu_baseline / --impl reference: the oracle port of the reference's torch-CPU step, timed on the
            host cores (the reference itself is pure Python on torch and is not shipped to the GPU box)
N > 1: training does not shard without changing the math ("replicas only"), so every rank
trains an independent replica; evaluation shards users; `sharded_train` row-shards the cfg5 tables.
"""
import argparse
import json
import os
import random
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[2]: ML-20M shape, dim 128, batch 8192, loss 'bpr' (which the reference wires to
    # adaptive_hinge_loss, implicit.py:194-199), 1 negative pair per positive, Adam(0.5, 0.999), lr 1e-3, l2 1e-5
    'cfg3': dict(name='ML-20M-shape implicit MF (138493x26744, dim 128), loss=bpr->adaptive_hinge, batch 8192, '
                      'n_neg 1, Adam(0.5,0.999) lr 1e-3 l2 1e-5',
                 U=138493, I=26744, D=128, B=8192, n_neg=1, loss='bpr', lr=1e-3, l2=1e-5, n_train=16200213),
    'cfg2': dict(name='ML-1M-shape implicit MF (6040x3706, dim 64), loss=pointwise, batch 1024, n_neg 5, Adam',
                 U=6040, I=3706, D=64, B=1024, n_neg=5, loss='pointwise', lr=1e-3, l2=1e-5, n_train=810169),
    'cfg1': dict(name='ML-100K-shape implicit MF (943x1682, dim 32), loss=bpr->adaptive_hinge, batch 256, n_neg 1',
                 U=943, I=1682, D=32, B=256, n_neg=1, loss='bpr', lr=1e-3, l2=1e-5, n_train=81000),
}
L2_POLICY = 'inputs larger than L2 (tables + Adam state 256 MB, rows gathered at random)'


def loss_kind(name):
    return {'pointwise': 'pointwise', 'hinge': 'hinge'}.get(name, 'adaptive_hinge')   # implicit.py:194-199


def config_of(w, zipf):
    """The `config` object: identical in the native and the reference arm."""
    return {'workload': w['name'], 'items': 'zipf(1.05)' if zipf else 'uniform',
            'negative_population': w['n_train'], 'l2_policy': L2_POLICY}


def algorithmic_bytes_per_interaction(D, n_neg, adam=True):
    """rows per positive r = 2(1+n) (pairs mode), each row (D+1) fp32 read p,m,v and
    written p,m,v under Adam, plus the positive's two int64 ids."""
    r = 2 * (1 + n_neg)
    return (6 if adam else 2) * r * (D + 1) * 4 + 16


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p['hbm_gbs'], bf16_tflops=p['bf16_tflops'], source='measured')
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, source='fallback')    # B200_PROFILING.md fallback


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {'hw_slowdown': getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8),
                 'hw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40),
                 'sw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20),
                 'sw_power_cap': getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4),
                 'hw_power_brake': getattr(nv, 'nvmlClocksEventReasonHwPowerBrakeSlowdown', 0x80)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.025)      # NVML queries contend with kernel launches for driver locks: poll sparsely

    def summary(self):
        self.stop_flag = True
        self.join(timeout=1.0)
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(self.samples)}


def synth_ids(rs, n, hi, zipf):
    if zipf:
        p = 1.0 / np.arange(1, hi + 1) ** 1.05
        return rs.choice(hi, n, p=p / p.sum()).astype(np.int64)
    return rs.randint(0, hi, n).astype(np.int64)


def stats(xs):
    xs = sorted(float(x) for x in xs)
    return dict(median=float(np.median(xs)), min=xs[0], max=xs[-1], n=len(xs))


# ------------------------------------------------------------------------------------------------
# CPU arm: oracle port of the reference step (torch CPU ops, all host threads)
# ------------------------------------------------------------------------------------------------
def cpu_reference_steps(w, steps, warmup, seed=0, max_seconds=150.0):
    """run_train_iteration (implicit.py:347-364) as the reference executes it on the CPU: python random.choices over
    the list of (user, item) tuples (len(neg_examples) == len(train), data_provider.py:81), zip, np.array,
    from_numpy, two forwards, loss, backward, dense torch.optim.Adam."""
    import torch
    from oracle import mf_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    rs = np.random.RandomState(seed)
    tabs = O.init_tables(w['U'], w['I'], w['D'], torch_seed=0)
    B, n_neg, pop = w['B'], w['n_neg'], w['n_train']
    model = O.OracleMF(*tabs, loss=w['loss'], optimizer='adam', lr=w['lr'], l2=w['l2'], batch_size=B,
                       num_negative_samples=n_neg)
    neg_list = list(zip(rs.randint(0, w['U'], pop).tolist(), rs.randint(0, w['I'], pop).tolist()))
    rng = random.Random(0)
    users = torch.from_numpy(synth_ids(rs, (warmup + steps) * B, w['U'], False))
    items = torch.from_numpy(synth_ids(rs, (warmup + steps) * B, w['I'], w['zipf']))

    def one(s):
        nu, ni = zip(*rng.choices(neg_list, k=n_neg * B))          # implicit.py:352-354
        nu, ni = torch.from_numpy(np.array(nu)).long(), torch.from_numpy(np.array(ni)).long()
        return model.train_step(users[s * B:(s + 1) * B], items[s * B:(s + 1) * B], nu, ni)

    for s in range(warmup):
        one(s)
    t0 = time.perf_counter()
    done = 0
    for s in range(warmup, warmup + steps):
        one(s)
        done += 1
        if time.perf_counter() - t0 > max_seconds:
            break
    dt = time.perf_counter() - t0
    return dict(value=done * B / dt, steps=done, seconds=dt, cores=torch.get_num_threads(),
                sample='%d full minibatch steps (B=%d) of the same workload after %d warm-up steps, negative '
                       'population %d pairs' % (done, B, warmup, pop))


def cpu_reference_eval(tables, test_csr, train_csr, n_users=200):
    """precision_recall_score's per-user loop (evaluation.py:155-180: predict -> mask -> argsort -> set
    intersections) on the first n_users rows, through the oracle's literal restatement (ranking='reference')."""
    import torch
    from oracle import mf_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    model = O.OracleMF(*[torch.from_numpy(np.ascontiguousarray(t)) for t in tables])
    sub_test, sub_train = test_csr[:n_users], train_csr[:n_users]
    evaluated = int((np.diff(sub_test.indptr) > 0).sum())
    O.precision_recall_score(model, sub_test[:8], sub_train[:8], k=np.array([5, 10, 20]), ranking='reference')
    t0 = time.perf_counter()
    O.precision_recall_score(model, sub_test, sub_train, k=np.array([5, 10, 20]), ranking='reference')
    dt = time.perf_counter() - t0
    return dict(value=evaluated / dt, unit='users/s', cores=torch.get_num_threads(), kind='port',
                sample='%d users with test items (first %d rows), full catalog, train mask, k=[5,10,20]'
                       % (evaluated, n_users))


# ------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------
class _EvalModel(object):
    """What spotlight.evaluation needs from a fitted ImplicitFactorizationModel."""

    def __init__(self, net, eng, num_items):
        self._net, self._engine_, self._num_items = net, eng, num_items


def train_measure(args, w, zipf, rank, dev, dist, lib):
    """K timed steps (device clock, median over repeats), the same steps end to end from pinned host ids, and the
    per-kernel device times of a profiled pass.  Returns (measurements, net, engine)."""
    import torch
    from recommendation_gans_b200.engine import MFEngine
    from spotlight.factorization.representations import BilinearNet
    import spotlight.optimizers as optimizers
    K, W = args.steps, args.warmup
    B, n_neg, D = w['B'], w['n_neg'], w['D']
    kind = loss_kind(w['loss'])
    R = max(args.repeats, 1)
    rs = np.random.RandomState(rank)
    torch.manual_seed(rank)
    net = BilinearNet(w['U'], w['I'], D).cuda()
    opt = optimizers.adam_optimizer(net.parameters(), lr=w['lr'], weight_decay=w['l2'])
    eng = MFEngine(net, opt, fast_math=bool(args.fast_math))
    n_steps_total = W + K * R
    users_h = synth_ids(rs, n_steps_total * B, w['U'], False)
    items_h = synth_ids(rs, n_steps_total * B, w['I'], zipf)
    pop_u = torch.from_numpy(synth_ids(rs, w['n_train'], w['U'], False)).to(dev)
    pop_i = torch.from_numpy(synth_ids(rs, w['n_train'], w['I'], False)).to(dev)
    users_d, items_d = torch.from_numpy(users_h).to(dev), torch.from_numpy(items_h).to(dev)
    rng = random.Random(rank)
    eng.rng_seed(rng)

    def run(lo, hi):
        # negatives are drawn on the device from the model's MT19937 stream, chunk by chunk (part of the step)
        return eng.train_epoch(kind, users_d[lo * B:hi * B], items_d[lo * B:hi * B], B, n_neg, pop_u, pop_i)

    run(0, W)                                   # warm-up (>= 3 steps)
    eng.rng_sync(rng)
    clocks = ClockSampler(dev.index)
    clocks.start()
    times, launches, final_loss = [], 0, float('nan')
    for rep in range(R):                        # fresh K steps each time: training simply continues
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        launches0 = eng.launches + lib.mfb_library_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        losses = run(W + rep * K, W + (rep + 1) * K)
        e1.record()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        times.append(e0.elapsed_time(e1))
        launches = eng.launches + lib.mfb_library_launches() - launches0
        final_loss = float(losses[-1].item())
    eng.rng_sync(rng)
    clk = clocks.summary()

    # ---- end to end: pinned host ids -> H2D -> steps -> D2H losses, through the host-buffer entry point
    pu = torch.from_numpy(users_h[W * B:(W + K) * B]).pin_memory()
    pi = torch.from_numpy(items_h[W * B:(W + K) * B]).pin_memory()
    eng.train_epoch_host(kind, pu.numpy()[:4 * B], pi.numpy()[:4 * B], B, n_neg, pop_u, pop_i, rng=rng)   # warm
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    e2e = []
    for rep in range(R):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        host_losses = eng.train_epoch_host(kind, pu.numpy(), pi.numpy(), B, n_neg, pop_u, pop_i, rng=rng)
        torch.cuda.synchronize()
        e2e.append(time.perf_counter() - t0)
        assert np.isfinite(host_losses).all()

    # ---- per-kernel device time (separate profiled pass; events bracket every launch)
    P = min(K, 256)
    eng.profile(True)
    eng.train_epoch_host(kind, pu.numpy()[:P * B], pi.numpy()[:P * B], B, n_neg, pop_u, pop_i, rng=rng)
    prof = eng.profile_read()
    eng.profile(False)
    return dict(ms=stats(times), e2e_s=stats(e2e), launches=launches, final_loss=final_loss, clocks=clk, prof=prof,
                prof_steps=P), net, eng


def eval_measure(args, eng, net, w, zipf, rank, world, dev, dist):
    """cfg4: precision/recall@k over the users with test items (train mask on), users sharded over the ranks."""
    import torch
    import scipy.sparse as sp
    from recommendation_gans_b200.sharding import shard_range
    from spotlight.interactions import Interactions
    from spotlight import evaluation as E
    U, I = w['U'], w['I']
    rs = np.random.RandomState(1234)                     # same interactions on every rank
    n_tr, n_te = 117 * U, 14 * U                          # ML-20M: ~117 train / ~14 test interactions per user
    train = Interactions(rs.randint(0, U, n_tr).astype(np.int32), synth_ids(rs, n_tr, I, zipf).astype(np.int32),
                         num_users=U, num_items=I)
    test = Interactions(rs.randint(0, U, n_te).astype(np.int32), synth_ids(rs, n_te, I, zipf).astype(np.int32),
                        num_users=U, num_items=I)
    train_csr, test_csr = train.csr_matrix, test.csr_matrix      # built once by Interactions (interactions.py:115)
    for c in (train_csr, test_csr):
        c.sum_duplicates()
        c.sort_indices()
    all_users = np.nonzero(np.diff(test_csr.indptr))[0].astype(np.int64)
    lo, hi = shard_range(len(all_users), rank, world)
    indptr = torch.from_numpy(train_csr.indptr.astype(np.int64)).to(dev)
    indices = torch.from_numpy(train_csr.indices.astype(np.int32)).to(dev)
    users = torch.from_numpy(all_users[lo:hi]).to(dev)
    eng.topk(users, 20, indptr, indices)                 # warm-up pass (also brings the clocks back up)
    # Passes over the SAME (user list, train CSR), as model.test() and per-epoch validation make them: the pair is named
    # by a plan key, so the model-independent train-mask images are built by the first pass only (mfb_topk_keyed);
    # `first_call` is that pass, `unkeyed` a pass that rebuilds them every time (mfb_topk).
    def timed(**kw):
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.topk(users, 20, indptr, indices, **kw)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 1e3
    first_call = timed(plan_key=0x5eed0001 + rank)
    times = [timed(plan_key=0x5eed0001 + rank) for rep in range(max(args.repeats, 1))]
    unkeyed = float(np.median([timed() for rep in range(3)]))
    redo = eng.topk_last_redo
    # end to end through the reference-facing call on HOST Interactions: CSR upload, top-k, hit counts, D2H
    model = _EvalModel(net, eng, I)
    fn = E.precision_recall_score_sharded if world > 1 else E.precision_recall_score
    ks = np.array([5, 10, 20])
    e2e, pr = [], None
    devnull = open(os.devnull, 'w')
    for rep in range(1 + max(args.repeats // 2, 3)):
        for c in (train_csr, test_csr):                  # drop the device copies: the upload is part of the call
            if hasattr(c, '_mfb_device'):
                del c._mfb_device
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        so, sys.stdout = sys.stdout, devnull             # the reference prints "Cold start users"
        try:
            pr = fn(model, test, train=train, k=ks)
        finally:
            sys.stdout = so
        torch.cuda.synchronize()
        if rep:                                          # first call warms the allocator
            e2e.append(time.perf_counter() - t0)
    kern = ('k_tc_gemm (TMA + tcgen05 kind::f16 with fp16 operands, fp32 TMEM accumulators) + exact fp32 re-score; %d users redone by k_topk_exact'
            % redo if os.environ.get('MFB_TC', '1') != '0' else 'k_topk_exact (fp32 CUDA cores)')
    h2d = (train_csr.indptr.size + test_csr.indptr.size) * 8 + (train_csr.nnz + test_csr.nnz) * 4 + (hi - lo) * 8
    out = dict(seconds=stats(times), first_call=first_call, unkeyed=unkeyed, users=hi - lo, users_total=len(all_users),
               kernel=kern, e2e_s=stats(e2e),
               precision_recall=[float(pr[0]), float(pr[1])], h2d_bytes=int(h2d), d2h_bytes=int((hi - lo) * 4 * 4))
    return out, (train_csr, test_csr)


def reduce_max(dist, dev, values):
    import torch
    if not dist:
        return [float(v) for v in values]
    t = torch.tensor(list(values), device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def gather_ranks(dist, dev, value):
    """[value of rank 0, value of rank 1, ...] on every rank."""
    import torch
    if not dist:
        return [float(value)]
    t = torch.tensor([float(value)], device=dev, dtype=torch.float64)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [float(x.item()) for x in out]


def reduce_sum(dist, dev, values):
    import torch
    if not dist:
        return [float(v) for v in values]
    t = torch.tensor(list(values), device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()]


def native_bench(args, w, rank, world):
    import torch
    import recommendation_gans_b200  # noqa: F401
    from recommendation_gans_b200 import _native as N

    local_rank = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    lib = N.load_library()
    K, W = args.steps, args.warmup
    B, n_neg, D = w['B'], w['n_neg'], w['D']
    peaks = measured_peaks()
    alg_b = algorithmic_bytes_per_interaction(D, n_neg)
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get('k_update_dram_bytes_per_launch')

    def one_distribution(zipf, with_cpu):
        tm, net, eng = train_measure(args, w, zipf, rank, dev, dist, lib)
        ev, csrs = eval_measure(args, eng, net, w, zipf, rank, world, dev, dist)
        t_train, t_train_min, t_train_max, t_e2e, t_eval, t_eval_e2e = reduce_max(
            dist, dev, [tm['ms']['median'] / 1e3, tm['ms']['min'] / 1e3, tm['ms']['max'] / 1e3, tm['e2e_s']['median'],
                        ev['seconds']['median'], ev['e2e_s']['median']])
        n_eval, launches = [int(x) for x in reduce_sum(dist, dev, [ev['users'], tm['launches']])]
        eval_rank_s = gather_ranks(dist, dev, ev['seconds']['median'])
        prof, P = tm['prof'], tm['prof_steps']
        upd_ms, upd_n = prof.get('update', (0.0, 1))
        upd_us = upd_ms * 1e3 / max(upd_n, 1)
        upd_bytes = (alg_b - 16) * B                  # every row's p,m,v read + p,m,v write happens in k_update
        step_bytes = alg_b * B
        flops = 2.0 * n_eval * w['I'] * D
        res = {
            'value': world * K * B / t_train, 'ms_per_step': t_train * 1e3 / K,
            'timing': {'statistic': 'median of %d timed regions of K steps each (max over ranks)' % tm['ms']['n'],
                       'ms_per_step_min': t_train_min * 1e3 / K, 'ms_per_step_max': t_train_max * 1e3 / K},
            'clocks': tm['clocks'], 'final_loss': tm['final_loss'],
            'e2e': {'value': world * K * B / t_e2e, 'unit': 'interactions/s', 'h2d_bytes_per_step': 16 * B,
                    'd2h_bytes_per_step': 4},
            'gpu_launches': launches,
            'roofline': {'bound': 'hbm', 'kernel': 'k_update',
                         'achieved': upd_bytes / (upd_us * 1e-6) / 1e9 if upd_us else None,
                         'peak': peaks['hbm_gbs'], 'peak_source': peaks['source'], 'unit': 'GB/s',
                         'frac': (upd_bytes / (upd_us * 1e-6) / 1e9 / peaks['hbm_gbs']) if upd_us else None,
                         'traffic': traffic, 'us_per_launch': upd_us, 'algorithmic_bytes_per_launch': upd_bytes},
            'roofline_step': {'bound': 'hbm', 'achieved': step_bytes / (t_train / K) / 1e9, 'peak': peaks['hbm_gbs'],
                              'unit': 'GB/s', 'frac': step_bytes / (t_train / K) / 1e9 / peaks['hbm_gbs'],
                              'algorithmic_bytes_per_step': step_bytes},
            'kernel_us_per_step': {k: v[0] * 1e3 / P for k, v in prof.items()},
            'eval': {'metric': 'top-k eval users/s (k=20, train mask, full catalog)', 'value': n_eval / t_eval,
                     'unit': 'users/s', 'users': n_eval, 'seconds': t_eval,
                     'timing': {'statistic': 'median of %d passes over the same (users, train CSR), named by a plan key: '
                                             'the train-mask images are built by the first pass only'
                                             % ev['seconds']['n'],
                                'seconds_min': ev['seconds']['min'], 'seconds_max': ev['seconds']['max'],
                                'first_call_seconds': ev['first_call'], 'unkeyed_seconds': ev['unkeyed'],
                                'rank_seconds': eval_rank_s},
                     'kernel': ev['kernel'],
                     'e2e': {'value': ev['users_total'] / t_eval_e2e, 'unit': 'users/s', 'seconds': t_eval_e2e,
                             'call': 'spotlight.evaluation.precision_recall_score%s(model, test, train, k=[5,10,20]) '
                                     'from host Interactions' % ('_sharded' if world > 1 else ''),
                             'h2d_bytes_per_step': ev['h2d_bytes'], 'd2h_bytes_per_step': ev['d2h_bytes'],
                             'precision_recall': ev['precision_recall']},
                     'roofline': {'bound': 'tensor', 'achieved': flops / t_eval / 1e12,
                                  'peak': peaks['bf16_tflops'] * world, 'unit': 'TFLOP/s',
                                  'frac': flops / t_eval / 1e12 / (peaks['bf16_tflops'] * world)}},
        }
        if with_cpu and rank == 0 and world == 1 and not args.no_cpu_baseline:
            from tests.gpu_helpers import tables_of
            eng.flush()
            res['eval']['cpu_baseline'] = cpu_reference_eval(tables_of(net), csrs[1], csrs[0], n_users=args.cpu_eval_users)
        eng.close()
        del eng, net
        torch.cuda.empty_cache()
        return res

    main = one_distribution(w['zipf'], True)
    twin = None if args.no_twin else one_distribution(not w['zipf'], False)

    # ---- cfg5: scaled catalog, tables row-sharded over the ranks (all-to-all of rows over NVLink); strong scaling
    sh = None
    if not args.no_sharded:
        from tools.shard_bench import run as shard_run
        sh = shard_run(steps=args.sharded_steps, warmup=args.sharded_warmup, zipf=w['zipf'],
                       fast_math=bool(args.fast_math))
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return None

    out = {
        'metric': 'train interactions/s (ML-20M-shape BPR MF)', 'value': main['value'], 'unit': 'interactions/s',
        'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': main['ms_per_step'], 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': config_of(w, w['zipf']),
        'run': {'parallelism': 'replicas x%d (training does not shard; eval shards users)' % world,
                'optimizer_math': 'fast (MUFU sqrt/rcp, ftz)' if args.fast_math else 'ieee',
                'final_loss': main['final_loss']},
    }
    for key in ('timing', 'clocks', 'e2e', 'gpu_launches', 'roofline', 'roofline_step', 'kernel_us_per_step', 'eval'):
        out[key] = main[key]
    if twin is not None:
        name = 'uniform' if w['zipf'] else 'zipf'
        out[name] = {k: twin[k] for k in ('value', 'ms_per_step', 'timing', 'e2e', 'roofline', 'roofline_step',
                                          'kernel_us_per_step', 'eval')}
        out[name]['items'] = 'uniform' if w['zipf'] else 'zipf(1.05)'
    if sh is not None:
        sh['roofline'] = {'bound': 'hbm', 'achieved': sh['hbm_gbs_algorithmic_total'], 'peak': peaks['hbm_gbs'] * world,
                          'unit': 'GB/s', 'frac': sh['hbm_gbs_algorithmic_total'] / (peaks['hbm_gbs'] * world),
                          'note': 'algorithmic bytes of the whole step over all GPUs; the step is bound by the '
                                  'dense-optimiser replay arithmetic (MUFU), see DESIGN.md'}
        out['sharded_train'] = sh
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_reference_steps(w, steps=args.cpu_steps, warmup=max(min(W, 5), 1))
        out['cpu_baseline'] = {'value': cb['value'], 'unit': 'interactions/s', 'cores': cb['cores'], 'kind': 'port',
                               'sample': cb['sample']}
    if dist:
        dist.destroy_process_group()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=494)       # a quarter of an ML-20M epoch (ceil(16.2M / 8192) = 1978)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--workload', default='cfg3', choices=sorted(WORKLOADS))
    ap.add_argument('--items', default='uniform', choices=['uniform', 'zipf'])
    ap.add_argument('--fast-math', type=int, default=1)
    ap.add_argument('--cpu-steps', type=int, default=120)
    ap.add_argument('--cpu-eval-users', type=int, default=300)
    ap.add_argument('--repeats', type=int, default=11, help='timed regions per measurement; the median is reported')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-twin', action='store_true', help='skip the second item-id distribution')
    ap.add_argument('--no-sharded', action='store_true', help='skip the cfg5 row-sharded training object')
    ap.add_argument('--sharded-steps', type=int, default=128)
    ap.add_argument('--sharded-warmup', type=int, default=512,
                    help='optimiser steps before the cfg5 timed region (the lazy replay length grows until every '
                         'row has been touched once: steady state needs a few hundred steps)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    w = dict(WORKLOADS[args.workload])
    w['zipf'] = args.items == 'zipf'
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))

    if args.impl == 'reference':
        if rank != 0:
            return 0
        cb = cpu_reference_steps(w, steps=args.steps, warmup=args.warmup)
        line = {'impl': 'reference', 'metric': 'train interactions/s (ML-20M-shape BPR MF)', 'value': cb['value'],
                'unit': 'interactions/s', 'n_gpus': args.gpus, 'steps': cb['steps'], 'warmup': args.warmup,
                'ms_per_step': cb['seconds'] * 1e3 / max(cb['steps'], 1), 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': config_of(w, w['zipf']),
                'cpu_baseline': {'value': cb['value'], 'unit': 'interactions/s', 'cores': cb['cores'], 'kind': 'port',
                                 'sample': cb['sample']},
                'e2e': {'value': cb['value'], 'unit': 'interactions/s', 'h2d_bytes_per_step': 0,
                        'd2h_bytes_per_step': 0}}
        print(json.dumps(line))
        return 0
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner, ...) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        out = native_bench(args, w, rank, world)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if out is not None:
        print(json.dumps(out))
        sys.stdout.flush()
    return 0


if __name__ == '__main__':
    sys.exit(main())
