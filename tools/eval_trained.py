# evaluation top-k on a model TRAINED on Zipf-skewed items (bench.py --items zipf shape): candidate statistics of
# the tensor-core path, redo count, and ids against the exact kernel
import os, sys, random, numpy as np, torch
sys.path.insert(0, '.')
import recommendation_gans_b200  # noqa
import bench
from recommendation_gans_b200.engine import MFEngine
from spotlight.factorization.representations import BilinearNet
import spotlight.optimizers as optimizers

w = dict(bench.WORKLOADS['cfg3']); w['zipf'] = True
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
rs = np.random.RandomState(0); torch.manual_seed(0)
net = BilinearNet(w['U'], w['I'], w['D']).cuda()
opt = optimizers.adam_optimizer(net.parameters(), lr=w['lr'], weight_decay=w['l2'])
eng = MFEngine(net, opt)
B = w['B']
users = torch.from_numpy(bench.synth_ids(rs, steps * B, w['U'], False)).cuda()
items = torch.from_numpy(bench.synth_ids(rs, steps * B, w['I'], True)).cuda()
pop_u = torch.from_numpy(bench.synth_ids(rs, w['n_train'], w['U'], False)).cuda()
pop_i = torch.from_numpy(bench.synth_ids(rs, w['n_train'], w['I'], False)).cuda()
rng = random.Random(0); eng.rng_seed(rng)
losses = eng.train_epoch(bench.loss_kind(w['loss']), users, items, B, w['n_neg'], pop_u, pop_i)
eng.flush()
print('final loss', float(losses[-1]))
for name, p in net.named_parameters():
    x = p.detach()
    n = x.norm(dim=1) if x.shape[1] > 1 else x.abs().squeeze(1)
    print('%-26s finite=%s  row norm: median %.4g  p99 %.4g  max %.4g' % (name, bool(torch.isfinite(x).all()), n.median(), n.quantile(0.99) if n.numel() < 1e7 else -1, n.max()))
ev = bench.eval_bench(eng, w, 0, 1, torch.device('cuda'), rs)
print(ev)
print('stats', eng.debug_tc_stats(w['U']))
