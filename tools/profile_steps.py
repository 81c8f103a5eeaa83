# in-situ per-kernel-class timing of the training pipeline at a BASELINE config shape
import sys, random, argparse, numpy as np, torch
sys.path.insert(0, '.')
import recommendation_gans_b200
from tests.gpu_helpers import make_engine
ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=200)
ap.add_argument('--n_neg', type=int, default=1)
ap.add_argument('--loss', default='adaptive_hinge')
ap.add_argument('--fast', type=int, default=0)
ap.add_argument('--zipf', type=int, default=0)
ap.add_argument('--shape', default='cfg3')
a = ap.parse_args()
U, I, D, B = {'cfg1': (943, 1682, 32, 256), 'cfg2': (6040, 3706, 64, 1024), 'cfg3': (138493, 26744, 128, 8192)}[a.shape]
rs = np.random.RandomState(0)
tabs = [rs.normal(0, 1.0 / D, (U, D)).astype(np.float32), rs.normal(0, 1.0 / D, (I, D)).astype(np.float32),
        np.zeros((U, 1), np.float32), np.zeros((I, 1), np.float32)]
net, opt, eng = make_engine(tabs, 'adam', 1e-3, 1e-5, fast_math=bool(a.fast))
steps = a.steps
def ids(n, hi, zipf):
    if zipf:
        p = 1.0 / np.arange(1, hi + 1) ** 1.05
        return rs.choice(hi, n, p=p / p.sum())
    return rs.randint(0, hi, n)
users = torch.from_numpy(ids(steps * B, U, 0)).cuda(); items = torch.from_numpy(ids(steps * B, I, a.zipf)).cuda()
pop_u = torch.from_numpy(rs.randint(0, U, 1000000)).cuda(); pop_i = torch.from_numpy(rs.randint(0, I, 1000000)).cuda()
random.seed(0)
for rep in range(3):
    eng.profile(rep == 2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nu, ni = eng.draw_negative_pairs(pop_u, pop_i, steps * a.n_neg * B)
    e0.record()
    losses = eng.train_steps(a.loss, users, items, B, a.n_neg, nu, ni)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print('rep %d: %.2f ms, %.1f us/step, %.1f M inter/s' % (rep, ms, ms * 1e3 / steps, steps * B / ms / 1e3))
prof = eng.profile_read()
tot = sum(v[0] for v in prof.values())
for k, (ms, cnt) in prof.items():
    print('%-8s %8.3f ms total  %6d launches  %8.2f us/launch  %5.1f%%  (%.2f us/step)' % (k, ms, cnt, ms * 1e3 / cnt, 100 * ms / tot, ms * 1e3 / steps))
