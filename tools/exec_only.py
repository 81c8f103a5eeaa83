# exec-only timing: one planner chunk (no concurrent planning), events around train_steps minus plan time
import sys, random, numpy as np, torch
sys.path.insert(0, '.')
import recommendation_gans_b200
from tests.gpu_helpers import make_engine
U, I, D, B, n_neg = 138493, 26744, 128, 8192, 1
rs = np.random.RandomState(0)
tabs = [rs.normal(0, 1.0 / D, (U, D)).astype(np.float32), rs.normal(0, 1.0 / D, (I, D)).astype(np.float32),
        np.zeros((U, 1), np.float32), np.zeros((I, 1), np.float32)]
net, opt, eng = make_engine(tabs, 'adam', 1e-3, 1e-5, fast_math=True)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 64
users = torch.from_numpy(rs.randint(0, U, 8 * steps * B)).cuda(); items = torch.from_numpy(rs.randint(0, I, 8 * steps * B)).cuda()
pop_u = torch.from_numpy(rs.randint(0, U, 1000000)).cuda(); pop_i = torch.from_numpy(rs.randint(0, I, 1000000)).cuda()
random.seed(0)
for rep in range(8):
    eng.profile(rep % 2 == 1)
    nu, ni = eng.draw_negative_pairs(pop_u, pop_i, steps * n_neg * B)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.train_steps('adaptive_hinge', users[rep*steps*B:(rep+1)*steps*B], items[rep*steps*B:(rep+1)*steps*B], B, n_neg, nu, ni)
    e1.record(); torch.cuda.synchronize()
    tot = e0.elapsed_time(e1)
    if rep % 2 == 1:
        prof = eng.profile_read()
        plan = prof['pack'][0] + prof['sort'][0]
        print('rep %d (profiled): total %.2f ms, plan %.2f ms, exec %.1f us/step | fwd %.1f upd %.1f' % (rep, tot, plan, (tot - plan) * 1e3 / steps, prof['forward'][0]*1e3/steps, prof['update'][0]*1e3/steps))
    else:
        print('rep %d: total %.2f ms -> %.1f us/step incl. plan' % (rep, tot, tot * 1e3 / steps))
