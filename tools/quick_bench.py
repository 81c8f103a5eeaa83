# first-contact timing of the V1 pipeline at cfg3 shape (not the bench contract yet)
import sys, time, random, numpy as np, torch
sys.path.insert(0, '.')
import recommendation_gans_b200
from tests.gpu_helpers import make_engine
U, I, D, B, n_neg = 138493, 26744, 128, 8192, 1
rs = np.random.RandomState(0)
tabs = [rs.normal(0, 1.0 / D, (U, D)).astype(np.float32), rs.normal(0, 1.0 / D, (I, D)).astype(np.float32),
        np.zeros((U, 1), np.float32), np.zeros((I, 1), np.float32)]
for fast in (False, True):
    net, opt, eng = make_engine(tabs, 'adam', 1e-3, 1e-5, fast_math=fast)
    steps = 200
    users = torch.from_numpy(rs.randint(0, U, steps * B)).cuda()
    items = torch.from_numpy(rs.randint(0, I, steps * B)).cuda()
    pop_u = torch.from_numpy(rs.randint(0, U, 1000000)).cuda(); pop_i = torch.from_numpy(rs.randint(0, I, 1000000)).cuda()
    random.seed(0)
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e0.record()
        nu, ni = eng.draw_negative_pairs(pop_u, pop_i, steps * n_neg * B)
        e1.record()
        losses = eng.train_steps('adaptive_hinge', users, items, B, n_neg, nu, ni)
        e2.record(); torch.cuda.synchronize()
        print('fast=%d rep %d: draw %.2f ms, train %.2f ms (%.1f us/step, %.1f M inter/s), wall %.1f ms, loss %.5f' % (
            fast, rep, e0.elapsed_time(e1), e1.elapsed_time(e2), e1.elapsed_time(e2) * 1e3 / steps,
            steps * B / e1.elapsed_time(e2) / 1e3, (time.time() - t0) * 1e3, losses[-1].item()))
    t0 = time.time(); eng.flush(); torch.cuda.synchronize(); print('flush %.2f ms' % ((time.time() - t0) * 1e3))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    uids = torch.arange(0, 20000).cuda()
    eng.topk(uids[:64], 20); ev0.record(); top = eng.topk(uids, 20); ev1.record(); torch.cuda.synchronize()
    print('topk exact: %d users in %.2f ms -> %.0f users/s' % (len(uids), ev0.elapsed_time(ev1), len(uids) / ev0.elapsed_time(ev1) * 1e3))
