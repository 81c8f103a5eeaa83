# per-rank time of the user-sharded evaluation pass on ONE GPU: rank r of G evaluates users [r*U/G, (r+1)*U/G) of the
# cfg4 shape with the whole item table (what each rank of `bench.py --gpus G` does), keyed like bench.py's timed passes.
import sys, os, numpy as np, torch, scipy.sparse as sp
sys.path.insert(0, '.')
import recommendation_gans_b200
from tests.gpu_helpers import make_net
from recommendation_gans_b200.engine import MFEngine
U, I, D, k = 138493, 26744, 128, 20
rs = np.random.RandomState(0)
tabs = [rs.normal(0, 1 / 128, (U, D)).astype(np.float32), rs.normal(0, 1 / 128, (I, D)).astype(np.float32),
        rs.normal(0, 1 / 128, (U, 1)).astype(np.float32), rs.normal(0, 1 / 128, (I, 1)).astype(np.float32)]
n_tr = 117 * U
tu = np.sort(rs.randint(0, U, n_tr)); ti = rs.randint(0, I, n_tr)
csr = sp.coo_matrix((np.ones(n_tr), (tu, ti)), shape=(U, I)).tocsr(); csr.sum_duplicates(); csr.sort_indices()
indptr = torch.from_numpy(csr.indptr.astype(np.int64)).cuda(); indices = torch.from_numpy(csr.indices.astype(np.int32)).cuda()
eng = MFEngine(make_net(tabs))
eng.profile(False)
for G in (1, 2, 4, 8):
    n = (U + G - 1) // G
    users = torch.arange(0, n, device='cuda', dtype=torch.int64)
    key = 0x77 + G
    eng.topk(users, k, indptr, indices, plan_key=key)
    ts = []
    for rep in range(7):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.topk(users, k, indptr, indices, plan_key=key); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = float(np.median(ts))
    if G == 1: t1 = t
    print('G=%d users/rank %6d: %.3f ms per pass (keyed)  -> strong-scaling efficiency %.2f' % (G, n, t, t1 / (G * t)))
