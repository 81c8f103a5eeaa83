# timing of the evaluation top-k pass at cfg4 shape: tensor-core path vs exact kernel
import sys, os, numpy as np, torch, scipy.sparse as sp
sys.path.insert(0, '.')
import recommendation_gans_b200
from tests.gpu_helpers import make_net
from recommendation_gans_b200.engine import MFEngine
U, I, D, k = int(os.environ.get("EVAL_U", 138493)), 26744, 128, 20
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0 / 128
rs = np.random.RandomState(0)
tabs = [rs.normal(0, scale, (U, D)).astype(np.float32), rs.normal(0, scale, (I, D)).astype(np.float32),
        rs.normal(0, scale, (U, 1)).astype(np.float32), rs.normal(0, scale, (I, 1)).astype(np.float32)]
if os.environ.get('EVAL_SKEW'):
    # trained-model shape: a heavy tail of item norms and a common direction that popular items share with users
    pop = (1.0 + np.arange(I)) ** -0.6
    if os.environ['EVAL_SKEW'] != '2': rs.shuffle(pop)     # 2: popularity sorted by item id
    common = rs.normal(0, 1, D).astype(np.float32); common /= np.linalg.norm(common)
    tabs[1] = (tabs[1] * (1 + 40 * pop[:, None]) + 8 * scale * np.sqrt(D) * pop[:, None] * common).astype(np.float32)
    tabs[0] = (tabs[0] + 2 * scale * np.sqrt(D) * rs.rand(U, 1).astype(np.float32) * common).astype(np.float32)
    tabs[3] = (tabs[3] + 4 * scale * pop[:, None]).astype(np.float32)
n_tr = 117 * U
tu = np.sort(rs.randint(0, U, n_tr)); ti = rs.randint(0, I, n_tr)
csr = sp.coo_matrix((np.ones(n_tr), (tu, ti)), shape=(U, I)).tocsr(); csr.sum_duplicates(); csr.sort_indices()
indptr = torch.from_numpy(csr.indptr.astype(np.int64)).cuda(); indices = torch.from_numpy(csr.indices.astype(np.int32)).cuda()
users = torch.arange(U, device='cuda', dtype=torch.int64)
res = {}
for mode in (('1',) if os.environ.get('MFB_TC_DBG') or os.environ.get('EVAL_TC_ONLY') else ('1', '0')):
    os.environ['MFB_TC'] = mode
    eng = MFEngine(make_net(tabs))
    eng.profile(False)
    ids = eng.topk(users, k, indptr, indices)
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ids = eng.topk(users, k, indptr, indices); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[mode] = ids.cpu().numpy()
    if mode == '1':
        print('candidate stats', eng.debug_tc_stats(U))
    print('MFB_TC=%s: %.3f ms per pass -> %.2f M users/s, %.1f TFLOP/s algorithmic, redo %d' % (
        mode, best, U / best / 1e3, 2.0 * U * I * D / best / 1e9, eng.topk_last_redo))
if '0' in res: print('ids identical:', (res['1'] == res['0']).all())
print({k: v for k, v in eng.profile_read().items()} if False else '')
