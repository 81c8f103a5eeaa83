# timings of the "next rows"  at ML-20M shape: offline negative-pair generator, mrr_score, hit_ratio
import sys, time, numpy as np, torch, scipy.sparse as sp
sys.path.insert(0, '.')
import recommendation_gans_b200
from tests.gpu_helpers import make_net
from spotlight.interactions import Interactions
from spotlight.sampling import get_negative_samples_arrays
from spotlight.evaluation import mrr_score, hit_ratio
U, I, D = 138493, 26744, 128
rs = np.random.RandomState(0)
n_train, n_test = 16200213, 2000026
p = 1.0 / np.arange(1, I + 1) ** 1.05
tu, ti = rs.randint(0, U, n_train).astype(np.int32), rs.choice(I, n_train, p=p / p.sum()).astype(np.int32)
train = Interactions(tu, ti, num_users=U, num_items=I)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    nu, ni = get_negative_samples_arrays(train, n_train, np.random.RandomState(1))
    torch.cuda.synchronize(); dt = time.time() - t0
    print('get_negative_samples: %d pairs in %.3f s (%.1f M pairs/s), CSR nnz %d' % (n_train, dt, n_train / dt / 1e6, train.tocsr().nnz))
tabs = [rs.normal(0, 1.0 / 8, (U, D)).astype(np.float32), rs.normal(0, 1.0 / 8, (I, D)).astype(np.float32),
        rs.normal(0, 0.1, (U, 1)).astype(np.float32), rs.normal(0, 0.1, (I, 1)).astype(np.float32)]
class M(object):
    _net = make_net(tabs); _num_items = I
test = Interactions(rs.randint(0, U, n_test).astype(np.int32), rs.randint(0, I, n_test).astype(np.int32), num_users=U, num_items=I)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    m = mrr_score(M(), test, train=train)
    torch.cuda.synchronize(); dt = time.time() - t0
    print('mrr_score (train mask): %d users in %.3f s (%.0f users/s), mean %.5f' % (len(m), dt, len(m) / dt, m.mean()))
loo = Interactions(np.arange(U, dtype=np.int32), rs.randint(0, I, U).astype(np.int32), num_users=U, num_items=I)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    h = hit_ratio(M(), loo, k=10)
    torch.cuda.synchronize(); dt = time.time() - t0
    print('hit_ratio@10: %d users in %.3f s (%.0f users/s), value %.5f' % (U, dt, U / dt, h))
