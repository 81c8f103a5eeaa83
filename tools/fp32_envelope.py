#!/usr/bin/env python
"""How far is the fp32 REFERENCE from itself?  (CPU only; evidence for tests/config_parity.table_report.)

Runs the oracle restatement of the reference step (same torch CPU ops, dense torch.optim.Adam) on a BASELINE config
three times: in fp32, in float64, and in fp32 with every minibatch's positives permuted -- the same multiset of
interactions per step, i.e. the same mathematics, only a different fp32 summation order inside
embedding_dense_backward.  Prints the max-norm relative deviation of each table between the runs.

    python tools/fp32_envelope.py [cfg3] [steps] [uniform|zipf]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import config_parity as P  # noqa: E402
from oracle import mf_oracle as O  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else 'cfg3'
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    zipf = (sys.argv[3] if len(sys.argv) > 3 else 'zipf') == 'zipf'
    P.CONFIGS[cfg]['steps'] = steps
    c = P.CONFIGS[cfg]
    users, items, neg_u, neg_i = P.synth(cfg, zipf)
    init = [t.numpy().copy() for t in O.init_tables(c['U'], c['I'], c['D'], torch_seed=0)]

    def run(dtype, permute=False):
        model = O.OracleMF(*[torch.from_numpy(t).to(dtype) for t in init], loss=c['loss'], optimizer='adam', lr=P.LR,
                           l2=P.L2, batch_size=c['B'], num_negative_samples=c['n_neg'])
        B, k = c['B'], c['n_neg'] * c['B']
        tu, ti, nu, ni = [torch.from_numpy(x) for x in (users, items, neg_u, neg_i)]
        rs = np.random.RandomState(5)
        for s in range(steps):
            a, b, cn, d = tu[s * B:(s + 1) * B], ti[s * B:(s + 1) * B], nu[s * k:(s + 1) * k], ni[s * k:(s + 1) * k]
            if permute:
                p = torch.from_numpy(rs.permutation(len(a)))
                a, b = a[p], b[p]
            model.train_step(a, b, cn, d)
        return [t.detach().double().numpy() for t in model.tables]

    t0 = time.time()
    r32 = run(torch.float32)
    r64 = run(torch.float64)
    r32p = run(torch.float32, permute=True)
    print('%s, %d steps, items %s, %d host threads, %.0f s' % (cfg, steps, 'zipf(1.05)' if zipf else 'uniform',
                                                               torch.get_num_threads(), time.time() - t0))
    names = ['user_emb', 'item_emb', 'user_bias', 'item_bias']
    for title, a, b in (('fp32 reference vs float64 run', r32, r64),
                        ('fp32 reference vs fp32 reference with permuted minibatches', r32, r32p),
                        ('permuted fp32 reference vs float64 run', r32p, r64)):
        print(title)
        for n, x, y in zip(names, a, b):
            d = np.abs(x - y)
            scale = np.abs(y).max()
            print('  %-10s max|diff|/max|ref| = %.2e   elements outside 1e-5: %d of %d' % (n, d.max() / scale,
                                                                                            (d > 1e-5 * scale).sum(), d.size))


if __name__ == '__main__':
    main()
