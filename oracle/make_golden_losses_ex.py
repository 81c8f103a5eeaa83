"""Freezes the reference's loss functions on the inputs the model driver never produces but the function signatures
accept (spotlight/losses.py:20-172): the `mask` argument and 2-D negatives [n, b] for adaptive_hinge_loss (the upstream
per-positive maximum).  Runs the REAL reference from /root/reference; writes tests/golden/losses_ex.npz.

    python oracle/make_golden_losses_ex.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import import_reference, OUT   # noqa: E402


def main():
    out_dir = os.path.abspath(OUT)
    _, _, losses, _, _, _, _ = import_reference()
    rs = np.random.RandomState(23)
    b, n = 96, 4
    d = {}
    pos0 = rs.uniform(0.01, 0.99, b).astype(np.float32)
    neg1 = rs.uniform(0.01, 0.99, b).astype(np.float32)
    neg2 = rs.uniform(0.01, 0.99, (n, b)).astype(np.float32)
    neg2[1, 7] = neg2[3, 7] = 0.999          # tie inside a column: the first row gets the gradient
    neg2[:, 9] = 0.5                         # whole column tied
    mask = (rs.uniform(0, 1, b) < 0.7).astype(np.float32)
    mask[7] = 1.0
    d.update(pos=pos0, neg1=neg1, neg2=neg2, mask=mask)
    cases = [('pointwise', 'neg1', False), ('pointwise', 'neg1', True), ('bpr', 'neg1', True), ('hinge', 'neg1', True),
             ('adaptive_hinge', 'neg1', True), ('adaptive_hinge', 'neg2', False), ('adaptive_hinge', 'neg2', True),
             # pairwise losses broadcast over the rows of [n, b] negatives (mean over n*b; with a mask: / mask.sum())
             ('bpr', 'neg2', False), ('bpr', 'neg2', True), ('hinge', 'neg2', False), ('hinge', 'neg2', True)]
    for name, negkey, use_mask in cases:
        pos = torch.from_numpy(pos0.copy()).requires_grad_(True)
        neg = torch.from_numpy(d[negkey].copy()).requires_grad_(True)
        fn = getattr(losses, name + '_loss')
        val = fn(pos, neg, mask=torch.from_numpy(mask) if use_mask else None)
        val.backward()
        tag = '%s_%s_%s' % (name, negkey, 'mask' if use_mask else 'nomask')
        d['loss_' + tag] = val.detach().numpy()
        d['dpos_' + tag] = pos.grad.numpy().copy()
        d['dneg_' + tag] = neg.grad.numpy().copy()
    d['cases'] = np.array(['%s|%s|%d' % c for c in cases])
    np.savez_compressed(os.path.join(out_dir, 'losses_ex.npz'), **d)
    print('wrote', os.path.join(out_dir, 'losses_ex.npz'), sorted(d))


if __name__ == '__main__':
    main()
