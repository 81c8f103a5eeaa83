"""Pins the `representation=` modules (spotlight/dnn_models/mlp.py:5-46, neuMF.py:7-61) to the REAL reference: the
reference's MLP / NeuMF are built under a fixed torch seed; their state dicts and eval-mode outputs on fixed id pairs
are frozen -> tests/golden/dnn_models.npz.  Run in the build container only: python oracle/make_golden_dnn_models.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import OUT, import_reference  # noqa: E402


def main():
    import_reference()
    from spotlight.dnn_models.mlp import MLP
    from spotlight.dnn_models.neuMF import NeuMF
    torch.set_num_threads(1)
    U, I = 40, 55
    rs = np.random.RandomState(1)
    users, items = rs.randint(0, U, 64), rs.randint(0, I, 64)
    out = dict(users=users, items=items, meta=np.array([U, I]))
    torch.manual_seed(5)
    mlp = MLP(layers=[32, 16, 8], num_users=U, num_items=I, embedding_dim=16)
    torch.manual_seed(6)
    neumf = NeuMF(mlp_layers=[24, 12], num_users=U, num_items=I, mf_embedding_dim=10, mlp_embedding_dim=12)
    for name, net in (('mlp', mlp), ('neumf', neumf)):
        net.eval()
        sd = net.state_dict()
        out[name + '_keys'] = np.array(list(sd.keys()))
        for k, v in sd.items():
            out['%s/%s' % (name, k)] = v.numpy()
        with torch.no_grad():
            out[name + '_out'] = net(torch.from_numpy(users), torch.from_numpy(items)).numpy()
    np.savez_compressed(os.path.join(OUT, 'dnn_models.npz'), **out)
    print({k: v.shape for k, v in out.items() if k.endswith('_out')})


if __name__ == '__main__':
    main()
