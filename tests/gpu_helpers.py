"""Shared helpers for the -m gpu parity tests (CUDA path vs oracle / golden vectors)."""
import numpy as np
import torch

import recommendation_gans_b200  # noqa: F401  (puts the drop-in modules on sys.path)
from recommendation_gans_b200.engine import MFEngine
from spotlight.factorization.representations import BilinearNet
import spotlight.optimizers as optimizers


def make_net(tables):
    ue, ie, ub, ib = [np.asarray(t, dtype=np.float32) for t in tables]
    net = BilinearNet(ue.shape[0], ie.shape[0], ue.shape[1])
    with torch.no_grad():
        net.user_embeddings.weight.copy_(torch.from_numpy(ue))
        net.item_embeddings.weight.copy_(torch.from_numpy(ie))
        net.user_biases.weight.copy_(torch.from_numpy(ub.reshape(-1, 1)))
        net.item_biases.weight.copy_(torch.from_numpy(ib.reshape(-1, 1)))
    return net.cuda()


def make_engine(tables, opt_name='adam', lr=1e-3, l2=0.0, fast_math=False):
    # parity tests default to the IEEE arithmetic; the fast-math mode has its own parametrised cases
    net = make_net(tables)
    opt = None
    if opt_name is not None:
        opt = getattr(optimizers, opt_name + '_optimizer')(net.parameters(), lr=lr, weight_decay=l2)
    return net, opt, MFEngine(net, opt, fast_math=fast_math)


def tables_of(net):
    sd = net.state_dict()
    return [sd[k].detach().cpu().numpy() for k in
            ('user_embeddings.weight', 'item_embeddings.weight', 'user_biases.weight', 'item_biases.weight')]


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
