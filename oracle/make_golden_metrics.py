"""Pins map_at_k and rmse_score (spotlight/evaluation.py:187-190, 278-353) to the REAL reference.

Run in the build container only:   python oracle/make_golden_metrics.py
Loads the final tables of the committed fit goldens (tests/golden/fit_*.npz, themselves produced by the
reference's own fit), puts them into the reference's BilinearNet / ImplicitFactorizationModel and freezes what the
reference's evaluation functions return for the same test split -> tests/golden/metrics.npz.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import OUT, import_reference  # noqa: E402


def main():
    ref_implicit, sampling, losses, evaluation, optimizers, Interactions, BilinearNet = import_reference()
    torch.set_num_threads(1)
    out = {}
    for name in ('fit_pointwise', 'fit_bpr'):
        g = np.load(os.path.join(OUT, name + '.npz'))
        U, I, D, B, n_neg, n_epochs = [int(x) for x in g['meta']]
        a, b = [int(x) for x in g['split']]
        users, items = g['users'], g['items']
        test = Interactions(users[b:].astype(np.int32), items[b:].astype(np.int32), num_users=U, num_items=I)
        net = BilinearNet(U, I, D, sparse=False)
        with torch.no_grad():
            net.user_embeddings.weight.copy_(torch.from_numpy(g['final0']))
            net.item_embeddings.weight.copy_(torch.from_numpy(g['final1']))
            net.user_biases.weight.copy_(torch.from_numpy(g['final2']))
            net.item_biases.weight.copy_(torch.from_numpy(g['final3']))
        model = ref_implicit.ImplicitFactorizationModel(embedding_dim=D, representation=net, batch_size=B,
                                                        experiment_name='golden_metrics_' + name)
        model.set_users(U, I)
        net.eval()
        for k in (1, 5, 10):
            out['%s_map_k%d' % (name, k)] = np.array(evaluation.map_at_k(model, test, k=k))
        # model.test (implicit.py:428-437): rmse_score summed over minibatches of the test pairs
        tu, ti = torch.from_numpy(users[b:]).long(), torch.from_numpy(items[b:]).long()
        parts = [float(evaluation.rmse_score(net, tu[s:s + B], ti[s:s + B])) for s in range(0, len(tu), B)]
        out['%s_rmse_parts' % name] = np.array(parts)
        out['%s_bce' % name] = np.array(np.sqrt(sum(parts) / len(tu)))
        print(name, {k: float(v) for k, v in out.items() if k.startswith(name) and v.ndim == 0})
    np.savez_compressed(os.path.join(OUT, 'metrics.npz'), **out)


if __name__ == '__main__':
    main()
