#!/usr/bin/env python
"""Localises table deviations of a config-level parity run (tests/config_parity.py): steps the oracle and the
CUDA path one minibatch at a time, compares all four tables on the device after every step and reports, for every
row whose deviation first exceeds the bound, the step and the slots of that step that address the row.

    python tools/parity_debug.py [cfg] [uniform|zipf] [ieee|fast] [steps]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import config_parity as P  # noqa: E402
from oracle import mf_oracle as O  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else 'cfg3'
    zipf = (sys.argv[2] if len(sys.argv) > 2 else 'zipf') == 'zipf'
    fast = (sys.argv[3] if len(sys.argv) > 3 else 'ieee') == 'fast'
    c = dict(P.CONFIGS[cfg])
    if len(sys.argv) > 4:
        c['steps'] = int(sys.argv[4])
        P.CONFIGS[cfg]['steps'] = c['steps']
    stepwise = os.environ.get('STEPWISE', '1') != '0'
    from tests.gpu_helpers import make_engine
    users, items, neg_u, neg_i = P.synth(cfg, zipf)
    init = [t.numpy().copy() for t in O.init_tables(c['U'], c['I'], c['D'], torch_seed=0)]
    model = O.OracleMF(*[torch.from_numpy(t) for t in init], loss=c['loss'], optimizer='adam', lr=P.LR, l2=P.L2,
                       batch_size=c['B'], num_negative_samples=c['n_neg'])
    net, _, eng = make_engine(init, 'adam', P.LR, P.L2, fast)
    B, k = c['B'], c['n_neg'] * c['B']
    kind = P.loss_kind(c['loss'])
    names = ['user_emb', 'item_emb', 'user_bias', 'item_bias']
    params = [net.user_embeddings.weight, net.item_embeddings.weight, net.user_biases.weight, net.item_biases.weight]
    flagged = [set(), set(), set(), set()]
    bound = float(os.environ.get('BOUND', '2e-7'))
    tu, ti, tnu, tni = [torch.from_numpy(x) for x in (users, items, neg_u, neg_i)]
    events = []
    chunk = 1 if stepwise else int(os.environ.get('CHUNK', '50'))
    for s0 in range(0, c['steps'], chunk):
        s1 = min(s0 + chunk, c['steps'])
        ref_losses = [model.train_step(tu[s * B:(s + 1) * B], ti[s * B:(s + 1) * B], tnu[s * k:(s + 1) * k],
                                       tni[s * k:(s + 1) * k]).item() for s in range(s0, s1)]
        got = eng.train_steps(kind, users[s0 * B:s1 * B], items[s0 * B:s1 * B], B, c['n_neg'], neg_u[s0 * k:s1 * k],
                              neg_i[s0 * k:s1 * k]).cpu().numpy()
        eng.flush()
        torch.cuda.synchronize()
        lrel = float(np.max(np.abs(got - np.array(ref_losses)) / np.abs(ref_losses)))
        for t, (p, r) in enumerate(zip(params, model.tables)):
            d = (p.detach() - r.detach().cuda()).abs().amax(dim=1)
            bad = torch.nonzero(d > bound).reshape(-1).cpu().numpy()
            new = [int(x) for x in bad if int(x) not in flagged[t]]
            for row in new[:6]:
                flagged[t].add(row)
                ev = dict(step=s1 - 1, table=names[t], row=row, dev=float(d[row].item()), loss_rel=lrel)
                if t in (0, 2):
                    ids = np.concatenate([users[(s1 - 1) * B:s1 * B], neg_u[(s1 - 1) * k:s1 * k]])
                    other = np.concatenate([items[(s1 - 1) * B:s1 * B], neg_i[(s1 - 1) * k:s1 * k]])
                else:
                    ids = np.concatenate([items[(s1 - 1) * B:s1 * B], neg_i[(s1 - 1) * k:s1 * k]])
                    other = np.concatenate([users[(s1 - 1) * B:s1 * B], neg_u[(s1 - 1) * k:s1 * k]])
                slots = np.nonzero(ids == row)[0]
                ev['slots'] = [int(x) for x in slots[:8]]
                ev['n_slots'] = int(len(slots))
                ev['partners'] = [int(other[x]) for x in slots[:8]]
                if t in (0, 2) and len(slots):
                    oth_all = np.concatenate([items[(s1 - 1) * B:s1 * B], neg_i[(s1 - 1) * k:s1 * k]])
                    ev['partner_item_multiplicity'] = [int((oth_all == other[x]).sum()) for x in slots[:8]]
                with torch.no_grad():
                    diff = (p.detach()[row] - r.detach()[row].cuda()).cpu().numpy()
                ev['n_elems_over'] = int((np.abs(diff) > bound).sum())
                ev['ref_row_absmax'] = float(r.detach()[row].abs().max())
                events.append(ev)
                print(json.dumps(ev), flush=True)
            for row in new[6:]:
                flagged[t].add(row)
            if len(new) > 6:
                print(json.dumps(dict(step=s1 - 1, table=names[t], more_new_rows=len(new) - 6)), flush=True)
    print('flagged rows per table:', [len(f) for f in flagged])


if __name__ == '__main__':
    main()
