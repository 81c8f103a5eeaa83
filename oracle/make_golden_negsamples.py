"""Freezes the reference's offline negative-pair generator (spotlight/sampling.py:46-70) on small seeded inputs:
runs the REAL get_negative_samples from /root/reference under np.random.seed and writes
tests/golden/neg_samples.npz (inputs, the pairs, and numpy's global MT19937 state afterwards).

    python oracle/make_golden_negsamples.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import import_reference, OUT   # noqa: E402


def main():
    out_dir = os.path.abspath(OUT)
    _, sampling, _, _, _, Interactions, _ = import_reference()
    d = {}
    cases = []
    # name, users, items, n interactions, samples, seed: dense enough that many pairs collide; a duplicate
    # interaction (stored value 2 -> has_key false); a user who interacted with all items but one
    for name, U, I, n, N, seed in (('dense', 40, 25, 500, 3000, 3), ('sparse', 300, 200, 2000, 5000, 11)):
        rs = np.random.RandomState(seed)
        users = rs.randint(0, U, n).astype(np.int32)
        items = rs.randint(0, I, n).astype(np.int32)
        users[users == 0] = 2                            # user 0 gets exactly the row below
        users[:2], items[:2] = 1, 3                      # duplicate pair
        full = np.arange(I - 1, dtype=np.int32)          # user 0: every item except the last
        users = np.concatenate([users, np.zeros(len(full), np.int32)])
        items = np.concatenate([items, full])
        train = Interactions(users, items, num_users=U, num_items=I)
        np.random.seed(seed)
        pairs = sampling.get_negative_samples(train, N)
        st = np.random.get_state()
        d[name + '_users'], d[name + '_items'] = users, items
        d[name + '_meta'] = np.array([U, I, N, seed])
        d[name + '_pairs'] = np.array(pairs, dtype=np.int64)
        d[name + '_state_key'], d[name + '_state_pos'] = st[1], np.array(st[2])
        cases.append(name)
        print(name, len(pairs), pairs[:3], st[2])
    d['cases'] = np.array(cases)
    np.savez_compressed(os.path.join(out_dir, 'neg_samples.npz'), **d)


if __name__ == '__main__':
    main()
