"""Pins the data path in front of fit to the REAL reference: make_implicit, the time-based / random
splits, shuffle, the cache files data_provider writes and what it reads back, argparse defaults.

Run in the build container only:   python oracle/make_golden_data_path.py   -> tests/golden/data_path.npz
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import OUT, import_reference  # noqa: E402


def main():
    import_reference()
    import pandas as pd
    from spotlight.interactions import Interactions
    from spotlight.dataset_manilupation import (random_train_test_split, shuffle_interactions,
                                                train_test_timebased_split)
    from utils.helper_functions import make_implicit
    from utils.data_provider import data_provider
    from utils import arg_extractor
    rs = np.random.RandomState(7)
    U, I, n = 60, 45, 1200
    users, items = rs.randint(0, U, n).astype(np.int64), rs.randint(0, I, n).astype(np.int64)
    ratings = rs.randint(1, 6, n).astype(np.float64)
    timestamps = rs.randint(0, 5000, n).astype(np.int64)          # many equal timestamps: argsort tie order matters
    out = dict(users=users, items=items, ratings=ratings, timestamps=timestamps, meta=np.array([U, I]))

    def mk():
        return Interactions(users.copy(), items.copy(), ratings.copy(), timestamps.copy(), num_users=U, num_items=I)
    data = make_implicit(mk())
    out['implicit_ratings'] = np.asarray(data.ratings)
    train, test = train_test_timebased_split(data, test_percentage=0.1)
    train, valid = train_test_timebased_split(train, test_percentage=0.1)
    out['inplace_user_ids'] = data.user_ids
    for name, part in (('train', train), ('valid', valid), ('test', test)):
        for field in ('user_ids', 'item_ids', 'ratings', 'timestamps'):
            out['%s_%s' % (name, field)] = np.asarray(getattr(part, field))
    a, b = random_train_test_split(mk(), test_percentage=0.2, random_state=np.random.RandomState(3))
    out['rand_train_items'], out['rand_test_users'] = a.item_ids, b.user_ids
    out['shuffled_timestamps'] = shuffle_interactions(mk(), random_state=np.random.RandomState(4)).timestamps
    # cache files written by the reference's own writer, then read back by its own reader
    neg = [(int(u), int(i)) for u, i in zip(rs.randint(0, U, len(train)), rs.randint(0, I, len(train)))]
    pop = pd.Series(np.bincount(items, minlength=I)[:I], index=np.arange(I) + 1000)
    tmp = tempfile.mkdtemp(prefix='refcache_') + os.sep
    writer = data_provider.__new__(data_provider)
    writer.movies_to_keep = -1
    writer.save_statistics(tmp + 'movielens_100K', U, I, len(data))
    writer.create_cvs_files(tmp + 'movielens_100K', train, valid, test, neg, pop)
    for part in ('train', 'valid', 'test', 'popularity'):
        with open('%smovielens_100K_%s_-1.csv' % (tmp, part)) as f:
            out['csv_' + part] = np.array(f.read())
    with open(tmp + 'movielens_100K_statistics_-1.json') as f:
        out['statistics_json'] = np.array(f.read())
    out['neg_pairs'] = np.array(neg, dtype=np.int64)
    out['popularity_values'], out['popularity_index'] = pop.values, pop.index.values
    loaded = data_provider(tmp, '100K', 1, movies_to_keep=-1)
    tr, va, te, neg2, pop2 = loaded.get_timebased_data()
    assert neg2 == neg
    for name, part in (('train', tr), ('valid', va), ('test', te)):
        out['loaded_%s_ratings' % name] = np.asarray(part.ratings)
        assert (np.asarray(part.user_ids) == out[name + '_user_ids']).all()
    sys.argv = ['mf_spotlight.py']
    out['arg_defaults'] = np.array(json.dumps(vars(arg_extractor.get_args()), sort_keys=True))
    np.savez_compressed(os.path.join(OUT, 'data_path.npz'), **out)
    print('written', sorted(out))


if __name__ == '__main__':
    main()
