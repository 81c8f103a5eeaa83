"""MovieLens loader (reference: spotlight/datasets/movielens.py:34-142).

Reads `<path>movielens_<variant>.hdf5` (datasets /user_id, /item_id, /rating, /timestamp), keeps ratings above 3.5 and
users with at least `min_uc` of them, and re-numbers users and items in order of first appearance.  The per-rating
Python loops of the reference (dict lookups through map/lambda) are replaced by pandas.factorize, which assigns the
same first-appearance codes.  h5py is imported on use: the cached-CSV path of utils.data_provider never needs it."""
import logging

import numpy as np
import pandas as pd

from spotlight.interactions import Interactions

logging.basicConfig(format='%(message)s', level=logging.INFO)

VARIANTS = ('100K', '1M', '10M', '20M')


def _get_movielens(dataset):
    import h5py                                   # not needed unless the raw file is read
    path = dataset + '.hdf5'
    logging.info("Data will be read from file: " + path)
    with h5py.File(path, 'r') as data:
        return data['/user_id'][:], data['/item_id'][:], data['/rating'][:], data['/timestamp'][:]


def get_count(tp, id):
    return tp[[id]].groupby(id, as_index=False).size()


def _counts(tp, column):
    """Interactions per id as a Series indexed by id (what the reference's groupby(...).size() was under the pandas
    of its time)."""
    return tp.groupby(column).size()


def filter_triplets(tp, min_uc=5, min_sc=0):
    """movielens.py:72-86: drop items with fewer than min_sc users, then users with fewer than min_uc items."""
    if min_sc > 0:
        itemcount = _counts(tp, 'movieId')
        tp = tp[tp['movieId'].isin(itemcount.index[itemcount >= min_sc])]
    if min_uc > 0:
        usercount = _counts(tp, 'userId')
        tp = tp[tp['userId'].isin(usercount.index[usercount >= min_uc])]
    return tp, _counts(tp, 'userId'), _counts(tp, 'movieId')


def keep_top_k(dataset, k):
    """movielens.py:60-62 (keeps the 1000 most frequent movies whatever k is, like the reference)."""
    valid = _counts(dataset, 'movieId').sort_values(ascending=False)[:1000].index.values
    return dataset.loc[dataset['movieId'].isin(valid)]


def get_movielens_dataset(variant='100K', path=None, min_uc=5, min_sc=0, movies_to_keep=-1):
    if variant not in VARIANTS:
        raise ValueError('Variant must be one of {}, got {}.'.format(VARIANTS, variant))
    url = 'movielens_{}'.format(variant)
    if path:
        url = path + url
    users, items, ratings, timestamps = _get_movielens(url)
    return interactions_from_ratings(users, items, ratings, timestamps, min_uc=min_uc, min_sc=min_sc,
                                     movies_to_keep=movies_to_keep)


def interactions_from_ratings(users, items, ratings, timestamps, min_uc=5, min_sc=0, movies_to_keep=-1):
    """movielens.py:113-142 on raw rating arrays -> (Interactions, item popularity Series)."""
    dataset = pd.DataFrame({'userId': users, 'movieId': items, 'rating': ratings, 'timestamps': timestamps})
    dataset = dataset[dataset['rating'] > 3.5]
    dataset, _, itemcount = filter_triplets(dataset, min_uc=min_uc, min_sc=min_sc)
    if movies_to_keep != -1 and movies_to_keep < itemcount.size:
        dataset = keep_top_k(dataset, movies_to_keep)
    uid, user_index = pd.factorize(dataset.userId.values)        # codes in order of first appearance
    sid, item_index = pd.factorize(dataset.movieId.values)
    num_users, num_items = len(user_index), len(item_index)
    logging.info("{} users and {} items".format(num_users, num_items))
    return Interactions(np.asarray(uid, dtype=np.int64), np.asarray(sid, dtype=np.int64), dataset.rating.values,
                        dataset.timestamps.values, num_users=num_users, num_items=num_items), itemcount
