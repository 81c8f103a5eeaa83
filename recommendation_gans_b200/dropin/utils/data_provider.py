"""Dataset cache of the training scripts (reference: utils/data_provider.py:21-178).

`data_provider(path, variant, negative_per_positive, movies_to_keep).get_timebased_data()` returns
`(train, valid, test, neg_examples, item_popularity)` exactly as the reference does, from the same on-disk cache:

    <path>movielens_<variant>_{train,valid,test}_<keep>.csv   columns userId,movieId,rating,timestamp
    <path>movielens_<variant>_popularity_<keep>.csv           headerless (movieId, count)
    <path>movielens_<variant>_statistics_<keep>.json          {"num_users", "num_items", "interactions"}
    <path>movielens_<variant>_ngt_<keep>.pkl                  pickled list of (user, item) negative pairs

When the cache is missing it is built from the raw hdf5 file: rating filter, implicit ratings, two time-based splits
(81 % / 9 % / 10 %) and `len(train)` negative pairs -- the pairs come from the GPU generator
(spotlight.sampling.get_negative_samples -> mfb_negative_pairs: the same pairs, in seconds instead of the reference's
per-sample Python loop).
"""
import json
import logging
import os
import pickle
import time

import pandas as pd

from spotlight.dataset_manilupation import train_test_timebased_split
from spotlight.interactions import Interactions
from spotlight.sampling import get_negative_samples
from utils.helper_functions import make_implicit

logging.basicConfig(format='%(message)s', level=logging.INFO)


class data_provider(object):

    def __init__(self, path, variant, negative_per_positive, movies_to_keep=-1):
        self.movies_to_keep = movies_to_keep
        rel_path = path + 'movielens_' + variant
        start = time.time()
        if self.exists(rel_path):
            logging.info("Data exists, loading from file ... ")
            stats = self.read_statistics(rel_path)
            sets = []
            for part in ('train', 'valid', 'test'):
                df = pd.read_csv(self._file(rel_path, part, 'csv'))
                sets.append(make_implicit(self.create_interactions(df, stats['num_users'], stats['num_items'])))
            train_set, valid_set, test_set = sets
            item_popularity = pd.read_csv(self._file(rel_path, 'popularity', 'csv'), header=None).iloc[:, 1]
            neg_examples = self.read_negative_examples(self._file(rel_path, 'ngt', 'pkl'))
        else:
            logging.info('Dataset is not set, creating csv files')
            from spotlight.datasets.movielens import get_movielens_dataset
            dataset, item_popularity = get_movielens_dataset(variant=variant, path=path, movies_to_keep=movies_to_keep)
            self.save_statistics(rel_path, dataset.num_users, dataset.num_items, len(dataset))
            dataset = make_implicit(dataset)
            train_set, test_set = train_test_timebased_split(dataset, test_percentage=0.1)
            train_set, valid_set = train_test_timebased_split(train_set, test_percentage=0.1)
            neg_examples = get_negative_samples(dataset, len(train_set))
            self.create_cvs_files(rel_path, train_set, valid_set, test_set, neg_examples, item_popularity)
        logging.info("Took %d seconds" % (time.time() - start))
        logging.info("{} user and {} items".format(train_set.num_users, train_set.num_items))
        self.config = {'train_set': train_set, 'valid_set': valid_set, 'test_set': test_set,
                       'item_popularity': item_popularity, 'neg_examples': neg_examples}

    def get_timebased_data(self):
        """(train, valid, test, neg_examples, item_popularity) -- data_provider.py:97-119."""
        c = self.config
        return c['train_set'], c['valid_set'], c['test_set'], c['neg_examples'], c['item_popularity']

    # -- cache files ---------------------------------------------------------------------------------
    def _file(self, rel_path, part, ext):
        return '%s_%s_%s.%s' % (rel_path, part, self.movies_to_keep, ext)

    def save_statistics(self, path, num_users, num_items, interactions):
        with open(self._file(path, 'statistics', 'json'), 'w') as fp:
            json.dump({'num_users': num_users, 'num_items': num_items, 'interactions': interactions}, fp)

    def read_statistics(self, path):
        with open(self._file(path, 'statistics', 'json'), 'r') as fp:
            return json.load(fp)

    def create_interactions(self, df, num_users, num_items):
        return Interactions(df.userId.values, df.movieId.values, df.rating.values, df.timestamp.values,
                            num_users=num_users, num_items=num_items)

    def create_cvs_files(self, rel_path, train, valid, test, neg_examples, item_popularity):
        with open(self._file(rel_path, 'ngt', 'pkl'), 'wb') as f:
            p = pickle.Pickler(f)
            p.fast = True
            p.dump(neg_examples)
        for part, data in (('train', train), ('valid', valid), ('test', test)):
            frame = pd.DataFrame({'userId': data.user_ids, 'movieId': data.item_ids, 'rating': data.ratings,
                                  'timestamp': data.timestamps}, columns=['userId', 'movieId', 'rating', 'timestamp'])
            frame.to_csv(self._file(rel_path, part, 'csv'), index=False)
        item_popularity.to_csv(self._file(rel_path, 'popularity', 'csv'), header=False)

    def read_negative_examples(self, target):
        with open(target, 'rb') as f:
            return pickle.load(f)

    def exists(self, path):
        return all(os.path.exists(self._file(path, part, ext)) for part, ext in
                   (('train', 'csv'), ('popularity', 'csv'), ('valid', 'csv'), ('test', 'csv'), ('ngt', 'pkl')))
