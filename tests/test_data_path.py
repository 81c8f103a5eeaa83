"""the data path in front of fit -- implicit ratings, time-based split, the on-disk cache
(utils/data_provider.py:43-64,157-178), the MovieLens filter / re-numbering, the CLI flags -- against vectors frozen
from the reference's own functions (tests/golden/data_path.npz, oracle/make_golden_data_path.py).  CPU only."""
import json
import os
import pickle
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, 'recommendation_gans_b200', 'dropin')
if DROPIN not in sys.path:
    sys.path.insert(0, DROPIN)


@pytest.fixture(scope='module')
def g(golden_dir):
    return np.load(os.path.join(golden_dir, 'data_path.npz'), allow_pickle=False)


def _mk(g, ratings=None):
    from spotlight.interactions import Interactions
    return Interactions(g['users'].copy(), g['items'].copy(), (g['ratings'] if ratings is None else ratings).copy(),
                        g['timestamps'].copy(), num_users=int(g['meta'][0]), num_items=int(g['meta'][1]))


def test_make_implicit_matches_reference(g):
    from utils.helper_functions import make_implicit
    out = make_implicit(_mk(g))
    assert out.ratings.dtype == g['implicit_ratings'].dtype
    np.testing.assert_array_equal(out.ratings, g['implicit_ratings'])


def test_timebased_split_matches_reference(g):
    """Including the reference's quirks: the input object is re-ordered in place, ratings are NOT re-ordered."""
    from spotlight.dataset_manilupation import train_test_timebased_split
    from utils.helper_functions import make_implicit
    data = make_implicit(_mk(g))
    train, test = train_test_timebased_split(data, test_percentage=0.1)
    train, valid = train_test_timebased_split(train, test_percentage=0.1)
    for name, part in (('train', train), ('valid', valid), ('test', test)):
        for field in ('user_ids', 'item_ids', 'ratings', 'timestamps'):
            np.testing.assert_array_equal(getattr(part, field), g['%s_%s' % (name, field)], err_msg=name + field)
        assert (part.num_users, part.num_items) == (int(g['meta'][0]), int(g['meta'][1]))
    np.testing.assert_array_equal(data.user_ids, g['inplace_user_ids'])


def test_random_split_and_shuffle_match_reference(g):
    from spotlight.dataset_manilupation import random_train_test_split, shuffle_interactions
    a, b = random_train_test_split(_mk(g), test_percentage=0.2, random_state=np.random.RandomState(3))
    np.testing.assert_array_equal(a.item_ids, g['rand_train_items'])
    np.testing.assert_array_equal(b.user_ids, g['rand_test_users'])
    s = shuffle_interactions(_mk(g), random_state=np.random.RandomState(4))
    np.testing.assert_array_equal(s.timestamps, g['shuffled_timestamps'])


def test_cache_files_round_trip_and_match_the_reference_layout(g, tmp_path):
    """The drop-in writes the cache the reference's data_provider reads (same file names, CSV header and columns,
    statistics json, pickled list of pairs) and reads what the reference wrote (committed fixture strings)."""
    import pandas as pd
    from utils.data_provider import data_provider
    from spotlight.dataset_manilupation import train_test_timebased_split
    from utils.helper_functions import make_implicit
    path = str(tmp_path) + os.sep
    data = make_implicit(_mk(g))
    train, test = train_test_timebased_split(data, test_percentage=0.1)
    train, valid = train_test_timebased_split(train, test_percentage=0.1)
    neg = [(int(u), int(i)) for u, i in g['neg_pairs']]
    pop = pd.Series(g['popularity_values'], index=g['popularity_index'])
    writer = data_provider.__new__(data_provider)
    writer.movies_to_keep = -1
    writer.save_statistics(path + 'movielens_100K', int(g['meta'][0]), int(g['meta'][1]), len(data))
    writer.create_cvs_files(path + 'movielens_100K', train, valid, test, neg, pop)
    # byte-level layout pinned to what the reference's own writer produced for the same data
    for part in ('train', 'valid', 'test', 'popularity'):
        with open('%smovielens_100K_%s_-1.csv' % (path, part)) as f:
            assert f.read() == str(g['csv_' + part]), part
    with open(path + 'movielens_100K_statistics_-1.json') as f:
        assert json.load(f) == json.loads(str(g['statistics_json']))
    with open(path + 'movielens_100K_ngt_-1.pkl', 'rb') as f:
        assert pickle.load(f) == neg
    loaded = data_provider(path, '100K', 1, movies_to_keep=-1)
    tr, va, te, neg2, pop2 = loaded.get_timebased_data()
    assert neg2 == neg
    np.testing.assert_array_equal(pop2.values, g['popularity_values'])
    for name, part in (('train', tr), ('valid', va), ('test', te)):
        np.testing.assert_array_equal(part.user_ids, g[name + '_user_ids'])
        np.testing.assert_array_equal(part.item_ids, g[name + '_item_ids'])
        np.testing.assert_array_equal(part.timestamps, g[name + '_timestamps'])
        # the reference re-applies make_implicit to the cached 0/1 ratings: everything becomes 0 (data_provider.py:57-59)
        np.testing.assert_array_equal(part.ratings, g['loaded_%s_ratings' % name])
        assert (part.num_users, part.num_items) == (int(g['meta'][0]), int(g['meta'][1]))


def test_movielens_filter_and_renumbering():
    """interactions_from_ratings (movielens.py:113-142): rating > 3.5, users with >= min_uc kept ratings, ids
    re-numbered by first appearance.  The reference's own loader does not run under this image's pandas (its
    groupby(as_index=False).size() indexing, movielens.py:64-85) and needs h5py, so this is checked against an
    independent numpy restatement: parity unpinned for this function."""
    from spotlight.datasets.movielens import interactions_from_ratings
    rs = np.random.RandomState(0)
    n = 5000
    users, items = rs.randint(100, 400, n), rs.randint(1000, 1300, n)
    ratings, stamps = rs.randint(1, 6, n).astype(np.float64), rs.randint(0, 10 ** 6, n)
    data, itemcount = interactions_from_ratings(users, items, ratings, stamps, min_uc=5)
    keep = ratings > 3.5
    u, i, r, t = users[keep], items[keep], ratings[keep], stamps[keep]
    ids, counts = np.unique(u, return_counts=True)
    ok = np.isin(u, ids[counts >= 5])
    u, i, r, t = u[ok], i[ok], r[ok], t[ok]

    def first_appearance_codes(x):
        _, first = np.unique(x, return_index=True)
        order = {v: c for c, v in enumerate(x[np.sort(first)])}
        return np.array([order[v] for v in x])
    np.testing.assert_array_equal(data.user_ids, first_appearance_codes(u))
    np.testing.assert_array_equal(data.item_ids, first_appearance_codes(i))
    np.testing.assert_array_equal(data.ratings, r)
    np.testing.assert_array_equal(data.timestamps, t)
    assert data.num_users == len(np.unique(u)) and data.num_items == len(np.unique(i))
    iid, icount = np.unique(i, return_counts=True)
    np.testing.assert_array_equal(itemcount.sort_index().values, icount)


def test_cli_flags_match_reference(g, monkeypatch):
    from utils.arg_extractor import get_args
    monkeypatch.setattr(sys, 'argv', ['mf_spotlight.py'])
    defaults = vars(get_args())
    assert json.dumps(defaults, sort_keys=True) == str(g['arg_defaults'])
    monkeypatch.setattr(sys, 'argv', ['x', '--use_gpu', 'True', '--dataset', '20M', '--k', '10', '--rmse', 'no',
                                      '--optim', 'rms', '--learning_rate', '0.01'])
    a = get_args()
    assert (a.use_gpu, a.dataset, a.k, a.rmse, a.optim, a.learning_rate) == (True, '20M', 10, False, 'rms', 0.01)


def test_mlp_and_neumf_modules_match_reference(golden_dir):
    """the drop-in MLP / NeuMF (`representation=` modules) have the reference's constructor, the same
    parameter names and shapes, the same initialisation under a fixed torch seed, and the same eval-mode outputs."""
    import torch
    from spotlight.dnn_models.mlp import MLP
    from spotlight.dnn_models.neuMF import NeuMF
    torch.set_num_threads(1)
    g = np.load(os.path.join(golden_dir, 'dnn_models.npz'))
    U, I = [int(x) for x in g['meta']]
    torch.manual_seed(5)
    mlp = MLP(layers=[32, 16, 8], num_users=U, num_items=I, embedding_dim=16)
    torch.manual_seed(6)
    neumf = NeuMF(mlp_layers=[24, 12], num_users=U, num_items=I, mf_embedding_dim=10, mlp_embedding_dim=12)
    users, items = torch.from_numpy(g['users']), torch.from_numpy(g['items'])
    for name, net in (('mlp', mlp), ('neumf', neumf)):
        net.eval()
        sd = net.state_dict()
        assert list(sd.keys()) == [str(k) for k in g[name + '_keys']]
        for k, v in sd.items():
            np.testing.assert_array_equal(v.numpy(), g['%s/%s' % (name, k)], err_msg=k)     # same init draws
        with torch.no_grad():
            np.testing.assert_array_equal(net(users, items).numpy(), g[name + '_out'])
