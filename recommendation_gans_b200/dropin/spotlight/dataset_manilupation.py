"""Splitting and shuffling of interaction sets (reference: spotlight/dataset_manilupation.py:11-236).
Host-side preprocessing: numpy only.  (The module name keeps the reference's spelling.)"""
import numpy as np

from spotlight.interactions import Interactions


def _take(interactions, index):
    """A new Interactions holding the selected entries of every per-interaction array."""
    pick = lambda arr: None if arr is None else arr[index]      # noqa: E731
    return Interactions(interactions.user_ids[index], interactions.item_ids[index], ratings=pick(interactions.ratings),
                        timestamps=pick(interactions.timestamps), weights=pick(interactions.weights),
                        num_users=interactions.num_users, num_items=interactions.num_items)


def shuffle_interactions(interactions, random_state=None):
    """dataset_manilupation.py:20-58: one permutation from `random_state.shuffle(arange(n))`."""
    if random_state is None:
        random_state = np.random.RandomState()
    order = np.arange(len(interactions.user_ids))
    random_state.shuffle(order)
    return _take(interactions, order)


def _cut(interactions, test_percentage):
    cutoff = int((1.0 - test_percentage) * len(interactions))
    return _take(interactions, slice(None, cutoff)), _take(interactions, slice(cutoff, None))


def random_train_test_split(interactions, test_percentage=0.2, random_state=None):
    """dataset_manilupation.py:60-119: shuffle, then cut at int((1 - test_percentage) * n)."""
    return _cut(shuffle_interactions(interactions, random_state=random_state), test_percentage)


def user_based_train_test_split(interactions, test_percentage=0.2, random_state=None):
    """dataset_manilupation.py:121-175: a user's whole history goes to one side, decided by a seeded murmur hash."""
    from sklearn.utils import murmurhash3_32
    if random_state is None:
        random_state = np.random.RandomState()
    seed = random_state.randint(np.iinfo(np.uint32).min, np.iinfo(np.uint32).max, dtype=np.int64)
    in_test = (murmurhash3_32(interactions.user_ids, seed=seed, positive=True) % 100 / 100.0) < test_percentage
    return _take(interactions, np.logical_not(in_test)), _take(interactions, in_test)


def train_test_timebased_split(interactions, test_percentage=0.2):
    """dataset_manilupation.py:177-236: the earliest (1 - test_percentage) of the interactions train, the rest test.
    As in the reference the input object is re-ordered IN PLACE by `timestamps.argsort()` -- user ids, item ids and
    timestamps, but NOT ratings or weights, which are sliced in their original order."""
    order = interactions.timestamps.argsort()
    interactions.user_ids = interactions.user_ids[order]
    interactions.item_ids = interactions.item_ids[order]
    interactions.timestamps = interactions.timestamps[order]
    return _cut(interactions, test_percentage)


# ---- helpers of the slate (GAN) data path: outside the accelerated hot path, provided so that the reference's other
# ---- scripts keep importing this module's names when the drop-in shadows it -------------------------------------
def delete_rows_csr(mat, row_indices=[], col_indices=[]):
    """dataset_manilupation.py:238-268: the CSR matrix without the listed rows / columns (axes are re-indexed)."""
    from scipy.sparse import csr_matrix
    if not isinstance(mat, csr_matrix):
        raise ValueError("works only for CSR format -- use .tocsr() first")
    rows, cols = list(row_indices or []), list(col_indices or [])
    if rows:
        keep = np.ones(mat.shape[0], dtype=bool)
        keep[rows] = False
        mat = mat[keep]
    if cols:
        keep = np.ones(mat.shape[1], dtype=bool)
        keep[cols] = False
        mat = mat[:, keep]
    return mat


def create_slates(interactions, n=5, padding_value=0):
    """dataset_manilupation.py:270-318: every user's last n interactions (by timestamp) become that user's slate and
    leave the interaction set; users with fewer than n interactions are dropped.  Returns (CSR of the remaining
    interactions without the slate-less users' rows, slates [users_with_slates, n])."""
    num_users = interactions.num_users
    slates = np.zeros((num_users, n))
    order = np.lexsort((interactions.timestamps, interactions.user_ids))       # by user, then time (stable)
    users_sorted = interactions.user_ids[order]
    starts = np.searchsorted(users_sorted, np.arange(num_users), side='left')
    ends = np.searchsorted(users_sorted, np.arange(num_users), side='right')
    drop = []
    for user in range(num_users):
        idx = order[starts[user]:ends[user]]
        if len(idx) == 0:
            continue
        if len(idx) < n:
            drop.extend(idx.tolist())
            continue
        slates[user] = interactions.item_ids[idx[-n:]]
        drop.extend(idx[-n:].tolist())
    for name in ('user_ids', 'item_ids', 'timestamps', 'ratings'):
        setattr(interactions, name, np.delete(getattr(interactions, name), drop))
    empty = np.where(~slates.any(axis=1))[0]
    slates = np.delete(slates, empty, axis=0)
    return delete_rows_csr(interactions.tocsr(), row_indices=list(empty)), slates


def train_test_split(interactions, test_percentage=0.2):
    """dataset_manilupation.py:320-364: per user, a random test_percentage of the interactions (np.random.choice WITH
    replacement over the user's items, as the reference draws them) moves to the test set; dense intermediate."""
    from scipy.sparse import coo_matrix
    dense = np.asarray(interactions.tocsr().todense())
    test = np.zeros(dense.shape)
    train = dense.copy()
    for user in range(dense.shape[0]):
        items = dense[user].nonzero()[0]
        picked = np.random.choice(items, int(items.shape[0] * test_percentage))
        train[user, picked] = 0
        test[user, picked] = dense[user, picked]
    assert np.all(train * test == 0)
    out = []
    for part in (train, test):
        coo = coo_matrix(part)
        out.append(Interactions(coo.row, coo.col, coo.data, timestamps=None, weights=None, num_users=dense.shape[0],
                                num_items=dense.shape[1]))
    return out[0], out[1]
