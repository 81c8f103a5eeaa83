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
