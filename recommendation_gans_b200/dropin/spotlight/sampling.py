"""Negative sampling (reference: spotlight/sampling.py:9-70).

`sample_items` keeps the fork's five-argument signature (the first two arguments are ignored there
too) and returns exactly the ids `random_state.randint(0, num_items, shape, dtype=np.int64)` would:
the MT19937 stream of the given numpy RandomState is continued on the GPU (masked rejection on
32-bit draws, mfb_mt_sample_items) and the RandomState is left where numpy would have left it.
"""
import logging
import time

import numpy as np

from recommendation_gans_b200.engine import negative_pairs_device, sample_items_device


def sample_items(interaction, user_ids, num_items, shape, random_state=None):
    if random_state is None:
        random_state = np.random.RandomState()
    if not isinstance(random_state, np.random.RandomState):
        # duck-typed generator supplied by the caller (sampling.py:31-33 only calls .randint)
        return random_state.randint(0, num_items, shape, dtype=np.int64)
    count = int(np.prod(shape))
    out = sample_items_device(num_items, count, random_state)
    return out.cpu().numpy().reshape(shape)


def negsamp_vectorized_bsearch_preverif(pos_inds, n_items, n_samp=32):
    """Uniform ids outside the sorted `pos_inds` (sampling.py:37-44); host preprocessing helper."""
    raw = np.random.randint(0, n_items - len(pos_inds), size=n_samp)
    shifted = pos_inds - np.arange(len(pos_inds))
    return raw + np.searchsorted(shifted, raw, side='right')


def get_negative_samples_arrays(train, num_samples, random_state=None):
    """get_negative_samples as two int64 numpy arrays (users, items).  The whole job runs on the GPU
    (mfb_negative_pairs): two uniform index streams, a CSR membership test per pair, and for the pairs that are known
    interactions a re-draw inside the user's complement -- consuming numpy's legacy MT19937 stream (the global
    generator unless `random_state` is given) exactly as the reference's Python loop does."""
    csr = train.tocsr()
    csr.sum_duplicates()
    csr.sort_indices()
    keys = csr.copy()
    keys.data = (keys.data == 1).astype(np.float64)      # Interactions.has_key: stored value == 1 (interactions.py:159)
    keys.eliminate_zeros()
    rs = random_state if random_state is not None else np.random.mtrand._rand
    users, items, _ = negative_pairs_device(keys, csr, train.num_users, train.num_items, num_samples, rs)
    return users.cpu().numpy(), items.cpu().numpy()


def get_negative_samples(train, num_samples):
    """Offline (user, item) negative-pair list (sampling.py:46-70): uniform pairs, re-drawn inside the
    user's complement when the pair is a known interaction.  Same pairs and same consumption of numpy's global
    generator as the reference; returns the reference's list of tuples."""
    logging.info("Generating %d Samples" % num_samples)
    start = time.time()
    users, items = get_negative_samples_arrays(train, num_samples)
    pairs = list(zip(users, items))
    logging.info("Took %d seconds" % (time.time() - start))
    return pairs
