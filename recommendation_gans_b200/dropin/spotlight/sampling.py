"""Negative sampling (reference: spotlight/sampling.py:9-70).

`sample_items` keeps the fork's five-argument signature (the first two arguments are ignored there
too) and returns exactly the ids `random_state.randint(0, num_items, shape, dtype=np.int64)` would:
the MT19937 stream of the given numpy RandomState is continued on the GPU (masked rejection on
32-bit draws, mfb_mt_sample_items) and the RandomState is left where numpy would have left it.
"""
import logging
import time

import numpy as np

from recommendation_gans_b200.engine import sample_items_device


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


def get_negative_samples(train, num_samples):
    """Offline (user, item) negative-pair list (sampling.py:46-70): uniform pairs, re-drawn inside the
    user's complement when the pair is a known interaction.  One-off host preprocessing; same global
    numpy RNG consumption order as the reference."""
    csr = train.tocsr()
    logging.info("Generating %d Samples" % num_samples)
    start = time.time()
    users = np.random.choice(train.num_users, num_samples)
    items = np.random.choice(train.num_items, num_samples)
    pairs = []
    for u, i in zip(users, items):
        if train.has_key(u, i):
            i = negsamp_vectorized_bsearch_preverif(csr[u, :].toarray().nonzero()[1], train.num_items, 1)[0]
        pairs.append((u, i))
    logging.info("Took %d seconds" % (time.time() - start))
    return pairs
