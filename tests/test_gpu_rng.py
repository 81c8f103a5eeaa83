"""Device MT19937 index streams vs the oracle, live CPython/numpy generators and golden vectors.
Integer work: bit-exact, including the generator state left behind."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import mt19937_ref as R

pytestmark = pytest.mark.gpu


def _eng():
    import recommendation_gans_b200  # noqa: F401
    from recommendation_gans_b200 import engine
    return engine


def test_raw_words_and_state():
    E = _eng()
    # (the last case is long enough for the model-less entry points to take the multi-CTA generator)
    for seed, pos0, n in ((0, 624, 5000), (3, 624, 1), (9, 100, 624 * 7 + 5), (4, 0, 624), (6, 77, 9 * 65536 + 123)):
        gen = R.MT19937.from_numpy_seed(seed)
        if pos0 != 624:
            gen.words(624 + pos0)          # move into the middle of a block
        state = np.empty(625, dtype=np.uint32)
        state[:624] = gen.mt
        state[624] = gen.pos
        got = E.mt_words_device(state, n).cpu().numpy().view(np.uint32)
        assert (got == gen.words(n)).all()
        assert (state[:624] == gen.mt).all() and state[624] == gen.pos


def test_parallel_generator_matches_the_sequential_stream():
    """The multi-CTA generator (shares whose starting states come from a GF(2) jump-ahead) emits the words of the one
    sequential stream and leaves the same state behind: every share count, ragged last shares, positions inside a
    block, and the fallback for short streams."""
    E = _eng()
    cases = ((0, 624, 16384 * 20, 16384),        # cfg3: 20 steps of 16 384 words -> 5 shares of 4 steps
             (3, 17, 16384 * 7 + 5, 16384),      # ragged tail, starts inside a block
             (5, 624, 131072 * 3, 131072),       # cfg5: one step per share
             (7, 300, 40000 * 9, 40000),         # share = one step of 40 000 words
             (9, 1, 20000 * 2 + 1, 20000),       # two shares, the second holds a single word... plus one
             (2, 624, 5000, 1024))               # below the threshold: the one-CTA generator
    for seed, pos0, n, wps in cases:
        gen = R.MT19937.from_numpy_seed(seed)
        if pos0 != 624:
            gen.words(624 + pos0)
        state = np.empty(625, dtype=np.uint32)
        state[:624] = gen.mt
        state[624] = gen.pos
        state_seq = state.copy()
        want = E.mt_words_device(state_seq, n).cpu().numpy().view(np.uint32)
        got = E.mt_words_device_parallel(state, n, wps).cpu().numpy().view(np.uint32)
        assert (got == want).all(), (seed, pos0, n, wps, int(np.argmax(got != want)))
        assert (state == state_seq).all()
        assert (want == gen.words(n)).all()


def test_choices_stream_matches_cpython():
    E = _eng()
    for seed in (0, 5, 2 ** 70 + 17):
        rng_dev, rng_ref = random.Random(seed), random.Random(seed)
        for L, k in ((81000, 5120), (7, 50), (16200213, 16384), (3, 1)):
            got = E.choices_indices_device(L, k, rng=rng_dev).cpu().numpy()
            assert got.tolist() == rng_ref.choices(range(L), k=k)
        assert rng_dev.getstate() == rng_ref.getstate()


def test_sample_items_matches_numpy(golden_dir):
    import recommendation_gans_b200  # noqa: F401
    from spotlight.sampling import sample_items
    g = np.load(os.path.join(golden_dir, 'rng_streams.npz'))
    for seed, n, cnt in ((0, 1682, 4096), (1, 3706, 4096), (2, 26744, 8192), (3, 2000000, 4096),
                         (4, 1, 16), (5, 2, 64), (6, 1025, 2000)):
        rs = np.random.RandomState(seed)
        got = sample_items(None, None, n, (cnt,), rs)
        assert got.dtype == np.int64 and (got == g['sample_items_s%d_n%d' % (seed, n)]).all()
        ref = np.random.RandomState(seed)
        ref.randint(0, n, cnt, dtype=np.int64)
        # the RandomState continues exactly where numpy would have left it
        assert (rs.randint(0, 10 ** 6, 100) == ref.randint(0, 10 ** 6, 100)).all()
    got = sample_items(None, None, 1682, (5,), np.random.RandomState(0))
    assert got.tolist() == [684, 559, 1653, 1216, 835]          # SURVEY section 4 known answer
    # 2-D shape and large count
    rs, ref = np.random.RandomState(11), np.random.RandomState(11)
    assert (sample_items(None, None, 26744, (300, 700), rs) == ref.randint(0, 26744, (300, 700), dtype=np.int64)).all()


def test_negative_pairs_follow_global_random(golden_dir):
    E = _eng()
    from tests.gpu_helpers import make_engine
    g = np.load(os.path.join(golden_dir, 'steps_pointwise_adam.npz'))
    U, I, D, B, n_neg = [int(x) for x in g['meta']]
    _, _, eng = make_engine([g['init%d' % i] for i in range(4)], None)
    pop_u = torch.from_numpy(g['neg_pairs'][:, 0].copy()).cuda()
    pop_i = torch.from_numpy(g['neg_pairs'][:, 1].copy()).cuda()
    random.seed(int(g['py_seed']))
    nsteps = len(g['neg_idx'])
    nu, ni = eng.draw_negative_pairs(pop_u, pop_i, nsteps * n_neg * B)
    exp = g['neg_pairs'][g['neg_idx'].reshape(-1)]
    assert (nu.cpu().numpy() == exp[:, 0]).all() and (ni.cpu().numpy() == exp[:, 1]).all()
    ref = random.Random(int(g['py_seed']))
    ref.choices(range(10), k=nsteps * n_neg * B)
    assert random.getstate() == ref.getstate()


def test_negative_pair_generator_bit_exact(golden_dir):
    """spotlight.sampling.get_negative_samples on the GPU vs the pairs the reference produced (tests/golden/
    neg_samples.npz): bit-exact pairs, numpy's global generator left where the reference leaves it; a user who
    interacted with every item raises numpy's ValueError."""
    import recommendation_gans_b200  # noqa: F401
    from spotlight.interactions import Interactions
    from spotlight.sampling import get_negative_samples, get_negative_samples_arrays
    g = np.load(os.path.join(golden_dir, 'neg_samples.npz'))
    for name in g['cases']:
        name = str(name)
        U, I, N, seed = [int(x) for x in g[name + '_meta']]
        train = Interactions(g[name + '_users'], g[name + '_items'], num_users=U, num_items=I)
        np.random.seed(seed)
        pairs = get_negative_samples(train, N)
        assert isinstance(pairs, list) and len(pairs) == N and isinstance(pairs[0], tuple)
        assert (np.array(pairs, dtype=np.int64) == g[name + '_pairs']).all()
        st = np.random.get_state()
        assert (st[1] == g[name + '_state_key']).all() and st[2] == int(g[name + '_state_pos'])
    full = Interactions(np.zeros(6, np.int32), np.arange(6, dtype=np.int32), num_users=2, num_items=6)
    with pytest.raises(ValueError):
        get_negative_samples_arrays(full, 50, np.random.RandomState(0))
