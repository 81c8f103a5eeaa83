"""Oracle index streams vs CPython `random` / numpy `RandomState` (live) and vs golden vectors
produced by the reference (`oracle/make_golden.py`).  Integer work: bit-exact."""
import os
import random

import numpy as np

from oracle import mt19937_ref as R


def test_choices_matches_cpython_and_keeps_state():
    for seed in (0, 5, 2 ** 70 + 17):
        random.seed(seed)
        gen = R.MT19937.from_python_seed(seed)
        st = random.getstate()[1]
        assert tuple(gen.mt.tolist()) == st[:624] and st[624] == gen.pos
        for k in (1, 333, 5120):
            exp = random.choices(range(81000), k=k)
            assert R.choices_indices(gen, 81000, k).tolist() == exp
        st = random.getstate()[1]
        assert tuple(gen.mt.tolist()) == st[:624] and st[624] == gen.pos


def test_randint_matches_numpy_and_keeps_state():
    for seed, n in ((0, 1682), (1, 3706), (7, 26744), (3, 2000000), (4, 1), (5, 2), (6, 1025), (8, 1024)):
        rs = np.random.RandomState(seed)
        gen = R.MT19937.from_numpy_seed(seed)
        for cnt in (1, 700, 5000):
            exp = rs.randint(0, n, cnt, dtype=np.int64)
            assert (R.randint_masked(gen, n, cnt) == exp).all()
        s = rs.get_state()
        assert (s[1] == gen.mt).all() and s[2] == gen.pos


def test_shuffle_matches_numpy():
    rs = np.random.RandomState(0)
    idx = np.arange(2000)
    rs.shuffle(idx)
    assert (R.legacy_shuffle_indices(R.MT19937.from_numpy_seed(0), 2000) == idx).all()


def test_golden_streams(golden_dir):
    g = np.load(os.path.join(golden_dir, 'rng_streams.npz'))
    for seed, n, cnt in ((0, 1682, 4096), (1, 3706, 4096), (2, 26744, 8192), (3, 2000000, 4096),
                         (4, 1, 16), (5, 2, 64), (6, 1025, 2000)):
        got = R.randint_masked(R.MT19937.from_numpy_seed(seed), n, cnt)
        assert (got == g['sample_items_s%d_n%d' % (seed, n)]).all()
    for seed, L, k in ((0, 80000, 5120), (5, 81000, 3000), (2 ** 70 + 17, 16200213, 16384), (9, 7, 50)):
        gen = R.MT19937.from_python_seed(seed)
        assert (R.choices_indices(gen, L, k) == g['choices_s%d_L%d' % (seed % 1000, L)]).all()
        assert (R.choices_indices(gen, L, k // 2 + 1) == g['choices_s%d_L%d_second' % (seed % 1000, L)]).all()
    # known answers recorded in SURVEY.md section 4
    assert R.randint_masked(R.MT19937.from_numpy_seed(0), 1682, 5).tolist() == [684, 559, 1653, 1216, 835]
    assert R.choices_indices(R.MT19937.from_python_seed(5), 80000, 5).tolist() == [49832, 59342, 63615, 75396, 59191]
