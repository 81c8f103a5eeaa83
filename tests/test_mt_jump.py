"""CPU-only: the GF(2) jump polynomials behind the multi-CTA MT19937 generator (csrc/mfb_mt_jump.cu).

g_J = x^J mod phi (phi = characteristic polynomial of MT19937's one-word transition, found by Berlekamp-Massey in the
library) must satisfy, for the word stream w[.] of ANY MT19937 instance,
        w[n + J + j] == XOR over the set bits i of g_J of w[n + i + j]
(tempering is linear and word-wise, so the identity holds for the tempered outputs numpy hands out).  The stream here
is numpy's legacy RandomState -- the generator the reference's sampling code uses (spotlight/sampling.py:33)."""
import ctypes

import numpy as np
import pytest

from recommendation_gans_b200 import _native as N


def jump_poly_bits(lib, jump):
    buf = (ctypes.c_uint64 * 312)()
    rc = lib.mfb_mt_jump_poly(ctypes.c_int64(jump), ctypes.cast(buf, ctypes.c_void_p))
    assert rc == 0
    words = np.frombuffer(buf, dtype=np.uint64).copy()
    bits = np.unpackbits(words.view(np.uint8), bitorder='little')
    return bits


@pytest.mark.parametrize('jump', [0, 1, 623, 624, 19937, 65536, 1_000_003])
def test_jump_polynomial_reproduces_the_stream(jump):
    lib = N.load_library()
    bits = jump_poly_bits(lib, jump)
    assert bits.shape[0] == 312 * 64 and not bits[19937:].any()       # degree < 19937
    idx = np.nonzero(bits)[0]
    for seed, start in ((12345, 0), (7, 1001)):
        rs = np.random.RandomState(seed)
        n = start + jump + 19937 + 700
        w = rs.randint(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)   # one tempered word per draw
        for j in (0, 1, 311, 623):
            got = np.bitwise_xor.reduce(w[start + idx + j])
            assert got == w[start + jump + j], (jump, seed, j)


def test_small_jumps_are_monomials():
    lib = N.load_library()
    for jump in (0, 5, 19936):
        bits = jump_poly_bits(lib, jump)
        assert bits.sum() == 1 and bits[jump] == 1
