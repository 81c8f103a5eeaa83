"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement of the two MT19937 index streams the reference draws its
negatives from.  Parity status: PINNED -- `tests/test_oracle_rng.py` checks
every function here bit-for-bit against CPython's `random` and numpy's
`RandomState` (the third-party code the reference calls) and against the
golden vectors under `tests/golden/` that `oracle/make_golden.py` produced by
running the reference itself.

Reference call sites restated here
  * /root/reference/implicit.py:352,370
        random.choices(self.neg_examples, k=num_negative_samples*batch_size)
    -> CPython Lib/random.py `choices` (no weights):
        population[floor(random() * n)]  with n = len(population) + 0.0
       CPython Modules/_randommodule.c `random_random`:
        a = genrand_uint32() >> 5;  b = genrand_uint32() >> 6
        x = (a * 67108864.0 + b) * (1.0 / 9007199254740992.0)
  * /root/reference/spotlight/sampling.py:33
        random_state.randint(0, num_items, shape, dtype=np.int64)
    -> numpy legacy RandomState, masked rejection on 32-bit draws
       (numpy/random/src/distributions/distributions.c
        `buffered_bounded_masked_uint32`, rng = num_items-1 <= 0xFFFFFFFF):
        mask = smallest 2^k-1 >= rng;  draw w = next_uint32() & mask until w <= rng.
  * /root/reference/spotlight/torch_utils.py:38-55  `shuffle` (host side; uses
    RandomState.shuffle -- restated for completeness as `legacy_shuffle_indices`).

MT19937 itself (Matsumoto & Nishimura 1998) is restated from the published
recurrence: N=624, M=397, MATRIX_A=0x9908b0df, tempering (11; 7,0x9d2c5680;
15,0xefc60000; 18).
"""
import numpy as np

N = 624
M = 397
MATRIX_A = 0x9908B0DF
UPPER = 0x80000000
LOWER = 0x7FFFFFFF


def init_genrand(seed):
    """Knuth-style LCG fill used by numpy `RandomState(seed)` for integer seeds."""
    mt = np.zeros(N, dtype=np.uint64)
    mt[0] = seed & 0xFFFFFFFF
    for i in range(1, N):
        prev = int(mt[i - 1])
        mt[i] = (1812433253 * (prev ^ (prev >> 30)) + i) & 0xFFFFFFFF
    return mt.astype(np.uint32)


def init_by_array(key):
    """CPython `random.seed(int)` path: key = 32-bit little-endian chunks of abs(seed)."""
    mt = [int(x) for x in init_genrand(19650218)]
    i, j = 1, 0
    klen = len(key)
    for _ in range(max(N, klen)):
        mt[i] = ((mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525)) + key[j] + j) & 0xFFFFFFFF
        i += 1
        j += 1
        if i >= N:
            mt[0] = mt[N - 1]
            i = 1
        if j >= klen:
            j = 0
    for _ in range(N - 1):
        mt[i] = ((mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941)) - i) & 0xFFFFFFFF
        i += 1
        if i >= N:
            mt[0] = mt[N - 1]
            i = 1
    mt[0] = 0x80000000
    return np.array(mt, dtype=np.uint32)


def python_seed_key(seed):
    """32-bit chunks of abs(seed), least significant first (CPython random_seed)."""
    seed = abs(int(seed))
    key = []
    while True:
        key.append(seed & 0xFFFFFFFF)
        seed >>= 32
        if seed == 0:
            break
    return key


def twist(mt):
    """One full regeneration of the 624-word state (in place semantics, returns new array)."""
    mt = [int(x) for x in mt]
    for kk in range(N):
        y = (mt[kk] & UPPER) | (mt[(kk + 1) % N] & LOWER)
        mt[kk] = mt[(kk + M) % N] ^ (y >> 1) ^ (MATRIX_A if (y & 1) else 0)
    return np.array(mt, dtype=np.uint32)


def temper(y):
    y = np.asarray(y, dtype=np.uint32).copy()
    y ^= y >> np.uint32(11)
    y ^= (y << np.uint32(7)) & np.uint32(0x9D2C5680)
    y ^= (y << np.uint32(15)) & np.uint32(0xEFC60000)
    y ^= y >> np.uint32(18)
    return y


class MT19937:
    """State = (624 words, pos) exactly as `random.getstate()[1]` / `RandomState.get_state()[1:3]`."""

    def __init__(self, state, pos):
        self.mt = np.array(state, dtype=np.uint32).copy()
        self.pos = int(pos)

    @classmethod
    def from_python_seed(cls, seed):
        return cls(init_by_array(python_seed_key(seed)), N)

    @classmethod
    def from_numpy_seed(cls, seed):
        return cls(init_genrand(seed), N)

    def words(self, n):
        """Next n tempered 32-bit outputs."""
        out = np.empty(n, dtype=np.uint32)
        filled = 0
        while filled < n:
            if self.pos >= N:
                self.mt = twist(self.mt)
                self.pos = 0
            take = min(n - filled, N - self.pos)
            out[filled:filled + take] = temper(self.mt[self.pos:self.pos + take])
            self.pos += take
            filled += take
        return out


def choices_indices(gen, population_len, k):
    """Index stream of `random.choices(pop, k=k)` (implicit.py:352): 2 words per sample."""
    w = gen.words(2 * k).astype(np.uint64)
    a = w[0::2] >> np.uint64(5)
    b = w[1::2] >> np.uint64(6)
    x = (a.astype(np.float64) * 67108864.0 + b.astype(np.float64)) * (1.0 / 9007199254740992.0)
    return np.floor(x * (population_len + 0.0)).astype(np.int64)


def randint_masked(gen, num_items, count):
    """`RandomState.randint(0, num_items, count, dtype=np.int64)` (sampling.py:33)."""
    rng = num_items - 1
    if rng == 0:
        return np.zeros(count, dtype=np.int64)
    mask = rng
    for s in (1, 2, 4, 8, 16):
        mask |= mask >> s
    out = np.empty(count, dtype=np.int64)
    got = 0
    while got < count:
        # draw in bulk; rejected words are consumed, accepted words kept in order
        need = count - got
        w = gen.words(need)
        v = w & np.uint32(mask)
        ok = v <= rng
        acc = v[ok]
        out[got:got + len(acc)] = acc
        got += len(acc)
    return out


def legacy_shuffle_indices(gen, n):
    """`RandomState.shuffle(np.arange(n))` (torch_utils.py:50-51): Fisher-Yates from the top,
    j = random_interval(i) = masked rejection on 32-bit words (64-bit if i > 0xFFFFFFFF, not restated)."""
    idx = np.arange(n)
    for i in range(n - 1, 0, -1):
        mask = i
        for s in (1, 2, 4, 8, 16):
            mask |= mask >> s
        while True:
            j = int(gen.words(1)[0]) & mask
            if j <= i:
                break
        idx[i], idx[j] = idx[j], idx[i]
    return idx
