"""Philox4x32-10 counter-based streams: the RNG contract of the native env.

The reference draws from one shared ``numpy.random.Generator`` (plus the legacy
global ``np.random`` in ``traffic_generators.py:66,96-97``).  The native env
replaces that with one counter-based stream per (env, slice, purpose):

    key     = (seed_lo, seed_hi)            # the 64-bit base seed of the batch (create_env's rng / seed)
    counter = (draw_index, stream_id, slice_index, global env id)

(The env id sits in the counter, not in the key: batches created with adjacent integer seeds -- the
reference's replication pattern ``default_rng(seed=i) for i in range(RUNS)``, experiments_kbrl.py:47 --
share no streams, and results stay invariant to how a batch is sharded over devices.)

Every variate consumes ONE counter tick (``normal`` and ``random(2)`` consume two).
The uniform->variate transforms below are the definition; the C oracle
(``oracle/ranslice_oracle.c``) and the CUDA kernels (``csrc/philox.cuh``) restate
them, and ``tests/refharness.py`` injects :class:`PhiloxStream` objects into the
unmodified reference so that all three can be compared draw for draw.
"""
import math

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF

# stream ids (purpose of the draw inside one slice)
STREAM_RAN = 0    # SliceRANeMBB.rng: inter-arrival / holding times   (slice_ran.py:208,220,238,243)
STREAM_CHAN = 1   # SINRSelectiveFading.rng + NominalSINR.rng          (channel_models.py:164-167,180-181,72,91)
STREAM_L1RX = 2   # SliceL1eMBB.rng: Bernoulli reception               (slice_l1.py:223)
STREAM_VBR = 3    # global np.random in VbrSource                      (traffic_generators.py:66,96-97)
STREAM_MTC = 4    # SliceRANmMTC.rng: device population at reset       (slice_ran.py:97-100)
STREAM_KBRL = 5   # global np.random.choice tie-break in GaussianKernel.predict (kernel.py:27)
STREAM_POLICY = 6 # bench/test action policy (not part of the env)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> 32, p0 & MASK
        hi1, lo1 = p1 >> 32, p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def u01(x0, x1):
    """53-bit uniform in [0,1) from two 32-bit words (exact in fp64)."""
    return ((x0 >> 5) * 67108864.0 + (x1 >> 6)) / 9007199254740992.0


class PhiloxStream:
    """Duck-typed stand-in for ``numpy.random.Generator`` (methods the reference calls)."""

    def __init__(self, seed, slice_index, stream_id, counter=0, env=0):
        self.k0 = seed & MASK
        self.k1 = (seed >> 32) & MASK
        self.slice_index = slice_index
        self.stream_id = stream_id
        self.env = env & MASK
        self.n = counter

    def _raw(self):
        out = philox4x32_10(self.n & MASK, self.stream_id, self.slice_index, self.env, self.k0, self.k1)
        self.n += 1
        return out

    def _u(self):
        x = self._raw()
        return u01(x[0], x[1])

    def random(self, size=None):
        if size is None:
            return self._u()
        import numpy as np
        return np.array([self._u() for _ in range(int(size))], dtype=np.float64)

    def exponential(self, scale=1.0):
        return -math.log(1.0 - self._u()) * scale

    def integers(self, low, high=None):
        if high is None:
            low, high = 0, low
        n = int(high) - int(low)
        return int(low) + ((self._raw()[0] * n) >> 32)

    def choice(self, seq):
        return seq[self.integers(len(seq))]

    def normal(self, mu=0.0, sigma=1.0):
        u1 = self._u()
        u2 = self._u()
        z = math.sqrt(-2.0 * math.log(1.0 - u1)) * math.cos(6.283185307179586 * u2)
        return mu + sigma * z
