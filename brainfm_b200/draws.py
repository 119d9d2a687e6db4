"""Random-draw sources of the generator.

The reference draws from three global generators (numpy, python `random`, torch) in a fixed order
(SURVEY.md 8a "RNG draw inventory").  Every draw site of this package goes through one of the two
classes below so that the order is explicit and the draws can be injected:

  HostDraws    -- production: the same global generators, called in the same order as the reference for
                  every scalar / small tensor; the two volume-sized normal fields (GMM eps, noise eps)
                  are NOT materialised: they are generated in-kernel by Philox4x32-10 keyed on a
                  per-sample seed.
  ReplayDraws  -- parity tests: replays a recorded list of (tag, value) pairs (the oracle's log),
                  checking the tag of every draw, so any divergence in draw order fails loudly.
"""
import random as _pyrandom

import numpy as np
import torch


class HostDraws:
    replay = False

    def rand(self, tag):
        return np.random.rand()

    def rand1(self, tag):
        return np.random.rand(1)

    def rand3(self, tag):
        return np.random.rand(3)

    def randn(self, tag):
        return np.random.randn()

    def randn1(self, tag):
        return np.random.randn(1)[0]

    def randint(self, tag, n):
        return np.random.randint(n)

    def rand_array(self, tag, shape):
        return np.random.rand(*shape)

    def choice(self, tag, seq):
        return _pyrandom.choice(seq)

    def torch_rand(self, tag, shape, dtype=torch.float32, out=None):
        """`out` (a contiguous CPU tensor of that shape, e.g. a view of the pinned plan arena) receives the draw
        in place: same generator consumption and same values as the allocating form."""
        return torch.rand(shape, dtype=dtype) if out is None else torch.rand(shape, dtype=dtype, out=out)

    def torch_randn(self, tag, shape, out=None):
        return torch.randn(shape, dtype=torch.float32) if out is None else torch.randn(shape, dtype=torch.float32,
                                                                                       out=out)

    def field_randn(self, tag, shape=None):
        """Volume-sized N(0,1) field: None => generated in-kernel (Philox)."""
        return None

    _seed_base, _seed_count = None, 0

    def seed64(self):
        """Philox key of one sample's volume-sized fields.  Derived from ONE draw of the numpy generator (taken
        the first time a key is needed) and a counter (splitmix64), so that it does not interleave extra draws
        with the reference's scalar draw sequence."""
        if self._seed_base is None:
            self._seed_base = int(np.random.randint(0, 2 ** 62))
        self._seed_count += 1
        z = (self._seed_base + 0x9E3779B97F4A7C15 * self._seed_count) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)


class ReplayDraws(HostDraws):
    replay = True

    def __init__(self, log):
        self.log = list(log)
        self.pos = 0

    def _next(self, tag):
        if self.pos >= len(self.log):
            raise AssertionError("draw log exhausted at %r" % tag)
        t, v = self.log[self.pos]
        if t != tag:
            raise AssertionError("draw order mismatch: wanted %r, log has %r (position %d)" % (tag, t, self.pos))
        self.pos += 1
        return v

    def rand(self, tag):
        return self._next(tag)

    rand1 = rand3 = randn = randn1 = rand

    def randint(self, tag, n):
        return self._next(tag)

    def rand_array(self, tag, shape):
        return np.asarray(self._next(tag)).reshape(shape)

    def choice(self, tag, seq):
        return self._next(tag)

    def torch_rand(self, tag, shape, dtype=torch.float32, out=None):
        v = self._next(tag)
        return v.clone() if out is None else out.copy_(v.reshape(out.shape))

    def torch_randn(self, tag, shape, out=None):
        v = self._next(tag)
        return v.clone() if out is None else out.copy_(v.reshape(out.shape))

    def field_randn(self, tag, shape=None):
        return self._next(tag)

    def seed64(self):
        return 0

    def done(self):
        return self.pos == len(self.log)
