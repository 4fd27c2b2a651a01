"""Philox4x32-10 counter-based generator: the ONE random stream shared by the
CPU oracle (Python shim back-end "philox", the C restatement) and the CUDA
level-generation kernel.

TEST INFRASTRUCTURE ONLY (see oracle/README.md): nothing in babyai_b200/ may
import this module.

Stream definition (our design; the reference uses numpy's MT19937 through
gym.utils.seeding -- see oracle/shim/gym/utils/seeding.py for that back-end):

  key      = (seed & 0xffffffff, seed >> 32)          per-env 64-bit seed
  counter  = (blk & 0xffffffff, blk >> 32, 0, 0)      blk = draw_index >> 2
  u32(i)   = philox4x32_10(counter, key)[i & 3]       i = draw_index, 0,1,2,...

  randint(lo, hi): n = hi - lo; if n == 1 -> lo, NO draw is consumed (mirrors
                   numpy legacy RandomState.randint(a, a+1), which consumes
                   nothing); else lo + ((u32 * n) >> 32)   (one draw)
  uniform(lo, hi): lo + (hi - lo) * (u32 * 2**-32) in float64 (one draw)

The constants are the Random123 / cuRAND Philox4x32-10 ones.
"""

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> 32, p0 & MASK
        hi1, lo1 = p1 >> 32, p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return (c0, c1, c2, c3)


class PhiloxRandom:
    """Drop-in for the three np.random.RandomState methods MiniGrid uses."""

    def __init__(self, seed=0, draws=0):
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.draws = int(draws)

    def _u32(self):
        i = self.draws
        self.draws += 1
        blk = i >> 2
        out = philox4x32_10((blk & MASK, (blk >> 32) & MASK, 0, 0),
                            (self.seed & MASK, self.seed >> 32))
        return out[i & 3]

    def randint(self, low, high=None):
        if high is None:
            low, high = 0, low
        n = int(high) - int(low)
        assert n >= 1
        if n == 1:
            return int(low)
        return int(low) + ((self._u32() * n) >> 32)

    def uniform(self, low=0.0, high=1.0):
        return low + (high - low) * (self._u32() * 2.0 ** -32)
