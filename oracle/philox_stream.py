"""TEST INFRASTRUCTURE (oracle side). Counter-based sample stream shared by the CPU oracle and the CUDA path.

The reference draws from CPython's global `random` (MT19937) at `rrt_star.py:65-66`, `neural_rrt.py:68-69,83,225`,
`ThreeD_rrt_star.py:66-68`, `ThreeD_neural_rrt.py:72-74,91,241`.  A Mersenne Twister cannot be indexed per task on a
GPU, so both sides of every parity test are driven by the SAME counter-based stream instead (SURVEY.md §8d "Sample
stream"): Philox-4x32-10, key = 64-bit seed, counter = (block, task, tag), two 53-bit uniforms per block.  The
device implementation is `mgr_sampling_cnn_b200/csrc/philox.cuh`; the C one is in `oracle/planner_oracle.c`.

    draw k of task t:  blk = k >> 1;  (o0,o1,o2,o3) = philox4x32_10(ctr=(blk_lo, blk_hi, t, 0x4E525254), key=(seed_lo, seed_hi))
                       (a,b) = (o0,o1) if k even else (o2,o3)
                       u = ((a >> 5) * 2**26 + (b >> 6)) / 2**53            in [0,1)
    random()      = u
    uniform(a,b)  = a + (b-a)*u                                            (CPython's formula, random.py `uniform`)
    randint(a,b)  = a + min(floor(u * (b-a+1)), b-a)                       (fp64 multiply, then floor)

`StreamRandom` is assigned over the reference module's `random` attribute (`module.random = StreamRandom(...)`), which
is the injection point named in SURVEY.md §8b.
"""
import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
STREAM_TAG = 0x4E525254  # 'NRRT'
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox-4x32-10.  All arguments broadcastable uint64 arrays holding 32-bit values."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64)
    c3 = np.asarray(c3, dtype=np.uint64)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = PHILOX_M0 * c0  # < 2**64, no overflow
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n1 = lo1
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        n3 = lo0
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def uniforms(seed: int, task: int, first: int, count: int) -> np.ndarray:
    """fp64 uniforms for draws [first, first+count) of `task` under `seed`."""
    k = np.arange(first, first + count, dtype=np.uint64)
    blk = k >> np.uint64(1)
    o0, o1, o2, o3 = philox4x32_10(blk & _MASK32, blk >> np.uint64(32), np.uint64(task & 0xFFFFFFFF),
                                   np.uint64(STREAM_TAG), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    odd = (k & np.uint64(1)).astype(bool)
    a = np.where(odd, o2, o0)
    b = np.where(odd, o3, o1)
    r53 = (a >> np.uint64(5)) * np.uint64(67108864) + (b >> np.uint64(6))
    return r53.astype(np.float64) / 9007199254740992.0


class StreamRandom(object):
    """Drop-in for the `random` module attribute of the reference planner modules."""

    CHUNK = 4096

    def __init__(self, seed: int, task: int):
        self.seed = int(seed)
        self.task = int(task)
        self.draws = 0
        self._buf = None
        self._buf_first = 0

    def _next(self) -> float:
        k = self.draws
        if self._buf is None or k >= self._buf_first + len(self._buf):
            self._buf_first = k
            self._buf = uniforms(self.seed, self.task, k, self.CHUNK)
        self.draws = k + 1
        return float(self._buf[k - self._buf_first])

    def random(self) -> float:
        return self._next()

    def uniform(self, a, b) -> float:
        return a + (b - a) * self._next()

    def randint(self, a: int, b: int) -> int:
        n = b - a + 1
        u = self._next()
        return a + min(int(np.floor(u * n)), n - 1)
