"""Seeded synthetic (x, mean, scale) generators shared by the tests and the benchmark.

`test`  : the reference's own self-test distribution, rans/test.py:8-10
          (mean = randint(-256,256)/256, scale = exp(U(-5,5))/256, x = round((mean+scale*U(-5,5))*256)/256)
`coder` : coder.py:45-47 (mean in [-32,32]/256, scale ~ 1, x within half a scale of the mean)
`wide`  : logistic samples with scale from e^-12 to e^4 clipped to the window: exercises tiny
          frequencies (freq down to 1-3), saturated tails and window edges
`edges` : symbols pinned to the first / last bins of the window
`needle`: scales from e^-20 to e^-9 (2e-9 .. 1.2e-4, far below a bin): the CDF argument at the
          symbol's own bin edges runs from ~30 to beyond 10^5, past the [-128, 128] limit of the
          float argument and past what the decoder's unlimited first try accepts (680)
"""
import numpy as np


def c_round(v):
    """C round(): half away from zero (numpy's round is half to even)."""
    v = np.asarray(v, np.float64)
    return np.where(v >= 0, np.floor(v + 0.5), np.ceil(v - 0.5))


def gen(kind: str, n: int, seed: int):
    rng = np.random.default_rng(seed)
    if kind == "test":
        mean = (rng.integers(-256, 257, n) / 256).astype(np.float32)
        scale = (np.exp(10 * rng.random(n) - 5) / 256).astype(np.float32)
        x = np.round((mean.astype(np.float64) + scale.astype(np.float64) * (10 * rng.random(n) - 5)) * 256) / 256
    elif kind == "coder":
        mean = (rng.integers(-32, 33, n) / 256).astype(np.float32)
        scale = np.exp(rng.random(n) * 0.01 - 0.005).astype(np.float32)
        x = np.round((mean.astype(np.float64) + scale.astype(np.float64) * (rng.random(n) - 0.5)) * 256) / 256
    elif kind == "wide":
        mean = rng.normal(0, 1.5, n).astype(np.float32)
        scale = np.exp(rng.uniform(-12, 4, n)).astype(np.float32)
        u = rng.random(n)
        x = np.round((mean.astype(np.float64) + scale.astype(np.float64) * np.log(u / (1 - u))) * 256) / 256
        lo = c_round(mean.astype(np.float64) * 256 - 1024)
        x = np.clip(x * 256, lo, lo + 2047) / 256
    elif kind == "edges":
        mean = rng.normal(0, 3.0, n).astype(np.float32)
        scale = np.exp(rng.uniform(-3, 3, n)).astype(np.float32)
        lo = c_round(mean.astype(np.float64) * 256 - 1024)
        pick = rng.integers(0, 4, n)
        x = np.where(pick == 0, lo, np.where(pick == 1, lo + 2047, np.where(pick == 2, lo + 1, lo + 2046))) / 256
    elif kind == "needle":
        mean = rng.normal(0, 1.5, n).astype(np.float32)
        scale = np.exp(rng.uniform(-20, -9, n)).astype(np.float32)
        x = (c_round(mean.astype(np.float64) * 256) + rng.integers(-1, 2, n) * (rng.random(n) < 0.1)) / 256
    else:
        raise ValueError(kind)
    return x.astype(np.float32), mean, scale


def ragged_offsets(n: int, n_streams: int, seed: int, allow_empty: bool = True):
    """Random partition of n symbols into n_streams contiguous streams (some empty)."""
    rng = np.random.default_rng(seed)
    cuts = np.sort(rng.integers(0, n + 1, n_streams - 1)) if n_streams > 1 else np.zeros(0, np.int64)
    off = np.concatenate([[0], cuts, [n]]).astype(np.int64)
    if not allow_empty:
        off = np.unique(off)
    return off
