"""One 20 M-symbol stream through the reference-shaped single-stream entry points (development aid):
exercises slot growth in the host codec and shows the latency-bound single-stream rates
(B200: encode 9.4 M symbols/s, decode 3.0 M symbols/s; a stream is one serial chain)."""
import sys, time, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flic_b200 import rans
n = 20_000_000
rng = np.random.default_rng(0)
mean = (rng.integers(-256, 257, n) / 256).astype(np.float32)
scale = (np.exp(10 * rng.random(n) - 5) / 256).astype(np.float32)
x = (np.round((mean.astype(np.float64) + scale.astype(np.float64) * (10 * rng.random(n) - 5)) * 256) / 256).astype(np.float32)
codec = rans.HostCodec(n, 1)
t = time.time(); st, buf = codec.encode_single(1 << 32, n, x, mean, scale); t1 = time.time() - t
t = time.time(); end, msg = codec.decode_single(st, buf[::-1].copy(), n, mean[::-1].copy(), scale[::-1].copy()); t2 = time.time() - t
print("single stream 20M symbols: encode %.2f s (%.1f Msym/s), decode %.2f s (%.1f Msym/s), words %d, ok %s end %d" % (t1, n/t1/1e6, t2, n/t2/1e6, buf.size, bool(np.array_equal(msg[::-1], x)), end))
