"""The reference's two self-tests, restated (not copied) so that they can run where the reference
is absent, driven through the reference's own import lines against compat/rans:

    rans/test.py:1-36     `from rans import encode, decode`; n random symbols (mean in [-1, 1] on the
                          1/256 grid, scale = exp(U(-5,5))/256, msg within 5 scales of the mean);
                          encode from 1<<32; decode with buffer / mean / scale reversed; print the
                          final state and count |msg - rec| > 1e-6
    coder.py:15,41-73     `from rans.rans import encode, decode`; 500 k symbols with mean in
                          [-32, 32]/256 and scale ~ 1, same round trip

Prints one JSON line.  Run with PYTHONPATH=<repo>/compat (tests/test_gpu_flow.py does)."""
import hashlib
import json
import math
import random
import struct
import sys
import time

n_test = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
n_coder = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000


def round_trip(encode, decode, mean, scale, msg):
    n = len(msg)
    t0 = time.time()
    state, buf = encode(1 << 32, n, msg, mean, scale)
    t1 = time.time()
    end, rec = decode(state, buf[::-1], n, mean[::-1], scale[::-1])
    t2 = time.time()
    rec = rec[::-1]
    errors = sum(1 for a, b in zip(msg, rec) if abs(a - b) > 1e-6)
    return {"n": n, "words": len(buf), "sha256": hashlib.sha256(struct.pack("<%dI" % len(buf), *buf)).hexdigest(), "final_state": end, "errors": errors, "bits_per_symbol": (64 + 32 * len(buf)) / n,
            "encode_s": round(t1 - t0, 4), "decode_s": round(t2 - t1, 4), "state": state,
            "list_types": [type(buf).__name__, type(rec).__name__, type(buf[0]).__name__ if buf else "int",
                           type(rec[0]).__name__]}


out = {}
random.seed(0)
from rans import decode, encode  # noqa: E402  (rans/test.py:1)
mean = [random.randint(-256, 256) / 256. for _ in range(n_test)]
scale = [math.exp(10 * random.random() - 5) / 256. for _ in range(n_test)]
msg = [round((mean[i] + scale[i] * (10 * random.random() - 5)) * 256) / 256. for i in range(n_test)]
out["rans_test_py"] = round_trip(encode, decode, mean, scale, msg)

from rans.rans import decode as decode2, encode as encode2  # noqa: E402  (coder.py:15, trainer.py:32)
assert encode2 is encode and decode2 is decode
random.seed(1)
mean = [random.randint(-32, 32) / 256. for _ in range(n_coder)]
scale = [math.exp(random.random() * 0.01 - 0.005) for _ in range(n_coder)]
msg = [round((mean[i] + scale[i] * (1. * random.random() - .5)) * 256) / 256. for i in range(n_coder)]
out["coder_py_main"] = round_trip(encode2, decode2, mean, scale, msg)
print(json.dumps(out))
