import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from flic_b200 import rans, _lib
n = 1572864
xs, ms, ss = bench.synth_numpy(n, 7)
xl, ml, sl = xs.tolist(), ms.tolist(), ss.tolist()
rans.encode(1 << 32, 1000, xl[:1000], ml[:1000], sl[:1000])
for rep in range(3):
    t1 = time.perf_counter()
    x, mean, scale = rans._f32(xl, n), rans._f32(ml, n), rans._f32(sl, n)
    t2 = time.perf_counter()
    c = rans._codec_for(n)
    t3 = time.perf_counter()
    st, buf = c.encode_single(1 << 32, n, x, mean, scale)
    t4 = time.perf_counter()
    b = buf.tolist()
    t5 = time.perf_counter()
    print(f"rep {rep}: conv {t2-t1:.3f} codec {t3-t2:.3f} encode_single {t4-t3:.3f} tolist {t5-t4:.3f} kernel {_lib.lib().flic_last_coder_kernel(0).decode()}")
t = time.perf_counter(); s2, b2 = rans.encode(1 << 32, n, xl, ml, sl); print("whole", time.perf_counter() - t)
t = time.perf_counter(); end, msg = rans.decode(s2, b2[::-1], n, ml[::-1], sl[::-1]); print("decode whole", time.perf_counter() - t)
