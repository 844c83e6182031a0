"""Path statistics of the CTA / cluster-per-stream decoder's consumer (development aid): run with a
library built with -DFLIC_COOP_STATS=1 (tools/variants.py build stats:-DFLIC_COOP_STATS=1), e.g.

    FLIC_B200_LIB=tools/_build/var_stats/libflic_b200.so python tools/coop_stats.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flic_b200 import rans, _lib

n = 400_000
for name, lo, hi in (("narrow", -5.0, 0.0), ("test.py", -5.0, 5.0), ("wide", 3.0, 5.0)):
    g = torch.Generator(device="cuda").manual_seed(3)
    mean = torch.randint(-256, 257, (n,), device="cuda", generator=g).float() / 256
    scale = torch.exp((hi - lo) * torch.rand(n, device="cuda", generator=g) + lo) / 256
    u = 10 * torch.rand(n, device="cuda", generator=g, dtype=torch.float64) - 5
    x = (torch.round((mean.double() + scale.double() * u) * 256) / 256).float()
    off = torch.tensor([0, n], device="cuda", dtype=torch.int64)
    enc = rans.encode_streams(x, mean, scale, off)
    for kern in (1, 2, 8):
        _lib.lib().flic_set_decode_kernel(kern)
        print(name, "kernel", kern, flush=True)
        xr, end, st = rans.decode_streams(enc, mean, scale, off)
        torch.cuda.synchronize()
        assert torch.equal(xr, x)
