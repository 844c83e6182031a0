"""Per-class single-stream decode rates of the CTA-per-stream decoder (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flic_b200 import rans, _lib

def run(name, n, lo, hi):
    g = torch.Generator(device="cuda").manual_seed(3)
    mean = torch.randint(-256, 257, (n,), device="cuda", generator=g).float() / 256
    scale = torch.exp((hi - lo) * torch.rand(n, device="cuda", generator=g) + lo) / 256
    u = 10 * torch.rand(n, device="cuda", generator=g, dtype=torch.float64) - 5
    x = (torch.round((mean.double() + scale.double() * u) * 256) / 256).float()
    off = torch.tensor([0, n], device="cuda", dtype=torch.int64)
    enc = rans.encode_streams(x, mean, scale, off)
    out = torch.empty(n, device="cuda")
    def timed(fn):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b)
    td = timed(lambda: rans.decode_streams(enc, mean, scale, off, out=out))
    xr, end, st = rans.decode_streams(enc, mean, scale, off)
    ok = torch.equal(xr, x) and not st.any().item()
    print(f"{name}: {n / td / 1e3:.2f} Msym/s = {td * 1e-3 * 1.965e9 / n:.0f} cycles/symbol  {_lib.lib().flic_last_coder_kernel(1).decode()} ok={ok}", flush=True)

n = 1_000_000
import math
run("single  c in (0.007, 3)", n, -5.0, math.log(3.0))
run("single, some misses c in (3.2, 5)", n, math.log(3.2), math.log(5.0))
run("multi n=2..4  c in (5.5, 10)", n, math.log(5.5), math.log(10.0))
run("multi n=8..16 c in (24, 47)", n, math.log(24.0), math.log(47.0))
run("none  c in (50, 148)", n, math.log(50.0), 5.0)
