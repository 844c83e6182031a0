"""Encode + decode of the sweep workload with ONE stream per image (12 288 contiguous symbols), for ncu
captures and quick timings (development aid).

    python tools/prof_target_chained.py [images] [--time]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flic_b200 import rans

args = [a for a in sys.argv[1:] if not a.startswith("--")]
imgs = int(args[0]) if args else 131072
n = imgs * 12288
g = torch.Generator(device="cuda").manual_seed(1)
mean = torch.randint(-256, 257, (n,), device="cuda", generator=g).float() / 256
scale = torch.exp(10 * torch.rand(n, device="cuda", generator=g) - 5) / 256
x = (torch.round((mean.double() + scale.double() * (10 * torch.rand(n, device="cuda", generator=g).double() - 5)) * 256) / 256).float()
off = torch.arange(imgs + 1, device="cuda", dtype=torch.int64) * 12288
ws = rans.Workspace()
out = torch.empty(n, device="cuda")
for _ in range(2):
    enc = rans.encode_streams(x, mean, scale, off, workspace=ws, own_output=False)
    xr, end, st = rans.decode_streams(enc, mean, scale, off, out=out)
torch.cuda.synchronize()
print("ok", torch.equal(xr, x), int(st.any()))
if "--time" in sys.argv:
    def timed(fn, reps=5):
        best = 1e9
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best
    te = timed(lambda: rans.encode_streams(x, mean, scale, off, workspace=ws, own_output=False))
    td = timed(lambda: rans.decode_streams(enc, mean, scale, off, out=out))
    print(f"images {imgs}: encode {n / te / 1e6:.1f} decode {n / td / 1e6:.1f} G symbols/s")
