"""Minimal encode+decode of the sweep workload for ncu captures (development aid).

    ncu --set full --clock-control none --import-source on -k regex:rans_ -s 2 -c 2 -o gpurun_out/prof python tools/prof_target.py [images]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flic_b200 import rans

imgs = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
n = imgs * 12288
g = torch.Generator(device="cuda").manual_seed(1)
mean = torch.randint(-256, 257, (n,), device="cuda", generator=g).float() / 256
scale = torch.exp(10 * torch.rand(n, device="cuda", generator=g) - 5) / 256
x = (torch.round((mean.double() + scale.double() * (10 * torch.rand(n, device="cuda", generator=g).double() - 5)) * 256) / 256).float()
segs = torch.cat([torch.full((imgs,), sg, device="cuda") for sg in (6144, 3072, 3072)])
off = torch.cat([torch.zeros(1, dtype=torch.int64, device="cuda"), torch.cumsum(segs, 0)])
ws = rans.Workspace()
out = torch.empty(n, device="cuda")
for _ in range(2):
    enc = rans.encode_streams(x, mean, scale, off, workspace=ws)
    xr, end, st = rans.decode_streams(enc, mean, scale, off, out=out)
torch.cuda.synchronize()
print("ok", torch.equal(xr, x), int(st.any()), enc.bits() / n)
