"""K1 (cdf_tables) and K5 (coupling add/round) on ImageNet64-shaped data, for ncu captures (development aid).

    ncu --set full --clock-control none --import-source on -k regex:"cdf_tables|couple_add" -s 2 -c 2 -o gpurun_out/prof_k1 python tools/prof_k1.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flic_b200 import rans, couplelib
n = 32768 * 12288
g = torch.Generator(device="cuda").manual_seed(1)
mean = torch.randint(-256, 257, (n,), device="cuda", generator=g).float() / 256
scale = torch.exp(10 * torch.rand(n, device="cuda", generator=g) - 5) / 256
x = (torch.round((mean.double() + scale.double() * (10 * torch.rand(n, device="cuda", generator=g).double() - 5)) * 256) / 256).float()
xa = torch.round(torch.randn(4096, 12, 32, 32, device="cuda") * 64) / 256
t = torch.randn(4096, 3, 32, 32, device="cuda")
for _ in range(2):
    start, freq, status = rans.cdf_tables(x, mean, scale)
    couplelib.couple_add_round(xa, t, 9, +1, 8)
torch.cuda.synchronize()
print("ok", int(status.item()), n, xa.numel())
