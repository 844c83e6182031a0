import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flic_b200 import rans, _lib
lo_, hi_ = (float(sys.argv[1]), float(sys.argv[2])) if len(sys.argv) > 2 else (-5.0, 0.0)
n = 400_000
g = torch.Generator(device="cuda").manual_seed(3)
mean = torch.randint(-256, 257, (n,), device="cuda", generator=g).float() / 256
scale = torch.exp((hi_ - lo_) * torch.rand(n, device="cuda", generator=g) + lo_) / 256
u = 10 * torch.rand(n, device="cuda", generator=g, dtype=torch.float64) - 5
x = (torch.round((mean.double() + scale.double() * u) * 256) / 256).float()
off = torch.tensor([0, n], device="cuda", dtype=torch.int64)
enc = rans.encode_streams(x, mean, scale, off)
if len(sys.argv) > 3:
    _lib.lib().flic_set_decode_kernel(int(sys.argv[3]))   # 0 lane, 1 CTA per stream, 2 / 4 / 8 cluster per stream
for _ in range(2):
    xr, end, st = rans.decode_streams(enc, mean, scale, off)
torch.cuda.synchronize()
print(torch.equal(xr, x))
