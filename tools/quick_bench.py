"""Quick device-resident timing of the coder kernels (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flic_b200 import rans, _lib

def synth(n, seed=1):
    g = torch.Generator(device="cuda").manual_seed(seed)
    mean = torch.randint(-256, 257, (n,), device="cuda", generator=g).float() / 256
    scale = torch.exp(10 * torch.rand(n, device="cuda", generator=g) - 5) / 256
    x = torch.round((mean.double() + scale.double() * (10 * torch.rand(n, device="cuda", generator=g).double() - 5)) * 256) / 256
    return x.float(), mean, scale

def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)

imgs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
per = 12288
n = imgs * per
x, mean, scale = synth(n)
def level_major(imgs, segs):
    return torch.cat([torch.full((imgs,), sg, device="cuda") for sg in segs])
cases = (("level-major 6144/3072/3072", level_major(imgs, [6144, 3072, 3072])),
         ("equal 3072 (4/image)", torch.full((n // 3072,), 3072, device="cuda")),
         ("per image (12288)", torch.full((imgs,), 12288, device="cuda")),
         ("768-symbol streams", torch.full((n // 768,), 768, device="cuda")))
for name, segs in cases:
    off = torch.cat([torch.zeros(1, dtype=torch.int64, device="cuda"), torch.cumsum(segs, 0)])
    ws = rans.Workspace()
    enc = rans.encode_streams(x, mean, scale, off, workspace=ws)
    te = timeit(lambda: rans.encode_streams(x, mean, scale, off, workspace=ws, own_output=False))
    out = torch.empty(n, device="cuda")
    td = timeit(lambda: rans.decode_streams(enc, mean, scale, off, out=out))
    xr, end, st = rans.decode_streams(enc, mean, scale, off)
    ok = torch.equal(xr, x) and not st.any().item()
    bps = enc.bits() / n
    print(f"{name}: streams={off.numel()-1} enc {te[0]:.3f} ms ({n/te[0]/1e6:.1f} Gsym/s, {n*(12+bps/8)/te[0]/1e6:.0f} GB/s alg) "
          f"dec {td[0]:.3f} ms ({n/td[0]/1e6:.1f} Gsym/s) bits/sym {bps:.4f} ok={ok}")
start = torch.empty(n, dtype=torch.int32, device="cuda")
tt = timeit(lambda: rans.cdf_tables(x, mean, scale))
print(f"cdf_tables: {tt[0]:.3f} ms ({n/tt[0]/1e6:.1f} Gsym/s, {n*20/tt[0]/1e6:.0f} GB/s alg)")
