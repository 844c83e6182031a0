"""Few-stream decode rates by kernel choice: lane (0), CTA per stream (1), cluster of 2 / 4 / 8 CTAs per
stream, on three distributions and several stream counts (development aid).

    python tools/cluster_dev.py [symbols per stream]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flic_b200 import rans, _lib


def run(name, per, lo, hi, streams, kernels):
    n = per * streams
    g = torch.Generator(device="cuda").manual_seed(3)
    mean = torch.randint(-256, 257, (n,), device="cuda", generator=g).float() / 256
    scale = torch.exp((hi - lo) * torch.rand(n, device="cuda", generator=g) + lo) / 256
    u = 10 * torch.rand(n, device="cuda", generator=g, dtype=torch.float64) - 5
    x = (torch.round((mean.double() + scale.double() * u) * 256) / 256).float()
    off = torch.arange(streams + 1, device="cuda", dtype=torch.int64) * per
    enc = rans.encode_streams(x, mean, scale, off)
    out = torch.empty(n, device="cuda")

    def timed(fn):
        fn(); torch.cuda.synchronize()
        best = 1e30
        for _ in range(2):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best
    row = []
    for k in kernels:
        old = _lib.lib().flic_set_decode_kernel(k)
        try:
            td = timed(lambda: rans.decode_streams(enc, mean, scale, off, out=out))
            xr, end, st = rans.decode_streams(enc, mean, scale, off)
            ok = torch.equal(xr, x) and not st.any().item()
            kn = _lib.lib().flic_last_coder_kernel(1).decode()
        finally:
            _lib.lib().flic_set_decode_kernel(old)
        row.append(f"{k}:{n / td / 1e3:8.2f}{'' if ok else ' BAD'}")
        last = kn
    print(f"{name:28s} {streams:5d} x {per:8d}  Msym/s by kernel  " + "  ".join(row) + f"   (auto -> {last})", flush=True)


print('cluster capacity (streams at once):', {c: _lib.lib().flic_decode_cluster_capacity(c) for c in (2, 4, 8)}, flush=True)
per = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
for streams in (1, 3, 16, 48):
    p = per if streams <= 3 else max(per // 16, 4096)
    ks = (0, 1, 2, 4, 8, -1)
    run("narrow  scale=e^U(-5,0)/256", p, -5.0, 0.0, streams, ks)
    run("test.py scale=e^U(-5,5)/256", p, -5.0, 5.0, streams, ks)
    run("wide    scale=e^U(3,5)/256", p, 3.0, 5.0, streams, ks)
