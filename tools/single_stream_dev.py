"""Device-resident single-stream decode rates on three distributions (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flic_b200 import rans, _lib

def run(name, n, lo, hi, streams=1):
    g = torch.Generator(device="cuda").manual_seed(3)
    mean = torch.randint(-256, 257, (n,), device="cuda", generator=g).float() / 256
    scale = torch.exp((hi - lo) * torch.rand(n, device="cuda", generator=g) + lo) / 256
    u = 10 * torch.rand(n, device="cuda", generator=g, dtype=torch.float64) - 5
    x = (torch.round((mean.double() + scale.double() * u) * 256) / 256).float()
    off = torch.arange(streams + 1, device="cuda", dtype=torch.int64) * (n // streams)
    enc = rans.encode_streams(x, mean, scale, off)
    out = torch.empty(n, device="cuda")
    def timed(fn):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b)
    td = timed(lambda: rans.decode_streams(enc, mean, scale, off, out=out))
    te = timed(lambda: rans.encode_streams(x, mean, scale, off))
    xr, end, st = rans.decode_streams(enc, mean, scale, off)
    ok = torch.equal(xr, x) and not st.any().item()
    print(f"{name}: {streams} stream(s) x {n // streams}: decode {n / td / 1e3:.2f} Msym/s ({td:.2f} ms) encode {n / te / 1e3:.2f} Msym/s  "
          f"kernel {_lib.lib().flic_last_coder_kernel(1).decode()} ok={ok} bits/sym {enc.bits() / n:.3f}", flush=True)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
run("narrow  scale=e^U(-5,0)/256", n, -5.0, 0.0)
run("test.py scale=e^U(-5,5)/256", n, -5.0, 5.0)
run("wide    scale=e^U(3,5)/256", n, 3.0, 5.0)
run("test.py, 3 streams", 3 * (n // 2), -5.0, 5.0, 3)
run("test.py, 48 streams", 48 * 65536, -5.0, 5.0, 48)
run("test.py, 768 streams", 768 * 4096, -5.0, 5.0, 768)
