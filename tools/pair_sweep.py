"""Two lanes per stream in the lane decoder (FLIC_DEC_PAIR_MAX_WARPS, warps per SM up to which it is used; 0 = never)
against one lane per stream: decode time by stream count, every run checked against the input."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import torch
    from flic_b200 import rans, _lib
    for name, lo, hi in (("test.py", -5.0, 5.0), ("image", 1.0, 3.5)):
        for streams, per in ((444, 4096), (768, 4096), (1536, 4096), (3072, 4096), (4736, 4096), (9472, 4096), (9936, 192),
                             (14208, 4096), (18944, 4096), (37888, 2048)):
            n = streams * per
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
                best = 1e9
                for _ in range(5):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); fn(); b.record(); torch.cuda.synchronize()
                    best = min(best, a.elapsed_time(b))
                return best
            td = timed(lambda: rans.decode_streams(enc, mean, scale, off, out=out))
            xr, end, st = rans.decode_streams(enc, mean, scale, off)
            ok = bool(torch.equal(xr, x)) and not bool(st.any()) and bool((end == (1 << 32)).all())
            print(f"{name:8s} {streams:6d} x {per:5d}: {n / td / 1e3:9.1f} Msym/s  {td:7.3f} ms  ok={ok}  "
                  f"{_lib.lib().flic_last_coder_kernel(1).decode()}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
    else:
        for v in ("0", "100000"):
            print("== FLIC_DEC_PAIR_MAX_WARPS =", v, flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env={**os.environ, "FLIC_DEC_PAIR_MAX_WARPS": v})
