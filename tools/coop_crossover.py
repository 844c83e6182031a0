"""Where the CTA-per-stream decoder stops paying: stream-count sweep, both kernels, two distributions."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def one():
    import torch
    from flic_b200 import rans, _lib
    for name, lo, hi in (("test.py", -5.0, 5.0), ("narrow", -5.0, 1.0)):
        for streams in (24, 48, 96, 148, 296, 444, 592, 768, 1536):
            per = 8192
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
                for _ in range(3):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); fn(); b.record(); torch.cuda.synchronize()
                    best = min(best, a.elapsed_time(b))
                return best
            td = timed(lambda: rans.decode_streams(enc, mean, scale, off, out=out))
            print(f"{name:8s} {streams:5d} streams: {n / td / 1e3:9.1f} Msym/s  {td:7.3f} ms  {_lib.lib().flic_last_coder_kernel(1).decode()}", flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        one()
    else:
        for v in ("100000", "0"):
            print("FLIC_DEC_COOP_MAX_STREAMS =", v, flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env={**os.environ, "FLIC_DEC_COOP_MAX_STREAMS": v})
