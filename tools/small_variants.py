"""Variant builds (tools/variants.py build ...) timed where the lane decoder is latency-bound: one warp per SM or less."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import torch
    from flic_b200 import rans, _lib
    _lib.lib().flic_set_decode_kernel(0)
    for streams, per in ((1, 400000), (3, 400000), (48, 16384), (768, 4096), (4736, 4096), (9936, 192)):
        n = streams * per
        g = torch.Generator(device="cuda").manual_seed(3)
        mean = torch.randint(-256, 257, (n,), device="cuda", generator=g).float() / 256
        scale = torch.exp(10 * torch.rand(n, device="cuda", generator=g) - 5) / 256
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
        te = timed(lambda: rans.encode_streams(x, mean, scale, off))
        xr, end, st = rans.decode_streams(enc, mean, scale, off)
        ok = bool(torch.equal(xr, x)) and not bool(st.any())
        print(f"  {streams:6d} x {per:5d}: decode {td:7.3f} ms ({td * 1e-3 * 1.9e9 / per:6.0f} cycles/symbol at 1.9 GHz)  encode {te:7.3f} ms  ok={ok}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
    else:
        pkg = os.path.join(ROOT, "finalproject-losslessimagecompression_b200")
        out = os.path.join(ROOT, "tools", "_build")
        libs = [("product", os.path.join(pkg, "libflic_b200.so"))]
        for d in sorted(os.listdir(out)) if os.path.isdir(out) else []:
            p = os.path.join(out, d, "libflic_b200.so")
            if d.startswith("var_") and os.path.exists(p):
                libs.append((d[4:], p))
        for name, lib in libs:
            print("==", name, flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env={**os.environ, "FLIC_B200_LIB": lib})
