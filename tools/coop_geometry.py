"""CTA-per-stream decoder built with other geometries (tools/variants.py build name:-DFLIC_COOP1_WARPS=..,-DFLIC_COOP1_SLOTS=..):
decode rate against stream count, beside the lane kernel, on two distributions; every run checked against the input."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import torch
    from flic_b200 import rans, _lib
    for name, lo, hi in (("test.py", -5.0, 5.0), ("narrow", -5.0, 1.0), ("image", 1.0, 3.5)):
        for streams in (296, 444, 592, 768, 888, 1536, 3072):
            per = 4096
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
            xr, end, st = rans.decode_streams(enc, mean, scale, off)
            ok = bool(torch.equal(xr, x)) and not bool(st.any())
            print(f"{name:8s} {streams:5d} streams: {n / td / 1e3:9.1f} Msym/s  {td:7.3f} ms  ok={ok}  "
                  f"{_lib.lib().flic_last_coder_kernel(1).decode()}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
    else:
        pkg = os.path.join(ROOT, "finalproject-losslessimagecompression_b200")
        out = os.path.join(ROOT, "tools", "_build")
        libs = [("lane kernel", os.path.join(pkg, "libflic_b200.so"), "0"), ("product geometry", os.path.join(pkg, "libflic_b200.so"), "100000")]
        for d in sorted(os.listdir(out)) if os.path.isdir(out) else []:
            p = os.path.join(out, d, "libflic_b200.so")
            if d.startswith("var_") and os.path.exists(p):
                libs.append((d[4:], p, "100000"))
        for name, lib, cap in libs:
            print(f"== {name} (FLIC_DEC_COOP_MAX_STREAMS={cap})", flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__), "one"],
                           env={**os.environ, "FLIC_B200_LIB": lib, "FLIC_DEC_COOP_MAX_STREAMS": cap})
