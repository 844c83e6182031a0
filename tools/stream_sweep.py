"""Device-timed encode / decode rates as a function of the stream partition (development aid;
bench.py's `stream_count_sweep` object is the contract and calls sweep() below).

    python tools/stream_sweep.py [--json]

Partitions: the reference-native one (3 streams: one per level
over a batch of 256 ImageNet64-shaped images, trainer.py:308-315), configs[0] (16 images x 3
levels of config1 @32x32), configs[1] (256 images x 3 levels), configs[2] (9 936 patch streams of
192 symbols), and the large lane-per-stream partitions of the sweep workload.
Symbols follow rans/test.py:8-10.
"""
from __future__ import annotations

import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def partitions():
    lv = (6144, 3072, 3072)
    return [
        ("native: 3 streams (levels of 256 images)", [256 * s for s in lv]),
        ("configs[0]: 48 streams (16 images x 3 levels @32x32)", [s for s in (1536, 768, 768) for _ in range(16)]),
        ("configs[1]: 768 streams (256 images x 3 levels)", [s for s in lv for _ in range(256)]),
        ("configs[2]: 9936 streams x 192 symbols", [192] * 9936),
        ("98304 streams (32768 images x 3 levels)", [s for s in lv for _ in range(32768)]),
        ("393216 streams (131072 images x 3 levels)", [s for s in lv for _ in range(131072)]),
    ]


def sweep(dev=None, reps: int = 3, max_symbols: int | None = None, verbose: bool = False):
    import numpy as np
    import torch

    from flic_b200 import _lib, rans
    dev = dev or torch.device("cuda", torch.cuda.current_device())
    out = []
    for name, lens in partitions():
        n = int(sum(lens))
        if max_symbols is not None and n > max_symbols:
            continue
        g = torch.Generator(device=dev).manual_seed(7)
        mean = torch.randint(-256, 257, (n,), device=dev, generator=g).float() / 256
        scale = torch.exp(10 * torch.rand(n, device=dev, generator=g) - 5) / 256
        u = 10 * torch.rand(n, device=dev, generator=g, dtype=torch.float64) - 5
        x = (torch.round((mean.double() + scale.double() * u) * 256) / 256).float()
        del u
        off = torch.from_numpy(np.concatenate([[0], np.cumsum(np.asarray(lens, dtype=np.int64))])).to(dev)
        ws = rans.Workspace()
        xo = torch.empty(n, dtype=torch.float32, device=dev)
        enc = rans.encode_streams(x, mean, scale, off, workspace=ws, own_output=False, validate=False)
        xr, end, st = rans.decode_streams(enc, mean, scale, off, out=xo, validate=False)
        ok = bool(torch.equal(xr, x)) and not bool(st.any().item()) and bool((end == (1 << 32)).all().item())
        bits = enc.bits() / n

        def timed(fn):
            fn()
            torch.cuda.synchronize(dev)
            best = 1e30
            for _ in range(reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record()
                torch.cuda.synchronize(dev)
                best = min(best, a.elapsed_time(b))
            return best
        te = timed(lambda: rans.encode_streams(x, mean, scale, off, workspace=ws, own_output=False, validate=False))
        ek = _lib.lib().flic_last_coder_kernel(0).decode()
        td = timed(lambda: rans.decode_streams(enc, mean, scale, off, out=xo, validate=False))
        dk = _lib.lib().flic_last_coder_kernel(1).decode()
        rec = {"partition": name, "streams": len(lens), "symbols": n, "round_trip_exact": ok,
               "bits_per_symbol": round(bits, 5),
               "encode_ms": round(te, 4), "decode_ms": round(td, 4),
               "encode_Msym_per_s": round(n / te / 1e3, 2), "decode_Msym_per_s": round(n / td / 1e3, 2),
               "encode_kernel": ek, "decode_kernel": dk}
        out.append(rec)
        if verbose:
            print(json.dumps(rec), flush=True)
        del x, mean, scale, xo, enc, xr, ws
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    res = sweep(verbose="--json" not in sys.argv)
    if "--json" in sys.argv:
        print(json.dumps(res))
