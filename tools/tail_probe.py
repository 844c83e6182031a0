"""How much of a lane-kernel launch is its last, partly filled wave: time against stream count.

Equal streams of 3072 symbols; the encoder holds 4 CTAs x 128 streams per SM (75 776 streams per
wave on 148 SMs), the decoder 7 (132 608).  A launch whose time grows in steps of whole waves pays
for its tail; one whose time is linear in the stream count does not.
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flic_b200 import rans


def synth(n, seed=1):
    g = torch.Generator(device="cuda").manual_seed(seed)
    mean = torch.randint(-256, 257, (n,), device="cuda", generator=g).float() / 256
    scale = torch.exp(10 * torch.rand(n, device="cuda", generator=g) - 5) / 256
    u = 10 * torch.rand(n, device="cuda", generator=g, dtype=torch.float64) - 5
    x = (torch.round((mean.double() + scale.double() * u) * 256) / 256).float()
    return x, mean, scale


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


def main():
    per = int(sys.argv[1]) if len(sys.argv) > 1 else 3072
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    wave_e, wave_d = sms * 4 * 128, sms * 7 * 128
    most = int(6.0 * wave_d)
    x, mean, scale = synth(most * per)
    ws = rans.Workspace()
    for kind, wave, ks in (("encode", wave_e, (4.0, 5.0, 5.19, 5.5, 6.0, 8.0, 10.0)),
                           ("decode", wave_d, (2.0, 2.97, 3.0, 3.25, 3.5, 4.0, 6.0))):
        for k in ks:
            streams = int(k * wave) // 128 * 128
            n = streams * per
            off = torch.arange(streams + 1, device="cuda", dtype=torch.int64) * per
            xs, ms, ss = x[:n], mean[:n], scale[:n]
            if kind == "encode":
                t = timeit(lambda: rans.encode_streams(xs, ms, ss, off, workspace=ws, own_output=False))
            else:
                enc = rans.encode_streams(xs, ms, ss, off, workspace=ws)
                out = torch.empty(n, device="cuda")
                t = timeit(lambda: rans.decode_streams(enc, ms, ss, off, out=out))
                del enc, out
            print(f"{kind} {k:5.2f} waves {streams:8d} streams: {t:7.3f} ms  {n / t / 1e6:6.1f} G symbols/s  "
                  f"{t / k:6.3f} ms per wave", flush=True)


if __name__ == "__main__":
    main()
