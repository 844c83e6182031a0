"""The HBM-bound kernels around the coder on ImageNet64-shaped batches: K1 (cdf_tables), K5 (coupling
add/round), N1 (permute, squeeze), a16 (u8 <-> grid), N4 (log_prob sums).  Two uses:

    python tools/prof_flowops.py                 # CUDA-event timings, achieved GB/s against MEASURED_PEAKS.json
    ncu --set full --clock-control none --import-source on -k regex:"cdf_tables|couple_add|permute_ch|squeeze|u8_to_grid|grid_to_u8|dlogistic" \
        -s 7 -c 7 -o gpurun_out/prof_flowops python tools/prof_flowops.py --once

Shapes: 32 768 images of 3x64x64 (402.7 M symbols); the flow ops see a level-0 activation
(B, 12, 32, 32) with the reference's 0.75 split (9 + 3 channels).
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flic_b200 import rans, couplelib, invertible, extenddim, flows, distlib

once = "--once" in sys.argv
B = 32768
n = B * 12288
g = torch.Generator(device="cuda").manual_seed(1)
mean = torch.randint(-256, 257, (n,), device="cuda", generator=g).float() / 256
scale = torch.exp(10 * torch.rand(n, device="cuda", generator=g) - 5) / 256
x = (torch.round((mean.double() + scale.double() * (10 * torch.rand(n, device="cuda", generator=g).double() - 5)) * 256) / 256).float()
img = torch.randint(0, 256, (B, 3, 64, 64), device="cuda", dtype=torch.uint8, generator=g)
act = torch.round(torch.randn(B, 12, 32, 32, device="cuda", generator=g) * 64) / 256
t = torch.randn(B, 3, 32, 32, device="cuda", generator=g)
perm = torch.randperm(12, device="cuda", generator=g).to(torch.int32)
grid = flows.u8_to_grid(img)
logscale = torch.log(scale).view(B, -1)
dist = distlib.DLogistic()

# (name, callable, algorithmic bytes per call)
ops = [
    ("cdf_tables_kernel (K1)", lambda: rans.cdf_tables(x, mean, scale), n * 20),
    ("couple_add_round_kernel (K5)", lambda: couplelib.couple_add_round(act, t, 9, +1, 8), t.numel() * 12),
    ("permute_channels_kernel (N1)", lambda: invertible.permute_channels(act, perm), act.numel() * 8),
    ("squeeze_kernel (N1)", lambda: extenddim.squeeze(grid, 2, +1), grid.numel() * 8),
    ("u8_to_grid_kernel (a16)", lambda: flows.u8_to_grid(img), img.numel() * 5),
    ("grid_to_u8_kernel (a16)", lambda: flows.grid_to_u8(grid), img.numel() * 5),
    ("dlogistic_log_prob_kernel, per-image sums (N4)", lambda: dist.log_prob_sums(x.view(B, -1), mean.view(B, -1), logscale), n * 12),
]
if once:
    for name, fn, _ in ops:
        fn()
    torch.cuda.synchronize()
    print("ok")
    sys.exit(0)

try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
res = []
for name, fn, nbytes in ops:
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    reps = 10
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    res.append({"kernel": name, "ms": round(ms, 4), "algorithmic_GB": round(nbytes / 1e9, 3),
                "achieved_GBps": round(nbytes / ms / 1e6, 1), "frac_of_hbm_peak": round(nbytes / ms / 1e6 / peak, 3)})
    print(json.dumps(res[-1]), flush=True)
print(json.dumps({"peak_GBps": peak, "images": B, "what": "CUDA events around 10 back-to-back calls through the Python API (allocation of the outputs included); inputs exceed the L2"}))
