"""The HBM-bound kernels around the coder on ImageNet64-shaped batches: K1 (cdf_tables), K5 (coupling
add/round), N1 (permute, squeeze), a16 (u8 <-> grid), N4 (log_prob sums).  Three uses:

    python tools/prof_flowops.py                 # CUDA-event timings, achieved GB/s against MEASURED_PEAKS.json
    ncu --set full --clock-control none --import-source on -k regex:"cdf_tables|couple_add|permute_ch|squeeze|u8_to_grid|grid_to_u8|dlogistic" \
        -s 1 -c 7 -o gpurun_out/prof_flowops python tools/prof_flowops.py --once
    bench.py imports measure() for the `flow_kernels` object of its line.

Shapes: 32 768 images of 3x64x64 (402.7 M symbols); the flow ops see a level-0 activation
(B, 12, 32, 32) with the reference's 0.75 split (9 + 3 channels).
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_ops(B=32768, device="cuda"):
    """[(name, callable, algorithmic bytes per call)] on freshly generated inputs."""
    import torch
    from flic_b200 import rans, couplelib, invertible, extenddim, flows, distlib
    n = B * 12288
    g = torch.Generator(device=device).manual_seed(1)
    mean = torch.randint(-256, 257, (n,), device=device, generator=g).float() / 256
    scale = torch.exp(10 * torch.rand(n, device=device, generator=g) - 5) / 256
    x = (torch.round((mean.double() + scale.double() * (10 * torch.rand(n, device=device, generator=g).double() - 5)) * 256) / 256).float()
    img = torch.randint(0, 256, (B, 3, 64, 64), device=device, dtype=torch.uint8, generator=g)
    act = torch.round(torch.randn(B, 12, 32, 32, device=device, generator=g) * 64) / 256
    t = torch.randn(B, 3, 32, 32, device=device, generator=g)
    perm = torch.randperm(12, device=device, generator=g).to(torch.int32)
    grid = flows.u8_to_grid(img)
    logscale = torch.log(scale).view(B, -1)
    dist = distlib.DLogistic()
    return [
        ("cdf_tables_kernel (K1)", lambda: rans.cdf_tables(x, mean, scale), n * 20),
        ("couple_add_round_kernel (K5)", lambda: couplelib.couple_add_round(act, t, 9, +1, 8), t.numel() * 12),
        ("permute_channels_kernel (N1)", lambda: invertible.permute_channels(act, perm), act.numel() * 8),
        ("squeeze2_vec_kernel (N1)", lambda: extenddim.squeeze(grid, 2, +1), grid.numel() * 8),
        ("u8_to_grid_vec_kernel (a16)", lambda: flows.u8_to_grid(img), img.numel() * 5),
        ("grid_to_u8_vec_kernel (a16)", lambda: flows.grid_to_u8(grid), img.numel() * 5),
        ("dlogistic_log_prob_kernel, per-image sums (N4)", lambda: dist.log_prob_sums(x.view(B, -1), mean.view(B, -1), logscale), n * 12),
    ]


def peak_gbps():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def measure(B=32768, device="cuda", reps=10):
    """CUDA events around `reps` back-to-back calls through the Python API (output allocation included)."""
    import torch
    peak = peak_gbps()
    res = []
    for name, fn, nbytes in build_ops(B, device):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        res.append({"kernel": name, "ms": round(ms, 4), "algorithmic_GB": round(nbytes / 1e9, 3),
                    "achieved_GBps": round(nbytes / ms / 1e6, 1), "frac_of_hbm_peak": round(nbytes / ms / 1e6 / peak, 3)})
    return {"peak_GBps": peak, "images": B, "kernels": res,
            "what": "CUDA events around 10 back-to-back calls through the Python API (allocation of the outputs included); "
                    "inputs exceed the L2; algorithmic bytes as in DESIGN.md section 5"}


if __name__ == "__main__":
    import torch
    if "--once" in sys.argv:
        for name, fn, _ in build_ops():
            fn()
        torch.cuda.synchronize()
        print("ok")
    else:
        out = measure()
        for r in out["kernels"]:
            print(json.dumps(r), flush=True)
        print(json.dumps({k: v for k, v in out.items() if k != "kernels"}))
