"""N3 evidence: chunks of compress / decompress on rotating CUDA streams (pipeline=k) -- step times for k = 1, 2, 4 and,
from a torch.profiler (kineto) trace of one step, how much of the coder kernels' time runs while another stream
has a kernel in flight.  imagenet64.yaml model, 256 images, codec_batch 64 (four chunks).

    python tools/pipeline_overlap.py > profiles/<tag>_pipeline_overlap.json
"""
import json, os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from flic_b200 import flows

layer = dict(name="DenseLayer", act="ReLU")
block = dict(name="DenseBlock", growth_channel=512, depth=12, layer=layer)
cfg = dict(name="IDFlows", nflows=8, nbits=8, nsplit=3, H=64, W=64, C=3,
           couple=dict(name="AdditiveCouple", split=0.75, nn=block, round=dict(name="Round", nbits=8)),
           extenddim=dict(name="ExtendDim", scale=2), prior=dict(name="Prior", round=dict(name="Round", nbits=8), nn=block),
           distribution=dict(name="DLogistic"), round=dict(name="Round", nbits=8))


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def overlap(events):
    """events: (start, end, stream, name) of every kernel.  Returns per-stream busy time, the time during which at
    least two streams are busy, and the coder kernels' total and overlapped time (all in ms)."""
    per_stream = {}
    for s, e, st, _ in events:
        per_stream.setdefault(st, []).append((s, e))
    edges = []
    for st, iv in per_stream.items():
        iv.sort()
        merged = []
        for s, e in iv:
            if merged and s <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], e)
            else:
                merged.append([s, e])
        per_stream[st] = merged
        for s, e in merged:
            edges += [(s, 1), (e, -1)]
    edges.sort()
    busy2 = busy1 = 0.0
    depth, last = 0, None
    for t, d in edges:
        if last is not None:
            if depth >= 2: busy2 += t - last
            if depth >= 1: busy1 += t - last
        depth += d; last = t
    coder_total = coder_over = 0.0
    for s, e, st, name in events:
        if "flic::rans_" not in name:
            continue
        coder_total += e - s
        for other, iv in per_stream.items():
            if other == st:
                continue
            for a, b in iv:
                lo, hi = max(a, s), min(b, e)
                if hi > lo:
                    coder_over += hi - lo
    return ({str(k): round(sum(e - s for s, e in v) / 1e3, 3) for k, v in per_stream.items()}, round(busy1 / 1e3, 3),
            round(busy2 / 1e3, 3), round(coder_total / 1e3, 3), round(min(coder_over, coder_total) / 1e3, 3))


def main():
    torch.manual_seed(0); random.seed(0)
    model = flows.build_model(cfg); flows.perturb_heads(model, 0.02); model = model.cuda().eval()
    img = torch.randint(0, 256, (256, 3, 64, 64), dtype=torch.uint8, generator=torch.Generator().manual_seed(1234)).cuda()
    out = {"model": "imagenet64.yaml (growth 512, depth 12), 256 images, codec_batch 64 = 4 chunks", "steps": {}}
    blobs = {}
    for k in (1, 2, 4):
        cb = model.compress(img, codec_batch=64, check=False, pipeline=k)
        blobs[k] = cb.to_bytes()
        assert torch.equal(model.decompress(cb, check=False, pipeline=k), img)
        tc = timed(lambda: model.compress(img, codec_batch=64, check=False, pipeline=k))
        td = timed(lambda: model.decompress(cb, check=False, pipeline=k))
        out["steps"][f"pipeline={k}"] = {"compress_ms": round(tc, 2), "decompress_ms": round(td, 2)}
    out["container_bytes_identical_for_every_pipeline"] = blobs[1] == blobs[2] == blobs[4]
    for k in (1, 2):
        cb = model.compress(img, codec_batch=64, check=False, pipeline=k)
        for what, fn in (("compress", lambda: model.compress(img, codec_batch=64, check=False, pipeline=k)),
                         ("decompress", lambda: model.decompress(cb, check=False, pipeline=k))):
            fn(); torch.cuda.synchronize()
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                fn(); torch.cuda.synchronize()
            ev = []
            for e in prof.events():
                if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start:
                    ev.append((e.time_range.start, e.time_range.end, getattr(e, "device_resource_id", getattr(e, "device_index", 0)), e.name))
            streams, busy1, busy2, ct, co = overlap(ev)
            out.setdefault("trace", {})[f"{what}, pipeline={k}"] = {
                "kernels": len(ev), "kernel_busy_ms_per_stream": streams, "ms_with_at_least_one_stream_busy": busy1,
                "ms_with_two_or_more_streams_busy": busy2, "coder_kernel_ms": ct,
                "coder_kernel_ms_while_another_stream_has_a_kernel_running": co}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
