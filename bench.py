#!/usr/bin/env python
"""Benchmark of the entropy-coding hot path (contract: see the repo prompt / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]  # the reference's CPU coder

Workload (BASELINE.json configs[4], the rANS-only sweep): ImageNet64-shaped images, 12 288
symbols each = the imagenet64.yaml latent levels 6144 / 3072 / 3072, with
precomputed logistic parameters drawn as rans/test.py:8-10.  One rANS stream per image x level
(3 per image), the partition IDFlows.compress uses.  A step is one pass of the hot path over one
chunk of `--images` images per GPU: fused table evaluation + rANS encode + stream concatenation,
then rANS decode back to symbols.  Ranks shard the image range; there is no data-path collective
(weak scaling: every rank codes its own chunk).

metric  = MB/s of raw pixels (1 symbol = 1 sub-pixel = 1 byte, MB = 1e6 B) through
          encode + decode, i.e. symbols / (t_encode + t_decode).
value   = inputs resident in HBM, CUDA-event timed.   e2e = same through the host-buffer C ABI
          (flic_codec_encode / flic_codec_decode) with pinned host arrays, copies inside the timing.
`--workload full` times BASELINE.json configs[1] instead (imagenet64.yaml model, batch 256,
IDFlows.compress + decompress, convolutions in PyTorch fp32).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEGMENTS = (6144, 3072, 3072)      # imagenet64.yaml latents (6,32,32) (12,16,16) (48,8,8)
PER_IMAGE = sum(SEGMENTS)
METRIC = "encode+decode throughput, raw pixels"
UNIT = "MB/s"


# ------------------------------------------------------------------------------------------------
# synthetic inputs (rans/test.py:8-10 distribution), numpy for the CPU arms, torch for the GPU arm
# ------------------------------------------------------------------------------------------------

def synth_numpy(n: int, seed: int):
    import numpy as np
    rng = np.random.default_rng(seed)
    mean = (rng.integers(-256, 257, n) / 256).astype(np.float32)
    scale = (np.exp(10 * rng.random(n) - 5) / 256).astype(np.float32)
    x = np.round((mean.astype(np.float64) + scale.astype(np.float64) * (10 * rng.random(n) - 5)) * 256) / 256
    return x.astype(np.float32), mean, scale


def synth_torch(n: int, seed: int, device):
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    mean = torch.randint(-256, 257, (n,), device=device, generator=g).float() / 256
    scale = torch.exp(10 * torch.rand(n, device=device, generator=g) - 5) / 256
    u = 10 * torch.rand(n, device=device, generator=g, dtype=torch.float64) - 5
    x = (torch.round((mean.double() + scale.double() * u) * 256) / 256).float()
    return x, mean, scale


def level_major_offsets(n_images: int):
    """Streams in level-major order (all images' level 0, then level 1, then level 2), the order in
    which the reference codes a batch (one call per level, trainer.py:308); image-major inside."""
    import numpy as np
    off = [0]
    for seg in SEGMENTS:
        off.extend((off[-1] + seg * (np.arange(n_images) + 1)).tolist())
    return np.asarray(off, dtype=np.int64)


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------

class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 3.0:   # nvidia-smi takes ~0.2 s to print its first row
                time.sleep(0.01)
            self.rows.clear()                                  # keep only samples taken inside the timed region
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# the reference arm: the reference's own Cython coder (oracle/_ref) on the host cores
# ------------------------------------------------------------------------------------------------

def _ref_worker(args):
    """One process = one shard of images, coded exactly like trainer.py:309-323: per level a fresh
    state, .tolist() inputs, encode, decode with reversed inputs, rebuild and compare."""
    seed, n_images, use_ref = args
    import numpy as np
    from oracle import pyoracle
    ref = pyoracle.ref_rans() if use_ref else None
    n = n_images * PER_IMAGE
    x, mean, scale = synth_numpy(n, seed)
    off = level_major_offsets(n_images)
    t_enc = t_dec = 0.0
    words = 0
    errors = 0
    for s in range(off.size - 1):
        a, b = int(off[s]), int(off[s + 1])
        if ref is not None:
            t1 = time.perf_counter()
            xi, mi, si = x[a:b].tolist(), mean[a:b].tolist(), scale[a:b].tolist()
            state, buf = ref.encode(1 << 32, b - a, xi, mi, si)
            t3 = time.perf_counter()
            end, msg = ref.decode(state, buf[::-1], b - a, mi[::-1], si[::-1])
            rec = np.asarray(msg[::-1], dtype=np.float32)
            t4 = time.perf_counter()
        else:
            t1 = time.perf_counter()
            state, buf = pyoracle.encode(1 << 32, b - a, x[a:b], mean[a:b], scale[a:b])
            t3 = time.perf_counter()
            end, msg = pyoracle.decode(state, buf[::-1], b - a, mean[a:b][::-1], scale[a:b][::-1])
            rec = msg[::-1]
            t4 = time.perf_counter()
        t_enc += t3 - t1
        t_dec += t4 - t3
        words += len(buf)
        errors += int((rec != x[a:b]).sum()) + int(end != 1 << 32)
    return t_enc, t_dec, words, errors


def run_reference(args) -> dict:
    """Times the reference's CPU implementation of the path.  kind = "reference" when the
    re-cythonised rans.pyx (oracle/_ref) is present, else "port" (oracle/rans_oracle.c)."""
    import multiprocessing as mp
    from oracle import pyoracle
    use_ref = pyoracle.ref_rans() is not None
    cores = args.cpu_procs or max(1, len(os.sched_getaffinity(0)))
    # per process and step: ~2 s of single-core work at the surveyed ~0.45 MB/s (reference) or ~1.3 MB/s (port)
    imgs_per_proc = args.cpu_images_per_proc or (64 if use_ref else 160)
    ctx = mp.get_context("spawn")
    times = []
    words = errors = 0
    with ctx.Pool(cores) as pool:
        for step in range(args.warmup + args.steps):
            jobs = [(1000 * step + p, imgs_per_proc, use_ref) for p in range(cores)]
            t0 = time.perf_counter()
            res = pool.map(_ref_worker, jobs)
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
                words += sum(r[2] for r in res)
                errors += sum(r[3] for r in res)
                enc_cpu = sum(r[0] for r in res)
                dec_cpu = sum(r[1] for r in res)
    n_sym = cores * imgs_per_proc * PER_IMAGE
    total = sum(times)
    value = n_sym * len(times) / total / 1e6
    one_core = n_sym / (enc_cpu + dec_cpu) / 1e6      # per-core rate inside the last step
    kind = "reference" if use_ref else "port"
    sample = (f"{cores} processes x {imgs_per_proc} images x {PER_IMAGE} symbols per step, each coded as "
              f"trainer.py:309-323 (tolist + encode, decode + rebuild), rans/test.py:8-10 distribution")
    return {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * total / max(len(times), 1), 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64/f32 + u64", "data": "synthetic",
        "config": workload_config(args, cores * imgs_per_proc),
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "value_per_core": round(one_core, 4), "encode_MBps_per_core": round(n_sym / enc_cpu / 1e6 * 1, 4),
                         "decode_MBps_per_core": round(n_sym / dec_cpu / 1e6 * 1, 4)},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "errors": errors, "bits_per_symbol": round((64 * cores * imgs_per_proc * 3 * len(times) + 32 * words) / (n_sym * len(times)), 5),
    }


def workload_config(args, images_per_step_per_rank):
    return {"workload": "configs[4] rANS-only sweep: ImageNet64-shaped images (12288 symbols = latent levels "
                        "6144/3072/3072), logistic params as rans/test.py:8-10, one stream per image x level",
            "images_per_step_per_gpu": int(images_per_step_per_rank), "symbols_per_image": PER_IMAGE,
            "streams_per_image": 3, "partition": "level-major, image-major inside a level",
            "l2": "inputs (12 B/symbol) exceed the 126 MB L2 many times over; no flush needed"}


# ------------------------------------------------------------------------------------------------
# the CUDA arm
# ------------------------------------------------------------------------------------------------

def cuda_events():
    import torch
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def run_flic(args) -> dict | None:
    import numpy as np
    import torch

    from flic_b200 import _lib, rans, sharding

    rank, world, local = sharding.init_process_group()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.lib()

    if args.workload == "full":
        return run_full(args, rank, world, dev)

    n_img = args.images
    n = n_img * PER_IMAGE
    x, mean, scale = synth_torch(n, 1234 + rank, dev)
    off_np = level_major_offsets(n_img)
    off = torch.from_numpy(off_np).to(dev)
    ns = off.numel() - 1
    ws = rans.Workspace()
    out = torch.empty(n, dtype=torch.float32, device=dev)

    # ---- parity gate before timing (small slice against the oracle; whole chunk round trip)
    parity = {}
    if rank == 0:
        from oracle import pyoracle
        k = 48
        sl = slice(0, k * SEGMENTS[0])
        xs, ms, ss = (t[sl].cpu().numpy() for t in (x, mean, scale))
        o = np.arange(k + 1, dtype=np.int64) * SEGMENTS[0]
        enc_s = rans.encode_streams(x[sl], mean[sl], scale[sl], torch.from_numpy(o).to(dev))
        w_o, wo_o, st_o, _ = pyoracle.encode_streams(xs, ms, ss, o)
        parity["bitstreams_equal_oracle"] = bool(
            np.array_equal(enc_s.words.cpu().numpy().view(np.uint32), w_o)
            and np.array_equal(enc_s.final_states.cpu().numpy().view(np.uint64), st_o))
        # cost of the finer partition against the reference's native one (one stream per
        # level over a batch of 256 images, trainer.py:308-315) on the same symbols
        nb = min(256, n_img)
        sel = torch.cat([torch.arange(nb * seg, device=dev) + base for seg, base in
                         zip(SEGMENTS, np.cumsum([0] + [s_ * n_img for s_ in SEGMENTS[:-1]]).tolist())])
        xb, mb, sb = x[sel].contiguous(), mean[sel].contiguous(), scale[sel].contiguous()
        native_off = torch.tensor(np.cumsum([0] + [s_ * nb for s_ in SEGMENTS]), dtype=torch.int64, device=dev)
        fine_off = torch.from_numpy(level_major_offsets(nb)).to(dev)
        nat = rans.encode_streams(xb, mb, sb, native_off)
        fin = rans.encode_streams(xb, mb, sb, fine_off)
        nsym_b = nb * PER_IMAGE
        bits_native, bits_fine = nat.bits() / nsym_b, fin.bits() / nsym_b
        parity["bits_per_symbol_256_images"] = {"native_one_stream_per_level": round(bits_native, 5),
                                                "one_stream_per_image_x_level": round(bits_fine, 5),
                                                "overhead_pct": round(100 * (bits_fine / bits_native - 1), 4)}
        del xb, mb, sb, nat, fin, sel
    enc = rans.encode_streams(x, mean, scale, off, workspace=ws)
    xr, end, st = rans.decode_streams(enc, mean, scale, off, out=out)
    parity["round_trip_exact"] = bool(torch.equal(xr, x)) and not bool(st.any().item()) and not bool(enc.status.any().item())
    parity["all_streams_end_at_1<<32"] = bool((end == (1 << 32)).all().item())
    n_words = enc.n_words()
    bits_per_symbol = (64 * ns + 32 * n_words) / n
    if not all(v for v in parity.values() if isinstance(v, bool)):
        raise SystemExit(f"bench.py: parity gate failed: {parity}")

    def step(timers=None):
        if timers is not None:
            e0, e1 = cuda_events(); d0, d1 = cuda_events()
            e0.record()
        enc_l = rans.encode_streams(x, mean, scale, off, workspace=ws, own_output=False)
        if timers is not None:
            e1.record(); d0.record()
        rans.decode_streams(enc_l, mean, scale, off, out=out)
        if timers is not None:
            d1.record()
            timers.append((e0, e1, d0, d1))

    for _ in range(args.warmup):
        step()
    launches0 = _lib.kernel_launches()
    sampler = ClockSampler(local)
    timers = []
    sharding.barrier(dev)
    torch.cuda.synchronize()
    sampler.start()
    t0, t1 = cuda_events()
    t0.record()
    for _ in range(args.steps):
        step(timers)
    t1.record()
    torch.cuda.synchronize()
    sharding.barrier(dev)
    clocks = sampler.stop()
    launches = _lib.kernel_launches() - launches0
    # which variant of each coder kernel the timed launches used (chosen by stream count / alignment)
    enc_kernel = _lib.lib().flic_last_coder_kernel(0).decode() or "rans_encode_kernel"
    dec_kernel = _lib.lib().flic_last_coder_kernel(1).decode() or "rans_decode_kernel"
    ms_total = sharding.max_over_ranks(t0.elapsed_time(t1), dev)
    enc_ms = sum(a.elapsed_time(b) for a, b, _, _ in timers) / len(timers)
    dec_ms = sum(c.elapsed_time(d) for _, _, c, d in timers) / len(timers)
    ms_step = ms_total / args.steps
    value = world * n / (ms_step * 1e-3) / 1e6

    # ---- BASELINE.json configs[4], second partition: one stream per image (12288 symbols each)
    off_img = torch.arange(n_img + 1, dtype=torch.int64, device=dev) * PER_IMAGE
    enc_i = rans.encode_streams(x, mean, scale, off_img, workspace=ws, own_output=False)
    xr_i, end_i, st_i = rans.decode_streams(enc_i, mean, scale, off_img, out=out)
    per_image_ok = bool(torch.equal(xr_i, x)) and not bool(st_i.any().item()) and bool((end_i == (1 << 32)).all().item())
    bits_img = (64 * n_img + 32 * enc_i.n_words()) / n
    pe0, pe1 = cuda_events(); pd0, pd1 = cuda_events()
    reps = max(1, min(args.steps, 5))
    pe0.record()
    for _ in range(reps):
        enc_i = rans.encode_streams(x, mean, scale, off_img, workspace=ws, own_output=False)
    pe1.record(); pd0.record()
    for _ in range(reps):
        rans.decode_streams(enc_i, mean, scale, off_img, out=out)
    pd1.record()
    torch.cuda.synchronize()
    pi_enc_ms = sharding.max_over_ranks(pe0.elapsed_time(pe1) / reps, dev)
    pi_dec_ms = sharding.max_over_ranks(pd0.elapsed_time(pd1) / reps, dev)
    per_image = {"streams_per_image": 1, "round_trip_exact": per_image_ok, "bits_per_symbol": round(bits_img, 5),
                 "encode_MBps": round(world * n / (pi_enc_ms * 1e-3) / 1e6, 1),
                 "decode_MBps": round(world * n / (pi_dec_ms * 1e-3) / 1e6, 1),
                 "value": round(world * n / ((pi_enc_ms + pi_dec_ms) * 1e-3) / 1e6, 1)}

    # ---- e2e through the host-buffer C ABI (pinned host arrays; copies inside the timed region)
    e2e = run_e2e(args, dev, rank, world)

    totals = sharding.gather_totals([n_words * 4 + 8 * ns, n], dev)

    # ---- BASELINE.json configs[1] beside it: the whole model (PyTorch fp32 convolutions + this coder)
    full = None
    if not args.no_full:
        del x, mean, scale, out, xr, xr_i, enc, enc_i, ws
        torch.cuda.empty_cache()
        try:
            f = run_full(args, rank, world, dev, steps=3, warmup=3)
            if f is not None:
                full = {k: f[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "config",
                                          "bits_per_dim_container", "e2e", "gpu_launches", "dtype")}
        except Exception as e:  # the sweep numbers stand on their own; say why this part is absent
            full = {"value": None, "error": repr(e)[:200]}
    if rank != 0:
        return None

    peaks = read_peaks()
    alg_bytes = 12.0 + bits_per_symbol / 8.0                      # per symbol, each direction
    dom_name, dom_ms = (dec_kernel, dec_ms) if dec_ms >= enc_ms else (enc_kernel + " (+scan, pack)", enc_ms)
    achieved = n * alg_bytes / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": round(achieved, 2), "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": round(achieved / peaks["hbm_gbs"], 5), "peak_source": peaks["source"],
                "algorithmic_bytes_per_symbol": round(alg_bytes, 4), "symbols_per_launch": n,
                "avg_launch_ms": round(dom_ms, 4), "traffic": read_traffic(dom_name, n), "issue": read_issue(dom_name),
                "encode_ms": round(enc_ms, 4), "decode_ms": round(dec_ms, 4),
                "encode_kernel": enc_kernel, "decode_kernel": dec_kernel,
                "encode_GBps_algorithmic": round(n * alg_bytes / (enc_ms * 1e-3) / 1e9, 2),
                "decode_GBps_algorithmic": round(n * alg_bytes / (dec_ms * 1e-3) / 1e9, 2)}
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64/f32 + u64", "data": "synthetic", "config": workload_config(args, n_img),
        "encode_MBps": round(world * n / (enc_ms * 1e-3) / 1e6, 1), "decode_MBps": round(world * n / (dec_ms * 1e-3) / 1e6, 1),
        "bits_per_symbol": round(bits_per_symbol, 5), "parity": parity, "roofline": roofline, "e2e": e2e,
        "gpu_launches": int(launches), "clocks": clocks,
        "compressed_bytes_all_ranks": int(sum(t[0] for t in totals)),
        "partition_one_stream_per_image": per_image,
        "full_model_configs1": full,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_subprocess(args)
    return line


def run_e2e(args, dev, rank, world) -> dict:
    """The same metric through flic_codec_encode / flic_codec_decode with pinned HOST buffers."""
    import ctypes as C
    import numpy as np
    import torch

    from flic_b200 import _lib, sharding
    L = _lib.lib()
    n_img = min(args.e2e_images, args.images)
    n = n_img * PER_IMAGE
    xs, ms, ss = synth_numpy(n, 99 + rank)
    off = level_major_offsets(n_img)
    ns = off.size - 1

    def pinned(arr):
        t = torch.from_numpy(arr).pin_memory()
        return t
    hx, hm, hs = pinned(xs), pinned(ms), pinned(ss)
    hoff = pinned(off)
    hwords = torch.empty(n, dtype=torch.int32).pin_memory()
    hwoff = torch.empty(ns + 1, dtype=torch.int64).pin_memory()
    hstates = torch.empty(ns, dtype=torch.int64).pin_memory()
    hstatus = torch.empty(ns, dtype=torch.int32).pin_memory()
    hout = torch.empty(n, dtype=torch.float32).pin_memory()
    hend = torch.empty(ns, dtype=torch.int64).pin_memory()
    codec = C.c_void_p()
    _lib.check(L.flic_codec_create(dev.index, n, ns, C.byref(codec)), "flic_codec_create")
    nw = C.c_int64(0)

    def step():
        _lib.check(L.flic_codec_encode(codec, hx.data_ptr(), hm.data_ptr(), hs.data_ptr(), hoff.data_ptr(), ns,
                                       hwords.data_ptr(), n, hwoff.data_ptr(), hstates.data_ptr(), hstatus.data_ptr(),
                                       C.byref(nw)), "flic_codec_encode")
        _lib.check(L.flic_codec_decode(codec, hwords.data_ptr(), hwoff.data_ptr(), hstates.data_ptr(), hm.data_ptr(),
                                       hs.data_ptr(), hoff.data_ptr(), ns, hout.data_ptr(), hend.data_ptr(),
                                       hstatus.data_ptr()), "flic_codec_decode")
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    assert torch.equal(hout, hx) and not bool(hstatus.any()), "e2e round trip failed"
    k = max(1, min(args.steps, 5))
    sharding.barrier(dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(k):
        step()
    torch.cuda.synchronize()
    dt = sharding.max_over_ranks(time.perf_counter() - t0, dev)
    L.flic_codec_destroy(codec)
    words_b = int(nw.value) * 4
    meta = (ns + 1) * 8 + ns * 8
    h2d = 12 * n + (ns + 1) * 8 + words_b + meta + 8 * n + (ns + 1) * 8          # encode inputs + decode inputs
    d2h = words_b + meta + ns * 4 + 4 * n + ns * 12                              # encode outputs + decode outputs
    return {"value": round(world * n * k / dt / 1e6, 2), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "images_per_step_per_gpu": n_img, "ms_per_step": round(1e3 * dt / k, 3),
            "api": "flic_codec_encode + flic_codec_decode (include/flic_b200.h), pinned host buffers"}


def run_full(args, rank, world, dev, steps=None, warmup=None) -> dict | None:
    """BASELINE.json configs[1]: imagenet64.yaml model, batch 256, full compress + decompress."""
    import random

    import torch

    from flic_b200 import _lib, flows, sharding
    layer = dict(name="DenseLayer", act="ReLU")
    block = dict(name="DenseBlock", growth_channel=args.growth, depth=args.depth, layer=layer)
    cfg = dict(name="IDFlows", nflows=8, nbits=8, nsplit=3, H=64, W=64, C=3,
               couple=dict(name="AdditiveCouple", split=0.75, nn=block, round=dict(name="Round", nbits=8)),
               extenddim=dict(name="ExtendDim", scale=2),
               prior=dict(name="Prior", round=dict(name="Round", nbits=8), nn=block),
               distribution=dict(name="DLogistic"), round=dict(name="Round", nbits=8))
    torch.manual_seed(0)
    random.seed(0)
    model = flows.build_model(cfg)
    flows.perturb_heads(model, 0.02)
    model = model.to(dev).eval()
    B = args.batch
    himg = torch.randint(0, 256, (B, 3, 64, 64), dtype=torch.uint8, generator=torch.Generator().manual_seed(1234 + rank)).pin_memory()
    img = himg.to(dev)

    def step_device():
        return model.decompress(model.compress(img, codec_batch=args.codec_batch, check=False), check=False)

    def step_host():
        blob = model.compress(himg.to(dev, non_blocking=True), codec_batch=args.codec_batch, check=False).to_bytes()
        return model.decompress(blob, check=False).cpu(), len(blob)

    steps = args.steps if steps is None else steps
    warmup = args.warmup if warmup is None else warmup
    rec = step_device()
    assert torch.equal(rec, img), "full-path round trip is not lossless"
    for _ in range(max(0, warmup - 1)):
        step_device()
    launches0 = _lib.kernel_launches()
    sampler = ClockSampler(dev.index)
    sharding.barrier(dev)
    torch.cuda.synchronize()
    sampler.start()
    t0, t1 = cuda_events()
    t0.record()
    for _ in range(steps):
        step_device()
    t1.record()
    torch.cuda.synchronize()
    sharding.barrier(dev)
    clocks = sampler.stop()
    launches = _lib.kernel_launches() - launches0
    ms_step = sharding.max_over_ranks(t0.elapsed_time(t1), dev) / steps
    k = max(1, min(steps, 3))
    torch.cuda.synchronize()
    h0 = time.perf_counter()
    for _ in range(k):
        rec_h, nbytes = step_host()
    torch.cuda.synchronize()
    dt = sharding.max_over_ranks(time.perf_counter() - h0, dev)
    assert torch.equal(rec_h, himg)
    if rank != 0:
        return None
    raw = B * 3 * 64 * 64
    return {"metric": METRIC, "value": round(world * raw / (ms_step * 1e-3) / 1e6, 3), "unit": UNIT, "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32 conv + f64/u64 coder", "data": "synthetic",
            "config": {"workload": "configs[1] imagenet64.yaml model (random init, heads N(0,0.02)), batch "
                                   f"{B} uniform-random uint8 3x64x64, IDFlows.compress + decompress",
                       "codec_batch": args.codec_batch, "growth": args.growth, "depth": args.depth,
                       "l2": "activations of one pass exceed L2"},
            "bits_per_dim_container": round(8 * nbytes / raw, 4),
            "e2e": {"value": round(world * raw * k / dt / 1e6, 3), "unit": UNIT, "h2d_bytes_per_step": int(raw),
                    "d2h_bytes_per_step": int(nbytes + raw), "api": "IDFlows.compress(u8 host) -> bytes -> decompress -> u8 host"},
            "gpu_launches": int(launches), "clocks": clocks}


def read_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return {"hbm_gbs": float(p["hbm_gbs"]), "source": "MEASURED_PEAKS.json (measured copy bandwidth)"}
    except Exception:
        return {"hbm_gbs": 6650.0, "source": "fallback 6.65 TB/s (B200_PROFILING.md)"}


def read_issue(kernel: str):
    """Instruction-issue evidence of the dominant kernel from the committed ncu capture: the coder is
    bound by issue slots on exact FP64 arithmetic, not by HBM, so this is what explains roofline.frac."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        rec = t[kernel.split(" ")[0]]
        wi = float(rec["warp_instructions_per_symbol"])
        return {"warp_instructions_per_symbol": wi, "issue_slots_busy_pct": rec.get("issue_slots_busy_pct"),
                "issue_ceiling_Gsymbols_per_s": round(148 * 4 * 1.965 / wi, 1), "source": rec["from"] + " (ncu --set full)"}
    except Exception:
        return None


def read_traffic(kernel: str, n_symbols: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full
    capture (profiles/traffic.json), scaled per symbol to this launch; None when not captured."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        name = kernel.split(" ")[0]
        if name in t:
            return round(t[name]["dram_bytes_per_symbol"] * n_symbols)
    except Exception:
        pass
    return None


def cpu_baseline_subprocess(args) -> dict:
    """The reference coder on the host cores, in a fresh process (no fork after CUDA init)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "1"]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, "WORLD_SIZE": "1", "RANK": "0"})
        line = json.loads(out.stdout.strip().splitlines()[-1])
        return line["cpu_baseline"]
    except Exception as e:  # the GPU numbers stand on their own; say why the baseline is absent
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": f"failed: {e!r}"[:200]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="flic", choices=["flic", "reference"])
    ap.add_argument("--workload", default="sweep", choices=["sweep", "full"])
    ap.add_argument("--images", type=int, default=131072, help="images per step per GPU (sweep)")
    ap.add_argument("--e2e-images", type=int, default=8192)
    ap.add_argument("--batch", type=int, default=256, help="images per step per GPU (full)")
    ap.add_argument("--codec-batch", type=int, default=64)
    ap.add_argument("--growth", type=int, default=512)
    ap.add_argument("--depth", type=int, default=12)
    ap.add_argument("--cpu-procs", type=int, default=0)
    ap.add_argument("--cpu-images-per-proc", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full", action="store_true", help="skip the configs[1] whole-model leg of the sweep line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "flic" else max(args.warmup, 1)

    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return 0                                  # under torchrun only rank 0 runs the CPU arm
        print(json.dumps(run_reference(args)), flush=True)
        return 0
    line = run_flic(args)
    if line is not None:
        print(json.dumps(line), flush=True)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
