#!/usr/bin/env python
"""Benchmark of the entropy-coding hot path (contract: see the repo prompt / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W]                   # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]  # the reference's CPU coder

Headline workload (BASELINE.json configs[4], the rANS-only sweep): ImageNet64-shaped images, 12 288
symbols each = the imagenet64.yaml latent levels 6144 / 3072 / 3072, with
precomputed logistic parameters drawn as rans/test.py:8-10.  One rANS stream per image x level.
A step is one pass of the hot path over one chunk of `--images` images per GPU: fused table
evaluation + rANS encode + stream concatenation, then rANS decode back to symbols.  Ranks shard
the image range; there is no data-path collective (weak scaling: every rank codes its own chunk).

metric  = MB/s of raw pixels (1 symbol = 1 sub-pixel = 1 byte, MB = 1e6 B) through
          encode + decode, i.e. symbols / (t_encode + t_decode).
value   = inputs resident in HBM, CUDA-event timed.   e2e = same through the host-buffer C ABI
          (flic_codec_encode / flic_codec_decode) with pinned host arrays, copies inside the timing,
          next to `host_ceiling_MBps`: the same copies with no kernel (flic_codec_probe_copies).

The same line carries, at N = 1, the other BASELINE.json configs and the reference-shaped calls:
  stream_count_sweep     device-timed encode / decode at 3 / 48 / 768 / 9936 / 98 304 / 393 216 streams
  flow_kernels           achieved HBM GB/s of K1, K5, permute, squeeze, u8 <-> grid, log_prob sums
  list_api               the drop-in rans.encode / rans.decode (Python lists) on 1.5 M symbols
  configs                configs[0] (config1.yaml @32x32), configs[2] (resflows_smallpatch),
                         configs[3] (config_twolevel): whole model, compress + decompress
  full_model_configs1    configs[1] (imagenet64.yaml, batch 256) with the same path on the host CPU
                         (oracle/cpu_flow.py: forward + trainer.py:308-327 loop + inverse) beside it
and at every N:
  sweep1m                configs[4] as a STRONG-scaling run: 1 048 576 images in fixed chunks generated
                         on the device from the chunk index, sharded over the ranks
`--workload full|config1|patches|twolevel|sweep1m|streams` makes one of those the headline instead.
"""
from __future__ import annotations

import argparse
import hashlib
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
SWEEP_CHUNK = 131072               # images per chunk of the strong-scaling sweep (393 216 streams, 8 chunks)


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
    # x = round((mean + scale U(-5, 5)) 256) / 256; float32 products of a 9-bit mean grid and a 24-bit
    # scale stay far inside the window either way, and float64 temporaries would double the footprint
    u = 10 * torch.rand(n, device=device, generator=g) - 5
    x = torch.round((mean + scale * u) * 256) / 256
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
        "config": workload_config(), "images_per_step": int(cores * imgs_per_proc),
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "value_per_core": round(one_core, 4), "encode_MBps_per_core": round(n_sym / enc_cpu / 1e6 * 1, 4),
                         "decode_MBps_per_core": round(n_sym / dec_cpu / 1e6 * 1, 4)},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "errors": errors, "bits_per_symbol": round((64 * cores * imgs_per_proc * 3 * len(times) + 32 * words) / (n_sym * len(times)), 5),
    }


def workload_config():
    """The same object in both arms (the step size is a top-level key of its own: MB/s does not
    depend on it, and the CPU arm takes a bounded sample of the workload)."""
    return {"workload": "configs[4] rANS-only sweep: ImageNet64-shaped images (12288 symbols = latent levels "
                        "6144/3072/3072), logistic params as rans/test.py:8-10, one stream per image x level",
            "symbols_per_image": PER_IMAGE, "streams_per_image": 3,
            "partition": "level-major, image-major inside a level",
            "l2": "inputs (12 B/symbol) exceed the 126 MB L2 many times over; no flush needed"}


# ------------------------------------------------------------------------------------------------
# the CUDA arm
# ------------------------------------------------------------------------------------------------

def cuda_events():
    import torch
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


_flush_buf = None


def flush_l2(dev):
    """Write a buffer larger than the 126 MB L2 (between timed iterations of workloads that fit in it)."""
    import torch
    global _flush_buf
    if _flush_buf is None or _flush_buf.device != dev:
        _flush_buf = torch.empty(64 << 20, dtype=torch.float32, device=dev)     # 256 MB
    _flush_buf.fill_(1.0)


def timed_each(fn, steps: int, warmup: int, dev, flush: bool = True):
    """Per-iteration CUDA-event timing with an L2 flush (untimed) in between.  Returns ms per step."""
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(dev)
    total = 0.0
    for _ in range(steps):
        if flush:
            flush_l2(dev)
        a, b = cuda_events()
        a.record(); fn(); b.record()
        torch.cuda.synchronize(dev)
        total += a.elapsed_time(b)
    return total / max(steps, 1)


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

    if args.workload in MODEL_LEGS:
        line = run_model(args, args.workload, rank, world, dev)
        if line is not None and world == 1 and not args.no_cpu_baseline and args.workload == "full":
            line["cpu_baseline"] = cpu_full_path_baseline(args)
        return line
    if args.workload == "sweep1m":
        s = run_sweep1m(args, rank, world, dev)
        if s is None:
            return None
        return {"metric": METRIC, "value": s["value"], "unit": UNIT, "n_gpus": world, "steps": s["steps"], "warmup": s["warmup"],
                "ms_per_step": s["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64/f32 + u64", "data": "synthetic", "config": s["config"], "sweep1m": s,
                "gpu_launches": s["gpu_launches"], "clocks": s["clocks"]}
    if args.workload == "streams":
        sweep = stream_count_sweep(dev)
        if rank != 0:
            return None
        big = sweep[-1]
        return {"metric": METRIC, "value": round(big["symbols"] / (big["encode_ms"] + big["decode_ms"]) / 1e3, 2), "unit": UNIT,
                "n_gpus": world, "steps": 3, "warmup": 1, "ms_per_step": round(big["encode_ms"] + big["decode_ms"], 4),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64/f32 + u64",
                "data": "synthetic", "config": workload_config(), "stream_count_sweep": sweep}

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
        # cost of the partition against the reference's native one (one stream per level
        # over a batch of 256 images, trainer.py:308-315) on the same symbols.  north_star: within 0.1 %.
        nb = min(256, n_img)
        sel = torch.cat([torch.arange(nb * seg, device=dev) + base for seg, base in
                         zip(SEGMENTS, np.cumsum([0] + [s_ * n_img for s_ in SEGMENTS[:-1]]).tolist())])
        xb, mb, sb = x[sel].contiguous(), mean[sel].contiguous(), scale[sel].contiguous()
        native_off = torch.tensor(np.cumsum([0] + [s_ * nb for s_ in SEGMENTS]), dtype=torch.int64, device=dev)
        fine_off = torch.from_numpy(level_major_offsets(nb)).to(dev)
        nat = rans.encode_streams(xb, mb, sb, native_off)
        fin = rans.encode_streams(xb, mb, sb, fine_off)
        # one CHAINED stream per image (coder.py:18-27; IDFlows.compress's default): the state runs through the levels
        levels, carried, base_ = [], None, 0
        for seg in SEGMENTS:
            e = rans.encode_streams(xb[base_:base_ + nb * seg], mb[base_:base_ + nb * seg], sb[base_:base_ + nb * seg],
                                    rans.uniform_offsets(nb, seg, dev), init_states=carried, workspace=rans.Workspace(),
                                    own_output=False)
            carried = e.final_states
            levels.append(e)
            base_ += nb * seg
        chained = rans.chain_levels(levels)
        nsym_b = nb * PER_IMAGE
        bits_native, bits_fine, bits_chain = nat.bits() / nsym_b, fin.bits() / nsym_b, chained.bits() / nsym_b
        parity["bits_per_symbol_256_images"] = {
            "native_one_stream_per_level": round(bits_native, 5),
            "one_stream_per_image_x_level": round(bits_fine, 5),
            "one_chained_stream_per_image": round(bits_chain, 5),
            "overhead_pct": round(100 * (bits_fine / bits_native - 1), 4),
            "chained_overhead_pct": round(100 * (bits_chain / bits_native - 1), 4),
            "chained_within_0.1pct_of_reference_partition": bool(bits_chain / bits_native - 1 < 1e-3)}
        del xb, mb, sb, nat, fin, sel, levels, chained
    enc = rans.encode_streams(x, mean, scale, off, workspace=ws, validate=False)
    xr, end, st = rans.decode_streams(enc, mean, scale, off, out=out, validate=False)
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
        enc_l = rans.encode_streams(x, mean, scale, off, workspace=ws, own_output=False, validate=False)
        if timers is not None:
            e1.record(); d0.record()
        rans.decode_streams(enc_l, mean, scale, off, out=out, validate=False)
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

    # ---- BASELINE.json configs[4], second partition: one stream per image (12288 symbols each), which is
    # ---- what IDFlows.compress's chained streams amount to on the coder's side
    off_img = torch.arange(n_img + 1, dtype=torch.int64, device=dev) * PER_IMAGE
    enc_i = rans.encode_streams(x, mean, scale, off_img, workspace=ws, own_output=False, validate=False)
    xr_i, end_i, st_i = rans.decode_streams(enc_i, mean, scale, off_img, out=out, validate=False)
    per_image_ok = bool(torch.equal(xr_i, x)) and not bool(st_i.any().item()) and bool((end_i == (1 << 32)).all().item())
    bits_img = (64 * n_img + 32 * enc_i.n_words()) / n
    pe0, pe1 = cuda_events(); pd0, pd1 = cuda_events()
    reps = max(1, min(args.steps, 5))
    pe0.record()
    for _ in range(reps):
        enc_i = rans.encode_streams(x, mean, scale, off_img, workspace=ws, own_output=False, validate=False)
    pe1.record(); pd0.record()
    for _ in range(reps):
        rans.decode_streams(enc_i, mean, scale, off_img, out=out, validate=False)
    pd1.record()
    torch.cuda.synchronize()
    pi_enc_ms = sharding.max_over_ranks(pe0.elapsed_time(pe1) / reps, dev)
    pi_dec_ms = sharding.max_over_ranks(pd0.elapsed_time(pd1) / reps, dev)
    per_image = {"streams_per_image": 1, "round_trip_exact": per_image_ok, "bits_per_symbol": round(bits_img, 5),
                 "encode_MBps": round(world * n / (pi_enc_ms * 1e-3) / 1e6, 1),
                 "decode_MBps": round(world * n / (pi_dec_ms * 1e-3) / 1e6, 1),
                 "value": round(world * n / ((pi_enc_ms + pi_dec_ms) * 1e-3) / 1e6, 1)}

    totals = sharding.gather_totals([n_words * 4 + 8 * ns, n], dev)
    del x, mean, scale, out, xr, xr_i, enc, enc_i, ws
    torch.cuda.empty_cache()

    # ---- e2e through the host-buffer C ABI (pinned host arrays; copies inside the timed region)
    e2e = run_e2e(args, dev, rank, world)

    # ---- configs[4] as a strong-scaling run (every N)
    sweep1m = None
    if not args.no_sweep1m:
        try:
            sweep1m = run_sweep1m(args, rank, world, dev, steps=1)
        except Exception as e:
            sweep1m = {"value": None, "error": repr(e)[:200]}

    # ---- the other configs, the reference-shaped calls and the whole model (N = 1 only: they are per-GPU replicas)
    extras = {}
    full = None
    if world == 1 and not args.no_extras:
        for name, fn in (("stream_count_sweep", lambda: stream_count_sweep(dev)),
                         ("flow_kernels", lambda: flow_kernels(dev)),
                         ("list_api", lambda: list_api(dev)),
                         ("configs", lambda: {leg: _leg_summary(run_model(args, leg, rank, world, dev, steps=2, warmup=3))
                                              for leg in ("config1", "patches", "twolevel")})):
            try:
                extras[name] = fn()
            except Exception as e:  # the headline stands on its own; say why a leg is absent
                extras[name] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()
    if not args.no_full and world == 1:
        try:
            f = run_model(args, "full", rank, world, dev, steps=3, warmup=3)
            if f is not None:
                full = _leg_summary(f)
                if not args.no_cpu_baseline:
                    full["cpu_baseline"] = cpu_full_path_baseline(args)
                    cb = full["cpu_baseline"]
                    if cb.get("value"):
                        full["speedup_vs_cpu_full_path"] = {"device": round(f["value"] / cb["value"], 1),
                                                            "e2e": round(f["e2e"]["value"] / cb["value"], 1)}
        except Exception as e:
            full = {"value": None, "error": repr(e)[:200]}
    if rank != 0:
        return None

    peaks = read_peaks()
    alg_bytes = 12.0 + bits_per_symbol / 8.0                      # per symbol, each direction
    dom_name, dom_ms = (dec_kernel, dec_ms) if dec_ms >= enc_ms else (enc_kernel + " (+scan, pack)", enc_ms)
    achieved = n * alg_bytes / (dom_ms * 1e-3) / 1e9
    prof = read_profile(dom_name)
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": round(achieved, 2), "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": round(achieved / peaks["hbm_gbs"], 5), "peak_source": peaks["source"],
                "algorithmic_bytes_per_symbol": round(alg_bytes, 4), "symbols_per_launch": n,
                "avg_launch_ms": round(dom_ms, 4),
                "traffic": round(prof["dram_bytes_per_symbol"] * n) if prof else None,
                "issue": ({"warp_instructions_per_symbol": prof["warp_instructions_per_symbol"],
                           "issue_slots_busy_pct": prof.get("issue_slots_busy_pct"),
                           "issue_ceiling_Gsymbols_per_s": round(148 * 4 * 1.965 / prof["warp_instructions_per_symbol"], 1),
                           "source": prof["from"] + " (ncu --set full; kernel sources unchanged since)"} if prof else None),
                "encode_ms": round(enc_ms, 4), "decode_ms": round(dec_ms, 4),
                "encode_kernel": enc_kernel, "decode_kernel": dec_kernel,
                "encode_GBps_algorithmic": round(n * alg_bytes / (enc_ms * 1e-3) / 1e9, 2),
                "decode_GBps_algorithmic": round(n * alg_bytes / (dec_ms * 1e-3) / 1e9, 2)}
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64/f32 + u64", "data": "synthetic", "config": workload_config(),
        "images_per_step": int(n_img) * world,
        "encode_MBps": round(world * n / (enc_ms * 1e-3) / 1e6, 1), "decode_MBps": round(world * n / (dec_ms * 1e-3) / 1e6, 1),
        "bits_per_symbol": round(bits_per_symbol, 5), "parity": parity, "roofline": roofline, "e2e": e2e,
        "gpu_launches": int(launches), "clocks": clocks,
        "compressed_bytes_all_ranks": int(sum(t[0] for t in totals)),
        "partition_one_stream_per_image": per_image,
        "sweep1m": sweep1m,
        "full_model_configs1": full,
    }
    line.update(extras)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_subprocess(args)
        la, cb = line.get("list_api"), line["cpu_baseline"]
        if isinstance(la, dict) and "encode_MBps" in la and cb.get("encode_MBps_per_core"):
            la["reference_one_core"] = {"encode_MBps": cb["encode_MBps_per_core"], "decode_MBps": cb["decode_MBps_per_core"],
                                        "what": "the reference's coder called the same way (tolist + encode, decode + rebuild) "
                                                "on one host core, from the cpu_baseline run"}
    return line


def run_e2e(args, dev, rank, world) -> dict:
    """The same metric through flic_codec_encode / flic_codec_decode with pinned HOST buffers, and the
    ceiling of that call pattern on this host: the same copies without the kernels."""
    import ctypes as C
    import torch

    from flic_b200 import _lib, sharding
    L = _lib.lib()
    n_img = min(args.e2e_images, args.images)
    n = n_img * PER_IMAGE
    xs, ms, ss = synth_numpy(n, 99 + rank)
    off = level_major_offsets(n_img)
    ns = off.size - 1

    def pinned(arr):
        return torch.from_numpy(arr).pin_memory()
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

    def probe():
        _lib.check(L.flic_codec_probe_copies(codec, hx.data_ptr(), hm.data_ptr(), hs.data_ptr(), hoff.data_ptr(), ns,
                                             hwords.data_ptr(), int(nw.value), hout.data_ptr()), "flic_codec_probe_copies")
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
    keep_words = hwords[: int(nw.value)].clone()
    probe()
    sharding.barrier(dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(k):
        probe()
    torch.cuda.synchronize()
    dt_probe = sharding.max_over_ranks(time.perf_counter() - t0, dev)
    L.flic_codec_destroy(codec)
    del keep_words
    words_b = int(nw.value) * 4
    meta = (ns + 1) * 8 + ns * 8
    h2d = 12 * n + (ns + 1) * 8 + words_b + meta + 8 * n + (ns + 1) * 8          # encode inputs + decode inputs
    d2h = words_b + meta + ns * 4 + 4 * n + ns * 12                              # encode outputs + decode outputs
    value = world * n * k / dt / 1e6
    ceiling = world * n * k / dt_probe / 1e6
    return {"value": round(value, 2), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "images_per_step_per_gpu": n_img, "ms_per_step": round(1e3 * dt / k, 3),
            "host_ceiling_MBps": round(ceiling, 2), "fraction_of_host_ceiling": round(value / ceiling, 4),
            "host_ceiling": "flic_codec_probe_copies: the same chunked copies on the same three streams and pinned "
                            "buffers with no kernel in between, all ranks at once",
            "link_GBps_at_ceiling": round((h2d + d2h) * k / dt_probe / 1e9, 2),
            "api": "flic_codec_encode + flic_codec_decode (include/flic_b200.h), pinned host buffers"}


# ------------------------------------------------------------------------------------------------
# reference-shaped calls
# ------------------------------------------------------------------------------------------------

def stream_count_sweep(dev):
    """Device-timed coder rates as a function of the stream partition (tools/stream_sweep.py)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_stream_sweep", os.path.join(ROOT, "tools", "stream_sweep.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.sweep(dev, reps=3)


def flow_kernels(dev):
    """Achieved HBM bandwidth of the kernels around the coder -- K1 (CDF tables), K5 (coupling add / round),
    permute, squeeze, u8 <-> grid, fused log_prob sums -- on 32 768 ImageNet64-shaped images
    (tools/prof_flowops.py: measure())."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_prof_flowops", os.path.join(ROOT, "tools", "prof_flowops.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.measure(32768, dev)


def list_api(dev, n: int = 1_572_864):
    """The drop-in itself: rans.encode / rans.decode over Python lists (trainer.py:311-318), one call =
    one stream.  n = level 0 of a batch of 256 ImageNet64 images, the reference's own largest call."""
    import numpy as np
    from flic_b200 import rans
    xs, ms, ss = synth_numpy(n, 7)
    t0 = time.perf_counter()
    xl, ml, sl = xs.tolist(), ms.tolist(), ss.tolist()                          # what trainer.py:311-313 does
    t_tolist = time.perf_counter() - t0
    # untimed first call at full size: library load, codec creation and growth, first-touch of the staging
    # buffers (the reference arm's processes are warm when they are timed, too)
    st_w, buf_w = rans.encode(1 << 32, n, xl, ml, sl)
    rans.decode(st_w, buf_w[::-1], n, ml[::-1], sl[::-1])
    del st_w, buf_w
    import torch
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    state, buf = rans.encode(1 << 32, n, xl, ml, sl)
    t2 = time.perf_counter()
    end, msg = rans.decode(state, buf[::-1], n, ml[::-1], sl[::-1])
    rec = np.asarray(msg[::-1], dtype=np.float32)
    t3 = time.perf_counter()
    return {"symbols": n, "streams": 1, "round_trip_exact": bool(np.array_equal(rec, xs)) and end == 1 << 32,
            "tolist_s": round(t_tolist, 4), "encode_s": round(t2 - t1, 4), "decode_s": round(t3 - t2, 4),
            "encode_MBps": round(n / (t_tolist + t2 - t1) / 1e6, 3), "decode_MBps": round(n / (t3 - t2) / 1e6, 3),
            "what": "rans.encode(state, n, x_, mean_, scale_) / rans.decode(...) with Python lists, timed like "
                    "trainer.py:309-323 (encode includes the .tolist() calls, decode the rebuild of the array)"}


# ------------------------------------------------------------------------------------------------
# whole models (PyTorch fp32 convolutions + this repository's kernels)
# ------------------------------------------------------------------------------------------------

MODEL_LEGS = ("full", "config1", "patches", "twolevel")


def _idflows_cfg(H, W, nsplit, nflows, couple_g, couple_d, prior_g, prior_d, act, scale=2):
    layer = dict(name="DenseLayer", act=act)
    return dict(name="IDFlows", nflows=nflows, nbits=8, nsplit=nsplit, H=H, W=W, C=3,
                couple=dict(name="AdditiveCouple", split=0.75, round=dict(name="Round", nbits=8),
                            nn=dict(name="DenseBlock", growth_channel=couple_g, depth=couple_d, layer=dict(layer))),
                extenddim=dict(name="ExtendDim", scale=scale),
                prior=dict(name="Prior", round=dict(name="Round", nbits=8),
                           nn=dict(name="DenseBlock", growth_channel=prior_g, depth=prior_d, layer=dict(layer))),
                distribution=dict(name="DLogistic"), round=dict(name="Round", nbits=8))


def model_leg(args, name):
    """(config dict, uint8 host images, description, compress kwargs) of one BASELINE.json config."""
    import torch
    g = torch.Generator().manual_seed(1234)
    if name == "full":          # configs[1]: configs/imagenet64.yaml:3-42, batch 256
        cfg = _idflows_cfg(64, 64, 3, 8, args.growth, args.depth, args.growth, args.depth, "ReLU")
        img = torch.randint(0, 256, (args.batch, 3, 64, 64), dtype=torch.uint8, generator=g)
        what = (f"configs[1] imagenet64.yaml model (random init, heads N(0,0.02)), batch {args.batch} uniform-random "
                "uint8 3x64x64, IDFlows.compress + decompress")
        return cfg, img, what, dict(codec_batch=args.codec_batch)
    if name == "config1":       # configs[0]: configs/config1.yaml:3-42 with H = W = 32, 16 images
        cfg = _idflows_cfg(32, 32, 3, 8, 384, 8, 512, 12, "LeakyReLU")
        img = torch.randint(0, 256, (16, 3, 32, 32), dtype=torch.uint8, generator=g)
        what = ("configs[0] config1.yaml model at H = W = 32 (random init, heads N(0,0.02)), 16 uniform-random uint8 "
                "3x32x32 images, IDFlows.compress + decompress (16 chained streams, 48 level segments)")
        return cfg, img, what, dict()
    if name == "patches":       # configs[2]: configs/resflows_smallpatch.yaml:3-43,73-75
        cfg = _idflows_cfg(8, 8, 1, 12, 512, 12, 256, 4, "ReLU")
        big = torch.randint(0, 256, (16, 3, 216, 184), dtype=torch.uint8, generator=g)
        img = big.unfold(2, 8, 8).unfold(3, 8, 8).permute(0, 2, 3, 1, 4, 5).reshape(-1, 3, 8, 8).contiguous()
        what = ("configs[2] resflows_smallpatch.yaml flow (random init, heads N(0,0.02)) on the 8x8 patches of 16 "
                "uniform-random uint8 3x216x184 images: 9936 rANS streams of 192 symbols")
        return cfg, img, what, dict(codec_batch=2484)
    if name == "twolevel":      # configs[3]: configs/config_twolevel.yaml:3-93
        fine = _idflows_cfg(8, 8, 1, 12, 512, 8, 512, 8, "ReLU")
        rough = _idflows_cfg(27, 23, 1, 12, 512, 8, 512, 8, "ReLU", scale=1)
        cfg = dict(name="TwoLevelFlows", H=215, W=178, C=3, pad=[1, 6], fine_flows=fine, rough_flows=rough, batchsize=1536)
        img = torch.randint(0, 256, (8, 3, 215, 178), dtype=torch.uint8, generator=g)
        what = ("configs[3] config_twolevel.yaml model (random init, heads N(0,0.02)), 8 uniform-random uint8 3x215x178 "
                "images: rough 27x23 image + 621 fine 8x8 residual patches each (8 + 4968 streams)")
        return cfg, img, what, dict()
    raise ValueError(name)


def run_model(args, leg, rank, world, dev, steps=None, warmup=None) -> dict | None:
    """One BASELINE.json model config: compress + decompress of a batch, device-resident and through
    the public API with host buffers.  Small batches fit the L2, so it is flushed between steps."""
    import random

    import torch

    from flic_b200 import _lib, flows, sharding
    cfg, himg, what, kw = model_leg(args, leg)
    torch.manual_seed(0)
    random.seed(0)
    model = flows.build_model(cfg)
    flows.perturb_heads(model, 0.02)
    model = model.to(dev).eval()
    himg = himg.pin_memory()
    img = himg.to(dev)
    two = leg == "twolevel"

    def step_device():
        if two:
            return model.decompress(model.compress(img, check=False), check=False)
        return model.decompress(model.compress(img, check=False, **kw), check=False)

    def step_host():
        if two:
            blob = model.compress(himg.to(dev, non_blocking=True), check=False)
        else:
            blob = model.compress(himg.to(dev, non_blocking=True), check=False, **kw).to_bytes()
        return model.decompress(blob, check=False).cpu(), len(blob)

    steps = args.steps if steps is None else steps
    warmup = max(3, args.warmup if warmup is None else warmup)
    rec = step_device()
    assert torch.equal(rec, img), f"{leg}: round trip is not lossless"
    for _ in range(warmup - 1):
        step_device()
    launches0 = _lib.kernel_launches()
    sampler = ClockSampler(dev.index)
    sharding.barrier(dev)
    torch.cuda.synchronize()
    sampler.start()
    ms_local = timed_each(step_device, steps, 0, dev, flush=True)
    sharding.barrier(dev)
    clocks = sampler.stop()
    launches = (_lib.kernel_launches() - launches0) // max(steps, 1)
    ms_step = sharding.max_over_ranks(ms_local, dev)
    k = max(1, min(steps, 3))
    torch.cuda.synchronize()
    h0 = time.perf_counter()
    for _ in range(k):
        rec_h, nbytes = step_host()
    torch.cuda.synchronize()
    dt = sharding.max_over_ranks(time.perf_counter() - h0, dev)
    assert torch.equal(rec_h, himg)
    coder = None if two else coder_share(model, img, kw, dev)
    if rank != 0:
        return None
    raw = himg.numel()
    return {"metric": METRIC, "value": round(world * raw / (ms_step * 1e-3) / 1e6, 3), "unit": UNIT, "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32 conv + f64/u64 coder", "data": "synthetic",
            "config": {"workload": what, **{k_: v for k_, v in kw.items()},
                       "l2": "a buffer larger than the L2 is written between timed steps"},
            "bits_per_dim_container": round(8 * nbytes / raw, 4),
            "e2e": {"value": round(world * raw * k / dt / 1e6, 3), "unit": UNIT, "h2d_bytes_per_step": int(raw),
                    "d2h_bytes_per_step": int(nbytes + raw), "api": "compress(u8 host) -> bytes -> decompress -> u8 host"},
            "gpu_launches": int(launches), "clocks": clocks, "coder": coder}


def coder_share(model, img, kw, dev):
    """The coder's part of one model step: the same (z, mean, scale) coded once more on their own, at the
    model's partition (one chained stream per image), device-timed with an L2 flush in between."""
    import torch

    from flic_b200 import rans
    stats = []
    batch = model.compress(img, check=False, stats=stats, **kw)
    n_levels = model.nsplit
    cbs = batch.codec_batch
    chunk = stats[:n_levels]                                   # the first chunk's levels
    n_real = min(cbs, img.shape[0])
    levels = []
    for level, (z, mean, logscale) in enumerate(chunk):
        seg = math.prod(model.latents_shape[level])
        n = n_real * seg
        levels.append((z.reshape(-1)[:n].contiguous(), mean.contiguous().reshape(-1)[:n].contiguous(),
                       torch.exp(logscale.contiguous()).reshape(-1)[:n].contiguous(), rans.uniform_offsets(n_real, seg, dev)))
    wss = [rans.Workspace() for _ in levels]

    def encode():
        carried, out = None, []
        for (z, m, s, off), ws in zip(levels, wss):
            e = rans.encode_streams(z, m, s, off, init_states=carried, workspace=ws, own_output=False, validate=False)
            carried = e.final_states
            out.append(e)
        return rans.chain_levels(out)
    enc = encode()

    def decode():
        st = left = None
        for lvl in reversed(range(len(levels))):
            z, m, s, off = levels[lvl]
            _, st, _, left = rans.decode_streams(enc, m, s, off, check_end=lvl == 0, validate=False, states=st,
                                                 words_left=left, return_words_left=True)
    enc_ms = timed_each(encode, 3, 2, dev)
    dec_ms = timed_each(decode, 3, 2, dev)
    nsym = sum(l[0].numel() for l in levels)
    bits = enc.bits() / nsym
    alg = 12.0 + bits / 8.0
    peak = read_peaks()["hbm_gbs"]
    return {"symbols": nsym, "streams": n_real, "levels": n_levels, "bits_per_symbol": round(bits, 4),
            "encode_ms": round(enc_ms, 4), "decode_ms": round(dec_ms, 4),
            "encode_Msym_per_s": round(nsym / enc_ms / 1e3, 1), "decode_Msym_per_s": round(nsym / dec_ms / 1e3, 1),
            "roofline_frac_encode": round(nsym * alg / (enc_ms * 1e-3) / 1e9 / peak, 5),
            "roofline_frac_decode": round(nsym * alg / (dec_ms * 1e-3) / 1e9 / peak, 5),
            "what": "first codec chunk, one chained stream per image; launches this small are latency-bound, the "
                    "roofline fractions say how far from the HBM bound"}


def _leg_summary(f):
    if f is None:
        return None
    return {k: f[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "config", "bits_per_dim_container", "e2e",
                              "gpu_launches", "dtype", "coder") if k in f}


def cpu_full_path_baseline(args) -> dict:
    """configs[1] on the host CPU, the reference's way: IDFlows.forward + the trainer.py:308-327 coding
    loop + generated_from_latents (oracle/cpu_flow.py), on a bounded sample of the same batch."""
    import random

    import torch

    from flic_b200 import flows
    from oracle import cpu_flow
    try:
        cfg, himg, what, _ = model_leg(args, "full")
        torch.manual_seed(0)
        random.seed(0)
        model = flows.build_model(cfg)
        flows.perturb_heads(model, 0.02)
        model = model.eval()
        n = min(args.cpu_full_images, himg.shape[0])
        cpu_flow.full_path(model, himg[:1])                                     # warm-up (thread pools, allocator)
        r = cpu_flow.full_path(model, himg[:n])
        raw = n * 3 * 64 * 64
        return {"value": round(raw / r["total_s"] / 1e6, 5), "unit": UNIT, "cores": int(torch.get_num_threads()),
                "host_cores": len(os.sched_getaffinity(0)), "kind": "port",
                "sample": f"{n} of the {himg.shape[0]} images, one batch: torch-CPU forward {r['forward_s']:.2f} s, coding loop "
                          f"{r['coding_s']:.2f} s ({r['coder']}; one thread, as the reference), inverse {r['inverse_s']:.2f} s",
                "lossless": r["lossless"], "errors": r["errors"], "real_bpd": round(r["real_bpd"], 4),
                "seconds_per_image": round(r["total_s"] / n, 4)}
    except Exception as e:
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": f"failed: {e!r}"[:200]}


# ------------------------------------------------------------------------------------------------
# configs[4] as a strong-scaling run
# ------------------------------------------------------------------------------------------------

def run_sweep1m(args, rank, world, dev, steps=None) -> dict | None:
    """BASELINE.json configs[4]: 1 M synthetic ImageNet64-shaped images with precomputed logistic
    parameters, sharded over the ranks.  The set is cut into fixed chunks of SWEEP_CHUNK images; chunk c is
    generated on the device from a generator seeded with c (so its content, and its compressed size, do not
    depend on which rank codes it) and the chunks are dealt out by sharding.shard_range.  Timed: the coder
    (encode + decode of every chunk, CUDA events around each chunk, summed), max over ranks; generation is
    not the path and is left out.  value = all images' symbols / that time."""
    import torch

    from flic_b200 import _lib, rans, sharding
    total_images = args.sweep_images
    n_chunks = (total_images + SWEEP_CHUNK - 1) // SWEEP_CHUNK
    lo, hi = sharding.shard_range(n_chunks, rank, world)
    steps = max(1, min(args.steps, 2)) if steps is None else steps
    ws = rans.Workspace()
    off_cache = {}
    launches0 = _lib.kernel_launches()
    sampler = ClockSampler(dev.index)
    bytes_rank = words_rank = 0
    ok = True
    coder_ms = gen_s = 0.0
    sharding.barrier(dev)
    torch.cuda.synchronize()
    sampler.start()
    wall0 = time.perf_counter()
    for step in range(steps + 1):                       # pass 0 is the warm-up pass
        ms_pass = 0.0
        for c in range(lo, hi):
            imgs = min(SWEEP_CHUNK, total_images - c * SWEEP_CHUNK)
            n = imgs * PER_IMAGE
            g0 = time.perf_counter()
            x, mean, scale = synth_torch(n, 7_000_000 + c, dev)
            if imgs not in off_cache:
                off_cache[imgs] = torch.from_numpy(level_major_offsets(imgs)).to(dev)
            off = off_cache[imgs]
            out = torch.empty(n, dtype=torch.float32, device=dev)
            torch.cuda.synchronize()
            gen_s += time.perf_counter() - g0
            a, b = cuda_events()
            a.record()
            enc = rans.encode_streams(x, mean, scale, off, workspace=ws, own_output=False, validate=False)
            xr, end, st = rans.decode_streams(enc, mean, scale, off, out=out, validate=False)
            b.record()
            torch.cuda.synchronize()
            ms_pass += a.elapsed_time(b)
            if step == 0:
                ok = ok and bool(torch.equal(xr, x)) and not bool(st.any().item()) and bool((end == (1 << 32)).all().item())
                nw = enc.n_words()
                words_rank += nw
                bytes_rank += 4 * nw + 8 * (off.numel() - 1)
            del x, mean, scale, out, enc, xr, end, st
        if step > 0:
            coder_ms += ms_pass
    wall = time.perf_counter() - wall0
    sharding.barrier(dev)
    clocks = sampler.stop()
    launches = _lib.kernel_launches() - launches0
    ms_step = sharding.max_over_ranks(coder_ms / steps, dev)
    per_rank = sharding.gather_totals([bytes_rank, (hi - lo), int(ok), int(round(coder_ms / steps * 1000))], dev)
    if rank != 0:
        return None
    n_total = total_images * PER_IMAGE
    return {"value": round(n_total / (ms_step * 1e-3) / 1e6, 2), "unit": UNIT, "scaling": "strong", "n_gpus": world,
            "steps": steps, "warmup": 1, "ms_per_step": round(ms_step, 3),
            "images_total": total_images, "symbols_total": n_total, "chunk_images": SWEEP_CHUNK, "chunks": n_chunks,
            "compressed_bytes_total": int(sum(r[0] for r in per_rank)),
            "compressed_bytes_per_rank": [int(r[0]) for r in per_rank], "chunks_per_rank": [int(r[1]) for r in per_rank],
            "coder_ms_per_rank": [r[3] / 1000 for r in per_rank],
            "round_trip_exact": all(bool(r[2]) for r in per_rank),
            "bits_per_symbol": round(8 * sum(r[0] for r in per_rank) / n_total, 5),
            "wall_s_incl_generation": round(wall, 2), "generation_s_this_rank": round(gen_s, 2),
            "gpu_launches": int(launches), "clocks": clocks,
            "config": {"workload": "configs[4] rANS-only sweep, strong scaling: 1M ImageNet64-shaped images in fixed chunks "
                                   "generated on the device from the chunk index, one stream per image x level; timed: "
                                   "encode + decode of every chunk (CUDA events), max over ranks; generation excluded",
                       "l2": "a chunk's inputs (19 GB) exceed the L2"}}


# ------------------------------------------------------------------------------------------------
# files the line quotes
# ------------------------------------------------------------------------------------------------

def read_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return {"hbm_gbs": float(p["hbm_gbs"]), "source": "MEASURED_PEAKS.json (measured copy bandwidth)"}
    except Exception:
        return {"hbm_gbs": 6650.0, "source": "fallback 6.65 TB/s (B200_PROFILING.md)"}


KERNEL_SOURCES = {"rans_encode_lane_kernel": "rans_encode.cu", "rans_encode_kernel": "rans_encode.cu",
                  "rans_decode_lane_kernel": "rans_decode.cu", "rans_decode_kernel": "rans_decode.cu",
                  "rans_decode_coop_kernel": "rans_decode_coop.cu", "cdf_tables_kernel": "cdf_tables.cu"}


def kernel_source_hash(kernel: str = "") -> str:
    """SHA-256 over the CUDA sources `kernel` is built from (its .cu file, every shared header and the
    build script with the nvcc flags; all sources when the kernel is not in KERNEL_SOURCES):
    profiles/traffic.json records it when an ncu capture is summarised, and the bench quotes the
    capture only while it still matches."""
    h = hashlib.sha256()
    pkg = os.path.join(ROOT, "finalproject-losslessimagecompression_b200")
    d = os.path.join(pkg, "csrc")
    own = KERNEL_SOURCES.get(kernel.split(" ")[0].split("<")[0])
    for name in sorted(os.listdir(d)):
        if name.endswith(".cuh") or (name.endswith(".cu") and (own is None or name == own)):
            h.update(name.encode())
            h.update(open(os.path.join(d, name), "rb").read())
    h.update(open(os.path.join(pkg, "build.py"), "rb").read())
    return h.hexdigest()[:16]


def read_profile(kernel: str):
    """DRAM bytes and instructions per symbol of the dominant kernel from the committed `ncu --set full`
    capture (profiles/traffic.json) -- or None when the kernel sources have changed since it was taken."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        rec = t[kernel.split(" ")[0]]
        if rec.get("source_hash") != kernel_source_hash(kernel):
            return None
        return rec
    except Exception:
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
    ap.add_argument("--workload", default="sweep", choices=["sweep", "sweep1m", "streams", *MODEL_LEGS])
    ap.add_argument("--images", type=int, default=131072, help="images per step per GPU (sweep)")
    ap.add_argument("--e2e-images", type=int, default=8192)
    ap.add_argument("--sweep-images", type=int, default=1 << 20, help="images of the strong-scaling sweep (all ranks)")
    ap.add_argument("--batch", type=int, default=256, help="images per step per GPU (full)")
    ap.add_argument("--codec-batch", type=int, default=64)
    ap.add_argument("--growth", type=int, default=512)
    ap.add_argument("--depth", type=int, default=12)
    ap.add_argument("--cpu-procs", type=int, default=0)
    ap.add_argument("--cpu-images-per-proc", type=int, default=0)
    ap.add_argument("--cpu-full-images", type=int, default=8, help="images of the whole-model CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full", action="store_true", help="skip the configs[1] whole-model leg of the sweep line")
    ap.add_argument("--no-extras", action="store_true", help="skip stream_count_sweep / list_api / configs legs")
    ap.add_argument("--no-sweep1m", action="store_true", help="skip the strong-scaling leg")
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
