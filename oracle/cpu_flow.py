"""CPU restatement of the reference's whole coding path for an IDFlows model.  TEST / BASELINE
INFRASTRUCTURE ONLY (bench.py's `cpu_baseline` of the full-model workloads; the product never
imports oracle/).

It does, on torch CPU tensors and with the operations the reference uses, what one evaluation
step of the reference does when `test_coding` is on, followed by the inverse pass:

    forward      flows.py:87-116      per level: ExtendDim (extenddim.py:23-29: view / permute /
                                      contiguous), [Permute, AdditiveCouple] x nflows + Permute
                                      (invertible.py:38-42: NCHW->NHWC copy, F.linear with the
                                      permutation matrix, NHWC->NCHW copy; couplelib.py:47-53 with
                                      Round = rint(x 256)/256 through the straight-through form of
                                      roundlib.py:18-38, then torch.cat), split, Prior
                                      (priorlib.py:36-47)
    coding loop  trainer.py:308-327   per level: state 1<<32, .tolist() x 3 (torch.exp of the
                                      logscales first), encode, decode with reversed inputs,
                                      rebuild the tensor, count errors, real bpd by the
                                      reference's formula
    inverse      flows.py:139-152     generated_from_latents: cat, flows.backward, extend.backward

The convolutions are the mirror model's own DenseBlock modules (plain torch.nn, identical to the
reference's nnblock.py), run on the CPU; the coder is the reference's own Cython module
(oracle/_ref) when it has been built, else the C restatement (oracle/liboracle.so).
"""
from __future__ import annotations

import time

import numpy as np
import torch
from torch.nn import functional as F

from . import pyoracle


def _squeeze(x, s):
    b, c, h, w = x.shape
    x = x.view(b, c, h // s, s, w // s, s).permute(0, 1, 3, 5, 2, 4).contiguous()
    return x.view(b, c * s * s, h // s, w // s)


def _unsqueeze(x, s):
    b, c, h, w = x.shape
    x = x.view(b, c // s // s, s, s, h, w).permute(0, 1, 4, 2, 5, 3).contiguous()
    return x.view(b, c // s // s, h * s, w * s)


def _permute(x, P):
    x = x.permute(0, 2, 3, 1).contiguous()
    x = F.linear(x, P)
    return x.permute(0, 3, 1, 2).contiguous()


def _round(x, nbits=8):
    y = x * (2.0 ** nbits)
    y = y + (torch.round(y) - y).detach()
    return y / (2.0 ** nbits)


def _flow(block, x, direction):
    from flic_b200.couplelib import AdditiveCouple
    mods = list(block["flows"])
    for m in (mods if direction > 0 else reversed(mods)):
        if isinstance(m, AdditiveCouple):
            xa, xb = x[:, :m.a_ch], x[:, m.a_ch:]
            t = _round(m.dense(xa))
            x = torch.cat([xa, xb + t if direction > 0 else xb - t], dim=1)
        else:
            x = _permute(x, m.P if direction > 0 else m.inv_P)
    return x


@torch.no_grad()
def forward(model, x):
    """flows.py:87-116 -> (latents, means, logscales)."""
    latents, means, logscales = [], [], []
    for level in range(model.nsplit):
        block = model.blocks[level]
        x = _squeeze(x, block["extend"].scale)
        x = _flow(block, x, +1)
        if level < model.nsplit - 1:
            half = x.shape[1] // 2
            z, x = x[:, :half], x[:, half:]
            mean, logscale = block["prior"](x)
        else:
            z = x
            mean, logscale = block["prior"](x)
        latents.append(z)
        means.append(mean)
        logscales.append(logscale)
    return latents, means, logscales


@torch.no_grad()
def generated_from_latents(model, latents):
    """flows.py:139-152."""
    x = None
    for level in reversed(range(model.nsplit)):
        block = model.blocks[level]
        z = latents[level]
        x = z if level == model.nsplit - 1 else torch.cat((z, x), dim=1)
        x = _flow(block, x, -1)
        x = _unsqueeze(x, block["extend"].scale)
    return x


def coding_loop(latents, means, logscales):
    """trainer.py:308-327.  Returns (decoded latents, words, errors, t_encode, t_decode) with the
    reference's own split of the time (t3 - t1 includes the .tolist() calls, t4 - t3 the rebuild)."""
    ref = pyoracle.ref_rans()
    out, words, errors, t_en, t_de = [], 0, 0, 0.0, 0.0
    for x, m, ls in zip(latents, means, logscales):
        t1 = time.time()
        state = 1 << 32
        if ref is not None:
            xi = x.reshape(-1).tolist()
            mi = m.reshape(-1).tolist()
            si = torch.exp(ls).reshape(-1).tolist()
            state, buf = ref.encode(state, len(xi), xi, mi, si)
            t3 = time.time()
            state_, msg = ref.decode(state, buf[::-1], len(xi), mi[::-1], si[::-1])
            latent = torch.tensor(msg[::-1]).reshape(*x.shape).to(x)
        else:
            xi = x.reshape(-1).numpy()
            mi = m.reshape(-1).contiguous().numpy()
            si = torch.exp(ls).reshape(-1).contiguous().numpy()
            state, buf = pyoracle.encode(state, xi.size, xi, mi, si)
            t3 = time.time()
            state_, msg = pyoracle.decode(state, buf[::-1], xi.size, mi[::-1], si[::-1])
            latent = torch.from_numpy(np.ascontiguousarray(msg[::-1])).reshape(*x.shape).to(x)
        t4 = time.time()
        errors += int(torch.sum(x != latent)) + int(state_ != 1 << 32)
        words += len(buf)
        t_en += t3 - t1
        t_de += t4 - t3
        out.append(latent)
    return out, words, errors, t_en, t_de


def full_path(model_cpu, images_u8: torch.Tensor) -> dict:
    """One batch through forward + coding loop + inverse on the CPU; timings in seconds."""
    x = torch.from_numpy(pyoracle.quantise_input_u8(images_u8.numpy()))
    t0 = time.time()
    latents, means, logscales = forward(model_cpu, x)
    t1 = time.time()
    decoded, words, errors, t_en, t_de = coding_loop(latents, means, logscales)
    t2 = time.time()
    rec = generated_from_latents(model_cpu, decoded)
    t3 = time.time()
    lossless = bool(torch.equal(rec, x))
    return {"forward_s": t1 - t0, "coding_s": t2 - t1, "encode_s": t_en, "decode_s": t_de, "inverse_s": t3 - t2,
            "total_s": t3 - t0, "words": words, "errors": errors, "lossless": lossless,
            "real_bpd": (64 * len(latents) + 32 * words) / images_u8.numel(),
            "coder": "reference rans.pyx (oracle/_ref)" if pyoracle.ref_rans() is not None else "C restatement (oracle/liboracle.so)"}
