"""GPU tests of everything around the coder kernels: coupling add/round (K5), index kernels (N1),
the flow mirror against the reference's golden outputs, compress/decompress, the drop-in list
API, the coder.py wrappers, the host-buffer C ABI and the exhaustive expf sweep."""
import json
import os
import struct

import numpy as np
import pytest
import torch

from _data import gen, ragged_offsets
from test_host_logic import GOLDEN, build_tiny, tiny_cfg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


# ---- K5 / N1 against the reference's own module outputs ------------------------------------

def test_couple_add_round_matches_reference_module(golden, oracle):
    from flic_b200.couplelib import couple_add_round
    a = int(golden["k5.a_ch"])
    x = torch.from_numpy(golden["k5.x"]).cuda()
    t = torch.from_numpy(golden["k5.t"]).cuda()
    z = couple_add_round(x.clone(), t, a, +1)
    assert np.array_equal(z.cpu().numpy(), golden["k5.z"])          # AdditiveCouple.forward, couplelib.py:47-53
    back = couple_add_round(z.clone(), t, a, -1)
    assert torch.equal(back, x)                                      # .backward, couplelib.py:55-61


@pytest.mark.parametrize("shape,a_ch", [((5, 12, 32, 32), 9), ((3, 24, 16, 16), 18), ((2, 48, 8, 8), 36),
                                        ((3, 3, 27, 23), 2), ((1, 4, 1, 1), 3), ((2, 12, 4, 4), 9)])
def test_couple_add_round_ties_and_shapes(oracle, shape, a_ch):
    """Ties-to-even on every k/512 input, odd sizes (scalar path) and aligned ones (float4 path)."""
    from flic_b200.couplelib import couple_add_round
    g = torch.Generator().manual_seed(shape[1])
    B, C, H, W = shape
    x = torch.randint(-4096, 4096, shape, generator=g).float() / 256
    t = torch.randint(-4096, 4096, (B, C - a_ch, H, W), generator=g).float() / 512        # half of them exact ties
    t[0].mul_(1.0000001)
    for direction, ref in ((+1, oracle.couple_forward), (-1, oracle.couple_backward)):
        out = couple_add_round(x.cuda().clone(), t.cuda(), a_ch, direction)
        want = x.numpy().copy()
        want[:, a_ch:] = ref(x.numpy()[:, a_ch:], t.numpy())
        assert np.array_equal(out.cpu().numpy(), want)


def test_round_trip_of_coupling_is_exact_at_full_size():
    """imagenet64 x 256 level-0 coupling shape: forward then backward restores x bit for bit."""
    from flic_b200.couplelib import couple_add_round
    x = torch.round(torch.randn(256, 12, 32, 32, device="cuda") * 64) / 256
    t = torch.randn(256, 3, 32, 32, device="cuda")
    z = couple_add_round(x.clone(), t, 9, +1)
    assert not torch.equal(z, x)
    assert torch.equal(couple_add_round(z, t, 9, -1), x)


def test_permute_and_squeeze_match_reference_modules(golden):
    from flic_b200 import extenddim
    model = build_tiny().cuda()
    perm = model.blocks[0]["flows"][0]
    x = torch.from_numpy(golden["k5.x"]).cuda()
    assert np.array_equal(perm.P.cpu().numpy(), golden["n1.perm_P"])
    assert np.array_equal(perm.forward(x, None)[0].cpu().numpy(), golden["n1.perm_fwd"])     # invertible.py:38-42
    assert np.array_equal(perm.backward(x).cpu().numpy(), golden["n1.perm_bwd"])             # invertible.py:44-48
    img = torch.from_numpy(golden["id.x"]).cuda()
    assert np.array_equal(extenddim.squeeze(img, 2, +1).cpu().numpy(), golden["n1.squeeze_fwd"])   # extenddim.py:23-29
    assert np.array_equal(extenddim.squeeze(x, 2, -1).cpu().numpy(), golden["n1.squeeze_bwd"])     # extenddim.py:31-37


@pytest.mark.parametrize("shape,scale", [((3, 3, 64, 64), 2), ((2, 12, 32, 32), 2), ((2, 5, 9, 6), 3), ((4, 3, 27, 23), 1),
                                         ((2, 4, 6, 12), 2), ((1, 2, 10, 8), 2), ((5, 3, 216, 184), 2)])
def test_squeeze_against_torch_expression(shape, scale):
    from flic_b200 import extenddim
    x = torch.randn(shape, device="cuda")
    B, C, H, W = shape
    want = x.view(B, C, H // scale, scale, W // scale, scale).permute(0, 1, 3, 5, 2, 4).contiguous() \
        .view(B, C * scale * scale, H // scale, W // scale)
    got = extenddim.squeeze(x, scale, +1)
    assert torch.equal(got, want)
    assert torch.equal(extenddim.squeeze(got, scale, -1), x)


def test_u8_grid_conversion(oracle):
    from flic_b200 import flows
    u8 = torch.arange(256, dtype=torch.uint8).repeat(7)[:1777].cuda()
    grid = flows.u8_to_grid(u8)
    assert np.array_equal(grid.cpu().numpy(), oracle.quantise_input_u8(u8.cpu().numpy()))
    back, status = flows.grid_to_u8(grid)
    assert torch.equal(back, u8) and int(status.item()) == 0
    _, status = flows.grid_to_u8(torch.tensor([0.5, 0.3], device="cuda"))      # 128/256 and an off-grid value
    assert int(status.item()) != 0
    # the 16-pixels-per-thread bodies: one bad value anywhere in a vector is reported; unaligned views
    # take the scalar kernels and give the same answer
    for bad in (0.5, 0.3, -1.0 / 256, 257.0 / 256):
        g2 = grid[:1024].clone()
        g2[517] = bad
        assert int(flows.grid_to_u8(g2)[1].item()) != 0
    big = torch.randint(0, 256, (4 * 3 * 64 * 64 + 3,), dtype=torch.uint8, device="cuda")
    want = oracle.quantise_input_u8(big.cpu().numpy())
    assert np.array_equal(flows.u8_to_grid(big).cpu().numpy(), want)
    assert np.array_equal(flows.u8_to_grid(big[1:]).cpu().numpy(), want[1:])
    gbig = torch.from_numpy(want).cuda()
    for view in (gbig, gbig[1:], gbig[4:]):
        back, status = flows.grid_to_u8(view)
        assert torch.equal(back, big[big.numel() - view.numel():]) and int(status.item()) == 0


# ---- the flow mirror against the reference's forward ----------------------------------------

def _close_latents(got, want):
    """Latents are k/256 values; a coupling's t = dense(xa) is computed by cuDNN here and by the
    CPU conv in the reference, so t can differ in the last ulp and flip a rounding tie.  Bit
    equality is required of all but a tiny fraction, and a flipped value must be one grid step away."""
    diff = np.abs(got - want)
    frac = float((diff != 0).mean())
    return frac, float(diff.max())


def test_idflows_forward_matches_reference_golden(golden):
    model = build_tiny().cuda()
    x = torch.from_numpy(golden["id.x"]).cuda()
    lat, means, logs, _ = model.forward(x, None)
    for i in range(2):
        frac, mx = _close_latents(lat[i].cpu().numpy(), golden[f"id.latent{i}"])
        assert frac < 5e-3 and mx <= 4 / 256 + 1e-6, (i, frac, mx)
        assert np.allclose(means[i].cpu().numpy(), golden[f"id.mean{i}"], atol=2e-3)          # fp32 conv, tolerance 2e-3
        assert np.allclose(logs[i].cpu().numpy(), golden[f"id.logscale{i}"], atol=2e-3)
    assert torch.equal(model.generated_from_latents(lat), x)                                  # flows.py:139-152
    logp, _ = model.log_likelihood(lat, means, logs)
    assert np.allclose(logp.cpu().numpy(), golden["id.logp"], rtol=2e-3)


def test_zero_head_model_is_bit_exact_with_reference_structure(golden):
    """With the reference's zero-initialised heads every coupling adds Round(0): latents are pure
    permutations / squeezes of the input and must match a torch re-expression bit for bit."""
    import random
    from flic_b200 import flows
    torch.manual_seed(0)
    random.seed(0)
    model = flows.build_model(tiny_cfg()).cuda().eval()
    x = torch.from_numpy(golden["id.x"]).cuda()
    lat, means, logs, _ = model.forward(x, None)
    assert all(float(m.abs().max()) == 0 and float(l.abs().max()) == 0 for m, l in zip(means, logs))
    cur = x
    for level in range(2):
        B, C, H, W = cur.shape
        cur = cur.view(B, C, H // 2, 2, W // 2, 2).permute(0, 1, 3, 5, 2, 4).reshape(B, C * 4, H // 2, W // 2)
        for m in model.blocks[level]["flows"]:
            if hasattr(m, "P"):
                cur = torch.nn.functional.linear(cur.permute(0, 2, 3, 1), m.P).permute(0, 3, 1, 2).contiguous()
        if level == 0:
            half = cur.shape[1] // 2
            assert torch.equal(lat[0], cur[:, :half])
            cur = cur[:, half:]
        else:
            assert torch.equal(lat[1], cur)


def test_conditional_flows_forward(golden):
    model = build_tiny("ConditionalFlows", conv_for_cond=False).cuda()
    x = torch.from_numpy(golden["id.x"]).cuda()
    cond = torch.from_numpy(golden["cond.cond"]).cuda()
    lat, means, logs, _ = model.forward(x, None, cond)
    for i in range(2):
        frac, mx = _close_latents(lat[i].cpu().numpy(), golden[f"cond.latent{i}"])
        assert frac < 5e-3 and mx <= 4 / 256 + 1e-6
        assert np.allclose(means[i].cpu().numpy(), golden[f"cond.mean{i}"], atol=2e-3)
        assert np.allclose(logs[i].cpu().numpy(), golden[f"cond.logscale{i}"], atol=2e-3)


# ---- compress / decompress -------------------------------------------------------------------

@pytest.mark.parametrize("n,codec_batch,sps", [(6, None, 1), (5, 4, 1), (7, 3, 2), (4, 4, 0), (1, 8, 1)])
def test_compress_decompress_lossless(oracle, n, codec_batch, sps):
    model = build_tiny().cuda()
    img = torch.randint(0, 256, (n, 3, 16, 16), dtype=torch.uint8, generator=torch.Generator().manual_seed(n)).cuda()
    stats = []
    batch = model.compress(img, codec_batch=codec_batch, streams_per_segment=sps, stats=stats, chain_levels=False)
    blob = batch.to_bytes()
    rec = model.decompress(blob)
    assert torch.equal(rec, img)
    # every stream is what the reference coder produces for the same slice of (z, mean, scale)
    cbs = batch.codec_batch
    k = 0
    for ci, chunk in enumerate(batch.sections):
        n_real = min(cbs, n - ci * cbs)
        for level, enc in enumerate(chunk):
            z, mean, logscale = stats[k]
            k += 1
            nsym = n_real * int(np.prod(model.latents_shape[level]))
            scale = torch.exp(logscale.contiguous())
            off = model._segment_offsets(level, n_real, sps, "cpu").numpy()
            words, woff, states, status = oracle.encode_streams(
                z.reshape(-1)[:nsym].cpu().numpy(), mean.contiguous().reshape(-1)[:nsym].cpu().numpy(),
                scale.reshape(-1)[:nsym].cpu().numpy(), off)
            assert np.array_equal(enc.words.cpu().numpy().view(np.uint32)[:words.size], words)
            assert np.array_equal(enc.final_states.cpu().numpy().view(np.uint64), states)
    # bits/dim: reference accounting vs the ideal code length (within 0.1% + the per-stream constant)
    lat, means, logs, _ = model.forward(model_input(img), None)
    logp, _ = model.log_likelihood(lat, means, logs)
    ideal_bpd = float((-logp / np.log(2)).mean())
    real_bpd = batch.bits_per_dim()
    stream_bpd = 64 * batch.n_streams() / img.numel()
    assert real_bpd >= ideal_bpd - 1e-3
    assert real_bpd - stream_bpd <= ideal_bpd * 1.001 + 32 * batch.n_streams() / img.numel()


@pytest.mark.parametrize("n,codec_batch", [(6, None), (5, 4), (1, 8), (9, 2)])
def test_chained_streams_match_the_reference_coder_chained(oracle, n, codec_batch):
    """The default partition: ONE stream per image whose state runs through all latent levels, as the
    reference's coder.Encode chains it (coder.py:18-27: encode level 0, 1, ... each from the state
    the previous one ended in, buffers appended).  Every image's (final state, words) must be what
    the reference coder returns when called that way on the image's slices of (z, mean, scale); the
    decoder continues each stream level by level, last level first (coder.py:29-38), and the cost
    is 64 bits per IMAGE plus the words (trainer.py:326-327's formula with one stream per image)."""
    model = build_tiny().cuda()
    img = torch.randint(0, 256, (n, 3, 16, 16), dtype=torch.uint8, generator=torch.Generator().manual_seed(40 + n)).cuda()
    stats = []
    batch = model.compress(img, codec_batch=codec_batch, stats=stats)
    assert batch.chained and batch.n_streams() == n
    assert torch.equal(model.decompress(batch.to_bytes()), img)
    assert torch.equal(model.decompress(batch), img)
    cbs = batch.codec_batch
    n_levels = model.nsplit
    total_words = 0
    for ci, chunk in enumerate(batch.sections):
        assert len(chunk) == 1
        enc = chunk[0]
        n_real = min(cbs, n - ci * cbs)
        words = enc.words.cpu().numpy().view(np.uint32)
        woff = enc.word_offsets.cpu().numpy()
        states = enc.final_states.cpu().numpy().view(np.uint64)
        for b in range(n_real):
            state, bufs = 1 << 32, []
            for level in range(n_levels):
                z, mean, logscale = stats[ci * n_levels + level]
                seg = int(np.prod(model.latents_shape[level]))
                sl = slice(b * seg, (b + 1) * seg)
                scale = torch.exp(logscale.contiguous())
                state, buf = oracle.encode(state, seg, z.reshape(-1)[sl].cpu().numpy(),
                                           mean.contiguous().reshape(-1)[sl].cpu().numpy(), scale.reshape(-1)[sl].cpu().numpy())
                bufs.append(buf)
            want = np.concatenate(bufs) if bufs else np.zeros(0, np.uint32)
            assert int(states[b]) == state
            assert np.array_equal(words[woff[b]:woff[b + 1]], want)
            total_words += want.size
    assert batch.reference_bits() == 64 * n + 32 * total_words
    # against one stream per image x level: the same words up to a few per image, 64 bits fewer per extra level
    per_level = model.compress(img, codec_batch=codec_batch, chain_levels=False)
    assert per_level.n_streams() == n * n_levels
    assert batch.reference_bits() < per_level.reference_bits()


def model_input(img):
    from flic_b200 import flows
    return flows.u8_to_grid(img)


def test_reference_native_partition_matches_trainer_loop(golden, oracle):
    """streams_per_segment=0 is the reference's partition: one stream per level over the whole
    batch from state 1<<32 (trainer.py:308-315); real bpd by its formula (trainer.py:326-327)."""
    model = build_tiny().cuda()
    img = torch.from_numpy(golden["id.u8"]).cuda()
    stats = []
    batch = model.compress(img, streams_per_segment=0, stats=stats)
    total_words = 0
    for level, enc in enumerate(batch.sections[0]):
        z, mean, logscale = stats[level]
        scale = torch.exp(logscale.contiguous())
        state, buf = oracle.encode(1 << 32, z.numel(), z.reshape(-1).cpu().numpy(),
                                   mean.contiguous().reshape(-1).cpu().numpy(), scale.reshape(-1).cpu().numpy())
        assert int(enc.final_states.cpu().numpy().view(np.uint64)[0]) == state
        assert np.array_equal(enc.words.cpu().numpy().view(np.uint32)[:buf.size], buf)
        total_words += buf.size
    assert batch.bits_per_dim() == (64 * 2 + 32 * total_words) / img.numel()
    # within 0.1% of the reference's own run on its CPU latents (cuDNN vs CPU conv differ in the last ulp)
    assert abs(batch.bits_per_dim() - float(golden["id.real_bpd"])) / float(golden["id.real_bpd"]) < 1e-3
    assert torch.equal(model.decompress(batch), img)


def test_conditional_compress_decompress(golden):
    model = build_tiny("ConditionalFlows", conv_for_cond=False).cuda()
    img = torch.from_numpy(golden["id.u8"]).cuda()
    cond = torch.from_numpy(golden["cond.cond"]).cuda()
    batch = model.compress(img, cond=cond, codec_batch=4)
    assert torch.equal(model.decompress(batch.to_bytes(), cond=cond), img)


def test_corrupt_container_is_detected():
    model = build_tiny().cuda()
    img = torch.randint(0, 256, (3, 3, 16, 16), dtype=torch.uint8).cuda()
    blob = bytearray(model.compress(img).to_bytes())
    blob[-9] ^= 0x40
    try:
        rec = model.decompress(bytes(blob))
    except ValueError:
        return                        # reported through a stream status (bad end state / no symbol)
    assert not torch.equal(rec, img)  # rANS is bijective: a flipped bit cannot decode to the same image


# ---- drop-in list API, coder.py wrappers, host codec ------------------------------------------

def test_drop_in_encode_decode_lists():
    from flic_b200 import rans
    kat = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat8.json")))
    state, buf = rans.encode(1 << 32, 8, kat["x"], kat["mean"], kat["scale"])
    assert (state, buf) == (kat["state"], kat["buf"])
    assert isinstance(buf, list) and all(isinstance(w, int) for w in buf)
    end, msg = rans.decode(state, buf[::-1], 8, kat["mean"][::-1], kat["scale"][::-1])
    assert end == 1 << 32 and msg[::-1] == kat["x"] and all(isinstance(v, float) for v in msg)
    with pytest.raises(ZeroDivisionError):
        rans.encode(1 << 32, 1, [0.0], [0.0], [0.0])
    with pytest.raises(ValueError):
        rans.encode(1 << 32, 1, [100.0], [0.0], [1.0])


def test_drop_in_matches_reference_goldens_like_rans_test_py():
    """rans/test.py's flow (encode, reverse, decode, compare, final state 1<<32) on its distribution."""
    import hashlib
    import math
    import random
    from flic_b200 import rans
    kat = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat_random.json")))["200000"]
    n = 200_000
    random.seed(0)
    mean = [random.randint(-256, 256) / 256 for _ in range(n)]
    scale = [math.exp(10 * random.random() - 5) / 256 for _ in range(n)]
    msg = [round((mean[i] + scale[i] * (10 * random.random() - 5)) * 256) / 256 for i in range(n)]
    x, buf = rans.encode(1 << 32, n, msg, mean, scale)
    assert x == kat["state"] and len(buf) == kat["n_words"]
    assert hashlib.sha256(struct.pack("<%dI" % len(buf), *buf)).hexdigest() == kat["sha256"]
    x, rec = rans.decode(x, buf[::-1], n, mean[::-1], scale[::-1])
    assert x == 1 << 32 and rec[::-1] == msg


def test_coder_wrappers_chain_state_like_coder_py(oracle):
    """coder.Encode / Decode chain the rANS state from level to level (coder.py:25,36).  The
    reference's chain only decodes correctly when the first symbol of every later level does not
    push a word; level 1 therefore starts with a near-certain symbol
    (freq ~ 2^24, so freq << 40 can never be reached).  The buggy case is checked to be
    *reported* rather than silently mis-decoded."""
    from flic_b200 import coder
    g = torch.Generator().manual_seed(3)
    shapes = [(2, 6, 8, 8), (2, 24, 4, 4)]
    means = [torch.randint(-32, 33, s, generator=g).float().div(256) for s in shapes]
    logscales = [(torch.rand(s, generator=g) * 0.01 - 0.005) for s in shapes]
    lat = [torch.round((m + torch.exp(l) * (torch.rand(m.shape, generator=g) - 0.5)) * 256) / 256
           for m, l in zip(means, logscales)]
    logscales[1].view(-1)[0] = -20.0
    lat[1].view(-1)[0] = means[1].view(-1)[0]
    means, logscales, lat = ([t.cuda() for t in ts] for ts in (means, logscales, lat))
    x, buffers = coder.Encode(lat, means, logscales)
    st = 1 << 32
    scales = [torch.exp(l).reshape(-1).cpu().numpy() for l in logscales]
    for i in range(2):                                  # coder.py:18-27 with the oracle as rans
        st, buf = oracle.encode(st, lat[i].numel(), lat[i].reshape(-1).cpu().numpy(),
                                means[i].reshape(-1).cpu().numpy(), scales[i])
        assert buffers[i] == buf.tolist()
    assert x == st
    x2, rec = coder.Decode(buffers, means, logscales, x)
    assert x2 == 1 << 32 and all(torch.equal(a, b) for a, b in zip(rec, lat))
    so = x
    for i in (1, 0):                                    # coder.py:29-38 with the oracle as rans
        so, msg = oracle.decode(so, np.asarray(buffers[i], np.uint32)[::-1], lat[i].numel(),
                                means[i].reshape(-1).cpu().numpy()[::-1], scales[i][::-1])
        assert np.array_equal(msg[::-1], lat[i].reshape(-1).cpu().numpy())
    assert so == x2
    # the reference's latent bug: make level 1's first symbol push a word -> its chain cannot be decoded
    logscales[1].view(-1)[0] = 0.0
    x, buffers = coder.Encode(lat, means, logscales)
    try:
        _, rec = coder.Decode(buffers, means, logscales, x)
        assert not all(torch.equal(a, b) for a, b in zip(rec, lat))
    except ValueError:
        pass


def test_host_codec_c_abi_with_host_buffers(oracle):
    from flic_b200 import rans
    n, ns = 120_000, 77
    x, mean, scale = gen("test", n, 17)
    off = ragged_offsets(n, ns, 5)
    codec = rans.HostCodec(n, ns)
    words, woff, states, status = codec.encode(x, mean, scale, off)
    words_o, woff_o, states_o, _ = oracle.encode_streams(x, mean, scale, off, n_threads=4)
    assert not status.any()
    assert np.array_equal(words, words_o) and np.array_equal(woff, woff_o) and np.array_equal(states, states_o)
    xr, end, status = codec.decode(words, woff, states, mean, scale, off)
    assert np.array_equal(xr, x) and (end == 1 << 32).all() and not status.any()
    with pytest.raises(Exception):
        codec.encode(np.zeros(n + 1, np.float32), np.zeros(n + 1, np.float32), np.ones(n + 1, np.float32), [0, n + 1])
    codec.close()


# ---- expf: the device restatement against the host libm over the whole reachable domain ------

def test_device_expf_exhaustive_against_host_libm(oracle):
    """Every float with |x| <= 104 (2.24e9 inputs): the device function must equal this host's
    expf except where the host itself departs from the published algorithm (2 inputs, both with
    part1 saturated), and must equal the restatement everywhere."""
    import ctypes as C
    from concurrent.futures import ThreadPoolExecutor
    from flic_b200 import _lib
    hi = struct.unpack("<I", struct.pack("<f", 104.0))[0] + 1
    chunk = 1 << 27
    bad_host, bad_rest, firsts = 0, 0, []
    L = _lib.lib()
    for base in (0, 0x80000000):
        for lo in range(0, hi, chunk):
            n = min(chunk, hi - lo)
            bits = torch.arange(base + lo, base + lo + n, dtype=torch.int64, device="cuda").to(torch.int32)
            xin = bits.view(torch.float32)
            y = torch.empty(n, dtype=torch.float32, device="cuda")
            _lib.check(L.flic_debug_expf(xin.data_ptr(), y.data_ptr(), n, None))
            got = y.cpu().numpy()
            parts = 8
            step = (n + parts - 1) // parts

            def run(k):
                a, b = k * step, min((k + 1) * step, n)
                return oracle.expf_compare(base + lo + a, got[a:b]) if a < b else (0, 0, np.zeros(0, np.uint32))
            with ThreadPoolExecutor(parts) as ex:
                for bh, br, first in ex.map(run, range(parts)):
                    bad_host += bh
                    bad_rest += br
                    firsts += first.tolist()
    assert bad_rest == 0
    assert set(firsts) <= {0x4202422f, 0xc27c65d9} and bad_host <= 2


@pytest.mark.parametrize("chunk", [4096, 70_000])
def test_host_codec_chunked_pipeline(oracle, monkeypatch, chunk):
    """A call larger than one slot is cut into chunks of whole streams that rotate through the
    codec's slots; the result must not depend on where the cuts fall.  One stream is longer than a
    slot (the slots grow), some are empty, and the partition is ragged."""
    from flic_b200 import rans
    monkeypatch.setenv("FLIC_CODEC_CHUNK_SYMBOLS", str(chunk))
    n = 260_000
    x, mean, scale = gen("test", n, 31)
    rng = np.random.default_rng(7)
    cuts = np.sort(rng.integers(0, n - 90_000, 300))
    off = np.concatenate([[0], cuts, cuts[-1:], [cuts[-1] + 90_000], [n]]).astype(np.int64)   # empty + 90k-symbol streams
    ns = off.size - 1
    codec = rans.HostCodec(n, ns)
    words, woff, states, status = codec.encode(x, mean, scale, off)
    assert not status.any()
    w_o, wo_o, st_o, _ = oracle.encode_streams(x, mean, scale, off, n_threads=8)
    assert np.array_equal(words, w_o) and np.array_equal(woff, wo_o) and np.array_equal(states, st_o)
    rec, end, status = codec.decode(words, woff, states, mean, scale, off)
    assert not status.any() and np.array_equal(rec, x)
    assert (end == (1 << 32)).all()
    # single-stream drop-ins still work on the same codec after the slots grew
    st1, buf1 = codec.encode_single(1 << 32, 50_000, x[:50_000], mean[:50_000], scale[:50_000])
    st_ref, buf_ref = oracle.encode(1 << 32, 50_000, x[:50_000], mean[:50_000], scale[:50_000])
    assert st1 == st_ref and np.array_equal(buf1, buf_ref)
    codec.close()


def test_device_part1_exhaustive_over_every_float_argument(oracle):
    """part1 of CDF() is a function of the float argument alone once the division is done
    (rans.pyx:25-26,34).  Sweep every non-NaN float (4.28e9 arguments, both infinities included):
    the device chain -- glibc-expf body, roundings on the FP64 pipe or by conversion, reciprocal,
    scaling, roundf -- must give the integer the reference arithmetic gives with this host's libm."""
    import ctypes as C
    from concurrent.futures import ThreadPoolExecutor
    from flic_b200 import _lib
    L = _lib.lib()
    chunk = 1 << 27
    parts = min(16, max(4, len(os.sched_getaffinity(0))))
    bad, firsts = 0, []
    for base in (0, 0x80000000):
        for lo in range(0, 0x7f800001, chunk):          # up to and including +/-inf
            n = min(chunk, 0x7f800001 - lo)
            bits = torch.arange(base + lo, base + lo + n, dtype=torch.int64, device="cuda").to(torch.int32)
            y = torch.empty(n, dtype=torch.int32, device="cuda")
            _lib.check(L.flic_debug_part1(bits.view(torch.float32).data_ptr(), y.data_ptr(), n, None))
            got = y.cpu().numpy()
            step = (n + parts - 1) // parts

            def run(k):
                a, b = k * step, min((k + 1) * step, n)
                return oracle.part1_compare(base + lo + a, got[a:b]) if a < b else (0, np.zeros(0, np.uint32))
            with ThreadPoolExecutor(parts) as ex:
                for b_, first in ex.map(run, range(parts)):
                    bad += b_
                    firsts += first.tolist()
    assert bad == 0, [hex(f) for f in firsts[:16]]


# ---- N4: fused DLogistic log-probability / sampler (floating point: tolerance, not bit equality) --

@pytest.mark.parametrize("shape", [(7, 6, 16, 16), (3, 48, 4, 4), (1, 1, 1, 5)])
def test_fused_dlogistic_log_prob_matches_torch_formula(shape):
    """flic_dlogistic_log_prob against the reference's formula (distlib.py:40-55) evaluated by
    torch in fp32 on the same device: elementwise within 1e-5 (abs + rel; the kernel runs the same
    float operations, only the summation order of the reduction differs), per-image sums within
    1e-5 relative (the kernel accumulates in double, torch.sum in float)."""
    from flic_b200.distlib import DLogistic
    g = torch.Generator(device="cuda").manual_seed(5)
    mean = (torch.rand(shape, device="cuda", generator=g) - 0.5) * 2
    logscale = (torch.rand(shape, device="cuda", generator=g) - 0.5) * 8 - 3
    x = torch.round((mean + torch.exp(logscale) * torch.randn(shape, device="cuda", generator=g) * 2) * 256) / 256
    x[0].view(-1)[0] = 7.0                       # far tail: log(eps) floor
    dist = DLogistic()
    want = DLogistic._log_prob_torch(x, mean, logscale, 8)
    got = dist.log_prob(x, mean, logscale, 8)
    assert got.shape == want.shape
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)
    sums = dist.log_prob_sums(x, mean, logscale, 8)
    want_sums = want.double().flatten(1).sum(1)
    assert torch.allclose(sums.double(), want_sums, rtol=1e-5, atol=1e-4)
    # autograd still goes through the torch formula
    m2 = mean.clone().requires_grad_(True)
    dist.log_prob(x, m2, logscale, 8).sum().backward()
    assert m2.grad is not None and torch.isfinite(m2.grad).all()


def test_fused_log_likelihood_matches_reference_golden():
    """IDFlows.log_likelihood through the fused per-image reduction, fed the reference's own
    latents / means / logscales (tests/golden/flow_tiny.npz, produced by the reference's model on
    the CPU), against the log-likelihood the reference computed from them."""
    g = np.load(GOLDEN)
    model = build_tiny().cuda().eval()
    lat = [torch.from_numpy(g[f"id.latent{i}"]).cuda() for i in range(model.nsplit)]
    means = [torch.from_numpy(g[f"id.mean{i}"]).cuda() for i in range(model.nsplit)]
    logs = [torch.from_numpy(g[f"id.logscale{i}"]).cuda() for i in range(model.nsplit)]
    with torch.no_grad():
        ll, per_level = model.log_likelihood(lat, means, logs)
    assert np.allclose(ll.cpu().numpy(), g["id.logp"], rtol=2e-5, atol=2e-5)
    assert len(per_level) == model.nsplit


def test_fused_dlogistic_sample_matches_torch_formula():
    from flic_b200.distlib import DLogistic
    shape = (4, 6, 16, 16)
    g = torch.Generator(device="cuda").manual_seed(9)
    mean = (torch.rand(shape, device="cuda", generator=g) - 0.5) * 2
    logscale = (torch.rand(shape, device="cuda", generator=g) - 0.5) * 4 - 3
    dist = DLogistic()
    torch.manual_seed(123)
    got = dist.sample(mean, logscale, 8)
    torch.manual_seed(123)
    u = torch.rand_like(mean)
    want = dist.round(torch.log(u / (1 - u)) * torch.exp(logscale) + mean, nbits=8)
    assert torch.equal(got * 256, torch.round(got * 256))                 # on the 1/256 grid
    diff = (got - want).abs()
    assert float(diff.max()) <= 1 / 256 and float((diff != 0).float().mean()) < 1e-4


def test_chunk_pipelining_does_not_change_the_bytes():
    """N3: chunks run on rotating CUDA streams (pipeline > 1).  The container must be byte-identical
    to the in-line result and decode to the pixels whichever way either side was run."""
    model = build_tiny().cuda().eval()
    img = torch.randint(0, 256, (11, 3, 16, 16), dtype=torch.uint8, generator=torch.Generator().manual_seed(3)).cuda()
    blobs = [model.compress(img, codec_batch=3, pipeline=p).to_bytes() for p in (1, 2, 4)]
    assert blobs[0] == blobs[1] == blobs[2]
    for p in (1, 3):
        assert torch.equal(model.decompress(blobs[0], pipeline=p), img)


# ---- BASELINE.json configs[2] and configs[3]: the shapes and stream partitions of the patch-wise
# ---- models (network width reduced; the coder sees the same symbol counts and stream counts) ----

def _check_streams_against_oracle(oracle, model, batch, stats, n, sps):
    cbs = batch.codec_batch
    k = 0
    for ci, chunk in enumerate(batch.sections):
        n_real = min(cbs, n - ci * cbs)
        for level, enc in enumerate(chunk):
            z, mean, logscale = stats[k]
            k += 1
            nsym = n_real * int(np.prod(model.latents_shape[level]))
            scale = torch.exp(logscale.contiguous())
            off = model._segment_offsets(level, n_real, sps, "cpu").numpy()
            words, woff, states, status = oracle.encode_streams(
                z.reshape(-1)[:nsym].cpu().numpy(), mean.contiguous().reshape(-1)[:nsym].cpu().numpy(),
                scale.reshape(-1)[:nsym].cpu().numpy(), off, n_threads=8)
            assert not status.any()
            assert np.array_equal(enc.word_offsets.cpu().numpy(), woff)
            assert np.array_equal(enc.words.cpu().numpy().view(np.uint32)[:words.size], words)
            assert np.array_equal(enc.final_states.cpu().numpy().view(np.uint64), states)


def test_config_resflows_smallpatch_patchwise_coding(oracle):
    """configs/resflows_smallpatch.yaml:3-43,73-75: IDFlows nsplit 1 on 8x8 patches (latent
    (12,4,4) = 192 symbols), 27 x 23 = 621 patches per 216x184 image, batch 16 -> 9936 independent
    rANS streams per batch.  Lossless, and every stream identical to the reference coder's."""
    import random
    from flic_b200 import flows
    layer = dict(name="DenseLayer", act="ReLU")
    cfg = dict(name="IDFlows", nflows=12, nbits=8, nsplit=1, H=8, W=8, C=3,
               couple=dict(name="AdditiveCouple", split=0.75, round=dict(name="Round", nbits=8),
                           nn=dict(name="DenseBlock", growth_channel=16, depth=2, layer=layer)),
               extenddim=dict(name="ExtendDim", scale=2),
               prior=dict(name="Prior", round=dict(name="Round", nbits=8),
                          nn=dict(name="DenseBlock", growth_channel=16, depth=2, layer=layer)),
               distribution=dict(name="DLogistic"), round=dict(name="Round", nbits=8))
    torch.manual_seed(0)
    random.seed(0)
    model = flows.build_model(cfg)
    flows.perturb_heads(model, 0.02)
    model = model.cuda().eval()
    assert model.latents_shape == [(12, 4, 4)]
    images = torch.randint(0, 256, (16, 3, 216, 184), dtype=torch.uint8, generator=torch.Generator().manual_seed(8))
    # Patching (trainer.py residual path): non-overlapping 8x8 tiles, row-major
    patches = images.unfold(2, 8, 8).unfold(3, 8, 8).permute(0, 2, 3, 1, 4, 5).reshape(-1, 3, 8, 8).contiguous().cuda()
    assert patches.shape[0] == 16 * 621
    stats = []
    batch = model.compress(patches, stats=stats)
    assert batch.n_streams() == 9936
    _check_streams_against_oracle(oracle, model, batch, stats, patches.shape[0], 1)
    assert torch.equal(model.decompress(batch.to_bytes()), patches)


def test_config_vqvae_conditional_patches(oracle):
    """configs/resflow-patches-vqvae.yaml:3-44: ConditionalFlows on 27x23 patches, ExtendDim scale 1,
    nsplit 1 (latent (3,27,23) = 1863 symbols, an odd stream length), 64 patches per image, the prior
    conditioned on a (here random) reconstruction (flows.py:303-327)."""
    import random
    from flic_b200 import flows
    layer = dict(name="DenseLayer", act="LeakyReLU")
    cfg = dict(name="ConditionalFlows", nflows=8, nbits=8, nsplit=1, H=27, W=23, C=3, conv_for_cond=False,
               couple=dict(name="AdditiveCouple", split=0.75, round=dict(name="Round", nbits=8),
                           nn=dict(name="DenseBlock", growth_channel=16, depth=2, layer=layer)),
               extenddim=dict(name="ExtendDim", scale=1),
               prior=dict(name="Prior", round=dict(name="Round", nbits=8),
                          nn=dict(name="DenseBlock", growth_channel=16, depth=2, layer=layer)),
               distribution=dict(name="DLogistic"), round=dict(name="Round", nbits=8))
    torch.manual_seed(0)
    random.seed(0)
    model = flows.build_model(cfg)
    flows.perturb_heads(model, 0.02)
    model = model.cuda().eval()
    assert model.latents_shape == [(3, 27, 23)]
    g = torch.Generator().manual_seed(4)
    n = 2 * 64
    patches = torch.randint(0, 256, (n, 3, 27, 23), dtype=torch.uint8, generator=g).cuda()
    cond = flows.u8_to_grid(torch.randint(0, 256, (n, 3, 27, 23), dtype=torch.uint8, generator=g).cuda())
    stats = []
    batch = model.compress(patches, cond=cond, codec_batch=48, stats=stats)      # 48 does not divide 128: padded last chunk
    assert batch.n_streams() == n
    _check_streams_against_oracle(oracle, model, batch, stats, n, 1)
    assert torch.equal(model.decompress(batch.to_bytes(), cond=cond), patches)
    # the residual path of the reference's ResidualTrainer (trainer.py:606-621): what is coded is
    # image - reconstruction, a grid-valued float tensor with negative entries, under the same cond
    resid = flows.u8_to_grid(patches) - cond
    assert float(resid.min()) < 0
    blob = model.compress_grid(resid, cond=cond, codec_batch=48).to_bytes()
    assert torch.equal(model.decompress_grid(blob, cond=cond), resid)


def test_twolevel_flows_compress_decompress(oracle):
    """configs/config_twolevel.yaml shapes (215 x 178 images padded to 216 x 184, rough 27 x 23,
    621 fine 8 x 8 patches per image; narrow networks): bytes -> pixels round trip, and the
    reference-style forward reports finite ideal bits per dimension."""
    from test_host_logic import twolevel_cfg
    import random
    from flic_b200 import flows
    torch.manual_seed(0)
    random.seed(0)
    model = flows.build_model(twolevel_cfg())
    flows.perturb_heads(model, 0.02)
    model = model.cuda().eval()
    img = torch.randint(0, 256, (3, 3, 215, 178), dtype=torch.uint8, generator=torch.Generator().manual_seed(2)).cuda()
    blob = model.compress(img)
    assert torch.equal(model.decompress(blob), img)
    lat, means, logs, bpd, bpd1, bpd2, _ = model.forward(flows.u8_to_grid(img))
    assert lat[0].shape == (3, 3, 27, 23) and lat[1].shape == (3 * 621, 12, 4, 4)
    assert np.isfinite(bpd) and bpd > 0
    real_bpd = 8 * len(blob) / img.numel()
    assert real_bpd >= bpd - 1e-3 and real_bpd < bpd * 1.02 + 64 * (3 + 3 * 621) / img.numel() + 0.05
    with pytest.raises(ValueError):
        model.decompress(b"XXXX" + blob[4:])


def test_reference_import_lines_resolve_to_the_cuda_coder(oracle):
    """compat/ on PYTHONPATH makes `from rans.rans import encode, decode` (trainer.py:32, coder.py:15)
    and `from rans import encode, decode` (rans/test.py:1) resolve, unmodified, to this repository's
    coder.  The reference's two self-tests (rans/test.py:6-36 with n patched down, coder.py:41-73),
    restated in tests/ref_style_selftest.py, run as a FILE in a fresh interpreter: no errors, final
    state 1<<32, and state / word count / SHA-256 of the words equal what the reference's own coder
    produced for the same lists (tests/golden/kat_random.json, made by tests/golden/make_golden.py)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = {**os.environ, "PYTHONPATH": os.path.join(root, "compat")}
    out = subprocess.run([sys.executable, os.path.join(root, "tests", "ref_style_selftest.py"), "200000", "100000"],
                         capture_output=True, text=True, env=env, cwd=root, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("rans_test_py", "coder_py_main"):
        r = res[key]
        assert r["errors"] == 0 and r["final_state"] == 1 << 32
        assert r["list_types"] == ["list", "list", "int", "float"]
    # the lists are built exactly as tests/golden/make_golden.py built them for the reference's coder
    kat = json.load(open(os.path.join(root, "tests", "golden", "kat_random.json")))
    for key, gold in (("rans_test_py", kat["200000"]), ("coder_py_main", kat["coder_100000"])):
        assert res[key]["state"] == gold["state"]
        assert res[key]["words"] == gold["n_words"] and res[key]["sha256"] == gold["sha256"]


def test_full_width_imagenet64_model_round_trip():
    """configs/imagenet64.yaml at its real width (growth 512, depth 12, 8 flows x 3 levels, 60 M
    parameters) on one codec batch of 64 images: compress -> bytes -> decompress must return the
    pixels, i.e. the convolutions of the two passes (cuDNN, deterministic, TF32 off) agree bit for
    bit on every one of the 786 432 latents' parameters.  (The reduced-width models of the other
    tests exercise the same code with far fewer accumulations per output.)"""
    import random
    from flic_b200 import flows
    layer = dict(name="DenseLayer", act="ReLU")
    block = dict(name="DenseBlock", growth_channel=512, depth=12, layer=layer)
    cfg = dict(name="IDFlows", nflows=8, nbits=8, nsplit=3, H=64, W=64, C=3,
               couple=dict(name="AdditiveCouple", split=0.75, nn=block, round=dict(name="Round", nbits=8)),
               extenddim=dict(name="ExtendDim", scale=2),
               prior=dict(name="Prior", round=dict(name="Round", nbits=8), nn=block),
               distribution=dict(name="DLogistic"), round=dict(name="Round", nbits=8))
    torch.manual_seed(0)
    random.seed(0)
    model = flows.build_model(cfg)
    flows.perturb_heads(model, 0.02)
    model = model.cuda().eval()
    img = torch.randint(0, 256, (64 + 7, 3, 64, 64), dtype=torch.uint8, generator=torch.Generator().manual_seed(5)).cuda()
    batch = model.compress(img, codec_batch=64)          # a full chunk and a padded one
    assert batch.chained and batch.n_streams() == 71
    blob = batch.to_bytes()
    assert torch.equal(model.decompress(blob), img)
    assert 9.5 < 8 * len(blob) / img.numel() < 10.8      # random-init model on uniform noise: ~10.1 bits/dim
