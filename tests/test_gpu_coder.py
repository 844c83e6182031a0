"""GPU parity tests of the coder kernels, through the C ABI, against the CPU oracle.

Bit-exact bar (BASELINE.json): frequency tables identical to the reference's quantisation;
each stream's (final_state, words) identical to the reference coder run on the same slice;
decoded symbols identical to the input.
"""
import numpy as np
import pytest
import torch

from _data import gen, ragged_offsets

pytestmark = pytest.mark.gpu

KINDS = ["test", "coder", "wide", "edges", "needle"]


def _cuda(*arrs):
    return [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrs]


def _u32(t):
    return t.cpu().numpy().view(np.uint32)


def _u64(t):
    return t.cpu().numpy().view(np.uint64)


@pytest.mark.parametrize("kind", KINDS)
def test_cdf_tables_bit_exact(oracle, kind):
    from flic_b200 import rans
    n = 1 << 20
    x, mean, scale = gen(kind, n, 11)
    _, st_o, fr_o = oracle.tables(x, mean, scale)
    xd, md, sd = _cuda(x, mean, scale)
    start, freq, status = rans.cdf_tables(xd, md, sd)
    assert int(status.item()) == 0
    assert np.array_equal(_u32(start), st_o.astype(np.uint32))
    assert np.array_equal(_u32(freq), fr_o.astype(np.uint32))
    assert int(_u32(freq).min()) >= 1


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("n,n_streams", [(200_000, 1), (200_000, 37), (300_000, 1000), (50_000, 4097)])
def test_streams_bit_exact_and_round_trip(oracle, kind, n, n_streams):
    from flic_b200 import rans
    x, mean, scale = gen(kind, n, 5 + n_streams)
    off = ragged_offsets(n, n_streams, 99 + n_streams)
    words_o, woff_o, states_o, status_o = oracle.encode_streams(x, mean, scale, off, n_threads=8)
    assert not status_o.any()
    xd, md, sd = _cuda(x, mean, scale)
    offd = torch.from_numpy(off).cuda()
    enc = rans.encode_streams(xd, md, sd, offd)
    assert not enc.status.any().item()
    assert np.array_equal(enc.word_offsets.cpu().numpy(), woff_o)
    assert np.array_equal(_u64(enc.final_states), states_o)
    assert np.array_equal(_u32(enc.words), words_o)
    # the reference's accounting: 64 bits per stream + 32 per word (trainer.py:326-327)
    assert enc.bits() == 64 * n_streams + 32 * words_o.size
    xr, end_states, status = rans.decode_streams(enc, md, sd, offd)
    assert not status.any().item()
    assert torch.equal(xr, xd)
    assert bool((end_states == (1 << 32)).all().item())


def test_uniform_partition_imagenet64_shape(oracle):
    """Per image x level streams of the imagenet64 model: 6144 / 3072 / 3072 symbols (App. B)."""
    from flic_b200 import rans
    imgs = 24
    for seg in (6144, 3072):
        n = imgs * seg
        x, mean, scale = gen("test", n, seg)
        off = np.arange(imgs + 1, dtype=np.int64) * seg
        words_o, woff_o, states_o, _ = oracle.encode_streams(x, mean, scale, off, n_threads=8)
        xd, md, sd = _cuda(x, mean, scale)
        enc = rans.encode_streams(xd, md, sd, rans.uniform_offsets(imgs, seg, "cuda"))
        assert np.array_equal(_u32(enc.words), words_o)
        assert np.array_equal(_u64(enc.final_states), states_o)
        xr, end, status = rans.decode_streams(enc, md, sd, torch.from_numpy(off).cuda())
        assert torch.equal(xr, xd) and not status.any().item()


def test_many_short_streams_exercise_the_large_grid_kernels(oracle):
    """Above ~28 k streams the encoder switches to 4 producer warps per CTA and above ~76 k the
    decoder to 4 warps per CTA; both variants must be bit-exact too (ragged, some empty streams)."""
    from flic_b200 import rans
    n, n_streams = 1_500_000, 90_000
    x, mean, scale = gen("test", n, 77)
    off = ragged_offsets(n, n_streams, 78)
    words_o, woff_o, states_o, _ = oracle.encode_streams(x, mean, scale, off, n_threads=8)
    xd, md, sd = _cuda(x, mean, scale)
    offd = torch.from_numpy(off).cuda()
    enc = rans.encode_streams(xd, md, sd, offd)
    assert np.array_equal(enc.word_offsets.cpu().numpy(), woff_o)
    assert np.array_equal(_u64(enc.final_states), states_o)
    assert np.array_equal(_u32(enc.words), words_o)
    xr, end, status = rans.decode_streams(enc, md, sd, offd)
    assert torch.equal(xr, xd) and not status.any().item() and bool((end == (1 << 32)).all().item())


def test_empty_and_tiny_inputs():
    from flic_b200 import rans
    e = torch.empty(0, dtype=torch.float32, device="cuda")
    enc = rans.encode_streams(e, e, e, torch.zeros(1, dtype=torch.int64, device="cuda"))
    assert enc.n_streams == 0 and enc.n_words() == 0
    # streams that are all empty
    enc = rans.encode_streams(e, e, e, torch.zeros(4, dtype=torch.int64, device="cuda"))
    assert enc.n_streams == 3 and enc.n_words() == 0
    assert bool((enc.final_states == (1 << 32)).all().item())
    xr, end, status = rans.decode_streams(enc, e, e, torch.zeros(4, dtype=torch.int64, device="cuda"))
    assert xr.numel() == 0 and not status.any().item()
    # one symbol
    x = torch.tensor([0.5], device="cuda")
    m = torch.tensor([0.0], device="cuda")
    s = torch.tensor([1.0], device="cuda")
    enc = rans.encode_streams(x, m, s)
    xr, end, status = rans.decode_streams(enc, m, s)
    assert torch.equal(xr, x) and int(end.item()) == 1 << 32


def test_kat8_known_answer():
    """KAT-8 from the reference build (tests/golden/kat8.json)."""
    import json
    import os
    from flic_b200 import rans
    kat = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat8.json")))
    x, m, s = (torch.tensor(kat[k], dtype=torch.float32, device="cuda") for k in ("x", "mean", "scale"))
    start, freq, status = rans.cdf_tables(x, m, s)
    assert _u32(start).tolist() == kat["start"] and _u32(freq).tolist() == kat["freq"]
    enc = rans.encode_streams(x, m, s)
    assert int(_u64(enc.final_states)[0]) == kat["state"]
    assert _u32(enc.words).tolist() == kat["buf"]


def test_status_zero_scale_and_out_of_window():
    from flic_b200 import rans, _lib
    x = torch.tensor([0.5, 0.25, 100.0, 0.0], device="cuda")
    m = torch.zeros(4, device="cuda")
    s = torch.tensor([1.0, 0.0, 1.0, 1.0], device="cuda")
    off = torch.tensor([0, 1, 2, 3, 4], device="cuda")
    enc = rans.encode_streams(x, m, s, off)
    st = enc.status.cpu().tolist()
    assert st[0] == 0 and st[3] == 0
    assert st[1] & _lib.ST_ZERO_SCALE
    assert st[2] & _lib.ST_OUT_OF_WINDOW
    with pytest.raises(ZeroDivisionError):
        enc.check()
    # off-grid symbol
    enc = rans.encode_streams(torch.tensor([0.3], device="cuda"), m[:1], s[:1])
    assert enc.status.cpu().tolist()[0] & _lib.ST_OUT_OF_WINDOW


def test_truncated_stream_is_reported():
    from flic_b200 import rans, _lib
    x, mean, scale = gen("test", 5000, 3)
    xd, md, sd = _cuda(x, mean, scale)
    enc = rans.encode_streams(xd, md, sd)
    nw = enc.n_words()
    cut = rans.EncodedStreams(enc.words[: nw - 3].clone(), torch.tensor([0, nw - 3], device="cuda"),
                              enc.final_states, enc.status, enc.n_symbols)
    xr, end, status = rans.decode_streams(cut, md, sd)
    assert status.cpu().tolist()[0] & (_lib.ST_UNDERRUN | _lib.ST_BAD_END_STATE | _lib.ST_NO_SYMBOL)


def test_init_states_chain_like_coder_py(oracle):
    """coder.Encode chains the state across levels (coder.py:25): encode(state_in) must accept any
    state the previous call returned."""
    from flic_b200 import rans
    x, mean, scale = gen("coder", 4000, 8)
    st1, buf1 = oracle.encode(1 << 32, 2000, x[:2000], mean[:2000], scale[:2000])
    st2, buf2 = oracle.encode(st1, 2000, x[2000:], mean[2000:], scale[2000:])
    xd, md, sd = _cuda(x[2000:], mean[2000:], scale[2000:])
    init = torch.tensor([np.uint64(st1).astype(np.int64)], device="cuda")
    enc = rans.encode_streams(xd, md, sd, None, init_states=init)
    assert int(_u64(enc.final_states)[0]) == st2
    assert np.array_equal(_u32(enc.words), buf2)


def test_full_size_round_trip_properties():
    """BASELINE config 5 chunk size (2048 images x 12288 symbols): size-independent properties --
    decode(encode(x)) == x, every stream ends at 1<<32, bits match the word count."""
    from flic_b200 import rans
    imgs, per = 2048, 12288
    n = imgs * per
    g = torch.Generator(device="cuda").manual_seed(1)
    mean = torch.randint(-256, 257, (n,), device="cuda", generator=g).float() / 256
    scale = torch.exp(10 * torch.rand(n, device="cuda", generator=g) - 5) / 256
    x = torch.round((mean.double() + scale.double() * (10 * torch.rand(n, device="cuda", generator=g).double() - 5)) * 256) / 256
    x = x.float()
    # per image x level segments 6144 / 3072 / 3072
    seg = torch.tensor([6144, 3072, 3072], device="cuda").repeat(imgs)
    off = torch.cat([torch.zeros(1, dtype=torch.int64, device="cuda"), torch.cumsum(seg, 0)])
    enc = rans.encode_streams(x, mean, scale, off)
    assert not enc.status.any().item()
    xr, end, status = rans.decode_streams(enc, mean, scale, off)
    assert not status.any().item()
    assert torch.equal(xr, x)
    assert bool((end == (1 << 32)).all().item())
    bps = enc.bits() / n
    assert 4.2 < bps < 4.7  # rans/test.py distribution codes at ~4.4 bits/symbol


@pytest.mark.parametrize("pads", [(0, 0, 0), (3, 3, 3), (5, 5, 5), (1, 2, 3), (0, 4, 0)])
def test_decode_on_any_array_alignment(oracle, pads):
    """The decoder stages 32-byte blocks per lane when mean, scale and the output share the same
    32-byte phase (any phase), and falls back to warp-wide 4-byte staging when they do not.  Both
    must return the symbols for ragged partitions whose streams start anywhere."""
    from flic_b200 import rans
    n, ns = 150_000, 333
    x, mean, scale = gen("test", n, 77)
    off = ragged_offsets(n, ns, 13)
    xd, md, sd = _cuda(x, mean, scale)
    offd = torch.from_numpy(off).cuda()
    enc = rans.encode_streams(xd, md, sd, offd)
    pm, ps, px = pads
    mbuf = torch.zeros(n + 8, device="cuda"); mbuf[pm:pm + n] = md
    sbuf = torch.ones(n + 8, device="cuda"); sbuf[ps:ps + n] = sd
    obuf = torch.full((n + 8,), -7.0, device="cuda")
    xr, end_states, status = rans.decode_streams(enc, mbuf[pm:pm + n], sbuf[ps:ps + n], offd, out=obuf[px:px + n])
    assert not status.any().item()
    assert torch.equal(obuf[px:px + n], xd)
    assert bool((obuf[:px] == -7.0).all().item()) and bool((obuf[px + n:] == -7.0).all().item())   # nothing outside
    assert bool((end_states == (1 << 32)).all().item())


@pytest.mark.parametrize("n_streams", [130_000, 60_000])
@pytest.mark.parametrize("pads", [(0, 0, 0), (5, 5, 5), (1, 2, 3)])
def test_lane_per_stream_encoder_is_bit_exact(oracle, pads, n_streams):
    """From 12 warps of streams per SM (56 832 streams on a 148-SM B200) the encoder runs one lane
    per stream, each lane staging 32-byte blocks of its own stream.  130 000 / 60 000 ragged streams
    (many shorter than a block, some empty) on arrays at phase 0, at a common odd phase, and at
    different phases (which falls back to the warp-cooperative kernel): every stream's (state, words)
    must be the reference coder's, and the decoder must return the symbols."""
    from flic_b200 import rans, _lib
    n = 2_400_000
    x, mean, scale = gen("test", n, 91)
    off = ragged_offsets(n, n_streams, 92)
    words_o, woff_o, states_o, _ = oracle.encode_streams(x, mean, scale, off, n_threads=8)
    bufs = []
    for arr, pad, fill in ((x, pads[0], 0.0), (mean, pads[1], 0.0), (scale, pads[2], 1.0)):
        b = torch.full((n + 8,), fill, device="cuda")
        b[pad:pad + n] = torch.from_numpy(arr).cuda()
        bufs.append(b[pad:pad + n])
    xd, md, sd = bufs
    offd = torch.from_numpy(off).cuda()
    enc = rans.encode_streams(xd, md, sd, offd)
    kernel = _lib.lib().flic_last_coder_kernel(0).decode()
    same_phase = len(set(pads)) == 1
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    if same_phase and (n_streams + 31) // 32 >= 12 * sms:
        assert kernel == "rans_encode_lane_kernel"
    if not same_phase:
        assert kernel == "rans_encode_kernel"
    assert not enc.status.any().item()
    assert np.array_equal(enc.word_offsets.cpu().numpy(), woff_o)
    assert np.array_equal(_u64(enc.final_states), states_o)
    assert np.array_equal(_u32(enc.words)[:words_o.size], words_o)
    xr, end, status = rans.decode_streams(enc, md, sd, offd)
    assert torch.equal(xr, xd) and not status.any().item() and bool((end == (1 << 32)).all().item())


def test_lane_per_stream_encoder_takes_initial_states(oracle):
    """init_states on the lane-per-stream kernel: code 130 000 streams once, then code them again
    starting from the states the first pass ended in (coder.py:25 chains states like that), and
    compare a sample of streams with the reference coder called with the same initial state."""
    from flic_b200 import rans, _lib
    n, n_streams = 2_400_000, 130_000
    x, mean, scale = gen("coder", n, 17)
    off = ragged_offsets(n, n_streams, 18)
    xd, md, sd = _cuda(x, mean, scale)
    offd = torch.from_numpy(off).cuda()
    first = rans.encode_streams(xd, md, sd, offd)
    second = rans.encode_streams(xd, md, sd, offd, init_states=first.final_states)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    if (n_streams + 31) // 32 >= 12 * sms:
        assert _lib.lib().flic_last_coder_kernel(0).decode() == "rans_encode_lane_kernel"
    st1 = _u64(first.final_states)
    st2 = _u64(second.final_states)
    woff = second.word_offsets.cpu().numpy()
    words = _u32(second.words)
    rng = np.random.default_rng(3)
    for s in rng.choice(n_streams, 300, replace=False):
        a, b = int(off[s]), int(off[s + 1])
        state, buf = oracle.encode(int(st1[s]), b - a, x[a:b], mean[a:b], scale[a:b])
        assert int(st2[s]) == state
        assert np.array_equal(words[woff[s]:woff[s + 1]], buf)


def test_reciprocal_divisions_are_exact_on_the_device():
    """The coder never divides: t4 / scale is a cubic-step reciprocal plus one exact-remainder
    correction, state / freq a low-biased reciprocal plus one integer correction (DESIGN.md
    section 2).  Both are checked on the GPU against the hardware's exact divisions over 2^31
    generated operand pairs each (realistic and adversarial ranges)."""
    import ctypes as C
    from flic_b200 import _lib
    L = _lib.lib()
    n = 1 << 31
    bad = torch.zeros(3, dtype=torch.int64, device="cuda")
    _lib.check(L.flic_debug_div_check(n, C.c_uint64(1), 0, bad[0:1].data_ptr(), None))
    _lib.check(L.flic_debug_div_check(n, C.c_uint64(2), 1, bad[1:2].data_ptr(), None))
    _lib.check(L.flic_debug_push_check(n, C.c_uint64(3), bad[2:3].data_ptr(), None))
    assert bad.tolist() == [0, 0, 0]


# ---- the CTA-per-stream decoder (few streams) ----------------------------------------------------

def _with_decode_kernel(which):
    """Context manager: force the lane-per-stream (0), the CTA-per-stream (1) or a cluster-per-stream
    (2, 4, 8 CTAs) decode kernel."""
    import contextlib
    from flic_b200 import _lib

    @contextlib.contextmanager
    def cm():
        old = _lib.lib().flic_set_decode_kernel(which)
        try:
            yield
        finally:
            _lib.lib().flic_set_decode_kernel(old)
    return cm()


def _narrow(n, seed):
    """Scales from 0.007 to 20 bins: every class of the CTA-per-stream decoder's windows (one chunk,
    two to four chunks, none) in one stream, symbols up to five scales from the mode."""
    rng = np.random.default_rng(seed)
    mean = rng.normal(0, 0.7, n).astype(np.float32)
    scale = (np.exp(rng.uniform(-5, 3, n)) / 256).astype(np.float32)
    x = np.round((mean.astype(np.float64) + scale.astype(np.float64) * (10 * rng.random(n) - 5)) * 256) / 256
    return x.astype(np.float32), mean, scale


@pytest.mark.parametrize("kind", KINDS + ["narrow"])
@pytest.mark.parametrize("n_streams", [1, 3, 48, 300])
def test_cta_per_stream_decoder_is_bit_exact(oracle, kind, n_streams):
    """Reference-native partitions (rans/test.py: 1 stream; trainer.py:308: 3 streams; configs[0]: 48)
    and a full wave of CTAs: the CTA-per-stream decoder must return the symbols, end every stream at
    1<<32, and leave exactly what the lane-per-stream kernel leaves (which the oracle pins)."""
    from flic_b200 import rans, _lib
    n = 60_000 if n_streams < 100 else 150_000
    x, mean, scale = _narrow(n, 5) if kind == "narrow" else gen(kind, n, 21 + n_streams)
    off = ragged_offsets(n, n_streams, 7 + n_streams)
    words_o, woff_o, states_o, _ = oracle.encode_streams(x, mean, scale, off, n_threads=8)
    xd, md, sd = _cuda(x, mean, scale)
    offd = torch.from_numpy(off).cuda()
    enc = rans.encode_streams(xd, md, sd, offd)
    assert np.array_equal(_u32(enc.words), words_o) and np.array_equal(_u64(enc.final_states), states_o)
    with _with_decode_kernel(1):
        xr, end, status = rans.decode_streams(enc, md, sd, offd)
        assert _lib.lib().flic_last_coder_kernel(1).decode() == "rans_decode_coop_kernel"
    assert not status.any().item()
    assert torch.equal(xr, xd)
    assert bool((end == (1 << 32)).all().item())
    with _with_decode_kernel(0):
        xr0, end0, status0 = rans.decode_streams(enc, md, sd, offd)
        assert _lib.lib().flic_last_coder_kernel(1).decode() != "rans_decode_coop_kernel"
    assert torch.equal(xr0, xr) and torch.equal(end0, end) and torch.equal(status0, status)
    # a thread-block cluster per stream (the other CTAs tabulate into the home CTA's shared memory)
    for cluster in (2, 4, 8):
        with _with_decode_kernel(cluster):
            xc, endc, statusc = rans.decode_streams(enc, md, sd, offd)
            assert _lib.lib().flic_last_coder_kernel(1).decode() == f"rans_decode_coop_kernel<cluster {cluster}>"
        assert torch.equal(xc, xr) and torch.equal(endc, end) and torch.equal(statusc, status), cluster


def test_cluster_per_stream_decoder_on_wide_and_mixed_windows(oracle):
    """Every window class of the cluster decoder in one stream -- one chunk, an index of 2 to 32 chunks,
    the 32-chunk cap (scales beyond 100 bins), symbols outside their window (up to 8 scale units from
    the mode), a group that overflows the slot memory (32 wide symbols in a row) -- against the oracle's
    bitstream, and the automatic choice for a handful of streams."""
    from flic_b200 import rans, _lib
    rng = np.random.default_rng(77)
    n = 40_000
    mean = rng.normal(0, 0.7, n).astype(np.float32)
    logc = rng.uniform(-5, 5.2, n)
    logc[10_000:10_640] = rng.uniform(4.0, 5.2, 640)          # twenty groups of wide symbols only
    scale = (np.exp(logc) / 256).astype(np.float32)
    u = rng.uniform(-8, 8, n)
    x = np.round((mean.astype(np.float64) + scale.astype(np.float64) * u) * 256) / 256
    lo = np.round(mean.astype(np.float64) * 256 - 1024)
    x = (np.clip(x * 256, lo + 1, lo + 2046) / 256).astype(np.float32)
    for n_streams in (1, 3, 5):
        off = ragged_offsets(n, n_streams, 3 + n_streams)
        words_o, woff_o, states_o, _ = oracle.encode_streams(x, mean, scale, off, n_threads=8)
        xd, md, sd = _cuda(x, mean, scale)
        offd = torch.from_numpy(off).cuda()
        enc = rans.encode_streams(xd, md, sd, offd)
        assert np.array_equal(_u32(enc.words), words_o) and np.array_equal(_u64(enc.final_states), states_o)
        xa, enda, sta = rans.decode_streams(enc, md, sd, offd)              # automatic choice
        assert _lib.lib().flic_last_coder_kernel(1).decode().startswith("rans_decode_coop_kernel<cluster")
        assert torch.equal(xa, xd) and not sta.any().item() and bool((enda == (1 << 32)).all().item())
        for which in (0, 1, 2, 4, 8):
            with _with_decode_kernel(which):
                xr, end, st = rans.decode_streams(enc, md, sd, offd)
            assert torch.equal(xr, xd) and torch.equal(end, enda) and torch.equal(st, sta), which


def test_cta_per_stream_decoder_reports_what_the_lane_kernel_reports():
    """Truncated buffers, zero / negative / non-finite scales and absurd means: same status bits from
    both decode kernels, no crash, no out-of-bounds access (run under compute-sanitizer by hand)."""
    from flic_b200 import rans, _lib
    x, mean, scale = gen("test", 20_000, 3)
    xd, md, sd = _cuda(x, mean, scale)
    off = torch.tensor([0, 5000, 5000, 12_345, 20_000], device="cuda")
    enc = rans.encode_streams(xd, md, sd, off)
    woff = enc.word_offsets.clone()
    # stream 0: three words short; stream 2: parameters damaged
    bad_m, bad_s = md.clone(), sd.clone()
    bad_s[6000] = 0.0
    bad_s[7000] = -1.0
    bad_s[8000] = float("nan")
    bad_m[9000] = 1.0e9
    words = enc.words.clone()
    cut = rans.EncodedStreams(words, woff, enc.final_states, enc.status, enc.n_symbols)
    cut.word_offsets = woff.clone()
    res = []
    for which in (0, 1, 2, 8):
        with _with_decode_kernel(which):
            # shorten stream 0 by pretending its words end three early (the words stay in place)
            wo = woff.clone()
            e = rans.EncodedStreams(torch.cat([words[:int(woff[1]) - 3], words[int(woff[1]):]]),
                                    torch.cat([wo[:1], wo[1:] - 3]), enc.final_states, enc.status, enc.n_symbols)
            xr, end, status = rans.decode_streams(e, bad_m, bad_s, off)
            res.append(status.cpu().tolist())
    assert res[0] == res[1] == res[2] == res[3]
    st = res[1]
    assert st[0] & (_lib.ST_UNDERRUN | _lib.ST_BAD_END_STATE | _lib.ST_NO_SYMBOL)
    assert st[1] == 0                      # the empty stream
    assert st[2] & _lib.ST_ZERO_SCALE and st[2] & _lib.ST_NONFINITE
    assert st[3] == 0


# ---- chained streams: decode continuation ----------------------------------------------------------

@pytest.mark.parametrize("n_streams", [1, 40, 5000])
@pytest.mark.parametrize("kind", ["test", "edges"])
def test_chained_levels_equal_one_stream_over_the_concatenation(oracle, kind, n_streams):
    """coder.Encode (coder.py:18-27) carries the state from one level into the next.  Coding three
    'levels' of every stream that way (init_states) and concatenating the words must give exactly
    the stream that codes the three segments back to back in one go, and decoding it level by level,
    last level first, each call continuing from the state and the unread words the previous one
    stopped at (flic_rans_decode_resume), must return every level -- including when a level's first
    symbol pushed a word, the case the reference's per-level buffers get wrong."""
    from flic_b200 import rans
    lens = [37, 64, 19]
    n_levels = len(lens)
    data = [gen(kind, n_streams * L, 100 + i) for i, L in enumerate(lens)]
    dev = [_cuda(*d) for d in data]
    offs = [torch.arange(n_streams + 1, device="cuda", dtype=torch.int64) * L for L in lens]
    levels, carried = [], None
    for (xd, md, sd), off in zip(dev, offs):
        e = rans.encode_streams(xd, md, sd, off, init_states=carried, workspace=rans.Workspace(), own_output=False)
        carried = e.final_states
        levels.append(e)
    chained = rans.chain_levels(levels)
    assert not chained.status.any().item()
    # the same symbols as one stream per image: [level 0 | level 1 | level 2] per stream
    cat = [np.concatenate([d[k].reshape(n_streams, L) for d, L in zip(data, lens)], axis=1).reshape(-1) for k in range(3)]
    total = sum(lens)
    off_cat = np.arange(n_streams + 1, dtype=np.int64) * total
    words_o, woff_o, states_o, _ = oracle.encode_streams(cat[0], cat[1], cat[2], off_cat, n_threads=8)
    nw = chained.n_words()
    assert np.array_equal(chained.word_offsets.cpu().numpy(), woff_o)
    assert np.array_equal(_u32(chained.words)[:nw], words_o)
    assert np.array_equal(_u64(chained.final_states), states_o)
    # decode with continuation, both kernels
    for which in (0, 1, 4):
        with _with_decode_kernel(which):
            st, left = None, None
            for lvl in reversed(range(n_levels)):
                xd, md, sd = dev[lvl]
                xr, st, status, left = rans.decode_streams(chained, md, sd, offs[lvl], check_end=lvl == 0, states=st,
                                                           words_left=left, return_words_left=True)
                assert not status.any().item()
                assert torch.equal(xr, xd)
            assert bool((st == (1 << 32)).all().item()) and not left.any().item()


@pytest.mark.parametrize("kind", KINDS)
def test_oracle_equals_the_reference_build_on_this_box(oracle, kind):
    """The checker itself, on the box whose libm defines expf for the run: the C restatement
    (oracle/liboracle.so) against the reference's own rans.pyx rebuilt into oracle/_ref (which
    travels here with the repository), same inputs, bit for bit -- the comparison
    tests/test_oracle_pinning.py makes in the build container, repeated where the GPU results are
    judged against the oracle."""
    ref = oracle.ref_rans()
    if ref is None:
        pytest.skip("oracle/_ref was not shipped; the committed goldens pin the oracle instead")
    n = 60_000
    x, mean, scale = gen(kind, n, 21)
    xl, ml, sl = x.tolist(), mean.tolist(), scale.tolist()
    state_r, buf_r = ref.encode(1 << 32, n, xl, ml, sl)
    state_o, buf_o = oracle.encode(1 << 32, n, x, mean, scale)
    assert state_r == state_o and buf_r == buf_o.tolist()
    end_r, msg_r = ref.decode(state_r, buf_r[::-1], n, ml[::-1], sl[::-1])
    assert end_r == 1 << 32 and msg_r[::-1] == xl
    # and the CUDA coder against the reference build directly
    from flic_b200 import rans
    xd, md, sd = _cuda(x, mean, scale)
    enc = rans.encode_streams(xd, md, sd)
    assert int(_u64(enc.final_states)[0]) == state_r and _u32(enc.words).tolist() == buf_r


def test_tables_and_streams_on_window_origin_ties(oracle):
    """The window origin is round(256 mean - 1024) with ties away from zero (rans.pyx:51,92); the kernels
    compute it on the float pipe.  Means exactly on a tie, one float either side of it, on both sides of
    mean = 4 (where 256 mean - 1024 changes sign): tables, bitstreams and the round trip against the oracle."""
    from flic_b200 import rans
    ks = np.arange(-1500, 1500, dtype=np.float64)
    ties = ((ks + 0.5) / 256.0).astype(np.float32)
    mean = np.concatenate([ties, np.nextafter(ties, np.float32(np.inf)), np.nextafter(ties, np.float32(-np.inf)),
                           (ks / 256.0).astype(np.float32)]).astype(np.float32)
    rng = np.random.default_rng(9)
    mean = np.tile(mean, 4)
    n = mean.size
    scale = (np.exp(rng.uniform(-4, 2, n)) / 256).astype(np.float32)
    lo = np.array([oracle.lib().flic_oracle_lower(__import__("ctypes").c_float(float(m))) for m in mean[: n // 4]])
    lo = np.tile(lo, 4)
    x = ((lo + rng.integers(900, 1150, n)) / 256.0).astype(np.float32)          # inside the window, near the mode
    _, st_o, fr_o = oracle.tables(x, mean, scale)
    xd, md, sd = _cuda(x, mean, scale)
    start, freq, status = rans.cdf_tables(xd, md, sd)
    assert int(status.item()) == 0
    assert np.array_equal(_u32(start), st_o.astype(np.uint32)) and np.array_equal(_u32(freq), fr_o.astype(np.uint32))
    off = ragged_offsets(n, 50, 4)
    words_o, woff_o, states_o, _ = oracle.encode_streams(x, mean, scale, off, n_threads=8)
    offd = torch.from_numpy(off).cuda()
    enc = rans.encode_streams(xd, md, sd, offd)
    assert np.array_equal(_u32(enc.words), words_o) and np.array_equal(_u64(enc.final_states), states_o)
    for which in (0, 1, 4):
        with _with_decode_kernel(which):
            xr, end, st = rans.decode_streams(enc, md, sd, offd)
            assert torch.equal(xr, xd) and not st.any().item() and bool((end == (1 << 32)).all().item())
