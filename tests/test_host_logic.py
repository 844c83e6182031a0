"""Host-side logic that needs no GPU: model mirror vs the reference's golden state, shapes,
stream partitions, container format, drop-in argument checking, sharding arithmetic."""
import os
import random

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "flow_tiny.npz")


def tiny_cfg(name="IDFlows", **extra):
    layer = dict(name="DenseLayer", act="LeakyReLU")
    cfg = dict(name=name, nflows=2, nbits=8, nsplit=2, H=16, W=16, C=3,
               couple=dict(name="AdditiveCouple", split=0.75, round=dict(name="Round", nbits=8),
                           nn=dict(name="DenseBlock", growth_channel=16, depth=2, layer=layer)),
               extenddim=dict(name="ExtendDim", scale=2),
               prior=dict(name="Prior", round=dict(name="Round", nbits=8),
                          nn=dict(name="DenseBlock", growth_channel=16, depth=2, layer=layer)),
               distribution=dict(name="DLogistic"), round=dict(name="Round", nbits=8))
    cfg.update(extra)
    return cfg


def build_tiny(name="IDFlows", **extra):
    from flic_b200 import flows
    torch.manual_seed(0)
    random.seed(0)
    model = flows.build_model(tiny_cfg(name, **extra))
    flows.perturb_heads(model, 0.02, seed=0)
    return model.eval()


@pytest.mark.parametrize("name,prefix", [("IDFlows", "id.sd."), ("ConditionalFlows", "cond.sd.")])
def test_same_seeds_same_weights_as_reference(name, prefix):
    """Construction order (and so RNG consumption), parameter names and shapes mirror the
    reference: the golden state_dict was produced by the reference's own classes."""
    g = np.load(GOLDEN)
    model = build_tiny(name, **({"conv_for_cond": False} if name == "ConditionalFlows" else {}))
    sd = model.state_dict()
    ref_keys = sorted(k[len(prefix):] for k in g.files if k.startswith(prefix))
    assert sorted(sd.keys()) == ref_keys
    for k in ref_keys:
        assert np.array_equal(sd[k].numpy(), g[prefix + k]), k
    # and a reference checkpoint loads
    model.load_state_dict({k: torch.from_numpy(g[prefix + k]) for k in ref_keys}, strict=True)


def test_latents_shape_of_named_configs():
    """imagenet64 (6,32,32)/(12,16,16)/(48,8,8); config1@32 (6,16,16)/(12,8,8)/(48,4,4);
    resflows_smallpatch (12,4,4)."""
    from flic_b200 import flows
    m = flows.build_model(tiny_cfg(nsplit=3, H=64, W=64))
    assert m.latents_shape == [(6, 32, 32), (12, 16, 16), (48, 8, 8)]
    assert sum(int(np.prod(s)) for s in m.latents_shape) == 12288
    m = flows.build_model(tiny_cfg(nsplit=3, H=32, W=32))
    assert m.latents_shape == [(6, 16, 16), (12, 8, 8), (48, 4, 4)]
    m = flows.build_model(tiny_cfg(nsplit=1, H=8, W=8))
    assert m.latents_shape == [(12, 4, 4)]
    c = m.blocks[0]["flows"][1]
    assert (c.a_ch, c.b_ch) == (9, 3)                   # couplelib.py:38 split of 12 channels
    m = flows.build_model(tiny_cfg("ConditionalFlows", nsplit=1, H=27, W=23, extenddim=dict(name="ExtendDim", scale=1)))
    assert m.latents_shape == [(3, 27, 23)]
    assert m.blocks[0]["prior"].cond_channel == 6       # 3 zero + 3 cond channels (flows.py:295-297)


def test_segment_offsets_partitions():
    from flic_b200 import flows
    m = flows.build_model(tiny_cfg(nsplit=3, H=64, W=64))
    off = m._segment_offsets(0, 4, 1, "cpu")
    assert off.tolist() == [0, 6144, 12288, 18432, 24576]
    off = m._segment_offsets(1, 2, 3, "cpu")
    assert off.tolist() == [0, 1024, 2048, 3072, 4096, 5120, 6144]
    assert m._segment_offsets(2, 5, 0, "cpu").tolist() == [0, 5 * 3072]   # reference-native: one stream per level


def test_container_round_trip_on_cpu():
    from flic_b200.container import CompressedBatch
    from flic_b200.rans import EncodedStreams
    rng = np.random.default_rng(0)
    cb = CompressedBatch(5, (3, 16, 16), 2, 4, 1, model_tag=0xabcdef)
    for chunk_imgs in (4, 1):
        chunk = []
        for level in range(2):
            counts = rng.integers(0, 9, chunk_imgs)
            woff = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
            words = rng.integers(0, 1 << 32, int(woff[-1]), dtype=np.uint64).astype(np.uint32)
            states = rng.integers(1 << 32, 1 << 63, chunk_imgs, dtype=np.uint64)
            chunk.append(EncodedStreams(torch.from_numpy(words.view(np.int32).copy()), torch.from_numpy(woff),
                                        torch.from_numpy(states.view(np.int64).copy()),
                                        torch.zeros(chunk_imgs, dtype=torch.int32), 0))
        cb.sections.append(chunk)
    blob = cb.to_bytes()
    assert len(blob) == 44 + sum(4 + 12 * e.n_streams + 4 * e.n_words() for ch in cb.sections for e in ch)
    back = CompressedBatch.from_bytes(blob, "cpu")
    assert (back.n_images, back.shape, back.n_levels, back.codec_batch, back.model_tag) == (5, (3, 16, 16), 2, 4, 0xabcdef)
    for a, b in zip(cb.sections, back.sections):
        for ea, eb in zip(a, b):
            assert torch.equal(ea.words[:ea.n_words()], eb.words) and torch.equal(ea.word_offsets, eb.word_offsets)
            assert torch.equal(ea.final_states, eb.final_states)
    assert back.reference_bits() == cb.reference_bits() == 64 * cb.n_streams() + 32 * cb.n_words()
    assert back.to_bytes() == blob
    for bad in (blob[:-1], blob + b"\0", b"XXXX" + blob[4:], blob[:30]):
        with pytest.raises(ValueError):
            CompressedBatch.from_bytes(bad, "cpu")
    # version 1 (no flags word) is still read
    v1 = blob[:4] + (1).to_bytes(2, "little") + blob[6:40] + blob[44:]
    old = CompressedBatch.from_bytes(v1, "cpu")
    assert not old.chained and old.n_words() == cb.n_words()
    # chained: one section per chunk, every image one stream
    ch = CompressedBatch(5, (3, 16, 16), 2, 4, 1, chained=True)
    for chunk in cb.sections:
        ch.sections.append(chunk[:1])
    back = CompressedBatch.from_bytes(ch.to_bytes(), "cpu")
    assert back.chained and [len(c) for c in back.sections] == [1, 1]
    assert back.reference_bits() == 64 * 5 + 32 * sum(c[0].n_words() for c in cb.sections)


def test_drop_in_argument_checks():
    """rans.encode / decode reject what the reference rejects, before touching the GPU
    (rans/rans.cpp:1585-1587 TypeError for non-list; :1571 OverflowError for a state >= 2^64)."""
    from flic_b200 import rans
    with pytest.raises(TypeError):
        rans.encode(1 << 32, 1, (0.0,), [0.0], [1.0])
    with pytest.raises(TypeError):
        rans.decode(1 << 32, np.zeros(1), 1, [0.0], [1.0])
    with pytest.raises(OverflowError):
        rans.encode(1 << 64, 1, [0.0], [0.0], [1.0])
    with pytest.raises(OverflowError):
        rans.decode(-1, [], 1, [0.0], [1.0])
    assert rans.encode(1 << 32, 0, [], [], []) == (1 << 32, [])           # n = 0: nothing coded
    assert rans.decode(12345 << 32, [], 0, None, None) == (12345 << 32, [])
    with pytest.raises(IndexError):
        rans.encode(1 << 32, 3, [0.0], [0.0], [1.0])


def test_uniform_offsets():
    from flic_b200 import rans
    assert rans.uniform_offsets(3, 5).tolist() == [0, 5, 10, 15]
    assert rans.uniform_offsets(2, 0).tolist() == [0, 0, 0]


def test_shard_ranges_cover_everything():
    from flic_b200.sharding import shard_range
    for n in (0, 1, 7, 8, 1_000_003):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_round_module_matches_reference_formula():
    """roundlib.Round on the CPU (ordinary torch code): golden from the reference's Round."""
    from flic_b200.roundlib import Round
    g = np.load(GOLDEN)
    out = Round(nbits=8)(torch.from_numpy(g["k5.ties"]))
    assert np.array_equal(out.numpy(), g["k5.round_ties"])


def test_numpy_oracle_pieces_match_reference_goldens(oracle):
    g = np.load(GOLDEN)
    a = int(g["k5.a_ch"])
    z = oracle.couple_forward(g["k5.x"][:, a:], g["k5.t"])
    assert np.array_equal(z, g["k5.z"][:, a:])
    assert np.array_equal(oracle.couple_backward(g["k5.z"][:, a:], g["k5.t"]), g["k5.x"][:, a:])
    assert np.array_equal(oracle.round_nbits(g["k5.ties"]), g["k5.round_ties"])
    assert np.array_equal(oracle.quantise_input_u8(g["id.u8"]), g["id.x"])


def test_oracle_reproduces_reference_coding_loop(oracle):
    """The reference's trainer.py:308-327 loop, run by the reference itself when the golden was
    made: words per level, final states, real bpd."""
    g = np.load(GOLDEN)
    total_words = 0
    for i in range(2):
        x, mean = g[f"id.latent{i}"].reshape(-1), g[f"id.mean{i}"].reshape(-1)
        scale = g[f"id.scale{i}"].reshape(-1)                         # exp(logscale) as the reference took it
        state, buf = oracle.encode(1 << 32, x.size, x, mean, scale)
        assert state == int(g["id.states"][i]) and np.array_equal(buf, g[f"id.words{i}"])
        total_words += buf.size
    assert (64 * 2 + 32 * total_words) / g["id.x"].size == float(g["id.real_bpd"])


# ---- TwoLevelFlows (flows.py:184-274, configs/config_twolevel.yaml) ------------------------------

def twolevel_cfg():
    layer = dict(name="DenseLayer", act="ReLU")

    def sub(H, W, scale, nflows):
        return dict(name="IDFlows", nflows=nflows, nbits=8, nsplit=1, H=H, W=W, C=3,
                    couple=dict(name="AdditiveCouple", split=0.75, round=dict(name="Round", nbits=8),
                                nn=dict(name="DenseBlock", growth_channel=16, depth=2, layer=dict(layer))),
                    extenddim=dict(name="ExtendDim", scale=scale),
                    prior=dict(name="Prior", round=dict(name="Round", nbits=8),
                               nn=dict(name="DenseBlock", growth_channel=16, depth=2, layer=dict(layer))),
                    distribution=dict(name="DLogistic"), round=dict(name="Round", nbits=8))
    return dict(name="TwoLevelFlows", H=215, W=178, C=3, pad=[1, 6], fine_flows=sub(8, 8, 2, 3),
                rough_flows=sub(27, 23, 1, 3), batchsize=1536)


def test_twolevel_structure_and_split_match_the_reference():
    """Same seeds -> the reference's state_dict (names, shapes, values: construction order), the
    same latents_shape, and the same rough image / residual patches for the same input
    (tests/golden/twolevel.json was produced by the reference's own TwoLevelFlows)."""
    import json
    from flic_b200 import flows
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "twolevel.json")))
    torch.manual_seed(0)
    random.seed(0)
    model = flows.build_model(twolevel_cfg()).eval()
    sd = model.state_dict()
    assert sorted(sd.keys()) == sorted(g["state"].keys())
    for k, (shape, total) in g["state"].items():
        assert list(sd[k].shape) == shape, k
        assert abs(float(sd[k].double().sum()) - total) <= 1e-9 * max(1.0, abs(total)), k
    assert [list(s) for s in model.latents_shape] == g["latents_shape"]
    u8 = torch.randint(0, 256, (2, 3, 215, 178), generator=torch.Generator().manual_seed(5), dtype=torch.uint8)
    x = torch.round(u8.float() / 255 * 256) / 256
    rx, px = model.split(x)
    assert list(px.shape) == g["px_shape"]
    assert float(rx.double().sum()) == g["rx_sum"] and rx[0, 0, 0, :8].tolist() == g["rx_head"]
    assert float(px.double().abs().sum()) == g["px_abs_sum"] and px[5].flatten().tolist() == g["px_patch5"]
    assert torch.equal(model.merge(rx, px), x)               # x = upsample(rx) + fx, exactly
