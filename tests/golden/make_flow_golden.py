"""Flow golden vectors from the REFERENCE's own modules, imported on CPU in the build container.

    python tests/golden/make_flow_golden.py        (needs /root/reference)

Writes tests/golden/flow_tiny.npz: for a tiny IDFlows and a tiny ConditionalFlows (same module
types as configs/config1.yaml, growth 16 / depth 2 so the fixture stays small)
  * the reference's state_dict after `torch.manual_seed(0); random.seed(0)` construction and
    N(0, 0.02) re-initialisation of the zero heads -- the mirror must reproduce it from the
    same seeds (construction-order parity) and load it (name parity);
  * input grid floats and the reference's latents / means / logscales / ideal log-likelihood;
  * AdditiveCouple.forward / backward and Round outputs for a tie-heavy input (K5 parity);
  * Permute / ExtendDim outputs (N1 parity);
  * the reference's trainer.py:308-327 coding loop run on those latents with the reference's
    own rans module: words per level, final states, real bpd.
"""
import os
import random
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference"

LAYER = dict(name="DenseLayer", act="LeakyReLU")
TINY = dict(nflows=2, nbits=8, nsplit=2, H=16, W=16, C=3,
            couple=dict(name="AdditiveCouple", split=0.75, round=dict(name="Round", nbits=8),
                        nn=dict(name="DenseBlock", growth_channel=16, depth=2, layer=LAYER)),
            extenddim=dict(name="ExtendDim", scale=2),
            prior=dict(name="Prior", round=dict(name="Round", nbits=8),
                       nn=dict(name="DenseBlock", growth_channel=16, depth=2, layer=LAYER)),
            distribution=dict(name="DLogistic"), round=dict(name="Round", nbits=8))


def _import_reference():
    shim = tempfile.mkdtemp(prefix="flic_shim_")
    os.makedirs(os.path.join(shim, "colorama"))
    with open(os.path.join(shim, "colorama", "__init__.py"), "w") as f:
        f.write("def reinit(*a, **k):\n    pass\n")   # roundlib.py:1 imports it and never uses it
    sys.path.insert(0, shim)
    sys.path.insert(0, REFERENCE)
    import flows as ref_flows          # noqa
    import nnblock as ref_nnblock      # noqa
    return ref_flows, ref_nnblock


def _perturb(model, DenseBlock, std=0.02, seed=0):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, DenseBlock):
                head = m.layers[-1]
                head.weight.copy_(torch.randn(head.weight.shape, generator=g) * std)
                head.bias.copy_(torch.randn(head.bias.shape, generator=g) * std)


def flow_vectors():
    from copy import deepcopy
    sys.path.insert(0, ROOT)
    from oracle import pyoracle
    ref_flows, ref_nnblock = _import_reference()
    out = {}
    torch.manual_seed(0)
    random.seed(0)
    model = ref_flows.IDFlows(**deepcopy(TINY)).eval()
    _perturb(model, ref_nnblock.DenseBlock)
    for k, v in model.state_dict().items():
        out["id.sd." + k] = v.numpy()
    u8 = torch.randint(0, 256, (6, 3, 16, 16), dtype=torch.uint8, generator=torch.Generator().manual_seed(1234))
    x = torch.round(u8.float() / 255 * 256) / 256      # ToTensor + Round(nbits=8), trainer.py:61,72
    out["id.u8"] = u8.numpy()
    out["id.x"] = x.numpy()
    with torch.no_grad():
        lat, means, logs, _ = model.forward(x, None)
        logp, _ = model.log_likelihood(lat, means, logs)
        rec = model.generated_from_latents(lat)
    assert torch.equal(rec, x)
    out["id.logp"] = logp.numpy()
    ref_rans = pyoracle.ref_rans()
    words, states = [], []
    for i in range(len(lat)):
        out[f"id.latent{i}"] = lat[i].contiguous().numpy()
        out[f"id.mean{i}"] = means[i].contiguous().numpy()
        out[f"id.logscale{i}"] = logs[i].contiguous().numpy()
        # the coder's actual input: exp taken by the reference's torch call (trainer.py:313); stored
        # because torch.exp may differ by an ulp between CPU ISAs and between CPU and CUDA
        out[f"id.scale{i}"] = torch.exp(logs[i]).contiguous().numpy()
        # trainer.py:308-323, verbatim call pattern
        xi = lat[i].reshape(-1).tolist()
        mi = means[i].reshape(-1).tolist()
        si = torch.exp(logs[i]).reshape(-1).tolist()
        state, buf = ref_rans.encode(1 << 32, len(xi), xi, mi, si)
        end, msg = ref_rans.decode(state, buf[::-1], len(xi), mi[::-1], si[::-1])
        assert end == 1 << 32 and msg[::-1] == xi
        out[f"id.words{i}"] = np.asarray(buf, dtype=np.uint32)
        states.append(state)
        words.append(len(buf))
    out["id.states"] = np.asarray(states, dtype=np.uint64)
    out["id.real_bpd"] = np.float64((64 * len(lat) + 32 * sum(words)) / x.numel())   # trainer.py:327

    # one coupling layer + Round on a tie-heavy input (values k/512 hit the rounding ties)
    couple = model.blocks[0]["flows"][1]
    g = torch.Generator().manual_seed(5)
    xin = torch.randint(-512, 512, (4, 12, 8, 8), generator=g).float() / 256
    with torch.no_grad():
        t = couple.dense(xin[:, :couple.a_ch])
        z, _ = couple.forward(xin, None)
        xb = couple.backward(z)
    assert torch.equal(xb, xin)
    ties = torch.randint(-2048, 2048, (4, 3, 8, 8), generator=g).float() / 512
    out["k5.x"] = xin.numpy()
    out["k5.t"] = t.numpy()
    out["k5.z"] = z.numpy()
    out["k5.ties"] = ties.numpy()
    out["k5.round_ties"] = couple.round(ties).numpy()
    out["k5.a_ch"] = np.int64(couple.a_ch)
    # Permute / ExtendDim
    perm = model.blocks[0]["flows"][0]
    with torch.no_grad():
        out["n1.perm_fwd"] = perm.forward(xin, None)[0].numpy()
        out["n1.perm_bwd"] = perm.backward(xin).numpy()
        out["n1.perm_P"] = perm.P.numpy()
        out["n1.squeeze_fwd"] = model.blocks[0]["extend"].forward(x, None)[0].numpy()
        out["n1.squeeze_bwd"] = model.blocks[0]["extend"].backward(xin).numpy()

    # ConditionalFlows (flows.py:277-361): prior sees a conditioning image
    torch.manual_seed(0)
    random.seed(0)
    cmodel = ref_flows.ConditionalFlows(conv_for_cond=False, **deepcopy(TINY)).eval()
    _perturb(cmodel, ref_nnblock.DenseBlock)
    for k, v in cmodel.state_dict().items():
        out["cond.sd." + k] = v.numpy()
    cond = torch.round(torch.rand(6, 3, 16, 16, generator=torch.Generator().manual_seed(7)) * 256) / 256
    out["cond.cond"] = cond.numpy()
    with torch.no_grad():
        lat, means, logs, _ = cmodel.forward(x, None, cond)
    for i in range(len(lat)):
        out[f"cond.latent{i}"] = lat[i].contiguous().numpy()
        out[f"cond.mean{i}"] = means[i].contiguous().numpy()
        out[f"cond.logscale{i}"] = logs[i].contiguous().numpy()
    np.savez_compressed(os.path.join(HERE, "flow_tiny.npz"), **out)
    print("flow vectors written:", len(out), "arrays")


if __name__ == "__main__":
    flow_vectors()
