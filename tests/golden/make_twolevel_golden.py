"""Golden structure of the REFERENCE's TwoLevelFlows (flows.py:184-274), imported on CPU here.

    python tests/golden/make_twolevel_golden.py        (needs /root/reference)

Writes tests/golden/twolevel.json: for a TwoLevelFlows with the shapes of configs/config_twolevel.yaml
(215 x 178 images, pad (1, 6), rough 27 x 23 with ExtendDim scale 1, fine 8 x 8 patches; network
width reduced to growth 16 / depth 2) built after torch.manual_seed(0); random.seed(0):
every state_dict key with its shape and the float64 sum of its values (construction-order and name
parity), latents_shape, and -- for one random grid image -- the rough image, the first fine patches
and their checksums as the reference's own forward computes them (pool / round / residual / patching).
"""
import json
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_flow_golden import _import_reference  # noqa: E402

LAYER = dict(name="DenseLayer", act="ReLU")


def sub(H, W, scale, nflows):
    return dict(name="IDFlows", nflows=nflows, nbits=8, nsplit=1, H=H, W=W, C=3,
                couple=dict(name="AdditiveCouple", split=0.75, round=dict(name="Round", nbits=8),
                            nn=dict(name="DenseBlock", growth_channel=16, depth=2, layer=dict(LAYER))),
                extenddim=dict(name="ExtendDim", scale=scale),
                prior=dict(name="Prior", round=dict(name="Round", nbits=8),
                           nn=dict(name="DenseBlock", growth_channel=16, depth=2, layer=dict(LAYER))),
                distribution=dict(name="DLogistic"), round=dict(name="Round", nbits=8))


def cfg():
    return dict(H=215, W=178, C=3, pad=[1, 6], fine_flows=sub(8, 8, 2, 3), rough_flows=sub(27, 23, 1, 3), batchsize=1536)


def main():
    ref_flows, _ = _import_reference()
    torch.manual_seed(0)
    random.seed(0)
    model = ref_flows.TwoLevelFlows(**cfg()).eval()
    out = {"state": {k: [list(v.shape), float(v.double().sum())] for k, v in model.state_dict().items()},
           "latents_shape": [list(s) for s in model.latents_shape]}
    g = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (2, 3, 215, 178), generator=g, dtype=torch.uint8)
    x = torch.round(u8.float() / 255 * 256) / 256
    with torch.no_grad():
        xp = model.pad2d(x)
        rx = model.round(model.pool(xp))
        fx = xp - model.invpool(rx)
        px, _ = model.patching(fx, None)
    out["rx_sum"] = float(rx.double().sum())
    out["rx_head"] = rx[0, 0, 0, :8].tolist()
    out["px_shape"] = list(px.shape)
    out["px_abs_sum"] = float(px.double().abs().sum())
    out["px_patch5"] = px[5].flatten().tolist()
    json.dump(out, open(os.path.join(HERE, "twolevel.json"), "w"))
    print("wrote", len(out["state"]), "tensors", out["latents_shape"], out["px_shape"])


if __name__ == "__main__":
    main()
