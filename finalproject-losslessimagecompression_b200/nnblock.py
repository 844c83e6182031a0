"""DenseLayer / DenseBlock: the coupling and prior sub-networks (reference: nnlayer.py:22-51,
nnblock.py:24-56).

These stay plain PyTorch fp32 convolutions (BASELINE.json north_star: "the coupling subnetworks'
convolutions stay PyTorch fp32 with deterministic algorithms"); they are here only because the
hot path cannot be driven without them.  Parameter names and construction order match the
reference so that its checkpoints load (`layers.<i>.layers.0/1.weight`, final 1x1 at
`layers.<depth>`) and so that the same torch seed gives the same weights.
"""
from copy import deepcopy

import torch
from torch import nn

from .moduleregister import Register


class NNLayer(Register):
    pass


class NNBlock(Register):
    pass


_ACTS = {"ReLU": nn.ReLU, "Tanh": nn.Tanh, "LeakyReLU": nn.LeakyReLU}


@NNLayer.register
class DenseLayer(nn.Module):
    """x -> cat(x, act(conv3x3(conv1x1(x)))), growing the channel count (nnlayer.py:22-51)."""

    def __init__(self, i_channel, o_channel, act="ReLU"):
        super().__init__()
        self.i_channel, self.o_channel = i_channel, o_channel
        self.act = _ACTS[act]() if act in _ACTS else Register.get(act)()
        self.layers = nn.Sequential(
            nn.Conv2d(i_channel, i_channel, kernel_size=1),
            nn.Conv2d(i_channel, o_channel - i_channel, kernel_size=3, padding=1),
            self.act,
        )

    def forward(self, x):
        return torch.cat((x, self.layers(x)), dim=1)


@NNBlock.register
class DenseBlock(nn.Module):
    """`depth` DenseLayers adding `growth_channel` channels in total, then a zero-initialised 1x1
    head to `o_channel` (nnblock.py:24-56; zero init at :50-51)."""

    def __init__(self, i_channel, o_channel, layer, growth_channel=512, depth=8):
        super().__init__()
        self.i_channel, self.o_channel = i_channel, o_channel
        self.growth_channel, self.depth = growth_channel, depth
        layer = deepcopy(layer)
        layer_type = NNLayer.get(layer.pop("name"))
        self.layers = nn.ModuleList()
        channel = i_channel
        for idx in range(depth):
            growth = (idx + 1) * growth_channel // depth - idx * growth_channel // depth
            self.layers.append(layer_type(i_channel=channel, o_channel=channel + growth, **deepcopy(layer)))
            channel += growth
        assert channel == i_channel + growth_channel
        self.layers.append(nn.Conv2d(channel, o_channel, 1))
        with torch.no_grad():
            self.layers[-1].weight.zero_()
            self.layers[-1].bias.zero_()

    def forward(self, x):
        for layer in self.layers:
            x = layer(x)
        return x
