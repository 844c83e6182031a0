"""Prior: conditioning tensor -> (mean, logscale) of the discretised logistic
(reference: priorlib.py:18-47).  A DenseBlock whose output is split in channel halves; the
unconditional (last-level) case feeds zeros (priorlib.py:42).  Plain PyTorch conv: this is the
producer of the coder's inputs, the boundary of the hot path.
"""
from copy import deepcopy

import torch
from torch import nn

from .moduleregister import Register
from .nnblock import NNBlock
from .roundlib import NNRound


class NNPrior(Register):
    pass


@NNPrior.register
class Prior(nn.Module):
    def __init__(self, out_channel, cond_channel, round=None, nn=None):
        super().__init__()
        self.out_channel, self.cond_channel = out_channel, cond_channel
        round, nn = deepcopy(round), deepcopy(nn)
        self.round = NNRound.get(round.pop("name"))(**round)
        block = NNBlock.get(nn.pop("name"))
        self.NN = block(cond_channel if cond_channel > 0 else out_channel, out_channel * 2, **nn)

    def forward(self, cond):
        params = self.NN(cond if self.cond_channel > 0 else torch.zeros_like(cond))
        return params[:, :self.out_channel], params[:, self.out_channel:]
