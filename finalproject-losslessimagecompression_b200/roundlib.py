"""Round: quantisation to the 2^-nbits grid (reference: roundlib.py:18-38).

Round(x) = STE(rint(x * 2^nbits)) / 2^nbits with torch.round's ties-to-even.  Inside the coupling
layers the rounding is fused into the add (couplelib.couple_add_round, CUDA kernel K5); this
standalone module serves the places where the reference rounds a tensor on its own
(input quantisation trainer.py:61,72, DLogistic.sample distlib.py:69) and is ordinary torch code
on whatever device the tensor lives on.  Inference only: the straight-through gradient is not
reproduced.
"""
import torch
from torch import nn

from .moduleregister import Register


class NNRound(Register):
    pass


class BaseRound(nn.Module):
    def forward(self, x):
        return x + (torch.round(x) - x).detach()  # roundlib.py:23-24, kept literally (sign of zero)


@NNRound.register
class Round(nn.Module):
    def __init__(self, nbits=None):
        super().__init__()
        self.nbits = nbits
        self.round = BaseRound()

    def forward(self, x, nbits=None):
        bins = 2 ** (nbits or self.nbits or 8)
        return self.round(x * bins) / bins
