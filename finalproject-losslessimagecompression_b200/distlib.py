"""DLogistic: the discretised logistic (reference: distlib.py:30-70).

log_prob is the *ideal* code length the coder's real cost is compared against
(trainer.py:269-272 vs :326-327); sample draws latents for generated_from_noise.  Elementwise
torch code, same formulas as the reference so the ideal bpd matches it to float rounding.
"""
from copy import deepcopy

import torch
from torch import nn
from torch.nn import functional as F

from .moduleregister import Register
from .roundlib import NNRound, Round


class NNDistribution(Register):
    pass


@NNDistribution.register
class DLogistic(nn.Module):
    def __init__(self, round=None):
        super().__init__()
        if round:
            round = deepcopy(round)
            self.round = NNRound.get(round.pop("name"))(**round)
        else:
            self.round = Round()

    def log_prob(self, x, mean, logscale, nbits=8, eps=1e-8):
        scale = torch.exp(logscale)
        half = 0.5 / (2 ** nbits)
        up = F.logsigmoid((x + half - mean) / scale)
        dn = F.logsigmoid((x - half - mean) / scale)
        return up + torch.log(1 - torch.exp(dn - up) + eps)   # distlib.py:52-55

    def sample(self, mean, logscale, nbits=8):
        u = torch.rand_like(mean)
        z = torch.log(u / (1 - u)) * torch.exp(logscale) + mean
        return self.round(z, nbits=nbits)
