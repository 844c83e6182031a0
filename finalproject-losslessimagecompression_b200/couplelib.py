"""AdditiveCouple: the integer additive-coupling step (reference: couplelib.py:24-61).

forward   z = cat(xa, xb + Round(dense(xa)))      couplelib.py:47-53
backward  x = cat(za, zb - Round(dense(za)))      couplelib.py:55-61

`dense` stays a PyTorch fp32 DenseBlock; everything after it -- scale by 2^nbits, round to
nearest even, straight-through form, rescale, add/subtract, concatenate -- is one CUDA kernel
(K5, csrc/couple_round.cu) that updates the b-channels of x in place.  Inference only.
"""
from copy import deepcopy

import torch

from . import _lib
from .invertible import InvertibleModule
from .moduleregister import Register
from .nnblock import NNBlock
from .roundlib import NNRound


class NNCouple(Register):
    pass


def couple_add_round(x: torch.Tensor, t: torch.Tensor, a_ch: int, direction: int, nbits: int = 8) -> torch.Tensor:
    """In place: x[:, a_ch:] += direction * Round_nbits(t).  x (B,C,H,W) float32 contiguous CUDA,
    t (B,C-a_ch,H,W).  Returns x.  Raises on CPU tensors: there is no CPU implementation."""
    if not (x.is_cuda and t.is_cuda):
        raise _lib.FlicError("couple_add_round needs CUDA tensors (no CPU fallback)")
    if x.dtype != torch.float32 or t.dtype != torch.float32:
        raise TypeError("couple_add_round works on float32")
    if not x.is_contiguous():
        raise ValueError("x must be contiguous (it is updated in place)")
    B, Cc, H, W = x.shape
    if t.shape != (B, Cc - a_ch, H, W):
        raise ValueError(f"t has shape {tuple(t.shape)}, expected {(B, Cc - a_ch, H, W)}")
    t = t.contiguous()
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().flic_couple_add_round(x.data_ptr(), t.data_ptr(), B, Cc, a_ch, H * W, int(direction),
                                                    int(nbits), torch.cuda.current_stream(x.device).cuda_stream),
                   "flic_couple_add_round")
    return x


@NNCouple.register
class AdditiveCouple(InvertibleModule):
    def __init__(self, channel, split=0.75, nn=None, round=None):
        super().__init__()
        self.channel, self.split = channel, split
        self.a_ch = int(channel * split)          # couplelib.py:38
        self.b_ch = channel - self.a_ch
        nn, round = deepcopy(nn), deepcopy(round)
        self.dense = NNBlock.get(nn.pop("name"))(i_channel=self.a_ch, o_channel=self.b_ch, **nn)
        self.round = NNRound.get(round.pop("name"))(**round)

    def _nbits(self, nbits):
        return nbits or getattr(self.round, "nbits", None) or 8

    @staticmethod
    def _writable(x, inplace):
        y = x.contiguous()
        return y.clone() if (not inplace and y.data_ptr() == x.data_ptr()) else y

    @torch.no_grad()
    def forward(self, x, logv, nbits=None, inplace=False):
        """inplace=True lets the kernel overwrite x's b-channels (the flow passes a fresh tensor)."""
        x = self._writable(x, inplace)
        t = self.dense(x[:, :self.a_ch].contiguous())
        return couple_add_round(x, t, self.a_ch, +1, self._nbits(nbits)), logv

    @torch.no_grad()
    def backward(self, z, nbits=None, inplace=False):
        z = self._writable(z, inplace)
        t = self.dense(z[:, :self.a_ch].contiguous())
        return couple_add_round(z, t, self.a_ch, -1, self._nbits(nbits))
