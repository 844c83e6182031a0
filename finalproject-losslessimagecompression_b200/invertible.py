"""Permute / InvertibleModuleList (reference: invertible.py:24-71).

The reference applies a channel permutation as NCHW->NHWC copy, F.linear with a dim x dim
permutation matrix, NHWC->NCHW copy.  Here it is one gather kernel (csrc/flow_index.cu).  The
matrices `P` / `inv_P` are kept as (frozen) parameters under the reference's names so that its
checkpoints load and `random.seed` reproduces the same permutations (invertible.py:33); the
index vectors are derived from them.
"""
import random

import torch
from torch import nn
from torch.nn.parameter import Parameter

from . import _lib


class InvertibleModule(nn.Module):
    def forward(self, *args, **kwargs):
        raise NotImplementedError

    def backward(self, *args, **kwargs):
        raise NotImplementedError

    def inverse(self, *args, **kwargs):  # a no-op in the reference too (invertible.py:20-21)
        pass


def permute_channels(x: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    """out[:, i] = x[:, perm[i]] for NCHW float32 CUDA x; perm int32 CUDA."""
    if not x.is_cuda:
        raise _lib.FlicError("permute_channels needs a CUDA tensor (no CPU fallback)")
    x = x.contiguous()
    B, Cc, H, W = x.shape
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().flic_permute_channels(x.data_ptr(), out.data_ptr(), perm.data_ptr(), B, Cc, H * W,
                                                    torch.cuda.current_stream(x.device).cuda_stream),
                   "flic_permute_channels")
    return out


class Permute(InvertibleModule):
    def __init__(self, dim):
        super().__init__()
        ids = list(range(dim))
        random.shuffle(ids)                      # same RNG call as invertible.py:33
        p = torch.zeros((dim, dim))
        p[torch.arange(dim), torch.tensor(ids)] = 1
        self.P = Parameter(p, requires_grad=False)
        self.inv_P = Parameter(p.t().clone(), requires_grad=False)
        self._cache = None

    def _indices(self, device):
        key = (self.P._version, self.P.data_ptr(), str(device))
        if self._cache is None or self._cache[0] != key:
            fwd = torch.argmax(self.P.detach(), dim=1).to(device=device, dtype=torch.int32)      # out[i] = x[ids[i]]
            bwd = torch.argmax(self.inv_P.detach(), dim=1).to(device=device, dtype=torch.int32)  # inverse gather
            self._cache = (key, fwd.contiguous(), bwd.contiguous())
        return self._cache[1], self._cache[2]

    def forward(self, x, logv):
        return permute_channels(x, self._indices(x.device)[0]), logv

    def backward(self, x):
        return permute_channels(x, self._indices(x.device)[1])


class InvertibleModuleList(InvertibleModule, nn.ModuleList):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)

    def inverse(self):
        for m in self:
            if isinstance(m, InvertibleModule):
                m.inverse()

    def forward(self, x, logv, *args, **kwargs):
        for m in self:
            x, logv = m.forward(x, logv, *args, **kwargs)
        return x, logv

    def backward(self, x, *args, **kwargs):
        for m in reversed(list(self)):
            x = m.backward(x, *args, **kwargs)
        return x


class LULinear(InvertibleModule):  # empty in the reference as well (invertible.py:74-76)
    pass
