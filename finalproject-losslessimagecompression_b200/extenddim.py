"""ExtendDim (space-to-depth) and Patching (reference: extenddim.py:17-67).

ExtendDim is a gather kernel (csrc/flow_index.cu); Patching is a pure view/permute/copy that runs
once per image at the edge of the path and stays a torch expression.
"""
import torch

from . import _lib
from .invertible import InvertibleModule
from .moduleregister import Register


class NNExtendDim(Register):
    pass


def squeeze(x: torch.Tensor, scale: int, direction: int) -> torch.Tensor:
    """direction +1: (B,C,H,W) -> (B,C*s*s,H/s,W/s); -1: the inverse.  float32 CUDA only."""
    if not x.is_cuda:
        raise _lib.FlicError("squeeze needs a CUDA tensor (no CPU fallback)")
    x = x.contiguous()
    B, Cc, H, W = x.shape
    s = int(scale)
    if direction > 0:
        if H % s or W % s:
            raise ValueError("H and W must be multiples of scale")
        out = torch.empty((B, Cc * s * s, H // s, W // s), dtype=x.dtype, device=x.device)
        big = (B, Cc, H, W)
    else:
        if Cc % (s * s):
            raise ValueError("channels must be a multiple of scale^2")
        out = torch.empty((B, Cc // (s * s), H * s, W * s), dtype=x.dtype, device=x.device)
        big = tuple(out.shape)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().flic_squeeze(x.data_ptr(), out.data_ptr(), *big, s, int(direction),
                                           torch.cuda.current_stream(x.device).cuda_stream), "flic_squeeze")
    return out


@NNExtendDim.register
class ExtendDim(InvertibleModule):
    def __init__(self, scale=2):
        super().__init__()
        self.scale = scale

    def forward(self, x, logv):
        return squeeze(x, self.scale, +1), logv

    def backward(self, x):
        return squeeze(x, self.scale, -1)


@Register.register
class Patching(InvertibleModule):
    """Image -> batch of h x w patches, row-major over the patch grid (extenddim.py:40-67)."""

    def __init__(self, H, W, h, w):
        assert H % h == 0 and W % w == 0
        super().__init__()
        self.H, self.W, self.h, self.w = H, W, h, w

    def forward(self, x, logv):
        B, Cc = x.shape[0], x.shape[1]
        x = x.reshape(B, Cc, self.H // self.h, self.h, self.W // self.w, self.w)
        return x.permute(0, 2, 4, 1, 3, 5).reshape(-1, Cc, self.h, self.w).contiguous(), logv

    def backward(self, x):
        hh, ww = self.H // self.h, self.W // self.w
        Cc = x.shape[1]
        x = x.reshape(x.shape[0] // (hh * ww), hh, ww, Cc, self.h, self.w)
        return x.permute(0, 3, 1, 4, 2, 5).reshape(-1, Cc, self.H, self.W).contiguous()
