"""DLogistic: the discretised logistic (reference: distlib.py:30-70).

log_prob is the *ideal* code length the coder's real cost is compared against
(trainer.py:269-272 vs :326-327); sample draws latents for generated_from_noise.

On CUDA tensors without autograd both run as one fused kernel each (csrc/logistic_prob.cu:
flic_dlogistic_log_prob / flic_dlogistic_sample), with the per-image
reduction of IDFlows.log_likelihood available in the same pass (log_prob_sums).  Only when a
gradient is needed (training, which is outside the coding path) does the same formula run as
differentiable torch ops; without autograd there is no CPU implementation -- CPU tensors raise.
"""
import ctypes as C
from copy import deepcopy

import torch
from torch import nn
from torch.nn import functional as F

from . import _lib
from .moduleregister import Register
from .roundlib import NNRound, Round


class NNDistribution(Register):
    pass


def _needs_autograd(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t.requires_grad for t in tensors)


def _fused_ok(*tensors) -> bool:
    """True: run the fused kernel.  False: autograd is needed, use the differentiable torch formula.
    Raises for inputs the kernel cannot take (CPU tensors, other dtypes): there is no CPU fallback."""
    if _needs_autograd(*tensors):
        return False
    if not all(t.is_cuda for t in tensors):
        raise _lib.FlicError("DLogistic needs CUDA tensors (no CPU implementation outside autograd)")
    if not all(t.dtype == torch.float32 for t in tensors):
        raise TypeError("DLogistic expects float32 tensors")
    return True


def _stream_ptr(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


@NNDistribution.register
class DLogistic(nn.Module):
    def __init__(self, round=None):
        super().__init__()
        if round:
            round = deepcopy(round)
            self.round = NNRound.get(round.pop("name"))(**round)
        else:
            self.round = Round()

    @staticmethod
    def _log_prob_torch(x, mean, logscale, nbits=8, eps=1e-8):
        """The formula of distlib.py:40-55 as differentiable torch ops.  Kept on purpose, for one
        caller only: a training loop that needs d(log_prob)/d(params) (loading a reference
        checkpoint and fine-tuning it must keep working).  It is NOT a second backend of the coding
        path: compress / decompress / log_likelihood run under torch.no_grad() and always take the
        CUDA kernel; CPU tensors raise (`_fused_ok`), they never come here."""
        scale = torch.exp(logscale)
        half = 0.5 / (2 ** nbits)
        up = F.logsigmoid((x + half - mean) / scale)
        dn = F.logsigmoid((x - half - mean) / scale)
        return up + torch.log(1 - torch.exp(dn - up) + eps)   # distlib.py:52-55

    def log_prob(self, x, mean, logscale, nbits=8, eps=1e-8):
        if not _fused_ok(x, mean, logscale):
            return self._log_prob_torch(x, mean, logscale, nbits, eps)
        x, mean, logscale = torch.broadcast_tensors(x, mean, logscale)
        if x.numel() == 0:
            return torch.empty_like(x)
        xv, mv, lv = x.contiguous(), mean.contiguous(), logscale.contiguous()
        out = torch.empty_like(xv)
        with torch.cuda.device(xv.device):
            _lib.check(_lib.lib().flic_dlogistic_log_prob(xv.data_ptr(), mv.data_ptr(), lv.data_ptr(), 1, xv.numel(),
                                                          int(nbits), C.c_float(eps), out.data_ptr(), None,
                                                          _stream_ptr(xv.device)), "flic_dlogistic_log_prob")
        return out

    def log_prob_sums(self, x, mean, logscale, nbits=8, eps=1e-8):
        """sum of log_prob over every dimension but the first, (B,) float32: the reduction of
        IDFlows.log_likelihood (flows.py:165-167) fused into the evaluation."""
        if not _fused_ok(x, mean, logscale):
            return self._log_prob_torch(x, mean, logscale, nbits, eps).flatten(1).sum(1)
        x, mean, logscale = torch.broadcast_tensors(x, mean, logscale)
        if x.numel() == 0:
            return torch.zeros(x.shape[0], dtype=torch.float32, device=x.device)
        xv, mv, lv = x.contiguous(), mean.contiguous(), logscale.contiguous()
        b = xv.shape[0]
        out = torch.empty(b, dtype=torch.float32, device=xv.device)
        with torch.cuda.device(xv.device):
            _lib.check(_lib.lib().flic_dlogistic_log_prob(xv.data_ptr(), mv.data_ptr(), lv.data_ptr(), b,
                                                          xv.numel() // b, int(nbits), C.c_float(eps), None,
                                                          out.data_ptr(), _stream_ptr(xv.device)),
                       "flic_dlogistic_log_prob")
        return out

    def sample(self, mean, logscale, nbits=8):
        u = torch.rand_like(mean)
        if _fused_ok(mean, logscale) and mean.shape == logscale.shape and mean.numel() > 0 and type(self.round) is Round:
            mv, lv, uv = mean.contiguous(), logscale.contiguous(), u.contiguous()
            out = torch.empty_like(mv)
            with torch.cuda.device(mv.device):
                _lib.check(_lib.lib().flic_dlogistic_sample(uv.data_ptr(), mv.data_ptr(), lv.data_ptr(), mv.numel(),
                                                            int(nbits), out.data_ptr(), _stream_ptr(mv.device)),
                           "flic_dlogistic_sample")
            return out
        z = torch.log(u / (1 - u)) * torch.exp(logscale) + mean
        return self.round(z, nbits=nbits)
