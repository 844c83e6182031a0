"""Encode / Decode over a list of latent levels (reference: coder.py:18-38).

Same signatures and return values as the reference's wrappers -- one rANS stream per level, the
state chained from one level into the next, buffers returned as Python lists of 32-bit words --
but the tensors never leave the GPU on the way in: each level is one single-stream call of the
CUDA coder with `init_states` carrying the chained state.

The reference's chaining has a latent bug: if a level's *first* symbol pushes
a word the decoder never pulls it back.  Decode reproduces the reference's behaviour exactly
(same pulls, same result) rather than fixing it; the live path resets the state per level
(trainer.py:310) and so does IDFlows.compress.
"""
import numpy as np
import torch

from . import rans


def _state_tensor(x: int, device) -> torch.Tensor:
    return torch.tensor([np.uint64(x).astype(np.int64)], dtype=torch.int64, device=device)


def Encode(latents, means, logscales, x=(1 << 32)):
    buffers = []
    for latent, mean, logscale in zip(latents, means, logscales):
        scale = torch.exp(logscale)
        enc = rans.encode_streams(latent.reshape(-1), mean.reshape(-1), scale.reshape(-1), None,
                                  init_states=_state_tensor(x, latent.device))
        enc.check()
        x = int(enc.final_states.cpu().numpy().view(np.uint64)[0])
        buffers.append(enc.words.cpu().numpy().view(np.uint32).tolist())
    return x, buffers


def Decode(buffers, means, logscales, x):
    latents = []
    for idx in reversed(range(len(means))):
        mean, logscale = means[idx], logscales[idx]
        dev = mean.device
        scale = torch.exp(logscale)
        words = torch.from_numpy(np.asarray(buffers[idx], dtype=np.uint32).view(np.int32).copy()).to(dev)
        enc = rans.EncodedStreams(words, torch.tensor([0, words.numel()], dtype=torch.int64, device=dev),
                                  _state_tensor(x, dev), torch.zeros(1, dtype=torch.int32, device=dev), mean.numel())
        z, end, status = rans.decode_streams(enc, mean.reshape(-1), scale.reshape(-1), None, check_end=False)
        rans.check_status(status)
        x = int(end.cpu().numpy().view(np.uint64)[0])
        latents.append(z.reshape(mean.shape).to(mean))
    return x, latents[::-1]
