"""IDFlows / ConditionalFlows with real compress / decompress (reference: flows.py:24-181, 277-361).

What is mirrored (same names, arguments and parameter layout, so the reference's YAML model
blocks and checkpoints work unchanged):
  IDFlows.__init__            flows.py:25-85     level structure, latents_shape
  IDFlows.forward             flows.py:87-116    -> (latents, means, logscales, logv)
  IDFlows.generated_from_*    flows.py:118-152
  IDFlows.log_likelihood      flows.py:154-169   ideal code length (bits/dim oracle)
  ConditionalFlows            flows.py:277-361   prior additionally sees a conditioning image
What is new (the reference's IDFlows.encode / .decode are empty stubs, flows.py:177-181, and its
only live call site, trainer.py:304-329, decodes with the means/logscales it already has):
  compress(images_u8)   -> CompressedBatch      flow forward + prior + rANS encode, level by level
  decompress(batch)     -> images_u8            decode last level, run the prior on what is known,
                                                decode the next level, ... exactly the control flow
                                                of generated_from_noise (flows.py:118-137)
encode / decode are aliases of these.  Everything elementwise on this path is a CUDA kernel of
libflic_b200.so; the DenseBlock convolutions are PyTorch fp32 (deterministic, TF32 off).
"""
from __future__ import annotations

import contextlib
import math
from copy import deepcopy

import torch
from torch import nn

from . import _lib, rans
from .container import CompressedBatch
from .couplelib import AdditiveCouple, NNCouple
from .distlib import NNDistribution
from .extenddim import NNExtendDim
from .invertible import InvertibleModuleList, Permute
from .moduleregister import Register
from .priorlib import NNPrior
from .roundlib import NNRound


class NNFlows(Register):
    pass


@contextlib.contextmanager
def deterministic_convs():
    """Compress and decompress must see bit-identical network outputs:
    deterministic cuDNN algorithms, no autotuning, no TF32, and no gradient bookkeeping."""
    cd, cb = torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark
    t1, t2 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            yield
    finally:
        torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = cd, cb
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = t1, t2


def u8_to_grid(img: torch.Tensor) -> torch.Tensor:
    """uint8 pixels -> the reference's input grid, rint(k/255*256)/256 (trainer.py:61,72)."""
    if not img.is_cuda or img.dtype != torch.uint8:
        raise TypeError("u8_to_grid expects a CUDA uint8 tensor")
    img = img.contiguous()
    out = torch.empty(img.shape, dtype=torch.float32, device=img.device)
    with torch.cuda.device(img.device):
        _lib.check(_lib.lib().flic_u8_to_grid(img.data_ptr(), out.data_ptr(), img.numel(),
                                              torch.cuda.current_stream(img.device).cuda_stream), "flic_u8_to_grid")
    return out


def grid_to_u8(x: torch.Tensor):
    """Inverse of u8_to_grid; returns (uint8 tensor, status word tensor)."""
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    status = torch.zeros(1, dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().flic_grid_to_u8(x.data_ptr(), out.data_ptr(), x.numel(), status.data_ptr(),
                                              torch.cuda.current_stream(x.device).cuda_stream), "flic_grid_to_u8")
    return out, status


@NNFlows.register
class IDFlows(nn.Module):
    def __init__(self, nflows=8, nbits=8, nsplit=3, H=64, W=64, C=3, couple=None, extenddim=None, prior=None,
                 distribution=None, round=None, batch_squeeze=0):
        super().__init__()
        if batch_squeeze:
            raise NotImplementedError("batch_squeeze (flows.py:93-98) mixes images into one sample; not on the coding path")
        couple, extenddim, prior = deepcopy(couple), deepcopy(extenddim), deepcopy(prior)
        distribution, round = deepcopy(distribution), deepcopy(round)
        self.nflows, self.nbits, self.nsplit = nflows, nbits, nsplit
        self.C, self.H, self.W = C, H, W
        self.batch_squeeze = 0
        self.couple_type = NNCouple.get(couple.pop("name"))
        self.prior_type = NNPrior.get(prior.pop("name"))
        self.extenddim_type = NNExtendDim.get(extenddim.pop("name"))
        self.dist_type = NNDistribution.get(distribution.pop("name"))
        self.round_type = NNRound.get(round.pop("name"))
        self._prior_cfg = deepcopy(prior)
        self.blocks = nn.ModuleList()
        self.latents_shape = []
        channel, h, w = C, H, W
        s = extenddim.get("scale")
        for level in range(nsplit):
            channel, h, w = channel * s * s, h // s, w // s
            flow = InvertibleModuleList()
            for _ in range(nflows):   # construction order = RNG order of flows.py:68-71
                flow.append(Permute(dim=channel))
                flow.append(self.couple_type(channel=channel, **deepcopy(couple)))
            flow.append(Permute(dim=channel))
            if level < nsplit - 1:
                prior_nn = self.prior_type(channel // 2, channel - channel // 2, **deepcopy(prior))
                self.latents_shape.append((channel // 2, h, w))
                channel -= channel // 2
            else:
                prior_nn = self.prior_type(channel, 0, **deepcopy(prior))
                self.latents_shape.append((channel, h, w))
            self.blocks.append(nn.ModuleDict(dict(extend=self.extenddim_type(**deepcopy(extenddim)),
                                                  flows=flow, prior=prior_nn)))
        self.dist = self.dist_type(**distribution)
        self.round = self.round_type(**round)

    # ---- the reference's own entry points ------------------------------------------------------
    def _prior_input(self, level, x_rest, cond):
        """What block['prior'] is fed at this level (flows.py:106,112; ConditionalFlows :317,323)."""
        return x_rest

    def _flow_forward(self, block, x):
        for m in block["flows"]:
            if isinstance(m, AdditiveCouple):
                x, _ = m.forward(x, None, inplace=True)   # x is the fresh output of the Permute before it
            else:
                x, _ = m.forward(x, None)
        return x

    def _flow_backward(self, block, x):
        mods = list(block["flows"])
        fresh = False
        for m in reversed(mods):
            if isinstance(m, AdditiveCouple):
                x = m.backward(x, inplace=fresh)
            else:
                x = m.backward(x)
                fresh = True
        return x

    @torch.no_grad()
    def forward(self, x, logv=None, cond=None):
        latents, means, logscales = [], [], []
        for level in range(self.nsplit):
            block = self.blocks[level]
            x, _ = block["extend"](x, None)
            cond = self._cond_step(level, cond)
            x = self._flow_forward(block, x)
            if level < self.nsplit - 1:
                half = x.shape[1] // 2
                z, x = x[:, :half], x[:, half:]
                mean, logscale = block["prior"](self._prior_input(level, x, cond))
            else:
                z = x
                mean, logscale = block["prior"](self._prior_input(level, x, cond))
            latents.append(z)
            means.append(mean)
            logscales.append(logscale)
        return latents, means, logscales, logv

    def _cond_step(self, level, cond):
        return cond

    @torch.no_grad()
    def generated_from_latents(self, latents):
        x = None
        for level in reversed(range(self.nsplit)):
            block = self.blocks[level]
            z = latents[level]
            x = z if level == self.nsplit - 1 else torch.cat((z, x), dim=1)
            x = self._flow_backward(block, x.contiguous())
            x = block["extend"].backward(x)
        return x

    @torch.no_grad()
    def generated_from_noise(self, latents):
        x = None
        for level in reversed(range(self.nsplit)):
            block = self.blocks[level]
            z = latents[level]
            mean, logscale = block["prior"](x if level < self.nsplit - 1 else z)
            z = self.round(z * torch.exp(logscale) + mean)
            x = z if level == self.nsplit - 1 else torch.cat((z, x), dim=1)
            x = self._flow_backward(block, x.contiguous())
            x = block["extend"].backward(x)
        return x

    def log_likelihood(self, latents, means, logscales):
        """flows.py:154-169.  On CUDA without autograd each level is one fused kernel
        (evaluation + per-image sum, DLogistic.log_prob_sums); the per-level means are the sums
        divided by the level's element count."""
        log_Ps = []
        log_prob = torch.zeros(latents[0].shape[0], device=latents[0].device)
        for z, mean, logscale in zip(latents, means, logscales):
            if hasattr(self.dist, "log_prob_sums"):
                sums = self.dist.log_prob_sums(z, mean, logscale, self.nbits)
                log_Ps.append(sums / (z.numel() // z.shape[0]))
            else:
                logp = self.dist.log_prob(z, mean, logscale, self.nbits)
                log_Ps.append(torch.mean(logp, dim=(1, 2, 3)))
                sums = torch.sum(logp, dim=(1, 2, 3))
            log_prob = log_prob + sums
        return log_prob / (self.H * self.W * self.C), log_Ps

    def inverse(self):
        for block in self.blocks:
            block["extend"].inverse()
            block["flows"].inverse()

    # ---- compress / decompress -------------------------------------------------------------------
    def _segment_offsets(self, level: int, n_img: int, streams_per_segment: int, device):
        """Stream partition of one level of one chunk.  A level's symbols are the row-major
        flattening of (B, C_z, H, W) (trainer.py:311), image b owning a contiguous segment.
        streams_per_segment k >= 1: each image segment is cut into k near-equal streams;
        0: the reference's native partition, one stream for the whole level of the chunk."""
        c, h, w = self.latents_shape[level]
        seg = c * h * w
        if streams_per_segment == 0:
            return torch.tensor([0, n_img * seg], dtype=torch.int64, device=device)
        k = streams_per_segment
        cuts = torch.tensor([seg * j // k for j in range(k)], dtype=torch.int64, device=device)
        base = torch.arange(n_img, dtype=torch.int64, device=device) * seg
        off = (base[:, None] + cuts[None, :]).reshape(-1)
        return torch.cat([off, torch.tensor([n_img * seg], dtype=torch.int64, device=device)])

    def _chunk_forward(self, x, cond, n_real, sps, stats, chain, slot):
        """x: (codec_batch, C, H, W) grid floats -> list of EncodedStreams: one per level, or a single
        one holding each image's levels chained (coder.py:18-27).  Nothing here waits for the GPU:
        the word arrays keep their worst-case capacity until CompressedBatch.finalize() trims them."""
        out = []
        carried = None
        for level in range(self.nsplit):
            block = self.blocks[level]
            x, _ = block["extend"](x, None)
            cond = self._cond_step(level, cond)
            x = self._flow_forward(block, x)
            if level < self.nsplit - 1:
                half = x.shape[1] // 2
                z, rest = x[:, :half].contiguous(), x[:, half:].contiguous()
                mean, logscale = block["prior"](self._prior_input(level, rest, cond))
                x = rest
            else:
                z = x
                mean, logscale = block["prior"](self._prior_input(level, x, cond))
            mean, scale = mean.contiguous(), torch.exp(logscale.contiguous())
            if stats is not None:
                stats.append((z, mean, logscale))
            off = self._segment_offsets(level, n_real, sps, x.device)
            n = int(n_real * math.prod(self.latents_shape[level]))
            ws = self.__dict__.setdefault("_enc_ws", {}).setdefault((x.device.index, slot, level), rans.Workspace())
            enc = rans.encode_streams(z.reshape(-1)[:n], mean.reshape(-1)[:n], scale.reshape(-1)[:n], off,
                                      init_states=carried, workspace=ws, own_output=False, validate=False)
            if chain:
                carried = enc.final_states          # the next level continues every image's stream
            else:
                enc.words = enc.words[:max(n, 1)].clone()   # the workspace is reused by the next chunk of this slot
            out.append(enc)
        return [rans.chain_levels(out)] if chain else out

    def _pipes(self, dev, n_chunks: int, pipeline: int):
        """Side CUDA streams for chunk-level pipelining, or [None] for in-line.
        The streams are kept on the model: the caching allocator pools memory per stream, so fresh
        streams on every call would turn every activation into a cudaMalloc."""
        k = min(int(pipeline), n_chunks)
        if k <= 1:
            return [None]
        cache = self.__dict__.setdefault("_side_streams", {})
        have = cache.setdefault(dev.index, [])
        while len(have) < k:
            have.append(torch.cuda.Stream(dev))
        return have[:k]

    def compress(self, images: torch.Tensor, cond: torch.Tensor | None = None, codec_batch: int | None = None,
                 streams_per_segment: int = 1, check: bool = True, stats: list | None = None,
                 pipeline: int = 2, chain_levels: bool | None = None) -> CompressedBatch:
        """Lossless compression of uint8 images (N, C, H, W) on the GPU.

        codec_batch: images per network pass (default: all of them).  It is recorded in the
        result because the decoder has to run the networks on batches of the same shape to get
        bit-identical prior outputs; the last chunk is padded with zero images whose streams are
        not stored.  streams_per_segment: rANS streams per (image, level); 0 selects the
        reference's native partition (one stream per level per chunk, trainer.py:308-315).
        chain_levels: carry each image's rANS state from one latent level into the next, as the
        reference's coder.Encode does (coder.py:18-27): one stream -- one 64-bit final state --
        per image instead of one per image and level, which is what keeps the container within
        0.1 % of the reference's native partition (one stream per level per batch).  Default: on
        when streams_per_segment == 1.
        pipeline: chunks in flight.  Chunks are independent, so chunk i runs on CUDA stream
        i % pipeline: the coder kernels of one chunk (a few dozen serial streams, latency-bound,
        a handful of warps) overlap the convolutions of the next one instead of idling the GPU
        between them.  The bytes produced do not depend on it."""
        if images.dtype != torch.uint8 or images.dim() != 4:
            raise TypeError("images must be uint8 (N, C, H, W)")
        return self._compress(images, True, cond, codec_batch, streams_per_segment, check, stats, pipeline, chain_levels)

    def compress_grid(self, x: torch.Tensor, cond: torch.Tensor | None = None, codec_batch: int | None = None,
                      streams_per_segment: int = 1, check: bool = True, stats: list | None = None,
                      pipeline: int = 2, chain_levels: bool | None = None) -> CompressedBatch:
        """compress() for inputs that are already on the 2^-nbits grid as float32 (N, C, H, W) --
        residuals and pooled images of the two-level model (flows.py:212-214) are such tensors and
        are not 8-bit pixel values."""
        if x.dtype != torch.float32 or x.dim() != 4:
            raise TypeError("x must be float32 (N, C, H, W) with values on the 2^-nbits grid")
        return self._compress(x, False, cond, codec_batch, streams_per_segment, check, stats, pipeline, chain_levels)

    def _compress(self, images, from_u8, cond, codec_batch, streams_per_segment, check, stats, pipeline, chain_levels=None):
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise _lib.FlicError("compress needs the model on a CUDA device (no CPU fallback)")
        images = images.to(dev, non_blocking=True)
        n = images.shape[0]
        cbs = int(codec_batch or max(n, 1))
        chain = (int(streams_per_segment) == 1) if chain_levels is None else bool(chain_levels)
        if chain and int(streams_per_segment) != 1:
            raise ValueError("chain_levels needs one stream per image (streams_per_segment == 1)")
        batch = CompressedBatch(n, (self.C, self.H, self.W), self.nsplit, cbs, int(streams_per_segment), chained=chain)
        if cond is not None:
            cond = cond.to(dev)
        with deterministic_convs(), torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            pipes = self._pipes(dev, (n + cbs - 1) // cbs, pipeline)
            for ci, i0 in enumerate(range(0, n, cbs)):
                side = pipes[ci % len(pipes)]
                if side is not None and ci < len(pipes):
                    side.wait_stream(main)                      # inputs were produced on the main stream
                with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
                    chunk = images[i0:i0 + cbs]
                    n_real = chunk.shape[0]
                    x = u8_to_grid(chunk) if from_u8 else chunk.contiguous().clone()
                    c = None if cond is None else cond[i0:i0 + cbs]
                    if n_real < cbs:
                        x = torch.cat([x, x.new_zeros((cbs - n_real,) + tuple(x.shape[1:]))])
                        if c is not None:
                            c = torch.cat([c, c.new_zeros((cbs - n_real,) + tuple(c.shape[1:]))])
                    section = self._chunk_forward(x, c, n_real, int(streams_per_segment), stats, chain, ci % len(pipes))
                    if side is not None:
                        for e in section:
                            e.record_stream(main)
                    batch.sections.append(section)
                # word arrays keep their worst-case size until trimmed, which needs the word counts on
                # the host: trim the chunk two rounds back, whose kernels have long finished
                if ci >= 2 * len(pipes):
                    batch.finalize(ci - 2 * len(pipes))
            for side in pipes:
                if side is not None:
                    main.wait_stream(side)
        batch.finalize()
        if check:
            for ch in batch.sections:
                for e in ch:
                    e.check()
        return batch

    def decompress(self, batch, cond: torch.Tensor | None = None, check: bool = True, pipeline: int = 2) -> torch.Tensor:
        """Inverse of compress: CompressedBatch (or its bytes) -> uint8 images (N, C, H, W).

        Inside a chunk the levels are a strict chain (prior convolutions -> rANS decode -> inverse
        flow -> next level's prior), and the decode of a chunk is a few dozen serial streams; with
        pipeline > 1 chunk i runs on CUDA stream i % pipeline, so that chain of one chunk overlaps
        the convolutions of another (level-pipelined decompress)."""
        return self._decompress(batch, True, cond, check, pipeline)

    def decompress_grid(self, batch, cond: torch.Tensor | None = None, check: bool = True, pipeline: int = 2) -> torch.Tensor:
        """Inverse of compress_grid: float32 (N, C, H, W) on the 2^-nbits grid."""
        return self._decompress(batch, False, cond, check, pipeline)

    def _decompress(self, batch, to_u8, cond, check, pipeline):
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise _lib.FlicError("decompress needs the model on a CUDA device (no CPU fallback)")
        if isinstance(batch, (bytes, bytearray, memoryview)):
            batch = CompressedBatch.from_bytes(bytes(batch), dev)
        if tuple(batch.shape) != (self.C, self.H, self.W) or batch.n_levels != self.nsplit:
            raise ValueError("container does not match this model")
        n, cbs, sps = batch.n_images, batch.codec_batch, batch.streams_per_segment
        outs, statuses = [], []
        if cond is not None:
            cond = cond.to(dev)
        with deterministic_convs(), torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            pipes = self._pipes(dev, (n + cbs - 1) // cbs if cbs else 0, pipeline)
            for ci, i0 in enumerate(range(0, n, cbs)):
                side = pipes[ci % len(pipes)]
                if side is not None and ci < len(pipes):
                    side.wait_stream(main)
                with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
                    n_real = min(cbs, n - i0)
                    c = None if cond is None else cond[i0:i0 + cbs]
                    if c is not None and n_real < cbs:
                        c = torch.cat([c, c.new_zeros((cbs - n_real,) + tuple(c.shape[1:]))])
                    conds = self._cond_pyramid(c)
                    x = None
                    carried = left = None
                    for level in reversed(range(self.nsplit)):
                        block = self.blocks[level]
                        cz, h, w = self.latents_shape[level]
                        if level == self.nsplit - 1:
                            probe = torch.zeros((cbs, cz, h, w), dtype=torch.float32, device=dev)
                            mean, logscale = block["prior"](self._prior_input(level, probe, conds[level]))
                        else:
                            mean, logscale = block["prior"](self._prior_input(level, x, conds[level]))
                        mean, scale = mean.contiguous(), torch.exp(logscale.contiguous())
                        nsym = n_real * cz * h * w
                        off = self._segment_offsets(level, n_real, sps, dev)
                        z = torch.zeros((cbs, cz, h, w), dtype=torch.float32, device=dev)
                        if batch.chained:
                            # every image's one stream, continued: this level starts from the state and
                            # the unread words the later level's decode stopped at (coder.py:29-38)
                            _, carried, st, left = rans.decode_streams(
                                batch.sections[ci][0], mean.reshape(-1)[:nsym], scale.reshape(-1)[:nsym], off,
                                out=z.view(-1)[:nsym], check_end=level == 0, validate=False,
                                states=carried, words_left=left, return_words_left=True)
                        else:
                            _, _, st = rans.decode_streams(batch.sections[ci][level], mean.reshape(-1)[:nsym],
                                                           scale.reshape(-1)[:nsym], off, out=z.view(-1)[:nsym],
                                                           validate=False)
                        statuses.append(st)
                        x = z if level == self.nsplit - 1 else torch.cat((z, x), dim=1)
                        x = self._flow_backward(block, x)
                        x = block["extend"].backward(x)
                    if to_u8:
                        img, st8 = grid_to_u8(x[:n_real])
                        statuses.append(st8)
                    else:
                        img = x[:n_real].contiguous()
                    outs.append(img)
                    if side is not None:
                        for t in statuses[-(self.nsplit + (1 if to_u8 else 0)):]:
                            t.record_stream(main)
                        img.record_stream(main)
            for side in pipes:
                if side is not None:
                    main.wait_stream(side)
        if check:
            for st in statuses:
                rans.check_status(st)
        if outs:
            return torch.cat(outs)
        return torch.empty((0, self.C, self.H, self.W), dtype=torch.uint8 if to_u8 else torch.float32, device=dev)

    def _cond_pyramid(self, cond):
        return [None] * self.nsplit

    # the reference's stubs (flows.py:177-181), now real
    def encode(self, x, **kw):
        return self.compress(x, **kw)

    def decode(self, x, **kw):
        return self.decompress(x, **kw)


@NNFlows.register
class ConditionalFlows(IDFlows):
    """IDFlows whose priors also see a conditioning image `cond` (e.g. a VQ-VAE reconstruction,
    trainer.py:606-621), squeezed alongside x (flows.py:311) or passed through strided convs
    (flows.py:313).  The coder is unchanged; only the prior's input grows by `ch` channels."""

    def __init__(self, conv_for_cond=False, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.conv_for_cond = conv_for_cond
        if conv_for_cond:
            self.convs = nn.ModuleList()
        ch = self.C
        for level in range(self.nsplit):
            block = self.blocks[level]
            s = block["extend"].scale
            ch *= s * s
            old = block["prior"]
            block["prior"] = self.prior_type(
                old.out_channel, (old.cond_channel if old.cond_channel > 0 else old.out_channel) + ch,
                **deepcopy(self._prior_cfg))
            if conv_for_cond:
                self.convs.append(nn.Conv2d(ch // s // s, ch, 4, 2, 1))

    def _cond_step(self, level, cond):
        if cond is None:
            raise ValueError("ConditionalFlows needs cond")
        if self.conv_for_cond:
            return self.convs[level](cond)
        return self.blocks[level]["extend"](cond, None)[0]

    def _prior_input(self, level, x_rest, cond):
        if level == self.nsplit - 1:
            x_rest = torch.zeros_like(x_rest)        # flows.py:323
        return torch.cat((x_rest, cond), dim=1)

    def _cond_pyramid(self, cond):
        out = []
        for level in range(self.nsplit):
            cond = self._cond_step(level, cond)
            out.append(cond)
        return out

    @torch.no_grad()
    def generated_from_noise(self, latents, cond):
        conds = self._cond_pyramid(cond)
        x = None
        for level in reversed(range(self.nsplit)):
            block = self.blocks[level]
            z = latents[level]
            mean, logscale = block["prior"](self._prior_input(level, x if level < self.nsplit - 1 else z, conds[level]))
            z = self.round(z * torch.exp(logscale) + mean)
            x = z if level == self.nsplit - 1 else torch.cat((z, x), dim=1)
            x = self._flow_backward(block, x.contiguous())
            x = block["extend"].backward(x)
        return x


@NNFlows.register
class TwoLevelFlows(nn.Module):
    """The reference's two-level model (flows.py:184-274, configs/config_twolevel.yaml): the padded
    image is average-pooled to a rough image rx = Round(pool(x)) coded by `rough`, and the residual
    fx = x - upsample(rx), cut into fine.H x fine.W patches (extenddim.Patching), is coded by `fine`.
    Same constructor, sub-module names (`fine`, `rough`) and latents_shape as the reference, so its
    checkpoints load.  The reference only trains this wrapper (its forward calls loss.backward and
    has a typo at :242); here forward is the inference mirror and compress / decompress are real:
    both parts are multiples of 1/256, the pooling sums are exact in float32 and the upsampling is
    replication (216 = 8 x 27, 184 = 8 x 23), so x = upsample(rx) + fx holds bit for bit."""

    MAGIC = b"FL2L"

    def __init__(self, H, W, C, pad, fine_flows, rough_flows, batchsize=256, nbits=8):
        super().__init__()
        from .extenddim import Patching
        from .roundlib import Round
        fine_flows, rough_flows = deepcopy(fine_flows), deepcopy(rough_flows)
        self.H, self.W, self.C = H + pad[0], W + pad[1], C
        self.pad = list(pad)
        self.pad2d = nn.ReplicationPad2d(padding=(0, pad[1], 0, pad[0]))
        self.fine = NNFlows.get(fine_flows.pop("name"))(**fine_flows)        # construction order of flows.py:198-199
        self.rough = NNFlows.get(rough_flows.pop("name"))(**rough_flows)
        self.pool = nn.AdaptiveAvgPool2d((self.rough.H, self.rough.W))
        self.invpool = nn.AdaptiveAvgPool2d((self.H, self.W))
        self.patching = Patching(self.H, self.W, self.fine.H, self.fine.W)
        self.ratio = (self.H // self.fine.H, self.W // self.fine.W)
        fs = self.fine.latents_shape[0]
        self.latents_shape = [tuple(self.rough.latents_shape[0]), (fs[0] * self.ratio[0] * self.ratio[1], fs[1], fs[2])]
        self.batchsize = batchsize
        self.round = Round(nbits=nbits)
        if self.H % self.rough.H or self.W % self.rough.W:
            raise ValueError("padded size must be a multiple of the rough size (exact pooling / replication)")

    # ---- the split x -> (rx, fine patches) and its inverse ------------------------------------
    def split(self, x: torch.Tensor):
        """x: (B, C, H - pad0, W - pad1) grid floats -> rx (B, C, rough.H, rough.W), patches (B * n, C, h, w)."""
        x = self.pad2d(x)
        rx = self.round(self.pool(x))                     # flows.py:212
        fx = x - self.invpool(rx)                         # flows.py:213, exact on the grid
        px, _ = self.patching(fx, None)
        return rx.contiguous(), px

    def merge(self, rx: torch.Tensor, px: torch.Tensor) -> torch.Tensor:
        fx = self.patching.backward(px)
        x = self.invpool(rx) + fx                         # flows.py:261
        return x[:, :, :x.shape[2] - self.pad[0], :x.shape[3] - self.pad[1]].contiguous()

    @torch.no_grad()
    def forward(self, x, logv=None, train=False):
        """Inference mirror of flows.py:206-246: latents / means / logscales of both levels and
        the ideal bits per dimension (total, rough, fine)."""
        if train:
            raise NotImplementedError("training is outside the coding path")
        rx, px = self.split(x)
        rl, rm, rs, logv = self.rough.forward(rx, logv)
        logp, _ = self.rough.log_likelihood(rl, rm, rs)
        bpd1 = float(torch.mean(-logp)) / math.log(2)
        fl, fm, fs, bpd2 = [], [], [], 0.0
        for i0 in range(0, px.shape[0], self.batchsize):
            l, m, sc, logv = self.fine.forward(px[i0:i0 + self.batchsize], logv)
            logp, _ = self.fine.log_likelihood(l, m, sc)
            bpd2 += float(torch.sum(-logp)) / math.log(2)
            fl.append(l[0]); fm.append(m[0]); fs.append(sc[0])
        bpd2 /= px.shape[0]
        bpd = ((bpd1 * self.rough.H * self.rough.W + bpd2 * self.H * self.W)
               / (self.H - self.pad[0]) / (self.W - self.pad[1]))                       # flows.py:240
        return ([rl[0], torch.cat(fl)], [rm[0], torch.cat(fm)], [rs[0], torch.cat(fs)], bpd, bpd1, bpd2, logv)

    def inverse(self):
        self.fine.inverse()
        self.rough.inverse()

    # ---- compress / decompress ---------------------------------------------------------------------
    def compress(self, images: torch.Tensor, rough_batch: int | None = None, check: bool = True) -> bytes:
        """uint8 (N, C, H - pad0, W - pad1) -> bytes: the rough container followed by the fine one
        (one rANS stream per rough image and one per fine patch)."""
        if images.dtype != torch.uint8 or images.dim() != 4:
            raise TypeError("images must be uint8 (N, C, H, W)")
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise _lib.FlicError("compress needs the model on a CUDA device (no CPU fallback)")
        rx, px = self.split(u8_to_grid(images.to(dev)))
        rough = self.rough.compress_grid(rx, codec_batch=rough_batch or rx.shape[0], check=check).to_bytes()
        fine = self.fine.compress_grid(px, codec_batch=self.batchsize, check=check).to_bytes()
        return self.MAGIC + len(rough).to_bytes(8, "little") + rough + fine

    def decompress(self, blob: bytes, check: bool = True) -> torch.Tensor:
        if bytes(blob[:4]) != self.MAGIC or len(blob) < 12:
            raise ValueError("not a two-level container")
        n_rough = int.from_bytes(blob[4:12], "little")
        if 12 + n_rough > len(blob):
            raise ValueError("truncated two-level container")
        rx = self.rough.decompress_grid(bytes(blob[12:12 + n_rough]), check=check)
        px = self.fine.decompress_grid(bytes(blob[12 + n_rough:]), check=check)
        img, status = grid_to_u8(self.merge(rx, px))
        if check:
            rans.check_status(status)
        return img


def build_model(cfg: dict) -> nn.Module:
    """`train.model` block of a reference YAML config -> model (trainer.py:204)."""
    cfg = deepcopy(cfg)
    cfg.pop("load_path", None)
    return NNFlows.get(cfg.pop("name"))(**cfg)


def perturb_heads(model: nn.Module, std: float = 0.02, seed: int = 0) -> None:
    """The DenseBlock heads are zero-initialised (nnblock.py:50-51), which makes every coupling
    the identity and every prior mean=0, scale=1.  For tests and benchmarks on random-init
    models, re-draw them N(0, std) in module order (KAT-flow)."""
    from .nnblock import DenseBlock
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, DenseBlock):
                head = m.layers[-1]
                head.weight.copy_(torch.randn(head.weight.shape, generator=g) * std)
                head.bias.copy_(torch.randn(head.bias.shape, generator=g) * std)


def smoke_round_trip() -> None:
    """Tiny IDFlows on cuda:0: compress -> bytes -> decompress must return the pixels."""
    import random
    torch.manual_seed(0)
    random.seed(0)
    layer = dict(name="DenseLayer", act="LeakyReLU")
    cfg = dict(name="IDFlows", nflows=2, nbits=8, nsplit=2, H=16, W=16, C=3,
               couple=dict(name="AdditiveCouple", split=0.75, round=dict(name="Round", nbits=8),
                           nn=dict(name="DenseBlock", growth_channel=16, depth=2, layer=layer)),
               extenddim=dict(name="ExtendDim", scale=2),
               prior=dict(name="Prior", round=dict(name="Round", nbits=8),
                          nn=dict(name="DenseBlock", growth_channel=16, depth=2, layer=layer)),
               distribution=dict(name="DLogistic"), round=dict(name="Round", nbits=8))
    model = build_model(cfg)
    perturb_heads(model, 0.02)
    model = model.cuda().eval()
    img = torch.randint(0, 256, (5, 3, 16, 16), dtype=torch.uint8, generator=torch.Generator().manual_seed(1)).cuda()
    blob = model.compress(img, codec_batch=4).to_bytes()
    rec = model.decompress(blob)
    assert torch.equal(rec, img), "flow round trip is not lossless"
