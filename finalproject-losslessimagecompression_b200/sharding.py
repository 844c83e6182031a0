"""Sharding of independent images / streams across the GPUs of one box.

Images are independent at coding time (no batch statistics; weights and permutations are
read-only), so rank r of W codes the contiguous image range shard_range(n, r, W) with a full
replica of the model and no collective on the data path.  The only exchange is an optional
all_gather of per-rank (compressed_bytes, n_symbols) pairs for reporting.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) for `rank`; the first n_items % world ranks get one extra."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def env_rank_world() -> tuple[int, int, int]:
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)),
            int(os.environ.get("LOCAL_RANK", 0)))


def init_process_group(backend: str | None = None) -> tuple[int, int, int]:
    """torch.distributed over NCCL (GPU) or gloo (CPU tests); rendezvous from the torchrun env."""
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def gather_totals(values: list[int], device=None) -> list[list[int]]:
    """all_gather of a small int64 vector per rank -> one list per rank (rank order)."""
    t = torch.tensor(values, dtype=torch.int64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [t.tolist()]
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.tolist() for o in out]


def max_over_ranks(value: float, device=None) -> float:
    """Timing convention of the benchmark: a step takes as long as its slowest rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(device=None) -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if device is not None and torch.device(device).type == "cuda":
            dist.barrier(device_ids=[torch.device(device).index or 0])
        else:
            dist.barrier()
