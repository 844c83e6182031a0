"""The N>1 path on CPU: two gloo ranks shard a stream partition, code their shards (with the CPU
oracle standing in for the device, which is the checker's job in tests), exchange only the
per-rank totals, and together reproduce the single-rank result."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_images, per_image, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from _data import gen
    from flic_b200 import sharding
    from oracle import pyoracle
    r, w, _ = sharding.init_process_group("gloo")
    assert (r, w) == (rank, world)
    x, mean, scale = gen("test", n_images * per_image, 123)       # same data on every rank (seeded)
    lo, hi = sharding.shard_range(n_images, rank, world)
    sl = slice(lo * per_image, hi * per_image)
    off = np.arange(hi - lo + 1, dtype=np.int64) * per_image
    words, woff, states, status = pyoracle.encode_streams(x[sl], mean[sl], scale[sl], off)
    totals = sharding.gather_totals([int(words.size) * 4 + 8 * (hi - lo), (hi - lo) * per_image])
    t = sharding.max_over_ranks(1.0 + rank)
    sharding.barrier()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), words=words, states=states, totals=np.asarray(totals), t=t,
             lo=lo, hi=hi)
    dist.destroy_process_group()


def test_two_rank_sharding_reproduces_single_rank(tmp_path, oracle):
    from _data import gen
    n_images, per_image, world = 11, 384, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_images, per_image, str(tmp_path)), nprocs=world, join=True)
    x, mean, scale = gen("test", n_images * per_image, 123)
    off = np.arange(n_images + 1, dtype=np.int64) * per_image
    words, woff, states, _ = oracle.encode_streams(x, mean, scale, off)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert [(int(p["lo"]), int(p["hi"])) for p in parts] == [(0, 6), (6, 11)]
    assert np.array_equal(np.concatenate([p["words"] for p in parts]), words)       # shards concatenate
    assert np.array_equal(np.concatenate([p["states"] for p in parts]), states)
    for p in parts:                                                                 # every rank saw all totals
        assert p["totals"].shape == (2, 2)
        assert int(p["totals"][:, 1].sum()) == n_images * per_image
        assert int(p["totals"][:, 0].sum()) == words.size * 4 + 8 * n_images
        assert float(p["t"]) == 2.0                                                 # max over ranks
