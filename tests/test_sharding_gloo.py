"""Multi-process path on CPU: world_size-2 (and 3) gloo groups exercising the batch sharding + result gather that the
N>1 bench and ``process_sharded`` use.  The per-shard work is a stand-in checksum (the CUDA pipeline needs a GPU); what
is tested is the host-side partitioning, ragged/empty shards, ordering and the gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chessvision import sharding


def test_shard_ranges_tile_the_batch():
    for n in (0, 1, 2, 7, 8, 37, 65536):
        for world in (1, 2, 3, 4, 8):
            ranges = [sharding.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for (a, b), (c, d) in zip(ranges, ranges[1:]):
                assert b == c and a <= b
            sizes = sharding.shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1


def fake_pipeline(images):
    """Per-board outputs with the shapes/dtypes of cvb_outputs, derived deterministically from the image bytes."""
    n = len(images)
    a = torch.from_numpy(np.ascontiguousarray(images)).reshape(n, 8 * 8 * 3).to(torch.int64)
    s = a.sum(1)
    return {
        "found": (s % 2).to(torch.uint8),
        "quad": (s[:, None, None] + torch.arange(8).reshape(1, 4, 2)).to(torch.int32),
        "labels": ((s[:, None] + torch.arange(64)[None]) % 13).to(torch.uint8),
        "probs": (s[:, None, None].float() / 1e6).expand(n, 64, 13).contiguous(),
    }


def worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        images = rng.integers(0, 256, (n, 8, 8, 3), dtype=np.uint8)
        full = sharding.process_sharded(fake_pipeline, images)
        want = fake_pipeline(images)
        ok = all(torch.equal(full[k], want[k]) for k in want) and set(full) == set(want)
        lo, hi = sharding.shard_range(n, rank, world)
        q.put((rank, bool(ok), hi - lo))
    finally:
        dist.destroy_process_group()


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,n", [(2, 37), (2, 1), (2, 0), (3, 8)])
def test_sharded_equals_unsharded_gloo(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in results)
    assert sum(sz for _, _, sz in results) == n
