"""Batch sharding of the image->FEN path over the GPUs of one box.

Boards are independent units, so inference shards by contiguous batch ranges with NO collective on the data path: rank r
of G processes boards ``[r*N//G, (r+1)*N//G)`` on its own GPU with its own replica of the weights.  The only
communication is the optional gather of the small per-board results (~1.2 KB/board: quad, found, status, probabilities,
labels, FEN) onto one rank after the pipeline has finished; it runs through ``torch.distributed`` (NCCL on device
tensors, gloo on host tensors) and is not part of the timed hot path.

The reference has no multi-device code at all; this module is new.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous range of board indices owned by ``rank``; ranges differ in size by at most one board and tile [0, n)."""
    assert n >= 0 and world >= 1 and 0 <= rank < world
    return (rank * n) // world, ((rank + 1) * n) // world


def shard_sizes(n: int, world: int) -> list[int]:
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def gather_outputs(local: dict[str, torch.Tensor], n_total: int, group=None) -> dict[str, torch.Tensor]:
    """All-gather per-board outputs of every rank's shard into full ``[n_total, ...]`` tensors (same on every rank).

    ``local[k]`` has this rank's shard size as its first dimension.  Shards may be ragged (sizes differ by one) or empty,
    so every rank pads to the largest shard, one all_gather per output runs, and the padding is dropped on arrival.
    """
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(n_total, world)
    cap = max(max(sizes), 1)
    full = {}
    for key, t in local.items():
        assert t.shape[0] == sizes[rank], f"output '{key}': shard holds {t.shape[0]} boards, expected {sizes[rank]}"
        padded = torch.zeros((cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        padded[: t.shape[0]] = t
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded, group=group)
        full[key] = torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)
    return full


def process_sharded(run_shard, images, group=None) -> dict[str, torch.Tensor]:
    """Run ``run_shard(images[lo:hi]) -> dict of per-board tensors`` on this rank's range and gather the results.

    ``images`` is the full batch (every rank sees the same host array or its own view of a shared file); only the owned
    range is touched.  With a single process this is just ``run_shard(images)``.
    """
    n = len(images)
    if not dist.is_available() or not dist.is_initialized():
        return run_shard(images)
    lo, hi = shard_range(n, dist.get_rank(group), dist.get_world_size(group))
    return gather_outputs(run_shard(images[lo:hi]), n, group)
