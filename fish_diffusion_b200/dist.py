"""Batch sharding for 1..8 GPU inference (one process per GPU).

Both operators on the path are independent along the batch axis: rank r of G takes a
contiguous slice of items, weights are replicated, and there is NO collective on the data path -- only a barrier and
a max-over-ranks of the device time for measurement.  The reference has no inference data parallelism at all
(tools/diffusion/inference.py:133-160 is strictly B=1); this is the new caller the BASELINE configs ask for.
"""
from __future__ import annotations

import os

import torch


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced slice [lo, hi) of `n_items` for `rank` (first n_items % world ranks get one extra)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def item_seed(base_seed: int, global_item: int) -> int:
    """Per-item RNG stream id, a function of the GLOBAL item index so results do not depend on the world size."""
    return (int(base_seed) * 0x9E3779B97F4A7C15 + int(global_item) * 0xBF58476D1CE4E5B9) & (2 ** 62 - 1)


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend=None):
    """torch.distributed plumbing (nccl on GPUs, gloo on CPU tests); returns (rank, world, local_rank)."""
    rank, world, local = env_rank_world()
    if world > 1 and not torch.distributed.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        torch.distributed.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def max_over_ranks(value: float, device=None) -> float:
    """max of a host scalar over all ranks (used for device-timed measurements)."""
    if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def gather_batch(local: torch.Tensor, n_items: int):
    """all_gather of per-rank output slices back into the full batch (host-side convenience, not timed)."""
    if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
        return local
    world = torch.distributed.get_world_size()
    sizes = [shard_range(n_items, r, world) for r in range(world)]
    max_n = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((max_n,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    torch.distributed.all_gather(outs, pad)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, sizes)], dim=0)
