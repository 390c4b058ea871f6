"""Multi-GPU plumbing of the sampler: one process per GPU, independent datasets sharded over ranks
with no data-path collective, and ONE all-gather of the final samples (the
reference is single-process, magi_v2.py:412-422 just stacks samples on the host).

`torch.distributed` is the transport (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `n_items` datasets owned by `rank`; sizes differ by at most one and all
    chains of a dataset stay on one rank (they share its kernel matrices)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n_items: int, world: int) -> List[int]:
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


def chain_id0(n_items: int, chains_per_item: int, rank: int, world: int) -> int:
    """Global id of this rank's first chain: keeps the Philox streams independent of the sharding."""
    return shard_range(n_items, rank, world)[0] * chains_per_item


def gather_samples(local: torch.Tensor, dataset_dim: int = 1, sizes: List[int] = None) -> torch.Tensor:
    """All-gather per-rank sample tensors along the dataset axis (samples are [n_iter, B_local, R, ...]).
    Equal shards use one `all_gather_into_tensor`; ragged shards (sizes given) pad to the largest."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    x = local.movedim(dataset_dim, 0).contiguous()
    if sizes is None or len(set(sizes)) == 1:
        out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x)
    else:
        mx = max(sizes)
        pad = torch.zeros((mx,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        pad[: x.shape[0]] = x
        buf = torch.empty((world * mx,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(buf, pad)
        out = torch.cat([buf[r * mx: r * mx + sizes[r]] for r in range(world)], dim=0)
    return out.movedim(0, dataset_dim)
