"""Sharding of a batch of independent problems over ranks.

Problems are independent (one `Optimizer_` per problem in the reference, no cross-problem term),
so the data path has NO collective: rank g owns the contiguous block [lo, hi) and, because the
synthetic inputs are a pure function of (seed, problem index), materialises its shard locally.
The only communication is the gather of the per-problem results / solutions at the end
(`torch.distributed.all_gather`, NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_range(B: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of ceil(B / world) problems per rank (last ranks may be short or empty)."""
    per = -(-B // world)
    lo = min(B, rank * per)
    return lo, min(B, lo + per)


def gather_rows(local: torch.Tensor, B: int, rank: int, world: int) -> torch.Tensor:
    """All-gather row blocks sharded by `shard_range` back into a [B, ...] tensor on every rank."""
    if world == 1:
        return local
    per = -(-B // world)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat(out, dim=0)[:B]


def gather_results(local: np.ndarray, B: int, rank: int, world: int, device) -> np.ndarray:
    """Same for the structured tob200_result array (moved as bytes)."""
    if world == 1:
        return local
    raw = torch.from_numpy(local.view(np.uint8).reshape(local.shape[0], local.dtype.itemsize).copy()).to(device)
    full = gather_rows(raw, B, rank, world)
    return full.cpu().numpy().view(local.dtype).reshape(-1)
