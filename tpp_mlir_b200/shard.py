"""Batch sharding of the MLP harness across GPUs (one process per GPU).

The reference has no distributed code at all; the only way its hot path shards is over the
independent rows of the activation matrix: rank r owns rows [r*B/G, (r+1)*B/G), parameters are
replicated by ONE broadcast from rank 0 at setup, there is no collective inside the timed loop, and timing is the
max over ranks. Backend-agnostic torch.distributed (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, rank: int, world: int, tile_m: int = 1) -> tuple[int, int]:
    """Rows [lo, hi) of rank `rank`. Only clean splits are allowed (north_star: "only where the batch
    dimension splits cleanly"): global_batch % (world * tile_m) == 0."""
    if global_batch % (world * tile_m) != 0:
        raise ValueError(f"batch {global_batch} does not split cleanly over {world} ranks x tile {tile_m}")
    per = global_batch // world
    return rank * per, (rank + 1) * per


def broadcast_parameters(tensors, src: int = 0) -> None:
    """The one collective of this path: weights/biases (and the synthetic input) from rank `src`, once."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for t in tensors:
            # raw bytes: bf16 bits travel as int16 tensors, a dtype gloo does not broadcast
            dist.broadcast(t.view(torch.uint8) if t.is_contiguous() else t, src=src)


def max_over_ranks(value: float, device=None) -> float:
    """Slowest rank's time: the whole-job time of a weak-scaled step."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_rows(local: torch.Tensor, dst: int = 0):
    """Optional, outside the timed loop: collect every rank's output rows on `dst` for a parity check."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return local
    flat = local.contiguous().view(torch.uint8)
    out = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(out, flat)
    if dist.get_rank() != dst:
        return None
    return torch.cat([o.view(local.dtype).reshape(local.shape) for o in out], dim=0)
