"""MC-sample sharding across ranks: rank r of G owns the contiguous block of global sample ids
[r*S/G, (r+1)*S/G); eps is keyed by the GLOBAL id, every rank adds the (sample-independent) T*dKL term itself, and
one all-reduce AVERAGE of the flat gradient per step gives d/dtheta [ mean_s nll_s + T*KL ] on every rank."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_samples(mc_samples: int, rank: int, world_size: int):
    """(local sample count, global id of local sample 0)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    if mc_samples % world_size:
        raise ValueError(f"mc_samples={mc_samples} must be divisible by world_size={world_size}")
    s_local = mc_samples // world_size
    return s_local, rank * s_local


def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place average over ranks: ReduceOp.AVG on NCCL (one kernel, CUDA-graph capturable); SUM + scale on
    backends without AVG (gloo, used by the CPU tests)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return flat
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(dist.get_world_size(group))
    return flat
