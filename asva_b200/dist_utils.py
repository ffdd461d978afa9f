"""Multi-GPU plumbing of the clip-sharded path: clips are independent units, so ranks never
exchange data inside the denoising loop.  What is shared: which clips a rank owns, the max-over-ranks step time the
bench reports, and one all-gather of the finished latents."""
from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_items(items: Sequence, rank: int, world_size: int) -> List:
    """Round-robin ownership: rank r takes items r, r+W, r+2W, ... (every item exactly once)."""
    return list(items)[rank::world_size]


def max_over_ranks(value: float, device) -> float:
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def gather_latents(latents: torch.Tensor) -> torch.Tensor:
    """(1,C,F,h,w) per rank -> (W,C,F,h,w) on every rank (NCCL over NVLink on GPUs; gloo in the CPU tests)."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return latents.clone()
    out = [torch.empty_like(latents) for _ in range(dist.get_world_size())]
    dist.all_gather(out, latents.contiguous())
    return torch.cat(out, dim=0)
