"""torch.distributed plumbing shared by the trainer and the CPU (gloo) tests: the embedding all-gather that makes
the InfoNCE loss global (north_star; concept of CT_CLIP/ct_clip/distributed.py:9-34, which is dead code in the
reference) and the row partition each rank back-propagates."""
from __future__ import annotations

import torch
import torch.distributed as dist


def gather_latents(t_raw: torch.Tensor, i_raw: torch.Tensor):
    """All ranks contribute b rows of text and image latents; returns the (W*b, L) global matrices ordered by rank.
    One packed message per rank ([text | image], 32 KB at b=8, L=512)."""
    world = dist.get_world_size()
    packed = torch.cat([t_raw, i_raw], dim=1).contiguous()
    out = torch.empty(world * packed.shape[0], packed.shape[1], device=packed.device, dtype=packed.dtype)
    try:
        dist.all_gather_into_tensor(out, packed)
    except (RuntimeError, NotImplementedError):      # backends without the tensor variant (older gloo)
        parts = [torch.empty_like(packed) for _ in range(world)]
        dist.all_gather(parts, packed)
        out = torch.cat(parts, dim=0)
    L = t_raw.shape[1]
    return out[:, :L].contiguous(), out[:, L:].contiguous()


def rank_rows(rank: int, b: int):
    """Rows of the global batch owned (and differentiated) by `rank`."""
    return rank * b, b
