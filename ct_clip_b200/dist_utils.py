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


class PeerLatentExchange:
    """The embedding exchange in front of the global InfoNCE as ONE hand-written kernel over NVLink peer memory
    (csrc/loss_optim.cu: latent_exchange_kernel) instead of cat + NCCL all-gather + two slicing copies: every rank pushes its
    b rows straight into the gather buffers of all peers and synchronises through release/acquire flags.
    torch's symmetric memory only provides the plumbing (allocation + mapping of the peers' buffers); if it is unavailable on
    the box (no P2P between the GPUs, old driver) the setup raises and the caller keeps the NCCL path."""

    def __init__(self, device, group=None):
        self.dev, self.group = device, group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.key, self.step = None, 0

    def _setup(self, b, L):
        import torch.distributed._symmetric_memory as symm_mem
        n = 2 * 2 * self.world * b * L
        self.buf = symm_mem.empty(n, dtype=torch.float32, device=self.dev)
        self.flags = symm_mem.empty(2 * self.world, dtype=torch.int32, device=self.dev)
        self.flags.zero_()
        torch.cuda.synchronize(self.dev)
        hb = symm_mem.rendezvous(self.buf, self.group)
        hf = symm_mem.rendezvous(self.flags, self.group)
        self.peer_bufs = torch.tensor([int(p) for p in hb.buffer_ptrs], dtype=torch.int64, device=self.dev)
        self.peer_flags = torch.tensor([int(p) for p in hf.buffer_ptrs], dtype=torch.int64, device=self.dev)
        torch.cuda.synchronize(self.dev)
        dist.barrier(self.group)          # every rank has zeroed its flags before anyone publishes step 1
        self.key, self.step = (b, L), 0

    def __call__(self, t_raw, i_raw):
        from . import ops
        b, L = t_raw.shape
        if self.key is None:
            self._setup(b, L)
        if (b, L) != self.key:            # a ragged last batch: same on every rank, takes the NCCL path
            return gather_latents(t_raw, i_raw)
        self.step += 1
        ops.latent_exchange(t_raw.contiguous(), i_raw.contiguous(), b=b, L=L, rank=self.rank, world=self.world,
                            peer_bufs=self.peer_bufs, peer_flags=self.peer_flags, step=self.step)
        half = self.world * b * L
        base = (self.step & 1) * 2 * half
        return self.buf[base:base + half].view(self.world * b, L), self.buf[base + half:base + 2 * half].view(self.world * b, L)
