"""TEST INFRASTRUCTURE ONLY. Restatement of the cosine-similarity vector quantiser the reference
imports from the third-party package `vector-quantize-pytorch==1.1.2`
(transformer_maskgit/setup.py:19; constructed ctvit.py:188 as
`VectorQuantize(dim, codebook_size, use_cosine_sim=True)`, called ctvit.py:403, `.codebook` read
ctvit.py:275).

PARITY UNPINNED: the package is neither vendored in /root/reference nor installed in this image,
and the reference has no test or golden vector for it. The algorithm below is the published
v1.1.2 behaviour for the defaults the reference uses (heads=1, decay=0.8, eps=1e-5,
kmeans_init=False, threshold_ema_dead_code=0, commitment_weight=1.0, sample_codebook_temp=0,
sync_codebook=False, codebook_dim == dim so no in/out projection, accept_image_fmap=False):

  flatten = l2norm(x.float()); embed_n = l2norm(embed)
  dist    = flatten @ embed_n^T ; ind = argmax(dist)              (first max wins)
  quant   = embed[ind]                                            (the stored, un-renormalised buffer)
  training only (buffer side effects; do not change this step's output):
     bins        = histogram(ind)
     cluster_size <- cluster_size*decay + bins*(1-decay)
     embed_sum   = scatter_add(flatten by ind)
     embed_new   = l2norm(embed_sum / max(bins,1)) ; rows with bins == 0 keep embed_n
     embed       <- embed*decay + embed_new*(1-decay)
  training only: quant = x + (quant - x).detach()   (straight-through)
  loss = mse(quant.detach(), x) * commitment_weight (training) else 0   -- unused on the CT-CLIP path

State-dict layout: `_codebook.initted (1,)`, `_codebook.cluster_size (1,C)`, `_codebook.embed (1,C,D)`.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn


def l2norm(t):
    return F.normalize(t, p=2, dim=-1)


def vq_cosine_lookup(x: torch.Tensor, embed: torch.Tensor):
    """x (..., D) float, embed (C, D). Returns (quantized (..., D), indices (...), flatten_normed (N, D))."""
    shape = x.shape
    flat = l2norm(x.reshape(-1, shape[-1]).float())
    dist = flat @ l2norm(embed.float()).t()
    ind = dist.argmax(dim=-1)
    quant = embed[ind].reshape(shape)
    return quant, ind.reshape(shape[:-1]), flat


def vq_ema_update(flat: torch.Tensor, ind: torch.Tensor, embed: torch.Tensor, cluster_size: torch.Tensor,
                  decay: float = 0.8):
    """Returns (new_embed (C,D), new_cluster_size (C,)). Pure function of its inputs."""
    Cn, D = embed.shape
    ind = ind.reshape(-1)
    bins = torch.bincount(ind, minlength=Cn).to(flat.dtype)
    new_cluster = cluster_size * decay + bins * (1 - decay)
    zero = bins == 0
    embed_sum = torch.zeros(Cn, D, dtype=flat.dtype).index_add_(0, ind, flat)
    embed_new = l2norm(embed_sum / bins.masked_fill(zero, 1.0)[:, None])
    embed_new = torch.where(zero[:, None], l2norm(embed), embed_new)
    new_embed = embed * decay + embed_new * (1 - decay)
    return new_embed, new_cluster


class _CosineSimCodebook(nn.Module):
    def __init__(self, dim, codebook_size, decay=0.8):
        super().__init__()
        self.decay = decay
        embed = l2norm(nn.init.kaiming_uniform_(torch.empty(1, codebook_size, dim)))
        self.register_buffer("initted", torch.Tensor([True]))
        self.register_buffer("cluster_size", torch.zeros(1, codebook_size))
        self.register_buffer("embed", embed)

    @torch.no_grad()
    def forward(self, x):
        quant, ind, flat = vq_cosine_lookup(x, self.embed[0])
        if self.training:
            ne, nc = vq_ema_update(flat, ind, self.embed[0], self.cluster_size[0], self.decay)
            self.embed.data[0].copy_(ne)
            self.cluster_size.data[0].copy_(nc)
        return quant, ind


class VectorQuantize(nn.Module):
    def __init__(self, dim, codebook_size, use_cosine_sim=True, decay=0.8, commitment_weight=1.0, **_):
        super().__init__()
        assert use_cosine_sim, "the CT-CLIP path only uses the cosine-sim codebook"
        self.commitment_weight = commitment_weight
        self._codebook = _CosineSimCodebook(dim, codebook_size, decay)

    @property
    def codebook(self):
        return self._codebook.embed[0]

    def forward(self, x, mask=None):
        assert mask is None
        xf = x.float()
        quant, ind = self._codebook(xf.detach())
        if self.training:
            quant = xf + (quant - xf).detach()
        loss = torch.zeros(1, dtype=xf.dtype, requires_grad=self.training)
        if self.training and self.commitment_weight > 0:
            loss = loss + F.mse_loss(quant.detach(), xf) * self.commitment_weight
        return quant, ind, loss
