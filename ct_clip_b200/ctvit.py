"""CTViT -- drop-in for `transformer_maskgit.CTViT` (reference: transformer_maskgit/transformer_maskgit/ctvit.py:118)
on the CT-CLIP contrastive path (encode + vector-quantise, `return_encoded_tokens` /
`return_only_codebook_ids`). Same constructor keywords, same state-dict keys and shapes; the
arithmetic runs in the hand-written sm_100a kernels of libctclip_b200.so through CTViTEngine.

The nn.Module tree below exists only to own parameters under the reference's names -- none of the
torch layers is ever called. The GenerateCT leftovers of the reference class (decode, GAN / VGG
losses, pixel heads; ctvit.py:309-351, :414-525) are dead code there and raise NotImplementedError here;
their parameters (`to_pixels*`, `to_patch_emb_first_frame`) are kept so checkpoints load strictly.
"""
from __future__ import annotations

from pathlib import Path

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .engine import CTViTEngine, ViTGeom


def _pair(v):
    return v if isinstance(v, tuple) else (v, v)


class _Holder(nn.Module):
    """Parameter container with a given set of children (never executed)."""

    def __init__(self, **children):
        super().__init__()
        for k, v in children.items():
            if isinstance(v, nn.Module):
                self.add_module(k, v)
            elif isinstance(v, nn.Parameter):
                self.register_parameter(k, v)
            else:
                self.register_buffer(k, v)


class _NormGammaBeta(nn.Module):  # attention.py:28-35: gamma parameter, beta zero buffer
    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(dim))
        self.register_buffer("beta", torch.zeros(dim))


class _AttentionParams(nn.Module):  # attention.py:88-125 (self-attention: num_null_kv = 0)
    def __init__(self, dim, dim_head, heads):
        super().__init__()
        inner = dim_head * heads
        self.norm = _NormGammaBeta(dim)
        self.context_norm = _NormGammaBeta(dim)
        self.null_kv = nn.Parameter(torch.randn(heads, 0, dim_head))
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(dim, inner * 2, bias=False)
        self.q_scale = nn.Parameter(torch.ones(dim_head))
        self.k_scale = nn.Parameter(torch.ones(dim_head))
        self.to_out = nn.Linear(inner, dim, bias=False)


class _TransformerParams(nn.Module):  # attention.py:280-309 with peg=True, no cross attention
    def __init__(self, dim, depth, dim_head, heads):
        super().__init__()
        inner_ff = int(4 * (2 / 3) * dim)
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([
                _Holder(dsconv=nn.Conv3d(dim, dim, 3, groups=dim)),
                _AttentionParams(dim, dim_head, heads),
                None,
                nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, inner_ff * 2, bias=False), nn.Identity(), nn.Identity(),
                              nn.Linear(inner_ff, dim, bias=False)),
            ]))
        self.norm_out = _NormGammaBeta(dim)


class _Codebook(nn.Module):  # vector_quantize_pytorch 1.1.2 CosineSimCodebook buffers
    def __init__(self, dim, codebook_size):
        super().__init__()
        embed = F.normalize(nn.init.kaiming_uniform_(torch.empty(1, codebook_size, dim)), dim=-1)
        self.register_buffer("initted", torch.Tensor([True]))
        self.register_buffer("cluster_size", torch.zeros(1, codebook_size))
        self.register_buffer("embed", embed)


class _CTViTTokensFn(torch.autograd.Function):
    """video -> quantised tokens (b,T,H,W,D) with the straight-through gradient of the training-mode quantiser."""

    @staticmethod
    def forward(ctx, module, need_grad, video, names, *params):
        P = dict(zip(names, params))
        ectx = module._run_forward(video, P, save=need_grad)
        g = module.engine.g
        tokens = torch.empty(ectx["M"], g.dim, device=video.device)
        ops.vq_gather(ectx["idx"], ectx["P"]["vq._codebook.embed"], tokens, ectx["M"], g.dim)
        module._finish_quantize(ectx)
        ctx.module, ctx.names, ctx.ectx = module, names, ectx
        ctx.save_for_backward(*params)
        module._last_indices = ectx["idx"].view(ectx["b"], ectx["T"], g.H, g.W)
        return tokens.view(ectx["b"], ectx["T"], g.H, g.W, g.dim)

    @staticmethod
    def backward(ctx, dtokens):
        module, names = ctx.module, ctx.names
        params = ctx.saved_tensors
        P = dict(zip(names, params))
        G = {n: torch.zeros_like(p) for n, p in P.items() if p.requires_grad and p.numel() > 0}
        Gd = _GradDict(G)
        dtok = dtokens.contiguous().view(-1, module.engine.g.dim).float()
        module.engine.backward(ctx.ectx, dtok, P, Gd)
        ctx.ectx = None
        grads = tuple(G.get(n) if (p.requires_grad and n in G and n in Gd.touched) else None for n, p in zip(names, params))
        return (None, None, None, None) + grads


class _GradDict(dict):
    """dict that records which gradients the engine actually wrote (others are returned as None)."""

    def __init__(self, d):
        super().__init__(d)
        self.touched = set()

    def __getitem__(self, k):
        self.touched.add(k)
        return super().__getitem__(k)


class CTViT(nn.Module):
    def __init__(self, *, dim, codebook_size, image_size, patch_size, temporal_patch_size, spatial_depth,
                 temporal_depth, discr_base_dim=16, dim_head=64, heads=8, channels=1, use_vgg_and_gan=True, vgg=None,
                 discr_attn_res_layers=(16,), use_hinge_loss=True, attn_dropout=0., ff_dropout=0.):
        super().__init__()
        assert attn_dropout == 0. and ff_dropout == 0., "dropout is 0 on the CT-CLIP path (run_train.py:17-27)"
        self.image_size = _pair(image_size)
        self.patch_size = _pair(patch_size)
        ph, pw = self.patch_size
        self.temporal_patch_size = temporal_patch_size
        ih, iw = self.image_size
        assert ih % ph == 0 and iw % pw == 0
        self.dim, self.heads, self.dim_head, self.channels = dim, heads, dim_head, channels
        self.spatial_depth, self.temporal_depth, self.codebook_size = spatial_depth, temporal_depth, codebook_size

        # ---- parameters under the reference's state-dict names (ctvit.py:158-198)
        self.spatial_rel_pos_bias = _Holder(net=nn.ModuleList([
            nn.Sequential(nn.Linear(2, dim), nn.Identity()),
            nn.Sequential(nn.Linear(dim, dim), nn.Identity()),
            nn.Linear(dim, heads),
        ]))
        self.to_patch_emb_first_frame = nn.Sequential(
            nn.Identity(), nn.LayerNorm(channels * pw * ph), nn.Linear(channels * pw * ph, dim), nn.LayerNorm(dim))
        pv = channels * pw * ph * temporal_patch_size
        self.to_patch_emb = nn.Sequential(nn.Identity(), nn.LayerNorm(pv), nn.Linear(pv, dim), nn.LayerNorm(dim))
        self.enc_spatial_transformer = _TransformerParams(dim, spatial_depth, dim_head, heads)
        self.enc_temporal_transformer = _TransformerParams(dim, temporal_depth, dim_head, heads)
        self.vq = _Holder(_codebook=_Codebook(dim, codebook_size))
        self.to_pixels_first_frame = nn.Sequential(nn.Linear(dim, channels * pw * ph), nn.Identity())
        self.to_pixels = nn.Sequential(nn.Linear(dim, pv), nn.Identity())

        self._engine = None
        self._weights_version = None
        self._last_indices = None
        self.vq_decay = 0.8
        self.ema_all_reduce = None       # set by the trainer under data parallelism
        self._force_indices = None       # test hook: (b,T,H,W) int tensor overriding the argmax

    # ------------------------------------------------------------------------------------------
    @property
    def geom(self) -> ViTGeom:
        return ViTGeom(dim=self.dim, codebook_size=self.codebook_size, image_hw=self.image_size, patch_hw=self.patch_size,
                       temporal_patch=self.temporal_patch_size, spatial_depth=self.spatial_depth,
                       temporal_depth=self.temporal_depth, dim_head=self.dim_head, heads=self.heads, channels=self.channels)

    @property
    def patch_height_width(self):  # ctvit.py:278-280
        return self.image_size[0] // self.patch_size[0], self.image_size[1] // self.patch_size[1]

    @property
    def image_num_tokens(self):
        h, w = self.patch_height_width
        return h * w

    @property
    def engine(self) -> CTViTEngine:
        if self._engine is None:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("ct_clip_b200.CTViT runs on a CUDA (sm_100a) device only -- there is no CPU path; "
                                   "move the module to cuda first")
            self._engine = CTViTEngine(self.geom, dev)
            self._weights_version = None
        return self._engine

    def _apply(self, fn, *a, **k):  # device moves invalidate the engine's operand copies
        self._engine = None
        return super()._apply(fn, *a, **k)

    def named_live_tensors(self):
        """(names, tensors) the engine consumes: parameters and the buffers it reads/writes."""
        names, tensors = [], []
        for n, p in self.named_parameters():
            names.append(n)
            tensors.append(p)
        for n, bfr in self.named_buffers():
            names.append(n)
            tensors.append(bfr)
        return names, tensors

    def mark_weights_dirty(self):
        self._weights_version = None

    def _param_dict(self, names, tensors):
        P = {}
        for n, t in zip(names, tensors):
            if n == "vq._codebook.embed" or n == "vq._codebook.cluster_size":
                t = t[0]
            if "dsconv.weight" in n:
                t = t.view(t.shape[0], 27)
            P[n] = t
        return P

    def _ensure_weights(self, P):
        ver = sum(t._version for t in P.values()) + sum(t.data_ptr() % 1000003 for t in P.values())
        if self._weights_version != ver:
            self.engine.prepare_weights(P)
            self._weights_version = ver

    def _run_forward(self, video, P, *, save, taps=None):
        P = self._param_dict(P.keys(), P.values())
        for t in P.values():
            assert t.is_contiguous()
        self._ensure_weights(P)
        if video.dtype not in (torch.float32, torch.int16):
            video = video.float()
        ectx = self.engine.forward(video.contiguous(), P, save=save, taps=taps)
        if self._force_indices is not None:
            ectx["idx"] = self._force_indices.to(device=video.device, dtype=torch.int32).reshape(-1).contiguous()
        ectx["P"] = P
        self._last_indices = ectx["idx"].view(ectx["b"], ectx["T"], self.engine.g.H, self.engine.g.W)   # code-book ids of this pass
        return ectx

    def _finish_quantize(self, ectx):
        """Training-mode code-book EMA. MUST run after the caller has gathered the quantised tokens: the reference
        looks codes up in the pre-update code-book (CosineSimCodebook.forward: batched_embedding before ema_inplace)."""
        if self.training:
            with torch.no_grad():
                self.engine.vq_ema(ectx, ectx["P"], self.vq_decay, self.ema_all_reduce)
                self.engine.prepare_codebook(ectx["P"])

    # ------------------------------------------------------------------------------------------
    def forward(self, video, mask=None, return_recons=False, return_recons_only=False, return_discr_loss=False,
                apply_grad_penalty=True, return_only_codebook_ids=False, return_encoded_tokens=False):
        assert video.ndim in {4, 5}
        if video.ndim == 4:  # ctvit.py:366-371
            video = video.unsqueeze(2)
            assert mask is None
        assert mask is None, "frame masks are not used on the CT-CLIP path"
        assert tuple(video.shape[-2:]) == tuple(self.image_size)  # ctvit.py:375
        if not (return_only_codebook_ids or return_encoded_tokens):
            raise NotImplementedError("reconstruction / GAN branches of CTViT (ctvit.py:414-525) are dead code in the "
                                      "reference CT-CLIP path and are not part of this build")
        names, tensors = self.named_live_tensors()
        if return_only_codebook_ids:
            with torch.no_grad():
                ectx = self._run_forward(video, dict(zip(names, tensors)), save=False)
                self._finish_quantize(ectx)
            g = self.engine.g
            return ectx["idx"].view(ectx["b"], ectx["T"], g.H, g.W).long()
        need_grad = torch.is_grad_enabled() and any(t.requires_grad for t in tensors)
        return _CTViTTokensFn.apply(self, need_grad, video, tuple(names), *tensors)

    # reference API surface (ctvit.py:259-276)
    def state_dict(self, *args, **kwargs):
        return super().state_dict(*args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.mark_weights_dirty()
        return super().load_state_dict(*args, **kwargs)

    def load(self, path):
        path = Path(path)
        assert path.exists()
        self.load_state_dict(torch.load(str(path), map_location="cpu"))

    @property
    def codebook(self):
        return self.vq._codebook.embed[0]

    def decode(self, *a, **k):
        raise NotImplementedError("CTViT.decode is dead code in the reference (ctvit.py:309-351 uses undefined dec_* modules)")
