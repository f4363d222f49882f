"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the CT-CLIP contrastive hot path.

A functional, dependency-free (torch CPU, fp32 or fp64) restatement of the reference algorithm.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it, and only as the checker / reported baseline. The product path (ct_clip_b200/*) never
imports anything under oracle/.

Every function cites the reference lines it restates (paths relative to /root/reference).
Parameters are addressed by the reference's own state-dict keys, so one flat dict drives the
reference (load_state_dict), this oracle and the CUDA implementation alike.

Pinning status:
  * image tower, heads, loss: pinned against the unmodified reference run through
    oracle/ref_shims.py (tests/test_oracle_cpu.py::test_oracle_matches_reference, build container
    only) and against the committed golden vectors tests/golden/*.pt (everywhere);
  * BERT text tower: pinned against transformers.BertModel (eager attention);
  * vector quantiser: restated from the published vector-quantize-pytorch 1.1.2 algorithm,
    PARITY UNPINNED (see oracle/vq_restated.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F

from .vq_restated import vq_cosine_lookup, vq_ema_update


@dataclass
class CTViTConfig:
    dim: int = 512
    codebook_size: int = 8192
    image_size: int = 480
    patch_size: int = 20
    temporal_patch_size: int = 10
    spatial_depth: int = 4
    temporal_depth: int = 4
    dim_head: int = 32
    heads: int = 8
    channels: int = 1

    @property
    def ff_inner(self) -> int:  # attention.py:45
        return int(4 * (2 / 3) * self.dim)

    @property
    def grid_hw(self) -> int:
        return self.image_size // self.patch_size

    @property
    def patch_voxels(self) -> int:
        return self.channels * self.temporal_patch_size * self.patch_size * self.patch_size


@dataclass
class BertConfigLite:
    vocab_size: int = 30522
    hidden: int = 768
    layers: int = 12
    heads: int = 12
    intermediate: int = 3072
    max_pos: int = 512
    type_vocab: int = 2
    eps: float = 1e-12


@dataclass
class CTCLIPConfig:
    vit: CTViTConfig = field(default_factory=CTViTConfig)
    bert: BertConfigLite = field(default_factory=BertConfigLite)
    dim_text: int = 768
    dim_latent: int = 512

    def dim_image(self) -> int:  # run_train.py:34 (294912 = 24*24*512)
        return self.vit.grid_hw * self.vit.grid_hw * self.vit.dim


# ------------------------------------------------------------------------------------------------
# image tower
# ------------------------------------------------------------------------------------------------
def _ln(x, gamma, beta, eps=1e-5):
    return F.layer_norm(x, x.shape[-1:], gamma, beta, eps)


def patch_embed(video, sd, pre, cfg: CTViTConfig):
    """ctvit.py:170-175: Rearrange 'b c (t pt) (h p1) (w p2) -> b t h w (c pt p1 p2)' -> LN(P) -> Linear -> LN(D)."""
    b, c, f, H, W = video.shape
    pt, p = cfg.temporal_patch_size, cfg.patch_size
    t, h, w = f // pt, H // p, W // p
    x = video.reshape(b, c, t, pt, h, p, w, p).permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(b, t, h, w, c * pt * p * p)
    x = _ln(x, sd[pre + "to_patch_emb.1.weight"], sd[pre + "to_patch_emb.1.bias"])
    x = F.linear(x, sd[pre + "to_patch_emb.2.weight"], sd[pre + "to_patch_emb.2.bias"])
    return _ln(x, sd[pre + "to_patch_emb.3.weight"], sd[pre + "to_patch_emb.3.bias"])


def cpb_bias(sd, pre, h, w):
    """attention.py:257-276 ContinuousPositionBias (num_dims=2, layers=2, log_dist): (heads, h*w, h*w) fp32."""
    dt = sd[pre + "spatial_rel_pos_bias.net.0.0.weight"].dtype
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    grid = torch.stack([ys, xs], dim=-1).reshape(-1, 2)
    rel = (grid[:, None, :] - grid[None, :, :]).to(dt)
    rel = torch.sign(rel) * torch.log(rel.abs() + 1)
    x = rel
    x = F.leaky_relu(F.linear(x, sd[pre + "spatial_rel_pos_bias.net.0.0.weight"], sd[pre + "spatial_rel_pos_bias.net.0.0.bias"]), 0.1)
    x = F.leaky_relu(F.linear(x, sd[pre + "spatial_rel_pos_bias.net.1.0.weight"], sd[pre + "spatial_rel_pos_bias.net.1.0.bias"]), 0.1)
    x = F.linear(x, sd[pre + "spatial_rel_pos_bias.net.2.weight"], sd[pre + "spatial_rel_pos_bias.net.2.bias"])
    return x.permute(2, 0, 1)


def peg(x, shape, weight, bias):
    """attention.py:63-84 PEG with causal=True: reshape (NOT rearrange) to `shape`, pad (1,1),(1,1),(2,0), depthwise conv3d."""
    orig = x.shape
    x = x.reshape(*shape, -1).permute(0, 4, 1, 2, 3)
    x = F.pad(x, (1, 1, 1, 1, 2, 0), value=0.0)
    x = F.conv3d(x, weight, bias, groups=weight.shape[0])
    return x.permute(0, 2, 3, 4, 1).reshape(orig)


def attention(x, sd, pre, heads, attn_bias=None, scale=8.0):
    """attention.py:127-181 (self-attention; num_null_kv=0, no mask, not causal).
    QUIRK (attention.py:139-145): `kv_input = default(context, x)` is bound BEFORE `x = self.norm(x)`,
    so K and V are projected from the RAW residual stream and only Q sees the LayerNorm."""
    xn = _ln(x, sd[pre + "norm.gamma"], sd[pre + "norm.beta"])
    q = F.linear(xn, sd[pre + "to_q.weight"])
    k, v = F.linear(x, sd[pre + "to_kv.weight"]).chunk(2, dim=-1)
    B, n, inner = q.shape
    dh = inner // heads
    q, k, v = (t.reshape(B, n, heads, dh).transpose(1, 2) for t in (q, k, v))
    q = F.normalize(q, dim=-1) * sd[pre + "q_scale"]
    k = F.normalize(k, dim=-1) * sd[pre + "k_scale"]
    sim = q @ k.transpose(-1, -2) * scale
    if attn_bias is not None:
        sim = sim + attn_bias
    attn = sim.softmax(dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(B, n, inner)
    return F.linear(out, sd[pre + "to_out.weight"])


def feedforward(x, sd, pre):
    """attention.py:39-52: LayerNorm -> Linear(D, 2F) -> gelu(gate)*x (x = first half) -> Linear(F, D)."""
    h = _ln(x, sd[pre + "0.weight"], sd[pre + "0.bias"])
    h = F.linear(h, sd[pre + "1.weight"])
    a, gate = h.chunk(2, dim=-1)
    return F.linear(F.gelu(gate) * a, sd[pre + "4.weight"])


def transformer(x, sd, pre, depth, heads, video_shape, attn_bias=None, taps=None):
    """attention.py:312-333."""
    for i in range(depth):
        lp = f"{pre}layers.{i}."
        x = peg(x, video_shape, sd[lp + "0.dsconv.weight"], sd[lp + "0.dsconv.bias"]) + x
        if taps is not None:
            taps[lp + "peg"] = x
        x = attention(x, sd, lp + "1.", heads, attn_bias) + x
        if taps is not None:
            taps[lp + "attn"] = x
        x = feedforward(x, sd, lp + "3.") + x
        if taps is not None:
            taps[lp + "ff"] = x
    return _ln(x, sd[pre + "norm_out.gamma"], sd[pre + "norm_out.beta"])


def ctvit_encode(tokens, sd, pre, cfg: CTViTConfig, taps=None):
    """ctvit.py:282-307: spatial stack over (b t) x (h w), temporal stack over (b h w) x t.
    NOTE (SURVEY trap T1): both stacks receive the SAME video_shape (b,t,h,w)."""
    b, t, h, w, d = tokens.shape
    video_shape = (b, t, h, w)
    x = tokens.reshape(b * t, h * w, d)
    bias = cpb_bias(sd, pre, h, w)
    if taps is not None:
        taps["cpb_bias"] = bias
    x = transformer(x, sd, pre + "enc_spatial_transformer.", cfg.spatial_depth, cfg.heads, video_shape, bias, taps)
    x = x.reshape(b, t, h, w, d)
    if taps is not None:
        taps["spatial_out"] = x
    x = x.permute(0, 2, 3, 1, 4).reshape(b * h * w, t, d)
    x = transformer(x, sd, pre + "enc_temporal_transformer.", cfg.temporal_depth, cfg.heads, video_shape, None, taps)
    return x.reshape(b, h, w, t, d).permute(0, 3, 1, 2, 4)


def ctvit_forward(video, sd, pre, cfg: CTViTConfig, training: bool, taps=None, force_indices=None):
    """ctvit.py:353-412 with return_encoded_tokens=True. Returns (tokens (b,t,h,w,d), indices (b,t,h,w), ema)
    where ema = (new_embed, new_cluster_size) in training mode (the buffer side effect) else None."""
    tokens = patch_embed(video, sd, pre, cfg)
    if taps is not None:
        taps["patch_tokens"] = tokens
    tokens = ctvit_encode(tokens, sd, pre, cfg, taps)
    b, t, h, w, d = tokens.shape
    if taps is not None:
        taps["pre_vq"] = tokens
    x = tokens.reshape(b, t * h * w, d)
    embed = sd[pre + "vq._codebook.embed"][0]
    with torch.no_grad():
        quant, ind, flat = vq_cosine_lookup(x.detach(), embed)
        if force_indices is not None:
            ind = force_indices.reshape(ind.shape).long()
            quant = embed[ind]
        ema = None
        if training:
            ema = vq_ema_update(flat, ind, embed.float(), sd[pre + "vq._codebook.cluster_size"][0].float())
    if training:
        quant = x + (quant - x).detach()
    return quant.reshape(b, t, h, w, d), ind.reshape(b, t, h, w), ema


# ------------------------------------------------------------------------------------------------
# text tower (HF BertModel restated: transformers/models/bert/modeling_bert.py, eager attention)
# ------------------------------------------------------------------------------------------------
def bert_forward(input_ids, attention_mask, sd, pre, cfg: BertConfigLite, dropout=None):
    """Returns last_hidden_state (b, n, hidden). Called at ct_clip.py:685-686.
    dropout (training mode, CTCLIPTrainer.py:254 with CXR-BERT's p = 0.1): None, or dict(p_hidden, p_attn, masks) where
    masks[(layer, kind)] is the boolean KEEP mask of a site -- HF applies `nn.Dropout` at: the embeddings after their LayerNorm
    ((-1, 3), BertEmbeddings.forward), the attention probabilities ((i, 0), BertSelfAttention: softmax -> dropout -> @ V),
    BertSelfOutput ((i, 1): dense -> dropout -> + residual -> LayerNorm) and BertOutput ((i, 2), same shape). A dropped
    element is zero, a kept one is scaled by 1/(1-p) (torch.nn.functional.dropout)."""
    def drop(t, site, p):
        if dropout is None or p <= 0:
            return t
        return t * dropout["masks"][site].to(t.dtype).reshape(t.shape) / (1.0 - p)
    ph = dropout["p_hidden"] if dropout else 0.0
    pa = dropout["p_attn"] if dropout else 0.0
    b, n = input_ids.shape
    x = (sd[pre + "embeddings.word_embeddings.weight"][input_ids]
         + sd[pre + "embeddings.position_embeddings.weight"][:n][None]
         + sd[pre + "embeddings.token_type_embeddings.weight"][0][None, None])
    x = _ln(x, sd[pre + "embeddings.LayerNorm.weight"], sd[pre + "embeddings.LayerNorm.bias"], cfg.eps)
    x = drop(x, (-1, 3), ph)
    dh = cfg.hidden // cfg.heads
    neg = torch.finfo(x.dtype).min
    add_mask = (1.0 - attention_mask[:, None, None, :].to(x.dtype)) * neg
    for i in range(cfg.layers):
        lp = f"{pre}encoder.layer.{i}."
        q = F.linear(x, sd[lp + "attention.self.query.weight"], sd[lp + "attention.self.query.bias"])
        k = F.linear(x, sd[lp + "attention.self.key.weight"], sd[lp + "attention.self.key.bias"])
        v = F.linear(x, sd[lp + "attention.self.value.weight"], sd[lp + "attention.self.value.bias"])
        q, k, v = (t.reshape(b, n, cfg.heads, dh).transpose(1, 2) for t in (q, k, v))
        s = q @ k.transpose(-1, -2) / math.sqrt(dh) + add_mask
        a = (drop(s.softmax(dim=-1), (i, 0), pa) @ v).transpose(1, 2).reshape(b, n, cfg.hidden)
        a = F.linear(a, sd[lp + "attention.output.dense.weight"], sd[lp + "attention.output.dense.bias"])
        x = _ln(drop(a, (i, 1), ph) + x, sd[lp + "attention.output.LayerNorm.weight"], sd[lp + "attention.output.LayerNorm.bias"], cfg.eps)
        m = F.gelu(F.linear(x, sd[lp + "intermediate.dense.weight"], sd[lp + "intermediate.dense.bias"]))
        m = F.linear(m, sd[lp + "output.dense.weight"], sd[lp + "output.dense.bias"])
        x = _ln(drop(m, (i, 2), ph) + x, sd[lp + "output.LayerNorm.weight"], sd[lp + "output.LayerNorm.bias"], cfg.eps)
    return x


# ------------------------------------------------------------------------------------------------
# CLIP heads + loss
# ------------------------------------------------------------------------------------------------
def clip_latents(enc_text, enc_image_tokens, sd):
    """ct_clip.py:724 (mean over t), :740 (flatten), :762 (CLS row), :765-771 (projections + l2norm)."""
    b = enc_image_tokens.shape[0]
    enc_image = enc_image_tokens.mean(dim=1).reshape(b, -1)
    text_lat = F.linear(enc_text[:, 0, :], sd["to_text_latent.weight"])
    img_lat = F.linear(enc_image, sd["to_visual_latent.weight"])
    return F.normalize(text_lat, dim=-1), F.normalize(img_lat, dim=-1)


def clip_loss(text_lat, img_lat, temperature):
    """ct_clip.py:796, :845-878 (single view, no DCL, no extra projection): symmetric InfoNCE in the
    reference's exp/log form (log(x) := log(x + 1e-20), no max-subtraction)."""
    temp = temperature.exp()
    t2i = text_lat @ img_lat.t() * temp
    i2t = t2i.t()
    e1, e2 = t2i.exp(), i2t.exp()
    pos1, pos2 = torch.diagonal(e1), torch.diagonal(e2)
    den1, den2 = e1.sum(-1), e2.sum(-1)
    l1 = (-torch.log(pos1 + 1e-20) + torch.log(den1 + 1e-20)).mean()
    l2 = (-torch.log(pos2 + 1e-20) + torch.log(den2 + 1e-20)).mean()
    return (l1 + l2) / 2


def clip_similarity(text_lat, img_lat, temperature):
    """ct_clip.py:805-807 inference path: einsum('b d, b d -> b') * exp(T) with broadcasting."""
    return (text_lat * img_lat).sum(-1) * temperature.exp()


def ctclip_forward(sd, cfg: CTCLIPConfig, input_ids, attention_mask, video, *, training=True, return_loss=True,
                   taps=None, force_indices=None):
    """CTCLIP.forward (ct_clip.py:614-901) on the script path (use_mlm/visual_ssl/extra projection/multiview off).
    Returns dict(loss|sims, text_latents, image_latents, tokens, indices, ema)."""
    enc_text = bert_forward(input_ids, attention_mask, sd, "text_transformer.", cfg.bert)
    tokens, ind, ema = ctvit_forward(video, sd, "visual_transformer.", cfg.vit, training, taps, force_indices)
    tl, il = clip_latents(enc_text, tokens, sd)
    out = dict(text_latents=tl, image_latents=il, tokens=tokens, indices=ind, ema=ema, enc_text=enc_text)
    if return_loss:
        out["loss"] = clip_loss(tl, il, sd["temperature"])
    else:
        out["sims"] = clip_similarity(tl, il, sd["temperature"])
    return out


# ------------------------------------------------------------------------------------------------
# deterministic synthetic parameters / inputs (shared by golden generation, tests and bench)
# ------------------------------------------------------------------------------------------------
def _key_seed(key: str, seed: int) -> int:
    h = 1469598103934665603
    for ch in (key + f"#{seed}").encode():
        h = ((h ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h & 0x7FFFFFFF


def synth_tensor(key: str, shape, seed: int = 0) -> torch.Tensor:
    """Deterministic value for one state-dict entry, independent of construction order."""
    g = torch.Generator().manual_seed(_key_seed(key, seed))
    shape = tuple(shape)
    leaf = key.split(".")[-1]
    if key.endswith("vq._codebook.embed"):
        return F.normalize(torch.randn(shape, generator=g), dim=-1)
    if key.endswith("vq._codebook.cluster_size"):
        return torch.zeros(shape)
    if key.endswith("initted"):
        return torch.ones(shape)
    if key == "temperature":
        return torch.tensor(1.0)
    if key.endswith("position_ids"):
        return torch.arange(shape[-1]).reshape(shape)
    if key.endswith("token_type_ids"):
        return torch.zeros(shape, dtype=torch.long)
    if leaf in ("q_scale", "k_scale", "gamma") or (leaf == "weight" and len(shape) == 1):
        return 1.0 + 0.1 * torch.randn(shape, generator=g)            # norm scales
    if leaf == "beta" and "norm" in key and "visual_transformer" in key and ".3.0." not in key:
        return torch.zeros(shape)                                        # attention.py:32 zero buffer
    if leaf in ("bias", "beta"):
        return 0.02 * torch.randn(shape, generator=g)
    if leaf == "null_kv":
        return torch.randn(shape, generator=g)
    if "dsconv.weight" in key:
        return 0.15 * torch.randn(shape, generator=g)
    if "spatial_rel_pos_bias" in key:
        return torch.randn(shape, generator=g) / math.sqrt(max(shape[-1], 1))
    if "embeddings" in key and len(shape) == 2:
        return 0.05 * torch.randn(shape, generator=g)
    fan_in = shape[-1] if len(shape) >= 2 else 1
    return torch.randn(shape, generator=g) / math.sqrt(fan_in)


def synth_state_dict(shapes: dict, seed: int = 0) -> dict:
    return {k: synth_tensor(k, s, seed) for k, s in shapes.items()}


def synth_inputs(b, frames, image, n_text, seed=1234, vocab=30522):
    """int16 HU volumes clip(round(N(-300, 450^2)), -1000, 1000) consumed as x/1000;
    token ids uniform in [5, vocab), [CLS]=2 first, [SEP]=3 last valid, lengths in [n/4, n], pad 0."""
    g = torch.Generator().manual_seed(seed)
    hu = (torch.randn(b, 1, frames, image, image, generator=g) * 450.0 - 300.0).round().clamp(-1000, 1000).to(torch.int16)
    g2 = torch.Generator().manual_seed(seed + 3087)
    ids = torch.randint(5, vocab, (b, n_text), generator=g2)
    lens = torch.randint(max(2, n_text // 4), n_text + 1, (b,), generator=g2)
    mask = (torch.arange(n_text)[None, :] < lens[:, None]).long()
    ids[:, 0] = 2
    ids[torch.arange(b), lens - 1] = 3
    ids = ids * mask
    return hu, ids, mask


# ------------------------------------------------------------------------------------------------
# dataset pre-processing: restatement of scripts/data.py:12-34 (resize_array) + :92-162 (nii_img_to_tensor)
# from the point where nibabel has produced the voxel array (`img_data = nii_img.get_fdata()`, float64, (x, y, z))
# ------------------------------------------------------------------------------------------------
def ct_preprocess(img_data, slope, intercept, xy_spacing, z_spacing, target_shape=(480, 480, 240)):
    """Returns the (1, D, H, W) float32 tensor in [-1, 1] the reference dataset yields."""
    import numpy as np
    current = (z_spacing, xy_spacing, xy_spacing)
    target = (1.5, 0.75, 0.75)
    img_data = slope * np.asarray(img_data, dtype=np.float64) + intercept            # data.py:109
    img_data = img_data.transpose(2, 0, 1)                                           # data.py:111
    tensor = torch.tensor(img_data).unsqueeze(0).unsqueeze(0)
    original_shape = tensor.shape[2:]
    new_shape = [int(original_shape[i] * (current[i] / target[i])) for i in range(3)]        # data.py:26-31
    resized = F.interpolate(tensor, size=new_shape, mode="trilinear", align_corners=False).cpu().numpy()
    img_data = np.transpose(resized[0][0], (1, 2, 0))                                # data.py:117
    img_data = (np.clip(img_data, -1000, 1000) / 1000).astype(np.float32)            # data.py:119-123
    tensor = torch.tensor(img_data)
    h, w, d = tensor.shape
    dh, dw, dd = target_shape
    h_start = max((h - dh) // 2, 0)
    h_end = min(h_start + dh, h)
    w_start = max((w - dw) // 2, 0)
    w_end = min(w_start + dw, w)
    d_start = max((d - dd) // 2, 0)
    d_end = min(d_start + dd, d)
    tensor = tensor[h_start:h_end, w_start:w_end, d_start:d_end]
    pad_h_before = (dh - tensor.size(0)) // 2
    pad_h_after = dh - tensor.size(0) - pad_h_before
    pad_w_before = (dw - tensor.size(1)) // 2
    pad_w_after = dw - tensor.size(1) - pad_w_before
    pad_d_before = (dd - tensor.size(2)) // 2
    pad_d_after = dd - tensor.size(2) - pad_d_before
    tensor = F.pad(tensor, (pad_d_before, pad_d_after, pad_w_before, pad_w_after, pad_h_before, pad_h_after), value=-1)
    return tensor.permute(2, 0, 1).unsqueeze(0)                                      # data.py:160-162
