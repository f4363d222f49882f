"""Native text tower: runs a HuggingFace `BertModel`'s parameters (the text encoder every reference script
injects, run_train.py:7-9 / ct_clip.py:685-686) on the sm_100a kernels, forward and backward.

Architecture restated from transformers' BertModel (post-LN encoder, learned absolute positions, erf-GELU,
eps from config): embeddings = word + position + token_type(0) -> LayerNorm; per layer
  a = SelfAttention(x) (heads x 64, 1/sqrt(64) scaling, additive padding mask)  ;  x = LN(x + a Wo^T + bo)
  m = gelu(x Wi^T + bi) Wo2^T + bo2                                              ;  x = LN(x + m)
Only the CLS row of the last layer feeds CT-CLIP (ct_clip.py:762); the pooler is never evaluated.

Dropout (CXR-BERT trains with hidden_dropout_prob = attention_probs_dropout_prob = 0.1, run_train.py:7-9 +
CTCLIPTrainer.py:254): the four HF sites -- embeddings, attention probabilities, BertSelfOutput, BertOutput -- run on the
native kernels with counter-based Philox masks (csrc/rng.cuh) that the backward pass regenerates from (seed, site offset).
`forward(..., dropout=dict(p_hidden=, p_attn=, seed=))`; None = eval mode / p = 0.
"""
from __future__ import annotations

import math

import torch

from . import ops


def supports(bert) -> bool:
    """True if `bert` is a HF BertModel-like module this engine can run."""
    try:
        cfg = bert.config
        emb, enc = bert.embeddings, bert.encoder.layer
        ok = (cfg.hidden_size % 128 == 0 and cfg.hidden_size // cfg.num_attention_heads == 64
              and cfg.hidden_act == "gelu" and getattr(cfg, "position_embedding_type", "absolute") == "absolute"
              and not getattr(cfg, "is_decoder", False) and len(enc) == cfg.num_hidden_layers
              and hasattr(emb, "word_embeddings") and cfg.intermediate_size % 8 == 0)
        return bool(ok)
    except Exception:
        return False


def dropout_active(bert) -> bool:
    cfg = bert.config
    return bert.training and (cfg.hidden_dropout_prob > 0 or cfg.attention_probs_dropout_prob > 0)


def dropout_config(bert, seed: int):
    """The dropout argument of BertEngine.forward for this module's current mode (None when nothing is dropped)."""
    if not dropout_active(bert):
        return None
    return dict(p_hidden=float(bert.config.hidden_dropout_prob), p_attn=float(bert.config.attention_probs_dropout_prob), seed=int(seed))


def site_offset(layer: int, kind: int) -> int:
    """Philox counter base of a dropout site: embeddings = (layer -1, kind 3) -> 0; layer i: kind 0 attention probabilities,
    1 BertSelfOutput, 2 BertOutput. Sites are 2^40 counters apart (a site has < 2^38 groups of 4 elements)."""
    return (4 * (layer + 1) + kind - 3) << 40


class BertEngine:
    def __init__(self, bert, device):
        cfg = bert.config
        self.H, self.heads, self.layers, self.inter = cfg.hidden_size, cfg.num_attention_heads, cfg.num_hidden_layers, cfg.intermediate_size
        self.eps = cfg.layer_norm_eps
        self.device = device
        bf = dict(dtype=torch.bfloat16, device=device)
        H, I = self.H, self.inter
        self.w = [dict(qkv=torch.empty(3 * H, H, **bf), bqkv=torch.empty(3 * H, device=device), wo=torch.empty(H, H, **bf),
                       wi=torch.empty(I, H, **bf), wo2=torch.empty(H, I, **bf)) for _ in range(self.layers)]
        self._version = None

    @staticmethod
    def param_names(bert):
        return [n for n, _ in bert.named_parameters()]

    def prepare(self, P):
        ver = sum(t._version for t in P.values()) + sum(t.data_ptr() % 1000003 for t in P.values())
        if ver == self._version:
            return
        H, I = self.H, self.inter
        if getattr(self, "_prep", None) is None:
            self._prep = ops.PrepBatch()
        pb = self._prep
        for i, w in enumerate(self.w):
            lp = f"encoder.layer.{i}."
            for j, nm in enumerate(("query", "key", "value")):
                pb.weight(P[lp + f"attention.self.{nm}.weight"], w["qkv"][j * H:(j + 1) * H], K=H, Np=H, Kp=H)
                w["bqkv"][j * H:(j + 1) * H].copy_(P[lp + f"attention.self.{nm}.bias"])      # 768-float memcpy
            pb.weight(P[lp + "attention.output.dense.weight"], w["wo"], K=H, Np=H, Kp=H)
            pb.weight(P[lp + "intermediate.dense.weight"], w["wi"], K=H, Np=I, Kp=H)
            pb.weight(P[lp + "output.dense.weight"], w["wo2"], K=I, Np=H, Kp=I)
        pb.run()      # all operands of the tower in one launch
        self._version = ver

    def mark_dirty(self):
        self._version = None

    # ------------------------------------------------------------------------------------------
    def forward(self, input_ids, attention_mask, P, *, save, dropout=None):
        """-> (last_hidden_state fp32 [b, n, H], ctx). dropout: None or dict(p_hidden, p_attn, seed) (training mode)."""
        ph = float(dropout["p_hidden"]) if dropout else 0.0
        pa = float(dropout["p_attn"]) if dropout else 0.0
        seed = int(dropout["seed"]) if dropout else 0
        dev, H, I, heads = self.device, self.H, self.inter, self.heads
        b, n = input_ids.shape
        M = b * n
        bf = dict(dtype=torch.bfloat16, device=dev)
        self.prepare(P)
        ids = input_ids.contiguous().view(-1).long()
        mask = attention_mask.to(torch.int32).contiguous()
        geom = dict(n=n, heads=heads, num_seqs=b, seq_inner=1, seq_outer_stride=n, tok_stride=1, dim_head=64,
                    scale=1.0 / math.sqrt(64.0), key_mask=mask)
        e = torch.empty(M, H, device=dev)
        ops.bert_embed(ids, P["embeddings.word_embeddings.weight"], P["embeddings.position_embeddings.weight"],
                       P["embeddings.token_type_embeddings.weight"], e, M, n, H)
        x = torch.empty(M, H, device=dev)
        xb = torch.empty(M, H, **bf)
        sv0 = dict(xhat=torch.empty(M, H, **bf), rstd=torch.empty(M, device=dev)) if save else dict(xhat=None, rstd=None)
        ops.ln_fwd(e, M, H, eps=self.eps, gamma=P["embeddings.LayerNorm.weight"], beta=P["embeddings.LayerNorm.bias"],
                   y_f32=x, y_bf16=xb, xhat=sv0["xhat"], rstd=sv0["rstd"])
        if ph > 0:    # BertEmbeddings: dropout after the LayerNorm
            ops.dropout(x, n=M * H, p=ph, seed=seed, offset=site_offset(-1, 3), y_f32=x, y_bf16=xb)
        saved = []
        for i, w in enumerate(self.w):
            lp = f"encoder.layer.{i}."
            qkv = torch.empty(M, 3 * H, **bf)
            ops.gemm(xb, w["qkv"], M=M, N=3 * H, K=H, epilogue=ops.EPI_BF16, C_out=qkv, bias=w["bqkv"])
            ao = torch.empty(M, H, **bf)
            lse = torch.empty(M, heads, device=dev) if save else None
            ops.attn_fwd(qkv, qkv[:, H:], qkv[:, 2 * H:], ao, lse, ldq=3 * H, ldk=3 * H, ldv=3 * H, ldo=H, dropout_p=pa,
                         dropout_seed=seed, dropout_offset=site_offset(i, 0), **geom)
            y = torch.empty(M, H, device=dev)
            if ph > 0:   # BertSelfOutput: LayerNorm(x + dropout(dense(ao)))
                ops.gemm(ao, w["wo"], M=M, N=H, K=H, epilogue=ops.EPI_F32, C_out=y, bias=P[lp + "attention.output.dense.bias"])
                ops.dropout(y, n=M * H, p=ph, seed=seed, offset=site_offset(i, 1), resid=x, y_f32=y)
            else:
                ops.gemm(ao, w["wo"], M=M, N=H, K=H, epilogue=ops.EPI_RESID_F32, C_out=y, resid=x,
                         bias=P[lp + "attention.output.dense.bias"])
            x1 = torch.empty(M, H, device=dev)
            x1b = torch.empty(M, H, **bf)
            s1 = dict(xhat=torch.empty(M, H, **bf), rstd=torch.empty(M, device=dev)) if save else dict(xhat=None, rstd=None)
            ops.ln_fwd(y, M, H, eps=self.eps, gamma=P[lp + "attention.output.LayerNorm.weight"],
                       beta=P[lp + "attention.output.LayerNorm.bias"], y_f32=x1, y_bf16=x1b, xhat=s1["xhat"], rstd=s1["rstd"])
            act = torch.empty(M, I, **bf)
            pre = torch.empty(M, I, **bf) if save else None
            ops.gemm(x1b, w["wi"], M=M, N=I, K=H, epilogue=ops.EPI_BIAS_GELU, C_out=act, C2=pre,
                     bias=P[lp + "intermediate.dense.bias"])
            if ph > 0:   # BertOutput: LayerNorm(x1 + dropout(dense(act)))
                ops.gemm(act, w["wo2"], M=M, N=H, K=I, epilogue=ops.EPI_F32, C_out=y, bias=P[lp + "output.dense.bias"])
                ops.dropout(y, n=M * H, p=ph, seed=seed, offset=site_offset(i, 2), resid=x1, y_f32=y)
            else:
                ops.gemm(act, w["wo2"], M=M, N=H, K=I, epilogue=ops.EPI_RESID_F32, C_out=y, resid=x1,
                         bias=P[lp + "output.dense.bias"])
            x2 = torch.empty(M, H, device=dev)
            x2b = torch.empty(M, H, **bf)
            s2 = dict(xhat=torch.empty(M, H, **bf), rstd=torch.empty(M, device=dev)) if save else dict(xhat=None, rstd=None)
            ops.ln_fwd(y, M, H, eps=self.eps, gamma=P[lp + "output.LayerNorm.weight"], beta=P[lp + "output.LayerNorm.bias"],
                       y_f32=x2, y_bf16=x2b, xhat=s2["xhat"], rstd=s2["rstd"])
            if save:
                saved.append(dict(xb=xb, qkv=qkv, ao=ao, lse=lse, ln1=s1, x1b=x1b, act=act, pre=pre, ln2=s2))
            x, xb = x2, x2b
        ctx = dict(b=b, n=n, M=M, ids=ids, geom=geom, emb=sv0, saved=saved, drop=(ph, pa, seed)) if save else None
        return x.view(b, n, H), ctx

    def backward(self, ctx, d_last, P, G):
        """d_last: fp32 [b*n, H] gradient w.r.t. last_hidden_state. Accumulates into G (keys like P)."""
        dev, H, I, heads = self.device, self.H, self.inter, self.heads
        M, n = ctx["M"], ctx["n"]
        ph, pa, seed = ctx.get("drop", (0.0, 0.0, 0))
        bf = dict(dtype=torch.bfloat16, device=dev)
        dx = d_last
        for i in reversed(range(self.layers)):
            w, sv = self.w[i], ctx["saved"][i]
            lp = f"encoder.layer.{i}."
            # x2 = LN2(y2), y2 = x1 + act Wo2^T + bo2
            dy = torch.empty(M, H, device=dev)
            dyb = torch.empty(M, H, **bf)
            ops.ln_bwd(M, H, g_f32=dx, gamma=P[lp + "output.LayerNorm.weight"], xhat=sv["ln2"]["xhat"], rstd=sv["ln2"]["rstd"],
                       dx_f32=dy, dx_bf16=dyb, dgamma=G[lp + "output.LayerNorm.weight"], dbeta=G[lp + "output.LayerNorm.bias"])
            dsrc = dy
            if ph > 0:   # gradient w.r.t. the dense output = keep/(1-p) * dy (the residual branch keeps the full dy)
                dsrc = torch.empty(M, H, device=dev)
                ops.dropout(dy, n=M * H, p=ph, seed=seed, offset=site_offset(i, 2), y_f32=dsrc, y_bf16=dyb)
            ops.colsum(dsrc, G[lp + "output.dense.bias"], M=M, N=H)       # bias gradients from the fp32 gradient, not its bf16 copy
            self._wgrad(dyb, sv["act"], G[lp + "output.dense.weight"], n_out=H, k_out=I, rows=M)
            dact = torch.empty(M, I, **bf)
            ops.gemm(dyb, w["wo2"], M=M, N=I, K=H, b_major=1, epilogue=ops.EPI_BF16, C_out=dact)
            ops.gelu_bwd(dact, sv["pre"], M=M, N=I, colsum_out=G[lp + "intermediate.dense.bias"])      # dact <- d(pre)
            self._wgrad(dact, sv["x1b"], G[lp + "intermediate.dense.weight"], n_out=I, k_out=H, rows=M)
            ops.gemm(dact, w["wi"], M=M, N=H, K=I, b_major=1, epilogue=ops.EPI_RESID_F32, C_out=dy, resid=dy)   # dx1
            # x1 = LN1(y1), y1 = x + ao Wo^T + bo
            d1 = torch.empty(M, H, device=dev)
            d1b = torch.empty(M, H, **bf)
            ops.ln_bwd(M, H, g_f32=dy, gamma=P[lp + "attention.output.LayerNorm.weight"], xhat=sv["ln1"]["xhat"],
                       rstd=sv["ln1"]["rstd"], dx_f32=d1, dx_bf16=d1b, dgamma=G[lp + "attention.output.LayerNorm.weight"],
                       dbeta=G[lp + "attention.output.LayerNorm.bias"])
            dsrc = d1
            if ph > 0:
                dsrc = torch.empty(M, H, device=dev)
                ops.dropout(d1, n=M * H, p=ph, seed=seed, offset=site_offset(i, 1), y_f32=dsrc, y_bf16=d1b)
            ops.colsum(dsrc, G[lp + "attention.output.dense.bias"], M=M, N=H)
            self._wgrad(d1b, sv["ao"], G[lp + "attention.output.dense.weight"], n_out=H, k_out=H, rows=M)
            dao = torch.empty(M, H, **bf)
            ops.gemm(d1b, w["wo"], M=M, N=H, K=H, b_major=1, epilogue=ops.EPI_BF16, C_out=dao)
            dqkv = torch.empty(M, 3 * H, **bf)
            delta = torch.empty(M, heads, device=dev)
            qkv = sv["qkv"]
            ops.attn_bwd(qkv, qkv[:, H:], qkv[:, 2 * H:], sv["ao"], sv["lse"], dao, delta, dqkv, dqkv[:, H:], dqkv[:, 2 * H:],
                         ldq=3 * H, ldk=3 * H, ldv=3 * H, ldo=H, ld_dq=3 * H, ld_dk=3 * H, ld_dv=3 * H, total_rows=M,
                         dropout_p=pa, dropout_seed=seed, dropout_offset=site_offset(i, 0), **ctx["geom"])
            for j, nm in enumerate(("query", "key", "value")):
                dj = dqkv[:, j * H:(j + 1) * H]
                ops.colsum(dj, G[lp + f"attention.self.{nm}.bias"], M=M, N=H, ld=3 * H)
                self._wgrad(dj, sv["xb"], G[lp + f"attention.self.{nm}.weight"], n_out=H, k_out=H, rows=M, lda=3 * H)
            ops.gemm(dqkv, w["qkv"], M=M, N=H, K=3 * H, b_major=1, epilogue=ops.EPI_RESID_F32, C_out=d1, resid=d1)  # dx
            dx = d1
            ctx["saved"][i] = None
        # embeddings: x0 = dropout(LN(e))
        if ph > 0:
            ops.dropout(dx, n=M * H, p=ph, seed=seed, offset=site_offset(-1, 3), y_f32=dx)
        de = torch.empty(M, H, device=dev)
        ops.ln_bwd(M, H, g_f32=dx, gamma=P["embeddings.LayerNorm.weight"], xhat=ctx["emb"]["xhat"], rstd=ctx["emb"]["rstd"],
                   dx_f32=de, dgamma=G["embeddings.LayerNorm.weight"], dbeta=G["embeddings.LayerNorm.bias"])
        ops.bert_embed_bwd(ctx["ids"], de, G["embeddings.word_embeddings.weight"], G["embeddings.position_embeddings.weight"],
                           M, n, H)
        ops.colsum(de, G["embeddings.token_type_embeddings.weight"], M=M, N=H)      # row 0 of the (2, H) table

    def _wgrad(self, dY, X, out, *, n_out, k_out, rows, lda=None):
        tiles = ((n_out + 127) // 128) * ((k_out + 127) // 128)
        ops.gemm(dY, X, M=n_out, N=k_out, K=rows, a_major=1, b_major=1, epilogue=ops.EPI_ATOMIC_F32, C_out=out,
                 ldc=out.stride(0), lda=lda, splits=ops.wgrad_splits(rows, tiles))
