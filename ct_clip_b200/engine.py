"""CTViT encoder engine: orchestrates the sm_100a kernels for the forward and the hand-written
backward of the factorised 3-D ViT (reference: transformer_maskgit/transformer_maskgit/ctvit.py:282-307,
:353-412 and attention.py:312-333).

Data layout in HBM (per rank, b volumes, M = b*T*H*W tokens):
  * ONE canonical token order (b, t, h, w) for the whole encoder; the spatial and temporal stacks
    address it with strides (no '(b t)(h w) d <-> (b h w) t d' copies, ctvit.py:291/297/301/305);
  * residual stream fp32 [M, D] (what the reference keeps under autocast), GEMM operands bf16;
  * LayerNorm affines are folded into the following Linear at weight-preparation time, so the
    saved-for-backward tensor of a pre-norm is the standardised row x_hat (bf16) + rstd.
This module contains no arithmetic of its own: every tensor op is a call into libctclip_b200.so.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import ops


@dataclass
class ViTGeom:
    dim: int
    codebook_size: int
    image_hw: tuple
    patch_hw: tuple
    temporal_patch: int
    spatial_depth: int
    temporal_depth: int
    dim_head: int
    heads: int
    channels: int

    @property
    def H(self):
        return self.image_hw[0] // self.patch_hw[0]

    @property
    def W(self):
        return self.image_hw[1] // self.patch_hw[1]

    @property
    def S(self):
        return self.H * self.W

    @property
    def inner(self):
        return self.heads * self.dim_head

    @property
    def ff_inner(self):  # attention.py:45
        return int(4 * (2 / 3) * self.dim)

    @property
    def ff_pad(self):  # zero-padded so that GEMM tiles / vector accesses stay aligned (multiple of 64)
        return (self.ff_inner + 63) // 64 * 64

    @property
    def patch_voxels(self):
        return self.channels * self.temporal_patch * self.patch_hw[0] * self.patch_hw[1]


class _LayerW:
    """bf16 GEMM operands + fp32 side vectors of one transformer layer (rebuilt after every optimiser step)."""
    __slots__ = ("wq", "bq", "wkv", "wo", "w1", "b1", "w2", "qkb")


class _Saved:
    """Activations one layer keeps for its backward."""
    __slots__ = ("x_in", "xhat1", "xb1", "rstd1", "q_raw", "qh", "kv_raw", "kh", "o", "lse", "xhat2", "rstd2", "h", "g")


class CTViTEngine:
    def __init__(self, geom: ViTGeom, device):
        g = self.g = geom
        self.device = device
        assert g.dim % 128 == 0 and g.dim <= 768, "dim must be a multiple of 128 (<= 768)"
        assert g.dim_head == 32, "the sm_100a attention kernels are specialised for dim_head = 32 (run_train.py:25)"
        assert g.patch_hw[1] % 2 == 0 and (g.patch_voxels * 2) % 16 == 0
        F, Fp = g.ff_inner, g.ff_pad
        # GEGLU interleave: prepared row 2j = value_j (orig row j), 2j+1 = gate_j (orig row F+j); rows >= 2F are zero
        rm = torch.full((2 * Fp,), -1, dtype=torch.int32)
        j = torch.arange(F, dtype=torch.int32)
        rm[0:2 * F:2] = j
        rm[1:2 * F:2] = F + j
        self.geglu_map = rm.to(device)
        self.spatial_w = [_LayerW() for _ in range(g.spatial_depth)]
        self.temporal_w = [_LayerW() for _ in range(g.temporal_depth)]
        self.wp = self.bp = self.ehat = None
        self._alloc_weights()
        R = (2 * g.H - 1) * (2 * g.W - 1)
        self.cpb_x = torch.empty(R, 2, device=device)
        ops.cpb_inputs(self.cpb_x, g.H, g.W)
        self._canon = {}
        self._prep = ops.PrepBatch()
        self.fused_geglu_bwd = False
        # spatial attention on tcgen05 / TMEM (csrc/attention_tc.cu) whenever the token grid allows it (bit 0 forward, bit 1 backward);
        # CTCLIP_ATTN_TC=0 keeps the mma.sync kernels of csrc/attention.cu (A/B measurements, debugging)
        import os
        tc = ops.attn_tc_supported(g.S, g.H, g.W, g.dim_head) if os.environ.get("CTCLIP_ATTN_TC", "1") != "0" else 0
        self.tc_fwd, self.tc_bwd = bool(tc & 1), bool(tc & 2)
        # table gradient of the tcgen05 backward: False = bf16 spill + reduction over the sequences, True = fp32 red.global.add into
        # an L2-resident transposed [heads, S, S] table (one binning pass per step)
        self.tc_dbias_red = os.environ.get("CTCLIP_ATTN_DBIAS_RED", "0") == "1"
        # data parallelism: callable(prefix) invoked when every gradient of the parameters under `prefix` is final (the trainer
        # starts their all-reduce while the rest of the backward is still running)
        self.on_grads_ready = None

    def _canon_table(self, T):
        """canon(f) of the temporal stack's PEG (SURVEY trap T1) as an int32 lookup table (index prep, built once per T)."""
        if T not in self._canon:
            g = self.g
            f = torch.arange(T * g.H * g.W, dtype=torch.int64)
            it, iw, ih = f % T, (f // T) % g.W, f // (T * g.W)
            self._canon[T] = ((it * g.H + ih) * g.W + iw).to(torch.int32).to(self.device)
        return self._canon[T]

    # ------------------------------------------------------------------------------------------
    def _alloc_weights(self):
        g, dev = self.g, self.device
        bf = dict(dtype=torch.bfloat16, device=dev)
        for lw in self.spatial_w + self.temporal_w:
            lw.wq = torch.empty(g.inner, g.dim, **bf)
            lw.bq = torch.empty(g.inner, device=dev)
            lw.wkv = torch.empty(2 * g.inner, g.dim, **bf)
            lw.wo = torch.empty(g.dim, g.inner, **bf)
            lw.w1 = torch.empty(2 * g.ff_pad, g.dim, **bf)
            lw.b1 = torch.empty(2 * g.ff_pad, device=dev)
            lw.w2 = torch.empty(g.dim, g.ff_pad, **bf)
            lw.qkb = torch.empty(1, device=dev)
        self.wp = torch.empty(g.dim, g.patch_voxels, **bf)
        self.bp = torch.empty(g.dim, device=dev)
        self.ehat = torch.empty(g.codebook_size, g.dim, **bf)

    def prepare_weights(self, P: dict):
        """fp32 master parameters (reference state-dict names, prefix stripped) -> bf16 GEMM operands; ONE launch."""
        g = self.g
        D, I, F, Fp = g.dim, g.inner, g.ff_inner, g.ff_pad
        pb = self._prep
        for stack, lws in (("enc_spatial_transformer", self.spatial_w), ("enc_temporal_transformer", self.temporal_w)):
            for i, lw in enumerate(lws):
                a, f = f"{stack}.layers.{i}.1.", f"{stack}.layers.{i}.3."
                pb.weight(P[a + "to_q.weight"], lw.wq, K=D, Np=I, Kp=D, gamma=P[a + "norm.gamma"])
                pb.bias(P[a + "to_q.weight"], lw.bq, K=D, Np=I, beta=P[a + "norm.beta"])
                pb.weight(P[a + "to_kv.weight"], lw.wkv, K=D, Np=2 * I, Kp=D)
                pb.weight(P[a + "to_out.weight"], lw.wo, K=I, Np=D, Kp=I)
                pb.weight(P[f + "1.weight"], lw.w1, K=D, Np=2 * Fp, Kp=D, gamma=P[f + "0.weight"], rowmap=self.geglu_map)
                pb.bias(P[f + "1.weight"], lw.b1, K=D, Np=2 * Fp, beta=P[f + "0.bias"], rowmap=self.geglu_map)
                pb.weight(P[f + "4.weight"], lw.w2, K=F, Np=D, Kp=Fp)
        pb.weight(P["to_patch_emb.2.weight"], self.wp, K=g.patch_voxels, Np=D, Kp=g.patch_voxels,
                  gamma=P["to_patch_emb.1.weight"])
        pb.bias(P["to_patch_emb.2.weight"], self.bp, K=g.patch_voxels, Np=D, beta=P["to_patch_emb.1.bias"],
                bias_in=P["to_patch_emb.2.bias"])
        pb.run()
        if self.tc_fwd:
            for i, lw in enumerate(self.spatial_w):   # logit bound of the fixed-reference softmax (attention_tc.cu)
                a = f"enc_spatial_transformer.layers.{i}.1."
                ops.qk_bound(P[a + "q_scale"], P[a + "k_scale"], lw.qkb, g.dim_head)
        self.prepare_codebook(P)

    def prepare_codebook(self, P):
        g = self.g
        ops.l2norm_rows_bf16(P["vq._codebook.embed"], self.ehat, g.codebook_size, g.dim)

    # ------------------------------------------------------------------------------------------
    def _ds_scratch(self, b, T):
        """bf16 [b*T*heads, S, S] scratch for the d-logits spill of the spatial backward (1 GB at configs[1]); kept across steps."""
        g = self.g
        n = b * T * g.heads * g.S * g.S
        if getattr(self, "_ds_buf", None) is None or self._ds_buf.numel() < n:
            self._ds_buf = torch.empty(n, dtype=torch.bfloat16, device=self.device)
        return self._ds_buf

    def _attn_geom(self, b, T, temporal):
        g = self.g
        if not temporal:
            return dict(n=g.S, heads=g.heads, num_seqs=b * T, seq_inner=1, seq_outer_stride=g.S, tok_stride=1)
        return dict(n=T, heads=g.heads, num_seqs=b * g.S, seq_inner=g.S, seq_outer_stride=T * g.S, tok_stride=g.S)

    def _cpb_forward(self, P, save, want_dense=False):
        """Continuous position bias: MLP on the (2H-1)(2W-1) distinct offsets -> table [R, heads]. The tcgen05 attention
        kernels read the table itself; the mma.sync kernels need it expanded to [heads, S, S] (+ fragment-ordered copies)."""
        g, dev = self.g, self.device
        R = self.cpb_x.shape[0]
        pre = "spatial_rel_pos_bias.net."
        h1 = torch.empty(R, g.dim, device=dev)
        h2 = torch.empty(R, g.dim, device=dev)
        tab = torch.empty(R, g.heads, device=dev)
        ops.sgemm(self.cpb_x, P[pre + "0.0.weight"], h1, M=R, N=g.dim, K=2, trans_b=True, bias=P[pre + "0.0.bias"], act=1)
        ops.sgemm(h1, P[pre + "1.0.weight"], h2, M=R, N=g.dim, K=g.dim, trans_b=True, bias=P[pre + "1.0.bias"], act=1)
        ops.sgemm(h2, P[pre + "2.weight"], tab, M=R, N=g.heads, K=g.dim, trans_b=True, bias=P[pre + "2.bias"])
        bias = frags = None
        need_dense = not (self.tc_fwd and (self.tc_bwd or not save))
        if need_dense or want_dense:
            bias = torch.empty(g.heads, g.S, g.S, dtype=torch.bfloat16, device=dev)     # natural layout: dbias kernel, taps
            ops.cpb_expand(tab, g.heads, g.H, g.W, bias, None)
        if need_dense:
            nfrag = ops.frag_elems(g.heads, g.S)
            bias_frag = torch.empty(nfrag, dtype=torch.bfloat16, device=dev)             # MMA-fragment order (fwd / dQ)
            bias_t_frag = torch.empty(nfrag, dtype=torch.bfloat16, device=dev)           # transposed, fragment order (dK/dV)
            ops.cpb_expand_frag(tab, g.heads, g.H, g.W, bias_frag, bias_t_frag)
            frags = (bias_frag, bias_t_frag)
        return tab, bias, frags, (h1, h2)

    def _cpb_backward(self, P, G, dbias, hs, dtab=None):
        """dbias: fp32 [heads,S,S] (mma.sync path) or None with dtab [R, heads] already reduced (tcgen05 path)."""
        g, dev = self.g, self.device
        R = self.cpb_x.shape[0]
        pre = "spatial_rel_pos_bias.net."
        h1, h2 = hs
        if dtab is None:
            dtab = torch.empty(R, g.heads, device=dev)
            ops.cpb_reduce(dbias, g.heads, g.H, g.W, dtab)
        # layer 2: tab = h2 W2^T + b2
        ops.sgemm(dtab, h2, G[pre + "2.weight"], M=g.heads, N=g.dim, K=R, trans_a=True, accumulate=True)
        ops.colsum(dtab, G[pre + "2.bias"], M=R, N=g.heads)
        dh2 = torch.empty(R, g.dim, device=dev)
        ops.sgemm(dtab, P[pre + "2.weight"], dh2, M=R, N=g.dim, K=g.heads, mask_ref=h2)
        ops.sgemm(dh2, h1, G[pre + "1.0.weight"], M=g.dim, N=g.dim, K=R, trans_a=True, accumulate=True)
        ops.colsum(dh2, G[pre + "1.0.bias"], M=R, N=g.dim)
        dh1 = torch.empty(R, g.dim, device=dev)
        ops.sgemm(dh2, P[pre + "1.0.weight"], dh1, M=R, N=g.dim, K=g.dim, mask_ref=h1)
        ops.sgemm(dh1, self.cpb_x, G[pre + "0.0.weight"], M=g.dim, N=2, K=R, trans_a=True, accumulate=True)
        ops.colsum(dh1, G[pre + "0.0.bias"], M=R, N=g.dim)

    # ------------------------------------------------------------------------------------------
    def _layer_forward(self, x, P, pre, lw, b, T, temporal, bias, save, frags=None, tab=None):
        """x: fp32 stream [M, D] (consumed). Returns (new stream, saved-or-None)."""
        g, dev = self.g, self.device
        M, D, I, Fp = x.shape[0], g.dim, g.inner, g.ff_pad
        bf = dict(dtype=torch.bfloat16, device=dev)
        sv = _Saved() if save else None
        x1 = torch.empty_like(x)
        ops.peg_fwd(x, x1, P[pre + "0.dsconv.weight"], P[pre + "0.dsconv.bias"], B=b, T=T, H=g.H, W=g.W, D=D,
                    temporal=temporal, canon_table=self._canon_table(T) if temporal else None)
        xhat1 = torch.empty(M, D, **bf)
        xb1 = torch.empty(M, D, **bf)
        rstd1 = torch.empty(M, device=dev) if save else None
        ops.ln_fwd(x1, M, D, xhat=xhat1, raw=xb1, rstd=rstd1)
        q_raw = torch.empty(M, I, **bf) if save else None
        qh = torch.empty(M, I, **bf)
        ops.gemm(xhat1, lw.wq, M=M, N=I, K=D, epilogue=ops.EPI_L2NORM, C_out=q_raw, ldc=I, C2=qh, bias=lw.bq,
                 norm_cols=I, norm_scale=P[pre + "1.q_scale"])
        kv_raw = torch.empty(M, 2 * I, **bf)
        kh = torch.empty(M, I, **bf)
        ops.gemm(xb1, lw.wkv, M=M, N=2 * I, K=D, epilogue=ops.EPI_L2NORM, C_out=kv_raw, C2=kh, norm_cols=I,
                 norm_scale=P[pre + "1.k_scale"])
        o = torch.empty(M, I, **bf)
        lse = torch.empty(M, g.heads, device=dev) if save else None
        v = kv_raw[:, I:]
        if tab is not None and self.tc_fwd and not temporal:     # tcgen05 / TMEM kernel, bias from the table
            # (also when only the FORWARD kernel takes the grid, e.g. 32 x 32: both families store lse = log2 sum 2^logit per
            # (row, head), so the mma.sync backward runs on the tcgen05 forward's outputs)
            ops.attn_fwd(qh, kh, v, o, lse, ldq=I, ldk=I, ldv=2 * I, ldo=I, cpb_table=tab, grid_hw=(g.H, g.W), qk_bound=lw.qkb,
                         **self._attn_geom(b, T, temporal))
        else:
            ops.attn_fwd(qh, kh, v, o, lse, ldq=I, ldk=I, ldv=2 * I, ldo=I, bias=bias, bias_frag=frags[0] if frags else None,
                         bias_t_frag=frags[1] if frags else None, **self._attn_geom(b, T, temporal))
        ops.gemm(o, lw.wo, M=M, N=D, K=I, epilogue=ops.EPI_RESID_F32, C_out=x1, resid=x1)          # x2 (in place)
        xhat2 = torch.empty(M, D, **bf)
        rstd2 = torch.empty(M, device=dev) if save else None
        ops.ln_fwd(x1, M, D, xhat=xhat2, rstd=rstd2)
        h = torch.empty(M, 2 * Fp, **bf) if save else None
        gg = torch.empty(M, Fp, **bf)
        ops.gemm(xhat2, lw.w1, M=M, N=2 * Fp, K=D, epilogue=ops.EPI_GEGLU, C_out=h, ldc=2 * Fp, C2=gg, bias=lw.b1)
        ops.gemm(gg, lw.w2, M=M, N=D, K=Fp, epilogue=ops.EPI_RESID_F32, C_out=x1, resid=x1)         # x3 (in place)
        if save:
            sv.x_in, sv.xhat1, sv.xb1, sv.rstd1 = x, xhat1, xb1, rstd1
            sv.q_raw, sv.qh, sv.kv_raw, sv.kh, sv.o, sv.lse = q_raw, qh, kv_raw, kh, o, lse
            sv.xhat2, sv.rstd2, sv.h, sv.g = xhat2, rstd2, h, gg
        return x1, sv

    def forward(self, video, P, *, save: bool, taps: dict | None = None):
        """video [b, c, F, Hi, Wi] (fp32 in [-1,1] or int16 HU) -> ctx with pre-VQ tokens and code indices."""
        g, dev = self.g, self.device
        b, c, Fr, Hi, Wi = video.shape
        assert (Hi, Wi) == tuple(g.image_hw) and c == g.channels and Fr % g.temporal_patch == 0
        T = Fr // g.temporal_patch
        M, D, Pv = b * T * g.S, g.dim, g.patch_voxels
        bf = dict(dtype=torch.bfloat16, device=dev)
        ctx = dict(b=b, T=T, M=M, save=save)
        # --- patch embedding (ctvit.py:170-175)
        xhat_p = torch.empty(M, Pv, **bf)
        ops.patchify(video, xhat_p, B=b, Cc=c, F=Fr, H=Hi, W=Wi, pt=g.temporal_patch, p1=g.patch_hw[0], p2=g.patch_hw[1])
        y0 = torch.empty(M, D, device=dev)
        ops.gemm(xhat_p, self.wp, M=M, N=D, K=Pv, epilogue=ops.EPI_F32, C_out=y0, bias=self.bp)
        x = torch.empty(M, D, device=dev)
        xhat3 = torch.empty(M, D, **bf) if save else None
        rstd3 = torch.empty(M, device=dev) if save else None
        ops.ln_fwd(y0, M, D, gamma=P["to_patch_emb.3.weight"], beta=P["to_patch_emb.3.bias"], y_f32=x, xhat=xhat3, rstd=rstd3)
        del y0
        if taps is not None:
            taps["patch_tokens"] = x.clone()
        ctx.update(xhat_p=xhat_p if save else None, xhat3=xhat3, rstd3=rstd3)
        # --- spatial stack (ctvit.py:291-297)
        tab, bias, bias_t, cpb_h = self._cpb_forward(P, save, want_dense=taps is not None)
        if taps is not None:
            taps["cpb_bias"] = bias.float()
        ctx.update(bias=bias, bias_t=bias_t, cpb_tab=tab, cpb_h=cpb_h if save else None)
        saved_s, saved_t = [], []
        for i, lw in enumerate(self.spatial_w):
            x, sv = self._layer_forward(x, P, f"enc_spatial_transformer.layers.{i}.", lw, b, T, False, bias, save, frags=bias_t,
                                        tab=tab)
            saved_s.append(sv)
            if taps is not None:
                taps[f"spatial.{i}"] = x.clone()
        y = torch.empty_like(x)
        xhat_ns = torch.empty(M, D, **bf) if save else None
        rstd_ns = torch.empty(M, device=dev) if save else None
        ops.ln_fwd(x, M, D, gamma=P["enc_spatial_transformer.norm_out.gamma"], beta=P["enc_spatial_transformer.norm_out.beta"],
                   y_f32=y, xhat=xhat_ns, rstd=rstd_ns)
        x = y
        if taps is not None:
            taps["spatial_out"] = x.clone()
        # --- temporal stack (ctvit.py:301-305), same canonical layout, strided sequences
        for i, lw in enumerate(self.temporal_w):
            x, sv = self._layer_forward(x, P, f"enc_temporal_transformer.layers.{i}.", lw, b, T, True, None, save)
            saved_t.append(sv)
            if taps is not None:
                taps[f"temporal.{i}"] = x.clone()
        y = torch.empty_like(x)
        yb = torch.empty(M, D, **bf)
        xhat_nt = torch.empty(M, D, **bf) if save else None
        rstd_nt = torch.empty(M, device=dev) if save else None
        ops.ln_fwd(x, M, D, gamma=P["enc_temporal_transformer.norm_out.gamma"], beta=P["enc_temporal_transformer.norm_out.beta"],
                   y_f32=y, y_bf16=yb, xhat=xhat_nt, rstd=rstd_nt)
        del x
        # --- vector quantiser lookup (ctvit.py:403): argmax of cosine similarity fused in the GEMM epilogue
        # (bf16 operands decide ~5 % of the near-ties differently from the fp32 quantiser: the epilogue also reports the runner-up
        # and ctclip_vq_rerank re-ranks the pair in fp32 against the un-rounded token and code-book rows)
        idx = torch.empty(M, dtype=torch.int32, device=dev)
        idx2 = torch.empty(M, dtype=torch.int32, device=dev)
        ops.gemm(yb, self.ehat, M=M, N=g.codebook_size, K=D, epilogue=ops.EPI_ARGMAX, arg_out=idx, arg2_out=idx2)
        ops.vq_rerank(y, P["vq._codebook.embed"], idx, idx2, M, D)
        ctx.update(saved_s=saved_s, saved_t=saved_t, xhat_ns=xhat_ns, rstd_ns=rstd_ns, xhat_nt=xhat_nt, rstd_nt=rstd_nt,
                   pre_vq=y, idx=idx)
        return ctx

    def vq_ema(self, ctx, P, decay=0.8, all_reduce=None):
        """Training-mode code-book EMA (vector_quantize_pytorch 1.1.2 CosineSimCodebook.forward, training branch).
        all_reduce: optional callable(tensor) summing statistics across data-parallel ranks."""
        g, dev = self.g, self.device
        stats = torch.zeros(g.codebook_size * (g.dim + 1), device=dev)       # [bins | embed_sum]: ONE all-reduce under data parallelism
        bins = stats[:g.codebook_size]
        esum = stats[g.codebook_size:].view(g.codebook_size, g.dim)
        ops.vq_ema_accum(ctx["pre_vq"], ctx["idx"], bins, esum, ctx["M"], g.dim)
        if all_reduce is not None:
            all_reduce(stats)
        ops.vq_ema_update(P["vq._codebook.embed"], P["vq._codebook.cluster_size"], bins, esum, g.codebook_size, g.dim, decay)

    # ------------------------------------------------------------------------------------------
    def _wgrad(self, dY, X, out, *, n_out, k_out, rows, ld_out=None):
        """out[n_out, k_out] (fp32, accumulated) += dY[rows, n_out]^T @ X[rows, k_out]   (both bf16, token-major)."""
        tiles = ((n_out + 127) // 128) * ((k_out + 127) // 128)
        ops.gemm(dY, X, M=n_out, N=k_out, K=rows, a_major=1, b_major=1, epilogue=ops.EPI_ATOMIC_F32, C_out=out,
                 ldc=ld_out if ld_out is not None else out.stride(0), splits=ops.wgrad_splits(rows, tiles))

    def _layer_backward(self, dres, dxb, sv, P, G, pre, lw, b, T, temporal, bias, bias_t, dbias, tab=None, dtab=None):
        """dres: fp32 [M,D] gradient w.r.t. the layer output (consumed); dxb: its bf16 copy.
        Returns (gradient w.r.t. the layer input fp32, its bf16 copy)."""
        g, dev = self.g, self.device
        M, D, I, F, Fp = dres.shape[0], g.dim, g.inner, g.ff_inner, g.ff_pad
        bf = dict(dtype=torch.bfloat16, device=dev)
        a, f = pre + "1.", pre + "3."
        # ---- feed-forward: x3 = x2 + g W2^T,  (h, g) = GEGLU(xhat2 W1'^T + b1')
        self._wgrad(dxb, sv.g, G[f + "4.weight"], n_out=D, k_out=F, rows=M)
        s1 = torch.zeros(2 * Fp, device=dev)
        # GEMM epilogue 8 fuses these two kernels (tests/test_gemm_gpu.py::test_gemm_geglu_bwd_fused) but measured 0.66 ms
        # per layer against 0.15 + 0.33 ms for the pair on B200 (the 8-warp epilogue becomes the bottleneck: ~25
        # instructions per (value, gate) pair + column-sum shuffles): kept opt-in.
        if self.fused_geglu_bwd and Fp <= 1536:    # dg = dxb W2 and the GEGLU backward in ONE kernel
            ops.gemm(dxb, lw.w2, M=M, N=Fp, K=D, b_major=1, epilogue=ops.EPI_GEGLU_BWD, C_out=sv.h, ldc=2 * Fp, colsum=s1)
        else:             # wider feed-forward (dim 768): the shared-memory column-sum accumulator does not fit
            dg = torch.empty(M, Fp, **bf)
            ops.gemm(dxb, lw.w2, M=M, N=Fp, K=D, b_major=1, epilogue=ops.EPI_BF16, C_out=dg)
            ops.geglu_bwd(dg, sv.h, M=M, n_pairs=Fp, colsum_out=s1)          # sv.h now holds dh
            del dg
        G1 = torch.zeros(2 * Fp, D, device=dev)
        self._wgrad(sv.h, sv.xhat2, G1, n_out=2 * Fp, k_out=D, rows=M)
        ops.unprep_wgrad(G1, P[f + "1.weight"], G[f + "1.weight"], K=D, Np=2 * Fp, gamma=P[f + "0.weight"],
                         rowmap=self.geglu_map, s=s1, dgamma=G[f + "0.weight"], dbeta=G[f + "0.bias"])
        dxh = torch.empty(M, D, **bf)
        ops.gemm(sv.h, lw.w1, M=M, N=D, K=2 * Fp, b_major=1, epilogue=ops.EPI_BF16, C_out=dxh)
        sv.h = None
        ops.ln_bwd(M, D, g_bf16=dxh, xhat=sv.xhat2, rstd=sv.rstd2, dres_in=dres, dx_f32=dres, dx_bf16=dxb)
        # ---- attention output projection: x2 = x1 + o Wo^T
        self._wgrad(dxb, sv.o, G[a + "to_out.weight"], n_out=D, k_out=I, rows=M)
        d_o = torch.empty(M, I, **bf)
        ops.gemm(dxb, lw.wo, M=M, N=I, K=D, b_major=1, epilogue=ops.EPI_BF16, C_out=d_o)
        # ---- attention core
        dqh = torch.empty(M, I, **bf)
        dkv = torch.empty(M, 2 * I, **bf)       # [dk_hat | dv]
        delta = torch.empty(M, g.heads, device=dev)
        if dtab is not None and self.tc_dbias_red:      # tcgen05 / TMEM kernel; d bias accumulated by fp32 reductions into dtab = [h][j][i]
            ops.attn_bwd(sv.qh, sv.kh, sv.kv_raw[:, I:], sv.o, sv.lse, d_o, delta, dqh, dkv, dkv[:, I:], ldq=I, ldk=I,
                         ldv=2 * I, ldo=I, ld_dq=I, ld_dk=2 * I, ld_dv=2 * I, total_rows=M, cpb_table=tab, grid_hw=(g.H, g.W),
                         dbias=dtab, **self._attn_geom(b, T, temporal))
        elif dtab is not None:      # tcgen05 / TMEM kernel: one pass, d logits spilled (bf16) + reduced over the sequences into dtab
            ops.attn_bwd(sv.qh, sv.kh, sv.kv_raw[:, I:], sv.o, sv.lse, d_o, delta, dqh, dkv, dkv[:, I:], ldq=I, ldk=I,
                         ldv=2 * I, ldo=I, ld_dq=I, ld_dk=2 * I, ld_dv=2 * I, total_rows=M, cpb_table=tab, grid_hw=(g.H, g.W),
                         dcpb_table=dtab, ds_scratch=self._ds_scratch(b, T), **self._attn_geom(b, T, temporal))
        else:
            ops.attn_bwd(sv.qh, sv.kh, sv.kv_raw[:, I:], sv.o, sv.lse, d_o, delta, dqh, dkv, dkv[:, I:], ldq=I, ldk=I,
                         ldv=2 * I, ldo=I, ld_dq=I, ld_dk=2 * I, ld_dv=2 * I, total_rows=M, bias=bias,
                         bias_frag=bias_t[0] if bias_t else None, bias_t_frag=bias_t[1] if bias_t else None,
                         dbias=dbias, ds_scratch=self._ds_scratch(b, T) if dbias is not None else None,
                         **self._attn_geom(b, T, temporal))
        # ---- l2norm * scale backward (attention.py:152-154); dq/dk overwritten with raw-projection gradients
        ops.l2norm_bwd(dqh, I, sv.q_raw, I, P[a + "q_scale"], dqh, I, G[a + "q_scale"], M, g.heads)
        ops.l2norm_bwd(dkv, 2 * I, sv.kv_raw, 2 * I, P[a + "k_scale"], dkv, 2 * I, G[a + "k_scale"], M, g.heads)
        # ---- projections: q = xhat1 Wq'^T (LayerNorm path), [k|v] = x1 Wkv^T (raw stream, attention.py:139-145)
        Gq = torch.zeros(I, D, device=dev)
        self._wgrad(dqh, sv.xhat1, Gq, n_out=I, k_out=D, rows=M)
        ops.unprep_wgrad(Gq, P[a + "to_q.weight"], G[a + "to_q.weight"], K=D, Np=I, gamma=P[a + "norm.gamma"],
                         dgamma=G[a + "norm.gamma"])
        self._wgrad(dkv, sv.xb1, G[a + "to_kv.weight"], n_out=2 * I, k_out=D, rows=M)
        ops.gemm(dqh, lw.wq, M=M, N=D, K=I, b_major=1, epilogue=ops.EPI_BF16, C_out=dxh)
        ops.ln_bwd(M, D, g_bf16=dxh, xhat=sv.xhat1, rstd=sv.rstd1, dres_in=dres, dx_f32=dres)
        ops.gemm(dkv, lw.wkv, M=M, N=D, K=2 * I, b_major=1, epilogue=ops.EPI_RESID_F32, C_out=dres, resid=dres)
        # ---- PEG: x1 = x0 + conv(x0)
        ct = self._canon_table(T) if temporal else None
        ops.peg_bwd_weight(sv.x_in, dres, G[pre + "0.dsconv.weight"], G[pre + "0.dsconv.bias"], B=b, T=T, H=g.H, W=g.W,
                           D=D, temporal=temporal, canon_table=ct)
        dx0 = torch.empty_like(dres)
        ops.peg_bwd_data(dres, dx0, P[pre + "0.dsconv.weight"], dx_bf16=dxb, B=b, T=T, H=g.H, W=g.W, D=D, temporal=temporal,
                         canon_table=ct)
        return dx0, dxb

    def backward(self, ctx, dtok, P, G):
        """dtok: fp32 [M, D] gradient w.r.t. the (straight-through) quantiser input, i.e. the temporal norm_out output.
        Accumulates parameter gradients into G (fp32 tensors keyed like P)."""
        g, dev = self.g, self.device
        b, T, M, D = ctx["b"], ctx["T"], ctx["M"], g.dim
        bf = dict(dtype=torch.bfloat16, device=dev)
        dres = torch.empty(M, D, device=dev)
        dxb = torch.empty(M, D, **bf)
        ops.ln_bwd(M, D, g_f32=dtok, gamma=P["enc_temporal_transformer.norm_out.gamma"], xhat=ctx["xhat_nt"],
                   rstd=ctx["rstd_nt"], dx_f32=dres, dx_bf16=dxb, dgamma=G["enc_temporal_transformer.norm_out.gamma"])
        for i in reversed(range(g.temporal_depth)):
            dres, dxb = self._layer_backward(dres, dxb, ctx["saved_t"][i], P, G, f"enc_temporal_transformer.layers.{i}.",
                                             self.temporal_w[i], b, T, True, None, None, None)
            ctx["saved_t"][i] = None
            if self.on_grads_ready is not None:
                self.on_grads_ready(f"enc_temporal_transformer.layers.{i}.")
        d2 = torch.empty_like(dres)
        ops.ln_bwd(M, D, g_f32=dres, gamma=P["enc_spatial_transformer.norm_out.gamma"], xhat=ctx["xhat_ns"],
                   rstd=ctx["rstd_ns"], dx_f32=d2, dx_bf16=dxb, dgamma=G["enc_spatial_transformer.norm_out.gamma"])
        dres = d2
        dbias = dtab = None
        if self.tc_bwd and self.tc_dbias_red:
            dtab = torch.zeros(g.heads, g.S, g.S, device=dev)        # TRANSPOSED d bias [h][j][i], L2-resident (10.6 MB at S = 576)
        elif self.tc_bwd:
            dtab = torch.zeros(self.cpb_x.shape[0], g.heads, device=dev)
        else:
            dbias = torch.zeros(g.heads, g.S, g.S, device=dev)
        for i in reversed(range(g.spatial_depth)):
            dres, dxb = self._layer_backward(dres, dxb, ctx["saved_s"][i], P, G, f"enc_spatial_transformer.layers.{i}.",
                                             self.spatial_w[i], b, T, False, ctx["bias"], ctx["bias_t"], dbias,
                                             tab=ctx["cpb_tab"], dtab=dtab)
            ctx["saved_s"][i] = None
            if self.on_grads_ready is not None:
                self.on_grads_ready(f"enc_spatial_transformer.layers.{i}.")
        if self.tc_bwd and self.tc_dbias_red:
            dbt, dtab = dtab, torch.zeros(self.cpb_x.shape[0], g.heads, device=dev)
            ops.cpb_reduce_t(dbt, g.heads, g.H, g.W, dtab)
        self._cpb_backward(P, G, dbias, ctx["cpb_h"], dtab=dtab)
        # patch embedding: x = LN_D(xhat_p Wp'^T + bp')
        ops.ln_bwd(M, D, g_f32=dres, gamma=P["to_patch_emb.3.weight"], xhat=ctx["xhat3"], rstd=ctx["rstd3"], dx_bf16=dxb,
                   dgamma=G["to_patch_emb.3.weight"], dbeta=G["to_patch_emb.3.bias"])
        sp = torch.zeros(D, device=dev)
        ops.colsum(dxb, sp, M=M, N=D)
        Gp = torch.zeros(D, g.patch_voxels, device=dev)
        self._wgrad(dxb, ctx["xhat_p"], Gp, n_out=D, k_out=g.patch_voxels, rows=M)
        ops.unprep_wgrad(Gp, P["to_patch_emb.2.weight"], G["to_patch_emb.2.weight"], K=g.patch_voxels, Np=D,
                         gamma=P["to_patch_emb.1.weight"], s=sp, dgamma=G["to_patch_emb.1.weight"],
                         dbeta=G["to_patch_emb.1.bias"], dbias=G["to_patch_emb.2.bias"])
