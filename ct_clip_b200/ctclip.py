"""CTCLIP -- drop-in for `ct_clip.CTCLIP` (reference: CT_CLIP/ct_clip/ct_clip.py:407-901) on the path the
reference scripts exercise: injected image encoder (CTViT) + injected text encoder (HF BertModel),
`to_text_latent` / `to_visual_latent` projections, learned temperature, symmetric InfoNCE.

Same constructor keywords, forward signature, return conventions and state-dict layout. The
contrastive step (image tower, pooling, projections, loss, and the whole backward) runs in
libctclip_b200.so; see DESIGN.md for what is out of scope (the x-clip default towers and the
MLM / SimSiam / SimCLR / FILIP / DCL / CLOOB / multiview variants, all disabled in every reference
script -- run_train.py:31-42).
"""
from __future__ import annotations

import copy
from pathlib import Path

import torch
from torch import nn

from . import bert as bert_native
from . import ops
from .ctvit import CTViT, _GradDict

_UNSUPPORTED = ("use_all_token_embeds", "downsample_image_embeds", "decoupled_contrastive_learning",
                "extra_latent_projection", "use_mlm", "use_visual_ssl")


class _Ctx:
    pass


class _ClipStepFn(torch.autograd.Function):
    """(text, volume, parameters) -> contrastive loss, with the hand-written backward.

    text tower: native kernels (cls is None, token ids in `text`) or an external module's CLS output (`cls`).
    The loss kernel produces d(latents) together with the loss, so backward() only has to push those through the
    projections and the towers."""

    @staticmethod
    def forward(ctx, module, need_grad, cls, text, video, names, *params):
        P = dict(zip(names, params))
        vit: CTViT = module.visual_transformer
        PV = {n[len("visual_transformer."):]: P[n] for n in names if n.startswith("visual_transformer.")}
        tctx = None
        if cls is None:   # native BERT
            PT = {n[len("text_transformer."):]: P[n] for n in names if n.startswith("text_transformer.")}
            last, tctx = module._bert_engine().forward(text.input_ids, text.attention_mask, PT, save=need_grad,
                                                       dropout=module._text_dropout())
            cls_in = last[:, 0, :]
            ctx.PT = PT
        else:
            cls_in = cls
        ectx = vit._run_forward(video, PV, save=need_grad)
        st = module._heads_forward(cls_in, ectx, P, want_loss=True, want_grads=need_grad)
        ctx.module, ctx.names, ctx.ectx, ctx.st, ctx.tctx = module, names, ectx, st, tctx
        ctx.native_text = cls is None
        ctx.save_for_backward(cls_in if cls is None else cls, *params)
        return st.loss.view(())

    @staticmethod
    def backward(ctx, gout):
        module, names, st, ectx = ctx.module, ctx.names, ctx.st, ctx.ectx
        cls, *params = ctx.saved_tensors
        P = dict(zip(names, params))
        sink = getattr(module, "_grad_sink", None)
        if sink is not None:   # trainer mode: accumulate straight into the flat gradient arena
            Gd, G = _GradDict(sink), None
        else:
            G = {n: torch.zeros_like(p) for n, p in P.items() if p.requires_grad and p.numel() > 0}
            Gd = _GradDict(G)
        text_bwd = None
        if ctx.native_text:
            tctx, PT = ctx.tctx, ctx.PT

            def text_bwd(dcls_):
                d_last = torch.zeros(tctx["M"], dcls_.shape[1], device=dcls_.device)
                d_last.view(tctx["b"], tctx["n"], -1)[:, 0, :] = dcls_          # only the CLS rows carry gradient (ct_clip.py:762)
                module._bert_engine().backward(tctx, d_last, PT, _PrefixView(Gd, "text_transformer."))
                if module.dp_grad_ready is not None:
                    module.dp_grad_ready("text_transformer.")
        dcls = module._backward_into(st, ectx, P, Gd, cls, float(gout), text_backward=text_bwd)
        if ctx.native_text:
            dcls = None
        ctx.ectx = ctx.st = ctx.tctx = None
        if sink is not None:
            return (None, None, dcls, None, None, None) + (None,) * len(names)
        grads = tuple(G[n] if (n in G and n in Gd.touched) else None for n in names)
        return (None, None, dcls, None, None, None) + grads


class _ClipLatentsFn(torch.autograd.Function):
    """(text, volume, parameters) -> (raw text latents [n_text, L], raw image latents [b, L]) with the hand-written backward of
    both towers: the differentiable building block of everything that is NOT the contrastive step -- fine-tuning on the
    similarity logits (scripts/ct_vocabfine_train.py:108-122) or on the latents. l2-normalisation, temperature and the tiny
    task loss on top are ordinary autograd ops on [n, L] tensors."""

    @staticmethod
    def forward(ctx, module, need_grad, text, video, names, *params):
        P = dict(zip(names, params))
        vit: CTViT = module.visual_transformer
        PV = {n[len("visual_transformer."):]: P[n] for n in names if n.startswith("visual_transformer.")}
        PT = {n[len("text_transformer."):]: P[n] for n in names if n.startswith("text_transformer.")}
        last, tctx = module._bert_engine().forward(text.input_ids, text.attention_mask, PT, save=need_grad,
                                                   dropout=module._text_dropout())
        cls_in = last[:, 0, :]
        ectx = vit._run_forward(video, PV, save=need_grad)
        st = module._heads_forward(cls_in, ectx, P, want_loss=False, want_grads=False, want_raw=True)
        ctx.module, ctx.names, ctx.ectx, ctx.st, ctx.tctx, ctx.PT = module, names, ectx, st, tctx, PT
        ctx.save_for_backward(cls_in, *params)
        return st.t_raw, st.i_raw

    @staticmethod
    def backward(ctx, d_t_raw, d_i_raw):
        module, names, st, ectx, tctx = ctx.module, ctx.names, ctx.st, ctx.ectx, ctx.tctx
        cls, *params = ctx.saved_tensors
        P = dict(zip(names, params))
        sink = getattr(module, "_grad_sink", None)
        if sink is not None:
            Gd, G = _GradDict(sink), None
        else:
            G = {n: torch.zeros_like(p) for n, p in P.items() if p.requires_grad and p.numel() > 0}
            Gd = _GradDict(G)
        dev = cls.device
        st.d_t_raw = (d_t_raw if d_t_raw is not None else torch.zeros_like(st.t_raw)).float().contiguous()
        st.d_i_raw = (d_i_raw if d_i_raw is not None else torch.zeros_like(st.i_raw)).float().contiguous()
        st.dtemp = torch.zeros(1, device=dev)          # the temperature enters downstream of the latents (plain autograd)
        dcls = module._backward_into(st, ectx, P, Gd, cls, 1.0)
        d_last = torch.zeros(tctx["M"], dcls.shape[1], device=dev)
        d_last.view(tctx["b"], tctx["n"], -1)[:, 0, :] = dcls
        module._bert_engine().backward(tctx, d_last, ctx.PT, _PrefixView(Gd, "text_transformer."))
        ctx.ectx = ctx.st = ctx.tctx = None
        if sink is not None:
            return (None, None, None, None, None) + (None,) * len(names)
        grads = tuple(G[n] if (n in G and n in Gd.touched) else None for n in names)
        return (None, None, None, None, None) + grads


class CTCLIP(nn.Module):
    def __init__(self, *, image_encoder=None, text_encoder=None, dim_text=512, dim_image=512, dim_latent=512,
                 num_text_tokens=28897, text_enc_depth=6, text_seq_len=256, text_heads=8, text_dim_head=64,
                 text_has_cls_token=False, text_pad_id=0, text_rotary_pos_emb=False, text_causal_mask=False,
                 text_eos_id=None, text_encode_without_mask=False, visual_enc_depth=6, visual_heads=8,
                 visual_dim_head=64, visual_image_size=256, visual_patch_size=32, visual_patch_dropout=0.5,
                 visual_has_cls_token=False, channels=3, use_all_token_embeds=False, downsample_image_embeds=False,
                 decoupled_contrastive_learning=False, extra_latent_projection=False, use_mlm=False,
                 text_ssl_loss_weight=0.05, use_visual_ssl=False, visual_ssl=None, visual_ssl_type='simsiam',
                 visual_ssl_hidden_layer=-1, simclr_temperature=0.1, image_ssl_loss_weight=0.05,
                 multiview_loss_weight=0.1, checkpoint_during_training=False, **kwargs):
        super().__init__()
        self.dtype = torch.float32
        flags = dict(use_all_token_embeds=use_all_token_embeds, downsample_image_embeds=downsample_image_embeds,
                     decoupled_contrastive_learning=decoupled_contrastive_learning,
                     extra_latent_projection=extra_latent_projection, use_mlm=use_mlm,
                     use_visual_ssl=use_visual_ssl or visual_ssl is not None)
        on = [k for k in _UNSUPPORTED if flags[k]]
        if on:
            raise NotImplementedError(f"CTCLIP options {on} are disabled in every reference script (run_train.py:31-42) "
                                      "and are outside the B200 hot-path build")
        if image_encoder is None or text_encoder is None:
            raise NotImplementedError("the x-clip default towers (ct_clip.py:476-508) are not part of this build: pass "
                                      "image_encoder=CTViT(...) and text_encoder=BertModel(...) as the reference scripts do")
        assert not (text_causal_mask and text_eos_id is None)
        if text_causal_mask:
            raise NotImplementedError("text_causal_mask is unused on the CT-CLIP path")
        self.dim_text, self.dim_image, self.dim_latent = dim_text, dim_image, dim_latent
        self.image_channels, self.image_size = channels, visual_image_size
        self.text_pad_id, self.text_has_cls_token, self.text_seq_len = text_pad_id, text_has_cls_token, text_seq_len
        self.text_encode_without_mask = text_encode_without_mask
        self.text_causal_mask, self.text_eos_id = text_causal_mask, text_eos_id
        self.text_transformer = text_encoder
        self.visual_has_cls_token = visual_has_cls_token
        self.visual_transformer = image_encoder
        self.use_mlm, self.text_ssl_loss_weight = False, 0
        self.use_visual_ssl, self.image_ssl_loss_weight = False, 0
        self.to_text_latent = nn.Linear(dim_text, dim_latent, bias=False)
        self.to_visual_latent = nn.Linear(dim_image, dim_latent, bias=False)
        self.temperature = nn.Parameter(torch.tensor(1.))
        self.use_all_token_embeds = False
        self.decoupled_contrastive_learning = False
        self.extra_latent_projection = False
        self.to_text_latent_extra = copy.deepcopy(self.to_text_latent)          # ct_clip.py:579-581 (never used)
        self.to_visual_latent_extra = copy.deepcopy(self.to_visual_latent)
        self.multiview_loss_weight = multiview_loss_weight
        self.tokenizer = None
        try:  # ct_clip.py:585 -- needs the HF hub; optional here (callers pass token ids)
            from transformers import BertTokenizer
            self.tokenizer = BertTokenizer.from_pretrained('microsoft/BiomedVLP-CXR-BERT-specialized', do_lower_case=True)
        except Exception:
            self.tokenizer = None
        # data-parallel hooks (set by CTClipTrainer): gather latents of all ranks before the loss
        self.dp_rank, self.dp_world = 0, 1
        self.dp_all_gather = None
        self.dp_early_reduce = None      # callable(param name): start the gradient all-reduce of that tensor right away
        self.dp_grad_ready = None        # callable(name prefix): every gradient under the prefix is final (bucketed all-reduce)
        self._wv_bf16 = None
        self._wv_version = None
        self._grad_sink = None
        self._bert = None
        self.force_torch_text = False     # debugging / comparison switch: run the text encoder as a torch module

    # ------------------------------------------------------------------------------------------
    def state_dict(self, *args, **kwargs):
        return super().state_dict(*args, **kwargs)

    def load_state_dict(self, state_dict, *args, **kwargs):
        sd = dict(state_dict)
        # checkpoints written with transformers <= 4.30 carry a persistent position_ids buffer; newer BertModel does not
        key = "text_transformer.embeddings.position_ids"
        if key in sd and key not in super().state_dict():
            sd.pop(key)
        self.mark_weights_dirty()
        return super().load_state_dict(sd, *args, **kwargs)

    def load(self, path, tolerant=True):
        """ct_clip.py:593-597 (torch.load + strict load_state_dict). tolerant=True first normalises wrapper prefixes, the
        position_ids buffer of old transformers releases, absent *_extra copies and foreign GenerateCT keys
        (ct_clip_b200.checkpoint.tolerant_state_dict); the load itself stays strict."""
        path = Path(path)
        assert path.exists()
        sd = torch.load(str(path), map_location="cpu")
        if tolerant:
            from .checkpoint import tolerant_state_dict
            sd, self.last_load_report = tolerant_state_dict(sd, super().state_dict())
        self.load_state_dict(sd)

    def mark_weights_dirty(self):
        self._wv_version = None
        if self._bert is not None:
            self._bert.mark_dirty()
        if isinstance(self.visual_transformer, CTViT):
            self.visual_transformer.mark_weights_dirty()

    def tokenize(self, prompt):
        if self.tokenizer is None:
            raise RuntimeError("no tokenizer available offline; pass token ids")
        return self.tokenizer(prompt, return_tensors="pt", padding="max_length", truncation=True, max_length=512).to("cuda")

    # ------------------------------------------------------------------------------------------
    def _fast_path(self):
        return isinstance(self.visual_transformer, CTViT)

    def _live(self):
        """names/tensors handed to the fused step (own parameters + the image tower's)."""
        names = ["to_text_latent.weight", "to_visual_latent.weight", "temperature"]
        tensors = [self.to_text_latent.weight, self.to_visual_latent.weight, self.temperature]
        vn, vt = self.visual_transformer.named_live_tensors()
        names += ["visual_transformer." + n for n in vn]
        tensors += vt
        if self._text_native():
            for n, p in self.text_transformer.named_parameters():
                names.append("text_transformer." + n)
                tensors.append(p)
        return names, tensors

    def _text_native(self):
        """True when the injected text encoder is a HF BertModel this build runs on its own kernels."""
        tt = self.text_transformer
        return (not self.force_torch_text) and bert_native.supports(tt)

    def _text_dropout(self):
        """dropout argument of the native text tower for this call: a fresh Philox seed per training-mode forward, drawn from
        torch's CPU generator (reproducible under torch.manual_seed, no device synchronisation); None in eval mode / p = 0."""
        tt = self.text_transformer
        if not bert_native.dropout_active(tt):
            return None
        seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
        return bert_native.dropout_config(tt, seed)

    def _bert_engine(self):
        dev = self.to_text_latent.weight.device
        if self._bert is None or self._bert.device != dev:
            self._bert = bert_native.BertEngine(self.text_transformer, dev)
        return self._bert

    def _visual_weight_bf16(self, W):
        ver = (W._version, W.data_ptr())
        if self._wv_version != ver or self._wv_bf16 is None or self._wv_bf16.device != W.device:
            if self._wv_bf16 is None or self._wv_bf16.shape != W.shape or self._wv_bf16.device != W.device:
                self._wv_bf16 = torch.empty(W.shape, dtype=torch.bfloat16, device=W.device)
            ops.cast_bf16(W, self._wv_bf16, W.numel())
            self._wv_version = ver
        return self._wv_bf16

    def _text_cls(self, text):
        """ct_clip.py:685-686 + :762: run the text encoder, take the CLS row. Returns (enc_text, cls)."""
        if self._text_native():
            with torch.no_grad():
                PT = dict(self.text_transformer.named_parameters())
                enc_text, _ = self._bert_engine().forward(text.input_ids, text.attention_mask, PT, save=False,
                                                          dropout=self._text_dropout())
            return enc_text, enc_text[:, 0, :]
        out = self.text_transformer(text.input_ids, attention_mask=text.attention_mask)
        enc_text = out[0]
        return enc_text, enc_text[:, 0, :]

    def _heads_forward(self, cls, ectx, P, *, want_loss, want_grads, want_raw=False):
        """pool + projections (+ loss fwd/bwd). cls: fp32 [b, dim_text]."""
        vit: CTViT = self.visual_transformer
        g = vit.engine.g
        dev = cls.device
        b, T, L = ectx["b"], ectx["T"], self.dim_latent
        K = g.S * g.dim
        assert K == self.dim_image, f"dim_image={self.dim_image} but the image tower yields {g.S}x{g.dim}={K} (ct_clip.py:740,564)"
        st = _Ctx()
        st.pooled_bf16 = torch.empty(b, K, dtype=torch.bfloat16, device=dev)
        ops.vq_gather_pool(ectx["idx"], ectx["P"]["vq._codebook.embed"], B=b, T=T, S=g.S, D=g.dim, pooled_bf16=st.pooled_bf16)
        vit._finish_quantize(ectx)      # code-book EMA (training mode) only after the codes have been read
        wv = self._visual_weight_bf16(P["to_visual_latent.weight"])
        i_raw = torch.zeros(b, L, device=dev)
        kb = (K + 63) // 64
        ops.gemm(st.pooled_bf16, wv, M=b, N=L, K=K, epilogue=ops.EPI_ATOMIC_F32, C_out=i_raw,
                 splits=max(1, min(kb, 2 * 148 // max(1, (L + 255) // 256))))
        cls32 = cls.detach().float().contiguous()
        bt = cls32.shape[0]
        t_raw = torch.empty(bt, L, device=dev)
        ops.sgemm(cls32, P["to_text_latent.weight"], t_raw, M=bt, N=L, K=self.dim_text, trans_b=True)
        st.cls32, st.b, st.L, st.K = cls32, b, L, K
        if want_raw:      # differentiable-latents path: the caller normalises (autograd) and owns the loss
            st.t_raw, st.i_raw = t_raw, i_raw
            return st
        if not want_loss:
            st.t_hat, st.i_hat = self._l2(t_raw), self._l2(i_raw)
            return st
        assert bt == b, "training needs one report per volume"
        if self.dp_world > 1:
            tg, ig = self.dp_all_gather(t_raw, i_raw)
        else:
            tg, ig = t_raw, i_raw
        B = tg.shape[0]
        st.t_hat, st.i_hat = torch.empty(B, L, device=dev), torch.empty(B, L, device=dev)
        inv, sim = torch.empty(2 * B, device=dev), torch.empty(B, B, device=dev)
        if want_loss:
            st.loss, st.dtemp = torch.empty(1, device=dev), torch.empty(1, device=dev)
            st.d_t_raw = torch.empty(b, L, device=dev) if want_grads else None
            st.d_i_raw = torch.empty(b, L, device=dev) if want_grads else None
            ops.clip_loss(tg, ig, P["temperature"], B=B, L=L, t_hat=st.t_hat, i_hat=st.i_hat, inv_norm=inv, sim=sim,
                          loss=st.loss, dtemperature=st.dtemp, d_t_raw=st.d_t_raw, d_i_raw=st.d_i_raw,
                          row0=self.dp_rank * b, nrows=b if want_grads else 0)
        return st

    def _backward_into(self, st, ectx, P, G, cls, gscale, text_backward=None):
        """Accumulate parameter gradients into G; returns d(cls). text_backward(dcls): optional callable that runs the text tower's
        backward as soon as d(cls) exists, i.e. BEFORE the image tower's backward."""
        vit: CTViT = self.visual_transformer
        g = vit.engine.g
        dev = cls.device
        b, L, K = st.b, st.L, st.K
        bt = st.cls32.shape[0]                          # reports in the text batch (== b for the contrastive step)
        d_t, d_i = st.d_t_raw, st.d_i_raw
        if gscale != 1.0:
            d_t, d_i = d_t * gscale, d_i * gscale      # scalar rescale of two [b, L] tensors (loss weighting)
        # one scalar. Under data parallelism every rank evaluates d loss / d temperature over the WHOLE global similarity matrix
        # (the tower gradients are row-partitioned, this one is not) and the arena is SUM-all-reduced: take 1/world of it here.
        G["temperature"].add_(st.dtemp.view(()), alpha=gscale / max(1, self.dp_world))
        # text projection: t_raw = cls Wt^T
        ops.sgemm(d_t, st.cls32, G["to_text_latent.weight"], M=L, N=self.dim_text, K=bt, trans_a=True, accumulate=True)
        dcls = torch.empty(bt, self.dim_text, device=dev)
        ops.sgemm(d_t, P["to_text_latent.weight"], dcls, M=bt, N=self.dim_text, K=L)
        # visual projection: i_raw = pooled Wv^T  (294912 -> 512: HBM-bound, weight streamed once per GEMM)
        d_i_bf = torch.empty(b, L, dtype=torch.bfloat16, device=dev)
        ops.cast_bf16(d_i, d_i_bf, b * L)
        ops.gemm(d_i_bf, st.pooled_bf16, M=L, N=K, K=b, a_major=1, b_major=1, epilogue=ops.EPI_ATOMIC_F32,
                 C_out=G["to_visual_latent.weight"], ldc=K)
        if self.dp_early_reduce is not None:   # 604 MB of the 1.14 GB gradient are final here: all-reduce them behind the towers' backward
            self.dp_early_reduce("to_visual_latent.weight")
        if text_backward is not None:          # the text tower first: its 0.44 GB of gradients then travel behind the long image-tower backward
            text_backward(dcls)
        dpooled = torch.empty(b, K, device=dev)
        ops.gemm(d_i_bf, self._visual_weight_bf16(P["to_visual_latent.weight"]), M=b, N=K, K=L, b_major=1,
                 epilogue=ops.EPI_F32, C_out=dpooled)
        # mean over t (ct_clip.py:724) + straight-through quantiser -> gradient at the temporal norm_out output
        dtok = torch.empty(ectx["M"], g.dim, device=dev)
        ops.pool_bwd(dpooled, dtok, B=b, T=ectx["T"], S=g.S, D=g.dim)
        PV = ectx["P"]
        GV = _PrefixView(G, "visual_transformer.")
        vit.engine.backward(ectx, dtok, PV, GV)
        return dcls.to(cls.dtype)

    # ------------------------------------------------------------------------------------------
    def forward(self, text, image, device=None, return_loss=False, return_encodings=False, return_latents=False,
                freeze_image_encoder=False, freeze_text_encoder=False, text_to_image=True, aug_text=None, aug_image=None):
        if aug_text is not None or aug_image is not None:
            raise NotImplementedError("multiview augmentation (ct_clip.py:651-675) is unused by the reference scripts")
        if not self._fast_path():
            return self._forward_foreign(text, image, return_loss, return_encodings, return_latents)
        names, tensors = self._live()
        if return_loss:
            if self._text_native() and not freeze_text_encoder:
                return _ClipStepFn.apply(self, torch.is_grad_enabled(), None, text, image, tuple(names), *tensors)
            _, cls = self._text_cls(text)
            if freeze_text_encoder:
                cls = cls.detach()
            return _ClipStepFn.apply(self, torch.is_grad_enabled(), cls, None, image, tuple(names), *tensors)
        if torch.is_grad_enabled() and self._text_native() and not return_encodings and \
                any(t.requires_grad for t in tensors):
            # fine-tuning on the similarity logits / latents (scripts/ct_vocabfine_train.py:108, ct_lipro_train.py:26-27): same
            # outputs as below, differentiable w.r.t. both towers through _ClipLatentsFn
            tl, il = self.latents_with_grad(text, image)
            if return_latents:
                return tl, il, None
            return (tl * il).sum(-1) * self.temperature.exp()                         # ct_clip.py:805-807
        enc_text, cls = self._text_cls(text)
        # inference / export paths: no gradient
        with torch.no_grad():
            P = dict(zip(names, tensors))
            vit: CTViT = self.visual_transformer
            vn, vt = vit.named_live_tensors()
            ectx = vit._run_forward(image, dict(zip(vn, vt)), save=False)
            g = vit.engine.g
            toks = None
            if return_latents:    # enc_image_send, ct_clip.py:721 (gathered before any code-book EMA)
                toks = torch.empty(ectx["M"], g.dim, device=cls.device)
                ops.vq_gather(ectx["idx"], ectx["P"]["vq._codebook.embed"], toks, ectx["M"], g.dim)
            st = self._heads_forward(cls, ectx, P, want_loss=False, want_grads=False)
            if return_encodings:  # ct_clip.py:746-747: (enc_text, mean-pooled + flattened image tokens)
                return enc_text, st.pooled_bf16.float()
            if return_latents:    # ct_clip.py:788-792
                return st.t_hat, st.i_hat, toks.view(ectx["b"], ectx["T"], g.H, g.W, g.dim)
            Bt, Bi = st.t_hat.shape[0], st.i_hat.shape[0]
            out = torch.empty(max(Bt, Bi), device=cls.device)
            ops.clip_sims(st.t_hat, Bt, st.i_hat, Bi, self.dim_latent, P["temperature"], out)   # ct_clip.py:805-807
            return out

    def latents_with_grad(self, text, image):
        """l2-normalised (text latents [n_text, L], image latents [b, L]), differentiable w.r.t. every live parameter."""
        names, tensors = self._live()
        t_raw, i_raw = _ClipLatentsFn.apply(self, torch.is_grad_enabled(), text, image, tuple(names), *tensors)
        return torch.nn.functional.normalize(t_raw, dim=-1), torch.nn.functional.normalize(i_raw, dim=-1)

    def _forward_foreign(self, text, image, return_loss, return_encodings, return_latents):
        raise NotImplementedError("CTCLIP here drives ct_clip_b200.CTViT as its image encoder (the encoder every reference "
                                  "script injects); arbitrary image encoders are outside the hot-path build")

    # ---- helpers for the zero-shot path (scripts/zero_shot.py): encode once, reuse
    @torch.no_grad()
    def encode_text_latents(self, text):
        _, cls = self._text_cls(text)
        t_raw = torch.empty(cls.shape[0], self.dim_latent, device=cls.device)
        ops.sgemm(cls.float().contiguous(), self.to_text_latent.weight, t_raw, M=cls.shape[0], N=self.dim_latent,
                  K=self.dim_text, trans_b=True)
        return self._l2(t_raw)

    def _l2(self, x):
        B, L = x.shape
        hat, dummy = torch.empty_like(x), torch.empty_like(x)
        inv, sim = torch.empty(2 * B, device=x.device), torch.empty(1, device=x.device)
        ops.clip_loss(x, x, self.temperature, B=B, L=L, t_hat=hat, i_hat=dummy, inv_norm=inv, sim=sim)
        return hat

    @torch.no_grad()
    def encode_image_latents(self, image):
        vit: CTViT = self.visual_transformer
        vn, vt = vit.named_live_tensors()
        ectx = vit._run_forward(image, dict(zip(vn, vt)), save=False)
        g = vit.engine.g
        b, K, L = ectx["b"], g.S * g.dim, self.dim_latent
        pooled = torch.empty(b, K, dtype=torch.bfloat16, device=image.device)
        ops.vq_gather_pool(ectx["idx"], ectx["P"]["vq._codebook.embed"], B=b, T=ectx["T"], S=g.S, D=g.dim, pooled_bf16=pooled)
        i_raw = torch.zeros(b, L, device=image.device)
        kb = (K + 63) // 64
        ops.gemm(pooled, self._visual_weight_bf16(self.to_visual_latent.weight), M=b, N=L, K=K,
                 epilogue=ops.EPI_ATOMIC_F32, C_out=i_raw, splits=max(1, min(kb, 148)))
        return self._l2(i_raw)


class _PrefixView:
    """G['x'] -> underlying['prefix' + 'x'] (so the image tower writes straight into the CLIP-level gradient dict)."""

    def __init__(self, d, prefix):
        self.d, self.prefix = d, prefix

    def __getitem__(self, k):
        return self.d[self.prefix + k]
