"""Torch-tensor front ends of the C-ABI kernels (raw device pointers + current stream).

Every function launches on torch's current CUDA stream and never synchronises. PyTorch is used
for device memory and streams only; all arithmetic happens inside libctclip_b200.so.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import (EPI_ARGMAX, EPI_ATOMIC_F32, EPI_BF16, EPI_BIAS_GELU, EPI_F32, EPI_GEGLU, EPI_GEGLU_BWD, EPI_L2NORM,  # noqa: F401
                   EPI_RESID_F32, AttnArgs, GemmArgs, LnBwdArgs, LnFwdArgs, LossArgs, PatchifyArgs, PegArgs,
                   SgemmArgs, call)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


# ------------------------------------------------------------------------------------------------
def gemm(A, B, *, M, N, K, a_major=0, b_major=0, epilogue=EPI_BF16, C_out=None, bias=None, resid=None, C2=None,
         arg_out=None, argval_out=None, splits=1, lda=None, ldb=None, ldc=None, ldc2=None, norm_cols=0,
         norm_scale=None, colsum=None, arg2_out=None):
    """C[M,N] = sum_k A(m,k) B(n,k); see include/ctclip_b200.h for the epilogues."""
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16 and A.is_cuda and B.is_cuda
    a = GemmArgs()
    a.M, a.N, a.K = M, N, K
    a.a_major, a.b_major = a_major, b_major
    a.A, a.lda = A.data_ptr(), (lda if lda is not None else A.stride(0))
    a.B, a.ldb = B.data_ptr(), (ldb if ldb is not None else B.stride(0))
    a.epilogue, a.splits = epilogue, splits
    a.C = _ptr(C_out)
    a.ldc = ldc if ldc is not None else (C_out.stride(0) if C_out is not None else 0)
    a.bias = _ptr(bias)
    a.resid = _ptr(resid)
    a.ldr = resid.stride(0) if resid is not None else 0
    a.C2 = _ptr(C2)
    a.ldc2 = ldc2 if ldc2 is not None else (C2.stride(0) if C2 is not None else 0)
    a.arg_out = _ptr(arg_out)
    a.argval_out = _ptr(argval_out)
    a.arg2_out = _ptr(arg2_out)
    a.norm_cols = norm_cols
    a.norm_scale = _ptr(norm_scale)
    a.colsum = _ptr(colsum)
    call("ctclip_gemm_bf16", C.byref(a), _stream(), tag=f"{M}x{N}x{K} a{a_major}b{b_major} epi{epilogue} s{splits}",
         work=("FB", 2.0 * M * N * K, _gemm_bytes(M, N, K, epilogue, norm_cols, C_out is not None)))


def _gemm_bytes(M, N, K, epilogue, norm_cols, has_c):
    """Algorithmic HBM bytes of one GEMM launch: both bf16 operands once + what the epilogue reads / writes."""
    ops_b = 2.0 * K * (M + N)
    mn = float(M) * N
    out = {EPI_BF16: 2 * mn, EPI_F32: 4 * mn, EPI_RESID_F32: 8 * mn, EPI_GEGLU: (2 * mn if has_c else 0) + mn,
           EPI_ATOMIC_F32: 4 * mn, EPI_ARGMAX: 4.0 * M, EPI_L2NORM: (2 * mn if has_c else 0) + 2.0 * M * norm_cols,
           EPI_BIAS_GELU: 4 * mn, EPI_GEGLU_BWD: 8 * mn}.get(epilogue, 2 * mn)
    return ops_b + out


def wgrad_splits(k_red: int, out_tiles: int) -> int:
    """Split-K factor for a weight-gradient GEMM: 0 = let the library choose (it minimises waves x k-blocks per unit for
    the tile shape it actually launches; a fixed '4 waves' guess left the 2816x512 gradient at 2.08 waves = 3 wave times)."""
    return 0


def ln_fwd(x, M, D, *, eps=1e-5, gamma=None, beta=None, xhat=None, raw=None, y_f32=None, y_bf16=None, rstd=None):
    a = LnFwdArgs()
    a.x, a.M, a.D, a.eps = x.data_ptr(), M, D, eps
    a.gamma, a.beta = _ptr(gamma), _ptr(beta)
    a.xhat_bf16, a.raw_bf16, a.y_f32, a.y_bf16, a.rstd_out = _ptr(xhat), _ptr(raw), _ptr(y_f32), _ptr(y_bf16), _ptr(rstd)
    per_elem = 4 + 2 * (xhat is not None) + 2 * (raw is not None) + 4 * (y_f32 is not None) + 2 * (y_bf16 is not None)
    call("ctclip_ln_fwd", C.byref(a), _stream(), tag=f"D{D} {per_elem}B/elem", work=("B", float(M) * D * per_elem))


def ln_bwd(M, D, *, g_f32=None, g_bf16=None, gamma=None, xhat, rstd, dres_in=None, add_bf16=None, dx_f32=None,
           dx_bf16=None, dgamma=None, dbeta=None):
    a = LnBwdArgs()
    a.M, a.D = M, D
    a.g_f32, a.g_bf16, a.gamma = _ptr(g_f32), _ptr(g_bf16), _ptr(gamma)
    a.xhat, a.rstd = xhat.data_ptr(), rstd.data_ptr()
    a.dres_in, a.add_bf16 = _ptr(dres_in), _ptr(add_bf16)
    a.dx_f32, a.dx_bf16, a.dgamma, a.dbeta = _ptr(dx_f32), _ptr(dx_bf16), _ptr(dgamma), _ptr(dbeta)
    per_elem = (2 + 4 * (g_f32 is not None) + 2 * (g_bf16 is not None) + 4 * (dres_in is not None) + 2 * (add_bf16 is not None)
                + 4 * (dx_f32 is not None) + 2 * (dx_bf16 is not None))
    call("ctclip_ln_bwd", C.byref(a), _stream(), tag=f"D{D} {per_elem}B/elem", work=("B", float(M) * D * per_elem))


def patchify(video, xhat, *, B, Cc, F, H, W, pt, p1, p2, eps=1e-5, ld_out=None):
    a = PatchifyArgs()
    a.video = video.data_ptr()
    if video.dtype == torch.int16:
        a.dtype, a.scale = 1, 1.0 / 1000.0
    elif video.dtype == torch.float32:
        a.dtype, a.scale = 0, 1.0
    else:
        raise TypeError(f"patchify: unsupported volume dtype {video.dtype} (need float32 or int16 HU)")
    a.B, a.C, a.F, a.H, a.W = B, Cc, F, H, W
    a.pt, a.p1, a.p2, a.eps = pt, p1, p2, eps
    a.xhat = xhat.data_ptr()
    a.ld_out = ld_out if ld_out is not None else xhat.stride(0)
    call("ctclip_patchify", C.byref(a), _stream(), tag=str(video.dtype).replace("torch.", ""),
         work=("B", float(video.numel()) * (video.element_size() + 2)))


def _peg_args(x, *, B, T, H, W, D, temporal, weight=None, bias=None, y=None, y_bf16=None, dy=None, dweight=None,
              dbias=None, lines=4, canon_table=None):
    a = PegArgs()
    a.x, a.dy, a.y, a.y_bf16 = x.data_ptr(), _ptr(dy), _ptr(y), _ptr(y_bf16)
    a.weight, a.bias, a.dweight, a.dbias = _ptr(weight), _ptr(bias), _ptr(dweight), _ptr(dbias)
    a.B, a.T, a.H, a.W, a.D, a.temporal, a.lines = B, T, H, W, D, int(temporal), lines
    a.canon_table = _ptr(canon_table) if temporal else None
    return a


def _peg_tag(kw):
    return "temporal" if kw.get("temporal") else "spatial"


def peg_fwd(x, y, weight, bias, **kw):
    call("ctclip_peg_fwd", C.byref(_peg_args(x, y=y, weight=weight, bias=bias, **kw)), _stream(), tag=_peg_tag(kw),
         work=("B", 8.0 * x.numel()))


def peg_bwd_data(dy, dx, weight, dx_bf16=None, **kw):
    call("ctclip_peg_bwd_data", C.byref(_peg_args(dy, y=dx, y_bf16=dx_bf16, weight=weight, **kw)), _stream(),
         tag=_peg_tag(kw), work=("B", (8.0 + 2 * (dx_bf16 is not None)) * dy.numel()))


def peg_bwd_weight(x, dy, dweight, dbias, **kw):
    call("ctclip_peg_bwd_weight", C.byref(_peg_args(x, dy=dy, dweight=dweight, dbias=dbias, **kw)), _stream(),
         tag=_peg_tag(kw), work=("B", 8.0 * x.numel()))


def _attn_args(q, k, v, o, lse, *, ldq, ldk, ldv, ldo, n, heads, num_seqs, seq_inner, seq_outer_stride, tok_stride,
               bias=None, bias_t=None, scale=8.0, dim_head=32, key_mask=None, bias_frag=None, bias_t_frag=None,
               cpb_table=None, grid_hw=None, qk_bound=None, dropout_p=0.0, dropout_seed=0, dropout_offset=0):
    """cpb_table (fp32 [(2h-1)(2w-1), heads]) + grid_hw=(h, w) select the tcgen05 / TMEM kernels (csrc/attention_tc.cu)."""
    a = AttnArgs()
    a.dropout_p, a.dropout_seed, a.dropout_offset = dropout_p, dropout_seed, dropout_offset
    if cpb_table is not None:
        a.cpb_table, a.qk_bound = cpb_table.data_ptr(), _ptr(qk_bound)
        a.grid_h, a.grid_w = grid_hw
    a.q, a.ldq, a.k, a.ldk, a.v, a.ldv = q.data_ptr(), ldq, k.data_ptr(), ldk, v.data_ptr(), ldv
    a.o, a.ldo, a.lse = o.data_ptr(), ldo, _ptr(lse)
    a.bias, a.bias_t = _ptr(bias), _ptr(bias_t)
    a.n, a.heads, a.dim_head, a.num_seqs, a.seq_inner = n, heads, dim_head, num_seqs, seq_inner
    a.seq_outer_stride, a.tok_stride, a.scale = seq_outer_stride, tok_stride, scale
    a.key_mask = _ptr(key_mask)
    a.bias_frag, a.bias_t_frag = _ptr(bias_frag), _ptr(bias_t_frag)
    return a


def _attn_flops(kw):
    """QK^T + PV of every (sequence, head): 4 n^2 dh"""
    return 4.0 * kw["num_seqs"] * kw["heads"] * kw["n"] * kw["n"] * kw.get("dim_head", 32)


def attn_fwd(q, k, v, o, lse, **kw):
    call("ctclip_attn_fwd", C.byref(_attn_args(q, k, v, o, lse, **kw)), _stream(),
         tag=f"n{kw['n']} dh{kw.get('dim_head', 32)}" + (" tc" if kw.get("cpb_table") is not None else ""), work=("F", _attn_flops(kw)))


def attn_tc_supported(n, grid_h, grid_w, dim_head=32) -> int:
    """bit 0: the tcgen05 forward kernel takes this geometry, bit 1: the tcgen05 backward kernel does."""
    return int(_lib.lib().ctclip_attn_tc_supported(n, grid_h, grid_w, dim_head))


def qk_bound(q_scale, k_scale, out, dim_head=32):
    """out[0] = max_d |q_scale_d * k_scale_d| (the logit bound of the fixed-reference softmax of the tcgen05 kernels)."""
    call("ctclip_qk_bound", q_scale.data_ptr(), k_scale.data_ptr(), dim_head, out.data_ptr(), _stream())


def attn_bwd(q, k, v, o, lse, d_o, delta, dq, dk, dv, *, ld_dq, ld_dk, ld_dv, total_rows, dbias=None, ds_scratch=None,
             dcpb_table=None, **kw):
    a = _attn_args(q, k, v, o, lse, **kw)
    a.ds_scratch = _ptr(ds_scratch)
    a.dcpb_table = _ptr(dcpb_table)
    a.d_o, a.delta = d_o.data_ptr(), delta.data_ptr()
    a.dq, a.ld_dq, a.dk, a.ld_dk, a.dv, a.ld_dv = dq.data_ptr(), ld_dq, dk.data_ptr(), ld_dk, dv.data_ptr(), ld_dv
    a.dbias = _ptr(dbias)
    a.total_rows = total_rows
    # algorithmic backward = 5 contractions (S, dP, dV, dQ, dK) = 2.5x the forward (the three kernels recompute S/dP)
    call("ctclip_attn_bwd", C.byref(a), _stream(),
         tag=f"n{kw['n']} dh{kw.get('dim_head', 32)}" + (" tc" if kw.get("cpb_table") is not None else "")
         + (" +dbias" if (dbias is not None or dcpb_table is not None) else ""), work=("F", 2.5 * _attn_flops(kw)))


def l2norm_bwd(dxh, ld_dxh, xraw, ld_x, scale, dx, ld_dx, dscale, rows, heads, dim_head=32):
    call("ctclip_l2norm_bwd", dxh.data_ptr(), ld_dxh, xraw.data_ptr(), ld_x, scale.data_ptr(), dx.data_ptr(), ld_dx,
         dscale.data_ptr(), rows, heads, dim_head, _stream(), work=("B", float(rows) * heads * dim_head * 6))


def sgemm(A, B, Cm, *, M, N, K, trans_a=False, trans_b=False, lda=None, ldb=None, ldc=None, bias=None, act=0,
          mask_ref=None, accumulate=False):
    a = SgemmArgs()
    a.M, a.N, a.K = M, N, K
    a.A, a.lda, a.trans_a = A.data_ptr(), (lda if lda is not None else A.stride(0)), int(trans_a)
    a.B, a.ldb, a.trans_b = B.data_ptr(), (ldb if ldb is not None else B.stride(0)), int(trans_b)
    a.C, a.ldc = Cm.data_ptr(), (ldc if ldc is not None else Cm.stride(0))
    a.bias, a.act = _ptr(bias), act
    a.mask_ref, a.ld_mask = _ptr(mask_ref), (mask_ref.stride(0) if mask_ref is not None else 0)
    a.accumulate = int(accumulate)
    call("ctclip_sgemm_f32", C.byref(a), _stream())


def colsum(x, out, *, M, N, ld=None):
    call("ctclip_colsum", x.data_ptr(), int(x.dtype == torch.bfloat16), ld if ld is not None else x.stride(0), M, N,
         out.data_ptr(), _stream())


def cast_bf16(x, y, n):
    call("ctclip_cast_f32_bf16", x.data_ptr(), y.data_ptr(), n, _stream())


def cpb_inputs(X, h, w):
    call("ctclip_cpb_inputs", X.data_ptr(), h, w, _stream())


def cpb_expand(table, heads, h, w, bias, bias_t=None):
    call("ctclip_cpb_expand", table.data_ptr(), heads, h, w, bias.data_ptr(), _ptr(bias_t), _stream())


def cpb_expand_frag(table, heads, h, w, bias_frag, bias_t_frag):
    call("ctclip_cpb_expand_frag", table.data_ptr(), heads, h, w, bias_frag.data_ptr(), bias_t_frag.data_ptr(), _stream())


def frag_elems(heads, n):
    """number of bf16 elements of a fragment-ordered bias table"""
    n_pad = (n + 15) // 16 * 16
    return heads * (n_pad // 16) * ((n_pad + 63) // 64) * 32 * 32


def cpb_reduce(dbias, heads, h, w, dtable):
    call("ctclip_cpb_reduce", dbias.data_ptr(), heads, h, w, dtable.data_ptr(), _stream())


def cpb_reduce_t(dbias_t, heads, h, w, dtable):
    """transposed fp32 table gradient [heads, n(j), n(i)] (tcgen05 backward, red.add path) -> dtable += (mirrored offsets)"""
    call("ctclip_cpb_reduce_t", dbias_t.data_ptr(), heads, h, w, dtable.data_ptr(), _stream())


def geglu_bwd(dg, h, *, M, n_pairs, colsum_out=None, ld_dg=None, ld_h=None):
    call("ctclip_geglu_bwd", dg.data_ptr(), ld_dg if ld_dg is not None else dg.stride(0), h.data_ptr(),
         ld_h if ld_h is not None else h.stride(0), M, n_pairs, _ptr(colsum_out), _stream(),
         work=("B", float(M) * n_pairs * 10))


def l2norm_rows_bf16(x, y, rows, D):
    call("ctclip_l2norm_rows_bf16", x.data_ptr(), y.data_ptr(), rows, D, _stream())


def dropout(x, *, n, p, seed, offset, resid=None, y_f32=None, y_bf16=None):
    """y = resid + keep * x / (1 - p) with the Philox mask of (seed, offset) (csrc/rng.cuh); the backward of a site is the same
    call on the upstream gradient with resid=None. In-place (y_f32 is x) is fine."""
    call("ctclip_dropout", x.data_ptr(), _ptr(resid), _ptr(y_f32), _ptr(y_bf16), n, p, seed, offset, _stream())


def vq_rerank(x, embed, idx, idx2, M, D):
    """fp32 re-ranking of the bf16 argmax GEMM's top-2 code candidates (in place on idx)."""
    call("ctclip_vq_rerank", x.data_ptr(), embed.data_ptr(), idx.data_ptr(), idx2.data_ptr(), M, D, _stream(),
         work=("B", float(M) * D * 4))


def vq_gather(idx, embed, out, M, D):
    call("ctclip_vq_gather", idx.data_ptr(), embed.data_ptr(), out.data_ptr(), M, D, _stream())


def vq_gather_pool(idx, embed, *, B, T, S, D, pooled_f32=None, pooled_bf16=None):
    call("ctclip_vq_gather_pool", idx.data_ptr(), embed.data_ptr(), B, T, S, D, _ptr(pooled_f32), _ptr(pooled_bf16),
         _stream())


def pool_bwd(dpooled, dtok, *, B, T, S, D):
    call("ctclip_pool_bwd", dpooled.data_ptr(), B, T, S, D, dtok.data_ptr(), _stream())


def vq_ema_accum(x, idx, bins, embed_sum, M, D):
    call("ctclip_vq_ema_accum", x.data_ptr(), idx.data_ptr(), M, D, bins.data_ptr(), embed_sum.data_ptr(), _stream())


def vq_ema_update(embed, cluster_size, bins, embed_sum, Cn, D, decay=0.8):
    call("ctclip_vq_ema_update", embed.data_ptr(), cluster_size.data_ptr(), bins.data_ptr(), embed_sum.data_ptr(), Cn, D,
         decay, _stream())


def prep_weight(W, out, *, K, Np, Kp, gamma=None, rowmap=None, ldw=None):
    call("ctclip_prep_weight", W.data_ptr(), ldw if ldw is not None else W.stride(0), K, _ptr(gamma), _ptr(rowmap), Np,
         Kp, out.data_ptr(), _stream())


class PrepBatch:
    """Collects prep_weight / prep_bias work items and runs them in ONE launch (ctclip_prep_batched). The descriptor table
    lives on the device and is rebuilt only when a tensor address changes (parameters in the trainer's arena never move)."""

    def __init__(self):
        self.items, self._key, self._table, self._keep = [], None, None, None

    def weight(self, W, out, *, K, Np, Kp, gamma=None, rowmap=None, ldw=None):
        self.items.append((0, W, ldw if ldw is not None else W.stride(0), gamma, None, None, rowmap, out, K, Np, Kp))

    def bias(self, W, out, *, K, Np, beta=None, bias_in=None, rowmap=None, ldw=None):
        self.items.append((1, W, ldw if ldw is not None else W.stride(0), None, beta, bias_in, rowmap, out, K, Np, 0))

    def run(self):
        import numpy as np
        items, self.items = self.items, []
        if not items:
            return
        key = tuple((it[0], it[1].data_ptr(), it[2], _ptr(it[3]) or 0, _ptr(it[4]) or 0, _ptr(it[5]) or 0, _ptr(it[6]) or 0,
                     it[7].data_ptr(), it[8], it[9], it[10]) for it in items)
        if key != self._key:
            arr = (_lib.PrepDesc * len(items))()
            for d, it in zip(arr, items):
                kind, W, ldw, gamma, beta, bias_in, rowmap, out, K, Np, Kp = it
                d.W, d.ldw, d.gamma, d.beta, d.bias_in = W.data_ptr(), ldw, _ptr(gamma), _ptr(beta), _ptr(bias_in)
                d.rowmap, d.out, d.K, d.Np, d.Kp, d.kind = _ptr(rowmap), out.data_ptr(), K, Np, Kp, kind
            host = torch.from_numpy(np.frombuffer(bytes(arr), dtype=np.uint8).copy())
            self._table = host.to(items[0][7].device)
            self._key = key
        self._keep = items      # keep the tensors alive until the next run
        call("ctclip_prep_batched", self._table.data_ptr(), len(items), 64, _stream(), tag=f"{len(items)} operands")


def prep_bias(W, out, *, K, Np, beta=None, bias_in=None, rowmap=None, ldw=None):
    call("ctclip_prep_bias", W.data_ptr(), ldw if ldw is not None else W.stride(0), K, _ptr(beta), _ptr(bias_in),
         _ptr(rowmap), Np, out.data_ptr(), _stream())


def unprep_wgrad(G, W, dW, *, K, Np, ldg=None, ldw=None, gamma=None, rowmap=None, s=None, dgamma=None, dbeta=None,
                 dbias=None):
    call("ctclip_unprep_wgrad", G.data_ptr(), ldg if ldg is not None else G.stride(0), W.data_ptr(),
         ldw if ldw is not None else W.stride(0), K, _ptr(gamma), _ptr(rowmap), Np, _ptr(s), dW.data_ptr(),
         _ptr(dgamma), _ptr(dbeta), _ptr(dbias), _stream())


def clip_loss(t_raw, i_raw, temperature, *, B, L, t_hat, i_hat, inv_norm, sim, loss=None, dtemperature=None,
              d_t_raw=None, d_i_raw=None, row0=0, nrows=0, loss_scale=1.0):
    a = LossArgs()
    a.t_raw, a.i_raw, a.B, a.L = t_raw.data_ptr(), i_raw.data_ptr(), B, L
    a.temperature = temperature.data_ptr()
    a.t_hat, a.i_hat, a.inv_norm, a.sim = t_hat.data_ptr(), i_hat.data_ptr(), inv_norm.data_ptr(), sim.data_ptr()
    a.loss, a.dtemperature = _ptr(loss), _ptr(dtemperature)
    a.d_t_raw, a.d_i_raw = _ptr(d_t_raw), _ptr(d_i_raw)
    a.row0, a.nrows, a.loss_scale = row0, nrows, loss_scale
    call("ctclip_clip_loss", C.byref(a), _stream())


def latent_exchange(t_raw, i_raw, *, b, L, rank, world, peer_bufs, peer_flags, step):
    """one kernel: push this rank's raw latents into every peer's gather buffer over NVLink, release / acquire the step flags"""
    call("ctclip_latent_exchange", t_raw.data_ptr(), i_raw.data_ptr(), b, L, rank, world, peer_bufs.data_ptr(),
         peer_flags.data_ptr(), step, _stream())


def clip_sims(t_hat, Bt, i_hat, Bi, L, temperature, out):
    call("ctclip_clip_sims", t_hat.data_ptr(), Bt, i_hat.data_ptr(), Bi, L, temperature.data_ptr(), out.data_ptr(),
         _stream())


def grad_sumsq(g, n, out):
    call("ctclip_grad_sumsq", g.data_ptr(), n, out.data_ptr(), _stream(), work=("B", 4.0 * n))


def adam_step(p, g, m, v, n, *, lr, beta1=0.9, beta2=0.99, eps=1e-8, step, max_norm=0.0, sumsq=None, grad_scale=1.0,
              weight_decay=0.0, n_decay=0):
    """weight_decay > 0: AdamW with decoupled decay on the first n_decay elements (the ndim >= 2 tensors, laid out first)."""
    call("ctclip_adam_step", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, lr, beta1, beta2, eps, step,
         max_norm, _ptr(sumsq), grad_scale, weight_decay, n_decay, _stream(), work=("B", 28.0 * n))


def bert_embed(ids, word, pos, type0, out, rows, n, H):
    call("ctclip_bert_embed", ids.data_ptr(), word.data_ptr(), pos.data_ptr(), type0.data_ptr(), out.data_ptr(), rows, n, H,
         _stream())


def bert_embed_bwd(ids, g, dword, dpos, rows, n, H):
    call("ctclip_bert_embed_bwd", ids.data_ptr(), g.data_ptr(), dword.data_ptr(), dpos.data_ptr(), rows, n, H, _stream())


def gelu_bwd(dy, pre, *, M, N, colsum_out=None):
    call("ctclip_gelu_bwd", dy.data_ptr(), dy.stride(0), pre.data_ptr(), pre.stride(0), M, N, _ptr(colsum_out), _stream())


def zero_shot_probs(img, txt, temperature, probs):
    call("ctclip_zero_shot_probs", img.data_ptr(), txt.data_ptr(), img.shape[0], txt.shape[0], img.shape[1],
         temperature.data_ptr(), probs.data_ptr(), _stream())
