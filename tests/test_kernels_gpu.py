"""Unit parity of the individual sm_100a kernels against fp32 restatements (oracle functions where one
exists, plain torch math otherwise), on the same bf16-rounded inputs. Run on the B200 box (-m gpu)."""
import math

import pytest
import torch
import torch.nn.functional as F

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _randn(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(DEV)


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("D", [512, 768])
def test_layernorm_fwd_bwd(D):
    from ct_clip_b200 import ops
    M = 1000
    x = _randn(M, D, seed=1) * 2 + 0.5
    gamma, beta = 1 + 0.1 * _randn(D, seed=2), 0.1 * _randn(D, seed=3)
    xhat = torch.empty(M, D, dtype=torch.bfloat16, device=DEV)
    raw = torch.empty_like(xhat)
    y = torch.empty(M, D, device=DEV)
    rstd = torch.empty(M, device=DEV)
    ops.ln_fwd(x, M, D, gamma=gamma, beta=beta, xhat=xhat, raw=raw, y_f32=y, rstd=rstd)
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yref = F.layer_norm(xr, (D,), gr, br, 1e-5)
    assert rel_err(y, yref) < 1e-5
    assert rel_err(raw, x) < 1e-2 and rel_err(xhat, F.layer_norm(x, (D,))) < 1e-2
    # affine backward
    dy = _randn(M, D, seed=4)
    yref.backward(dy)
    dx = torch.empty(M, D, device=DEV)
    dg, db = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    xhat32 = F.layer_norm(x, (D,)).to(torch.bfloat16)
    ops.ln_bwd(M, D, g_f32=dy, gamma=gamma, xhat=xhat32, rstd=rstd, dx_f32=dx, dgamma=dg, dbeta=db)
    assert rel_err(dx, xr.grad) < 1e-2
    assert rel_err(dg, gr.grad) < 1e-2 and rel_err(db, br.grad) < 1e-3
    # folded (no gamma) backward with residual accumulate and bf16 upstream
    dxh = _randn(M, D, seed=5).to(torch.bfloat16)
    res = _randn(M, D, seed=6)
    xr2 = x.clone().requires_grad_(True)
    (F.layer_norm(xr2, (D,)) * dxh.float()).sum().backward()
    out = res.clone()
    outb = torch.empty(M, D, dtype=torch.bfloat16, device=DEV)
    ops.ln_bwd(M, D, g_bf16=dxh, xhat=xhat32, rstd=rstd, dres_in=out, dx_f32=out, dx_bf16=outb)
    assert rel_err(out, xr2.grad + res) < 1e-2
    assert rel_err(outb, xr2.grad + res) < 2e-2


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("temporal,T,H,W", [(False, 4, 4, 4), (True, 4, 4, 4), (False, 5, 3, 6), (True, 5, 3, 6), (True, 6, 4, 4),
                                            (False, 3, 10, 24), (True, 6, 10, 24), (False, 7, 12, 20), (True, 8, 8, 8),
                                            (False, 26, 24, 24), (True, 24, 24, 24)])
def test_peg_fwd_bwd(temporal, T, H, W):
    """fp32 PEG kernels against the fp32 oracle (1e-5): forward, data gradient (+ bf16 copy), weight and bias gradients."""
    from ct_clip_b200 import ops
    from oracle import ctclip_oracle as O
    # which kernels run: spatial grids with W <= 24 and temporal grids with T == H == W take the plane-streaming
    # kernels (csrc/peg_stream.cu; the 24^3 cases cut the CTA ranges mid-column: priming steps, several columns per CTA),
    # the other temporal grids the general kernels of csrc/peg.cu; test_peg_kernel_families_agree pins the two against each other
    b, D = 2, 512
    x = _randn(b, T, H, W, D, seed=7)            # canonical layout
    w = 0.2 * _randn(D, 1, 3, 3, 3, seed=8)
    bias = 0.1 * _randn(D, seed=9)
    y = torch.empty_like(x)
    kw = dict(B=b, T=T, H=H, W=W, D=D, temporal=temporal)
    if temporal and T == 6:     # also exercise the precomputed canon(f) table
        f = torch.arange(T * H * W)
        kw["canon_table"] = (((f % T) * H + f // (T * W)) * W + (f // T) % W).to(torch.int32).to(DEV)
    ops.peg_fwd(x.view(-1, D), y.view(-1, D), w.view(D, 27), bias, **kw)

    def conv(xc, wc, bc):    # conv term only, canonical layout in / out
        if temporal:   # reference memory order of the temporal stack: (b h w) t d, reshaped as (b,T,H,W)
            xr = xc.permute(0, 2, 3, 1, 4).reshape(b * H * W, T, D)
            return O.peg(xr, (b, T, H, W), wc, bc).reshape(b, H, W, T, D).permute(0, 3, 1, 2, 4)
        return O.peg(xc.reshape(b * T, H * W, D), (b, T, H, W), wc, bc).reshape(b, T, H, W, D)

    tol = 1e-5
    xcpu, wcpu, bcpu, = x.cpu(), w.cpu(), bias.cpu()
    xc, wc, bc = xcpu.clone().requires_grad_(True), wcpu.clone().requires_grad_(True), bcpu.clone().requires_grad_(True)
    cref = conv(xc, wc, bc)
    assert rel_err(y, cref.detach() + xcpu) < tol
    assert rel_err(y, (conv(xcpu, wcpu, bcpu) + xcpu)) < 1e-2
    dy = _randn(b, T, H, W, D, seed=10)
    cref.backward(dy.cpu())
    dx = torch.empty_like(x)
    dxb = torch.empty(x.shape, dtype=torch.bfloat16, device=DEV)
    ops.peg_bwd_data(dy.view(-1, D), dx.view(-1, D), w.view(D, 27), dx_bf16=dxb.view(-1, D), **kw)
    dx_ref = xc.grad + dy.cpu()          # dx = dy + conv^T(dy)
    assert rel_err(dx, dx_ref) < tol
    assert rel_err(dxb, dx_ref) < 1e-2
    dw, db = torch.zeros(D, 27, device=DEV), torch.zeros(D, device=DEV)
    ops.peg_bwd_weight(x.view(-1, D), dy.view(-1, D), dw, db, **kw)
    assert rel_err(dw, wc.grad.view(D, 27)) < 1e-4
    assert rel_err(db, bc.grad) < 1e-4


@pytest.mark.parametrize("temporal,T,H,W", [(False, 24, 24, 24), (True, 24, 24, 24), (False, 9, 13, 22), (True, 12, 12, 12)])
def test_peg_kernel_families_agree(temporal, T, H, W):
    """plane-streaming kernels (default) vs the general kernels on the same inputs: forward, data gradient (+ bf16 copy), weight
    and bias gradients; b = 3 so that CTA ranges start and end mid-column in both families."""
    from ct_clip_b200 import _lib, ops
    b, D = 3, 512
    x = _randn(b * T * H * W, D, seed=11)
    dy = _randn(b * T * H * W, D, seed=12)
    w = 0.2 * _randn(D, 27, seed=13)
    bias = 0.1 * _randn(D, seed=14)
    kw = dict(B=b, T=T, H=H, W=W, D=D, temporal=temporal)
    res = {}
    try:
        for variant in (0, 1):
            _lib.check(_lib.lib().ctclip_debug_set_peg_variant(variant), "variant")
            y, dx = torch.empty_like(x), torch.empty_like(x)
            dxb = torch.empty(x.shape, dtype=torch.bfloat16, device=DEV)
            dw, db = torch.zeros(D, 27, device=DEV), torch.zeros(D, device=DEV)
            ops.peg_fwd(x, y, w, bias, **kw)
            ops.peg_bwd_data(dy, dx, w, dx_bf16=dxb, **kw)
            ops.peg_bwd_weight(x, dy, dw, db, **kw)
            ops.peg_bwd_weight(x, dy, dw, db, **kw)            # gradients ACCUMULATE
            torch.cuda.synchronize()
            res[variant] = (y, dx, dxb.float(), dw, db)
    finally:
        _lib.check(_lib.lib().ctclip_debug_set_peg_variant(0), "restore default")
    for a, r, tol in zip(res[0], res[1], (1e-5, 1e-5, 1e-2, 1e-4, 1e-4)):
        assert rel_err(a, r) < tol, (rel_err(a, r), tol)


# ------------------------------------------------------------------------------------------------
def _attn_ref(q, k, v, bias, scale=8.0):
    # q,k,v: [seqs, heads, n, dh] fp32
    sim = q @ k.transpose(-1, -2) * scale
    if bias is not None:
        sim = sim + bias
    return sim.softmax(-1) @ v


def _to_frag(bias):
    """dense [heads, n, n] -> MMA-fragment order [heads, n_pad/16, ceil(n_pad/64), 32 lanes, 8 n-tiles, 4] (zero padded)."""
    heads, n, _ = bias.shape
    n_pad = (n + 15) // 16 * 16
    RT, CB = n_pad // 16, (n_pad + 63) // 64
    big = torch.zeros(heads, RT * 16, CB * 64, dtype=bias.dtype, device=bias.device)
    big[:, :n, :n] = bias
    lane = torch.arange(32, device=bias.device)
    e = torch.arange(4, device=bias.device)
    nt = torch.arange(8, device=bias.device)
    row_in = (lane[:, None, None] // 4) + 8 * (e[None, None, :] // 2)                       # [32,1,4]
    col_in = nt[None, :, None] * 8 + 2 * (lane[:, None, None] % 4) + (e[None, None, :] % 2)   # [32,8,4]
    rows = (torch.arange(RT, device=bias.device)[:, None, None, None, None] * 16 + row_in[None, None])    # [RT,1,32,1,4]
    cols = (torch.arange(CB, device=bias.device)[None, :, None, None, None] * 64 + col_in[None, None])   # [1,CB,32,8,4]
    rows = rows.expand(RT, CB, 32, 8, 4)
    cols = cols.expand(RT, CB, 32, 8, 4)
    return big[:, rows, cols].contiguous().view(-1)


@pytest.mark.parametrize("mode,b,T,S", [("spatial", 2, 3, 16), ("spatial", 1, 2, 576), ("spatial", 1, 2, 100),
                                        ("spatial_frag", 1, 2, 576), ("spatial_frag", 2, 2, 100),
                                        ("temporal", 2, 24, 16), ("temporal", 1, 5, 8)])
def test_attention_fwd_bwd(mode, b, T, S):
    use_frag = mode == "spatial_frag"
    mode = "spatial" if use_frag else mode
    from ct_clip_b200 import ops
    heads, dh = 8, 32
    I = heads * dh
    M = b * T * S
    q = F.normalize(_randn(M, heads, dh, seed=11), dim=-1).mul(1.1).to(torch.bfloat16).view(M, I)
    k = F.normalize(_randn(M, heads, dh, seed=12), dim=-1).mul(0.9).to(torch.bfloat16).view(M, I)
    kv = torch.empty(M, 2 * I, dtype=torch.bfloat16, device=DEV)
    kv[:, I:] = _randn(M, I, seed=13).to(torch.bfloat16)
    v = kv[:, I:]
    if mode == "spatial":
        n, nseq = S, b * T
        geom = dict(n=n, heads=heads, num_seqs=nseq, seq_inner=1, seq_outer_stride=S, tok_stride=1)
        bias32 = 0.5 * _randn(heads, n, n, seed=14)
        bias = bias32.to(torch.bfloat16)
        bias_t = bias.transpose(1, 2).contiguous()

        def to_seq(x):   # [M, I] -> [nseq, heads, n, dh]
            return x.float().view(nseq, n, heads, dh).permute(0, 2, 1, 3)

        def from_seq(x):
            return x.permute(0, 2, 1, 3).reshape(M, I)
    else:
        n, nseq = T, b * S
        geom = dict(n=n, heads=heads, num_seqs=nseq, seq_inner=S, seq_outer_stride=T * S, tok_stride=S)
        bias = bias_t = None

        def to_seq(x):   # canonical (b,t,s) rows -> [(b s), heads, t, dh]
            return x.float().view(b, T, S, heads, dh).permute(0, 2, 3, 1, 4).reshape(nseq, heads, T, dh)

        def from_seq(x):
            return x.reshape(b, S, heads, T, dh).permute(0, 3, 1, 2, 4).reshape(M, I)

    o = torch.empty(M, I, dtype=torch.bfloat16, device=DEV)
    lse = torch.empty(M, heads, device=DEV)
    if use_frag:    # fragment-ordered bias tables (what the encoder uses); the dense table still feeds the dbias kernel
        log2e = 1.4426950408889634   # the fragment tables hold bias * log2(e) (ctclip_cpb_expand_frag)
        geom["bias_frag"] = _to_frag((bias.float() * log2e).to(bias.dtype))
        geom["bias_t_frag"] = _to_frag((bias_t.float() * log2e).to(bias_t.dtype))
        bias_t = None
    ops.attn_fwd(q, k, v, o, lse, ldq=I, ldk=I, ldv=2 * I, ldo=I, bias=bias, **geom)
    qs, ks, vs = (to_seq(t).requires_grad_(True) for t in (q, k, v))
    bref = bias.float().clone().requires_grad_(True) if bias is not None else None
    oref = _attn_ref(qs, ks, vs, bref)
    assert rel_err(o, from_seq(oref)) < 1e-2
    # backward
    d_o = _randn(M, I, seed=15).to(torch.bfloat16)
    oref.backward(to_seq(d_o))
    dq = torch.empty(M, I, dtype=torch.bfloat16, device=DEV)
    dkv = torch.zeros(M, 2 * I, dtype=torch.bfloat16, device=DEV)
    delta = torch.empty(M, heads, device=DEV)
    dbias = torch.zeros(heads, n, n, device=DEV) if bias is not None else None
    ops.attn_bwd(q, k, v, o, lse, d_o, delta, dq, dkv, dkv[:, I:], ldq=I, ldk=I, ldv=2 * I, ldo=I, ld_dq=I, ld_dk=2 * I,
                 ld_dv=2 * I, total_rows=M, bias=bias, bias_t=bias_t, dbias=dbias, **geom)
    assert rel_err(dq, from_seq(qs.grad)) < 2e-2
    assert rel_err(dkv[:, :I], from_seq(ks.grad)) < 2e-2
    assert rel_err(dkv[:, I:], from_seq(vs.grad)) < 2e-2
    if bias is not None:
        assert rel_err(dbias, bref.grad) < 2e-2
        if n % 2 == 0:   # second route to dbias: d logits spilled by the dQ kernel + streaming reduction over the sequences
            num_items = geom["num_seqs"] * heads
            scratch = torch.full((num_items * n * n,), float("nan"), dtype=torch.bfloat16, device=DEV)
            dbias2 = torch.zeros(heads, n, n, device=DEV)
            dq2 = torch.empty_like(dq)
            ops.attn_bwd(q, k, v, o, lse, d_o, delta, dq2, dkv, dkv[:, I:], ldq=I, ldk=I, ldv=2 * I, ldo=I, ld_dq=I,
                         ld_dk=2 * I, ld_dv=2 * I, total_rows=M, bias=bias, bias_t=bias_t, dbias=dbias2, ds_scratch=scratch,
                         **geom)
            assert torch.equal(dq2, dq)
            assert rel_err(dbias2, bref.grad) < 2e-2


def test_l2norm_bwd_and_epilogue():
    from ct_clip_b200 import ops
    M, heads, dh, D = 512, 8, 32, 512
    I = heads * dh
    x = _randn(M, D, seed=16).to(torch.bfloat16)
    w = (_randn(2 * I, D, seed=17) / math.sqrt(D)).to(torch.bfloat16)
    scale = 1 + 0.1 * _randn(dh, seed=18)
    raw = torch.empty(M, 2 * I, dtype=torch.bfloat16, device=DEV)
    kh = torch.empty(M, I, dtype=torch.bfloat16, device=DEV)
    ops.gemm(x, w, M=M, N=2 * I, K=D, epilogue=ops.EPI_L2NORM, C_out=raw, C2=kh, norm_cols=I, norm_scale=scale)
    ref = x.float() @ w.float().t()
    assert rel_err(raw, ref) < 1e-2
    kref = F.normalize(ref[:, :I].view(M, heads, dh), dim=-1) * scale
    assert rel_err(kh, kref.view(M, I)) < 1e-2
    # backward of normalise*scale given the raw projection
    xr = raw[:, :I].float().view(M, heads, dh).clone().requires_grad_(True)
    sr = scale.clone().requires_grad_(True)
    g = _randn(M, I, seed=19).to(torch.bfloat16)
    (F.normalize(xr, dim=-1) * sr * g.float().view(M, heads, dh)).sum().backward()
    dx = torch.empty(M, I, dtype=torch.bfloat16, device=DEV)
    ds = torch.zeros(dh, device=DEV)
    ops.l2norm_bwd(g, I, raw, 2 * I, scale, dx, I, ds, M, heads)
    assert rel_err(dx, xr.grad.view(M, I)) < 2e-2
    assert rel_err(ds, sr.grad) < 1e-2


def test_geglu_bwd_and_colsum():
    from ct_clip_b200 import ops
    M, Fp = 700, 1408
    h = _randn(M, 2 * Fp, seed=20).to(torch.bfloat16)
    dg = _randn(M, Fp, seed=21).to(torch.bfloat16)
    hr = h.float().clone().requires_grad_(True)
    (F.gelu(hr[:, 1::2]) * hr[:, 0::2] * dg.float()).sum().backward()
    cs = torch.zeros(2 * Fp, device=DEV)
    hh = h.clone()
    ops.geglu_bwd(dg, hh, M=M, n_pairs=Fp, colsum_out=cs)
    assert rel_err(hh, hr.grad) < 1e-2
    assert rel_err(cs, hr.grad.sum(0)) < 1e-2
    cs2 = torch.zeros(2 * Fp, device=DEV)
    ops.colsum(hh, cs2, M=M, N=2 * Fp)
    assert rel_err(cs2, hh.float().sum(0)) < 1e-4


def test_sgemm_small():
    from ct_clip_b200 import ops
    A, B = _randn(100, 70, seed=22), _randn(45, 70, seed=23)
    Cm = torch.empty(100, 45, device=DEV)
    bias = _randn(45, seed=24)
    ops.sgemm(A, B, Cm, M=100, N=45, K=70, trans_b=True, bias=bias, act=1)
    assert rel_err(Cm, F.leaky_relu(A @ B.t() + bias, 0.1)) < 1e-5
    C2 = torch.ones(70, 45, device=DEV)
    Bn = _randn(100, 45, seed=25)
    ops.sgemm(A, Bn, C2, M=70, N=45, K=100, trans_a=True, accumulate=True)
    assert rel_err(C2, A.t() @ Bn + 1) < 1e-5


def test_loss_kernel_matches_oracle():
    from ct_clip_b200 import ops
    from oracle import ctclip_oracle as O
    B, L = 16, 512
    t_raw, i_raw = _randn(B, L, seed=26), _randn(B, L, seed=27)
    temp = torch.tensor([1.3], device=DEV)
    t_hat, i_hat = torch.empty(B, L, device=DEV), torch.empty(B, L, device=DEV)
    inv, sim = torch.empty(2 * B, device=DEV), torch.empty(B, B, device=DEV)
    loss, dtemp = torch.empty(1, device=DEV), torch.empty(1, device=DEV)
    dt, di = torch.empty(B, L, device=DEV), torch.empty(B, L, device=DEV)
    ops.clip_loss(t_raw, i_raw, temp, B=B, L=L, t_hat=t_hat, i_hat=i_hat, inv_norm=inv, sim=sim, loss=loss,
                  dtemperature=dtemp, d_t_raw=dt, d_i_raw=di, row0=0, nrows=B)
    tr, ir = t_raw.cpu().requires_grad_(True), i_raw.cpu().requires_grad_(True)
    tp = temp.cpu()[0].clone().requires_grad_(True)
    lref = O.clip_loss(F.normalize(tr, dim=-1), F.normalize(ir, dim=-1), tp)
    lref.backward()
    assert abs(loss.item() - lref.item()) < 1e-5 * max(1, abs(lref.item()))
    assert rel_err(dt, tr.grad) < 1e-4 and rel_err(di, ir.grad) < 1e-4
    assert abs(dtemp.item() - tp.grad.item()) < 1e-4 * max(1e-3, abs(tp.grad.item()))


def test_adam_and_clip():
    from ct_clip_b200 import ops
    n = 100003
    p = _randn(n, seed=28)
    g = _randn(n, seed=29) * 0.01
    pr = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([pr], lr=1e-3, betas=(0.9, 0.99), eps=1e-8)
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    pw = p.clone()
    for step in range(1, 4):
        gi = g * step
        pr.grad = gi.clone()
        torch.nn.utils.clip_grad_norm_([pr], 0.5)
        opt.step()
        ss = torch.zeros(1, device=DEV)
        ops.grad_sumsq(gi, n, ss)
        ops.adam_step(pw, gi, m, v, n, lr=1e-3, step=step, max_norm=0.5, sumsq=ss)
    assert rel_err(pw, pr.data) < 1e-5


def test_adamw_decoupled_decay_on_leading_range():
    """AdamW of the reference's get_optimizer (optimizer.py:26-34): weight decay on the ndim >= 2 tensors only, which the arena
    lays out first; against torch.optim.AdamW with the same two parameter groups."""
    from ct_clip_b200 import ops
    n_w, n_b = 4096 * 8, 1000                       # a [4096, 8] weight (decayed) followed by a bias vector (not decayed)
    w0, b0 = _randn(n_w, seed=40), _randn(n_b, seed=41)
    gw, gb = _randn(n_w, seed=42) * 0.01, _randn(n_b, seed=43) * 0.01
    wr, br = torch.nn.Parameter(w0.clone().view(4096, 8)), torch.nn.Parameter(b0.clone())
    opt = torch.optim.AdamW([{"params": [wr]}, {"params": [br], "weight_decay": 0}], lr=1e-2, weight_decay=0.1, betas=(0.9, 0.99), eps=1e-8)
    n = n_w + n_b
    p = torch.cat([w0, b0]).clone()
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for step in range(1, 4):
        wr.grad, br.grad = (gw * step).view(4096, 8).clone(), (gb * step).clone()
        opt.step()
        g = torch.cat([gw, gb]) * step
        ops.adam_step(p, g, m, v, n, lr=1e-2, step=step, weight_decay=0.1, n_decay=n_w)
    assert rel_err(p[:n_w], wr.data.view(-1)) < 1e-5 and rel_err(p[n_w:], br.data) < 1e-5
