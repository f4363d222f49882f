"""tcgen05 / TMEM spatial attention (csrc/attention_tc.cu) against fp32 softmax attention with the continuous position
bias expanded from its table (attention.py:152-178, :245-282), forward and backward, incl. the table gradient."""
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import rel_err, rms_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel_index(H, W):
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    y, x = ys.reshape(-1), xs.reshape(-1)
    return (y[:, None] - y[None, :] + H - 1) * (2 * W - 1) + (x[:, None] - x[None, :] + W - 1)      # [n, n], rel(i, j)


def _inputs(nseq, H, W, heads=8, dh=32, seed=0):
    n, I = H * W, heads * dh
    M = nseq * n
    g = torch.Generator().manual_seed(seed)
    qs = (1.0 + 0.2 * torch.rand(dh, generator=g)).to(DEV)
    ks = (1.0 + 0.2 * torch.rand(dh, generator=g)).to(DEV)
    q = (F.normalize(torch.randn(M, heads, dh, generator=g).to(DEV), dim=-1) * qs).to(torch.bfloat16).view(M, I)
    k = (F.normalize(torch.randn(M, heads, dh, generator=g).to(DEV), dim=-1) * ks).to(torch.bfloat16).view(M, I)
    kv = torch.empty(M, 2 * I, dtype=torch.bfloat16, device=DEV)
    kv[:, I:] = torch.randn(M, I, generator=g).to(torch.bfloat16).to(DEV)
    tab = (0.7 * torch.randn((2 * H - 1) * (2 * W - 1), heads, generator=g)).to(DEV)
    d_o = torch.randn(M, I, generator=g).to(torch.bfloat16).to(DEV)
    return q, k, kv, tab, d_o, qs, ks


@pytest.mark.parametrize("bwd_warps", [8, 16, 108, 116])
@pytest.mark.parametrize("nseq,H,W", [(2, 24, 24), (3, 8, 24), (2, 16, 24), (1, 32, 24), (1, 32, 32), (20, 24, 24)])
def test_attention_tc_fwd_bwd(nseq, H, W, bwd_warps):
    from ct_clip_b200 import _lib, ops
    _lib.check(_lib.lib().ctclip_debug_set_attn_bwd_warps(bwd_warps), "set warps")
    heads, dh = 8, 32
    n, I = H * W, heads * dh
    M = nseq * n
    sup = ops.attn_tc_supported(n, H, W, dh)
    assert sup & 1
    assert ops.attn_tc_supported(96, 4, 24, dh) == 0        # 96 keys: not a whole number of 64-row TMA boxes -> mma.sync path
    q, k, kv, tab, d_o, qs, ks = _inputs(nseq, H, W)
    v = kv[:, I:]
    qkb = torch.empty(1, device=DEV)
    ops.qk_bound(qs, ks, qkb)
    assert abs(qkb.item() - (qs * ks).abs().max().item()) < 1e-6
    geom = dict(n=n, heads=heads, num_seqs=nseq, seq_inner=1, seq_outer_stride=n, tok_stride=1)
    o = torch.zeros(M, I, dtype=torch.bfloat16, device=DEV)
    lse = torch.zeros(M, heads, device=DEV)
    ops.attn_fwd(q, k, v, o, lse, ldq=I, ldk=I, ldv=2 * I, ldo=I, cpb_table=tab, grid_hw=(H, W), qk_bound=qkb, **geom)
    torch.cuda.synchronize()

    def to_seq(x):
        return x.float().view(nseq, n, heads, dh).permute(0, 2, 1, 3)

    def from_seq(x):
        return x.permute(0, 2, 1, 3).reshape(M, I)

    rel = _rel_index(H, W).to(DEV)
    tabr = tab.clone().requires_grad_(True)
    bias = tabr[rel].permute(2, 0, 1)                                       # [heads, n, n]
    qr, kr, vr = (to_seq(t).requires_grad_(True) for t in (q, k, v))
    sim = qr @ kr.transpose(-1, -2) * 8.0 + bias
    oref = sim.softmax(-1) @ vr
    lse_ref = torch.logsumexp(sim, dim=-1) * 1.4426950408889634              # log2 domain, [nseq, heads, n]
    assert rel_err(o, from_seq(oref)) < 1e-2
    assert (lse - lse_ref.permute(0, 2, 1).reshape(M, heads)).abs().max().item() < 2e-3
    if not (sup & 2):
        return
    oref.backward(to_seq(d_o))
    dq = torch.zeros(M, I, dtype=torch.bfloat16, device=DEV)
    dkv = torch.zeros(M, 2 * I, dtype=torch.bfloat16, device=DEV)
    delta = torch.empty(M, heads, device=DEV)
    dtab = torch.zeros_like(tab)
    scratch = torch.full((nseq * heads * n * n,), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.attn_bwd(q, k, v, o, lse, d_o, delta, dq, dkv, dkv[:, I:], ldq=I, ldk=I, ldv=2 * I, ldo=I, ld_dq=I, ld_dk=2 * I,
                 ld_dv=2 * I, total_rows=M, cpb_table=tab, grid_hw=(H, W), dcpb_table=dtab, ds_scratch=scratch, **geom)
    torch.cuda.synchronize()
    assert rel_err(dq, from_seq(qr.grad)) < 2e-2
    assert rel_err(dkv[:, :I], from_seq(kr.grad)) < 2e-2
    assert rel_err(dkv[:, I:], from_seq(vr.grad)) < 2e-2
    assert rms_err(dtab, tabr.grad) < 2e-2
    # without the table gradient (no spill)
    dq2 = torch.zeros_like(dq)
    ops.attn_bwd(q, k, v, o, lse, d_o, delta, dq2, dkv, dkv[:, I:], ldq=I, ldk=I, ldv=2 * I, ldo=I, ld_dq=I, ld_dk=2 * I,
                 ld_dv=2 * I, total_rows=M, cpb_table=tab, grid_hw=(H, W), **geom)
    assert torch.equal(dq2, dq)
    # table gradient through fp32 reductions into the transposed [h][j][i] table + mirrored binning
    dbt = torch.zeros(heads, n, n, device=DEV)
    ops.attn_bwd(q, k, v, o, lse, d_o, delta, dq2, dkv, dkv[:, I:], ldq=I, ldk=I, ldv=2 * I, ldo=I, ld_dq=I, ld_dk=2 * I,
                 ld_dv=2 * I, total_rows=M, cpb_table=tab, grid_hw=(H, W), dbias=dbt, **geom)
    dtab2 = torch.zeros_like(tab)
    ops.cpb_reduce_t(dbt, heads, H, W, dtab2)
    assert rms_err(dtab2, tabr.grad) < 1e-2
    _lib.check(_lib.lib().ctclip_debug_set_attn_bwd_warps(108), "restore default")


def test_attention_tc_matches_mma_sync_path():
    """Both attention implementations on the same inputs (the mma.sync kernels take the expanded bf16 bias)."""
    from ct_clip_b200 import ops
    nseq, H, W, heads, dh = 2, 24, 24, 8, 32
    n, I = H * W, heads * dh
    M = nseq * n
    q, k, kv, tab, d_o, qs, ks = _inputs(nseq, H, W, seed=5)
    v = kv[:, I:]
    geom = dict(n=n, heads=heads, num_seqs=nseq, seq_inner=1, seq_outer_stride=n, tok_stride=1)
    o1 = torch.empty(M, I, dtype=torch.bfloat16, device=DEV)
    o2 = torch.empty_like(o1)
    l1 = torch.empty(M, heads, device=DEV)
    l2 = torch.empty_like(l1)
    qkb = torch.empty(1, device=DEV)
    ops.qk_bound(qs, ks, qkb)
    ops.attn_fwd(q, k, v, o1, l1, ldq=I, ldk=I, ldv=2 * I, ldo=I, cpb_table=tab, grid_hw=(H, W), qk_bound=qkb, **geom)
    bias = torch.empty(heads, n, n, dtype=torch.bfloat16, device=DEV)
    ops.cpb_expand(tab, heads, H, W, bias, None)
    ops.attn_fwd(q, k, v, o2, l2, ldq=I, ldk=I, ldv=2 * I, ldo=I, bias=bias, **geom)
    assert rel_err(o1, o2) < 1e-2
    assert (l1 - l2).abs().max().item() < 3e-2     # the bf16 bias of the mma.sync path moves the logits by up to ~1e-2
