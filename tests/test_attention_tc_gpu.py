"""EXPERIMENTAL tcgen05 / TMEM attention forward (csrc/attention_tc.cu) against fp32 softmax attention.

Skipped unless CTCLIP_EXPERIMENTAL=1: the kernel was written after the GPU budget of round 1 was spent and has not run on
hardware yet; the default path (mma.sync kernels, tests/test_kernels_gpu.py) does not depend on it."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("CTCLIP_EXPERIMENTAL") != "1", reason="experimental kernel: set CTCLIP_EXPERIMENTAL=1")]


@pytest.mark.parametrize("S,with_bias", [(64, False), (192, True), (576, True)])
def test_attention_tc_forward(S, with_bias):
    from ct_clip_b200 import ops
    dev = "cuda"
    b, T, heads, dh = 1, 2, 8, 32
    I, M = heads * dh, 1 * 2 * S
    g = torch.Generator(device="cpu").manual_seed(3)
    q = torch.nn.functional.normalize(torch.randn(M, heads, dh, generator=g), dim=-1).reshape(M, I).to(torch.bfloat16).to(dev)
    kv = torch.randn(M, 2 * I, generator=g)
    kv[:, :I] = torch.nn.functional.normalize(kv[:, :I].reshape(M, heads, dh), dim=-1).reshape(M, I)
    kv = kv.to(torch.bfloat16).to(dev)
    k, v = kv[:, :I], kv[:, I:]
    bias = (0.5 * torch.randn(heads, S, S, generator=g)).to(torch.bfloat16).to(dev) if with_bias else None
    o = torch.empty(M, I, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(M, heads, device=dev)
    geom = dict(n=S, heads=heads, num_seqs=b * T, seq_inner=1, seq_outer_stride=S, tok_stride=1)
    ops.attn_fwd_tc(q, k, v, o, lse, ldq=I, ldk=2 * I, ldv=2 * I, ldo=I, bias=bias, **geom)
    torch.cuda.synchronize()

    def to_seq(x):      # [M, heads*dh] -> [seqs, heads, S, dh]
        return x.float().reshape(b * T, S, heads, dh).permute(0, 2, 1, 3)
    sim = to_seq(q) @ to_seq(k).transpose(-1, -2) * 8.0
    if bias is not None:
        sim = sim + bias.float()
    ref = (sim.softmax(-1) @ to_seq(v)).permute(0, 2, 1, 3).reshape(M, I)
    err = ((o.float() - ref).abs().max() / ref.abs().max()).item()
    assert err < 1e-2, err
    lse_ref = torch.logsumexp(sim, dim=-1).permute(0, 2, 1).reshape(M, heads) * 1.4426950408889634   # log2 domain
    assert ((lse - lse_ref).abs().max() / lse_ref.abs().max()).item() < 1e-3
