"""GEMM family (tcgen05/TMA) against a plain fp32 matmul of the same bf16-rounded operands.

Tolerance: operands are identical bf16 values on both sides, accumulation is fp32 on both, so
the only difference is summation order -> rel 2e-3 of the output scale (bf16 outputs: + 2^-8).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(rows, cols, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(rows, cols, generator=g) * scale).to(torch.bfloat16).cuda()


def _close(got, ref, tol):
    got = got.float()
    ref = ref.float()
    denom = ref.abs().max().clamp_min(1e-6)
    err = ((got - ref).abs().max() / denom).item()
    assert err < tol, f"rel err {err:.3e} >= {tol}"
    return err


SHAPES = [
    (128, 256, 64), (256, 512, 512), (384, 768, 512), (1000, 2816, 512), (4096, 512, 1408),
    (130, 96, 200), (128, 64, 4000), (256, 8192, 512), (8, 512, 768),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_nt_bf16_out(M, N, K):
    from ct_clip_b200 import ops
    A, B = _mk(M, K, 1), _mk(N, K, 2)
    Cb = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    ops.gemm(A, B, M=M, N=N, K=K, epilogue=ops.EPI_BF16, C_out=Cb)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    _close(Cb, ref, 1e-2)


@pytest.mark.parametrize("M,N,K", [(256, 512, 512), (1000, 2816, 512), (130, 96, 200)])
def test_gemm_f32_bias_resid(M, N, K):
    from ct_clip_b200 import ops
    A, B = _mk(M, K, 3), _mk(N, K, 4)
    bias = torch.randn(N, device="cuda")
    resid = torch.randn(M, N, device="cuda")
    Cf = torch.empty(M, N, device="cuda")
    ops.gemm(A, B, M=M, N=N, K=K, epilogue=ops.EPI_F32, C_out=Cf, bias=bias)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t() + bias
    _close(Cf, ref, 2e-3)
    out = resid.clone()
    ops.gemm(A, B, M=M, N=N, K=K, epilogue=ops.EPI_RESID_F32, C_out=out, resid=out, bias=bias)
    torch.cuda.synchronize()
    _close(out, ref + resid, 2e-3)


@pytest.mark.parametrize("M,N,K", [(256, 2816, 512), (384, 256, 128)])
def test_gemm_geglu(M, N, K):
    from ct_clip_b200 import ops
    A, B = _mk(M, K, 5), _mk(N, K, 6, 0.05)
    H = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    G = torch.empty(M, N // 2, dtype=torch.bfloat16, device="cuda")
    ops.gemm(A, B, M=M, N=N, K=K, epilogue=ops.EPI_GEGLU, C_out=H, C2=G)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    _close(H, ref, 1e-2)
    refg = torch.nn.functional.gelu(ref[:, 1::2]) * ref[:, 0::2]
    _close(G, refg, 1e-2)


@pytest.mark.parametrize("M,N,K", [(1000, 1408, 512), (384, 64, 128), (300, 96, 256)])
def test_gemm_geglu_bwd_fused(M, N, K):
    """epilogue 8: dg = A W (W stored [K, N], MN-major as in the engine) fused with the GEGLU backward on the interleaved
    pre-activation h (in place) and the column sums of the result (attention.py:39-42 backward)."""
    from ct_clip_b200 import ops
    A, W = _mk(M, K, 21), _mk(K, N, 22, 0.05)
    h = _mk(M, 2 * N, 23)
    h0 = h.clone()
    cs = torch.zeros(2 * N, device="cuda")
    ops.gemm(A, W, M=M, N=N, K=K, b_major=1, epilogue=ops.EPI_GEGLU_BWD, C_out=h, ldc=2 * N, colsum=cs)
    torch.cuda.synchronize()
    dg = A.float() @ W.float()
    value, gate = h0.float()[:, 0::2].clone().requires_grad_(True), h0.float()[:, 1::2].clone().requires_grad_(True)
    (torch.nn.functional.gelu(gate) * value * dg).sum().backward()
    ref = torch.empty(M, 2 * N, device="cuda")
    ref[:, 0::2], ref[:, 1::2] = value.grad, gate.grad
    _close(h, ref, 1e-2)
    _close(cs, ref.sum(0), 1e-2)


@pytest.mark.parametrize("amaj,bmaj", [(1, 1), (0, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K,splits", [(768, 512, 4096, 1), (2816, 512, 8192, 8), (512, 4000, 1024, 3), (256, 128, 200, 2)])
def test_gemm_mn_major_atomic(amaj, bmaj, M, N, K, splits):
    """Weight-gradient form: operands stored [K, rows]; split-K accumulation with red.add."""
    from ct_clip_b200 import ops
    At, Bt = _mk(K, M, 7), _mk(K, N, 8)  # stored reduction-major
    A = At if amaj == 1 else At.t().contiguous()
    B = Bt if bmaj == 1 else Bt.t().contiguous()
    Cf = torch.zeros(M, N, device="cuda")
    ops.gemm(A, B, M=M, N=N, K=K, a_major=amaj, b_major=bmaj, epilogue=ops.EPI_ATOMIC_F32, C_out=Cf, splits=splits)
    torch.cuda.synchronize()
    ref = At.float().t() @ Bt.float()
    _close(Cf, ref, 2e-3)


def test_gemm_argmax():
    from ct_clip_b200 import ops
    M, N, K = 1000, 8192, 512
    A, B = _mk(M, K, 9), _mk(N, K, 10)
    idx = torch.empty(M, dtype=torch.int32, device="cuda")
    val = torch.empty(M, device="cuda")
    ops.gemm(A, B, M=M, N=N, K=K, epilogue=ops.EPI_ARGMAX, arg_out=idx, argval_out=val)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    rv, ri = ref.max(dim=1)
    # values must agree; indices may differ only where the top-2 gap is below fp32 summation noise
    _close(val, rv, 2e-3)
    agree = (idx.long() == ri).float().mean().item()
    assert agree > 0.995, agree
    picked = ref.gather(1, idx.long()[:, None])[:, 0]
    assert ((rv - picked).abs() <= 1e-3 * rv.abs().max()).all()


def test_gemm_argmax_top2_and_fp32_rerank():
    """ARGMAX epilogue with the runner-up index + ctclip_vq_rerank: the pair must be the true top-2 of the bf16 scores, and
    after the fp32 re-ranking the index must equal the fp32 cosine code-book argmax (vector_quantize_pytorch 1.1.2 ranking)
    except where even the fp32 top-3 is closer than summation noise."""
    from ct_clip_b200 import ops
    M, C, D = 3000, 8192, 512
    g = torch.Generator().manual_seed(21)
    x = torch.randn(M, D, generator=g).cuda()                       # quantiser input (fp32)
    embed = torch.nn.functional.normalize(torch.randn(C, D, generator=g), dim=-1).cuda() * (0.8 + 0.4 * torch.rand(C, 1, generator=g).cuda())
    xb = x.to(torch.bfloat16)
    ehat = torch.empty(C, D, dtype=torch.bfloat16, device="cuda")
    ops.l2norm_rows_bf16(embed, ehat, C, D)
    idx = torch.empty(M, dtype=torch.int32, device="cuda")
    idx2 = torch.empty(M, dtype=torch.int32, device="cuda")
    ops.gemm(xb, ehat, M=M, N=C, K=D, epilogue=ops.EPI_ARGMAX, arg_out=idx, arg2_out=idx2)
    torch.cuda.synchronize()
    sb = xb.float() @ ehat.float().t()                               # the scores the GEMM ranks
    top = sb.topk(3, dim=1)
    got1, got2 = sb.gather(1, idx.long()[:, None])[:, 0], sb.gather(1, idx2.long()[:, None])[:, 0]
    tol = 1e-3 * sb.abs().max()
    assert (idx != idx2).all()
    assert ((top.values[:, 0] - got1).abs() <= tol).all() and ((top.values[:, 1] - got2).abs() <= tol).all()
    before = (idx.long() == (torch.nn.functional.normalize(x, dim=-1) @ torch.nn.functional.normalize(embed, dim=-1).t()).argmax(1)).float().mean().item()
    ops.vq_rerank(x, embed, idx, idx2, M, D)
    torch.cuda.synchronize()
    ref = (torch.nn.functional.normalize(x, dim=-1) @ torch.nn.functional.normalize(embed, dim=-1).t()).argmax(1)
    after = (idx.long() == ref).float().mean().item()
    print(f"index agreement with the fp32 quantiser: {before:.4f} (bf16 argmax) -> {after:.4f} (top-2 + fp32 re-rank)")
    assert after >= 0.998 and after >= before, (before, after)
