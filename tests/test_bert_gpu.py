"""Native BERT text tower (sm_100a kernels) vs the CPU oracle restatement of transformers.BertModel:
last_hidden_state and every parameter gradient for a CLS-only upstream gradient (what CT-CLIP feeds it)."""
import pytest
import torch

from tests.helpers import rel_err, rms_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("b,n,layers", [(2, 32, 2), (8, 128, 2), (3, 100, 1)])
def test_bert_engine_fwd_bwd(b, n, layers):
    from transformers import BertConfig, BertModel

    from ct_clip_b200.bert import BertEngine, supports
    from oracle import ctclip_oracle as O
    bert = BertModel(BertConfig(num_hidden_layers=layers, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0))
    assert supports(bert)
    shapes = {k: tuple(v.shape) for k, v in bert.state_dict().items()}
    sd = O.synth_state_dict({"text_transformer." + k: v for k, v in shapes.items()}, 1)
    bert.load_state_dict({k[len("text_transformer."):]: v for k, v in sd.items()}, strict=True)
    bert = bert.cuda()
    _, ids, mask = O.synth_inputs(b, 4, 16, n, seed=5)
    # oracle
    sdp = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    ref = O.bert_forward(ids, mask, sdp, "text_transformer.", O.BertConfigLite(layers=layers))
    gcls = torch.randn(b, 768, generator=torch.Generator().manual_seed(3))
    (ref[:, 0, :] * gcls).sum().backward()
    # engine
    eng = BertEngine(bert, torch.device("cuda"))
    P = dict(bert.named_parameters())
    last, ctx = eng.forward(ids.cuda(), mask.cuda(), P, save=True)
    valid = mask.bool()
    assert rms_err(last.cpu()[valid], ref.detach()[valid]) < 1e-2          # padded positions are don't-care
    assert rms_err(last[:, 0, :], ref[:, 0, :]) < 1e-2
    G = {k: torch.zeros_like(v) for k, v in P.items()}
    d_last = torch.zeros(b * n, 768, device="cuda")
    d_last.view(b, n, 768)[:, 0, :] = gcls.cuda()
    eng.backward(ctx, d_last, P, G)
    torch.cuda.synchronize()
    gmax = max(v.grad.abs().max().item() for v in sdp.values() if v.is_floating_point() and v.grad is not None)
    bad = []
    for k, g in G.items():
        r = sdp["text_transformer." + k].grad
        if r is None or r.abs().max().item() < 1e-6 * gmax:
            assert g.abs().max().item() < 1e-2 * gmax, k
            continue
        e = rms_err(g, r)
        if e > 2e-2:
            bad.append((k, e))
    assert not bad, bad


def _oracle_masks(seed, b, n, heads, H, layers, ph, pa):
    """KEEP masks of every dropout site, rebuilt on the CPU from the same (seed, site offset) counters (tests/philox_ref.py)."""
    from ct_clip_b200.bert import site_offset
    from tests.philox_ref import keep_mask
    masks = {(-1, 3): torch.from_numpy(keep_mask(seed, site_offset(-1, 3), b * n * H, ph)).view(b, n, H)}
    for i in range(layers):
        masks[(i, 0)] = torch.from_numpy(keep_mask(seed, site_offset(i, 0), b * heads * n * n, pa)).view(b, heads, n, n)
        masks[(i, 1)] = torch.from_numpy(keep_mask(seed, site_offset(i, 1), b * n * H, ph)).view(b, n, H)
        masks[(i, 2)] = torch.from_numpy(keep_mask(seed, site_offset(i, 2), b * n * H, ph)).view(b, n, H)
    return masks


@pytest.mark.parametrize("b,n,layers", [(2, 32, 2), (4, 128, 2)])
def test_bert_engine_dropout_matches_oracle_with_same_masks(b, n, layers):
    """Training-mode BERT (hidden / attention dropout 0.1, the CXR-BERT configuration of run_train.py:7-9) on the native kernels:
    the Philox masks the kernels generate are rebuilt on the CPU and handed to the oracle; outputs and all gradients must agree
    to the same tolerance as without dropout."""
    from transformers import BertConfig, BertModel

    from ct_clip_b200.bert import BertEngine, dropout_config, supports
    from oracle import ctclip_oracle as O
    ph = pa = 0.1
    bert = BertModel(BertConfig(num_hidden_layers=layers, hidden_dropout_prob=ph, attention_probs_dropout_prob=pa))
    assert supports(bert)
    shapes = {k: tuple(v.shape) for k, v in bert.state_dict().items()}
    sd = O.synth_state_dict({"text_transformer." + k: v for k, v in shapes.items()}, 1)
    bert.load_state_dict({k[len("text_transformer."):]: v for k, v in sd.items()}, strict=True)
    bert = bert.cuda().train()
    _, ids, mask = O.synth_inputs(b, 4, 16, n, seed=5)
    seed = 0x1234567890AB
    drop = dropout_config(bert, seed)
    assert drop is not None and drop["p_hidden"] == ph
    masks = _oracle_masks(seed, b, n, 12, 768, layers, ph, pa)
    keep_frac = masks[(0, 0)].float().mean().item()
    assert abs(keep_frac - 0.9) < 0.01, keep_frac
    sdp = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    ref = O.bert_forward(ids, mask, sdp, "text_transformer.", O.BertConfigLite(layers=layers),
                         dropout=dict(p_hidden=ph, p_attn=pa, masks=masks))
    gcls = torch.randn(b, 768, generator=torch.Generator().manual_seed(3))
    (ref[:, 0, :] * gcls).sum().backward()
    eng = BertEngine(bert, torch.device("cuda"))
    P = dict(bert.named_parameters())
    last, ctx = eng.forward(ids.cuda(), mask.cuda(), P, save=True, dropout=drop)
    valid = mask.bool()
    assert rms_err(last.cpu()[valid], ref.detach()[valid]) < 1e-2
    assert rms_err(last[:, 0, :], ref[:, 0, :]) < 1e-2
    # no dropout at all would be far away: the test really exercises the masks
    last0, _ = eng.forward(ids.cuda(), mask.cuda(), P, save=False)
    assert rms_err(last0[:, 0, :], ref[:, 0, :]) > 5e-2
    G = {k: torch.zeros_like(v) for k, v in P.items()}
    d_last = torch.zeros(b * n, 768, device="cuda")
    d_last.view(b, n, 768)[:, 0, :] = gcls.cuda()
    eng.backward(ctx, d_last, P, G)
    torch.cuda.synchronize()
    gmax = max(v.grad.abs().max().item() for v in sdp.values() if v.is_floating_point() and v.grad is not None)
    bad = []
    for k, g in G.items():
        r = sdp["text_transformer." + k].grad
        if r is None or r.abs().max().item() < 1e-6 * gmax:
            continue
        e = rms_err(g, r)
        if e > 2e-2:
            bad.append((k, e))
    assert not bad, bad


def test_ctclip_text_tower_is_native_with_dropout():
    """CTCLIP keeps BertModel(hidden_dropout_prob=0.1) on the native kernels in training mode (round 1 fell back to HF eager)."""
    from transformers import BertConfig, BertModel

    from ct_clip_b200 import CTCLIP, CTViT
    vit = CTViT(dim=512, codebook_size=256, image_size=32, patch_size=16, temporal_patch_size=4, spatial_depth=1, temporal_depth=1,
                dim_head=32, heads=8)
    bert = BertModel(BertConfig(num_hidden_layers=1))        # HF defaults: both dropouts 0.1
    clip = CTCLIP(image_encoder=vit, text_encoder=bert, dim_text=768, dim_image=4 * 512, dim_latent=128).cuda().train()
    assert clip._text_native() and clip._text_dropout() is not None
    clip.eval()
    assert clip._text_native() and clip._text_dropout() is None
