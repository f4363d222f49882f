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
