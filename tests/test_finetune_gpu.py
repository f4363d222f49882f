"""fine-tuning fast paths against the reference scripts' literal per-prompt loops run through the SAME model."""
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import rms_err
from tests.test_ctclip_gpu import _Tok

pytestmark = pytest.mark.gpu

SMALL_VIT = dict(dim=512, codebook_size=256, image_size=32, patch_size=16, temporal_patch_size=4, spatial_depth=1, temporal_depth=1,
                 dim_head=32, heads=8)


def _build(seed=0):
    from transformers import BertConfig, BertModel

    from ct_clip_b200 import CTCLIP, CTViT
    torch.manual_seed(seed)
    vit = CTViT(**SMALL_VIT)
    bert = BertModel(BertConfig(num_hidden_layers=1, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, vocab_size=1000))
    return CTCLIP(image_encoder=vit, text_encoder=bert, dim_text=768, dim_image=4 * 512, dim_latent=128).cuda()


def _bank(n_prompts, n=16, seed=3):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(5, 1000, (n_prompts, n), generator=g)
    ids[:, 0] = 2
    mask = torch.ones(n_prompts, n, dtype=torch.long)
    mask[:, 12:] = 0
    return _Tok((ids * mask).cuda(), mask.cuda())


def test_vocabfine_fast_path_matches_the_per_prompt_loop():
    from ct_clip_b200.finetune import vocabfine_step
    C = 6
    clip = _build()
    clip.eval()                                   # no code-book EMA between the repeated forwards of the literal loop
    for p in clip.parameters():
        p.requires_grad_(True)
    vol = torch.randn(1, 1, 8, 32, 32, generator=torch.Generator().manual_seed(1)).clamp(-1, 1).cuda()
    labels = torch.tensor([1, 0, 0, 1, 1, 0])
    bank = _bank(2 * C)
    # ---- literal loop of ct_vocabfine_train.py:86-119: one model call per pathology, (text_yes, text_no) ordered by the label
    logits_list = []
    for l in range(C):
        yes, no = (2 * l, 2 * l + 1) if labels[l] == 1 else (2 * l + 1, 2 * l)
        tok = _Tok(bank.input_ids[[yes, no]], bank.attention_mask[[yes, no]])
        out = clip(tok, vol, device="cuda")                      # (2,) similarities, differentiable
        assert out.shape == (2,) and out.requires_grad
        logits_list.append(F.softmax(out, dim=0))
    loss_ref = F.mse_loss(torch.cat(logits_list), torch.tensor([1.0, 0.0] * C).cuda())
    loss_ref.backward()
    ref = {n: p.grad.clone() for n, p in clip.named_parameters() if p.grad is not None}
    clip.zero_grad(set_to_none=True)
    # ---- fast path: one text batch, one image forward, one backward
    loss = vocabfine_step(clip, vol, labels, bank)
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - loss_ref.item()) < 2e-3 * max(1.0, abs(loss_ref.item())), (loss.item(), loss_ref.item())
    gmax = max(v.abs().max().item() for v in ref.values())
    checked = 0
    for n, p in clip.named_parameters():
        if n not in ref or ref[n].abs().max().item() < 1e-3 * gmax or n.endswith("spatial_rel_pos_bias.net.2.bias"):
            continue          # (net.2.bias: a per-head constant added to the logits, softmax-invariant -> analytically zero gradient)
        assert p.grad is not None, n
        assert rms_err(p.grad, ref[n]) < 5e-2, (n, rms_err(p.grad, ref[n]))       # the loop accumulates 6 bf16 backward passes
        checked += 1
    assert checked > 20


def test_lipro_classifier_uses_only_the_image_tower():
    from ct_clip_b200.finetune import ImageLatentsClassifier, lipro_loss
    clip = _build(1)
    head = ImageLatentsClassifier(clip, 128, 18).cuda().train()
    assert all(not p.requires_grad for p in clip.parameters())
    vols = torch.randn(3, 1, 8, 32, 32, generator=torch.Generator().manual_seed(2)).clamp(-1, 1).cuda()
    tok = _bank(1)
    head.dropout.p = 0.0
    logits = head(tok, vols, device="cuda", return_latents=True)
    with torch.no_grad():
        clip.eval()
        _, il, _ = clip(_Tok(tok.input_ids.expand(3, -1), tok.attention_mask.expand(3, -1)), vols, device="cuda", return_latents=True)
    ref = head.classifier(torch.relu(il))
    assert logits.shape == (3, 18) and torch.allclose(logits, ref, atol=1e-5)
    labels = (torch.rand(3, 18, generator=torch.Generator().manual_seed(4)) < 0.3).float().cuda()
    lipro_loss(logits, labels).backward()
    assert head.classifier.weight.grad is not None and all(p.grad is None for p in clip.parameters())
