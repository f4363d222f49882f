"""CTClipTrainer (flat arena, fused clip+Adam, prefetching train_step) and CTClipInference (zero-shot path) on a GPU."""
import numpy as np
import pytest
import torch

from tests.helpers import CFG1_VIT, oracle_vit_cfg
from tests.test_ctclip_gpu import _Tok, build_clip

pytestmark = pytest.mark.gpu


def test_trainer_step_matches_oracle_adam(tmp_path):
    from ct_clip_b200.data import SyntheticCTReportDataset
    from ct_clip_b200.trainer import CTClipTrainer
    from oracle import ctclip_oracle as O
    kw = dict(CFG1_VIT, spatial_depth=1, temporal_depth=1)
    clip, sd, cfg = build_clip(kw, bert_layers=1)
    ds = SyntheticCTReportDataset(8, frames=32, image=64, n_text=32)
    lr = 1e-3
    tr = CTClipTrainer(clip, num_train_steps=2, batch_size=2, train_dataset=ds, num_workers=0, lr=lr, save_model_every=0,
                       results_folder=str(tmp_path))
    hu, ids, mask = O.synth_inputs(2, 32, 64, 32)
    video = hu.float() / 1000.0
    # oracle: gradients -> clip_grad_norm_(0.5) -> Adam(betas=(0.9, 0.99)) first step
    sdp = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    out = O.ctclip_forward(sdp, cfg, ids, mask, video, training=True)
    out["loss"].backward()
    trainable = {k for k, _ in clip.named_parameters()}     # buffers (LayerNorm.beta, attention.py:32; code-book) are not trained
    live = [k for k, v in sdp.items() if v.is_floating_point() and v.grad is not None and not k.endswith("_extra.weight")
            and k in trainable]
    ps = [torch.nn.Parameter(sd[k].clone()) for k in live]
    for p_, k in zip(ps, live):
        p_.grad = sdp[k].grad.clone()
    torch.nn.utils.clip_grad_norm_(ps, 0.5)
    torch.optim.Adam(ps, lr=lr, betas=(0.9, 0.99), eps=1e-8).step()
    clip.visual_transformer._force_indices = out["indices"]
    loss = tr.step_on_batch(hu.cuda(), _Tok(ids.cuda(), mask.cuda()))
    torch.cuda.synchronize()
    assert abs(loss.item() - out["loss"].item()) < 1e-2 * abs(out["loss"].item())
    now = dict(clip.named_parameters())
    agree, total = 0, 0
    for p_, k in zip(ps, live):
        d_ref = (p_.data - sd[k]).reshape(-1)
        d_got = (now[k].detach().cpu() - sd[k]).reshape(-1)
        big = sdp[k].grad.reshape(-1).abs() > 1e-3 * sdp[k].grad.abs().max()       # elements whose update is not eps-dominated
        assert d_got.abs().max().item() <= lr * 1.001 + 1e-9, k                    # an Adam step never exceeds lr per element
        agree += (torch.sign(d_ref[big]) == torch.sign(d_got[big])).sum().item()
        total += int(big.sum())
    assert agree / total > 0.97, agree / total
    # parameters without gradient on this path stay untouched (torch's Adam skips them too)
    assert torch.equal(clip.to_visual_latent_extra.weight.cpu(), sd["to_visual_latent_extra.weight"])
    # public loop: two prefetched steps run and log a finite loss
    clip.visual_transformer._force_indices = None
    logs = tr.train_step()
    assert np.isfinite(logs["loss"])
    tr.save(str(tmp_path / "ck.pt"))
    tr.load(str(tmp_path / "ck.pt"))


def test_zero_shot_inference_matches_oracle(tmp_path):
    from ct_clip_b200.inference import PATHOLOGIES, CTClipInference
    from oracle import ctclip_oracle as O
    kw = dict(CFG1_VIT, spatial_depth=1, temporal_depth=1)
    clip, sd, cfg = build_clip(kw, bert_layers=1)
    nvol = 3
    hu, _, _ = O.synth_inputs(nvol, 32, 64, 32)
    _, pids, pmask = O.synth_inputs(36, 4, 16, 32, seed=99)      # stand-in token ids of the 36 prompts (no tokenizer offline)

    class DS(torch.utils.data.Dataset):
        def __len__(self):
            return nvol

        def __getitem__(self, i):
            return hu[i], "report", np.zeros(18, dtype=np.float32), f"acc{i}"

    inf = CTClipInference(clip, dataset=DS(), prompt_tokens=dict(input_ids=pids, attention_mask=pmask),
                          results_folder=str(tmp_path))
    probs = inf.infer()
    assert probs.shape == (nvol, len(PATHOLOGIES))
    # reference semantics (zero_shot.py:133-143): per volume and pathology, softmax over the 2 prompt similarities
    video = hu.float() / 1000.0
    with torch.no_grad():
        ref = np.zeros((nvol, 18), dtype=np.float32)
        for v in range(nvol):
            for p_ in range(18):
                o = O.ctclip_forward(sd, cfg, pids[2 * p_:2 * p_ + 2], pmask[2 * p_:2 * p_ + 2], video[v:v + 1], training=False,
                                     return_loss=False)
                ref[v, p_] = torch.softmax(o["sims"], dim=0)[0].item()
    # VQ index flips (bf16 encoder) perturb the image latent; probabilities are compared loosely, ranking strictly enough
    assert np.abs(probs - ref).max() < 5e-2, np.abs(probs - ref).max()
    assert (tmp_path / "predicted_weights.npz").exists()
