"""CPU suite part 1: pin the oracle. (a) against the committed golden vectors produced by the unmodified reference,
(b) against the reference itself when /root/reference is mounted (build container only), (c) BERT restatement
against transformers.BertModel, (d) properties of the restated vector quantiser."""
from pathlib import Path

import pytest
import torch

from oracle import ctclip_oracle as O
from oracle import ref_shims, vq_restated

GOLD = Path(__file__).resolve().parent / "golden"


def _sub_close(t, ref, tol):
    f = t.detach().reshape(-1).float()
    k = ref["head"].numel()
    scale = max(ref["norm"] / max(ref["numel"], 1) ** 0.5, 1e-12)
    assert f.numel() == ref["numel"]
    assert (f[:k] - ref["head"]).abs().max().item() <= tol * max(scale, ref["head"].abs().max().item())
    assert (f[-k:] - ref["tail"]).abs().max().item() <= tol * max(scale, ref["tail"].abs().max().item())
    assert abs(f.norm().item() - ref["norm"]) <= tol * max(ref["norm"], 1e-12)


def _cfg(case):
    v = case["vit"]
    return O.CTCLIPConfig(vit=O.CTViTConfig(dim=v["dim"], codebook_size=v["codebook_size"], image_size=v["image_size"],
                                            patch_size=v["patch_size"], temporal_patch_size=v["temporal_patch_size"],
                                            spatial_depth=v["spatial_depth"], temporal_depth=v["temporal_depth"],
                                            dim_head=v["dim_head"], heads=v["heads"]),
                          bert=O.BertConfigLite(layers=case["bert_layers"]))


@pytest.mark.parametrize("name", ["cfg1", "scramble", "cfg5_small"])
def test_oracle_matches_golden_reference_outputs(name):
    g = torch.load(GOLD / f"{name}.pt", weights_only=False)
    case, cfg = g["case"], _cfg(g["case"])
    shapes = {k: tuple(v) for k, v in g["shapes"].items()}
    sd = O.synth_state_dict(shapes, 0)
    hu, ids, mask = O.synth_inputs(case["b"], case["frames"], case["vit"]["image_size"], case["n_text"])
    video = hu.float() / 1000.0
    with torch.no_grad():
        ev = O.ctclip_forward(sd, cfg, ids, mask, video, training=False, return_loss=False)
        assert torch.equal(ev["indices"], g["eval"]["indices"])
        assert torch.allclose(ev["text_latents"], g["eval"]["text_latents"], atol=2e-6)
        assert torch.allclose(ev["image_latents"], g["eval"]["image_latents"], atol=2e-6)
        _sub_close(ev["tokens"], g["eval"]["tokens"], 1e-5)
        s2 = O.ctclip_forward(sd, cfg, ids[:2], mask[:2], video[:1], training=False, return_loss=False)["sims"]
        assert torch.allclose(s2, g["eval"]["sims"], atol=1e-5)
    sdp = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    out = O.ctclip_forward(sdp, cfg, ids, mask, video, training=True)
    out["loss"].backward()
    assert abs(out["loss"].item() - g["train"]["loss"]) < 1e-5
    gmax = max(r["norm"] for r in g["train"]["grads"].values())
    for n, ref in g["train"]["grads"].items():
        got = sdp[n].grad
        assert got is not None, n
        if ref["norm"] < 1e-6 * gmax:      # analytically-zero gradients (key bias, last CPB bias): round-off only
            assert got.norm().item() < 1e-5 * gmax, n
            continue
        _sub_close(got, ref, 2e-3)
    for n in g["train"]["no_grad"]:        # parameters the reference never reaches must not be reached here either
        gg = sdp[n].grad
        assert gg is None or gg.abs().max().item() == 0, n
    _sub_close(out["ema"][0], g["train"]["embed"], 1e-5)
    _sub_close(out["ema"][1], g["train"]["cluster_size"], 1e-6)


@pytest.mark.skipif(not ref_shims.reference_available(), reason="reference checkout not mounted (GPU box)")
def test_oracle_matches_reference_live():
    from transformers import BertConfig, BertModel
    CTViT, CTCLIP = ref_shims.load_reference()
    kw = dict(dim=256, codebook_size=256, image_size=32, patch_size=8, temporal_patch_size=4, spatial_depth=2, temporal_depth=1,
              dim_head=32, heads=8)
    vit = CTViT(**kw)
    bert = BertModel(BertConfig(num_hidden_layers=1, attn_implementation="eager", hidden_dropout_prob=0.0,
                                attention_probs_dropout_prob=0.0))
    clip = CTCLIP(image_encoder=vit, text_encoder=bert, dim_text=768, dim_image=16 * 256, dim_latent=512)
    sd = O.synth_state_dict({k: tuple(v.shape) for k, v in clip.state_dict().items()}, 3)
    clip.load_state_dict(sd, strict=True)
    hu, ids, mask = O.synth_inputs(3, 12, 32, 16, seed=7)
    video = hu.float() / 1000.0

    class Tok:
        input_ids, attention_mask = ids, mask
    clip.train()
    loss = clip(Tok, video, device="cpu", return_loss=True)
    loss.backward()
    cfg = O.CTCLIPConfig(vit=O.CTViTConfig(dim=256, codebook_size=256, image_size=32, patch_size=8, temporal_patch_size=4,
                                           spatial_depth=2, temporal_depth=1), bert=O.BertConfigLite(layers=1))
    sdp = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    out = O.ctclip_forward(sdp, cfg, ids, mask, video, training=True)
    out["loss"].backward()
    assert abs(out["loss"].item() - loss.item()) < 1e-6
    gmax = max(p.grad.abs().max().item() for p in clip.parameters() if p.grad is not None and p.grad.numel())
    for n, p in clip.named_parameters():
        if p.grad is None or p.grad.numel() == 0:
            continue
        assert (p.grad - sdp[n].grad).abs().max().item() < 2e-5 * gmax, n
    assert torch.allclose(vit.vq._codebook.embed[0], out["ema"][0], atol=1e-6)


def test_bert_restatement_matches_transformers():
    from transformers import BertConfig, BertModel
    torch.manual_seed(0)
    bert = BertModel(BertConfig(num_hidden_layers=2, attn_implementation="eager")).eval()
    sd = {"t." + k: v for k, v in bert.state_dict().items()}
    _, ids, mask = O.synth_inputs(3, 4, 16, 24)
    with torch.no_grad():
        ref = bert(ids, attention_mask=mask)[0]
        got = O.bert_forward(ids, mask, sd, "t.", O.BertConfigLite(layers=2))
    assert torch.allclose(ref, got, atol=2e-5)


def test_vq_restated_semantics():
    torch.manual_seed(1)
    embed = torch.nn.functional.normalize(torch.randn(8, 4), dim=-1)
    embed[5] = embed[2]                                   # exact tie: first index must win (torch.argmax semantics)
    x = embed[[5, 0, 7]] * torch.tensor([[3.0], [0.5], [2.0]])   # positive scaling must not change the winner
    q, ind, flat = vq_restated.vq_cosine_lookup(x, embed)
    assert ind.tolist() == [2, 0, 7]
    assert torch.equal(q, embed[ind])                    # un-renormalised rows of the stored buffer
    new_e, new_c = vq_restated.vq_ema_update(flat, ind, embed, torch.zeros(8), decay=0.8)
    assert torch.allclose(new_c, torch.tensor([0.2, 0, 0.2, 0, 0, 0, 0, 0.2]))
    untouched = [1, 3, 4, 5, 6]
    assert torch.allclose(new_e[untouched], embed[untouched], atol=1e-6)        # 0.8 e + 0.2 l2norm(e) == e for unit rows
    vq = vq_restated.VectorQuantize(dim=4, codebook_size=8)
    vq.train()
    xx = torch.randn(2, 5, 4, requires_grad=True)
    qq, ii, loss = vq(xx)
    qq.sum().backward()
    assert torch.allclose(xx.grad, torch.ones_like(xx))  # straight-through estimator
    assert ii.shape == (2, 5) and loss.shape == (1,)


def test_oracle_bert_dropout_sites_match_transformers_train_mode():
    """Pins the oracle's dropout placement (embeddings / attention probabilities / BertSelfOutput / BertOutput) against
    transformers.BertModel in TRAINING mode: torch.nn.functional.dropout is replaced by a function that applies the oracle's
    keep masks in call order, so both sides drop exactly the same elements."""
    import torch.nn.functional as F
    from transformers import BertConfig, BertModel

    from oracle import ctclip_oracle as O
    layers, b, n, p = 2, 2, 12, 0.1
    torch.manual_seed(0)
    bert = BertModel(BertConfig(num_hidden_layers=layers, attn_implementation="eager", hidden_dropout_prob=p,
                                attention_probs_dropout_prob=p)).train()
    sd = {"t." + k: v for k, v in bert.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(5, 1000, (b, n), generator=g)
    mask = torch.ones(b, n, dtype=torch.long)
    mask[1, 8:] = 0
    masks = {(-1, 3): torch.rand(b, n, 768, generator=g) > p}
    order = [(-1, 3)]
    for i in range(layers):
        masks[(i, 0)] = torch.rand(b, 12, n, n, generator=g) > p
        masks[(i, 1)] = torch.rand(b, n, 768, generator=g) > p
        masks[(i, 2)] = torch.rand(b, n, 768, generator=g) > p
        order += [(i, 0), (i, 1), (i, 2)]
    ref = O.bert_forward(ids, mask, sd, "t.", O.BertConfigLite(layers=layers), dropout=dict(p_hidden=p, p_attn=p, masks=masks))
    calls = []
    real = F.dropout

    def fake(x, p=0.5, training=True, inplace=False):
        if not training or p == 0:
            return x
        site = order[len(calls)]
        calls.append(site)
        return x * masks[site].to(x.dtype).reshape(x.shape) / (1.0 - p)
    F.dropout = fake
    torch.nn.functional.dropout = fake
    try:
        hf = bert(ids, attention_mask=mask)[0]
    finally:
        F.dropout = real
        torch.nn.functional.dropout = real
    assert calls == order, calls
    valid = mask.bool()
    assert (hf[valid] - ref[valid]).abs().max().item() < 1e-4


def _preprocess_case():
    """A seeded stored-value volume (H, W, D order of a NIfTI array) that exercises resampling in all three axes, the HU clip on
    both sides, a centre crop in-plane and padding in depth."""
    import numpy as np
    rng = np.random.default_rng(11)
    raw = np.round(rng.normal(900.0, 700.0, size=(300, 280, 60))).clip(0, 4000)
    return raw, dict(slope=1.0, intercept=-1024.0, xy=1.3, z=2.0)


def _digest(t):
    """shape, statistics and 512 sampled voxels (fixed positions): robust against last-bit differences of F.interpolate between
    CPU instruction sets, which a hash of all 55 M voxels would not be."""
    import numpy as np
    a = t.detach().cpu().contiguous().numpy().astype(np.float64)
    pos = np.random.default_rng(5).integers(0, a.size, size=512)
    return dict(shape=list(a.shape), mean=float(a.mean()), std=float(a.std()), min=float(a.min()), max=float(a.max()),
                frac_padding=float((a == -1.0).mean()), sample_positions_seed=5, samples=[float(v) for v in a.reshape(-1)[pos]])


def _digest_close(d, want):
    assert d["shape"] == want["shape"]
    for k in ("mean", "std", "frac_padding"):
        assert abs(d[k] - want[k]) < 1e-6, (k, d[k], want[k])
    assert d["min"] == want["min"] and d["max"] == want["max"]
    assert max(abs(x - y) for x, y in zip(d["samples"], want["samples"])) < 1e-6


@pytest.mark.skipif(not ref_shims.reference_available(), reason="reference checkout not mounted (GPU box)")
def test_ct_preprocess_matches_reference_dataset_live():
    """oracle.ct_preprocess against the UNMODIFIED CTReportDataset.nii_img_to_tensor (scripts/data.py:92-162),
    bit for bit, with nibabel.load stubbed by an in-memory array and a one-row metadata frame."""
    import pandas as pd
    mod = ref_shims.load_reference_dataset_class()
    raw, m = _preprocess_case()

    class _Img:
        def get_fdata(self):
            return raw.astype("float64")
    mod.nib.load = lambda path: _Img()
    ds = object.__new__(mod.CTReportDataset)
    df = pd.DataFrame({"VolumeName": ["case.nii.gz"], "RescaleSlope": [m["slope"]], "RescaleIntercept": [m["intercept"]],
                       "XYSpacing": [f"[{m['xy']}, {m['xy']}]"], "ZSpacing": [m["z"]]})
    ref = ds.nii_img_to_tensor("/data/case.nii.gz", df)
    got = O.ct_preprocess(raw, m["slope"], m["intercept"], m["xy"], m["z"])
    assert tuple(ref.shape) == tuple(got.shape) == (1, 240, 480, 480)
    assert torch.equal(ref, got)
    # the committed digest (tests/golden/preprocess_digest.json) was produced by such a reference run
    import json
    from pathlib import Path
    _digest_close(_digest(ref), json.loads((Path(__file__).parent / "golden" / "preprocess_digest.json").read_text()))


def test_ct_preprocess_matches_committed_reference_digest():
    """Same case without the reference checkout (GPU box): the oracle must reproduce the digest of the reference's output."""
    import json
    from pathlib import Path
    raw, m = _preprocess_case()
    got = O.ct_preprocess(raw, m["slope"], m["intercept"], m["xy"], m["z"])
    _digest_close(_digest(got), json.loads((Path(__file__).parent / "golden" / "preprocess_digest.json").read_text()))
