"""Latent dump + GPU retrieval.

* `dump_latents`: the batched version of scripts/forward_data.py:114-150 (one `return_latents` forward per BATCH instead of per
  volume; same files: `<folder>/text/<accession>.npz` and `<folder>/image/<accession>.npz`, key `arr`, shape (1, dim_latent)).
* `topk`: scores = fp32 GEMM on the device + per-row top-k (csrc/retrieval.cu) instead of the nested Python loops and
  `sorted(enumerate(...))` of scripts/report_to_volume_new.py:47-63 / scripts/volume_to_volume_new.py:80-96.
* `report_to_volume_recall` / `volume_to_volume_overlap`: the two metrics those scripts print, on top of `topk`.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from . import ops
from ._lib import call


def _stream():
    return torch.cuda.current_stream().cuda_stream


def topk(queries: torch.Tensor, gallery: torch.Tensor, k: int, *, cosine: bool = False):
    """queries [Q, L], gallery [G, L] fp32 CUDA -> (indices int32 [Q, k], scores fp32 [Q, k]), best first.
    cosine=True normalises both sides first (volume_to_volume_new.py:86-90); False is the plain dot product of
    report_to_volume_new.py:57 (the saved latents are already unit vectors)."""
    assert queries.is_cuda and gallery.is_cuda and queries.dtype == gallery.dtype == torch.float32
    Q, L = queries.shape
    G = gallery.shape[0]
    assert gallery.shape[1] == L and 0 < k <= G
    q, g = queries.contiguous(), gallery.contiguous()
    if cosine:
        qn, gn = torch.empty_like(q), torch.empty_like(g)
        call("ctclip_l2norm_rows_f32", q.data_ptr(), qn.data_ptr(), Q, L, _stream())
        call("ctclip_l2norm_rows_f32", g.data_ptr(), gn.data_ptr(), G, L, _stream())
        q, g = qn, gn
    scores = torch.empty(Q, G, device=q.device)
    ops.sgemm(q, g, scores, M=Q, N=G, K=L, trans_b=True)
    idx = torch.empty(Q, k, dtype=torch.int32, device=q.device)
    val = torch.empty(Q, k, device=q.device)
    call("ctclip_topk_rows", scores.data_ptr(), G, Q, G, k, idx.data_ptr(), val.data_ptr(), _stream())
    return idx, val


def report_to_volume_recall(text_latents: torch.Tensor, image_latents: torch.Tensor, ks=(5, 10, 50, 100)):
    """recall@k of report -> volume retrieval: fraction of reports whose own volume is among the k best
    (report_to_volume_new.py:51-66: `if i in top_k_indices`)."""
    n = text_latents.shape[0]
    idx, _ = topk(text_latents, image_latents, max(ks))
    own = torch.arange(n, device=idx.device, dtype=torch.int32)[:, None]
    hit = idx == own                                   # [n, kmax] (tiny boolean bookkeeping on the result)
    return {int(k): hit[:, :k].any(dim=1).float().mean().item() for k in ks}


def volume_to_volume_overlap(image_latents: torch.Tensor, labels: torch.Tensor, ks=(1, 5, 10, 50)):
    """volume_to_volume_new.py:57-99: every volume queries the gallery of volumes with at least one positive label (cosine
    similarity); the score of a hit is |a AND b| / (|a AND b| + |a XOR b|) of the binary label vectors (`calc_similarity`);
    returns, per k, the RUNNING mean over all k so far -- the script never resets `ratios_external` between k values."""
    labels = labels.to(image_latents.device).float()
    keep = labels.sum(dim=1) != 0
    gal, gal_labels = image_latents[keep].contiguous(), labels[keep]
    idx, _ = topk(image_latents, gal, max(ks), cosine=True)
    out, ratios_external = {}, []
    for k in ks:
        sel = gal_labels[idx[:, :k].long()]                      # [n, k, C]
        a = labels[:, None, :]
        both = (a * sel).sum(-1)
        differ = (a != sel).float().sum(-1)
        ratio = both / (both + differ)                           # 0/0 -> nan, as the script's ZeroDivisionError would be fatal there
        ratios_external += ratio.mean(dim=1).tolist()
        out[int(k)] = float(np.mean(np.array(ratios_external)))
    return out


@torch.no_grad()
def dump_latents(clip, batches, results_folder, device="cuda"):
    """batches: iterable of (volumes [b,1,F,H,W], tokens-like with .input_ids/.attention_mask, accession names[b]).
    Writes one npz per accession and side, like forward_data.py:133-146, and returns (text_latents, image_latents, names)."""
    root = Path(results_folder)
    (root / "text").mkdir(parents=True, exist_ok=True)
    (root / "image").mkdir(parents=True, exist_ok=True)
    clip.eval()
    all_t, all_i, names = [], [], []
    for vols, tok, accs in batches:
        tl, il, _ = clip(tok, vols.to(device), device=device, return_latents=True)
        tl_c, il_c = tl.float().cpu().numpy(), il.float().cpu().numpy()
        for j, acc in enumerate(accs):
            np.savez(root / "text" / f"{acc}.npz", arr=tl_c[j:j + 1])
            np.savez(root / "image" / f"{acc}.npz", arr=il_c[j:j + 1])
        all_t.append(tl)
        all_i.append(il)
        names += list(accs)
    return torch.cat(all_t), torch.cat(all_i), names
