"""GPU top-k retrieval + the two evaluation metrics against literal restatements of the reference scripts'
nested loops (scripts/report_to_volume_new.py:6-17, 47-63; scripts/volume_to_volume_new.py:9-31, 57-99)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _find_top_k_indices(values, k):            # report_to_volume_new.py:6-17
    s = sorted(enumerate(values), key=lambda x: x[1], reverse=True)
    return [i for i, _ in s[:k]]


def _calc_similarity(a1, a2):                  # volume_to_volume_new.py:16-31
    oneandone = oneorzero = 0
    for k in range(len(a1)):
        if a1[k] == 1 and a2[k] == 1:
            oneandone += 1
        if a1[k] != a2[k]:
            oneorzero += 1
    return oneandone / (oneandone + oneorzero)


def test_topk_matches_python_sorted_including_ties():
    from ct_clip_b200.retrieval import topk
    g = torch.Generator().manual_seed(0)
    q = torch.randn(37, 64, generator=g)
    gal = torch.randn(501, 64, generator=g)
    gal[100] = gal[7]                     # exact ties: Python's stable sort keeps the lower index first
    gal[300] = gal[7]
    idx, val = topk(q.cuda(), gal.cuda(), 25)
    scores = (q.double() @ gal.double().t()).float()
    for i in range(q.shape[0]):
        got = idx[i].cpu().tolist()
        # fp32 GEMM order vs float64: compare through the scores (ties / near-ties may swap neighbours)
        ref = _find_top_k_indices((q[i] @ gal.t()).tolist(), 25)
        assert torch.allclose(scores[i, got], scores[i, ref], atol=1e-4)
        assert sorted(val[i].cpu().tolist(), reverse=True) == val[i].cpu().tolist()
    # the tied triple must come out in index order
    s7 = (q @ gal[7]).argmax().item()
    row = idx[s7].cpu().tolist()
    if 7 in row and 100 in row and 300 in row:
        assert row.index(7) < row.index(100) < row.index(300)


def test_retrieval_metrics_match_reference_scripts():
    from ct_clip_b200.retrieval import report_to_volume_recall, volume_to_volume_overlap
    g = torch.Generator().manual_seed(1)
    n, L, C = 60, 32, 18
    img = torch.nn.functional.normalize(torch.randn(n, L, generator=g), dim=-1)
    txt = torch.nn.functional.normalize(img + 0.8 * torch.randn(n, L, generator=g), dim=-1)
    labels = (torch.rand(n, C, generator=g) < 0.25).long()
    labels[3] = 0                          # a volume without findings: excluded from the gallery, still a query
    # ---- report -> volume (report_to_volume_new.py:47-66)
    ks = (5, 10, 50)
    ref = {}
    for k in ks:
        hit = 0
        for i in range(n):
            crosses = [float(txt[i] @ img[j]) for j in range(n)]
            if i in _find_top_k_indices(crosses, k):
                hit += 1
        ref[k] = hit / n
    got = report_to_volume_recall(txt.cuda(), img.cuda(), ks)
    assert all(abs(got[k] - ref[k]) < 1e-6 for k in ks), (got, ref)
    # ---- volume -> volume (volume_to_volume_new.py:57-99, incl. the never-reset running mean)
    lab = labels.numpy()
    second = [j for j in range(n) if lab[j].sum() != 0]
    ks2 = (1, 5, 10)
    ratios_external, ref2 = [], {}
    x = img.numpy().astype(np.float64)
    for k in ks2:
        for i in range(n):
            crosses = [float(np.dot(x[i], x[j]) / (np.linalg.norm(x[i]) * np.linalg.norm(x[j]))) for j in second]
            top = _find_top_k_indices(crosses, k)
            rr = []
            for t in top:
                a, b = lab[i], lab[second[t]]
                inter = int(((a == 1) & (b == 1)).sum())
                diff = int((a != b).sum())
                rr.append(inter / (inter + diff) if inter + diff > 0 else float("nan"))
            ratios_external.append(np.mean(np.array(rr)))
        ref2[k] = float(np.mean(np.array(ratios_external)))
    got2 = volume_to_volume_overlap(img.cuda(), labels, ks2)
    for k in ks2:
        if np.isnan(ref2[k]):
            assert np.isnan(got2[k])
        else:
            assert abs(got2[k] - ref2[k]) < 1e-5, (k, got2, ref2)
