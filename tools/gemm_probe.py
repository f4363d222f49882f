#!/usr/bin/env python
"""Time the GEMM shapes of one contrastive step in isolation (CUDA events, L2 flushed between launches) under the
debug knobs of csrc/gemm_tcgen05.cu. Usage on the GPU box:  python tools/gemm_probe.py [--variants]
Each variant runs in its own process because the knobs are read once per process."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

M = 110592
# (M, N, K, a_major, b_major, epilogue, splits, label)
SHAPES = [
    (M, 2816, 512, 0, 0, 3, 1, "FF1 fwd GEGLU"),
    (M, 1408, 512, 0, 1, 0, 1, "FF2 dgrad (dg) BF16"),
    (M, 512, 2816, 0, 1, 0, 1, "FF1 dgrad BF16"),
    (M, 512, 1408, 0, 0, 2, 1, "FF2 fwd RESID"),
    (M, 512, 512, 0, 1, 2, 1, "kv dgrad RESID"),
    (M, 512, 256, 0, 0, 2, 1, "out-proj fwd RESID"),
    (M, 512, 512, 0, 0, 6, 1, "kv-proj fwd L2NORM"),
    (M, 256, 512, 0, 0, 6, 1, "q-proj fwd L2NORM"),
    (2816, 512, M, 1, 1, 4, 0, "FF1 wgrad"),
    (512, 1365, M, 1, 1, 4, 0, "FF2 wgrad"),
    (M, 512, 4000, 0, 0, 1, 1, "patch fwd F32"),
    (512, 512, M, 1, 1, 4, 0, "kv wgrad"),
    (512, 4000, M, 1, 1, 4, 0, "patch wgrad"),
]


def run_one():
    import torch
    from ct_clip_b200 import ops
    dev = torch.device("cuda", 0)
    bf = dict(dtype=torch.bfloat16, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = []
    for (m, n, k, am, bm, epi, splits, label) in SHAPES:
        A = torch.randn((m, k) if am == 0 else (k, m), **bf) * 0.05
        npad = (n + 7) // 8 * 8
        B = torch.randn((n, k) if bm == 0 else (k, npad), **bf) * 0.05
        kw = dict(M=m, N=n, K=k, a_major=am, b_major=bm, epilogue=epi, splits=splits)
        if epi in (0,):
            kw["C_out"] = torch.empty(m, npad, **bf)
        elif epi in (1, 2, 4):
            kw["C_out"] = torch.zeros(m, n, device=dev)
            if epi == 2:
                kw["resid"] = kw["C_out"]
            if epi == 1:
                kw["bias"] = torch.zeros(n, device=dev)
        elif epi == 3:
            kw["C_out"] = torch.empty(m, n, **bf)
            kw["C2"] = torch.empty(m, n // 2, **bf)
            kw["bias"] = torch.zeros(n, device=dev)
        elif epi == 6:
            kw["C_out"] = torch.empty(m, n, **bf)
            kw["C2"] = torch.empty(m, min(n, 256), **bf)
            kw["norm_cols"] = min(n, 256)
            kw["norm_scale"] = torch.ones(32, device=dev)
            kw["bias"] = torch.zeros(n, device=dev) if n == 256 else None
        for _ in range(2):
            ops.gemm(A, B, **kw)
        ts = []
        for _ in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.gemm(A, B, **kw)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        t = ts[len(ts) // 2]
        out.append((label, f"{m}x{n}x{k}", round(t * 1e3, 1), round(2.0 * m * n * k / (t * 1e-3) / 1e12, 1)))
        del A, B, kw
    print(json.dumps(out))


VARIANTS = [
    ("default (v2 epilogue)", {}),
    ("v1 epilogue", {"CTCLIP_GEMM_OLD_EPI": "1"}),
    ("no epilogue stores (mainloop ceiling)", {"CTCLIP_GEMM_EPI_NONE": "1"}),
    ("BN=128 forced", {"CTCLIP_GEMM_BN": "128"}),
]
if "--quick" in sys.argv:
    VARIANTS = [VARIANTS[0], VARIANTS[2]]

if __name__ == "__main__":
    if "--one" in sys.argv:
        run_one()
        sys.exit(0)
    res = {}
    for name, env in VARIANTS:
        e = dict(os.environ, **env)
        r = subprocess.run([sys.executable, __file__, "--one"], env=e, capture_output=True, text=True, timeout=600)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("[")]
        res[name] = json.loads(line[-1]) if line else ("FAILED: " + r.stderr[-400:])
    names = list(res)
    print("| shape | " + " | ".join(names) + " |")
    print("|---|" + "---|" * len(names))
    for i, sh in enumerate(SHAPES):
        cells = []
        for nme in names:
            v = res[nme]
            cells.append(f"{v[i][2]} us / {v[i][3]} TF/s" if isinstance(v, list) else "fail")
        print(f"| {sh[7]} {sh[0]}x{sh[1]}x{sh[2]} | " + " | ".join(cells) + " |")
    for nme in names:
        if not isinstance(res[nme], list):
            print(nme, res[nme])
