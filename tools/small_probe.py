#!/usr/bin/env python
"""A/B of the patchify kernels at configs[1] size on one GPU: v2 vs v3 (CTCLIP_PATCHIFY_V3=0/1). Checks that they agree
and prints median times."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
from ct_clip_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


# ---- patchify
B, Fr, H, W = 8, 240, 480, 480
g = torch.Generator(device="cpu").manual_seed(1)
vol = torch.randint(-1000, 1000, (B, 1, Fr, H, W), generator=g, dtype=torch.int16).to(dev)
M, P = B * 24 * 24 * 24, 4000
outs = {}
for v in ("0", "1"):
    os.environ["CTCLIP_PATCHIFY_V3"] = v
    out = torch.empty(M, P, dtype=torch.bfloat16, device=dev)
    t = timeit(lambda: ops.patchify(vol, out, B=B, Cc=1, F=Fr, H=H, W=W, pt=10, p1=20, p2=20))
    outs[v] = out
    print(f"patchify v3={v}: {t:8.1f} us  ({vol.numel() * 4 / t / 1e3:.0f} GB/s)", flush=True)
d = (outs["0"].float() - outs["1"].float()).abs().max().item()
print("patchify max |v2 - v3| =", d)
assert d < 2e-2
volf = (vol[:1].float() / 1000.0)
o0, o1 = torch.empty(M // B, P, dtype=torch.bfloat16, device=dev), torch.empty(M // B, P, dtype=torch.bfloat16, device=dev)
os.environ["CTCLIP_PATCHIFY_V3"] = "0"
ops.patchify(volf, o0, B=1, Cc=1, F=Fr, H=H, W=W, pt=10, p1=20, p2=20)
os.environ["CTCLIP_PATCHIFY_V3"] = "1"
ops.patchify(volf, o1, B=1, Cc=1, F=Fr, H=H, W=W, pt=10, p1=20, p2=20)
print("patchify fp32 input max diff =", (o0.float() - o1.float()).abs().max().item())
del vol, outs, out

