#!/usr/bin/env python
"""Time ctclip_peg_fwd at configs[1] size under the probing knobs of csrc/peg_mma.cu (CTCLIP_PEG_DEBUG bits:
1 = no plane loads, 2 = no MMAs/stores, 4 = no output stores, 8 = no conversion) and for the fp32 stencil path."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
from ct_clip_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
B, T, H, W, D = 8, 24, 24, 24, 512
x = torch.randn(B * T * H * W, D, device=dev)
y = torch.empty_like(x)
w = 0.1 * torch.randn(D, 27, device=dev)
bias = torch.zeros(D, device=dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def run(label, dbg, exact=False, temporal=False):
    os.environ["CTCLIP_PEG_DEBUG"] = str(dbg)
    kw = dict(B=B, T=T, H=H, W=W, D=D, temporal=temporal, mma=not exact)
    for _ in range(2):
        ops.peg_fwd(x, y, w, bias, **kw)
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.peg_fwd(x, y, w, bias, **kw)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"{label:55s} {ts[len(ts) // 2] * 1e3:8.1f} us", flush=True)


run("mma full", 0)
run("mma, no output stores (4)", 4)
run("mma, no MMAs / stores (2)", 2)
run("mma, no plane loads in steady state (1)", 1)
run("mma, no loads, no compute (3)", 3)
run("mma, no loads, no compute, no convert (11)", 11)
run("mma, no convert (8)", 8)
run("fp32 stencil (exact)", 0, exact=True)
run("mma full, temporal addressing (identity table absent)", 0, temporal=True)
