#!/usr/bin/env python
"""Time the three PEG entry points at configs[1] size (8 volumes, 24x24x24 token grid, D = 512) for both kernel families:
variant 0 = plane-streaming kernels (csrc/peg_stream.cu, the default), variant 1 = general kernels (csrc/peg.cu), on the
spatial and the temporal stack, with an L2 flush before every timed launch; also checks that the two families agree."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
from ct_clip_b200 import _lib, ops  # noqa: E402

QUICK = "--quick" in sys.argv      # one launch of each plane-streaming kernel, spatial stack only (for ncu)
dev = torch.device("cuda", 0)
B, T, H, W, D = 8, 24, 24, 24, 512
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(B * T * H * W, D, device=dev, generator=g)
dy = torch.randn(B * T * H * W, D, device=dev, generator=g)
w = 0.1 * torch.randn(D, 27, device=dev, generator=g)
bias = 0.1 * torch.randn(D, device=dev, generator=g)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
HBM = 6.57e12   # MEASURED_PEAKS.json copy bandwidth is read by bench.py; here only for a rough fraction


def canon_table():
    f = torch.arange(T * H * W)
    return (((f % T) * H + f // (T * W)) * W + (f // T) % W).to(torch.int32).to(dev)


def timed(fn, reps=7):
    if QUICK:
        fn()
        torch.cuda.synchronize()
        return 1.0
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


results = {}
for temporal in ((False,) if QUICK else (False, True)):
    kw = dict(B=B, T=T, H=H, W=W, D=D, temporal=temporal)
    if temporal:
        kw["canon_table"] = canon_table()
    outs = {}
    for variant in ((0,) if QUICK else (1, 0)):
        _lib.check(_lib.lib().ctclip_debug_set_peg_variant(variant), "variant")
        y = torch.empty_like(x)
        dx = torch.empty_like(x)
        dxb = torch.empty(x.shape, dtype=torch.bfloat16, device=dev)
        dw, db = torch.zeros(D, 27, device=dev), torch.zeros(D, device=dev)
        t_f = timed(lambda: ops.peg_fwd(x, y, w, bias, **kw))
        t_d = timed(lambda: ops.peg_bwd_data(dy, dx, w, dx_bf16=dxb, **kw))
        t_w = timed(lambda: ops.peg_bwd_weight(x, dy, dw, db, **kw))
        dw.zero_(); db.zero_()
        ops.peg_bwd_weight(x, dy, dw, db, **kw)
        torch.cuda.synchronize()
        outs[variant] = (y.clone(), dx.clone(), dw.clone(), db.clone())
        name = ("temporal" if temporal else "spatial") + (" stream" if variant == 0 else " general")
        nb = x.numel()
        print(f"{name:18s} fwd {t_f * 1e6:7.1f} us ({8 * nb / t_f / 1e9:6.0f} GB/s)   bwd_data {t_d * 1e6:7.1f} us "
              f"({10 * nb / t_d / 1e9:6.0f} GB/s)   bwd_weight {t_w * 1e6:7.1f} us ({8 * nb / t_w / 1e9:6.0f} GB/s)", flush=True)
    for i, nm in enumerate(() if QUICK else ("y", "dx", "dweight", "dbias")):
        a, b_ = outs[0][i], outs[1][i]
        err = ((a - b_).abs().max() / b_.abs().max()).item()
        print(f"   stream vs general {nm:8s} max rel diff {err:.2e}")
_lib.check(_lib.lib().ctclip_debug_set_peg_variant(0), "variant")
