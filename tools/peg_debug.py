#!/usr/bin/env python
"""Locate disagreements between the plane-streaming and the general PEG kernels (debug aid)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
from ct_clip_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda", 0)
B, T, H, W, D = [int(v) for v in sys.argv[1:5]] + [512]
REPS = int(sys.argv[5]) if len(sys.argv) > 5 else 20
g = torch.Generator(device="cuda").manual_seed(1)
dy = torch.randn(B * T * H * W, D, device=dev, generator=g)
w = 0.2 * torch.randn(D, 27, device=dev, generator=g)
bias = 0.1 * torch.randn(D, device=dev, generator=g)
kw = dict(B=B, T=T, H=H, W=W, D=D, temporal=False)


def run(variant, mode):
    _lib.check(_lib.lib().ctclip_debug_set_peg_variant(variant), "variant")
    out = torch.full_like(dy, float("nan"))
    if mode == 1:
        dxb = torch.empty(dy.shape, dtype=torch.bfloat16, device=dev)
        ops.peg_bwd_data(dy, out, w, dx_bf16=dxb, **kw)
    else:
        ops.peg_fwd(dy, out, w, bias, **kw)
    torch.cuda.synchronize()
    return out.view(B, T, H, W, D)


x5 = dy.view(B, T, H, W, D)
for mode in (1, 0):
    ref = run(1, mode)
    nbad_runs = 0
    for rep in range(REPS):
        got = run(0, mode)
        bad = ((got - ref).abs() > 1e-4 * ref.abs().max()) | torch.isnan(got)
        n = int(bad.sum())
        if not n:
            continue
        nbad_runs += 1
        idx = bad.nonzero()
        desc = []
        for name, col in zip("b t h w".split(), range(4)):
            desc.append(f"{name}={sorted(set(idx[:, col].tolist()))}")
        cs = sorted(set(idx[:, 4].tolist()))
        print(f"mode {mode} rep {rep}: {n} bad; " + " ".join(desc) + f" c={cs[0]}..{cs[-1]}")
        if nbad_runs <= 3:
            # which single tap explains the difference? out[p] = in[p] + sum_k w[k] in[p + off(k)] (mode 0: off = (k0-2,k1-1,k2-1);
            # mode 1: out[p] = in[p] + sum_k w[k] in[p - off(k)])
            for i in idx[:: max(1, len(idx) // 6)][:6].tolist():
                b_, t_, h_, w_, c_ = i
                delta = float(got[b_, t_, h_, w_, c_] - ref[b_, t_, h_, w_, c_])
                best = None
                for k in range(27):
                    k0, k1, k2 = k // 9, (k // 3) % 3, k % 3
                    o = (k0 - 2, k1 - 1, k2 - 1)
                    sgn = 1 if mode == 0 else -1
                    tt, hh, ww = t_ + sgn * o[0], h_ + sgn * o[1], w_ + sgn * o[2]
                    if not (0 <= tt < T and 0 <= hh < H and 0 <= ww < W):
                        continue
                    contrib = float(w[c_, k] * x5[b_, tt, hh, ww, c_])
                    if best is None or abs(delta + contrib) < best[0]:
                        best = (abs(delta + contrib), k, (tt, hh, ww), contrib)
                print(f"     {i}: got-ref {delta:+.5f}; closest 'missing tap': k={best[1]} (k0,k1,k2)={best[1] // 9, (best[1] // 3) % 3, best[1] % 3} "
                      f"input {best[2]} contribution {best[3]:+.5f} residual {best[0]:.2e}")
    print(f"mode {mode}: {nbad_runs} of {REPS} runs with bad elements")
_lib.check(_lib.lib().ctclip_debug_set_peg_variant(0), "variant")
