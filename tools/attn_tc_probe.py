"""GPU box: spatial attention at configs[1] size (192 sequences x 8 heads x 576 tokens, dim_head 32), tcgen05 kernels
(csrc/attention_tc.cu) vs the mma.sync kernels (csrc/attention.cu); CUDA-event timing, L2 flushed between repetitions.
  python tools/attn_tc_probe.py [--reps 10] [--only tc|mma] [--fwd-only]        (ncu: wrap with -k regex:attn_tc)"""
import argparse
import sys

import torch

sys.path.insert(0, ".")
from ct_clip_b200 import ops  # noqa: E402
from tests.test_attention_tc_gpu import _inputs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--nseq", type=int, default=192)
    ap.add_argument("--only", default="both")
    ap.add_argument("--fwd-only", action="store_true")
    ap.add_argument("--bwd-warps", type=int, default=0, help="8 or 16 (default: time both)")
    a = ap.parse_args()
    nseq, H, W, heads, dh = a.nseq, 24, 24, 8, 32
    n, I = H * W, heads * dh
    M = nseq * n
    q, k, kv, tab, d_o, qs, ks = _inputs(nseq, H, W)
    v = kv[:, I:]
    qkb = torch.empty(1, device="cuda")
    ops.qk_bound(qs, ks, qkb)
    geom = dict(n=n, heads=heads, num_seqs=nseq, seq_inner=1, seq_outer_stride=n, tok_stride=1)
    o = torch.empty(M, I, dtype=torch.bfloat16, device="cuda")
    lse = torch.empty(M, heads, device="cuda")
    dq = torch.empty(M, I, dtype=torch.bfloat16, device="cuda")
    dkv = torch.empty(M, 2 * I, dtype=torch.bfloat16, device="cuda")
    delta = torch.empty(M, heads, device="cuda")
    dtab = torch.zeros_like(tab)
    scratch = torch.empty(nseq * heads * n * n, dtype=torch.bfloat16, device="cuda")
    bias = torch.empty(heads, n, n, dtype=torch.bfloat16, device="cuda")
    ops.cpb_expand(tab, heads, H, W, bias, None)
    nfrag = ops.frag_elems(heads, n)
    bf, btf = torch.empty(nfrag, dtype=torch.bfloat16, device="cuda"), torch.empty(nfrag, dtype=torch.bfloat16, device="cuda")
    ops.cpb_expand_frag(tab, heads, H, W, bf, btf)
    dbias = torch.zeros(heads, n, n, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    flops = 4.0 * nseq * heads * n * n * dh

    def timeit(fn):
        ts = []
        for _ in range(a.reps + 2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts = sorted(ts[2:])
        return ts[len(ts) // 2]

    def fwd_tc():
        ops.attn_fwd(q, k, v, o, lse, ldq=I, ldk=I, ldv=2 * I, ldo=I, cpb_table=tab, grid_hw=(H, W), qk_bound=qkb, **geom)

    def fwd_mma():
        ops.attn_fwd(q, k, v, o, lse, ldq=I, ldk=I, ldv=2 * I, ldo=I, bias=bias, bias_frag=bf, bias_t_frag=btf, **geom)

    def bwd_tc(with_tab=True):
        ops.attn_bwd(q, k, v, o, lse, d_o, delta, dq, dkv, dkv[:, I:], ldq=I, ldk=I, ldv=2 * I, ldo=I, ld_dq=I, ld_dk=2 * I, ld_dv=2 * I,
                     total_rows=M, cpb_table=tab, grid_hw=(H, W), dcpb_table=dtab if with_tab else None,
                     ds_scratch=scratch if with_tab else None, **geom)

    def bwd_mma():
        ops.attn_bwd(q, k, v, o, lse, d_o, delta, dq, dkv, dkv[:, I:], ldq=I, ldk=I, ldv=2 * I, ldo=I, ld_dq=I, ld_dk=2 * I, ld_dv=2 * I,
                     total_rows=M, bias=bias, bias_frag=bf, bias_t_frag=btf, dbias=dbias, ds_scratch=scratch, **geom)

    if a.only in ("both", "tc"):
        t = timeit(fwd_tc)
        print(f"attn_fwd tc       : {t:.3f} ms  {flops / t / 1e9:.0f} TFLOP/s")
        if not a.fwd_only:
            from ct_clip_b200 import _lib
            for ew in ([a.bwd_warps] if a.bwd_warps else [8, 16, 108, 116]):
                _lib.check(_lib.lib().ctclip_debug_set_attn_bwd_warps(ew), "set warps")
                t = timeit(bwd_tc)
                print(f"attn_bwd tc +dtab [variant {ew:3d}]: {t:.3f} ms  {2.5 * flops / t / 1e9:.0f} TFLOP/s (incl. delta + table-gradient reduction)")
                dbt = torch.zeros(heads, n, n, device="cuda")
                t = timeit(lambda: ops.attn_bwd(q, k, v, o, lse, d_o, delta, dq, dkv, dkv[:, I:], ldq=I, ldk=I, ldv=2 * I, ldo=I, ld_dq=I,
                                                ld_dk=2 * I, ld_dv=2 * I, total_rows=M, cpb_table=tab, grid_hw=(H, W), dbias=dbt, **geom))
                print(f"attn_bwd tc +red  [variant {ew:3d}]: {t:.3f} ms  {2.5 * flops / t / 1e9:.0f} TFLOP/s (table gradient by fp32 red.add into the L2-resident table)")
                t = timeit(lambda: bwd_tc(False))
                print(f"attn_bwd tc       [variant {ew:3d}]: {t:.3f} ms  {2.5 * flops / t / 1e9:.0f} TFLOP/s (no table gradient / spill)")
    if a.only in ("both", "mma"):
        fwd_mma()
        t = timeit(fwd_mma)
        print(f"attn_fwd mma.sync : {t:.3f} ms  {flops / t / 1e9:.0f} TFLOP/s")
        if not a.fwd_only:
            t = timeit(bwd_mma)
            print(f"attn_bwd mma.sync : {t:.3f} ms  {2.5 * flops / t / 1e9:.0f} TFLOP/s")


if __name__ == "__main__":
    main()
