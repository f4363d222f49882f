"""GPU bring-up of csrc/attention_tc.cu: forward and backward at several geometries, each in its own process (a trap in one
kernel must not hide the others), with error magnitudes printed instead of asserted."""
import subprocess
import sys


def one(nseq, H, W, do_bwd):
    import torch
    import torch.nn.functional as F
    sys.path.insert(0, ".")
    from ct_clip_b200 import ops
    from tests.test_attention_tc_gpu import _inputs, _rel_index
    from tests.helpers import rel_err, rms_err
    heads, dh = 8, 32
    n, I = H * W, heads * dh
    M = nseq * n
    q, k, kv, tab, d_o, qs, ks = _inputs(nseq, H, W)
    v = kv[:, I:]
    qkb = torch.empty(1, device="cuda")
    ops.qk_bound(qs, ks, qkb)
    geom = dict(n=n, heads=heads, num_seqs=nseq, seq_inner=1, seq_outer_stride=n, tok_stride=1)
    o = torch.zeros(M, I, dtype=torch.bfloat16, device="cuda")
    lse = torch.zeros(M, heads, device="cuda")
    ops.attn_fwd(q, k, v, o, lse, ldq=I, ldk=I, ldv=2 * I, ldo=I, cpb_table=tab, grid_hw=(H, W), qk_bound=qkb, **geom)
    torch.cuda.synchronize()
    to_seq = lambda x: x.float().view(nseq, n, heads, dh).permute(0, 2, 1, 3)
    from_seq = lambda x: x.permute(0, 2, 1, 3).reshape(M, I)
    rel = _rel_index(H, W).cuda()
    tabr = tab.clone().requires_grad_(True)
    bias = tabr[rel].permute(2, 0, 1)
    qr, kr, vr = (to_seq(t).requires_grad_(True) for t in (q, k, v))
    sim = qr @ kr.transpose(-1, -2) * 8.0 + bias
    oref = sim.softmax(-1) @ vr
    lse_ref = (torch.logsumexp(sim, dim=-1) * 1.4426950408889634).permute(0, 2, 1).reshape(M, heads)
    print(f"  fwd n={n} nseq={nseq}: o rel_err {rel_err(o, from_seq(oref)):.3e}  lse abs err {(lse - lse_ref).abs().max().item():.3e}", flush=True)
    if not do_bwd:
        return
    oref.backward(to_seq(d_o))
    # backward with the REFERENCE o / lse so that a forward bug does not mask the backward result
    o_in = from_seq(oref.detach()).to(torch.bfloat16).contiguous()
    lse_in = lse_ref.detach().contiguous()
    dq = torch.zeros(M, I, dtype=torch.bfloat16, device="cuda")
    dkv = torch.zeros(M, 2 * I, dtype=torch.bfloat16, device="cuda")
    delta = torch.empty(M, heads, device="cuda")
    dtab = torch.zeros_like(tab)
    scratch = torch.zeros(nseq * heads * n * n, dtype=torch.bfloat16, device="cuda")
    ops.attn_bwd(q, k, v, o_in, lse_in, d_o, delta, dq, dkv, dkv[:, I:], ldq=I, ldk=I, ldv=2 * I, ldo=I, ld_dq=I, ld_dk=2 * I,
                 ld_dv=2 * I, total_rows=M, cpb_table=tab, grid_hw=(H, W), dcpb_table=dtab, ds_scratch=scratch, **geom)
    torch.cuda.synchronize()
    print(f"  bwd n={n}: dq {rel_err(dq, from_seq(qr.grad)):.3e}  dk {rel_err(dkv[:, :I], from_seq(kr.grad)):.3e}  "
          f"dv {rel_err(dkv[:, I:], from_seq(vr.grad)):.3e}  dtab rms {rms_err(dtab, tabr.grad):.3e}", flush=True)
    # spill check: dS^T[item][j][i] against autograd's d sim
    # (d sim = P o (dP - delta)): recompute in fp32
    P = sim.detach().softmax(-1)
    dP = to_seq(d_o) @ vr.detach().transpose(-1, -2)
    dS = P * (dP - (dP * P).sum(-1, keepdim=True))                         # [nseq, heads, n(i), n(j)]
    sp = scratch.view(nseq, heads, n, n).float()                            # [.., j, i]
    print(f"  bwd spill rel_err {rel_err(sp, dS.transpose(-1, -2)):.3e}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4] == "1")
        sys.exit(0)
    for cfg in [(1, 4, 24, 0), (2, 8, 24, 0), (2, 24, 24, 0), (1, 32, 32, 0), (2, 8, 24, 1), (2, 24, 24, 1), (20, 24, 24, 1)]:
        print(f"== attn_tc nseq={cfg[0]} H={cfg[1]} W={cfg[2]} bwd={cfg[3]}", flush=True)
        r = subprocess.run(["timeout", "120", sys.executable, __file__] + [str(c) for c in cfg], capture_output=True, text=True)
        print(r.stdout[-3000:], flush=True)
        if r.returncode != 0:
            print(f"  rc={r.returncode}\n" + r.stderr[-2500:], flush=True)
