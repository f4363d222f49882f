"""CPU suite part 3: host-side logic -- drop-in surface (import names, constructor kwargs, state-dict layout),
parameter arena layout, synthetic data contract, and the data-parallel math on 2 gloo ranks."""
import os
import socket
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ctclip_oracle as O

GOLD = Path(__file__).resolve().parent / "golden"


def test_dropin_import_names_and_constructor_surface():
    from ct_clip import CTCLIP            # scripts/run_train.py:3
    from transformer_maskgit import CTViT  # scripts/run_train.py:1
    from transformers import BertConfig, BertModel
    vit = CTViT(dim=512, codebook_size=8192, image_size=480, patch_size=20, temporal_patch_size=10, spatial_depth=1,
                temporal_depth=1, dim_head=32, heads=8)                               # run_train.py:17-27 keywords
    assert vit.patch_height_width == (24, 24) and vit.image_num_tokens == 576
    bert = BertModel(BertConfig(num_hidden_layers=1))
    clip = CTCLIP(image_encoder=vit, text_encoder=bert, dim_text=768, dim_image=294912, dim_latent=512,
                  extra_latent_projection=False, use_mlm=False, downsample_image_embeds=False, use_all_token_embeds=False)  # :31-42
    sd = clip.state_dict()
    assert sd["to_visual_latent.weight"].shape == (512, 294912) and sd["temperature"].shape == ()
    assert "to_visual_latent_extra.weight" in sd and "visual_transformer.vq._codebook.embed" in sd
    with pytest.raises(NotImplementedError):
        CTCLIP(image_encoder=vit, text_encoder=bert, use_mlm=True)
    with pytest.raises(NotImplementedError):
        CTCLIP(dim_text=768)
    import sys
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "scripts"))
    from CTCLIPTrainer import CTClipTrainer  # noqa: F401  (run_train.py:4)
    from zero_shot import CTClipInference  # noqa: F401    (run_zero_shot.py:4)


@pytest.mark.parametrize("name", ["cfg1", "scramble", "cfg5_small"])
def test_state_dict_layout_matches_reference_golden(name):
    """Key set and shapes equal the reference's own state_dict (recorded in the golden file)."""
    from transformers import BertConfig, BertModel

    from ct_clip_b200 import CTCLIP, CTViT
    g = torch.load(GOLD / f"{name}.pt", weights_only=False)
    c = g["case"]
    vit = CTViT(**c["vit"])
    bert = BertModel(BertConfig(num_hidden_layers=c["bert_layers"]))
    hw = c["vit"]["image_size"] // c["vit"]["patch_size"]
    clip = CTCLIP(image_encoder=vit, text_encoder=bert, dim_text=768, dim_image=hw * hw * c["vit"]["dim"], dim_latent=512)
    mine = {k: list(v.shape) for k, v in clip.state_dict().items()}
    assert mine == g["shapes"]
    clip.load_state_dict(O.synth_state_dict({k: tuple(v) for k, v in g["shapes"].items()}, 0), strict=True)


def test_no_cpu_fallback():
    """The product path must fail loudly off-GPU instead of silently computing elsewhere."""
    from ct_clip_b200 import CTViT
    vit = CTViT(dim=128, codebook_size=64, image_size=16, patch_size=8, temporal_patch_size=2, spatial_depth=1,
                temporal_depth=1, dim_head=32, heads=4)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            vit(torch.zeros(1, 1, 4, 16, 16), return_encoded_tokens=True)


def test_param_arena_layout():
    from ct_clip_b200.trainer import ParamArena
    ps = [("a", torch.nn.Parameter(torch.randn(3, 5))), ("b", torch.nn.Parameter(torch.randn(()))),
          ("c", torch.nn.Parameter(torch.randn(7)))]
    orig = [p.detach().clone() for _, p in ps]
    ar = ParamArena(ps, torch.device("cpu"))
    assert ar.offsets == [0, 16, 20] and ar.numel == 28
    for (n, p), o in zip(ps, orig):
        assert torch.equal(p.detach(), o)
        assert p.data.untyped_storage().data_ptr() == ar.p.untyped_storage().data_ptr()
        assert p.grad.untyped_storage().data_ptr() == ar.g.untyped_storage().data_ptr()
    ps[0][1].grad.add_(1.0)
    assert ar.g[:15].eq(1).all() and ar.g[15] == 0
    ar.zero_grad()
    assert ar.g.abs().sum() == 0


def test_synthetic_dataset_contract():
    from ct_clip_b200.data import SyntheticCTReportDataset
    ds = SyntheticCTReportDataset(4, frames=8, image=16, n_text=12)
    v, t = ds[1]
    v2, _ = ds[1]
    assert v.shape == (1, 8, 16, 16) and v.dtype == torch.int16 and torch.equal(v, v2)
    assert v.min() >= -1000 and v.max() <= 1000
    assert t["input_ids"][0] == 2 and t["attention_mask"].sum() >= 3
    last = int(t["attention_mask"].sum()) - 1
    assert t["input_ids"][last] == 3 and (t["input_ids"][last + 1:] == 0).all()
    vols, toks = ds.collate([ds[0], ds[1]])
    assert vols.shape == (2, 1, 8, 16, 16) and toks["input_ids"].shape == (2, 12)


# ---------------------------------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, b, L, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ct_clip_b200.dist_utils import gather_latents, rank_rows
    g = torch.Generator().manual_seed(11)
    T, I = torch.randn(world * b, L, generator=g), torch.randn(world * b, L, generator=g)     # the global batch
    W = torch.randn(L, L, generator=g) / L ** 0.5                                               # a shared parameter
    r0, nr = rank_rows(rank, b)
    t_loc = T[r0:r0 + nr].clone().requires_grad_(True)
    Wl = W.clone().requires_grad_(True)
    i_loc = I[r0:r0 + nr] @ Wl
    tg, ig = gather_latents(t_loc.detach(), i_loc.detach())
    # every rank evaluates the GLOBAL loss, differentiating only through its own rows
    tg = torch.cat([tg[:r0], t_loc, tg[r0 + nr:]])
    ig = torch.cat([ig[:r0], i_loc, ig[r0 + nr:]])
    norm = torch.nn.functional.normalize
    temp = torch.tensor(1.3, requires_grad=True)       # the learned temperature (ct_clip.py:553): NOT row-partitioned
    loss = O.clip_loss(norm(tg, dim=-1), norm(ig, dim=-1), temp)
    loss.backward()
    gw = Wl.grad.clone()
    dist.all_reduce(gw, op=dist.ReduceOp.SUM)          # gradients are SUMMED across ranks (not averaged)
    # every rank holds d loss / d temperature of the WHOLE global matrix: CTCLIP._backward_into adds 1/world of it to the arena
    gt = temp.grad.clone() / world
    dist.all_reduce(gt, op=dist.ReduceOp.SUM)
    ret[rank] = (loss.item(), t_loc.grad.clone(), gw, gt.item())
    dist.destroy_process_group()


def _bucket_worker(rank, world, port, ret):
    """early all-reduce of one tensor of the flat gradient arena + the remaining slices == one all-reduce of the arena"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ct_clip_b200.trainer import ParamArena, rest_slices
    g = torch.Generator().manual_seed(5)
    ps = [(n, torch.nn.Parameter(torch.randn(shp, generator=g))) for n, shp in
          (("head", (3, 5)), ("big", (7, 9)), ("temperature", ()), ("tail", (11,)))]
    ar = ParamArena(ps, torch.device("cpu"))
    ar.g.copy_(torch.randn(ar.numel, generator=torch.Generator().manual_seed(100 + rank)))
    whole = ar.g.clone()
    dist.all_reduce(whole, op=dist.ReduceOp.SUM)
    for name in ("big", "head", "tail"):
        flat = ar.g.clone()
        o = ar.offsets[ar.names.index(name)]
        k = dict(ps)[name].numel()
        work = dist.all_reduce(flat[o:o + k], op=dist.ReduceOp.SUM, async_op=True)      # what CTCLIP.dp_early_reduce starts
        for lo, hi in rest_slices(ar.names, ar.offsets, ar.numel, name):
            dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM)
        work.wait()
        # padding elements between tensors may be reduced twice or never: they are zero gradients of no parameter
        for n, p in ps:
            oo = ar.offsets[ar.names.index(n)]
            assert torch.allclose(flat[oo:oo + p.numel()], whole[oo:oo + p.numel()]), (name, n)
    ret[rank] = True
    dist.destroy_process_group()


def _bucketer_worker(rank, world, port, ret):
    """GradBucketer: layers announced in reverse order + early tensor + complement == ONE all-reduce of the whole arena"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ct_clip_b200.trainer import GradBucketer, ParamArena
    g = torch.Generator().manual_seed(5)
    shapes = [("head.weight", (6, 10)), ("big.weight", (40, 50)), ("temperature", ())]
    for i in range(4):
        shapes += [(f"tower.layers.{i}.w1", (16, 16)), (f"tower.layers.{i}.b1", (16,)), (f"tower.layers.{i}.w2", (16, 3, 3))]
    shapes += [("tower.norm.gamma", (16,)), ("text.emb", (30, 8))]
    ps = [(n, torch.nn.Parameter(torch.randn(shp, generator=g))) for n, shp in shapes]
    ar = ParamArena(ps, torch.device("cpu"))
    ar.g.copy_(torch.randn(ar.numel, generator=torch.Generator().manual_seed(100 + rank)))
    whole = ar.g.clone()
    dist.all_reduce(whole, op=dist.ReduceOp.SUM)
    bk = GradBucketer(ar, bucket_bytes=2 * (16 * 16 + 16 * 9) * 4)       # two layers per bucket
    bk.tensor_ready("big.weight")
    bk.prefix_ready("text.")
    for i in reversed(range(4)):
        bk.prefix_ready(f"tower.layers.{i}.")
    launched_before_finish = bk.launches
    bk.finish()
    ok = True
    for n, p in ps:      # padding between tensors is no parameter's gradient: compare tensor by tensor
        o = ar.offsets[ar.names.index(n)]
        ok = ok and torch.allclose(ar.g[o:o + p.numel()], whole[o:o + p.numel()])
    ret[rank] = (ok, launched_before_finish, bk.launches)
    dist.destroy_process_group()


def test_grad_bucketer_matches_single_all_reduce_two_ranks_gloo():
    from ct_clip_b200.trainer import complement_ranges, merge_ranges
    assert merge_ranges([(8, 12), (0, 4), (4, 8), (20, 24)]) == [(0, 12), (20, 24)]
    assert complement_ranges([(0, 12), (20, 24)], 30) == [(12, 20), (24, 30)]
    assert complement_ranges([], 8) == [(0, 8)]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_bucketer_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    for r in (0, 1):
        ok, before, total = ret[r]
        assert ok
        assert before >= 3          # early tensor + two 2-layer buckets were on the wire before finish()
        assert total <= before + 4


def test_bucketed_gradient_all_reduce_two_ranks_gloo():
    from ct_clip_b200.trainer import rest_slices
    assert rest_slices(["a", "b", "c"], [0, 16, 20], 28, "a") == [(16, 28)]
    assert rest_slices(["a", "b", "c"], [0, 16, 20], 28, "b") == [(0, 16), (20, 28)]
    assert rest_slices(["a", "b", "c"], [0, 16, 20], 28, "c") == [(0, 20)]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_bucket_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1)


def test_data_parallel_global_loss_two_ranks_gloo():
    world, b, L = 2, 3, 32
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_dp_worker, args=(world, _free_port(), b, L, ret), nprocs=world, join=True)
    # single-process reference at the global batch size
    g = torch.Generator().manual_seed(11)
    T, I = torch.randn(world * b, L, generator=g), torch.randn(world * b, L, generator=g)
    W = torch.randn(L, L, generator=g) / L ** 0.5
    T.requires_grad_(True)
    W.requires_grad_(True)
    norm = torch.nn.functional.normalize
    temp = torch.tensor(1.3, requires_grad=True)
    loss = O.clip_loss(norm(T, dim=-1), norm(I @ W, dim=-1), temp)
    loss.backward()
    for r in range(world):
        l_r, dt_r, gw_r, gt_r = ret[r]
        assert abs(l_r - loss.item()) < 1e-6
        assert torch.allclose(dt_r, T.grad[r * b:(r + 1) * b], atol=1e-6)
        assert torch.allclose(gw_r, W.grad, atol=1e-5)
        assert abs(gt_r - temp.grad.item()) < 1e-6      # the un-partitioned scalar must not be counted world times


# ---------------------------------------------------------------------------------------------------------------
def test_tolerant_state_dict_and_async_writer(tmp_path):
    """checkpoints written by other wrappers / library versions load strictly after normalisation;
    the asynchronous writer produces a loadable file (CPU tensors here; the CUDA path adds pinned buffers + a side stream)."""
    from transformers import BertConfig, BertModel

    from ct_clip_b200 import CTCLIP, CTViT
    from ct_clip_b200.checkpoint import AsyncCheckpointWriter, tolerant_state_dict
    vit = CTViT(dim=128, codebook_size=64, image_size=16, patch_size=8, temporal_patch_size=2, spatial_depth=1, temporal_depth=1,
                dim_head=32, heads=4)
    bert = BertModel(BertConfig(num_hidden_layers=1, hidden_size=64, num_attention_heads=2, intermediate_size=128))
    clip = CTCLIP(image_encoder=vit, text_encoder=bert, dim_text=64, dim_image=4 * 128, dim_latent=32)
    ref = {k: v.clone() for k, v in clip.state_dict().items()}
    # a checkpoint as accelerate(unwrap=False) + transformers 4.30 + a GenerateCT-era CTViT would have written it
    old = {"module." + k: v for k, v in ref.items() if not k.startswith(("to_text_latent_extra", "to_visual_latent_extra"))}
    old["module.text_transformer.embeddings.position_ids"] = torch.arange(512)[None]
    old["module.visual_transformer.discr.layers.0.weight"] = torch.zeros(3, 3)
    clean, report = tolerant_state_dict(old, clip.state_dict())
    assert set(clean) == set(ref) and all(torch.equal(clean[k], ref[k]) for k in ref if not k.endswith("_extra.weight"))
    assert "stripped 'module.'" in report["renamed"] and len(report["dropped"]) == 2 and len(report["filled"]) == 2
    clip.load_state_dict(clean, strict=True)
    # a genuinely different architecture still fails loudly
    bad = dict(ref)
    bad["visual_transformer.to_patch_emb.2.weight"] = torch.zeros(7, 7)
    with pytest.raises(RuntimeError):
        clip.load_state_dict(tolerant_state_dict(bad, clip.state_dict())[0], strict=True)
    # asynchronous writer + CTCLIP.load round trip
    w = AsyncCheckpointWriter()
    w.save(clip.state_dict(), tmp_path / "a.pt")
    w.save({k: v + 1 if v.is_floating_point() else v for k, v in clip.state_dict().items()}, tmp_path / "b.pt")   # joins the first
    w.wait()
    a, b = torch.load(tmp_path / "a.pt"), torch.load(tmp_path / "b.pt")
    assert all(torch.equal(a[k], ref[k]) for k in ref)
    k0 = "to_text_latent.weight"
    assert torch.equal(b[k0], ref[k0] + 1) and not (tmp_path / "b.pt.tmp").exists()
    torch.save({"module." + k: v for k, v in ref.items()}, tmp_path / "wrapped.pt")
    clip.load(tmp_path / "wrapped.pt")
    assert clip.last_load_report["renamed"] == ["stripped 'module.'"]


def test_bench_stage_table_rates_each_stage_against_its_roofline():
    """bench.stage_table: HBM stages vs the copy bandwidth, tensor stages vs the sustained GEMM rate, GEMMs against whichever
    of the two bounds the launch (fake events: no GPU needed)."""
    import bench

    class Ev:
        def __init__(self, t):
            self.t = t

        def elapsed_time(self, other):
            return other.t - self.t
    peaks = dict(hbm=6000.0, tf_sus=1500.0)
    M = 110592
    rec = [("ctclip_ln_fwd", "D512", ("B", 6.0e8), Ev(0.0), Ev(0.2)),                       # 600 MB in 0.2 ms = 3000 GB/s
           ("ctclip_gemm_bf16", "big-K", ("FB", 2.0 * M * 512 * 2816, 1.0e9), Ev(0.0), Ev(0.25)),
           ("ctclip_gemm_bf16", "resid", ("FB", 2.0 * M * 512 * 256, 5.1e8), Ev(0.0), Ev(0.1)),
           ("ctclip_cpb_expand", None, None, Ev(0.0), Ev(0.01))]
    rows = {r[0]: r for r in bench.stage_table(rec, peaks)}
    assert rows["ln_fwd [D512]"][3] == "hbm" and abs(rows["ln_fwd [D512]"][6] - 0.5) < 1e-6
    assert rows["gemm_bf16 [big-K]"][3] == "tensor" and rows["gemm_bf16 [big-K]"][5] == "TFLOP/s"
    assert rows["gemm_bf16 [resid]"][3] == "hbm" and abs(rows["gemm_bf16 [resid]"][4] - 5100.0) < 1e-6
    assert rows["cpb_expand"][3] == "latency"
