#!/usr/bin/env python
"""bench.py -- CT volumes/sec through the full contrastive step (BASELINE.json metric).

  python bench.py --gpus 1 --steps K --warmup W                  our arm (B200 kernels)
  torchrun --nproc-per-node N ... bench.py --gpus N ...          data parallel, one rank per GPU
  python bench.py --impl reference ...                           reference arm: the reference algorithm's own CPU
                                                                 PyTorch path (oracle port; the Python reference
                                                                 cannot travel to the GPU box) on the host cores

A "step" = forward + InfoNCE loss + backward + gradient all-reduce + clip(0.5) + Adam on one synthetic batch
(BASELINE.json configs[1]: CTViT dim 512, 12+12 layers, 480x480x240 int16 volumes, patch (20,20,10), 128-token
text, bs 8 per GPU). `value` is timed with inputs resident in HBM; `e2e` adds the pinned-host -> device copy of
every batch and the loss read-back, through the public CTClipTrainer API.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

_OUT_FD = 1
ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("HF_HUB_OFFLINE", "1")
os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--depth", type=int, default=12, help="spatial_depth = temporal_depth (configs[1]: 12; reference scripts: 4)")
    ap.add_argument("--batch", type=int, default=8, help="volumes per GPU")
    ap.add_argument("--dim", type=int, default=512, help="CTViT width (configs[1]: 512; configs[4]: 768 with --image 512 --frames 320 --depth 24)")
    ap.add_argument("--image", type=int, default=480)
    ap.add_argument("--frames", type=int, default=240)
    ap.add_argument("--text-len", type=int, default=128)
    ap.add_argument("--bert-layers", type=int, default=12)
    ap.add_argument("--config", default="train", choices=["train", "zero_shot"],
                    help="train: BASELINE configs[1] contrastive step (the headline); zero_shot: configs[3], 18 x 2 prompt bank x N volumes")
    ap.add_argument("--volumes", type=int, default=1304, help="zero_shot: number of synthetic volumes (reference validation set: 1304)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stages", action="store_true", help="skip the per-stage roofline pass (one extra untimed step)")
    ap.add_argument("--cpu-sample-volumes", type=int, default=2)
    ap.add_argument("--cpu-sample-frames", type=int, default=None,
                    help="crop the CPU arm's volumes to this many frames (default: whole volumes)")
    return ap.parse_args()


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf_burst=d.get("bf16_tflops", 1590.0),
                    tf_sus=d.get("bf16_tflops_sustained", 1400.0), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [c.strip() for c in line.split(",")])

    def window(self, t0, t1):
        """keep only the samples taken inside [t0, t1] (the sampler is started early: nvidia-smi needs ~0.3 s to come up)"""
        inside = [r for r in self.rows if t0 <= r[0] <= t1]
        if inside:
            self.rows = inside

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            r = r[1:]
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=(max(mx) if mx else None), reasons=sorted(reasons),
                    samples=len(sm))


def stage_table(rec, peaks):
    """Aggregate (entry point, tag) -> [stage, launches, ms, bound, achieved, unit, frac] from one instrumented step.
    HBM-bound stages are rated against the measured copy bandwidth, tensor stages against the measured sustained bf16 rate."""
    agg = {}
    for name, tag, work, e0, e1 in rec:
        key = (name.replace("ctclip_", ""), tag or "", work[0] if work else "")
        a = agg.setdefault(key, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += e0.elapsed_time(e1)
        a[2] += work[1] if work else 0.0
        a[3] += work[2] if work and len(work) > 2 else 0.0
    rows = []
    for (name, tag, kind), (n, t_ms, w, w2) in agg.items():
        if kind == "FB":   # GEMMs carry both: rate the launch against whichever roofline bounds it (K = 512 residual
            # epilogues move 8 bytes per output and are HBM-bound, the wide-K shapes are tensor-bound)
            kind = "B" if w2 / (peaks["hbm"] * 1e9) > w / (peaks["tf_sus"] * 1e12) else "F"
            if kind == "B":
                w = w2
        if kind == "B":
            ach, unit, peak, bound = w / (t_ms * 1e-3) / 1e9, "GB/s", peaks["hbm"], "hbm"
        elif kind == "F":
            ach, unit, peak, bound = w / (t_ms * 1e-3) / 1e12, "TFLOP/s", peaks["tf_sus"], "tensor"
        else:
            ach, unit, peak, bound = 0.0, "", 1.0, "latency"
        rows.append([f"{name} [{tag}]" if tag else name, n, t_ms, bound, ach, unit, ach / peak if peak else 0.0, ("ctclip_" + name, tag)])
    rows.sort(key=lambda r: -r[2])
    return rows


def write_stage_table(path, rows, step_ms):
    tot = sum(r[2] for r in rows)
    with open(path, "w") as f:
        f.write(f"# per-stage roofline, one instrumented step (CUDA events around every C-ABI call; sum {tot:.1f} ms, "
                f"un-instrumented step {step_ms:.1f} ms)\n")
        f.write("| stage (entry point [tag]) | launches | ms/step | share | bound | achieved | of measured peak |\n|---|---|---|---|---|---|---|\n")
        for r in rows:
            f.write(f"| {r[0]} | {r[1]} | {r[2]:.3f} | {100 * r[2] / tot:.1f}% | {r[3]} | {r[4]:.1f} {r[5]} | {100 * r[6]:.1f}% |\n")


def flops_per_volume(cfgd):
    """SURVEY 8(d) algorithmic FLOPs of one forward pass per volume (and the x3 step estimate)."""
    D, I, F, P, C, L = cfgd.get("D", 512), 256, int(cfgd.get("D", 512) * 4 * 2 / 3), cfgd["P"], 8192, 512
    N, S, T = cfgd["N"], cfgd["S"], cfgd["T"]
    layer_s = 54 * N * D + 8 * N * D * I + 4 * N * S * I + 6 * N * D * F
    layer_t = 54 * N * D + 8 * N * D * I + 4 * N * T * I + 6 * N * D * F
    fwd = 2 * N * P * D + cfgd["depth"] * (layer_s + layer_t) + 2 * N * D * C + 2 * S * D * L
    return fwd, 3 * (fwd - 2 * N * D * C) + 2 * N * D * C


def build_model(args, device):
    from transformers import BertConfig, BertModel

    from ct_clip_b200 import CTCLIP, CTViT
    torch.manual_seed(0)
    vit = CTViT(dim=args.dim, codebook_size=8192, image_size=args.image, patch_size=args.image // 24 if args.image % 24 == 0 else 16,
                temporal_patch_size=args.frames // 24 if args.frames % 24 == 0 else 8, spatial_depth=args.depth,
                temporal_depth=args.depth, dim_head=32, heads=8)
    bert = BertModel(BertConfig(num_hidden_layers=args.bert_layers, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0))
    h, w = vit.patch_height_width
    clip = CTCLIP(image_encoder=vit, text_encoder=bert, dim_text=768, dim_image=h * w * args.dim, dim_latent=512)
    return clip.to(device)


def run_b200(args):
    import torch.distributed as dist

    from ct_clip_b200 import _lib, ops
    from ct_clip_b200.data import SyntheticCTReportDataset
    from ct_clip_b200.trainer import CTClipTrainer, _Tokens
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N > 1)"
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    clip = build_model(args, device)
    ds = SyntheticCTReportDataset(10 ** 9, frames=args.frames, image=args.image, n_text=args.text_len, seed=1234 + 7919 * rank)
    trainer = CTClipTrainer(clip, num_train_steps=10 ** 9, batch_size=args.batch, train_dataset=ds, num_workers=0,
                            save_model_every=0, save_results_every=0, results_folder="/tmp/ctclip_bench")
    # two distinct synthetic batches in pinned host memory (885 MB of int16 per batch: larger than the 126 MB L2,
    # so consecutive steps cannot reuse cached inputs)
    host = []
    for j in range(2):
        items = [ds[j * args.batch + i] for i in range(args.batch)]
        vol, tok = ds.collate(items)
        host.append((vol.pin_memory(), tok["input_ids"].pin_memory(), tok["attention_mask"].pin_memory()))
    dev = [(v.to(device), _Tokens(i.to(device), m.to(device))) for v, i, m in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident steps (value)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()      # early: nvidia-smi takes a few hundred ms to deliver its first sample
    for w_ in range(args.warmup):
        trainer.step_on_batch(*dev[w_ % 2])
    barrier()
    # CPU cost of enqueueing the forward pass (about a third of a step's launches) from an EMPTY launch queue: tells whether
    # the Python/ctypes launch path could become the bottleneck (the whole-step figure below includes queue back-pressure)
    clip.train()
    torch.cuda.synchronize()
    t_f0 = time.perf_counter()
    l0 = _lib.launch_count
    with torch.no_grad():
        clip(dev[0][1], dev[0][0], return_loss=True, device=device)
    host_fwd_ms = (time.perf_counter() - t_f0) * 1e3
    host_fwd_launches = _lib.launch_count - l0
    barrier()
    # ---- per-stage pass BEFORE the timed region (untimed): CUDA events around every C-ABI call of ONE extra step. It names the
    # dominant stage, which is then timed INSIDE the timed region (below). EVERY rank runs the extra step (it contains the
    # embedding all-gather and the gradient all-reduce: a rank-0-only step would dead-lock at N > 1); only rank 0 records.
    stage_rows, dom_key = None, None
    if not args.no_stages:
        rec = [] if rank == 0 else None
        _lib.STAGE_TIMER = rec
        trainer.step_on_batch(*dev[0])
        barrier()
        _lib.STAGE_TIMER = None
        if rank == 0:
            stage_rows = stage_table(rec, load_peaks())
            dom_key = stage_rows[0][7] if stage_rows else None       # (entry point, tag) of the stage with the largest ms
    in_region = [] if rank == 0 else None
    if rank == 0:
        _lib.STAGE_FILTER = (lambda name, tag: name == "ctclip_gemm_bf16" or (dom_key is not None and (name, tag or "") == dom_key))
        _lib.STAGE_TIMER = in_region
    launches0 = _lib.launch_count
    t_wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_host0 = time.perf_counter()
    for s in range(args.steps):
        loss = trainer.step_on_batch(*dev[s % 2])
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / max(1, args.steps)   # CPU time to enqueue one step (no sync)
    e1.record()
    barrier()
    _lib.STAGE_TIMER = None
    _lib.STAGE_FILTER = None
    ms = e0.elapsed_time(e1)
    launches = (_lib.launch_count - launches0) // max(1, args.steps)
    if rank == 0:
        sampler.window(t_wall0, time.time())
    clocks = sampler.stop() if rank == 0 else None
    loss_val = float(loss.item())

    # ---- end-to-end through the public API path (CTClipTrainer.train_step): every step copies its batch from pinned host
    # memory (overlapped with the previous step by the trainer's DevicePrefetcher) and reads the loss back to the host
    from ct_clip_b200.trainer import DevicePrefetcher

    def host_batches():
        i = 0
        while True:
            v, ii, mm = host[i % 2]
            yield v, dict(input_ids=ii, attention_mask=mm)
            i += 1
    trainer._prefetcher = DevicePrefetcher(host_batches(), device, trainer._tokenize)
    trainer.print = lambda msg: None
    trainer.train_step()                       # primes the copy pipeline (untimed)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for s in range(args.steps):
        logs = trainer.train_step()            # H2D copy of its batch + step + loss.item()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    if rank == 0 and stage_rows is not None and os.environ.get("CTCLIP_BENCH_STAGE_TABLE"):
        write_stage_table(os.environ["CTCLIP_BENCH_STAGE_TABLE"], stage_rows, ms / args.steps)
    t = torch.tensor([ms, ms_e2e], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()

    # ---- in-region events: the GEMM family (all launches) and the dominant stage
    gsum_ms, gflops, n_g = 0.0, 0.0, 0
    by_shape = {}
    dom = [0, 0.0, 0.0, 0.0, ""]     # launches, ms, flops-or-bytes, bytes (GEMM), kind
    for name, tag, work, ev0, ev1 in (in_region or []):
        t_ = ev0.elapsed_time(ev1)
        if name == "ctclip_gemm_bf16":
            gsum_ms += t_
            gflops += work[1]
            n_g += 1
            a_ = by_shape.setdefault(tag, [0, 0.0, 0.0])
            a_[0] += 1
            a_[1] += t_
            a_[2] += work[1]
        if dom_key is not None and (name, tag or "") == dom_key:
            dom[0] += 1
            dom[1] += t_
            dom[2] += work[1] if work else 0.0
            dom[3] += work[2] if work and len(work) > 2 else 0.0
            dom[4] = work[0] if work else ""
    if rank == 0 and os.environ.get("CTCLIP_BENCH_GEMM_TABLE"):
        with open(os.environ["CTCLIP_BENCH_GEMM_TABLE"], "w") as f:
            f.write("MxNxK a_major b_major epi splits | launches total_ms TFLOP/s\n")
            for shp, (c_, t_, fl_) in sorted(by_shape.items(), key=lambda kv: -kv[1][1]):
                f.write(f"{shp} | {c_} {t_:.3f} {fl_ / (t_ * 1e-3) / 1e12 if t_ > 0 else 0:.1f}\n")
    peaks = load_peaks()
    achieved = (gflops / (gsum_ms * 1e-3) / 1e12) if gsum_ms > 0 else 0.0
    # dominant GEMM shape (kept as roofline.gemm_family.dominant_shape)
    dom_shape, (dom_n, dom_ms, dom_fl) = max(by_shape.items(), key=lambda kv: kv[1][1]) if by_shape else (None, (0, 0.0, 0.0))
    dom_achieved = dom_fl / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    roof = None
    if dom_key is not None and dom[1] > 0:
        kind, w = dom[4], dom[2]
        if kind == "FB":    # a GEMM: whichever roofline bounds the launch (same rule as the stage table)
            kind = "B" if dom[3] / (peaks["hbm"] * 1e9) > dom[2] / (peaks["tf_sus"] * 1e12) else "F"
            w = dom[3] if kind == "B" else dom[2]
        if kind == "B":
            ach, unit, peak, bound, psrc = w / (dom[1] * 1e-3) / 1e9, "GB/s", peaks["hbm"], "hbm", "hbm_gbs (STREAM-style copy)"
        else:
            ach, unit, peak, bound, psrc = w / (dom[1] * 1e-3) / 1e12, "TFLOP/s", peaks["tf_sus"], "tensor", "bf16_tflops_sustained (kernel timed inside a long step)"
        traffic = None
        tpath = ROOT / "profiles" / "ncu_traffic.json"
        stage_name = f"{dom_key[0].replace('ctclip_', '')} [{dom_key[1]}]" if dom_key[1] else dom_key[0].replace("ctclip_", "")
        if tpath.exists():
            traffic = json.loads(tpath.read_text()).get(stage_name)
        roof = {"kernel": f"{stage_name}: the C-ABI entry point (all kernels it launches) with the largest share of the step in the "
                          "per-stage table; timed with CUDA events around each of its calls inside the timed region",
                "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak if peak else None,
                "traffic": traffic, "traffic_note": "dram__bytes_read+write per launch from profiles/ncu_traffic.json (ncu --set full capture)",
                "peak_source": f"{peaks['src']} {psrc}",
                "launches_per_step": dom[0] // max(1, args.steps), "share_of_step": dom[1] / ms if ms > 0 else None}
    vit = clip.visual_transformer
    g = vit.geom
    T = args.frames // g.temporal_patch
    cfgd = dict(P=g.patch_voxels, N=T * g.S, S=g.S, T=T, depth=args.depth, D=args.dim)
    fwd_flops, step_flops = flops_per_volume(cfgd)
    if world > 1:      # leave the NCCL communicator cleanly on every rank (NCCL warns about leaked process groups otherwise)
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    vols = args.batch * world * args.steps
    h2d = sum(x.numel() * x.element_size() for x in host[0])
    out = {
        "metric": f"CT volumes/sec contrastive step @ {args.image}x{args.image}x{args.frames}, bs{args.batch}/GPU",
        "value": vols / (ms * 1e-3), "unit": "volumes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"BASELINE configs[{1 if args.dim == 512 else 4}]: CTViT dim{args.dim} depth{args.depth}+{args.depth}, {args.image}x{args.image}x{args.frames} "
                               f"int16 volumes, patch ({g.patch_hw[0]},{g.patch_hw[1]},{g.temporal_patch}), {args.text_len}-token text, "
                               f"BERT-base({args.bert_layers}L, random init), bs{args.batch}/GPU",
                   "global_batch": args.batch * world, "parallelism": f"dp{world}",
                   "l2": "two alternating input batches of 885 MB each (> 126 MB L2); activations of a step exceed L2 by >100x",
                   "text_tower": "BERT-base on the native kernels (ct_clip_b200/bert.py), weights from the injected HF BertModel",
                   "step_tflop_per_volume_algorithmic": step_flops / 1e12},
        "e2e": {"value": vols / (ms_e2e * 1e-3), "unit": "volumes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "loss": loss_val,
        "clocks": clocks,
        "roofline": dict(roof or {"kernel": None, "bound": "tensor", "achieved": None, "peak": peaks["tf_sus"], "unit": "TFLOP/s",
                                  "frac": None, "traffic": None},
                         gemm_family={"achieved": achieved, "unit": "TFLOP/s", "frac": achieved / peaks["tf_sus"] if peaks["tf_sus"] else None,
                                      "launches_per_step": n_g // max(1, args.steps), "share_of_step": gsum_ms / ms if ms > 0 else None,
                                      "dominant_shape": dom_shape, "dominant_shape_tflops": dom_achieved,
                                      "dominant_shape_share_of_step": dom_ms / ms if ms > 0 else None}),
        "host_enqueue_ms_per_step": host_enqueue_ms,
        "host_enqueue_fwd": {"ms": host_fwd_ms, "launches": host_fwd_launches,
                             "note": "CPU time to enqueue one forward pass from an empty queue (no back-pressure)"},
        "model_tflops": step_flops * vols / (ms * 1e-3) / 1e12,
    }
    if stage_rows is not None:
        out["stages"] = [dict(stage=r[0], launches=r[1], ms=round(r[2], 3), bound=r[3], achieved=round(r[4], 1), unit=r[5],
                              frac=round(r[6], 3)) for r in stage_rows[:16]]
    if not args.no_cpu_baseline and world == 1:      # the CPU leg runs on rank 0 at N = 1 only
        out["cpu_baseline"] = cpu_baseline_guarded(args)
    _emit(_OUT_FD, out)
    if world > 1:
        dist.destroy_process_group()


def _pin_threads(cores):
    """Pin this process to the first `cores` hardware threads (one NUMA node on the 2-socket bench host) and tell torch to use
    exactly that many: removes the run-to-run spread of a floating 32-thread team on a 128-thread machine."""
    try:
        avail = sorted(os.sched_getaffinity(0))
        os.sched_setaffinity(0, set(avail[:cores]))
        return avail
    except Exception:
        return None


def cpu_baseline(args, sample_volumes=2, timed_steps=1, sample_frames=None, warmup_steps=0):
    """The reference algorithm's own CPU PyTorch path (oracle port: the Python reference cannot travel to the GPU box) on the
    host cores, through the WHOLE optimiser step of CTCLIPTrainer.py:244-263: forward + InfoNCE loss + backward +
    clip_grad_norm_(0.5) + Adam. Default sample: `sample_volumes` WHOLE volumes (all frames, full width and depth) -- the
    same model/config as the GPU arm at a smaller batch (b = 2). `sample_frames` crops the temporal
    extent (then scaled back to whole volumes and reported as such in `sample`); value = median over `timed_steps`."""
    from oracle import ctclip_oracle as O
    # torch's CPU eager kernels stop scaling (and then regress) beyond a few tens of threads on these shapes: measured on
    # the 128-thread bench host, the 128-thread run was 3x slower than an 8-thread run. Use min(cores, 32) pinned threads.
    cores = min(os.cpu_count() or 1, 32)
    old_aff = _pin_threads(cores)
    old_threads = torch.get_num_threads()
    torch.set_num_threads(cores)
    p = args.image // 24 if args.image % 24 == 0 else 16
    pt = args.frames // 24 if args.frames % 24 == 0 else 8
    if sample_frames is None:
        sample_frames = args.frames
    cfg = O.CTCLIPConfig(vit=O.CTViTConfig(image_size=args.image, patch_size=p, temporal_patch_size=pt, spatial_depth=args.depth,
                                           temporal_depth=args.depth), bert=O.BertConfigLite(layers=args.bert_layers))
    shapes = oracle_shapes(cfg)
    sd = {k: (v.requires_grad_(True) if v.is_floating_point() and "_codebook" not in k else v)
          for k, v in O.synth_state_dict(shapes, 0).items()}
    params = [v for v in sd.values() if v.is_floating_point() and v.requires_grad]
    opt = torch.optim.Adam(params, lr=1.25e-6, betas=(0.9, 0.99))          # optimizer.py:23-34 with wd = 0
    hu, ids, mask = O.synth_inputs(max(2, sample_volumes), sample_frames, args.image, args.text_len)
    hu, ids, mask = hu[:sample_volumes], ids[:sample_volumes], mask[:sample_volumes]
    video = hu.float() / 1000.0
    times = []
    for it in range(warmup_steps + timed_steps):
        t0 = time.time()
        opt.zero_grad(set_to_none=True)
        out = O.ctclip_forward(sd, cfg, ids, mask, video, training=True)
        out["loss"].backward()
        torch.nn.utils.clip_grad_norm_(params, 0.5)                        # CTCLIPTrainer.py:259-260
        opt.step()
        if it >= warmup_steps:
            times.append(time.time() - t0)
    times.sort()
    dt = times[len(times) // 2]
    torch.set_num_threads(old_threads)
    if old_aff is not None:
        try:
            os.sched_setaffinity(0, set(old_aff))
        except Exception:
            pass
    frac = sample_frames / args.frames
    cpu_model = "unknown CPU"
    try:
        with open("/proc/cpuinfo") as f:
            cpu_model = next(ln.split(":", 1)[1].strip() for ln in f if ln.startswith("model name"))
    except Exception:
        pass
    what = "WHOLE volumes" if sample_frames == args.frames else f"slabs of {sample_frames}/{args.frames} frames (scaled x{1 / frac:.0f})"
    return {"value": sample_volumes * frac / dt, "unit": "volumes/s", "cores": cores, "kind": "port",
            "cpu": f"{cpu_model}, {os.cpu_count()} hardware threads on the host, {cores} used (pinned)",
            "sample": f"median of {timed_steps} step(s) (+{warmup_steps} warm-up) of forward + loss + backward + clip_grad_norm_(0.5) + Adam on "
                      f"{sample_volumes} {what} ({args.image}x{args.image}x{sample_frames}, depth {args.depth}+{args.depth}, BERT "
                      f"{args.bert_layers}L, {args.text_len} tokens), fp32, torch CPU, {cores} pinned threads",
            "same_model_config": sample_frames == args.frames, "seconds": dt, "all_seconds": [round(t, 2) for t in times]}


def _mem_available_gb():
    try:
        with open("/proc/meminfo") as f:
            for ln in f:
                if ln.startswith("MemAvailable"):
                    return int(ln.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


def run_zero_shot(args):
    """BASELINE configs[3] (scripts/run_zero_shot.py / zero_shot.py:106-171): the 36-prompt text bank is encoded once, every volume
    goes through the image tower once. `value`: volumes resident in HBM; `e2e`: CTClipInference.infer() on host batches (pinned
    memory, the H2D copy of every batch and the D2H read-back of its (b, 18) probabilities inside the timed region)."""
    import numpy as np

    from ct_clip_b200 import _lib
    from ct_clip_b200.inference import CTClipInference, prompts
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(device)
    clip = build_model(args, device).eval()
    n_prompts = len(prompts())
    g = torch.Generator().manual_seed(7)
    ids = torch.randint(5, 30522, (n_prompts, args.text_len), generator=g)
    lens = torch.randint(8, 16, (n_prompts,), generator=g)
    mask = (torch.arange(args.text_len)[None, :] < lens[:, None]).long()
    ids[:, 0] = 2
    ids = ids * mask
    bsz = args.batch
    pool = [(torch.randn(bsz, 1, args.frames, args.image, args.image, generator=g) * 450.0 - 300.0).round_().clamp_(-1000, 1000)
            .to(torch.int16).pin_memory() for _ in range(2)]          # 2 x 885 MB: alternating batches > L2

    class Pool(torch.utils.data.Dataset):
        def __init__(self, n):
            self.n = n

        def __len__(self):
            return self.n

        def __getitem__(self, i):
            return pool[(i // bsz) & 1][i % bsz], "", np.zeros(18, dtype=np.float32), f"vol{i}"
    n_vol = (args.volumes + bsz - 1) // bsz * bsz
    inf = CTClipInference(clip, dataset=Pool(n_vol), prompt_tokens=dict(input_ids=ids, attention_mask=mask), batch_size=bsz,
                          results_folder="/tmp/ctclip_zero_shot")
    sampler = ClockSampler(device.index or 0)
    sampler.start()
    # ---- device-resident: text bank once + image tower per batch
    dev_pool = [p.to(device) for p in pool]
    from ct_clip_b200 import ops
    with torch.no_grad():
        for w in range(max(3, args.warmup)):
            clip.encode_image_latents(dev_pool[w & 1])
        torch.cuda.synchronize()
        l0 = _lib.launch_count
        t_wall0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        text_lat = clip.encode_text_latents(inf._bank())
        for b in range(n_vol // bsz):
            img = clip.encode_image_latents(dev_pool[b & 1])
            probs = torch.empty(bsz, 18, device=device)
            ops.zero_shot_probs(img, text_lat, clip.temperature, probs)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count - l0
    sampler.window(t_wall0, time.time())
    clocks = sampler.stop()
    # ---- end to end through the public API (DataLoader over pinned host volumes -> H2D -> infer -> probabilities on the host)
    inf.infer()                                  # warm-up pass over the pool (untimed)
    torch.cuda.synchronize()
    t0 = time.time()
    pred = inf.infer()
    torch.cuda.synchronize()
    dt = time.time() - t0
    out = {"metric": "CT volumes/sec zero-shot inference (18 pathologies x 2 prompts), 480x480x240", "value": n_vol / (ms * 1e-3),
           "unit": "volumes/s", "n_gpus": 1, "steps": n_vol // bsz, "warmup": max(3, args.warmup), "ms_per_step": ms / (n_vol // bsz),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
           "config": {"workload": f"BASELINE configs[3]: 36-prompt text bank x {n_vol} synthetic {args.image}x{args.image}x{args.frames} int16 volumes, "
                                  f"CTViT dim512 depth{args.depth}+{args.depth}, batches of {bsz}", "parallelism": "dp1",
                      "l2": "two alternating 885 MB input batches (> 126 MB L2)"},
           "e2e": {"value": n_vol / dt, "unit": "volumes/s", "h2d_bytes_per_step": int(pool[0].numel() * 2), "d2h_bytes_per_step": bsz * 18 * 4,
                   "note": "wall clock around CTClipInference.infer(): DataLoader collation of pinned volumes + H2D + image tower + D2H"},
           "gpu_launches": launches // max(1, n_vol // bsz), "clocks": clocks, "probs_shape": list(pred.shape)}
    _emit(_OUT_FD, out)


def cpu_baseline_guarded(args):
    """cpu_baseline leg of the GPU arm: the reference arm in a CHILD process (the eager CPU autograd of 2 whole volumes at
    12+12 layers keeps ~45 GB of activations: an out-of-memory kill or a slow host must not take the bench line with it).
    Falls back to temporally cropped volumes (said so in `sample`) when the host has < 96 GB available or the child fails."""
    frames = args.cpu_sample_frames
    if frames is None and _mem_available_gb() < 96:
        frames = min(args.frames, 4 * (args.frames // 24 if args.frames % 24 == 0 else 8))
    for attempt_frames in ([frames] if frames is not None else [None, 4 * (args.frames // 24 if args.frames % 24 == 0 else 8)]):
        cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--depth", str(args.depth),
               "--image", str(args.image), "--frames", str(args.frames), "--text-len", str(args.text_len), "--bert-layers", str(args.bert_layers)]
        if attempt_frames is not None:
            cmd += ["--cpu-sample-frames", str(attempt_frames)]
        try:
            env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
            for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
                env.pop(k, None)
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            if r.returncode == 0 and line:
                return json.loads(line[-1])["cpu_baseline"]
        except Exception:
            pass
    return {"value": None, "unit": "volumes/s", "cores": 0, "kind": "port", "sample": "CPU baseline child process failed twice (see stderr)"}


def oracle_shapes(cfg):
    """State-dict shapes of the reference layout for an oracle config (no module construction needed)."""
    v, b = cfg.vit, cfg.bert
    D, I, F, P, Hd = v.dim, v.dim_head * v.heads, v.ff_inner, v.patch_voxels, v.heads
    s = {"temperature": (), "to_text_latent.weight": (cfg.dim_latent, cfg.dim_text),
         "to_visual_latent.weight": (cfg.dim_latent, cfg.dim_image())}
    pre = "visual_transformer."
    s.update({pre + "spatial_rel_pos_bias.net.0.0.weight": (D, 2), pre + "spatial_rel_pos_bias.net.0.0.bias": (D,),
              pre + "spatial_rel_pos_bias.net.1.0.weight": (D, D), pre + "spatial_rel_pos_bias.net.1.0.bias": (D,),
              pre + "spatial_rel_pos_bias.net.2.weight": (Hd, D), pre + "spatial_rel_pos_bias.net.2.bias": (Hd,),
              pre + "to_patch_emb.1.weight": (P,), pre + "to_patch_emb.1.bias": (P,), pre + "to_patch_emb.2.weight": (D, P),
              pre + "to_patch_emb.2.bias": (D,), pre + "to_patch_emb.3.weight": (D,), pre + "to_patch_emb.3.bias": (D,),
              pre + "vq._codebook.embed": (1, v.codebook_size, D), pre + "vq._codebook.cluster_size": (1, v.codebook_size)})
    for stack, depth in (("enc_spatial_transformer", v.spatial_depth), ("enc_temporal_transformer", v.temporal_depth)):
        for i in range(depth):
            lp = f"{pre}{stack}.layers.{i}."
            s.update({lp + "0.dsconv.weight": (D, 1, 3, 3, 3), lp + "0.dsconv.bias": (D,), lp + "1.norm.gamma": (D,),
                      lp + "1.norm.beta": (D,), lp + "1.q_scale": (v.dim_head,), lp + "1.k_scale": (v.dim_head,),
                      lp + "1.to_q.weight": (I, D), lp + "1.to_kv.weight": (2 * I, D), lp + "1.to_out.weight": (D, I),
                      lp + "3.0.weight": (D,), lp + "3.0.bias": (D,), lp + "3.1.weight": (2 * F, D), lp + "3.4.weight": (D, F)})
        s.update({f"{pre}{stack}.norm_out.gamma": (D,), f"{pre}{stack}.norm_out.beta": (D,)})
    t = "text_transformer."
    s.update({t + "embeddings.word_embeddings.weight": (b.vocab_size, b.hidden), t + "embeddings.position_embeddings.weight": (b.max_pos, b.hidden),
              t + "embeddings.token_type_embeddings.weight": (b.type_vocab, b.hidden), t + "embeddings.LayerNorm.weight": (b.hidden,),
              t + "embeddings.LayerNorm.bias": (b.hidden,)})
    for i in range(b.layers):
        lp = f"{t}encoder.layer.{i}."
        for nm in ("attention.self.query", "attention.self.key", "attention.self.value", "attention.output.dense"):
            s.update({lp + nm + ".weight": (b.hidden, b.hidden), lp + nm + ".bias": (b.hidden,)})
        s.update({lp + "attention.output.LayerNorm.weight": (b.hidden,), lp + "attention.output.LayerNorm.bias": (b.hidden,),
                  lp + "intermediate.dense.weight": (b.intermediate, b.hidden), lp + "intermediate.dense.bias": (b.intermediate,),
                  lp + "output.dense.weight": (b.hidden, b.intermediate), lp + "output.dense.bias": (b.hidden,),
                  lp + "output.LayerNorm.weight": (b.hidden,), lp + "output.LayerNorm.bias": (b.hidden,)})
    return s


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path (oracle port), whole optimiser step, on 2 WHOLE volumes
    (the reference's CPU-runnable batch), 1 warm-up + median of up to 3 timed steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vols_per_step = 2      # the InfoNCE loss needs >= 2 samples to be non-trivial
    r = cpu_baseline(args, sample_volumes=vols_per_step, timed_steps=max(1, min(args.steps, 3)), warmup_steps=max(0, min(args.warmup, 1)),
                     sample_frames=args.cpu_sample_frames)
    out = {"impl": "reference", "metric": "CT volumes/sec contrastive step @ 480x480x240, bs8/GPU", "value": r["value"],
           "unit": "volumes/s", "n_gpus": args.gpus, "steps": len(r["all_seconds"]), "warmup": min(args.warmup, 1),
           "ms_per_step": 1e3 * r["seconds"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic",
           "config": {"workload": "same model/config as the b200 arm (BASELINE configs[1]) at batch 2; each step = " + r["sample"],
                      "parallelism": "cpu", "global_batch": vols_per_step},
           "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "cpu", "sample", "same_model_config")},
           "e2e": {"value": r["value"], "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    _emit(_OUT_FD, out)


def _quiet_stdout():
    """Native libraries (NCCL prints its version banner) write to file descriptor 1: route fd 1 to stderr for the whole run
    and keep a private duplicate for the ONE JSON line the contract asks for."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return real


def _emit(fd, obj):
    os.write(fd, (json.dumps(obj) + "\n").encode())


if __name__ == "__main__":
    a = parse()
    _OUT_FD = _quiet_stdout()
    if a.impl == "reference":
        run_reference(a)
    elif a.config == "zero_shot":
        run_zero_shot(a)
    else:
        run_b200(a)
