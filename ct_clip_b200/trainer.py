"""CTClipTrainer -- drop-in for scripts/CTCLIPTrainer.py:113-348 (reference), B200-native step loop.

What one `train_step()` does (reference lines in brackets):
  batch -> device [CTCLIPTrainer.py:244-251] -> contrastive forward + loss [:254-255] -> backward [:257]
  -> gradient all-reduce across ranks [accelerate DDP] -> clip_grad_norm_(0.5) [:259-260] -> Adam step [:262-263].

B200-first differences (all documented in DESIGN.md):
  * one process per GPU with torch.distributed/NCCL (launched by torchrun) instead of HF accelerate;
  * every trainable tensor lives in ONE flat fp32 arena (params / grads / Adam m / Adam v), so the
    gradient all-reduce is a single NCCL call over NVLink/NVSwitch and clip + Adam are two kernels;
  * the loss is the GLOBAL-batch InfoNCE: latents are all-gathered before the similarity matrix
    (north_star; the reference's DDP loss is rank-local) and gradients are summed, which equals the
    single-process reference at the global batch size;
  * parameters that can never receive a gradient on this path (the *_extra projection copies,
    ct_clip.py:579-581) are left out of the arena, like torch's Adam skips grad-less tensors.
"""
from __future__ import annotations

import os
import warnings
from pathlib import Path

import torch
import torch.distributed as dist
from torch import nn

from . import ops
from .ctclip import CTCLIP
from .data import SyntheticCTReportDataset, cycle  # noqa: F401


def _dist_env():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


def merge_ranges(ranges):
    """Sorted union of half-open [lo, hi) ranges (touching ranges are joined)."""
    out = []
    for a, b in sorted(ranges):
        if out and a <= out[-1][1]:
            out[-1] = (out[-1][0], max(out[-1][1], b))
        else:
            out.append((a, b))
    return out


def complement_ranges(merged, numel):
    """[0, numel) minus a sorted, disjoint list of ranges."""
    out, pos = [], 0
    for a, b in merged:
        if a > pos:
            out.append((pos, a))
        pos = max(pos, b)
    if pos < numel:
        out.append((pos, numel))
    return out


def rest_slices(names, offsets, numel, name):
    """[lo, hi) ranges of a flat arena (tensor `names[i]` starts at `offsets[i]`) that are NOT covered by tensor `name`:
    what remains to be all-reduced after `name` went out early."""
    i = names.index(name)
    lo = offsets[i]
    hi = offsets[i + 1] if i + 1 < len(offsets) else numel
    return [(a, b) for a, b in ((0, lo), (hi, numel)) if b > a]


class ParamArena:
    """Flat fp32 storage for parameters, gradients and Adam moments; module parameters become views into it."""

    def __init__(self, named_params, device):
        self.names, self.params, self.offsets = [], [], []
        off = 0
        # tensors with ndim >= 2 first: they are the weight-decayed group of the reference's AdamW (optimizer.py:3-8, 26-34), so
        # decoupled weight decay is "scale the first n_decay elements" inside the Adam kernel -- no mask, no extra pass
        named_params = list(named_params)
        ordered = [(n, p) for n, p in named_params if p.ndim >= 2] + [(n, p) for n, p in named_params if p.ndim < 2]
        self.n_decay = 0
        for n, p in ordered:
            self.names.append(n)
            self.params.append(p)
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4          # keep every tensor 16-byte aligned
            if p.ndim >= 2:
                self.n_decay = off
        self.numel = max(off, 4)
        self.p = torch.zeros(self.numel, device=device)
        self.g = torch.zeros(self.numel, device=device)
        self.m = torch.zeros(self.numel, device=device)
        self.v = torch.zeros(self.numel, device=device)
        self.grad_views = {}
        for n, p, o in zip(self.names, self.params, self.offsets):
            k = p.numel()
            self.p[o:o + k].copy_(p.data.reshape(-1))
            p.data = self.p[o:o + k].view(p.shape)
            gv = self.g[o:o + k].view(p.shape)
            p.grad = gv
            self.grad_views[n] = gv
        self.sumsq = torch.zeros(1, device=device)
        self.step = 0

    def zero_grad(self):
        self.g.zero_()
        for p, n in zip(self.params, self.names):   # autograd may have replaced .grad; pin it to the arena view again
            if p.grad is None or p.grad.data_ptr() != self.grad_views[n].data_ptr():
                p.grad = self.grad_views[n]

    def adam_step(self, *, lr, max_norm, grad_scale=1.0, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.0):
        """weight_decay == 0: Adam (optimizer.py:23-24); > 0: AdamW on the ndim >= 2 tensors (optimizer.py:26-34)."""
        self.step += 1
        self.sumsq.zero_()
        ops.grad_sumsq(self.g, self.numel, self.sumsq)
        ops.adam_step(self.p, self.g, self.m, self.v, self.numel, lr=lr, beta1=betas[0], beta2=betas[1], eps=eps,
                      step=self.step, max_norm=(max_norm or 0.0), sumsq=self.sumsq, grad_scale=grad_scale,
                      weight_decay=weight_decay, n_decay=self.n_decay)

    def state_dict(self):
        return dict(step=self.step, names=list(self.names), offsets=list(self.offsets), exp_avg=self.m.cpu(),
                    exp_avg_sq=self.v.cpu())

    def load_state_dict(self, sd):
        assert list(sd["names"]) == list(self.names), "optimizer state does not match this model"
        self.step = int(sd["step"])
        self.m.copy_(sd["exp_avg"])
        self.v.copy_(sd["exp_avg_sq"])


class GradBucketer:
    """Gradient all-reduce overlapped with the backward pass (reverse-forward order, like DDP's buckets) over the flat arena:
      * the arena lays the ndim >= 2 tensors out first, in module order: the tensors of one transformer layer are contiguous and
        consecutive layers adjacent, so "layer i is done" extends a pending [lo, hi) range downwards;
      * a pending range goes on the wire (async all-reduce on the process group's own stream: it starts once the kernels enqueued
        so far are done and overlaps everything enqueued after it) as soon as it holds >= bucket_bytes;
      * whatever was never announced (1-D tensors, heads, patch embedding ...) is reduced by finish() as the complement."""

    def __init__(self, arena, bucket_bytes=48 << 20, group=None):
        self.arena, self.bucket_bytes, self.group = arena, bucket_bytes, group
        self.pending, self.done, self.works = [], [], []
        self.launches = 0

    def _span(self, i):
        o = self.arena.offsets[i]
        return o, o + (self.arena.params[i].numel() + 3) // 4 * 4

    def _launch(self, lo, hi):
        self.works.append(dist.all_reduce(self.arena.g[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.done.append((lo, hi))
        self.launches += 1

    def tensor_ready(self, name):
        """one (large) tensor is final: put it on the wire right away (the 604 MB visual projection)"""
        self._launch(*self._span(self.arena.names.index(name)))

    def prefix_ready(self, prefix):
        spans = [self._span(i) for i, (n, p) in enumerate(zip(self.arena.names, self.arena.params)) if n.startswith(prefix) and p.ndim >= 2]
        if not spans:
            return
        self.pending = merge_ranges(self.pending + [(min(a for a, _ in spans), max(b for _, b in spans))])
        keep = []
        for a, b in self.pending:
            if (b - a) * 4 >= self.bucket_bytes:
                self._launch(a, b)
            else:
                keep.append((a, b))
        self.pending = keep

    def finish(self):
        """Reduce every arena range that is not on the wire yet, then wait for all collectives of this step."""
        for a, b in self.pending:
            self._launch(a, b)
        self.pending = []
        for a, b in complement_ranges(merge_ranges(self.done), self.arena.numel):
            self._launch(a, b)
        for w in self.works:
            w.wait()
        self.works, self.done = [], []


class CTClipTrainer(nn.Module):
    def __init__(self, CTClip: CTCLIP, *, num_train_steps, batch_size, data_train="train", data_valid="valid",
                 reports_file_train="data_reports.xslx", reports_file_valid="data_reports.xslx",
                 train_meta_file="meta_data.csv", valid_meta_file="meta_data.csv", labels="labels.csv", tokenizer=None,
                 lr=1.25e-6, wd=0., max_grad_norm=0.5, save_results_every=1, save_model_every=1,
                 results_folder='./ctclip/', num_workers=8, accelerate_kwargs: dict = dict(),
                 train_dataset=None, valid_dataset=None, text_max_length=512, async_checkpoints=False):
        super().__init__()
        self.wd = float(wd)      # 0 (reference default): Adam; > 0: AdamW on the ndim >= 2 tensors (optimizer.py:10-34)
        if not torch.cuda.is_available():
            raise RuntimeError("CTClipTrainer needs a CUDA (sm_100a) device: there is no CPU training path")
        self.world, self.rank, local = _dist_env()
        self.device = torch.device("cuda", local)
        torch.cuda.set_device(self.device)
        if self.world > 1 and not dist.is_initialized():
            dist.init_process_group("nccl", device_id=self.device)
        self.CTClip = CTClip.to(self.device)
        self.tokenizer = tokenizer if tokenizer is not None else CTClip.tokenizer
        self.text_max_length = text_max_length
        self.register_buffer('steps', torch.Tensor([0]))
        self.num_train_steps, self.batch_size = num_train_steps, batch_size
        self.max_grad_norm, self.lr = max_grad_norm, lr
        self.save_model_every, self.save_results_every = save_model_every, save_results_every
        self.async_checkpoints = async_checkpoints

        # ---- data (scripts/data.py output contract: (1,F,H,W) fp32 in [-1,1] or int16 HU + report text / token ids)
        if train_dataset is None:
            from .data import load_reference_dataset
            train_dataset = load_reference_dataset(data_train, reports_file_train, train_meta_file)
        self.ds = train_dataset
        self.valid_ds = valid_dataset
        if valid_dataset is not None or save_results_every not in (0, 1, None):
            import warnings
            warnings.warn("CTClipTrainer: the validation pass of CTCLIPTrainer.py:266-329 (AUROC on valid_dataset every "
                          "save_results_every steps) is outside the hot-path build; valid_dataset / save_results_every are ignored")
        # every rank draws a DISJOINT shard (accelerate.prepare(dl) does the same in the reference): identical batches on all
        # ranks would put world-size copies of each (volume, report) pair into the all-gathered global batch = false negatives
        self.sampler = None
        if self.world > 1:
            from torch.utils.data.distributed import DistributedSampler
            self.sampler = DistributedSampler(self.ds, num_replicas=self.world, rank=self.rank, shuffle=True, drop_last=True)
        self.dl = torch.utils.data.DataLoader(self.ds, num_workers=num_workers, batch_size=batch_size,
                                              shuffle=self.sampler is None, sampler=self.sampler, pin_memory=True,
                                              drop_last=self.world > 1, collate_fn=getattr(self.ds, "collate", None))
        self.dl_iter = cycle(self.dl, on_epoch=self.sampler.set_epoch if self.sampler is not None else None)
        self._prefetcher = None

        # ---- flat arena over everything that can receive a gradient
        live = [(n, p) for n, p in self.CTClip.named_parameters()
                if not n.startswith(("to_text_latent_extra", "to_visual_latent_extra"))]
        self.arena = ParamArena(live, self.device)
        self.CTClip.mark_weights_dirty()
        self.CTClip._grad_sink = {n: self.arena.grad_views[n] for n in self.arena.names}
        self._setup_dp()

        self.results_folder = Path(results_folder)
        self.results_folder.mkdir(parents=True, exist_ok=True)

    # ------------------------------------------------------------------------------------------
    def _setup_dp(self):
        clip = self.CTClip
        clip.dp_rank, clip.dp_world = self.rank, self.world
        if self.world == 1:
            return
        # identical initial weights on every rank (DDP broadcasts module state at construction)
        dist.broadcast(self.arena.p, src=0)
        for _, b in clip.named_buffers():
            if b.is_floating_point() and b.numel() > 0:
                dist.broadcast(b, src=0)
        clip.mark_weights_dirty()

        from .dist_utils import PeerLatentExchange, gather_latents as gather
        clip.dp_all_gather = gather
        self.latent_exchange = None
        if os.environ.get("CTCLIP_DP_EXCHANGE", "p2p") != "nccl":
            # one kernel over NVLink peer memory; the mapping is set up at the first step. Every rank takes the same branch: a
            # failure of the (collective) setup is agreed on with an all-reduce before the fallback to NCCL is chosen.
            xch = PeerLatentExchange(self.device)

            def exchange(t_raw, i_raw, _x=xch):
                if self.latent_exchange is False:
                    return gather(t_raw, i_raw)
                if _x.key is None:
                    ok = torch.ones(1, device=self.device)
                    try:
                        _x._setup(*t_raw.shape)
                    except Exception as e:      # noqa: BLE001  (no P2P / no symmetric memory on this box)
                        ok.zero_()
                        warnings.warn(f"peer-memory latent exchange unavailable ({e!r}); using the NCCL all-gather")
                    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                    if ok.item() < 1:
                        self.latent_exchange, _x.key = False, None
                        return gather(t_raw, i_raw)
                    self.latent_exchange = _x
                return _x(t_raw, i_raw)

            clip.dp_all_gather = exchange
        self.bucketer = GradBucketer(self.arena)
        clip.dp_early_reduce = self.bucketer.tensor_ready
        clip.dp_grad_ready = self.bucketer.prefix_ready
        clip.visual_transformer.engine.on_grads_ready = lambda prefix: self.bucketer.prefix_ready("visual_transformer." + prefix)
        clip.visual_transformer.ema_all_reduce = lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM)

    @property
    def is_main(self):
        return self.rank == 0

    def print(self, msg):
        if self.is_main:
            print(msg, flush=True)

    def save(self, path):
        if not self.is_main:
            return
        # parameters are views into ONE arena storage: clone them, or every entry would serialise the whole 1.1 GB storage
        model = {k: v.detach().clone() for k, v in self.CTClip.state_dict().items()}
        torch.save(dict(model=model, optim=self.arena.state_dict()), path)

    def load(self, path):
        """Reads a package written by save(); a reference package (CTCLIPTrainer.py:210-217: 'optim' is a torch.optim.Adam
        state dict keyed by parameter index) restores the model and leaves the Adam moments at zero, with a warning."""
        path = Path(path)
        assert path.exists()
        pkg = torch.load(path, map_location="cpu")
        self.CTClip.load_state_dict(pkg['model'])
        optim = pkg.get('optim')
        if isinstance(optim, dict) and "names" in optim and "exp_avg" in optim:
            self.arena.load_state_dict(optim)
        elif optim is not None:
            import warnings
            warnings.warn("CTClipTrainer.load: optimizer state is not in the flat-arena format of this trainer (a "
                          "torch.optim.Adam state dict from the reference?); model restored, Adam moments start from zero")
        self.CTClip.mark_weights_dirty()

    # ------------------------------------------------------------------------------------------
    def _tokenize(self, text):
        if hasattr(text, "input_ids"):
            return text
        if isinstance(text, dict):
            return _Tokens(text["input_ids"], text["attention_mask"])
        if self.tokenizer is None:
            raise RuntimeError("no tokenizer (offline): the dataset must yield token ids")
        return self.tokenizer(list(text), return_tensors="pt", padding="max_length", truncation=True,
                              max_length=self.text_max_length)

    def step_on_batch(self, video, text_tokens):
        """forward + backward + all-reduce + clip + Adam on a batch already resident on the device."""
        self.CTClip.train()
        self.arena.zero_grad()
        loss = self.CTClip(text_tokens, video, return_loss=True, device=self.device)
        loss.backward()
        if self.world > 1:
            self.bucketer.finish()
        probe = getattr(self, "_grad_probe", None)
        if probe is not None:    # tests (tests/dp_check_multigpu.py): the reduced gradient arena, before clipping / Adam
            probe(self.arena)
        ev = getattr(self, "_ckpt_event", None)
        if ev is not None:       # an asynchronous checkpoint copy may still be reading the parameters
            torch.cuda.current_stream().wait_event(ev)
            self._ckpt_event = None
        self.arena.adam_step(lr=self.lr, max_norm=self.max_grad_norm, weight_decay=self.wd)
        self.CTClip.mark_weights_dirty()
        return loss

    def train_step(self):
        steps = int(self.steps.item())
        logs = {}
        if self._prefetcher is None:   # H2D of the next batch overlaps this step's compute
            self._prefetcher = DevicePrefetcher(self.dl_iter, self.device, self._tokenize)
        video, tok = self._prefetcher.next()
        loss = self.step_on_batch(video, tok)
        logs['loss'] = loss.item()                      # device->host sync, as the reference (:258)
        self.print(f"{steps}: loss: {logs['loss']}")
        if self.is_main and self.save_model_every and not (steps % self.save_model_every):
            # CTCLIPTrainer.py:331-337 writes 1.75 GB synchronously here. async_checkpoints=True (opt-in, not yet exercised
            # on a GPU): copy to pinned memory on a side stream + torch.save on a background thread
            model_path = str(self.results_folder / f'CTClip.{steps}.pt')
            if self.async_checkpoints:
                self._save_async(model_path)
            else:
                torch.save({k: v.detach().clone() for k, v in self.CTClip.state_dict().items()}, model_path)
            self.print(f'{steps}: saving model to {str(self.results_folder)}')
        self.steps += 1
        return logs

    def _ckpt_writer(self):
        if getattr(self, "_writer", None) is None:
            from .checkpoint import AsyncCheckpointWriter
            self._writer = AsyncCheckpointWriter()
        return self._writer

    def _save_async(self, path):
        """Parameters are views into the arena: the next Adam step must not overwrite them while the device-to-host copy is in
        flight (step_on_batch waits for the writer's copy event before the optimiser kernels); buffers that the next FORWARD
        mutates (code-book EMA) are cloned on the compute stream first."""
        live = set(self.arena.names)
        sd = {k: (v if k in live else v.clone()) for k, v in self.CTClip.state_dict().items()}
        w = self._ckpt_writer()
        w.save(sd, path)
        self._ckpt_event = w.copy_event

    def train(self, log_fn=lambda logs: None):
        while self.steps < self.num_train_steps:
            logs = self.train_step()
            log_fn(logs)
        if getattr(self, "_writer", None) is not None:
            self._writer.wait()          # the last checkpoint is on disk before train() returns
        self.print('training complete')


class DevicePrefetcher:
    """Overlaps the pinned-host -> device copy of batch i+1 with the compute of batch i (copy stream + events).
    `it` yields (volume, tokens-like) with CPU tensors (pinned for true asynchrony)."""

    def __init__(self, it, device, to_tokens):
        self.it, self.device, self.to_tokens = it, device, to_tokens
        self.stream = torch.cuda.Stream(device=device)
        self._next = None
        self._preload()

    def _preload(self):
        try:
            video, text = next(self.it)
        except StopIteration:
            self._next = None
            return
        tok = self.to_tokens(text)
        with torch.cuda.stream(self.stream):
            vd = video.to(self.device, non_blocking=True)
            td = _Tokens(tok.input_ids.to(self.device, non_blocking=True), tok.attention_mask.to(self.device, non_blocking=True))
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._next = (vd, td, ev)

    def next(self):
        if self._next is None:
            raise StopIteration
        vd, td, ev = self._next
        torch.cuda.current_stream().wait_event(ev)
        for t_ in (vd, td.input_ids, td.attention_mask):
            t_.record_stream(torch.cuda.current_stream())
        self._preload()
        return vd, td


class _Tokens:
    def __init__(self, input_ids, attention_mask):
        self.input_ids, self.attention_mask = input_ids, attention_mask

    def to(self, device):
        return _Tokens(self.input_ids.to(device), self.attention_mask.to(device))
