"""Checkpoint I/O for the drop-in classes.

Reference behaviour being replaced:
  * scripts/CTCLIPTrainer.py:331-337 -- `accelerator.save(state_dict, 'CTClip.{steps}.pt')` on the main rank, every
    `save_model_every` steps (default: EVERY step, 1.75 GB, synchronously inside train_step);
  * scripts/CTCLIPTrainer.py:205-222 -- `save` / `load` of {'model', 'optim'} packages;
  * CT_CLIP/ct_clip/ct_clip.py:593-597 -- `CTCLIP.load(path)`: torch.load + strict load_state_dict.

What this module adds:
  * `tolerant_state_dict`: makes checkpoints written by other wrappers / library versions loadable without weakening the
    strictness for everything else: 'module.' prefixes (DDP / accelerate with unwrap=False), the persistent
    `embeddings.position_ids` buffer of transformers <= 4.30, the never-used `*_extra` projection copies, and keys of the
    GenerateCT stacks that share the CTViT module in some checkpoints (discriminator / VGG / pixel decoders);
  * `AsyncCheckpointWriter`: snapshot -> pinned host memory on a side stream -> `torch.save` on a background thread, so
    that a per-step checkpoint costs the training loop one device-to-host copy it does not wait for.
"""
from __future__ import annotations

import threading
from pathlib import Path

import torch

# key prefixes the CT-CLIP forward path never reads: may be absent from / surplus in a checkpoint
_OPTIONAL_PREFIXES = ("to_text_latent_extra.", "to_visual_latent_extra.")
_FOREIGN_PREFIXES = ("visual_transformer.discr.", "visual_transformer.vgg.", "visual_transformer.discr_optim.")
_DROPPABLE_KEYS = ("text_transformer.embeddings.position_ids",)


def tolerant_state_dict(state_dict, model_state_dict):
    """Return (clean, report): `clean` can be passed to a STRICT load_state_dict of the model whose own state dict is
    `model_state_dict`; `report` lists what was renamed / dropped / filled in. Unknown or shape-mismatched keys are left in
    place so that the strict load still fails loudly on a genuinely different architecture."""
    report = dict(renamed=[], dropped=[], filled=[])
    sd = dict(state_dict)
    if "model" in sd and isinstance(sd["model"], dict) and "optim" in sd:      # trainer package (CTCLIPTrainer.py:209-213)
        sd = dict(sd["model"])
        report["renamed"].append("unwrapped trainer package ['model']")
    # 1. wrapper prefixes
    for prefix in ("module.", "CTClip.", "_orig_mod."):
        if sd and all(k.startswith(prefix) for k in sd):
            sd = {k[len(prefix):]: v for k, v in sd.items()}
            report["renamed"].append(f"stripped '{prefix}'")
    # 2. buffers / sub-modules that newer libraries or this build do not have
    for k in list(sd):
        if k in model_state_dict:
            continue
        if k in _DROPPABLE_KEYS or k.startswith(_FOREIGN_PREFIXES):
            sd.pop(k)
            report["dropped"].append(k)
    # 3. tensors the path never reads: take the model's own value when the checkpoint lacks them
    for k, v in model_state_dict.items():
        if k not in sd and k.startswith(_OPTIONAL_PREFIXES):
            sd[k] = v.detach().clone()
            report["filled"].append(k)
    # 4. the code-book's `initted` flag changed shape between vector-quantize-pytorch versions (scalar <-> [1])
    k = "visual_transformer.vq._codebook.initted"
    if k in sd and k in model_state_dict and sd[k].numel() == model_state_dict[k].numel() and sd[k].shape != model_state_dict[k].shape:
        sd[k] = sd[k].reshape(model_state_dict[k].shape)
        report["renamed"].append(f"reshaped {k}")
    return sd, report


class AsyncCheckpointWriter:
    """One checkpoint in flight at a time. `save(tensors, path)` copies every tensor to (pinned, when CUDA is present) host
    memory on a private stream and returns; a worker thread waits for the copy and runs torch.save. `wait()` joins."""

    def __init__(self):
        self._thread = None
        self._stream = None
        self._host = {}
        self.last_error = None
        self.copy_event = None     # CUDA event recorded after the device-to-host copies of the latest save (None on CPU)

    def _host_buffer(self, name, t):
        buf = self._host.get(name)
        if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
            buf = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=t.is_cuda)
            self._host[name] = buf
        return buf

    def wait(self):
        if self._thread is not None:
            self._thread.join()
            self._thread = None
        if self.last_error is not None:
            err, self.last_error = self.last_error, None
            raise err

    def save(self, tensors: dict, path, extra: dict | None = None):
        """tensors: name -> tensor (device or host). The previous save is joined first (its host buffers are reused)."""
        self.wait()
        path = Path(path)
        event = None
        any_cuda = any(t.is_cuda for t in tensors.values() if torch.is_tensor(t))
        staged = {}
        if any_cuda:
            dev = next(t.device for t in tensors.values() if torch.is_tensor(t) and t.is_cuda)
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=dev)
            self._stream.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(self._stream):
                for n, t in tensors.items():
                    staged[n] = self._host_buffer(n, t).copy_(t.detach(), non_blocking=True) if torch.is_tensor(t) else t
                    if torch.is_tensor(t) and t.is_cuda:
                        # the caller may drop `t` (a clone made on the compute stream) as soon as save() returns: tell the caching
                        # allocator that this side stream still reads it, or the block could be re-used under the queued copy
                        t.record_stream(self._stream)
                event = torch.cuda.Event()
                event.record(self._stream)
            self.copy_event = event
        else:
            for n, t in tensors.items():
                staged[n] = self._host_buffer(n, t).copy_(t.detach()) if torch.is_tensor(t) else t

        def work():
            try:
                if event is not None:
                    event.synchronize()
                payload = dict(staged) if extra is None else dict(model=dict(staged), **extra)
                tmp = path.with_suffix(path.suffix + ".tmp")
                torch.save(payload, tmp)
                tmp.replace(path)            # readers never see a half-written file
            except Exception as e:           # surfaced by the next wait()
                self.last_error = e
        self._thread = threading.Thread(target=work, daemon=True)
        self._thread.start()
