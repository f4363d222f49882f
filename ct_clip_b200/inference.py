"""CTClipInference -- drop-in for scripts/zero_shot.py:53-171 (reference zero-shot evaluation).

Reference loop: for every validation volume (bs=1) and each of 18 pathologies it tokenises the two prompts
"{p} is present." / "{p} is not present." and re-runs the WHOLE model (18 image-tower passes per volume,
zero_shot.py:133-138), then softmaxes the two similarities and keeps P(present) (:140-143).

B200 path: the 36-prompt text bank is encoded once, every volume goes through the image tower once, and the
(volumes x 36) similarity + pairwise softmax reproduces exactly the same (n_volumes, 18) matrix (eval mode has
no quantiser side effects; only fp reduction order differs).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch
from torch import nn

from .ctclip import CTCLIP

PATHOLOGIES = ['Medical material', 'Arterial wall calcification', 'Cardiomegaly', 'Pericardial effusion',
               'Coronary artery wall calcification', 'Hiatal hernia', 'Lymphadenopathy', 'Emphysema', 'Atelectasis',
               'Lung nodule', 'Lung opacity', 'Pulmonary fibrotic sequela', 'Pleural effusion',
               'Mosaic attenuation pattern', 'Peribronchial thickening', 'Consolidation', 'Bronchiectasis',
               'Interlobular septal thickening']  # zero_shot.py:124


def prompts():
    out = []
    for p in PATHOLOGIES:
        out += [f"{p} is present.", f"{p} is not present."]  # zero_shot.py:134
    return out


class _Tokens:
    def __init__(self, input_ids, attention_mask):
        self.input_ids, self.attention_mask = input_ids, attention_mask


class CTClipInference(nn.Module):
    def __init__(self, CTClip: CTCLIP, *, data_folder="external_valid", reports_file="data_reports.xslx",
                 meta_file="meta.csv", results_folder='./results', labels="labels.csv", accelerate_kwargs: dict = dict(),
                 dataset=None, prompt_tokens=None, tokenizer=None, batch_size=1, num_workers=0):
        super().__init__()
        if not torch.cuda.is_available():
            raise RuntimeError("CTClipInference needs a CUDA (sm_100a) device")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.CTClip = CTClip.to(self.device)
        self.tokenizer = tokenizer if tokenizer is not None else CTClip.tokenizer
        self.prompt_tokens = prompt_tokens
        if dataset is None:
            raise RuntimeError("pass dataset=<Dataset yielding (volume (1,F,H,W), report, onehot labels, accession)> -- "
                               "NIfTI loading (scripts/data_inference_nii.py) is outside the hot-path build")
        self.ds = dataset
        self.dl = torch.utils.data.DataLoader(self.ds, num_workers=num_workers, batch_size=batch_size, shuffle=False)
        self.results_folder = Path(results_folder)
        self.results_folder.mkdir(parents=True, exist_ok=True)
        self.register_buffer('steps', torch.Tensor([0]))

    def _bank(self):
        tok = self.prompt_tokens
        if tok is None:
            if self.tokenizer is None:
                raise RuntimeError("no tokenizer available offline: pass prompt_tokens=dict(input_ids, attention_mask) for "
                                   "the 36 prompts returned by ct_clip_b200.inference.prompts()")
            enc = self.tokenizer(prompts(), return_tensors="pt", padding="max_length", truncation=True, max_length=512)
            tok = dict(input_ids=enc.input_ids, attention_mask=enc.attention_mask)
        return _Tokens(tok["input_ids"].to(self.device), tok["attention_mask"].to(self.device))

    def _host_batches(self):
        """(pinned volume batch, release(), rest of the batch) with the collation of batch k+1 running in a background thread while the
        GPU works on batch k: the volumes are stacked straight into one of three reusable PINNED buffers (one memcpy; the default
        collate builds a pageable tensor, whose host-to-device copy is synchronous), and a buffer is only reused after the event
        recorded behind its H2D copy has fired."""
        import queue
        import threading
        q = queue.Queue(maxsize=1)
        bufs, events = [None] * 3, [None] * 3

        def work():
            try:
                it = iter(torch.utils.data.DataLoader(self.ds, num_workers=self.dl.num_workers, batch_size=self.dl.batch_size,
                                                      shuffle=False, collate_fn=lambda items: items))
                for k, items in enumerate(it):
                    vols = [torch.as_tensor(x[0]) for x in items]
                    slot = k % 3
                    if events[slot] is not None:
                        events[slot].synchronize()
                    shape = (len(vols),) + tuple(vols[0].shape)
                    if bufs[slot] is None or bufs[slot].shape[1:] != shape[1:] or bufs[slot].dtype != vols[0].dtype or bufs[slot].shape[0] < shape[0]:
                        bufs[slot] = torch.empty(shape, dtype=vols[0].dtype).pin_memory()
                    dst = bufs[slot][:len(vols)]
                    torch.stack(vols, out=dst)
                    rest = [[x[j] for x in items] for j in range(1, len(items[0]))]
                    q.put((slot, dst, rest))
                q.put(None)
            except BaseException as e:      # noqa: BLE001  (re-raised in the consumer)
                q.put(e)

        th = threading.Thread(target=work, daemon=True)
        th.start()
        while True:
            item = q.get()
            if item is None:
                break
            if isinstance(item, BaseException):
                raise item
            slot, dst, rest = item

            def copied(slot=slot):
                events[slot] = torch.cuda.Event()
                events[slot].record()
            yield dst, copied, rest
        th.join()

    @torch.no_grad()
    def infer(self, log_fn=lambda logs: None):
        self.CTClip.eval()
        from . import ops
        text_lat = self.CTClip.encode_text_latents(self._bank())              # (36, L), once
        predicted, real, names = [], [], []
        for host_vol, copied, rest in self._host_batches():
            vol = host_vol.to(self.device, non_blocking=True)
            copied()
            img_lat = self.CTClip.encode_image_latents(vol)                     # (b, L), one image pass per volume
            probs = torch.empty(vol.shape[0], len(PATHOLOGIES), device=self.device)
            ops.zero_shot_probs(img_lat, text_lat, self.CTClip.temperature, probs)   # sims * exp(T) + pair softmax (zero_shot.py:140-143)
            predicted.append(probs)                                             # read back once, after the last batch
            if len(rest) > 1:
                real.append(np.asarray([np.asarray(r) for r in rest[1]]).reshape(vol.shape[0], -1))
            if len(rest) > 2:
                names += [str(n) for n in rest[2]]
        predicted = torch.cat(predicted, dim=0).cpu().numpy()
        np.savez(self.results_folder / "predicted_weights.npz", data=predicted)    # zero_shot.py:152-165
        if real:
            np.savez(self.results_folder / "labels_weights.npz", data=np.concatenate(real, axis=0))
        with open(self.results_folder / "accessions.txt", "w") as f:
            for n in names:
                f.write(str(n) + "\n")
        log_fn(dict(n_volumes=int(predicted.shape[0])))
        return predicted
