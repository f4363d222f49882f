"""Dataset glue. The reference's NIfTI pipeline (scripts/data.py:92-174: nibabel -> HU -> trilinear resample ->
clip -> /1000 -> crop/pad 480x480x240) is outside the hot-path scope; what the
trainer needs is its OUTPUT CONTRACT: (1, F, H, W) volumes in [-1, 1] (fp32) -- or the same data as int16 HU,
consumed as x/1000 by the patch-embed kernel -- plus the report text / token ids.
"""
from __future__ import annotations

import torch


def cycle(dl, on_epoch=None):  # scripts/CTCLIPTrainer.py:55-58
    """on_epoch(epoch) is called before each pass (DistributedSampler.set_epoch: a different shuffle per epoch)."""
    epoch = 0
    while True:
        if on_epoch is not None:
            on_epoch(epoch)
        for data in dl:
            yield data
        epoch += 1


class SyntheticCTReportDataset(torch.utils.data.Dataset):
    """Synthetic CT volumes + token ids with the statistics SURVEY 8(d) prescribes:
    int16 HU = clip(round(N(-300, 450^2)), -1000, 1000); ids uniform in [5, vocab), [CLS]=2, [SEP]=3, pad 0."""

    def __init__(self, length, frames=240, image=480, n_text=128, vocab=30522, seed=1234, as_int16=True):
        self.length, self.frames, self.image, self.n_text, self.vocab = length, frames, image, n_text, vocab
        self.seed, self.as_int16 = seed, as_int16

    def __len__(self):
        return self.length

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed + i)
        hu = (torch.randn(1, self.frames, self.image, self.image, generator=g) * 450.0 - 300.0).round_().clamp_(-1000, 1000)
        vol = hu.to(torch.int16) if self.as_int16 else hu / 1000.0
        n = self.n_text
        ids = torch.randint(5, self.vocab, (n,), generator=g)
        ln = int(torch.randint(max(2, n // 4), n + 1, (1,), generator=g))
        mask = (torch.arange(n) < ln).long()
        ids[0] = 2
        ids[ln - 1] = 3
        return vol, dict(input_ids=ids * mask, attention_mask=mask)

    @staticmethod
    def collate(batch):
        vols = torch.stack([b[0] for b in batch])
        return vols, dict(input_ids=torch.stack([b[1]["input_ids"] for b in batch]),
                          attention_mask=torch.stack([b[1]["attention_mask"] for b in batch]))


def load_reference_dataset(data_folder, reports_file, meta_file):
    """The reference's CTReportDataset needs nibabel + the CT-RATE files; wire your own Dataset with the same output
    contract through CTClipTrainer(train_dataset=...) when those are not available."""
    raise RuntimeError(
        "NIfTI loading (scripts/data.py) is not part of the B200 hot-path build: pass train_dataset=<Dataset yielding "
        "((1,F,H,W) volume, report text or token ids)> to CTClipTrainer, e.g. the reference's CTReportDataset itself "
        "or ct_clip_b200.data.SyntheticCTReportDataset")
