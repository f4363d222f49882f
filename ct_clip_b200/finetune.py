"""Fine-tuning heads on the same forward.

* `ImageLatentsClassifier` -- ClassFine / LiPro (scripts/ct_lipro_train.py:17-32, 89-106): a linear probe on the frozen model's image
  latents. The reference runs the WHOLE CTCLIP.forward per step -- BERT on a 512-token " " prompt, the image tower with autograd
  bookkeeping -- only to read `image_latents`. Fast path: `CTCLIP.encode_image_latents` (image tower forward only, no text tower,
  nothing saved for backward); only the 512 x 18 head sees autograd.
* `vocabfine_step` -- VocabFine (scripts/ct_vocabfine_train.py:77-123): per volume the reference calls the model 18 times (one
  (yes, no) prompt pair per pathology) and back-propagates three times, i.e. 18 image-tower forwards + 18 backwards for ONE volume.
  Fast path: the 36 prompts of the label-independent text bank go through the text tower as one batch, the image tower runs once,
  and the three MSE terms are summed before ONE backward (gradients add, so the update is the same).
The towers run on the sm_100a kernels through `CTCLIP.latents_with_grad`; the arithmetic left to torch is the [36, L] x [1, L]
similarity, a 2-way softmax and the 18 x 512 classifier.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

PATHOLOGIES = ['Medical material', 'Arterial wall calcification', 'Cardiomegaly', 'Pericardial effusion',
               'Coronary artery wall calcification', 'Hiatal hernia', 'Lymphadenopathy', 'Emphysema', 'Atelectasis', 'Lung nodule',
               'Lung opacity', 'Pulmonary fibrotic sequela', 'Pleural effusion', 'Mosaic attenuation pattern',
               'Peribronchial thickening', 'Consolidation', 'Bronchiectasis', 'Interlobular septal thickening']   # ct_vocabfine_train.py:71-75

# class weights of ct_lipro_train.py:76-80 (BCEWithLogitsLoss pos_weight)
LIPRO_POS_WEIGHT = (9.211362733, 2.384068466, 8.295479204, 32.8629776, 2.992233613, 6.064870808, 3.176470588, 4.187083754,
                    3.022222222, 1.216071737, 1.677849552, 3.152851834, 7.123261694, 18.16629381, 13.8480647, 6.335045662,
                    10.81701149, 13.40695067)


class ImageLatentsClassifier(nn.Module):
    """ct_lipro_train.py:17-32. Same constructor / forward / save / load; the frozen model is only asked for image latents."""

    def __init__(self, trained_model, latent_dim, num_classes, dropout_prob=0.3):
        super().__init__()
        self.trained_model = trained_model
        for param in self.trained_model.parameters():
            param.requires_grad = False
        self.dropout = nn.Dropout(dropout_prob)
        self.relu = nn.ReLU()
        self.classifier = nn.Linear(latent_dim, num_classes)

    def forward(self, *args, **kwargs):
        # reference call: model(text_tokens, inputs, device=..., return_latents=True); text_tokens (a " " prompt) is unused
        image = kwargs.get("image", args[1] if len(args) > 1 else None)
        if image is None:
            raise TypeError("ImageLatentsClassifier.forward(text_tokens, volumes, ...): volumes missing")
        was_training = self.trained_model.training
        self.trained_model.eval()          # frozen tower: no code-book EMA, no dropout inside the towers
        with torch.no_grad():
            image_latents = self.trained_model.encode_image_latents(image.to(self.classifier.weight.device))
        self.trained_model.train(was_training)
        return self.classifier(self.dropout(self.relu(image_latents)))

    def save(self, file_path):
        torch.save(self.state_dict(), file_path)

    def load(self, file_path):
        self.load_state_dict(torch.load(file_path))


def lipro_loss(logits, labels):
    """ct_lipro_train.py:80: BCEWithLogitsLoss(pos_weight=class weights)."""
    w = torch.tensor(LIPRO_POS_WEIGHT, device=logits.device, dtype=logits.dtype)[: logits.shape[-1]]
    return F.binary_cross_entropy_with_logits(logits, labels, pos_weight=w)


def vocabfine_prompts(pathologies=PATHOLOGIES):
    """The label-independent text bank: for every pathology ("<p> is present. ", "<p> is not present. ") -- ct_vocabfine_train.py:96-103."""
    out = []
    for p in pathologies:
        out += [f"{p} is present. ", f"{p} is not present. "]
    return out


def vocabfine_loss_from_latents(text_latents, image_latents, temperature, labels, group=6):
    """text_latents [2*C, L] (present, not-present per pathology), image_latents [1, L], labels [C] in {0, 1}.
    Per pathology: logits = softmax([sim(correct text), sim(wrong text)]) against (1, 0); MSELoss(mean) per group of `group`
    pathologies (ct_vocabfine_train.py:86-119), summed over the groups (the script calls backward() once per group before one
    optimizer step: gradients add)."""
    C = labels.shape[0]
    sims = (text_latents * image_latents).sum(-1) * temperature.exp()            # [2C]  (ct_clip.py:805-807, broadcast over the bank)
    sims = sims.view(C, 2)                                                        # (present, not present)
    lab = labels.to(sims.device).long().view(C, 1)
    correct = torch.where(lab == 1, sims[:, 0:1], sims[:, 1:2])                  # text_yes of the script
    wrong = torch.where(lab == 1, sims[:, 1:2], sims[:, 0:1])
    probs = torch.softmax(torch.cat([correct, wrong], dim=1), dim=1)             # softmax over the (yes, no) pair
    target = torch.tensor([1.0, 0.0], device=probs.device).expand_as(probs)
    total = probs.new_zeros(())
    for g0 in range(0, C, group):
        total = total + F.mse_loss(probs[g0:g0 + group].reshape(-1), target[g0:g0 + group].reshape(-1))
    return total


def vocabfine_step(clip, volume, labels, bank_tokens):
    """One VocabFine training example: volume [1,1,F,H,W], labels [C], bank_tokens = tokenised `vocabfine_prompts()` (object with
    input_ids / attention_mask [2C, n]). Returns the summed loss (call .backward() on it, then step the optimiser)."""
    tl, il = clip.latents_with_grad(bank_tokens, volume)
    return vocabfine_loss_from_latents(tl, il, clip.temperature, labels)
