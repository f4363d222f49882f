"""Drop-in import name of the reference package (`from transformer_maskgit import CTViT`, scripts/run_train.py:1).
Only the CT-CLIP image tower is provided; the GenerateCT trainers the reference re-exports are out of scope."""
from ct_clip_b200.ctvit import CTViT  # noqa: F401
