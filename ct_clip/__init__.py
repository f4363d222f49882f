"""Drop-in import name of the reference package (`from ct_clip import CTCLIP`, scripts/run_train.py:3)."""
from ct_clip_b200.ctclip import CTCLIP  # noqa: F401
