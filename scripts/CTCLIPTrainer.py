"""Drop-in module name used by the reference entry script (`from CTCLIPTrainer import CTClipTrainer`, run_train.py:4)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ct_clip_b200.trainer import CTClipTrainer  # noqa: E402,F401
