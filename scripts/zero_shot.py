"""Drop-in module name used by the reference entry script (`from zero_shot import CTClipInference`, run_zero_shot.py:4)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ct_clip_b200.inference import CTClipInference  # noqa: E402,F401
