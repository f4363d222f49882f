"""ct_clip_b200 -- B200-native (sm_100a) implementation of the CT-CLIP contrastive hot path.

Public surface mirrors the reference (ibrahimethemhamamci/CT-CLIP):
    CTViT            transformer_maskgit/transformer_maskgit/ctvit.py:118
    CTCLIP           CT_CLIP/ct_clip/ct_clip.py:407
    CTClipTrainer    scripts/CTCLIPTrainer.py:113
    CTClipInference  scripts/zero_shot.py:53
Heavy imports are lazy so that `import ct_clip_b200` works on a CPU-only box.
"""
__all__ = ["CTViT", "CTCLIP", "CTClipTrainer", "CTClipInference"]


def __getattr__(name):
    if name == "CTViT":
        from .ctvit import CTViT
        return CTViT
    if name == "CTCLIP":
        from .ctclip import CTCLIP
        return CTCLIP
    if name == "CTClipTrainer":
        from .trainer import CTClipTrainer
        return CTClipTrainer
    if name == "CTClipInference":
        from .inference import CTClipInference
        return CTClipInference
    raise AttributeError(name)
