"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (ct_clip_b200/*).

Loads the UNMODIFIED reference (ibrahimethemhamamci/CT-CLIP, mounted read-only at
/root/reference in the build container) on CPU so that
  * the oracle restatement (oracle/ctclip_oracle.py) can be validated against it, and
  * golden vectors (tests/golden/*.pt, made by tests/golden/make_golden.py) can be generated.
The reference does not exist on the GPU box; nothing on the `-m gpu` / bench / smoke path uses
this file.

What is shimmed (zero edits to reference sources):
  1. stub modules for packages the reference imports but this image lacks and the contrastive
     path never calls: accelerate, ema_pytorch, nibabel (pulled in by
     transformer_maskgit/__init__.py:1-3 via the GenerateCT trainers);
  2. `vector_quantize_pytorch` (pinned ==1.1.2 in transformer_maskgit/setup.py:19, not vendored):
     replaced by oracle.vq_restated.VectorQuantize -- a restatement of the published algorithm.
     PARITY UNPINNED for this one boundary: the reference holds no test / golden vector for it.
  3. the hard-coded torch.device('cuda') in attention.py:135,171,195,219,260 and
     ctvit.py:110,236,292,374 is redirected to CPU by replacing the module-global name `torch`
     inside those two modules with a delegating proxy.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from pathlib import Path

REFERENCE_ROOT = Path(os.environ.get("CTCLIP_REFERENCE_ROOT", "/root/reference"))


def reference_available() -> bool:
    return (REFERENCE_ROOT / "CT_CLIP" / "ct_clip" / "ct_clip.py").exists()


class _TorchCpuProxy:
    """Delegates to torch, except device('cuda') -> device('cpu')."""

    def __init__(self, torch_mod):
        object.__setattr__(self, "_t", torch_mod)

    def __getattr__(self, name):
        return getattr(object.__getattribute__(self, "_t"), name)

    def device(self, *args, **kwargs):
        t = object.__getattribute__(self, "_t")
        if args and isinstance(args[0], str) and args[0].startswith("cuda"):
            return t.device("cpu")
        return t.device(*args, **kwargs)


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs() -> None:
    import torch
    import transformers  # noqa: F401  (must be imported before the accelerate stub exists)

    class _Dummy:
        def __init__(self, *a, **k):
            pass

    if "accelerate" not in sys.modules:
        acc = _stub("accelerate", Accelerator=_Dummy, DistributedDataParallelKwargs=_Dummy,
                    DistributedType=_Dummy)
        acc.utils = _stub("accelerate.utils", InitProcessGroupKwargs=_Dummy)
        acc.__path__ = []
    if "ema_pytorch" not in sys.modules:
        _stub("ema_pytorch", EMA=_Dummy)
    if "nibabel" not in sys.modules:
        _stub("nibabel", load=lambda *a, **k: (_ for _ in ()).throw(RuntimeError("nibabel stub")))
    from oracle import vq_restated
    _stub("vector_quantize_pytorch", VectorQuantize=vq_restated.VectorQuantize)
    del torch


_loaded = None


def load_reference():
    """Returns (CTViT, CTCLIP) classes of the unmodified reference, runnable on CPU."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    os.environ.setdefault("HF_HUB_OFFLINE", "1")
    os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")
    _install_stubs()
    # Our own drop-in shims live at the repo root under the same import names
    # (transformer_maskgit/, ct_clip/); make sure the REFERENCE packages win here.
    for name in list(sys.modules):
        if name.split(".")[0] in ("transformer_maskgit", "ct_clip"):
            del sys.modules[name]
    paths = [str(REFERENCE_ROOT / "transformer_maskgit"), str(REFERENCE_ROOT / "CT_CLIP")]
    saved = list(sys.path)
    sys.path[:0] = paths
    try:
        import torch
        tm = importlib.import_module("transformer_maskgit")
        att = importlib.import_module("transformer_maskgit.attention")
        ctv = importlib.import_module("transformer_maskgit.ctvit")
        proxy = _TorchCpuProxy(torch)
        att.torch = proxy
        ctv.torch = proxy
        cc = importlib.import_module("ct_clip")
        ref = (ctv.CTViT, cc.CTCLIP)
    finally:
        sys.path[:] = saved
    # keep the reference modules under private aliases, and free the public names again so that
    # later `import transformer_maskgit` resolves to the repo's drop-in package.
    for name in list(sys.modules):
        root = name.split(".")[0]
        if root in ("transformer_maskgit", "ct_clip"):
            sys.modules["_reference_" + name] = sys.modules.pop(name)
    del tm
    _loaded = ref
    return ref


def load_reference_dataset_class():
    """The unmodified reference dataset class (scripts/data.py: CTReportDataset) with a nibabel stub whose `load` the caller
    replaces: used to pin oracle.ctclip_oracle.ct_preprocess against data.py:92-162 (nii_img_to_tensor)."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    import importlib.util
    if "nibabel" not in sys.modules:
        _stub("nibabel", load=lambda *a, **k: (_ for _ in ()).throw(RuntimeError("nibabel stub")))
    if "tqdm" not in sys.modules:
        try:
            import tqdm  # noqa: F401
        except ImportError:
            _stub("tqdm", tqdm=lambda x, **k: x)
    spec = importlib.util.spec_from_file_location("_reference_scripts_data", str(REFERENCE_ROOT / "scripts" / "data.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
