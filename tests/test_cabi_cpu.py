"""CPU suite part 2: the C-ABI shared library loads without a GPU/driver and exports exactly what
include/ctclip_b200.h declares; argument errors are reported through the ABI (no compute is launched here)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "ctclip_b200.h"


def _declared():
    txt = HEADER.read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ctclip_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    from ct_clip_b200 import _lib
    from ct_clip_b200.build import build_lib
    build_lib()
    return _lib.lib()


def test_library_exports_every_declared_symbol(lib):
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in ctclip_b200.h but not exported"


def test_python_binding_covers_header(lib):
    from ct_clip_b200 import _lib
    declared = set(_declared()) - {"ctclip_version", "ctclip_last_error"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_version_and_error_reporting(lib):
    from ct_clip_b200 import _lib
    assert lib.ctclip_version() == 100
    a = _lib.GemmArgs()           # all-zero arguments: rejected before any CUDA call
    rc = lib.ctclip_gemm_bf16(C.byref(a), None)
    assert rc != 0
    assert b"gemm" in lib.ctclip_last_error()
    with pytest.raises(_lib.CtclipError):
        _lib.call("ctclip_ln_fwd", C.byref(_lib.LnFwdArgs()), None)


def test_struct_layouts_match_c(lib):
    """ctypes mirrors must have the C struct sizes (compiled probe)."""
    import subprocess
    import tempfile

    from ct_clip_b200 import _lib
    src = '#include <stdio.h>\n#include "ctclip_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",' \
          'sizeof(ctclip_gemm_args),sizeof(ctclip_ln_fwd_args),sizeof(ctclip_ln_bwd_args),sizeof(ctclip_patchify_args),' \
          'sizeof(ctclip_peg_args),sizeof(ctclip_attn_args),sizeof(ctclip_sgemm_args),sizeof(ctclip_loss_args));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        (Path(d) / "p.c").write_text(src)
        subprocess.run(["gcc", "-I", str(ROOT / "include"), str(Path(d) / "p.c"), "-o", str(Path(d) / "p")], check=True)
        out = subprocess.run([str(Path(d) / "p")], capture_output=True, text=True, check=True).stdout.split()
    py = [C.sizeof(s) for s in (_lib.GemmArgs, _lib.LnFwdArgs, _lib.LnBwdArgs, _lib.PatchifyArgs, _lib.PegArgs, _lib.AttnArgs,
                                _lib.SgemmArgs, _lib.LossArgs)]
    assert [int(x) for x in out] == py
