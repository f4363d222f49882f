"""ctypes binding of libctclip_b200.so (the C ABI declared in include/ctclip_b200.h).

The product path has NO fallback: if the shared object is missing or a call fails, an exception
is raised. Nothing here imports the oracle.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libctclip_b200.so"
_lib = None


class CtclipError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("a_major", C.c_int32), ("b_major", C.c_int32),
        ("A", C.c_void_p), ("lda", C.c_int64),
        ("B", C.c_void_p), ("ldb", C.c_int64),
        ("epilogue", C.c_int32), ("splits", C.c_int32),
        ("C", C.c_void_p), ("ldc", C.c_int64),
        ("bias", C.c_void_p),
        ("resid", C.c_void_p), ("ldr", C.c_int64),
        ("C2", C.c_void_p), ("ldc2", C.c_int64),
        ("arg_out", C.c_void_p), ("argval_out", C.c_void_p),
    ]


EPI_BF16, EPI_F32, EPI_RESID_F32, EPI_GEGLU, EPI_ATOMIC_F32, EPI_ARGMAX = range(6)


def lib() -> C.CDLL:
    """Load the shared object (building nothing: use __graft_entry__.build() / build.py first)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise CtclipError(
                f"{LIB_PATH} is missing: run `python -m ct_clip_b200.build` (needs nvcc). "
                "There is no CPU / PyTorch fallback for the hot path.")
        _lib = C.CDLL(str(LIB_PATH))
        _lib.ctclip_version.restype = C.c_int
        _lib.ctclip_last_error.restype = C.c_char_p
        for name in dir(_sigs):
            if name.startswith("ctclip_"):
                fn = getattr(_lib, name)
                fn.restype = C.c_int
                fn.argtypes = getattr(_sigs, name)
    return _lib


class _sigs:
    ctclip_gemm_bf16 = [C.POINTER(GemmArgs), C.c_void_p]


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().ctclip_last_error().decode(errors="replace")
        raise CtclipError(f"{what} failed (rc={rc}): {msg}")
