"""Torch-tensor front ends of the C-ABI kernels (raw device pointers + current stream).

Every function launches on torch's current CUDA stream and never synchronises.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import (EPI_ARGMAX, EPI_ATOMIC_F32, EPI_BF16, EPI_F32, EPI_GEGLU, EPI_RESID_F32,  # noqa: F401
                   GemmArgs, check)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def gemm(A: torch.Tensor, B: torch.Tensor, *, M: int, N: int, K: int, a_major: int = 0, b_major: int = 0,
         epilogue: int = EPI_BF16, C_out: torch.Tensor | None = None, bias: torch.Tensor | None = None,
         resid: torch.Tensor | None = None, C2: torch.Tensor | None = None, arg_out: torch.Tensor | None = None,
         argval_out: torch.Tensor | None = None, splits: int = 1, lda: int | None = None, ldb: int | None = None,
         ldc: int | None = None) -> None:
    """C[M,N] = sum_k A(m,k) B(n,k); see include/ctclip_b200.h for the epilogues."""
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16 and A.is_cuda and B.is_cuda
    a = GemmArgs()
    a.M, a.N, a.K = M, N, K
    a.a_major, a.b_major = a_major, b_major
    a.A, a.lda = A.data_ptr(), (lda if lda is not None else A.stride(0))
    a.B, a.ldb = B.data_ptr(), (ldb if ldb is not None else B.stride(0))
    a.epilogue, a.splits = epilogue, splits
    a.C = _ptr(C_out)
    a.ldc = ldc if ldc is not None else (C_out.stride(0) if C_out is not None else 0)
    a.bias = _ptr(bias)
    a.resid = _ptr(resid)
    a.ldr = resid.stride(0) if resid is not None else 0
    a.C2 = _ptr(C2)
    a.ldc2 = C2.stride(0) if C2 is not None else 0
    a.arg_out = _ptr(arg_out)
    a.argval_out = _ptr(argval_out)
    check(_lib.lib().ctclip_gemm_bf16(C.byref(a), _stream()), "ctclip_gemm_bf16")
