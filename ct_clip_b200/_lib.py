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

P = C.c_void_p
I32 = C.c_int32
I64 = C.c_int64
F32 = C.c_float
U64 = C.c_uint64


class CtclipError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("M", I32), ("N", I32), ("K", I32), ("a_major", I32), ("b_major", I32),
        ("A", P), ("lda", I64), ("B", P), ("ldb", I64),
        ("epilogue", I32), ("splits", I32),
        ("C", P), ("ldc", I64), ("bias", P), ("resid", P), ("ldr", I64),
        ("C2", P), ("ldc2", I64), ("arg_out", P), ("argval_out", P),
        ("norm_cols", I32), ("norm_scale", P), ("colsum", P), ("arg2_out", P),
    ]


class LnFwdArgs(C.Structure):
    _fields_ = [("x", P), ("M", I64), ("D", I32), ("eps", F32), ("gamma", P), ("beta", P),
                ("xhat_bf16", P), ("raw_bf16", P), ("y_f32", P), ("y_bf16", P), ("rstd_out", P)]


class LnBwdArgs(C.Structure):
    _fields_ = [("M", I64), ("D", I32), ("g_f32", P), ("g_bf16", P), ("gamma", P), ("xhat", P), ("rstd", P),
                ("dres_in", P), ("add_bf16", P), ("dx_f32", P), ("dx_bf16", P), ("dgamma", P), ("dbeta", P)]


class PatchifyArgs(C.Structure):
    _fields_ = [("video", P), ("dtype", I32), ("scale", F32), ("B", I32), ("C", I32), ("F", I32), ("H", I32),
                ("W", I32), ("pt", I32), ("p1", I32), ("p2", I32), ("eps", F32), ("xhat", P), ("ld_out", I64)]


class PegArgs(C.Structure):
    _fields_ = [("x", P), ("dy", P), ("y", P), ("y_bf16", P), ("weight", P), ("bias", P), ("dweight", P),
                ("dbias", P), ("B", I32), ("T", I32), ("H", I32), ("W", I32), ("D", I32), ("temporal", I32),
                ("lines", I32), ("canon_table", P)]


class AttnArgs(C.Structure):
    _fields_ = [("q", P), ("ldq", I64), ("k", P), ("ldk", I64), ("v", P), ("ldv", I64), ("o", P), ("ldo", I64),
                ("lse", P), ("bias", P), ("bias_t", P),
                ("n", I32), ("heads", I32), ("dim_head", I32), ("num_seqs", I32), ("seq_inner", I32),
                ("seq_outer_stride", I64), ("tok_stride", I64), ("scale", F32),
                ("d_o", P), ("delta", P), ("dq", P), ("ld_dq", I64), ("dk", P), ("ld_dk", I64),
                ("dv", P), ("ld_dv", I64), ("dbias", P), ("total_rows", I64), ("key_mask", P),
                ("bias_frag", P), ("bias_t_frag", P), ("ds_scratch", P),
                ("cpb_table", P), ("grid_h", I32), ("grid_w", I32), ("qk_bound", P), ("dcpb_table", P),
                ("dropout_p", F32), ("dropout_seed", U64), ("dropout_offset", U64)]


class PreprocessArgs(C.Structure):
    _fields_ = [("raw", P), ("raw_dtype", I32), ("X", I32), ("Y", I32), ("Z", I32), ("slope", F32), ("intercept", F32),
                ("xy_spacing", C.c_double), ("z_spacing", C.c_double), ("target_xy", C.c_double), ("target_z", C.c_double),
                ("out_d", I32), ("out_h", I32), ("out_w", I32), ("out", P), ("out_dtype", I32), ("pad_value", F32)]


class SgemmArgs(C.Structure):
    _fields_ = [("M", I32), ("N", I32), ("K", I32), ("A", P), ("lda", I64), ("trans_a", I32),
                ("B", P), ("ldb", I64), ("trans_b", I32), ("C", P), ("ldc", I64), ("bias", P), ("act", I32),
                ("mask_ref", P), ("ld_mask", I64), ("accumulate", I32)]


class PrepDesc(C.Structure):
    _fields_ = [("W", P), ("ldw", I64), ("gamma", P), ("beta", P), ("bias_in", P), ("rowmap", P), ("out", P),
                ("K", I32), ("Np", I32), ("Kp", I32), ("kind", I32)]


class LossArgs(C.Structure):
    _fields_ = [("t_raw", P), ("i_raw", P), ("B", I32), ("L", I32), ("temperature", P), ("t_hat", P), ("i_hat", P),
                ("inv_norm", P), ("sim", P), ("loss", P), ("dtemperature", P), ("d_t_raw", P), ("d_i_raw", P),
                ("row0", I32), ("nrows", I32), ("loss_scale", F32)]


EPI_BF16, EPI_F32, EPI_RESID_F32, EPI_GEGLU, EPI_ATOMIC_F32, EPI_ARGMAX, EPI_L2NORM, EPI_BIAS_GELU, EPI_GEGLU_BWD = range(9)

# name -> argtypes (every entry point returns int and takes the stream last)
SIGNATURES = {
    "ctclip_gemm_bf16": [C.POINTER(GemmArgs), P],
    "ctclip_ln_fwd": [C.POINTER(LnFwdArgs), P],
    "ctclip_ln_bwd": [C.POINTER(LnBwdArgs), P],
    "ctclip_patchify": [C.POINTER(PatchifyArgs), P],
    "ctclip_peg_fwd": [C.POINTER(PegArgs), P],
    "ctclip_peg_bwd_data": [C.POINTER(PegArgs), P],
    "ctclip_peg_bwd_weight": [C.POINTER(PegArgs), P],
    "ctclip_debug_set_peg_variant": [I32],
    "ctclip_attn_fwd": [C.POINTER(AttnArgs), P],
    "ctclip_attn_bwd": [C.POINTER(AttnArgs), P],
    "ctclip_attn_tc_supported": [I32, I32, I32, I32],
    "ctclip_qk_bound": [P, P, I32, P, P],
    "ctclip_debug_set_attn_bwd_warps": [I32],
    "ctclip_l2norm_bwd": [P, I64, P, I64, P, P, I64, P, I64, I32, I32, P],
    "ctclip_sgemm_f32": [C.POINTER(SgemmArgs), P],
    "ctclip_colsum": [P, I32, I64, I64, I32, P, P],
    "ctclip_cast_f32_bf16": [P, P, I64, P],
    "ctclip_cpb_inputs": [P, I32, I32, P],
    "ctclip_cpb_expand": [P, I32, I32, I32, P, P, P],
    "ctclip_cpb_reduce": [P, I32, I32, I32, P, P],
    "ctclip_cpb_reduce_t": [P, I32, I32, I32, P, P],
    "ctclip_cpb_expand_frag": [P, I32, I32, I32, P, P, P],
    "ctclip_geglu_bwd": [P, I64, P, I64, I64, I32, P, P],
    "ctclip_l2norm_rows_bf16": [P, P, I32, I32, P],
    "ctclip_dropout": [P, P, P, P, I64, F32, U64, U64, P],
    "ctclip_ct_preprocess": [C.POINTER(PreprocessArgs), P],
    "ctclip_topk_rows": [P, I64, I32, I32, I32, P, P, P],
    "ctclip_l2norm_rows_f32": [P, P, I32, I32, P],
    "ctclip_vq_rerank": [P, P, P, P, I64, I32, P],
    "ctclip_vq_gather": [P, P, P, I64, I32, P],
    "ctclip_vq_gather_pool": [P, P, I32, I32, I32, I32, P, P, P],
    "ctclip_pool_bwd": [P, I32, I32, I32, I32, P, P],
    "ctclip_vq_ema_accum": [P, P, I64, I32, P, P, P],
    "ctclip_vq_ema_update": [P, P, P, P, I32, I32, F32, P],
    "ctclip_prep_weight": [P, I64, I32, P, P, I32, I32, P, P],
    "ctclip_prep_bias": [P, I64, I32, P, P, P, I32, P, P],
    "ctclip_prep_batched": [P, I32, I32, P],
    "ctclip_unprep_wgrad": [P, I64, P, I64, I32, P, P, I32, P, P, P, P, P, P],
    "ctclip_clip_loss": [C.POINTER(LossArgs), P],
    "ctclip_latent_exchange": [P, P, I32, I32, I32, I32, P, P, C.c_uint32, P],
    "ctclip_clip_sims": [P, I32, P, I32, I32, P, P, P],
    "ctclip_grad_sumsq": [P, I64, P, P],
    "ctclip_adam_step": [P, P, P, P, I64, F32, F32, F32, F32, I32, F32, P, F32, F32, I64, P],
    "ctclip_bert_embed": [P, P, P, P, P, I64, I32, I32, P],
    "ctclip_bert_embed_bwd": [P, P, P, P, I64, I32, I32, P],
    "ctclip_gelu_bwd": [P, I64, P, I64, I64, I32, P, P],
    "ctclip_zero_shot_probs": [P, P, I32, I32, I32, P, P, P],
}

# launches of our own kernels issued through this binding (bench.py reports it as gpu_launches)
launch_count = 0


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
        for name, argtypes in SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype = C.c_int
            fn.argtypes = argtypes
    return _lib


# bench.py sets this to a list to collect (entry point, tag, work, start_event, end_event) per call (per-stage roofline
# table); work = ("B", algorithmic bytes) | ("F", algorithmic flops) | None. Never enabled inside a timed region.
STAGE_TIMER = None
# optional predicate (name, tag) -> bool: record only matching calls (bench.py times the dominant stage + the GEMM family INSIDE
# its timed region with it; None = record every call)
STAGE_FILTER = None


def call(name: str, *args, tag=None, work=None) -> None:
    """Invoke an entry point; raise CtclipError with the library's message on failure."""
    global launch_count
    if STAGE_TIMER is not None and (STAGE_FILTER is None or STAGE_FILTER(name, tag)):
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib(), name)(*args)
        e1.record()
        STAGE_TIMER.append((name, tag, work, e0, e1))
    else:
        rc = getattr(lib(), name)(*args)
    launch_count += 1
    if rc != 0:
        msg = lib().ctclip_last_error().decode(errors="replace")
        raise CtclipError(f"{name} failed (rc={rc}): {msg}")


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().ctclip_last_error().decode(errors="replace")
        raise CtclipError(f"{what} failed (rc={rc}): {msg}")
