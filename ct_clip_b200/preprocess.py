"""GPU input pipeline: scripts/data.py:92-162 `nii_img_to_tensor` on the device.

The NIfTI container itself (gzip + header) is still decoded on the host (nibabel in the reference); everything after
`nii_img.get_fdata()` -- rescale to HU, trilinear resampling to 0.75 x 0.75 x 1.5 mm, clipping, scaling, centre crop / pad,
axis permutation -- is one kernel launch per volume (`ctclip_ct_preprocess`). Feeding the raw int16 voxels keeps the H2D copy at
the size of the scan (a 512x512x300 chest CT: 157 MB) and the output can be produced directly as int16 HU, the layout the
patch-embed kernel reads (half the bytes of the reference's fp32 volume)."""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import PreprocessArgs, call

TARGET_SHAPE = (480, 480, 240)      # (h, w, d) of data.py:127
TARGET_SPACING = (0.75, 1.5)        # (xy, z) of data.py:103-105


def preprocess_ct(raw: torch.Tensor, *, slope: float, intercept: float, xy_spacing: float, z_spacing: float,
                  target_shape=TARGET_SHAPE, target_spacing=TARGET_SPACING, out_dtype=torch.float32, out: torch.Tensor | None = None,
                  pad_value: float = -1.0) -> torch.Tensor:
    """raw: CUDA tensor [X, Y, Z] (float32 or int16), the array nibabel's get_fdata() returns, before slope / intercept.
    Returns the (1, D, H, W) volume of the dataset contract (fp32 in [-1, 1], or int16 HU when out_dtype=torch.int16)."""
    if not raw.is_cuda:
        raise RuntimeError("preprocess_ct runs on the GPU: move the raw voxel array to the device first (there is no CPU path)")
    if raw.dtype not in (torch.float32, torch.int16):
        raise TypeError(f"raw voxels must be float32 or int16, got {raw.dtype}")
    if out_dtype not in (torch.float32, torch.int16):
        raise TypeError("out_dtype must be torch.float32 or torch.int16")
    raw = raw.contiguous()
    X, Y, Z = raw.shape
    h, w, d = target_shape
    if out is None:
        out = torch.empty(1, d, h, w, dtype=out_dtype, device=raw.device)
    assert out.shape == (1, d, h, w) and out.dtype == out_dtype and out.is_contiguous()
    a = PreprocessArgs()
    a.raw, a.raw_dtype, a.X, a.Y, a.Z = raw.data_ptr(), int(raw.dtype == torch.int16), X, Y, Z
    a.slope, a.intercept, a.xy_spacing, a.z_spacing = slope, intercept, xy_spacing, z_spacing
    a.target_xy, a.target_z = target_spacing
    a.out_d, a.out_h, a.out_w = d, h, w
    a.out, a.out_dtype, a.pad_value = out.data_ptr(), int(out_dtype == torch.int16), pad_value
    call("ctclip_ct_preprocess", C.byref(a), torch.cuda.current_stream().cuda_stream, tag=str(out_dtype).replace("torch.", ""),
         work=("B", float(raw.numel() * raw.element_size() + out.numel() * out.element_size())))
    return out
