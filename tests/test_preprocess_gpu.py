"""GPU input pipeline: ctclip_ct_preprocess vs the oracle restatement of scripts/data.py:92-162."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,xy,z,target", [
    ((64, 64, 40), 0.9, 2.0, (48, 48, 24)),        # up-sampled in every axis, then cropped
    ((40, 44, 30), 0.6, 1.2, (48, 48, 32)),        # down-sampled, then padded with -1
    ((50, 38, 21), 0.75, 3.1, (32, 48, 40)),       # crop one axis, pad another, z stretched 2.07x
])
@pytest.mark.parametrize("raw_dtype", [torch.int16, torch.float32])
def test_ct_preprocess_matches_reference_pipeline(shape, xy, z, target, raw_dtype):
    from ct_clip_b200.preprocess import preprocess_ct
    from oracle import ctclip_oracle as O
    g = torch.Generator().manual_seed(7)
    raw = (torch.randn(*shape, generator=g) * 400 + 900).round().clamp(0, 3000)       # stored CT values (before rescale)
    slope, intercept = 1.0, -1024.0
    ref = O.ct_preprocess(raw.numpy().astype(np.float64), slope, intercept, xy, z, target_shape=target)
    out = preprocess_ct(raw.to(raw_dtype).cuda(), slope=slope, intercept=intercept, xy_spacing=xy, z_spacing=z, target_shape=target)
    torch.cuda.synchronize()
    assert out.shape == ref.shape == (1, target[2], target[0], target[1])
    assert (out.cpu() - ref).abs().max().item() < 2e-5          # fp32 interpolation vs the reference's float64
    out16 = preprocess_ct(raw.to(raw_dtype).cuda(), slope=slope, intercept=intercept, xy_spacing=xy, z_spacing=z, target_shape=target,
                          out_dtype=torch.int16)
    assert (out16.cpu().float() / 1000.0 - ref).abs().max().item() <= 5.1e-4      # int16 HU: half a Hounsfield unit
    assert (out16.cpu()[ref == -1.0] == -1000).all()              # padding = -1 = -1000 HU
