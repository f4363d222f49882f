"""Shared helpers for the parity tests (tests only: the oracle is imported here, never by the product)."""
import torch

from oracle import ctclip_oracle as O

CFG1_VIT = dict(dim=512, codebook_size=8192, image_size=64, patch_size=16, temporal_patch_size=8, spatial_depth=4,
                temporal_depth=4, dim_head=32, heads=8)


def rel_err(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()


def rms_err(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt().clamp_min(1e-12)).item()


def oracle_vit_cfg(kw):
    return O.CTViTConfig(dim=kw["dim"], codebook_size=kw["codebook_size"], image_size=kw["image_size"],
                         patch_size=kw["patch_size"], temporal_patch_size=kw["temporal_patch_size"],
                         spatial_depth=kw["spatial_depth"], temporal_depth=kw["temporal_depth"],
                         dim_head=kw["dim_head"], heads=kw["heads"])


def temporal_to_canonical(x, b, h, w):
    """oracle temporal-stack tensors are ((b h w), t, d); canonical is (b, t, h, w, d)."""
    t, d = x.shape[1], x.shape[2]
    return x.reshape(b, h, w, t, d).permute(0, 3, 1, 2, 4)
