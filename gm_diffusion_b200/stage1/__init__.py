"""Stage-1 numerics (Eq.(1), tone-mapping operators, gamut) — same names as gm_diffusion/stage1/__init__.py:8-16."""
from .augmentations import RandomExposureAdjust
from .tone_mapping import (
    apply_gm_to_sdr,
    fix_mulog_tmo,
    gamut_compress,
    hard_clip_tmo,
    linear_scale_tmo,
    random_tmo_cuda,
    reconstruct_for_disk,
    reconstruct_hdr,
    rgbe_encode,
    tmo_cuda,
)

__all__ = ["RandomExposureAdjust", "apply_gm_to_sdr", "fix_mulog_tmo", "gamut_compress", "hard_clip_tmo", "linear_scale_tmo",
           "random_tmo_cuda", "tmo_cuda", "reconstruct_hdr", "reconstruct_for_disk", "rgbe_encode"]
