"""gm_diffusion_b200 — B200-native (sm_100a) Stage-3 hot path of GM-Diffusion.

Mirrors the reference's top-level exports (gm_diffusion/__init__.py:16-24): the three pipelines and the stage-1
tone-mapping functions.  Everything numerical runs in hand-written CUDA behind the C-ABI of include/gmd_b200.h
(`gm_diffusion_b200/_C/libgmd_b200.so`); there is no CPU / PyTorch fallback — calls fail loudly without it.
"""
from .stage1 import (apply_gm_to_sdr, fix_mulog_tmo, gamut_compress, hard_clip_tmo, linear_scale_tmo, random_tmo_cuda,
                     reconstruct_for_disk, reconstruct_hdr, rgbe_encode, tmo_cuda)
from .hdr_io import pack_radiance, save_hdr_image
from .stage1 import RandomExposureAdjust
from .pipelines import (StableDiffusionDualUNetImprovedPipeline, StableDiffusionDualUNetPipeline, StableDiffusionGMPipeline)
from .schedulers import DDIMScheduler, DDPMScheduler, DPMSolverMultistepScheduler, PNDMScheduler
from .unet import B200UNet
from .vae import B200Vae, B200VaeDecoder
from .text_encoder import B200ClipTextEncoder

__version__ = "0.1.0"
__all__ = ["RandomExposureAdjust", "apply_gm_to_sdr", "fix_mulog_tmo", "gamut_compress", "hard_clip_tmo", "linear_scale_tmo", "random_tmo_cuda",
           "tmo_cuda", "reconstruct_hdr", "reconstruct_for_disk", "rgbe_encode", "save_hdr_image", "pack_radiance", "StableDiffusionDualUNetPipeline", "StableDiffusionDualUNetImprovedPipeline",
           "StableDiffusionGMPipeline", "PNDMScheduler", "DDIMScheduler", "DDPMScheduler", "DPMSolverMultistepScheduler", "B200UNet", "B200VaeDecoder", "B200Vae", "B200ClipTextEncoder"]
