"""Pipelines — same exports as gm_diffusion/pipelines/__init__.py:5-19."""
from .stable_diffusion_dual_unet import StableDiffusionDualUNetPipeline
from .stable_diffusion_dual_unet_improved import StableDiffusionDualUNetImprovedPipeline
from .stable_diffusion_gm import StableDiffusionGMPipeline

__all__ = ["StableDiffusionDualUNetPipeline", "StableDiffusionDualUNetImprovedPipeline", "StableDiffusionGMPipeline"]
