"""Drop-in for gm_diffusion/pipelines/stable_diffusion_dual_unet_improved.py.  The reference file is byte-identical to
stable_diffusion_dual_unet.py except for the class name (:156) and a commented-out experiment (:1079-1084), so one
implementation serves both names."""
from .stable_diffusion_dual_unet import StableDiffusionDualUNetPipeline


class StableDiffusionDualUNetImprovedPipeline(StableDiffusionDualUNetPipeline):
    pass
