"""Drop-in for gm_diffusion/stage1/augmentations.py: `RandomExposureAdjust` with the same constructor, attributes, helper
methods, random-number consumption (Python `random`, then torch's global CPU generator: randint, normal, normal) and errors,
but the ten elementwise torch ops of `__call__` (:63-65) run as ONE launch of `gmd_exposure_adjust` (csrc/hdr.cu) on CUDA
tensors.  No CPU fallback."""
from __future__ import annotations

import random
from typing import Dict, Tuple, Union

import torch

from .. import _lib as L

_CURVE, _QUANT, _EXPOSE = 1, 2, 4


def _launch(img: torch.Tensor, stages: int, n: float = 1.0, sigma: float = 0.0, exposure: float = 1.0, gamma: float = 1.0) -> torch.Tensor:
    L.require_cuda(img)
    x = img.to(torch.float32).contiguous()
    out = torch.empty_like(x)
    L.check(L.lib().gmd_exposure_adjust(x.data_ptr(), out.data_ptr(), x.numel(), stages, float(n), float(sigma), float(exposure), float(gamma),
                                        L.current_stream()), "gmd_exposure_adjust")
    return out


class RandomExposureAdjust:
    """augmentations.py:13-79."""

    def __init__(self, gamma: float = 2.2, prob: float = 1.0):
        self.gamma = gamma
        self.prob = prob
        self.exposure_levels = torch.tensor([0.1, 0.25, 0.5, 1.0, 4.0, 8.0, 16.0], dtype=torch.float32)

    def hdr_to_ldr(self, img: torch.Tensor, exposure: float) -> torch.Tensor:
        return _launch(img, _EXPOSE, exposure=exposure, gamma=self.gamma)

    @staticmethod
    def sample_camera_curve() -> Tuple[float, float]:
        n = float(torch.clamp(torch.normal(mean=0.65, std=0.1, size=()), 0.4, 0.9))
        sigma = float(torch.clamp(torch.normal(mean=0.6, std=0.1, size=()), 0.4, 0.8))
        return n, sigma

    @staticmethod
    def apply_inv_sigmoid_curve(y: torch.Tensor, n: float, sigma: float) -> torch.Tensor:
        return _launch(y, _CURVE, n=n, sigma=sigma)

    @staticmethod
    def discretize_to_uint16(img: torch.Tensor) -> torch.Tensor:
        return _launch(img, _QUANT)

    def __call__(self, imgs: torch.Tensor, *, return_metadata: bool = False) -> Union[torch.Tensor, Tuple[torch.Tensor, Dict[str, float]]]:
        if random.random() > self.prob:
            return (imgs, {"exposure": 1.0, "n": 1.0, "sigma": 0.0}) if return_metadata else imgs
        exposure = float(self.exposure_levels[torch.randint(len(self.exposure_levels), (1,))])
        n, sigma = self.sample_camera_curve()
        if imgs.dim() not in (3, 4):
            raise ValueError("RandomExposureAdjust expects a tensor with shape (C,H,W) or (N,C,H,W)")
        if imgs.dtype != torch.float32:
            raise TypeError(f"RandomExposureAdjust expects float32 tensors, received {imgs.dtype}")
        ldr_img = _launch(imgs, _CURVE | _QUANT | _EXPOSE, n=n, sigma=sigma, exposure=exposure, gamma=self.gamma)
        if return_metadata:
            return ldr_img, {"exposure": exposure, "n": n, "sigma": sigma}
        return ldr_img

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(gamma={self.gamma}, prob={self.prob}, exposure_levels={self.exposure_levels.tolist()})"


__all__ = ["RandomExposureAdjust"]
