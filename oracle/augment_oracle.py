"""Oracle for the stage-1 augmentation `RandomExposureAdjust` (gm_diffusion/stage1/augmentations.py:13-79): CPU restatement, op for op.
TEST INFRASTRUCTURE ONLY.  PINNED: tests/golden/exposure_reference.npz holds outputs and metadata of the REAL reference class
(executed by file path — it is pure torch) for fixed seeds (oracle/make_golden.py:exposure_golden)."""
from __future__ import annotations

import random

import torch

EXPOSURE_LEVELS = torch.tensor([0.1, 0.25, 0.5, 1.0, 4.0, 8.0, 16.0], dtype=torch.float32)


def hdr_to_ldr(img, exposure, gamma=2.2):                       # :24-26
    return torch.pow(torch.clamp(img * exposure, 0.0, 1.0), 1.0 / gamma)


def sample_camera_curve():                                      # :28-32
    n = float(torch.clamp(torch.normal(mean=0.65, std=0.1, size=()), 0.4, 0.9))
    sigma = float(torch.clamp(torch.normal(mean=0.6, std=0.1, size=()), 0.4, 0.8))
    return n, sigma


def apply_inv_sigmoid_curve(y, n, sigma):                       # :34-36
    return torch.pow((sigma * y) / (1 + sigma - y + 1e-8), 1.0 / n)


def discretize_to_uint16(img):                                  # :38-41
    max_int = 2 ** 16 - 1
    return torch.clamp(img * max_int, 0, max_int).round() / max_int


def random_exposure_adjust(imgs, gamma=2.2, prob=1.0):          # :43-73, returns (ldr, metadata)
    if random.random() > prob:
        return imgs, {"exposure": 1.0, "n": 1.0, "sigma": 0.0}
    exposure = float(EXPOSURE_LEVELS[torch.randint(len(EXPOSURE_LEVELS), (1,))])
    n, sigma = sample_camera_curve()
    x = discretize_to_uint16(apply_inv_sigmoid_curve(imgs, n, sigma))
    return hdr_to_ldr(x, exposure, gamma), {"exposure": exposure, "n": n, "sigma": sigma}
