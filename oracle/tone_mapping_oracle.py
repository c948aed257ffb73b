"""Oracle for kernel (d): restatement of gm_diffusion/stage1/tone_mapping.py (reference lines cited per
function) in plain torch on the CPU, plus the numpy Eq.(1) twin the inference scripts really execute
(scripts/inference/experiments/formal_improved.py:34-45).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import importlib.util
import math
import os

import numpy as np
import torch

REFERENCE_TM = "/root/reference/gm_diffusion/stage1/tone_mapping.py"

# tone_mapping.py:80-82
BT2020_TO_709 = ((1.660491, -0.587641, -0.072850),
                 (-0.124550, 1.132900, -0.008349),
                 (-0.018151, -0.100579, 1.118730))


def linear_scale_tmo(img: torch.Tensor, qmax: float) -> torch.Tensor:
    """tone_mapping.py:14-18."""
    return img / (qmax + 1)


def hard_clip_tmo(hdr_img: torch.Tensor, qmax: float) -> torch.Tensor:
    """tone_mapping.py:21-26 (qmax ignored)."""
    return torch.clamp(hdr_img, 0, 1)


def fix_mulog_tmo(hdr_img: torch.Tensor, qmax: float, mu: float = 500) -> torch.Tensor:
    """tone_mapping.py:29-36 (mu fixed to 500 there; random_tmo_cuda :50-57 draws it from U(500, 5000))."""
    hdr_img = hdr_img / (qmax + 1)
    tm = torch.log1p(mu * hdr_img) / math.log1p(mu)
    return torch.clamp(tm, 0, 1)


def tmo_cuda(hdr_img: torch.Tensor) -> torch.Tensor:
    """tone_mapping.py:39-47."""
    hdr_img = torch.clamp(hdr_img / 10, 0, 1)
    if not torch.all((0 <= hdr_img) & (hdr_img <= 1)):
        raise ValueError("HDR image values should be in the range [0, 1]")
    mu = 5_000.0
    return torch.log1p(mu * hdr_img) / math.log1p(mu)


def apply_gm_to_sdr(gm: torch.Tensor, sdr: torch.Tensor, qmax: float = 9, eps: float = 1 / 64) -> torch.Tensor:
    """tone_mapping.py:60-71: Eq.(1) with the output clamp."""
    sdr_linear = torch.clamp(sdr, 0, 1) ** 2.2
    hdr = (sdr_linear + eps) * (1 + gm * qmax) - eps
    return torch.clamp(hdr, 0, qmax + 1)


def apply_gm_to_sdr_numpy(gm: np.ndarray, sdr: np.ndarray, qmax: float = 99, eps: float = 1 / 64) -> np.ndarray:
    """formal_improved.py:34-45 (numpy twin run on the host by the inference scripts; NO output clamp)."""
    sdr = np.clip(sdr, 0, 1)
    sdr_linear = sdr ** 2.2
    return (sdr_linear + eps) * (1 + gm * qmax) - eps


def gamut_compress(tmo_hdr_img: torch.Tensor) -> torch.Tensor:
    """tone_mapping.py:74-90: rgb' = M rgb per pixel, clamp(0,1); input [B,3,H,W]."""
    conversion = torch.tensor(BT2020_TO_709, device=tmo_hdr_img.device, dtype=tmo_hdr_img.dtype).t()
    img = tmo_hdr_img.permute(0, 2, 3, 1)
    img = torch.matmul(img, conversion)
    img = img.permute(0, 3, 1, 2)
    return torch.clamp(img, 0, 1)


def denormalize(img: torch.Tensor) -> torch.Tensor:
    """scripts/inference/generate_hdr.py:227,232: (x / 2 + 0.5).clamp(0, 1)."""
    return (img / 2 + 0.5).clamp(0, 1)


def save_hdr_scale(hdr: torch.Tensor, qmax: float) -> torch.Tensor:
    """scripts/inference/generate_hdr.py:27-30 numerics (divide by qmax+1, RGB->BGR); the file write stays OpenCV."""
    return (hdr / (qmax + 1)).flip(-1)


def load_reference_tm():
    """Execute the REAL reference file by path (pure torch; the package import needs diffusers).  Only
    available in the build container; returns None elsewhere (e.g. on the GPU box)."""
    if not os.path.exists(REFERENCE_TM):
        return None
    spec = importlib.util.spec_from_file_location("ref_tone_mapping", REFERENCE_TM)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
