"""Host tail of the inference scripts: `save_hdr_image` with the same signature as
scripts/inference/generate_hdr.py:27-30, but the `/ (qmax+1)` division and the Radiance RGBE quantisation run on the GPU
(csrc/hdr.cu `rgbe_pack`), the device->host copy is 4 B/px, and the container is written here — no OpenCV on the path.

A file written here decodes (with cv2.imread or any Radiance reader) to exactly the floats a file written by the reference's
`cv2.imwrite` decodes to, because the RGBE bytes are identical; only the lossless scanline packing differs (literal runs
instead of OpenCV's run-length search)."""
from __future__ import annotations

import os
from typing import Union

import numpy as np
import torch

from .stage1.tone_mapping import rgbe_encode

_HEADER = b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n"


def pack_radiance(rgbe: np.ndarray) -> bytes:
    """uint8 [H,W,4] (R,G,B,E) -> Radiance .hdr file bytes.  Scanlines use the "new RLE" container (2,2,W_hi,W_lo, then the four
    byte planes) with literal runs only, which every reader accepts and which — unlike flat pixels — cannot be mistaken for
    a run-length header; widths outside [8, 32767] must be flat by the format's own rule."""
    if rgbe.ndim != 3 or rgbe.shape[2] != 4 or rgbe.dtype != np.uint8:
        raise ValueError(f"pack_radiance expects uint8 [H,W,4], got {rgbe.dtype} {rgbe.shape}")
    H, W, _ = rgbe.shape
    head = _HEADER + f"-Y {H} +X {W}\n".encode()
    if W < 8 or W > 32767:
        return head + rgbe.tobytes()
    planes = np.ascontiguousarray(rgbe.transpose(0, 2, 1))                    # [H,4,W]
    full, tail = divmod(W, 128)
    parts = []
    if full:
        body = planes[:, :, :full * 128].reshape(H, 4, full, 128)
        cnt = np.full((H, 4, full, 1), 128, np.uint8)
        parts.append(np.concatenate([cnt, body], axis=3).reshape(H, 4, full * 129))
    if tail:
        cnt = np.full((H, 4, 1), tail, np.uint8)
        parts.append(np.concatenate([cnt, planes[:, :, full * 128:]], axis=2))
    lines = np.concatenate(parts, axis=2).reshape(H, -1)
    marker = np.empty((H, 4), np.uint8)
    marker[:] = (2, 2, W >> 8, W & 255)
    return head + np.concatenate([marker, lines], axis=1).tobytes()


def save_hdr_image(apply_HDR: Union[torch.Tensor, np.ndarray], save_dir: str, filename: str, qmax: float) -> str:
    """generate_hdr.py:27-30: `cv2.imwrite(join(save_dir, filename), (apply_HDR / (qmax+1)).astype(float32)[:, :, [2,1,0]])`.
    `apply_HDR` is the RGB HDR image [H,W,3] (CUDA tensor, or a numpy array which is uploaded first)."""
    if not str(filename).lower().endswith(".hdr"):
        raise ValueError("save_hdr_image writes Radiance .hdr files (the only format the reference scripts use)")
    t = torch.as_tensor(apply_HDR)
    if t.dim() != 3 or t.shape[-1] != 3:
        raise ValueError(f"save_hdr_image expects [H,W,3], got {tuple(t.shape)}")
    if not t.is_cuda:
        t = t.to("cuda")
    rgbe = rgbe_encode(t.to(torch.float32), float(qmax) + 1.0, channels_last=True)
    path = os.path.join(save_dir, f"{filename}")
    with open(path, "wb") as f:
        f.write(pack_radiance(rgbe.cpu().numpy()))
    return path


__all__ = ["save_hdr_image", "pack_radiance"]
