"""Oracle for the host tail: the Radiance RGBE quantisation `cv2.imwrite("x.hdr", bgr)` applies inside the
reference's save_hdr_image (scripts/inference/generate_hdr.py:27-30) and the uint8 conversion of :243-244.
TEST INFRASTRUCTURE ONLY.

The encoder lives in a third-party dependency (OpenCV's HdrEncoder, i.e. Greg Ward's published float2rgbe; opencv-python is
unpinned in the reference).  PINNED: tests/golden/rgbe_cv2.npz holds RGBE bytes parsed out of files written by the real
cv2.imwrite (4.13.0) in this container (oracle/make_golden.py), and tests/test_cpu_oracle.py checks this restatement against
them bit for bit."""
from __future__ import annotations

import numpy as np


def float2rgbe(rgb: np.ndarray) -> np.ndarray:
    """float32 [...,3] (R,G,B) -> uint8 [...,4] (R,G,B,E).  Ward: v = max; v < 1e-32 -> 0; else (m,e) = frexp(v);
    s = m*256/v; bytes = (uchar)(c*s), e+128.  The C comparison promotes v to double."""
    rgb = np.asarray(rgb, np.float32)
    v = rgb.max(-1)
    m, e = np.frexp(v)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        s = (m.astype(np.float64) * 256.0 / v.astype(np.float64)).astype(np.float32)
        q = rgb * s[..., None]
    out = np.zeros(rgb.shape[:-1] + (4,), np.uint8)
    out[..., :3] = np.clip(np.nan_to_num(q, nan=0.0, posinf=255.0, neginf=0.0).astype(np.int64), 0, 255)
    out[..., 3] = (e + 128).astype(np.int64) & 255
    out[v.astype(np.float64) < 1e-32] = 0
    return out


def rgbe2float(rgbe: np.ndarray) -> np.ndarray:
    """OpenCV's rgbe2float: f = ldexp(1, E - 136); c * f (no +0.5); E == 0 -> 0."""
    rgbe = np.asarray(rgbe, np.uint8)
    f = np.ldexp(np.float32(1.0), rgbe[..., 3].astype(np.int32) - 136).astype(np.float32)
    out = rgbe[..., :3].astype(np.float32) * f[..., None]
    out[rgbe[..., 3] == 0] = 0
    return out


def quantize_u8(x: np.ndarray) -> np.ndarray:
    """generate_hdr.py:243-244: (x * 255).astype(np.uint8) for x in [0,1] (float32 product, truncation)."""
    return np.clip((np.asarray(x, np.float32) * np.float32(255.0)).astype(np.int64), 0, 255).astype(np.uint8)


def save_hdr_pixels(apply_HDR: np.ndarray, qmax: float) -> np.ndarray:
    """The bytes save_hdr_image puts in the file for an RGB [H,W,3] array: float2rgbe((hdr / (qmax+1)).astype(float32))."""
    return float2rgbe((np.asarray(apply_HDR) / (qmax + 1)).astype(np.float32))


def parse_radiance(raw: bytes) -> np.ndarray:
    """Radiance .hdr file bytes -> uint8 [H,W,4]; understands flat and new-RLE scanlines (literal and repeat runs)."""
    pos = raw.index(b"\n\n") + 2
    end = raw.index(b"\n", pos)
    tok = raw[pos:end].split()
    assert tok[0] == b"-Y" and tok[2] == b"+X", tok
    H, W = int(tok[1]), int(tok[3])
    data = np.frombuffer(raw[end + 1:], np.uint8)
    out = np.zeros((H, W, 4), np.uint8)
    p = 0
    for y in range(H):
        if 8 <= W < 32768 and data[p] == 2 and data[p + 1] == 2 and ((int(data[p + 2]) << 8) | int(data[p + 3])) == W:
            p += 4
            for c in range(4):
                x = 0
                while x < W:
                    n = int(data[p]); p += 1
                    if n > 128:
                        n -= 128
                        out[y, x:x + n, c] = data[p]; p += 1
                    else:
                        out[y, x:x + n, c] = data[p:p + n]; p += n
                    x += n
        else:
            out[y] = data[p:p + 4 * W].reshape(W, 4); p += 4 * W
    assert p == len(data), (p, len(data))
    return out
