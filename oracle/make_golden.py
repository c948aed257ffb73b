"""Generate tests/golden/tm_*.npz by running the REAL reference gm_diffusion/stage1/tone_mapping.py
(executed by file path in the build container: it is pure torch).  The vectors pin oracle/tone_mapping_oracle.py
and kernel (d).  Run: python -m oracle.make_golden"""
from __future__ import annotations

import math
from pathlib import Path

import numpy as np
import torch

from . import tone_mapping_oracle as tmo

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def main():
    ref = tmo.load_reference_tm()
    if ref is None:
        raise SystemExit("reference not present; golden vectors can only be generated in the build container")
    OUT.mkdir(parents=True, exist_ok=True)
    g = torch.Generator().manual_seed(0)
    # domain: sdr, gm in [0,1] (train_vqgan_lora.py:1129-1131), plus out-of-range / edge values
    sdr = torch.rand(2, 3, 24, 20, generator=g)
    gm = torch.rand(2, 3, 24, 20, generator=g)
    edge = torch.tensor([0.0, 1.0, 0.5, 1e-3, 1e-6, 0.999999, -0.25, 1.25])
    sdr.view(-1)[: edge.numel()] = edge
    gm.view(-1)[: edge.numel()] = edge.flip(0)
    out = dict(sdr=sdr.numpy(), gm=gm.numpy())
    for qmax in (9.0, 49.0, 99.0):
        hdr = ref.apply_gm_to_sdr(gm, sdr, qmax=qmax, eps=1 / 64)
        q = int(qmax)
        out[f"hdr_q{q}"] = hdr.numpy()
        out[f"linear_q{q}"] = ref.linear_scale_tmo(hdr, qmax).numpy()
        out[f"hardclip_q{q}"] = ref.hard_clip_tmo(hdr, qmax).numpy()
        out[f"mulog_q{q}"] = ref.fix_mulog_tmo(hdr, qmax).numpy()
        out[f"mulog_gamut_q{q}"] = ref.gamut_compress(ref.fix_mulog_tmo(hdr, qmax)).numpy()
        out[f"tmo_cuda_q{q}"] = ref.tmo_cuda(hdr).numpy()
    out["hdr_default"] = ref.apply_gm_to_sdr(gm, sdr).numpy()  # qmax=9, eps=1/64 defaults (tone_mapping.py:63-64)
    out["gamut_only"] = ref.gamut_compress(sdr).numpy()
    np.savez_compressed(OUT / "tm_reference.npz", **out)
    # the SURVEY §8c spot values (seeded [1,3,4,4])
    torch.manual_seed(0)
    s = torch.rand(1, 3, 4, 4); m = torch.rand(1, 3, 4, 4)
    h = ref.apply_gm_to_sdr(m, s, qmax=99)
    np.savez_compressed(OUT / "tm_spot.npz", sdr=s.numpy(), gm=m.numpy(), hdr=h.numpy(),
                        mulog=ref.fix_mulog_tmo(h, 99).numpy(), mulog_gamut=ref.gamut_compress(ref.fix_mulog_tmo(h, 99)).numpy())
    print("wrote", sorted(p.name for p in OUT.glob("tm_*.npz")), "hdr max", float(h.max()),
          "mulog mean", float(ref.fix_mulog_tmo(h, 99).mean()))


def rgbe_golden():
    """tests/golden/rgbe_cv2.npz: RGBE bytes parsed out of files the REAL cv2.imwrite wrote through the reference's own
    save_hdr_image recipe (scripts/inference/generate_hdr.py:27-30: /(qmax+1), float32, RGB->BGR flip)."""
    import os
    import tempfile

    import cv2

    from . import rgbe_oracle as ro
    rng = np.random.default_rng(0)
    H, W = 24, 40
    hdr = (rng.random((H, W, 3), dtype=np.float32) * np.exp(rng.uniform(-14, 5, (H, W, 1))).astype(np.float32)) * 100.0
    hdr[0, :8] = [[0, 0, 0], [1e-31, 0, 0], [100, 100, 100], [50, 25, 12.5], [100, 0, 0], [0, 100, 0], [0, 0, 100], [3e-31, 2e-31, 1e-31]]
    hdr[1, :4] = [[1.0, 2.0, 4.0], [127.99, 128.0, 128.01], [1e-30, 1e-30, 1e-30], [99.999, 1e-3, 1e-6]]
    narrow = hdr[:5, :6].copy()  # W < 8: OpenCV writes flat pixels
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for name, img in (("wide", hdr), ("narrow", narrow)):
            q = 99
            x = (img / (q + 1)).astype(np.float32)
            cv2.imwrite(os.path.join(d, f"{name}.hdr"), x[:, :, [2, 1, 0]])
            raw = open(os.path.join(d, f"{name}.hdr"), "rb").read()
            out[f"{name}_hdr"] = img
            out[f"{name}_rgbe"] = ro.parse_radiance(raw)
            out[f"{name}_decoded"] = cv2.imread(os.path.join(d, f"{name}.hdr"), cv2.IMREAD_UNCHANGED)[:, :, ::-1].copy()
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(OUT / "rgbe_cv2.npz", **out)
    print("wrote rgbe_cv2.npz (cv2", cv2.__version__, ")")


def grad_golden():
    """tests/golden/tm_grads.npz: gradients of the stage-1 training chain (scripts/stage1/train_vqgan_lora.py:1134-1141:
    apply_gm_to_sdr(qmax=49) -> fix_mulog_tmo -> gamut_compress) and of each function alone, from torch autograd through the
    REAL reference tone_mapping.py."""
    ref = tmo.load_reference_tm()
    if ref is None:
        raise SystemExit("reference not present")
    g = torch.Generator().manual_seed(7)
    shape = (2, 3, 16, 12)
    sdr = (torch.rand(shape, generator=g) * 0.96 + 0.02)
    gm = (torch.rand(shape, generator=g) * 0.96 + 0.02)
    w = torch.randn(shape, generator=g)
    out = dict(sdr=sdr.numpy(), gm=gm.numpy(), w=w.numpy())

    def grads(fn, *xs):
        xs = [x.clone().requires_grad_(True) for x in xs]
        y = fn(*xs)
        (y * w).sum().backward()
        return y.detach().numpy(), [x.grad.numpy() for x in xs]

    y, (g_gm, g_sdr) = grads(lambda a, b: ref.gamut_compress(ref.fix_mulog_tmo(ref.apply_gm_to_sdr(a, b, qmax=49), 49)), gm, sdr)
    out.update(chain_out=y, chain_g_gm=g_gm, chain_g_sdr=g_sdr)
    y, (g_gm, g_sdr) = grads(lambda a, b: ref.apply_gm_to_sdr(a, b, qmax=49), gm, sdr)
    out.update(eq1_out=y, eq1_g_gm=g_gm, eq1_g_sdr=g_sdr)
    hdr = torch.from_numpy(y)
    for name, fn in (("mulog", lambda x: ref.fix_mulog_tmo(x, 49)), ("linear", lambda x: ref.linear_scale_tmo(x, 49)),
                     ("hardclip", lambda x: ref.hard_clip_tmo(x * 0.03, 49)), ("gamut", lambda x: ref.gamut_compress(x * 0.02)),
                     ("tmo_cuda", lambda x: ref.tmo_cuda(x * 0.3))):
        y2, (gx,) = grads(fn, hdr)
        out[f"{name}_out"], out[f"{name}_g"] = y2, gx
    np.savez_compressed(OUT / "tm_grads.npz", **out)
    print("wrote tm_grads.npz")


def exposure_golden():
    """tests/golden/exposure_reference.npz: the REAL gm_diffusion/stage1/augmentations.py (pure torch) executed by file path."""
    import importlib.util
    import random
    path = "/root/reference/gm_diffusion/stage1/augmentations.py"
    spec = importlib.util.spec_from_file_location("ref_aug", path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    out = {}
    random.seed(3); torch.manual_seed(3)
    aug = m.RandomExposureAdjust()
    g = torch.Generator().manual_seed(9)
    x = torch.rand(2, 3, 20, 28, generator=g)
    x.view(-1)[:6] = torch.tensor([0.0, 1.0, 0.5, 1e-6, 0.999999, 0.25])
    out["x"] = x.numpy()
    for i in range(4):
        y, meta = aug(x, return_metadata=True)
        out[f"y{i}"] = y.numpy()
        out[f"meta{i}"] = np.array([meta["exposure"], meta["n"], meta["sigma"]], np.float64)
    out["curve"] = m.RandomExposureAdjust.apply_inv_sigmoid_curve(x, 0.7, 0.55).numpy()
    out["quant"] = m.RandomExposureAdjust.discretize_to_uint16(x * 1.2 - 0.1).numpy()
    out["ldr"] = aug.hdr_to_ldr(x * 3, 0.5).numpy()
    np.savez_compressed(OUT / "exposure_reference.npz", **out)
    print("wrote exposure_reference.npz")


if __name__ == "__main__":
    rgbe_golden()
    grad_golden()
    exposure_golden()
    main()
