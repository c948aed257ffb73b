"""Drop-in for gm_diffusion/stage1/tone_mapping.py: same function names, argument meaning and error
behaviour, each routed to the fused sm_100a kernel `gmd_hdr_reconstruct` (csrc/hdr.cu).  CUDA tensors only —
there is no CPU or PyTorch fallback.  `reconstruct_hdr` exposes the whole fused chain (de-normalise ->
Eq.(1) -> TMO -> gamut -> min/max) as ONE launch; the reference spends ~15 launches and a host round-trip on it
(scripts/inference/generate_hdr.py:225-268)."""
from __future__ import annotations

import ctypes as C
import random
from typing import Optional, Tuple

import torch

from .. import _lib as L

_TMO = {None: L.TMO_NONE, "none": L.TMO_NONE, "linear_scale": L.TMO_LINEAR, "hard_clip": L.TMO_HARD_CLIP,
        "fix_mulog": L.TMO_MULOG, "mulog": L.TMO_MULOG, "tmo_cuda": L.TMO_CUDA}


def _prep(t: torch.Tensor) -> Tuple[torch.Tensor, int]:
    if t.dtype == torch.bfloat16:
        return t.contiguous(), L.BF16
    return t.to(torch.float32).contiguous(), L.F32


@L.on_tensor_device
def _run(sdr, gm, *, flags, tmo, qmax, eps, mu, want_hdr, want_tmo, want_minmax, layout=None, want_rgbe=False, want_u8=False,
         rgbe_div=1.0):
    """Returns (hdr, tmo, minmax) and, when byte outputs are requested, (hdr, tmo, minmax, rgbe, sdr_u8, gm_u8)."""
    L.require_cuda(sdr, gm)
    sdr, dt = _prep(sdr)
    if gm is not None:
        gm, dt2 = _prep(gm)
        if dt2 != dt:
            sdr, gm, dt = sdr.float(), gm.float(), L.F32
        if gm.shape != sdr.shape:
            gm, sdr = (t.contiguous() for t in torch.broadcast_tensors(gm, sdr))
    p = L.HdrParams()
    if layout is None:
        layout = L.LAYOUT_PLANAR3 if (sdr.dim() == 4 and sdr.shape[1] == 3 and (flags & L.HDR_GAMUT)) else L.LAYOUT_FLAT
    if layout == L.LAYOUT_PLANAR3:
        p.n_px, p.batch = sdr.shape[2] * sdr.shape[3], sdr.shape[0]
    elif layout == L.LAYOUT_INTERLEAVED3:
        p.n_px, p.batch = sdr.numel() // 3, 1
    else:
        p.n_px, p.batch = sdr.numel(), 1
    if sdr.numel() == 0:  # empty input: nothing to launch
        e = torch.empty(sdr.shape, dtype=torch.float32, device=sdr.device)
        mm0 = torch.tensor([0x7F800000, -2139095041], dtype=torch.int32, device=sdr.device) if want_minmax else None  # (+inf, -inf)
        if want_rgbe or want_u8:
            u = torch.empty(sdr.shape, dtype=torch.uint8, device=sdr.device)
            r = torch.empty((*_px_shape(sdr, layout), 4), dtype=torch.uint8, device=sdr.device) if want_rgbe else None
            return (e if want_hdr else None), (e.clone() if want_tmo else None), mm0, r, (u if want_u8 else None), (u.clone() if want_u8 and gm is not None else None)
        return (e if want_hdr else None), (e.clone() if want_tmo else None), mm0
    hdr = torch.empty(sdr.shape, dtype=torch.float32, device=sdr.device) if want_hdr else None
    out = torch.empty(sdr.shape, dtype=torch.float32, device=sdr.device) if want_tmo else None
    mm = torch.empty(2, dtype=torch.int32, device=sdr.device) if want_minmax else None
    p.sdr, p.gm, p.hdr_out, p.tmo_out, p.minmax = L.ptr(sdr), L.ptr(gm), L.ptr(hdr), L.ptr(out), L.ptr(mm)
    p.layout, p.in_dtype, p.flags, p.tmo = layout, dt, flags, tmo
    p.qmax, p.eps, p.mu = float(qmax), float(eps), float(mu)
    rgbe = s8 = g8 = None
    if want_rgbe:
        if layout == L.LAYOUT_FLAT:
            raise ValueError("RGBE output needs [B,3,H,W] or channels_last [...,3]")
        rgbe = torch.empty((*_px_shape(sdr, layout), 4), dtype=torch.uint8, device=sdr.device)
        p.rgbe_out, p.rgbe_div = rgbe.data_ptr(), float(rgbe_div)
    if want_u8:
        s8 = torch.empty(sdr.shape, dtype=torch.uint8, device=sdr.device)
        p.sdr_u8_out = s8.data_ptr()
        if gm is not None:
            g8 = torch.empty(sdr.shape, dtype=torch.uint8, device=sdr.device)
            p.gm_u8_out = g8.data_ptr()
    L.check(L.lib().gmd_hdr_reconstruct(C.byref(p), L.current_stream()), "gmd_hdr_reconstruct")
    if want_rgbe or want_u8:
        return hdr, out, mm, rgbe, s8, g8
    return hdr, out, mm


def _px_shape(t: torch.Tensor, layout: int):
    """Shape of the pixel grid: [B,3,H,W] -> (B,H,W); [...,3] -> (...)."""
    return (t.shape[0], *t.shape[2:]) if layout == L.LAYOUT_PLANAR3 else tuple(t.shape[:-1])


def _needs_grad(*ts) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in ts)


class _HdrChainFn(torch.autograd.Function):
    """Autograd support for stage-1 training (scripts/stage1/train_vqgan_lora.py:1134-1141 back-propagates through
    apply_gm_to_sdr -> TMO -> gamut_compress): forward = the fused kernel, backward = `gmd_hdr_reconstruct_bwd`, one launch each."""

    @staticmethod
    def forward(ctx, sdr, gm, cfg):
        if sdr.dtype != torch.float32 or (gm is not None and (gm.dtype != torch.float32 or gm.shape != sdr.shape)):
            raise NotImplementedError("differentiable Eq.(1)/TMO needs fp32 inputs of identical shape (the training path)")
        sdr_c = sdr.detach().contiguous()
        gm_c = gm.detach().contiguous() if gm is not None else None
        hdr, out, _ = _run(sdr_c, gm_c, flags=cfg["flags"], tmo=cfg["tmo"], qmax=cfg["qmax"], eps=cfg["eps"], mu=cfg["mu"],
                           want_hdr=cfg["which"] == "hdr", want_tmo=cfg["which"] == "tmo", want_minmax=False, layout=cfg["layout"])
        ctx.save_for_backward(sdr_c, gm_c if gm_c is not None else sdr_c.new_empty(0))
        ctx.cfg = cfg
        ctx.has_gm = gm is not None
        return hdr if cfg["which"] == "hdr" else out

    @staticmethod
    def backward(ctx, grad):
        sdr, gm = ctx.saved_tensors
        cfg = ctx.cfg
        grad = grad.to(torch.float32).contiguous()
        p = L.HdrParams()
        layout = cfg["layout"]
        if layout is None:
            layout = L.LAYOUT_PLANAR3 if (sdr.dim() == 4 and sdr.shape[1] == 3 and (cfg["flags"] & L.HDR_GAMUT)) else L.LAYOUT_FLAT
        if layout == L.LAYOUT_PLANAR3:
            p.n_px, p.batch = sdr.shape[2] * sdr.shape[3], sdr.shape[0]
        elif layout == L.LAYOUT_INTERLEAVED3:
            p.n_px, p.batch = sdr.numel() // 3, 1
        else:
            p.n_px, p.batch = sdr.numel(), 1
        p.sdr, p.gm = sdr.data_ptr(), (gm.data_ptr() if ctx.has_gm else None)
        p.layout, p.in_dtype, p.flags, p.tmo = layout, L.F32, cfg["flags"], cfg["tmo"]
        p.qmax, p.eps, p.mu = float(cfg["qmax"]), float(cfg["eps"]), float(cfg["mu"])
        g_sdr = torch.empty_like(sdr) if ctx.needs_input_grad[0] else None
        g_gm = torch.empty_like(sdr) if (ctx.has_gm and ctx.needs_input_grad[1]) else None
        if g_sdr is None and g_gm is None:
            return None, None, None
        if sdr.numel():
            L.check(L.lib().gmd_hdr_reconstruct_bwd(C.byref(p), grad.data_ptr(), L.ptr(g_sdr), L.ptr(g_gm), int(cfg["which"] == "tmo"),
                                                    L.current_stream()), "gmd_hdr_reconstruct_bwd")
        return g_sdr, g_gm, None


def _chain(sdr, gm, *, flags, tmo, qmax, eps, mu, which, layout=None):
    """One differentiable output of the fused chain (`which` = "hdr" | "tmo")."""
    L.require_cuda(sdr, gm)
    # The kernels compute in fp32 on same-shape operands.  Mixed precision (accelerate fp16 / bf16 in train_vqgan_lora.py) and a
    # broadcast gain map go through differentiable torch casts / expands around the fused launch, and the result comes back in the
    # dtype torch's own promotion would give the reference expression.
    out_dtype = sdr.dtype if gm is None else torch.result_type(sdr, gm)
    if gm is not None and gm.shape != sdr.shape:
        sdr, gm = torch.broadcast_tensors(sdr, gm)
    s32 = sdr.to(torch.float32)
    g32 = gm.to(torch.float32) if gm is not None else None
    y = _HdrChainFn.apply(s32, g32, dict(flags=flags, tmo=tmo, qmax=qmax, eps=eps, mu=mu, which=which, layout=layout))
    return y if y.dtype == out_dtype or not out_dtype.is_floating_point else y.to(out_dtype)


def _decode_minmax(mm: torch.Tensor) -> Tuple[float, float]:
    lo, hi = mm.tolist()  # device -> host sync, like hdr.max()/hdr.min() at generate_hdr.py:268
    return float(L.lib().gmd_decode_ordered(lo)), float(L.lib().gmd_decode_ordered(hi))


def linear_scale_tmo(img: torch.Tensor, qmax: float) -> torch.Tensor:
    """tone_mapping.py:14-18."""
    if _needs_grad(img):
        return _chain(img, None, flags=0, tmo=L.TMO_LINEAR, qmax=qmax, eps=0, mu=0, which="tmo")
    return _run(img, None, flags=0, tmo=L.TMO_LINEAR, qmax=qmax, eps=0, mu=0, want_hdr=False, want_tmo=True,
                want_minmax=False)[1]


def hard_clip_tmo(hdr_img: torch.Tensor, qmax: float) -> torch.Tensor:
    """tone_mapping.py:21-26 (qmax ignored, kept for API compatibility)."""
    del qmax
    if _needs_grad(hdr_img):
        return _chain(hdr_img, None, flags=0, tmo=L.TMO_HARD_CLIP, qmax=0, eps=0, mu=0, which="tmo")
    return _run(hdr_img, None, flags=0, tmo=L.TMO_HARD_CLIP, qmax=0, eps=0, mu=0, want_hdr=False, want_tmo=True,
                want_minmax=False)[1]


def fix_mulog_tmo(hdr_img: torch.Tensor, qmax: float) -> torch.Tensor:
    """tone_mapping.py:29-36 (mu = 500)."""
    if _needs_grad(hdr_img):
        return _chain(hdr_img, None, flags=0, tmo=L.TMO_MULOG, qmax=qmax, eps=0, mu=500.0, which="tmo")
    return _run(hdr_img, None, flags=0, tmo=L.TMO_MULOG, qmax=qmax, eps=0, mu=500.0, want_hdr=False, want_tmo=True,
                want_minmax=False)[1]


def tmo_cuda(hdr_img: torch.Tensor) -> torch.Tensor:
    """tone_mapping.py:39-47; raises ValueError when the clamped image is not inside [0,1] (NaN input)."""
    _, out, mm = _run(hdr_img, None, flags=0, tmo=L.TMO_CUDA, qmax=0, eps=0, mu=5000.0, want_hdr=False, want_tmo=True,
                      want_minmax=True)
    lo, hi = _decode_minmax(mm)
    if hdr_img.numel() and hi != hi:  # the kernel reports the maximum of a tensor with NaNs as NaN; +-inf are clamped into [0, 1] like the reference does
        raise ValueError("HDR image values should be in the range [0, 1]")
    if _needs_grad(hdr_img):
        return _chain(hdr_img, None, flags=0, tmo=L.TMO_CUDA, qmax=0, eps=0, mu=5000.0, which="tmo")
    return out


def random_tmo_cuda(hdr_img: torch.Tensor, qmax: float) -> torch.Tensor:
    """tone_mapping.py:50-57: mu ~ U(500, 5000) from Python's `random`, as in the reference."""
    mu = random.uniform(500, 5_000)
    if _needs_grad(hdr_img):
        return _chain(hdr_img, None, flags=0, tmo=L.TMO_MULOG, qmax=qmax, eps=0, mu=mu, which="tmo")
    return _run(hdr_img, None, flags=0, tmo=L.TMO_MULOG, qmax=qmax, eps=0, mu=mu, want_hdr=False, want_tmo=True,
                want_minmax=False)[1]


def apply_gm_to_sdr(gm: torch.Tensor, sdr: torch.Tensor, qmax: float = 9, eps: float = 1 / 64) -> torch.Tensor:
    """tone_mapping.py:60-71: hdr = clamp((clamp(sdr,0,1)^2.2 + eps) * (1 + gm*qmax) - eps, 0, qmax+1)."""
    if _needs_grad(sdr, gm):
        return _chain(sdr, gm, flags=L.HDR_EQ1 | L.HDR_CLAMP_OUT, tmo=L.TMO_NONE, qmax=qmax, eps=eps, mu=0, which="hdr")
    return _run(sdr, gm, flags=L.HDR_EQ1 | L.HDR_CLAMP_OUT, tmo=L.TMO_NONE, qmax=qmax, eps=eps, mu=0, want_hdr=True,
                want_tmo=False, want_minmax=False)[0]


def gamut_compress(tmo_hdr_img: torch.Tensor) -> torch.Tensor:
    """tone_mapping.py:74-90: BT.2020 -> BT.709 matrix + clamp for [B,3,H,W]."""
    if tmo_hdr_img.dim() != 4 or tmo_hdr_img.shape[1] != 3:
        # the reference's permute(0,2,3,1) @ [3,3] fails on anything else
        raise RuntimeError(f"gamut_compress expects [B,3,H,W], got {tuple(tmo_hdr_img.shape)}")
    if _needs_grad(tmo_hdr_img):
        return _chain(tmo_hdr_img, None, flags=L.HDR_GAMUT, tmo=L.TMO_NONE, qmax=0, eps=0, mu=0, which="tmo", layout=L.LAYOUT_PLANAR3)
    return _run(tmo_hdr_img, None, flags=L.HDR_GAMUT, tmo=L.TMO_NONE, qmax=0, eps=0, mu=0, want_hdr=False, want_tmo=True,
                want_minmax=False, layout=L.LAYOUT_PLANAR3)[1]


def reconstruct_hdr(sdr: torch.Tensor, gm: torch.Tensor, qmax: float = 99.0, eps: float = 1 / 64, *, tmo: Optional[str] = None,
                    gamut: bool = False, denormalize: bool = False, clamp: bool = True, return_hdr: bool = True,
                    return_minmax: bool = False, mu: float = 500.0, channels_last: bool = False, exp_gain: bool = False):
    """The whole post-VAE chain in one launch.  Returns (hdr | None, tmo | None[, (min, max)]).

    sdr, gm: [B,3,H,W] (or [...,3] with channels_last) fp32/bf16, in [0,1] or, with `denormalize`, the raw VAE
    output in [-1,1] (generate_hdr.py:227,232).  `clamp=False` gives the numpy-twin behaviour
    (formal_improved.py:34-45)."""
    if tmo not in _TMO:
        raise ValueError(f"unknown tmo {tmo!r}; choose from {sorted(k for k in _TMO if k)}")
    flags = L.HDR_EQ1 | (L.HDR_DENORM if denormalize else 0) | (L.HDR_CLAMP_OUT if clamp else 0) | \
        (L.HDR_GAMUT if gamut else 0) | (L.HDR_EXP_GAIN if exp_gain else 0)
    want_tmo = tmo is not None or gamut
    if channels_last:
        if sdr.shape[-1] != 3:
            raise ValueError("channels_last expects [...,3]")
        layout = L.LAYOUT_INTERLEAVED3
    elif sdr.dim() == 4 and sdr.shape[1] == 3:
        layout = L.LAYOUT_PLANAR3
    else:
        if gamut:
            raise ValueError("gamut compression needs [B,3,H,W] or channels_last [...,3]")
        layout = L.LAYOUT_FLAT
    if _needs_grad(sdr, gm):
        # training: the whole chain as ONE differentiable op (forward one launch, backward one launch)
        if denormalize or exp_gain or return_minmax or not want_tmo:
            raise NotImplementedError("differentiable reconstruct_hdr: [0,1] inputs, linear gain, a TMO and/or gamut output, no min/max")
        out = _chain(sdr, gm, flags=flags, tmo=_TMO[tmo], qmax=qmax, eps=eps, mu=mu, which="tmo", layout=layout)
        return None, out
    hdr, out, mm = _run(sdr, gm, flags=flags, tmo=_TMO[tmo], qmax=qmax, eps=eps, mu=mu, want_hdr=return_hdr,
                        want_tmo=want_tmo, want_minmax=return_minmax, layout=layout)
    if return_minmax:
        return hdr, out, _decode_minmax(mm)
    return hdr, out


def reconstruct_for_disk(sdr: torch.Tensor, gm: torch.Tensor, qmax: float = 99.0, eps: float = 1 / 64, *, denormalize: bool = True,
                         clamp: bool = False, channels_last: bool = False, return_hdr: bool = False):
    """The inference scripts' whole host tail in ONE launch (generate_hdr.py:225-268 + save_hdr_image :27-30): de-normalise the two
    VAE images, Eq.(1) (numpy-twin behaviour: no output clamp by default), and emit what goes to disk — PNG-ready uint8 SDR and
    GM (`(x * 255).astype(np.uint8)`, :243-244) and the Radiance RGBE pixels of `hdr / (qmax + 1)` exactly as
    `cv2.imwrite("x.hdr", ...)` quantises them — so the device->host copy is 4 + 3 + 3 B/px instead of 2 x 12 B/px of fp32.

    Returns (rgbe uint8 [B,H,W,4] in R,G,B,E order, sdr_u8, gm_u8[, hdr fp32]); sdr_u8/gm_u8 keep the input's layout."""
    flags = L.HDR_EQ1 | (L.HDR_DENORM if denormalize else 0) | (L.HDR_CLAMP_OUT if clamp else 0)
    if channels_last:
        if sdr.shape[-1] != 3:
            raise ValueError("channels_last expects [...,3]")
        layout = L.LAYOUT_INTERLEAVED3
    elif sdr.dim() == 4 and sdr.shape[1] == 3:
        layout = L.LAYOUT_PLANAR3
    else:
        raise ValueError("reconstruct_for_disk needs [B,3,H,W] or channels_last [...,3]")
    hdr, _, _, rgbe, s8, g8 = _run(sdr, gm, flags=flags, tmo=L.TMO_NONE, qmax=qmax, eps=eps, mu=0, want_hdr=return_hdr, want_tmo=False,
                                   want_minmax=False, layout=layout, want_rgbe=True, want_u8=True, rgbe_div=float(qmax) + 1.0)
    return (rgbe, s8, g8, hdr) if return_hdr else (rgbe, s8, g8)


def rgbe_encode(hdr: torch.Tensor, divisor: float = 1.0, *, channels_last: bool = True) -> torch.Tensor:
    """Radiance RGBE pixels of `hdr / divisor` (uint8 [...,4], R,G,B,E) — the quantisation `cv2.imwrite("x.hdr", bgr)` applies."""
    if channels_last:
        if hdr.shape[-1] != 3:
            raise ValueError("channels_last expects [...,3]")
        layout = L.LAYOUT_INTERLEAVED3
    elif hdr.dim() == 4 and hdr.shape[1] == 3:
        layout = L.LAYOUT_PLANAR3
    else:
        raise ValueError("rgbe_encode needs [B,3,H,W] or channels_last [...,3]")
    return _run(hdr, None, flags=0, tmo=L.TMO_NONE, qmax=0, eps=0, mu=0, want_hdr=False, want_tmo=False, want_minmax=False,
                layout=layout, want_rgbe=True, rgbe_div=float(divisor))[3]


__all__ = ["linear_scale_tmo", "hard_clip_tmo", "fix_mulog_tmo", "tmo_cuda", "random_tmo_cuda", "apply_gm_to_sdr",
           "gamut_compress", "reconstruct_hdr", "reconstruct_for_disk", "rgbe_encode"]
