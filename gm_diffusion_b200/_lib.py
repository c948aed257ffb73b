"""ctypes binding of libgmd_b200.so (the C-ABI declared in include/gmd_b200.h).

There is NO fallback: if the library is missing or a call is made without a CUDA device the binding raises.
Structures mirror the header field for field; tests check every declared symbol is exported.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from pathlib import Path

_ROOT = Path(__file__).resolve().parent
LIB_PATH = _ROOT / "_C" / "libgmd_b200.so"
HEADER_PATH = _ROOT.parent / "include" / "gmd_b200.h"

# enums (include/gmd_b200.h)
LAYOUT_FLAT, LAYOUT_PLANAR3, LAYOUT_INTERLEAVED3 = 0, 1, 2
F32, BF16 = 0, 1
TMO_NONE, TMO_LINEAR, TMO_HARD_CLIP, TMO_MULOG, TMO_CUDA = 0, 1, 2, 3, 4
HDR_EQ1, HDR_DENORM, HDR_CLAMP_OUT, HDR_GAMUT, HDR_EXP_GAIN = 1, 2, 4, 8, 16
SCHED_LINEAR, SCHED_DDIM, SCHED_DDPM, SCHED_DPMPP = 0, 1, 2, 3
EPI_BIAS, EPI_ROW_BIAS, EPI_RESIDUAL, EPI_GEGLU, EPI_OUT_F32, EPI_SCALE, EPI_RESIDUAL_F32 = 1, 2, 4, 8, 16, 32, 64
CONV_PAD_END = 256

_vp, _i64, _i32, _f32 = C.c_void_p, C.c_int64, C.c_int32, C.c_float


class HdrParams(C.Structure):
    _fields_ = [("sdr", _vp), ("gm", _vp), ("hdr_out", _vp), ("tmo_out", _vp), ("minmax", _vp),
                ("n_px", _i64), ("batch", _i64), ("layout", _i32), ("in_dtype", _i32), ("flags", _i32),
                ("tmo", _i32), ("qmax", _f32), ("eps", _f32), ("mu", _f32), ("rgbe_div", _f32),
                ("rgbe_out", _vp), ("sdr_u8_out", _vp), ("gm_u8_out", _vp)]


class SchedParams(C.Structure):
    _fields_ = [("eps_uncond", _vp), ("eps_cond", _vp), ("x", _vp), ("x_stash", _vp), ("hist", _vp * 3),
                ("noise", _vp), ("x_next", _vp), ("stash_out", _vp), ("eps_out", _vp), ("unet_in_next", _vp),
                ("concat_out", _vp), ("concat_tail", _vp), ("concat_lead", _vp), ("x0_out", _vp),
                ("n_px", _i64), ("px_per_sample", _i64), ("unet_in_ch", _i32), ("unet_in_dup", _i32), ("concat_dup", _i32), ("concat_self", _i32), ("mode", _i32),
                ("use_stash", _i32), ("guidance_scale", _f32), ("guidance_rescale", _f32),
                ("rescale_stats", _vp), ("sqrt_alpha_t", _f32), ("sqrt_1m_alpha_t", _f32), ("plms_kind", _i32),
                ("c_sample", _f32), ("c_num", _f32), ("c_denom", _f32), ("ddim_sqrt_alpha_t", _f32), ("ddim_sqrt_1m_alpha_t", _f32),
                ("ddim_sqrt_alpha_prev", _f32), ("ddim_dir_coeff", _f32), ("ddim_sigma", _f32)]


class GemmParams(C.Structure):
    _fields_ = [("a", _vp), ("lda", _i64), ("w", _vp), ("ldw", _i64), ("out", _vp), ("ldo", _i64),
                ("bias", _vp), ("row_bias", _vp), ("ld_row_bias", _i64), ("rows_per_sample", _i64),
                ("residual", _vp), ("ldr", _i64), ("M", _i64), ("N", _i64), ("K", _i64), ("batch", _i64),
                ("stride_a", _i64), ("stride_w", _i64), ("stride_o", _i64), ("flags", _i32), ("alpha", _f32),
                ("workspace", _vp), ("workspace_bytes", _i64), ("w_tiled", _i32), ("gn_sums", _vp), ("gn_rows_per_sample", _i64),
                ("ln_out_sums", _vp), ("ln_out_copy", _vp), ("ln_in_sums", _vp), ("ln_in_c", _vp), ("ln_eps", _f32)]


class ConvParams(C.Structure):
    _fields_ = [("x0", _vp), ("C0", _i32), ("x1", _vp), ("C1", _i32), ("w", _vp), ("out", _vp), ("bias", _vp),
                ("row_bias", _vp), ("ld_row_bias", _i64), ("residual", _vp), ("N", _i32), ("H", _i32), ("W", _i32),
                ("Cout", _i32), ("Cout_pad", _i32), ("ksize", _i32), ("stride", _i32), ("upsample", _i32),
                ("flags", _i32), ("workspace", _vp), ("workspace_bytes", _i64), ("w_tiled", _i32), ("gn_sums", _vp)]


class AttnParams(C.Structure):
    _fields_ = [("q", _vp), ("q_stride_b", _i64), ("q_stride_n", _i64), ("q_stride_h", _i64),
                ("k", _vp), ("k_stride_b", _i64), ("k_stride_n", _i64), ("k_stride_h", _i64),
                ("v", _vp), ("v_stride_b", _i64), ("v_stride_n", _i64), ("v_stride_h", _i64),
                ("o", _vp), ("o_stride_b", _i64), ("o_stride_n", _i64), ("o_stride_h", _i64),
                ("B", _i32), ("H", _i32), ("Nq", _i32), ("Nk", _i32), ("d", _i32), ("scale", _f32)]


_lib = None


def declared_symbols() -> list[str]:
    """Every function name declared in include/gmd_b200.h."""
    text = HEADER_PATH.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gmd_[a-z0-9_]+)\s*\(", text)))


def _header_version():
    m = re.search(r"#define\s+GMD_VERSION\s+(\d+)", HEADER_PATH.read_text()) if HEADER_PATH.exists() else None
    return int(m.group(1)) if m else None


def _check_fresh() -> None:
    """A library built from other sources than the ones in the tree (and so, possibly, with other struct layouts than the ctypes
    mirrors below) must not be used silently: compare the build stamp with the digest of csrc/ + the header."""
    if LIB_PATH.parent != _ROOT / "_C" or os.environ.get("GMD_SKIP_DIGEST_CHECK") == "1":
        return   # an explicitly chosen A/B library (profiles/*: GMD_AB_LIB)
    stamp = LIB_PATH.parent / "build.sha256"
    from . import build as _build
    if not (_build.CSRC.exists() and HEADER_PATH.exists()):
        return
    if not stamp.exists() or stamp.read_text().strip() != _build._digest():
        raise RuntimeError(f"{LIB_PATH} is stale (csrc/ or include/gmd_b200.h changed since it was built): rebuild with "
                           "`python -m gm_diffusion_b200.build`")


def lib() -> C.CDLL:
    """Load the shared library (once). Raises if it has not been built — there is no CPU path."""
    global _lib, LIB_PATH
    if _lib is not None:
        return _lib
    if os.environ.get("GMD_AB_LIB") and LIB_PATH == _ROOT / "_C" / "libgmd_b200.so":
        LIB_PATH = Path(os.environ["GMD_AB_LIB"]).resolve()   # an alternative build of the library for A/B measurements (profiles/, bench.py)
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m gm_diffusion_b200.build` "
            "(gm_diffusion_b200 has no CPU or PyTorch fallback)")
    _check_fresh()
    L = C.CDLL(str(LIB_PATH), mode=os.RTLD_GLOBAL if hasattr(os, "RTLD_GLOBAL") else 0)
    want = _header_version()
    if want is not None and int(L.gmd_version()) != want:
        raise RuntimeError(f"{LIB_PATH} reports gmd_version() = {int(L.gmd_version())} but include/gmd_b200.h declares {want}: "
                           "rebuild with `python -m gm_diffusion_b200.build --force`")
    L.gmd_version.restype = C.c_int
    L.gmd_last_error.restype = C.c_char_p
    L.gmd_launch_count.restype = C.c_int64
    L.gmd_reset_launch_count.restype = None
    L.gmd_add_launch_count.restype = None
    L.gmd_add_launch_count.argtypes = [C.c_int64]
    L.gmd_decode_ordered.restype = C.c_float
    L.gmd_decode_ordered.argtypes = [C.c_int32]
    L.gmd_hdr_reconstruct.argtypes = [C.POINTER(HdrParams), _vp]
    L.gmd_exposure_adjust.argtypes = [_vp, _vp, _i64, _i32, C.c_double, _f32, _f32, C.c_double, _vp]
    L.gmd_hdr_reconstruct_bwd.argtypes = [C.POINTER(HdrParams), _vp, _vp, _vp, _i32, _vp]
    L.gmd_cfg_sched_step.argtypes = [C.POINTER(SchedParams), _vp]
    L.gmd_latents_nchw_to_px.argtypes = [_vp, _vp, _i64, _i64, _vp]
    L.gmd_latents_px_to_nchw.argtypes = [_vp, _vp, _i64, _i64, _vp]
    L.gmd_pack_unet_input.argtypes = [_vp, _vp, _vp, _i64, _i32, _vp]
    L.gmd_pack_image_nchw.argtypes = [_vp, _vp, _i64, _i64, _i32, _vp]
    L.gmd_vae_sample.argtypes = [_vp, _vp, _vp, _i64, _f32, _vp]
    L.gmd_gemm_fwd.argtypes = [C.POINTER(GemmParams), _vp]
    L.gmd_conv_fwd.argtypes = [C.POINTER(ConvParams), _vp]
    L.gmd_conv_gn_sums_ok.argtypes = [C.POINTER(ConvParams)]
    L.gmd_gemm_gn_sums_ok.argtypes = [C.POINTER(GemmParams), _i64]
    L.gmd_groupnorm_apply.argtypes = [_vp, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _f32, _i32, _i32, _vp]
    L.gmd_groupnorm_silu.argtypes = [_vp, _i32, _vp, _i32, _vp, _vp, _vp, _i32, _i32, _i32, _f32, _i32, _i32, _vp, _vp]
    L.gmd_layernorm.argtypes = [_vp, _vp, _vp, _vp, _i64, _i32, _f32, _i32, _vp]
    L.gmd_softmax_rows.argtypes = [_vp, _vp, _i64, _i64, _f32, _vp]
    L.gmd_softmax_rows_masked.argtypes = [_vp, _vp, _i64, _i64, _f32, _i32, _i32, _vp]
    L.gmd_timestep_embedding.argtypes = [_f32, _vp, _i32, _i32, _vp]
    L.gmd_silu.argtypes = [_vp, _vp, _i64, _vp]
    L.gmd_attn_fwd.argtypes = [C.POINTER(AttnParams), _vp]
    for name in declared_symbols():
        fn = getattr(L, name)  # AttributeError here == header/library drift
        if fn.restype is C.c_int and name not in ("gmd_version",):
            pass
    _lib = L
    return L


def check(rc: int, what: str = "") -> None:
    """Turn a C-ABI return code into the Python exception the reference API would raise."""
    if rc == 0:
        return
    msg = lib().gmd_last_error().decode("utf-8", "replace")
    if rc == -1:
        raise ValueError(f"{what}: {msg}" if what else msg)
    if rc == -3:
        raise NotImplementedError(f"{what}: {msg}" if what else msg)
    raise RuntimeError(f"{what}: {msg} (rc={rc})" if what else f"{msg} (rc={rc})")


def current_stream() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors) -> None:
    """The product path is CUDA-only; refuse anything else loudly."""
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("gm_diffusion_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    cur = torch.cuda.current_device()
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("gm_diffusion_b200: expected CUDA tensors (no CPU fallback)")
        if t.device.index != cur:
            # kernels launch on the CURRENT device's current stream: a tensor elsewhere would be read through a wrong context
            raise RuntimeError(f"gm_diffusion_b200: tensor lives on cuda:{t.device.index} but the current CUDA device is cuda:{cur}; "
                               "run the call under `with torch.cuda.device(tensor.device):` (the pipelines, B200UNet and B200Vae do)")


def on_own_device(method):
    """Decorator for methods of objects with a `.device`: run under `torch.cuda.device(self.device)` so every C-ABI launch inside
    goes to the stream and kernel attributes of THAT device (one process may hold pipelines on several GPUs)."""
    import functools

    @functools.wraps(method)
    def wrapper(self, *a, **kw):
        import torch

        dev = torch.device(self.device)
        if dev.type != "cuda" or dev.index is None or dev.index == torch.cuda.current_device():
            return method(self, *a, **kw)
        with torch.cuda.device(dev):
            return method(self, *a, **kw)
    return wrapper


def on_tensor_device(fn):
    """Decorator for free functions whose first CUDA tensor argument decides the device."""
    import functools

    @functools.wraps(fn)
    def wrapper(*a, **kw):
        import torch

        for t in list(a) + list(kw.values()):
            if isinstance(t, torch.Tensor) and t.is_cuda:
                if t.device.index == torch.cuda.current_device():
                    break
                with torch.cuda.device(t.device):
                    return fn(*a, **kw)
        return fn(*a, **kw)
    return wrapper


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def launch_count() -> int:
    return int(lib().gmd_launch_count())


def reset_launch_count() -> None:
    lib().gmd_reset_launch_count()
