"""Thin tensor-level wrappers over the C-ABI (one C call each; no math happens in Python).

Layouts: activations bf16 NHWC `[N,H,W,C]` or token-major `[M,C]`; conv weights `[Cout, kh*kw*Cin]` with
k = (r*kw + s)*Cin + c; linear weights `[N,K]` (torch Linear layout); biases / norm affine fp32.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L

bf16 = torch.bfloat16

_WS = {}


def pick_bn(n_rows: int, geglu: bool = False) -> int:
    """N tile the GEMM kernel uses for `n_rows` weight rows (mirror of pick_bn in csrc/gemm.cu)."""
    if n_rows % 160 == 0:
        return 160
    if n_rows % 128 == 0 or n_rows > 128:
        return 128
    if n_rows > 32 or geglu:
        return 64
    return 32


class TiledWeight:
    """Weights pre-tiled for the tcgen05 kernels: `[ceil(N/T)][ceil(K/64)][T][64]` bf16 (zero padded), so every operand tile the
    TMA fetches is ONE contiguous 128*T-byte read.  With the plain `[N, K]` layout a tile is T separate 128-byte rows, each in
    a different DRAM page: weight-streaming-bound layers (8x8 / 16x16 resolution) ran at ~0.4 TB/s."""

    def __init__(self, data: torch.Tensor, n: int, k: int, bn: int, swizzled: bool = False):
        self.data, self.n, self.k, self.bn, self.swizzled = data, n, k, bn, swizzled

    @property
    def code(self) -> int:
        """value of the C-ABI `w_tiled` field"""
        return self.bn + (1000 if self.swizzled else 0)

    @property
    def shape(self):
        return (self.n, self.k)

    def to(self, device):
        return TiledWeight(self.data.to(device), self.n, self.k, self.bn, self.swizzled)


def tile_weight(w: torch.Tensor, geglu: bool = False, swizzle: bool = True) -> TiledWeight:
    """[N, K] (any float dtype) -> TiledWeight.  `swizzle`: store each [T][64] tile as its SWIZZLE_128B shared-memory image so
    the kernel fetches it with one bulk copy."""
    n, k = w.shape
    bn = pick_bn(n, geglu)
    tn, tk = (n + bn - 1) // bn, (k + 63) // 64
    wp = torch.zeros((tn * bn, tk * 64), dtype=bf16, device=w.device)
    wp[:n, :k] = w.to(bf16)
    data = wp.view(tn, bn, tk, 64).permute(0, 2, 1, 3).contiguous().view(tn * tk * bn, 64)
    if swizzle:
        # 16-byte chunk c (8 bf16) of tile row r moves to chunk c ^ (r & 7); bn is a multiple of 8, so r & 7 == global row & 7
        rows = torch.arange(data.shape[0], device=data.device) & 7
        src_chunk = torch.arange(8, device=data.device)[None, :] ^ rows[:, None]          # dst chunk j holds src chunk j ^ (r & 7)
        data = data.view(-1, 8, 8).gather(1, src_chunk[:, :, None].expand(-1, -1, 8)).reshape(-1, 64).contiguous()
    return TiledWeight(data, n, k, bn, swizzle)



_SLOT = 0


class scratch_slot:
    """Static scratch (split-K workspace, GroupNorm accumulator arena) is per (device, slot).  Work that runs CONCURRENTLY with other
    work of this library on the same device — the GM UNet on a side stream beside the SDR UNet of the next step — selects a slot of
    its own with `with ops.scratch_slot(1): ...`; everything else uses slot 0."""

    def __init__(self, slot: int):
        self.slot = int(slot)

    def __enter__(self):
        global _SLOT
        self.prev, _SLOT = _SLOT, self.slot
        return self

    def __exit__(self, *exc):
        global _SLOT
        _SLOT = self.prev
        return False


def _scratch_key(device):
    return (torch.device(device).index or 0, _SLOT)


def splitk_workspace(device) -> torch.Tensor:
    """fp32 scratch for the split-K partial sums of the few-tile / long-K layers (allocated once per device and slot, static address)."""
    key = _scratch_key(device)
    ws = _WS.get(key)
    if ws is None:
        ws = torch.empty(128 << 20, dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws


LN_FOLD = __import__("os").environ.get("GMD_LN_FOLD", "0") == "1"   # 1: norm1 / norm2 of the transformer blocks folded into the GEMMs around them (measured neutral: DESIGN.md)
GN_EPILOGUE = __import__("os").environ.get("GMD_GN_EPILOGUE", "1") != "0"   # 0: statistics by the standalone GroupNorm kernel (A/B measurements)
_GN_ARENA = {}


class gn_arena:
    """Zeroed int64 accumulators for the GroupNorm statistics the convolution / GEMM epilogues add into (gmd_b200.h "GroupNorm fused
    into the producing convolution / GEMM"), bump-allocated from one static buffer per device and zeroed by ONE memset per forward
    pass instead of one per tensor:

        with ops.gn_arena(device, words_hint) as ar:   # zeroes words_hint words (everything, when the hint is None)
            ... conv2d(..., gn_stats=True) ...
        words_hint = ar.used

    Outside such a block every statistics tensor is a fresh torch.zeros.  The buffer is static (CUDA-graph friendly)."""
    WORDS = 8 << 20

    def __init__(self, device, words_hint=None):
        self.key = _scratch_key(device)
        self.device, self.hint, self.used = device, words_hint, 0

    def __enter__(self):
        st = _GN_ARENA.setdefault(self.key, {"buf": None, "active": None})
        if st["buf"] is None:
            st["buf"] = torch.zeros(self.WORDS, dtype=torch.int64, device=self.device)
        self.buf = st["buf"]
        n = self.WORDS if self.hint is None else min(self.WORDS, int(self.hint))
        if n > 0:
            self.buf[:n].zero_()
        self.zeroed = n
        self.prev, st["active"] = st["active"], self
        return self

    def __exit__(self, *exc):
        _GN_ARENA[self.key]["active"] = self.prev
        return False

    def alloc(self, n: int, c: int) -> torch.Tensor:
        words = n * c                       # [n, c/2, 2]
        if self.used + words > self.zeroed:
            if self.used + words <= self.WORDS and self.zeroed < self.WORDS:
                self.buf[self.zeroed:].zero_()          # the hint was too small (first pass at a larger batch): zero the rest
                self.zeroed = self.WORDS
            else:
                return torch.zeros((n, c // 2, 2), dtype=torch.int64, device=self.device)
        t = self.buf[self.used:self.used + words].view(n, c // 2, 2)
        self.used += (words + 1) // 2 * 2   # keep 16-byte alignment
        return t


def _gn_sums_alloc(device, n: int, c: int) -> torch.Tensor:
    st = _GN_ARENA.get(_scratch_key(device))
    if st is not None and st["active"] is not None:
        return st["active"].alloc(n, c)
    return torch.zeros((n, c // 2, 2), dtype=torch.int64, device=device)


def dup_rows(t: torch.Tensor) -> torch.Tensor:
    """[t ; t] along dim 0 (the two CFG halves of a still-shared activation) as two device-to-device memcpys — no copy kernel."""
    t = t.contiguous()
    n = t.shape[0]
    out = torch.empty((2 * n,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    out[:n].copy_(t)
    out[n:].copy_(t)
    return out


def gemm(a: torch.Tensor, w: torch.Tensor, *, bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
         row_bias: Optional[torch.Tensor] = None, rows_per_sample: int = 0, geglu: bool = False, out: Optional[torch.Tensor] = None,
         out_f32: bool = False, alpha: Optional[float] = None, splitk: bool = False, gn_rows_per_sample: int = 0,
         ln_out: bool = False, ln_in=None):
    """out[M,N] = a[M,K] @ w[N,K]^T (+bias +row_bias[sample] +residual | GEGLU).  3-D inputs are batched.
    `gn_rows_per_sample` > 0: also return the GroupNorm statistics of the output, accumulated by the epilogue (`(out, sums)`, sums
    int64 fixed point [samples, N/2, 2]; None where the kernel cannot provide them).
    LayerNorm folding (gmd_b200.h): `ln_out=True` also returns a bf16 copy of the output and its per-row statistics
    (`(out, copy, sums)`); `ln_in=(sums, c, eps)` makes this GEMM the projection behind that LayerNorm (a = the bf16 copy,
    w = W * diag(gamma), bias = b + W beta, c = row sums of w)."""
    tiled = isinstance(w, TiledWeight)
    wt = w.data if tiled else w
    L.require_cuda(a, wt)
    assert a.dtype == bf16 and wt.dtype == bf16, "gemm operands must be bf16"
    batched = a.dim() == 3
    assert not (tiled and batched)
    if batched:
        Bt, M, K = a.shape
        N = w.shape[1]
        assert w.shape[0] == Bt and w.shape[2] == K
    else:
        M, K = a.shape
        N = w.shape[0]
        assert w.shape[1] == K
        Bt = 1
    assert a.stride(-1) == 1 and wt.stride(-1) == 1
    n_out = N // 2 if geglu else N
    if out is None:
        shape = (Bt, M, n_out) if batched else (M, n_out)
        out = torch.empty(shape, dtype=torch.float32 if out_f32 else bf16, device=a.device)
    p = L.GemmParams()
    p.a, p.lda = a.data_ptr(), a.stride(-2)
    p.w, p.ldw = wt.data_ptr(), (K if tiled else w.stride(-2))
    p.w_tiled = w.code if tiled else 0
    p.out, p.ldo = out.data_ptr(), out.stride(-2)
    p.M, p.N, p.K, p.batch = M, N, K, Bt
    if batched:
        p.stride_a, p.stride_w, p.stride_o = a.stride(0), wt.stride(0), out.stride(0)
    flags = 0
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
        flags |= L.EPI_BIAS
        p.bias = bias.data_ptr()
    if row_bias is not None:
        assert row_bias.dtype == torch.float32 and rows_per_sample > 0
        flags |= L.EPI_ROW_BIAS
        p.row_bias, p.ld_row_bias, p.rows_per_sample = row_bias.data_ptr(), row_bias.stride(0), rows_per_sample
    if residual is not None:
        assert residual.dtype in (bf16, torch.float32) and residual.stride(-1) == 1
        flags |= L.EPI_RESIDUAL | (L.EPI_RESIDUAL_F32 if residual.dtype == torch.float32 else 0)
        p.residual, p.ldr = residual.data_ptr(), residual.stride(-2)
    if geglu:
        flags |= L.EPI_GEGLU
    if out_f32 or out.dtype == torch.float32:
        flags |= L.EPI_OUT_F32
    if alpha is not None:
        flags |= L.EPI_SCALE
        p.alpha = float(alpha)
    p.flags = flags
    if splitk and not batched:
        # opt-in for GEMMs: here the number of K splits depends on the tile count, i.e. on M — it would make a sample's result
        # depend on the batch it is computed in (summation order), which the sharded pipelines must not do.  (The convolution
        # entry point uses a batch-independent rule instead and is on by default, see conv2d.)
        ws = splitk_workspace(a.device)
        p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel()
    ln_copy = ln_sums = None
    if ln_out:
        assert not batched and not geglu
        ln_copy = torch.empty((M, N), dtype=bf16, device=a.device)
        ln_sums = _gn_sums_alloc(a.device, M, 2).view(M, 2)
        p.ln_out_sums, p.ln_out_copy = ln_sums.data_ptr(), ln_copy.data_ptr()
    if ln_in is not None:
        s_in, c_in, eps_in = ln_in
        assert s_in.dtype == torch.int64 and s_in.shape == (M, 2) and s_in.is_contiguous() and c_in.dtype == torch.float32 and c_in.numel() == N
        p.ln_in_sums, p.ln_in_c, p.ln_eps = s_in.data_ptr(), c_in.data_ptr(), float(eps_in)
    sums = None
    if gn_rows_per_sample > 0 and GN_EPILOGUE and L.lib().gmd_gemm_gn_sums_ok(C.byref(p), gn_rows_per_sample):
        sums = _gn_sums_alloc(a.device, M // gn_rows_per_sample, N)
        p.gn_sums, p.gn_rows_per_sample = sums.data_ptr(), gn_rows_per_sample
    L.check(L.lib().gmd_gemm_fwd(C.byref(p), L.current_stream()), "gmd_gemm_fwd")
    if ln_out:
        return out, ln_copy, ln_sums
    if gn_rows_per_sample > 0:
        return out, sums
    return out


def conv2d(x: torch.Tensor, w: torch.Tensor, cout: int, *, ksize: int = 3, stride: int = 1, upsample: bool = False,
           x1: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None, row_bias: Optional[torch.Tensor] = None,
           residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, out_f32: bool = False, splitk: bool = True,
           pad_end: bool = False, gn_stats: bool = False):
    """NHWC bf16 convolution as implicit GEMM (3x3 pad 1 stride 1/2, optional folded nearest-2x upsample).  `pad_end` (stride 2):
    the AutoencoderKL encoder's asymmetric padding — no leading pad, one zero row/column at the bottom/right.
    `gn_stats`: also return the GroupNorm statistics of the output, accumulated by the epilogue (`(out, sums)`, sums int64 fixed
    point [N, Cout/2, 2]; None where the kernel cannot provide them: split-K layers, ragged tiles, images smaller than a tile)."""
    tiled = isinstance(w, TiledWeight)
    wt = w.data if tiled else w
    L.require_cuda(x, wt)
    assert x.dtype == bf16 and wt.dtype == bf16 and x.is_contiguous() and wt.is_contiguous()
    N, H, W, C0 = x.shape
    C1 = 0
    if x1 is not None:
        assert x1.dtype == bf16 and x1.is_contiguous() and x1.shape[:3] == x.shape[:3]
        C1 = x1.shape[3]
    if tiled:
        assert w.n == cout and w.k == ksize * ksize * ((C0 + C1 + 63) // 64 * 64), (w.shape, ksize, C0, C1)
    else:
        assert w.shape[1] == ksize * ksize * (C0 + C1), (w.shape, ksize, C0, C1)
    Ho, Wo = (H // 2, W // 2) if stride == 2 else ((2 * H, 2 * W) if upsample else (H, W))
    if out is None:
        out = torch.empty((N, Ho, Wo, cout), dtype=torch.float32 if out_f32 else bf16, device=x.device)
    p = L.ConvParams()
    p.x0, p.C0 = x.data_ptr(), C0
    p.x1, p.C1 = (x1.data_ptr(), C1) if x1 is not None else (None, 0)
    p.w, p.out = wt.data_ptr(), out.data_ptr()
    p.w_tiled = w.code if tiled else 0
    p.N, p.H, p.W, p.Cout, p.Cout_pad = N, H, W, cout, w.shape[0]
    p.ksize, p.stride, p.upsample = ksize, stride, int(upsample)
    flags = 0
    if bias is not None:
        assert bias.dtype == torch.float32
        flags |= L.EPI_BIAS
        p.bias = bias.data_ptr()
    if row_bias is not None:
        assert row_bias.dtype == torch.float32
        flags |= L.EPI_ROW_BIAS
        p.row_bias, p.ld_row_bias = row_bias.data_ptr(), row_bias.stride(0)
    if residual is not None:
        assert residual.dtype in (bf16, torch.float32) and residual.is_contiguous()
        flags |= L.EPI_RESIDUAL | (L.EPI_RESIDUAL_F32 if residual.dtype == torch.float32 else 0)
        p.residual = residual.data_ptr()
    if out.dtype == torch.float32:
        flags |= L.EPI_OUT_F32
    if pad_end:
        assert stride == 2
        flags |= L.CONV_PAD_END
    p.flags = flags
    if splitk:
        # the library splits K four ways for the <= 8x8-resolution layers (32-64 tiles on 148 SMs otherwise: 79 -> 32 us per conv,
        # profiles/prof_splitk.py); the rule looks only at the per-image geometry, so results do not depend on the batch
        ws = splitk_workspace(x.device)
        p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel()
    sums = None
    if gn_stats and GN_EPILOGUE and L.lib().gmd_conv_gn_sums_ok(C.byref(p)):
        sums = _gn_sums_alloc(x.device, N, cout)
        p.gn_sums = sums.data_ptr()
    L.check(L.lib().gmd_conv_fwd(C.byref(p), L.current_stream()), "gmd_conv_fwd")
    if gn_stats:
        return out, sums
    return out


def gn_workspace(n_max: int, groups: int, device) -> torch.Tensor:
    """Zero-initialised GroupNorm scratch (partials, finals, arrival counters) for up to `n_max` samples."""
    return torch.zeros(1024 + n_max * 34 * groups * 2, dtype=torch.float32, device=device)


def groupnorm_silu(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, *, x1: Optional[torch.Tensor] = None, groups: int = 32,
                   eps: float = 1e-5, silu: bool = True, out: Optional[torch.Tensor] = None, stats_ws: Optional[torch.Tensor] = None,
                   sums: Optional[torch.Tensor] = None, sums1: Optional[torch.Tensor] = None) -> torch.Tensor:
    """GroupNorm(+SiLU) over NHWC bf16; with `x1` the channels are [x | x1] and the output is the concatenation.
    `sums` / `sums1`: statistics of x / x1 from their producers' epilogues (conv2d / gemm with gn_stats) — the activation is then
    read ONCE (gmd_groupnorm_apply) instead of being reduced first."""
    L.require_cuda(x)
    N, C0 = x.shape[0], x.shape[-1]
    HW = x.numel() // (N * C0)
    C1 = x1.shape[-1] if x1 is not None else 0
    if out is None:
        out = torch.empty(x.shape[:-1] + (C0 + C1,), dtype=bf16, device=x.device)
    if sums is not None and (x1 is None or sums1 is not None) and ((C0 + C1) // groups) % 2 == 0 and (C0 + C1) // 8 <= 320:
        assert sums.shape == (N, C0 // 2, 2) and (x1 is None or sums1.shape == (N, C1 // 2, 2))
        L.check(L.lib().gmd_groupnorm_apply(x.data_ptr(), C0, sums.data_ptr(), L.ptr(x1), C1, L.ptr(sums1) if x1 is not None else None,
                                            gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), N, HW, groups, float(eps), int(silu),
                                            L.F32 if x.dtype == torch.float32 else L.BF16, L.current_stream()), "gmd_groupnorm_apply")
        return out
    if stats_ws is None:
        stats_ws = gn_workspace(N, groups, x.device)
    assert stats_ws.numel() >= 1024 + N * 34 * groups * 2, "groupnorm workspace too small"
    L.check(L.lib().gmd_groupnorm_silu(x.data_ptr(), C0, L.ptr(x1), C1, gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), N, HW,
                                       groups, float(eps), int(silu), L.F32 if x.dtype == torch.float32 else L.BF16, stats_ws.data_ptr(),
                                       L.current_stream()), "gmd_groupnorm_silu")
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    L.require_cuda(x)
    Cc = x.shape[-1]
    M = x.numel() // Cc
    if out is None:
        out = torch.empty(x.shape, dtype=bf16, device=x.device)
    L.check(L.lib().gmd_layernorm(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), M, Cc, float(eps),
                                  L.F32 if x.dtype == torch.float32 else L.BF16, L.current_stream()), "gmd_layernorm")
    return out


def softmax_rows(x: torch.Tensor, scale: float, out: Optional[torch.Tensor] = None, n_valid: Optional[int] = None, causal_period: int = 0) -> torch.Tensor:
    """softmax(scale * x) over the last dim (bf16).  `n_valid`: columns beyond it are key padding (probability 0); `causal_period` > 0:
    row r additionally sees only columns <= r % causal_period (one causal matrix of that many rows per head)."""
    L.require_cuda(x)
    N = x.shape[-1]
    if out is None:
        out = torch.empty_like(x)
    if n_valid is None and causal_period == 0:
        L.check(L.lib().gmd_softmax_rows(x.data_ptr(), out.data_ptr(), x.numel() // N, N, float(scale), L.current_stream()), "gmd_softmax_rows")
    else:
        L.check(L.lib().gmd_softmax_rows_masked(x.data_ptr(), out.data_ptr(), x.numel() // N, N, float(scale), int(n_valid or N), int(causal_period),
                                                L.current_stream()), "gmd_softmax_rows_masked")
    return out


def timestep_embedding(t: float, batch: int, dim: int, device, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    if out is None:
        out = torch.empty((batch, dim), dtype=bf16, device=device)
    L.check(L.lib().gmd_timestep_embedding(float(t), out.data_ptr(), batch, dim, L.current_stream()), "gmd_timestep_embedding")
    return out


def silu(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    L.require_cuda(x)
    if out is None:
        out = torch.empty_like(x)
    L.check(L.lib().gmd_silu(x.data_ptr(), out.data_ptr(), x.numel(), L.current_stream()), "gmd_silu")
    return out


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, scale: Optional[float] = None,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q: [B,Nq,H*d] view, k/v: [B,Nk,H*d] views (last dim contiguous, arbitrary row strides) -> [B,Nq,H*d]."""
    L.require_cuda(q, k, v)
    B, Nq, Cc = q.shape
    Nk = k.shape[1]
    d = Cc // heads
    assert q.dtype == bf16 and k.dtype == bf16 and v.dtype == bf16
    assert q.stride(2) == 1 and k.stride(2) == 1 and v.stride(2) == 1
    if out is None:
        out = torch.empty((B, Nq, Cc), dtype=bf16, device=q.device)
    p = L.AttnParams()
    p.q, p.q_stride_b, p.q_stride_n, p.q_stride_h = q.data_ptr(), q.stride(0), q.stride(1), d
    p.k, p.k_stride_b, p.k_stride_n, p.k_stride_h = k.data_ptr(), k.stride(0), k.stride(1), d
    p.v, p.v_stride_b, p.v_stride_n, p.v_stride_h = v.data_ptr(), v.stride(0), v.stride(1), d
    p.o, p.o_stride_b, p.o_stride_n, p.o_stride_h = out.data_ptr(), out.stride(0), out.stride(1), d
    p.B, p.H, p.Nq, p.Nk, p.d = B, heads, Nq, Nk, d
    p.scale = float(scale if scale is not None else d ** -0.5)
    L.check(L.lib().gmd_attn_fwd(C.byref(p), L.current_stream()), "gmd_attn_fwd")
    return out


# ---- weight repacking helpers (run once at load) ------------------------------------------------------------
def pack_conv_weight_tiled(w_oihw: torch.Tensor, cin_pad: Optional[int] = None) -> TiledWeight:
    """OIHW -> TiledWeight with k = (r*kw + s)*Cpad + c, Cpad = channels zero-padded to a multiple of 64 (one k block never
    straddles two taps).  `cin_pad` first pads the logical input channels (conv_in: 4 -> 8)."""
    co, ci, kh, kw = w_oihw.shape
    w = w_oihw.permute(0, 2, 3, 1)
    ci2 = max(ci, cin_pad or 0)
    cpad = (ci2 + 63) // 64 * 64
    if cpad > ci:
        w = torch.nn.functional.pad(w, (0, cpad - ci))
    return tile_weight(w.reshape(co, kh * kw * cpad))


def pack_conv_weight(w_oihw: torch.Tensor, cin_pad: Optional[int] = None) -> torch.Tensor:
    """OIHW -> [Cout, kh*kw*Cin(_pad)] bf16, k = (r*kw + s)*Cin + c."""
    co, ci, kh, kw = w_oihw.shape
    w = w_oihw.permute(0, 2, 3, 1)
    if cin_pad is not None and cin_pad > ci:
        w = torch.nn.functional.pad(w, (0, cin_pad - ci))
    return w.reshape(co, -1).to(bf16).contiguous()


def pack_geglu_weight(w: torch.Tensor, b: torch.Tensor, tile: int = 160):
    """diffusers GEGLU proj [2*inner, K] (rows: value | gate) -> rows interleaved per N tile: [value half-tile | gate half-tile]."""
    inner = w.shape[0] // 2
    half = tile // 2
    assert inner % half == 0, (inner, tile)
    wv, wg = w[:inner].reshape(inner // half, half, -1), w[inner:].reshape(inner // half, half, -1)
    wi = torch.cat([wv, wg], dim=1).reshape(2 * inner, -1)
    return wi.to(bf16).contiguous(), b.float().contiguous()  # bias stays [value | gate]; the kernel indexes both halves


def pack_geglu_weight_tiled(w: torch.Tensor, b: torch.Tensor):
    wi, bi = pack_geglu_weight(w, b, tile=160)
    return tile_weight(wi, geglu=True), bi
