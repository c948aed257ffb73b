"""B200 executor for the SD VAE decoder the reference's callers run after the loop
(`vae.decode(latent / 0.18215)` scripts/inference/generate_hdr.py:225-233; formal_baseline.py:228-237;
diffusers AutoencoderKL decoder).  Same kernels as the UNet: tcgen05 implicit-GEMM
convs (nearest-2x upsample folded into the gather), GroupNorm+SiLU, and the single-head d=512 mid-block
attention as two batched tcgen05 GEMMs around a row softmax.  Output is the raw decoder image in [-1,1],
bf16 NHWC, which kernel (d) consumes directly (de-normalise + Eq.(1) fused)."""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib as L
from . import ops
from .unet import _Resnet, _f32, _w

bf16 = torch.bfloat16
SCALING_FACTOR = 0.18215


class B200VaeDecoder:
    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda", scaling_factor: float = SCALING_FACTOR):
        sd, dev = state_dict, torch.device(device)
        self.device = dev
        self.scaling_factor = scaling_factor
        self.config = dict(scaling_factor=scaling_factor)
        # post_quant_conv (1x1, 4->4) with 1/scaling_factor folded in; in/out channels padded to 8
        wpq = torch.zeros(8, 8)
        wpq[:4, :4] = sd["post_quant_conv.weight"].reshape(4, 4).float() / scaling_factor
        bpq = torch.zeros(8)
        bpq[:4] = sd["post_quant_conv.bias"].float()
        self.w_pq, self.b_pq = _w(wpq, dev), _f32(bpq, dev)
        d = "decoder."
        self.w_in = ops.pack_conv_weight_tiled(sd[d + "conv_in.weight"].to(dev), cin_pad=8)
        self.b_in = _f32(sd[d + "conv_in.bias"], dev)
        self.c_mid = sd[d + "conv_in.weight"].shape[0]
        self.mid_res = [_Resnet(sd, d + f"mid_block.resnets.{j}.", dev, None, eps=1e-6) for j in range(2)]
        a = d + "mid_block.attentions.0."
        self.a_norm = (_f32(sd[a + "group_norm.weight"], dev), _f32(sd[a + "group_norm.bias"], dev))
        self.w_qk = _w(torch.cat([sd[a + "to_q.weight"], sd[a + "to_k.weight"]], 0), dev)
        self.b_qk = _f32(torch.cat([sd[a + "to_q.bias"], sd[a + "to_k.bias"]], 0), dev)
        # to_v is used as the A operand (V^T = Wv . y^T), so it stays in the plain [N, K] layout
        self.w_v, self.b_v = sd[a + "to_v.weight"].detach().to(device=dev, dtype=bf16).contiguous(), _f32(sd[a + "to_v.bias"], dev)
        self.w_o, self.b_o = _w(sd[a + "to_out.0.weight"], dev), _f32(sd[a + "to_out.0.bias"], dev)
        self.ups = []
        i = 0
        while d + f"up_blocks.{i}.resnets.0.conv1.weight" in sd:
            res = []
            j = 0
            while d + f"up_blocks.{i}.resnets.{j}.conv1.weight" in sd:
                res.append(_Resnet(sd, d + f"up_blocks.{i}.resnets.{j}.", dev, None, eps=1e-6))
                j += 1
            us = None
            if d + f"up_blocks.{i}.upsamplers.0.conv.weight" in sd:
                us = (ops.pack_conv_weight_tiled(sd[d + f"up_blocks.{i}.upsamplers.0.conv.weight"].to(dev)),
                      _f32(sd[d + f"up_blocks.{i}.upsamplers.0.conv.bias"], dev))
            self.ups.append((res, us))
            i += 1
        self.n_out = (_f32(sd[d + "conv_norm_out.weight"], dev), _f32(sd[d + "conv_norm_out.bias"], dev))
        self.w_out = ops.pack_conv_weight_tiled(sd[d + "conv_out.weight"].to(dev))
        self.b_out = _f32(sd[d + "conv_out.bias"], dev)
        self._gn_ws = ops.gn_workspace(64, 32, dev)  # up to 64 samples per forward

    @classmethod
    def from_module(cls, module, device="cuda", **kw) -> "B200VaeDecoder":
        sf = kw.pop("scaling_factor", None)
        if sf is None:
            cfg = getattr(module, "config", None)
            sf = (cfg.get("scaling_factor") if isinstance(cfg, dict) else getattr(cfg, "scaling_factor", None)) or SCALING_FACTOR
        return cls({k: v for k, v in module.state_dict().items()}, device=device, scaling_factor=sf, **kw)

    def _attention(self, x):
        B, H, W, c = x.shape
        n, m = H * W, B * H * W
        ws = self._gn_ws
        y = ops.groupnorm_silu(x, *self.a_norm, eps=1e-6, silu=False, stats_ws=ws).view(m, c)
        qk = ops.gemm(y, self.w_qk, bias=self.b_qk).view(B, n, 2 * c)
        # V^T per image straight out of a GEMM with swapped operand roles: Vt[c, token] = Wv[c,:] . y[token,:]
        vt = torch.empty((B, c, n), dtype=bf16, device=x.device)
        y3 = y.view(B, n, c)
        for b in range(B):
            ops.gemm(self.w_v, y3[b], out=vt[b])
        s = ops.gemm(qk[..., :c], qk[..., c:])                     # [B, n, n] logits
        p = ops.softmax_rows(s, c ** -0.5, out=s)
        # rows of P sum to 1, so the V bias passes through the attention average unchanged: add it after P.V
        o = ops.gemm(p, vt, bias=self.b_v).view(m, c)
        return ops.gemm(o, self.w_o, bias=self.b_o, residual=x.view(m, c)).view(B, H, W, c)

    def decode_px(self, latents_px: torch.Tensor, B: int, h: int, w: int) -> torch.Tensor:
        """latents_px: fp32 pixel-major [B*h*w, 4] SCALED latents (the division by scaling_factor is folded into
        post_quant_conv).  Returns the decoder image, bf16 NHWC [B, 8h, 8w, 3], range ~[-1,1]."""
        n_px = B * h * w
        z = torch.empty((n_px, 8), dtype=bf16, device=self.device)
        L.check(L.lib().gmd_pack_unet_input(latents_px.data_ptr(), None, z.data_ptr(), n_px, 8, L.current_stream()), "gmd_pack_unet_input")
        z = ops.gemm(z, self.w_pq, bias=self.b_pq).view(B, h, w, 8)
        ws = self._gn_ws
        x = ops.conv2d(z, self.w_in, self.c_mid, bias=self.b_in)
        x = self.mid_res[0](x, None, None, ws)
        x = self._attention(x)
        x = self.mid_res[1](x, None, None, ws)
        for res, us in self.ups:
            for r in res:
                x = r(x, None, None, ws)
            if us is not None:
                x = ops.conv2d(x, us[0], us[0].shape[0], upsample=True, bias=us[1])
        x = ops.groupnorm_silu(x, *self.n_out, eps=1e-6, stats_ws=ws)
        return ops.conv2d(x, self.w_out, self.w_out.shape[0], bias=self.b_out)

    @torch.no_grad()
    def decode(self, z_nchw: torch.Tensor) -> torch.Tensor:
        """diffusers convention: `vae.decode(latents / scaling_factor)` -> image NCHW (model dtype bf16 here).
        `z_nchw` is the UNSCALED latent, exactly what the reference passes at generate_hdr.py:226,231."""
        B, c, h, w = z_nchw.shape
        zs = (z_nchw.to(self.device, torch.float32) * self.scaling_factor).contiguous()  # undo: the fold divides again
        px = torch.empty((B * h * w, 4), dtype=torch.float32, device=self.device)
        L.check(L.lib().gmd_latents_nchw_to_px(zs.data_ptr(), px.data_ptr(), B, h * w, L.current_stream()), "gmd_latents_nchw_to_px")
        return self.decode_px(px, B, h, w).permute(0, 3, 1, 2)
