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


class _Config(dict):
    """diffusers FrozenDict look-alike: `vae.config.scaling_factor` (generate_hdr.py:209) and `vae.config["scaling_factor"]`."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class _VaeAttention:
    """Single-head, d = C mid-block attention of AutoencoderKL (encoder and decoder): two batched tcgen05 GEMMs around a row softmax."""

    def __init__(self, sd, a, dev):
        self.a_norm = (_f32(sd[a + "group_norm.weight"], dev), _f32(sd[a + "group_norm.bias"], dev))
        self.w_qk = _w(torch.cat([sd[a + "to_q.weight"], sd[a + "to_k.weight"]], 0), dev)
        self.b_qk = _f32(torch.cat([sd[a + "to_q.bias"], sd[a + "to_k.bias"]], 0), dev)
        # to_v is used as the A operand (V^T = Wv . y^T), so it stays in the plain [N, K] layout
        self.w_v, self.b_v = sd[a + "to_v.weight"].detach().to(device=dev, dtype=bf16).contiguous(), _f32(sd[a + "to_v.bias"], dev)
        self.w_o, self.b_o = _w(sd[a + "to_out.0.weight"], dev), _f32(sd[a + "to_out.0.bias"], dev)

    def __call__(self, x, ws, xs=None):
        """xs: GroupNorm statistics of x from its producer's epilogue (or None).  Returns (output, statistics of the output)."""
        B, H, W, c = x.shape
        n, m = H * W, B * H * W
        y = ops.groupnorm_silu(x, *self.a_norm, eps=1e-6, silu=False, stats_ws=ws, sums=xs).view(m, c)
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
        out, os_ = ops.gemm(o, self.w_o, bias=self.b_o, residual=x.view(m, c), gn_rows_per_sample=n)
        return out.view(B, H, W, c), os_


class B200VaeDecoder:
    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda", scaling_factor: float = SCALING_FACTOR):
        sd, dev = state_dict, torch.device(device)
        self.device = dev
        self.scaling_factor = scaling_factor
        self.config = _Config(scaling_factor=scaling_factor, latent_channels=4, block_out_channels=(128, 256, 512, 512))
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
        self.mid_attn = _VaeAttention(sd, d + "mid_block.attentions.0.", dev)
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

    def _attention(self, x, xs=None):
        return self.mid_attn(x, self._gn_ws, xs)


    @L.on_own_device
    def decode_px(self, latents_px: torch.Tensor, B: int, h: int, w: int) -> torch.Tensor:
        """latents_px: fp32 pixel-major [B*h*w, 4] SCALED latents (the division by scaling_factor is folded into
        post_quant_conv).  Returns the decoder image, bf16 NHWC [B, 8h, 8w, 3], range ~[-1,1]."""
        n_px = B * h * w
        z = torch.empty((n_px, 8), dtype=bf16, device=self.device)
        L.check(L.lib().gmd_pack_unet_input(latents_px.data_ptr(), None, z.data_ptr(), n_px, 8, L.current_stream()), "gmd_pack_unet_input")
        z = ops.gemm(z, self.w_pq, bias=self.b_pq).view(B, h, w, 8)
        ws = self._gn_ws
        hints = self.__dict__.setdefault("_gn_words", {})
        with ops.gn_arena(self.device, hints.get(("dec", B, h, w))) as arena:   # one memset zeroes every GroupNorm accumulator of the pass
            # every activation travels with the GroupNorm statistics its producer's epilogue formed (None where unavailable)
            x, xs = ops.conv2d(z, self.w_in, self.c_mid, bias=self.b_in, gn_stats=True)
            x, xs = self.mid_res[0](x, None, None, ws, xs=xs)
            x, xs = self._attention(x, xs)
            x, xs = self.mid_res[1](x, None, None, ws, xs=xs)
            for res, us in self.ups:
                for r in res:
                    x, xs = r(x, None, None, ws, xs=xs)
                if us is not None:
                    x, xs = ops.conv2d(x, us[0], us[0].shape[0], upsample=True, bias=us[1], gn_stats=True)
            x = ops.groupnorm_silu(x, *self.n_out, eps=1e-6, stats_ws=ws, sums=xs)
            img = ops.conv2d(x, self.w_out, self.w_out.shape[0], bias=self.b_out)
        hints[("dec", B, h, w)] = max(arena.used, hints.get(("dec", B, h, w), 0))
        return img

    @torch.no_grad()
    @L.on_own_device
    def decode(self, z_nchw: torch.Tensor) -> torch.Tensor:
        """diffusers convention: `vae.decode(latents / scaling_factor)` -> image NCHW (model dtype bf16 here).
        `z_nchw` is the UNSCALED latent, exactly what the reference passes at generate_hdr.py:226,231."""
        B, c, h, w = z_nchw.shape
        zs = (z_nchw.to(self.device, torch.float32) * self.scaling_factor).contiguous()  # undo: the fold divides again
        px = torch.empty((B * h * w, 4), dtype=torch.float32, device=self.device)
        L.check(L.lib().gmd_latents_nchw_to_px(zs.data_ptr(), px.data_ptr(), B, h * w, L.current_stream()), "gmd_latents_nchw_to_px")
        return self.decode_px(px, B, h, w).permute(0, 3, 1, 2)


def _randn(shape, generator, device):
    """diffusers randn_tensor: a CPU generator draws on the CPU and the sample moves to the device."""
    gdev = generator.device if generator is not None else torch.device(device)
    where = "cpu" if gdev.type == "cpu" and torch.device(device).type != "cpu" else device
    return torch.randn(shape, generator=generator, device=where, dtype=torch.float32).to(device)


class LatentDistribution:
    """diffusers DiagonalGaussianDistribution over the encoder moments (`vae.encode(x).latent_dist`, generate_hdr.py:208).
    `mean` / `logvar` are NCHW views of the fp32 moments; `sample` and `mode` run `gmd_vae_sample`."""

    def __init__(self, moments_nhwc: torch.Tensor):
        self.parameters = moments_nhwc                       # fp32 [B, h, w, 8] = (mean[4], logvar[4])
        m = moments_nhwc.permute(0, 3, 1, 2)
        self.mean, self.logvar = m[:, :4], m[:, 4:]

    def _draw(self, noise_nchw: Optional[torch.Tensor]) -> torch.Tensor:
        B, h, w, _ = self.parameters.shape
        n_px, dev = B * h * w, self.parameters.device
        stream = L.current_stream()
        npx = None
        if noise_nchw is not None:
            npx = torch.empty((n_px, 4), dtype=torch.float32, device=dev)
            L.check(L.lib().gmd_latents_nchw_to_px(noise_nchw.contiguous().data_ptr(), npx.data_ptr(), B, h * w, stream), "gmd_latents_nchw_to_px")
        px = torch.empty((n_px, 4), dtype=torch.float32, device=dev)
        L.check(L.lib().gmd_vae_sample(self.parameters.data_ptr(), L.ptr(npx), px.data_ptr(), n_px, 1.0, stream), "gmd_vae_sample")
        out = torch.empty((B, 4, h, w), dtype=torch.float32, device=dev)
        L.check(L.lib().gmd_latents_px_to_nchw(px.data_ptr(), out.data_ptr(), B, h * w, stream), "gmd_latents_px_to_nchw")
        return out

    def sample(self, generator: Optional[torch.Generator] = None, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        if noise is None:
            noise = _randn(tuple(self.mean.shape), generator, self.parameters.device)
        return self._draw(noise.to(self.parameters.device, torch.float32))

    def mode(self) -> torch.Tensor:
        return self._draw(None)


class EncoderOutput:
    def __init__(self, latent_dist: LatentDistribution):
        self.latent_dist = latent_dist

    def __getitem__(self, i):  # `vae.encode(x, return_dict=False)[0]`
        return (self.latent_dist,)[i]


class B200Vae(B200VaeDecoder):
    """Decoder + encoder of AutoencoderKL on the same kernels.  The encoder is the entry of the reference's SDR->HDR CLI
    (`pipeline.vae.encode(sdr_image).latent_dist.sample() * scaling_factor`, scripts/inference/generate_hdr.py:207-209):
    conv_in, 4 x [2 resnets, stride-2 conv with the bottom/right zero pad], mid (resnet, attention, resnet), GN+SiLU, conv_out
    with quant_conv (1x1, 8->8) folded into its weights; the moments leave the last conv in fp32."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda", scaling_factor: float = SCALING_FACTOR):
        super().__init__(state_dict, device=device, scaling_factor=scaling_factor)
        sd, dev = state_dict, self.device
        e = "encoder."
        self.e_w_in = ops.pack_conv_weight_tiled(sd[e + "conv_in.weight"].to(dev), cin_pad=8)
        self.e_b_in = _f32(sd[e + "conv_in.bias"], dev)
        self.e_c0 = sd[e + "conv_in.weight"].shape[0]
        self.e_downs = []
        i = 0
        while e + f"down_blocks.{i}.resnets.0.conv1.weight" in sd:
            res, j = [], 0
            while e + f"down_blocks.{i}.resnets.{j}.conv1.weight" in sd:
                res.append(_Resnet(sd, e + f"down_blocks.{i}.resnets.{j}.", dev, None, eps=1e-6))
                j += 1
            ds = None
            if e + f"down_blocks.{i}.downsamplers.0.conv.weight" in sd:
                ds = (ops.pack_conv_weight_tiled(sd[e + f"down_blocks.{i}.downsamplers.0.conv.weight"].to(dev)),
                      _f32(sd[e + f"down_blocks.{i}.downsamplers.0.conv.bias"], dev))
            self.e_downs.append((res, ds))
            i += 1
        self.e_mid_res = [_Resnet(sd, e + f"mid_block.resnets.{j}.", dev, None, eps=1e-6) for j in range(2)]
        self.e_attn = _VaeAttention(sd, e + "mid_block.attentions.0.", dev)
        self.e_n_out = (_f32(sd[e + "conv_norm_out.weight"], dev), _f32(sd[e + "conv_norm_out.bias"], dev))
        # quant_conv(conv_out(x)) == one 3x3 conv with W' = Wq . Wc and b' = Wq . bc + bq (exact in exact arithmetic)
        wq = sd["quant_conv.weight"].reshape(8, 8).to(dev, torch.float32)
        wc = sd[e + "conv_out.weight"].to(dev, torch.float32)
        w = torch.einsum("om,mcrs->ocrs", wq, wc)
        b = wq @ sd[e + "conv_out.bias"].to(dev, torch.float32) + sd["quant_conv.bias"].to(dev, torch.float32)
        self.e_w_out, self.e_b_out = ops.pack_conv_weight_tiled(w), _f32(b, dev)

    @L.on_own_device
    def moments(self, image: torch.Tensor) -> torch.Tensor:
        """image: [B,3,H,W] in [-1,1] (any float dtype) -> fp32 moments NHWC [B, H/8, W/8, 8]."""
        B, c, H, W = image.shape
        if c != 3 or H % 8 or W % 8:
            raise ValueError(f"encode expects [B,3,H,W] with H and W divisible by 8, got {tuple(image.shape)}")
        img = image.to(self.device, torch.float32).contiguous()
        x = torch.empty((B, H, W, 8), dtype=bf16, device=self.device)
        L.check(L.lib().gmd_pack_image_nchw(img.data_ptr(), x.data_ptr(), B, H * W, 3, L.current_stream()), "gmd_pack_image_nchw")
        ws = self._gn_ws
        hints = self.__dict__.setdefault("_gn_words", {})
        with ops.gn_arena(self.device, hints.get(("enc", B, H, W))) as arena:
            x, xs = ops.conv2d(x, self.e_w_in, self.e_c0, bias=self.e_b_in, gn_stats=True)
            for res, ds in self.e_downs:
                for r in res:
                    x, xs = r(x, None, None, ws, xs=xs)
                if ds is not None:
                    x, xs = ops.conv2d(x, ds[0], ds[0].shape[0], stride=2, pad_end=True, bias=ds[1], gn_stats=True)
            x, xs = self.e_mid_res[0](x, None, None, ws, xs=xs)
            x, xs = self.e_attn(x, ws, xs)
            x, xs = self.e_mid_res[1](x, None, None, ws, xs=xs)
            x = ops.groupnorm_silu(x, *self.e_n_out, eps=1e-6, stats_ws=ws, sums=xs)
            mom = ops.conv2d(x, self.e_w_out, 8, bias=self.e_b_out, out_f32=True)
        hints[("enc", B, H, W)] = max(arena.used, hints.get(("enc", B, H, W), 0))
        return mom

    @torch.no_grad()
    @L.on_own_device
    def encode(self, image: torch.Tensor, return_dict: bool = True):
        out = EncoderOutput(LatentDistribution(self.moments(image)))
        return out if return_dict else (out.latent_dist,)

    @classmethod
    def from_module(cls, module, device="cuda", **kw):
        return super().from_module(module, device=device, **kw)


__all__ = ["B200VaeDecoder", "B200Vae", "LatentDistribution"]
