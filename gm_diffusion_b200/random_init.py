"""Random-init weights in diffusers state_dict form for the SD1.5-architecture UNet (config literal
scripts/inference/generate_hdr.py:116-135) and the SD VAE decoder.  There is no network for checkpoints, so the
benchmark and smoke test run the real architecture on synthetic weights (BASELINE.json: "random-init SD1.5-arch
UNet"); real checkpoints load through the same `B200UNet(state_dict)` / `B200VaeDecoder(state_dict)` path.
Initialisation follows torch's Linear/Conv2d defaults (U(-1/sqrt(fan_in), +1/sqrt(fan_in)) for weight and bias;
norm affine = (1, 0)), generated directly on the target device."""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch


class _Gen:
    def __init__(self, seed: int, device):
        self.device = torch.device(device)
        self.g = torch.Generator(device=self.device).manual_seed(seed)
        self.sd: Dict[str, torch.Tensor] = {}

    def _u(self, shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=self.g, device=self.device) * 2 - 1) * b

    def linear(self, name, cin, cout, bias=True):
        self.sd[name + ".weight"] = self._u((cout, cin), cin)
        if bias:
            self.sd[name + ".bias"] = self._u((cout,), cin)

    def conv(self, name, cin, cout, k):
        self.sd[name + ".weight"] = self._u((cout, cin, k, k), cin * k * k)
        self.sd[name + ".bias"] = self._u((cout,), cin * k * k)

    def norm(self, name, c):
        self.sd[name + ".weight"] = torch.ones(c, device=self.device)
        self.sd[name + ".bias"] = torch.zeros(c, device=self.device)

    def resnet(self, p, cin, cout, temb):
        self.norm(p + "norm1", cin)
        self.conv(p + "conv1", cin, cout, 3)
        if temb:
            self.linear(p + "time_emb_proj", temb, cout)
        self.norm(p + "norm2", cout)
        self.conv(p + "conv2", cout, cout, 3)
        if cin != cout:
            self.conv(p + "conv_shortcut", cin, cout, 1)

    def transformer(self, p, c, ctx):
        self.norm(p + "norm", c)
        self.conv(p + "proj_in", c, c, 1)
        t = p + "transformer_blocks.0."
        for i in (1, 2, 3):
            self.norm(t + f"norm{i}", c)
        for a, kd in (("attn1", c), ("attn2", ctx)):
            self.linear(t + a + ".to_q", c, c, bias=False)
            self.linear(t + a + ".to_k", kd, c, bias=False)
            self.linear(t + a + ".to_v", kd, c, bias=False)
            self.linear(t + a + ".to_out.0", c, c)
        self.linear(t + "ff.net.0.proj", c, 8 * c)
        self.linear(t + "ff.net.2", 4 * c, c)
        self.conv(p + "proj_out", c, c, 1)


def sd15_unet_state_dict(in_channels: int = 4, seed: int = 0, device="cuda", block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280),
                         cross_attention_dim: int = 768, out_channels: int = 4) -> Dict[str, torch.Tensor]:
    g = _Gen(seed, device)
    ch = list(block_out_channels)
    temb = ch[0] * 4
    g.conv("conv_in", in_channels, ch[0], 3)
    g.linear("time_embedding.linear_1", ch[0], temb)
    g.linear("time_embedding.linear_2", temb, temb)
    cout = ch[0]
    for i in range(4):
        cin, cout = cout, ch[i]
        for j in range(2):
            g.resnet(f"down_blocks.{i}.resnets.{j}.", cin if j == 0 else cout, cout, temb)
            if i < 3:
                g.transformer(f"down_blocks.{i}.attentions.{j}.", cout, cross_attention_dim)
        if i < 3:
            g.conv(f"down_blocks.{i}.downsamplers.0.conv", cout, cout, 3)
    g.resnet("mid_block.resnets.0.", ch[-1], ch[-1], temb)
    g.transformer("mid_block.attentions.0.", ch[-1], cross_attention_dim)
    g.resnet("mid_block.resnets.1.", ch[-1], ch[-1], temb)
    rev = list(reversed(ch))
    cout = rev[0]
    for i in range(4):
        cprev, cout, cin = cout, rev[i], rev[min(i + 1, 3)]
        for j in range(3):
            skip = cin if j == 2 else cout
            rin = cprev if j == 0 else cout
            g.resnet(f"up_blocks.{i}.resnets.{j}.", rin + skip, cout, temb)
            if i > 0:
                g.transformer(f"up_blocks.{i}.attentions.{j}.", cout, cross_attention_dim)
        if i < 3:
            g.conv(f"up_blocks.{i}.upsamplers.0.conv", cout, cout, 3)
    g.norm("conv_norm_out", ch[0])
    g.conv("conv_out", ch[0], out_channels, 3)
    return g.sd


def sd_vae_decoder_state_dict(seed: int = 0, device="cuda", block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)) -> Dict[str, torch.Tensor]:
    g = _Gen(seed, device)
    rev = list(reversed(block_out_channels))
    g.conv("post_quant_conv", 4, 4, 1)
    d = "decoder."
    g.conv(d + "conv_in", 4, rev[0], 3)
    g.resnet(d + "mid_block.resnets.0.", rev[0], rev[0], 0)
    a = d + "mid_block.attentions.0."
    g.norm(a + "group_norm", rev[0])
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        g.linear(a + n, rev[0], rev[0])
    g.resnet(d + "mid_block.resnets.1.", rev[0], rev[0], 0)
    c = rev[0]
    for i, co in enumerate(rev):
        for j in range(3):
            g.resnet(d + f"up_blocks.{i}.resnets.{j}.", c if j == 0 else co, co, 0)
        if i < len(rev) - 1:
            g.conv(d + f"up_blocks.{i}.upsamplers.0.conv", co, co, 3)
        c = co
    g.norm(d + "conv_norm_out", block_out_channels[0])
    g.conv(d + "conv_out", block_out_channels[0], 3, 3)
    return g.sd


def sd_vae_state_dict(seed: int = 0, device="cuda", block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)) -> Dict[str, torch.Tensor]:
    """Full AutoencoderKL (encoder + quant_conv + decoder) with diffusers key names."""
    sd = sd_vae_decoder_state_dict(seed, device, block_out_channels)
    g = _Gen(seed + 1, device)
    ch = block_out_channels
    e = "encoder."
    g.conv(e + "conv_in", 3, ch[0], 3)
    c = ch[0]
    for i, co in enumerate(ch):
        for j in range(2):
            g.resnet(e + f"down_blocks.{i}.resnets.{j}.", c if j == 0 else co, co, 0)
        if i < len(ch) - 1:
            g.conv(e + f"down_blocks.{i}.downsamplers.0.conv", co, co, 3)
        c = co
    g.resnet(e + "mid_block.resnets.0.", c, c, 0)
    a = e + "mid_block.attentions.0."
    g.norm(a + "group_norm", c)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        g.linear(a + n, c, c)
    g.resnet(e + "mid_block.resnets.1.", c, c, 0)
    g.norm(e + "conv_norm_out", c)
    g.conv(e + "conv_out", c, 8, 3)
    g.conv("quant_conv", 8, 8, 1)
    sd.update(g.sd)
    return sd


def widen_conv_in_state_dict(sd4: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """scripts/stage2/train_gm_unet.py:658-677 `_replace_unet_conv_in`: 4 -> 8 input channels (tiled x0.5)."""
    sd = dict(sd4)
    w = sd4["conv_in.weight"]
    sd["conv_in.weight"] = torch.cat([w, w], dim=1) * 0.5
    return sd
