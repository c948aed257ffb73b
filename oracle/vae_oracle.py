"""Oracle for the VAE decoder called by the reference's scripts after the loop
(scripts/inference/generate_hdr.py:225-233; formal_baseline.py:228-237): plain-PyTorch restatement of the
decoder half of diffusers `AutoencoderKL` (SD1.5 config, scaling_factor 0.18215), diffusers key names.
TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (third-party diffusers class; see oracle/__init__.py)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .unet_oracle import ResnetBlock2D, Upsample2D

SCALING_FACTOR = 0.18215


class VaeAttention(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.group_norm = nn.GroupNorm(32, c, eps=1e-6)
        self.to_q = nn.Linear(c, c)
        self.to_k = nn.Linear(c, c)
        self.to_v = nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Identity()])

    def forward(self, x):
        b, c, h, w = x.shape
        y = self.group_norm(x).view(b, c, h * w).transpose(1, 2)
        q, k, v = self.to_q(y), self.to_k(y), self.to_v(y)
        s = torch.matmul(q, k.transpose(-1, -2)) * (c ** -0.5)
        o = torch.matmul(torch.softmax(s, dim=-1), v)
        o = self.to_out[0](o).transpose(1, 2).reshape(b, c, h, w)
        return o + x


class VaeMid(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, temb_dim=None, eps=1e-6), ResnetBlock2D(c, c, temb_dim=None, eps=1e-6)])
        self.attentions = nn.ModuleList([VaeAttention(c)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class VaeUp(nn.Module):
    def __init__(self, cin, cout, add_up):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb_dim=None, eps=1e-6) for i in range(3)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class Decoder(nn.Module):
    def __init__(self, ch=(128, 256, 512, 512), latent=4, out=3):
        super().__init__()
        rev = list(reversed(ch))
        self.conv_in = nn.Conv2d(latent, rev[0], 3, padding=1)
        self.mid_block = VaeMid(rev[0])
        ups, c = [], rev[0]
        for i, co in enumerate(rev):
            ups.append(VaeUp(c, co, add_up=i < len(rev) - 1))
            c = co
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(32, ch[0], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[0], out, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for u in self.up_blocks:
            x = u(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class VaeDecoderOracle(nn.Module):
    """`decode(z)` == AutoencoderKL.decode(z, return_dict=False)[0]; the caller divides by scaling_factor."""

    def __init__(self, ch=(128, 256, 512, 512)):
        super().__init__()
        self.post_quant_conv = nn.Conv2d(4, 4, 1)
        self.decoder = Decoder(ch)
        self.config = dict(scaling_factor=SCALING_FACTOR, block_out_channels=tuple(ch))

    def decode(self, z):
        return self.decoder(self.post_quant_conv(z))

    forward = decode
