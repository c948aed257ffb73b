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


class VaeDown(nn.Module):
    """diffusers DownEncoderBlock2D: resnets without temb, then Downsample2D(padding=0): F.pad(x, (0,1,0,1)) + 3x3 stride-2 conv."""

    def __init__(self, cin, cout, add_down):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb_dim=None, eps=1e-6) for i in range(2)])
        self.downsamplers = nn.ModuleList([nn.Module()]) if add_down else None
        if add_down:
            self.downsamplers[0].conv = nn.Conv2d(cout, cout, 3, stride=2, padding=0)

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0].conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))
        return x


class Encoder(nn.Module):
    def __init__(self, ch=(128, 256, 512, 512), cin=3, latent=4):
        super().__init__()
        self.conv_in = nn.Conv2d(cin, ch[0], 3, padding=1)
        downs, c = [], ch[0]
        for i, co in enumerate(ch):
            downs.append(VaeDown(c, co, add_down=i < len(ch) - 1))
            c = co
        self.down_blocks = nn.ModuleList(downs)
        self.mid_block = VaeMid(ch[-1])
        self.conv_norm_out = nn.GroupNorm(32, ch[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[-1], 2 * latent, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for d in self.down_blocks:
            x = d(x)
        return self.conv_out(F.silu(self.conv_norm_out(self.mid_block(x))))


class DiagonalGaussian:
    """diffusers DiagonalGaussianDistribution (`vae.encode(x).latent_dist`, generate_hdr.py:208)."""

    def __init__(self, moments):
        self.mean, logvar = torch.chunk(moments, 2, dim=1)
        self.logvar = torch.clamp(logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator=None, noise=None):
        if noise is None:
            noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean


class VaeOracle(nn.Module):
    """Full AutoencoderKL (SD1.5 config): `encode(x)` -> DiagonalGaussian, `decode(z)` as VaeDecoderOracle."""

    def __init__(self, ch=(128, 256, 512, 512)):
        super().__init__()
        self.encoder = Encoder(ch)
        self.quant_conv = nn.Conv2d(8, 8, 1)
        self.post_quant_conv = nn.Conv2d(4, 4, 1)
        self.decoder = Decoder(ch)
        self.config = dict(scaling_factor=SCALING_FACTOR, block_out_channels=tuple(ch))

    def moments(self, x):
        return self.quant_conv(self.encoder(x))

    def encode(self, x):
        return DiagonalGaussian(self.moments(x))

    def decode(self, z):
        return self.decoder(self.post_quant_conv(z))


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
