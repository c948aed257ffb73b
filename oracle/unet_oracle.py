"""Oracle for the UNet (kernels a + b): plain-PyTorch restatement of diffusers `UNet2DConditionModel` with the
reference's config literal (scripts/inference/generate_hdr.py:116-135; in_channels=8 variant :138-142 /
scripts/stage2/train_gm_unet.py:658-677).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the class lives in third-party diffusers (not in /root/reference).  The restatement uses
diffusers' state_dict key names so real checkpoints load, and is pinned by the
public parameter counts 859 520 964 (in=4) / 859 532 484 (in=8) in tests/test_unet_oracle.py.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

SD15_CONFIG = dict(
    sample_size=64, in_channels=4, out_channels=4, layers_per_block=2,
    block_out_channels=(320, 640, 1280, 1280),
    down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
    up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
    cross_attention_dim=768, attention_head_dim=8, norm_num_groups=32, norm_eps=1e-5,
    flip_sin_to_cos=True, freq_shift=0, act_fn="silu", time_cond_proj_dim=None,
)


def timestep_embedding(t: torch.Tensor, dim: int = 320) -> torch.Tensor:
    """diffusers get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0), fp32."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, cin, dim):
        super().__init__()
        self.linear_1 = nn.Linear(cin, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_dim=1280, groups=32, eps=1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_dim, cout) if temb_dim else None
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb=None):
        h = self.conv1(F.silu(self.norm1(x)))
        if self.time_emb_proj is not None:
            h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    def __init__(self, dim, ctx_dim=None, heads=8):
        super().__init__()
        ctx_dim = ctx_dim or dim
        self.heads = heads
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(ctx_dim, dim, bias=False)
        self.to_v = nn.Linear(ctx_dim, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Identity()])

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        b, n, c = x.shape
        h, d = self.heads, c // self.heads
        q = self.to_q(x).view(b, n, h, d).transpose(1, 2)
        k = self.to_k(ctx).view(b, -1, h, d).transpose(1, 2)
        v = self.to_v(ctx).view(b, -1, h, d).transpose(1, 2)
        s = torch.matmul(q, k.transpose(-1, -2)) * (d ** -0.5)
        o = torch.matmul(torch.softmax(s, dim=-1), v)
        o = o.transpose(1, 2).reshape(b, n, c)
        return self.to_out[0](o)


class GEGLU(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Identity(), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, ctx_dim, heads):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, ctx_dim, heads)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, ctx):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), ctx)
        x = x + self.ff(self.norm3(x))
        return x


class Transformer2DModel(nn.Module):
    def __init__(self, dim, ctx_dim, heads, groups=32):
        super().__init__()
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Conv2d(dim, dim, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, ctx_dim, heads)])
        self.proj_out = nn.Conv2d(dim, dim, 1)

    def forward(self, x, ctx):
        b, c, h, w = x.shape
        res = x
        y = self.proj_in(self.norm(x))
        y = y.permute(0, 2, 3, 1).reshape(b, h * w, c)
        for blk in self.transformer_blocks:
            y = blk(y, ctx)
        y = y.reshape(b, h, w, c).permute(0, 3, 1, 2)
        return self.proj_out(y) + res


class Downsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cin, cout, attn, ctx_dim, heads, add_down, temb=1280):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb) for i in range(2)])
        if attn:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, ctx_dim, heads) for _ in range(2)])
        else:
            self.attentions = None
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, x, temb, ctx):
        outs = []
        for i, r in enumerate(self.resnets):
            x = r(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, ctx)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, c, ctx_dim, heads, temb=1280):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, temb), ResnetBlock2D(c, c, temb)])
        self.attentions = nn.ModuleList([Transformer2DModel(c, ctx_dim, heads)])

    def forward(self, x, temb, ctx):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, ctx)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    def __init__(self, cin, cout, cprev, attn, ctx_dim, heads, add_up, temb=1280):
        super().__init__()
        res = []
        for i in range(3):
            skip = cin if i == 2 else cout
            rin = cprev if i == 0 else cout
            res.append(ResnetBlock2D(rin + skip, cout, temb))
        self.resnets = nn.ModuleList(res)
        self.attentions = nn.ModuleList([Transformer2DModel(cout, ctx_dim, heads) for _ in range(3)]) if attn else None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x, skips, temb, ctx):
        for i, r in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = r(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, ctx)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class UNet2DConditionOracle(nn.Module):
    """SD1.5-architecture epsilon predictor; `forward(sample, t, encoder_hidden_states)` -> eps (same shape
    conventions as the call at stable_diffusion_dual_unet.py:1052-1060 with return_dict=False -> [0])."""

    def __init__(self, in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280),
                 cross_attention_dim=768, heads=8):
        super().__init__()
        ch = list(block_out_channels)
        self.config = dict(SD15_CONFIG, in_channels=in_channels, out_channels=out_channels,
                           block_out_channels=tuple(ch), cross_attention_dim=cross_attention_dim)
        self.in_channels = in_channels
        temb = ch[0] * 4
        self.conv_in = nn.Conv2d(in_channels, ch[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(ch[0], temb)
        downs = []
        cout = ch[0]
        for i in range(4):
            cin, cout = cout, ch[i]
            downs.append(DownBlock(cin, cout, attn=i < 3, ctx_dim=cross_attention_dim, heads=heads, add_down=i < 3, temb=temb))
        self.down_blocks = nn.ModuleList(downs)
        self.mid_block = MidBlock(ch[-1], cross_attention_dim, heads, temb)
        rev = list(reversed(ch))
        ups = []
        cout = rev[0]
        for i in range(4):
            cprev = cout
            cout = rev[i]
            cin = rev[min(i + 1, 3)]
            ups.append(UpBlock(cin, cout, cprev, attn=i > 0, ctx_dim=cross_attention_dim, heads=heads, add_up=i < 3, temb=temb))
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(32, ch[0], eps=1e-5)
        self.conv_out = nn.Conv2d(ch[0], out_channels, 3, padding=1)

    def forward(self, sample, timestep, encoder_hidden_states, **kw):
        b = sample.shape[0]
        t = torch.as_tensor(timestep, device=sample.device).reshape(-1).expand(b)
        temb = self.time_embedding(timestep_embedding(t, self.conv_in.out_channels).to(sample.dtype))
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, temb, encoder_hidden_states)
            skips.extend(outs)
        x = self.mid_block(x, temb, encoder_hidden_states)
        for blk in self.up_blocks:
            x = blk(x, skips, temb, encoder_hidden_states)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        return x


def widen_conv_in(unet4: UNet2DConditionOracle) -> UNet2DConditionOracle:
    """scripts/stage2/train_gm_unet.py:658-677 `_replace_unet_conv_in`: 4 -> 8 input channels, the old
    kernel tiled twice along the input dim and halved."""
    u8 = UNet2DConditionOracle(in_channels=8, block_out_channels=unet4.config["block_out_channels"],
                               cross_attention_dim=unet4.config["cross_attention_dim"])
    sd = {k: v.clone() for k, v in unet4.state_dict().items()}
    w = sd["conv_in.weight"]
    sd["conv_in.weight"] = torch.cat([w, w], dim=1) * 0.5
    u8.load_state_dict(sd)
    return u8


def count_params(m: nn.Module) -> int:
    return sum(p.numel() for p in m.parameters())
