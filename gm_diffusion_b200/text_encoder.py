"""CLIP text encoder (the `text_encoder` the reference pipelines call at stable_diffusion_dual_unet.py:400-427 — transformers'
`CLIPTextModel`, ViT-L/14 text tower for SD1.5: 12 layers, width 768, 12 heads of 64, 77 tokens, causal attention, quick-GELU MLP) on
the library's kernels: LayerNorm, tcgen05 GEMMs (QKV fused; per-head score and P.V products as batched GEMMs;
V^T straight out of a GEMM with swapped operand roles), the masked row softmax, SiLU.  quick_gelu(x) = x sigmoid(1.702 x) is SiLU with
1.702 folded into fc1 and 1/1.702 into fc2; the V bias passes through the attention average unchanged (rows of P sum to 1) and is
folded into the out-projection bias.  The token stream stays fp32 like the UNet's.

Accepts the module itself (`B200ClipTextEncoder.from_module(clip_text_model)`) and answers the two call forms the pipelines use:
`enc(ids)[0]` and `enc(ids, output_hidden_states=True)[-1][-(clip_skip + 1)]` + `enc.text_model.final_layer_norm(h)`."""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List

import torch

from . import _lib as L
from . import ops

bf16 = torch.bfloat16


def _f32(t, dev):
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


class _Layer:
    def __init__(self, sd: Dict[str, torch.Tensor], pre: str, dev):
        g = lambda k: sd[pre + k].detach().to(dev, torch.float32)
        self.ln1 = (_f32(g("layer_norm1.weight"), dev), _f32(g("layer_norm1.bias"), dev))
        self.ln2 = (_f32(g("layer_norm2.weight"), dev), _f32(g("layer_norm2.bias"), dev))
        wq, wk, wv, wo = g("self_attn.q_proj.weight"), g("self_attn.k_proj.weight"), g("self_attn.v_proj.weight"), g("self_attn.out_proj.weight")
        self.w_qk = ops.tile_weight(torch.cat([wq, wk], 0).to(bf16))
        self.b_qk = torch.cat([g("self_attn.q_proj.bias"), g("self_attn.k_proj.bias")]).contiguous()
        self.w_v = wv.to(bf16).contiguous()          # A operand of the V^T GEMM: plain [768, 768]
        self.w_o = ops.tile_weight(wo.to(bf16))
        self.b_o = (g("self_attn.out_proj.bias") + wo @ g("self_attn.v_proj.bias")).contiguous()
        self.w_fc1 = ops.tile_weight((1.702 * g("mlp.fc1.weight")).to(bf16))
        self.b_fc1 = (1.702 * g("mlp.fc1.bias")).contiguous()
        self.w_fc2 = ops.tile_weight((g("mlp.fc2.weight") / 1.702).to(bf16))
        self.b_fc2 = g("mlp.fc2.bias").contiguous()


class B200ClipTextEncoder:
    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda", heads: int = 12, eps: float = 1e-5):
        dev = torch.device(device)
        self.device, self.heads, self.eps = dev, heads, eps
        sd = {k[len("text_model."):] if k.startswith("text_model.") else k: v for k, v in state_dict.items()}
        self.tok = _f32(sd["embeddings.token_embedding.weight"], dev)
        self.pos = _f32(sd["embeddings.position_embedding.weight"], dev)
        self.width = self.tok.shape[1]
        if self.width % (8 * heads):
            raise NotImplementedError("head dim must be a multiple of 8")
        self.layers: List[_Layer] = []
        i = 0
        while f"encoder.layers.{i}.layer_norm1.weight" in sd:
            self.layers.append(_Layer(sd, f"encoder.layers.{i}.", dev))
            i += 1
        self.ln_f = (_f32(sd["final_layer_norm.weight"], dev), _f32(sd["final_layer_norm.bias"], dev))
        # the duck-typed surface the reference pipelines touch
        self.text_model = SimpleNamespace(final_layer_norm=self.final_layer_norm)
        self.config = SimpleNamespace(hidden_size=self.width, num_hidden_layers=len(self.layers), max_position_embeddings=self.pos.shape[0])
        self.dtype = torch.float32

    @classmethod
    def from_module(cls, module, device="cuda", **kw) -> "B200ClipTextEncoder":
        cfg = getattr(module, "config", None)
        act = getattr(cfg, "hidden_act", "quick_gelu")
        if act != "quick_gelu":
            raise NotImplementedError(f"CLIP text encoder with hidden_act={act!r} (SD1.5's is quick_gelu)")
        return cls({k: v for k, v in module.state_dict().items()}, device=device, heads=getattr(cfg, "num_attention_heads", 12),
                   eps=getattr(cfg, "layer_norm_eps", 1e-5), **kw)

    def parameters(self):
        yield self.tok

    def final_layer_norm(self, h: torch.Tensor) -> torch.Tensor:
        shp = h.shape
        return ops.layernorm(h.to(self.device, torch.float32).reshape(-1, shp[-1]).contiguous(), *self.ln_f, eps=self.eps).view(shp).float()

    @torch.no_grad()
    @L.on_own_device
    def __call__(self, input_ids: torch.Tensor, attention_mask=None, output_hidden_states: bool = False, **kw):
        """input_ids [B, T] (T <= max positions; the pipelines pad to 77).  Returns a tuple-like: [0] last_hidden_state fp32 [B, T, width];
        with output_hidden_states, [-1] is the tuple of hidden states (embeddings + every layer, before the final LayerNorm)."""
        ids = input_ids.to(self.device)
        B, T = ids.shape
        C, H = self.width, self.heads
        d = C // H
        Tp = (T + 7) // 8 * 8                                                  # key dimension padded to a multiple of 8 (TMA rows)
        x = (torch.nn.functional.embedding(ids, self.tok) + self.pos[:T]).reshape(B * T, C).contiguous()   # fp32 token stream (a table lookup)
        hidden = [x.view(B, T, C)] if output_hidden_states else None
        scores = torch.empty((H, T, Tp), dtype=bf16, device=self.device)
        vt = torch.zeros((C, Tp), dtype=bf16, device=self.device)              # V^T of one prompt; the padding columns stay zero
        attn = torch.empty((B * T, C), dtype=bf16, device=self.device)
        for ly in self.layers:
            h = ops.layernorm(x, *ly.ln1, eps=self.eps)
            qk = ops.gemm(h, ly.w_qk, bias=ly.b_qk)                            # [B*T, 2C]
            for b in range(B):
                rows = slice(b * T, (b + 1) * T)
                q3 = qk[rows, :C].view(T, H, d).permute(1, 0, 2)               # [H, T, d] views: heads are the GEMM batch
                k3 = qk[rows, C:].view(T, H, d).permute(1, 0, 2)
                ops.gemm(q3, k3, out=scores)                                   # S = Q K^T per head -> [H, T, Tp] (columns >= T unwritten)
                ops.softmax_rows(scores, d ** -0.5, out=scores, n_valid=T, causal_period=T)
                ops.gemm(ly.w_v, h[rows], out=vt)                              # V^T[c, token] = Wv[c, :] . h[token, :]   (bias folded into b_o)
                ops.gemm(scores, vt.view(H, d, Tp), out=attn[rows].view(T, H, d).permute(1, 0, 2))
            x = ops.gemm(attn, ly.w_o, bias=ly.b_o, residual=x, out_f32=True)
            h = ops.layernorm(x, *ly.ln2, eps=self.eps)
            u = ops.silu(ops.gemm(h, ly.w_fc1, bias=ly.b_fc1))
            x = ops.gemm(u, ly.w_fc2, bias=ly.b_fc2, residual=x, out_f32=True)
            if output_hidden_states:
                hidden.append(x.view(B, T, C))
        last = ops.layernorm(x, *self.ln_f, eps=self.eps).view(B, T, C).float()
        if output_hidden_states:
            return (last, tuple(hidden))
        return (last,)
