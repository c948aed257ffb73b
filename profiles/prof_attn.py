"""ncu target: level-0 self-attention (N=4096, 8 heads, d=40) on 4 samples, a few launches."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from gm_diffusion_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(4, 4096, 960, device="cuda", generator=g).to(torch.bfloat16)
for _ in range(4):
    ops.attention(qkv[..., :320], qkv[..., 320:640], qkv[..., 640:], 8)
torch.cuda.synchronize()
print("ok")
