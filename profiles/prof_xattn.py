"""ncu target: level-0 text cross-attention (Nq=4096, Nk=77, 8 heads, d=40) on 16 samples, a few launches."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from gm_diffusion_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.randn(16, 4096, 320, device="cuda", generator=g).to(torch.bfloat16)
kv = torch.randn(16, 77, 640, device="cuda", generator=g).to(torch.bfloat16)
for _ in range(4):
    ops.attention(q, kv[..., :320], kv[..., 320:], 8)
torch.cuda.synchronize()
print("ok")
