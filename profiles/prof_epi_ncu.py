"""ncu target: epilogue-dominated GEMM (M=65536, N=320, K=64, bias, bf16 out)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from gm_diffusion_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
M, N, K = 65536, 320, 64
a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
w = ops.tile_weight((torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16))
b = torch.randn(N, device="cuda", generator=g)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    ops.gemm(a, w, bias=b, out=out)
torch.cuda.synchronize()
print("ok")
