import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from gm_diffusion_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
w = ops.tile_weight((torch.randn(1280, 11520, device="cuda", generator=g) / 107).to(torch.bfloat16))
b = torch.randn(1280, device="cuda", generator=g)
a1 = torch.randn(512, 11520, device="cuda", generator=g).to(torch.bfloat16)
for _ in range(3):
    ops.gemm(a1, w, bias=b); ops.gemm(a1, w, bias=b, splitk=True)
torch.cuda.synchronize(); print("ok")
