"""ncu target: the two dominant tensor-core shapes (3x3 conv 64x64 320->320 at batch 16, GEGLU GEMM M=65536) a few times."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from gm_diffusion_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(16, 64, 64, 320, device="cuda", generator=g).to(torch.bfloat16)
w = ops.pack_conv_weight_tiled((torch.randn(320, 320, 3, 3, device="cuda", generator=g) / 54).to(torch.bfloat16))   # production layout -> halo main loop
b = torch.randn(320, device="cuda", generator=g)
a = torch.randn(65536, 320, device="cuda", generator=g).to(torch.bfloat16)
w2, b2 = ops.pack_geglu_weight_tiled((torch.randn(2560, 320, device="cuda", generator=g) / 18).to(torch.bfloat16), torch.randn(2560, device="cuda", generator=g))
for _ in range(3):
    ops.conv2d(x, w, 320, bias=b)
    ops.gemm(a, w2, bias=b2, geglu=True)
torch.cuda.synchronize()
print("ok")
