"""ncu target: the small-K token GEMMs (M=65536, K=320): plain bf16-out+bias, fp32-out + fp32 residual, GEGLU."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from gm_diffusion_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn(65536, 320, device="cuda", generator=g).to(torch.bfloat16)
w = (torch.randn(320, 320, device="cuda", generator=g) / 18).to(torch.bfloat16)
b = torch.randn(320, device="cuda", generator=g)
res = torch.randn(65536, 320, device="cuda", generator=g)
w2 = (torch.randn(2560, 320, device="cuda", generator=g) / 18).to(torch.bfloat16)
b2 = torch.randn(2560, device="cuda", generator=g)
for _ in range(3):
    ops.gemm(a, w, bias=b)
    ops.gemm(a, w, bias=b, residual=res, out_f32=True)
    ops.gemm(a, w2, bias=b2, geglu=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in [("plain", lambda: ops.gemm(a, w, bias=b)), ("f32res", lambda: ops.gemm(a, w, bias=b, residual=res, out_f32=True)),
                 ("geglu", lambda: ops.gemm(a, w2, bias=b2, geglu=True))]:
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, e0.elapsed_time(e1) / 20 * 1000, "us")
