"""Microbenchmarks for the small-M / long-K layers (8x8 resolution): conv vs the equivalent GEMM, plain vs tiled weights, with and
without the split-K workspace."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from gm_diffusion_b200 import ops, _lib as L
g = torch.Generator(device="cuda").manual_seed(0)
def bench(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
SPLIT = False
def nosplit():
    global SPLIT
    SPLIT = False
def split():
    global SPLIT
    SPLIT = True
w = (torch.randn(1280, 11520, device="cuda", generator=g) / 107).to(torch.bfloat16)
wt = ops.tile_weight(w)
b = torch.randn(1280, device="cuda", generator=g)
for B in (8, 16, 32):
    x = torch.randn(B, 8, 8, 1280, device="cuda", generator=g).to(torch.bfloat16)
    a = torch.randn(B * 64, 11520, device="cuda", generator=g).to(torch.bfloat16)
    res = {}
    for name, fn in [("conv plainW", lambda: ops.conv2d(x, w, 1280, bias=b, splitk=SPLIT)), ("conv tiledW", lambda: ops.conv2d(x, wt, 1280, bias=b, splitk=SPLIT)),
                     ("gemm plainW", lambda: ops.gemm(a, w, bias=b, splitk=SPLIT)), ("gemm tiledW", lambda: ops.gemm(a, wt, bias=b, splitk=SPLIT))]:
        nosplit(); t0 = bench(fn); split(); t1 = bench(fn)
        res[name] = (t0, t1)
    print(f"B={B} M={B*64}: " + " | ".join(f"{k}: {v[0]:.1f} / split {v[1]:.1f} us" for k, v in res.items()))
