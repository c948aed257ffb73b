"""Self-/cross-attention shapes of one batch-8 denoise step (CFG: 16 samples in the SDR UNet, 8 in the GM UNet), timed alone
with CUDA events; TFLOP/s with the true head dim and key count (4*B*H*Nq*Nk*d)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import os
from gm_diffusion_b200 import _lib, ops
if os.environ.get('GMD_AB_LIB'):   # A/B runs against an alternative build of the library
    _lib.LIB_PATH = Path(os.environ['GMD_AB_LIB']).resolve()

g = torch.Generator(device="cuda").manual_seed(0)
shapes = [(16, 4096, 4096, 320), (8, 4096, 4096, 320), (16, 1024, 1024, 640), (8, 1024, 1024, 640), (16, 256, 256, 1280),
          (16, 4096, 77, 320), (16, 1024, 77, 640), (1, 16384, 16384, 320)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B, Nq, Nk, C in shapes:
    q = torch.randn(B, Nq, C, device="cuda", generator=g).to(torch.bfloat16)
    k = torch.randn(B, Nk, C, device="cuda", generator=g).to(torch.bfloat16)
    v = torch.randn(B, Nk, C, device="cuda", generator=g).to(torch.bfloat16)
    for _ in range(3):
        ops.attention(q, k, v, 8)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.attention(q, k, v, 8); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    print(f"attn B={B} Nq={Nq} Nk={Nk} C={C} d={C // 8}: {ms * 1e3:8.1f} us  {4 * B * Nq * Nk * C / ms / 1e9:7.1f} TFLOP/s", flush=True)
