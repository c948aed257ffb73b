"""LayerNorm of the fp32 token stream at the UNet's shapes, timed alone (L2 flushed between launches)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from gm_diffusion_b200 import _lib, ops
if os.environ.get('GMD_AB_LIB'):
    _lib.LIB_PATH = Path(os.environ['GMD_AB_LIB']).resolve()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
g = torch.Generator(device="cuda").manual_seed(0)
for M, C in [(65536, 320), (32768, 320), (16384, 640), (8192, 640), (4096, 1280), (2048, 1280)]:
    x = torch.randn(M, C, device="cuda", generator=g)
    ga, be = torch.randn(C, device="cuda", generator=g), torch.randn(C, device="cuda", generator=g)
    out = torch.empty(M, C, dtype=torch.bfloat16, device="cuda")
    for _ in range(3): ops.layernorm(x, ga, be, out=out)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.layernorm(x, ga, be, out=out); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); ms = ts[5]
    print(f"layernorm fp32 ({M}, {C}): {ms * 1e3:7.1f} us  {M * C * 6 / ms / 1e6:7.1f} GB/s (read fp32 + write bf16)", flush=True)
