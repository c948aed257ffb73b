"""GroupNorm(+SiLU) per-call time inside a CUDA graph (how the pipelines run it): 50 calls on rotating buffers captured once, replayed.
GMD_GN_TWO_PASS=1 selects the stats + apply pair for an A/B."""
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from gm_diffusion_b200 import ops

g = torch.Generator(device="cuda").manual_seed(0)
ws = ops.gn_workspace(64, 32, "cuda")
shapes = [((16, 64, 64, 320), torch.bfloat16), ((16, 64, 64, 320), torch.float32), ((8, 64, 64, 320), torch.bfloat16), ((16, 64, 64, 960), torch.bfloat16),
          ((16, 32, 32, 640), torch.bfloat16), ((16, 32, 32, 640), torch.float32), ((16, 16, 16, 1280), torch.bfloat16),
          ((16, 16, 16, 1280), torch.float32), ((16, 8, 8, 1280), torch.bfloat16), ((8, 8, 8, 1280), torch.float32), ((16, 8, 8, 2560), torch.bfloat16)]
SILU = os.environ.get("GN_SILU", "1") == "1"
print("two-pass" if os.environ.get("GMD_GN_TWO_PASS") == "1" else "one-pass cluster kernel", "silu" if SILU else "no silu")
for shp, dt in shapes:
    nbuf = 4
    xs = [torch.randn(shp, device="cuda", generator=g).to(dt) for _ in range(nbuf)]
    outs = [torch.empty(shp, device="cuda", dtype=torch.bfloat16) for _ in range(nbuf)]
    gam, bet = torch.randn(shp[-1], device="cuda", generator=g), torch.randn(shp[-1], device="cuda", generator=g)
    def run(n):
        for i in range(n):
            ops.groupnorm_silu(xs[i % nbuf], gam, bet, stats_ws=ws, out=outs[i % nbuf], silu=SILU)
    run(4); torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        run(48)
    graph.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        graph.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / (5 * 48)
    nbytes = xs[0].numel() * (xs[0].element_size() + 2)
    print(f"{str(shp):22s} {str(dt):15s} {us:7.2f} us/call   {nbytes / us / 1e6:6.2f} TB/s algorithmic (read once + write once)")
