"""ncu target for the GroupNorm kernels: one call per representative shape of the 512x512 UNet (after a warm-up call each)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from gm_diffusion_b200 import ops

g = torch.Generator(device="cuda").manual_seed(0)
ws = ops.gn_workspace(64, 32, "cuda")
shapes = [((16, 64, 64, 320), torch.bfloat16), ((16, 64, 64, 320), torch.float32), ((16, 32, 32, 640), torch.bfloat16),
          ((16, 16, 16, 1280), torch.bfloat16), ((16, 8, 8, 1280), torch.float32)]
if len(sys.argv) > 1:
    shapes = [shapes[int(i)] for i in sys.argv[1].split(",")]
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 10
xs = []
for shp, dt in shapes:
    x = torch.randn(shp, device="cuda", generator=g).to(dt)
    gam, bet = torch.randn(shp[-1], device="cuda", generator=g), torch.randn(shp[-1], device="cuda", generator=g)
    xs.append((x, gam, bet))
    ops.groupnorm_silu(x, gam, bet, stats_ws=ws)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for x, gam, bet in xs:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        ops.groupnorm_silu(x, gam, bet, stats_ws=ws)
    e1.record(); torch.cuda.synchronize()
    print(tuple(x.shape), x.dtype, f"{e0.elapsed_time(e1) * 1000 / REPS:.1f} us per call ({REPS} back-to-back)")
torch.cuda.cudart().cudaProfilerStop()
