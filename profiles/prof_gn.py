"""ncu target: GroupNorm(+SiLU) on the level-0 / level-1 UNet activations, a few launches each (GMD_GN_MODE selects the path)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from gm_diffusion_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
ws = ops.gn_workspace(64, 32, "cuda")
for shp, dt in [((16, 64, 64, 320), torch.bfloat16), ((16, 64, 64, 320), torch.float32), ((16, 32, 32, 640), torch.bfloat16), ((16, 16, 16, 1280), torch.bfloat16)]:
    xs = [torch.randn(shp, device="cuda", generator=g).to(dt) for _ in range(4)]
    out = torch.empty(shp, device="cuda", dtype=torch.bfloat16)
    gam, bet = torch.randn(shp[-1], device="cuda", generator=g), torch.randn(shp[-1], device="cuda", generator=g)
    for i in range(4):
        ops.groupnorm_silu(xs[i], gam, bet, stats_ws=ws, out=out)
torch.cuda.synchronize()
print("ok")
