"""compute-sanitizer target: every hand-written kernel family once on small shapes — the tcgen05 GEMM and implicit-GEMM
conv (mbarrier pipelines, TMA, TMEM), the flash-attention kernels (self d = 40 / 80, text cross, d = 160), GroupNorm, LayerNorm,
kernel (c) and kernel (d) — with the results checked against fp32 torch.
  compute-sanitizer --tool memcheck  python profiles/sanitize_target.py
  compute-sanitizer --tool racecheck python profiles/sanitize_target.py      (one tool per gpurun call, B200_PROFILING.md)"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import __graft_entry__ as E
from gm_diffusion_b200 import ops, schedulers as S

dev = torch.device("cuda:0")
E._first_launches(dev)
# a few more shapes: d = 160 attention, ragged key count, 2-source GroupNorm, stride-2 / upsample / 2-source conv, fused scheduler step
g = torch.Generator().manual_seed(1)
bf = torch.bfloat16
q = torch.randn(1, 256, 3 * 1280, generator=g).to(bf).to(dev)
ops.attention(q[..., :1280], q[..., 1280:2560], q[..., 2560:], 8)
q = torch.randn(1, 300, 3 * 640, generator=g).to(bf).to(dev)
ops.attention(q[:, :, :640], q[:, :200, 640:1280], q[:, :200, 1280:], 8)
x = torch.randn(1, 16, 16, 320, generator=g).to(bf).to(dev)
x1 = torch.randn(1, 16, 16, 320, generator=g).to(bf).to(dev)
ops.groupnorm_silu(x, torch.ones(640, device=dev), torch.zeros(640, device=dev), x1=x1)
w = ops.pack_conv_weight_tiled((torch.randn(320, 320, 3, 3, generator=g) * 0.02).to(dev))
ops.conv2d(x, w, 320, stride=2)
ops.conv2d(x, w, 320, upsample=True)
w2 = ops.pack_conv_weight_tiled((torch.randn(320, 640, 3, 3, generator=g) * 0.02).to(dev))
ops.conv2d(x, w2, 320, x1=x1, residual=x)
sched = S.PNDMScheduler(); sched.set_timesteps(4)
st = S.BranchState(2 * 256, dev); st.x.normal_()
eps = torch.randn(2 * 2 * 256, 4, device=dev)
uin = torch.zeros(2 * 256, 8, dtype=bf, device=dev)
for t in sched.timesteps.tolist():
    S.fused_step(sched.plan_step(t), st, eps[512:], eps[:512], guidance_scale=7.5, px_per_sample=256, x0_coeffs=sched.x0_coeffs(t), unet_in_next=uin)
torch.cuda.synchronize()
print("sanitize target ok")
