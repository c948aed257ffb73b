"""ncu target: ONE eager dual-branch denoise step at the benchmark shape (SDR UNet on 16 samples + fused CFG/PLMS step + GM UNet
on 8 samples + fused step), i.e. the unit bench.py repeats 51x per batch.  A warm-up step runs first; the measured step is
bracketed by cudaProfilerStart/Stop (use `ncu --profile-from-start off`)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
from gm_diffusion_b200 import schedulers as S

dev = torch.device("cuda:0")
pipe = bench.build_pipeline(dev)
B, h, w = 8, 64, 64
g = torch.Generator(device=dev).manual_seed(5)
ctx2 = torch.randn(2 * B, 77, 768, device=dev, generator=g)
kv_s, kv_g = pipe.unet.project_context(ctx2), pipe.gm_unet.project_context(ctx2[B:])
sched = S.PNDMScheduler(); sched.set_timesteps(50)
gs = S.clone_scheduler(sched)
tb_s, tb_g = pipe.unet.timestep_table([981]), pipe.gm_unet.timestep_table([981])
n_px = B * h * w
sdr, gm = S.BranchState(n_px, dev), S.BranchState(n_px, dev)
sdr.x.normal_(generator=g); gm.x.copy_(sdr.x)
unet_in = torch.zeros(B, h, w, 8, dtype=torch.bfloat16, device=dev)
gm_in = torch.zeros(B, h, w, 8, dtype=torch.bfloat16, device=dev)
eps_s = torch.empty(2 * B, h, w, 4, device=dev); eps_g = torch.empty(B, h, w, 4, device=dev)

def step(t):
    pipe.unet.forward(unet_in, tb_s, kv_s, out=eps_s, cfg_shared=True)
    S.fused_step(sched.plan_step(t), sdr, eps_s[B:].reshape(-1, 4), eps_s[:B].reshape(-1, 4), guidance_scale=7.5, px_per_sample=h * w,
                 x0_coeffs=sched.x0_coeffs(t), unet_in_next=unet_in, unet_in_dup=1, concat_out=gm_in, concat_tail=gm.x)
    pipe.gm_unet.forward(gm_in, tb_g, kv_g, out=eps_g)
    S.fused_step(gs.plan_step(t), gm, eps_g.reshape(-1, 4), x0_coeffs=gs.x0_coeffs(t))

step(981); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(961); e1.record(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("eager denoise step ms", e0.elapsed_time(e1))
