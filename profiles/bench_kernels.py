"""Per-kernel roofline measurements for the HBM-bound kernels (c), (d) and the secondary BASELINE.json configs.
CUDA events on the launching stream, >=3 warm-ups; inputs far larger than the 126 MB L2 (or stated otherwise).
Run on a B200:  python profiles/bench_kernels.py > gpurun_out/kernels_r01.json"""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import gm_diffusion_b200 as G
from gm_diffusion_b200 import schedulers as S
from gm_diffusion_b200.stage1 import tone_mapping as TM

PEAK = 6555.8
try:
    PEAK = float(json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass
dev = torch.device("cuda:0")
out = {"hbm_peak_gbs": PEAK, "kernel_d": [], "kernel_c": [], "configs": {}}

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

# ---- kernel (d): BASELINE.json config 4: 4K frames, batch 32 (sweep 1..32), qmax 99 ----
g = torch.Generator(device=dev).manual_seed(0)
for B in (1, 4, 16, 32):
    sdr = torch.rand(B, 3, 2160, 3840, device=dev, generator=g)
    gm = torch.rand(B, 3, 2160, 3840, device=dev, generator=g)
    px = B * 2160 * 3840
    variants = [
        ("eq1 only (36 B/px)", 36, lambda: TM.reconstruct_hdr(sdr, gm, 99.0)),
        ("eq1 -> fix_mulog -> gamut, one output (36 B/px)", 36, lambda: TM.reconstruct_hdr(sdr, gm, 99.0, tmo="fix_mulog", gamut=True, return_hdr=False)),
        ("eq1 + fix_mulog + gamut, both outputs + min/max (48 B/px)", 48, lambda: TM.reconstruct_hdr(sdr, gm, 99.0, tmo="fix_mulog", gamut=True, return_minmax=False)),
        ("fix_mulog standalone (24 B/px)", 24, lambda: TM.fix_mulog_tmo(sdr, 99.0)),
    ]
    for name, bpp, fn in variants:
        ms = timeit(fn, n=5 if B >= 16 else 10)
        gbs = px * bpp / ms / 1e6
        out["kernel_d"].append({"batch": B, "variant": name, "ms": round(ms, 4), "GBps": round(gbs, 1), "frac_of_hbm_peak": round(gbs / PEAK, 3),
                                "algorithmic_bytes": px * bpp})
    del sdr, gm
torch.cuda.empty_cache()

# ---- kernel (c): SDR half (32 E bytes) + GM half (28 E bytes) per image per step, E = 16384 (512x512) ----
for B in (8, 64, 512, 4096):
    h = w = 64
    n_px = B * h * w
    sched = S.PNDMScheduler(); sched.set_timesteps(50)
    gs = S.clone_scheduler(sched)
    sdr, gmb = S.BranchState(n_px, dev), S.BranchState(n_px, dev)
    sdr.x.normal_(); gmb.x.normal_()
    eps = torch.randn(2 * n_px, 4, device=dev); epg = torch.randn(n_px, 4, device=dev)
    uin = torch.zeros(n_px, 8, dtype=torch.bfloat16, device=dev); gin = torch.zeros(n_px, 8, dtype=torch.bfloat16, device=dev)
    ts = sched.timesteps.tolist()
    for t in ts[:6]:  # reach the steady 4-term PLMS state
        S.fused_step(sched.plan_step(t), sdr, eps[n_px:], eps[:n_px], guidance_scale=7.5, px_per_sample=h * w, x0_coeffs=sched.x0_coeffs(t),
                     unet_in_next=uin, concat_out=gin, concat_tail=gmb.x)
        S.fused_step(gs.plan_step(t), gmb, epg, x0_coeffs=gs.x0_coeffs(t))
    plan, gplan = sched.plan_step(ts[6]), gs.plan_step(ts[6])
    def both():
        S.fused_step(plan, sdr, eps[n_px:], eps[:n_px], guidance_scale=7.5, px_per_sample=h * w, x0_coeffs=(0.9, 0.4), unet_in_next=uin, concat_out=gin, concat_tail=gmb.x)
        S.fused_step(gplan, gmb, epg, x0_coeffs=(0.9, 0.4))
    ms = timeit(both, n=20)
    # actual bytes of this implementation: fp32 eps (the UNet's conv_out writes fp32): SDR: 2 eps + x + 3 hist + gm_x reads, x + eps_out writes (16 B each) + 2 x 16 B bf16 rows
    actual = n_px * ((2 + 1 + 3 + 1) * 16 + 2 * 16 + 2 * 16) + n_px * ((1 + 1 + 3) * 16 + 2 * 16)
    alg = B * 60 * 16384
    out["kernel_c"].append({"batch": B, "ms_sdr_plus_gm_step": round(ms, 5), "algorithmic_bytes": alg, "GBps_algorithmic": round(alg / ms / 1e6, 1),
                            "frac_of_hbm_peak": round(alg / ms / 1e6 / PEAK, 4), "actual_bytes": actual, "GBps_actual": round(actual / ms / 1e6, 1),
                            "note": "two launches; launch-latency-bound at batch 8 (the benchmark's batch); includes ~2x5 us of host launch overhead when not graph-captured"})

# ---- BASELINE.json config 3: single (SDR -> GM) pipeline at 1024x1024 (latent 128x128, 16384 tokens at level 0), batch 1 ----
from gm_diffusion_b200 import random_init as R
sd8 = R.widen_conv_in_state_dict(R.sd15_unet_state_dict(4, seed=0, device=dev))
pipe = G.StableDiffusionGMPipeline(vae=None, text_encoder=None, tokenizer=None, unet=G.B200UNet(sd8, device=dev), scheduler=G.PNDMScheduler(), device=dev)
del sd8
pe = torch.randn(1, 77, 768, device=dev); ne = torch.randn(1, 77, 768, device=dev)
sl = 0.18215 * torch.randn(1, 4, 128, 128, device=dev)
run = lambda: pipe(sl, prompt_embeds=pe, negative_prompt_embeds=ne, num_inference_steps=10, guidance_scale=7.5, output_type="latent")
o = run(); assert torch.isfinite(o.images).all()
ms = timeit(run, n=2, warm=1)
evals = 11
out["configs"]["config3_single_1024"] = {"resolution": "1024x1024", "batch": 1, "steps": "10 PNDM (11 evals)", "ms_per_unet_eval_pair": round(ms / evals, 2),
                                         "algorithmic_tflop_per_eval": 2 * 4.6744, "achieved_tflops": round(2 * 4674.4 / (ms / evals), 1),
                                         "extrapolated_s_per_image_50_steps": round(ms / evals * 51 / 1000, 2)}
print(json.dumps(out))
