#!/usr/bin/env python
"""Benchmark of the Stage-3 hot path: HDR img/s (512x512, 50-step dual-branch) — BASELINE.json's metric.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

One "step" = one batch of 8 images per GPU through the WHOLE path: prompt embeddings -> 51-eval PNDM dual-branch loop
(SDR UNet under CFG 7.5 + GM UNet, bf16 tensor-core kernels, fp32 latents) -> 2x VAE decode -> Eq.(1) qmax=99 -> HDR fp32
[B,512,512,3].  Weak scaling: every rank runs its own batch of 8 (global batch 8N), no collective inside the loop, one
NCCL all-gather of the HDR outputs per step.  Synthetic data: random-init SD1.5-architecture weights, N(0,1) embeddings.

`value`  : images/s with inputs resident in HBM, device-timed (CUDA events), max over ranks.
`e2e`    : same metric through the public pipeline call with HOST (pinned) inputs and a device->host read of the HDR batch.
`roofline`: tensor-core kernel family (tcgen05 GEMM + implicit-GEMM conv) measured live with CUDA events on an instrumented
            eager denoise step after the timed region; algorithmic FLOPs from the layer shapes; peak from MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference`: the oracle port of the reference loop on the host cores (diffusers is not installable
            offline, so the reference pipelines cannot run), bounded sample, extrapolated to img/s.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "HDR img/s (512x512, 50-step dual-branch)"
UNIT = "img/s"
BATCH_PER_GPU = 8
HEIGHT = WIDTH = 512
STEPS_INFER = 50
GUIDANCE = 7.5
QMAX = 99.0
# algorithmic work: 2*MACs of convs + linears + attention with true head dims / true Nk
GFLOP_UNET4, GFLOP_UNET8, GFLOP_VAE = 803.3, 803.4, 2514.5
EVALS = 51


def workload_config(n_gpus):
    return {"workload": "text-to-HDR dual-branch SD1.5-arch (BASELINE.json configs[1])", "resolution": "512x512",
            "scheduler": "PNDM 50 steps (51 UNet evals per branch)", "guidance_scale": GUIDANCE, "batch_per_gpu": BATCH_PER_GPU,
            "global_batch": BATCH_PER_GPU * n_gpus, "parallelism": f"dp{n_gpus} (image sharding, all-gather of HDR outputs)",
            "includes": "prompt-embeds -> dual loop -> 2x VAE decode -> Eq.(1) qmax=99", "l2": "per-step working set (>1.7 GB weights per UNet + activations) exceeds the 126 MB L2"}


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d.get("bf16_tflops_sustained", 1376.7)), float(d.get("hbm_gbs", 6555.8)), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/), or None."""
    p = ROOT / "profiles" / "roofline_traffic.json"
    try:
        return json.loads(p.read_text())
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU reference arm: the oracle port of the reference loop on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(steps: int, warmup: int):
    """Bounded sample of the SAME workload on the host: batch 1, fp32, 512x512 -> one dual-branch denoise step (SDR UNet on
    2 samples under CFG + GM UNet) per timed step; one VAE decode + Eq.(1) timed once; extrapolated to 51 evals + 2 decodes."""
    from oracle import pipeline_oracle  # noqa: F401  (the CPU port; bench's reference leg is allowed to use it)
    from oracle import tone_mapping_oracle as TMO
    from oracle.schedulers_oracle import PNDMOracle
    from oracle.unet_oracle import UNet2DConditionOracle, widen_conv_in
    from oracle.vae_oracle import VaeDecoderOracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    u4 = UNet2DConditionOracle(4).eval()
    u8 = widen_conv_in(u4).eval()
    vae = VaeDecoderOracle().eval()
    g = torch.Generator().manual_seed(1)
    pe, ne = torch.randn(1, 77, 768, generator=g), torch.randn(1, 77, 768, generator=g)
    lat = torch.randn(1, 4, HEIGHT // 8, WIDTH // 8, generator=g)
    sched = PNDMOracle(); sched.set_timesteps(STEPS_INFER)
    gsched = PNDMOracle(); gsched.set_timesteps(STEPS_INFER)
    emb = torch.cat([ne, pe])

    def denoise_step(x, gx, t):
        with torch.no_grad():
            e = u4(torch.cat([x, x]), t, encoder_hidden_states=emb)
            u, c = e.chunk(2)
            e = u + GUIDANCE * (c - u)
            a = sched.alphas_cumprod[t]
            x0 = (x - (1 - a).sqrt() * e) / a.sqrt()
            xn = sched.step(e, t, x)[0]
            ge = u8(torch.cat([x0, gx], 1), t, encoder_hidden_states=pe)
            return xn, gsched.step(ge, t, gx)[0]

    x, gx = lat, lat.clone()
    ts = sched.timesteps.tolist()
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        x, gx = denoise_step(x, gx, ts[i % len(ts)])
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    t_step = sum(times) / len(times)
    t0 = time.perf_counter()
    with torch.no_grad():
        img = TMO.denormalize(vae.decode(x / 0.18215))
        hdr = TMO.apply_gm_to_sdr_numpy(img.numpy(), img.numpy(), QMAX)  # the numpy twin the scripts run on the host
    t_tail = 2 * (time.perf_counter() - t0)  # two decodes (SDR + GM) per image
    assert hdr.shape[1] == 3
    s_per_img = EVALS * t_step + t_tail
    model = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip(); break
    except OSError:
        pass
    return {"value": 1.0 / s_per_img, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"batch 1 fp32 512x512: {steps} timed dual-branch denoise steps of 51 ({t_step:.2f} s each, 3 UNet forward-equivalents = 2.41 TFLOP) "
                      f"+ 1 of 2 VAE decodes with Eq.(1) ({t_tail / 2:.2f} s); extrapolated to 51 evals + 2 decodes = {s_per_img:.1f} s/image",
            "cpu_model": model, "torch_threads": torch.get_num_threads(), "s_per_step": t_step,
            "cpu_gflops": (2 * GFLOP_UNET4 + GFLOP_UNET8) / t_step}, t_step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    base, t_step = cpu_reference_sample(max(1, args.steps), max(0, args.warmup))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 / base["value"] * BATCH_PER_GPU, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic (random-init SD1.5-arch weights, N(0,1) embeddings)",
            "config": workload_config(args.gpus), "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference pipelines need diffusers (not installable offline); this arm times the oracle port of the same loop on the host cores"}
    emit(line)
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------------
def build_pipeline(device):
    import gm_diffusion_b200 as G
    from gm_diffusion_b200 import random_init as R
    sd4 = R.sd15_unet_state_dict(4, seed=0, device=device)
    unet = G.B200UNet(sd4, device=device)
    gm_unet = G.B200UNet(R.widen_conv_in_state_dict(sd4), device=device)
    del sd4
    vae = G.B200VaeDecoder(R.sd_vae_decoder_state_dict(seed=4, device=device), device=device)
    torch.cuda.empty_cache()
    return G.StableDiffusionDualUNetPipeline(vae=vae, text_encoder=None, tokenizer=None, unet=unet, gm_unet=gm_unet,
                                             scheduler=G.PNDMScheduler(), device=device)


def instrumented_step(pipe, B):
    """One eager denoise step (SDR UNet 2B + GM UNet B) with CUDA events around every C-ABI call: per-family device time and
    the tensor-core family's achieved FLOP/s (algorithmic FLOPs from the layer shapes)."""
    from gm_diffusion_b200 import ops
    dev = pipe.device
    fams = {"gemm": "gemm", "conv2d": "gemm", "attention": "attn", "groupnorm_silu": "norm", "layernorm": "norm"}
    orig = {n: getattr(ops, n) for n in fams}
    records = []

    def flops_of(name, a, k, out):
        if name == "gemm":
            x, w = a[0], a[1]
            return 2.0 * x.shape[-2] * w.shape[-2] * x.shape[-1] * (x.shape[0] if x.dim() == 3 else 1)
        if name == "conv2d":
            x, w = a[0], a[1]
            return 2.0 * (out.numel() // out.shape[-1]) * a[2] * w.shape[1]  # output pixels x Cout x (taps * Cin)
        if name == "attention":
            q, kk, heads = a[0], a[1], a[3]
            return 4.0 * q.shape[0] * q.shape[1] * kk.shape[1] * q.shape[2]
        return 0.0

    def wrap(name):
        def f(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = orig[name](*a, **k)
            e1.record()
            records.append((fams[name], flops_of(name, a, k, out), e0, e1))
            return out
        return f

    h = w = HEIGHT // 8
    try:
        for n in fams:
            setattr(ops, n, wrap(n))
        g = torch.Generator(device=dev).manual_seed(5)
        ctx2 = torch.randn(2 * B, 77, 768, device=dev, generator=g)
        kv_s, kv_g = pipe.unet.project_context(ctx2), pipe.gm_unet.project_context(ctx2[:B])
        tb_s, tb_g = pipe.unet.timestep_table([501]), pipe.gm_unet.timestep_table([501])
        xs = torch.randn(B, h, w, 8, device=dev, generator=g).to(torch.bfloat16)  # one copy: the CFG halves share the prefix
        xs[..., 4:] = 0
        xg = torch.randn(B, h, w, 8, device=dev, generator=g).to(torch.bfloat16)
        records.clear()
        torch.cuda.synchronize()
        pipe.unet.forward(xs, tb_s, kv_s, cfg_shared=True)
        pipe.gm_unet.forward(xg, tb_g, kv_g)
        torch.cuda.synchronize()
    finally:
        for n in fams:
            setattr(ops, n, orig[n])
    agg = {}
    for fam, fl, e0, e1 in records:
        a = agg.setdefault(fam, {"ms": 0.0, "flop": 0.0, "launches": 0})
        a["ms"] += e0.elapsed_time(e1); a["flop"] += fl; a["launches"] += 1
    return agg


def run_b200(args):
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    from gm_diffusion_b200 import _lib as L
    from gm_diffusion_b200 import dist as D
    rank, local, world = D.init_from_env()
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B = BATCH_PER_GPU
    total = B * world
    pipe = build_pipeline(dev)
    g = torch.Generator().manual_seed(1000 + rank)
    pe_h = torch.randn(B, 77, 768, generator=g).pin_memory()
    ne_h = torch.randn(B, 77, 768, generator=g).pin_memory()
    lat_h = torch.randn(B, 4, HEIGHT // 8, WIDTH // 8, generator=g).pin_memory()
    hdr_h = torch.empty(B, HEIGHT, WIDTH, 3, dtype=torch.float32).pin_memory()
    pe_d, ne_d, lat_d = pe_h.to(dev), ne_h.to(dev), lat_h.to(dev)
    kw = dict(height=HEIGHT, width=WIDTH, num_inference_steps=STEPS_INFER, guidance_scale=GUIDANCE, output_type="hdr", qmax=QMAX)

    def step_resident():
        hdr, _, _ = pipe(prompt_embeds=pe_d, negative_prompt_embeds=ne_d, latents=lat_d, **kw)
        return D.gather_outputs(hdr, total) if world > 1 else hdr

    def step_e2e():
        hdr, _, _ = pipe(prompt_embeds=pe_h.to(dev, non_blocking=True), negative_prompt_embeds=ne_h.to(dev, non_blocking=True),
                         latents=lat_h.to(dev, non_blocking=True), **kw)
        hdr_h.copy_(hdr, non_blocking=True)
        return hdr

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        return D.max_over_ranks(e0.elapsed_time(e1) / 1000.0, dev), out

    for _ in range(max(3, args.warmup)):
        out = step_resident()
    assert torch.isfinite(out).all(), "non-finite HDR output"
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    L.reset_launch_count(); pipe.graph_launches = 0
    t_res, out = timed(step_resident, args.steps)
    launches = L.launch_count() + pipe.graph_launches
    step_e2e(); torch.cuda.synchronize()
    t_e2e, _ = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    kw_lat = dict(kw, output_type="latent")
    t_lat, _ = timed(lambda: pipe(prompt_embeds=pe_d, negative_prompt_embeds=ne_d, latents=lat_d, **kw_lat)[0], 1)
    value = total * args.steps / t_res
    e2e_value = total * args.steps / t_e2e
    ms_batch = 1000.0 * t_res / args.steps
    if rank != 0:
        if world > 1:
            torch.distributed.barrier()
        return 0

    # ---- roofline of the dominant kernel family, measured live (instrumented eager step, outside the timed region) ----
    peak_tf, peak_hbm, peak_src = load_peaks()
    fam = instrumented_step(pipe, B)
    tot_ms = sum(v["ms"] for v in fam.values())
    gm_ = fam.get("gemm", {"ms": 1e-9, "flop": 0.0, "launches": 1})
    ach_tf = gm_["flop"] / (gm_["ms"] * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": "gmd::gemm_kernel<BN,STAGES> (tcgen05 GEMM + implicit-GEMM conv; all instantiations)",
                "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf, "peak_source": peak_src,
                "traffic": ncu_traffic(), "launches_per_denoise_step": gm_["launches"], "avg_launch_ms": gm_["ms"] / gm_["launches"],
                "algorithmic_gflop_per_denoise_step": gm_["flop"] / 1e9, "share_of_step": gm_["ms"] / tot_ms}
    at = fam.get("attn", {"ms": 1e-9, "flop": 0.0, "launches": 1})
    kernels = {k: {"ms_per_denoise_step": round(v["ms"], 3), "share": round(v["ms"] / tot_ms, 4), "launches": v["launches"],
                   "tflops": round(v["flop"] / (v["ms"] * 1e-3) / 1e12, 1) if v["flop"] else None} for k, v in fam.items()}
    unet_tflop_per_img = EVALS * (2 * GFLOP_UNET4 + GFLOP_UNET8) / 1000.0
    total_tflop_per_img = unet_tflop_per_img + 2 * GFLOP_VAE / 1000.0
    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu_base, _ = cpu_reference_sample(1, 1)
        except Exception as ex:  # the baseline is reported, never load-bearing
            cpu_base = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_batch, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic (random-init SD1.5-arch UNets + VAE decoder, N(0,1) prompt embeddings and latents)",
            "config": workload_config(world),
            "ms_per_denoise_step": 1000.0 * t_lat / EVALS,
            "ms_tail_vae_x2_plus_eq1": ms_batch - 1000.0 * t_lat,
            "whole_path_tflops_per_gpu": total_tflop_per_img * value / world,
            "frac_of_tensor_roofline_whole_path": total_tflop_per_img * value / world / peak_tf,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": (pe_h.numel() + ne_h.numel() + lat_h.numel()) * 4,
                    "d2h_bytes_per_step": hdr_h.numel() * 4, "api": "StableDiffusionDualUNetPipeline.__call__(output_type='hdr') with pinned host tensors"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernels": kernels,
            "attention": {"achieved_tflops": at["flop"] / (at["ms"] * 1e-3) / 1e12, "note": "true head dims 40/80/160, true Nk; padding not credited"},
            "cpu_baseline": cpu_base}
    emit(line)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return 0


_JSON_OUT = None


def emit(line: dict) -> None:
    """The ONE stdout line of the contract.  File descriptor 1 is pointed at stderr for the whole run (see main), so nothing a
    library prints on stdout — NCCL's version banner at process-group creation — can land next to it."""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _JSON_OUT
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--comparators", action="store_true",
                    help="GPU-side comparators (SURVEY 8d): torch SDPA / flash_attn / cuDNN / un-fused torch chains / torch-eager UNet "
                         "on the same shapes, next to the hand-written kernels; prints one JSON object")
    args = ap.parse_args()
    if args.comparators:
        sys.path.insert(0, str(Path(__file__).resolve().parent / "profiles"))
        import bench_comparators
        from oracle import tone_mapping_oracle, unet_oracle   # torch restatements, timed on the GPU as the reference's stack
        _JSON_OUT.write(json.dumps(bench_comparators.run(tone_mapping_oracle, unet_oracle, build_pipeline), indent=1) + "\n")
        _JSON_OUT.flush()
        return 0
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
