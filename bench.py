#!/usr/bin/env python
"""Benchmark of the Stage-3 hot path: HDR img/s (512x512, 50-step dual-branch) — BASELINE.json's metric.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

One "step" = one batch of 8 images per GPU through the WHOLE path: prompt embeddings -> 51-eval PNDM dual-branch loop
(SDR UNet under CFG 7.5 + GM UNet, bf16 tensor-core kernels, fp32 latents) -> 2x VAE decode -> Eq.(1) qmax=99 -> HDR fp32
[B,512,512,3].  The loop is one CUDA graph on two streams (GM branch of step i beside the SDR UNet of step i+1; GMD_TWO_STREAMS=0
keeps one stream).  Weak scaling: every rank runs its own batch of 8 (global batch 8N), no collective inside the loop, one
NCCL all-gather of the HDR outputs per step.  Synthetic data: random-init SD1.5-architecture weights, N(0,1) embeddings.

`value`  : images/s with inputs resident in HBM, device-timed (CUDA events), max over ranks.
`e2e`    : same metric through the public pipeline call with HOST (pinned) inputs and a device->host read of the HDR batch.
`roofline`: tensor-core kernel family (tcgen05 GEMM + implicit-GEMM conv) measured live after the timed region: the family's calls
            of one denoise step replayed from a CUDA graph of their own (CUDA events); algorithmic FLOPs from the layer shapes; peak
            from MEASURED_PEAKS.json.  `gpu_comparator`: torch-eager bf16 (cuDNN / cuBLAS / cuDNN SDPA) on the same box.
            `extra_configs`: BASELINE.json configs[3] (1024x1024 single pipeline) and configs[4] (4K x 32 Eq.(1) sweep).
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
            "includes": "prompt-embeds -> dual loop -> 2x VAE decode -> Eq.(1) qmax=99", "l2": "per-step working set (>1.7 GB weights per UNet + activations) exceeds the 126 MB L2",
            "execution": "whole denoising loop as one CUDA graph; GM branch of step i on a side stream beside the SDR UNet of step i+1"
                         + (" (GMD_TWO_STREAMS=0: one stream)" if os.environ.get("GMD_TWO_STREAMS", "1") == "0" else "")}


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d.get("bf16_tflops_sustained", 1376.7)), float(d.get("hbm_gbs", 6555.8)), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes of the dominant kernel family from the committed ncu launch list of one warm denoise step (profiles/), or None."""
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


def family_times(pipe, B):
    """Per-kernel-family device time of ONE denoise step (SDR UNet 2B + GM UNet B), measured the way the step really runs: from
    CUDA-graph replay.  One eager pass records every C-ABI call of the step; then, per family, a graph holding ONLY that family's
    calls (same tensors, same order) is captured and its replay is timed with CUDA events (3 warm-ups, median of 7).  The three
    family graphs add up to the step (the step is a serial chain of kernels), so the per-family numbers are no longer upper
    bounds inflated by eager launch gaps as in round 1.  Also returned: the single hottest GEMM / conv instantiation (its calls
    of the step in a graph of their own)."""
    from gm_diffusion_b200 import ops
    dev = pipe.device
    fams = {"gemm": "gemm", "conv2d": "gemm", "attention": "attn", "groupnorm_silu": "norm", "layernorm": "norm"}
    orig = {n: getattr(ops, n) for n in fams}
    calls = []   # (family, flops, signature, closure)

    def flops_of(name, a, out):
        if name == "gemm":
            x, w = a[0], a[1]
            return 2.0 * x.shape[-2] * w.shape[-2] * x.shape[-1] * (x.shape[0] if x.dim() == 3 else 1)
        if name == "conv2d":
            return 2.0 * (out.numel() // out.shape[-1]) * a[2] * a[1].shape[1]  # output pixels x Cout x (taps * Cin)
        if name == "attention":
            q, kk = a[0], a[1]
            return 4.0 * q.shape[0] * q.shape[1] * kk.shape[1] * q.shape[2]
        return 0.0

    def bytes_of(name, a, k, out):
        """operands + output of the call once (weights as stored: padded tiles)"""
        n = out.numel() * out.element_size()
        for t in list(a) + [k.get("x1"), k.get("residual")]:
            if isinstance(t, torch.Tensor):
                n += t.numel() * t.element_size()
            elif hasattr(t, "data") and isinstance(getattr(t, "data"), torch.Tensor):    # ops.TiledWeight
                n += t.data.numel() * t.data.element_size()
        return float(n)

    def sig_of(name, a, k, out):
        if name == "gemm":
            f = "".join(t for t, on in (("+geglu", k.get("geglu")), ("+res", k.get("residual") is not None), ("+f32out", out.dtype == torch.float32)) if on)
            return f"gemm M={a[0].shape[-2]} N={a[1].shape[-2]} K={a[0].shape[-1]}{f}"
        if name == "conv2d":
            x = a[0]
            f = "".join(t for t, on in (("+s2", k.get("stride", 1) == 2), ("+up", k.get("upsample")), ("+2src", k.get("x1") is not None),
                                        ("+res", k.get("residual") is not None), ("+f32out", out.dtype == torch.float32)) if on)
            return f"conv{k.get('ksize', 3)} {tuple(x.shape[:3])} Cin={x.shape[3] + (k['x1'].shape[3] if k.get('x1') is not None else 0)} Cout={a[2]}{f}"
        return name

    def wrap(name):
        def f(*a, **k):
            ret = orig[name](*a, **k)
            out = ret[0] if isinstance(ret, tuple) else ret            # (conv2d / gemm with gn_stats return (out, statistics))
            k2 = {kk: vv for kk, vv in k.items() if kk != "out"}     # the replayed call allocates its own output
            calls.append((fams[name], flops_of(name, a, out), sig_of(name, a, k, out), lambda a=a, k2=k2, name=name: orig[name](*a, **k2),
                          bytes_of(name, a, k, out)))
            return ret
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
        calls.clear()
        pipe.unet.forward(xs, tb_s, kv_s, cfg_shared=True)
        pipe.gm_unet.forward(xg, tb_g, kv_g)
        torch.cuda.synchronize()
    finally:
        for n in fams:
            setattr(ops, n, orig[n])

    def graph_ms(closures):
        with ops.gn_arena(dev, None):          # (as inside a UNet forward: one memset for all GroupNorm accumulators, not one per call)
            for c in closures:
                c()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            with ops.gn_arena(dev, None):
                for c in closures:
                    c()
        ts = []
        for i in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        del gr
        return statistics.median(ts)

    agg = {}
    for fam in ("gemm", "attn", "norm"):
        sel = [c for c in calls if c[0] == fam]
        agg[fam] = {"ms": graph_ms([c[3] for c in sel]), "flop": sum(c[1] for c in sel), "launches": len(sel), "bytes": sum(c[4] for c in sel)}
    # hottest GEMM / conv instantiation by algorithmic work share -> its own graph
    groups = {}
    for fam, fl, sg, cl, _ in calls:
        if fam == "gemm":
            gp = groups.setdefault(sg, {"flop": 0.0, "calls": []})
            gp["flop"] += fl; gp["calls"].append(cl)
    top = sorted(groups.items(), key=lambda kv: -kv[1]["flop"])[:6]
    hot = []
    for sg, gp in top:
        ms = graph_ms(gp["calls"])
        hot.append({"what": sg, "calls_per_denoise_step": len(gp["calls"]), "ms": round(ms, 4), "tflops": round(gp["flop"] / ms / 1e9, 1)})
    torch.cuda.empty_cache()
    return agg, hot


def sdr_forward_ms(pipe, B):
    """One SDR UNet forward of the CFG batch (2B samples) alone on the GPU, replayed from a CUDA graph (median of 7, L2 flushed)."""
    dev = pipe.device
    g = torch.Generator(device=dev).manual_seed(5)
    h, w = HEIGHT // 8, WIDTH // 8
    kv = pipe.unet.project_context(torch.randn(2 * B, 77, 768, device=dev, generator=g))
    tb = pipe.unet.timestep_table([501])
    x = torch.randn(B, h, w, 8, device=dev, generator=g).to(torch.bfloat16)
    eps = torch.empty(2 * B, h, w, 4, device=dev)
    fwd = lambda: pipe.unet.forward(x, tb, kv, out=eps, cfg_shared=True)
    fwd(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fwd()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for i in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def gpu_comparator(B):
    """The reference's own GPU stack on this box, outside every timed region: ONE SDR UNet forward of the CFG batch
    (2B samples, 64x64 latents) as torch eager bf16 channels_last on cuDNN / cuBLAS with SDPA attention — what diffusers runs —
    and cuDNN SDPA on the step's two big self-attention shapes.  The torch network is the oracle's restatement (random init)."""
    import torch.nn.functional as F
    from torch.nn.attention import SDPBackend, sdpa_kernel
    from oracle import unet_oracle as UO      # comparator only: never on the product path
    dev = torch.device("cuda", torch.cuda.current_device())
    bf = torch.bfloat16
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn, n=7, warm=3):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    out = {"what": "torch eager bf16 on cuDNN / cuBLAS / cuDNN-SDPA on the same box, L2 flushed between iterations, median of 7"}
    g = torch.Generator(device=dev).manual_seed(0)
    from gm_diffusion_b200 import ops
    att = []
    for Bq, N, C in ((2 * B, 4096, 320), (2 * B, 1024, 640)):
        q, k, v = (torch.randn(Bq, N, C, device=dev, generator=g).to(bf) for _ in range(3))
        q4, k4, v4 = (t.view(Bq, N, 8, C // 8).transpose(1, 2) for t in (q, k, v))
        def cudnn():
            with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
                return F.scaled_dot_product_attention(q4, k4, v4)
        row = {"B": Bq, "N": N, "d": C // 8, "ms_ours": round(timeit(lambda: ops.attention(q, k, v, 8)), 4)}
        try:
            row["ms_cudnn_sdpa"] = round(timeit(cudnn), 4)
        except Exception as e:
            row["ms_cudnn_sdpa"] = f"unavailable: {type(e).__name__}"
        att.append(row)
        del q, k, v
    out["self_attention"] = att

    def sdpa_forward(self, x, ctx=None):   # diffusers' AttnProcessor2_0
        ctx = x if ctx is None else ctx
        b, n, c = x.shape
        hd, d = self.heads, c // self.heads
        q = self.to_q(x).view(b, n, hd, d).transpose(1, 2)
        k = self.to_k(ctx).view(b, -1, hd, d).transpose(1, 2)
        v = self.to_v(ctx).view(b, -1, hd, d).transpose(1, 2)
        return self.to_out[0](F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, n, c))

    old_fwd, old_bm = UO.Attention.forward, torch.backends.cudnn.benchmark
    try:
        UO.Attention.forward = sdpa_forward
        torch.backends.cudnn.benchmark = True
        with torch.device(dev):
            ref = UO.UNet2DConditionOracle(4).eval().to(bf).to(memory_format=torch.channels_last)
        lat = torch.randn(2 * B, 4, HEIGHT // 8, WIDTH // 8, device=dev, generator=g).to(bf).contiguous(memory_format=torch.channels_last)
        ctx = torch.randn(2 * B, 77, 768, device=dev, generator=g).to(bf)
        with torch.no_grad():
            out["unet_forward_2B_samples_ms_torch_eager"] = round(timeit(lambda: ref(lat, 501, ctx), n=5), 3)
        del ref
    finally:
        UO.Attention.forward, torch.backends.cudnn.benchmark = old_fwd, old_bm
    torch.cuda.empty_cache()
    return out


def extra_configs(pipe):
    """BASELINE.json configs[3] and configs[4] as extra keys of the same line (not bench lines of their own)."""
    import gm_diffusion_b200 as G
    from gm_diffusion_b200.stage1 import tone_mapping as TM
    dev = pipe.device
    out = {}
    # configs[3]: SDR->HDRTV up-conversion, single pipeline at 1024x1024 (latent 128x128: 16384 tokens at level 0), batch 1, 50 steps, qmax 99
    single = G.StableDiffusionGMPipeline(vae=pipe.vae, text_encoder=None, tokenizer=None, unet=pipe.gm_unet, scheduler=G.PNDMScheduler(), device=dev)
    g = torch.Generator().manual_seed(3)
    pe, ne = torch.randn(1, 77, 768, generator=g).to(dev), torch.randn(1, 77, 768, generator=g).to(dev)
    sdr_lat = (0.18215 * torch.randn(1, 4, 128, 128, generator=g)).to(dev)
    call = lambda: single(sdr_lat, prompt_embeds=pe, negative_prompt_embeds=ne, num_inference_steps=STEPS_INFER, guidance_scale=GUIDANCE, output_type="latent")
    call(); call(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); call(); e1.record(); torch.cuda.synchronize()
    s = e0.elapsed_time(e1) / 1000.0
    out["config3_single_pipeline_1024x1024_batch1"] = {"s_per_image_50_steps": round(s, 4), "ms_per_cfg_unet_eval": round(1000 * s / EVALS, 3),
                                                        "tflops": round(EVALS * 2 * 4674.4 / s / 1000, 1)}
    del single
    torch.cuda.empty_cache()
    # configs[4]: Eq.(1) reconstruction (+ fix_mulog) standalone, 4K frames, batch 32, fp32 planar
    gg = torch.Generator(device=dev).manual_seed(0)
    sdr = torch.rand(32, 3, 2160, 3840, device=dev, generator=gg)
    gm = torch.rand(32, 3, 2160, 3840, device=dev, generator=gg)
    px = 32 * 2160 * 3840
    _, hbm, _ = load_peaks()
    rows = {}
    for name, bpp, fn in (("eq1 (36 B/px)", 36, lambda: TM.reconstruct_hdr(sdr, gm, QMAX)),
                          ("eq1 -> fix_mulog -> gamut, one output (36 B/px)", 36, lambda: TM.reconstruct_hdr(sdr, gm, QMAX, tmo="fix_mulog", gamut=True, return_hdr=False)),
                          ("fix_mulog standalone (24 B/px)", 24, lambda: TM.fix_mulog_tmo(sdr, QMAX))):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        rows[name] = {"ms": round(ms, 3), "GBps": round(px * bpp / ms / 1e6, 1), "frac_of_hbm_peak": round(px * bpp / ms / 1e6 / hbm, 3)}
    out["config4_eq1_4k_batch32"] = rows
    del sdr, gm
    torch.cuda.empty_cache()
    return out


def run_config2(pipe, dev, rank, world, kw):
    """BASELINE.json configs[2] as written: global batch 64 sharded over the N ranks (32 / 16 / 8 images per GPU), first as plain
    image sharding, then with the CFG halves split across GPU pairs (ranks 2k, 2k+1 work on the same 64 / (N/2) images; SURVEY.md
    §8e).  One warm-up batch (graph capture at the new batch size), one timed batch, barrier + synchronize on both sides, max over
    ranks.  Every rank takes part; rank 0 reports."""
    import torch.distributed as dist
    from gm_diffusion_b200 import dist as D
    G_BATCH = 64
    out = {"global_batch": G_BATCH}

    def timed_once(fn):
        fn(); torch.cuda.synchronize()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        dist.barrier(); torch.cuda.synchronize()
        return D.max_over_ranks(e0.elapsed_time(e1) / 1000.0, dev)

    def inputs(n, seed):
        g = torch.Generator().manual_seed(seed)
        return (torch.randn(n, 77, 768, generator=g).to(dev), torch.randn(n, 77, 768, generator=g).to(dev),
                torch.randn(n, 4, HEIGHT // 8, WIDTH // 8, generator=g).to(dev))

    if G_BATCH % world == 0:
        Bl = G_BATCH // world
        pe, ne, lat = inputs(Bl, 2000 + rank)
        t = timed_once(lambda: D.gather_outputs(pipe(prompt_embeds=pe, negative_prompt_embeds=ne, latents=lat, **kw)[0], G_BATCH))
        out["image_sharding"] = {"images_per_gpu": Bl, "img_per_s": G_BATCH / t, "s_per_global_batch": t, "scaling": "strong (total work fixed at 64 images)"}
        del pe, ne, lat
    if world % 2 == 0 and G_BATCH % (world // 2) == 0:
        groups = [dist.new_group([2 * k_, 2 * k_ + 1]) for k_ in range(world // 2)]     # (every rank creates every group)
        Bp = G_BATCH // (world // 2)
        pe, ne, lat = inputs(Bp, 3000 + rank // 2)                                      # both ranks of a pair: the same images
        pipe.enable_cfg_pair(groups[rank // 2])
        try:
            t = timed_once(lambda: pipe(prompt_embeds=pe, negative_prompt_embeds=ne, latents=lat, **kw)[0])
        finally:
            pipe.cfg_pair = None
        out["cfg_pair"] = {"images_per_gpu_pair": Bp, "img_per_s": G_BATCH / t, "s_per_global_batch": t,
                           "note": "CFG halves of the SDR UNet split across each GPU pair, GM UNet split by images, one eps all-gather per UNet per step "
                                   "(NCCL, eager stepping); both ranks of a pair hold the full result, no output gather timed"}
    torch.cuda.empty_cache()
    return out


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
    config2 = None
    if world > 1 and not args.no_config2:
        try:
            config2 = run_config2(pipe, dev, rank, world, kw)
        except Exception as ex:
            config2 = {"unavailable": f"{type(ex).__name__}: {ex}"}
    if rank != 0:
        if world > 1:
            torch.distributed.barrier()
        return 0

    # ---- roofline of the dominant kernel family, measured live from CUDA-graph replay (outside the timed region) ----
    peak_tf, peak_hbm, peak_src = load_peaks()
    ms_denoise = 1000.0 * t_lat / EVALS
    fam, hot = family_times(pipe, B)
    tot_ms = sum(v["ms"] for v in fam.values())
    gm_ = fam.get("gemm", {"ms": 1e-9, "flop": 0.0, "launches": 1, "bytes": 0.0})
    ach_tf = gm_["flop"] / (gm_["ms"] * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": "gmd::gemm_kernel<BN,STAGES> (tcgen05 GEMM + implicit-GEMM conv; all instantiations)",
                "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf, "peak_source": peak_src,
                "traffic": (ncu_traffic() or {}).get("dram_bytes_per_launch"), "traffic_detail": ncu_traffic(),
                "algorithmic_bytes_per_launch": gm_.get("bytes", 0.0) / gm_["launches"],
                "launches_per_denoise_step": gm_["launches"], "avg_launch_ms": gm_["ms"] / gm_["launches"],
                "algorithmic_gflop_per_denoise_step": gm_["flop"] / 1e9, "share_of_step": gm_["ms"] / tot_ms,
                "how": "the family's calls of one denoise step replayed from a CUDA graph of their own (CUDA events, median of 7); the three "
                       "family graphs sum to families_sum_ms = the step on ONE stream; ms_per_denoise_step (whole loop, one graph, GM branch "
                       "of step i on a side stream beside the SDR UNet of step i+1) is shorter because the two UNets fill each other's kernel tails",
                "families_sum_ms": tot_ms, "hottest_instantiations": [dict(h_, frac=round(h_["tflops"] / peak_tf, 3)) for h_ in hot]}
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
    if config2 is not None:
        line["config2_global_batch_64"] = config2
    if world == 1 and not args.no_comparator:
        try:
            line["gpu_comparator"] = gpu_comparator(B)
            # our SDR forward of the same 2B samples, alone on the GPU, from a CUDA graph of its own (inside the loop graph it shares
            # the GPU with the GM branch of the previous step, so the step is shorter than SDR + GM forward)
            line["gpu_comparator"]["unet_forward_2B_samples_ms_ours_in_graph"] = round(sdr_forward_ms(pipe, B), 3)
        except Exception as ex:
            line["gpu_comparator"] = {"unavailable": f"{type(ex).__name__}: {ex}"}
        try:
            line["extra_configs"] = extra_configs(pipe)
        except Exception as ex:
            line["extra_configs"] = {"unavailable": f"{type(ex).__name__}: {ex}"}
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
    ap.add_argument("--no-config2", action="store_true", help="N > 1: skip the BASELINE configs[2] block (global batch 64, image sharding and CFG pairs)")
    ap.add_argument("--no-comparator", action="store_true", help="skip the GPU comparator block and the extra configs (quick runs)")
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
