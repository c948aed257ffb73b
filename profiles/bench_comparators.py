"""GPU-side comparators on the same box: what the reference's own stack (torch eager bf16 on cuDNN / cuBLAS / SDPA,
which is what diffusers runs) gets on this B200 for the same shapes, next to the hand-written kernels.
  (a) attention: torch SDPA (default / flash / cuDNN / efficient backends) and flash_attn 2.x
  (b) 3x3 convolution: cuDNN (channels_last bf16, benchmark mode) and GroupNorm+SiLU as torch ops
  (c) CFG + x0 + PLMS update + concat as the un-fused torch op chain of the oracle
  (d) Eq.(1) (+ fix_mulog + gamut) as the un-fused torch ops of the tone-mapping oracle
  (U) one SDR UNet forward on the CFG batch (16 samples, 64x64 latents): the oracle network in bf16 torch eager with SDPA
CUDA events on the current stream, 3 warm-ups, median of 10, a 256 MB write between iterations to flush L2.
The torch restatements are the oracle's modules (tone-mapping functions, SD1.5-architecture UNet); they are handed in by `bench.py
--comparators` (the one place allowed to execute `oracle/`), this file never imports them.
Run on a B200:  python bench.py --comparators > gpurun_out/comparators_r01.json"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.nn.functional as F

import gm_diffusion_b200 as G
from gm_diffusion_b200 import ops
from gm_diffusion_b200.stage1 import tone_mapping as TM


def run(TMO, UO, build_pipeline):
    """TMO / UO: the oracle's tone-mapping and UNet modules (torch restatements, run here ON THE GPU as comparators)."""
    dev = torch.device("cuda:0")
    bf16 = torch.bfloat16
    torch.backends.cudnn.benchmark = True
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {"box": torch.cuda.get_device_name(0), "torch": torch.__version__, "attention": [], "conv3x3": [], "groupnorm_silu": [], "kernel_c": [],
           "kernel_d": [], "unet_forward": {}}


    def timeit(fn, n=10, warm=3):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]


    def try_time(fn, **kw):
        try:
            return round(timeit(fn, **kw), 4)
        except Exception as e:  # a backend that refuses the shape (e.g. head dim) is reported, not hidden
            torch.cuda.synchronize()
            return f"unavailable: {type(e).__name__}: {str(e)[:80]}"


    g = torch.Generator(device=dev).manual_seed(0)
    # ---- (a) attention -------------------------------------------------------------------------------------------------------
    from torch.nn.attention import SDPBackend, sdpa_kernel
    try:
        from flash_attn import flash_attn_func
    except Exception as e:  # pragma: no cover
        flash_attn_func = None
        out["flash_attn_import"] = f"{type(e).__name__}: {e}"
    for B, Nq, Nk, C in [(16, 4096, 4096, 320), (16, 1024, 1024, 640), (16, 256, 256, 1280), (16, 4096, 77, 320), (1, 16384, 16384, 320)]:
        H, d = 8, C // 8
        q = torch.randn(B, Nq, C, device=dev, generator=g).to(bf16)
        k = torch.randn(B, Nk, C, device=dev, generator=g).to(bf16)
        v = torch.randn(B, Nk, C, device=dev, generator=g).to(bf16)
        q4, k4, v4 = (t.view(B, -1, H, d).transpose(1, 2) for t in (q, k, v))   # [B,H,N,d] views, as diffusers' AttnProcessor2_0 builds them
        fl = 4.0 * B * Nq * Nk * C
        row = {"B": B, "Nq": Nq, "Nk": Nk, "d": d, "gflop": round(fl / 1e9, 2), "ms": {}}
        row["ms"]["gm_diffusion_b200"] = try_time(lambda: ops.attention(q, k, v, H))
        row["ms"]["torch_sdpa_default"] = try_time(lambda: F.scaled_dot_product_attention(q4, k4, v4))
        for name, be in (("torch_sdpa_flash", SDPBackend.FLASH_ATTENTION), ("torch_sdpa_cudnn", SDPBackend.CUDNN_ATTENTION),
                         ("torch_sdpa_efficient", SDPBackend.EFFICIENT_ATTENTION)):
            def run_be(be=be):
                with sdpa_kernel(be):
                    return F.scaled_dot_product_attention(q4, k4, v4)
            row["ms"][name] = try_time(run_be)
        if flash_attn_func is not None:
            qf, kf, vf = (t.view(B, -1, H, d) for t in (q, k, v))
            row["ms"]["flash_attn_2"] = try_time(lambda: flash_attn_func(qf, kf, vf))
        row["tflops"] = {n: round(fl / ms / 1e9, 1) for n, ms in row["ms"].items() if isinstance(ms, float)}
        out["attention"].append(row)
        del q, k, v
    print(json.dumps(out["attention"]), file=sys.stderr, flush=True)

    # ---- (b) 3x3 convolution + GroupNorm/SiLU ----------------------------------------------------------------------------------
    for N, Hh, Cin, Cout in [(16, 64, 320, 320), (16, 32, 640, 640), (16, 16, 1280, 1280), (16, 8, 1280, 1280)]:
        x = torch.randn(N, Hh, Hh, Cin, device=dev, generator=g).to(bf16)
        w = (torch.randn(Cout, Cin, 3, 3, device=dev, generator=g) * 0.02).to(bf16)
        b = torch.randn(Cout, device=dev, generator=g)
        wt = ops.pack_conv_weight_tiled(w)
        x_cl = x.permute(0, 3, 1, 2)   # NCHW view of NHWC memory = channels_last
        w_cl = w.contiguous(memory_format=torch.channels_last)
        bb = b.to(bf16)
        fl = 2.0 * N * Hh * Hh * Cout * 9 * Cin
        ms = {"gm_diffusion_b200": try_time(lambda: ops.conv2d(x, wt, Cout, bias=b)),
              "cudnn_channels_last": try_time(lambda: F.conv2d(x_cl, w_cl, bb, padding=1)),
              "cudnn_nchw": try_time(lambda xx=x_cl.contiguous(), ww=w.contiguous(): F.conv2d(xx, ww, bb, padding=1))}
        out["conv3x3"].append({"N": N, "H": Hh, "Cin": Cin, "Cout": Cout, "gflop": round(fl / 1e9, 1), "ms": ms,
                               "tflops": {n: round(fl / t / 1e9, 1) for n, t in ms.items() if isinstance(t, float)}})
        gamma, beta = torch.randn(Cin, device=dev, generator=g), torch.randn(Cin, device=dev, generator=g)
        gb, bb2 = gamma.to(bf16), beta.to(bf16)
        msn = {"gm_diffusion_b200": try_time(lambda: ops.groupnorm_silu(x, gamma, beta)),
               "torch_group_norm_then_silu_channels_last": try_time(lambda: F.silu(F.group_norm(x_cl, 32, gb, bb2, 1e-5))),
               "torch_group_norm_then_silu_nchw": try_time(lambda xx=x_cl.contiguous(): F.silu(F.group_norm(xx, 32, gb, bb2, 1e-5)))}
        out["groupnorm_silu"].append({"N": N, "H": Hh, "C": Cin, "bytes_algorithmic": 2 * x.numel() * 2, "ms": msn})
        del x, w, wt
    print(json.dumps(out["conv3x3"]), file=sys.stderr, flush=True)

    # ---- (c) CFG + x0 + PLMS + concat: the oracle's un-fused torch chain vs the fused kernel ------------------------------------
    for B in (8, 512):
        h = w = 64
        n_px = B * h * w
        x = torch.randn(B, 4, h, w, device=dev, generator=g)
        eu, ec = torch.randn(B, 4, h, w, device=dev, generator=g), torch.randn(B, 4, h, w, device=dev, generator=g)
        hist = [torch.randn(B, 4, h, w, device=dev, generator=g) for _ in range(3)]
        a_t, a_p = 0.35, 0.40

        def torch_chain():
            e = eu + 7.5 * (ec - eu)
            x0 = (x - (1 - a_t) ** 0.5 * e) / a_t ** 0.5
            e4 = (55 * e - 59 * hist[0] + 37 * hist[1] - 9 * hist[2]) / 24
            xn = (a_p / a_t) ** 0.5 * x - (a_p - a_t) * e4 / (a_t * (1 - a_p) ** 0.5 + (a_t * (1 - a_t) * a_p) ** 0.5)
            gm_in = torch.cat([x0, x], dim=1).to(bf16)
            sdr_in = torch.cat([xn] * 2).to(bf16)
            return xn, gm_in, sdr_in

        out["kernel_c"].append({"batch": B, "ms_torch_unfused_sdr_half": try_time(torch_chain),
                                "note": "fused kernel: profiles/kernels_r01.json (SDR + GM halves: 0.036 ms at batch 8, launch-bound)"})

    # ---- (d) Eq.(1) [+ fix_mulog + gamut] -------------------------------------------------------------------------------------
    for B in (1, 4):
        sdr = torch.rand(B, 3, 2160, 3840, device=dev, generator=g)
        gm = torch.rand(B, 3, 2160, 3840, device=dev, generator=g)
        px = B * 2160 * 3840
        ms = {"gm_diffusion_b200 eq1": try_time(lambda: TM.apply_gm_to_sdr(gm, sdr, 99.0)),
              "torch eq1 (reference op sequence)": try_time(lambda: TMO.apply_gm_to_sdr(gm, sdr, 99.0)),
              "gm_diffusion_b200 eq1->fix_mulog->gamut": try_time(lambda: TM.reconstruct_hdr(sdr, gm, 99.0, tmo="fix_mulog", gamut=True, return_hdr=False)),
              "torch eq1->fix_mulog->gamut (reference op sequence)": try_time(lambda: TMO.gamut_compress(TMO.fix_mulog_tmo(TMO.apply_gm_to_sdr(gm, sdr, 99.0), 99.0)))}
        out["kernel_d"].append({"batch": B, "px": px, "ms": ms, "GBps_36B_per_px": {n: round(px * 36 / t / 1e6, 1) for n, t in ms.items() if isinstance(t, float)}})
        del sdr, gm
    torch.cuda.empty_cache()

    # ---- (U) one SDR UNet forward on the CFG batch ---------------------------------------------------------------------------
    def sdpa_forward(self, x, ctx=None):   # diffusers' AttnProcessor2_0 (SDPA) instead of the oracle's explicit softmax(QK^T)V
        ctx = x if ctx is None else ctx
        b, n, c = x.shape
        h, d = self.heads, c // self.heads
        q = self.to_q(x).view(b, n, h, d).transpose(1, 2)
        k = self.to_k(ctx).view(b, -1, h, d).transpose(1, 2)
        v = self.to_v(ctx).view(b, -1, h, d).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, n, c)
        return self.to_out[0](o)


    UO.Attention.forward = sdpa_forward
    B = 8
    with torch.device(dev):
        ref = UO.UNet2DConditionOracle(4).eval().to(bf16).to(memory_format=torch.channels_last)
    lat = torch.randn(2 * B, 4, 64, 64, device=dev, generator=g).to(bf16).contiguous(memory_format=torch.channels_last)
    ctx = torch.randn(2 * B, 77, 768, device=dev, generator=g).to(bf16)
    with torch.no_grad():
        ms_ref = try_time(lambda: ref(lat, 501, ctx), n=5)
    del ref
    torch.cuda.empty_cache()
    pipe = build_pipeline(dev)
    kv = pipe.unet.project_context(ctx.float())
    tb = pipe.unet.timestep_table([501])
    xs = torch.randn(B, 64, 64, 8, device=dev, generator=g).to(bf16)
    ms_ours = try_time(lambda: pipe.unet.forward(xs, tb, kv, cfg_shared=True), n=5)
    out["unet_forward"] = {"samples": 2 * B, "latent": "64x64", "gflop": round(2 * B * 803.3, 1),
                           "ms": {"gm_diffusion_b200 (eager launches, CFG-shared prefix, cached text K/V and timestep table)": ms_ours,
                                  "torch eager bf16 channels_last, cuDNN benchmark, SDPA (the reference's stack)": ms_ref},
                           "note": "the bench replays the same forward from a CUDA graph (the bench line: ms_per_denoise_step covers SDR(16) + GM(8) forwards, gpu_comparator.unet_forward_2B_samples_ms_ours_in_graph is this forward alone)"}
    return out
