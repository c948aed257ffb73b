"""Drop-in for gm_diffusion/pipelines/stable_diffusion_dual_unet.py: `StableDiffusionDualUNetPipeline` with the
reference's `__call__` signature (:784-812) and return convention (:1122-1132), executing the joint SDR /
gain-map denoising loop (:1040-1093) on hand-written sm_100a kernels:

    per step:  SDR UNet (2B under CFG)  ->  fused CFG + x0 + PLMS/DDIM + concat  ->  GM UNet (B)  ->  fused step

Differences from the reference, all deliberate:
  * GM-branch text conditioning uses the batch-correct conditional slice
    (`prompt_embeds[negative_prompt_embeds.shape[0]:]`, visualize_latents.py:274); identical to the reference's
    `prompt_embeds[1:]` (:1086) at batch 1, and the only form that works for batch > 1.
  * `output_type="latent"` returns `(latents, gm_latents)` exactly like the reference.  Other output types are broken in
    the reference (:1118-1131 index the batch of a single SDR decode); here they decode BOTH latents:
    "pt"/"np"/"pil" -> `(sdr_images, gm_images)`, and "hdr" additionally applies Eq.(1) on the GPU and returns
    `(hdr [B,H,W,3] fp32, sdr, gm)`; "disk" returns what the scripts write to files, produced by the same single launch:
    `(rgbe uint8 [B,H,W,4], sdr_u8 [B,H,W,3], gm_u8 [B,H,W,3])` (generate_hdr.py:243-244 and save_hdr_image :27-30).
  * cross-attention K/V and the timestep-embedding MLP are hoisted out of the loop; latents / PLMS history / CFG math
    stay in fp32 regardless of the UNet compute dtype (bf16).
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Union

import torch

from .. import _lib as L
from .. import schedulers as S
from ..stage1 import tone_mapping as TM
from ._common import (PipelineBase, StableDiffusionPipelineOutput, as_b200_unet, bf16, randn_tensor, retrieve_timesteps)


class _Workspace:
    """Static device buffers for one (batch, h, w): everything the loop touches has a fixed address (CUDA graphs)."""

    def __init__(self, B, h, w, cfg_mult, device, temb_cols_sdr, temb_cols_gm):
        n_px = B * h * w
        self.B, self.h, self.w, self.n_px = B, h, w, n_px
        self.sdr = S.BranchState(n_px, device)
        self.gm = S.BranchState(n_px, device)
        self.unet_in = torch.zeros((B, h, w, 8), dtype=bf16, device=device)  # ONE copy: the CFG halves share it (B200UNet cfg_shared)
        self.gm_in = torch.zeros((B, h, w, 8), dtype=bf16, device=device)
        self.eps_sdr = torch.empty((cfg_mult * B, h, w, 4), dtype=torch.float32, device=device)
        self.eps_gm = torch.empty((B, h, w, 4), dtype=torch.float32, device=device)
        self.temb_sdr = torch.empty((1, temb_cols_sdr), dtype=torch.float32, device=device)
        self.temb_gm = torch.empty((1, temb_cols_gm), dtype=torch.float32, device=device)
        self.rescale_ws = torch.zeros(B * 4, dtype=torch.float32, device=device)
        self.kv_sdr: Optional[List[torch.Tensor]] = None
        self.kv_gm: Optional[List[torch.Tensor]] = None

    def set_context(self, which: str, kv: List[torch.Tensor]):
        cur = getattr(self, which)
        if cur is None or any(a.shape != b.shape for a, b in zip(cur, kv)):
            setattr(self, which, kv)
        else:
            for a, b in zip(cur, kv):
                a.copy_(b)


class StableDiffusionDualUNetPipeline(PipelineBase):
    model_cpu_offload_seq = "text_encoder->image_encoder->unet->vae"

    def __init__(self, vae, text_encoder, tokenizer, unet, gm_unet, scheduler, safety_checker=None, feature_extractor=None,
                 image_encoder=None, requires_safety_checker: bool = True, device="cuda"):
        """Constructor kwargs as stable_diffusion_dual_unet.py:202-214.  `unet` / `gm_unet` may be `B200UNet` objects or
        any module exposing a diffusers UNet2DConditionModel state_dict (repacked once, here)."""
        self._init_common(vae, text_encoder, tokenizer, scheduler, safety_checker, feature_extractor, image_encoder,
                          requires_safety_checker, device)
        self.unet = as_b200_unet(unet, self.device)
        self.gm_unet = as_b200_unet(gm_unet, self.device)
        if self.unet.in_channels != 4 or self.gm_unet.in_channels != 8:
            raise ValueError(f"dual pipeline expects a 4-channel SDR UNet and an 8-channel GM UNet, got "
                             f"{self.unet.in_channels} / {self.gm_unet.in_channels}")
        self.gm_scheduler = None
        self._ws: Dict[Any, _Workspace] = {}
        self._pair_ws: Dict[Any, Dict[str, Any]] = {}
        self._loop_graphs: Dict[Any, Any] = {}
        self.use_loop_graph = True   # capture the whole denoising loop as one CUDA graph when nothing needs the host between steps
        # Inside the loop graph the GM branch of step i (GM UNet + its scheduler step) runs on a side stream beside the SDR UNet of
        # step i+1: the SDR forward never reads GM state, only the SDR scheduler step does (it writes the GM UNet's input, whose tail is
        # the GM latents).  Same kernels on the same data — bit-identical to the one-stream order — but the tail of every kernel of one
        # UNet is filled by CTAs of the other.  GMD_TWO_STREAMS=0 keeps one stream (A/B).
        self.use_two_streams = __import__("os").environ.get("GMD_TWO_STREAMS", "1") != "0"
        self._side_stream = None
        self.cfg_pair = None   # gm_diffusion_b200.dist.CfgPair: split the CFG halves / the GM images over two ranks (latency mode)

    def enable_cfg_pair(self, group=None):
        """SURVEY.md §8e optional mode: see `gm_diffusion_b200.dist.CfgPair`.  Both ranks must call the pipeline with identical
        arguments; both return the full result."""
        from ..dist import CfgPair
        self.cfg_pair = CfgPair(group)
        return self

    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    @L.on_own_device
    def __call__(
        self,
        prompt: Union[str, List[str]] = None,
        height: Optional[int] = None,
        width: Optional[int] = None,
        num_inference_steps: int = 50,
        timesteps: List[int] = None,
        sigmas: List[float] = None,
        guidance_scale: float = 7.5,
        negative_prompt: Optional[Union[str, List[str]]] = None,
        num_images_per_prompt: Optional[int] = 1,
        eta: float = 0.0,
        generator: Optional[Union[torch.Generator, List[torch.Generator]]] = None,
        latents: Optional[torch.Tensor] = None,
        prompt_embeds: Optional[torch.Tensor] = None,
        negative_prompt_embeds: Optional[torch.Tensor] = None,
        ip_adapter_image=None,
        ip_adapter_image_embeds: Optional[List[torch.Tensor]] = None,
        output_type: Optional[str] = "pil",
        return_dict: bool = True,
        cross_attention_kwargs: Optional[Dict[str, Any]] = None,
        guidance_rescale: float = 0.0,
        clip_skip: Optional[int] = None,
        callback_on_step_end: Optional[Callable] = None,
        callback_on_step_end_tensor_inputs: List[str] = ["latents"],
        **kwargs,
    ):
        callback = kwargs.pop("callback", None)          # legacy callback(step, t, latents), dual_unet.py:1108-1110
        callback_steps = kwargs.pop("callback_steps", None)
        qmax = kwargs.pop("qmax", 99.0)                  # only used by the output_type="hdr" extension
        hdr_eps = kwargs.pop("hdr_eps", 1 / 64)
        # unknown kwargs (e.g. noise_level=0.0, formal_baseline.py:221) are swallowed like the reference's **kwargs (:811)

        # 0. defaults (:930-932): unet.config.sample_size (64) * vae_scale_factor
        height = height or 64 * self.vae_scale_factor
        width = width or 64 * self.vae_scale_factor
        # 1. check inputs (:934-945)
        self.check_inputs(prompt, height, width, callback_steps, negative_prompt, prompt_embeds, negative_prompt_embeds,
                          ip_adapter_image, ip_adapter_image_embeds, callback_on_step_end_tensor_inputs)
        if cross_attention_kwargs and set(cross_attention_kwargs) - {"scale"}:
            raise NotImplementedError("only cross_attention_kwargs={'scale': s} is accepted: the B200 UNets are built from plain state dicts "
                                      "and carry no LoRA layers, so the LoRA scale is a no-op exactly as in the reference without an adapter "
                                      "(formal_improved.py:268); other attention-processor kwargs are not accelerated")
        self._guidance_scale = guidance_scale
        self._guidance_rescale = guidance_rescale
        self._clip_skip = clip_skip
        self._cross_attention_kwargs = cross_attention_kwargs
        self._interrupt = False
        # 2. batch size (:953-959)
        if prompt is not None and isinstance(prompt, str):
            batch_size = 1
        elif prompt is not None and isinstance(prompt, list):
            batch_size = len(prompt)
        else:
            batch_size = prompt_embeds.shape[0]
        device = self.device
        # 3. encode prompt (:968-984)
        prompt_embeds, negative_prompt_embeds = self.encode_prompt(
            prompt, device, num_images_per_prompt, self.do_classifier_free_guidance, negative_prompt,
            prompt_embeds=prompt_embeds, negative_prompt_embeds=negative_prompt_embeds, clip_skip=self.clip_skip)
        out_dtype = prompt_embeds.dtype
        do_cfg = self.do_classifier_free_guidance
        if (do_cfg and self.skip_identical_cfg and guidance_rescale == 0.0 and negative_prompt_embeds is not None
                and negative_prompt_embeds.shape == prompt_embeds.shape and torch.equal(negative_prompt_embeds, prompt_embeds)):
            # the negative prompt IS the prompt — e.g. the SDR->HDR CLI's prompt=[""] with the default negative
            # (generate_hdr.py:212-218).  Both CFG halves are then the same forward and eps_u + g*(eps_c - eps_u) == eps_c bit for
            # bit (the kernels are deterministic and batch-independent), so the unconditional half is not run.
            do_cfg = False
        B = batch_size * num_images_per_prompt
        # 4. timesteps (:996-998)
        timesteps, num_inference_steps = retrieve_timesteps(self.scheduler, num_inference_steps, device, timesteps, sigmas)
        # 5. latents (:1002-1012); gm_latents starts as a clone of the SDR noise
        latents = self.prepare_latents(B, 4, height, width, torch.float32, device, generator, latents)
        h, w = latents.shape[-2:]
        self._num_timesteps = len(timesteps)
        self.gm_scheduler = S.clone_scheduler(self.scheduler)  # :1036-1037

        pair = self.cfg_pair if do_cfg else None          # without CFG there is nothing to split between the two ranks
        ws = self._workspace(B, h, w, 2 if do_cfg else 1)
        stream = L.current_stream()
        lat32 = latents.to(device=device, dtype=torch.float32).contiguous()
        L.check(L.lib().gmd_latents_nchw_to_px(lat32.data_ptr(), ws.sdr.x.data_ptr(), B, h * w, stream), "gmd_latents_nchw_to_px")
        ws.gm.x.copy_(ws.sdr.x)
        ws.sdr.reset(); ws.gm.reset()
        L.check(L.lib().gmd_pack_unet_input(ws.sdr.x.data_ptr(), None, ws.unet_in.data_ptr(), ws.n_px, 8, stream), "gmd_pack_unet_input")
        # :1045 torch.cat([latents] * 2) is never materialised: both CFG halves read the same buffer (B200UNet cfg_shared)
        # step-invariant work hoisted out of the loop: text K/V per layer, timestep-embedding tables
        sdr_ctx = torch.cat([negative_prompt_embeds, prompt_embeds]) if do_cfg else prompt_embeds  # :983-984
        gm_ctx = prompt_embeds  # conditional half only, no CFG on the GM branch (:1086; batch-correct form VIS:274)
        ts = [int(t) for t in timesteps]
        table_sdr = self.unet.timestep_table(ts)
        table_gm = self.gm_unet.timestep_table(ts)
        extra = self.prepare_extra_step_kwargs(generator, eta)
        # The WHOLE loop as one CUDA graph: every scheduler coefficient, time-embedding row and history-ring
        # slot of every step is a function of (scheduler, timesteps, guidance) only, so the 51 x (row select, UNet forward, fused
        # step) x 2 sequence is captured once per such key and a call is ONE graph launch.  Eager stepping (with one graph per UNet
        # forward) remains for everything that needs the host between steps: callbacks (the reference's cooperative cancel can only
        # be requested from one), per-step ancestral noise drawn from the caller's generator (DDPM, DDIM eta > 0), the CFG-pair
        # exchange, and use_cuda_graph / use_loop_graph = False.
        stochastic = isinstance(self.scheduler, S.DDPMScheduler) or (isinstance(self.scheduler, S.DDIMScheduler) and extra["eta"] > 0)
        loop_graph_ok = (self.use_cuda_graph and self.use_loop_graph and pair is None and callback is None and callback_on_step_end is None
                         and not stochastic)
        loop_key = (B, h, w, do_cfg, type(self.scheduler).__name__, tuple(ts), float(guidance_scale), float(guidance_rescale if do_cfg else 0.0),
                    table_sdr.data_ptr(), table_gm.data_ptr(), self.use_two_streams)
        if pair is None:
            ws.set_context("kv_sdr", self.unet.project_context(sdr_ctx))
            ws.set_context("kv_gm", self.gm_unet.project_context(gm_ctx))
            if loop_graph_ok:
                # (torch refuses to replay a graph inside a capture, so the loop graph holds the UNet kernels themselves)
                run_sdr = lambda: self.unet.forward(ws.unet_in, ws.temb_sdr, ws.kv_sdr, out=ws.eps_sdr, cfg_shared=do_cfg)
                run_gm = lambda: self.gm_unet.forward(ws.gm_in, ws.temb_gm, ws.kv_gm, out=ws.eps_gm)
            else:
                run_sdr = self._unet_runner(("sdr", B, h, w, do_cfg), self.unet, ws.unet_in, ws.temb_sdr, ws.kv_sdr, ws.eps_sdr, cfg_shared=do_cfg)
                run_gm = self._unet_runner(("gm", B, h, w), self.gm_unet, ws.gm_in, ws.temb_gm, ws.kv_gm, ws.eps_gm)
            gm_lo, gm_hi = 0, B
        else:
            # CFG-pair mode: this rank's CFG half of the SDR UNet on all B images, and its half of the images for the GM UNet;
            # ws.eps_sdr [2B] is then filled by an all-gather (rank 0 = uncond rows, rank 1 = cond rows), ws.eps_gm likewise
            half_ctx = negative_prompt_embeds if pair.rank == 0 else prompt_embeds
            gm_lo, gm_hi = (0, B) if B % 2 else (pair.rank * (B // 2), (pair.rank + 1) * (B // 2))
            pw = self._pair_ws.setdefault((B, h, w), {})
            if not pw:
                f32 = dict(dtype=torch.float32, device=device)
                pw.update(eps_half=torch.empty((B, h, w, 4), **f32), eps_gm_part=torch.empty((gm_hi - gm_lo, h, w, 4), **f32), kv_sdr=None, kv_gm=None)
            for name, kv in (("kv_sdr", self.unet.project_context(half_ctx)), ("kv_gm", self.gm_unet.project_context(gm_ctx[gm_lo:gm_hi]))):
                if pw[name] is None:
                    pw[name] = kv
                else:
                    for a_, b_ in zip(pw[name], kv):
                        a_.copy_(b_)
            run_sdr = self._unet_runner(("sdr-pair", B, h, w, pair.rank), self.unet, ws.unet_in, ws.temb_sdr, pw["kv_sdr"], pw["eps_half"])
            run_gm = self._unet_runner(("gm-pair", B, h, w, pair.rank), self.gm_unet, ws.gm_in[gm_lo:gm_hi], ws.temb_gm, pw["kv_gm"], pw["eps_gm_part"])
        eps_u = ws.eps_sdr[:B].reshape(-1, 4) if do_cfg else None
        eps_c = (ws.eps_sdr[B:] if do_cfg else ws.eps_sdr).reshape(-1, 4)
        eps_g = ws.eps_gm.reshape(-1, 4)

        # 7a. the loop on two streams (loop graph only): stream order within a branch, events between the branches
        def denoise_loop_two_streams(progress_bar):
            from .. import ops
            main = torch.cuda.current_stream()
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream(device=device)
            side = self._side_stream
            gm_done = None
            for i, t in enumerate(ts):
                ws.temb_sdr.copy_(table_sdr[i:i + 1])
                run_sdr()                                                       # SDR eps of step i  ||  GM branch of step i-1
                if gm_done is not None:
                    main.wait_event(gm_done)                                    # GM latents of step i-1 written, ws.gm_in no longer read
                S.fused_step(self.scheduler.plan_step(t, extra["eta"]), ws.sdr, eps_c, eps_u, guidance_scale=guidance_scale,
                             guidance_rescale=guidance_rescale if do_cfg else 0.0, px_per_sample=h * w,
                             x0_coeffs=self.scheduler.x0_coeffs(t), unet_in_next=ws.unet_in, unet_in_dup=1,
                             concat_out=ws.gm_in, concat_tail=ws.gm.x, rescale_ws=ws.rescale_ws)
                sdr_stepped = torch.cuda.Event()
                sdr_stepped.record(main)
                side.wait_event(sdr_stepped)
                with torch.cuda.stream(side), ops.scratch_slot(1):
                    ws.temb_gm.copy_(table_gm[i:i + 1])
                    run_gm()
                    S.fused_step(self.gm_scheduler.plan_step(t, extra["eta"]), ws.gm, eps_g, x0_coeffs=self.gm_scheduler.x0_coeffs(t))
                    gm_done = torch.cuda.Event()
                    gm_done.record(side)
                progress_bar.update()
            if gm_done is not None:
                main.wait_event(gm_done)

        # 7. denoising loop (:1040-1113)
        def denoise_loop(progress_bar):
            for i, t in enumerate(ts):
                if self.interrupt:
                    continue
                ws.temb_sdr.copy_(table_sdr[i:i + 1])
                ws.temb_gm.copy_(table_gm[i:i + 1])
                run_sdr()                                                       # :1052-1060  SDR eps (uncond | cond)
                if pair is not None:
                    pair.gather(pw["eps_half"], ws.eps_sdr)                     # [uncond rows of rank 0 | cond rows of rank 1]
                plan = self.scheduler.plan_step(t, extra["eta"])
                if plan.needs_noise:                                            # DDIM eta > 0: SDR draw first, then GM (§8a-Q7)
                    ws.sdr.noise = self._draw_noise(B, h, w, generator)
                S.fused_step(plan, ws.sdr, eps_c, eps_u, guidance_scale=guidance_scale,        # :1063-1080
                             guidance_rescale=guidance_rescale if do_cfg else 0.0, px_per_sample=h * w,
                             x0_coeffs=self.scheduler.x0_coeffs(t), unet_in_next=ws.unet_in, unet_in_dup=1,
                             concat_out=ws.gm_in, concat_tail=ws.gm.x, rescale_ws=ws.rescale_ws)
                run_gm()                                                        # :1083-1092  GM eps, no CFG
                if pair is not None:
                    if gm_hi - gm_lo == B:
                        ws.eps_gm.copy_(pw["eps_gm_part"])                      # odd batch: both ranks ran all images
                    else:
                        pair.gather(pw["eps_gm_part"], ws.eps_gm)
                gplan = self.gm_scheduler.plan_step(t, extra["eta"])
                if gplan.needs_noise:
                    ws.gm.noise = self._draw_noise(B, h, w, generator)
                S.fused_step(gplan, ws.gm, eps_g, x0_coeffs=self.gm_scheduler.x0_coeffs(t))    # :1093
                progress_bar.update()
                if callback is not None and callback_steps and i % callback_steps == 0:
                    callback(i, t, self._latents_nchw(ws.sdr.x, B, h, w))

        with self.progress_bar(total=num_inference_steps) as progress_bar:
            if not loop_graph_ok:
                denoise_loop(progress_bar)
            else:
                ent = self._loop_graphs.get(loop_key)
                if ent is None:
                    two = self.use_two_streams and not self.interrupt
                    ws.temb_sdr.copy_(table_sdr[0:1]); ws.temb_gm.copy_(table_gm[0:1])
                    run_sdr()                          # eager warm-up: lazily allocated scratch exists before the capture; eps buffers are rewritten
                    if two:
                        from .. import ops
                        with ops.scratch_slot(1):
                            run_gm()
                    else:
                        run_gm()
                    torch.cuda.synchronize()
                    n0 = L.launch_count()
                    lg = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(lg):
                        (denoise_loop_two_streams if two else denoise_loop)(progress_bar)
                    n_kernels = L.launch_count() - n0      # launches recorded by the capture pass = kernels per replay
                    L.lib().gmd_add_launch_count(-n_kernels)   # (they did not run)
                    if len(self._loop_graphs) >= 4:
                        self._loop_graphs.clear()
                    self._loop_graphs[loop_key] = ent = (lg, n_kernels, (table_sdr, table_gm))  # (the tables must outlive their cache entry)
                ent[0].replay()
                self.graph_launches += ent[1]

        sdr_lat = self._latents_nchw(ws.sdr.x, B, h, w)
        gm_lat = self._latents_nchw(ws.gm.x, B, h, w)
        if output_type == "latent":
            return (sdr_lat.to(out_dtype), gm_lat.to(out_dtype))  # :1122-1132
        return self._decode_outputs(ws, B, h, w, output_type, qmax, hdr_eps)

    # ------------------------------------------------------------------------------------------------------------
    def _workspace(self, B, h, w, cfg_mult) -> _Workspace:
        key = (B, h, w, cfg_mult)
        ws = self._ws.get(key)
        if ws is None:
            ws = _Workspace(B, h, w, cfg_mult, self.device, self.unet.w_temb.shape[0], self.gm_unet.w_temb.shape[0])
            self._ws[key] = ws
        return ws

    def _draw_noise(self, B, h, w, generator):
        z = randn_tensor((B, 4, h, w), generator=generator, device=self.device, dtype=torch.float32)
        out = torch.empty((B * h * w, 4), dtype=torch.float32, device=self.device)
        L.check(L.lib().gmd_latents_nchw_to_px(z.contiguous().data_ptr(), out.data_ptr(), B, h * w, L.current_stream()))
        return out

    def _latents_nchw(self, x_px, B, h, w):
        out = torch.empty((B, 4, h, w), dtype=torch.float32, device=self.device)
        L.check(L.lib().gmd_latents_px_to_nchw(x_px.data_ptr(), out.data_ptr(), B, h * w, L.current_stream()), "gmd_latents_px_to_nchw")
        return out

    def _decode_outputs(self, ws, B, h, w, output_type, qmax, hdr_eps):
        if self.vae is None:
            raise ValueError(f"output_type={output_type!r} needs a VAE; pass output_type='latent' or construct the pipeline with one")
        sdr_img = self.vae.decode_px(ws.sdr.x, B, h, w)   # bf16 NHWC in [-1,1]
        gm_img = self.vae.decode_px(ws.gm.x, B, h, w)
        if output_type == "hdr":
            # de-normalise + Eq.(1) in ONE kernel (generate_hdr.py:227,232,256-265 run this on the host in numpy; no clamp there)
            hdr, _ = TM.reconstruct_hdr(sdr_img, gm_img, qmax=qmax, eps=hdr_eps, denormalize=True, clamp=False, channels_last=True)
            return hdr, sdr_img, gm_img
        if output_type == "disk":
            # same launch, but emitting what the scripts write to disk (generate_hdr.py:243-244 PNG uint8, :27-30 Radiance RGBE of
            # hdr/(qmax+1)): 10 B/px leave the GPU instead of 24 B/px of fp32, and `hdr_io.pack_radiance` only adds the container
            return TM.reconstruct_for_disk(sdr_img, gm_img, qmax=qmax, eps=hdr_eps, denormalize=True, clamp=False, channels_last=True)
        outs = []
        for img in (sdr_img, gm_img):
            x = (img.float() / 2 + 0.5).clamp(0, 1)       # VaeImageProcessor.postprocess de-normalise
            if output_type == "pt":
                outs.append(x.permute(0, 3, 1, 2))
            elif output_type == "np":
                outs.append(x.cpu().numpy())
            elif output_type == "pil":
                from PIL import Image
                arr = (x.cpu().numpy() * 255).round().astype("uint8")
                outs.append([Image.fromarray(a) for a in arr])
            else:
                raise ValueError(f"unknown output_type {output_type!r}")
        return tuple(outs)
