"""Drop-in for gm_diffusion/pipelines/stable_diffusion_gm.py: `StableDiffusionGMPipeline` (SDR latent -> gain-map
latent) with the reference's `__call__(sdr_latent, prompt, ...)` signature (:782-811) and return convention
(:1106-1114).  One 8-channel UNet on `cat([sdr_latent, latents], 1)` under CFG (:1045-1064), scheduler step (:1071),
live `callback_on_step_end` (:1073-1081) — on the sm_100a kernels.  Latents are sized from `sdr_latent`
(:1009-1010; `height` / `width` are ignored exactly like the reference, §8a-Q8)."""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Union

import torch

from .. import _lib as L
from .. import schedulers as S
from ._common import PipelineBase, StableDiffusionPipelineOutput, as_b200_unet, bf16, retrieve_timesteps


class StableDiffusionGMPipeline(PipelineBase):
    def __init__(self, vae, text_encoder, tokenizer, unet, scheduler, safety_checker=None, feature_extractor=None,
                 image_encoder=None, requires_safety_checker: bool = True, device="cuda"):
        self._init_common(vae, text_encoder, tokenizer, scheduler, safety_checker, feature_extractor, image_encoder,
                          requires_safety_checker, device)
        self.unet = as_b200_unet(unet, self.device)
        if self.unet.in_channels != 8:
            raise ValueError(f"StableDiffusionGMPipeline expects the 8-channel GM UNet, got in_channels={self.unet.in_channels}")
        self._ws: Dict[Any, dict] = {}
        self._loop_graphs: Dict[Any, Any] = {}
        self.use_loop_graph = True   # capture the whole denoising loop as one CUDA graph when nothing needs the host between steps

    @torch.no_grad()
    @L.on_own_device
    def __call__(
        self,
        sdr_latent: torch.Tensor,
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
        callback_steps = kwargs.pop("callback_steps", None)
        kwargs.pop("callback", None)
        height = height or 64 * self.vae_scale_factor
        width = width or 64 * self.vae_scale_factor
        self.check_inputs(prompt, height, width, callback_steps, negative_prompt, prompt_embeds, negative_prompt_embeds,
                          ip_adapter_image, ip_adapter_image_embeds, callback_on_step_end_tensor_inputs)
        if cross_attention_kwargs and set(cross_attention_kwargs) - {"scale"}:
            raise NotImplementedError("only cross_attention_kwargs={'scale': s} is accepted: the B200 UNets are built from plain state dicts "
                                      "and carry no LoRA layers, so the LoRA scale is a no-op exactly as in the reference without an adapter "
                                      "(formal_improved.py:268); other attention-processor kwargs are not accelerated")
        self._guidance_scale, self._guidance_rescale, self._clip_skip = guidance_scale, guidance_rescale, clip_skip
        self._interrupt = False
        if prompt is not None and isinstance(prompt, str):
            batch_size = 1
        elif prompt is not None and isinstance(prompt, list):
            batch_size = len(prompt)
        else:
            batch_size = prompt_embeds.shape[0]
        device = self.device
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
        timesteps, num_inference_steps = retrieve_timesteps(self.scheduler, num_inference_steps, device, timesteps, sigmas)
        # gm.py:1005-1015: 4 latent channels, spatial size from sdr_latent (height/width ignored)
        hh, ww = sdr_latent.shape[-2] * self.vae_scale_factor, sdr_latent.shape[-1] * self.vae_scale_factor
        latents = self.prepare_latents(B, 4, hh, ww, torch.float32, device, generator, latents)
        h, w = latents.shape[-2:]
        if sdr_latent.shape[0] != B:
            raise ValueError(f"sdr_latent batch {sdr_latent.shape[0]} != effective batch {B}")
        self._num_timesteps = len(timesteps)
        n_px = B * h * w
        mult = 2 if do_cfg else 1
        key = (B, h, w, mult)
        ws = self._ws.get(key)
        if ws is None:
            f32 = dict(dtype=torch.float32, device=device)
            ws = dict(state=S.BranchState(n_px, device), sdr_px=torch.empty(n_px, 4, **f32),
                      unet_in=torch.zeros((B, h, w, 8), dtype=bf16, device=device),  # one copy; CFG halves share it
                      eps=torch.empty((mult * B, h, w, 4), **f32), temb=torch.empty((1, self.unet.w_temb.shape[0]), **f32),
                      rescale=torch.zeros(B * 4, **f32), kv=None)
            self._ws[key] = ws
        st: S.BranchState = ws["state"]
        stream = L.current_stream()
        lat32 = latents.to(device=device, dtype=torch.float32).contiguous()
        sdr32 = sdr_latent.to(device=device, dtype=torch.float32).contiguous()
        L.check(L.lib().gmd_latents_nchw_to_px(lat32.data_ptr(), st.x.data_ptr(), B, h * w, stream))
        L.check(L.lib().gmd_latents_nchw_to_px(sdr32.data_ptr(), ws["sdr_px"].data_ptr(), B, h * w, stream))
        st.reset()
        L.check(L.lib().gmd_pack_unet_input(ws["sdr_px"].data_ptr(), st.x.data_ptr(), ws["unet_in"].data_ptr(), n_px, 8, stream))
        # gm.py:1047 torch.cat([x] * 2) is never materialised (B200UNet cfg_shared)
        ctx = torch.cat([negative_prompt_embeds, prompt_embeds]) if do_cfg else prompt_embeds
        kv = self.unet.project_context(ctx)
        if ws["kv"] is None or any(a.shape != b.shape for a, b in zip(ws["kv"], kv)):
            ws["kv"] = kv
        else:
            for a, b in zip(ws["kv"], kv):
                a.copy_(b)
        ts = [int(t) for t in timesteps]
        table = self.unet.timestep_table(ts)
        extra = self.prepare_extra_step_kwargs(generator, eta)
        # the whole loop as one CUDA graph when nothing needs the host between steps (same rule as the dual pipeline,
        # stable_diffusion_dual_unet.py here): no callbacks, no per-step ancestral noise
        stochastic = isinstance(self.scheduler, S.DDPMScheduler) or (isinstance(self.scheduler, S.DDIMScheduler) and extra["eta"] > 0)
        loop_graph_ok = self.use_cuda_graph and self.use_loop_graph and callback_on_step_end is None and not stochastic
        if loop_graph_ok:
            run = lambda: self.unet.forward(ws["unet_in"], ws["temb"], ws["kv"], out=ws["eps"], cfg_shared=do_cfg)
        else:
            run = self._unet_runner(("gm1", B, h, w, do_cfg), self.unet, ws["unet_in"], ws["temb"], ws["kv"], ws["eps"], cfg_shared=do_cfg)
        eps_u = ws["eps"][:B].reshape(-1, 4) if do_cfg else None
        eps_c = (ws["eps"][B:] if do_cfg else ws["eps"]).reshape(-1, 4)

        def denoise_loop():
            for i, t in enumerate(ts):  # gm.py:1040-1091
                if self.interrupt:
                    continue
                ws["temb"].copy_(table[i:i + 1])
                run()
                plan = self.scheduler.plan_step(t, extra["eta"])
                if plan.needs_noise:
                    z = torch.randn((B, 4, h, w), generator=generator, device=device if generator is None or generator.device.type == "cuda" else "cpu").to(device)
                    st.noise = torch.empty(n_px, 4, dtype=torch.float32, device=device)
                    L.check(L.lib().gmd_latents_nchw_to_px(z.contiguous().data_ptr(), st.noise.data_ptr(), B, h * w, stream))
                S.fused_step(plan, st, eps_c, eps_u, guidance_scale=guidance_scale, guidance_rescale=guidance_rescale if do_cfg else 0.0,
                             px_per_sample=h * w, x0_coeffs=self.scheduler.x0_coeffs(t), concat_out=ws["unet_in"], concat_lead=ws["sdr_px"],
                             concat_self=True, concat_dup=1, rescale_ws=ws["rescale"])
                if callback_on_step_end is not None:  # gm.py:1073-1081
                    cb_kwargs = {}
                    for k in callback_on_step_end_tensor_inputs:
                        cb_kwargs[k] = {"latents": self._latents_nchw(st.x, B, h, w), "prompt_embeds": prompt_embeds,
                                        "negative_prompt_embeds": negative_prompt_embeds}[k]
                    cb_out = callback_on_step_end(self, i, t, cb_kwargs) or {}
                    if "latents" in cb_out:
                        new = cb_out["latents"].to(device=device, dtype=torch.float32).contiguous()
                        L.check(L.lib().gmd_latents_nchw_to_px(new.data_ptr(), st.x.data_ptr(), B, h * w, stream))
                        L.check(L.lib().gmd_pack_unet_input(ws["sdr_px"].data_ptr(), st.x.data_ptr(), ws["unet_in"].data_ptr(), n_px, 8, stream))

        if not loop_graph_ok:
            denoise_loop()
        else:
            loop_key = (B, h, w, do_cfg, type(self.scheduler).__name__, tuple(ts), float(guidance_scale), float(guidance_rescale if do_cfg else 0.0),
                        table.data_ptr(), tuple(t_.data_ptr() for t_ in ws["kv"]))
            ent = self._loop_graphs.get(loop_key)
            if ent is None:
                ws["temb"].copy_(table[0:1])
                run()                              # eager warm-up: lazily allocated scratch exists before the capture; eps is rewritten
                torch.cuda.synchronize()
                n0 = L.launch_count()
                lg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(lg):
                    denoise_loop()
                n_kernels = L.launch_count() - n0
                L.lib().gmd_add_launch_count(-n_kernels)   # (the capture pass recorded them; they did not run)
                if len(self._loop_graphs) >= 4:
                    self._loop_graphs.clear()
                self._loop_graphs[loop_key] = ent = (lg, n_kernels, table)   # (the table must outlive its cache entry)
            ent[0].replay()
            self.graph_launches += ent[1]
        image = self._latents_nchw(st.x, B, h, w).to(out_dtype)
        if output_type != "latent":
            if self.vae is None:
                raise ValueError(f"output_type={output_type!r} needs a VAE; pass output_type='latent'")
            img = self.vae.decode_px(st.x, B, h, w)  # gm.py:1093-1096
            x = (img.float() / 2 + 0.5).clamp(0, 1)
            if output_type == "pt":
                image = x.permute(0, 3, 1, 2)
            elif output_type == "np":
                image = x.cpu().numpy()
            elif output_type == "pil":
                from PIL import Image
                image = [Image.fromarray(a) for a in (x.cpu().numpy() * 255).round().astype("uint8")]
            else:
                raise ValueError(f"unknown output_type {output_type!r}")
        if not return_dict:
            return (image, None)
        return StableDiffusionPipelineOutput(images=image, nsfw_content_detected=None)

    def _latents_nchw(self, x_px, B, h, w):
        out = torch.empty((B, 4, h, w), dtype=torch.float32, device=self.device)
        L.check(L.lib().gmd_latents_px_to_nchw(x_px.data_ptr(), out.data_ptr(), B, h * w, L.current_stream()))
        return out
