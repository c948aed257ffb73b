"""Oracle for the loops: line-by-line CPU restatement of the reference pipelines' denoising loops
(stable_diffusion_dual_unet.py:1002-1012,1033-1037,1040-1093,1117-1132 and
stable_diffusion_gm.py:1003-1015,1037-1071) and of the callers' post-processing
(scripts/inference/generate_hdr.py:225-233,256-265).  TEST INFRASTRUCTURE ONLY.

The pipelines themselves cannot be imported here (every import block pulls diffusers,
stable_diffusion_dual_unet.py:22-42) -> PARITY UNPINNED for the loop; quirks of SURVEY.md §8a-Q are kept.
"""
from __future__ import annotations

import copy

import torch

from . import tone_mapping_oracle as tm
from .schedulers_oracle import rescale_noise_cfg


@torch.no_grad()
def dual_unet_loop(unet, gm_unet, scheduler, prompt_embeds, negative_prompt_embeds, latents, num_inference_steps=50,
                   guidance_scale=7.5, guidance_rescale=0.0, eta=0.0, generator=None, trace=None):
    """Returns (latents, gm_latents) like output_type='latent' (dual_unet.py:1122-1132)."""
    do_cfg = guidance_scale > 1  # dual_unet.py:767-768 (time_cond_proj_dim is None for SD1.5)
    if do_cfg:
        embeds = torch.cat([negative_prompt_embeds, prompt_embeds])  # :983-984
    else:
        embeds = prompt_embeds
    scheduler.set_timesteps(num_inference_steps)  # :996-998
    timesteps = scheduler.timesteps
    latents = latents * scheduler.init_noise_sigma  # :717
    gm_latents = latents.clone()  # :1012  (same initial noise)
    gm_scheduler = copy.deepcopy(scheduler)  # :1036-1037
    # batch-correct conditional slice (visualize_latents.py:274); identical to prompt_embeds[1:] at B=1 (:1086)
    gm_embeds = embeds[negative_prompt_embeds.shape[0]:] if do_cfg else embeds
    extra = {}
    if "eta" in scheduler.step.__code__.co_varnames:
        extra["eta"] = eta
    if "generator" in scheduler.step.__code__.co_varnames:
        extra["generator"] = generator
    for i, t in enumerate(timesteps):
        latents_in = latents
        latent_model_input = torch.cat([latents] * 2) if do_cfg else latents  # :1045
        latent_model_input = scheduler.scale_model_input(latent_model_input, t)  # :1047
        gm_latents = gm_scheduler.scale_model_input(gm_latents, t)  # :1048
        sdr_noise_pred = unet(latent_model_input, t, encoder_hidden_states=embeds)  # :1052-1060
        raw_sdr = sdr_noise_pred
        if do_cfg:
            u, c = sdr_noise_pred.chunk(2)  # :1064
            sdr_noise_pred = u + guidance_scale * (c - u)  # :1065
            if guidance_rescale > 0.0:
                sdr_noise_pred = rescale_noise_cfg(sdr_noise_pred, c, guidance_rescale)  # :1067-1069
        a = scheduler.alphas_cumprod.to(sdr_noise_pred.device)[t].view(-1, 1, 1, 1)  # :1072
        x0_latent = (latents - (1 - a).sqrt() * sdr_noise_pred) / a.sqrt()  # :1073-1075
        latents = scheduler.step(sdr_noise_pred, t, latents, **extra)[0]  # :1077
        gm_latent_input = torch.cat([x0_latent, gm_latents], dim=1)  # :1080
        gm_noise_pred = gm_unet(gm_latent_input, t, encoder_hidden_states=gm_embeds)  # :1083-1092 (no CFG)
        gm_latents = gm_scheduler.step(gm_noise_pred, t, gm_latents, **extra)[0]  # :1093
        if trace is not None:
            trace.append(dict(t=int(t), x_in=latents_in.clone(), sdr_raw=raw_sdr.clone(), sdr_eps=sdr_noise_pred.clone(), gm_eps=gm_noise_pred.clone(),
                              latents=latents.clone(), gm_latents=gm_latents.clone(), x0=x0_latent.clone()))
    return latents, gm_latents


@torch.no_grad()
def single_gm_loop(unet8, scheduler, sdr_latent, prompt_embeds, negative_prompt_embeds, latents, num_inference_steps=50,
                   guidance_scale=7.5, guidance_rescale=0.0, eta=0.0, generator=None, trace=None):
    """stable_diffusion_gm.py:1040-1071: 8-channel UNet on cat([sdr_latent, latents]) under CFG."""
    do_cfg = guidance_scale > 1
    embeds = torch.cat([negative_prompt_embeds, prompt_embeds]) if do_cfg else prompt_embeds
    scheduler.set_timesteps(num_inference_steps)
    latents = latents * scheduler.init_noise_sigma
    extra = {}
    if "eta" in scheduler.step.__code__.co_varnames:
        extra["eta"] = eta
    if "generator" in scheduler.step.__code__.co_varnames:
        extra["generator"] = generator
    for i, t in enumerate(scheduler.timesteps):
        x = torch.cat([sdr_latent, latents], dim=1)  # gm.py:1045
        x = torch.cat([x] * 2) if do_cfg else x  # :1047
        x = scheduler.scale_model_input(x, t)  # :1048
        noise_pred = unet8(x, t, encoder_hidden_states=embeds)  # :1051-1059
        if do_cfg:
            u, c = noise_pred.chunk(2)
            noise_pred = u + guidance_scale * (c - u)  # :1062-1064
            if guidance_rescale > 0.0:
                noise_pred = rescale_noise_cfg(noise_pred, c, guidance_rescale)
        latents = scheduler.step(noise_pred, t, latents, **extra)[0]  # :1071
        if trace is not None:
            trace.append(dict(t=int(t), eps=noise_pred.clone(), latents=latents.clone()))
    return latents


@torch.no_grad()
def decode_and_reconstruct(vae, sdr_latent, gm_latent, qmax=99.0, eps=1 / 64, clamp=False):
    """generate_hdr.py:225-233 (decode both latents, de-normalise) + :256-265 (Eq.(1), numpy twin: no clamp).
    Returns (sdr [B,3,H,W], gm [B,3,H,W], hdr [B,3,H,W]) fp32."""
    sf = vae.config["scaling_factor"]
    sdr = tm.denormalize(vae.decode(sdr_latent / sf)).float()
    gm = tm.denormalize(vae.decode(gm_latent / sf)).float()
    hdr = (torch.clamp(sdr, 0, 1) ** 2.2 + eps) * (1 + gm * qmax) - eps
    if clamp:
        hdr = torch.clamp(hdr, 0, qmax + 1)
    return sdr, gm, hdr
