"""Shared host logic of the three pipelines: argument checking, prompt encoding, timestep retrieval, latent
preparation and the B200 denoising engine.  Function names and error behaviour follow
gm_diffusion/pipelines/stable_diffusion_dual_unet.py (cited per function)."""
from __future__ import annotations

import inspect
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Union

import torch

from .. import _lib as L
from .. import schedulers as S
from ..unet import B200UNet
from ..vae import B200Vae, B200VaeDecoder

bf16 = torch.bfloat16


@dataclass
class StableDiffusionPipelineOutput:
    """Stand-in for diffusers' output class (stable_diffusion_gm.py:1111-1114)."""
    images: Any
    nsfw_content_detected: Optional[List[bool]] = None

    def __getitem__(self, i):
        return (self.images, self.nsfw_content_detected)[i]


def retrieve_timesteps(scheduler, num_inference_steps=None, device=None, timesteps=None, sigmas=None, **kwargs):
    """stable_diffusion_dual_unet.py:97-153."""
    if timesteps is not None and sigmas is not None:
        raise ValueError("Only one of `timesteps` or `sigmas` can be passed. Please choose one to set custom values")
    if timesteps is not None:
        if "timesteps" not in set(inspect.signature(scheduler.set_timesteps).parameters.keys()):
            raise ValueError(
                f"The current scheduler class {scheduler.__class__}'s `set_timesteps` does not support custom"
                f" timestep schedules. Please check whether you are using the correct scheduler.")
        scheduler.set_timesteps(timesteps=timesteps, device=device, **kwargs)
        timesteps = scheduler.timesteps
        num_inference_steps = len(timesteps)
    elif sigmas is not None:
        if "sigmas" not in set(inspect.signature(scheduler.set_timesteps).parameters.keys()):
            raise ValueError(
                f"The current scheduler class {scheduler.__class__}'s `set_timesteps` does not support custom"
                f" sigmas schedules. Please check whether you are using the correct scheduler.")
        scheduler.set_timesteps(sigmas=sigmas, device=device, **kwargs)
        timesteps = scheduler.timesteps
        num_inference_steps = len(timesteps)
    else:
        scheduler.set_timesteps(num_inference_steps, device=device, **kwargs)
        timesteps = scheduler.timesteps
    return timesteps, num_inference_steps


def as_b200_unet(unet, device) -> B200UNet:
    if isinstance(unet, B200UNet):
        return unet
    if hasattr(unet, "state_dict"):
        return B200UNet.from_module(unet, device=device)
    raise TypeError(f"unet must be a B200UNet or expose a diffusers-style state_dict(), got {type(unet)}")


def as_b200_vae(vae, device) -> Optional[B200VaeDecoder]:
    if vae is None or isinstance(vae, B200VaeDecoder):
        return vae
    if hasattr(vae, "state_dict"):
        has_encoder = any(k.startswith("encoder.") for k in vae.state_dict())
        return (B200Vae if has_encoder else B200VaeDecoder).from_module(vae, device=device)
    raise TypeError(f"vae must be a B200VaeDecoder or expose a diffusers-style state_dict(), got {type(vae)}")


def as_b200_text_encoder(text_encoder, device):
    """transformers' CLIPTextModel (the only text encoder the reference pipelines are built with, dual_unet.py:19,211) -> the same
    network on this library's kernels.  Anything else (None, an already converted encoder, a caller's own
    callable) is kept as it is."""
    from ..text_encoder import B200ClipTextEncoder
    if text_encoder is None or isinstance(text_encoder, B200ClipTextEncoder):
        return text_encoder
    if type(text_encoder).__name__ == "CLIPTextModel" and hasattr(text_encoder, "state_dict"):
        return B200ClipTextEncoder.from_module(text_encoder, device=device)
    return text_encoder


def as_b200_scheduler(scheduler):
    """Accept our schedulers, or any object with a diffusers scheduler `config` whose class name we support."""
    if isinstance(scheduler, (S.PNDMScheduler, S.DDIMScheduler, S.DDPMScheduler, S.DPMSolverMultistepScheduler)):
        return scheduler
    name = type(scheduler).__name__
    cfg = getattr(scheduler, "config", None)
    if cfg is not None and "PNDM" in name:
        return S.PNDMScheduler.from_config(cfg)
    if cfg is not None and "DDIM" in name:
        return S.DDIMScheduler.from_config(cfg)
    if cfg is not None and "DDPM" in name:
        return S.DDPMScheduler.from_config(cfg)
    if cfg is not None and "DPMSolverMultistep" in name:
        get = cfg.get if hasattr(cfg, "get") else lambda k, d=None: getattr(cfg, k, d)
        if (get("algorithm_type", "dpmsolver++") != "dpmsolver++" or get("solver_order", 2) != 2 or get("solver_type", "midpoint") != "midpoint"
                or get("use_karras_sigmas", False) or get("final_sigmas_type", "zero") != "zero" or get("thresholding", False)):
            raise NotImplementedError("only DPM-Solver++(2M, midpoint, final sigma zero) — the from_config defaults the reference uses — is accelerated")
        return S.DPMSolverMultistepScheduler.from_config(cfg)
    raise NotImplementedError(
        f"scheduler {name} is not accelerated: the fused step kernel implements PNDM (PLMS), DDIM, DDPM and DPM-Solver++(2M)")


class PipelineBase:
    """Common constructor / helpers (stable_diffusion_dual_unet.py:202-334)."""

    _callback_tensor_inputs = ["latents", "prompt_embeds", "negative_prompt_embeds"]
    _optional_components = ["safety_checker", "feature_extractor", "image_encoder"]

    def _init_common(self, vae, text_encoder, tokenizer, scheduler, safety_checker, feature_extractor, image_encoder,
                     requires_safety_checker, device):
        if not torch.cuda.is_available():
            raise RuntimeError("gm_diffusion_b200 pipelines need a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device)
        self._execution_device = self.device
        self.vae = as_b200_vae(vae, self.device)
        self.text_encoder = text_encoder
        self.tokenizer = tokenizer
        self.scheduler = scheduler   # (property: converts diffusers-style schedulers, also when assigned after construction)
        self.safety_checker = None  # accepted at the signature level, never run
        self.feature_extractor = feature_extractor
        self.image_encoder = image_encoder
        self.vae_scale_factor = 8  # 2 ** (len(vae.config.block_out_channels) - 1), dual_unet.py:300
        self._guidance_scale = 7.5
        self._guidance_rescale = 0.0
        self._clip_skip = None
        self._cross_attention_kwargs = None
        self._interrupt = False
        self._num_timesteps = 0
        self.use_cuda_graph = True
        self.skip_identical_cfg = True  # drop the uncond forward when negative_prompt_embeds == prompt_embeds (bit-exact)
        self._graphs: Dict[Any, Any] = {}
        self.graph_launches = 0  # kernels executed through CUDA-graph replays (the C-ABI counter only sees eager launches)

    # The reference scripts swap the scheduler AFTER construction (`pipeline.scheduler = DPMSolverMultistepScheduler.from_config(
    # pipeline.scheduler.config)`: formal_improved.py:195, rebuttal_r2q2.py:195, formal_improved_ablation.py:195, rebuttal_visual.py:270):
    # every assignment goes through the same conversion as the constructor argument.
    @property
    def text_encoder(self):
        return self._text_encoder

    @text_encoder.setter
    def text_encoder(self, value):
        self._text_encoder = as_b200_text_encoder(value, getattr(self, "device", "cuda"))
        self.__dict__.pop("_text_cache", None)

    @property
    def scheduler(self):
        return self._scheduler

    @scheduler.setter
    def scheduler(self, value):
        self._scheduler = as_b200_scheduler(value)

    # properties, dual_unet.py:751-780
    @property
    def guidance_scale(self):
        return self._guidance_scale

    @property
    def guidance_rescale(self):
        return self._guidance_rescale

    @property
    def clip_skip(self):
        return self._clip_skip

    @property
    def do_classifier_free_guidance(self):
        return self._guidance_scale > 1  # time_cond_proj_dim is None for the SD1.5 config (dual_unet.py:767-768)

    @property
    def cross_attention_kwargs(self):
        return self._cross_attention_kwargs

    @property
    def num_timesteps(self):
        return self._num_timesteps

    @property
    def interrupt(self):
        return self._interrupt

    def to(self, device):
        if torch.device(device).type != "cuda":
            raise RuntimeError("gm_diffusion_b200 pipelines are CUDA-only")
        return self

    def progress_bar(self, total):
        class _Bar:
            def __enter__(self_inner):
                return self_inner

            def __exit__(self_inner, *a):
                return False

            def update(self_inner, n=1):
                pass
        return _Bar()

    # ---- dual_unet.py:629-696 -------------------------------------------------------------------------------
    def check_inputs(self, prompt, height, width, callback_steps, negative_prompt=None, prompt_embeds=None,
                     negative_prompt_embeds=None, ip_adapter_image=None, ip_adapter_image_embeds=None,
                     callback_on_step_end_tensor_inputs=None):
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        if callback_steps is not None and (not isinstance(callback_steps, int) or callback_steps <= 0):
            raise ValueError(f"`callback_steps` has to be a positive integer but is {callback_steps} of type {type(callback_steps)}.")
        if callback_on_step_end_tensor_inputs is not None and not all(
                k in self._callback_tensor_inputs for k in callback_on_step_end_tensor_inputs):
            raise ValueError(
                f"`callback_on_step_end_tensor_inputs` has to be in {self._callback_tensor_inputs}, but found "
                f"{[k for k in callback_on_step_end_tensor_inputs if k not in self._callback_tensor_inputs]}")
        if prompt is not None and prompt_embeds is not None:
            raise ValueError(f"Cannot forward both `prompt`: {prompt} and `prompt_embeds`: {prompt_embeds}. Please make sure to"
                             " only forward one of the two.")
        elif prompt is None and prompt_embeds is None:
            raise ValueError("Provide either `prompt` or `prompt_embeds`. Cannot leave both `prompt` and `prompt_embeds` undefined.")
        elif prompt is not None and (not isinstance(prompt, str) and not isinstance(prompt, list)):
            raise ValueError(f"`prompt` has to be of type `str` or `list` but is {type(prompt)}")
        if negative_prompt is not None and negative_prompt_embeds is not None:
            raise ValueError(f"Cannot forward both `negative_prompt`: {negative_prompt} and `negative_prompt_embeds`:"
                             f" {negative_prompt_embeds}. Please make sure to only forward one of the two.")
        if prompt_embeds is not None and negative_prompt_embeds is not None:
            if prompt_embeds.shape != negative_prompt_embeds.shape:
                raise ValueError("`prompt_embeds` and `negative_prompt_embeds` must have the same shape when passed directly, but"
                                 f" got: `prompt_embeds` {prompt_embeds.shape} != `negative_prompt_embeds`"
                                 f" {negative_prompt_embeds.shape}.")
        if ip_adapter_image is not None or ip_adapter_image_embeds is not None:
            raise NotImplementedError("IP-Adapter inputs are accepted at the signature level only (SURVEY.md §8b)")

    # ---- dual_unet.py:336-516 -------------------------------------------------------------------------------
    def encode_prompt(self, prompt, device, num_images_per_prompt, do_classifier_free_guidance, negative_prompt=None,
                      prompt_embeds=None, negative_prompt_embeds=None, lora_scale=None, clip_skip=None):
        # lora_scale: no LoRA layers exist on the text encoder handed to this pipeline unless the caller loaded an adapter; the
        # reference's scale_lora_layers/unscale_lora_layers pair (dual_unet.py:379-386,511-514) is then a no-op and is not restated.
        if prompt is not None and isinstance(prompt, str):
            batch_size = 1
        elif prompt is not None and isinstance(prompt, list):
            batch_size = len(prompt)
        else:
            batch_size = prompt_embeds.shape[0]

        def _encode(texts):
            if self.tokenizer is None or self.text_encoder is None:
                raise ValueError("string prompts need `tokenizer` and `text_encoder`; pass `prompt_embeds` / "
                                 "`negative_prompt_embeds` instead")
            # the text encoder depends on the strings only — the default negative prompt "" (and any repeated prompt,
            # e.g. the SDR->HDR CLI's prompt=[""], generate_hdr.py:212-218) is encoded once per (text, clip_skip) and reused
            cache = self.__dict__.setdefault("_text_cache", {})
            keys = [(t_, clip_skip, id(self.text_encoder)) for t_ in texts]
            if all(k_ in cache for k_ in keys):
                return torch.cat([cache[k_] for k_ in keys], 0)
            out = _encode_uncached(texts)
            if len(cache) > 256:
                cache.clear()
            for k_, row in zip(keys, out):
                cache[k_] = row[None].detach()
            return out

        def _encode_uncached(texts):
            ids = self.tokenizer(texts, padding="max_length", max_length=self.tokenizer.model_max_length, truncation=True,
                                 return_tensors="pt").input_ids
            enc_dev = next(self.text_encoder.parameters()).device
            if clip_skip is None:
                out = self.text_encoder(ids.to(enc_dev))[0]
            else:
                o = self.text_encoder(ids.to(enc_dev), output_hidden_states=True)
                out = self.text_encoder.text_model.final_layer_norm(o[-1][-(clip_skip + 1)])
            return out

        if prompt_embeds is None:
            prompt_embeds = _encode([prompt] if isinstance(prompt, str) else prompt)
        prompt_embeds = prompt_embeds.to(device=device)
        bs, seq, _ = prompt_embeds.shape
        prompt_embeds = prompt_embeds.repeat(1, num_images_per_prompt, 1).view(bs * num_images_per_prompt, seq, -1)
        if do_classifier_free_guidance and negative_prompt_embeds is None:
            if negative_prompt is None:
                uncond = [""] * batch_size
            elif prompt is not None and type(prompt) is not type(negative_prompt):
                raise TypeError(f"`negative_prompt` should be the same type to `prompt`, but got {type(negative_prompt)} !="
                                f" {type(prompt)}.")
            elif isinstance(negative_prompt, str):
                uncond = [negative_prompt]
            elif batch_size != len(negative_prompt):
                raise ValueError(f"`negative_prompt`: {negative_prompt} has batch size {len(negative_prompt)}, but `prompt`:"
                                 f" {prompt} has batch size {batch_size}. Please make sure that passed `negative_prompt` matches"
                                 " the batch size of `prompt`.")
            else:
                uncond = negative_prompt
            negative_prompt_embeds = _encode(uncond)
        if do_classifier_free_guidance:
            negative_prompt_embeds = negative_prompt_embeds.to(device=device)
            seq = negative_prompt_embeds.shape[1]
            negative_prompt_embeds = negative_prompt_embeds.repeat(1, num_images_per_prompt, 1).view(
                batch_size * num_images_per_prompt, seq, -1)
        return prompt_embeds, negative_prompt_embeds

    # ---- dual_unet.py:698-718 -------------------------------------------------------------------------------
    def prepare_latents(self, batch_size, num_channels_latents, height, width, dtype, device, generator, latents=None):
        shape = (batch_size, num_channels_latents, int(height) // self.vae_scale_factor, int(width) // self.vae_scale_factor)
        if isinstance(generator, list) and len(generator) != batch_size:
            raise ValueError(f"You have passed a list of generators of length {len(generator)}, but requested an effective batch"
                             f" size of {batch_size}. Make sure the batch size matches the length of the generators.")
        if latents is None:
            latents = randn_tensor(shape, generator=generator, device=device, dtype=dtype)
        else:
            latents = latents.to(device)
        return latents * self.scheduler.init_noise_sigma

    def prepare_extra_step_kwargs(self, generator, eta):
        """dual_unet.py:612-627: eta only reaches schedulers that accept it (DDIM)."""
        return {"eta": eta if isinstance(self.scheduler, S.DDIMScheduler) else 0.0, "generator": generator}

    # ---- CUDA-graph cache for one UNet at one shape ------------------------------------------------------------
    def _unet_runner(self, key, unet: B200UNet, sample, temb_row, ctx_kv, eps_out, cfg_shared: bool = False):
        """Returns a zero-arg callable running unet.forward on the given STATIC buffers; captured into a CUDA graph
        on first use (the forward is ~900 launches; replaying it removes the Python/ctypes launch overhead)."""
        if not self.use_cuda_graph:
            return lambda: unet.forward(sample, temb_row, ctx_kv, out=eps_out, cfg_shared=cfg_shared)
        def make_replay(g, n_kernels):
            def replay():
                g.replay()
                self.graph_launches += n_kernels
            return replay

        ent = self._graphs.get(key)
        if ent is not None:
            g, bufs, n_kernels = ent
            same = all(a.data_ptr() == b for a, b in zip([sample, temb_row, eps_out] + list(ctx_kv), bufs))
            if same:
                return make_replay(g, n_kernels)
        n0 = L.launch_count()
        unet.forward(sample, temb_row, ctx_kv, out=eps_out, cfg_shared=cfg_shared)  # warm-up: attribute setup, allocator pools
        n_kernels = L.launch_count() - n0
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            unet.forward(sample, temb_row, ctx_kv, out=eps_out, cfg_shared=cfg_shared)
        self._graphs[key] = (g, [t.data_ptr() for t in [sample, temb_row, eps_out] + list(ctx_kv)], n_kernels)
        return make_replay(g, n_kernels)


def randn_tensor(shape, generator=None, device=None, dtype=None):
    """diffusers.utils.torch_utils.randn_tensor: CPU generators draw on the CPU and move (reproducible across devices)."""
    device = torch.device(device) if device is not None else torch.device("cpu")
    if isinstance(generator, list):
        shape1 = (1,) + tuple(shape[1:])
        return torch.cat([randn_tensor(shape1, g, device, dtype) for g in generator], 0)
    gdev = generator.device if generator is not None else device
    rand_device = torch.device("cpu") if (gdev.type == "cpu" and device.type != "cpu") else device
    return torch.randn(shape, generator=generator, device=rand_device, dtype=dtype).to(device)
