"""Host side of kernel (c): scheduler objects with the attribute surface the reference pipelines touch
(`set_timesteps`, `timesteps`, `order`, `init_noise_sigma`, `scale_model_input`, `alphas_cumprod`, `config`;
stable_diffusion_dual_unet.py:151-152,717,1033,1047,1072,1077) whose `step` is not a chain of ~15 torch
launches but a *plan*: a handful of fp32 coefficients computed on the host for the fused CUDA kernel
`gmd_cfg_sched_step`.  Device state (latents, PLMS history ring, stash) lives in `BranchState`.

Coefficient arithmetic mirrors diffusers PNDMScheduler._get_prev_sample / DDIMScheduler.step op for op in
fp32 (torch CPU scalars) so the fused kernel lands within ~1e-6 of the unfused chain.
"""
from __future__ import annotations

import copy
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch

from . import _lib as L


class SchedulerConfig(dict):
    """Scheduler configuration that reads like diffusers' FrozenDict: attribute access (`config.steps_offset`), mapping access
    (`config["steps_offset"]`, `.get`, `dict(config)`) and `vars()`-free copying — so the reference scripts' post-construction swap
    `DPMSolverMultistepScheduler.from_config(pipeline.scheduler.config)` (formal_improved.py:195, rebuttal_r2q2.py:195) works with
    our classes and with the real diffusers ones."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None

    def __setattr__(self, k, v):
        self[k] = v


_SD15_DEFAULTS = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                      skip_prk_steps=True, set_alpha_to_one=False, steps_offset=1, prediction_type="epsilon",
                      clip_sample=False, timestep_spacing="leading",
                      # keys the fused step kernel does not implement: accepted only at the value the SD1.5 scheduler_config.json has
                      trained_betas=None, variance_type="fixed_small", thresholding=False, rescale_betas_zero_snr=False)


def _sd15_config(**over):
    cfg = dict(_SD15_DEFAULTS)
    cfg.update(over)
    return SchedulerConfig(cfg)


def _config_items(config) -> dict:
    """dict view of a diffusers FrozenDict / dict / namespace-like scheduler config."""
    if isinstance(config, dict):
        return dict(config)
    if hasattr(config, "items"):
        return dict(config.items())
    return dict(vars(config))


@dataclass
class StepPlan:
    """Everything kernel (c) needs for one scheduler update, as host scalars."""
    mode: int = L.SCHED_LINEAR
    plms_kind: int = 0                   # multistep combination (include/gmd_b200.h)
    n_hist: int = 0                      # history tensors read
    c_sample: float = 1.0
    c_num: float = 0.0
    c_denom: float = 1.0
    use_stash: bool = False              # PLMS counter 1: update from the stashed sample
    write_stash: bool = False            # PLMS counter 0
    push_eps: bool = True                # append the (post-CFG) eps to the history ring
    ddim: tuple = (1.0, 0.0, 1.0, 0.0, 0.0)  # sqrt_a_t, sqrt_1m_a_t, sqrt_a_prev, dir_coeff, sigma
    needs_noise: bool = False


class _SchedulerBase:
    order = 1
    init_noise_sigma = 1.0
    _uses_clip_sample = False      # DDIM / DDPM clamp the x0 prediction when clip_sample is set; PNDM / DPM-Solver have no such key
    _uses_variance_type = False    # DDPM only

    def __init__(self, **over):
        self.config = _sd15_config(**over)
        c = self.config
        if c.beta_schedule == "scaled_linear":
            betas = torch.linspace(c.beta_start ** 0.5, c.beta_end ** 0.5, c.num_train_timesteps, dtype=torch.float32) ** 2
        elif c.beta_schedule == "linear":
            betas = torch.linspace(c.beta_start, c.beta_end, c.num_train_timesteps, dtype=torch.float32)
        else:
            raise NotImplementedError(f"beta_schedule {c.beta_schedule}")
        if c.prediction_type != "epsilon":
            raise NotImplementedError("only epsilon prediction is on the reference path (SD1.5 scheduler config)")
        # Options the fused step kernel does not implement must not be dropped silently: a diffusers DDIMScheduler() / DDPMScheduler()
        # built with LIBRARY defaults has clip_sample=True and would produce different latents from the reference after conversion.
        # (The SD1.5 scheduler_config.json the reference loads — generate_hdr.py:162-176 — sets every one of these as accepted here.)
        if c.clip_sample and self._uses_clip_sample:
            raise NotImplementedError("clip_sample=True is not implemented by the fused scheduler kernel (the SD1.5 scheduler config sets "
                                      "clip_sample=false); pass clip_sample=False")
        if c.timestep_spacing != "leading":
            raise NotImplementedError(f"timestep_spacing={c.timestep_spacing!r}: only the SD1.5 config's 'leading' spacing is on the reference path")
        if c.trained_betas is not None:
            raise NotImplementedError("trained_betas is not supported (the reference uses the scaled_linear schedule)")
        if c.thresholding:
            raise NotImplementedError("dynamic thresholding is not implemented by the fused scheduler kernel")
        if c.rescale_betas_zero_snr:
            raise NotImplementedError("rescale_betas_zero_snr is not implemented")
        if self._uses_variance_type and c.variance_type != "fixed_small":
            raise NotImplementedError(f"variance_type={c.variance_type!r}: the fused DDPM step implements 'fixed_small' (the SD1.5 config)")
        self.betas = betas
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if c.set_alpha_to_one else self.alphas_cumprod[0]
        self.timesteps: Optional[torch.Tensor] = None
        self.num_inference_steps: Optional[int] = None

    @classmethod
    def from_config(cls, config, **over):
        d = _config_items(config)
        d.update(over)
        return cls(**{k: v for k, v in d.items() if k in _SD15_DEFAULTS})

    def scale_model_input(self, sample, timestep=None):
        return sample  # identity for PNDM / DDIM / DDPM (dual_unet.py:1047-1048)

    def x0_coeffs(self, t: int):
        """sqrt(alpha_t), sqrt(1 - alpha_t) for the x0 prediction at dual_unet.py:1072-1075."""
        a = self.alphas_cumprod[int(t)]
        return float(a.sqrt()), float((1 - a).sqrt())

    def reset(self):
        pass


class PNDMScheduler(_SchedulerBase):
    """PLMS (skip_prk_steps) — diffusers PNDMScheduler.step_plms as a coefficient plan."""

    def __init__(self, **over):
        super().__init__(**over)
        if not self.config.skip_prk_steps:
            raise NotImplementedError("PRK warm-up steps are not on the reference path (SD1.5 config sets skip_prk_steps)")
        self.counter = 0
        self.n_ets = 0

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = int(num_inference_steps)
        ratio = self.config.num_train_timesteps // self.num_inference_steps
        base = (np.arange(0, self.num_inference_steps) * ratio).round() + self.config.steps_offset
        plms = np.concatenate([base[:-1], base[-2:-1], base[-1:]])[::-1].copy()
        self.timesteps = torch.from_numpy(plms.astype(np.int64))
        self.reset()

    def reset(self):
        self.counter = 0
        self.n_ets = 0

    def _prev_coeffs(self, timestep: int, prev_timestep: int):
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        b_t, b_p = 1 - a_t, 1 - a_p
        sample_coeff = (a_p / a_t) ** 0.5
        denom = a_t * b_p ** 0.5 + (a_t * b_t * a_p) ** 0.5
        return float(sample_coeff), float(a_p - a_t), float(denom)

    def plan_step(self, timestep: int, eta: float = 0.0) -> StepPlan:
        timestep = int(timestep)
        ratio = self.config.num_train_timesteps // self.num_inference_steps
        prev = timestep - ratio
        plan = StepPlan()
        if self.counter != 1:
            self.n_ets = min(self.n_ets, 3) + 1
            plan.push_eps = True
        else:
            prev = timestep
            timestep = timestep + ratio
            plan.push_eps = False
        if self.n_ets == 1 and self.counter == 0:
            plan.plms_kind, plan.n_hist, plan.write_stash = 0, 0, True
        elif self.n_ets == 1 and self.counter == 1:
            plan.plms_kind, plan.n_hist, plan.use_stash = 1, 1, True
        elif self.n_ets == 2:
            plan.plms_kind, plan.n_hist = 2, 1
        elif self.n_ets == 3:
            plan.plms_kind, plan.n_hist = 3, 2
        else:
            plan.plms_kind, plan.n_hist = 4, 3
        plan.c_sample, plan.c_num, plan.c_denom = self._prev_coeffs(timestep, prev)
        self.counter += 1
        return plan


class DDIMScheduler(_SchedulerBase):
    _uses_clip_sample = True

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = int(num_inference_steps)
        ratio = self.config.num_train_timesteps // self.num_inference_steps
        ts = (np.arange(0, self.num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64) + self.config.steps_offset
        self.timesteps = torch.from_numpy(ts)

    def plan_step(self, timestep: int, eta: float = 0.0) -> StepPlan:
        timestep = int(timestep)
        prev = timestep - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        variance = ((1 - a_p) / (1 - a_t)) * (1 - a_t / a_p)
        std = eta * variance ** 0.5
        plan = StepPlan(mode=L.SCHED_DDIM, push_eps=False)
        plan.ddim = (float(a_t ** 0.5), float(b_t ** 0.5), float(a_p ** 0.5), float((1 - a_p - std ** 2) ** 0.5), float(std))
        plan.needs_noise = eta > 0
        return plan


class DDPMScheduler(_SchedulerBase):
    """diffusers DDPMScheduler (epsilon prediction, fixed_small variance, leading spacing) — the scheduler every reference CLI
    actually passes (scripts/inference/generate_hdr.py:162-176; formal_baseline.py:175-191).  Ancestral noise is drawn by the
    pipeline from the caller's generator in the reference's order (SDR branch first, then GM branch, each step)."""
    _uses_clip_sample = True
    _uses_variance_type = True

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = int(num_inference_steps)
        ratio = self.config.num_train_timesteps // self.num_inference_steps
        ts = (np.arange(0, self.num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64) + self.config.steps_offset
        self.timesteps = torch.from_numpy(ts)

    def plan_step(self, timestep: int, eta: float = 0.0) -> StepPlan:
        t = int(timestep)
        prev = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev] if prev >= 0 else torch.tensor(1.0)
        b_t, b_p = 1 - a_t, 1 - a_p
        cur_alpha = a_t / a_p
        cur_beta = 1 - cur_alpha
        c_x0 = (a_p ** 0.5 * cur_beta) / b_t
        c_xt = cur_alpha ** 0.5 * b_p / b_t
        sigma = 0.0
        if t > 0:
            variance = torch.clamp((1 - a_p) / (1 - a_t) * cur_beta, min=1e-20)
            sigma = float(variance ** 0.5)
        plan = StepPlan(mode=L.SCHED_DDPM, push_eps=False)
        plan.ddim = (float(a_t ** 0.5), float(b_t ** 0.5), float(c_x0), float(c_xt), sigma)
        plan.needs_noise = sigma != 0.0
        return plan


class DPMSolverMultistepScheduler(_SchedulerBase):
    """diffusers DPMSolverMultistepScheduler at the defaults `from_config(pndm_config)` gives it — dpmsolver++, solver_order 2,
    midpoint, lower_order_final, final_sigmas_type "zero", leading spacing — which scripts/inference/experiments/
    formal_improved.py:195 swaps into the dual pipeline.  The history ring holds x0 predictions, not eps."""

    def __init__(self, **over):
        super().__init__(**over)
        self.sigmas: Optional[torch.Tensor] = None
        self.step_index = 0
        self.lower_order_nums = 0

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = n = int(num_inference_steps)
        c = self.config
        if c.timestep_spacing != "leading":
            raise NotImplementedError("only the SD1.5 config's leading timestep spacing is on the reference path")
        step_ratio = c.num_train_timesteps // (n + 1)
        ts = (np.arange(0, n + 1) * step_ratio).round()[::-1][:-1].copy().astype(np.int64) + c.steps_offset
        sig = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0.0]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts)
        self.reset()

    def reset(self):
        self.step_index = 0
        self.lower_order_nums = 0

    @staticmethod
    def _alpha_sigma(sigma: torch.Tensor):
        alpha_t = 1 / ((sigma ** 2 + 1) ** 0.5)
        return alpha_t, sigma * alpha_t

    def plan_step(self, timestep: int, eta: float = 0.0) -> StepPlan:
        i = self.step_index
        n = len(self.timesteps)
        final_first_order = i == n - 1  # final_sigmas_type "zero" forces a first-order last step
        alpha_s0, sigma_s0 = self._alpha_sigma(self.sigmas[i])
        alpha_t, sigma_t = self._alpha_sigma(self.sigmas[i + 1])
        lambda_t = torch.log(alpha_t) - torch.log(sigma_t)
        lambda_s0 = torch.log(alpha_s0) - torch.log(sigma_s0)
        h = lambda_t - lambda_s0
        plan = StepPlan(mode=L.SCHED_DPMPP, push_eps=True)
        plan.ddim = (float(alpha_s0), float(sigma_s0), 1.0, 0.0, 0.0)
        plan.c_sample = float(sigma_t / sigma_s0)
        plan.c_num = float(alpha_t * (torch.exp(-h) - 1.0))
        if self.lower_order_nums < 1 or final_first_order:
            plan.plms_kind, plan.n_hist = 0, 0
        else:
            alpha_s1, sigma_s1 = self._alpha_sigma(self.sigmas[i - 1])
            lambda_s1 = torch.log(alpha_s1) - torch.log(sigma_s1)
            r0 = (lambda_s0 - lambda_s1) / h
            plan.plms_kind, plan.n_hist = 1, 1
            plan.c_denom = float(1.0 / r0)
        if self.lower_order_nums < 2:
            self.lower_order_nums += 1
        self.step_index += 1
        return plan


class BranchState:
    """Device state of one denoising branch (SDR or GM): fp32 pixel-major latents [B*h*w, 4], a 4-slot eps ring
    and the PLMS stash.  Allocated once; every step is in-place (CUDA-graph friendly)."""

    def __init__(self, n_px: int, device):
        f = dict(dtype=torch.float32, device=device)
        self.n_px = n_px
        self.x = torch.empty(n_px, 4, **f)
        self.stash = torch.empty(n_px, 4, **f)
        self.ring = [torch.empty(n_px, 4, **f) for _ in range(4)]
        self.order: List[int] = []  # ring slots, most recent first
        self.noise: Optional[torch.Tensor] = None

    def reset(self):
        self.order = []

    def hist(self, k: int) -> Optional[torch.Tensor]:
        return self.ring[self.order[k]] if k < len(self.order) else None

    def free_slot(self) -> int:
        used = set(self.order[:3])
        for s in range(4):
            if s not in used:
                return s
        raise AssertionError


def fused_step(plan: StepPlan, state: BranchState, eps_cond: torch.Tensor, eps_uncond: Optional[torch.Tensor] = None, *,
               guidance_scale: float = 1.0, guidance_rescale: float = 0.0, px_per_sample: int = 0, x0_coeffs=(1.0, 0.0),
               unet_in_next: Optional[torch.Tensor] = None, concat_out: Optional[torch.Tensor] = None,
               concat_tail: Optional[torch.Tensor] = None, concat_lead: Optional[torch.Tensor] = None,
               x0_out: Optional[torch.Tensor] = None, unet_in_dup: int = 1, concat_dup: int = 1, concat_self: bool = False, rescale_ws: Optional[torch.Tensor] = None, stream: Optional[int] = None) -> None:
    """One launch of `gmd_cfg_sched_step` (csrc/sched.cu) for one branch; updates `state` in place."""
    p = L.SchedParams()
    p.eps_uncond, p.eps_cond = L.ptr(eps_uncond), eps_cond.data_ptr()
    p.x, p.x_next = state.x.data_ptr(), state.x.data_ptr()
    p.x_stash = state.stash.data_ptr() if plan.use_stash else None
    p.stash_out = state.stash.data_ptr() if plan.write_stash else None
    for k in range(3):
        h = state.hist(k) if k < plan.n_hist else None
        p.hist[k] = L.ptr(h)
    slot = None
    if plan.push_eps and plan.mode in (L.SCHED_LINEAR, L.SCHED_DPMPP):
        slot = state.free_slot()
        p.eps_out = state.ring[slot].data_ptr()
    if plan.needs_noise:
        if state.noise is None:
            raise ValueError("this scheduler step needs pre-drawn variance noise in BranchState.noise")
        p.noise = state.noise.data_ptr()
    p.unet_in_next, p.concat_out = L.ptr(unet_in_next), L.ptr(concat_out)
    p.concat_tail, p.concat_lead, p.x0_out = L.ptr(concat_tail), L.ptr(concat_lead), L.ptr(x0_out)
    ch = None
    for t in (unet_in_next, concat_out):
        if t is not None:
            ch = t.shape[-1]
    p.unet_in_ch = ch or 8
    p.unet_in_dup, p.concat_dup, p.concat_self = unet_in_dup, concat_dup, int(concat_self)
    p.n_px, p.px_per_sample = state.n_px, px_per_sample or state.n_px
    p.mode, p.use_stash = plan.mode, int(plan.use_stash)
    p.guidance_scale, p.guidance_rescale = float(guidance_scale), float(guidance_rescale)
    p.rescale_stats = L.ptr(rescale_ws)
    p.sqrt_alpha_t, p.sqrt_1m_alpha_t = x0_coeffs
    p.plms_kind = plan.plms_kind
    p.c_sample, p.c_num, p.c_denom = plan.c_sample, plan.c_num, plan.c_denom
    (p.ddim_sqrt_alpha_t, p.ddim_sqrt_1m_alpha_t, p.ddim_sqrt_alpha_prev, p.ddim_dir_coeff, p.ddim_sigma) = plan.ddim
    L.check(L.lib().gmd_cfg_sched_step(C.byref(p), L.current_stream() if stream is None else stream), "gmd_cfg_sched_step")
    if slot is not None:
        state.order.insert(0, slot)
        del state.order[3:]


def clone_scheduler(s):
    """`copy.deepcopy(self.scheduler)` at dual_unet.py:1036-1037."""
    return copy.deepcopy(s)
