"""Oracle for the scheduler half of kernel (c): CPU restatement of the diffusers schedulers the reference
pipelines call (`scheduler.set_timesteps` at stable_diffusion_dual_unet.py:151-152, `scheduler.step` at
:1077/:1093, `alphas_cumprod` at :1072).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: diffusers (>=0.33, README.md:54) is not vendored in /root/reference and cannot be
installed offline; the formulas below restate the published PNDMScheduler.step_plms /
DDIMScheduler.step / DDPMScheduler.step / DPMSolverMultistepScheduler.step algorithms with the SD1.5 scheduler config
and are pinned by invariants in tests/test_schedulers_oracle.py.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch


def sd15_scheduler_config(**over):
    cfg = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
               skip_prk_steps=True, set_alpha_to_one=False, steps_offset=1, prediction_type="epsilon",
               clip_sample=False, timestep_spacing="leading")
    cfg.update(over)
    return SimpleNamespace(**cfg)


def make_alphas_cumprod(cfg) -> torch.Tensor:
    if cfg.beta_schedule == "scaled_linear":
        betas = torch.linspace(cfg.beta_start ** 0.5, cfg.beta_end ** 0.5, cfg.num_train_timesteps, dtype=torch.float32) ** 2
    elif cfg.beta_schedule == "linear":
        betas = torch.linspace(cfg.beta_start, cfg.beta_end, cfg.num_train_timesteps, dtype=torch.float32)
    else:
        raise NotImplementedError(cfg.beta_schedule)
    return torch.cumprod(1.0 - betas, dim=0)


class _Base:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, **over):
        self.config = sd15_scheduler_config(**over)
        self.alphas_cumprod = make_alphas_cumprod(self.config)
        self.final_alpha_cumprod = torch.tensor(1.0) if self.config.set_alpha_to_one else self.alphas_cumprod[0]
        self.timesteps = None
        self.num_inference_steps = None

    def scale_model_input(self, sample, timestep=None):
        return sample


class PNDMOracle(_Base):
    """diffusers PNDMScheduler with skip_prk_steps (PLMS only)."""

    def __init__(self, **over):
        super().__init__(**over)
        self.ets = []
        self.counter = 0
        self.cur_sample = None

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        base = (np.arange(0, num_inference_steps) * ratio).round() + self.config.steps_offset
        assert self.config.skip_prk_steps, "PRK steps are not on the reference path (SD1.5 config skips them)"
        plms = np.concatenate([base[:-1], base[-2:-1], base[-1:]])[::-1].copy()
        self.timesteps = torch.from_numpy(plms.astype(np.int64))
        self.ets = []
        self.counter = 0
        self.cur_sample = None

    def _get_prev_sample(self, sample, timestep, prev_timestep, model_output):
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        b_p = 1 - a_p
        sample_coeff = (a_p / a_t) ** 0.5
        denom = a_t * b_p ** 0.5 + (a_t * b_t * a_p) ** 0.5
        return sample_coeff * sample - (a_p - a_t) * model_output / denom

    def step(self, model_output, timestep, sample, return_dict=False, **kw):
        timestep = int(timestep)
        ratio = self.config.num_train_timesteps // self.num_inference_steps
        prev_timestep = timestep - ratio
        if self.counter != 1:
            self.ets = self.ets[-3:]
            self.ets.append(model_output)
        else:
            prev_timestep = timestep
            timestep = timestep + ratio
        if len(self.ets) == 1 and self.counter == 0:
            self.cur_sample = sample
        elif len(self.ets) == 1 and self.counter == 1:
            model_output = (model_output + self.ets[-1]) / 2
            sample = self.cur_sample
            self.cur_sample = None
        elif len(self.ets) == 2:
            model_output = (3 * self.ets[-1] - self.ets[-2]) / 2
        elif len(self.ets) == 3:
            model_output = (23 * self.ets[-1] - 16 * self.ets[-2] + 5 * self.ets[-3]) / 12
        else:
            model_output = (1 / 24) * (55 * self.ets[-1] - 59 * self.ets[-2] + 37 * self.ets[-3] - 9 * self.ets[-4])
        prev = self._get_prev_sample(sample, timestep, prev_timestep, model_output)
        self.counter += 1
        return (prev,)


class DDIMOracle(_Base):
    """diffusers DDIMScheduler (epsilon prediction, clip_sample False, leading spacing)."""

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64) + self.config.steps_offset
        self.timesteps = torch.from_numpy(ts)

    def step(self, model_output, timestep, sample, eta=0.0, generator=None, variance_noise=None, return_dict=False, **kw):
        timestep = int(timestep)
        prev_timestep = timestep - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        pred_x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        variance = ((1 - a_p) / (1 - a_t)) * (1 - a_t / a_p)
        std = eta * variance ** 0.5
        direction = (1 - a_p - std ** 2) ** 0.5 * model_output
        prev = a_p ** 0.5 * pred_x0 + direction
        if eta > 0:
            if variance_noise is None:
                variance_noise = _randn_like(model_output, generator)
            prev = prev + std * variance_noise
        return (prev,)


class DDPMOracle(_Base):
    """diffusers DDPMScheduler (fixed_small variance, epsilon prediction) — what the inference CLIs pass
    (scripts/inference/generate_hdr.py:162-176).  Ancestral noise comes from the shared generator."""

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64) + self.config.steps_offset
        self.timesteps = torch.from_numpy(ts)

    def step(self, model_output, timestep, sample, generator=None, variance_noise=None, return_dict=False, **kw):
        t = int(timestep)
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev_t] if prev_t >= 0 else torch.tensor(1.0)
        b_t, b_p = 1 - a_t, 1 - a_p
        cur_alpha = a_t / a_p
        cur_beta = 1 - cur_alpha
        pred_x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        c_x0 = (a_p ** 0.5 * cur_beta) / b_t
        c_xt = cur_alpha ** 0.5 * b_p / b_t
        prev = c_x0 * pred_x0 + c_xt * sample
        if t > 0:
            if variance_noise is None:
                variance_noise = _randn_like(model_output, generator)
            variance = torch.clamp((1 - a_p) / (1 - a_t) * cur_beta, min=1e-20)
            prev = prev + variance ** 0.5 * variance_noise
        return (prev,)


class DPMSolverOracle(_Base):
    """diffusers DPMSolverMultistepScheduler as `from_config(<SD1.5 PNDM config>)` builds it (formal_improved.py:195):
    algorithm_type dpmsolver++, solver_order 2, solver_type midpoint, lower_order_final True, final_sigmas_type zero,
    no karras sigmas, no thresholding, epsilon prediction, leading spacing.  Written tensor-op for tensor-op."""

    solver_order = 2

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        last_timestep = self.config.num_train_timesteps  # lambda_min_clipped = -inf clips nothing
        step_ratio = last_timestep // (num_inference_steps + 1)
        ts = (np.arange(0, num_inference_steps + 1) * step_ratio).round()[::-1][:-1].copy().astype(np.int64) + self.config.steps_offset
        sigmas = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        sigmas = np.interp(ts, np.arange(0, len(sigmas)), sigmas)
        sigmas = np.concatenate([sigmas, [0.0]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sigmas)
        self.timesteps = torch.from_numpy(ts)
        self.model_outputs = [None] * self.solver_order
        self.lower_order_nums = 0
        self.step_index = 0

    @staticmethod
    def _sigma_to_alpha_sigma_t(sigma):
        alpha_t = 1 / ((sigma ** 2 + 1) ** 0.5)
        return alpha_t, sigma * alpha_t

    def convert_model_output(self, model_output, sample):
        alpha_t, sigma_t = self._sigma_to_alpha_sigma_t(self.sigmas[self.step_index])
        return (sample - sigma_t * model_output) / alpha_t

    def _first_order(self, m0, sample):
        sigma_t, sigma_s = self.sigmas[self.step_index + 1], self.sigmas[self.step_index]
        alpha_t, sigma_t = self._sigma_to_alpha_sigma_t(sigma_t)
        alpha_s, sigma_s = self._sigma_to_alpha_sigma_t(sigma_s)
        h = (torch.log(alpha_t) - torch.log(sigma_t)) - (torch.log(alpha_s) - torch.log(sigma_s))
        return (sigma_t / sigma_s) * sample - (alpha_t * (torch.exp(-h) - 1.0)) * m0

    def _second_order(self, outputs, sample):
        sigma_t, sigma_s0, sigma_s1 = self.sigmas[self.step_index + 1], self.sigmas[self.step_index], self.sigmas[self.step_index - 1]
        alpha_t, sigma_t = self._sigma_to_alpha_sigma_t(sigma_t)
        alpha_s0, sigma_s0 = self._sigma_to_alpha_sigma_t(sigma_s0)
        alpha_s1, sigma_s1 = self._sigma_to_alpha_sigma_t(sigma_s1)
        lambda_t = torch.log(alpha_t) - torch.log(sigma_t)
        lambda_s0 = torch.log(alpha_s0) - torch.log(sigma_s0)
        lambda_s1 = torch.log(alpha_s1) - torch.log(sigma_s1)
        m0, m1 = outputs[-1], outputs[-2]
        h, h_0 = lambda_t - lambda_s0, lambda_s0 - lambda_s1
        r0 = h_0 / h
        D0, D1 = m0, (1.0 / r0) * (m0 - m1)
        return ((sigma_t / sigma_s0) * sample - (alpha_t * (torch.exp(-h) - 1.0)) * D0
                - 0.5 * (alpha_t * (torch.exp(-h) - 1.0)) * D1)

    def step(self, model_output, timestep, sample, return_dict=False, **kw):
        n = len(self.timesteps)
        lower_order_final = self.step_index == n - 1  # final_sigmas_type == "zero"
        m = self.convert_model_output(model_output, sample)
        self.model_outputs = self.model_outputs[1:] + [m]
        sample = sample.to(torch.float32)
        if self.lower_order_nums < 1 or lower_order_final:
            prev = self._first_order(m, sample)
        else:
            prev = self._second_order(self.model_outputs, sample)
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1
        self.step_index += 1
        return (prev,)


def _randn_like(ref: torch.Tensor, generator):
    """diffusers.utils.torch_utils.randn_tensor: a CPU generator draws on the CPU, then the sample moves to the tensor's device."""
    dev = generator.device if generator is not None else ref.device
    return torch.randn(ref.shape, generator=generator, device=dev, dtype=ref.dtype).to(ref.device)


def rescale_noise_cfg(noise_cfg, noise_pred_text, guidance_rescale=0.0):
    """stable_diffusion_dual_unet.py:71-94."""
    std_text = noise_pred_text.std(dim=list(range(1, noise_pred_text.ndim)), keepdim=True)
    std_cfg = noise_cfg.std(dim=list(range(1, noise_cfg.ndim)), keepdim=True)
    rescaled = noise_cfg * (std_text / std_cfg)
    return guidance_rescale * rescaled + (1 - guidance_rescale) * noise_cfg
