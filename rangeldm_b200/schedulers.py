"""`DDPMScheduler`, `DDIMScheduler`, `DPMSolverMultistepScheduler` with the diffusers surface the
reference uses (`ldm/inference.py:126-127`, `ldm/pipelines.py:99,106,139,227,244-246,336-341,356,362`;
construction `ldm/train_unconditional.py:347-352`; algorithm SURVEY.md App. A.4).

Host side: integer timestep tables (bit-exact with the reference schedulers) and a per-step table
of 7 affine coefficients, uploaded once by `set_timesteps`.  Device side: every `step` is ONE
launch of `rldm_sched_step` (include/rldm.h):

    x0   = k0*x + k1*eps
    x'   = k2*x + k3*x0 + k4*x0_prev + k5*eps + k6*noise

which replaces the ~10 elementwise launches with CPU-scalar coefficients of the reference path.
"""
import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib
from .configuration import ConfigMixin


@dataclass
class SchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


def _betas(num_train_timesteps, beta_start, beta_end, beta_schedule):
    if beta_schedule == "linear":
        return torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    if beta_schedule == "scaled_linear":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    raise NotImplementedError(f"beta_schedule {beta_schedule!r}")


class _SchedulerBase(ConfigMixin):
    config_name = "scheduler_config.json"
    order = 1

    def _setup(self):
        c = self.config
        if c.prediction_type != "epsilon":
            raise NotImplementedError("only prediction_type='epsilon' (the reference's) is implemented")
        self.betas = _betas(c.num_train_timesteps, c.beta_start, c.beta_end, c.beta_schedule)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)     # fp32, as the reference
        self._ac = self.alphas_cumprod.double().numpy()
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, c.num_train_timesteps)[::-1].copy().astype(np.int64))
        self._coef_dev = None
        self._coef_host = None
        self._index = {}

    def scale_model_input(self, sample, timestep=None):
        return sample

    def _spaced(self, n):
        c = self.config
        N = c.num_train_timesteps
        if n > N:
            raise ValueError(f"num_inference_steps {n} > num_train_timesteps {N}")
        if c.timestep_spacing == "leading":
            ts = (np.arange(0, n) * (N // n)).round()[::-1].copy().astype(np.int64) + c.steps_offset
        elif c.timestep_spacing == "linspace":
            ts = np.linspace(0, N - 1, n).round()[::-1].copy().astype(np.int64)
        elif c.timestep_spacing == "trailing":
            ts = np.round(np.arange(N, 0, -N / n)).astype(np.int64) - 1
        else:
            raise ValueError(c.timestep_spacing)
        return ts

    def _install(self, ts, coefs):
        self.timesteps = torch.from_numpy(np.ascontiguousarray(ts))
        self._index = {int(t): i for i, t in enumerate(ts)}
        self._coef_host = torch.tensor(np.asarray(coefs, dtype=np.float64), dtype=torch.float32)
        self._coef_dev = None
        self._state = {}
        self._steps_key = None          # set by the pipelines after THEIR set_timesteps call (see _RangePipeline._set_steps)

    def coef_table(self, device):
        """(steps, 8) fp32 coefficient table on `device` (7 used, padded to 8 for 32 B rows)."""
        if self._coef_dev is None or self._coef_dev.device != device:
            self._coef_dev = self._coef_host.to(device)
        return self._coef_dev

    def _launch(self, i, sample, eps, x0_prev, noise, want_x0):
        if not sample.is_cuda:
            raise RuntimeError("scheduler.step runs on CUDA only (no CPU fallback)")
        sample = sample.contiguous()
        eps = eps.contiguous()
        out = torch.empty_like(sample)
        x0 = torch.empty_like(sample) if want_x0 else None
        k = self.coef_table(sample.device)[i]
        _lib.call("rldm_sched_step", _lib.ptr(k), _lib.ptr(sample), _lib.ptr(eps), _lib.ptr(x0_prev), _lib.ptr(noise),
                  _lib.ptr(out), _lib.ptr(x0), sample.numel())
        return out, x0

    def _step_index(self, timestep):
        t = int(timestep)
        if t not in self._index:
            raise ValueError(f"timestep {t} is not in the schedule set by set_timesteps")
        return self._index[t]

    def add_noise(self, original, noise, timesteps):
        ac = self.alphas_cumprod.to(original.device)[timesteps].view(-1, *([1] * (original.ndim - 1)))
        return ac.sqrt() * original + (1 - ac).sqrt() * noise


class DDIMScheduler(_SchedulerBase):
    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 trained_betas=None, clip_sample=True, set_alpha_to_one=True, steps_offset=0,
                 prediction_type="epsilon", thresholding=False, dynamic_thresholding_ratio=0.995,
                 clip_sample_range=1.0, sample_max_value=1.0, timestep_spacing="leading",
                 rescale_betas_zero_snr=False):
        self._capture_init(locals())
        if clip_sample or thresholding or trained_betas is not None or rescale_betas_zero_snr:
            raise NotImplementedError("DDIMScheduler: clip_sample/thresholding/trained_betas are not used by the "
                                      "reference (clip_sample=False, `ldm/train_unconditional.py:351`)")
        self._setup()
        self.final_alpha_cumprod = 1.0 if set_alpha_to_one else float(self._ac[0])

    def set_timesteps(self, num_inference_steps, device=None, eta=0.0):
        self.num_inference_steps = num_inference_steps
        self._eta = eta
        ts = self._spaced(num_inference_steps)
        N = self.config.num_train_timesteps
        coefs = []
        for t in ts:
            p = int(t) - N // num_inference_steps
            a_t = self._ac[t]
            a_p = self._ac[p] if p >= 0 else self.final_alpha_cumprod
            var = ((1 - a_p) / (1 - a_t)) * (1 - a_t / a_p)
            std = eta * math.sqrt(max(var, 0.0))
            coefs.append([1 / math.sqrt(a_t), -math.sqrt(1 - a_t) / math.sqrt(a_t), 0.0, math.sqrt(a_p), 0.0,
                          math.sqrt(max(1 - a_p - std * std, 0.0)), std, 0.0])
        self._install(ts, coefs)

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output=None,
             generator=None, variance_noise=None, return_dict: bool = True):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' first")
        if eta != getattr(self, "_eta", 0.0):
            self.set_timesteps(self.num_inference_steps, eta=eta)
        noise = None
        if eta > 0:
            from .pipelines import randn_tensor
            noise = variance_noise if variance_noise is not None else randn_tensor(
                model_output.shape, generator=generator, device=model_output.device, dtype=model_output.dtype)
        out, x0 = self._launch(self._step_index(timestep), sample, model_output, None, noise, True)
        return SchedulerOutput(out, x0) if return_dict else (out,)


class DDPMScheduler(_SchedulerBase):
    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 trained_betas=None, variance_type="fixed_small", clip_sample=True, prediction_type="epsilon",
                 thresholding=False, dynamic_thresholding_ratio=0.995, clip_sample_range=1.0, sample_max_value=1.0,
                 timestep_spacing="leading", steps_offset=0):
        self._capture_init(locals())
        self._clip = clip_sample
        if thresholding or trained_betas is not None or variance_type != "fixed_small":
            raise NotImplementedError("DDPMScheduler: only variance_type='fixed_small' without thresholding")
        self._setup()

    def set_timesteps(self, num_inference_steps, device=None):
        if self._clip:
            raise NotImplementedError("DDPMScheduler(clip_sample=True): the reference trains and samples with "
                                      "clip_sample=False (`ldm/train_unconditional.py:351`)")
        self.num_inference_steps = num_inference_steps
        ts = self._spaced(num_inference_steps)
        N = self.config.num_train_timesteps
        coefs = []
        for t in ts:
            p = int(t) - N // num_inference_steps
            a_t = self._ac[t]
            a_p = self._ac[p] if p >= 0 else 1.0
            cur_a = a_t / a_p
            cur_b = 1 - cur_a
            c0 = math.sqrt(a_p) * cur_b / (1 - a_t)
            cx = math.sqrt(cur_a) * (1 - a_p) / (1 - a_t)
            var = max((1 - a_p) / (1 - a_t) * cur_b, 1e-20)
            coefs.append([1 / math.sqrt(a_t), -math.sqrt(1 - a_t) / math.sqrt(a_t), cx, c0, 0.0, 0.0,
                          math.sqrt(var) if t > 0 else 0.0, 0.0])
        self._install(ts, coefs)

    def step(self, model_output, timestep, sample, generator=None, variance_noise=None, return_dict: bool = True):
        if self.num_inference_steps is None:
            self.set_timesteps(self.config.num_train_timesteps)
        noise = None
        if int(timestep) > 0:
            from .pipelines import randn_tensor
            noise = variance_noise if variance_noise is not None else randn_tensor(
                model_output.shape, generator=generator, device=model_output.device, dtype=model_output.dtype)
        out, x0 = self._launch(self._step_index(timestep), sample, model_output, None, noise, True)
        return SchedulerOutput(out, x0) if return_dict else (out,)


class DPMSolverMultistepScheduler(_SchedulerBase):
    """DPM-Solver++(2M), midpoint, epsilon prediction (BASELINE's sampler).  The final step follows
    `final_sigmas_type="zero"` (diffusers >= 0.26; == the in-tree `DPMPP2MSampler` with append_zero,
    `vae/sgm/modules/diffusionmodules/sampling.py:333-335`)."""
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 trained_betas=None, solver_order=2, prediction_type="epsilon", thresholding=False,
                 dynamic_thresholding_ratio=0.995, sample_max_value=1.0, algorithm_type="dpmsolver++",
                 solver_type="midpoint", lower_order_final=True, euler_at_final=False, use_karras_sigmas=False,
                 lambda_min_clipped=-float("inf"), variance_type=None, timestep_spacing="linspace", steps_offset=0,
                 final_sigmas_type="zero"):
        self._capture_init(locals())
        if (solver_order != 2 or algorithm_type != "dpmsolver++" or solver_type != "midpoint" or thresholding
                or use_karras_sigmas or trained_betas is not None or euler_at_final):
            raise NotImplementedError("DPMSolverMultistepScheduler: only solver_order=2, dpmsolver++, midpoint")
        self._setup()

    def set_timesteps(self, num_inference_steps, device=None):
        c = self.config
        n, N = num_inference_steps, c.num_train_timesteps
        self.num_inference_steps = n
        if c.timestep_spacing == "linspace":
            ts = np.linspace(0, N - 1, n + 1).round()[::-1][:-1].copy().astype(np.int64)
        elif c.timestep_spacing == "leading":
            ts = (np.arange(0, n + 1) * (N // (n + 1))).round()[::-1][:-1].copy().astype(np.int64) + c.steps_offset
        elif c.timestep_spacing == "trailing":
            ts = np.arange(N, 0, -N / n).round().copy().astype(np.int64) - 1
        else:
            raise ValueError(c.timestep_spacing)
        sig_all = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()    # fp32, as diffusers
        sig = np.interp(ts, np.arange(0, N), sig_all)
        last = math.sqrt((1 - self._ac[0]) / self._ac[0]) if c.final_sigmas_type == "sigma_min" else 0.0
        sig = np.concatenate([sig, [last]]).astype(np.float32).astype(np.float64)   # reference keeps fp32 sigmas
        self.sigmas = torch.from_numpy(sig.astype(np.float32))
        alpha = 1.0 / np.sqrt(sig ** 2 + 1.0)
        sigma = sig * alpha
        with np.errstate(divide="ignore"):
            lam = np.log(alpha) - np.log(sigma)
        coefs = []
        for i in range(n):
            final = (i == n - 1) and ((c.lower_order_final and n < 15) or c.final_sigmas_type == "zero")
            a_s, s_s, a_t, s_t = alpha[i], sigma[i], alpha[i + 1], sigma[i + 1]
            h = lam[i + 1] - lam[i]
            cc = -a_t * math.expm1(-h) if np.isfinite(h) else a_t      # -(alpha_t (e^{-h} - 1))
            k = [1 / a_s, -s_s / a_s, s_t / s_s, cc, 0.0, 0.0, 0.0, 0.0]
            # solver_order == 2: diffusers' `lower_order_second` only demotes a THIRD-order solver, so step n-2 stays
            # second order; only the first step and the final step are first order (== in-tree DPMPP2MSampler)
            if not (i == 0 or final):
                r0 = (lam[i] - lam[i - 1]) / h
                k[3] = cc * (1 + 0.5 / r0)
                k[4] = -cc * 0.5 / r0
            coefs.append(k)
        self._install(ts, coefs)
        self._x0_prev = None
        self._counter = 0

    def step(self, model_output, timestep, sample, generator=None, return_dict: bool = True):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' first")
        i = self._counter                 # diffusers' _step_index: advances by one per call
        if i >= self.num_inference_steps:
            raise IndexError("DPMSolverMultistepScheduler.step called more often than num_inference_steps")
        out, x0 = self._launch(i, sample, model_output, self._x0_prev, None, True)
        self._x0_prev = x0
        self._counter += 1
        return SchedulerOutput(out, x0) if return_dict else (out,)
