"""Oracle schedulers (TEST INFRASTRUCTURE ONLY): restatement of diffusers' DDPMScheduler,
DDIMScheduler and DPMSolverMultistepScheduler (third-party `diffusers`, un-vendored, un-pinned,
0.21 <= v <= ~0.26) as the reference configures and calls them:

* construction: `ldm/train_unconditional.py:347-352` (1000 train steps, linear betas 1e-4..0.02,
  epsilon prediction, clip_sample=False, timestep_spacing="leading");
* call sites: `ldm/pipelines.py:99,106,227,244-246,336,338,356,362`, `ldm/inference.py:126-127`.

Algorithm: SURVEY.md App. A.4.  DPM-Solver++(2M) final-step convention frozen to
`final_sigmas_type="zero"` (diffusers >= 0.26, identical to the in-tree `DPMPP2MSampler` with
`append_zero`, `vae/sgm/modules/diffusionmodules/sampling.py:333-335`); the older convention is
available as `final_sigmas_type="sigma_min"`.  PARITY UNPINNED for DDIM/DDPM (no reference
vectors); the DPM++2M update is pinned against the reference's DPMPP2MSampler in tests.
"""
import numpy as np
import torch


class _Base:
    def __init__(self, num_train_timesteps=1000, beta_start=1e-4, beta_end=0.02, timestep_spacing="leading",
                 steps_offset=0):
        self.num_train_timesteps = num_train_timesteps
        self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.timestep_spacing = timestep_spacing
        self.steps_offset = steps_offset
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def scale_model_input(self, sample, timestep=None):
        return sample

    def _leading(self, n):
        step_ratio = self.num_train_timesteps // n
        ts = (np.arange(0, n) * step_ratio).round()[::-1].copy().astype(np.int64)
        return ts + self.steps_offset

    def set_timesteps(self, n):
        self.num_inference_steps = n
        if self.timestep_spacing == "leading":
            ts = self._leading(n)
        elif self.timestep_spacing == "linspace":
            ts = np.linspace(0, self.num_train_timesteps - 1, n).round()[::-1].copy().astype(np.int64)
        elif self.timestep_spacing == "trailing":
            ts = np.round(np.arange(self.num_train_timesteps, 0, -self.num_train_timesteps / n)).astype(np.int64) - 1
        else:
            raise ValueError(self.timestep_spacing)
        self.timesteps = torch.from_numpy(ts)


class OracleDDIMScheduler(_Base):
    final_alpha_cumprod = torch.tensor(1.0)

    def step(self, model_output, timestep, sample, eta=0.0, variance_noise=None):
        t = int(timestep)
        prev = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        var = ((1 - a_p) / (1 - a_t)) * (1 - a_t / a_p)
        std = eta * var ** 0.5
        direction = (1 - a_p - std ** 2) ** 0.5 * model_output
        x = a_p ** 0.5 * x0 + direction
        if eta > 0:
            x = x + std * variance_noise
        return x


class OracleDDPMScheduler(_Base):
    def step(self, model_output, timestep, sample, variance_noise=None):
        t = int(timestep)
        prev = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev] if prev >= 0 else torch.tensor(1.0)
        b_t, b_p = 1 - a_t, 1 - a_p
        cur_a = a_t / a_p
        cur_b = 1 - cur_a
        x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        c0 = (a_p ** 0.5 * cur_b) / b_t
        cx = cur_a ** 0.5 * b_p / b_t
        x = c0 * x0 + cx * sample
        if t > 0:
            var = torch.clamp((1 - a_p) / (1 - a_t) * cur_b, min=1e-20)
            x = x + (var ** 0.5) * variance_noise
        return x


class OracleDPMSolverMultistepScheduler(_Base):
    """solver_order=2, dpmsolver++, midpoint, epsilon prediction, no karras sigmas."""

    def __init__(self, final_sigmas_type="zero", lower_order_final=True, **kw):
        kw.setdefault("timestep_spacing", "linspace")
        super().__init__(**kw)
        self.final_sigmas_type = final_sigmas_type
        self.lower_order_final = lower_order_final

    def set_timesteps(self, n):
        self.num_inference_steps = n
        last = self.num_train_timesteps
        if self.timestep_spacing == "linspace":
            ts = np.linspace(0, last - 1, n + 1).round()[::-1][:-1].copy().astype(np.int64)
        elif self.timestep_spacing == "leading":
            ts = (np.arange(0, n + 1) * (last // (n + 1))).round()[::-1][:-1].copy().astype(np.int64) + self.steps_offset
        elif self.timestep_spacing == "trailing":
            ts = np.arange(last, 0, -self.num_train_timesteps / n).round().copy().astype(np.int64) - 1
        else:
            raise ValueError(self.timestep_spacing)
        sig = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        sigmas = np.interp(ts, np.arange(0, len(sig)), sig)
        if self.final_sigmas_type == "sigma_min":
            sigma_last = ((1 - self.alphas_cumprod[0]) / self.alphas_cumprod[0]) ** 0.5
        else:
            sigma_last = 0
        self.sigmas = torch.from_numpy(np.concatenate([sigmas, [sigma_last]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts)
        self.model_outputs = [None, None]
        self.lower_order_nums = 0
        self.step_index = 0

    @staticmethod
    def _alpha_sigma(sigma):
        alpha_t = 1 / ((sigma ** 2 + 1) ** 0.5)
        return alpha_t, sigma * alpha_t

    def step(self, model_output, timestep, sample):
        i, n = self.step_index, len(self.timesteps)
        final = (i == n - 1) and ((self.lower_order_final and n < 15) or self.final_sigmas_type == "zero")
        second_last = (i == n - 2) and self.lower_order_final and n < 15
        a_s0, s_s0 = self._alpha_sigma(self.sigmas[i])
        x0 = (sample - s_s0 * model_output) / a_s0
        self.model_outputs = [self.model_outputs[1], x0]
        a_t, s_t = self._alpha_sigma(self.sigmas[i + 1])
        lam_t = torch.log(a_t) - torch.log(s_t)
        lam_s0 = torch.log(a_s0) - torch.log(s_s0)
        h = lam_t - lam_s0
        if self.lower_order_nums < 1 or final or second_last:
            x = (s_t / s_s0) * sample - (a_t * (torch.exp(-h) - 1.0)) * x0
        else:
            a_s1, s_s1 = self._alpha_sigma(self.sigmas[i - 1])
            lam_s1 = torch.log(a_s1) - torch.log(s_s1)
            h0 = lam_s0 - lam_s1
            r0 = h0 / h
            m0, m1 = self.model_outputs[-1], self.model_outputs[-2]
            D0, D1 = m0, (1.0 / r0) * (m0 - m1)
            x = (s_t / s_s0) * sample - (a_t * (torch.exp(-h) - 1.0)) * D0 \
                - 0.5 * (a_t * (torch.exp(-h) - 1.0)) * D1
        if self.lower_order_nums < 2:
            self.lower_order_nums += 1
        self.step_index += 1
        return x
