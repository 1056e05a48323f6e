"""oracle/schedulers.py — TEST INFRASTRUCTURE (see oracle/__init__.py).

CPU restatement of the scheduler arithmetic the reference calls:
  diffusers==0.17.1 (reference requirements.txt:30, NOT vendored => parity unpinned vs the package)
    schedulers/scheduling_ddpm.py :: DDPMScheduler.{__init__, set_timesteps, step, add_noise}
    schedulers/scheduling_ddim.py :: DDIMScheduler.{__init__, set_timesteps, step}
  call sites: models/diffusion_ddpm.py:65-70,167,204-211,257-262,268-274;
              models/diffusion_ddim.py:57-72; generate.py:28-33.
Semantics restated from the published 0.17.1 algorithm: all coefficient math in
fp32 0-dim CPU tensors, in the library's operation order.
"""
import torch


class _Out:
    def __init__(self, prev_sample, pred_original_sample=None):
        self.prev_sample = prev_sample
        self.pred_original_sample = pred_original_sample


def _betas(num_train_timesteps, beta_start, beta_end, beta_schedule):
    if beta_schedule == "linear":
        return torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    if beta_schedule == "scaled_linear":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    raise NotImplementedError(beta_schedule)


class RefDDPMScheduler:
    """DDPMScheduler(variance_type='fixed_small', prediction_type='epsilon')."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 clip_sample=True, prediction_type="epsilon"):
        assert prediction_type == "epsilon"
        self.num_train_timesteps = num_train_timesteps
        self.clip_sample = clip_sample
        self.betas = _betas(num_train_timesteps, beta_start, beta_end, beta_schedule)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.num_inference_steps = None
        self.timesteps = torch.arange(0, num_train_timesteps).flip(0)

    def set_timesteps(self, num_inference_steps):
        if num_inference_steps > self.num_train_timesteps:
            raise ValueError("num_inference_steps > num_train_timesteps")
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        self.timesteps = (torch.arange(0, num_inference_steps) * ratio).round().flip(0).to(torch.int64)

    def coefficients(self, t):
        """Returns dict of fp32 0-dim tensors for timestep t (restated op order)."""
        t = int(t)
        n = self.num_inference_steps if self.num_inference_steps else self.num_train_timesteps
        prev_t = t - self.num_train_timesteps // n
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        beta_prod_t = 1 - a_t
        beta_prod_prev = 1 - a_prev
        cur_alpha = a_t / a_prev
        cur_beta = 1 - cur_alpha
        c0 = (a_prev ** 0.5 * cur_beta) / beta_prod_t
        cx = cur_alpha ** 0.5 * beta_prod_prev / beta_prod_t
        var = torch.clamp((1 - a_prev) / (1 - a_t) * cur_beta, min=1e-20)
        return dict(sqrt_beta_prod=beta_prod_t ** 0.5, sqrt_alpha_prod=a_t ** 0.5, c0=c0, cx=cx,
                    sigma=var ** 0.5 if t > 0 else torch.tensor(0.0))

    def step(self, model_output, timestep, sample, noise=None, generator=None):
        t = int(timestep)
        c = self.coefficients(t)
        x0 = (sample - c["sqrt_beta_prod"] * model_output) / c["sqrt_alpha_prod"]
        if self.clip_sample:
            x0 = x0.clamp(-1, 1)
        prev = c["c0"] * x0 + c["cx"] * sample
        if t > 0:
            if noise is None:
                noise = torch.randn(model_output.shape, generator=generator, dtype=model_output.dtype)
            prev = prev + c["sigma"] * noise
        return _Out(prev, x0)

    def add_noise(self, original_samples, noise, timesteps):
        ac = self.alphas_cumprod.to(dtype=original_samples.dtype)
        sa = ac[timesteps] ** 0.5
        sb = (1 - ac[timesteps]) ** 0.5
        sa = sa.flatten()
        sb = sb.flatten()
        while sa.dim() < original_samples.dim():
            sa = sa.unsqueeze(-1)
            sb = sb.unsqueeze(-1)
        return sa * original_samples + sb * noise


class RefDDIMScheduler:
    """DDIMScheduler(set_alpha_to_one=True, steps_offset=0, prediction_type='epsilon'), eta=0."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 clip_sample=True, prediction_type="epsilon", set_alpha_to_one=True, steps_offset=0):
        assert prediction_type == "epsilon"
        self.num_train_timesteps = num_train_timesteps
        self.clip_sample = clip_sample
        self.steps_offset = steps_offset
        self.betas = _betas(num_train_timesteps, beta_start, beta_end, beta_schedule)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_inference_steps = None
        self.timesteps = torch.arange(0, num_train_timesteps).flip(0)

    def set_timesteps(self, num_inference_steps):
        if num_inference_steps > self.num_train_timesteps:
            raise ValueError("num_inference_steps > num_train_timesteps")
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        self.timesteps = ((torch.arange(0, num_inference_steps) * ratio).round().flip(0).to(torch.int64)
                          + self.steps_offset)

    def coefficients(self, t):
        t = int(t)
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        beta_prod_t = 1 - a_t
        return dict(sqrt_beta_prod=beta_prod_t ** 0.5, sqrt_alpha_prod=a_t ** 0.5,
                    sqrt_alpha_prev=a_prev ** 0.5, dir_coef=(1 - a_prev - 0.0) ** 0.5)

    def step(self, model_output, timestep, sample, eta=0.0):
        assert eta == 0.0
        c = self.coefficients(timestep)
        x0 = (sample - c["sqrt_beta_prod"] * model_output) / c["sqrt_alpha_prod"]
        if self.clip_sample:
            x0 = x0.clamp(-1, 1)
        direction = c["dir_coef"] * model_output
        prev = c["sqrt_alpha_prev"] * x0 + direction
        return _Out(prev, x0)

    add_noise = RefDDPMScheduler.add_noise


# utils/schedulers.py:6-40 restated (the reference functions are unbound "methods" reading self.device)
def ref_linear_beta_schedule(steps):
    scale = 1000 / steps
    return torch.linspace(scale * 0.0001, scale * 0.02, steps, dtype=torch.float32)


def ref_linear_beta_schedule_v2(steps):
    scale = 500 / steps
    return torch.linspace(scale * 0.0001, scale * 0.02, steps, dtype=torch.float32)


def ref_cosine_beta_schedule(timesteps, s=0.008, dtype=torch.float32):
    import numpy as np
    steps = timesteps + 1
    x = np.linspace(0, steps, steps)
    ac = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = 1 - (ac[1:] / ac[:-1])
    return torch.tensor(np.clip(betas, a_min=0, a_max=0.999), dtype=dtype)
