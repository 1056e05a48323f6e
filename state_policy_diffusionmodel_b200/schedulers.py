"""Noise schedules and scheduler objects of the denoising path.

* `linear_beta_schedule`, `linear_beta_schedule_v2`, `cosine_beta_schedule` keep the signatures of the
  reference's utils/schedulers.py:6-40 (they are written there as unbound methods that read
  `self.device`; `self` may be any object with a `.device`, or None).
* `DDPMScheduler` / `DDIMScheduler` present the subset of the diffusers==0.17.1 interface that the
  reference touches (models/diffusion_ddpm.py:65-70,167,204-211; generate.py:28-33): constructor
  kwargs, `.set_timesteps(n)`, `.timesteps`, `.step(eps, t, x).prev_sample`, `.add_noise(x0, noise, t)`.
  All coefficient arithmetic is fp32 on the host in the library's operation order; the tensor math
  of `.step` / `.add_noise` runs in the fused CUDA kernels of libspdm (no CPU fallback for tensors on
  the GPU path).  `coef_table()` exports the per-step coefficients the graphed loop consumes.
"""
import numpy as np
import torch


# ---------------------------------------------------------------------------------------------
# utils/schedulers.py drop-ins
# ---------------------------------------------------------------------------------------------
def _dev(self):
    return getattr(self, "device", None) if self is not None else None


def linear_beta_schedule(self, steps):
    """utils/schedulers.py:6-15 — linear betas scaled by 1000/steps."""
    scale = 1000 / steps
    return torch.linspace(scale * 0.0001, scale * 0.02, steps, dtype=torch.float32, device=_dev(self))


def linear_beta_schedule_v2(self, steps):
    """utils/schedulers.py:17-26 — same with scale 500/steps."""
    scale = 500 / steps
    return torch.linspace(scale * 0.0001, scale * 0.02, steps, dtype=torch.float32, device=_dev(self))


def cosine_beta_schedule(self, timesteps, s=0.008, dtype=torch.float32):
    """utils/schedulers.py:28-40 — squared-cosine alpha-bar, betas clipped to [0, 0.999] (numpy f64 -> dtype)."""
    steps = timesteps + 1
    x = np.linspace(0, steps, steps)
    alphas_cumprod = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    alphas_cumprod = alphas_cumprod / alphas_cumprod[0]
    betas = 1 - (alphas_cumprod[1:] / alphas_cumprod[:-1])
    return torch.tensor(np.clip(betas, a_min=0, a_max=0.999), dtype=dtype)


# ---------------------------------------------------------------------------------------------
# scheduler objects
# ---------------------------------------------------------------------------------------------
class SchedulerOutput:
    def __init__(self, prev_sample, pred_original_sample=None):
        self.prev_sample = prev_sample
        self.pred_original_sample = pred_original_sample


class _SchedulerBase:
    kind = None

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 trained_betas=None, clip_sample=True, prediction_type="epsilon", clip_sample_range=1.0, thresholding=False,
                 **unused):
        if prediction_type != "epsilon":
            raise NotImplementedError("only prediction_type='epsilon' (the reference's setting) is implemented")
        if thresholding:
            raise NotImplementedError("dynamic thresholding is not on the reference's path")
        if clip_sample and not float(clip_sample_range) > 0:
            raise ValueError("clip_sample_range must be positive")
        if trained_betas is not None:
            self.betas = torch.as_tensor(trained_betas, dtype=torch.float32).cpu()
        elif beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        self.num_train_timesteps = int(self.betas.numel())
        self.clip_sample = bool(clip_sample)
        self.clip_sample_range = float(clip_sample_range)
        self.prediction_type = prediction_type
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.arange(0, self.num_train_timesteps).flip(0).to(torch.int64)
        self._plan = None  # a DenoisePlan borrowed for the elementwise kernels of step()/add_noise()

    # diffusers 0.17.1 "leading" spacing: arange(n) * (T // n), reversed
    def set_timesteps(self, num_inference_steps, device=None):
        if num_inference_steps > self.num_train_timesteps:
            raise ValueError("`num_inference_steps` cannot be larger than `num_train_timesteps`")
        self.num_inference_steps = int(num_inference_steps)
        ratio = self.num_train_timesteps // self.num_inference_steps
        self.timesteps = (torch.arange(0, self.num_inference_steps) * ratio).round().flip(0).to(torch.int64)

    def _ratio(self):
        n = self.num_inference_steps if self.num_inference_steps else self.num_train_timesteps
        return self.num_train_timesteps // n

    def scale_model_input(self, sample, timestep=None):
        return sample

    def coef_row(self, t):
        raise NotImplementedError

    def coef_table(self):
        """(K, 8) fp32 rows {c0, c1, k_x0, k_x, k_eps, k_noise, clip, 0} for the current `timesteps`:
           x0 = (x - c0*eps)/c1 ;  clip > 0 (clip_sample=True): x0 = x0.clamp(-clip, clip) ;
           x_prev = k_x0*x0 + k_x*x + k_eps*eps + k_noise*z."""
        return torch.stack([self.coef_row(int(t)) for t in self.timesteps])

    def _clip(self):
        return torch.tensor(self.clip_sample_range if self.clip_sample else 0.0)

    # -- tensor math: fused CUDA kernels ---------------------------------------------------------
    def _borrow_plan(self, sample):
        from .engine import DenoisePlan
        rows, dim = int(sample.shape[-2]), int(sample.shape[-1])
        p = self._plan
        if p is None or (p.rows, p.dim) != (rows, dim) or p.device != sample.device:
            p = DenoisePlan(attention=False, precision="fp32", batch_max=1, rows=rows, dim=dim, obs_horizon=1, cond_dim=0,
                            inpaint_rows=0, device=sample.device, graph_steps=0, scheduler_only=True)
            self._plan = p
        return p

    def step(self, model_output, timestep, sample, eta=0.0, generator=None, variance_noise=None, return_dict=True, **unused):
        if eta != 0.0:
            raise NotImplementedError("eta != 0 is not on the reference's path")
        if not sample.is_cuda:
            raise RuntimeError("spdm schedulers run their tensor math on the GPU only (no CPU fallback)")
        t = int(timestep)
        row = self.coef_row(t)
        p = self._borrow_plan(sample)
        p.set_schedule(self.kind, row[None], torch.tensor([t]))
        noise = None
        if float(row[5]) != 0.0:
            noise = variance_noise if variance_noise is not None else torch.randn(
                model_output.shape, generator=generator, device=model_output.device, dtype=model_output.dtype)
        shp = sample.shape
        B = int(np.prod(shp[:-2])) if sample.dim() > 2 else 1
        prev = p.step(sample.reshape(B, 1, shp[-2], shp[-1]), model_output.reshape(B, 1, shp[-2], shp[-1]), 0,
                      noise=None if noise is None else noise.reshape(B, 1, shp[-2], shp[-1])).reshape(shp)
        out = SchedulerOutput(prev.to(sample.dtype))
        return out if return_dict else (out.prev_sample,)

    def add_noise(self, original_samples, noise, timesteps):
        if not original_samples.is_cuda:
            raise RuntimeError("spdm schedulers run their tensor math on the GPU only (no CPU fallback)")
        ac = self.alphas_cumprod
        p = self._borrow_plan(original_samples)
        shp = original_samples.shape
        B = shp[0]
        t = timesteps.reshape(-1).to(torch.int64)
        if t.numel() == 1 and B > 1:
            t = t.expand(B)
        x0 = original_samples.reshape(B, -1)
        # the kernel treats each batch row as one "sample" of rows*dim elements
        if x0.shape[1] != p.rows * p.dim:
            raise ValueError("add_noise expects (B, ..., rows, dim) samples")
        out = p.add_noise(x0, noise.reshape(B, -1), t, ac ** 0.5, (1 - ac) ** 0.5)
        return out.reshape(shp).to(original_samples.dtype)

    def __len__(self):
        return self.num_train_timesteps


class DDPMScheduler(_SchedulerBase):
    """diffusers 0.17.1 DDPMScheduler(variance_type='fixed_small'), epsilon prediction; clip_sample as in the library
    (default True, range 1.0; the reference passes False)."""
    kind = "ddpm"

    def coef_row(self, t):
        prev_t = t - self._ratio()
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        beta_prod_t = 1 - a_t
        beta_prod_prev = 1 - a_prev
        cur_alpha = a_t / a_prev
        cur_beta = 1 - cur_alpha
        k_x0 = (a_prev ** 0.5 * cur_beta) / beta_prod_t
        k_x = cur_alpha ** 0.5 * beta_prod_prev / beta_prod_t
        if t > 0:
            var = torch.clamp((1 - a_prev) / (1 - a_t) * cur_beta, min=1e-20)
            sigma = var ** 0.5
        else:
            sigma = torch.tensor(0.0)
        z = torch.tensor(0.0)
        return torch.stack([beta_prod_t ** 0.5, a_t ** 0.5, k_x0, k_x, z, sigma, self._clip(), z]).to(torch.float32)


class DDIMScheduler(_SchedulerBase):
    """diffusers 0.17.1 DDIMScheduler(set_alpha_to_one=True, steps_offset=0), eta = 0."""
    kind = "ddim"

    def __init__(self, *args, set_alpha_to_one=True, steps_offset=0, **kwargs):
        super().__init__(*args, **kwargs)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.steps_offset = steps_offset

    def set_timesteps(self, num_inference_steps, device=None):
        super().set_timesteps(num_inference_steps, device)
        self.timesteps = self.timesteps + self.steps_offset

    def coef_row(self, t):
        prev_t = t - self._ratio()
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        beta_prod_t = 1 - a_t
        z = torch.tensor(0.0)
        # (use_clipped_model_output=False, diffusers' default: the direction term keeps the unclipped model output)
        return torch.stack([beta_prod_t ** 0.5, a_t ** 0.5, a_prev ** 0.5, z, (1 - a_prev - 0.0) ** 0.5, z, self._clip(), z]).to(torch.float32)
