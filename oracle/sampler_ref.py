"""oracle/sampler_ref.py — TEST INFRASTRUCTURE (see oracle/__init__.py).

CPU restatement of the reference sampling loops and the training forward:
  models/diffusion_ddpm.py:223-277 (Diffusion_DDPM.sample), :176-214 (validate), :128-173
  (process_single_batch), :216-219 (add_constraints); models/diffusion_ddim.py:23-74.
The reference loops run batch element 0 only; `sample_ref` takes any batch (the batched extension
the new path adds behind the same method) — with B=1 it is exactly the reference loop.
"""
import torch

from . import unet_ref
from .schedulers import RefDDIMScheduler, RefDDPMScheduler


def make_scheduler(kind, num_train_timesteps):
    cls = RefDDPMScheduler if kind == "ddpm" else RefDDIMScheduler
    return cls(num_train_timesteps=num_train_timesteps, beta_schedule="linear", clip_sample=False,
               prediction_type="epsilon")


def add_constraints(x_t, x_inpaint, inpaint_horizon):
    """models/diffusion_ddpm.py:216-219 (in place)."""
    x_t[:, :, :inpaint_horizon, :] = x_inpaint
    return x_t


@torch.no_grad()
def sample_ref(sd, scheduler, n_steps, x_T, obs_cond, inpaint, inpaint_horizon, attention=True, noise=None,
               history=False, max_steps=None, forward=None):
    """Loop of models/diffusion_ddpm.py:268-276 / diffusion_ddim.py:67-73.

    obs_cond (B,1,T,cond_dim); inpaint (B,1,ih,dim); noise (K,B,1,rows,dim) injected per step index
    (used where the scheduler adds noise, i.e. DDPM t>0)."""
    x_t = x_T.clone()
    hist = [x_t.clone()]
    scheduler.set_timesteps(n_steps)
    for i, t in enumerate(scheduler.timesteps):
        if max_steps is not None and i >= max_steps:
            break
        if forward is not None:   # another noise estimator (oracle/simple_unet_ref.unet_forward)
            est = forward(sd, x_t, torch.tensor([int(t)]), obs_cond)
        else:
            est = unet_ref.unet_forward(sd, x_t, torch.tensor([int(t)]), obs_cond, attention=attention)
        if isinstance(scheduler, RefDDPMScheduler):
            x_t = scheduler.step(est, t, x_t, noise=None if noise is None else noise[i]).prev_sample
        else:
            x_t = scheduler.step(est, t, x_t).prev_sample
        x_t = add_constraints(x_t, inpaint, inpaint_horizon)
        if history:
            hist.append(x_t.clone())
    return hist if history else x_t


def training_forward_ref(sd, esd, scheduler, batch, obs_horizon, inpaint_horizon, t, noise, attention=True):
    """models/diffusion_ddpm.py:128-173 with injected t (B,) and noise: returns (loss, x_noisy, noise_est)."""
    obs = {k: v[:, :obs_horizon].float() for k, v in batch.items()}
    pred = {k: v[:, obs_horizon:].float() for k, v in batch.items()}
    cond = unet_ref.obs_cond(esd, obs).unsqueeze(1)
    x0 = torch.cat([pred["position"], pred["action"]], dim=-1).unsqueeze(1)
    inp = unet_ref.inpaint_vector(obs, inpaint_horizon).unsqueeze(1)
    vec = torch.cat([inp, x0], dim=2)
    x_noisy = scheduler.add_noise(vec, noise, t)
    x_noisy = add_constraints(x_noisy, inp, inpaint_horizon)
    est = unet_ref.unet_forward(sd, x_noisy, t, cond, attention=attention)
    loss = torch.nn.functional.mse_loss(noise, est)
    return loss, x_noisy, est
