"""oracle/ref_runner.py — TEST INFRASTRUCTURE (see oracle/__init__.py): the CPU arm of bench.py.

Times the reference's OWN modules (oracle/_ref: unmodified copies of models/Unet_FiLmLayer*.py, models/diffusion_dd{pm,im}.py,
models/encoder/autoencoder.py made by oracle/build_ref.py) on the host cores: `Diffusion_DDIM` / `Diffusion_DDPM` are
constructed through oracle/shim.py, their `prepare_obs_cond_vectors` / `prepare_inpaint_vectors` / `add_constraints` /
`noise_estimator` / `vision_encoder` are called as they are, and the K-step loop of models/diffusion_ddim.py:67-74
(models/diffusion_ddpm.py:268-276) is run over ALL rows of the batch -- the reference's `sample()` itself keeps batch element 0
only, so the loop body is restated here with the batch dimension left in; everything inside it is the reference's code.
When oracle/_ref is absent (a fresh clone on a box that never saw /root/reference) the functional port of oracle/unet_ref.py is
timed instead and the result is labelled kind = "port".
"""
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF, "models", "diffusion_ddim.py"))


_model_cache = {}


def _model(kind, attention, K, dim, rows, sd, esd):
    """The reference wrapper with the fixture weights loaded (strict)."""
    key = (kind, attention, K, dim, rows)
    if key in _model_cache:
        return _model_cache[key]
    from . import shim
    shim.install(root=REF)
    from models.diffusion_ddim import Diffusion_DDIM
    from models.diffusion_ddpm import Diffusion_DDPM
    from .schedulers import RefDDIMScheduler
    name = "UNet_Film" if attention else "UNet_FilmnoAttention"
    cls = Diffusion_DDIM if kind == "ddim" else Diffusion_DDPM
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):   # the constructor prints its configuration
        m = cls(noise_steps=K, obs_horizon=10, pred_horizon=rows - 1, observation_dim=135, prediction_dim=dim, model=name,
                inpaint_horizon=1)
    if kind == "ddim":  # generate.py:28-35
        m.noise_scheduler = RefDDIMScheduler(num_train_timesteps=K, beta_schedule="linear", clip_sample=False, prediction_type="epsilon")
        m.noise_steps = K
    m.noise_estimator.load_state_dict(sd, strict=True)
    m.vision_encoder.load_state_dict(esd, strict=True)
    m.eval()
    _model_cache[key] = m
    return m


@torch.no_grad()
def sample_rate(kind, attention, K, batch, x_T, sd, esd, dim=5, max_steps=None, noise=None):
    """One sampling call over the whole batch on the reference modules.  Returns (trajectories/s, seconds, description).
    `max_steps` < K times only that many denoising steps and extrapolates the loop to K (flagged in the description)."""
    B, rows = x_T.shape[0], x_T.shape[2]
    m = _model(kind, attention, K, dim, rows, sd, esd)
    t0 = time.perf_counter()
    obs_cond = m.prepare_obs_cond_vectors(batch).unsqueeze(1)               # (B, 1, obs_horizon, obs_dim)
    inpaint_vector = m.prepare_inpaint_vectors(batch).unsqueeze(1)[..., :dim]
    t_enc = time.perf_counter() - t0
    x_t = x_T.clone()
    m.noise_scheduler.set_timesteps(m.noise_steps)
    n, t_loop = 0, 0.0
    for i, t in enumerate(m.noise_scheduler.timesteps):                     # ddim:67-74 / ddpm:268-276
        if max_steps is not None and i >= max_steps:
            break
        t0 = time.perf_counter()
        est_noise = m.noise_estimator(x_t, torch.tensor([t]), obs_cond)
        if kind == "ddpm":
            x_t = m.noise_scheduler.step(est_noise, t, x_t, noise=None if noise is None else noise[i]).prev_sample
        else:
            x_t = m.noise_scheduler.step(est_noise, t, x_t).prev_sample
        x_t = m.add_constraints(x_t, inpaint_vector)
        t_loop += time.perf_counter() - t0
        n += 1
    total = t_enc + t_loop * (K / n)
    what = "%d trajectories on the reference modules (oracle/_ref): conditioning encode %.2fs + %d denoising steps (%.3fs each)%s" % (
        B, t_enc, n, t_loop / n, "" if n == K else ", loop extrapolated to %d steps" % K)
    return B / total, total, what, x_t


def training_rate(attention, full, t, noise, sd, esd, reps=2, lr=1e-4):
    """The reference's training step on its own LightningModule: process_single_batch's arithmetic (ddpm:128-173) with the drawn
    t / noise injected, loss.backward(), clip_grad_norm_(0.5) (train.py:107) and torch.optim.Adam.step() (ddpm:115-125)."""
    B = t.numel()
    m = _model("ddpm", attention, 1000, 5, 31, sd, esd)
    m.train()
    m.vision_encoder.eval()
    opt = torch.optim.Adam(m.parameters(), lr=lr)
    times = []
    for i in range(reps + 1):
        t0 = time.perf_counter()
        observation_batch = m.prepare_observation_batch(full)
        prediction_batch = m.prepare_prediction_batch(full)
        obs_cond = m.prepare_obs_cond_vectors(observation_batch).unsqueeze(1)
        x_0 = m.prepare_prediction_vectors(prediction_batch).unsqueeze(1)
        x_0_inpaint = m.prepare_inpaint_vectors(observation_batch).unsqueeze(1)
        prediction_vector = torch.cat([x_0_inpaint, x_0], dim=2)
        x_noisy = m.noise_scheduler.add_noise(prediction_vector, noise, t)
        x_noisy = m.add_constraints(x_noisy, x_0_inpaint)
        est = m.noise_estimator(x_noisy, t, obs_cond)
        loss = m.loss(noise, est)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 0.5)
        opt.step()
        if i > 0:
            times.append(time.perf_counter() - t0)
    per = sum(times) / len(times)
    m.eval()
    return B / per, per, "%d samples per step on the reference LightningModule (oracle/_ref), %d timed steps of fwd + loss.backward() + clip + Adam (%.2fs each), 1 warm-up" % (
        B, len(times), per)
