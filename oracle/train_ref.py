"""oracle/train_ref.py — TEST INFRASTRUCTURE (see oracle/__init__.py).

CPU restatement of the reference's training step: loss + gradients of `process_single_batch`
(models/diffusion_ddpm.py:128-173) for every trainable tensor the reference hands to Adam
(`configure_optimizers`, ddpm:115-125: the U-Net AND the vision encoder), the gradient-norm clip of
`pl.Trainer(gradient_clip_val=0.5)` (train.py:104-107) and one `torch.optim.Adam` update.
Gradients come from torch autograd over the functional restatement (unet_ref / sampler_ref), pinned
against `loss.backward()` on the reference's own modules by oracle/make_golden.py ->
tests/golden/train_grads.npz (checked in tests/test_oracle_golden.py).
"""
import torch

from . import sampler_ref

ENC_PREFIX = "vision_encoder."


def loss_and_grads(sd, esd, scheduler, batch, obs_horizon, inpaint_horizon, t, noise, attention=True):
    """Returns (loss, grads) with grads keyed like the reference's `named_parameters()`:
    U-Net tensors under their state_dict names, encoder tensors under `vision_encoder.<k>`."""
    sd_g = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    esd_g = {k: v.detach().clone().requires_grad_(True) for k, v in esd.items()}
    loss, _, _ = sampler_ref.training_forward_ref(sd_g, esd_g, scheduler, batch, obs_horizon, inpaint_horizon, t, noise,
                                                  attention=attention)
    names = list(sd_g) + [ENC_PREFIX + k for k in esd_g]
    leaves = list(sd_g.values()) + list(esd_g.values())
    gs = torch.autograd.grad(loss, leaves, allow_unused=True)
    grads = {n: (g if g is not None else torch.zeros_like(p)) for n, g, p in zip(names, gs, leaves)}
    return loss.detach(), grads


def clip_grad_norm(grads, max_norm=0.5):
    """torch.nn.utils.clip_grad_norm_ (what Lightning's gradient_clip_val uses, algorithm 'norm'):
    total = ||all grads||_2 ; coef = min(1, max_norm / (total + 1e-6)).  Returns (total_norm, clipped grads)."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return total, {k: g * coef for k, g in grads.items()}


def adam_step(params, grads, m, v, step, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam (default flags) single-tensor update, `step` counted from 1."""
    out_p, out_m, out_v = {}, {}, {}
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    for k, p in params.items():
        g = grads[k]
        mk = beta1 * m[k] + (1 - beta1) * g
        vk = beta2 * v[k] + (1 - beta2) * g * g
        denom = vk.sqrt() / (bc2 ** 0.5) + eps
        out_p[k] = p - (lr / bc1) * mk / denom
        out_m[k], out_v[k] = mk, vk
    return out_p, out_m, out_v


def summary(t, n_probe=8):
    """Small per-tensor fingerprint used by the golden file: [sum, abs-sum, l2, probes...] at fixed strided indices."""
    f = t.detach().flatten().double()
    idx = torch.linspace(0, f.numel() - 1, n_probe).long()
    return torch.cat([torch.stack([f.sum(), f.abs().sum(), f.norm()]), f[idx]]).float()
