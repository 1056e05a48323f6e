"""GPU parity of the training step: spdm_train_fwd_bwd / spdm_adam_step through the C ABI vs
the CPU oracle (oracle/train_ref.py: autograd over the restatement, pinned to the reference's loss.backward() + Adam by
tests/golden/train_grads*.npz).

The fixtures are seeded so that no ReLU pre-activation of the encoder's first conv lies within fp32 rounding of zero (with
seed 5 one does, at 1.7e-8: its sign then depends on the summation order and moves 3e-3 of one channel's gradient -- a kink of
the function, not an error of either implementation).

Tolerances (rel = max|a-b| / max|b| per tensor):
  fp32 path : every parameter gradient rel <= 1e-4 (measured ~8e-6; fp32 atomics, different summation order)
  bf16 path : global relative L2 over all gradients <= 1e-2 and per-tensor cosine >= 0.995 (measured 4e-3 / 0.9987):
              activations AND activation gradients are stored in bf16, accumulation is fp32.
"""
import os

import pytest
import torch

from oracle import fixtures, train_ref
from oracle.schedulers import RefDDPMScheduler

pytestmark = pytest.mark.gpu


def _case(B, seed=778):
    g = torch.Generator().manual_seed(seed)
    full = {"image": torch.rand((B, 40, 3, 96, 96), generator=g), "position": 0.3 * torch.randn((B, 40, 2), generator=g),
            "velocity": 2 * torch.rand((B, 40, 2), generator=g) - 1, "action": 2 * torch.rand((B, 40, 3), generator=g) - 1}
    t = torch.randint(0, 1000, (B,), generator=g)
    noise = torch.randn((B, 1, 31, 5), generator=g)
    return full, t, noise


def _named(sd, esd):
    named = dict(sd)
    named.update({"vision_encoder." + k: v for k, v in esd.items()})
    return named


def _gpu_grads(precision, attention, sd, esd, full, t, noise):
    import state_policy_diffusionmodel_b200 as spdm
    B = t.numel()
    plan = spdm.DenoisePlan(attention=attention, precision=precision, batch_max=B, inpaint_rows=1)
    plan.enable_training(_named(sd, esd))
    ac = RefDDPMScheduler(num_train_timesteps=1000, beta_schedule="linear", clip_sample=False).alphas_cumprod
    obs = {k: v[:, :10] for k, v in full.items()}
    pred = {k: v[:, 10:] for k, v in full.items()}
    inp = torch.cat([obs["position"][:, -1:], obs["action"][:, -1:]], dim=-1)
    vec = torch.cat([inp.unsqueeze(1), torch.cat([pred["position"], pred["action"]], dim=-1).unsqueeze(1)], dim=2)
    loss = plan.train_fwd_bwd(obs["image"], obs["position"], obs["action"], obs["velocity"], vec, noise, t, ac ** 0.5, (1 - ac) ** 0.5,
                              inpaint=inp.reshape(B, -1))
    torch.cuda.synchronize()
    return plan, float(loss.item()), {k: plan.grad_view(k).detach().cpu().clone() for k in plan.train_offsets}


def _oracle(attention, sd, esd, full, t, noise):
    sched = RefDDPMScheduler(num_train_timesteps=1000, beta_schedule="linear", clip_sample=False)
    torch.set_num_threads(os.cpu_count() or 1)
    return train_ref.loss_and_grads(sd, esd, sched, full, 10, 1, t, noise, attention=attention)


@pytest.mark.parametrize("attention", [True, False])
def test_fp32_gradients_match_oracle(attention):
    sd = fixtures.make_unet_weights(attention=attention, seed=0)
    esd = fixtures.make_encoder_weights()
    full, t, noise = _case(3)
    want_loss, want = _oracle(attention, sd, esd, full, t, noise)
    plan, loss, got = _gpu_grads("fp32", attention, sd, esd, full, t, noise)
    assert abs(loss - float(want_loss)) <= 1e-5 * abs(float(want_loss))
    assert sorted(got) == sorted(want)
    for k, w in want.items():
        rel = float((got[k] - w).abs().max() / w.abs().max().clamp_min(1e-30))
        assert rel <= 1e-4, (k, rel)
    plan.close()


def test_bf16_gradients_match_oracle():
    sd = fixtures.make_unet_weights(attention=True, seed=0)
    esd = fixtures.make_encoder_weights()
    full, t, noise = _case(32)
    want_loss, want = _oracle(True, sd, esd, full, t, noise)
    plan, loss, got = _gpu_grads("bf16", True, sd, esd, full, t, noise)
    assert abs(loss - float(want_loss)) <= 1e-3 * abs(float(want_loss))
    num = den = 0.0
    for k, w in want.items():
        a, b = got[k].double(), w.double()
        num += float(((a - b) ** 2).sum())
        den += float((b ** 2).sum())
        cos = float((a * b).sum() / (a.norm() * b.norm()).clamp_min(1e-30))
        assert cos >= 0.995, (k, cos)
    assert (num / den) ** 0.5 <= 1e-2
    plan.close()


def test_bf16_ragged_batch_is_padded_with_zero_weight_samples():
    """the reference's DataLoader has no drop_last (utils/load_data.py:174), so the last batch of an epoch is ragged.
    The raw C ABI still refuses a batch that is not a multiple of the tile granularity, but says how to declare padding
    (spdm_train_set_valid); DenoisePlan.train_fwd_bwd pads with zero-weight samples: loss and gradients are those of the real
    samples -- checked against the fp32 path on exactly the real samples."""
    import ctypes
    import state_policy_diffusionmodel_b200 as spdm
    from state_policy_diffusionmodel_b200._lib import SpdmError
    sd = fixtures.make_unet_weights(attention=False, seed=0)
    esd = fixtures.make_encoder_weights()
    full, t, noise = _case(5)
    plan32, loss32, g32 = _gpu_grads("fp32", False, sd, esd, full, t, noise)
    plan32.close()
    B = 5
    plan = spdm.DenoisePlan(attention=False, precision="bf16", batch_max=32, inpaint_rows=1)
    assert plan.batch_multiple == 32
    plan.enable_training(_named(sd, esd))
    ac = RefDDPMScheduler(num_train_timesteps=1000, beta_schedule="linear", clip_sample=False).alphas_cumprod
    obs = {k: v[:, :10] for k, v in full.items()}
    inp = torch.cat([obs["position"][:, -1:], obs["action"][:, -1:]], dim=-1)
    vec = torch.cat([inp.unsqueeze(1), torch.cat([full["position"][:, 10:], full["action"][:, 10:]], dim=-1).unsqueeze(1)], dim=2)
    loss = plan.train_fwd_bwd(obs["image"], obs["position"], obs["action"], obs["velocity"], vec, noise, t, ac ** 0.5, (1 - ac) ** 0.5,
                              inpaint=inp.reshape(B, -1))
    torch.cuda.synchronize()
    assert abs(float(loss) - loss32) <= 2e-3 * abs(loss32)
    num = den = 0.0
    for k, b in g32.items():
        a, b = plan.grad_view(k).detach().cpu().double(), b.double()
        num += float(((a - b) ** 2).sum())
        den += float((b ** 2).sum())
    assert (num / den) ** 0.5 <= 2e-2   # 5 samples: bf16 rounding averages over far fewer terms than in the batch-32 test
    # a full batch afterwards resets the padding declaration
    full32, t32, n32 = _case(32, seed=9)
    obs = {k: v[:, :10] for k, v in full32.items()}
    inp = torch.cat([obs["position"][:, -1:], obs["action"][:, -1:]], dim=-1)
    vec = torch.cat([inp.unsqueeze(1), torch.cat([full32["position"][:, 10:], full32["action"][:, 10:]], dim=-1).unsqueeze(1)], dim=2)
    la = plan.train_fwd_bwd(obs["image"], obs["position"], obs["action"], obs["velocity"], vec, n32, t32, ac ** 0.5, (1 - ac) ** 0.5,
                            inpaint=inp.reshape(32, -1))
    assert torch.isfinite(la).all() and plan._valid == 0
    # the raw ABI refuses an undeclared ragged batch loudly
    x = torch.zeros(8, device="cuda")
    rc = plan.lib.spdm_train_fwd_bwd(plan._h, *([ctypes.c_void_p(x.data_ptr())] * 11), 3, None)
    assert rc < 0 and b"multiple of 32" in plan.lib.spdm_last_error()
    assert SpdmError is not None
    plan.close()


@pytest.mark.parametrize("set_to_none", [True, False])
def test_lightning_order_step_zero_grad_backward_updates_weights(set_to_none):
    """Lightning's automatic optimization (and the common plain loop) runs training_step -> optimizer.zero_grad()
    -> loss.backward() -> clip -> optimizer.step().  The gradients of the native step must survive the zero_grad and reach the
    optimizer: the weights move exactly as with the order zero_grad -> step -> backward."""
    attention = False
    full, t, noise = _case(3)
    dev = {k: v.cuda() for k, v in full.items()}
    results = []
    for order in ("lightning", "classic"):
        m = _module("fp32", attention).train()
        opt = m.configure_optimizers()["optimizer"]
        before = {k: p.detach().clone() for k, p in m.named_trainable().items()}
        for it in range(2):     # second iteration: .grad tensors of the first one are still attached
            if order == "classic":
                opt.zero_grad(set_to_none=set_to_none)
            loss = m.process_single_batch(dev, t=t.cuda(), noise=noise.cuda())
            if order == "lightning":
                opt.zero_grad(set_to_none=set_to_none)
            loss.backward()
            assert all(p.grad is not None for p in m.parameters())
            torch.nn.utils.clip_grad_norm_(m.parameters(), 0.5)
            opt.step()
        after = {k: p.detach().clone() for k, p in m.named_trainable().items()}
        moved = sum(float((after[k] - before[k]).abs().sum()) for k in before)
        assert moved > 0, "no parameter was updated"
        results.append(after)
    # The two runs are not bit-identical: weight / bias gradients are accumulated with atomics whose order varies from launch to launch,
    # and Adam's normalised step turns last-bit gradient noise into up to a few 1e-7 per weight after two steps (measured over 12
    # repetitions: 0.2-0.8 of a 1e-7 + 1e-5 |w| bound).  A lost gradient would show as ~1e-4 (lr) per weight: 1e-6 + 1e-4 |w| separates them.
    for k in results[0]:
        torch.testing.assert_close(results[0][k], results[1][k], rtol=1e-4, atol=1e-6)


def test_gradient_accumulation_adds():
    """Two backward() calls without zero_grad accumulate (autograd semantics), one scaled by 0.5 via (0.5 * loss).backward()."""
    full, t, noise = _case(3)
    dev = {k: v.cuda() for k, v in full.items()}
    m = _module("fp32", False).train()
    loss = m.process_single_batch(dev, t=t.cuda(), noise=noise.cuda())
    loss.backward()
    g1 = {k: p.grad.detach().clone() for k, p in m.named_trainable().items()}
    loss = m.process_single_batch(dev, t=t.cuda(), noise=noise.cuda())
    (0.5 * loss).backward()
    for k, p in m.named_trainable().items():
        torch.testing.assert_close(p.grad, 1.5 * g1[k], rtol=1e-4, atol=1e-9)


def test_module_position_only_training_and_validation():
    """BASELINE configs[0] geometry (prediction_dim = 2) through the module surface: process_single_batch in training
    and in eval mode; the training loss equals the eval-mode (forward-only) loss on the same draws."""
    import state_policy_diffusionmodel_b200 as spdm
    torch.manual_seed(0)
    m = spdm.Diffusion_DDPM(noise_steps=1000, obs_horizon=10, pred_horizon=30, observation_dim=135, prediction_dim=2,
                            model='UNet_FilmnoAttention', inpaint_horizon=1).cuda()
    m.configure(precision="fp32")
    full, t, _ = _case(3)
    dev = {k: v.cuda() for k, v in full.items()}
    noise = torch.randn((3, 1, 31, 2), device="cuda")
    m.train()
    loss = m.process_single_batch(dev, t=t.cuda(), noise=noise)
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    m.eval()
    with torch.no_grad():
        loss_eval = m.process_single_batch(dev, t=t.cuda(), noise=noise)
    assert abs(float(loss) - float(loss_eval)) <= 1e-5 * abs(float(loss_eval))
    out = m.sample({k: v[:, :10].clone() for k, v in dev.items()})
    assert tuple(out.shape) == (1, 1, 31, 2)


def test_fused_optimizer_state_survives_plan_rebuild_and_logs_lr():
    """ADVICE r1 (low + medium): a larger batch rebuilds the training plan -- Adam moments and the step count carry over; the
    training step logs 'lr' (train.py's EarlyStopping monitors it) and validation_step(batch, 0) runs validate()."""
    m = _module("bf16", False).train()
    logged = {}
    m.log = lambda name, value, **kw: logged.__setitem__(name, value)
    full, _, _ = _case(32, seed=5)
    dev = {k: v.cuda() for k, v in full.items()}
    torch.manual_seed(0)
    m.training_step(dev, 0)
    m.optimizer_step()
    m.training_step(dev, 1)
    m.optimizer_step()
    assert logged["lr"] == m.lr and "train_loss" in logged
    st = m._tplan.optimizer_state_dict()
    assert st["step"] == 2
    full64, _, _ = _case(64, seed=6)
    m.training_step({k: v.cuda() for k, v in full64.items()}, 2)          # batch_max 32 -> 64: new plan
    st2 = m._tplan.optimizer_state_dict()
    assert st2["step"] == 2
    for k in st["exp_avg"]:
        assert torch.equal(st["exp_avg"][k], st2["exp_avg"][k]) and torch.equal(st["exp_avg_sq"][k], st2["exp_avg_sq"][k])
    m.optimizer_step()
    assert m._tplan.adam_steps == 3
    m.eval()
    with torch.no_grad():
        m.validation_step(dev, 0)
    assert "val_loss" in logged and tuple(m.last_validation["prediction"].shape) == (1, 1, 31, 5)


def test_clip_and_adam_match_oracle():
    """spdm_adam_step on synthetic flat gradients vs clip_grad_norm_ + torch.optim.Adam restated in oracle/train_ref.py
    (pinned to the real optimizer by tests/golden/train_grads.npz), three consecutive steps, with and without clipping."""
    import state_policy_diffusionmodel_b200 as spdm
    sd = fixtures.make_unet_weights(attention=False, seed=0)
    esd = fixtures.make_encoder_weights()
    plan = spdm.DenoisePlan(attention=False, precision="fp32", batch_max=1, inpaint_rows=1)
    plan.enable_training(_named(sd, esd))
    g = torch.Generator().manual_seed(5)
    n = plan.train_total
    p = {"w": plan.params_flat.detach().cpu().clone()}
    m = {"w": torch.zeros(n)}
    v = {"w": torch.zeros(n)}
    for step, (scale, max_norm) in enumerate([(1e-3, 0.5), (1e-6, 0.5), (1e-2, 0.0)], start=1):
        grads = scale * torch.randn(n, generator=g)
        plan.grads_flat.copy_(grads)
        plan.adam_step(lr=1e-3, max_norm=max_norm, sync=False)
        gd = {"w": grads}
        if max_norm > 0:
            total, gd = train_ref.clip_grad_norm(gd, max_norm)
        p, m, v = train_ref.adam_step(p, gd, m, v, step, lr=1e-3)
        torch.cuda.synchronize()
        torch.testing.assert_close(plan.params_flat.cpu(), p["w"], rtol=1e-5, atol=2e-7)
        torch.testing.assert_close(plan.grads_flat.cpu(), gd["w"], rtol=1e-5, atol=1e-12)  # clipped in place, like torch
    plan.close()


def _module(precision, attention=True, seed=0):
    import state_policy_diffusionmodel_b200 as spdm
    m = spdm.Diffusion_DDPM(noise_steps=1000, obs_horizon=10, pred_horizon=30, observation_dim=135, prediction_dim=5,
                            model='UNet_Film' if attention else 'UNet_FilmnoAttention', inpaint_horizon=1, learning_rate=1e-4)
    m.noise_estimator.load_state_dict(fixtures.make_unet_weights(attention=attention, seed=seed), strict=True)
    m.vision_encoder.load_state_dict(fixtures.make_encoder_weights(), strict=True)
    m = m.cuda()
    m.configure(precision=precision)
    return m


def test_module_training_step_is_a_drop_in():
    """Diffusion_DDPM.training_step -> loss.backward() -> torch.optim.Adam.step() exactly as Lightning drives the reference
    (ddpm:92-125), against the oracle's gradients and update; then the fused optimizer_step on a twin module."""
    attention = False
    sd = fixtures.make_unet_weights(attention=attention, seed=0)
    esd = fixtures.make_encoder_weights()
    full, t, noise = _case(3)
    want_loss, want = _oracle(attention, sd, esd, full, t, noise)
    m = _module("fp32", attention).train()
    opt = m.configure_optimizers()["optimizer"]
    opt.zero_grad()
    loss = m.process_single_batch({k: v.cuda() for k, v in full.items()}, t=t.cuda(), noise=noise.cuda())
    assert loss.requires_grad and loss.dim() == 0
    loss.backward()
    named = m.named_trainable()
    assert sorted(named) == sorted(want)
    for k, p in named.items():
        rel = float((p.grad.cpu() - want[k]).abs().max() / want[k].abs().max().clamp_min(1e-30))
        assert rel <= 1e-4, (k, rel)
    before = {k: p.detach().cpu().clone() for k, p in named.items()}
    total = torch.nn.utils.clip_grad_norm_(m.parameters(), 0.5)
    opt.step()
    want_total, clipped = train_ref.clip_grad_norm(want, 0.5)
    assert abs(float(total) - float(want_total)) <= 1e-4 * float(want_total)
    # twin module, fused clip + Adam
    m2 = _module("fp32", attention).train()
    m2.process_single_batch({k: v.cuda() for k, v in full.items()}, t=t.cuda(), noise=noise.cuda())
    m2.optimizer_step(lr=1e-4, gradient_clip_val=0.5)
    torch.cuda.synchronize()
    moved = 0.0
    for k, p in named.items():
        p2 = m2.named_trainable()[k]
        d1, d2 = (p.detach().cpu() - before[k]), (p2.detach().cpu() - before[k])
        moved += float(d1.abs().sum())
        # Adam's first update is lr * g / (|g| + eps): identical wherever the two gradient computations agree in sign
        frac_bad = float(((d1 - d2).abs() > 2e-5).float().mean())
        assert frac_bad < 1e-3, (k, frac_bad)
    assert moved > 0
    # the sampling path sees the updated weights (the inference plan reloads on the version bump)
    m.eval()
    batch = fixtures.make_batch(2, seed=11)
    torch.manual_seed(0)
    out_after = m.sample({k: v.clone() for k, v in batch.items()})
    assert torch.isfinite(out_after).all()


def test_bf16_loss_decreases():
    """Twenty native bf16 training steps (fused clip + Adam) on one fixed synthetic batch: the loss must fall."""
    m = _module("bf16", True).train()
    B = 32
    full, _, _ = _case(B, seed=5)
    batch = {k: v.cuda() for k, v in full.items()}
    torch.manual_seed(0)
    losses = []
    for _ in range(20):
        loss = m.training_step(batch, 0)
        m.optimizer_step(lr=1e-4, gradient_clip_val=0.5)
        losses.append(float(loss))
    assert all(l == l for l in losses)
    assert sum(losses[-5:]) / 5 < 0.9 * sum(losses[:5]) / 5, losses


@pytest.mark.parametrize("attention,rows,dim,B,precision", [
    (False, 31, 5, 32, "bf16"),   # no-attention variant on the tensor-core path
    (True, 61, 5, 2, "fp32"),     # 2x prediction horizon: 64x8 maps, GroupNorm-backward clusters of 2 CTAs, L = 512 attention
    (False, 31, 2, 3, "fp32"),    # position-only prediction (prediction_dim = 2, pads lw = uw = 3)
    (True, 61, 5, 32, "bf16"),    # 2x horizon on the tensor-core path
])
def test_training_gradients_other_geometries(attention, rows, dim, B, precision):
    """Other BASELINE geometries (configs[0] position-only, configs[4] 2x horizon, no-attention variant): same tolerances."""
    import state_policy_diffusionmodel_b200 as spdm
    sd = fixtures.make_unet_weights(attention=attention, seed=0)
    esd = fixtures.make_encoder_weights()
    g = torch.Generator().manual_seed(91)
    T = 10 + rows - 1
    full = {"image": torch.rand((B, T, 3, 96, 96), generator=g), "position": 0.3 * torch.randn((B, T, 2), generator=g),
            "velocity": 2 * torch.rand((B, T, 2), generator=g) - 1, "action": 2 * torch.rand((B, T, 3), generator=g) - 1}
    t = torch.randint(0, 1000, (B,), generator=g)
    noise = torch.randn((B, 1, rows, dim), generator=g)
    sched = RefDDPMScheduler(num_train_timesteps=1000, beta_schedule="linear", clip_sample=False)
    obs = {k: v[:, :10] for k, v in full.items()}
    pred = {k: v[:, 10:] for k, v in full.items()}
    if dim == 5:
        inp = torch.cat([obs["position"][:, -1:], obs["action"][:, -1:]], dim=-1)
        x0 = torch.cat([pred["position"], pred["action"]], dim=-1)
    else:  # position-only variant (the one models/diffusion_ddim.py:75-86 keeps commented out)
        inp = obs["position"][:, -1:]
        x0 = pred["position"]
    vec = torch.cat([inp.unsqueeze(1), x0.unsqueeze(1)], dim=2)
    # oracle: the same restated forward (unet_ref / add_noise / inpaint), autograd for the gradients
    from oracle import sampler_ref, unet_ref
    sd_g = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    esd_g = {k: v.clone().requires_grad_(True) for k, v in esd.items()}
    torch.set_num_threads(os.cpu_count() or 1)
    cond = unet_ref.obs_cond(esd_g, obs).unsqueeze(1)
    x_noisy = sampler_ref.add_constraints(sched.add_noise(vec, noise, t), inp.unsqueeze(1), 1)
    est = unet_ref.unet_forward(sd_g, x_noisy, t, cond, attention=attention)
    want_loss = torch.nn.functional.mse_loss(noise, est)
    names = list(sd_g) + ["vision_encoder." + k for k in esd_g]
    gs = torch.autograd.grad(want_loss, list(sd_g.values()) + list(esd_g.values()))
    want = dict(zip(names, gs))
    plan = spdm.DenoisePlan(attention=attention, precision=precision, batch_max=B, rows=rows, dim=dim, inpaint_rows=1)
    plan.enable_training(_named(sd, esd))
    ac = sched.alphas_cumprod
    loss = plan.train_fwd_bwd(obs["image"], obs["position"], obs["action"], obs["velocity"], vec, noise, t, ac ** 0.5, (1 - ac) ** 0.5,
                              inpaint=inp.reshape(B, -1))
    torch.cuda.synchronize()
    assert abs(float(loss) - float(want_loss)) <= (1e-5 if precision == "fp32" else 2e-3) * abs(float(want_loss))
    num = den = 0.0
    for k, w in want.items():
        a, b = plan.grad_view(k).detach().cpu().double(), w.double()
        if precision == "fp32":
            rel = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
            assert rel <= 1e-4, (k, rel)
        else:
            cos = float((a * b).sum() / (a.norm() * b.norm()).clamp_min(1e-30))
            # 0.99 here: the 8x2 / 4x1-level tensors see only 16 / 4 pixels per sample, so at batch 32 the bf16 rounding of the
            # activation gradients averages out less than in the headline test (worst measured: 0.9945)
            assert cos >= 0.99, (k, cos)
        num += float(((a - b) ** 2).sum())
        den += float((b ** 2).sum())
    assert (num / den) ** 0.5 <= (1e-4 if precision == "fp32" else 1e-2)
    plan.close()


def test_strided_observation_window_matches_contiguous():
    """bf16 path: the observation window passed as a strided view of the full recording (what `batch['image'][:, :10]` is)
    gives the same gradients as a contiguous copy of it."""
    import state_policy_diffusionmodel_b200 as spdm
    sd = fixtures.make_unet_weights(attention=False, seed=0)
    esd = fixtures.make_encoder_weights()
    B = 32
    full, t, noise = _case(B, seed=17)
    plan = spdm.DenoisePlan(attention=False, precision="bf16", batch_max=B, inpaint_rows=1)
    plan.enable_training(_named(sd, esd))
    ac = RefDDPMScheduler(num_train_timesteps=1000, beta_schedule="linear", clip_sample=False).alphas_cumprod
    dev = {k: v.cuda() for k, v in full.items()}
    obs = {k: v[:, :10] for k, v in dev.items()}
    inp = torch.cat([obs["position"][:, -1:], obs["action"][:, -1:]], dim=-1)
    vec = torch.cat([inp.unsqueeze(1), torch.cat([dev["position"][:, 10:], dev["action"][:, 10:]], dim=-1).unsqueeze(1)], dim=2)
    outs = []
    for image in (obs["image"], obs["image"].contiguous()):
        assert image.is_contiguous() == (len(outs) == 1)
        loss = plan.train_fwd_bwd(image, obs["position"], obs["action"], obs["velocity"], vec, noise, t, ac ** 0.5, (1 - ac) ** 0.5,
                                  inpaint=inp.reshape(B, -1))
        torch.cuda.synchronize()
        outs.append((float(loss), plan.grads_flat.clone()))
    assert abs(outs[0][0] - outs[1][0]) <= 1e-6 * abs(outs[1][0])
    rel = float((outs[0][1] - outs[1][1]).norm() / outs[1][1].norm())
    assert rel <= 1e-4, rel   # fp32 atomics only
    plan.close()
