"""oracle/make_golden.py — TEST INFRASTRUCTURE, build-container only.

Runs the UNMODIFIED reference modules from /root/reference (U-Nets as-is; Lightning wrappers behind
oracle/shim.py) on the deterministic fixtures of oracle/fixtures.py and writes small golden
input/output vectors to tests/golden/*.npz.  Usage:  python -m oracle.make_golden
"""
import os
import sys

import numpy as np
import torch

from . import fixtures, shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def _tap_summary(t):
    f = t.detach().flatten(1)
    return torch.stack([f.mean(1), f.std(1), f.abs().max(1).values, f[:, 0], f[:, -1]], dim=1)


def golden_unet(attention, t_vals, B, seed, name, rows=31, dim=5):
    if attention:
        from models.Unet_FiLmLayer import UNet_Film as Net
    else:
        from models.Unet_FiLmLayer_noAttention import UNet_Film_noAttention as Net
    sd = fixtures.make_unet_weights(attention=attention, seed=seed)
    net = Net(in_channels=1, out_channels=1, noise_steps=1000, global_cond_dim=1350, time_dim=256).eval()
    net.load_state_dict(sd, strict=True)  # pins key names + shapes
    g = torch.Generator().manual_seed(seed + 100)
    x = torch.rand((B, 1, rows, dim), generator=g)
    y = torch.randn((B, 1, 10, 135), generator=g)
    t = torch.tensor(t_vals, dtype=torch.long)
    taps = {}
    hooks = []
    for modname in ("inc", "down1", "down2", "down3", "bot1", "bot2", "bot3", "up1", "up2", "up3",
                    "sa1", "sa2", "sa3", "sa4", "sa5", "sa6"):
        if hasattr(net, modname):
            hooks.append(getattr(net, modname).register_forward_hook(
                lambda m, i, o, n=modname: taps.__setitem__(n, o.detach().clone())))
    with torch.no_grad():
        out = net(x, t, y)
    for h in hooks:
        h.remove()
    with torch.no_grad():
        out_nocond = net(x, t, None)
    d = {"x": x, "y": y, "t": t, "out": out, "out_nocond": out_nocond, "seed": seed, "attention": int(attention)}
    for k, v in taps.items():
        d["tap_" + k] = _tap_summary(v)
    # full activation of two layers to pin layout-sensitive ops (upsample/concat, attention)
    d["full_down1"] = taps["down1"]
    d["full_up1"] = taps["up1"]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(d))
    print(name, "out", tuple(out.shape), float(out.abs().mean()))


def golden_encoder():
    from models.encoder.autoencoder import Autoencoder
    esd = fixtures.make_encoder_weights()
    enc = Autoencoder().encoder.eval()
    enc.load_state_dict(esd, strict=True)
    g = torch.Generator().manual_seed(5)
    img = torch.rand((4, 3, 96, 96), generator=g)
    with torch.no_grad():
        out = enc(img)
    np.savez_compressed(os.path.join(OUT, "encoder.npz"), **_np({"img_seed": 5, "out": out}))
    print("encoder", tuple(out.shape))


def _wrapper(kind, model_name, noise_steps, pred_dim, attention, seed):
    from models.diffusion_ddpm import Diffusion_DDPM
    from models.diffusion_ddim import Diffusion_DDIM
    from oracle.schedulers import RefDDIMScheduler
    cls = Diffusion_DDPM if kind == "ddpm" else Diffusion_DDIM
    m = cls(noise_steps=noise_steps if kind == "ddpm" else 1000, obs_horizon=10, pred_horizon=30, observation_dim=135,
            prediction_dim=pred_dim, model=model_name, inpaint_horizon=1).eval()
    m.noise_estimator.load_state_dict(fixtures.make_unet_weights(attention=attention, seed=seed), strict=True)
    m.vision_encoder.load_state_dict(fixtures.make_encoder_weights(), strict=True)
    if kind == "ddim":  # generate.py:28-35
        m.noise_scheduler = RefDDIMScheduler(num_train_timesteps=noise_steps, beta_schedule="linear",
                                             clip_sample=False, prediction_type="epsilon")
        m.noise_steps = noise_steps
    return m


def golden_sample(kind, model_name, attention, noise_steps, pred_dim, name, seed=0, B=2):
    m = _wrapper(kind, model_name, noise_steps, pred_dim, attention, seed)
    batch = fixtures.make_batch(B, seed=4321)
    if pred_dim == 2:
        # position-only prediction: the committed wrapper still concatenates pos+act for the inpaint vector
        # (models/diffusion_ddpm.py:340-348; the position-only variant is commented out at diffusion_ddim.py:75-86),
        # so a 2-wide x_t cannot take a 5-wide inpaint row.  Restate the commented-out variant for this config.
        m.prepare_inpaint_vectors = lambda ob: ob["position"][:, -m.inpaint_horizon:, :]
    obs = m.prepare_observation_batch(batch)
    with torch.no_grad():
        cond = m.prepare_obs_cond_vectors(obs)
        inp = m.prepare_inpaint_vectors(obs)
    torch.manual_seed(1000 + seed)
    hist = m.sample(batch={k: v.clone() for k, v in obs.items()}, option="sample_history")
    # replay the RNG stream the wrapper consumed: rand for x_T, then randn per step with t>0 (DDPM only)
    torch.manual_seed(1000 + seed)
    x_T = torch.rand(1, 1, 31, pred_dim)
    noises = []
    m.noise_scheduler.set_timesteps(noise_steps)
    for t in m.noise_scheduler.timesteps:
        if kind == "ddpm" and int(t) > 0:
            noises.append(torch.randn(1, 1, 31, pred_dim))
        else:
            noises.append(torch.zeros(1, 1, 31, pred_dim))
    assert torch.equal(hist[0], x_T)
    torch.manual_seed(1000 + seed)
    final = m.sample(batch={k: v.clone() for k, v in obs.items()})
    assert torch.allclose(final, hist[-1], atol=0, rtol=0)
    d = {"batch_seed": 4321, "B": B, "obs_cond": cond, "inpaint": inp, "x_T": x_T, "noise": torch.stack(noises),
         "history": torch.stack(hist), "timesteps": m.noise_scheduler.timesteps, "unet_seed": seed,
         "attention": int(attention), "noise_steps": noise_steps, "pred_dim": pred_dim}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(d))
    print(name, "final", float(hist[-1].abs().mean()), "steps", len(hist) - 1)


def golden_validate_and_train(seed=0):
    m = _wrapper("ddpm", "UNet_Film", 12, 5, True, seed)
    B = 3
    g = torch.Generator().manual_seed(777)
    full = {"image": torch.rand((B, 40, 3, 96, 96), generator=g), "position": 0.3 * torch.randn((B, 40, 2), generator=g),
            "velocity": 2 * torch.rand((B, 40, 2), generator=g) - 1, "action": 2 * torch.rand((B, 40, 3), generator=g) - 1}
    torch.manual_seed(55)
    x0, obs, inp = m.validate(full)
    torch.manual_seed(56)
    with torch.no_grad():
        loss = m.process_single_batch(full)
    torch.manual_seed(56)
    t = torch.randint(0, m.noise_steps, (B,)).long()
    noise = torch.randn(B, 1, 31, 5)
    torch.manual_seed(55)
    x_T = torch.rand(1, 1, 31, 5)
    noises = [torch.randn(1, 1, 31, 5) if int(tt) > 0 else torch.zeros(1, 1, 31, 5) for tt in range(11, -1, -1)]
    d = {"full_seed": 777, "validate_x0": x0, "validate_inpaint": inp, "validate_x_T": x_T,
         "validate_noise": torch.stack(noises), "train_loss": loss, "train_t": t, "train_noise": noise}
    np.savez_compressed(os.path.join(OUT, "validate_train.npz"), **_np(d))
    print("validate/train loss", float(loss))


def golden_train_grads(seed=0):
    """loss.backward() + clip_grad_norm_(0.5) + one torch.optim.Adam step on the reference's own Diffusion_DDPM
    (U-Net + vision encoder), injected t / noise: per-tensor fingerprints of the gradients and of the updated weights."""
    from . import train_ref
    for attention, model_name, name in ((True, "UNet_Film", "train_grads"), (False, "UNet_FilmnoAttention", "train_grads_noattn")):
        m = _wrapper("ddpm", model_name, 1000, 5, attention, seed).train()
        B = 3
        g = torch.Generator().manual_seed(778)
        full = {"image": torch.rand((B, 40, 3, 96, 96), generator=g), "position": 0.3 * torch.randn((B, 40, 2), generator=g),
                "velocity": 2 * torch.rand((B, 40, 2), generator=g) - 1, "action": 2 * torch.rand((B, 40, 3), generator=g) - 1}
        t = torch.tensor([3, 500, 997], dtype=torch.long)
        noise = torch.randn((B, 1, 31, 5), generator=g)
        # process_single_batch draws t then noise from the global RNG (ddpm:158-161): patch the two draws
        real_randint, real_randn_like = torch.randint, torch.randn_like
        torch.randint = lambda *a, **k: t.clone()
        torch.randn_like = lambda x, *a, **k: noise.clone()
        try:
            loss = m.process_single_batch(full)
        finally:
            torch.randint, torch.randn_like = real_randint, real_randn_like
        opt = torch.optim.Adam(m.parameters(), lr=1e-4)
        opt.zero_grad()
        loss.backward()
        named = [(k[len("noise_estimator."):] if k.startswith("noise_estimator.") else k, p) for k, p in m.named_parameters()]
        d = {"full_seed": 778, "t": t, "noise": noise, "loss": loss.detach(), "names": np.array([k for k, _ in named])}
        d["grad_fp"] = torch.stack([train_ref.summary(p.grad if p.grad is not None else torch.zeros_like(p)) for _, p in named])
        total = torch.nn.utils.clip_grad_norm_(m.parameters(), 0.5)
        opt.step()
        d["total_norm"] = total
        d["param_fp_after"] = torch.stack([train_ref.summary(p) for _, p in named])
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(d))
        print(name, "loss", float(loss), "total_norm", float(total), "tensors", len(named))


def golden_dataset(seed=0):
    """the reference's own CarRacingDataset (utils/load_data.py:11-144) on the synthetic dataset of
    oracle/data_ref.py -- only its zarr reader is replaced -- and utils/data_utils.unnormalize_position."""
    from . import data_ref
    shim.install()
    import utils.load_data as ld
    from utils.data_utils import unnormalize_position
    raw = data_ref.make_synthetic_dataset(seed)
    img = data_ref.image_chw_float(raw["img_u8"])

    def _load(self, dataset_path):
        return img, {"position": raw["position"], "velocity": raw["velocity"], "action": raw["action"]}, raw["episode_ends"]

    class DS(ld.CarRacingDataset):           # training dataset: computes the statistics, items are dicts
        _load_data = _load

    class DSI(ld.CarRacingDatasetForInference):   # inference dataset: given statistics, items carry translation / start / end
        _load_data = _load

    out = {"seed": seed}
    for tag, (obs_h, pred_h, step) in {"a": (3, 4, 2), "b": (2, 3, 1)}.items():
        ds = DS("unused", pred_h, obs_h, None, step_size=step)
        dsi = DSI("unused", pred_h, obs_h, ds.stats, step_size=step)
        assert dsi.indices == ds.indices
        idxs = list(range(0, len(ds), max(1, len(ds) // 5)))[:6]
        train_items = [ds[i] for i in idxs]
        items = [dsi[i] for i in idxs]
        for ti, it in zip(train_items, items):   # both classes normalise a window identically
            for k in ("position", "velocity", "action", "image"):
                assert np.array_equal(np.asarray(ti[k]), np.asarray(it[0][k])), k
        out[tag + "_cfg"] = np.asarray([obs_h, pred_h, step])
        out[tag + "_indices"] = np.asarray(ds.indices)
        out[tag + "_idxs"] = np.asarray(idxs)
        out[tag + "_pos_stats"] = np.asarray([ds.stats["position"]["min"], ds.stats["position"]["max"]])
        out[tag + "_vel_stats"] = np.stack([ds.stats["velocity"]["min"], ds.stats["velocity"]["max"]])
        out[tag + "_act_stats"] = np.stack([ds.stats["action"]["min"], ds.stats["action"]["max"]])
        out[tag + "_position"] = np.stack([it[0]["position"] for it in items])
        out[tag + "_velocity"] = np.stack([it[0]["velocity"] for it in items])
        out[tag + "_action"] = np.stack([it[0]["action"] for it in items])
        out[tag + "_translation"] = np.stack([it[1] for it in items])
        out[tag + "_start_end"] = np.asarray([[it[2], it[3]] for it in items])
        im = np.stack([it[0]["image"] for it in items]).astype(np.float64)      # (n, T, 3, H, W): fingerprints only
        out[tag + "_image_fp"] = np.stack([im.sum(axis=(2, 3, 4)), im[:, :, 0, 0, 0], im[:, :, 1, 5, 7], im[:, :, 2, -1, -1],
                                          (im * np.arange(im.shape[-1])).sum(axis=(2, 3, 4))], axis=-1)
        out[tag + "_unnorm"] = np.stack([unnormalize_position(it[0]["position"], it[1], ds.stats["position"]) for it in items])
        print("dataset", tag, len(ds), "windows", out[tag + "_position"].shape)
    np.savez_compressed(os.path.join(OUT, "dataset.npz"), **out)


def golden_simple_unet(seed=7):
    """The reference's legacy UNet (models/simple_Unet.py:260-339) in eval mode on the fixture weights: output + three skip taps."""
    from models.simple_Unet import UNet
    sd = fixtures.make_simple_unet_weights(seed=seed)
    net = UNet(in_channels=1, out_channels=1, noise_steps=1000, global_cond_dim=1350, time_dim=256).eval()
    ref_table = net.pos_encoding.pos_encoding.clone()
    net.load_state_dict(sd, strict=True)     # pins key names, shapes and the buffer
    assert torch.equal(ref_table, sd["pos_encoding.pos_encoding"]), "PositionalEncoding table restatement differs from the reference buffer"
    g = torch.Generator().manual_seed(seed + 100)
    B = 3
    x = torch.rand((B, 1, 31, 5), generator=g)
    y = torch.randn((B, 1, 10, 135), generator=g)
    t = torch.tensor([999, 500, 3], dtype=torch.long)
    taps = {}
    hooks = [getattr(net, m).register_forward_hook(lambda mod, i, o, n=m: taps.__setitem__(n, o.detach().clone())) for m in ("input_conv", "down3", "up1", "up3")]
    with torch.no_grad():
        out = net(x, t, y)
        out_t1 = net(x, torch.tensor([17]), y)    # a single timestep broadcast over the batch (models/diffusion_ddpm.py:209)
    for h in hooks:
        h.remove()
    d = {"x": x, "y": y, "t": t, "out": out, "out_t1": out_t1, "seed": seed}
    for k, v in taps.items():
        d["tap_" + k] = _tap_summary(v)
    np.savez_compressed(os.path.join(OUT, "simple_unet.npz"), **_np(d))
    print("simple_unet out", tuple(out.shape), float(out.abs().mean()))


def golden_resnet18gn(seed=2):
    """The reference's `VisionEncoder()` (models/Unet_FiLmLayer.py:383-386: torchvision resnet18, fc = Identity, BatchNorm ->
    GroupNorm(C // 16)) on the fixture weights: 512 features of 5 frames + the activation after layer1 / layer3."""
    from models.Unet_FiLmLayer import VisionEncoder
    sd = fixtures.make_resnet_weights(seed=seed)
    net = VisionEncoder().eval()
    net.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(seed + 100)
    img = torch.rand((5, 3, 96, 96), generator=g)
    taps = {}
    hooks = [getattr(net, m).register_forward_hook(lambda mod, i, o, n=m: taps.__setitem__(n, o.detach().clone())) for m in ("layer1", "layer3")]
    with torch.no_grad():
        out = net(img)
    for h in hooks:
        h.remove()
    # (the frames are torch.rand((5, 3, 96, 96)) under the CPU generator seeded seed + 100: regenerated by the tests, not stored)
    d = {"img_fingerprint": _tap_summary(img), "out": out, "seed": seed, "tap_layer1": _tap_summary(taps["layer1"]),
         "tap_layer3": _tap_summary(taps["layer3"])}
    np.savez_compressed(os.path.join(OUT, "resnet18gn.npz"), **_np(d))
    print("resnet18gn out", tuple(out.shape), float(out.abs().mean()))


def golden_beta_schedules():
    """utils/schedulers.py:6-40 of the reference, imported as is (the functions are unbound "methods" reading self.device)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_utils_schedulers", os.path.join(shim.REFERENCE_ROOT, "utils", "schedulers.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    class Dev:
        device = torch.device("cpu")
    out = {}
    for steps in (10, 50, 100, 1000):
        out["linear_%d" % steps] = mod.linear_beta_schedule(Dev(), steps)
        out["linear_v2_%d" % steps] = mod.linear_beta_schedule_v2(Dev(), steps)
        out["cosine_%d" % steps] = mod.cosine_beta_schedule(Dev(), steps)
    out["cosine_100_s02_f64"] = mod.cosine_beta_schedule(Dev(), 100, s=0.02, dtype=torch.float64)
    np.savez_compressed(os.path.join(OUT, "beta_schedules.npz"), **_np(out))
    print("beta_schedules", sorted(out))


def main():
    os.makedirs(OUT, exist_ok=True)
    shim.install()
    torch.set_num_threads(8)
    golden_unet(True, [999, 500, 3], 3, 0, "unet_attn")
    golden_unet(False, [7], 2, 3, "unet_noattn")
    golden_unet(True, [40], 2, 4, "unet_attn_rows61", rows=61)
    golden_unet(False, [123], 2, 5, "unet_noattn_pos2", dim=2)
    golden_encoder()
    golden_sample("ddim", "UNet_Film", True, 10, 5, "sample_ddim10_attn")
    golden_sample("ddpm", "UNet_FilmnoAttention", False, 20, 2, "sample_ddpm20_noattn_pos2", seed=3)
    golden_validate_and_train()
    golden_train_grads()
    golden_dataset()
    golden_beta_schedules()
    golden_simple_unet()
    golden_resnet18gn()


if __name__ == "__main__":
    sys.exit(main())
