"""CPU: the oracle restatement vs golden vectors produced by the REAL reference modules
(oracle/make_golden.py, run in the build container against /root/reference)."""
import os

import numpy as np
import pytest
import torch

from oracle import fixtures, sampler_ref, unet_ref
from oracle.schedulers import RefDDPMScheduler

TOL = dict(rtol=1e-4, atol=2e-5)


def _load(golden_dir, name):
    return {k: torch.from_numpy(v) if v.ndim else v for k, v in np.load(os.path.join(golden_dir, name + ".npz")).items()}


def _summary(t):
    f = t.flatten(1)
    return torch.stack([f.mean(1), f.std(1), f.abs().max(1).values, f[:, 0], f[:, -1]], dim=1)


@pytest.mark.parametrize("name", ["unet_attn", "unet_noattn", "unet_attn_rows61", "unet_noattn_pos2"])
def test_unet_oracle_matches_reference(golden_dir, name):
    g = _load(golden_dir, name)
    attention = bool(int(g["attention"]))
    sd = fixtures.make_unet_weights(attention=attention, seed=int(g["seed"]))
    taps = {}
    with torch.no_grad():
        out = unet_ref.unet_forward(sd, g["x"], g["t"], g["y"], attention=attention, taps=taps)
        out_nc = unet_ref.unet_forward(sd, g["x"], g["t"], None, attention=attention)
    torch.testing.assert_close(out, g["out"], **TOL)
    torch.testing.assert_close(out_nc, g["out_nocond"], **TOL)
    torch.testing.assert_close(taps["down1"], g["full_down1"], **TOL)
    torch.testing.assert_close(taps["up1"], g["full_up1"], **TOL)
    # module-output summaries: reference hook names -> oracle tap names
    ren = {"inc": "x1", "bot1": "bot1", "bot2": "bot2", "bot3": "x5", "down1": "down1", "down2": "down2",
           "down3": "down3", "up1": "up1", "up2": "up2", "up3": "up3", "sa1": "x2", "sa2": "x3", "sa3": "x4",
           "sa4": "u1", "sa5": "u2", "sa6": "u3"}
    for k, v in g.items():
        if k.startswith("tap_"):
            torch.testing.assert_close(_summary(taps[ren[k[4:]]]), v, rtol=1e-3, atol=1e-4)


def test_encoder_oracle(golden_dir):
    g = _load(golden_dir, "encoder")
    esd = fixtures.make_encoder_weights()
    img = torch.rand((4, 3, 96, 96), generator=torch.Generator().manual_seed(int(g["img_seed"])))
    torch.testing.assert_close(unet_ref.encoder_forward(esd, img), g["out"], **TOL)


@pytest.mark.parametrize("name,kind", [("sample_ddim10_attn", "ddim"), ("sample_ddpm20_noattn_pos2", "ddpm")])
def test_sampling_loop_oracle(golden_dir, name, kind):
    g = _load(golden_dir, name)
    attention = bool(int(g["attention"]))
    pred_dim = int(g["pred_dim"])
    K = int(g["noise_steps"])
    sd = fixtures.make_unet_weights(attention=attention, seed=int(g["unet_seed"]))
    esd = fixtures.make_encoder_weights()
    batch = fixtures.make_batch(int(g["B"]), seed=int(g["batch_seed"]))
    cond = unet_ref.obs_cond(esd, batch)
    torch.testing.assert_close(cond, g["obs_cond"], **TOL)
    if pred_dim == 2:
        inp = batch["position"][:, -1:, :]
    else:
        inp = unet_ref.inpaint_vector(batch, 1)
    torch.testing.assert_close(inp, g["inpaint"], rtol=0, atol=0)
    sched = sampler_ref.make_scheduler(kind, K)
    hist = sampler_ref.sample_ref(sd, sched, K, g["x_T"], cond[0:1].unsqueeze(1), inp[0:1].unsqueeze(1), 1,
                                  attention=attention, noise=g["noise"], history=True)
    assert [int(t) for t in sched.timesteps] == [int(t) for t in g["timesteps"]]
    torch.testing.assert_close(torch.stack(hist), g["history"], rtol=1e-4, atol=5e-5)


def test_validate_and_training_forward_oracle(golden_dir):
    g = _load(golden_dir, "validate_train")
    sd = fixtures.make_unet_weights(attention=True, seed=0)
    esd = fixtures.make_encoder_weights()
    B = 3
    gen = torch.Generator().manual_seed(int(g["full_seed"]))
    full = {"image": torch.rand((B, 40, 3, 96, 96), generator=gen), "position": 0.3 * torch.randn((B, 40, 2), generator=gen),
            "velocity": 2 * torch.rand((B, 40, 2), generator=gen) - 1, "action": 2 * torch.rand((B, 40, 3), generator=gen) - 1}
    obs = {k: v[:, :10] for k, v in full.items()}
    cond = unet_ref.obs_cond(esd, obs)
    inp = unet_ref.inpaint_vector(obs, 1)
    sched = RefDDPMScheduler(num_train_timesteps=12, beta_schedule="linear", clip_sample=False)
    x0 = sampler_ref.sample_ref(sd, sched, 12, g["validate_x_T"], cond[0:1].unsqueeze(1), inp[0:1].unsqueeze(1), 1,
                                attention=True, noise=g["validate_noise"])
    torch.testing.assert_close(x0, g["validate_x0"], rtol=1e-4, atol=5e-5)
    sched = RefDDPMScheduler(num_train_timesteps=12, beta_schedule="linear", clip_sample=False)
    with torch.no_grad():
        loss, _, _ = sampler_ref.training_forward_ref(sd, esd, sched, full, 10, 1, g["train_t"], g["train_noise"])
    torch.testing.assert_close(loss, torch.as_tensor(g["train_loss"]), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["train_grads", "train_grads_noattn"])
def test_training_gradients_and_adam_oracle(golden_dir, name):
    """oracle/train_ref.py (autograd over the restatement, clip, Adam) vs loss.backward() / clip_grad_norm_ /
    torch.optim.Adam.step() on the reference's own Diffusion_DDPM (oracle/make_golden.py::golden_train_grads)."""
    from oracle import train_ref
    raw = np.load(os.path.join(golden_dir, name + ".npz"))
    attention = name == "train_grads"
    sd = fixtures.make_unet_weights(attention=attention, seed=0)
    esd = fixtures.make_encoder_weights()
    B = 3
    gen = torch.Generator().manual_seed(int(raw["full_seed"]))
    full = {"image": torch.rand((B, 40, 3, 96, 96), generator=gen), "position": 0.3 * torch.randn((B, 40, 2), generator=gen),
            "velocity": 2 * torch.rand((B, 40, 2), generator=gen) - 1, "action": 2 * torch.rand((B, 40, 3), generator=gen) - 1}
    t, noise = torch.from_numpy(raw["t"]), torch.from_numpy(raw["noise"])
    sched = RefDDPMScheduler(num_train_timesteps=1000, beta_schedule="linear", clip_sample=False)
    loss, grads = train_ref.loss_and_grads(sd, esd, sched, full, 10, 1, t, noise, attention=attention)
    torch.testing.assert_close(loss, torch.as_tensor(raw["loss"]), rtol=1e-5, atol=1e-6)
    names = [str(n) for n in raw["names"]]
    assert sorted(names) == sorted(grads)  # exactly the tensors the reference hands to Adam (U-Net + vision encoder)
    got = torch.stack([train_ref.summary(grads[n]) for n in names])
    want = torch.from_numpy(raw["grad_fp"])
    scale = want[:, 2:3].clamp_min(1e-12)  # per-tensor l2 norm
    assert float(((got - want).abs() / scale).max()) < 2e-3
    total, clipped = train_ref.clip_grad_norm(grads, 0.5)
    torch.testing.assert_close(total, torch.as_tensor(raw["total_norm"]), rtol=1e-4, atol=1e-6)
    params = dict(sd)
    params.update({train_ref.ENC_PREFIX + k: v for k, v in esd.items()})
    zeros = {k: torch.zeros_like(v) for k, v in params.items()}
    new_p, _, _ = train_ref.adam_step(params, clipped, zeros, dict(zeros), 1, lr=1e-4)
    got_p = torch.stack([train_ref.summary(new_p[n]) for n in names])
    want_p = torch.from_numpy(raw["param_fp_after"])
    assert float(((got_p - want_p).abs() / want_p[:, 2:3].clamp_min(1e-12)).max()) < 1e-4


def test_dataset_oracle_vs_reference_golden(golden_dir):
    """oracle/data_ref.py against tests/golden/dataset.npz, which the reference's own CarRacingDataset /
    CarRacingDatasetForInference (utils/load_data.py) and unnormalize_position (utils/data_utils.py:35-40) produced."""
    import numpy as np
    from oracle import data_ref
    g = np.load(os.path.join(golden_dir, "dataset.npz"))
    raw = data_ref.make_synthetic_dataset(int(g["seed"]))
    data = {"image": data_ref.image_chw_float(raw["img_u8"]), "position": raw["position"], "velocity": raw["velocity"],
            "action": raw["action"]}
    for tag in ("a", "b"):
        obs_h, pred_h, step = (int(v) for v in g[tag + "_cfg"])
        ds = data_ref.RefWindowDataset(data, raw["episode_ends"], pred_h, obs_h, None, step)
        assert np.array_equal(np.asarray(ds.indices), g[tag + "_indices"])
        assert np.array_equal(np.asarray([ds.stats["position"]["min"], ds.stats["position"]["max"]]), g[tag + "_pos_stats"])
        assert np.array_equal(np.stack([ds.stats["velocity"]["min"], ds.stats["velocity"]["max"]]), g[tag + "_vel_stats"])
        assert np.array_equal(np.stack([ds.stats["action"]["min"], ds.stats["action"]["max"]]), g[tag + "_act_stats"])
        batch, tr, start, end = ds.collate(list(g[tag + "_idxs"]))
        for k in ("position", "velocity", "action"):
            assert batch[k].dtype == np.float32
            assert np.array_equal(batch[k], g[tag + "_" + k]), k          # bit-exact: same float32 operations
        assert np.array_equal(tr, g[tag + "_translation"])
        assert np.array_equal(np.stack([start, end], axis=1), g[tag + "_start_end"])
        im = batch["image"].astype(np.float64)
        fp = np.stack([im.sum(axis=(2, 3, 4)), im[:, :, 0, 0, 0], im[:, :, 1, 5, 7], im[:, :, 2, -1, -1],
                       (im * np.arange(im.shape[-1])).sum(axis=(2, 3, 4))], axis=-1)
        assert np.allclose(fp, g[tag + "_image_fp"], rtol=1e-12, atol=0)
        un = np.stack([data_ref.unnormalize_position(batch["position"][i], tr[i], ds.stats["position"]) for i in range(len(tr))])
        assert np.array_equal(un, g[tag + "_unnorm"])
        # round trip: un-normalising a window gives back the raw positions of its frames
        for i, s0 in enumerate(start):
            rawpos = raw["position"][s0:end[i]:step]
            assert np.allclose(un[i], rawpos, rtol=0, atol=2e-4 * float(np.abs(rawpos).max()))


def test_simple_unet_oracle_vs_reference_golden(golden_dir):
    """oracle/simple_unet_ref.py against the reference's own legacy UNet (models/simple_Unet.py:260-339; golden written by
    oracle/make_golden.py::golden_simple_unet, which also checks the PositionalEncoding buffer restatement bit for bit)."""
    import os
    import numpy as np
    from oracle import simple_unet_ref
    g = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(golden_dir, "simple_unet.npz")).items()}
    sd = fixtures.make_simple_unet_weights(seed=int(g["seed"]))
    with torch.no_grad():
        out = simple_unet_ref.unet_forward(sd, g["x"], g["t"], g["y"])
        out1 = simple_unet_ref.unet_forward(sd, g["x"], torch.tensor([17]), g["y"])
    assert float((out - g["out"]).abs().max()) <= 1e-6 * float(g["out"].abs().max())
    assert float((out1 - g["out_t1"]).abs().max()) <= 1e-6 * float(g["out_t1"].abs().max())


def test_resnet18gn_oracle_vs_reference_golden(golden_dir):
    """oracle/resnet_ref.py against the reference's own `VisionEncoder()` (models/Unet_FiLmLayer.py:316-386; golden written by
    oracle/make_golden.py::golden_resnet18gn): 512 features of 5 frames."""
    import os
    import numpy as np
    from oracle import resnet_ref
    g = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(golden_dir, "resnet18gn.npz")).items()}
    seed = int(g["seed"])
    img = torch.rand((5, 3, 96, 96), generator=torch.Generator().manual_seed(seed + 100))
    f = img.flatten(1)
    fp = torch.stack([f.mean(1), f.std(1), f.abs().max(1).values, f[:, 0], f[:, -1]], dim=1)
    assert torch.allclose(fp, g["img_fingerprint"]), "the regenerated frames differ from the ones the golden was made with"
    sd = fixtures.make_resnet_weights(seed=seed)
    with torch.no_grad():
        out = resnet_ref.encode(sd, img)
    assert float((out - g["out"]).abs().max()) <= 1e-5 * float(g["out"].abs().max())
