"""GPU parity of the dataset -> batch path: spdm_gather_windows / spdm_unnormalize_position through
state_policy_diffusionmodel_b200.DeviceWindowDataset against the oracle restatement of the reference's CarRacingDataset
(oracle/data_ref.py, itself pinned bit-exact to the reference classes by tests/golden/dataset.npz).  Integer / index work and
the float32 normalisation are bit-exact; images are bit-exact (one IEEE division per value)."""
import numpy as np
import pytest
import torch

from oracle import data_ref

pytestmark = pytest.mark.gpu


def _datasets(spdm, obs_h, pred_h, step, u8=True, seed=0, **kw):
    raw = data_ref.make_synthetic_dataset(seed, **kw)
    img_f = data_ref.image_chw_float(raw["img_u8"])
    ref = data_ref.RefWindowDataset({"image": img_f, "position": raw["position"], "velocity": raw["velocity"], "action": raw["action"]},
                                    raw["episode_ends"], pred_h, obs_h, None, step)
    dev = spdm.DeviceWindowDataset(raw["img_u8"] if u8 else img_f, raw["position"], raw["velocity"], raw["action"], raw["episode_ends"],
                                   pred_h, obs_h, None, step)
    return raw, ref, dev


@pytest.mark.parametrize("obs_h,pred_h,step,u8", [(3, 4, 2, True), (2, 3, 1, True), (3, 4, 2, False), (10, 30, 1, True)])
def test_gather_windows_bit_exact(obs_h, pred_h, step, u8):
    import state_policy_diffusionmodel_b200 as spdm
    kw = dict(n_frames=160, episode_ends=(70, 160)) if obs_h == 10 else {}
    raw, ref, dev = _datasets(spdm, obs_h, pred_h, step, u8, **kw)
    assert dev.indices == ref.indices and len(dev) == len(ref) > 0
    assert float(dev.stats["position"]["min"]) == float(ref.stats["position"]["min"])
    assert np.array_equal(dev.stats["velocity"]["max"], ref.stats["velocity"]["max"])
    idxs = [len(ref) - 1, 0, len(ref) // 2, 1, 0]                      # unsorted, with a repeat
    want, tr, start, end = ref.collate(idxs)
    got, gtr, gstart, gend = dev.batch(idxs)
    for k in ("position", "velocity", "action", "image"):
        assert got[k].dtype == torch.float32 and tuple(got[k].shape) == want[k].shape, k
        assert np.array_equal(got[k].cpu().numpy(), want[k].astype(np.float32)), k
    assert np.array_equal(gtr.cpu().numpy(), tr)
    assert np.array_equal(gstart.cpu().numpy(), start) and np.array_equal(gend.cpu().numpy(), end)
    # images only for the observation window (what the model reads): a prefix of the full item
    got_obs, _, _, _ = dev.batch(idxs, image_frames=obs_h)
    assert np.array_equal(got_obs["image"].cpu().numpy(), want["image"][:, :obs_h].astype(np.float32))
    no_img, _, _, _ = dev.batch(idxs, image_frames=0)
    assert "image" not in no_img and np.array_equal(no_img["position"].cpu().numpy(), want["position"])


def test_unnormalize_position_and_round_trip():
    import state_policy_diffusionmodel_b200 as spdm
    raw, ref, dev = _datasets(spdm, 3, 4, 2)
    idxs = list(range(0, len(ref), 3))
    want, tr, start, end = ref.collate(idxs)
    got, gtr, _, _ = dev.batch(idxs, image_frames=0)
    un = dev.unnormalize_position(got["position"], gtr).cpu().numpy()
    exp = np.stack([data_ref.unnormalize_position(want["position"][i], tr[i], ref.stats["position"]) for i in range(len(idxs))])
    assert np.array_equal(un, exp.astype(np.float32))
    for i, s0 in enumerate(start):   # size-independent property: windows un-normalise back to the raw track
        rawpos = raw["position"][s0:end[i]:2]
        assert np.allclose(un[i], rawpos, rtol=0, atol=2e-4 * float(np.abs(rawpos).max()))
    # a (K+1, B, rows, 2) history stack with one translation per sample, as sample_history produces
    hist = torch.stack([got["position"] * (1.0 + 0.1 * k) for k in range(3)])
    trs = gtr.unsqueeze(0).expand(3, -1, -1).reshape(-1, 2)
    un_h = dev.unnormalize_position(hist, trs).cpu().numpy()
    for k in range(3):
        e = np.stack([data_ref.unnormalize_position(hist[k, i].cpu().numpy(), tr[i], ref.stats["position"]) for i in range(len(idxs))])
        assert np.array_equal(un_h[k], e.astype(np.float32))


def test_gather_feeds_the_model_and_errors_are_loud():
    import state_policy_diffusionmodel_b200 as spdm
    raw, ref, dev = _datasets(spdm, 10, 30, 1, n_frames=160, episode_ends=(70, 160))
    batch, tr, _, _ = dev.batch([0, 5, 9, 40], image_frames=10)
    model = spdm.Diffusion_DDIM(noise_steps=1000, obs_horizon=10, pred_horizon=30, observation_dim=135, prediction_dim=5,
                                model="UNet_Film", inpaint_horizon=1).cuda().eval()
    model.use_ddim(5)
    out = model.sample({k: v for k, v in batch.items()}, batched=True, seed=3)
    assert tuple(out.shape) == (4, 1, 31, 5) and torch.isfinite(out).all()
    world = dev.unnormalize_position(out[:, 0, :, :2].contiguous(), tr)
    assert tuple(world.shape) == (4, 31, 2) and torch.isfinite(world).all()
    with pytest.raises(IndexError):
        dev.batch([len(dev)])
    with pytest.raises(ValueError):
        dev.batch([])
    with pytest.raises(ValueError):
        spdm.DeviceWindowDataset(raw["img_u8"][:, :, :94], raw["position"], raw["velocity"], raw["action"], raw["episode_ends"], 30, 10)
    with pytest.raises(RuntimeError):
        spdm.DeviceWindowDataset(raw["img_u8"], raw["position"], raw["velocity"], raw["action"], raw["episode_ends"], 30, 10, device="cpu")
    lib = spdm._lib.load()
    rc = lib.spdm_gather_windows(None, 0, 96, 96, None, None, None, None, 1, 1, 0, 1, None, None, None, None, None, None, None)
    assert rc < 0 and b"null" in lib.spdm_last_error()
