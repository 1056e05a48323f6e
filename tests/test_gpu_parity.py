"""GPU parity tests: the CUDA path (through the C ABI of libspdm.so) against the CPU oracle and the committed
golden vectors produced by the unmodified reference modules (tests/golden/*.npz, oracle/make_golden.py).

Tolerances (BASELINE.json north_star): fp32 path rel 1e-4 per U-Net forward and per denoising step;
bf16 path rel 1e-2 on the final trajectory.  rel(a, b) = max|a - b| / max|b|.
"""
import os

import numpy as np
import pytest
import torch

from oracle import fixtures, sampler_ref, unet_ref
from oracle.schedulers import RefDDIMScheduler, RefDDPMScheduler

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
BF16_FWD_TOL = 3e-2     # single forward, informational bound for the tensor-core path
BF16_FINAL_TOL = 1e-2   # final trajectory (the contract)


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


@pytest.fixture(scope="module")
def spdm():
    import state_policy_diffusionmodel_b200 as m
    assert torch.cuda.is_available()
    return m


def _golden(golden_dir, name):
    return {k: torch.from_numpy(v) if isinstance(v, np.ndarray) and v.dtype != object else v
            for k, v in np.load(os.path.join(golden_dir, name + ".npz")).items()}


UNET_CASES = [("unet_attn", True, 0, 31, 5), ("unet_noattn", False, 3, 31, 5), ("unet_attn_rows61", True, 4, 61, 5),
              ("unet_noattn_pos2", False, 5, 31, 2)]


@pytest.mark.parametrize("name,attention,seed,rows,dim", UNET_CASES)
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_unet_forward_vs_golden(spdm, golden_dir, name, attention, seed, rows, dim, precision):
    """UNet_Film / UNet_Film_noAttention forward against the reference's own output (golden) and the oracle."""
    g = _golden(golden_dir, name)
    sd = fixtures.make_unet_weights(attention=attention, seed=seed)
    x, y, t = g["x"], g["y"], g["t"]
    B = x.shape[0]
    plan = spdm.DenoisePlan(attention=attention, precision=precision, batch_max=B, rows=rows, dim=dim, graph_steps=0)
    plan.load_unet_state_dict(sd)
    assert [m for m in plan.missing_weights() if not m.startswith("vision_encoder")] == []
    tol = FP32_TOL if precision == "fp32" else BF16_FWD_TOL
    out = plan.unet_forward(x, t, y)
    assert out.shape == g["out"].shape
    assert rel(out, g["out"]) < tol
    out_nc = plan.unet_forward(x, t, None)
    assert rel(out_nc, g["out_nocond"]) < tol
    # layout-sensitive intermediate activations pinned by the golden file
    for tap in ("down1", "up1"):
        full = g["full_" + tap]
        _, got = plan.debug_forward(x, t, y, tap, tuple(full.shape))
        assert rel(got, full) < tol, tap
    with torch.no_grad():
        ref = unet_ref.unet_forward(sd, x, t, y, attention=attention)
    assert rel(out, ref) < tol
    assert plan.launch_count > 0
    plan.close()


def test_unet_batch_and_t_broadcast(spdm):
    """t of shape (1,) broadcasts over the batch (models/diffusion_ddpm.py:209); odd batch sizes work on both paths."""
    sd = fixtures.make_unet_weights(attention=True, seed=0)
    g = torch.Generator().manual_seed(9)
    x = torch.rand((5, 1, 31, 5), generator=g)
    y = torch.randn((5, 1, 10, 135), generator=g)
    t = torch.tensor([17])
    with torch.no_grad():
        ref = unet_ref.unet_forward(sd, x, t, y, attention=True)
    for precision, tol in (("fp32", FP32_TOL), ("bf16", BF16_FWD_TOL)):
        plan = spdm.DenoisePlan(attention=True, precision=precision, batch_max=40, graph_steps=0)
        plan.load_unet_state_dict(sd)
        out = plan.unet_forward(x, t, y)
        assert rel(out, ref) < tol
        # a smaller batch on the same plan must give the same rows
        out2 = plan.unet_forward(x[:2], t, y[:2])
        assert rel(out2, ref[:2]) < tol
        plan.close()


def test_encoder_and_cond_vs_golden(spdm, golden_dir):
    """Autoencoder.encoder (models/encoder/autoencoder.py:11-20) + prepare_obs_cond_vectors (diffusion_ddpm.py:317-330)."""
    g = _golden(golden_dir, "encoder")
    esd = fixtures.make_encoder_weights()
    gen = torch.Generator().manual_seed(5)
    img = torch.rand((4, 3, 96, 96), generator=gen)
    plan = spdm.DenoisePlan(attention=False, precision="fp32", batch_max=4, graph_steps=0)
    plan.load_encoder_state_dict(esd)
    out = plan.encode_images(img)
    assert rel(out, g["out"]) < FP32_TOL
    batch = fixtures.make_batch(3, seed=4321)
    plan.load_unet_state_dict(fixtures.make_unet_weights(attention=False, seed=3))
    cond = plan.encode_cond(batch["image"], batch["position"], batch["action"], batch["velocity"])
    with torch.no_grad():
        ref = unet_ref.obs_cond(esd, batch)
    assert rel(cond.reshape(3, 10, 135), ref) < FP32_TOL
    plan.close()


@pytest.mark.parametrize("kind,T,n", [("ddpm", 1000, 1000), ("ddpm", 20, 20), ("ddim", 50, 50), ("ddim", 100, 100), ("ddim", 1000, 50)])
def test_step_kernel_vs_oracle(spdm, kind, T, n):
    """One fused posterior update + inpaint per schedule index against the restated diffusers step."""
    Ref = RefDDPMScheduler if kind == "ddpm" else RefDDIMScheduler
    Mine = spdm.DDPMScheduler if kind == "ddpm" else spdm.DDIMScheduler
    kw = dict(num_train_timesteps=T, beta_schedule="linear", clip_sample=False, prediction_type="epsilon")
    ref, mine = Ref(**kw), Mine(**kw)
    ref.set_timesteps(n)
    mine.set_timesteps(n)
    assert torch.equal(ref.timesteps, mine.timesteps)
    plan = spdm.DenoisePlan(attention=False, precision="fp32", batch_max=1, rows=31, dim=5, inpaint_rows=1, graph_steps=0,
                            scheduler_only=True)
    plan.set_schedule(kind, mine.coef_table(), mine.timesteps)
    g = torch.Generator().manual_seed(3)
    B = 7
    x = torch.randn((B, 1, 31, 5), generator=g)
    eps = torch.randn((B, 1, 31, 5), generator=g)
    z = torch.randn((B, 1, 31, 5), generator=g)
    inp = torch.randn((B, 1, 1, 5), generator=g)
    for i in sorted(set([0, 1, n // 2, n - 2, n - 1])):
        t = int(ref.timesteps[i])
        if kind == "ddpm":
            want = ref.step(eps, t, x, noise=z).prev_sample
        else:
            want = ref.step(eps, t, x).prev_sample
        want = sampler_ref.add_constraints(want.clone(), inp, 1)
        got = plan.step(x, eps, i, noise=z, inpaint=inp.reshape(B, -1))
        assert rel(got, want) < 1e-5, (kind, i, t)
    plan.close()


def test_scheduler_objects_on_gpu(spdm):
    """DDPMScheduler/DDIMScheduler .step / .add_noise (the objects the reference assigns to `.noise_scheduler`)."""
    g = torch.Generator().manual_seed(11)
    x = torch.randn((3, 1, 31, 5), generator=g)
    eps = torch.randn((3, 1, 31, 5), generator=g)
    z = torch.randn((3, 1, 31, 5), generator=g)
    kw = dict(num_train_timesteps=100, beta_schedule="linear", clip_sample=False, prediction_type="epsilon")
    for Ref, Mine in ((RefDDPMScheduler, spdm.DDPMScheduler), (RefDDIMScheduler, spdm.DDIMScheduler)):
        ref, mine = Ref(**kw), Mine(**kw)
        ref.set_timesteps(100)
        mine.set_timesteps(100)
        for t in (99, 42, 0):
            want = ref.step(eps, t, x, noise=z).prev_sample if Ref is RefDDPMScheduler else ref.step(eps, t, x).prev_sample
            got = mine.step(eps.cuda(), t, x.cuda(), variance_noise=z.cuda()).prev_sample
            assert rel(got, want) < 1e-5
        tt = torch.tensor([5, 50, 99])
        assert rel(mine.add_noise(x.cuda(), z.cuda(), tt.cuda()), ref.add_noise(x, z, tt)) < 1e-6


SAMPLE_CASES = [("sample_ddim10_attn", "ddim", True, 0, 5), ("sample_ddpm20_noattn_pos2", "ddpm", False, 3, 2)]


@pytest.mark.parametrize("name,kind,attention,seed,dim", SAMPLE_CASES)
@pytest.mark.parametrize("graph_steps", [0, 1, 4])
def test_sample_fp32_per_step_vs_golden(spdm, golden_dir, name, kind, attention, seed, dim, graph_steps):
    """The K-step loop (spdm_sample) with injected x_T and noise against the history the unmodified reference wrapper
    produced (golden), per denoising step, fp32 path, rel 1e-4."""
    g = _golden(golden_dir, name)
    K = int(g["noise_steps"])
    sd = fixtures.make_unet_weights(attention=attention, seed=seed)
    Mine = spdm.DDPMScheduler if kind == "ddpm" else spdm.DDIMScheduler
    sch = Mine(num_train_timesteps=K, beta_schedule="linear", clip_sample=False, prediction_type="epsilon")
    sch.set_timesteps(K)
    assert torch.equal(sch.timesteps, g["timesteps"])
    plan = spdm.DenoisePlan(attention=attention, precision="fp32", batch_max=2, rows=31, dim=dim, inpaint_rows=1,
                            graph_steps=graph_steps)
    plan.load_unet_state_dict(sd)
    plan.set_schedule(kind, sch.coef_table(), sch.timesteps)
    plan.set_cond(g["obs_cond"][:1].reshape(1, -1))
    x0, hist = plan.sample(g["x_T"], noise=g["noise"], inpaint=g["inpaint"][:1].reshape(1, -1), history=True)
    want = g["history"]  # (K+1, 1, 1, 31, dim)
    assert hist.shape == want.shape
    for k in range(K + 1):
        assert rel(hist[k], want[k]) < FP32_TOL, "step %d" % k
    assert rel(x0, want[-1]) < FP32_TOL
    plan.close()


@pytest.mark.parametrize("kind,attention,K", [("ddim", True, 50), ("ddpm", False, 100)])
def test_sample_bf16_final_trajectory(spdm, kind, attention, K):
    """bf16 tensor-core path: final trajectory within rel 1e-2 of the fp32 oracle loop (same x_T, same injected noise)."""
    B = 4
    sd = fixtures.make_unet_weights(attention=attention, seed=0)
    esd = fixtures.make_encoder_weights()
    batch = fixtures.make_batch(B, seed=1234)
    x_T = fixtures.make_xT(B)
    noise = fixtures.make_noise(K, B)
    with torch.no_grad():
        cond = unet_ref.obs_cond(esd, batch).unsqueeze(1)
        inp = unet_ref.inpaint_vector(batch, 1).unsqueeze(1)
        ref_s = sampler_ref.make_scheduler(kind, K)
        want = sampler_ref.sample_ref(sd, ref_s, K, x_T, cond, inp, 1, attention=attention, noise=noise)
    Mine = spdm.DDPMScheduler if kind == "ddpm" else spdm.DDIMScheduler
    sch = Mine(num_train_timesteps=K, beta_schedule="linear", clip_sample=False, prediction_type="epsilon")
    sch.set_timesteps(K)
    results = {}
    for precision in ("fp32", "bf16"):
        plan = spdm.DenoisePlan(attention=attention, precision=precision, batch_max=B, inpaint_rows=1, graph_steps=5)
        plan.load_unet_state_dict(sd)
        plan.load_encoder_state_dict(esd)
        plan.set_schedule(kind, sch.coef_table(), sch.timesteps)
        plan.encode_cond(batch["image"], batch["position"], batch["action"], batch["velocity"])
        results[precision] = plan.sample(x_T, noise=noise, inpaint=inp.reshape(B, -1))
        plan.close()
    assert rel(results["fp32"], want) < 5e-4   # K compounded steps of the 1e-4 per-step path
    assert rel(results["bf16"], want) < BF16_FINAL_TOL


def test_sample_philox_noise_statistics(spdm):
    """Throughput mode draws z in-kernel (Philox4x32-10 + Box-Muller): check it is N(0,1), seed-reproducible and
    seed-sensitive, using a 1-step DDPM schedule whose update is x_prev = k_x*x + ... + sigma*z with eps-net output."""
    sd = fixtures.make_unet_weights(attention=False, seed=0)
    B, K = 64, 4
    sch = spdm.DDPMScheduler(num_train_timesteps=K, beta_schedule="linear", clip_sample=False, prediction_type="epsilon")
    sch.set_timesteps(K)
    plan = spdm.DenoisePlan(attention=False, precision="bf16", batch_max=B, cond_dim=0, inpaint_rows=0, graph_steps=2)
    plan.load_unet_state_dict({k: v for k, v in sd.items() if "cond_encoder" not in k})
    plan.set_schedule("ddpm", sch.coef_table(), sch.timesteps)
    x_T = fixtures.make_xT(B)
    a = plan.sample(x_T, seed=1)
    b = plan.sample(x_T, seed=1)
    c = plan.sample(x_T, seed=2)
    assert torch.equal(a, b)
    assert not torch.equal(a, c)
    # compare against the injected-noise run with z = 0: the difference is a linear image of the Philox draws
    zero = torch.zeros((K, B, 1, 31, 5))
    d = plan.sample(x_T, noise=zero)
    diff = (a - d).flatten()
    assert torch.isfinite(diff).all()
    assert abs(float(diff.mean())) < 0.05 * float(diff.std())
    plan.close()


def test_module_surface(spdm):
    """The nn.Module / wrapper surface the reference's scripts touch, end to end on the GPU."""
    torch.manual_seed(0)
    net = spdm.UNet_Film(in_channels=1, out_channels=1, noise_steps=1000, global_cond_dim=1350).cuda().eval()
    sd = fixtures.make_unet_weights(attention=True, seed=0)
    net.load_state_dict(sd, strict=True)
    net.configure(precision="fp32")
    g = torch.Generator().manual_seed(100)
    x = torch.rand((2, 1, 31, 5), generator=g)
    y = torch.randn((2, 1, 10, 135), generator=g)
    t = torch.tensor([10])
    with torch.no_grad():
        out = net(x.cuda(), t.cuda(), y.cuda())
        ref = unet_ref.unet_forward(sd, x, t, y, attention=True)
    assert rel(out, ref) < FP32_TOL
    with pytest.raises(RuntimeError):
        net(x, t, y)  # CPU tensors: no fallback

    m = spdm.Diffusion_DDIM(noise_steps=1000, obs_horizon=10, pred_horizon=30, observation_dim=135, prediction_dim=5,
                            model="UNet_Film", inpaint_horizon=1).cuda().eval()
    m.noise_estimator.load_state_dict(sd, strict=True)
    esd = fixtures.make_encoder_weights()
    m.vision_encoder.load_state_dict(esd, strict=True)
    m.configure(precision="fp32")
    m.use_ddim(10)  # generate.py:28-35
    batch = fixtures.make_batch(2, seed=4321)
    torch.manual_seed(5)
    hist = m.sample(batch={k: v.clone() for k, v in batch.items()}, option="sample_history")
    assert isinstance(hist, list) and len(hist) == 11 and tuple(hist[0].shape) == (1, 1, 31, 5)
    # same start sample through the oracle loop
    with torch.no_grad():
        cond = unet_ref.obs_cond(esd, batch)[:1].unsqueeze(1)
        inp = unet_ref.inpaint_vector(batch, 1)[:1].unsqueeze(1)
        want = sampler_ref.sample_ref(sd, sampler_ref.make_scheduler("ddim", 10), 10, hist[0].cpu(), cond, inp, 1, attention=True)
    assert rel(hist[-1], want) < 5e-4
    final = m.sample(batch={k: v.clone() for k, v in batch.items()}, x_T=hist[0])
    assert rel(final, want) < 5e-4
    x0, obs, inpaint = m.sample(batch={k: torch.cat([v, v, v, v], dim=1) for k, v in batch.items()}, mode="validation")
    assert tuple(x0.shape) == (1, 1, 31, 5) and tuple(inpaint.shape) == (1, 1, 1, 5) and obs["image"].shape[1] == 10
    outs = m.sample(batch={k: v.clone() for k, v in batch.items()}, batched=True)
    assert tuple(outs.shape) == (2, 1, 31, 5)


def test_training_forward_loss(spdm, golden_dir):
    """process_single_batch forward (q-sample + inpaint + U-Net + MSE, diffusion_ddpm.py:128-173) against the golden loss."""
    g = _golden(golden_dir, "validate_train")
    B = 3
    gen = torch.Generator().manual_seed(777)
    full = {"image": torch.rand((B, 40, 3, 96, 96), generator=gen), "position": 0.3 * torch.randn((B, 40, 2), generator=gen),
            "velocity": 2 * torch.rand((B, 40, 2), generator=gen) - 1, "action": 2 * torch.rand((B, 40, 3), generator=gen) - 1}
    m = spdm.Diffusion_DDPM(noise_steps=12, obs_horizon=10, pred_horizon=30, observation_dim=135, prediction_dim=5,
                            model="UNet_Film", inpaint_horizon=1).cuda().eval()
    m.noise_estimator.load_state_dict(fixtures.make_unet_weights(attention=True, seed=0), strict=True)
    m.vision_encoder.load_state_dict(fixtures.make_encoder_weights(), strict=True)
    m.configure(precision="fp32")
    with torch.no_grad():
        loss = m.process_single_batch(full, t=g["train_t"].cuda(), noise=g["train_noise"].cuda())
    assert abs(float(loss) - float(g["train_loss"])) < 1e-4 * abs(float(g["train_loss"]))


def test_concurrent_sub_batches_match(spdm):
    """SPDM_FLAG_SPLIT: running the U-Net of a denoising step as concurrent sub-batches must not change any sample
    (every op is per-sample; GroupNorm partial sums are per sample).  The split-K factor of the deep convs depends on
    the sub-batch size, i.e. on the fp32 summation order, so the comparison is to bf16 rounding, not bitwise."""
    B, K = 96, 6
    sd = fixtures.make_unet_weights(attention=True, seed=0)
    esd = fixtures.make_encoder_weights()
    batch = fixtures.make_batch(B, seed=21)
    x_T = fixtures.make_xT(B)
    sch = spdm.DDPMScheduler(num_train_timesteps=K, beta_schedule="linear", clip_sample=False, prediction_type="epsilon")
    sch.set_timesteps(K)
    outs = []
    for split in (1, 2, 3):
        plan = spdm.DenoisePlan(attention=True, precision="bf16", batch_max=B, inpaint_rows=1, graph_steps=3, split=split)
        plan.load_unet_state_dict(sd)
        plan.load_encoder_state_dict(esd)
        plan.set_schedule("ddpm", sch.coef_table(), sch.timesteps)
        plan.encode_cond(batch["image"], batch["position"], batch["action"], batch["velocity"])
        inp = torch.cat([batch["position"][:, -1:], batch["action"][:, -1:]], dim=-1).reshape(B, -1)
        outs.append(plan.sample(x_T, inpaint=inp, seed=5).cpu())
        plan.close()
    assert torch.isfinite(outs[0]).all()
    assert rel(outs[1], outs[0]) < 5e-3 and rel(outs[2], outs[0]) < 5e-3


def test_bf16_encoder_path(spdm, golden_dir):
    """bf16 plans run the encoder's Linear(9216 -> 128) on the tcgen05 GEMM: features within bf16 rounding of the golden."""
    g = _golden(golden_dir, "encoder")
    gen = torch.Generator().manual_seed(5)
    img = torch.rand((4, 3, 96, 96), generator=gen)
    plan = spdm.DenoisePlan(attention=False, precision="bf16", batch_max=4, graph_steps=0)
    plan.load_encoder_state_dict(fixtures.make_encoder_weights())
    out = plan.encode_images(img)
    assert rel(out, g["out"]) < 1e-2
    plan.close()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_unet_forward_4x_horizon(spdm, precision):
    """Prediction horizon 4x the repo default (rows = 121 -> 128x8 padded map, 1024-token attention at the top level):
    the geometry of BASELINE configs[4]; checked against the oracle (pinned to the reference at rows 31 and 61)."""
    sd = fixtures.make_unet_weights(attention=True, seed=2)
    g = torch.Generator().manual_seed(31)
    B = 2
    x = torch.rand((B, 1, 121, 5), generator=g)
    y = torch.randn((B, 1, 10, 135), generator=g)
    t = torch.tensor([250, 3])
    with torch.no_grad():
        ref = unet_ref.unet_forward(sd, x, t, y, attention=True)
    plan = spdm.DenoisePlan(attention=True, precision=precision, batch_max=B, rows=121, dim=5, graph_steps=0)
    plan.load_unet_state_dict(sd)
    out = plan.unet_forward(x, t, y)
    assert out.shape == ref.shape
    assert rel(out, ref) < (FP32_TOL if precision == "fp32" else BF16_FWD_TOL)
    plan.close()


def test_error_paths_are_loud(spdm):
    """No silent fallback: misuse of the C ABI reports an error message instead of computing something else."""
    from state_policy_diffusionmodel_b200._lib import SpdmError
    sd = fixtures.make_unet_weights(attention=False, seed=0)
    plan = spdm.DenoisePlan(attention=False, precision="bf16", batch_max=4, graph_steps=1)
    x = torch.rand((2, 1, 31, 5))
    with pytest.raises(SpdmError, match="weights missing"):
        plan.unet_forward(x, torch.tensor([1]), None)
    with pytest.raises(SpdmError, match="unexpected weight name"):
        plan.load_weight("inc.third.weight", torch.zeros(3))
    with pytest.raises(SpdmError, match="shape"):
        plan.load_weight("inc.second.weight", torch.zeros(64, 64, 1, 1))
    plan.load_unet_state_dict(sd)
    with pytest.raises(SpdmError, match="batch_max"):
        plan.unet_forward(torch.rand((5, 1, 31, 5)), torch.tensor([1]), None)
    with pytest.raises(SpdmError, match="conditioning"):
        plan.sample(x)
    plan.set_cond(torch.zeros((2, 1350)))
    with pytest.raises(SpdmError, match="no schedule"):
        plan.sample(x)
    sch = spdm.DDIMScheduler(num_train_timesteps=4, beta_schedule="linear", clip_sample=False, prediction_type="epsilon")
    sch.set_timesteps(4)
    plan.set_schedule("ddim", sch.coef_table(), sch.timesteps)
    assert torch.isfinite(plan.sample(x)).all()
    with pytest.raises(SpdmError, match="t_count"):
        plan.unet_forward(x, torch.tensor([1, 2, 3]), None)
    plan.close()
    with pytest.raises(NotImplementedError):
        spdm.DDPMScheduler(num_train_timesteps=10, thresholding=True)
    with pytest.raises(SpdmError, match="fp32 path only"):
        spdm.DenoisePlan(simple=True, precision="bf16", batch_max=1)


def test_sampling_pipeline_matches_sequential(spdm):
    """SamplingPipeline (independent batches in flight on separate plans/streams) returns what sample() returns."""
    m = spdm.Diffusion_DDIM(noise_steps=1000, obs_horizon=10, pred_horizon=30, observation_dim=135, prediction_dim=5,
                            model="UNet_FilmnoAttention", inpaint_horizon=1).cuda().eval()
    m.noise_estimator.load_state_dict(fixtures.make_unet_weights(attention=False, seed=3), strict=True)
    m.vision_encoder.load_state_dict(fixtures.make_encoder_weights(), strict=True)
    m.configure(precision="bf16", graph_steps=4, batch_max=8).use_ddim(8)
    batches = [fixtures.make_batch(8, seed=50 + i) for i in range(5)]
    xs = [fixtures.make_xT(8, seed=70 + i) for i in range(5)]
    want = [m.sample({k: v.clone() for k, v in b.items()}, batched=True, x_T=x, seed=1).cpu() for b, x in zip(batches, xs)]
    pipe = spdm.SamplingPipeline(m, depth=3, batch_max=8)
    tickets = [pipe.submit(b, x_T=x.cuda(), seed=1) for b, x in zip(batches, xs)]
    got = [pipe.result(t).cpu() for t in tickets]
    for g, w in zip(got, want):
        assert torch.equal(g, w)


@pytest.mark.parametrize("B", [24, 256, 512, 1536])
def test_bf16_forward_cluster_splitk_geometries(spdm, B):
    """The deep-level convs change kernel with the batch (cluster split-K + fused GroupNorm for few tiles: cluster sizes
    8 / 4 / 2 at these batches; plain tiles beyond): the bf16 forward must agree with the fp32 CUDA path (itself pinned to
    the golden vectors at rel 1e-4) at every one of them, including a batch that needs padding to the tile granularity.  The
    attention blocks change too: fused head + separate tail below ~300 (sa6) / ~1200 (sa5) trajectories, the whole C = 64 block in
    one launch above (B = 1536 covers both)."""
    sd = fixtures.make_unet_weights(attention=True, seed=0)
    g = torch.Generator().manual_seed(21 + B)
    x = torch.rand((B, 1, 31, 5), generator=g)
    y = torch.randn((B, 1, 10, 135), generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    outs = {}
    for precision in ("fp32", "bf16"):
        plan = spdm.DenoisePlan(attention=True, precision=precision, batch_max=B, graph_steps=0)
        plan.load_unet_state_dict(sd)
        outs[precision] = plan.unet_forward(x, t, y).float().cpu()
        plan.close()
    assert torch.isfinite(outs["bf16"]).all()
    assert rel(outs["bf16"], outs["fp32"]) < BF16_FWD_TOL
    # per-sample check: a wrong sample <-> statistics association would hide in a global max-norm
    per = (outs["bf16"] - outs["fp32"]).abs().flatten(1).max(dim=1).values / outs["fp32"].abs().flatten(1).max(dim=1).values
    assert float(per.max()) < 2 * BF16_FWD_TOL


def test_bf16_forward_folded_w2_convs(spdm, monkeypatch):
    """At batches whose tiles fill the machine the 3x3 convs of the H/4 x 2 level run folded (the two pixels of a row become
    channels: K = 3 x 2 Cin, N = 2 Cout, no zero-padding MMAs along W).  Same maths, other summation order: the folded bf16
    forward must agree with the unfolded bf16 forward (SPDM_NO_FOLD=1) and with the fp32 CUDA path."""
    B = 1024
    sd = fixtures.make_unet_weights(attention=True, seed=0)
    g = torch.Generator().manual_seed(5)
    x = torch.rand((B, 1, 31, 5), generator=g)
    y = torch.randn((B, 1, 10, 135), generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    outs = {}
    for tag, precision, nofold in (("fp32", "fp32", "0"), ("fold", "bf16", "0"), ("plain", "bf16", "1")):
        monkeypatch.setenv("SPDM_NO_FOLD", nofold)
        plan = spdm.DenoisePlan(attention=True, precision=precision, batch_max=B, graph_steps=0)
        plan.load_unet_state_dict(sd)
        outs[tag] = plan.unet_forward(x, t, y).float().cpu()
        if tag != "fp32":
            # the tap of a folded layer reads the same memory through the natural (B, C, H, W) view
            _, outs[tag + "_tap"] = plan.debug_forward(x, t, y, "up1", (B, 128, 8, 2))
        plan.close()
    assert torch.isfinite(outs["fold"]).all()
    assert rel(outs["fold"], outs["fp32"]) < BF16_FWD_TOL
    assert rel(outs["fold"], outs["plain"]) < BF16_FWD_TOL
    assert rel(outs["fold_tap"], outs["plain_tap"]) < BF16_FWD_TOL
    assert not torch.equal(outs["fold"], outs["plain"]), "the folded path was not taken"


def test_load_from_checkpoint_like_generate_py(spdm, tmp_path):
    """generate.py:23-36 `load_model`: Diffusion_DDIM.load_from_checkpoint(ckpt, hparams_file=...) on a Lightning-format
    checkpoint (state_dict with the `noise_estimator.` / `vision_encoder.` prefixes, hyper-parameters in hparams.yaml), then the
    scheduler swap; the loaded model must sample what a directly-constructed model with the same weights samples."""
    import yaml
    sd = fixtures.make_unet_weights(attention=True, seed=0)
    esd = fixtures.make_encoder_weights()
    hp = dict(noise_steps=1000, obs_horizon=10, pred_horizon=30, observation_dim=135, prediction_dim=5, learning_rate=1e-4,
              model="UNet_Film", vision_encoder=None, noise_scheduler_type="linear", inpaint_horizon=1, step_size=5)
    state = {"noise_estimator." + k: v for k, v in sd.items()}
    state.update({"vision_encoder." + k: v for k, v in esd.items()})
    ckpt = tmp_path / "epoch=4.ckpt"
    torch.save({"state_dict": state, "epoch": 4, "global_step": 100}, ckpt)
    hparams = tmp_path / "hparams.yaml"
    hparams.write_text(yaml.safe_dump(hp))
    # ---- generate.py:23-36 ----
    num_of_ddim_steps = 10
    model = spdm.Diffusion_DDIM.load_from_checkpoint(str(ckpt), hparams_file=str(hparams))
    noise_scheduler = spdm.DDIMScheduler(num_train_timesteps=num_of_ddim_steps, beta_schedule='linear', clip_sample=False,
                                         prediction_type='epsilon')
    model.noise_scheduler = noise_scheduler
    model.noise_steps = num_of_ddim_steps
    model.eval()
    # ----
    model = model.cuda()
    model.configure(precision="fp32")
    assert model.hparams["step_size"] == 5 and model.pred_horizon == 30
    batch = fixtures.make_batch(2, seed=4321)
    x_T = fixtures.make_xT(1)
    got = model.sample({k: v.clone() for k, v in batch.items()}, x_T=x_T.cuda())
    with torch.no_grad():
        cond = unet_ref.obs_cond(esd, batch)[:1].unsqueeze(1)
        inp = unet_ref.inpaint_vector(batch, 1)[:1].unsqueeze(1)
        want = sampler_ref.sample_ref(sd, sampler_ref.make_scheduler("ddim", 10), 10, x_T, cond, inp, 1, attention=True)
    assert rel(got, want) < 5e-4
    # uint8 HWC frames through the same public call (batched extension)
    u8 = (batch["image"] * 255).round().clamp(0, 255).to(torch.uint8).permute(0, 1, 3, 4, 2).contiguous()
    fl = (u8.float() / 255.0).permute(0, 1, 4, 2, 3).contiguous()
    xT2 = fixtures.make_xT(2).cuda()
    a = model.sample(dict(batch, image=fl), batched=True, x_T=xT2)
    b = model.sample(dict(batch, image=u8), batched=True, x_T=xT2)
    assert torch.equal(a, b)


def test_simple_unet_forward_and_sampling(spdm, golden_dir):
    """The legacy `UNet` (models/simple_Unet.py:260-339; the model='UNet' DEFAULT of the reference's Diffusion_DDPM constructor) on the
    fp32 path: forward against the reference's own output (golden) at rel 1e-4, per-sample and broadcast timesteps; then a DDPM
    sampling loop through `Diffusion_DDPM()` built with the reference's default model argument against the oracle loop."""
    from oracle import simple_unet_ref
    g = _golden(golden_dir, "simple_unet")
    sd = fixtures.make_simple_unet_weights(seed=int(g["seed"]))
    net = spdm.UNet(in_channels=1, out_channels=1, noise_steps=1000, global_cond_dim=1350).cuda().eval()
    assert set(net.state_dict()) == set(sd) and all(tuple(net.state_dict()[k].shape) == tuple(v.shape) for k, v in sd.items())
    net.load_state_dict(sd, strict=True)
    with torch.no_grad():
        out = net(g["x"].cuda(), g["t"].cuda(), g["y"].cuda())
        out1 = net(g["x"].cuda(), torch.tensor([17]).cuda(), g["y"].cuda())
    assert rel(out, g["out"]) < FP32_TOL and rel(out1, g["out_t1"]) < FP32_TOL
    with pytest.raises(ValueError):
        net(g["x"].cuda(), g["t"].cuda(), None)
    # the wrapper with the reference's default `model` argument
    K = 12
    m = spdm.Diffusion_DDPM(noise_steps=K, obs_horizon=10, pred_horizon=30, observation_dim=135, prediction_dim=5, inpaint_horizon=1).cuda().eval()
    assert isinstance(m.noise_estimator, spdm.UNet) and m.precision == "fp32"
    sd12 = fixtures.make_simple_unet_weights(seed=11, noise_steps=K)
    esd = fixtures.make_encoder_weights()
    m.noise_estimator.load_state_dict(sd12, strict=True)
    m.vision_encoder.load_state_dict(esd, strict=True)
    B = 3
    batch = fixtures.make_batch(B, seed=31)
    x_T = fixtures.make_xT(B)
    noise = fixtures.make_noise(K, B)
    got = m.sample({k: v.clone() for k, v in batch.items()}, batched=True, x_T=x_T.cuda(), noise=noise.cuda())
    with torch.no_grad():
        cond = unet_ref.obs_cond(esd, batch).unsqueeze(1)
        inp = unet_ref.inpaint_vector(batch, 1).unsqueeze(1)
        want = sampler_ref.sample_ref(sd12, sampler_ref.make_scheduler("ddpm", K), K, x_T, cond, inp, 1, noise=noise,
                                      forward=simple_unet_ref.unet_forward)
    assert rel(got, want) < 5e-4
    one = m.sample({k: v.clone() for k, v in batch.items()})           # reference default: batch element 0 only
    assert tuple(one.shape) == (1, 1, 31, 5)
    m.train()
    with pytest.raises(NotImplementedError):
        m.process_single_batch({k: torch.cat([v] * 4, dim=1).cuda() for k, v in batch.items()})


@pytest.mark.parametrize("attention,B", [(True, 3), (True, 40), (False, 64)])
def test_tf32_path_forward_and_final_trajectory(spdm, attention, B):
    """precision='tf32' (north_star: "the bf16/TF32 path"): fp32 activations, the 3x3 convs on tcgen05.mma.kind::tf32.  A single
    forward within 5e-3 of the fp32 oracle (TF32 keeps 10 mantissa bits), the final DDIM-20 trajectory within the contract's
    rel 1e-2 -- and the convs really left the CUDA cores (the result differs from the fp32 path's)."""
    sd = fixtures.make_unet_weights(attention=attention, seed=0)
    esd = fixtures.make_encoder_weights()
    g = torch.Generator().manual_seed(41 + B)
    x = torch.rand((B, 1, 31, 5), generator=g)
    y = torch.randn((B, 1, 10, 135), generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    outs = {}
    for precision in ("fp32", "tf32"):
        plan = spdm.DenoisePlan(attention=attention, precision=precision, batch_max=B, graph_steps=0)
        plan.load_unet_state_dict(sd)
        outs[precision] = plan.unet_forward(x, t, y).cpu()
        plan.close()
    assert torch.isfinite(outs["tf32"]).all()
    assert rel(outs["tf32"], outs["fp32"]) < 5e-3
    assert not torch.equal(outs["tf32"], outs["fp32"]), "the TF32 plan computed on the fp32 CUDA-core path"
    if B > 8:
        return
    K = 20
    batch = fixtures.make_batch(B, seed=7)
    x_T = fixtures.make_xT(B)
    with torch.no_grad():
        cond = unet_ref.obs_cond(esd, batch).unsqueeze(1)
        inp = unet_ref.inpaint_vector(batch, 1).unsqueeze(1)
        want = sampler_ref.sample_ref(sd, sampler_ref.make_scheduler("ddim", K), K, x_T, cond, inp, 1, attention=attention)
    sch = spdm.DDIMScheduler(num_train_timesteps=K, beta_schedule="linear", clip_sample=False, prediction_type="epsilon")
    sch.set_timesteps(K)
    plan = spdm.DenoisePlan(attention=attention, precision="tf32", batch_max=B, inpaint_rows=1, graph_steps=5)
    plan.load_unet_state_dict(sd)
    plan.load_encoder_state_dict(esd)
    plan.set_schedule("ddim", sch.coef_table(), sch.timesteps)
    plan.encode_cond(batch["image"], batch["position"], batch["action"], batch["velocity"])
    got = plan.sample(x_T, inpaint=inp.reshape(B, -1))
    assert rel(got, want) < BF16_FINAL_TOL
    plan.close()


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 3e-2)])
def test_resnet18gn_encoder(spdm, golden_dir, precision, tol):
    """`vision_encoder='resnet18'`: the ResNet18-GroupNorm `VisionEncoder()` of models/Unet_FiLmLayer.py:316-386 --
    features against the reference module's own output (golden), conditioning vector (cond_dim = 519) against the oracle, and a
    sampling call through `Diffusion_DDIM(vision_encoder='resnet18', observation_dim=519)`."""
    from oracle import resnet_ref
    g = _golden(golden_dir, "resnet18gn")
    seed = int(g["seed"])
    img = torch.rand((5, 3, 96, 96), generator=torch.Generator().manual_seed(seed + 100))
    rsd = fixtures.make_resnet_weights(seed=seed)
    plan = spdm.DenoisePlan(attention=False, precision=precision, batch_max=4, cond_dim=519, encoder="resnet18", graph_steps=0)
    plan.load_encoder_state_dict(rsd)
    out = plan.encode_images(img)
    assert tuple(out.shape) == (5, 512)
    assert rel(out, g["out"]) < tol
    # a frame count that is not a multiple of the 128-row tile and spans two chunks gives the same rows
    out2 = plan.encode_images(img.repeat(58, 1, 1, 1))     # 290 frames = one full chunk of 256 + a ragged one
    # (the GroupNorm statistics are accumulated with shared-memory atomics: the summation order, and with it a bf16 rounding here and
    # there, depends on the launch -- hence not bit-identical on the bf16 path)
    assert tuple(out2.shape) == (290, 512) and rel(out2, out.repeat(58, 1)) < (1e-5 if precision == "fp32" else 1e-2)
    plan.close()
    # wrapper: cond_dim 519, FiLM U-Net conditioned on ResNet features
    torch.manual_seed(0)
    with pytest.raises(ValueError):
        spdm.Diffusion_DDIM(noise_steps=1000, obs_horizon=10, pred_horizon=30, observation_dim=135, prediction_dim=5, model="UNet_Film",
                            vision_encoder="resnet18", inpaint_horizon=1)
    m = spdm.Diffusion_DDIM(noise_steps=1000, obs_horizon=10, pred_horizon=30, observation_dim=519, prediction_dim=5,
                            model="UNet_FilmnoAttention", vision_encoder="resnet18", inpaint_horizon=1).cuda().eval()
    assert set(m.vision_encoder.state_dict()) == set(rsd)
    m.vision_encoder.load_state_dict(rsd, strict=True)
    sd = fixtures.make_unet_weights(attention=False, cond_dim=5190, seed=3)
    m.noise_estimator.load_state_dict(sd, strict=True)
    m.configure(precision=precision, graph_steps=2).use_ddim(4)
    batch = fixtures.make_batch(2, seed=77)
    cond = m.prepare_obs_cond_vectors({k: v.cuda() for k, v in batch.items()})
    with torch.no_grad():
        feat = resnet_ref.encode(rsd, batch["image"].flatten(end_dim=1)).reshape(2, 10, 512)
        want_cond = torch.cat([batch["position"], batch["action"], batch["velocity"], feat], dim=-1)
    assert tuple(cond.shape) == (2, 10, 519) and rel(cond, want_cond) < tol
    x_T = fixtures.make_xT(2)
    got = m.sample({k: v.clone() for k, v in batch.items()}, batched=True, x_T=x_T.cuda())
    with torch.no_grad():
        inp = unet_ref.inpaint_vector(batch, 1).unsqueeze(1)
        want = sampler_ref.sample_ref(sd, sampler_ref.make_scheduler("ddim", 4), 4, x_T, want_cond.unsqueeze(1), inp, 1, attention=False)
    assert rel(got, want) < (5e-4 if precision == "fp32" else BF16_FINAL_TOL)
    m.train()
    with pytest.raises(NotImplementedError):
        m.process_single_batch({k: torch.cat([v] * 4, dim=1).cuda() for k, v in batch.items()})
