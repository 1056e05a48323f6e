"""CPU: scheduler restatement vs SURVEY.md A.4 known-answer values + self-consistency identities.
(diffusers 0.17.1 itself is not installable here: parity unpinned against the package.)"""
import pytest
import torch

from oracle.schedulers import (RefDDIMScheduler, RefDDPMScheduler, ref_cosine_beta_schedule,
                               ref_linear_beta_schedule, ref_linear_beta_schedule_v2)


def close(a, b, rel=2e-6):
    assert abs(float(a) - b) <= rel * abs(b) + 1e-12, (float(a), b)


def test_tables_known_answers():
    s = RefDDPMScheduler(1000, clip_sample=False)
    close(s.betas[0], 9.999999747e-05); close(s.betas[1], 1.1991991778e-04); close(s.betas[999], 1.9999999553e-02)
    close(s.alphas_cumprod[0], 0.9998999834); close(s.alphas_cumprod[500], 0.0777966529, 1e-5)
    close(s.alphas_cumprod[999], 4.0358303522e-05, 1e-4)
    s = RefDDPMScheduler(100, clip_sample=False)
    close(s.betas[1], 3.0101009179e-04); close(s.alphas_cumprod[50], 0.7692912817, 1e-5); close(s.alphas_cumprod[99], 0.3635632396, 1e-5)
    s = RefDDPMScheduler(50, clip_sample=False)
    close(s.betas[1], 5.0612242194e-04); close(s.alphas_cumprod[25], 0.8736623526, 1e-5); close(s.alphas_cumprod[49], 0.6029515862, 1e-5)


@pytest.mark.parametrize("t,vals", [
    (999, (0.9999797940, 0.0063528186, 1.2835137e-04, 0.9899486899, 0.1414212286)),
    (500, (0.9603142142, 0.2789205015, 3.0580366e-03, 0.9941043854, 0.1002560183)),
    (1, (0.0148304123, 0.9998900294, 0.5452302098, 0.4547152817, 0.0073847678)),
    (0, (0.0100008296, 0.9999499917, 1.0, 0.0, 0.0))])
def test_ddpm_coefficients(t, vals):
    s = RefDDPMScheduler(1000, clip_sample=False)
    s.set_timesteps(1000)
    c = s.coefficients(t)
    for k, v in zip(("sqrt_beta_prod", "sqrt_alpha_prod", "c0", "cx", "sigma"), vals):
        close(c[k], v, 2e-5)


def test_ddim_coefficients():
    s = RefDDIMScheduler(50, clip_sample=False); s.set_timesteps(50)
    c = s.coefficients(49)
    close(c["sqrt_alpha_prod"], 0.7764995694, 1e-5); close(c["sqrt_beta_prod"], 0.6301177740, 1e-5)
    close(c["sqrt_alpha_prev"], 0.7843829989, 1e-5); close(c["dir_coef"], 0.6202767491, 1e-5)
    c = s.coefficients(0)
    close(c["sqrt_alpha_prod"], 0.9999499917); close(c["sqrt_alpha_prev"], 1.0); assert float(c["dir_coef"]) == 0.0
    s = RefDDIMScheduler(100, clip_sample=False); s.set_timesteps(100)
    c = s.coefficients(99)
    close(c["sqrt_alpha_prod"], 0.6029620767, 1e-5); close(c["sqrt_alpha_prev"], 0.6090836525, 1e-5)
    assert [int(t) for t in s.timesteps[:3]] == [99, 98, 97]


def test_identities():
    torch.manual_seed(0)
    s = RefDDPMScheduler(1000, clip_sample=False); s.set_timesteps(1000)
    x = torch.randn(2, 1, 31, 5); eps = torch.randn_like(x)
    for t in (999, 321, 1):
        mu = s.step(eps, t, x, noise=torch.zeros_like(x)).prev_sample
        closed = (x - s.betas[t] / (1 - s.alphas_cumprod[t]) ** 0.5 * eps) / s.alphas[t] ** 0.5
        torch.testing.assert_close(mu, closed, rtol=1e-4, atol=1e-5)
    x0 = torch.randn(2, 1, 31, 5)
    xn = s.add_noise(x0, eps, torch.tensor([0, 0]))
    torch.testing.assert_close(s.step(eps, 0, xn).prev_sample, x0, rtol=1e-4, atol=1e-5)
    d = RefDDIMScheduler(1000, clip_sample=False); d.set_timesteps(50)
    assert int(d.timesteps[0]) == 980 and int(d.timesteps[-1]) == 0
    a = d.step(eps, 980, x).prev_sample; b = d.step(eps, 980, x).prev_sample
    assert torch.equal(a, b)
    with pytest.raises(ValueError):
        RefDDPMScheduler(10).set_timesteps(11)


def test_utils_schedulers_known_answers():
    b = ref_linear_beta_schedule(100)
    close(b[0], 1.0000000475e-03); close(b[99], 0.2000000030); close(torch.cumprod(1 - b, 0)[99], 2.0390087e-05, 1e-4)
    b = ref_linear_beta_schedule_v2(1000)
    close(b[0], 4.9999999e-05); close(b[999], 9.9999998e-03); close(torch.cumprod(1 - b, 0)[999], 6.4618289e-03, 1e-4)
    b = ref_cosine_beta_schedule(1000)
    close(b[0], 4.1284224e-05, 1e-5); close(b[500], 3.1556915e-03, 1e-5); close(b[999], 0.9990000129)


def test_beta_schedules_vs_reference_golden():
    """utils/schedulers.py:6-40: the oracle restatement AND the product's drop-in functions are bit-identical to what the
    reference's own file returns (tests/golden/beta_schedules.npz, written by oracle/make_golden.py::golden_beta_schedules from
    the unmodified /root/reference/utils/schedulers.py)."""
    import os
    import numpy as np
    import state_policy_diffusionmodel_b200 as spdm
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "beta_schedules.npz"))

    class Dev:
        device = torch.device("cpu")
    for steps in (10, 50, 100, 1000):
        for tag, ref, mine in (("linear", ref_linear_beta_schedule, spdm.linear_beta_schedule),
                               ("linear_v2", ref_linear_beta_schedule_v2, spdm.linear_beta_schedule_v2),
                               ("cosine", ref_cosine_beta_schedule, spdm.cosine_beta_schedule)):
            want = torch.from_numpy(g["%s_%d" % (tag, steps)])
            assert torch.equal(ref(steps), want), (tag, steps)
            assert torch.equal(mine(Dev(), steps), want), (tag, steps)
    want = torch.from_numpy(g["cosine_100_s02_f64"])
    assert want.dtype == torch.float64
    assert torch.equal(ref_cosine_beta_schedule(100, s=0.02, dtype=torch.float64), want)
    assert torch.equal(spdm.cosine_beta_schedule(None, 100, s=0.02, dtype=torch.float64), want)
