"""GPU diagnostic: per-tensor gradient error of spdm_train_fwd_bwd against the CPU oracle (oracle/train_ref.py).
Usage: python tools/train_diag.py [fp32|bf16] [B] [attn|noattn]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fixtures, train_ref  # noqa: E402
from oracle.schedulers import RefDDPMScheduler  # noqa: E402
import state_policy_diffusionmodel_b200 as spdm  # noqa: E402


def make_case(B, seed=778):
    g = torch.Generator().manual_seed(seed)
    full = {"image": torch.rand((B, 40, 3, 96, 96), generator=g), "position": 0.3 * torch.randn((B, 40, 2), generator=g),
            "velocity": 2 * torch.rand((B, 40, 2), generator=g) - 1, "action": 2 * torch.rand((B, 40, 3), generator=g) - 1}
    t = torch.randint(0, 1000, (B,), generator=g)
    noise = torch.randn((B, 1, 31, 5), generator=g)
    return full, t, noise


def run_gpu(precision, attention, sd, esd, full, t, noise, B):
    plan = spdm.DenoisePlan(attention=attention, precision=precision, batch_max=B, inpaint_rows=1)
    named = dict(sd)
    named.update({"vision_encoder." + k: v for k, v in esd.items()})
    plan.enable_training(named)
    sched = RefDDPMScheduler(num_train_timesteps=1000, beta_schedule="linear", clip_sample=False)
    ac = sched.alphas_cumprod
    obs = {k: v[:, :10] for k, v in full.items()}
    pred = {k: v[:, 10:] for k, v in full.items()}
    x0 = torch.cat([pred["position"], pred["action"]], dim=-1).unsqueeze(1)
    inp = torch.cat([obs["position"][:, -1:], obs["action"][:, -1:]], dim=-1)
    vec = torch.cat([inp.unsqueeze(1), x0], dim=2)
    loss = plan.train_fwd_bwd(obs["image"], obs["position"], obs["action"], obs["velocity"], vec, noise, t, ac ** 0.5, (1 - ac) ** 0.5,
                              inpaint=inp.reshape(B, -1))
    torch.cuda.synchronize()
    grads = {k: plan.grad_view(k).detach().cpu().clone() for k in named}
    return plan, float(loss.item()), grads


def main():
    precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    attention = (sys.argv[3] if len(sys.argv) > 3 else "attn") == "attn"
    sd = fixtures.make_unet_weights(attention=attention, seed=0)
    esd = fixtures.make_encoder_weights()
    full, t, noise = make_case(B)
    sched = RefDDPMScheduler(num_train_timesteps=1000, beta_schedule="linear", clip_sample=False)
    t0 = time.time()
    torch.set_num_threads(os.cpu_count())
    want_loss, want = train_ref.loss_and_grads(sd, esd, sched, full, 10, 1, t, noise, attention=attention)
    print("oracle: loss %.6f (%.1fs)" % (float(want_loss), time.time() - t0))
    plan, loss, got = run_gpu(precision, attention, sd, esd, full, t, noise, B)
    print("gpu %s: loss %.6f  launches %d  workspace %.1f MB" % (precision, loss, plan.launch_count, plan.workspace_bytes / 1e6))
    worst = []
    num = den = 0.0
    for k in want:
        a, b = got[k].double(), want[k].double()
        rel = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
        cos = float((a * b).sum() / (a.norm() * b.norm()).clamp_min(1e-30))
        num += float(((a - b) ** 2).sum())
        den += float((b ** 2).sum())
        worst.append((rel, cos, k, float(b.abs().max())))
    worst.sort(reverse=True)
    for rel, cos, k, mx in worst[:40]:
        print("  %-48s rel %.3e cos %.6f  max|g| %.3e" % (k, rel, cos, mx))
    print("global rel l2 %.3e over %d tensors; median rel %.3e" % ((num / den) ** 0.5, len(worst), worst[len(worst) // 2][0]))


if __name__ == "__main__":
    main()
