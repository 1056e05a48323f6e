"""GPU: time the training step (spdm_train_fwd_bwd + spdm_adam_step) on synthetic data.
Usage: python tools/train_bench.py [B] [steps] [precision] [attn|noattn]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fixtures  # noqa: E402  (weights/inputs only)
import state_policy_diffusionmodel_b200 as spdm  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    precision = sys.argv[3] if len(sys.argv) > 3 else "bf16"
    attention = (sys.argv[4] if len(sys.argv) > 4 else "attn") == "attn"
    dev = torch.device("cuda", 0)
    sd = fixtures.make_unet_weights(attention=attention, seed=0)
    esd = fixtures.make_encoder_weights()
    plan = spdm.DenoisePlan(attention=attention, precision=precision, batch_max=B, inpaint_rows=1)
    named = dict(sd)
    named.update({"vision_encoder." + k: v for k, v in esd.items()})
    plan.enable_training(named)
    g = torch.Generator(device=dev).manual_seed(1)
    img = torch.rand((B, 10, 3, 96, 96), device=dev, generator=g)
    pos = 0.3 * torch.randn((B, 10, 2), device=dev, generator=g)
    vel = 2 * torch.rand((B, 10, 2), device=dev, generator=g) - 1
    act = 2 * torch.rand((B, 10, 3), device=dev, generator=g) - 1
    x0 = torch.randn((B, 1, 31, 5), device=dev, generator=g) * 0.3
    inp = x0[:, 0, :1, :].reshape(B, -1).contiguous()
    sch = spdm.DDPMScheduler(num_train_timesteps=1000, beta_schedule="linear", clip_sample=False, prediction_type="epsilon")
    ac = sch.alphas_cumprod.to(dev)
    sa, sb = ac ** 0.5, (1 - ac) ** 0.5

    def step():
        t = torch.randint(0, 1000, (B,), device=dev)
        noise = torch.randn((B, 1, 31, 5), device=dev)
        loss = plan.train_fwd_bwd(img, pos, act, vel, x0, noise, t, sa, sb, inpaint=inp)
        plan.adam_step(lr=1e-4, max_norm=0.5)
        return loss

    losses = []
    for _ in range(3):
        losses.append(step())
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    l0 = plan.launch_count
    e0.record()
    for _ in range(steps):
        losses.append(step())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print("B=%d %s %s: %.2f ms/step, %.0f samples/s, launches/step %d, workspace %.0f MB, loss %s" % (
        B, precision, "attn" if attention else "noattn", ms, B / ms * 1e3, (plan.launch_count - l0) // steps, plan.workspace_bytes / 1e6,
        [round(float(x), 4) for x in losses[:3] + losses[-1:]]))


if __name__ == "__main__":
    main()
