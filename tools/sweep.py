#!/usr/bin/env python
"""Sweep of BASELINE configs[4]: sampler steps x batch x prediction horizon on one GPU (run on a B200:
`python tools/sweep.py > profiles/rNN_sweep.md`) or, under torchrun, on N GPUs (weak scaling: `batch` trajectories per GPU,
batch-sharded, max over ranks; the table then reports the aggregate).  For every point: trajectories/s with device-resident inputs
(conditioning encode + graphed K-step loop, CUDA events), ms per denoising step, and the kernel-class table of one
eager, event-timed denoising step (tensor TFLOP/s of the conv class against the measured bf16 peak, GB/s of the
GroupNorm-apply class against the measured HBM peak)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import state_policy_diffusionmodel_b200 as spdm  # noqa: E402
from bench import merge_conv, peaks, synth_batch  # noqa: E402


WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))


def run_point(model, K, B, rows, reps):
    import torch.distributed as dist
    dev = model.device
    model.pred_horizon = rows - 1
    model.use_ddim(K)
    host = synth_batch(B, 10, 1234)
    devb = {k: v.to(dev) for k, v in host.items()}
    x_T = torch.rand((B, 1, rows, 5), device=dev)
    plan = model._plan(B)
    model._bind_schedule(plan)
    inpaint = model.prepare_inpaint_vectors(devb).reshape(B, -1).contiguous()

    def step():
        plan.encode_cond(devb["image"], devb["position"], devb["action"], devb["velocity"])
        return plan.sample(x_T, inpaint=inpaint, seed=1)

    heavy = B * rows * K > 4096 * 61 * 50
    for _ in range(1 if heavy else 3):
        step()
    torch.cuda.synchronize()
    if WORLD > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    if WORLD > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    prof = plan.profile_step(B, reps=3)
    return ms, prof


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, nargs="*", default=[10, 25, 50, 100])
    ap.add_argument("--batches", type=int, nargs="*", default=[64, 256, 1024, 4096, 16384], help="trajectories per GPU")
    ap.add_argument("--rows", type=int, nargs="*", default=[31, 61, 121])
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    pk = peaks()
    if WORLD > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    torch.manual_seed(0)
    out = (lambda *a, **k: print(*a, **k)) if RANK == 0 else (lambda *a, **k: None)
    out("%d GPU(s), batch = trajectories per GPU, trajectories/s = aggregate over the GPUs\n" % WORLD)
    out("| rows | batch | DDIM steps | trajectories/s | ms / denoise step | conv3x3 TFLOP/s (frac of %.0f) | GN-apply GB/s (frac of %.0f) | attention core ms | launches/step |"
          % (pk["tf_burst"], pk["hbm"]))
    out("|---|---|---|---|---|---|---|---|---|")
    for rows in args.rows:
        model = spdm.Diffusion_DDIM(noise_steps=1000, obs_horizon=10, pred_horizon=rows - 1, observation_dim=135, prediction_dim=5,
                                    model="UNet_Film", inpaint_horizon=1).cuda().eval()   # (cuda:LOCAL_RANK under torchrun)
        model.configure(precision="bf16", graph_steps=10, batch_max=max(args.batches))
        for B in args.batches:
            for K in args.steps:
                if K != 50 and B not in (256, 4096):
                    continue
                ms, prof = run_point(model, K, B, rows, args.reps)
                conv, app = merge_conv(prof), prof["gn_apply"]
                tf = conv["flops"] / (conv["ms"] * 1e-3) / 1e12 if conv["ms"] > 0 else 0.0
                gbs = app["bytes"] / (app["ms"] * 1e-3) / 1e9 if app["ms"] > 0 else 0.0
                launches = sum(v["launches"] for v in prof.values())
                out("| %d | %d | %d | %.0f | %.3f | %.0f (%.2f) | %.0f (%.2f) | %.3f | %.0f |" % (
                    rows, B, K, B * WORLD / (ms * 1e-3), ms / K, tf, tf / pk["tf_burst"], gbs, gbs / pk["hbm"], prof["sdpa"]["ms"], launches),
                    flush=True)
        del model
        torch.cuda.empty_cache()
    out()
    out(json.dumps({"peaks": pk, "n_gpus": WORLD}))
    if WORLD > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
