#!/usr/bin/env python
"""In-graph marginal cost of each kernel group of one denoising step (run on a B200: `python tools/ablate.py --batch 256`).

ncu times every launch alone (cold caches, no programmatic-dependent-launch overlap), CUDA events around eager launches
add their own gaps; neither says what a kernel costs INSIDE the captured graph.  Here the graphed DDIM loop is timed with
one group of launches left out at a time (`SPDM_SKIP_IDX`, see csrc/plan.cu::timed) -- the outputs are garbage, the time
difference to the full loop is the group's marginal cost on the critical path.  One subprocess per configuration (the
switch is read when the plan is created)."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# launch indices of one UNet_Film forward at batch 256, bf16 path, in issue order: 64 launches (conv_in carries its GroupNorm + GELU) -- the 18 deep-level convs are
# cluster split-K launches with the GroupNorm apply fused behind them, every attention block is two launches (fused head + tail);
# run with SPDM_FUSE_MODE=0 (the default auto mode also folds 11 of the 13 apply launches into their swapped convs at this batch)
APPLY = [2, 5, 7, 9, 11, 43, 45, 47, 49, 54, 56, 58, 60]
GROUPS = {
    "none": [],
    "gn_apply, separate kernels, 32x8 + 16x4 levels (13)": APPLY,
    "attention heads sa1-sa5: LN + in_proj + core (5)": [12, 19, 26, 39, 50],
    "attention tails (6)": [13, 20, 27, 40, 51, 62],
    "attention heads: LN + in_proj + core (6)": [12, 19, 26, 39, 50, 61],
    "sa6: head, tail (2)": [61, 62],
    "attention whole (12)": [12, 13, 19, 20, 26, 27, 39, 40, 50, 51, 61, 62],
    "pool+upsample (6)": [3, 14, 21, 34, 41, 52],
    "conv 32x8 (5)": [1, 53, 55, 57, 59],
    "conv 16x4 (8)": [4, 6, 8, 10, 42, 44, 46, 48],
    "cluster conv+GN 8x2 (8)": [15, 16, 17, 18, 35, 36, 37, 38],
    "cluster conv+GN 4x1 (10)": [22, 23, 24, 25, 28, 29, 30, 31, 32, 33],
    "everything but conv_in/outc (62)": list(range(1, 63)),
}


def child(args):
    import torch
    sys.path.insert(0, ROOT)
    import state_policy_diffusionmodel_b200 as spdm
    from bench import synth_batch
    torch.manual_seed(0)
    B, K, rows = args.batch, 50, 31
    model = spdm.Diffusion_DDIM(noise_steps=1000, obs_horizon=10, pred_horizon=rows - 1, observation_dim=135, prediction_dim=5,
                                model="UNet_Film", inpaint_horizon=1).cuda().eval()
    model.configure(precision="bf16", graph_steps=10, batch_max=B)
    model.use_ddim(K)
    dev = model.device
    devb = {k: v.to(dev) for k, v in synth_batch(B, 10, 1234).items()}
    x_T = torch.rand((B, 1, rows, 5), device=dev)
    plan = model._plan(B)
    model._bind_schedule(plan)
    inpaint = model.prepare_inpaint_vectors(devb).reshape(B, -1).contiguous()
    plan.encode_cond(devb["image"], devb["position"], devb["action"], devb["velocity"])
    for _ in range(3):
        plan.sample(x_T, inpaint=inpaint, seed=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        plan.sample(x_T, inpaint=inpaint, seed=1)
    e1.record()
    torch.cuda.synchronize()
    print("ABLATE_MS %.5f" % (e0.elapsed_time(e1) / args.reps / K))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--child", action="store_true")
    args = ap.parse_args()
    if args.child:
        return child(args)
    base = None
    print("batch %d, DDIM-50 graphed loop; ms per denoising step with a group of launches left out" % args.batch)
    print("| group left out | ms / step | marginal cost (us) | share of step |")
    print("|---|---|---|---|")
    for name, idx in GROUPS.items():
        env = dict(os.environ)
        env.setdefault("SPDM_FUSE_MODE", "0")   # the index map below is the one of the unfused GroupNorm-apply launches
        if idx:
            env["SPDM_SKIP_IDX"] = ",".join(map(str, idx))
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--batch", str(args.batch), "--reps", str(args.reps)],
                             env=env, capture_output=True, text=True)
        ms = None
        for line in out.stdout.splitlines():
            if line.startswith("ABLATE_MS"):
                ms = float(line.split()[1])
        if ms is None:
            print("| %s | failed: %s | | |" % (name, (out.stderr or out.stdout).strip().splitlines()[-1:] ), flush=True)
            continue
        if base is None:
            base = ms
        print("| %s | %.4f | %.1f | %.1f%% |" % (name, ms, (base - ms) * 1e3, 100 * (base - ms) / base), flush=True)
    print(json.dumps({"batch": args.batch, "base_ms_per_step": base}))


if __name__ == "__main__":
    main()
