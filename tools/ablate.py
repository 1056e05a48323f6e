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

# launch indices of one UNet_Film forward, bf16 path, in issue order (profiles/r01_launches_ddim50_b256_final.csv)
APPLY = [1, 3, 6, 8, 10, 12, 19, 21, 23, 25, 32, 34, 36, 38, 44, 46, 48, 50, 52, 54, 57, 59, 61, 63, 70, 72, 74, 76, 83, 85, 87, 89]
GROUPS = {
    "none": [],
    "gn_apply (32)": APPLY,
    "layernorm (6)": [13, 26, 39, 64, 77, 90],
    "in_proj (6)": [14, 27, 40, 65, 78, 91],
    "sdpa (6)": [15, 28, 41, 66, 79, 92],
    "sdpa sa6 (1)": [92],
    "attn tail (6)": [16, 29, 42, 67, 80, 93],
    "sa6 whole (4)": [90, 91, 92, 93],
    "attention whole (24)": [13, 14, 15, 16, 26, 27, 28, 29, 39, 40, 41, 42, 64, 65, 66, 67, 77, 78, 79, 80, 90, 91, 92, 93],
    "pool+upsample (6)": [4, 17, 30, 55, 68, 81],
    "conv 32x8 (5)": [2, 82, 84, 86, 88],
    "conv 16x4 (8)": [5, 7, 9, 11, 69, 71, 73, 75],
    "conv 8x2 (8)": [18, 20, 22, 24, 56, 58, 60, 62],
    "conv 4x1 (10)": [31, 33, 35, 37, 43, 45, 47, 49, 51, 53],
    "level 4x1 convs+applies (20)": list(range(31, 39)) + list(range(43, 55)),
    "everything but conv_in/outc (93)": list(range(1, 94)),
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
