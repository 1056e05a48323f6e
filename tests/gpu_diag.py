"""Layer-by-layer diagnosis of the CUDA U-Net against the CPU oracle (run on a B200: `python tests/gpu_diag.py`).

Not a pytest file: prints the relative error of every tapped activation for the fp32 and bf16 plans,
so that one gpurun call localises a wrong kernel.  TEST INFRASTRUCTURE (uses oracle/)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fixtures, unet_ref  # noqa: E402
from state_policy_diffusionmodel_b200 import DenoisePlan  # noqa: E402

TAPS_ATTN = ["inc.first", "inc.second", "inc", "down1.doubleConv1.first", "down1.doubleConv1.second", "down1", "sa1", "down2", "sa2",
             "down3", "sa3", "bot1", "bot2", "bot3", "up1.doubleConv1.first", "up1", "sa4", "up2", "sa5", "up3", "sa6"]
ORACLE_NAME = {"sa1": "x2", "sa2": "x3", "sa3": "x4", "sa4": "u1", "sa5": "u2", "sa6": "u3", "bot3": "x5", "inc": "x1"}


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def main():
    attention = "--noattn" not in sys.argv
    precisions = [p for p in ("fp32", "bf16") if ("--" + p) in sys.argv] or ["fp32", "bf16"]
    B = 3
    torch.manual_seed(0)
    sd = fixtures.make_unet_weights(attention=attention, seed=0)
    g = torch.Generator().manual_seed(100)
    x = torch.rand((B, 1, 31, 5), generator=g)
    y = torch.randn((B, 1, 10, 135), generator=g)
    t = torch.tensor([999, 500, 3])
    taps = {}
    with torch.no_grad():
        ref = unet_ref.unet_forward(sd, x, t, y, attention=attention, taps=taps)
    for prec in precisions:
        print("=== precision", prec, "attention", attention, flush=True)
        plan = DenoisePlan(attention=attention, precision=prec, batch_max=B, rows=31, dim=5, graph_steps=0)
        plan.load_unet_state_dict(sd)
        print("missing:", [m for m in plan.missing_weights() if not m.startswith("vision")], "workspace MB", plan.workspace_bytes / 1e6)
        t0 = time.time()
        out = plan.unet_forward(x, t, y)
        torch.cuda.synchronize()
        print("forward ok in %.3fs; out rel err = %.3e" % (time.time() - t0, rel(out.cpu(), ref)), flush=True)
        names = TAPS_ATTN if attention else [n for n in TAPS_ATTN if not n.startswith("sa")]
        for name in names:
            oname = ORACLE_NAME.get(name, name)
            if not attention and name in ("down1", "down2", "down3"):
                oname = name
            r = taps[oname]
            _, got = plan.debug_forward(x, t, y, name, tuple(r.shape))
            print("  %-28s shape %-18s rel err %.3e" % (name, tuple(r.shape), rel(got.cpu(), r)), flush=True)
        plan.close()


if __name__ == "__main__":
    main()
