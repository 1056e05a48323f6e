"""Microbenchmarks of the tcgen05 implicit-GEMM kernel on a B200 (`python tests/perf_conv.py`): one launch per U-Net
conv shape at batch B, with parts of the kernel switched off to locate the bottleneck.  Not a pytest file."""
import ctypes
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from state_policy_diffusionmodel_b200 import _lib  # noqa: E402

SHAPES = [  # (name, H, W, Cin, Cout, taps)
    ("L0 128->128 3x3", 32, 8, 128, 128, 9), ("L0 128->64 3x3", 32, 8, 128, 64, 9), ("L0 64->64 3x3", 32, 8, 64, 64, 9),
    ("L1 256->256 3x3", 16, 4, 256, 256, 9), ("L2 512->512 3x3", 8, 2, 512, 512, 9), ("L3 512->512 3x3", 4, 1, 512, 512, 9),
    ("L0 64->192 1x1", 32, 8, 64, 192, 1), ("L0 64->64 1x1", 32, 8, 64, 64, 1), ("L1 128->128 1x1", 16, 4, 128, 128, 1)]
MODES = [(0, "real"), (32, "no epilogue"), (16, "mma free-run"), (16 + 32, "mma free-run, no epilogue"), (64, "no swap"),
         (1, "no weight loads (plain kernel)"), (2, "no pixel loads (plain kernel)")]
if len(sys.argv) > 2 and sys.argv[2] == "timing":   # per-role wait cycles of block 0 (plain kernel), printed by the kernel
    MODES = [(4096, "timing"), (4096 + 1, "timing, no weight loads"), (4096 + 2, "timing, no pixel loads")]
    SHAPES = SHAPES[3:4]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    lib = _lib.load()
    torch.cuda.init()
    torch.zeros(1, device="cuda")
    for name, H, W, Cin, Cout, taps in SHAPES:
        flops = 2.0 * B * H * W * Cin * Cout * taps
        row = []
        for dbg, label in MODES:
            ms = ctypes.c_float()
            rc = lib.spdm_microbench_conv(H, W, B, Cin, Cout, taps, dbg, 1 if dbg & 4096 else 20, ctypes.byref(ms))
            if rc:
                print(name, "error", lib.spdm_last_error())
                break
            row.append("%s %.1fus (%.0f TF)" % (label, ms.value * 1e3, flops / (ms.value * 1e-3) / 1e12))
        print("%-18s B=%d | " % (name, B) + " | ".join(row), flush=True)


if __name__ == "__main__":
    main()
