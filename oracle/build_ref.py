"""oracle/build_ref.py — TEST INFRASTRUCTURE, build-container only (needs /root/reference).

Recipe for `oracle/_ref/`: the reference's own hot-path modules, copied UNMODIFIED from the read-only checkout so that the
benchmark box (which has no /root/reference) can time the real reference on its host cores (`bench.py --impl reference`,
`cpu_baseline`).  `oracle/_ref/` is git-ignored -- it never enters the history -- but it is not gpurun-ignored, so it travels
to the GPU box like the built libspdm.so.  Run by `__graft_entry__.build()` whenever /root/reference is present:

    python -m oracle.build_ref

Only the files of SURVEY.md 8(a) are taken; their third-party imports (pytorch_lightning, diffusers, matplotlib, ...) are
absent on the box and are stood in for by oracle/shim.py at import time (the scheduler arithmetic is the restated
diffusers 0.17.1 of oracle/schedulers.py: parity unpinned, as everywhere).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ["models/Unet_FiLmLayer.py", "models/Unet_FiLmLayer_noAttention.py", "models/simple_Unet.py", "models/diffusion_ddpm.py",
         "models/diffusion_ddim.py", "models/encoder/autoencoder.py", "utils/schedulers.py", "utils/print_utils.py", "utils/plot_utils.py"]


def build(reference_root="/root/reference", verbose=True):
    if not os.path.isdir(reference_root):
        if verbose:
            print("oracle/_ref: %s is absent, keeping whatever is already in %s" % (reference_root, DEST))
        return os.path.isdir(DEST)
    manifest = {}
    for rel in FILES:
        src = os.path.join(reference_root, rel)
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": reference_root, "sha256": manifest}, f, indent=1)
    if verbose:
        print("oracle/_ref: %d reference files copied unmodified from %s" % (len(FILES), reference_root))
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
