"""Builds libspdm.so (hand-written sm_100a CUDA + the C ABI of include/spdm.h) in-tree with nvcc.

The library is compiled for sm_100a only (tcgen05 / TMEM / TMA instructions); there is no other
backend and no CPU fallback.  Every source is compiled to its own object (in parallel, only when it
or a header changed) and the objects are linked into the shared library.
"""
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libspdm.so")
SOURCES = ["kernels.cu", "simple_kernels.cu", "resnet.cu", "conv_tc.cu", "conv_tf32.cu", "sdpa_tc.cu", "attn_tc.cu", "attn_head.cu", "bwd_kernels.cu", "wgrad_tc.cu", "data_kernels.cu", "plan.cu"]
HEADERS = [os.path.join(CSRC, h) for h in ("common.cuh", "tc_ptx.cuh", "train.cuh", "train_impl.inl", "simple_unet.inl", "resnet.inl")] + [
    os.path.join(os.path.dirname(HERE), "include", "spdm.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libspdm.so cannot be built (set NVCC=/path/to/nvcc)")


def _mtime(path):
    return os.path.getmtime(path) if os.path.exists(path) else 0.0


def _obj(src):
    return os.path.join(OBJ, os.path.splitext(src)[0] + ".o")


def _obj_stale(src):
    t = _mtime(_obj(src))
    return t == 0.0 or any(_mtime(d) > t for d in [os.path.join(CSRC, src)] + HEADERS)


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(_mtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile csrc/*.cu into libspdm.so next to this file.  Returns the library path."""
    if not force and not is_stale():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)

    def compile_one(src):
        cmd = [nvcc] + NVCC_FLAGS + ["-c", "-o", _obj(src), os.path.join(CSRC, src)]
        if verbose:
            print(" ".join(cmd), flush=True)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s%s" % (src, res.stdout, res.stderr))

    todo = [s for s in SOURCES if force or _obj_stale(s)]
    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
        list(ex.map(compile_one, todo))
    cmd = [nvcc, "-shared", "-o", LIB] + [_obj(s) for s in SOURCES]
    if verbose:
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return LIB
