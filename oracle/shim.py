"""oracle/shim.py — TEST INFRASTRUCTURE, build-container only (needs /root/reference).

Lets the reference's Lightning wrappers (models/diffusion_ddpm.py, models/diffusion_ddim.py) import
in an environment without pytorch_lightning / diffusers / matplotlib / zarr by injecting minimal
stand-ins into sys.modules.  The diffusers scheduler classes resolve to oracle/schedulers.py (the
restated 0.17.1 arithmetic — the package itself is not installable here: parity unpinned).
Used by oracle/make_golden.py (on /root/reference) and oracle/ref_runner.py (on the copy in oracle/_ref, for bench.py's
CPU arm).
"""
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = "/root/reference"


class _Hparams(dict):
    __getattr__ = dict.get


class _LightningModule(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self.hparams = _Hparams()

    def save_hyperparameters(self, *a, **k):
        import inspect
        frame = inspect.currentframe().f_back
        args = inspect.getargvalues(frame)
        for name in args.args:
            if name != "self":
                self.hparams[name] = args.locals[name]

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    def log(self, *a, **k):
        pass

    @classmethod
    def load_from_checkpoint(cls, *a, **k):
        return cls()


def install(root=None):
    """`root`: directory holding the reference's `models/` and `utils/` (default: the read-only checkout)."""
    root = root or REFERENCE_ROOT
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")
        pl.LightningModule = _LightningModule
        pl.LightningDataModule = object
        pl.Trainer = object
        sys.modules["pytorch_lightning"] = pl
    from . import schedulers as S
    for name, attr, cls in (("scheduling_ddpm", "DDPMScheduler", S.RefDDPMScheduler),
                            ("scheduling_ddim", "DDIMScheduler", S.RefDDIMScheduler)):
        for modname in ("diffusers", "diffusers.schedulers", "diffusers.schedulers." + name):
            if modname not in sys.modules:
                sys.modules[modname] = types.ModuleType(modname)
        setattr(sys.modules["diffusers.schedulers." + name], attr, cls)
    for modname in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.animation", "zarr",
                    "numcodecs", "cv2"):
        try:
            __import__(modname)
        except Exception:
            m = types.ModuleType(modname)
            m.get_cmap = lambda *a, **k: None

            def _ga(name):
                if name.startswith("__"):
                    raise AttributeError(name)
                return lambda *a, **k: None
            m.__getattr__ = _ga
            sys.modules[modname] = m
    try:
        import PIL  # noqa: F401
    except Exception:
        pil = types.ModuleType("PIL")
        pil.Image = types.ModuleType("PIL.Image")
        sys.modules["PIL"] = pil
        sys.modules["PIL.Image"] = pil.Image
    if root not in sys.path:
        sys.path.insert(0, root)
