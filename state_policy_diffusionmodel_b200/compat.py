"""Drop-in aliases: makes `from models.diffusion_ddpm import *`, `from models.diffusion_ddim import *`,
`from models.Unet_FiLmLayer import *`, `from utils.schedulers import *` and
`from diffusers.schedulers.scheduling_dd{pm,im} import DD{PM,IM}Scheduler` resolve to this package, so
the reference's train.py / generate.py / evaluation/*.py run on the B200 path unchanged
(star-import contents listed in SURVEY.md 8b).

Only the LEAF modules of the hot path are replaced.  The reference's `models` / `utils` directories hold more than the hot
path (`utils.load_data`, `utils.print_utils`, `models.encoder.autoencoder`, ...; generate.py:4, train.py:9-11): when those
packages are importable (the scripts run from the reference checkout) they stay the real packages and keep resolving their
other sub-modules from disk; a stand-in package is created only where no real one exists.
"""
import importlib
import importlib.util
import sys
import types

_STANDIN = "_spdm_standin"


def _package(name):
    """The real package `name` when one is importable, else an (empty) stand-in package."""
    m = sys.modules.get(name)
    if m is not None and not getattr(m, _STANDIN, False):
        return m
    try:
        found = importlib.util.find_spec(name) if m is None else None
        if m is not None:  # a stand-in of an earlier call: look again, the reference may be on sys.path by now
            del sys.modules[name]
            try:
                found = importlib.util.find_spec(name)
            finally:
                sys.modules[name] = m
    except (ImportError, ValueError):
        found = None
    if found is not None:
        if m is not None:
            del sys.modules[name]
        try:
            return importlib.import_module(name)
        except Exception:
            if m is not None:
                sys.modules[name] = m
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []
        setattr(m, _STANDIN, True)
        sys.modules[name] = m
    return m


def _leaf(pkg, name, **attrs):
    """Registers `pkg.name` as a module holding `attrs` (replacing whatever the real package would import from disk)."""
    full = pkg.__name__ + "." + name
    m = types.ModuleType(full)
    m.__package__ = pkg.__name__
    for k, v in attrs.items():
        setattr(m, k, v)
    m.__all__ = [k for k in attrs if not k.startswith("_")]
    sys.modules[full] = m
    setattr(pkg, name, m)
    return m


def install_reference_aliases(override_diffusers=True):
    import numpy as np
    import torch
    import torch.nn as nn

    from . import diffusion, schedulers, unet
    models = _package("models")
    common = dict(torch=torch, nn=nn, np=np, UNet=unet.UNet, UNet_Film=unet.UNet_Film, UNet_Film_noAttention=unet.UNet_Film_noAttention,
                  DoubleConvolution=unet.DoubleConvolution, DownSample=unet.DownSample, UpSample=unet.UpSample,
                  SelfAttention=unet.SelfAttention)
    _leaf(models, "Unet_FiLmLayer", **common)
    _leaf(models, "Unet_FiLmLayer_noAttention", **common)
    _leaf(models, "simple_Unet", UNet=unet.UNet, torch=torch, nn=nn)
    # models/diffusion_ddpm.py:6-19 also star-exports pl / plt / datetime (scripts only rely on torch, np, nn and the classes)
    extra = dict(pl=diffusion.pl, datetime=diffusion.datetime)
    _leaf(models, "diffusion_ddpm", Diffusion_DDPM=diffusion.Diffusion_DDPM, DDPMScheduler=schedulers.DDPMScheduler, **extra, **common)
    _leaf(models, "diffusion_ddim", Diffusion_DDIM=diffusion.Diffusion_DDIM, Diffusion_DDPM=diffusion.Diffusion_DDPM,
          DDIMScheduler=schedulers.DDIMScheduler, **extra, **common)
    utils = _package("utils")
    _leaf(utils, "schedulers", linear_beta_schedule=schedulers.linear_beta_schedule,
          linear_beta_schedule_v2=schedulers.linear_beta_schedule_v2, cosine_beta_schedule=schedulers.cosine_beta_schedule,
          torch=torch, np=np)
    if override_diffusers:
        d = _package("diffusers")
        ds = sys.modules.get("diffusers.schedulers")
        if ds is None:
            ds = types.ModuleType("diffusers.schedulers")
            ds.__path__ = []
            sys.modules["diffusers.schedulers"] = ds
            setattr(d, "schedulers", ds)
        _leaf(ds, "scheduling_ddpm", DDPMScheduler=schedulers.DDPMScheduler)
        _leaf(ds, "scheduling_ddim", DDIMScheduler=schedulers.DDIMScheduler)
