"""Drop-in aliases: makes `from models.diffusion_ddpm import *`, `from models.diffusion_ddim import *`,
`from models.Unet_FiLmLayer import *`, `from utils.schedulers import *` and
`from diffusers.schedulers.scheduling_dd{pm,im} import DD{PM,IM}Scheduler` resolve to this package, so
the reference's train.py / generate.py / evaluation/*.py run on the B200 path unchanged
(star-import contents listed in SURVEY.md 8b)."""
import sys
import types


def _module(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def install_reference_aliases(override_diffusers=True):
    import numpy as np
    import torch
    import torch.nn as nn

    from . import diffusion, schedulers, unet
    models = _module("models")
    models.__path__ = getattr(models, "__path__", [])
    common = dict(torch=torch, nn=nn, np=np, UNet_Film=unet.UNet_Film, UNet_Film_noAttention=unet.UNet_Film_noAttention,
                  DoubleConvolution=unet.DoubleConvolution, DownSample=unet.DownSample, UpSample=unet.UpSample,
                  SelfAttention=unet.SelfAttention)
    _module("models.Unet_FiLmLayer", **common)
    _module("models.Unet_FiLmLayer_noAttention", **common)
    _module("models.diffusion_ddpm", Diffusion_DDPM=diffusion.Diffusion_DDPM, DDPMScheduler=schedulers.DDPMScheduler,
            pl=diffusion.pl, **common)
    _module("models.diffusion_ddim", Diffusion_DDIM=diffusion.Diffusion_DDIM, Diffusion_DDPM=diffusion.Diffusion_DDPM,
            DDIMScheduler=schedulers.DDIMScheduler, pl=diffusion.pl, **common)
    utils = _module("utils")
    utils.__path__ = getattr(utils, "__path__", [])
    _module("utils.schedulers", linear_beta_schedule=schedulers.linear_beta_schedule,
            linear_beta_schedule_v2=schedulers.linear_beta_schedule_v2, cosine_beta_schedule=schedulers.cosine_beta_schedule,
            torch=torch, np=np)
    if override_diffusers:
        _module("diffusers")
        _module("diffusers.schedulers")
        _module("diffusers.schedulers.scheduling_ddpm", DDPMScheduler=schedulers.DDPMScheduler)
        _module("diffusers.schedulers.scheduling_ddim", DDIMScheduler=schedulers.DDIMScheduler)
