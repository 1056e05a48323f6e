"""state_policy_diffusionmodel_b200 — B200-native denoising hot path of State_Policy_DiffusionModel.

Public surface (mirrors the reference's modules):
    UNet_Film, UNet_Film_noAttention        <- models/Unet_FiLmLayer.py, models/Unet_FiLmLayer_noAttention.py
    UNet                                    <- models/simple_Unet.py (legacy model='UNet' default; fp32 path, inference only)
    Diffusion_DDPM, Diffusion_DDIM          <- models/diffusion_ddpm.py, models/diffusion_ddim.py
    DDPMScheduler, DDIMScheduler            <- diffusers 0.17.1 objects the reference instantiates
    linear_beta_schedule, linear_beta_schedule_v2, cosine_beta_schedule   <- utils/schedulers.py
    DenoisePlan                             <- thin handle on the C ABI (include/spdm.h, libspdm.so)
    DeviceWindowDataset                     <- utils/load_data.py CarRacingDataset with the arrays resident in HBM
All compute runs in hand-written sm_100a CUDA kernels behind the C ABI; there is no CPU fallback.
"""
from .schedulers import (DDIMScheduler, DDPMScheduler, cosine_beta_schedule, linear_beta_schedule,  # noqa: F401
                         linear_beta_schedule_v2)
from .engine import DenoisePlan  # noqa: F401
from .unet import UNet, UNet_Film, UNet_Film_noAttention  # noqa: F401
from .diffusion import Diffusion_DDIM, Diffusion_DDPM, SamplingPipeline  # noqa: F401
from .compat import install_reference_aliases  # noqa: F401
from .data import DeviceWindowDataset, create_sample_indices_sparse  # noqa: F401

__all__ = ["UNet", "UNet_Film", "UNet_Film_noAttention", "Diffusion_DDPM", "Diffusion_DDIM", "DDPMScheduler", "DDIMScheduler",
           "linear_beta_schedule", "linear_beta_schedule_v2", "cosine_beta_schedule", "DenoisePlan", "SamplingPipeline",
           "install_reference_aliases", "DeviceWindowDataset", "create_sample_indices_sparse"]
