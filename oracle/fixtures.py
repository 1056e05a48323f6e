"""oracle/fixtures.py — TEST INFRASTRUCTURE (see oracle/__init__.py).

Deterministic synthetic weights and inputs shared by the oracle, the golden generator, the GPU
parity tests and bench.py.  Weights are drawn WITHOUT the reference modules (so they can be
regenerated on the GPU box, where /root/reference does not exist): shapes follow SURVEY.md A.2,
values are PyTorch-default-like (uniform +-1/sqrt(fan_in); norm affine perturbed off 1/0 so that
gamma/beta paths are exercised).  oracle/make_golden.py loads them into the REAL reference modules
with load_state_dict(strict=True), which also pins the key/shape contract.
"""
import math

import torch


def _unet_shapes(attention=True, cond_dim=1350, time_dim=256, in_ch=1, out_ch=1):
    s = {}

    def dc(p, ci, co):
        s[p + ".first.weight"] = (co, ci, 3, 3)
        s[p + ".second.weight"] = (co, co, 3, 3)
        s[p + ".norm.weight"] = (co,)
        s[p + ".norm.bias"] = (co,)

    def stage(p, ci, co):
        dc(p + ".doubleConv1", ci, ci)
        dc(p + ".doubleConv2", ci, co)
        s[p + ".emb_layer.1.weight"] = (co, time_dim)
        s[p + ".emb_layer.1.bias"] = (co,)
        if cond_dim is not None:
            s[p + ".cond_encoder.2.weight"] = (2 * co, cond_dim)
            s[p + ".cond_encoder.2.bias"] = (2 * co,)

    def sa(p, c):
        s[p + ".attention.in_proj_weight"] = (3 * c, c)
        s[p + ".attention.in_proj_bias"] = (3 * c,)
        s[p + ".attention.out_proj.weight"] = (c, c)
        s[p + ".attention.out_proj.bias"] = (c,)
        s[p + ".ln.weight"] = (c,)
        s[p + ".ln.bias"] = (c,)
        s[p + ".ff_self.0.weight"] = (c,)
        s[p + ".ff_self.0.bias"] = (c,)
        s[p + ".ff_self.1.weight"] = (c, c)
        s[p + ".ff_self.1.bias"] = (c,)
        s[p + ".ff_self.3.weight"] = (c, c)
        s[p + ".ff_self.3.bias"] = (c,)

    dc("inc", in_ch, 64)
    stage("down1", 64, 128)
    stage("down2", 128, 256)
    stage("down3", 256, 256)
    dc("bot1", 256, 512)
    dc("bot2", 512, 512)
    dc("bot3", 512, 256)
    stage("up1", 512, 128)
    stage("up2", 256, 64)
    stage("up3", 128, 64)
    if attention:
        for name, c in (("sa1", 128), ("sa2", 256), ("sa3", 256), ("sa4", 128), ("sa5", 64), ("sa6", 64)):
            sa(name, c)
    s["outc.weight"] = (out_ch, 64, 1, 1)
    s["outc.bias"] = (out_ch,)
    return s


def _draw(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k in sorted(shapes):
        shp = shapes[k]
        if (k.endswith("norm.weight") or k.endswith("ln.weight") or k.endswith("ff_self.0.weight")):
            sd[k] = 1.0 + 0.1 * (2 * torch.rand(shp, generator=g) - 1)
        elif (k.endswith("norm.bias") or k.endswith("ln.bias") or k.endswith("ff_self.0.bias")):
            sd[k] = 0.1 * (2 * torch.rand(shp, generator=g) - 1)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            if len(shp) == 1:  # linear/conv bias: use a modest fixed bound
                bound = 0.05
            else:
                bound = 1.0 / math.sqrt(fan_in)
            sd[k] = bound * (2 * torch.rand(shp, generator=g) - 1)
    return sd


def make_unet_weights(attention=True, cond_dim=1350, seed=0):
    return _draw(_unet_shapes(attention, cond_dim), seed)


def _simple_unet_shapes(cond_dim=1350, time_dim=256):
    s = {}

    def dc(p, ci, co):
        s[p + ".first.weight"] = (co, ci, 3, 3)
        s[p + ".second.weight"] = (co, co, 3, 3)
        s[p + ".norm.weight"] = (co,)
        s[p + ".norm.bias"] = (co,)

    dc("input_conv", 1, 16)
    for name, ci, co in (("down1", 16, 32), ("down2", 64, 128), ("down3", 160, 256), ("up1", 448, 128), ("up2", 224, 64), ("up3", 112, 32)):
        dc(name + ".doubleConv1", ci, ci)
        dc(name + ".doubleConv2", ci, co)
        s[name + ".emb_layer.1.weight"] = (co, time_dim)
        s[name + ".emb_layer.1.bias"] = (co,)
        s[name + ".cond_emb_layer.1.weight"] = (32, cond_dim)
        s[name + ".cond_emb_layer.1.bias"] = (32,)
    s["outc.weight"] = (1, 64, 1, 1)
    s["outc.bias"] = (1,)
    return s


def make_simple_unet_weights(cond_dim=1350, seed=0, noise_steps=1000):
    """State dict of the legacy simple U-Net (models/simple_Unet.py:260-339) incl. its PositionalEncoding buffer."""
    from .simple_unet_ref import pos_table
    sd = _draw(_simple_unet_shapes(cond_dim), seed)
    sd["pos_encoding.pos_encoding"] = pos_table(noise_steps + 1)
    return sd


def make_resnet_weights(seed=2):
    """State dict of the ResNet18-GroupNorm vision encoder (models/Unet_FiLmLayer.py:316-386, torchvision key names): convs
    kaiming-like (std sqrt(2 / fan_out), as torchvision initialises them), GroupNorm affine perturbed off 1 / 0."""
    from .resnet_ref import shapes
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in sorted(shapes().items()):
        if len(shp) == 4:
            fan_out = shp[0] * shp[2] * shp[3]
            sd[k] = math.sqrt(2.0 / fan_out) * torch.randn(shp, generator=g)
        elif k.endswith(".weight"):
            sd[k] = 1.0 + 0.1 * (2 * torch.rand(shp, generator=g) - 1)
        else:
            sd[k] = 0.1 * (2 * torch.rand(shp, generator=g) - 1)
    return sd


def make_encoder_weights(seed=1):
    shapes = {"0.weight": (16, 3, 2, 2), "0.bias": (16,), "2.weight": (32, 16, 2, 2), "2.bias": (32,),
              "4.weight": (64, 32, 2, 2), "4.bias": (64,), "7.weight": (128, 9216), "7.bias": (128,)}
    return _draw(shapes, seed)


def make_batch(B, T_obs=10, seed=1234):
    """SURVEY.md 8(d) synthetic CarRacing-shaped conditioning."""
    g = torch.Generator().manual_seed(seed)
    image = torch.rand((B, T_obs, 3, 96, 96), generator=g)
    dpos = 0.02 * torch.randn((B, T_obs, 2), generator=g)
    dpos[:, 0] = 0
    position = torch.cumsum(dpos, dim=1)
    velocity = 2 * torch.rand((B, T_obs, 2), generator=g) - 1
    action = 2 * torch.rand((B, T_obs, 3), generator=g) - 1
    return {"image": image, "position": position, "velocity": velocity, "action": action}


def make_xT(B, rows=31, dim=5, seed=77):
    g = torch.Generator().manual_seed(seed)
    return torch.rand((B, 1, rows, dim), generator=g)  # U[0,1) as models/diffusion_ddpm.py:252


def make_noise(K, B, rows=31, dim=5, seed=99):
    g = torch.Generator().manual_seed(seed)
    return torch.randn((K, B, 1, rows, dim), generator=g)
