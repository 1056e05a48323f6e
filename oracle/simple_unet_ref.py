"""oracle/simple_unet_ref.py — TEST INFRASTRUCTURE (see oracle/__init__.py).

Functional fp32 CPU restatement of the reference's legacy `UNet` (models/simple_Unet.py:260-339, the `model='UNet'` default of
Diffusion_DDPM, models/diffusion_ddpm.py:60-62), driven by a plain state_dict with the module's own key names.  Eval mode
(PositionalEncoding's dropout is the identity).  Pinned against the reference module by oracle/make_golden.py ->
tests/golden/simple_unet.npz, checked in tests/test_oracle_golden.py.
"""
import math

import torch
import torch.nn.functional as F

from .unet_ref import pad_to, unpad

STAGES = (("down1", 16, 32), ("down2", 64, 128), ("down3", 160, 256), ("up1", 448, 128), ("up2", 224, 64), ("up3", 112, 32))


def pos_table(max_len, embedding_dim=256):
    """models/simple_Unet.py:226-236 — the registered buffer `pos_encoding.pos_encoding`: sin on even, cos on odd columns."""
    pe = torch.zeros(max_len, embedding_dim)
    position = torch.arange(start=0, end=max_len).unsqueeze(1)
    div_term = torch.exp(-math.log(10000.0) * torch.arange(0, embedding_dim, 2).float() / embedding_dim)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def double_conv(sd, p, x, residual=False):
    """models/simple_Unet.py:104-125 — conv, GN(1,C), GELU, conv, the SAME GN, [+ input], GELU."""
    g, b = sd[p + ".norm.weight"], sd[p + ".norm.bias"]
    x_res = x
    x = F.conv2d(x, sd[p + ".first.weight"], padding=1)
    x = F.gelu(F.group_norm(x, 1, g, b))
    x = F.conv2d(x, sd[p + ".second.weight"], padding=1)
    x = F.group_norm(x, 1, g, b)
    return F.gelu(x + x_res) if residual else F.gelu(x)


def _tail(sd, p, x, t_emb, cond):
    """models/simple_Unet.py:152-166 — + Linear(SiLU(t)) broadcast over the map; cat the 32-channel Linear(SiLU(cond)) map."""
    e = F.linear(F.silu(t_emb), sd[p + ".emb_layer.1.weight"], sd[p + ".emb_layer.1.bias"])
    x = x + e[:, :, None, None]
    if cond is not None:
        ce = F.linear(F.silu(cond.reshape(cond.shape[0], -1)), sd[p + ".cond_emb_layer.1.weight"], sd[p + ".cond_emb_layer.1.bias"])
        x = torch.cat([x, ce[:, :, None, None].expand(-1, -1, x.shape[-2], x.shape[-1])], dim=1)
    return x


def down(sd, p, x, t_emb, cond):
    x = F.max_pool2d(x, 2, 2)
    x = double_conv(sd, p + ".doubleConv1", x, residual=True)
    x = double_conv(sd, p + ".doubleConv2", x)
    return _tail(sd, p, x, t_emb, cond)


def up(sd, p, x, x_res, t_emb, cond):
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    x = torch.cat([x, x_res], dim=1)
    x = double_conv(sd, p + ".doubleConv1", x, residual=True)
    x = double_conv(sd, p + ".doubleConv2", x)
    return _tail(sd, p, x, t_emb, cond)


def unet_forward(sd, x, t, y):
    """models/simple_Unet.py:307-327.  x (B,1,rows,dim); t (B,) or (1,) long; y (B,1,T,cond_dim)."""
    x, padding = pad_to(x, 8)
    t_emb = sd["pos_encoding.pos_encoding"][t].squeeze(-1)          # :238-242 (eval: no dropout)
    if t_emb.dim() == 1:
        t_emb = t_emb[None, :]
    x1 = double_conv(sd, "input_conv", x)
    x2 = down(sd, "down1", x1, t_emb, y)
    x3 = down(sd, "down2", x2, t_emb, y)
    x4 = down(sd, "down3", x3, t_emb, y)
    x = up(sd, "up1", x4, x3, t_emb, y)
    x = up(sd, "up2", x, x2, t_emb, y)
    x = up(sd, "up3", x, x1, t_emb, y)
    logits = F.conv2d(x, sd["outc.weight"], sd["outc.bias"])
    return unpad(logits, padding)
