"""oracle/unet_ref.py — TEST INFRASTRUCTURE (see oracle/__init__.py).

Functional fp32 CPU restatement of the reference noise estimator and vision encoder, driven by a
plain state_dict (SURVEY.md A.2 key names).  Each function cites the reference lines it follows
(paths relative to the reference repo root).  Pinned against the reference's own modules by
oracle/make_golden.py -> tests/golden/, checked in tests/test_oracle_golden.py.
"""
import math

import torch
import torch.nn.functional as F

STAGES = ("down1", "down2", "down3", "up1", "up2", "up3")


def pad_to(x, stride=8):
    """models/Unet_FiLmLayer.py:15-34 — zero pad H,W up to a multiple of `stride`, extra on the high side."""
    h, w = x.shape[-2:]
    new_h = h + stride - h % stride if h % stride > 0 else h
    new_w = w + stride - w % stride if w % stride > 0 else w
    lh, uh = int((new_h - h) / 2), int(new_h - h) - int((new_h - h) / 2)
    lw, uw = int((new_w - w) / 2), int(new_w - w) - int((new_w - w) / 2)
    pads = (lw, uw, lh, uh)
    return F.pad(x, pads, "constant", 0), pads


def unpad(x, pad):
    """models/Unet_FiLmLayer.py:36-41."""
    if pad[2] + pad[3] > 0:
        x = x[:, :, pad[2]:-pad[3], :]
    if pad[0] + pad[1] > 0:
        x = x[:, :, :, pad[0]:-pad[1]]
    return x


def pos_encoding(t, channels):
    """models/Unet_FiLmLayer.py:266-274.  t: (N,1) float."""
    inv_freq = 1.0 / (10000 ** (torch.arange(0, channels, 2) / channels))
    a = torch.sin(t.repeat(1, channels // 2) * inv_freq)
    b = torch.cos(t.repeat(1, channels // 2) * inv_freq)
    return torch.cat([a, b], dim=-1)


def double_conv(sd, p, x, taps=None):
    """models/Unet_FiLmLayer.py:85-115 — conv, GN(1,C), GELU, conv, the SAME GN again; no bias."""
    g, b = sd[p + ".norm.weight"], sd[p + ".norm.bias"]
    x = F.conv2d(x, sd[p + ".first.weight"], padding=1)
    if taps is not None:
        taps[p + ".first"] = x
    x = F.group_norm(x, 1, g, b)
    x = F.gelu(x)
    x = F.conv2d(x, sd[p + ".second.weight"], padding=1)
    if taps is not None:
        taps[p + ".second"] = x
    return F.group_norm(x, 1, g, b)


def _temb_film(sd, p, x, temb, cond):
    """models/Unet_FiLmLayer.py:165-177 — + Linear(SiLU(temb)); FiLM scale*x+bias from Linear(Mish(flatten(y)))."""
    e = F.linear(F.silu(temb), sd[p + ".emb_layer.1.weight"], sd[p + ".emb_layer.1.bias"])
    x = x + e[:, :, None, None]
    if cond is not None:
        c = F.linear(F.mish(cond).flatten(1), sd[p + ".cond_encoder.2.weight"], sd[p + ".cond_encoder.2.bias"])
        C = x.shape[1]
        x = c[:, :C, None, None] * x + c[:, C:, None, None]
    return x


def down(sd, p, x, temb, cond, taps=None):
    """models/Unet_FiLmLayer.py:158-179."""
    x = F.max_pool2d(x, 2)
    x = double_conv(sd, p + ".doubleConv1", x, taps)
    x = double_conv(sd, p + ".doubleConv2", x, taps)
    return _temb_film(sd, p, x, temb, cond)


def up(sd, p, x, skip, temb, cond, taps=None):
    """models/Unet_FiLmLayer.py:216-237 — bilinear x2 align_corners=True, cat([up, skip])."""
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    x = torch.cat([x, skip], dim=1)
    x = double_conv(sd, p + ".doubleConv1", x, taps)
    x = double_conv(sd, p + ".doubleConv2", x, taps)
    return _temb_film(sd, p, x, temb, cond)


def self_attention(sd, p, x, heads=4):
    """models/Unet_FiLmLayer.py:44-82 — LN, 4-head MHA (+res), LN, Linear, GELU, Linear (+res)."""
    B, C, H, W = x.shape
    L = H * W
    xt = x.reshape(B, C, L).swapaxes(1, 2)
    h = F.layer_norm(xt, (C,), sd[p + ".ln.weight"], sd[p + ".ln.bias"])
    qkv = F.linear(h, sd[p + ".attention.in_proj_weight"], sd[p + ".attention.in_proj_bias"])
    q, k, v = qkv.split(C, dim=-1)
    hd = C // heads
    q = q.reshape(B, L, heads, hd).transpose(1, 2)
    k = k.reshape(B, L, heads, hd).transpose(1, 2)
    v = v.reshape(B, L, heads, hd).transpose(1, 2)
    att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, L, C)
    o = F.linear(o, sd[p + ".attention.out_proj.weight"], sd[p + ".attention.out_proj.bias"])
    a = o + xt
    f = F.layer_norm(a, (C,), sd[p + ".ff_self.0.weight"], sd[p + ".ff_self.0.bias"])
    f = F.linear(f, sd[p + ".ff_self.1.weight"], sd[p + ".ff_self.1.bias"])
    f = F.gelu(f)
    f = F.linear(f, sd[p + ".ff_self.3.weight"], sd[p + ".ff_self.3.bias"])
    out = f + a
    return out.swapaxes(2, 1).reshape(B, C, H, W)


def unet_forward(sd, x, t, y=None, attention=True, time_dim=256, taps=None):
    """models/Unet_FiLmLayer.py:277-312 (attention) / Unet_FiLmLayer_noAttention.py:271-301.

    x (B,1,rows,dim) fp32; t (B,) or (1,) integer/float; y (B,1,T_obs,cond_dim) or None.
    `taps`: optional dict that receives named intermediate activations (NCHW)."""
    t = t.unsqueeze(-1).type(torch.float)
    temb = pos_encoding(t, time_dim)
    x, padding = pad_to(x, 8)
    sa = (lambda name, v: self_attention(sd, name, v)) if attention else (lambda name, v: v)

    def keep(name, v):
        if taps is not None:
            taps[name] = v
        return v

    x1 = keep("x1", double_conv(sd, "inc", x, taps))
    x2 = keep("down1", down(sd, "down1", x1, temb, y, taps))
    x2 = keep("x2", sa("sa1", x2))
    x3 = keep("down2", down(sd, "down2", x2, temb, y, taps))
    x3 = keep("x3", sa("sa2", x3))
    x4 = keep("down3", down(sd, "down3", x3, temb, y, taps))
    x4 = keep("x4", sa("sa3", x4))
    x5 = keep("bot1", double_conv(sd, "bot1", x4, taps))
    x5 = keep("bot2", double_conv(sd, "bot2", x5, taps))
    x5 = keep("x5", double_conv(sd, "bot3", x5, taps))
    u = keep("up1", up(sd, "up1", x5, x3, temb, y, taps))
    u = keep("u1", sa("sa4", u))
    u = keep("up2", up(sd, "up2", u, x2, temb, y, taps))
    u = keep("u2", sa("sa5", u))
    u = keep("up3", up(sd, "up3", u, x1, temb, y, taps))
    u = keep("u3", sa("sa6", u))
    out = F.conv2d(u, sd["outc.weight"], sd["outc.bias"])
    return unpad(out, padding)


def encoder_forward(esd, img, prefix=""):
    """models/encoder/autoencoder.py:11-20 — Autoencoder.encoder: (N,3,96,96) -> (N,128).
    Keys `{prefix}{0,2,4,7}.{weight,bias}`."""
    x = F.relu(F.conv2d(img, esd[prefix + "0.weight"], esd[prefix + "0.bias"], stride=2, padding=1))
    x = F.relu(F.conv2d(x, esd[prefix + "2.weight"], esd[prefix + "2.bias"], stride=2))
    x = F.relu(F.conv2d(x, esd[prefix + "4.weight"], esd[prefix + "4.bias"], stride=2))
    return F.linear(x.flatten(1), esd[prefix + "7.weight"], esd[prefix + "7.bias"])


def obs_cond(esd, batch, prefix=""):
    """models/diffusion_ddpm.py:317-330 — cat[pos(2), act(3), vel(2), encoder(img)(128)] -> (B,T,135)."""
    img = batch["image"]
    feat = encoder_forward(esd, img.flatten(end_dim=1), prefix).reshape(*img.shape[:2], -1)
    return torch.cat([batch["position"], batch["action"], batch["velocity"], feat], dim=-1)


def inpaint_vector(batch, inpaint_horizon):
    """models/diffusion_ddpm.py:340-348."""
    return torch.cat([batch["position"][:, -inpaint_horizon:, :], batch["action"][:, -inpaint_horizon:, :]], dim=-1)
