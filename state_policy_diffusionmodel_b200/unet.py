"""`UNet_Film` / `UNet_Film_noAttention` with the reference's constructor, forward signature and state_dict.

Mirrors models/Unet_FiLmLayer.py:240-312 and models/Unet_FiLmLayer_noAttention.py:240-301 of the
reference: same sub-module names, same parameter shapes, same construction order (so a given
`torch.manual_seed` yields the same initial weights and reference checkpoints load with strict=True).
The sub-modules are parameter containers only — `forward` hands the whole network to libspdm
(hand-written sm_100a kernels) through `DenoisePlan`; nothing is computed by torch operators.
"""
import torch
import torch.nn as nn

from .engine import DenoisePlan


class DoubleConvolution(nn.Module):
    """Parameters of models/Unet_FiLmLayer.py:85-115 (two bias-free 3x3 convs sharing one GroupNorm(1, C))."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.first = nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1, bias=False)
        self.second = nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1, bias=False)
        self.act = nn.GELU()
        self.norm = nn.GroupNorm(1, out_channels)


class _Stage(nn.Module):
    """Parameters of DownSample / UpSample (models/Unet_FiLmLayer.py:118-156, 183-214)."""

    def __init__(self, in_channels, out_channels, embeddedTime_dim=256, cond_dim=None):
        super().__init__()
        self.doubleConv1 = DoubleConvolution(in_channels, in_channels)
        self.doubleConv2 = DoubleConvolution(in_channels, out_channels)
        self.emb_layer = nn.Sequential(nn.SiLU(), nn.Linear(embeddedTime_dim, out_channels))
        if cond_dim is not None:
            self.out_channels = out_channels
            self.cond_encoder = nn.Sequential(nn.Mish(), nn.Flatten(1, -1), nn.Linear(cond_dim, out_channels * 2),
                                              nn.Unflatten(-1, (-1, 1)))


class DownSample(_Stage):
    pass


class UpSample(_Stage):
    pass


class SelfAttention(nn.Module):
    """Parameters of models/Unet_FiLmLayer.py:44-69 (4-head MHA + LN + 2-layer feed-forward)."""

    def __init__(self, channels):
        super().__init__()
        self.channels = channels
        self.attention = nn.MultiheadAttention(channels, 4, batch_first=True)
        self.ln = nn.LayerNorm([channels])
        self.ff_self = nn.Sequential(nn.LayerNorm([channels]), nn.Linear(channels, channels), nn.GELU(),
                                     nn.Linear(channels, channels))


class _UNetBase(nn.Module):
    _attention = True
    _simple = False

    def __init__(self, in_channels, out_channels, noise_steps, time_dim=256, global_cond_dim=None):
        super().__init__()
        if in_channels != 1 or out_channels != 1:
            raise NotImplementedError("the B200 path implements the reference's wiring in_channels = out_channels = 1 "
                                      "(models/diffusion_ddpm.py:76-82)")
        self.time_dim = time_dim
        self.noise_steps = noise_steps
        self.global_cond_dim = global_cond_dim
        sa = self._attention
        # construction order == reference, so that seeded default init matches
        self.inc = DoubleConvolution(in_channels, 64)
        self.down1 = DownSample(64, 128, cond_dim=global_cond_dim)
        if sa:
            self.sa1 = SelfAttention(128)
        self.down2 = DownSample(128, 256, cond_dim=global_cond_dim)
        if sa:
            self.sa2 = SelfAttention(256)
        self.down3 = DownSample(256, 256, cond_dim=global_cond_dim)
        if sa:
            self.sa3 = SelfAttention(256)
        self.bot1 = DoubleConvolution(256, 512)
        self.bot2 = DoubleConvolution(512, 512)
        self.bot3 = DoubleConvolution(512, 256)
        self.up1 = UpSample(512, 128, cond_dim=global_cond_dim)
        if sa:
            self.sa4 = SelfAttention(128)
        self.up2 = UpSample(256, 64, cond_dim=global_cond_dim)
        if sa:
            self.sa5 = SelfAttention(64)
        self.up3 = UpSample(128, 64, cond_dim=global_cond_dim)
        if sa:
            self.sa6 = SelfAttention(64)
        self.outc = nn.Conv2d(64, out_channels, kernel_size=1)
        # B200 execution options (not part of the state_dict)
        self.precision = "bf16"   # "bf16": tcgen05 implicit-GEMM path; "fp32": CUDA-core parity path
        self.batch_max = 0        # 0 = grow on demand
        self.split = 1            # sub-batches run concurrently inside the graphed sampling loop
        self.encoder = "autoencoder"   # vision encoder the plan carries: "autoencoder" | "resnet18" (set by the Diffusion_DDPM owner)
        self._plan = None
        self._plan_key = None
        self._weights_tag = None
        self._weights_epoch = 0   # bumped when the weights change behind torch's back (fused native optimizer step)

    # ----------------------------------------------------------------------------------------
    def configure(self, precision=None, batch_max=None):
        if precision is not None:
            self.precision = precision
        if batch_max is not None:
            self.batch_max = int(batch_max)
        return self

    def _tag(self):
        return (self._weights_epoch,) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    def plan_for(self, B, rows, dim, obs_horizon=None, cond_dim=None, inpaint_rows=None, graph_steps=None):
        """Returns a DenoisePlan able to run B samples of rows x dim, (re)building it and (re)loading weights when
        needed.  `inpaint_rows` / `graph_steps` = None means "whatever the current plan has"."""
        dev = self.outc.weight.device
        if dev.type != "cuda":
            raise RuntimeError("spdm U-Net: parameters are on %s — move the module to a CUDA device (no CPU fallback)" % dev)
        G = self.global_cond_dim
        if G is None:
            T, cd = 1, 0
        elif obs_horizon is not None and cond_dim is not None and obs_horizon * cond_dim == G:
            T, cd = int(obs_horizon), int(cond_dim)
        else:
            T, cd = 1, int(G)
        old = self._plan
        if old is not None:
            if inpaint_rows is None:
                inpaint_rows = old.inpaint_rows
            if graph_steps is None:
                graph_steps = old.cfg.graph_steps
            if obs_horizon is None and old.obs_horizon * old.cond_dim == T * cd:
                T, cd = old.obs_horizon, old.cond_dim
        inpaint_rows = 0 if inpaint_rows is None else int(inpaint_rows)
        graph_steps = 1 if graph_steps is None else int(graph_steps)
        key = (self.precision, rows, dim, T, cd, inpaint_rows, graph_steps, str(dev), self.split, self.encoder)
        if old is None or self._plan_key != key or old.batch_max < B:
            cap = max(int(B), self.batch_max)
            if old is not None:
                if self._plan_key == key:
                    cap = max(cap, old.batch_max)
                old.close()
            self._plan = DenoisePlan(attention=self._attention, precision=self.precision, batch_max=cap, rows=rows, dim=dim,
                                     obs_horizon=T, cond_dim=cd, inpaint_rows=inpaint_rows, time_dim=self.time_dim, device=dev,
                                     graph_steps=graph_steps, split=self.split, simple=self._simple, encoder=self.encoder)
            self._plan_key = key
            self._weights_tag = None
        tag = self._tag()
        if self._weights_tag != tag:
            self._plan.load_unet_state_dict(self.state_dict())
            self._weights_tag = tag
        return self._plan

    def forward(self, x, t, y=None):
        """x (B,1,rows,dim); t (B,) or (1,) integer timesteps; y (B,1,T_obs,cond_dim) or None -> (B,1,rows,dim)."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and self.training:
            raise NotImplementedError("spdm: the U-Net module alone has no autograd graph — the training step (forward + "
                                      "backward + optimizer) is Diffusion_DDPM.training_step / process_single_batch; "
                                      "call this forward under torch.no_grad() / .eval()")
        if not x.is_cuda:
            raise RuntimeError("spdm U-Net: input is on the CPU (no CPU fallback)")
        B, _, rows, dim = x.shape
        if y is not None and self.global_cond_dim is None:
            raise ValueError("conditioning passed to a U-Net built with global_cond_dim=None")
        T = cd = None
        if y is not None and y.dim() == 4:
            T, cd = int(y.shape[2]), int(y.shape[3])
        plan = self.plan_for(B, rows, dim, T, cd)
        out = plan.unet_forward(x, t.reshape(-1), y)
        return out.to(x.dtype)


class _FilmPosEncoding:
    def pos_encoding(self, t, channels):
        """models/Unet_FiLmLayer.py:266-274 (kept for API parity; the kernels compute it on the device)."""
        inv_freq = 1.0 / (10000 ** (torch.arange(0, channels, 2, device=t.device) / channels))
        a = torch.sin(t.repeat(1, channels // 2) * inv_freq)
        b = torch.cos(t.repeat(1, channels // 2) * inv_freq)
        return torch.cat([a, b], dim=-1)


class UNet_Film(_FilmPosEncoding, _UNetBase):
    """FiLM U-Net with six SelfAttention blocks (reference models/Unet_FiLmLayer.py:240)."""
    _attention = True


class UNet_Film_noAttention(_FilmPosEncoding, _UNetBase):
    """FiLM U-Net without attention (reference models/Unet_FiLmLayer_noAttention.py:240)."""
    _attention = False


# ---------------------------------------------------------------------------------------------------------------------
# the legacy simple U-Net: models/simple_Unet.py (the reference's `model='UNet'` default, models/diffusion_ddpm.py:60-62)
# ---------------------------------------------------------------------------------------------------------------------
class _SimpleDoubleConvolution(DoubleConvolution):
    """Parameters of models/simple_Unet.py:82-125 (same two bias-free 3x3 convs + one GroupNorm; `residual` changes the forward only)."""

    def __init__(self, in_channels, out_channels, residual=False):
        super().__init__(in_channels, out_channels)
        self.residual = residual


class _SimpleStage(nn.Module):
    """Parameters of the simple DownSample / UpSample (models/simple_Unet.py:128-211)."""

    def __init__(self, in_channels, out_channels, embeddedTime_dim=256, cond_dim=None):
        super().__init__()
        self.cond_dim = cond_dim
        self.doubleConv1 = _SimpleDoubleConvolution(in_channels, in_channels, residual=True)
        self.doubleConv2 = _SimpleDoubleConvolution(in_channels, out_channels)
        self.emb_layer = nn.Sequential(nn.SiLU(), nn.Linear(embeddedTime_dim, out_channels))
        if cond_dim is not None:
            self.cond_emb_layer = nn.Sequential(nn.SiLU(), nn.Linear(in_features=cond_dim, out_features=32))


class PositionalEncoding(nn.Module):
    """models/simple_Unet.py:214-242: the (max_len, embedding_dim) sin / cos table is a registered buffer (part of the state_dict);
    dropout applies in training mode only -- the B200 path implements the eval-mode lookup."""

    def __init__(self, embedding_dim, dropout=0.1, max_len=1000, apply_dropout=True):
        super().__init__()
        import math
        self.dropout = nn.Dropout(p=dropout)
        self.apply_dropout = apply_dropout
        pos_encoding = torch.zeros(max_len, embedding_dim)
        position = torch.arange(start=0, end=max_len).unsqueeze(1)
        div_term = torch.exp(-math.log(10000.0) * torch.arange(0, embedding_dim, 2).float() / embedding_dim)
        pos_encoding[:, 0::2] = torch.sin(position * div_term)
        pos_encoding[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer(name='pos_encoding', tensor=pos_encoding)


class UNet(_UNetBase):
    """Legacy simple U-Net (reference models/simple_Unet.py:260-339): 16/32/128/256-channel DoubleConvolutions with a trailing GELU
    (residual on the first of every stage), table positional encoding, a 32-channel conditioning map concatenated after every
    stage, no attention, no FiLM.  Runs on the fp32 CUDA-core path of libspdm (its channel counts are not multiples of the
    64-wide tensor-core operand tiles), inference only."""
    _attention = False
    _simple = True

    def __init__(self, in_channels, out_channels, noise_steps=1000, time_dim=256, global_cond_dim=None):
        nn.Module.__init__(self)
        if in_channels != 1 or out_channels != 1:
            raise NotImplementedError("the B200 path implements the reference's wiring in_channels = out_channels = 1")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.noise_steps, self.time_dim, self.global_cond_dim = noise_steps, time_dim, global_cond_dim
        self.pos_encoding = PositionalEncoding(embedding_dim=time_dim, max_len=self.noise_steps + 1)
        self.input_conv = _SimpleDoubleConvolution(in_channels, 16)
        self.down1 = _SimpleStage(16, 32, cond_dim=global_cond_dim)
        self.down2 = _SimpleStage(64, 128, cond_dim=global_cond_dim)
        self.down3 = _SimpleStage(160, 256, cond_dim=global_cond_dim)
        self.up1 = _SimpleStage(288 + 160, 128, cond_dim=global_cond_dim)
        self.up2 = _SimpleStage(160 + 64, 64, cond_dim=global_cond_dim)
        self.up3 = _SimpleStage(96 + 16, 32, cond_dim=global_cond_dim)
        self.outc = nn.Conv2d(in_channels=64, out_channels=out_channels, kernel_size=(1, 1))
        self.precision = "fp32"
        self.batch_max = 0
        self.split = 1
        self.encoder = "autoencoder"
        self._plan = None
        self._plan_key = None
        self._weights_tag = None
        self._weights_epoch = 0

    def configure(self, precision=None, batch_max=None):
        if precision not in (None, "fp32"):
            raise ValueError("the simple U-Net runs on the fp32 path only")
        if batch_max is not None:
            self.batch_max = int(batch_max)
        return self

    def _tag(self):
        return super()._tag() + ((self.pos_encoding.pos_encoding.data_ptr(), self.pos_encoding.pos_encoding._version),)

    def forward(self, x, t, y=None):
        if y is None:
            raise ValueError("the simple U-Net needs the conditioning `y`: its stage widths include the 32-channel cond_emb "
                             "(models/simple_Unet.py:268-276; the reference fails on a channel mismatch without it)")
        if self.training and torch.is_grad_enabled():
            raise NotImplementedError("spdm: the simple U-Net is inference-only on the B200 path (eval mode / torch.no_grad()); "
                                      "train the FiLM U-Nets (model='UNet_Film' / 'UNet_FilmnoAttention') natively")
        self.precision = "fp32"
        return super().forward(x, t, y)


# ---------------------------------------------------------------------------------------------------------------------
# ResNet18 with GroupNorm: `VisionEncoder()` of the reference (models/Unet_FiLmLayer.py:316-386)
# ---------------------------------------------------------------------------------------------------------------------
class _GNBasicBlock(nn.Module):
    """Parameters of torchvision's BasicBlock after replace_bn_with_gn (models/Unet_FiLmLayer.py:368-380): conv1 / bn1 / conv2 / bn2
    (+ downsample.{0,1}); `bnX` are GroupNorm(C // 16, C) -- the attribute names stay those of the replaced BatchNorm2d."""

    def __init__(self, inplanes, planes, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.GroupNorm(planes // 16, planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn2 = nn.GroupNorm(planes // 16, planes)
        self.downsample = None
        if stride != 1 or inplanes != planes:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes, kernel_size=1, stride=stride, bias=False),
                                            nn.GroupNorm(planes // 16, planes))


class ResNet18GN(nn.Module):
    """`VisionEncoder()` = get_resnet('resnet18') + replace_bn_with_gn (models/Unet_FiLmLayer.py:316-386): same sub-module names,
    parameter shapes and construction order as torchvision.models.resnet18 with fc = Identity, so its state_dict (and a checkpoint of
    it) loads with strict=True.  A parameter container: `forward` runs the encoder in libspdm (convs as tcgen05 GEMMs over patch rows).
    (N, 3, 96, 96) -> (N, 512); inference only."""
    feat_dim = 512

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.GroupNorm(64 // 16, 64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = nn.Sequential(_GNBasicBlock(64, 64), _GNBasicBlock(64, 64))
        self.layer2 = nn.Sequential(_GNBasicBlock(64, 128, 2), _GNBasicBlock(128, 128))
        self.layer3 = nn.Sequential(_GNBasicBlock(128, 256, 2), _GNBasicBlock(256, 256))
        self.layer4 = nn.Sequential(_GNBasicBlock(256, 512, 2), _GNBasicBlock(512, 512))
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Identity()
        for m in self.modules():   # torchvision's initialisation (resnet.py): kaiming_normal_ for the convs, GroupNorm 1 / 0
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        self._owner = None

    def forward(self, img):
        if self._owner is None:
            raise RuntimeError("ResNet18GN is driven through its Diffusion_DDPM owner")
        return self._owner()._plan().encode_images(img)
