"""Python handle on a libspdm plan (include/spdm.h) — device memory, streams and weights come from torch,
every computation is a C-ABI call into the hand-written CUDA kernels.

A `DenoisePlan` stands in for the compute of
  * `UNet_Film.forward` / `UNet_Film_noAttention.forward`   (reference models/Unet_FiLmLayer.py:277-312)
  * `Autoencoder.encoder` + `prepare_obs_cond_vectors`       (models/encoder/autoencoder.py:11-20,
                                                              models/diffusion_ddpm.py:317-330)
  * the K-step loop of `Diffusion_DDPM.sample` / `Diffusion_DDIM.sample`
                                                             (models/diffusion_ddpm.py:268-276, diffusion_ddim.py:67-73)
for one (variant, precision, rows x dim, batch_max) configuration on one GPU.
"""
import ctypes

import torch

from . import _lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def _grad_phase(name):
    """When, inside the backward pass, the gradient of parameter `name` is final (spdm_train_wait_phase): 0 = up path
    (outc, sa4-6, up1-3), 1 = bottleneck + down path, 2 = time-embedding / FiLM Linears and the vision encoder (they need the
    gradients of every stage)."""
    if name.startswith("vision_encoder.") or ".emb_layer." in name or ".cond_encoder." in name:
        return 2
    return 0 if name.split(".")[0] in ("outc", "sa4", "sa5", "sa6", "up1", "up2", "up3") else 1


class DenoisePlan:
    def __init__(self, attention=True, precision="bf16", batch_max=1, rows=31, dim=5, obs_horizon=10, cond_dim=135,
                 inpaint_rows=1, time_dim=256, device=None, graph_steps=1, scheduler_only=False, split=1, simple=False,
                 encoder="autoencoder"):
        """`attention`: UNet_Film (True) or UNet_Film_noAttention (False); `simple=True`: the legacy `UNet` of
        models/simple_Unet.py (the reference's model='UNet' default) -- fp32 path, inference only.  `encoder`: "autoencoder"
        (models/encoder/autoencoder.py, 128 features per frame, cond_dim 135) or "resnet18" (the ResNet18-GroupNorm `VisionEncoder()`
        of models/Unet_FiLmLayer.py:316-386, 512 features, cond_dim 519; inference only)."""
        if encoder not in ("autoencoder", "resnet18"):
            raise ValueError("encoder must be 'autoencoder' or 'resnet18'")
        self.encoder = encoder
        self.feat_dim = 512 if encoder == "resnet18" else 128
        if not torch.cuda.is_available():
            raise RuntimeError("spdm: no CUDA device — the B200 denoising path has no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if precision not in ("fp32", "bf16", "tf32"):
            raise ValueError("precision must be 'fp32', 'bf16' or 'tf32'")
        cfg = _lib.SpdmConfig(
            variant=_lib.VARIANT_SIMPLE_UNET if simple else (_lib.VARIANT_ATTENTION if attention else _lib.VARIANT_NO_ATTENTION),
            precision={"bf16": _lib.PRECISION_BF16, "tf32": _lib.PRECISION_TF32, "fp32": _lib.PRECISION_FP32}[precision],
            batch_max=int(batch_max), rows=int(rows), dim=int(dim), obs_horizon=int(obs_horizon),
            cond_dim=int(cond_dim or 0), inpaint_rows=int(inpaint_rows), time_dim=int(time_dim),
            device=self.device.index or 0, graph_steps=int(graph_steps),
            flags=(_lib.FLAG_SCHEDULER_ONLY if scheduler_only else 0) | ((int(split) & 0xF) << 8)
            | (_lib.FLAG_ENCODER_RESNET18 if encoder == "resnet18" else 0))
        self.cfg = cfg
        self.attention, self.precision, self.simple = attention and not simple, precision, bool(simple)
        self.batch_max, self.rows, self.dim = int(batch_max), int(rows), int(dim)
        self.obs_horizon, self.cond_dim, self.inpaint_rows = int(obs_horizon), int(cond_dim or 0), int(inpaint_rows)
        self.K = 0
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.spdm_plan_create(ctypes.byref(handle), ctypes.byref(cfg)))
        self._h = handle
        self._keep = []  # tensors that must outlive an enqueued call

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None):
            self.lib.spdm_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def load_weight(self, name, tensor):
        t = _f32c(tensor, self.device)
        shape = (ctypes.c_int64 * max(t.dim(), 1))(*(t.shape if t.dim() else (1,)))
        with torch.cuda.device(self.device):
            _lib.check(self.lib.spdm_plan_load_weight(self._h, name.encode(), _ptr(t), shape, max(t.dim(), 1), _stream()))
            torch.cuda.current_stream().synchronize()  # `t` may be a temporary

    def load_unet_state_dict(self, sd, prefix=""):
        """nn.Module.load_state_dict for the noise estimator (SURVEY A.2 key names)."""
        for k, v in sd.items():
            if prefix and not k.startswith(prefix):
                continue
            self.load_weight(k[len(prefix):], v)
        if self.simple:   # its positional encoding is a buffer of the state_dict (pos_encoding.pos_encoding), loaded above
            return
        # exact torch value of the sinusoidal frequencies (models/Unet_FiLmLayer.py:267-270)
        td = self.cfg.time_dim
        inv_freq = 1.0 / (10000 ** (torch.arange(0, td, 2) / td))
        self.load_weight("pos_encoding.inv_freq", inv_freq)

    def load_encoder_state_dict(self, esd, prefix=""):
        """Autoencoder.encoder weights: keys `{prefix}{0,2,4,7}.{weight,bias}`; ResNet18-GroupNorm: every floating-point entry of the
        torchvision state_dict (`conv1.weight`, `bn1.weight`, `layer1.0.conv1.weight`, ...)."""
        if self.encoder == "resnet18":
            for k, v in esd.items():
                if k.startswith(prefix) and torch.is_floating_point(v):
                    self.load_weight("vision_encoder." + k[len(prefix):], v)
            return
        for k in ("0.weight", "0.bias", "2.weight", "2.bias", "4.weight", "4.bias", "7.weight", "7.bias"):
            self.load_weight("vision_encoder." + k, esd[prefix + k])

    def missing_weights(self):
        n = self.lib.spdm_plan_missing_weights(self._h)
        names = self.lib.spdm_last_error().decode().split()
        return names if n else []

    # ------------------------------------------------------------------ schedule
    def set_schedule(self, kind, coef, timesteps):
        """coef: (K, 8) fp32 host rows {c0, c1, k_x0, k_x, k_eps, k_noise, 0, 0}; timesteps: (K,) int64."""
        coef = coef.detach().to("cpu", torch.float32).contiguous()
        ts = timesteps.detach().to("cpu", torch.int64).contiguous()
        assert coef.shape == (ts.numel(), 8)
        k = _lib.SCHED_DDPM if kind == "ddpm" else _lib.SCHED_DDIM
        with torch.cuda.device(self.device):
            _lib.check(self.lib.spdm_plan_set_schedule(self._h, k, ts.numel(), _ptr(coef), _ptr(ts), _stream()))
        self.K = ts.numel()

    # ------------------------------------------------------------------ conditioning
    def encode_images(self, images):
        img = _f32c(images, self.device).reshape(-1, 3, 96, 96)
        out = torch.empty((img.shape[0], self.feat_dim), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.spdm_encode_images(self._h, _ptr(img), _ptr(out), img.shape[0], _stream()))
        return out

    def encode_cond(self, image, position, action, velocity, return_cond=True):
        """Vision encoder + obs_cond + the six FiLM Linears, cached in the plan.  `image`: float (B,T,3,96,96) in [0,1] as the
        reference's batches carry it, or uint8 (B,T,96,96,3) HWC as the simulator stores it (decoded x / 255 on the device)."""
        B = image.shape[0]
        pos, act, vel = (_f32c(t, self.device) for t in (position, action, velocity))
        with torch.cuda.device(self.device):
            if image.dtype == torch.uint8:
                if image.dim() != 5 or tuple(image.shape[2:]) != (96, 96, 3):
                    raise ValueError("uint8 frames must be (B, T, 96, 96, 3) HWC")
                img = image.detach().to(self.device, non_blocking=True).contiguous()
                _lib.check(self.lib.spdm_encode_cond_u8(self._h, _ptr(img), _ptr(pos), _ptr(act), _ptr(vel), B, _stream()))
            else:
                img = _f32c(image, self.device)
                _lib.check(self.lib.spdm_encode_cond(self._h, _ptr(img), _ptr(pos), _ptr(act), _ptr(vel), B, _stream()))
        self._keep_cond = [img, pos, act, vel]
        return self.get_cond(B) if return_cond else None

    def set_cond(self, obs_cond):
        c = _f32c(obs_cond, self.device).reshape(obs_cond.shape[0], -1)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.spdm_set_cond(self._h, _ptr(c), c.shape[0], _stream()))
        self._keep = [c]

    def get_cond(self, B):
        out = torch.empty((B, self.obs_horizon * self.cond_dim), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.spdm_get_cond(self._h, _ptr(out), B, _stream()))
        return out

    # ------------------------------------------------------------------ U-Net forward
    def _prep_fwd(self, x, t, y):
        x = _f32c(x, self.device)
        B = x.shape[0]
        t = t.detach().to(self.device, torch.int64).reshape(-1).contiguous()
        yy = None if y is None else _f32c(y, self.device).reshape(B, -1)
        out = torch.empty((B, 1, self.rows, self.dim), device=self.device, dtype=torch.float32)
        return x, t, yy, out, B

    def unet_forward(self, x, t, y=None, use_cached_cond=False):
        x, t, yy, out, B = self._prep_fwd(x, t, y)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.spdm_unet_forward(self._h, _ptr(x), _ptr(t), t.numel(), _ptr(yy), int(use_cached_cond),
                                                  _ptr(out), B, _stream()))
        self._keep = [x, t, yy]
        return out

    def debug_forward(self, x, t, y, tap, tap_shape):
        """Runs the forward and returns (out, activation `tap` as (B, C, H, W) fp32)."""
        x, t, yy, out, B = self._prep_fwd(x, t, y)
        tap_out = torch.zeros(tap_shape, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            n = self.lib.spdm_debug_forward(self._h, _ptr(x), _ptr(t), t.numel(), _ptr(yy), 0, _ptr(out), B, tap.encode(),
                                            _ptr(tap_out), _stream())
            _lib.check(n)
            torch.cuda.synchronize()
        assert n == tap_out.numel(), (n, tap_out.shape)
        return out, tap_out

    # ------------------------------------------------------------------ scheduler step / sampling
    def step(self, x, eps, step_index, noise=None, inpaint=None):
        x, eps = _f32c(x, self.device), _f32c(eps, self.device)
        nz = None if noise is None else _f32c(noise, self.device)
        ip = None if inpaint is None else _f32c(inpaint, self.device)
        out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.spdm_step(self._h, _ptr(x), _ptr(eps), _ptr(nz), _ptr(ip), _ptr(out), int(step_index), x.shape[0],
                                          _stream()))
        self._keep = [x, eps, nz, ip]
        return out

    def sample(self, x_T, noise=None, inpaint=None, history=False, seed=0):
        """Runs all K schedule steps on the cached conditioning.  Returns x_0 (B,1,rows,dim) [, history (K+1,B,1,rows,dim)]."""
        x_T = _f32c(x_T, self.device)
        B = x_T.shape[0]
        nz = None if noise is None else _f32c(noise, self.device)
        ip = None if inpaint is None else _f32c(inpaint, self.device).reshape(B, -1)
        out = torch.empty_like(x_T)
        hist = torch.empty((self.K + 1,) + tuple(x_T.shape), device=self.device, dtype=torch.float32) if history else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.spdm_sample(self._h, _ptr(x_T), _ptr(nz), _ptr(ip), _ptr(out), _ptr(hist), int(seed) & (2**64 - 1),
                                            B, _stream()))
        self._keep = [x_T, nz, ip, hist]
        return (out, hist) if history else out

    def add_noise(self, x0, noise, t, sqrt_ab, sqrt_1mab, inpaint=None):
        x0, noise = _f32c(x0, self.device), _f32c(noise, self.device)
        t = t.detach().to(self.device, torch.int64).contiguous()
        sa, sb = _f32c(sqrt_ab, self.device), _f32c(sqrt_1mab, self.device)
        ip = None if inpaint is None else _f32c(inpaint, self.device)
        out = torch.empty_like(x0)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.spdm_add_noise(self._h, _ptr(x0), _ptr(noise), _ptr(t), _ptr(sa), _ptr(sb), _ptr(ip), _ptr(out),
                                               x0.shape[0], _stream()))
        self._keep = [x0, noise, t, sa, sb, ip]
        return out

    # ------------------------------------------------------------------ training step
    def enable_training(self, named_tensors):
        """Turns the plan into a training plan (reference: `process_single_batch` + `loss.backward()` + Adam,
        models/diffusion_ddpm.py:115-173).  `named_tensors`: ordered {name: tensor} of every trainable tensor — U-Net
        state_dict names, encoder tensors as `vision_encoder.<k>`.  Allocates the flat fp32 parameter / gradient /
        Adam-moment buffers (one NCCL all-reduce and one optimizer kernel cover all of them), uploads the weights and
        returns {name: (offset, shape)}."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.spdm_train_enable(self._h))
            # flat layout in gradient-completion order, so that each phase is one contiguous slice (bucket) of the buffer
            self.train_offsets, off = {}, 0
            self.train_buckets = []
            for phase in (0, 1, 2):
                lo = off
                for name, t in named_tensors.items():
                    if _grad_phase(name) == phase:
                        self.train_offsets[name] = (off, tuple(t.shape))
                        off += (t.numel() + 3) // 4 * 4
                self.train_buckets.append((lo, off))
            self.train_total = off
            self.params_flat = torch.zeros(off, device=self.device, dtype=torch.float32)
            self.grads_flat = torch.zeros_like(self.params_flat)
            self.adam_m = torch.zeros_like(self.params_flat)
            self.adam_v = torch.zeros_like(self.params_flat)
            self._adam_scratch = torch.zeros(1, device=self.device, dtype=torch.float32)
            self.adam_steps = 0
            for name, t in named_tensors.items():
                o, shp = self.train_offsets[name]
                self.params_flat[o:o + t.numel()].copy_(t.detach().reshape(-1).to(self.device, torch.float32))
                shape = (ctypes.c_int64 * max(len(shp), 1))(*(shp if len(shp) else (1,)))
                _lib.check(self.lib.spdm_train_bind(self._h, name.encode(), o, shape, max(len(shp), 1)))
            _lib.check(self.lib.spdm_train_set_buffers(self._h, _ptr(self.params_flat), _ptr(self.grads_flat), off))
            td = self.cfg.time_dim
            self.load_weight("pos_encoding.inv_freq", 1.0 / (10000 ** (torch.arange(0, td, 2) / td)))
            self.sync_weights()
        return self.train_offsets

    def param_view(self, name, flat=None):
        o, shp = self.train_offsets[name]
        n = 1
        for d in shp:
            n *= d
        return (self.params_flat if flat is None else flat)[o:o + n].view(shp)

    def grad_view(self, name):
        return self.param_view(name, self.grads_flat)

    def allreduce_gradients(self, group=None, comm_stream=None):
        """Data-parallel step: summing all-reduce of the flat gradient buffer in its three completion-phase buckets, each
        enqueued on `comm_stream` behind the event that marks its gradients final, so the first two run under the rest of the
        backward pass.  Returns 1 / world_size (fold it into `adam_step(grad_scale=...)`)."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return 1.0
        if group is None:
            group = self._gradient_group()
        if comm_stream is None:
            if getattr(self, "_comm_stream", None) is None:
                import os
                # high priority (SPDM_COMM_PRIORITY=0 switches it off): the block scheduler then places the all-reduce's CTAs ahead of
                # the next persistent compute launch instead of behind the whole queue of them
                prio = -1 if int(os.environ.get("SPDM_COMM_PRIORITY", "1")) else 0
                self._comm_stream = torch.cuda.Stream(device=self.device, priority=prio)
            comm_stream = self._comm_stream
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device), torch.cuda.stream(comm_stream):
            for phase, (lo, hi) in enumerate(self.train_buckets):
                if hi > lo:
                    _lib.check(self.lib.spdm_train_wait_phase(self._h, phase, ctypes.c_void_p(comm_stream.cuda_stream)))
                    dist.all_reduce(self.grads_flat[lo:hi], op=dist.ReduceOp.SUM, group=group)
        cur.wait_stream(comm_stream)
        return 1.0 / dist.get_world_size(group)

    def _gradient_group(self):
        """Process group of the gradient all-reduces.  `SPDM_GRAD_COMM_CTAS=n` (> 0) puts them on a communicator of their own whose
        collectives are capped at n CTAs (an experiment: fewer SMs taken from the persistent one-CTA-per-SM compute launches the
        all-reduces overlap with).  Measured on 8 GPUs: the step gets SLOWER (7.88 ms default, 8.12 ms at 4 CTAs, 9.05 ms at 2), i.e.
        the all-reduce time is on the critical path rather than hidden -- default 0 = the default group."""
        if getattr(self, "_grad_group", False) is not False:
            return self._grad_group
        import os
        import torch.distributed as dist
        ctas = int(os.environ.get("SPDM_GRAD_COMM_CTAS", "0"))
        self._grad_group = None
        if ctas > 0 and dist.get_backend() == "nccl":
            try:
                opts = dist.ProcessGroupNCCL.Options()
                opts.config.max_ctas = ctas
                opts.config.min_ctas = 1
                self._grad_group = dist.new_group(backend="nccl", pg_options=opts)
            except Exception:   # an older torch / NCCL without communicator configs: the default group
                self._grad_group = None
        return self._grad_group

    def sync_weights(self):
        """Repack the flat fp32 parameters into the kernel layouts (forward operands and data-gradient twins)."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.spdm_train_sync_weights(self._h, _stream()))

    @property
    def batch_multiple(self):
        """Granularity of the batch on the tensor-core path (whole samples per 128-row tile at the deepest level); 1 on fp32."""
        return int(self.lib.spdm_plan_batch_multiple(self._h))

    def train_fwd_bwd(self, image, position, action, velocity, x0, noise, t, sqrt_ab, sqrt_1mab, inpaint=None):
        """q-sample + inpaint + encoder + U-Net forward, MSE loss, full backward.  Gradients (fp32, PyTorch layout) are left
        in `grads_flat`; returns the loss as a 1-element device tensor.

        A batch that is not a multiple of `batch_multiple` (the last batch of an epoch: the reference's DataLoader has no
        drop_last, utils/load_data.py:174) is padded with copies of its first sample and the real count is declared to the
        library (spdm_train_set_valid): loss and gradients are those of the real samples only."""
        B = x0.shape[0]
        bm = self.batch_multiple
        valid = 0
        if B % bm:
            Bp = (B + bm - 1) // bm * bm
            if Bp > self.batch_max:
                raise _lib.SpdmError("ragged batch of %d needs %d slots (batch_multiple %d) but the plan holds %d" % (B, Bp, bm, self.batch_max))
            idx = torch.cat([torch.arange(B), torch.zeros(Bp - B, dtype=torch.long)])

            def pad(v):
                return None if v is None else v.to(self.device).index_select(0, idx.to(self.device))
            image, position, action, velocity, x0, noise, t, inpaint = (pad(v) for v in (image, position, action, velocity, x0, noise, t, inpaint))
            valid, B = B, Bp
        if valid != getattr(self, "_valid", 0):
            _lib.check(self.lib.spdm_train_set_valid(self._h, valid))
            self._valid = valid
        pos, act, vel = (_f32c(v, self.device) for v in (position, action, velocity))
        # the observation window is usually a slice [:, :obs_horizon] of the full recording: pass it as a strided view
        # (bf16 path) instead of copying 566 MB per step
        stride = 0
        if (self.precision == "bf16" and image.is_cuda and image.dtype == torch.float32 and image.dim() == 5
                and not image.is_contiguous() and image[0].is_contiguous()):
            img, stride = image.detach(), int(image.stride(0))
        else:
            img = _f32c(image, self.device)
        if stride != getattr(self, "_img_stride", 0):
            _lib.check(self.lib.spdm_train_set_image_stride(self._h, stride))
            self._img_stride = stride
        x0, noise = _f32c(x0, self.device), _f32c(noise, self.device)
        t = t.detach().to(self.device, torch.int64).contiguous()
        sa, sb = _f32c(sqrt_ab, self.device), _f32c(sqrt_1mab, self.device)
        ip = None if inpaint is None else _f32c(inpaint, self.device)
        loss = torch.empty(1, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.spdm_train_fwd_bwd(self._h, _ptr(img), _ptr(pos), _ptr(act), _ptr(vel), _ptr(x0), _ptr(noise), _ptr(t),
                                                   _ptr(sa), _ptr(sb), _ptr(ip), _ptr(loss), B, _stream()))
        self._keep = [img, pos, act, vel, x0, noise, t, sa, sb, ip]
        return loss

    def adam_step(self, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, max_norm=0.5, grad_scale=1.0, sync=True):
        """clip_grad_norm_(max_norm) (train.py:107) + torch.optim.Adam.step() (ddpm:115-125) on the flat buffers, then
        re-upload the weights.  grad_scale = 1 / world_size after a summing all-reduce."""
        self.adam_steps += 1
        with torch.cuda.device(self.device):
            _lib.check(self.lib.spdm_adam_step(_ptr(self.params_flat), _ptr(self.grads_flat), _ptr(self.adam_m), _ptr(self.adam_v),
                                               self.train_total, float(lr), float(betas[0]), float(betas[1]), float(eps),
                                               int(self.adam_steps), float(max_norm or 0.0), float(grad_scale),
                                               _ptr(self._adam_scratch), _stream()))
            if sync:
                self.sync_weights()

    def optimizer_state_dict(self):
        """The fused optimizer's state (Adam moments per parameter name + step count), for checkpoints and plan rebuilds."""
        return {"step": int(self.adam_steps),
                "exp_avg": {k: self.param_view(k, self.adam_m).detach().clone() for k in self.train_offsets},
                "exp_avg_sq": {k: self.param_view(k, self.adam_v).detach().clone() for k in self.train_offsets}}

    def load_optimizer_state_dict(self, state):
        self.adam_steps = int(state["step"])
        for k in self.train_offsets:
            if k in state["exp_avg"]:
                self.param_view(k, self.adam_m).copy_(state["exp_avg"][k].to(self.device))
                self.param_view(k, self.adam_v).copy_(state["exp_avg_sq"][k].to(self.device))

    def profile_step(self, B, reps=3):
        """Eager, CUDA-event-timed denoising step: {class: dict(ms, launches, flops, bytes)} averaged over `reps`."""
        n = len(_lib.PROFILE_CLASSES)
        buf = (ctypes.c_double * (n * 4))()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.spdm_profile_step(self._h, int(B), int(reps), buf, _stream()))
        return {name: dict(ms=buf[i * 4], launches=buf[i * 4 + 1], flops=buf[i * 4 + 2], bytes=buf[i * 4 + 3])
                for i, name in enumerate(_lib.PROFILE_CLASSES)}

    # ------------------------------------------------------------------ introspection
    @property
    def launch_count(self):
        return int(self.lib.spdm_plan_launch_count(self._h))

    @property
    def workspace_bytes(self):
        return int(self.lib.spdm_plan_workspace_bytes(self._h))
