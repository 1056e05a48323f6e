"""`Diffusion_DDPM` / `Diffusion_DDIM` with the reference's constructor, attributes and method surface
(models/diffusion_ddpm.py:22-348, models/diffusion_ddim.py:19-74), running on libspdm.

What stays exactly as in the reference (default behaviour of `sample(batch, option=None)`):
  * `sample`/`validate` use batch element 0 only and return (1,1,pred_h+ih,pred_dim)      (ddpm:196,246)
  * x_T ~ U[0,1) from `torch.rand` on the model device                                     (ddpm:205,252)
  * K = `self.noise_steps` steps of `self.noise_scheduler`, inpainting after every step    (ddpm:268-276)
  * `option='sample_history'` returns a Python list of K+1 tensors                         (ddpm:254-265)
What is added behind the same method (keyword-only, default off):
  * `batched=True` samples every row of the batch; `x_T=` / `noise=` inject the random draws;
    `mode='validation'` returns `validate(batch)`'s 3-tuple (the call shape the reference's
    evaluation/*.py scripts use).
The K-step loop is one C-ABI call (`spdm_sample`): conditioning encoder + FiLM GEMM once, then a
CUDA-graphed U-Net + posterior update + inpaint per step; no per-step host work.
"""
from datetime import datetime

import torch
import torch.nn as nn

from .schedulers import DDIMScheduler, DDPMScheduler
from .unet import ResNet18GN, UNet, UNet_Film, UNet_Film_noAttention

try:  # Lightning is optional: the reference subclasses pl.LightningModule, which is absent on the B200 image
    import pytorch_lightning as pl
    _Base = pl.LightningModule
except Exception:  # pragma: no cover - depends on the environment
    pl = None

    class _Hparams(dict):
        __getattr__ = dict.get

    class _Base(nn.Module):
        """Minimal stand-in for pl.LightningModule: hparams, .device, .log, load_from_checkpoint."""

        def __init__(self):
            super().__init__()
            self.hparams = _Hparams()

        def save_hyperparameters(self, **kw):
            self.hparams.update(kw)

        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        def log(self, *a, **k):
            pass

        @classmethod
        def load_from_checkpoint(cls, checkpoint_path, hparams_file=None, map_location=None, **kwargs):
            ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
            hp = dict(ckpt.get("hyper_parameters", {}))
            if hparams_file is not None:
                import yaml
                with open(hparams_file) as f:
                    hp.update(yaml.safe_load(f) or {})
            hp.update(kwargs)
            model = cls(**hp)
            model.load_state_dict(ckpt["state_dict"], strict=True)
            return model


class _NativeLoss(torch.autograd.Function):
    """The loss returned by a native training step.  spdm_train_fwd_bwd has already left d loss / d parameter in the plan's
    private flat gradient buffer; `loss.backward()` -- which Lightning's automatic optimization and plain user loops call
    AFTER `optimizer.zero_grad()` (training_step -> zero_grad -> backward -> step) -- publishes them into the parameters'
    `.grad` with autograd's semantics: `.grad = g * dL/dp` where it is None (zero_grad(set_to_none=True)), `.grad += g * dL/dp`
    otherwise (zero_grad(set_to_none=False), gradient accumulation)."""

    @staticmethod
    def forward(ctx, loss, anchor, owner):
        ctx.owner = owner
        return loss.clone()

    @staticmethod
    def backward(ctx, g):
        ctx.owner()._publish_gradients(g)
        return None, None, None


class VisionEncoder(nn.Sequential):
    """Parameters of Autoencoder.encoder (models/encoder/autoencoder.py:11-20); state_dict keys 0,2,4,7."""

    def __init__(self, channels=3, latent_dim=128):
        super().__init__(nn.Conv2d(channels, 16, 2, stride=2, padding=1), nn.ReLU(), nn.Conv2d(16, 32, 2, stride=2, padding=0),
                         nn.ReLU(), nn.Conv2d(32, 64, 2, stride=2, padding=0), nn.ReLU(), nn.Flatten(),
                         nn.Linear(64 * 12 * 12, latent_dim))
        self._owner = None

    def forward(self, img):
        """(N,3,96,96) -> (N,128) on the fused encoder kernels."""
        if self._owner is None:
            raise RuntimeError("VisionEncoder is driven through its Diffusion_DDPM owner")
        return self._owner()._plan().encode_images(img)


class Diffusion_DDPM(_Base):
    def __init__(self, noise_steps=1000, obs_horizon=10, pred_horizon=10, observation_dim=2, prediction_dim=2,
                 learning_rate=1e-4, model='UNet', vision_encoder=None, noise_scheduler_type='linear', inpaint_horizon=10,
                 step_size=1, noise_scheduler=None, autoencoder_checkpoint="./tb_logs_autoencoder/version_23/checkpoints/epoch=25.ckpt"):
        super().__init__()
        if noise_scheduler is not None:  # train.py:88 passes `noise_scheduler=` (a TypeError in the reference)
            noise_scheduler_type = noise_scheduler
        self.save_hyperparameters(noise_steps=noise_steps, obs_horizon=obs_horizon, pred_horizon=pred_horizon,
                                  observation_dim=observation_dim, prediction_dim=prediction_dim, learning_rate=learning_rate,
                                  model=model, vision_encoder=vision_encoder, noise_scheduler_type=noise_scheduler_type,
                                  inpaint_horizon=inpaint_horizon, step_size=step_size)
        self.date = datetime.today().strftime('%Y_%m_%d_%H-%M-%S')
        self.noise_steps = noise_steps
        self.NoiseScheduler = None
        self.obs_horizon = obs_horizon
        self.pred_horizon = pred_horizon
        self.observation_dim = observation_dim
        self.prediction_dim = prediction_dim
        self.inpaint_horizon = inpaint_horizon
        if model == 'UNet_Film':
            self.model = UNet_Film
        elif model == 'UNet_FilmnoAttention':
            self.model = UNet_Film_noAttention
        else:  # ddpm:60-62: anything else is the legacy simple U-Net (fp32 path, inference only here)
            self.model = UNet
        self.noise_scheduler = DDPMScheduler(num_train_timesteps=self.noise_steps, beta_schedule='linear', clip_sample=False,
                                             prediction_type='epsilon')
        self.lr = learning_rate
        self.loss = nn.MSELoss()
        self.noise_estimator = self.model(in_channels=1, out_channels=1, noise_steps=noise_steps,
                                          global_cond_dim=observation_dim * obs_horizon, time_dim=256)
        import os
        import weakref
        if vision_encoder in ('resnet18', 'ResNet18', 'resnet'):
            # `VisionEncoder()` of models/Unet_FiLmLayer.py:383-386 (ResNet18, GroupNorm instead of BatchNorm): 512 features per frame;
            # the reference's comment at ddpm:79 ("512 is the output dim of Resnet18") is this wiring
            if model not in ('UNet_Film', 'UNet_FilmnoAttention'):
                raise ValueError("vision_encoder='resnet18' is wired to the FiLM U-Nets")
            if observation_dim != 7 + ResNet18GN.feat_dim:
                raise ValueError("vision_encoder='resnet18' produces 512 image features: observation_dim must be 2 + 3 + 2 + 512 = 519, got %d"
                                 % observation_dim)
            self.vision_encoder = ResNet18GN()
            autoencoder_checkpoint = None
        else:
            self.vision_encoder = VisionEncoder()
        self.vision_encoder._owner = weakref.ref(self)
        self.noise_estimator.encoder = "resnet18" if isinstance(self.vision_encoder, ResNet18GN) else "autoencoder"
        if autoencoder_checkpoint and os.path.exists(autoencoder_checkpoint):  # ddpm:84-88
            sd = torch.load(autoencoder_checkpoint, map_location="cpu", weights_only=False)["state_dict"]
            enc = {k.split("encoder.", 1)[1]: v for k, v in sd.items() if "encoder." in k and "decoder" not in k}
            self.vision_encoder.load_state_dict(enc, strict=True)
        self.vision_encoder.eval()
        # B200 execution options
        self.precision = "fp32" if self.model is UNet else "bf16"
        self.graph_steps = 10   # denoising steps per captured CUDA graph (the K % graph_steps remainder runs as 1-step graphs)
        self.batch_max = 0
        self._enc_tag = None

    # ------------------------------------------------------------------------------------------
    # engine plumbing
    # ------------------------------------------------------------------------------------------
    def configure(self, precision=None, graph_steps=None, batch_max=None, split=None):
        if precision is not None:
            if self.model is UNet and precision != "fp32":
                raise ValueError("the simple U-Net (model='UNet') runs on the fp32 path only")
            self.precision = precision
        if graph_steps is not None:
            self.graph_steps = int(graph_steps)
        if batch_max is not None:
            self.batch_max = int(batch_max)
        if split is not None:
            self.noise_estimator.split = int(split)
        return self

    def _plan(self, B=1):
        ne = self.noise_estimator
        ne.precision = self.precision
        ne.batch_max = max(ne.batch_max, self.batch_max)
        cond_dim = self.observation_dim
        plan = ne.plan_for(B, self.pred_horizon + self.inpaint_horizon, self.prediction_dim, self.obs_horizon, cond_dim,
                           inpaint_rows=self.inpaint_horizon, graph_steps=self.graph_steps)
        tag = (id(plan),) + tuple((p.data_ptr(), p._version) for p in self.vision_encoder.parameters())
        if self._enc_tag != tag:
            plan.load_encoder_state_dict(self.vision_encoder.state_dict())
            self._enc_tag = tag
        return plan

    def _bind_schedule(self, plan):
        sch = self.noise_scheduler
        if not hasattr(sch, "coef_table"):
            raise TypeError("noise_scheduler must be a state_policy_diffusionmodel_b200.schedulers scheduler "
                            "(DDPMScheduler / DDIMScheduler)")
        key = (id(sch), id(plan), int(self.noise_steps), sch.num_train_timesteps)
        if getattr(self, "_sched_key", None) == key and plan.K == int(self.noise_steps):
            return  # same scheduler object, same step count, same plan: the device tables are current
        sch.set_timesteps(self.noise_steps)
        plan.set_schedule(sch.kind, sch.coef_table(), sch.timesteps)
        self._sched_key = key

    # ------------------------------------------------------------------------------------------
    # training / validation hooks (ddpm:92-125)
    # ------------------------------------------------------------------------------------------
    def _current_lr(self):
        """ddpm:95 logs `self.optimizers().param_groups[0]['lr']` (train.py's EarlyStopping monitors it); outside a Lightning
        trainer there is no attached optimizer and the configured rate is logged instead."""
        try:
            opt = self.optimizers()
            if isinstance(opt, (list, tuple)):
                opt = opt[0]
            return opt.param_groups[0]['lr']
        except Exception:
            return self.lr

    def training_step(self, batch, batch_idx):
        loss = self.process_single_batch(batch)
        self.log("train_loss", loss)
        self.log('lr', self._current_lr())
        return loss

    def validation_step(self, batch, batch_idx):
        if batch_idx == 0:  # ddpm:100-109: one sampled trajectory per validation run, handed to the plotting hook
            x_0_predicted, observation_batch, inpaint_vector = self.validate(batch)
            self.plt2tensorboard(batch=batch, prediction=x_0_predicted, inpaint_vector=inpaint_vector,
                                 observation_batch=observation_batch)
        loss = self.process_single_batch(batch)
        self.log("val_loss", loss, sync_dist=True)
        return loss

    def plt2tensorboard(self, batch, prediction, inpaint_vector, observation_batch):
        """ddpm:350-367 draws the sampled trajectory into the tensorboard logger.  Plotting is outside the hot path: the
        sampled trajectory is kept in `last_validation` (device tensors) and, when a Lightning logger and matplotlib are
        present, the predicted and ground-truth positions are added as a figure."""
        self.last_validation = {"prediction": prediction, "inpaint_vector": inpaint_vector, "observation_batch": observation_batch}
        logger = getattr(self, "logger", None)
        exp = getattr(logger, "experiment", None) if logger is not None else None
        if exp is None or not hasattr(exp, "add_figure"):
            return
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
        except Exception:
            return
        pred = prediction[0, 0, :, :2].detach().float().cpu().numpy()
        truth = batch['position'][0, self.obs_horizon:, :].detach().float().cpu().numpy()
        obs = observation_batch['position'][0].detach().float().cpu().numpy()
        fig = plt.figure()
        plt.plot(obs[:, 0], obs[:, 1], 'o', label='observation')
        plt.plot(truth[:, 0], truth[:, 1], 'o', label='ground truth')
        plt.plot(pred[:, 0], pred[:, 1], 'x', label='prediction')
        plt.legend()
        exp.add_figure("plot", fig, global_step=int(getattr(self, "global_step", 0)))
        plt.close(fig)

    def configure_optimizers(self):
        optimizer = torch.optim.Adam(self.parameters(), lr=self.lr)
        scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, 'min', patience=5)
        return {"optimizer": optimizer, "lr_scheduler": {"scheduler": scheduler, "monitor": "val_loss", "frequency": 1}}

    # ------------------------------------------------------------------------------------------
    # native training step: forward + backward + optimizer in libspdm (ddpm:115-173, train.py:104-107)
    # ------------------------------------------------------------------------------------------
    def named_trainable(self):
        """Every tensor the reference hands to Adam (`self.parameters()`: U-Net and vision encoder), under libspdm names."""
        named = {k: p for k, p in self.noise_estimator.named_parameters()}
        named.update({"vision_encoder." + k: p for k, p in self.vision_encoder.named_parameters()})
        return named

    def _training_plan(self, B):
        from .engine import DenoisePlan
        if self.model is UNet:
            raise NotImplementedError("spdm: the native training step covers model='UNet_Film' / 'UNet_FilmnoAttention'; the legacy simple "
                                      "U-Net (model='UNet') is inference-only on the B200 path")
        if isinstance(self.vision_encoder, ResNet18GN):
            raise NotImplementedError("spdm: the native training step covers the autoencoder vision encoder; vision_encoder='resnet18' is "
                                      "inference-only on the B200 path")
        plan = getattr(self, "_tplan", None)
        key = (self.precision, self.pred_horizon + self.inpaint_horizon, self.prediction_dim, self.obs_horizon, self.observation_dim,
               self.inpaint_horizon, str(self.device))
        if plan is None or self._tplan_key != key or plan.batch_max < B:
            named = self.named_trainable()
            carry = None
            if plan is not None:  # keep the current values (and the optimizer moments): they live in the old plan's flat buffers
                named = {k: p.detach().clone() for k, p in named.items()}
                carry = plan.optimizer_state_dict()
                plan.close()
            ne = self.noise_estimator
            cap = max(int(B), self.batch_max)
            if self.precision == "bf16":  # room to pad a ragged (last-of-epoch) batch to the tile granularity (engine.train_fwd_bwd)
                up8 = lambda v: (v + 7) // 8 * 8
                hw3 = (up8(self.pred_horizon + self.inpaint_horizon) // 8) * (up8(self.prediction_dim) // 8)
                bm = max(1, 128 // hw3)
                cap = (cap + bm - 1) // bm * bm
            plan = DenoisePlan(attention=ne._attention, precision=self.precision, batch_max=cap,
                               rows=self.pred_horizon + self.inpaint_horizon, dim=self.prediction_dim, obs_horizon=self.obs_horizon,
                               cond_dim=self.observation_dim, inpaint_rows=self.inpaint_horizon, time_dim=ne.time_dim, device=self.device)
            plan.enable_training(named)
            if carry is not None:
                plan.load_optimizer_state_dict(carry)
            # nn.Parameters now alias the plan's flat fp32 parameter buffer: any torch optimizer works on them unchanged.  Their
            # `.grad` is filled by `loss.backward()` (_publish_gradients) from the plan's private gradient buffer.
            for k, p in self.named_trainable().items():
                p.data = plan.param_view(k)
                p.grad = None
            self._grads_pub = None
            self._tplan, self._tplan_key = plan, key
            self._tplan_versions = None
        return plan

    def _param_versions(self):
        return tuple(p._version for p in self.parameters())

    def _native_training_step(self, batch, t=None, noise=None):
        observation_batch = self.prepare_observation_batch(batch)
        prediction_batch = self.prepare_prediction_batch(batch)
        B = observation_batch['position'].shape[0]
        plan = self._training_plan(B)
        if self._tplan_versions != self._param_versions():  # a torch optimizer (or load_state_dict) changed the weights
            plan.sync_weights()
            self._weights_changed()
        x_0 = self.prepare_prediction_vectors(prediction_batch).unsqueeze(1)
        x_0_inpaint = self.prepare_inpaint_vectors(observation_batch).unsqueeze(1)
        if t is None:
            t = torch.randint(0, self.noise_steps, (B,), device=self.device).long()
        prediction_vector = torch.cat([x_0_inpaint, x_0], dim=2)
        if noise is None:
            noise = torch.randn_like(prediction_vector)
        ac = self.noise_scheduler.alphas_cumprod
        loss = plan.train_fwd_bwd(observation_batch['image'], observation_batch['position'], observation_batch['action'],
                                  observation_batch['velocity'], prediction_vector, noise, t, ac ** 0.5, (1 - ac) ** 0.5,
                                  inpaint=x_0_inpaint.reshape(B, -1) if self.inpaint_horizon > 0 else None)
        anchor = next(iter(self.noise_estimator.parameters()))
        self._tplan_versions = self._param_versions()
        import weakref
        return _NativeLoss.apply(loss.reshape(()), anchor, weakref.ref(self))

    def _publish_gradients(self, g):
        """`loss.backward()` of a native step: .grad (+)= g * (the plan's gradient buffer), for every trainable parameter.
        The published gradients live in one flat buffer of their own (`.grad` tensors are views of it), so whatever the
        optimizer / `zero_grad` / `clip_grad_norm_` do to `.grad` in place never touches the buffer the kernels write."""
        plan = self._tplan
        named = self.named_trainable()
        if getattr(self, "_grads_pub", None) is None or self._grads_pub.data_ptr() == 0 or self._grads_pub.numel() != plan.grads_flat.numel():
            self._grads_pub = torch.empty_like(plan.grads_flat)
        pub = self._grads_pub
        views = {k: plan.param_view(k, pub) for k in named}
        state = [None if p.grad is None else p.grad.data_ptr() == views[k].data_ptr() for k, p in named.items()]
        g = g.to(pub.dtype)
        if all(st is None for st in state):          # zero_grad(set_to_none=True) (torch / Lightning default), or the first step
            torch.mul(plan.grads_flat, g, out=pub)
            for k, p in named.items():
                p.grad = views[k]
        elif all(st is True for st in state):        # zero_grad(set_to_none=False) or gradient accumulation: accumulate, flat
            pub.addcmul_(plan.grads_flat, g)
        else:                                        # mixed / foreign .grad tensors: per parameter, autograd's rule
            for k, p in named.items():
                inc = plan.grad_view(k) * g
                if p.grad is None:
                    p.grad = inc
                else:
                    p.grad.add_(inc)

    def _weights_changed(self):
        self.noise_estimator._weights_epoch += 1
        self._enc_tag = None

    def allreduce_gradients(self, group=None):
        """Data-parallel training: NCCL all-reduce (sum) of the flat gradient buffer in three completion-phase buckets that
        overlap the backward pass.  Returns the factor (1 / world_size) that `optimizer_step` folds into the update."""
        return self._tplan.allreduce_gradients(group=group)

    def optimizer_step(self, lr=None, betas=(0.9, 0.999), eps=1e-8, gradient_clip_val=0.5, grad_scale=1.0):
        """Fused clip_grad_norm_(gradient_clip_val) + Adam over all parameters (one kernel pair on the flat buffers);
        equivalent to `clip_grad_norm_` + `configure_optimizers()['optimizer'].step()` of the reference."""
        plan = self._tplan
        plan.adam_step(lr=self.lr if lr is None else lr, betas=betas, eps=eps, max_norm=gradient_clip_val, grad_scale=grad_scale)
        self._weights_changed()
        self._tplan_versions = self._param_versions()

    def process_single_batch(self, batch, t=None, noise=None):
        """ddpm:128-173 — q-sample + inpaint + U-Net + MSE.  In training mode with autograd enabled the whole step
        (forward, loss, backward) runs in libspdm and the parameter gradients are left in `.grad`; otherwise forward only.
        `t` / `noise` may be injected for parity runs."""
        if self.training and torch.is_grad_enabled():
            return self._native_training_step(batch, t=t, noise=noise)
        observation_batch = self.prepare_observation_batch(batch)
        prediction_batch = self.prepare_prediction_batch(batch)
        B = observation_batch['position'].shape[0]
        plan = self._plan(B)
        plan.encode_cond(observation_batch['image'], observation_batch['position'], observation_batch['action'],
                         observation_batch['velocity'])
        x_0 = self.prepare_prediction_vectors(prediction_batch).unsqueeze(1)
        x_0_inpaint = self.prepare_inpaint_vectors(observation_batch).unsqueeze(1)
        if t is None:
            t = torch.randint(0, self.noise_steps, (B,), device=self.device).long()
        prediction_vector = torch.cat([x_0_inpaint, x_0], dim=2)
        if noise is None:
            noise = torch.randn_like(prediction_vector)
        ac = self.noise_scheduler.alphas_cumprod
        x_noisy = plan.add_noise(prediction_vector, noise, t, ac ** 0.5, (1 - ac) ** 0.5,
                                 inpaint=x_0_inpaint.reshape(B, -1) if self.inpaint_horizon > 0 else None)
        with torch.no_grad():
            noise_estimated = plan.unet_forward(x_noisy, t, None, use_cached_cond=True)
        self._last = (x_noisy, noise_estimated)
        return self.loss(noise.to(noise_estimated.device), noise_estimated)

    # ------------------------------------------------------------------------------------------
    # sampling
    # ------------------------------------------------------------------------------------------
    def add_constraints(self, x_t, x_inpaint):
        """ddpm:216-219 (in place)."""
        x_t[:, :, :self.inpaint_horizon, :] = x_inpaint
        return x_t

    def _run_loop(self, observation_batch, first_only, history, x_T=None, noise=None, seed=None):
        if first_only:
            observation_batch = {k: v[:1] for k, v in observation_batch.items()}
        B = observation_batch['position'].shape[0]
        plan = self._plan(B)
        self._bind_schedule(plan)
        plan.encode_cond(observation_batch['image'], observation_batch['position'], observation_batch['action'],
                         observation_batch['velocity'], return_cond=False)
        inpaint = self.prepare_inpaint_vectors(observation_batch).unsqueeze(1)
        rows = self.pred_horizon + self.inpaint_horizon
        if x_T is None:
            x_T = torch.rand(B, 1, rows, self.prediction_dim, device=self.device)
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        res = plan.sample(x_T, noise=noise, inpaint=inpaint if self.inpaint_horizon > 0 else None, history=history, seed=seed)
        return res, inpaint

    def validate(self, batch):
        """ddpm:176-214 — full-window batch -> (x_0, observation_batch, inpaint_vector), batch element 0 only."""
        observation_batch = self.prepare_observation_batch(batch)
        x_0, inpaint = self._run_loop(observation_batch, True, False)
        return x_0, observation_batch, inpaint

    def sample(self, batch, option=None, *, mode=None, batched=False, x_T=None, noise=None, seed=None):
        """ddpm:223-277 / ddim:23-74."""
        if mode == 'validation':  # call shape of the reference's evaluation/*.py scripts
            return self.validate(batch)
        for key, tensor in batch.items():
            if torch.is_tensor(tensor):
                batch[key] = tensor.to(self.device, non_blocking=True)
        # frames may also arrive as the uint8 (B, T, 96, 96, 3) HWC the simulator stores: decoded on the device (a quarter of the H2D bytes)
        observation_batch = {k: batch[k] if batch[k].dtype == torch.uint8 else batch[k].float() for k in ('image', 'position', 'action', 'velocity')}
        res, _ = self._run_loop(observation_batch, not batched, option == 'sample_history', x_T=x_T, noise=noise, seed=seed)
        if option == 'sample_history':
            _, hist = res
            return [hist[i] for i in range(hist.shape[0])]
        return res

    # ------------------------------------------------------------------------------------------
    # batch-prep helpers (ddpm:283-348)
    # ------------------------------------------------------------------------------------------
    def _slice(self, batch, sl):
        # .to()/.float() are no-ops for fp32 tensors already on the device: the slices stay views of the caller's batch
        return {k: batch[k][:, sl].to(self.device).float() for k in ('image', 'position', 'action', 'velocity')}

    def prepare_observation_batch(self, batch):
        return self._slice(batch, slice(None, self.obs_horizon))

    def prepare_prediction_batch(self, batch):
        return self._slice(batch, slice(self.obs_horizon, None))

    def prepare_obs_cond_vectors(self, observation_batch):
        """encoder + cat[pos, act, vel, img_feat] -> (B, T_obs, 135), computed by the fused conditioning kernels."""
        B, T = observation_batch['position'].shape[:2]
        plan = self._plan(B)
        cond = plan.encode_cond(observation_batch['image'], observation_batch['position'], observation_batch['action'],
                                observation_batch['velocity'])
        return cond.reshape(B, T, -1)

    def prepare_prediction_vectors(self, prediction_batch):
        """ddpm:332-338; position only when prediction_dim == 2 (mirrors prepare_inpaint_vectors)."""
        if self.prediction_dim == prediction_batch['position'].shape[-1]:
            return prediction_batch['position']
        return torch.cat([prediction_batch['position'], prediction_batch['action']], dim=-1)

    def prepare_inpaint_vectors(self, observation_batch):
        """ddpm:340-348; for position-only prediction (prediction_dim == 2) the position rows alone
        (the variant the reference keeps commented out at diffusion_ddim.py:75-86)."""
        pos = observation_batch['position'][:, -self.inpaint_horizon:, :]
        if self.prediction_dim == pos.shape[-1]:
            return pos
        act = observation_batch['action'][:, -self.inpaint_horizon:, :]
        return torch.cat([pos, act], dim=-1)


class Diffusion_DDIM(Diffusion_DDPM):
    """models/diffusion_ddim.py:19 — same loop; generate.py:28-35 swaps in a DDIMScheduler and re-purposes
    `noise_steps` as the number of DDIM steps.  `use_ddim(K)` does exactly that."""

    def use_ddim(self, num_steps):
        self.noise_scheduler = DDIMScheduler(num_train_timesteps=num_steps, beta_schedule='linear', clip_sample=False,
                                             prediction_type='epsilon')
        self.noise_steps = num_steps
        return self


class SamplingPipeline:
    """Keeps `depth` independent sampling calls in flight on separate CUDA streams, each with its own DenoisePlan
    (workspace + graphs).  At small batch a denoising step is bound by its ~95 dependent kernel launches, not by the
    machine, so a second (third) batch overlaps almost for free: 4 850 -> 6 560 (7 490) trajectories/s at batch 256.

        pipe = SamplingPipeline(model, depth=2, batch_max=256)
        tickets = [pipe.submit(batch) for batch in batches]       # returns immediately
        outs = [pipe.result(t) for t in tickets]                  # (B, 1, pred_h + ih, pred_dim) each
    """

    def __init__(self, model, depth=2, batch_max=None):
        from .engine import DenoisePlan
        self.model = model
        ne = model.noise_estimator
        base = model._plan(batch_max or max(model.batch_max, 1))
        model._bind_schedule(base)
        sch = model.noise_scheduler
        self.lanes = []
        for j in range(int(depth)):
            if j == 0:
                plan = base
            else:
                plan = DenoisePlan(attention=ne._attention, precision=base.precision, batch_max=base.batch_max, rows=base.rows,
                                   dim=base.dim, obs_horizon=base.obs_horizon, cond_dim=base.cond_dim, inpaint_rows=base.inpaint_rows,
                                   time_dim=ne.time_dim, device=base.device, graph_steps=base.cfg.graph_steps, split=ne.split, simple=ne._simple, encoder=ne.encoder)
                plan.load_unet_state_dict(ne.state_dict())
                plan.load_encoder_state_dict(model.vision_encoder.state_dict())
                plan.set_schedule(sch.kind, sch.coef_table(), sch.timesteps)
            self.lanes.append((plan, torch.cuda.Stream(device=base.device)))
        self._next = 0
        self._pending = {}

    def submit(self, batch, x_T=None, noise=None, seed=None):
        m = self.model
        plan, stream = self.lanes[self._next % len(self.lanes)]
        ticket = self._next
        self._next += 1
        stream.wait_stream(torch.cuda.current_stream(plan.device))
        with torch.cuda.stream(stream):
            obs = {k: batch[k].to(m.device, non_blocking=True) for k in ('image', 'position', 'action', 'velocity')}
            obs = {k: v if v.dtype == torch.uint8 else v.float() for k, v in obs.items()}
            B = obs['position'].shape[0]
            plan.encode_cond(obs['image'], obs['position'], obs['action'], obs['velocity'], return_cond=False)
            inpaint = m.prepare_inpaint_vectors(obs).unsqueeze(1)
            if x_T is None:
                x_T = torch.rand(B, 1, m.pred_horizon + m.inpaint_horizon, m.prediction_dim, device=m.device)
            if seed is None:
                seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            out = plan.sample(x_T, noise=noise, inpaint=inpaint if m.inpaint_horizon > 0 else None, seed=seed)
            done = torch.cuda.Event()
            done.record(stream)
        self._pending[ticket] = (out, done, obs)
        return ticket

    def result(self, ticket):
        out, done, _ = self._pending.pop(ticket)
        done.synchronize()
        out.record_stream(torch.cuda.current_stream(out.device))  # allocated on the lane stream, consumed on the caller's
        return out
