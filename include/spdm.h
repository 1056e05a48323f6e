/*
 * spdm.h — C ABI of the B200-native denoising hot path (libspdm.so).
 *
 * The reference (rafaelsoStanford/State_Policy_DiffusionModel) has no FFI layer: its
 * boundary is the PyTorch module surface.  Every entry point below names the reference
 * call it stands in for (paths relative to /root/reference).  The Python mirror in
 * state_policy_diffusionmodel_b200/ binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - All tensor arguments are caller-owned, contiguous DEVICE memory (fp32 unless noted; timesteps int64).
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work on it.
 *   - Return 0 on success, negative on error; spdm_last_error() gives a thread-local message.
 *   - A plan is bound to one device and is not thread-safe.  There is no CPU fallback.
 */
#ifndef SPDM_H
#define SPDM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct spdm_plan spdm_plan;

enum { SPDM_VARIANT_ATTENTION = 0, SPDM_VARIANT_NO_ATTENTION = 1,
       SPDM_VARIANT_SIMPLE_UNET = 2 /* models/simple_Unet.py:260-339, the model='UNet' default; fp32 path, inference only */ };
enum { SPDM_PRECISION_FP32 = 0, SPDM_PRECISION_BF16 = 1,
       SPDM_PRECISION_TF32 = 2 /* fp32 activations; the 3x3 convs on tcgen05.mma.kind::tf32 (sampling / forward only) */ };
enum { SPDM_SCHED_DDPM = 0, SPDM_SCHED_DDIM = 1 };
enum { SPDM_FLAG_SCHEDULER_ONLY = 1,    /* plan without U-Net weights/workspace: spdm_step, spdm_add_noise only */
       SPDM_FLAG_ENCODER_RESNET18 = 2   /* vision encoder = ResNet18 with GroupNorm(C/16) (models/Unet_FiLmLayer.py:316-386, `VisionEncoder()`):
                                           weights "vision_encoder.<torchvision resnet18 key>", 512 features per frame => cond_dim = 519;
                                           inference only */ };
#define SPDM_FLAG_SPLIT(n) (((n) & 0xF) << 8) /* run n sub-batches of every denoising step concurrently (spdm_sample) */

typedef struct spdm_config {
  int32_t variant;       /* models/Unet_FiLmLayer.py:240 (0), Unet_FiLmLayer_noAttention.py:240 (1), simple_Unet.py:260 (2) */
  int32_t precision;     /* SPDM_PRECISION_*: fp32 = CUDA-core path, bf16 / tf32 = tcgen05 paths     */
  int32_t batch_max;     /* largest number of trajectories a call will carry                        */
  int32_t rows;          /* pred_horizon + inpaint_horizon   (models/diffusion_ddpm.py:252)         */
  int32_t dim;           /* prediction_dim                   (models/diffusion_ddpm.py:252)         */
  int32_t obs_horizon;   /* T_obs                            (models/diffusion_ddpm.py:283-298)     */
  int32_t cond_dim;      /* per-frame conditioning width; global_cond_dim = obs_horizon*cond_dim,
                            0 = unconditional U-Net (Unet_FiLmLayer.py:144,169)                     */
  int32_t inpaint_rows;  /* inpaint_horizon                  (models/diffusion_ddpm.py:216-219)     */
  int32_t time_dim;      /* 256                              (models/Unet_FiLmLayer.py:241)         */
  int32_t device;        /* CUDA device ordinal                                                     */
  int32_t graph_steps;   /* denoise steps unrolled per CUDA graph launch (0 = no graphs)            */
  int32_t flags;         /* SPDM_FLAG_*                                                            */
} spdm_config;

/* Diffusion_DDPM.__init__ (models/diffusion_ddpm.py:22-88): allocate weights + workspace. */
int spdm_plan_create(spdm_plan** out, const spdm_config* cfg);
int spdm_plan_destroy(spdm_plan* plan);

/* nn.Module.load_state_dict for noise_estimator.* / vision_encoder.* (SURVEY A.2 key names,
 * without the Lightning prefix for the U-Net; encoder keys are "vision_encoder.{0,2,4,7}.*").
 * `src` is fp32 device memory in PyTorch layout; it is repacked into the plan. */
int spdm_plan_load_weight(spdm_plan* plan, const char* name, const float* src,
                          const int64_t* shape, int32_t ndim, void* stream);
/* Number of expected weight tensors still missing (0 = ready); names via spdm_last_error(). */
int spdm_plan_missing_weights(spdm_plan* plan);

/* DDPMScheduler/DDIMScheduler.set_timesteps (call sites models/diffusion_ddpm.py:204,257;
 * diffusion_ddim.py:57,67).  coef is HOST memory, K rows of 8 floats:
 *   {sqrt(1-abar_t), sqrt(abar_t), k_x0, k_x, k_eps, k_noise, clip, 0}
 *   x0 = (x - c0*eps)/c1 ;  clip > 0: x0 = clamp(x0, -clip, clip)  (clip_sample=True, clip_sample_range; the
 *   reference passes clip_sample=False, models/diffusion_ddpm.py:68) ;  x_prev = k_x0*x0 + k_x*x + k_eps*eps + k_noise*z
 * timesteps is HOST int64[K] (descending), the value fed to the U-Net time embedding. */
int spdm_plan_set_schedule(spdm_plan* plan, int32_t kind, int32_t K, const float* coef,
                           const int64_t* timesteps, void* stream);

/* Autoencoder.encoder (models/encoder/autoencoder.py:11-20): images (n,3,96,96) -> (n,128); with SPDM_FLAG_ENCODER_RESNET18 the
 * ResNet18-GroupNorm encoder (models/Unet_FiLmLayer.py:316-386): -> (n,512). */
int spdm_encode_images(spdm_plan* plan, const float* images, float* out, int32_t n, void* stream);

/* prepare_obs_cond_vectors (models/diffusion_ddpm.py:317-330) + the six FiLM cond_encoder
 * linears (Unet_FiLmLayer.py:149-154).  images (B,T,3,96,96), position (B,T,2), action (B,T,3),
 * velocity (B,T,2).  Leaves the conditioning cached in the plan for spdm_sample. */
int spdm_encode_cond(spdm_plan* plan, const float* images, const float* position,
                     const float* action, const float* velocity, int32_t B, void* stream);
/* Same with the frames as the simulator / dataset stores them (generateData/trajectory_control_utils.py:170 divides by 255 on
 * the way into the zarr file): uint8 (B,T,96,96,3) HWC, decoded x / 255 inside the encoder's first conv -- a quarter of the
 * bytes of the fp32 frames over PCIe and out of HBM. */
int spdm_encode_cond_u8(spdm_plan* plan, const uint8_t* images_hwc, const float* position, const float* action,
                        const float* velocity, int32_t B, void* stream);
/* Same, from an already built obs_cond (B, T*cond_dim). */
int spdm_set_cond(spdm_plan* plan, const float* obs_cond, int32_t B, void* stream);
/* Copy of the cached obs_cond (B, T*cond_dim) for inspection / parity tests. */
int spdm_get_cond(spdm_plan* plan, float* out, int32_t B, void* stream);

/* UNet_Film.forward / UNet_Film_noAttention.forward (Unet_FiLmLayer.py:277-312).
 * x (B,1,rows,dim); t int64 device, t_count = 1 (broadcast) or B; y (B, T*cond_dim) or NULL
 * (NULL = reuse the cached conditioning if use_cached_cond, else unconditional). */
int spdm_unet_forward(spdm_plan* plan, const float* x, const int64_t* t, int32_t t_count,
                      const float* y, int32_t use_cached_cond, float* out, int32_t B, void* stream);

/* One scheduler.step + add_constraints (models/diffusion_ddpm.py:211-213,216-219) for schedule
 * index `step` (0..K-1).  noise may be NULL (no noise term).  x_out may alias x. */
int spdm_step(spdm_plan* plan, const float* x, const float* eps, const float* noise,
              const float* inpaint, float* x_out, int32_t step, int32_t B, void* stream);

/* Diffusion_DDPM.sample / Diffusion_DDIM.sample loop (diffusion_ddpm.py:269-276,
 * diffusion_ddim.py:67-74) over all K schedule steps for B trajectories, using the cached
 * conditioning.  x_T (B,1,rows,dim) is the start sample; noise (K,B,1,rows,dim) or NULL
 * (NULL: Philox noise from `seed` where the schedule has a noise term); inpaint
 * (B,1,inpaint_rows,dim) or NULL; out (B,1,rows,dim); history (K+1,B,1,rows,dim) or NULL. */
int spdm_sample(spdm_plan* plan, const float* x_T, const float* noise, const float* inpaint,
                float* out, float* history, uint64_t seed, int32_t B, void* stream);

/* DDPMScheduler.add_noise + add_constraints (models/diffusion_ddpm.py:167-168):
 * x_noisy = sqrt_ab[t_b]*x0 + sqrt_1mab[t_b]*noise, first inpaint_rows rows <- inpaint.
 * sqrt_ab / sqrt_1mab are device tables indexed by t (int64 device, B entries). */
int spdm_add_noise(spdm_plan* plan, const float* x0, const float* noise, const int64_t* t,
                   const float* sqrt_ab, const float* sqrt_1mab, const float* inpaint,
                   float* out, int32_t B, void* stream);

/* ---- training step (models/diffusion_ddpm.py:115-173, train.py:104-107) ------------------------------------------
 * The caller owns two flat fp32 device buffers (parameters, gradients) holding every trainable tensor in PyTorch
 * layout; nn.Parameter storage can alias them, so torch optimizers and NCCL all-reduces see ordinary tensors. */
/* Allocate the data-gradient weight twins; every weight must be uploaded (again) afterwards. */
int spdm_train_enable(spdm_plan* plan);
/* Tensor `name` (load_weight names) lives at `offset` floats into both flat buffers. */
int spdm_train_bind(spdm_plan* plan, const char* name, int64_t offset, const int64_t* shape, int32_t ndim);
int spdm_train_set_buffers(spdm_plan* plan, float* params, float* grads, int64_t total);
/* load_state_dict from the flat parameter buffer (after every optimizer step). */
int spdm_train_sync_weights(spdm_plan* plan, void* stream);
/* Diffusion_DDPM.process_single_batch + loss.backward() (ddpm:128-173): images (B,T,3,96,96), position (B,T,2),
 * action (B,T,3), velocity (B,T,2) = the observation window; x0 (B,1,rows,dim) = cat[inpaint rows, prediction];
 * noise (B,1,rows,dim); t int64 (B,); sqrt_ab / sqrt_1mab = DDPMScheduler.add_noise tables; inpaint (B,inpaint_rows*dim).
 * Zeroes the gradient buffer, then leaves d loss / d parameter in it; loss_out = one device float (MSE, mean). */
int spdm_train_fwd_bwd(spdm_plan* plan, const float* images, const float* position, const float* action,
                       const float* velocity, const float* x0, const float* noise, const int64_t* t,
                       const float* sqrt_ab, const float* sqrt_1mab, const float* inpaint, float* loss_out,
                       int32_t B, void* stream);
/* `images` of the next spdm_train_fwd_bwd calls is a strided view: frame (b, t) at images + b*stride + t*3*96*96 floats
 * (0 = contiguous).  The reference slices the observation window out of the full recording (ddpm:283-298). bf16 plans only. */
int spdm_train_set_image_stride(spdm_plan* plan, int64_t stride);
/* Ragged batches (the reference's DataLoader has no drop_last, utils/load_data.py:174): the bf16 path needs B to be a multiple of
 * spdm_plan_batch_multiple(); the caller pads the batch with any valid samples and declares how many at the head are real.
 * The loss is then the mean over the real samples and the padding samples contribute no gradient.  0 = all real (default). */
int spdm_train_set_valid(spdm_plan* plan, int32_t valid);
/* Data-parallel overlap: make `stream` wait until every gradient of completion phase `phase` of the latest
 * spdm_train_fwd_bwd is final.  Phase 0: outc, sa4-sa6, up1-up3 (convs, norms, attention); phase 1: the rest of the U-Net
 * except the emb_layer / cond_encoder Linears; phase 2: everything (those Linears and the vision encoder). */
int spdm_train_wait_phase(spdm_plan* plan, int32_t phase, void* stream);
/* torch.nn.utils.clip_grad_norm_(max_norm) (Lightning gradient_clip_val=0.5, train.py:107; max_norm <= 0: off) followed
 * by torch.optim.Adam.step() (ddpm:115-125) over n floats; `step` counts from 1; grad_scale multiplies the gradient first
 * (1/world_size after a summing all-reduce); scratch = one device float. */
int spdm_adam_step(float* params, float* grads, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                   float eps, int32_t step, float max_norm, float grad_scale, float* scratch, void* stream);

/* Kernel-class profile of one denoising step (U-Net forward + posterior update) run EAGERLY with CUDA
 * events around every launch, averaged over `reps` steps: out[c*4 + {0,1,2,3}] = {ms, launches,
 * algorithmic FLOPs, algorithmic bytes} for class c.  Needs a schedule and (if conditional) cached cond. */
#define SPDM_PROFILE_CLASSES 10
enum { SPDM_PC_CONV3 = 0, SPDM_PC_GEMM1, SPDM_PC_APPLY, SPDM_PC_STATS, SPDM_PC_RESAMPLE, SPDM_PC_LN,
       SPDM_PC_SDPA, SPDM_PC_IO, SPDM_PC_STEP, SPDM_PC_CONV3_GN /* cluster split-K conv + fused GroupNorm apply */ };
int spdm_profile_step(spdm_plan* plan, int32_t B, int32_t reps, double* out, void* stream);

/* Microbenchmark of one tcgen05 implicit-GEMM launch (3x3 conv when taps == 9, Linear when taps == 1) on
 * zero-filled scratch buffers: average milliseconds over `iters` launches.  dbg = 0 runs the real kernel;
 * other values switch parts of it off (perf experiments, tests/perf_conv.py). */
int spdm_microbench_conv(int32_t H, int32_t W, int32_t B, int32_t Cin, int32_t Cout, int32_t taps,
                         int32_t dbg, int32_t iters, float* ms_out);

/* Introspection used by bench.py / tests. */
int64_t spdm_plan_launch_count(spdm_plan* plan);     /* kernels enqueued so far by this plan  */
int64_t spdm_plan_workspace_bytes(spdm_plan* plan);
int32_t spdm_plan_batch_multiple(spdm_plan* plan);  /* granularity of B on the tensor-core path (1 on the fp32 path) */
/* Debug tap (tests only): spdm_unet_forward that additionally copies the internal activation `tap_name`
 * ("inc", "down1", "sa1", "x2", "bot3", "up1", "u3", "<block>.first", ... ) to tap_out as fp32
 * (B, C, H, W).  Returns the element count written, or negative. */
int64_t spdm_debug_forward(spdm_plan* plan, const float* x, const int64_t* t, int32_t t_count,
                           const float* y, int32_t use_cached_cond, float* out, int32_t B,
                           const char* tap_name, float* tap_out, void* stream);

/* ---- dataset -> batch on the device (utils/load_data.py:11-144, utils/data_utils.py:18-62) ------------
 * The dataset arrays stay resident in device memory; one call gathers a batch of B strided windows, exactly as
 * CarRacingDataset.__getitem__ + the DataLoader's default collate build it on the host:
 *   images      uint8 (N, H, W, 3) as the simulator wrote them (decoded x / 255 in flight), or float (N, 3, H, W)
 *   starts      int64 [B], first frame of every window (create_sample_indices_sparse, data_utils.py:46-56)
 *   out_image   (B, T_img, 3, H, W): frames start + t*step, t < T_img <= T (T_img = T is the reference's item; obs_horizon is
 *               all the model reads, models/diffusion_ddpm.py:283-298; 0 = no images)
 *   out_position (B, T, 2): min-max normalised with the scalar position statistics, centred on the window's first point,
 *               halved (load_data.py:128-133); out_translation (B, 2) = that first point
 *   out_velocity (B, T, 2), out_action (B, T, 3): min-max normalised per dimension (load_data.py:78-81)
 * Arithmetic is fp32, one IEEE rounding per operation in the reference's order: bit-identical to the numpy float32 path. */
enum { SPDM_IMAGES_U8_HWC = 0, SPDM_IMAGES_F32_CHW = 1 };
typedef struct spdm_data_stats {
  float pos_min, pos_max;          /* scalars: mean of the per-window minima / maxima (load_data.py:58-71) */
  float vel_min[2], vel_max[2];    /* get_data_stats(velocity)  (data_utils.py:10-16)                       */
  float act_min[3], act_max[3];    /* get_data_stats(action)                                                */
} spdm_data_stats;
int spdm_gather_windows(const void* images, int32_t image_kind, int32_t H, int32_t W, const float* position,
                        const float* velocity, const float* action, const int64_t* starts, int32_t B, int32_t T,
                        int32_t T_img, int32_t step, const spdm_data_stats* stats /* host */, float* out_image,
                        float* out_position, float* out_velocity, float* out_action, float* out_translation, void* stream);
/* utils/data_utils.py:35-40 unnormalize_position on (n_samples, rows_per_sample, 2) device positions with the per-sample
 * translation (n_samples, 2): sampled trajectories / whole sample_history stacks without a host round trip. */
int spdm_unnormalize_position(const float* npos, const float* translation, float pos_min, float pos_max, int64_t n_samples,
                              int32_t rows_per_sample, float* out, void* stream);
int64_t spdm_data_launch_count(void);

const char* spdm_last_error(void);
const char* spdm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SPDM_H */
