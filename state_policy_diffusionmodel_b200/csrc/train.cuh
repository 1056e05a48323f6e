// train.cuh — launch descriptors of the training-step kernels (bwd_kernels.cu, wgrad_tc.cu).
//
// The training step is the reference's `process_single_batch` + `loss.backward()` + Adam
// (models/diffusion_ddpm.py:115-173, train.py:104-107).  Gradients of activations use the activation
// dtype T of the plan (fp32 parity path / bf16 tensor-core path); every parameter gradient is fp32 in
// PyTorch layout and is ACCUMULATED with atomics (the step zeroes the flat gradient buffer first).
#pragma once
#include "common.cuh"

// ---- GroupNorm(1, C) backward (+GELU, +time-embedding add, +FiLM), one block per sample ---------
struct GnBwdArgs {
  const void* dy; int ld_dy;     // upstream gradient of the block output [M, C]
  const void* raw; int ld_raw;   // conv output the forward normalised [M, C]
  void* dx; int ld_dx;           // gradient w.r.t. raw [M, C]
  const float* stats; int P;     // forward partial (sum, sumsq) [B][P][2]
  const float* gamma; const float* beta;
  float* dgamma; float* dbeta;   // [C], atomically accumulated
  float* part;                   // optional scratch [B][2C]: per-sample (d gamma | d beta) rows, summed over B by a finalize kernel
                                 // (512 blocks hammering the same C addresses with atomics cost 50 us per launch)
  int act;                       // ACT_GELU: out = gelu(gn(raw));  ACT_NONE: out = gn(raw) (+temb) (*film)
  const float* temb; int temb_off;  // [B][SPDM_TEMB_WIDTH] rows (per sample) or null
  const float* film; int film_off;  // [B][SPDM_FILM_WIDTH] or null
  float* d_temb;                 // [B][SPDM_TEMB_WIDTH]: row b, cols temb_off.. written (not accumulated)
  float* d_film;                 // [B][SPDM_FILM_WIDTH]: scale grads at film_off, bias grads at film_off + C
  int HW, C;
  float eps;
};
template <typename T> bool launch_gn_bwd(const GnBwdArgs& a, int B, cudaStream_t s);   // true: a.part needs launch_gn_part_finalize
void launch_gn_part_finalize(const float* part, float* dgamma, float* dbeta, int B, int C, cudaStream_t s);

// ---- weight gradient on CUDA cores: dw[co][ci][tap] += sum_m x[shift(m, tap)][ci] * dy[m][co] --------
struct WgradArgs {
  const void* x; int ld_x;       // forward input  [M, Cin]
  const void* dy; int ld_dy;     // output gradient [M, Cout]
  long long M;
  int Cin, Cout, H, W, taps;     // taps 9: 3x3 pad 1 over (H, W) maps; taps 1: Linear
  float* dw;                     // PyTorch layout (Cout, Cin, 3, 3) or (Cout, Cin); fp32, accumulated
};
template <typename TX, typename TDY> void launch_wgrad_simt(const WgradArgs& a, cudaStream_t s);
// out[n] += sum_m dy[m][n]   (bias gradients)
template <typename T> void launch_colsum(const T* dy, int ld, long long M, int N, float* out, cudaStream_t s);

// ---- resampling -------------------------------------------------------------------------------------
// MaxPool2d(2) backward: dx[b, 2ho+i, 2wo+j, c] = (first arg-max of the window ? dy : 0) + add (skip-connection gradient)
template <typename T>
void launch_pool_bwd(const T* x, int ld_x, const T* dy, int ld_dy, const T* add, int ld_add, T* dx, int ld_dx, int B, int Ho, int Wo, int C,
                     cudaStream_t s);
// bilinear x2 (align_corners=True) backward, gather form: dx [B, Hi, Wi, C] from dy [B, 2Hi, 2Wi, C]
template <typename T> void launch_upsample_bwd(const T* dy, int ld_dy, T* dx, int ld_dx, int B, int Hi, int Wi, int C, cudaStream_t s);

// ---- SelfAttention pieces -----------------------------------------------------------------------------
// LayerNorm backward over C: dx = ln'(x)^T (dy) + add;  dgamma/dbeta accumulated
template <typename T>
void launch_layernorm_bwd(const T* dy, int ld_dy, const T* x, int ld_x, const float* g, const T* add, int ld_add, T* dx, int ld_dx,
                          float* dgamma, float* dbeta, long long M, int C, cudaStream_t s);
// attention core backward: qkv [B*L][3C], o = forward output [B*L][C], d_o [B*L][C] -> d_qkv [B*L][3C]
template <typename T> void launch_sdpa_bwd(const T* qkv, const T* o, const T* d_o, T* d_qkv, int B, int L, int C, int heads, cudaStream_t s);
bool launch_sdpa_fwd_mma(const bf16* qkv, bf16* out, int B, int L, int C, int heads, cudaStream_t s);  // bf16, mma.sync; false = shape too large
template <typename T> void launch_gelu_fwd(const T* x, T* y, long long n, cudaStream_t s);              // y = gelu(x)
template <typename T> void launch_gelu_bwd(const T* dy, const T* pre, T* dx, long long n, cudaStream_t s);  // dx = dy * gelu'(pre)

// ---- ends of the network ------------------------------------------------------------------------------
// MSELoss(noise, eps_hat) (mean over B*rows*dim) + its gradient through outc (1x1 conv 64 -> 1, + unpad):
//   loss += sum (eps_hat - noise)^2 / N ;  d_act[M0, C] ; d_w[C], d_b accumulated.   (ddpm:171, Unet_FiLmLayer.py:264,310)
// B_valid > 0: only the first B_valid samples are real (the rest pad a ragged batch to the tile granularity): N = B_valid*rows*dim
// and the padding samples get a zero gradient, so nothing downstream of them reaches a parameter gradient.
template <typename T>
void launch_mse_outc_bwd(const float* eps_hat, const float* noise, const T* act, int ld, const float* w, T* d_act, float* d_w, float* d_b,
                         float* loss, int B, int H, int W, int C, int rows, int dim, int lh, int lw, cudaStream_t s, int B_valid = 0);
// inc.first weight gradient: dw[64][1][3][3] += sum x_pad[shift] * d_raw     (Unet_FiLmLayer.py:101, pad_to folded in)
template <typename T>
void launch_conv_in_wgrad(const float* x, const T* d_raw, float* dw, int B, int H, int W, int rows, int dim, int lh, int lw, cudaStream_t s);

// ---- conditioning -------------------------------------------------------------------------------------
void launch_posenc_silu(const long long* t, int n, const float* inv_freq, float* out, int time_dim, cudaStream_t s);  // silu(pos_encoding(t))
void launch_mish_bwd(const float* dy, const float* x, float* dx, long long n, cudaStream_t s);
// d_cond (B, T*cond_dim) -> gradient of the image features (B*T, cond_dim - 7)
void launch_gather_feat_grad(const float* d_cond, float* d_feat, int B, int T, int cond_dim, cudaStream_t s);
// Autoencoder.encoder conv stack backward (recomputes the three layers per 8-row strip):
//   d_feat [n][9216] (hwc order, gradient of the post-ReLU features), feat = forward features (ReLU mask)
//   weight grads accumulated in PyTorch layouts (16,3,2,2) (32,16,2,2) (64,32,2,2) + biases
void launch_enc_convs_bwd(const float* img, const float* w1, const float* b1, const float* w2t, const float* b2, const float* w3t,
                          const float* feat, const float* d_feat, float* dw1, float* db1, float* dw2, float* db2, float* dw3, float* db3,
                          int n, cudaStream_t s);
// dwl (128, 9216 chw) += tmp (128, 9216 hwc)
void launch_enc_linear_grad_permute(const float* tmp_hwc, float* dwl, cudaStream_t s);
// (128, 9216 chw) -> [128][9216 hwc] fp32 (B operand of the feature-gradient GEMM)
void launch_pack_enc_linear_t(const float* w, float* out, cudaStream_t s);

// ---- vision encoder as patch GEMMs on the tensor cores (bf16 training path; layouts in bwd_kernels.cu) ----------
// img: frame (b, t) at img + b * bstride + t * 3*96*96 (bstride = T*3*96*96 for a contiguous observation window)
void launch_enc_conv1_fwd(const float* img, const float* w1, const float* b1, bf16* c1p, int n, int T, long long bstride, cudaStream_t s);
// the same from uint8 HWC frames (n, 96, 96, 3), decoded x / 255 while the input strip is staged
void launch_enc_conv1_fwd_u8(const uint8_t* img_hwc, const float* w1, const float* b1, bf16* c1p, int n, cudaStream_t s);
void launch_decode_u8_hwc(const uint8_t* img, float* out, long long frames, int H, int W, cudaStream_t s);  // data_kernels.cu
void launch_enc_conv1_wgrad(const float* img, const bf16* d1, const bf16* act, float* dw1, float* db1, int n, int T, long long bstride, cudaStream_t s);  // act != null: ReLU mask applied on load
// over [rows][64] bf16; colsum64 (or null) += column sums of the masked gradient
void launch_relu_mask(const bf16* d, const bf16* act, bf16* out, long long n, float* colsum64, cudaStream_t s);
void launch_enc_pack_w2(const float* w2, bf16* w2p, bf16* w2pT, cudaStream_t s);  // (32,16,2,2) -> block-diagonal [64][128] and its transpose
void launch_enc_pack_w3(const float* w3, bf16* w3p, bf16* w3pT, cudaStream_t s);  // (64,32,2,2) -> [64][128] and its transpose
void launch_enc_pack_b2(const float* b2, float* b2p, cudaStream_t s);             // (32,) -> [64] (one copy per patch of the pair)
void launch_enc_pack_linear_T16(const float* w, bf16* out, cudaStream_t s);
void launch_enc_unpack_grads(const float* g2, const float* g3, const float* gb2, float* dw2, float* db2, float* dw3, cudaStream_t s);

// ---- dgrad weight packs (the data gradient of a conv/Linear is a forward GEMM with these weights) ----------
void launch_pack_conv_dgrad_f32(const float* oihw, float* out, int Cout, int Cin, cudaStream_t s);   // -> [tap'][Cout][Cin], tap' = 8 - tap
void launch_pack_conv_dgrad_bf16(const float* oihw, bf16* out, int Cout, int Cin, cudaStream_t s);  // -> [Cin][tap'][Cout]
void launch_pack_linear_dgrad_bf16(const float* nk, bf16* out, int N, int K, cudaStream_t s);       // -> [K][N]

// ---- optimizer ------------------------------------------------------------------------------------------
void launch_sumsq(const float* g, long long n, float* out, cudaStream_t s);  // out += sum g^2
// torch.optim.Adam (default flags) on flat buffers; if sumsq != null the gradient is first scaled by
// min(1, max_norm / (sqrt(*sumsq) + 1e-6))  (clip_grad_norm_, Lightning gradient_clip_val) and grad_scale
// (1 / world_size after a summing all-reduce).
void launch_adam(float* p, float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps, int step,
                 const float* sumsq, float max_norm, float grad_scale, cudaStream_t s);

// packed conv gradient [tap][Cout][Cin] -> PyTorch (Cout, Cin, 3, 3)
void launch_unpack_conv_grad(const float* packed, float* dst, int Cout, int Cin, cudaStream_t s);
// FiLM Linears on the tensor cores (bf16 training)
void launch_film_pack16(const float* src, bf16* wf, bf16* wb, int C2, int G, int GP, int off, cudaStream_t s);
void launch_mish_pad_bf16(const float* cond, bf16* out, int B, int Bpad, int G, int GP, cudaStream_t s);       // [Bpad][GP], zero padded
void launch_mish_bwd_bf16(const bf16* dy, int ld_dy, const float* x, float* dx, int B, int G, cudaStream_t s);
long long bwd_launch_count();
extern long long wgrad_tc_launch_count_value;

// ---- tcgen05 weight gradient (wgrad_tc.cu): same contract as launch_wgrad_simt for bf16 operands -------------
bool wgrad_tc_supported(int Cin, int Cout, int H, int W, int taps, int ld_x, int ld_dy, long long M);
// returns 0 on success; negative when the shape is unsupported (caller falls back to nothing: it is an error)
int wgrad_tc_launch(const bf16* x, int ld_x, const bf16* dy, int ld_dy, long long M, int Cin, int Cout, int H, int W, int taps, float* dw,
                    cudaStream_t s);
const char* wgrad_tc_last_error();
