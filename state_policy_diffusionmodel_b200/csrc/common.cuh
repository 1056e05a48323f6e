// common.cuh — shared types for the spdm CUDA sources (sm_100a only).
//
// Activation layout: every feature map of the U-Net lives in HBM channels-last,
//   act[b][h][w][c]  (c fastest),  row r = (b*H + h)*W + w,  element (r, c) at  base[r*ld + c]
// i.e. a row-major [M = B*H*W, C] matrix with leading dimension ld >= C (ld > C when the map is a
// channel slice of a concat buffer).  This is at the same time the token-major (B, L, C) view that
// SelfAttention needs (reference models/Unet_FiLmLayer.py:74) and the K-major A operand of the
// implicit-GEMM convolution: a 4-D TMA box {64 ch, W, Hb, Bt} at coordinate (c0, dx, h0+dy, b0)
// lands in shared memory as a 128-row x 128-byte K-major tile, with the conv's zero padding
// supplied by TMA out-of-bounds fill.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

typedef __nv_bfloat16 bf16;

enum : int { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2 };
enum : int { TEMB_NONE = 0, TEMB_ROW0 = 1, TEMB_PER_SAMPLE = 2, TEMB_STEP = 3 };
enum : int { EPI_STATS = 1, EPI_BIAS = 2, EPI_GELU = 4, EPI_RESID = 8, EPI_VT = 16, EPI_APPLY = 32, EPI_RELU = 64, EPI_MASK = 128 /* out = resid > 0 ? acc : 0 (ReLU backward) */ };

#define SPDM_FILM_WIDTH 1792 /* sum over the 6 stages of 2*C_out */
#define SPDM_TEMB_WIDTH 896  /* sum over the 6 stages of C_out   */
#define SPDM_MAX_PARTIALS 64 /* upper bound on GroupNorm partial slots per sample */

// Programmatic dependent launch: every kernel of the forward path is launched with the programmatic-stream-
// serialization attribute.  pdl_wait() blocks until the preceding kernel in the stream has completed and its
// writes are visible (so everything after it is ordinary stream order); whatever a kernel does before it (barrier
// init, TMEM allocation, descriptor prefetch) overlaps with the tail of its predecessor.  pdl_trigger() lets the
// successor start that prologue as soon as every CTA of this kernel is resident.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

extern int g_spdm_pdl;  // 1: launch with the PDL attribute (default), 0: plain launches (kernels.cu)
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_spdm_pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float gelu_exact(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// 8-wide vector load/store of activations (C and ld are always multiples of 8)
__device__ __forceinline__ void load8(const float* p, float v[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float v[8]) {
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
}
__device__ __forceinline__ void store8(float* p, const float v[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float v[8]) {
  uint4 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = t;
}

// ---------------------------------------------------------------------------------------------
// Launch descriptors (plain structs passed by value)
// ---------------------------------------------------------------------------------------------
struct GemmSimtArgs {
  const void* in;     // [M, Cin] activations (TIn), leading dim ld_in
  const float* w;     // [taps][Cin][Cout] fp32
  const float* bias;  // [Cout] or null
  const void* resid;  // [M, Cout] (TOut, leading dim ld_res) or null
  void* out;          // [M, Cout] (TOut, leading dim ld_out)
  int M, Cin, Cout;
  int ld_in, ld_out, ld_res;
  int H, W;     // spatial geometry of a sample (rows per sample = H*W); used when taps == 9
  int taps;     // 1 = flat GEMM, 9 = 3x3 conv pad 1
  int act;      // ACT_*
};

struct ApplyArgs {
  const void* raw;     // [M, C] raw conv output (TIn), ld_in
  void* out;           // [M, C] (TOut), ld_out
  const float* stats;  // [B][P][2] partial (sum, sumsq)
  int P;
  const float* gamma;
  const float* beta;
  const float* temb;   // [rows][SPDM_TEMB_WIDTH] or null
  int temb_mode, temb_off;
  const int* step_ptr; // device int: schedule index (TEMB_STEP); the row used is *step_ptr + step_off
  int step_off;        // position of this launch's denoising step inside a captured multi-step graph (the counter advances once per graph)
  const float* film;   // [B][SPDM_FILM_WIDTH] or null; scale at film_off, bias at film_off + C
  int film_off;
  int HW, C, ld_in, ld_out;
  int act;
  float eps;
};

// Device-resident per-call parameters of the graphed sampling loop: the captured graph only holds a
// pointer to this block, so one graph serves every call (any noise / inpaint / history / seed).
struct StepDyn {
  int step; int pad;
  unsigned long long seed;
  const float* noise;    // [K][B][n] or null
  const float* inpaint;  // [B][inpaint_elems] or null
  float* history;        // [(K+1)][B][n] or null (slot step+1 is written)
  int use_philox; int pad2;
};

struct StepArgs {
  const float* x;
  const float* eps;
  float* x_out;
  const float* coef;     // [K][8]
  const StepDyn* dyn;    // non-null: step index, noise, inpaint, history, seed come from device memory
  // dyn == null: explicit single step (spdm_step)
  const float* noise;    // [B][n] for this step, or null
  const float* inpaint;  // [B][inpaint_elems] or null
  int step_host;
  int n;                 // elements per sample (rows*dim)
  int inpaint_elems;     // inpaint_rows*dim
  int B;                 // samples handled by this launch: [b0, b0 + B) of a batch of B_total
  int b0, B_total;
  int step_off;          // dyn != null: schedule index = dyn->step + step_off (see ApplyArgs::step_off)
};

// launchers implemented in kernels.cu ------------------------------------------------------------
template <typename TI, typename TO> void launch_gemm_simt(const GemmSimtArgs& a, cudaStream_t s);
template <typename TI, typename TO> void launch_apply(const ApplyArgs& a, int B, cudaStream_t s);
// GroupNorm (statistics + apply, +GELU/temb/FiLM) of a split-K conv: sums the S fp32 partial tiles per element first.
// One block per sample; HW*C <= 16384 and a multiple of 1024.  a.raw is ignored, partial = [S][M][C] fp32.
void launch_apply_partial(const ApplyArgs& a, const float* partial, int S, long long M, int B, cudaStream_t s);
template <typename T> void launch_stats(const T* raw, float* stats, int B, int HW, int C, int ld, cudaStream_t s);  // P = 1
template <typename T> void launch_pool(const T* in, int ld_in, T* out, int ld_out, int B, int Ho, int Wo, int C, cudaStream_t s);
template <typename T> void launch_upsample(const T* in, int ld_in, T* out, int ld_out, int B, int Hi, int Wi, int C, cudaStream_t s);
template <typename T> void launch_layernorm(const T* in, int ld_in, T* out, int ld_out, const float* g, const float* b, long long M, int C, cudaStream_t s);
template <typename T> void launch_sdpa(const T* qkv, T* out, int B, int L, int C, int heads, cudaStream_t s);
// inc.first + the sample's GroupNorm partial sums (P = 1) in one launch
template <typename T> void launch_conv_in(const float* x, const float* w, T* out, float* stats, int B, int H, int W, int rows, int dim, int lh, int lw, cudaStream_t s);
// bf16 path, H*W == 256: inc.first + GroupNorm + GELU in one launch, out = activated map [B*256][64]
void launch_conv_in_gn(const float* x, const float* w, const float* gamma, const float* beta, bf16* out, int B, int H, int W, int rows,
                       int dim, int lh, int lw, cudaStream_t s);
// outc fused with the posterior update of the graphed loop (a.dyn != null); act = last activation [B*H*W][ld]
template <typename T> void launch_outc_step(const StepArgs& a, const T* act, int ld, const float* w, const float* bias, int H, int W, int C, int rows, int dim, int lh, int lw, cudaStream_t s);
template <typename T> void launch_outc(const T* x, int ld, const float* w, const float* bias, float* eps, int B, int H, int W, int C, int rows, int dim, int lh, int lw, cudaStream_t s);
template <typename T> void launch_to_nchw(const T* in, int ld, float* out, int B, int HW, int C, cudaStream_t s);
void launch_step(const StepArgs& a, cudaStream_t s);
void launch_advance(int* step_ptr, int delta, cudaStream_t s);
void launch_delay(long long cycles, cudaStream_t s);  // spin kernel (profiling: lets the host run ahead)
void launch_temb(const long long* t_dev, int n_t, const float* inv_freq, const float* w_cat, const float* b_cat, float* out, int time_dim, cudaStream_t s);
void launch_mish(const float* in, float* out, long long n, cudaStream_t s);
// w2t / w3t: conv weights transposed to [Cin*4][Cout]; feat: [n][9216] in (pixel, channel) order, fp32 or bf16
template <typename TO> void launch_enc_convs(const float* img, const float* w1, const float* b1, const float* w2t, const float* b2, const float* w3t, const float* b3, TO* feat, int n, cudaStream_t s);
void launch_cast_f32(const bf16* in, float* out, long long n, cudaStream_t s);
void launch_pack_enc_linear_bf16(const float* w, bf16* out, cudaStream_t s);  // (128, 9216 chw) -> bf16 [128][9216 hwc]
void launch_build_cond(const float* pos, const float* act, const float* vel, const float* feat, float* cond, int B, int T, int cond_dim, cudaStream_t s);
void launch_add_noise(const float* x0, const float* noise, const long long* t, const float* sa, const float* sb, const float* inpaint, float* out, int n, int inpaint_elems, int B, cudaStream_t s);
// weight repack (fp32 PyTorch layout -> kernel layout)
void launch_pack_conv_f32(const float* oihw, float* out, int Cout, int Cin, int k, cudaStream_t s);  // -> [tap][Cin][Cout]
void launch_pack_conv_bf16(const float* oihw, bf16* out, int Cout, int Cin, int k, cudaStream_t s); // -> [Cout][tap][Cin]
// 3x3, Cout and Cin multiples of 32 (else false, nothing launched): one pass -> [Cout][tap][Cin] and/or [Cin][8 - tap][Cout] (either may be null)
bool launch_pack_conv3_bf16(const float* oihw, bf16* out_fwd, bf16* out_dgrad, int Cout, int Cin, cudaStream_t s);
bool launch_unpack_conv3_grad(const float* packed, float* dst, int Cout, int Cin, cudaStream_t s);   // [tap][Cout][Cin] fp32 -> OIHW
void launch_pack_conv_fold2_bf16(const float* oihw, bf16* out, int Cout, int Cin, cudaStream_t s);    // W = 2 fold: [2 Cout][9][2 Cin]
void launch_pack_conv_pfold_bf16(const float* oihw, bf16* out, int Cin, cudaStream_t s);               // pair fold, Cout = 64: [128][3][4][Cin]
void launch_pack_linear_f32(const float* nk, float* out, int N, int K, int ld_out, int col_off, cudaStream_t s);  // -> [K][ld_out] at col_off
void launch_cast_bf16(const float* in, bf16* out, long long n, cudaStream_t s);
void launch_pack_enc_linear(const float* w, float* out, cudaStream_t s);  // (128, 64*12*12 chw) -> [9216 hwc][128]
// legacy simple U-Net (models/simple_Unet.py), fp32 (simple_kernels.cu)
void launch_su_conv_in(const float* x, const float* w, float* out, int B, int H, int W, int rows, int dim, int lh, int lw, cudaStream_t s);
void launch_su_apply(const float* raw, int ld_in, float* out, int ld_out, const float* stats, const float* gamma, const float* beta,
                     const float* resid, int ld_res, const float* temb, int temb_stride, int temb_mode, const int* step_ptr, int step_off,
                     int B, int HW, int C, cudaStream_t s);
void launch_su_temb(const long long* t_dev, int n_t, const float* table, int max_len, const float* w_cat, const float* b_cat, float* out,
                    int time_dim, int width, cudaStream_t s);
void launch_su_silu(const float* in, float* out, long long n, cudaStream_t s);
void launch_su_bcast(const float* emb, int emb_stride, float* out, int ld_out, int B, int HW, cudaStream_t s);
// ResNet18-GroupNorm vision encoder (resnet.cu): channels-last activations, convs as GEMMs over explicit patch rows
template <typename T> void launch_rn_im2col_img(const float* img, T* out, long long frames_pad, int n_real, cudaStream_t s);
template <typename T> void launch_rn_im2col(const T* in, T* out, long long frames, int H, int W, int C, int k, int stride, int pad, cudaStream_t s);
template <typename T> void launch_rn_gn(const T* x, T* out, float* stats, const float* gamma, const float* beta, const T* resid, int relu, long long frames,
                                        int HW, int C, cudaStream_t s);   // GroupNorm(C / 16) (+ residual) (+ ReLU)
template <typename T> void launch_rn_maxpool(const T* in, T* out, long long frames, int H, int W, int C, cudaStream_t s);
template <typename T> void launch_rn_avgpool(const T* in, float* out, long long frames, int HW, int C, cudaStream_t s);
void launch_rn_pack_conv1(const float* oihw, bf16* w16, float* w32, cudaStream_t s);
long long kernels_launch_count();
void kernels_count_launch();  // one more launch (kernels that live in other translation units)

// tcgen05 path (conv_tc.cu) -----------------------------------------------------------------------
struct TcGemm;  // opaque: tensor maps + geometry for one implicit-GEMM launch
// in: bf16 [Bcap*H*W, Cin] (ld_in); w_packed: bf16 [Cout][taps*Cin]; taps 1 or 9
TcGemm* tc_gemm_create(const bf16* in, int ld_in, const bf16* w_packed, int Cin, int Cout, int taps, int H, int W,
                       int Bcap);
// Pair-folded 3x3 conv with 64 output channels (weights from launch_pack_conv_pfold_bf16): same `in` / geometry arguments as
// tc_gemm_create; runs on the swapped-operand kernel only (EPI_STATS, optionally with the fused GroupNorm apply), B*H*W/2 must
// be a multiple of 256.  tc_gemm_launch takes the real output pointer / leading dimension.  Null if the geometry does not fit.
TcGemm* tc_gemm_create_pfold(const bf16* in, int ld_in, const bf16* w_pfold, int Cin, int H, int W, int Bcap);
bool tc_gemm_is_pfold(const TcGemm* g);
void tc_gemm_destroy(TcGemm* g);
// out bf16 [M, Cout] ld_out; stats partials [B][P][2] with EPI_STATS.  B must be a multiple of tc_batch_multiple.
// returns P, the number of GroupNorm partial slots per sample that EPI_STATS wrote
// EPI_VT (in_proj of an attention block, Cout = 3C): the V third of the output is written transposed to
// vt[row / vt_lk][C][row % vt_lk] for sdpa_tc instead of to `out`.
int tc_gemm_launch(const TcGemm* g, bf16* out, int ld_out, float* stats, const float* bias, const bf16* resid,
                   int ld_res, int flags, int B, cudaStream_t s, bf16* vt = nullptr, int vt_lk = 0,
                   const ApplyArgs* fuse = nullptr, int ksplit = 1, float* partial = nullptr);
// Split-K factor worth using for a 3x3 conv launch of B samples (1 = none).  With ksplit > 1 tc_gemm_launch writes
// fp32 partial tiles [ksplit][B*H*W][Cout] to `partial` instead of `out`; launch_apply_partial sums them.
int tc_gemm_split(const TcGemm* g, int B);
// True when a launch for B samples keeps whole samples and all channels inside one tile, so that GroupNorm apply
// (+GELU, +temb, +FiLM) can run in the conv epilogue (pass `fuse`; `out` then receives the activated map).
bool tc_gemm_can_fuse_apply(const TcGemm* g, int B);
bool tc_gemm_fuse_apply_pays(const TcGemm* g, int B);   // ... and it is measured faster than the separate apply kernel
// Cluster split-K with GroupNorm apply fused behind it (deep levels at small batch): K slices per tile (0 = not
// applicable at this geometry / batch; 1 = the cluster only shares the GroupNorm statistics of a tile's N tiles) and the launch; `out` receives the activated bf16 map, no partial tiles.
int tc_gemm_cluster_split(const TcGemm* g, int B);
int tc_gemm_launch_cluster(const TcGemm* g, bf16* out, int ld_out, const ApplyArgs* ap, int B, int ks, cudaStream_t s);
// Chain of cluster convs (conv_chain_kernel): a run of [3x3 conv + GroupNorm (+GELU / time embedding / FiLM)] layers of one deep
// level (whole samples per 128-row tile) in one launch, optionally with the MaxPool2d(2) (pre_kind 1) / bilinear x2 upsample
// (pre_kind 2) that feeds the first layer; pre_out is the first layer's input buffer (a channel slice of it for the upsample +
// concat).  Returns null when the run is not a small-batch cluster case (the caller then issues the layers one by one).
struct TcChain;
struct TcChainLayerDesc { const TcGemm* g; bf16* out; int ld_out; ApplyArgs ap; };
TcChain* tc_chain_create(const TcChainLayerDesc* layers, int n, int B, int pre_kind, const bf16* pre_in, int pre_ld_in, bf16* pre_out,
                         int pre_ld_out, int pre_C);
void tc_chain_destroy(TcChain* c);
int tc_chain_launch(const TcChain* c, cudaStream_t s, int step_off);   // step_off: ApplyArgs::step_off for every layer of this launch
const char* tc_last_error();
void tc_set_debug(int v);  // microbenchmark switches, see TcParams::dbg
int tc_batch_multiple(int H, int W);  // granularity of B required by the 128-row M tiling at geometry HxW
long long tc_launch_count();

// tcgen05 kind::tf32 3x3 convolution on fp32 activations (conv_tf32.cu): the TF32 precision mode ---------------------------
struct TfGemm;
TfGemm* tf32_conv_create(const float* in, int ld_in, const float* w_packed, int Cin, int Cout, int H, int W, int Bcap);
void tf32_conv_destroy(TfGemm* g);
void tf32_conv_launch(const TfGemm* g, float* out, int ld_out, int B, cudaStream_t s);   // raw fp32 conv output
const char* tf32_last_error();
void launch_pack_conv_tf32(const float* oihw, float* out, int Cout, int Cin, cudaStream_t s);   // -> [Cout][tap][Cin] fp32

// tcgen05 attention core (sdpa_tc.cu) ------------------------------------------------------------------
struct SdpaTc;
bool sdpa_tc_supported(int L, int C, int heads);
int sdpa_tc_keys_per_tile(int L);
// qkv: bf16 [Mcap][3C] (Q | K | unused); vt: bf16 [Mcap / LK][C][LK] written by the in_proj GEMM (EPI_VT)
SdpaTc* sdpa_tc_create(const bf16* qkv, const bf16* vt, int C, int L, int heads, long long Mcap);
void sdpa_tc_destroy(SdpaTc* g);
void sdpa_tc_launch(const SdpaTc* g, bf16* out, long long M, cudaStream_t s);  // out bf16 [M][C], M % 128 == 0

// fused SelfAttention head (attn_head.cu): att = SDPA(in_proj(LayerNorm(x))) for maps of L <= 128 tokens, one launch --------
struct AttnHead;
bool attn_head_supported(int L, int C, int heads);
// out_proj / ff_self weights (bf16 [C][C], K-major) and constants: with them and C == 64 (one CTA owns all channels of a tile) the
// kernel runs the WHOLE SelfAttention block -- attn_head_merges_tail() -- and writes the block output instead of att
struct AttnHeadTail { const bf16 *wo, *w1, *w2; const float *bo, *b1, *b2, *ln_g, *ln_b; };
// w_in_proj: bf16 [3C][C] (K-major), bias fp32 [3C], ln_g / ln_b fp32 [C]
AttnHead* attn_head_create(const bf16* w_in_proj, const float* bias, const float* ln_g, const float* ln_b, int C, int L, int heads,
                           const AttnHeadTail* tail = nullptr);
bool attn_head_merges_tail(const AttnHead* g);
void attn_head_destroy(AttnHead* g);
// M % 128 == 0 (256 when L == 256); out = att [M][C], or with a merged tail y = block output [M][ld_y]
void attn_head_launch(const AttnHead* g, const bf16* x, int ld_x, bf16* out, long long M, cudaStream_t s, bf16* y = nullptr, int ld_y = 0);

// fused SelfAttention tail (attn_tc.cu): out = FF(LN(out_proj(att) + x)) + (out_proj(att) + x) ------------------
struct AttnTail;
bool attn_tail_supported(int C);
AttnTail* attn_tail_create(const bf16* att, long long Mcap, int C, const bf16* wo, const bf16* w1, const bf16* w2, const float* bo,
                           const float* b1, const float* b2, const float* ln_g, const float* ln_b);
void attn_tail_destroy(AttnTail* g);
void attn_tail_launch(const AttnTail* g, const bf16* x, int ld_x, bf16* out, int ld_out, long long M, cudaStream_t s);
