// common.cuh — shared types for the spdm CUDA sources (sm_100a only).
//
// Activation layout ("HBWC"): every feature map of the U-Net lives in HBM as
//   act[h][b][w][c]   (c fastest),  b in [0,Bp), Bp = batch padded to a multiple of the largest
//   batch tile.  Row index of a pixel:  r = (h*Bp + b)*W + w ; a feature map is an [M=H*Bp*W, C]
//   row-major matrix — the token-major (B,L,C) view SelfAttention needs
//   (models/Unet_FiLmLayer.py:74) and the K-major A operand of the implicit GEMM.
// H is outermost so that one TMA box {C=64, W, Bt, Ht+2} (nested strides) fetches the halo'd
// tile of a 3x3 convolution: the +-1 row taps become 1024-byte-aligned offsets into that tile.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

typedef __nv_bfloat16 bf16;

// A conv M-tile is Ht rows of h  x  Bt samples  x  all W columns = 128 pixels.
struct TileGeom {
  int H, W, Bp;
  int Ht, Bt;
};

__host__ __device__ inline int geom_tiles_m(const TileGeom& g) { return (g.H / g.Ht) * (g.Bp / g.Bt); }

// tile-local row r_t (0..127) -> global HBWC row
__device__ __forceinline__ int tile_row_global(const TileGeom& g, int h0, int b0, int r_t) {
  const int bw = g.Bt * g.W;
  const int hh = r_t / bw;
  const int rem = r_t - hh * bw;
  return ((h0 + hh) * g.Bp + b0) * g.W + rem;
}

enum : int {
  EPI_STATS = 1,   // conv: write raw output + per-sample (sum, sumsq) partials for GroupNorm(1,C)
  EPI_BIAS = 2,
  EPI_GELU = 4,
  EPI_RELU = 8,
  EPI_RESID = 16,
};

enum : int { ACT_NONE = 0, ACT_GELU = 1 };
enum : int { TEMB_NONE = 0, TEMB_ROW0 = 1, TEMB_PER_SAMPLE = 2, TEMB_STEP = 3 };

__device__ __forceinline__ float gelu_exact(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// 4-wide vector load/store of activations (C is always a multiple of 4)
__device__ __forceinline__ void load4(const float* p, float v[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const bf16* p, float v[4]) {
  const uint2 t = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
__device__ __forceinline__ void store4(float* p, const float v[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(bf16* p, const float v[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

// ---------------------------------------------------------------------------------------------
// Host-side launch descriptors (plain structs passed by value to kernels)
// ---------------------------------------------------------------------------------------------
struct GemmSimtArgs {
  const float* in;      // [M, Cin] fp32 activations (HBWC rows)
  const float* w;       // [taps][Cin][Cout] fp32
  const float* bias;    // [Cout] or null
  const void* resid;    // [M, Cout] (TOut) or null
  void* out;            // [M, Cout] TOut
  float* stats;         // [Bp][P][2] partials (EPI_STATS)
  int Cin, Cout, M;
  TileGeom g;           // used when taps == 9
  int taps;             // 1 (flat GEMM) or 9 (3x3, pad 1)
  int flags;
  int P;                // partial slots per sample
  int remap;            // 0 none, 1 = encoder conv2 output patch-major remap
};

struct ApplyArgs {
  const void* raw;      // [M, C] raw conv output
  void* out;            // [M, C]
  const float* stats;   // [Bp][P][2]
  int P;
  const float* gamma;
  const float* beta;
  const float* temb;    // [rows][temb_stride] or null
  int temb_mode, temb_stride, temb_off;
  const int* step_base; // device int: schedule index base (TEMB_STEP)
  int step_local;
  const float* film;    // [Bp][film_stride] or null; scale at film_off, bias at film_off + C
  int film_stride, film_off;
  int H, W, Bp, C;
  int act;
  float inv_n, eps;
};

// launchers implemented in kernels.cu -----------------------------------------------------------
template <typename TOut> void launch_gemm_simt(const GemmSimtArgs& a, cudaStream_t s);
template <typename T> void launch_apply(const ApplyArgs& a, cudaStream_t s);
template <typename T> void launch_pool(const T* in, T* out, int Ho, int Wo, int Bp, int C, cudaStream_t s);
template <typename T> void launch_upcat(const T* low, const T* skip, T* out, int Hi, int Wi, int Bp,
                                        int C1, int C2, cudaStream_t s);
template <typename T> void launch_layernorm(const T* in, T* out, const float* g, const float* b, int M, int C,
                                            cudaStream_t s);
template <typename T> void launch_sdpa(const T* qkv, T* out, int H, int W, int Bp, int C, int heads,
                                       cudaStream_t s);
template <typename T> void launch_outc(const T* x, const float* w, const float* bias, float* eps, int B, int Bp,
                                       int H, int W, int C, int rows, int dim, int lh, int lw, cudaStream_t s);
void launch_pad_input(const float* x, float* out, int B, int Bp, int H, int W, int rows, int dim, int lh, int lw,
                      cudaStream_t s);
void launch_step(const float* x, const float* eps, const float* noise, const float* inpaint, float* x_out,
                 float* history, const float* coef, const int* step_base, int step_local, int n_per_sample,
                 int inpaint_elems, int B, uint64_t seed, int use_philox, cudaStream_t s);
void launch_advance(int* step_base, int delta, cudaStream_t s);
void launch_temb(const int64_t* t_dev, const int* t_list_i32, int n_t, const float* inv_freq, const float* w_t,
                 const float* bias, float* out, int time_dim, int n_out, cudaStream_t s);
void launch_mish(const float* in, float* out, int64_t n, cudaStream_t s);
void launch_enc_conv1(const float* img, const float* w, const float* bias, float* out, int n, cudaStream_t s);
void launch_build_cond(const float* pos, const float* act, const float* vel, const float* feat, float* cond,
                       int B, int T, int cond_dim, cudaStream_t s);
void launch_add_noise(const float* x0, const float* noise, const int64_t* t, const float* sa, const float* sb,
                      const float* inpaint, float* out, int n_per_sample, int inpaint_elems, int B,
                      cudaStream_t s);
// weight repack
void launch_pack_conv_f32(const float* oihw, float* out, int Cout, int Cin, int k, cudaStream_t s);   // -> [tap][Cin][Cout]
void launch_pack_conv_bf16(const float* oihw, bf16* out, int Cout, int Cin, int k, cudaStream_t s);   // -> [Cout][tap][Cin]
void launch_pack_linear_f32(const float* nk, float* out, int N, int K, cudaStream_t s);               // -> [K][N]
void launch_cast_bf16(const float* in, bf16* out, int64_t n, cudaStream_t s);
void launch_pack_enc_linear(const float* w, float* out, cudaStream_t s);  // (128, 64*12*12 chw) -> [9216 hwc][128]
template <typename T> void launch_to_nchw(const T* hbwc, float* out, int B, int Bp, int H, int W, int C,
                                          cudaStream_t s);
void launch_from_f32(const float* in, bf16* out, int64_t n, cudaStream_t s);

// tcgen05 path (conv_tc.cu) ----------------------------------------------------------------------
struct TcLayer;  // opaque: tensor maps + geometry for one implicit-GEMM launch
TcLayer* tc_layer_create(const bf16* in, const bf16* w_packed, int Cin, int Cout, int taps, const TileGeom& g,
                         int M_flat, const bf16* in2, int Cin2);
void tc_layer_destroy(TcLayer* l);
// raw conv: out bf16 [M,Cout] + stats partials; linear: bias/gelu/resid epilogue
void tc_layer_launch(const TcLayer* l, bf16* out, float* stats, const float* bias, const bf16* resid, int flags,
                     cudaStream_t s);
int tc_layer_partials(const TcLayer* l);  // P (partial slots per sample) written by EPI_STATS
int simt_partials(const TileGeom& g, int Cout);
const char* tc_last_error();

TileGeom choose_geom(int H, int W, int Bp);
int geom_max_bt(int H0, int W0);
