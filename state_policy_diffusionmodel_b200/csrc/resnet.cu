// resnet.cu — kernels of the ResNet18-GroupNorm vision encoder (reference models/Unet_FiLmLayer.py:316-386: torchvision resnet18,
// fc = Identity, every BatchNorm2d replaced by GroupNorm(C / 16, C)); SURVEY.md 8(f) row 1, north_star item (1).  sm_100a.
//
// Activations are channels-last [n][H][W][C] (bf16 on the tensor-core plans, fp32 on the parity plan).  Every convolution is a
// GEMM over patch rows: an im2col kernel writes [n*Ho*Wo][kh*kw*Cin] (K order = (kh, kw, c), zero fill for the padding ring), the
// contraction itself is the flat tcgen05 GEMM of conv_tc.cu (bf16: TMA-fed, TMEM accumulators) or the fp32 CUDA-core GEMM.  The
// spatial sizes of the network (48, 24, 12, 6, 3) do not divide the 128-row tiles of the implicit-GEMM kernel, hence the explicit
// patch matrix; it costs ~9 MB of extra HBM traffic per frame against 0.67 GFLOP, i.e. the encoder stays tensor-bound.
//   conv1 7x7 s2 p3 (3 -> 64; K = 147 padded to 192)  -> GroupNorm(4) + ReLU -> MaxPool 3x3 s2 p1
//   layer1..4: BasicBlock x2 (3x3 convs, stride 2 + 1x1 stride-2 downsample at the start of layers 2-4), GroupNorm + ReLU,
//   residual add before the last ReLU -> global average pool -> 512 features.
#include "common.cuh"

static inline int cdiv_r(long long a, long long b) { return (int)((a + b - 1) / b); }

namespace {

template <typename T> __device__ __forceinline__ void zero8(T* p);
template <> __device__ __forceinline__ void zero8<float>(float* p) {
  *reinterpret_cast<float4*>(p) = make_float4(0.f, 0.f, 0.f, 0.f);
  *reinterpret_cast<float4*>(p + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
}
template <> __device__ __forceinline__ void zero8<bf16>(bf16* p) { *reinterpret_cast<uint4*>(p) = make_uint4(0u, 0u, 0u, 0u); }
template <typename T> __device__ __forceinline__ void copy8(T* dst, const T* src);
template <> __device__ __forceinline__ void copy8<float>(float* dst, const float* src) {
  *reinterpret_cast<float4*>(dst) = *reinterpret_cast<const float4*>(src);
  *reinterpret_cast<float4*>(dst + 4) = *reinterpret_cast<const float4*>(src + 4);
}
template <> __device__ __forceinline__ void copy8<bf16>(bf16* dst, const bf16* src) { *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src); }

// conv1 patch rows from the fp32 NCHW frames: out[(f*48 + ho)*48 + wo][(kh*7 + kw)*3 + c], k >= 147 zero.  One thread per 8 K entries.
template <typename T>
__global__ void __launch_bounds__(256) rn_im2col_img_kernel(const float* __restrict__ img, T* __restrict__ out, long long total_vec, int n_real) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= total_vec) return;
  const int k8 = (int)(v % 24) * 8;
  const long long row = v / 24;
  const int wo = (int)(row % 48), ho = (int)((row / 48) % 48);
  const long long f = row / 2304;
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = k8 + i;
    float val = 0.f;
    if (k < 147 && f < n_real) {
      const int tap = k / 3, c = k - tap * 3;
      const int kh = tap / 7, kw = tap - kh * 7;
      const int h = 2 * ho - 3 + kh, w = 2 * wo - 3 + kw;
      if (h >= 0 && h < 96 && w >= 0 && w < 96) val = __ldg(img + ((f * 3 + c) * 96 + h) * 96 + w);
    }
    x[i] = val;
  }
  store8(out + row * 192 + k8, x);
}

// generic NHWC im2col: out[(f*Ho + ho)*Wo + wo][(kh*k + kw)*C + c]; one thread per 8 channels of one (row, tap)
template <typename T>
__global__ void __launch_bounds__(256) rn_im2col_kernel(const T* __restrict__ in, T* __restrict__ out, long long total_vec, int H, int W, int C, int k,
                                                        int stride, int pad, int Ho, int Wo) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= total_vec) return;
  const int cv = C >> 3;
  const int c8 = (int)(v % cv) << 3;
  long long t = v / cv;
  const int tap = (int)(t % (k * k));
  const long long row = t / (k * k);
  const int wo = (int)(row % Wo), ho = (int)((row / Wo) % Ho);
  const long long f = row / ((long long)Ho * Wo);
  const int kh = tap / k, kw = tap - kh * k;
  const int h = ho * stride - pad + kh, w = wo * stride - pad + kw;
  T* dst = out + row * ((long long)k * k * C) + (long long)tap * C + c8;
  if (h >= 0 && h < H && w >= 0 && w < W) copy8<T>(dst, in + ((f * H + h) * W + w) * C + c8);
  else zero8<T>(dst);
}

// GroupNorm(C / 16, C) statistics: one block per frame; stats[f][g] = (mean, rstd)
template <typename T>
__global__ void __launch_bounds__(256) rn_gn_stats_kernel(const T* __restrict__ x, float* __restrict__ stats, int HW, int C, float eps) {
  __shared__ float s_sum[32], s_sq[32];
  const int f = blockIdx.x, tid = threadIdx.x;
  const int G = C >> 4;
  if (tid < 32) { s_sum[tid] = 0.f; s_sq[tid] = 0.f; }
  __syncthreads();
  const int cv = C >> 3;                 // 8..64: divides 256, so a thread always sees the same 8 channels
  const int c8 = (tid % cv) << 3;
  float s = 0.f, q = 0.f;
  const T* base = x + (size_t)f * HW * C;
  for (int v = tid; v < HW * cv; v += 256) {
    float t8[8];
    load8(base + (size_t)(v / cv) * C + c8, t8);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += t8[i]; q = fmaf(t8[i], t8[i], q); }
  }
  atomicAdd(&s_sum[c8 >> 4], s);
  atomicAdd(&s_sq[c8 >> 4], q);
  __syncthreads();
  if (tid < G) {
    const float inv_n = 1.0f / (float)(HW * 16);
    const float mean = s_sum[tid] * inv_n;
    const float var = fmaxf(s_sq[tid] * inv_n - mean * mean, 0.f);
    stats[((size_t)f * 32 + tid) * 2] = mean;
    stats[((size_t)f * 32 + tid) * 2 + 1] = rsqrtf(var + eps);
  }
}

// y = gn(x) * gamma + beta (+ resid) (ReLU); in place allowed
template <typename T>
__global__ void __launch_bounds__(256) rn_gn_apply_kernel(const T* __restrict__ x, T* __restrict__ out, const float* __restrict__ stats,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta, const T* __restrict__ resid,
                                                          int relu, long long total_vec, int HW, int C) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= total_vec) return;
  const int cv = C >> 3;
  const int c8 = (int)(v % cv) << 3;
  const long long row = v / cv;
  const long long f = row / HW;
  const float mean = stats[((size_t)f * 32 + (c8 >> 4)) * 2], rstd = stats[((size_t)f * 32 + (c8 >> 4)) * 2 + 1];
  float t8[8], g8[8], b8[8];
  load8(x + row * C + c8, t8);
  load8(gamma + c8, g8);
  load8(beta + c8, b8);
#pragma unroll
  for (int i = 0; i < 8; ++i) t8[i] = (t8[i] - mean) * rstd * g8[i] + b8[i];
  if (resid) {
    float r8[8];
    load8(resid + row * C + c8, r8);
#pragma unroll
    for (int i = 0; i < 8; ++i) t8[i] += r8[i];
  }
  if (relu) {
#pragma unroll
    for (int i = 0; i < 8; ++i) t8[i] = fmaxf(t8[i], 0.f);
  }
  store8(out + row * C + c8, t8);
}

// MaxPool2d(kernel 3, stride 2, padding 1)
template <typename T>
__global__ void __launch_bounds__(256) rn_maxpool_kernel(const T* __restrict__ in, T* __restrict__ out, long long total_vec, int H, int W, int C, int Ho,
                                                         int Wo) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= total_vec) return;
  const int cv = C >> 3;
  const int c8 = (int)(v % cv) << 3;
  const long long row = v / cv;
  const int wo = (int)(row % Wo), ho = (int)((row / Wo) % Ho);
  const long long f = row / ((long long)Ho * Wo);
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw) {
      const int h = 2 * ho - 1 + kh, w = 2 * wo - 1 + kw;
      if (h < 0 || h >= H || w < 0 || w >= W) continue;
      float t8[8];
      load8(in + ((f * H + h) * W + w) * C + c8, t8);
#pragma unroll
      for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], t8[i]);
    }
  store8(out + row * C + c8, m);
}

// AdaptiveAvgPool2d(1) + flatten: out[f][c] fp32
template <typename T>
__global__ void __launch_bounds__(256) rn_avgpool_kernel(const T* __restrict__ in, float* __restrict__ out, long long total, int HW, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long long f = i / C;
  float s = 0.f;
  for (int p = 0; p < HW; ++p) s += to_f32<T>(in[(f * HW + p) * C + c]);
  out[i] = s / (float)HW;
}

// conv1 weights (64, 3, 7, 7) -> K-major [64][192] (tensor-core B operand) / [192][64] (CUDA-core GEMM), K = (kh*7 + kw)*3 + c
__global__ void rn_pack_conv1_kernel(const float* __restrict__ oihw, bf16* __restrict__ w16, float* __restrict__ w32) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 192) return;
  const int k = i % 192, co = i / 192;
  float v = 0.f;
  if (k < 147) {
    const int tap = k / 3, c = k - tap * 3;
    v = oihw[(co * 3 + c) * 49 + tap];
  }
  if (w16) w16[co * 192 + k] = __float2bfloat16_rn(v);
  if (w32) w32[k * 64 + co] = v;
}
}  // namespace

template <typename T> void launch_rn_im2col_img(const float* img, T* out, long long frames_pad, int n_real, cudaStream_t s) {
  const long long total = frames_pad * 2304 * 24;
  rn_im2col_img_kernel<T><<<cdiv_r(total, 256), 256, 0, s>>>(img, out, total, n_real);
  kernels_count_launch();
}
template <typename T> void launch_rn_im2col(const T* in, T* out, long long frames, int H, int W, int C, int k, int stride, int pad, cudaStream_t s) {
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const long long total = frames * Ho * Wo * k * k * (C >> 3);
  rn_im2col_kernel<T><<<cdiv_r(total, 256), 256, 0, s>>>(in, out, total, H, W, C, k, stride, pad, Ho, Wo);
  kernels_count_launch();
}
template <typename T> void launch_rn_gn(const T* x, T* out, float* stats, const float* gamma, const float* beta, const T* resid, int relu, long long frames,
                                        int HW, int C, cudaStream_t s) {
  rn_gn_stats_kernel<T><<<(unsigned)frames, 256, 0, s>>>(x, stats, HW, C, 1e-5f);
  const long long total = frames * HW * (C >> 3);
  rn_gn_apply_kernel<T><<<cdiv_r(total, 256), 256, 0, s>>>(x, out, stats, gamma, beta, resid, relu, total, HW, C);
  kernels_count_launch();
  kernels_count_launch();
}
template <typename T> void launch_rn_maxpool(const T* in, T* out, long long frames, int H, int W, int C, cudaStream_t s) {
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long total = frames * Ho * Wo * (C >> 3);
  rn_maxpool_kernel<T><<<cdiv_r(total, 256), 256, 0, s>>>(in, out, total, H, W, C, Ho, Wo);
  kernels_count_launch();
}
template <typename T> void launch_rn_avgpool(const T* in, float* out, long long frames, int HW, int C, cudaStream_t s) {
  const long long total = frames * C;
  rn_avgpool_kernel<T><<<cdiv_r(total, 256), 256, 0, s>>>(in, out, total, HW, C);
  kernels_count_launch();
}
void launch_rn_pack_conv1(const float* oihw, bf16* w16, float* w32, cudaStream_t s) {
  rn_pack_conv1_kernel<<<cdiv_r(64 * 192, 256), 256, 0, s>>>(oihw, w16, w32);
}

#define RN_INST(T)                                                                                                                             \
  template void launch_rn_im2col_img<T>(const float*, T*, long long, int, cudaStream_t);                                                       \
  template void launch_rn_im2col<T>(const T*, T*, long long, int, int, int, int, int, int, cudaStream_t);                                      \
  template void launch_rn_gn<T>(const T*, T*, float*, const float*, const float*, const T*, int, long long, int, int, cudaStream_t);           \
  template void launch_rn_maxpool<T>(const T*, T*, long long, int, int, int, cudaStream_t);                                                    \
  template void launch_rn_avgpool<T>(const T*, float*, long long, int, int, cudaStream_t);
RN_INST(float)
RN_INST(bf16)
