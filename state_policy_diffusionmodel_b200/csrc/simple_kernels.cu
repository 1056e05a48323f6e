// simple_kernels.cu — CUDA-core kernels of the legacy `UNet` of models/simple_Unet.py:260-339 (the `model='UNet'` default of
// Diffusion_DDPM, models/diffusion_ddpm.py:60-62), fp32, sm_100a.
//
// The network reuses the fp32 implicit-GEMM (kernels.cu::gemm_simt_kernel), GroupNorm statistics, max-pool and bilinear
// upsample kernels of the FiLM U-Net; what differs lives here:
//   * channel counts that are multiples of 16 but not of 64 (16, 160, 288, 448, 224, 96, 112): a GroupNorm apply whose
//     thread <-> channel mapping does not assume a power-of-two channel count;
//   * DoubleConvolution ends with GELU, optionally on `norm(x) + x_res` (residual=True, :104-125);
//   * the time embedding is a table lookup (PositionalEncoding, :214-242: sin / cos interleaved) -> SiLU -> Linear;
//   * conditioning is a 32-channel broadcast map concatenated after every stage (:160-166, :203-209).
#include "common.cuh"

static inline int cdiv_s(long long a, long long b) { return (int)((a + b - 1) / b); }

namespace {

// input_conv.first: Conv2d(1 -> 16, 3x3, pad 1, no bias) on the zero-padded (pad_to 8, :14-33) sample; one thread per pixel.
__global__ void __launch_bounds__(256) su_conv_in_kernel(const float* __restrict__ x, const float* __restrict__ w /*[9][16]*/,
                                                         float* __restrict__ out, long long total_px, int H, int W, int rows, int dim,
                                                         int lh, int lw) {
  pdl_wait();
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_px) return;
  const int px = (int)(i % (H * W));
  const long long b = i / (H * W);
  const int hh = px / W, ww = px - hh * W;
  float acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = 0.f;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int sh = hh + tap / 3 - 1 - lh, sw = ww + tap % 3 - 1 - lw;
    if (sh < 0 || sh >= rows || sw < 0 || sw >= dim) continue;
    const float xv = x[((size_t)b * rows + sh) * dim + sw];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = fmaf(xv, __ldg(w + tap * 16 + c), acc[c]);
  }
  store8(out + i * 16, acc);
  store8(out + i * 16 + 8, acc + 8);
}

// GroupNorm(1, C) apply of the simple U-Net:  y = norm(raw) * gamma + beta ; y += resid (DoubleConvolution(residual=True)) ;
// y = gelu(y) (every DoubleConvolution ends with it, :123-125; also the GELU between its two convs) ; y += temb[row][c] (:152-155).
// One thread per 4 channels of a pixel; statistics are the single (sum, sumsq) partial written by stats_kernel.
__global__ void __launch_bounds__(256) su_apply_kernel(const float* __restrict__ raw, int ld_in, float* __restrict__ out, int ld_out,
                                                       const float* __restrict__ stats, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, const float* __restrict__ resid, int ld_res,
                                                       const float* __restrict__ temb, int temb_stride, int temb_mode, const int* step_ptr,
                                                       int step_off, int HW, int C, long long total4, float eps) {
  pdl_wait();
  pdl_trigger();
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= total4) return;
  const int c4 = C >> 2;
  const long long row = v / c4;
  const int c = (int)(v - row * c4) << 2;
  const long long b = row / HW;
  const double n = (double)HW * (double)C;
  const double s = (double)stats[2 * b], q = (double)stats[2 * b + 1];
  const double dmean = s / n;
  double var = q / n - dmean * dmean;
  if (var < 0.0) var = 0.0;
  const float mean = (float)dmean, rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float4 x = *reinterpret_cast<const float4*>(raw + row * ld_in + c);
  const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
  const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
  float y[4] = {(x.x - mean) * rstd * g.x + be.x, (x.y - mean) * rstd * g.y + be.y, (x.z - mean) * rstd * g.z + be.z,
                (x.w - mean) * rstd * g.w + be.w};
  if (resid) {
    const float4 r = *reinterpret_cast<const float4*>(resid + row * ld_res + c);
    y[0] += r.x; y[1] += r.y; y[2] += r.z; y[3] += r.w;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) y[i] = gelu_exact(y[i]);
  if (temb_mode != TEMB_NONE) {
    long long trow = 0;
    if (temb_mode == TEMB_PER_SAMPLE) trow = b;
    else if (temb_mode == TEMB_STEP) trow = *step_ptr + step_off;
    const float4 t = __ldg(reinterpret_cast<const float4*>(temb + trow * temb_stride + c));
    y[0] += t.x; y[1] += t.y; y[2] += t.z; y[3] += t.w;
  }
  *reinterpret_cast<float4*>(out + row * ld_out + c) = make_float4(y[0], y[1], y[2], y[3]);
}

// PositionalEncoding lookup (eval mode: no dropout) -> SiLU -> the six emb_layer Linears at once:
//   out[row][n] = silu(table[t_row]) @ w_cat[256][width] + b_cat     (:136-142, :184-190, :238-242)
__global__ void __launch_bounds__(256) su_temb_kernel(const long long* __restrict__ t_dev, const float* __restrict__ table, int max_len,
                                                      const float* __restrict__ w_cat, const float* __restrict__ b_cat,
                                                      float* __restrict__ out, int time_dim, int width) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float pe[];
  const int row = blockIdx.x;
  long long t = t_dev[row];
  if (t < 0) t = 0;
  if (t >= max_len) t = max_len - 1;
  for (int i = threadIdx.x; i < time_dim; i += blockDim.x) {
    const float v = table[(size_t)t * time_dim + i];
    pe[i] = v / (1.f + expf(-v));
  }
  __syncthreads();
  for (int n = threadIdx.x; n < width; n += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < time_dim; ++k) acc = fmaf(pe[k], __ldg(w_cat + (size_t)k * width + n), acc);
    out[(size_t)row * width + n] = acc + __ldg(b_cat + n);
  }
}

__global__ void su_silu_kernel(const float* __restrict__ in, float* __restrict__ out, long long n) {
  pdl_wait();
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = in[i];
  out[i] = x / (1.f + expf(-x));
}

// cond_emb (B, 32) of one stage repeated over every pixel into channels [0, 32) of `out` (already offset to the slot) (:160-166)
__global__ void su_bcast_kernel(const float* __restrict__ emb, int emb_stride, float* __restrict__ out, int ld_out, int HW, long long total) {
  pdl_wait();
  pdl_trigger();
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one float4 of 8 per pixel
  if (v >= total) return;
  const long long row = v >> 3;
  const int c = (int)(v & 7) << 2;
  const long long b = row / HW;
  *reinterpret_cast<float4*>(out + row * ld_out + c) = __ldg(reinterpret_cast<const float4*>(emb + b * emb_stride + c));
}
}  // namespace

void launch_su_conv_in(const float* x, const float* w, float* out, int B, int H, int W, int rows, int dim, int lh, int lw, cudaStream_t s) {
  const long long total = (long long)B * H * W;
  launch_pdl(su_conv_in_kernel, dim3(cdiv_s(total, 256)), dim3(256), 0, s, x, w, out, total, H, W, rows, dim, lh, lw);
  kernels_count_launch();
}
void launch_su_apply(const float* raw, int ld_in, float* out, int ld_out, const float* stats, const float* gamma, const float* beta,
                     const float* resid, int ld_res, const float* temb, int temb_stride, int temb_mode, const int* step_ptr, int step_off,
                     int B, int HW, int C, cudaStream_t s) {
  const long long total4 = (long long)B * HW * (C >> 2);
  launch_pdl(su_apply_kernel, dim3(cdiv_s(total4, 256)), dim3(256), 0, s, raw, ld_in, out, ld_out, stats, gamma, beta, resid, ld_res, temb,
             temb_stride, temb_mode, step_ptr, step_off, HW, C, total4, 1e-5f);
  kernels_count_launch();
}
void launch_su_temb(const long long* t_dev, int n_t, const float* table, int max_len, const float* w_cat, const float* b_cat, float* out,
                    int time_dim, int width, cudaStream_t s) {
  launch_pdl(su_temb_kernel, dim3(n_t), dim3(256), time_dim * sizeof(float), s, t_dev, table, max_len, w_cat, b_cat, out, time_dim, width);
  kernels_count_launch();
}
void launch_su_silu(const float* in, float* out, long long n, cudaStream_t s) {
  launch_pdl(su_silu_kernel, dim3(cdiv_s(n, 256)), dim3(256), 0, s, in, out, n);
  kernels_count_launch();
}
void launch_su_bcast(const float* emb, int emb_stride, float* out, int ld_out, int B, int HW, cudaStream_t s) {
  const long long total = (long long)B * HW * 8;
  launch_pdl(su_bcast_kernel, dim3(cdiv_s(total, 256)), dim3(256), 0, s, emb, emb_stride, out, ld_out, HW, total);
  kernels_count_launch();
}
