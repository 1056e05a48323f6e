// bwd_kernels.cu — CUDA-core kernels of the training step (sm_100a): everything of `loss.backward()` and the
// optimizer that is not a dense bf16 contraction.  GroupNorm / GELU / FiLM / time-embedding backward, max-pool and
// bilinear-upsample backward, LayerNorm and attention-core backward, the MSE loss + outc backward, the fp32 weight-
// gradient GEMM of the parity path, the vision-encoder backward and fused clip + Adam.
// Reference lines are cited per kernel (paths relative to the reference repo root).
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include "train.cuh"

static long long g_bwd_launches = 0;
long long bwd_launch_count() { return g_bwd_launches; }
#define COUNT_LAUNCH() (++g_bwd_launches)
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

namespace {
__device__ __forceinline__ float gelu_grad(float u) {  // d/du [0.5 u (1 + erf(u / sqrt 2))]
  return 0.5f * (1.0f + erff(u * 0.70710678118654752440f)) + u * 0.3989422804014327f * __expf(-0.5f * u * u);
}
// Derivative of the GELU the bf16 forward actually applies after GroupNorm (kernels.cu::gelu_tanh_fast / conv_tc.cu::gelu_fast_tc:
// 0.5 y (1 + tanh(k (y + c y^3)))): one MUFU + ~10 FMA-class instructions instead of the erf polynomial + exp (the GELU launches of
// the GroupNorm backward are issue-bound: ncu issue-active 42-48 %).  The fp32 parity path keeps the exact pair.
template <typename T> __device__ __forceinline__ float gelu_grad_as_forward(float u) {
  if constexpr (sizeof(T) == 4) {
    return gelu_grad(u);
  } else {
    const float u2 = u * u;
    const float inner = 0.7978845608028654f * u * fmaf(0.044715f, u2, 1.0f);
    float th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(inner));
    const float dinner = fmaf(0.7978845608028654f * 3.0f * 0.044715f, u2, 0.7978845608028654f);
    return fmaf(0.5f * u * dinner, fmaf(-th, th, 1.0f), fmaf(0.5f, th, 0.5f));
  }
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum over the 256 threads of a block, result broadcast to every thread (scratch: 8 floats)
__device__ __forceinline__ float block_sum_256(float v, float* scratch) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += scratch[i];
  return t;
}
}  // namespace

// =================================================================================================
// GroupNorm(1, C) backward (models/Unet_FiLmLayer.py:105,112-115) fused with what follows it in the forward:
//   first conv of a DoubleConvolution:  out = gelu(gn(raw))                              (:113)
//   last conv of a Down/Up stage:       out = scale * (gn(raw) + temb) + bias             (:165-177)
// One block per sample (the normalisation group is the whole sample):
//   pass 1: g = d out / d gn ; per-channel sums (d gamma, d beta, d FiLM scale/bias, d temb), per-sample
//           sums s1 = sum g*gamma, s2 = sum g*gamma*xhat
//   pass 2: dx = rstd * (g*gamma - s1/N - xhat * s2/N)
// =================================================================================================
namespace {
template <typename T>
__global__ void __launch_bounds__(256) gn_bwd_kernel(GnBwdArgs a) {
  __shared__ float red[256 * 8];
  __shared__ float scratch[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  double s = 0.0, q = 0.0;
  for (int p = 0; p < a.P; ++p) {
    const float2 sq = __ldg(reinterpret_cast<const float2*>(a.stats) + (size_t)b * a.P + p);
    s += (double)sq.x;
    q += (double)sq.y;
  }
  const double n = (double)a.HW * (double)a.C;
  const double dmean = s / n;
  double var = q / n - dmean * dmean;
  if (var < 0.0) var = 0.0;
  const float mean = (float)dmean;
  const float rstd = (float)(1.0 / sqrt(var + (double)a.eps));

  const T* __restrict__ dy = reinterpret_cast<const T*>(a.dy);
  const T* __restrict__ raw = reinterpret_cast<const T*>(a.raw);
  T* __restrict__ dx = reinterpret_cast<T*>(a.dx);
  const int vpr = a.C >> 3;           // 8-channel vectors per row: 8..64, divides 256
  const int lane_c = tid % vpr;       // this thread always handles the same 8 channels
  const int row0 = tid / vpr, row_step = 256 / vpr;
  const int c8 = lane_c << 3;
  float g[8], be[8], te[8], fs[8];
  load8(a.gamma + c8, g);
  load8(a.beta + c8, be);
#pragma unroll
  for (int i = 0; i < 8; ++i) { te[i] = 0.f; fs[i] = 1.f; }
  const bool has_film = a.film != nullptr, has_temb = a.temb != nullptr;
  if (has_temb) load8(a.temb + (size_t)b * SPDM_TEMB_WIDTH + a.temb_off + c8, te);
  if (has_film) load8(a.film + (size_t)b * SPDM_FILM_WIDTH + a.film_off + c8, fs);

  float acc_dg[8], acc_db[8], acc_fb[8], acc_fs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc_dg[i] = 0.f; acc_db[i] = 0.f; acc_fb[i] = 0.f; acc_fs[i] = 0.f; }
  float s1 = 0.f, s2 = 0.f;
  for (int row = row0; row < a.HW; row += row_step) {
    float d[8], x[8];
    load8(dy + ((size_t)b * a.HW + row) * a.ld_dy + c8, d);
    load8(raw + ((size_t)b * a.HW + row) * a.ld_raw + c8, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = (x[i] - mean) * rstd;
      const float u = xh * g[i] + be[i];
      float go = d[i];
      if (has_film) { acc_fb[i] += d[i]; acc_fs[i] = fmaf(d[i], u + te[i], acc_fs[i]); go = d[i] * fs[i]; }
      else if (has_temb) acc_fb[i] += d[i];
      if (a.act == ACT_GELU) go *= gelu_grad_as_forward<T>(u);
      acc_dg[i] = fmaf(go, xh, acc_dg[i]);
      acc_db[i] += go;
      const float gg = go * g[i];
      s1 += gg;
      s2 = fmaf(gg, xh, s2);
    }
  }
  s1 = block_sum_256(s1, scratch);
  s2 = block_sum_256(s2, scratch);
  const float inv_n = 1.0f / ((float)a.HW * (float)a.C);
  const float m1 = s1 * inv_n, m2 = s2 * inv_n;

  // per-channel reductions across the row groups: one quantity at a time through shared memory
  auto reduce_channels = [&](const float (&acc)[8], float (&out)[8]) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) red[tid * 8 + i] = acc[i];
    __syncthreads();
    if (tid < vpr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) out[i] = 0.f;
      for (int r = 0; r < row_step; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) out[i] += red[(r * vpr + tid) * 8 + i];
    }
  };
  float tot[8];
  reduce_channels(acc_dg, tot);
  if (tid < vpr)
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(a.dgamma + c8 + i, tot[i]);
  reduce_channels(acc_db, tot);
  if (tid < vpr) {
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(a.dbeta + c8 + i, tot[i]);
  }
  if (has_film || has_temb) {
    reduce_channels(acc_fb, tot);  // sum_hw dy
    if (tid < vpr) {
      if (has_film) store8(a.d_film + (size_t)b * SPDM_FILM_WIDTH + a.film_off + a.C + c8, tot);
      if (has_temb) {
        float dt[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) dt[i] = tot[i] * fs[i];  // d temb = scale * sum_hw dy
        store8(a.d_temb + (size_t)b * SPDM_TEMB_WIDTH + a.temb_off + c8, dt);
      }
    }
    if (has_film) {
      reduce_channels(acc_fs, tot);
      if (tid < vpr) store8(a.d_film + (size_t)b * SPDM_FILM_WIDTH + a.film_off + c8, tot);
    }
  }
  // pass 2
  for (int row = row0; row < a.HW; row += row_step) {
    float d[8], x[8];
    load8(dy + ((size_t)b * a.HW + row) * a.ld_dy + c8, d);
    load8(raw + ((size_t)b * a.HW + row) * a.ld_raw + c8, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = (x[i] - mean) * rstd;
      float go = d[i] * fs[i];
      if (a.act == ACT_GELU) go *= gelu_grad_as_forward<T>(xh * g[i] + be[i]);
      x[i] = rstd * (go * g[i] - m1 - xh * m2);
    }
    store8(dx + ((size_t)b * a.HW + row) * a.ld_dx + c8, x);
  }
}

// Register-cached variant (every shape of the U-Net): a cluster of CS CTAs owns one sample (CS = 1 up to 2048 8-channel
// vectors, else 2/4/8), each thread keeps its <= 4 vectors of (g = d out/d gn * gamma, xhat) in registers between the
// reduction pass and the dx pass, so dy / raw are read once and gelu' is evaluated once.  The two per-sample sums cross the
// cluster through distributed shared memory; per-channel sums go out as atomics.
template <typename T, int VPT>
__global__ void __launch_bounds__(512) gn_bwd_cached_kernel(GnBwdArgs a, int CS) {
  namespace cg = cooperative_groups;
  extern __shared__ float red[];       // [threads][8]
  __shared__ float wsum[2][16];
  __shared__ float cta_part[2];
  const int b = blockIdx.x / CS, part = blockIdx.x - b * CS;
  const int tid = threadIdx.x, nthreads = blockDim.x;
  double s = 0.0, q = 0.0;
  for (int p = 0; p < a.P; ++p) {
    const float2 sq = __ldg(reinterpret_cast<const float2*>(a.stats) + (size_t)b * a.P + p);
    s += (double)sq.x;
    q += (double)sq.y;
  }
  const double n = (double)a.HW * (double)a.C;
  const double dmean = s / n;
  double var = q / n - dmean * dmean;
  if (var < 0.0) var = 0.0;
  const float mean = (float)dmean;
  const float rstd = (float)(1.0 / sqrt(var + (double)a.eps));
  const T* __restrict__ dy = reinterpret_cast<const T*>(a.dy);
  const T* __restrict__ raw = reinterpret_cast<const T*>(a.raw);
  T* __restrict__ dx = reinterpret_cast<T*>(a.dx);
  const int vpr = a.C >> 3;            // divides nthreads
  const int c8 = (tid % vpr) << 3;
  const int row_step = nthreads / vpr;
  const int row0 = part * (a.HW / CS) + tid / vpr;
  float g[8], be[8], te[8], fs[8];
  load8(a.gamma + c8, g);
  load8(a.beta + c8, be);
#pragma unroll
  for (int i = 0; i < 8; ++i) { te[i] = 0.f; fs[i] = 1.f; }
  const bool has_film = a.film != nullptr, has_temb = a.temb != nullptr;
  if (has_temb) load8(a.temb + (size_t)b * SPDM_TEMB_WIDTH + a.temb_off + c8, te);
  if (has_film) load8(a.film + (size_t)b * SPDM_FILM_WIDTH + a.film_off + c8, fs);
  float gg[VPT][8], xh[VPT][8];
  float acc_dg[8], acc_db[8], acc_fb[8], acc_fs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc_dg[i] = 0.f; acc_db[i] = 0.f; acc_fb[i] = 0.f; acc_fs[i] = 0.f; }
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int row = row0 + k * row_step;
    float d[8], x[8];
    load8(dy + ((size_t)b * a.HW + row) * a.ld_dy + c8, d);
    load8(raw + ((size_t)b * a.HW + row) * a.ld_raw + c8, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xhv = (x[i] - mean) * rstd;
      const float u = xhv * g[i] + be[i];
      float go = d[i];
      if (has_film) { acc_fb[i] += d[i]; acc_fs[i] = fmaf(d[i], u + te[i], acc_fs[i]); go = d[i] * fs[i]; }
      else if (has_temb) acc_fb[i] += d[i];
      if (a.act == ACT_GELU) go *= gelu_grad_as_forward<T>(u);
      acc_dg[i] = fmaf(go, xhv, acc_dg[i]);
      acc_db[i] += go;
      const float gv = go * g[i];
      s1 += gv;
      s2 = fmaf(gv, xhv, s2);
      gg[k][i] = gv;
      xh[k][i] = xhv;
    }
  }
  // per-sample sums: warp -> CTA -> cluster
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if ((tid & 31) == 0) { wsum[0][tid >> 5] = s1; wsum[1][tid >> 5] = s2; }
  __syncthreads();
  float t1 = 0.f, t2 = 0.f;
  for (int i = 0; i < (nthreads >> 5); ++i) { t1 += wsum[0][i]; t2 += wsum[1][i]; }
  if (CS > 1) {
    cg::cluster_group cluster = cg::this_cluster();
    if (tid == 0) { cta_part[0] = t1; cta_part[1] = t2; }
    cluster.sync();
    t1 = 0.f; t2 = 0.f;
    for (int r = 0; r < CS; ++r) {
      const float* peer = cluster.map_shared_rank(cta_part, r);
      t1 += peer[0];
      t2 += peer[1];
    }
    cluster.sync();  // nobody leaves (or reuses cta_part) while a peer may still be reading it
  }
  const float inv_n = 1.0f / ((float)a.HW * (float)a.C);
  const float m1 = t1 * inv_n, m2 = t2 * inv_n;
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int row = row0 + k * row_step;
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = rstd * (gg[k][i] - m1 - xh[k][i] * m2);
    store8(dx + ((size_t)b * a.HW + row) * a.ld_dx + c8, o);
  }
  // per-channel sums of the four quantities in one pass: fold lanes that share channels inside each warp with shuffles, park
  // one row per warp (or per warp segment when vpr < 32) in shared memory, then vpr threads per quantity add the rows up.
  //   red layout: [quantity q][row r][vpr][8]
  const int seg = vpr < 32 ? vpr : 32;          // lanes l and l + seg (+ 2 seg ...) of a warp hold the same channels
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    if (off >= seg) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc_dg[i] += __shfl_xor_sync(0xffffffffu, acc_dg[i], off);
        acc_db[i] += __shfl_xor_sync(0xffffffffu, acc_db[i], off);
        acc_fb[i] += __shfl_xor_sync(0xffffffffu, acc_fb[i], off);
        acc_fs[i] += __shfl_xor_sync(0xffffffffu, acc_fs[i], off);
      }
    }
  }
  // after the fold, lanes 0..seg-1 of every warp hold the warp's sums for channel group (tid % vpr)
  const int lane = tid & 31, warp = tid >> 5;
  const int rows_per_q = (nthreads * seg / 32) / vpr;   // partial rows per quantity
  const size_t qstride = (size_t)rows_per_q * vpr * 8;
  const bool use_part = a.part != nullptr && CS == 1;
  const int n_pass = (has_film || has_temb) ? 2 : 1;    // pass 0: (d gamma, d beta); pass 1: (sum dy, sum dy*(gn+temb))
  for (int pass = 0; pass < n_pass; ++pass) {
    __syncthreads();
    if (lane < seg) {
      const int cgp = (warp * 32 + lane) % vpr;
      const int r = (warp * seg) / vpr;
      float* dst = red + ((size_t)r * vpr + cgp) * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        dst[i] = pass == 0 ? acc_dg[i] : acc_fb[i];
        dst[qstride + i] = pass == 0 ? acc_db[i] : acc_fs[i];
      }
    }
    __syncthreads();
    if (tid < 2 * vpr) {
      const int qi = 2 * pass + tid / vpr, cgp = tid % vpr;
      float tot[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int r = 0; r < rows_per_q; ++r) {
        const float* src = red + (qi & 1) * qstride + ((size_t)r * vpr + cgp) * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) tot[i] += src[i];
      }
      const int cc = cgp << 3;
      if (qi == 0) {
        if (use_part) store8(a.part + (size_t)b * 2 * a.C + cc, tot);
        else {
#pragma unroll
          for (int i = 0; i < 8; ++i) atomicAdd(a.dgamma + cc + i, tot[i]);
        }
      } else if (qi == 1) {
        if (use_part) store8(a.part + (size_t)b * 2 * a.C + a.C + cc, tot);
        else {
#pragma unroll
          for (int i = 0; i < 8; ++i) atomicAdd(a.dbeta + cc + i, tot[i]);
        }
      } else if (qi == 2) {   // sum_hw dy: FiLM bias gradient and (times the FiLM scale) the time-embedding gradient
        if (has_film) {
          float* dst = a.d_film + (size_t)b * SPDM_FILM_WIDTH + a.film_off + a.C + cc;
#pragma unroll
          for (int i = 0; i < 8; ++i) atomicAdd(dst + i, tot[i]);
        }
        if (has_temb) {
          float fsc[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
          if (has_film) load8(a.film + (size_t)b * SPDM_FILM_WIDTH + a.film_off + cc, fsc);
          float* dst = a.d_temb + (size_t)b * SPDM_TEMB_WIDTH + a.temb_off + cc;
#pragma unroll
          for (int i = 0; i < 8; ++i) atomicAdd(dst + i, tot[i] * fsc[i]);
        }
      } else if (has_film) {  // sum_hw dy * (gn + temb): FiLM scale gradient
        float* dst = a.d_film + (size_t)b * SPDM_FILM_WIDTH + a.film_off + cc;
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(dst + i, tot[i]);
      }
    }
  }
}

template <typename T, int VPT> void launch_gn_cached(const GnBwdArgs& a, int B, int CS, int threads, cudaStream_t s) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(B * CS);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = (size_t)2 * threads * 8 * sizeof(float);  // two quantities per pass, <= one row per warp segment
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, gn_bwd_cached_kernel<T, VPT>, a, CS);
}
// d gamma[c] += sum_b part[b][c], d beta[c] += sum_b part[b][C + c]: 32 columns x 32 row groups per block
__global__ void __launch_bounds__(1024) gn_part_finalize_kernel(const float* __restrict__ part, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                int B, int C) {
  __shared__ float sm[32][33];
  const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;
  float acc = 0.f;
  if (col < 2 * C) {
#pragma unroll 8
    for (int b = rg; b < B; b += 32) acc += __ldg(part + (size_t)b * 2 * C + col);
  }
  sm[rg][lane] = acc;
  __syncthreads();
  if (rg == 0 && col < 2 * C) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) t += sm[r][lane];
    if (col < C) atomicAdd(dgamma + col, t);
    else atomicAdd(dbeta + col - C, t);
  }
}
}  // namespace
// NOTE: d_film / d_temb are ACCUMULATED by the cached kernel (zero them first) and overwritten by the fallback kernel.
// Returns true when the per-sample (d gamma, d beta) partials in a.part still have to be summed over the batch
// (launch_gn_part_finalize -- nothing else of the step reads them, so the caller may put it on another stream).
template <typename T> bool launch_gn_bwd(const GnBwdArgs& a, int B, cudaStream_t s) {
  const int nvec = a.HW * (a.C >> 3), vpr = a.C >> 3;
  if ((nvec & (nvec - 1)) == 0 && nvec >= 64 && nvec <= 16384) {
    static int tmax = 0;   // SPDM_GN_BWD_T: threads per CTA (512 default; 256 = two CTAs per SM, twice the cluster size)
    if (!tmax) { const char* e = getenv("SPDM_GN_BWD_T"); tmax = e ? atoi(e) : 512; if (tmax != 256 && tmax != 512) tmax = 512; }
    const int cap = tmax * 4;
    int CS = nvec > cap ? nvec / cap : 1;
    if (CS > 8) CS = 8;
    const int nv = nvec / CS;
    const int threads = nv < tmax ? nv : tmax;
    const int vpt = nv / threads;
    if (threads % vpr == 0 && a.HW % CS == 0 && (a.HW / CS) * vpr == nv && (vpt == 1 || vpt == 2 || vpt == 4)) {
      if (vpt == 4) launch_gn_cached<T, 4>(a, B, CS, threads, s);
      else if (vpt == 2) launch_gn_cached<T, 2>(a, B, CS, threads, s);
      else launch_gn_cached<T, 1>(a, B, CS, threads, s);
      COUNT_LAUNCH();
      return a.part && CS == 1;
    }
  }
  gn_bwd_kernel<T><<<B, 256, 0, s>>>(a);
  COUNT_LAUNCH();
  return false;
}
template bool launch_gn_bwd<float>(const GnBwdArgs&, int, cudaStream_t);
template bool launch_gn_bwd<bf16>(const GnBwdArgs&, int, cudaStream_t);
void launch_gn_part_finalize(const float* part, float* dgamma, float* dbeta, int B, int C, cudaStream_t s) {
  gn_part_finalize_kernel<<<cdiv(2 * C, 32), 1024, 0, s>>>(part, dgamma, dbeta, B, C);
  COUNT_LAUNCH();
}

// =================================================================================================
// Weight gradient on CUDA cores (fp32 accumulate): 64(ci) x 64(co) tile per block, split over the pixel
// dimension; the partial tile is added to dw (PyTorch layout) with atomics.
//   nn.Conv2d(k=3, pad=1) weight grad (models/Unet_FiLmLayer.py:101,103) / nn.Linear weight grad.
// =================================================================================================
namespace {
template <typename TX, typename TDY>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(WgradArgs a, int ksplit) {
  __shared__ float Xs[16][64 + 4];
  __shared__ float Ds[16][64 + 4];
  const TX* __restrict__ x = reinterpret_cast<const TX*>(a.x);
  const TDY* __restrict__ dy = reinterpret_cast<const TDY*>(a.dy);
  const int tid = threadIdx.x;
  const int co_tiles = (a.Cout + 63) / 64;
  const int ci0 = (blockIdx.x / co_tiles) * 64, co0 = (blockIdx.x % co_tiles) * 64;
  const int tap = blockIdx.y;
  int ddy = 0, ddx = 0;
  if (a.taps == 9) { ddy = tap / 3 - 1; ddx = tap % 3 - 1; }
  if (a.taps == 9 && ((a.W == 1 && ddx != 0) || (a.H == 1 && ddy != 0))) return;  // structurally empty tap
  const long long per = ((a.M + ksplit - 1) / ksplit + 15) / 16 * 16;
  const long long m_begin = (long long)blockIdx.z * per;
  long long m_end = m_begin + per;
  if (m_end > a.M) m_end = a.M;
  const int tx = tid & 15, ty = tid >> 4;      // outputs: ci = ci0 + ty*4 + i, co = co0 + tx*4 + j
  const int lrow = tid >> 4, lc = (tid & 15) * 4;  // loads: row lrow of the 16-row chunk, 4 channels at lc
  const int HW = a.H * a.W;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long long m0 = m_begin; m0 < m_end; m0 += 16) {
    const long long m = m0 + lrow;
    float xv[4] = {0.f, 0.f, 0.f, 0.f}, dv[4] = {0.f, 0.f, 0.f, 0.f};
    if (m < m_end) {
      bool valid = true;
      long long src = m;
      if (a.taps == 9) {
        const int rem = (int)(m % HW);
        const int hh = rem / a.W + ddy, ww = rem % a.W + ddx;
        valid = hh >= 0 && hh < a.H && ww >= 0 && ww < a.W;
        src = m + ddy * a.W + ddx;
      }
      if (valid) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (ci0 + lc + i < a.Cin) xv[i] = to_f32<TX>(x[src * a.ld_x + ci0 + lc + i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (co0 + lc + i < a.Cout) dv[i] = to_f32<TDY>(dy[m * a.ld_dy + co0 + lc + i]);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) { Xs[lrow][lc + i] = xv[i]; Ds[lrow][lc + i] = dv[i]; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 xa = *reinterpret_cast<const float4*>(&Xs[k][ty * 4]);
      const float4 da = *reinterpret_cast<const float4*>(&Ds[k][tx * 4]);
      const float xr[4] = {xa.x, xa.y, xa.z, xa.w};
      const float dr[4] = {da.x, da.y, da.z, da.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xr[i], dr[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ci = ci0 + ty * 4 + i;
    if (ci >= a.Cin) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co >= a.Cout) continue;
      atomicAdd(a.dw + ((size_t)co * a.Cin + ci) * a.taps + tap, acc[i][j]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ dy, int ld, long long M, int N, float* __restrict__ out, long long rows_per_block) {
  __shared__ float sm[4][64];
  const int c = blockIdx.x * 64 + (threadIdx.x & 63), rg = threadIdx.x >> 6;
  const long long m0 = (long long)blockIdx.y * rows_per_block;
  long long m1 = m0 + rows_per_block;
  if (m1 > M) m1 = M;
  float acc = 0.f;
  if (c < N)
    for (long long m = m0 + rg; m < m1; m += 4) acc += to_f32<T>(dy[m * ld + c]);
  sm[rg][threadIdx.x & 63] = acc;
  __syncthreads();
  if (rg == 0 && c < N) atomicAdd(out + c, sm[0][threadIdx.x] + sm[1][threadIdx.x] + sm[2][threadIdx.x] + sm[3][threadIdx.x]);
}
// Column sums with 16-byte (bf16) / 32-byte (fp32) loads: a block covers 64 columns as 8 vectors x 32 row lanes, four independent
// loads in flight per thread (the scalar kernel above spends ~11 us on any bias gradient of the step, 36 of them per step).
template <typename T>
__global__ void __launch_bounds__(256) colsum8_kernel(const T* __restrict__ dy, int ld, long long M, int N, float* __restrict__ out, long long rows_per_block) {
  __shared__ float sm[32][65];
  const int cv = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c = blockIdx.x * 64 + cv * 8;
  const long long m0 = (long long)blockIdx.y * rows_per_block;
  long long m1 = m0 + rows_per_block;
  if (m1 > M) m1 = M;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c < N) {
    long long m = m0 + rl;
    for (; m + 96 < m1; m += 128) {
      float v[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) load8(dy + (m + 32 * u) * ld + c, v[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += v[u][i];
    }
    for (; m < m1; m += 32) {
      float v[8];
      load8(dy + m * ld + c, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sm[rl][cv * 8 + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < 64 && blockIdx.x * 64 + threadIdx.x < N) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) t += sm[r][threadIdx.x];
    atomicAdd(out + blockIdx.x * 64 + threadIdx.x, t);
  }
}
}  // namespace
template <typename TX, typename TDY> void launch_wgrad_simt(const WgradArgs& a, cudaStream_t s) {
  const int tiles = cdiv(a.Cin, 64) * cdiv(a.Cout, 64);
  long long ks = (2 * 148 + (long long)tiles * a.taps - 1) / ((long long)tiles * a.taps);
  const long long max_ks = (a.M + 127) / 128;
  if (ks > max_ks) ks = max_ks;
  if (ks < 1) ks = 1;
  wgrad_simt_kernel<TX, TDY><<<dim3(tiles, a.taps, (unsigned)ks), 256, 0, s>>>(a, (int)ks);
  COUNT_LAUNCH();
}
template void launch_wgrad_simt<float, float>(const WgradArgs&, cudaStream_t);
template void launch_wgrad_simt<bf16, bf16>(const WgradArgs&, cudaStream_t);
template void launch_wgrad_simt<float, bf16>(const WgradArgs&, cudaStream_t);
template <typename T> void launch_colsum(const T* dy, int ld, long long M, int N, float* out, cudaStream_t s) {
  int gy = (int)((M + 255) / 256);
  if (gy > 148 * 4) gy = 148 * 4;
  const long long rpb = (M + gy - 1) / gy;
  if (N % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(dy) & (8 * sizeof(T) - 1)) == 0)
    colsum8_kernel<T><<<dim3(cdiv(N, 64), gy), 256, 0, s>>>(dy, ld, M, N, out, rpb);
  else
    colsum_kernel<T><<<dim3(cdiv(N, 64), gy), 256, 0, s>>>(dy, ld, M, N, out, rpb);
  COUNT_LAUNCH();
}
template void launch_colsum<float>(const float*, int, long long, int, float*, cudaStream_t);
template void launch_colsum<bf16>(const bf16*, int, long long, int, float*, cudaStream_t);

// =================================================================================================
// MaxPool2d(2) backward (models/Unet_FiLmLayer.py:132) + skip-connection gradient add;
// bilinear x2 align_corners=True backward (:191) in gather form.
// =================================================================================================
namespace {
template <typename T>
__global__ void pool_bwd_kernel(const T* __restrict__ x, int ld_x, const T* __restrict__ dy, int ld_dy, const T* __restrict__ add, int ld_add,
                                T* __restrict__ dx, int ld_dx, long long total, int Ho, int Wo, int C) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= total) return;
  const int vpr = C >> 3;
  const long long orow = v / vpr;
  const int c8 = (int)(v - orow * vpr) << 3;
  const int wo = (int)(orow % Wo);
  const long long t = orow / Wo;
  const int ho = (int)(t % Ho);
  const long long b = t / Ho;
  const int Wi = Wo * 2, Hi = Ho * 2;
  const long long r00 = (b * Hi + 2 * ho) * Wi + 2 * wo;
  const long long rr[4] = {r00, r00 + 1, r00 + Wi, r00 + Wi + 1};
  float xin[4][8], d[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) load8(x + rr[k] * ld_x + c8, xin[k]);
  load8(dy + orow * ld_dy + c8, d);
  int arg[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {  // first maximum in window scan order (strict >), as ATen's max_pool2d
    float m = xin[0][i];
    int am = 0;
#pragma unroll
    for (int k = 1; k < 4; ++k)
      if (xin[k][i] > m) { m = xin[k][i]; am = k; }
    arg[i] = am;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float o[8];
    if (add) load8(add + rr[k] * ld_add + c8, o);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] += (arg[i] == k) ? d[i] : 0.f;
    store8(dx + rr[k] * ld_dx + c8, o);
  }
}

template <typename T>
__global__ void upsample_bwd_kernel(const T* __restrict__ dy, int ld_dy, T* __restrict__ dx, int ld_dx, long long total, int Hi, int Wi, int C) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= total) return;
  const int vpr = C >> 3;
  const long long irow = v / vpr;
  const int c8 = (int)(v - irow * vpr) << 3;
  const int wi = (int)(irow % Wi);
  const long long t = irow / Wi;
  const int hi = (int)(t % Hi);
  const long long b = t / Hi;
  const int Ho = Hi * 2, Wo = Wi * 2;
  const float sh = (Ho > 1) ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f;
  const float sw = (Wo > 1) ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
  // candidate outputs: src(o) = o*s in (i-1, i+1)
  int ho_lo = 0, ho_hi = Ho - 1, wo_lo = 0, wo_hi = Wo - 1;
  if (sh > 0.f) { ho_lo = max(0, (int)floorf((hi - 1) / sh) - 1); ho_hi = min(Ho - 1, (int)ceilf((hi + 1) / sh) + 1); }
  if (sw > 0.f) { wo_lo = max(0, (int)floorf((wi - 1) / sw) - 1); wo_hi = min(Wo - 1, (int)ceilf((wi + 1) / sw) + 1); }
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int ho = ho_lo; ho <= ho_hi; ++ho) {
    // identical arithmetic to upsample_kernel (kernels.cu)
    const float fh = sh * ho;
    const int h0 = (int)fh;
    const int h1 = h0 + ((h0 < Hi - 1) ? 1 : 0);
    const float lh1 = fh - h0, lh0 = 1.f - lh1;
    const float wh = (h0 == hi ? lh0 : 0.f) + (h1 == hi ? lh1 : 0.f);
    if (wh == 0.f) continue;
    for (int wo = wo_lo; wo <= wo_hi; ++wo) {
      const float fw = sw * wo;
      const int w0 = (int)fw;
      const int w1 = w0 + ((w0 < Wi - 1) ? 1 : 0);
      const float lw1 = fw - w0, lw0 = 1.f - lw1;
      const float ww = (w0 == wi ? lw0 : 0.f) + (w1 == wi ? lw1 : 0.f);
      if (ww == 0.f) continue;
      float d[8];
      load8(dy + ((b * Ho + ho) * Wo + wo) * ld_dy + c8, d);
      const float wgt = wh * ww;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(wgt, d[i], acc[i]);
    }
  }
  store8(dx + irow * ld_dx + c8, acc);
}
}  // namespace
template <typename T>
void launch_pool_bwd(const T* x, int ld_x, const T* dy, int ld_dy, const T* add, int ld_add, T* dx, int ld_dx, int B, int Ho, int Wo, int C,
                     cudaStream_t s) {
  const long long total = (long long)B * Ho * Wo * (C >> 3);
  pool_bwd_kernel<T><<<cdiv(total, 256), 256, 0, s>>>(x, ld_x, dy, ld_dy, add, ld_add, dx, ld_dx, total, Ho, Wo, C);
  COUNT_LAUNCH();
}
template <typename T> void launch_upsample_bwd(const T* dy, int ld_dy, T* dx, int ld_dx, int B, int Hi, int Wi, int C, cudaStream_t s) {
  const long long total = (long long)B * Hi * Wi * (C >> 3);
  upsample_bwd_kernel<T><<<cdiv(total, 256), 256, 0, s>>>(dy, ld_dy, dx, ld_dx, total, Hi, Wi, C);
  COUNT_LAUNCH();
}
template void launch_pool_bwd<float>(const float*, int, const float*, int, const float*, int, float*, int, int, int, int, int, cudaStream_t);
template void launch_pool_bwd<bf16>(const bf16*, int, const bf16*, int, const bf16*, int, bf16*, int, int, int, int, int, cudaStream_t);
template void launch_upsample_bwd<float>(const float*, int, float*, int, int, int, int, int, cudaStream_t);
template void launch_upsample_bwd<bf16>(const bf16*, int, bf16*, int, int, int, int, int, cudaStream_t);

// =================================================================================================
// LayerNorm backward (models/Unet_FiLmLayer.py:51,53): one warp per token row, C <= 256.
// =================================================================================================
namespace {
template <typename T, int LANES>  // LANES = C / 8 lanes per token row, 32 / LANES rows per warp (like layernorm_kernel)
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const T* __restrict__ dy, int ld_dy, const T* __restrict__ x, int ld_x,
                                                            const float* __restrict__ g, const T* __restrict__ add, int ld_add, T* __restrict__ dx,
                                                            int ld_dx, float* __restrict__ dgamma, float* __restrict__ dbeta, long long M) {
  constexpr int C = LANES * 8, ROWS_PER_WARP = 32 / LANES;
  __shared__ float sg[C], sb[C];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < C; i += 256) { sg[i] = 0.f; sb[i] = 0.f; }
  __syncthreads();
  const int sub = lane / LANES, l = lane % LANES;
  const int c0 = l * 8;
  float gg[8];
  load8(g + c0, gg);
  float adg[8], adb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { adg[i] = 0.f; adb[i] = 0.f; }
  constexpr float inv_c = 1.0f / (float)C;
  const long long rows_per_block_iter = 8LL * ROWS_PER_WARP;
  for (long long base = (long long)blockIdx.x * rows_per_block_iter; base < M; base += (long long)gridDim.x * rows_per_block_iter) {
    const long long row = base + warp * ROWS_PER_WARP + sub;
    const bool ok = row < M;
    float xv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (ok) { load8(x + row * ld_x + c0, xv); load8(dy + row * ld_dy + c0, d); }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += xv[i];
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * inv_c;
    float qq = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float t = xv[i] - mean; qq = fmaf(t, t, qq); }
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
    const float rstd = rsqrtf(qq * inv_c + 1e-5f);
    float a1 = 0.f, a2 = 0.f, xh[8], dh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      xh[i] = (xv[i] - mean) * rstd;
      dh[i] = d[i] * gg[i];
      a1 += dh[i];
      a2 = fmaf(dh[i], xh[i], a2);
      adg[i] = fmaf(d[i], xh[i], adg[i]);   // d == 0 for rows past M
      adb[i] += d[i];
    }
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) { a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o); }
    a1 *= inv_c;
    a2 *= inv_c;
    if (ok) {
      float o8[8];
      if (add) load8(add + row * ld_add + c0, o8);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o8[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) o8[i] += rstd * (dh[i] - a1 - xh[i] * a2);
      store8(dx + row * ld_dx + c0, o8);
    }
  }
  // lanes with the same l inside a warp hold the same channels: fold them before touching shared memory
#pragma unroll
  for (int o = LANES; o < 32; o <<= 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { adg[i] += __shfl_xor_sync(0xffffffffu, adg[i], o); adb[i] += __shfl_xor_sync(0xffffffffu, adb[i], o); }
  }
  if (sub == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { atomicAdd(&sg[c0 + i], adg[i]); atomicAdd(&sb[c0 + i], adb[i]); }
  }
  __syncthreads();
  for (int i = tid; i < C; i += 256) { atomicAdd(dgamma + i, sg[i]); atomicAdd(dbeta + i, sb[i]); }
}
}  // namespace
template <typename T>
void launch_layernorm_bwd(const T* dy, int ld_dy, const T* x, int ld_x, const float* g, const T* add, int ld_add, T* dx, int ld_dx,
                          float* dgamma, float* dbeta, long long M, int C, cudaStream_t s) {
  const int rows_per_iter = 8 * (256 / C);
  int grid = cdiv(M, rows_per_iter);
  if (grid > 148 * 4) grid = 148 * 4;
  if (C == 64) layernorm_bwd_kernel<T, 8><<<grid, 256, 0, s>>>(dy, ld_dy, x, ld_x, g, add, ld_add, dx, ld_dx, dgamma, dbeta, M);
  else if (C == 128) layernorm_bwd_kernel<T, 16><<<grid, 256, 0, s>>>(dy, ld_dy, x, ld_x, g, add, ld_add, dx, ld_dx, dgamma, dbeta, M);
  else layernorm_bwd_kernel<T, 32><<<grid, 256, 0, s>>>(dy, ld_dy, x, ld_x, g, add, ld_add, dx, ld_dx, dgamma, dbeta, M);
  COUNT_LAUNCH();
}
template void launch_layernorm_bwd<float>(const float*, int, const float*, int, const float*, const float*, int, float*, int, float*, float*,
                                          long long, int, cudaStream_t);
template void launch_layernorm_bwd<bf16>(const bf16*, int, const bf16*, int, const float*, const bf16*, int, bf16*, int, float*, float*,
                                         long long, int, cudaStream_t);

// =================================================================================================
// Attention core backward (nn.MultiheadAttention, 4 heads; models/Unet_FiLmLayer.py:50,76).
// One block per (sample, head): Q, K, V, dO of the group live in shared memory as fp32.
//   pass 1 (thread per query i): log-sum-exp of the scores, D_i = dO_i . O_i, dQ_i = scale * sum_j dS_ij K_j
//   pass 2 (thread per key j):   dV_j = sum_i P_ij dO_i ; dK_j = scale * sum_i dS_ij Q_i
//   with P_ij = exp(scale q_i.k_j - lse_i), dS_ij = P_ij (dO_i . V_j - D_i)
// =================================================================================================
namespace {
template <typename T, int HD>
__global__ void __launch_bounds__(128) sdpa_bwd_kernel(const T* __restrict__ qkv, const T* __restrict__ o, const T* __restrict__ d_o,
                                                       T* __restrict__ d_qkv, int L, int C, int heads) {
  extern __shared__ float sm[];
  float* Qs = sm;
  float* Ks = Qs + (size_t)L * HD;
  float* Vs = Ks + (size_t)L * HD;
  float* Gs = Vs + (size_t)L * HD;  // dO
  float* lse = Gs + (size_t)L * HD;
  float* Dd = lse + L;
  const int b = blockIdx.x / heads, h = blockIdx.x - b * heads;
  const int ld = 3 * C;
  const int tid = threadIdx.x;
  const int vpr = HD >> 3;
  for (int v = tid; v < L * vpr; v += 128) {
    const int j = v / vpr, d8 = (v - j * vpr) << 3;
    const T* row = qkv + ((size_t)b * L + j) * ld + h * HD + d8;
    float t8[8];
    load8(row, t8);
#pragma unroll
    for (int i = 0; i < 8; ++i) Qs[j * HD + d8 + i] = t8[i];
    load8(row + C, t8);
#pragma unroll
    for (int i = 0; i < 8; ++i) Ks[j * HD + d8 + i] = t8[i];
    load8(row + 2 * C, t8);
#pragma unroll
    for (int i = 0; i < 8; ++i) Vs[j * HD + d8 + i] = t8[i];
    load8(d_o + ((size_t)b * L + j) * C + h * HD + d8, t8);
#pragma unroll
    for (int i = 0; i < 8; ++i) Gs[j * HD + d8 + i] = t8[i];
  }
  __syncthreads();
  const float scale = rsqrtf((float)HD);
  for (int i = tid; i < L; i += 128) {
    float q[HD], g[HD], dq[HD];
    float Di = 0.f;
    const T* orow = o + ((size_t)b * L + i) * C + h * HD;
#pragma unroll
    for (int d = 0; d < HD; d += 8) {
      float t8[8];
      load8(orow + d, t8);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        q[d + e] = Qs[i * HD + d + e] * scale;
        g[d + e] = Gs[i * HD + d + e];
        dq[d + e] = 0.f;
        Di = fmaf(g[d + e], t8[e], Di);
      }
    }
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < L; ++j) {
      float sdot = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) sdot = fmaf(q[d], Ks[j * HD + d], sdot);
      if (sdot > m) { l = l * __expf(m - sdot) + 1.f; m = sdot; }
      else l += __expf(sdot - m);
    }
    const float ls = m + __logf(l);
    for (int j = 0; j < L; ++j) {
      float sdot = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) { sdot = fmaf(q[d], Ks[j * HD + d], sdot); dp = fmaf(g[d], Vs[j * HD + d], dp); }
      const float ds = __expf(sdot - ls) * (dp - Di);
#pragma unroll
      for (int d = 0; d < HD; ++d) dq[d] = fmaf(ds, Ks[j * HD + d], dq[d]);
    }
    lse[i] = ls;
    Dd[i] = Di;
    T* drow = d_qkv + ((size_t)b * L + i) * ld + h * HD;
#pragma unroll
    for (int d = 0; d < HD; d += 8) {
      float t8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) t8[e] = dq[d + e] * scale;
      store8(drow + d, t8);
    }
  }
  __syncthreads();
  for (int j = tid; j < L; j += 128) {
    float dk[HD], dv[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) { dk[d] = 0.f; dv[d] = 0.f; }
    for (int i = 0; i < L; ++i) {
      float sdot = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) {
        sdot = fmaf(Qs[i * HD + d], Ks[j * HD + d], sdot);
        dp = fmaf(Gs[i * HD + d], Vs[j * HD + d], dp);
      }
      const float p = __expf(sdot * scale - lse[i]);
      const float ds = p * (dp - Dd[i]) * scale;
#pragma unroll
      for (int d = 0; d < HD; ++d) { dv[d] = fmaf(p, Gs[i * HD + d], dv[d]); dk[d] = fmaf(ds, Qs[i * HD + d], dk[d]); }
    }
    T* drow = d_qkv + ((size_t)b * L + j) * ld + h * HD;
#pragma unroll
    for (int d = 0; d < HD; d += 8) {
      store8(drow + C + d, dk + d);
      store8(drow + 2 * C + d, dv + d);
    }
  }
}

template <typename T> __global__ void gelu_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long nvec) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nvec) return;
  float t[8];
  load8(x + v * 8, t);
#pragma unroll
  for (int i = 0; i < 8; ++i) t[i] = gelu_exact(t[i]);
  store8(y + v * 8, t);
}
template <typename T> __global__ void gelu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ pre, T* __restrict__ dx, long long nvec) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nvec) return;
  float d[8], u[8];
  load8(dy + v * 8, d);
  load8(pre + v * 8, u);
#pragma unroll
  for (int i = 0; i < 8; ++i) d[i] *= gelu_grad(u[i]);
  store8(dx + v * 8, d);
}
}  // namespace

// -------------------------------------------------------------------------------------------------
// Tensor-core version of the same backward for bf16 (warp-level mma.sync.m16n8k16, fp32 accumulate):
//   phase 1, a warp per 16-query tile: S = Q K^T twice (once for the log-sum-exp, once to form P), dP = dO V^T,
//            dS = P o (dP - D), dQ = scale * dS K      (the S / dS accumulator tiles become A fragments directly)
//   phase 2, a warp per 16-key tile:   S^T = K Q^T, dP^T = V dO^T, dV = P^T dO, dK = scale * dS^T Q
// Q, K, V, dO of a (sample, head) group sit in shared memory as bf16 rows padded by 16 bytes (conflict-free ldmatrix);
// groups with L <= 16 are packed four to a block (one warp each).  Rows / keys past L are zero-filled and masked.
// -------------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const bf16* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const bf16* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], const bf16* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t (&r)[2], const bf16* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}

template <int HD>
__global__ void __launch_bounds__(128) sdpa_bwd_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ o, const bf16* __restrict__ d_o,
                                                           bf16* __restrict__ d_qkv, int n_groups, int L, int Lp, int C, int heads, int gpb) {
  constexpr int STR = HD + 8;      // padded row stride (elements)
  constexpr int KS = HD / 16;      // k-steps over the head dimension
  constexpr int DN = HD / 8;       // 8-wide output tiles over the head dimension
  extern __shared__ __align__(16) unsigned char smraw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const size_t region = (size_t)4 * Lp * STR * sizeof(bf16) + (size_t)2 * Lp * sizeof(float);
  const int ld = 3 * C;
  const float scale = rsqrtf((float)HD);
  constexpr int VPR = HD / 8;
  // ---- stage Q, K, V, dO (zero rows past L) and D = rowsum(dO o O), lse = +inf (until phase 1 fills it) ----
  for (int idx = tid; idx < gpb * 4 * Lp * VPR; idx += 128) {
    const int v = idx % VPR;
    int r = idx / VPR;
    const int row = r % Lp; r /= Lp;
    const int arr = r & 3, gl = r >> 2;
    const int grp = blockIdx.x * gpb + gl;
    bf16* dst = reinterpret_cast<bf16*>(smraw + gl * region) + ((size_t)arr * Lp + row) * STR + v * 8;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (grp < n_groups && row < L) {
      const int b = grp / heads, h = grp - b * heads;
      const bf16* src = arr < 3 ? qkv + ((size_t)b * L + row) * ld + arr * C + h * HD + v * 8
                                : d_o + ((size_t)b * L + row) * C + h * HD + v * 8;
      val = *reinterpret_cast<const uint4*>(src);
    }
    *reinterpret_cast<uint4*>(dst) = val;
  }
  for (int idx = tid; idx < gpb * Lp; idx += 128) {
    const int row = idx % Lp, gl = idx / Lp;
    const int grp = blockIdx.x * gpb + gl;
    float* lse = reinterpret_cast<float*>(smraw + gl * region + (size_t)4 * Lp * STR * sizeof(bf16));
    float* Dd = lse + Lp;
    float dsum = 0.f;
    if (grp < n_groups && row < L) {
      const int b = grp / heads, h = grp - b * heads;
      const bf16* po = o + ((size_t)b * L + row) * C + h * HD;
      const bf16* pg = d_o + ((size_t)b * L + row) * C + h * HD;
#pragma unroll
      for (int d = 0; d < HD; d += 8) {
        float a8[8], b8[8];
        load8(po + d, a8);
        load8(pg + d, b8);
#pragma unroll
        for (int e = 0; e < 8; ++e) dsum = fmaf(a8[e], b8[e], dsum);
      }
    }
    Dd[row] = dsum;
    lse[row] = INFINITY;
  }
  __syncthreads();
  const int wpg = 4 / gpb;                         // warps per group
  const int gl = warp / wpg, wl = warp - gl * wpg;  // group inside the block, warp inside the group
  const int grp = blockIdx.x * gpb + gl;
  const bool active = grp < n_groups;
  bf16* Qs = reinterpret_cast<bf16*>(smraw + gl * region);
  bf16* Ks = Qs + (size_t)Lp * STR;
  bf16* Vs = Ks + (size_t)Lp * STR;
  bf16* Gs = Vs + (size_t)Lp * STR;
  float* lse = reinterpret_cast<float*>(Gs + (size_t)Lp * STR);
  float* Dd = lse + Lp;
  const int b = active ? grp / heads : 0, h = active ? grp - b * heads : 0;
  const int ntiles = Lp / 16;

  // ================= phase 1: queries =================
  for (int it = wl; it < ntiles && active; it += wpg) {
    const int r0 = it * 16;
    uint32_t qa[KS][4], ga[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      ldsm_x4(qa[ks], Qs + (size_t)(r0 + (lane & 15)) * STR + ks * 16 + (lane >> 4) * 8);
      ldsm_x4(ga[ks], Gs + (size_t)(r0 + (lane & 15)) * STR + ks * 16 + (lane >> 4) * 8);
    }
    float m0 = -INFINITY, l0 = 0.f, m1 = -INFINITY, l1 = 0.f;
    for (int nt = 0; nt < Lp / 8; ++nt) {
      float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t kb[2];
        ldsm_x2(kb, Ks + (size_t)(nt * 8 + (lane & 7)) * STR + ks * 16 + ((lane >> 3) & 1) * 8);
        mma16816(c, qa[ks], kb);
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool ok = nt * 8 + 2 * t + e < L;
        const float s0 = ok ? c[e] * scale : -INFINITY, s1 = ok ? c[2 + e] * scale : -INFINITY;
        if (s0 > m0) { l0 = l0 * __expf(m0 - s0) + 1.f; m0 = s0; } else if (ok) l0 += __expf(s0 - m0);
        if (s1 > m1) { l1 = l1 * __expf(m1 - s1) + 1.f; m1 = s1; } else if (ok) l1 += __expf(s1 - m1);
      }
    }
#pragma unroll
    for (int off = 1; off <= 2; off <<= 1) {  // merge the 4 lanes that share a row
      const float om0 = __shfl_xor_sync(0xffffffffu, m0, off), ol0 = __shfl_xor_sync(0xffffffffu, l0, off);
      const float om1 = __shfl_xor_sync(0xffffffffu, m1, off), ol1 = __shfl_xor_sync(0xffffffffu, l1, off);
      const float nm0 = fmaxf(m0, om0), nm1 = fmaxf(m1, om1);
      l0 = (m0 == -INFINITY ? 0.f : l0 * __expf(m0 - nm0)) + (om0 == -INFINITY ? 0.f : ol0 * __expf(om0 - nm0));
      l1 = (m1 == -INFINITY ? 0.f : l1 * __expf(m1 - nm1)) + (om1 == -INFINITY ? 0.f : ol1 * __expf(om1 - nm1));
      m0 = nm0; m1 = nm1;
    }
    const float lse0 = m0 + __logf(l0), lse1 = m1 + __logf(l1);
    const float D0 = Dd[r0 + g], D1 = Dd[r0 + g + 8];
    float dq[DN][4];
#pragma unroll
    for (int dn = 0; dn < DN; ++dn) { dq[dn][0] = 0.f; dq[dn][1] = 0.f; dq[dn][2] = 0.f; dq[dn][3] = 0.f; }
    for (int np = 0; np < ntiles; ++np) {
      float ds[2][4];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int nt = np * 2 + hh;
        float c[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint32_t kb[2], vb[2];
          ldsm_x2(kb, Ks + (size_t)(nt * 8 + (lane & 7)) * STR + ks * 16 + ((lane >> 3) & 1) * 8);
          ldsm_x2(vb, Vs + (size_t)(nt * 8 + (lane & 7)) * STR + ks * 16 + ((lane >> 3) & 1) * 8);
          mma16816(c, qa[ks], kb);
          mma16816(dp, ga[ks], vb);
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const bool ok = nt * 8 + 2 * t + e < L;
          const float p0 = ok ? __expf(c[e] * scale - lse0) : 0.f, p1 = ok ? __expf(c[2 + e] * scale - lse1) : 0.f;
          ds[hh][e] = p0 * (dp[e] - D0);
          ds[hh][2 + e] = p1 * (dp[2 + e] - D1);
        }
      }
      uint32_t dsa[4] = {pack_bf16x2(ds[0][0], ds[0][1]), pack_bf16x2(ds[0][2], ds[0][3]), pack_bf16x2(ds[1][0], ds[1][1]),
                         pack_bf16x2(ds[1][2], ds[1][3])};
#pragma unroll
      for (int dn = 0; dn < DN; ++dn) {
        uint32_t kb2[2];
        ldsm_x2_trans(kb2, Ks + (size_t)(np * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * STR + dn * 8);
        mma16816(dq[dn], dsa, kb2);
      }
    }
    const int row_a = r0 + g, row_b = r0 + g + 8;
    if (t == 0) {
      if (row_a < L) lse[row_a] = lse0;
      if (row_b < L) lse[row_b] = lse1;
    }
#pragma unroll
    for (int dn = 0; dn < DN; ++dn) {
      const int col = h * HD + dn * 8 + 2 * t;
      if (row_a < L) *reinterpret_cast<uint32_t*>(d_qkv + ((size_t)b * L + row_a) * ld + col) = pack_bf16x2(dq[dn][0] * scale, dq[dn][1] * scale);
      if (row_b < L) *reinterpret_cast<uint32_t*>(d_qkv + ((size_t)b * L + row_b) * ld + col) = pack_bf16x2(dq[dn][2] * scale, dq[dn][3] * scale);
    }
  }
  __syncthreads();
  // ================= phase 2: keys =================
  for (int jt = wl; jt < ntiles && active; jt += wpg) {
    const int k0 = jt * 16;
    uint32_t ka[KS][4], va[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      ldsm_x4(ka[ks], Ks + (size_t)(k0 + (lane & 15)) * STR + ks * 16 + (lane >> 4) * 8);
      ldsm_x4(va[ks], Vs + (size_t)(k0 + (lane & 15)) * STR + ks * 16 + (lane >> 4) * 8);
    }
    float dk[DN][4], dv[DN][4];
#pragma unroll
    for (int dn = 0; dn < DN; ++dn) {
#pragma unroll
      for (int e = 0; e < 4; ++e) { dk[dn][e] = 0.f; dv[dn][e] = 0.f; }
    }
    for (int np = 0; np < ntiles; ++np) {
      float pt[2][4], dst[2][4];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int nt = np * 2 + hh;
        float c[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint32_t qb[2], gb[2];
          ldsm_x2(qb, Qs + (size_t)(nt * 8 + (lane & 7)) * STR + ks * 16 + ((lane >> 3) & 1) * 8);
          ldsm_x2(gb, Gs + (size_t)(nt * 8 + (lane & 7)) * STR + ks * 16 + ((lane >> 3) & 1) * 8);
          mma16816(c, ka[ks], qb);
          mma16816(dp, va[ks], gb);
        }
        const float2 ls = *reinterpret_cast<const float2*>(lse + nt * 8 + 2 * t);   // +inf for queries past L -> p = 0
        const float2 dd = *reinterpret_cast<const float2*>(Dd + nt * 8 + 2 * t);
        pt[hh][0] = __expf(c[0] * scale - ls.x); pt[hh][1] = __expf(c[1] * scale - ls.y);
        pt[hh][2] = __expf(c[2] * scale - ls.x); pt[hh][3] = __expf(c[3] * scale - ls.y);
        dst[hh][0] = pt[hh][0] * (dp[0] - dd.x); dst[hh][1] = pt[hh][1] * (dp[1] - dd.y);
        dst[hh][2] = pt[hh][2] * (dp[2] - dd.x); dst[hh][3] = pt[hh][3] * (dp[3] - dd.y);
      }
      uint32_t pa[4] = {pack_bf16x2(pt[0][0], pt[0][1]), pack_bf16x2(pt[0][2], pt[0][3]), pack_bf16x2(pt[1][0], pt[1][1]),
                        pack_bf16x2(pt[1][2], pt[1][3])};
      uint32_t dsa[4] = {pack_bf16x2(dst[0][0], dst[0][1]), pack_bf16x2(dst[0][2], dst[0][3]), pack_bf16x2(dst[1][0], dst[1][1]),
                         pack_bf16x2(dst[1][2], dst[1][3])};
#pragma unroll
      for (int dn = 0; dn < DN; ++dn) {
        uint32_t gb2[2], qb2[2];
        ldsm_x2_trans(gb2, Gs + (size_t)(np * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * STR + dn * 8);
        ldsm_x2_trans(qb2, Qs + (size_t)(np * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * STR + dn * 8);
        mma16816(dv[dn], pa, gb2);
        mma16816(dk[dn], dsa, qb2);
      }
    }
    const int row_a = k0 + g, row_b = k0 + g + 8;
#pragma unroll
    for (int dn = 0; dn < DN; ++dn) {
      const int col = h * HD + dn * 8 + 2 * t;
      if (row_a < L) {
        bf16* base = d_qkv + ((size_t)b * L + row_a) * ld + col;
        *reinterpret_cast<uint32_t*>(base + C) = pack_bf16x2(dk[dn][0] * scale, dk[dn][1] * scale);
        *reinterpret_cast<uint32_t*>(base + 2 * C) = pack_bf16x2(dv[dn][0], dv[dn][1]);
      }
      if (row_b < L) {
        bf16* base = d_qkv + ((size_t)b * L + row_b) * ld + col;
        *reinterpret_cast<uint32_t*>(base + C) = pack_bf16x2(dk[dn][2] * scale, dk[dn][3] * scale);
        *reinterpret_cast<uint32_t*>(base + 2 * C) = pack_bf16x2(dv[dn][2], dv[dn][3]);
      }
    }
  }
}

// Forward attention core with the same building blocks (the training forward keeps Q|K|V row-major, which the tcgen05 core of
// inference -- V^T produced by the in_proj epilogue -- does not): a warp per 16-query tile, pass 1 = log-sum-exp of S = Q K^T,
// pass 2 = O = P V with P = exp(scale S - lse) formed straight from the accumulator tiles.
template <int HD>
__global__ void __launch_bounds__(128) sdpa_fwd_mma_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int n_groups, int L, int Lp, int C,
                                                           int heads, int gpb) {
  constexpr int STR = HD + 8, KS = HD / 16, DN = HD / 8, VPR = HD / 8;
  extern __shared__ __align__(16) unsigned char smraw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const size_t region = (size_t)3 * Lp * STR * sizeof(bf16);
  const int ld = 3 * C;
  const float scale = rsqrtf((float)HD);
  for (int idx = tid; idx < gpb * 3 * Lp * VPR; idx += 128) {
    const int v = idx % VPR;
    int r = idx / VPR;
    const int row = r % Lp; r /= Lp;
    const int arr = r % 3, gl = r / 3;
    const int grp = blockIdx.x * gpb + gl;
    bf16* dst = reinterpret_cast<bf16*>(smraw + gl * region) + ((size_t)arr * Lp + row) * STR + v * 8;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (grp < n_groups && row < L) {
      const int b = grp / heads, h = grp - b * heads;
      val = *reinterpret_cast<const uint4*>(qkv + ((size_t)b * L + row) * ld + arr * C + h * HD + v * 8);
    }
    *reinterpret_cast<uint4*>(dst) = val;
  }
  __syncthreads();
  const int wpg = 4 / gpb;
  const int gl = warp / wpg, wl = warp - gl * wpg;
  const int grp = blockIdx.x * gpb + gl;
  if (grp >= n_groups) return;
  const bf16* Qs = reinterpret_cast<const bf16*>(smraw + gl * region);
  const bf16* Ks = Qs + (size_t)Lp * STR;
  const bf16* Vs = Ks + (size_t)Lp * STR;
  const int b = grp / heads, h = grp - b * heads;
  const int ntiles = Lp / 16;
  for (int it = wl; it < ntiles; it += wpg) {
    const int r0 = it * 16;
    uint32_t qa[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) ldsm_x4(qa[ks], Qs + (size_t)(r0 + (lane & 15)) * STR + ks * 16 + (lane >> 4) * 8);
    float m0 = -INFINITY, l0 = 0.f, m1 = -INFINITY, l1 = 0.f;
    for (int nt = 0; nt < Lp / 8; ++nt) {
      float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t kb[2];
        ldsm_x2(kb, Ks + (size_t)(nt * 8 + (lane & 7)) * STR + ks * 16 + ((lane >> 3) & 1) * 8);
        mma16816(c, qa[ks], kb);
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool ok = nt * 8 + 2 * t + e < L;
        const float s0 = ok ? c[e] * scale : -INFINITY, s1 = ok ? c[2 + e] * scale : -INFINITY;
        if (s0 > m0) { l0 = l0 * __expf(m0 - s0) + 1.f; m0 = s0; } else if (ok) l0 += __expf(s0 - m0);
        if (s1 > m1) { l1 = l1 * __expf(m1 - s1) + 1.f; m1 = s1; } else if (ok) l1 += __expf(s1 - m1);
      }
    }
#pragma unroll
    for (int off = 1; off <= 2; off <<= 1) {
      const float om0 = __shfl_xor_sync(0xffffffffu, m0, off), ol0 = __shfl_xor_sync(0xffffffffu, l0, off);
      const float om1 = __shfl_xor_sync(0xffffffffu, m1, off), ol1 = __shfl_xor_sync(0xffffffffu, l1, off);
      const float nm0 = fmaxf(m0, om0), nm1 = fmaxf(m1, om1);
      l0 = (m0 == -INFINITY ? 0.f : l0 * __expf(m0 - nm0)) + (om0 == -INFINITY ? 0.f : ol0 * __expf(om0 - nm0));
      l1 = (m1 == -INFINITY ? 0.f : l1 * __expf(m1 - nm1)) + (om1 == -INFINITY ? 0.f : ol1 * __expf(om1 - nm1));
      m0 = nm0; m1 = nm1;
    }
    const float lse0 = m0 + __logf(l0), lse1 = m1 + __logf(l1);
    float oacc[DN][4];
#pragma unroll
    for (int dn = 0; dn < DN; ++dn) { oacc[dn][0] = 0.f; oacc[dn][1] = 0.f; oacc[dn][2] = 0.f; oacc[dn][3] = 0.f; }
    for (int np = 0; np < ntiles; ++np) {
      float pp[2][4];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int nt = np * 2 + hh;
        float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint32_t kb[2];
          ldsm_x2(kb, Ks + (size_t)(nt * 8 + (lane & 7)) * STR + ks * 16 + ((lane >> 3) & 1) * 8);
          mma16816(c, qa[ks], kb);
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const bool ok = nt * 8 + 2 * t + e < L;
          pp[hh][e] = ok ? __expf(c[e] * scale - lse0) : 0.f;
          pp[hh][2 + e] = ok ? __expf(c[2 + e] * scale - lse1) : 0.f;
        }
      }
      uint32_t pa[4] = {pack_bf16x2(pp[0][0], pp[0][1]), pack_bf16x2(pp[0][2], pp[0][3]), pack_bf16x2(pp[1][0], pp[1][1]),
                        pack_bf16x2(pp[1][2], pp[1][3])};
#pragma unroll
      for (int dn = 0; dn < DN; ++dn) {
        uint32_t vb[2];
        ldsm_x2_trans(vb, Vs + (size_t)(np * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * STR + dn * 8);
        mma16816(oacc[dn], pa, vb);
      }
    }
    const int row_a = r0 + g, row_b = r0 + g + 8;
#pragma unroll
    for (int dn = 0; dn < DN; ++dn) {
      const int col = h * HD + dn * 8 + 2 * t;
      if (row_a < L) *reinterpret_cast<uint32_t*>(out + ((size_t)b * L + row_a) * C + col) = pack_bf16x2(oacc[dn][0], oacc[dn][1]);
      if (row_b < L) *reinterpret_cast<uint32_t*>(out + ((size_t)b * L + row_b) * C + col) = pack_bf16x2(oacc[dn][2], oacc[dn][3]);
    }
  }
}
template <int HD> bool launch_sdpa_fwd_mma_t(const bf16* qkv, bf16* out, int B, int L, int C, int heads, cudaStream_t s) {
  const int Lp = (L + 15) / 16 * 16;
  const int gpb = Lp <= 16 ? 4 : 1;
  const size_t region = (size_t)3 * Lp * (HD + 8) * sizeof(bf16);
  const size_t smem = region * gpb;
  if (smem > 200 * 1024) return false;
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(sdpa_fwd_mma_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr_set = true; }
  const int n_groups = B * heads;
  sdpa_fwd_mma_kernel<HD><<<(n_groups + gpb - 1) / gpb, 128, smem, s>>>(qkv, out, n_groups, L, Lp, C, heads, gpb);
  return true;
}

template <int HD> bool launch_sdpa_bwd_mma(const bf16* qkv, const bf16* o, const bf16* d_o, bf16* d_qkv, int B, int L, int C, int heads, cudaStream_t s) {
  const int Lp = (L + 15) / 16 * 16;
  const int gpb = Lp <= 16 ? 4 : 1;
  const size_t region = (size_t)4 * Lp * (HD + 8) * sizeof(bf16) + (size_t)2 * Lp * sizeof(float);
  const size_t smem = region * gpb;
  if (smem > 200 * 1024 || region % 16) return false;
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(sdpa_bwd_mma_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr_set = true; }
  const int n_groups = B * heads;
  sdpa_bwd_mma_kernel<HD><<<(n_groups + gpb - 1) / gpb, 128, smem, s>>>(qkv, o, d_o, d_qkv, n_groups, L, Lp, C, heads, gpb);
  return true;
}
}  // namespace
template <typename T> void launch_sdpa_bwd(const T* qkv, const T* o, const T* d_o, T* d_qkv, int B, int L, int C, int heads, cudaStream_t s) {
  const int hd = C / heads;
  if constexpr (sizeof(T) == 2) {
    static int simt = -1;  // SPDM_SDPA_BWD_SIMT=1: CUDA-core kernel on the bf16 path too (A/B switch)
    if (simt < 0) { const char* e = getenv("SPDM_SDPA_BWD_SIMT"); simt = e ? atoi(e) : 0; }
    if (!simt) {
      bool done = false;
      if (hd == 16) done = launch_sdpa_bwd_mma<16>(qkv, o, d_o, d_qkv, B, L, C, heads, s);
      else if (hd == 32) done = launch_sdpa_bwd_mma<32>(qkv, o, d_o, d_qkv, B, L, C, heads, s);
      else if (hd == 64) done = launch_sdpa_bwd_mma<64>(qkv, o, d_o, d_qkv, B, L, C, heads, s);
      if (done) { COUNT_LAUNCH(); return; }
    }
  }
  const size_t smem = ((size_t)4 * L * hd + 2 * L) * sizeof(float);
#define SDPA_BWD_CASE(HD)                                                                                       \
  {                                                                                                             \
    static bool attr_set = false;                                                                               \
    if (!attr_set) {                                                                                            \
      cudaFuncSetAttribute(sdpa_bwd_kernel<T, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);    \
      attr_set = true;                                                                                          \
    }                                                                                                           \
    sdpa_bwd_kernel<T, HD><<<B * heads, 128, smem, s>>>(qkv, o, d_o, d_qkv, L, C, heads);                       \
  }
  if (hd == 16) SDPA_BWD_CASE(16)
  else if (hd == 32) SDPA_BWD_CASE(32)
  else if (hd == 64) SDPA_BWD_CASE(64)
#undef SDPA_BWD_CASE
  COUNT_LAUNCH();
}
template void launch_sdpa_bwd<float>(const float*, const float*, const float*, float*, int, int, int, int, cudaStream_t);
template void launch_sdpa_bwd<bf16>(const bf16*, const bf16*, const bf16*, bf16*, int, int, int, int, cudaStream_t);
template <typename T> void launch_gelu_fwd(const T* x, T* y, long long n, cudaStream_t s) {
  gelu_fwd_kernel<T><<<cdiv(n / 8, 256), 256, 0, s>>>(x, y, n / 8);
  COUNT_LAUNCH();
}
template <typename T> void launch_gelu_bwd(const T* dy, const T* pre, T* dx, long long n, cudaStream_t s) {
  gelu_bwd_kernel<T><<<cdiv(n / 8, 256), 256, 0, s>>>(dy, pre, dx, n / 8);
  COUNT_LAUNCH();
}
template void launch_gelu_fwd<float>(const float*, float*, long long, cudaStream_t);
template void launch_gelu_fwd<bf16>(const bf16*, bf16*, long long, cudaStream_t);
template void launch_gelu_bwd<float>(const float*, const float*, float*, long long, cudaStream_t);
template void launch_gelu_bwd<bf16>(const bf16*, const bf16*, bf16*, long long, cudaStream_t);

// =================================================================================================
// MSELoss + outc backward (models/diffusion_ddpm.py:171; Unet_FiLmLayer.py:264,310-311): one block per sample.
//   e = 2 (eps_hat - noise) / N on the unpadded window, zero on the pad ring
//   d_act[px][c] = e[px] * w[c] ;  d_w[c] += sum_px e[px] act[px][c] ;  d_b += sum e ;  loss += sum (eps_hat-noise)^2 / N
// inc.first weight gradient (:101 with pad_to :15-34 folded in): one block per sample.
// =================================================================================================
namespace {
template <typename T>
__global__ void __launch_bounds__(256) mse_outc_bwd_kernel(const float* __restrict__ eps_hat, const float* __restrict__ noise, const T* __restrict__ act,
                                                           int ld, const float* __restrict__ w, T* __restrict__ d_act, float* __restrict__ d_w,
                                                           float* __restrict__ d_b, float* __restrict__ loss, float inv_n, int H, int W, int C,
                                                           int rows, int dim, int lh, int lw, int B_valid) {
  extern __shared__ float se[];  // [H*W] e per padded pixel, then [256*8] reduction scratch
  float* red = se + H * W;
  __shared__ float scratch[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  float lsum = 0.f, esum = 0.f;
  for (int px = tid; px < H * W; px += 256) {
    const int hh = px / W - lh, ww = px % W - lw;
    float e = 0.f;
    if (b < B_valid && hh >= 0 && hh < rows && ww >= 0 && ww < dim) {  // samples past B_valid pad a ragged batch: no loss, no gradient
      const size_t i = ((size_t)b * rows + hh) * dim + ww;
      const float d = eps_hat[i] - noise[i];
      lsum = fmaf(d, d, lsum);
      e = 2.f * d * inv_n;
      esum += e;
    }
    se[px] = e;
  }
  lsum = block_sum_256(lsum, scratch);
  esum = block_sum_256(esum, scratch);
  if (tid == 0) { atomicAdd(loss, lsum * inv_n); atomicAdd(d_b, esum); }
  const int vpr = C >> 3, c8 = (tid % vpr) << 3, row_step = 256 / vpr;
  float wv[8], acc[8];
  load8(w + c8, wv);
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int px = tid / vpr; px < H * W; px += row_step) {
    const float e = se[px];
    float a8[8], o[8];
    load8(act + ((size_t)b * H * W + px) * ld + c8, a8);
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[i] = e * wv[i]; acc[i] = fmaf(e, a8[i], acc[i]); }
    store8(d_act + ((size_t)b * H * W + px) * C + c8, o);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) red[tid * 8 + i] = acc[i];
  __syncthreads();
  if (tid < vpr) {
    float t[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < row_step; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] += red[(r * vpr + tid) * 8 + i];
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(d_w + c8 + i, t[i]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) conv_in_wgrad_kernel(const float* __restrict__ x, const T* __restrict__ d_raw, float* __restrict__ dw, int H,
                                                            int W, int rows, int dim, int lh, int lw) {
  extern __shared__ float sx[];   // [rows*dim] sample, then [64*9] block accumulator
  float* sacc = sx + rows * dim;
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < rows * dim; i += 256) sx[i] = x[(size_t)b * rows * dim + i];
  for (int i = tid; i < 64 * 9; i += 256) sacc[i] = 0.f;
  __syncthreads();
  const int c8 = (tid & 7) << 3;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[t][i] = 0.f;
  for (int px = tid >> 3; px < H * W; px += 32) {
    const int hh = px / W, ww = px - hh * W;
    float d[8];
    load8(d_raw + ((size_t)b * H * W + px) * 64 + c8, d);
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int sh = hh + tap / 3 - 1 - lh, sw_ = ww + tap % 3 - 1 - lw;
      if (sh < 0 || sh >= rows || sw_ < 0 || sw_ >= dim) continue;
      const float xv = sx[sh * dim + sw_];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[tap][i] = fmaf(xv, d[i], acc[tap][i]);
    }
  }
#pragma unroll
  for (int tap = 0; tap < 9; ++tap)
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(&sacc[(c8 + i) * 9 + tap], acc[tap][i]);
  __syncthreads();
  for (int i = tid; i < 64 * 9; i += 256) atomicAdd(dw + i, sacc[i]);  // (64, 1, 3, 3) is [c][tap]
}
}  // namespace
template <typename T>
void launch_mse_outc_bwd(const float* eps_hat, const float* noise, const T* act, int ld, const float* w, T* d_act, float* d_w, float* d_b,
                         float* loss, int B, int H, int W, int C, int rows, int dim, int lh, int lw, cudaStream_t s, int B_valid) {
  const size_t smem = ((size_t)H * W + 256 * 8) * sizeof(float);
  if (B_valid <= 0 || B_valid > B) B_valid = B;
  const float inv_n = 1.0f / ((float)B_valid * rows * dim);   // MSELoss 'mean' over the real samples (ddpm:171)
  mse_outc_bwd_kernel<T><<<B, 256, smem, s>>>(eps_hat, noise, act, ld, w, d_act, d_w, d_b, loss, inv_n, H, W, C, rows, dim, lh, lw, B_valid);
  COUNT_LAUNCH();
}
template <typename T>
void launch_conv_in_wgrad(const float* x, const T* d_raw, float* dw, int B, int H, int W, int rows, int dim, int lh, int lw, cudaStream_t s) {
  const size_t smem = ((size_t)rows * dim + 64 * 9) * sizeof(float);
  conv_in_wgrad_kernel<T><<<B, 256, smem, s>>>(x, d_raw, dw, H, W, rows, dim, lh, lw);
  COUNT_LAUNCH();
}
template void launch_mse_outc_bwd<float>(const float*, const float*, const float*, int, const float*, float*, float*, float*, float*, int, int, int,
                                         int, int, int, int, int, cudaStream_t, int);
template void launch_mse_outc_bwd<bf16>(const float*, const float*, const bf16*, int, const float*, bf16*, float*, float*, float*, int, int, int,
                                        int, int, int, int, int, cudaStream_t, int);
template void launch_conv_in_wgrad<float>(const float*, const float*, float*, int, int, int, int, int, int, int, cudaStream_t);
template void launch_conv_in_wgrad<bf16>(const float*, const bf16*, float*, int, int, int, int, int, int, int, cudaStream_t);

// =================================================================================================
// Conditioning: silu(pos_encoding(t)) rows (Unet_FiLmLayer.py:266-274,136-142), Mish backward (:150),
// gather of the image-feature gradient out of d obs_cond (models/diffusion_ddpm.py:317-330).
// =================================================================================================
namespace {
__global__ void posenc_silu_kernel(const long long* __restrict__ t_dev, const float* __restrict__ inv_freq, float* __restrict__ out, int time_dim) {
  const int row = blockIdx.x;
  const float t = (float)t_dev[row];
  const int half = time_dim >> 1;
  for (int i = threadIdx.x; i < time_dim; i += blockDim.x) {
    const float arg = t * inv_freq[i < half ? i : i - half];
    const float v = i < half ? sinf(arg) : cosf(arg);
    out[(size_t)row * time_dim + i] = v / (1.f + expf(-v));
  }
}
__global__ void mish_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  const float sp = v > 20.f ? v : log1pf(expf(v));
  const float th = tanhf(sp);
  const float sig = 1.f / (1.f + expf(-v));
  dx[i] = dy[i] * (th + v * (1.f - th * th) * sig);
}
__global__ void gather_feat_grad_kernel(const float* __restrict__ d_cond, float* __restrict__ d_feat, long long total, int cond_dim) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int f = cond_dim - 7;
  const long long bt = i / f;
  const int j = (int)(i - bt * f);
  d_feat[i] = d_cond[bt * cond_dim + 7 + j];
}
}  // namespace
void launch_posenc_silu(const long long* t, int n, const float* inv_freq, float* out, int time_dim, cudaStream_t s) {
  posenc_silu_kernel<<<n, 256, 0, s>>>(t, inv_freq, out, time_dim);
  COUNT_LAUNCH();
}
void launch_mish_bwd(const float* dy, const float* x, float* dx, long long n, cudaStream_t s) {
  mish_bwd_kernel<<<cdiv(n, 256), 256, 0, s>>>(dy, x, dx, n);
  COUNT_LAUNCH();
}
namespace {
// FiLM Linears on the tensor cores (bf16 training): operand packs and the Mish on either side of the GEMMs.
//   wf16 [1792][GP]  = the six (2C, G) PyTorch weights stacked, columns padded with zeros to GP = ceil64(G): B operand of the forward GEMM
//   wb16 [GP][1792]  = its transpose: B operand of the data-gradient GEMM
__global__ void film_pack16_kernel(const float* __restrict__ src, bf16* __restrict__ wf, bf16* __restrict__ wb, int C2, int G, int GP, int off) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)C2 * G) return;
  const int j = (int)(i / G), g = (int)(i - (long long)j * G);
  const bf16 v = __float2bfloat16_rn(src[i]);
  wf[(size_t)(off + j) * GP + g] = v;
  wb[(size_t)g * SPDM_FILM_WIDTH + off + j] = v;
}
__global__ void mish_pad_bf16_kernel(const float* __restrict__ cond, bf16* __restrict__ out, int B, int G, int GP, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int b = (int)(i / GP), g = (int)(i - (long long)b * GP);
  float y = 0.f;
  if (b < B && g < G) {
    const float x = cond[(size_t)b * G + g];
    const float sp = x > 20.f ? x : log1pf(expf(x));
    y = x * tanhf(sp);
  }
  out[i] = __float2bfloat16_rn(y);
}
__global__ void mish_bwd_bf16_kernel(const bf16* __restrict__ dy, int ld_dy, const float* __restrict__ x, float* __restrict__ dx, int G, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = (int)(i / G), g = (int)(i - (long long)b * G);
  const float v = x[i];
  const float sp = v > 20.f ? v : log1pf(expf(v));
  const float th = tanhf(sp);
  const float sig = 1.f / (1.f + expf(-v));
  dx[i] = __bfloat162float(dy[(size_t)b * ld_dy + g]) * (th + v * (1.f - th * th) * sig);
}
}  // namespace
void launch_film_pack16(const float* src, bf16* wf, bf16* wb, int C2, int G, int GP, int off, cudaStream_t s) {
  film_pack16_kernel<<<cdiv((long long)C2 * G, 256), 256, 0, s>>>(src, wf, wb, C2, G, GP, off);
}
void launch_mish_pad_bf16(const float* cond, bf16* out, int B, int Bpad, int G, int GP, cudaStream_t s) {
  const long long total = (long long)Bpad * GP;
  mish_pad_bf16_kernel<<<cdiv(total, 256), 256, 0, s>>>(cond, out, B, G, GP, total);
  COUNT_LAUNCH();
}
void launch_mish_bwd_bf16(const bf16* dy, int ld_dy, const float* x, float* dx, int B, int G, cudaStream_t s) {
  const long long n = (long long)B * G;
  mish_bwd_bf16_kernel<<<cdiv(n, 256), 256, 0, s>>>(dy, ld_dy, x, dx, G, n);
  COUNT_LAUNCH();
}
void launch_gather_feat_grad(const float* d_cond, float* d_feat, int B, int T, int cond_dim, cudaStream_t s) {
  const long long total = (long long)B * T * (cond_dim - 7);
  gather_feat_grad_kernel<<<cdiv(total, 256), 256, 0, s>>>(d_cond, d_feat, total, cond_dim);
  COUNT_LAUNCH();
}

// =================================================================================================
// Vision encoder conv stack backward (models/encoder/autoencoder.py:11-17).  k == stride == 2: the three layers are
// non-overlapping patch contractions, so the backward is strip-local too.  A block walks frames; per 8-row input strip
// it recomputes conv1/conv2 (+ReLU) in shared memory exactly like enc_convs_kernel, then back-propagates
// d feat (12 x 64) -> d c2 (32 x 2 x 24) -> d c1 (16 x 4 x 48), accumulating every weight gradient in thread-owned
// registers across all strips and frames of the block; one atomicAdd per element per block at the end.
// =================================================================================================
namespace {
__global__ void __launch_bounds__(256) enc_convs_bwd_kernel(const float* __restrict__ img, const float* __restrict__ w1, const float* __restrict__ b1,
                                                            const float* __restrict__ w2t, const float* __restrict__ b2, const float* __restrict__ w3t,
                                                            const float* __restrict__ feat, const float* __restrict__ d_feat,
                                                            float* __restrict__ dw1, float* __restrict__ db1, float* __restrict__ dw2,
                                                            float* __restrict__ db2, float* __restrict__ dw3, float* __restrict__ db3, int n) {
  extern __shared__ float es[];
  float* w3s = es;                   // [128][64]  ((c*4+kk), o)
  float* w2s = w3s + 128 * 64;       // [64][32]
  float* w1s = w2s + 64 * 32;        // [12][16]
  float* s_in = w1s + 12 * 16;       // [3][8][97]
  float* s_c1 = s_in + 3 * 8 * 97;   // [16][4][48]  post-ReLU
  float* s_c2 = s_c1 + 16 * 4 * 48;  // [32][2][24]  post-ReLU
  float* s_d3 = s_c2 + 32 * 2 * 24;  // [12][64]     gradient of pre-ReLU conv3
  float* s_d2 = s_d3 + 12 * 64;      // [32][2][24]  gradient of pre-ReLU conv2
  float* s_d1 = s_d2 + 32 * 2 * 24;  // [16][4][48]  gradient of pre-ReLU conv1
  const int tid = threadIdx.x;
  for (int e = tid; e < 128 * 64; e += 256) w3s[e] = __ldg(w3t + e);
  for (int e = tid; e < 64 * 32; e += 256) w2s[e] = __ldg(w2t + e);
  for (int e = tid; e < 12 * 16; e += 256) w1s[e] = __ldg(w1 + (e % 16) * 12 + e / 16);
  const int ch1 = tid & 15, ch2 = tid & 31, ch3 = tid & 63;
  const float bias1 = __ldg(b1 + ch1), bias2 = __ldg(b2 + ch2);
  // thread-owned weight-gradient accumulators
  float a3[32];   // dW3 element (o = tid%64, ck = tid/64 + 4r)
  float a2[8];    // dW2 element (o2 = tid%32, ck = tid/32 + 8r), ck = c1*4+kk in [0,64)
  float a1 = 0.f; // dW1 element tid < 192: (o1 = tid%16, ck = tid/16), ck = c*4+kk in [0,12)
  float ab = 0.f; // bias gradient: tid < 64 -> db3[tid]; 64..95 -> db2; 96..111 -> db1
#pragma unroll
  for (int r = 0; r < 32; ++r) a3[r] = 0.f;
#pragma unroll
  for (int r = 0; r < 8; ++r) a2[r] = 0.f;

  for (int frame = blockIdx.x; frame < n; frame += gridDim.x) {
    const float* im = img + (size_t)frame * 3 * 96 * 96;
    for (int i = 0; i < 12; ++i) {
      __syncthreads();
      for (int e = tid; e < 3 * 8 * 97; e += 256) {
        const int col = e % 97 - 1;
        const int r = (e / 97) % 8;
        const int c = e / (97 * 8);
        const int gr = 8 * i - 1 + r;
        float v = 0.f;
        if (gr >= 0 && gr < 96 && col >= 0 && col < 96) v = im[((size_t)c * 96 + gr) * 96 + col];
        s_in[e] = v;
      }
      for (int e = tid; e < 12 * 64; e += 256) {  // d conv3 (pre-ReLU) = d feat where feat > 0
        const size_t gi = (size_t)frame * 9216 + (size_t)i * 12 * 64 + e;
        s_d3[e] = feat[gi] > 0.f ? d_feat[gi] : 0.f;
      }
      __syncthreads();
      // ---- recompute conv1 ----
#pragma unroll 4
      for (int k = 0; k < 12; ++k) {
        const int pos = (tid >> 4) + 16 * k;
        const int rr = pos / 48, cc = pos - rr * 48;
        float acc = bias1;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            acc = fmaf(s_in[(c * 8 + 2 * rr + (kk >> 1)) * 97 + 2 * cc + (kk & 1)], w1s[(c * 4 + kk) * 16 + ch1], acc);
        s_c1[(ch1 * 4 + rr) * 48 + cc] = fmaxf(acc, 0.f);
      }
      __syncthreads();
      // ---- recompute conv2 ----
      for (int k = 0; k < 6; ++k) {
        const int pos = (tid >> 5) + 8 * k;
        const int rr = pos / 24, cc = pos - rr * 24;
        float acc = bias2;
        for (int c = 0; c < 16; ++c)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            acc = fmaf(s_c1[(c * 4 + 2 * rr + (kk >> 1)) * 48 + 2 * cc + (kk & 1)], w2s[(c * 4 + kk) * 32 + ch2], acc);
        s_c2[(ch2 * 2 + rr) * 24 + cc] = fmaxf(acc, 0.f);
      }
      __syncthreads();
      // ---- dW3 / db3 ----
      {
        const int o = ch3;
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
          const int ck = (tid >> 6) + 4 * r, c = ck >> 2, kk = ck & 3;
          const float* c2p = s_c2 + (c * 2 + (kk >> 1)) * 24 + (kk & 1);
          float acc = a3[r];
#pragma unroll
          for (int j = 0; j < 12; ++j) acc = fmaf(s_d3[j * 64 + o], c2p[2 * j], acc);
          a3[r] = acc;
        }
        if (tid < 64) {
#pragma unroll
          for (int j = 0; j < 12; ++j) ab += s_d3[j * 64 + tid];
        }
      }
      // ---- d c2 = W3^T d3, masked by c2 > 0 ----
      for (int k = 0; k < 6; ++k) {
        const int pos = (tid >> 5) + 8 * k;   // 48 positions: rr = pos / 24, cc = pos % 24
        const int rr = pos / 24, cc = pos - rr * 24;
        const int j = cc >> 1, kk = rr * 2 + (cc & 1);
        const float* wp = w3s + (ch2 * 4 + kk) * 64;
        float acc = 0.f;
#pragma unroll 8
        for (int o = 0; o < 64; ++o) acc = fmaf(s_d3[j * 64 + o], wp[o], acc);
        const int idx = (ch2 * 2 + rr) * 24 + cc;
        s_d2[idx] = s_c2[idx] > 0.f ? acc : 0.f;
      }
      __syncthreads();
      // ---- dW2 / db2 ----
      {
        const int o2 = ch2;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int ck = (tid >> 5) + 8 * r, c = ck >> 2, kk = ck & 3;
          float acc = a2[r];
          for (int pos = 0; pos < 48; ++pos) {
            const int rr = pos / 24, cc = pos - rr * 24;
            acc = fmaf(s_d2[(o2 * 2 + rr) * 24 + cc], s_c1[(c * 4 + 2 * rr + (kk >> 1)) * 48 + 2 * cc + (kk & 1)], acc);
          }
          a2[r] = acc;
        }
        if (tid >= 64 && tid < 96) {
          const int o = tid - 64;
          for (int pos = 0; pos < 48; ++pos) ab += s_d2[o * 48 + pos];
        }
      }
      // ---- d c1 = W2^T d2, masked by c1 > 0 ----
      for (int k = 0; k < 12; ++k) {
        const int pos = (tid >> 4) + 16 * k;  // 192 positions: rr = pos / 48, cc = pos % 48
        const int rr = pos / 48, cc = pos - rr * 48;
        const int r2 = rr >> 1, c2 = cc >> 1, kk = (rr & 1) * 2 + (cc & 1);
        const float* wp = w2s + (ch1 * 4 + kk) * 32;
        float acc = 0.f;
#pragma unroll 8
        for (int o = 0; o < 32; ++o) acc = fmaf(s_d2[(o * 2 + r2) * 24 + c2], wp[o], acc);
        const int idx = (ch1 * 4 + rr) * 48 + cc;
        s_d1[idx] = s_c1[idx] > 0.f ? acc : 0.f;
      }
      __syncthreads();
      // ---- dW1 / db1 ----
      if (tid < 192) {
        const int o1 = tid & 15, ck = tid >> 4, c = ck >> 2, kk = ck & 3;
        float acc = a1;
        for (int pos = 0; pos < 192; ++pos) {
          const int rr = pos / 48, cc = pos - rr * 48;
          acc = fmaf(s_d1[(o1 * 4 + rr) * 48 + cc], s_in[(c * 8 + 2 * rr + (kk >> 1)) * 97 + 2 * cc + (kk & 1)], acc);
        }
        a1 = acc;
      }
      if (tid >= 96 && tid < 112) {
        const int o = tid - 96;
        for (int pos = 0; pos < 192; ++pos) ab += s_d1[o * 192 + pos];
      }
    }
  }
  // ---- flush: PyTorch layouts (64,32,2,2) = [o][ck], (32,16,2,2) = [o2][ck], (16,3,2,2) = [o1][ck] ----
#pragma unroll
  for (int r = 0; r < 32; ++r) atomicAdd(dw3 + ch3 * 128 + (tid >> 6) + 4 * r, a3[r]);
#pragma unroll
  for (int r = 0; r < 8; ++r) atomicAdd(dw2 + ch2 * 64 + (tid >> 5) + 8 * r, a2[r]);
  if (tid < 192) atomicAdd(dw1 + (tid & 15) * 12 + (tid >> 4), a1);
  if (tid < 64) atomicAdd(db3 + tid, ab);
  else if (tid < 96) atomicAdd(db2 + tid - 64, ab);
  else if (tid < 112) atomicAdd(db1 + tid - 96, ab);
}

__global__ void enc_linear_grad_permute_kernel(const float* __restrict__ tmp, float* __restrict__ dwl) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 128LL * 9216) return;
  const int k2 = (int)(i % 9216);   // hwc index p*64 + c
  const int nrow = (int)(i / 9216);
  const int c = k2 % 64, p = k2 / 64;
  dwl[(size_t)nrow * 9216 + c * 144 + p] += tmp[i];
}
__global__ void pack_enc_linear_t_kernel(const float* __restrict__ w, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 128LL * 9216) return;
  const int k2 = (int)(i % 9216);
  const int nrow = (int)(i / 9216);
  const int c = k2 % 64, p = k2 / 64;
  out[i] = w[(size_t)nrow * 9216 + c * 144 + p];
}
}  // namespace
void launch_enc_convs_bwd(const float* img, const float* w1, const float* b1, const float* w2t, const float* b2, const float* w3t,
                          const float* feat, const float* d_feat, float* dw1, float* db1, float* dw2, float* db2, float* dw3, float* db3,
                          int n, cudaStream_t s) {
  constexpr size_t smem = (128 * 64 + 64 * 32 + 12 * 16 + 3 * 8 * 97 + 2 * (16 * 4 * 48) + 2 * (32 * 2 * 24) + 12 * 64) * sizeof(float);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(enc_convs_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
  int grid = n < 148 * 2 ? n : 148 * 2;
  enc_convs_bwd_kernel<<<grid, 256, smem, s>>>(img, w1, b1, w2t, b2, w3t, feat, d_feat, dw1, db1, dw2, db2, dw3, db3, n);
  COUNT_LAUNCH();
}
void launch_enc_linear_grad_permute(const float* tmp_hwc, float* dwl, cudaStream_t s) {
  enc_linear_grad_permute_kernel<<<cdiv(128LL * 9216, 256), 256, 0, s>>>(tmp_hwc, dwl);
  COUNT_LAUNCH();
}
void launch_pack_enc_linear_t(const float* w, float* out, cudaStream_t s) {
  pack_enc_linear_t_kernel<<<cdiv(128LL * 9216, 256), 256, 0, s>>>(w, out);
}

// =================================================================================================
// dgrad weight packs: d x = conv(d y, W') with W'[tap'][co][ci] = W[co][ci][8 - tap'] (3x3, pad 1) and
// d x = d y W for a Linear y = x W^T.
// =================================================================================================
namespace {
__global__ void pack_conv_dgrad_f32_kernel(const float* __restrict__ oihw, float* __restrict__ out, int Cout, int Cin) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Cout * Cin * 9) return;
  const int ci = (int)(i % Cin);          // out index = (tap' * Cout + co) * Cin + ci
  const long long t = i / Cin;
  const int co = (int)(t % Cout);
  const int tp = (int)(t / Cout);
  out[i] = oihw[((size_t)co * Cin + ci) * 9 + (8 - tp)];
}
__global__ void pack_conv_dgrad_bf16_kernel(const float* __restrict__ oihw, bf16* __restrict__ out, int Cout, int Cin) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Cout * Cin * 9) return;
  const int co = (int)(i % Cout);         // out index = (ci * 9 + tap') * Cout + co
  const long long t = i / Cout;
  const int tp = (int)(t % 9);
  const int ci = (int)(t / 9);
  out[i] = __float2bfloat16_rn(oihw[((size_t)co * Cin + ci) * 9 + (8 - tp)]);
}
__global__ void pack_linear_dgrad_bf16_kernel(const float* __restrict__ nk, bf16* __restrict__ out, int N, int K) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * K) return;
  const int nn = (int)(i % N);            // out [K][N]
  const int k = (int)(i / N);
  out[i] = __float2bfloat16_rn(nk[(size_t)nn * K + k]);
}
}  // namespace
namespace {
__global__ void unpack_conv_grad_kernel(const float* __restrict__ packed, float* __restrict__ dst, int Cout, int Cin) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // index into packed [tap][co][ci]
  if (i >= (long long)Cout * Cin * 9) return;
  const int ci = (int)(i % Cin);
  const long long t = i / Cin;
  const int co = (int)(t % Cout);
  const int tap = (int)(t / Cout);
  dst[((size_t)co * Cin + ci) * 9 + tap] = packed[i];
}
}  // namespace
void launch_unpack_conv_grad(const float* packed, float* dst, int Cout, int Cin, cudaStream_t s) {
  unpack_conv_grad_kernel<<<cdiv((long long)Cout * Cin * 9, 256), 256, 0, s>>>(packed, dst, Cout, Cin);
  COUNT_LAUNCH();
}
void launch_pack_conv_dgrad_f32(const float* oihw, float* out, int Cout, int Cin, cudaStream_t s) {
  pack_conv_dgrad_f32_kernel<<<cdiv((long long)Cout * Cin * 9, 256), 256, 0, s>>>(oihw, out, Cout, Cin);
}
void launch_pack_conv_dgrad_bf16(const float* oihw, bf16* out, int Cout, int Cin, cudaStream_t s) {
  pack_conv_dgrad_bf16_kernel<<<cdiv((long long)Cout * Cin * 9, 256), 256, 0, s>>>(oihw, out, Cout, Cin);
}
void launch_pack_linear_dgrad_bf16(const float* nk, bf16* out, int N, int K, cudaStream_t s) {
  pack_linear_dgrad_bf16_kernel<<<cdiv((long long)N * K, 256), 256, 0, s>>>(nk, out, N, K);
}

// =================================================================================================
// Optimizer: clip_grad_norm_ (Lightning gradient_clip_val, train.py:107) + torch.optim.Adam (ddpm:115-125).
// =================================================================================================
namespace {
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  __shared__ float scratch[8];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) acc = fmaf(g[i], g[i], acc);
  acc = block_sum_256(acc, scratch);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                   long long n, float lr, float beta1, float beta2, float eps, float bc1, float bc2_sqrt,
                                                   const float* __restrict__ sumsq, float max_norm, float grad_scale) {
  float coef = grad_scale;
  if (sumsq) {
    const float total = sqrtf(*sumsq) * grad_scale;
    float c = max_norm / (total + 1e-6f);
    if (c > 1.f) c = 1.f;
    coef *= c;
  }
  const float step_size = lr / bc1;
  auto upd = [&](float& pi, float& gi, float& mi, float& vi) {
    gi *= coef;
    mi = beta1 * mi + (1.f - beta1) * gi;
    vi = beta2 * vi + (1.f - beta2) * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= step_size * (mi / denom);
  };
  // the step streams 32 bytes per parameter (read and write p, g, m, v): 16-byte accesses, the tail element-wise
  const long long n4 = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) ? 0 : n / 4;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    float4 P = reinterpret_cast<float4*>(p)[i], Gv = reinterpret_cast<float4*>(g)[i], Mv = reinterpret_cast<float4*>(m)[i], Vv = reinterpret_cast<float4*>(v)[i];
    upd(P.x, Gv.x, Mv.x, Vv.x); upd(P.y, Gv.y, Mv.y, Vv.y); upd(P.z, Gv.z, Mv.z, Vv.z); upd(P.w, Gv.w, Mv.w, Vv.w);
    reinterpret_cast<float4*>(p)[i] = P; reinterpret_cast<float4*>(g)[i] = Gv; reinterpret_cast<float4*>(m)[i] = Mv; reinterpret_cast<float4*>(v)[i] = Vv;
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    float pi = p[i], gi = g[i], mi = m[i], vi = v[i];
    upd(pi, gi, mi, vi);
    p[i] = pi; g[i] = gi; m[i] = mi; v[i] = vi;
  }
}
}  // namespace
void launch_sumsq(const float* g, long long n, float* out, cudaStream_t s) {
  int grid = cdiv(n, 256 * 8);
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  sumsq_kernel<<<grid, 256, 0, s>>>(g, n, out);
  COUNT_LAUNCH();
}
void launch_adam(float* p, float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps, int step,
                 const float* sumsq, float max_norm, float grad_scale, cudaStream_t s) {
  int grid = cdiv(n, 256 * 4);
  if (grid > 148 * 16) grid = 148 * 16;
  if (grid < 1) grid = 1;
  const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  adam_kernel<<<grid, 256, 0, s>>>(p, g, m, v, n, lr, beta1, beta2, eps, bc1, bc2_sqrt, sumsq, max_norm, grad_scale);
  COUNT_LAUNCH();
}

// =================================================================================================
// Vision encoder on the tensor cores (bf16 training path; models/encoder/autoencoder.py:11-20).
// k == stride == 2 makes conv2 / conv3 plain GEMMs once the activations are stored patch-major:
//   c1p [M2 = n*576][64]  : row = ((frame*144 + p3)*4 + kk3) = one 2x2 patch of conv1 pixels, col = kk2*16 + ch
//   c2  [M2/2][64] : row = a PAIR of conv1 patches (kk3 = 2j, 2j+1), col = (kk3 % 2)*32 + ch  ==  [M3 = n*144][128] : row = conv3
//                    patch p3, col = kk3*32 + ch   (conv2 runs two patches per GEMM row against a block-diagonal weight)
//   feat [M3][64] == [n][9216] in (pixel, channel) order
// so conv2 = [M2/2][128] c1p @ W2bd^T (128 -> 64), conv3 = c2 @ W3p^T (128 -> 64), both through conv_tc.cu with a
// bias + ReLU epilogue, and their backward is wgrad_tc.cu + the same GEMM with transposed weights.  What stays on CUDA
// cores: conv1 (3 -> 16, K = 12), its weight gradient, the ReLU masks and the tiny weight (un)packs.
// =================================================================================================
namespace {
// Stage the 8-row input strip of conv3 row r3 into shared memory as [3*8 rows][100]: image column j sits at index 4 + j
// (16-byte aligned float4 copies), index 3 is the zero padding column -1, row 0 of strip 0 is the zero padding row -1.
__device__ __forceinline__ void enc_stage_strip(const float* __restrict__ im, int r3, float* s_in, int tid) {
  for (int i = tid; i < 24 * 24; i += 192) {
    const int row = i / 24, q4 = i - row * 24;
    const int c = row >> 3, r8 = row & 7;
    const int gr = 8 * r3 - 1 + r8;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gr >= 0) v = __ldg(reinterpret_cast<const float4*>(im + ((size_t)c * 96 + gr) * 96) + q4);
    *reinterpret_cast<float4*>(s_in + row * 100 + 4 + 4 * q4) = v;
  }
  if (tid < 24) s_in[tid * 100 + 3] = 0.f;
}

// The same strip from a uint8 HWC frame (what the simulator / dataset stores): 8 rows x 288 bytes, decoded x / 255 (one IEEE
// division, the reference's `/ 255.0`) while staging; a thread handles 3 of the 576 32-bit words.
__device__ __forceinline__ void enc_stage_strip_u8(const uint8_t* __restrict__ im, int r3, float* s_in, int tid) {
  for (int i = tid; i < 8 * 72; i += 192) {
    const int r8 = i / 72, wq = i - r8 * 72;
    const int gr = 8 * r3 - 1 + r8;
    uint32_t w = 0;
    if (gr >= 0) w = __ldg(reinterpret_cast<const uint32_t*>(im + (size_t)gr * 288) + wq);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = 4 * wq + j, x = k / 3, c = k - 3 * x;
      s_in[(c * 8 + r8) * 100 + 4 + x] = gr >= 0 ? __fdiv_rn((float)((w >> (8 * j)) & 255u), 255.0f) : 0.f;
    }
  }
  if (tid < 24) s_in[tid * 100 + 3] = 0.f;
}

// one block per (frame, conv3 row r3): 192 conv1 pixels, one thread each, 16 channels per thread
template <bool U8>
__global__ void __launch_bounds__(192) enc_conv1_fwd_kernel(const void* __restrict__ img, const float* __restrict__ w1, const float* __restrict__ b1,
                                                            bf16* __restrict__ c1p, int T, long long bstride) {
  __shared__ __align__(16) float s_in[24 * 100];
  __shared__ __align__(16) float w1s[12 * 16];
  __shared__ __align__(16) float b1s[16];
  const int tid = threadIdx.x;
  const int frame = blockIdx.x / 12, r3 = blockIdx.x % 12;
  w1s[tid] = __ldg(w1 + (tid & 15) * 12 + (tid >> 4));
  if (tid < 16) b1s[tid] = __ldg(b1 + tid);
  if constexpr (U8) {
    enc_stage_strip_u8(reinterpret_cast<const uint8_t*>(img) + (size_t)frame * 96 * 96 * 3, r3, s_in, tid);   // contiguous (B*T, 96, 96, 3)
  } else {
    // frame = b*T + t; samples bstride apart
    enc_stage_strip(reinterpret_cast<const float*>(img) + (size_t)(frame / T) * bstride + (size_t)(frame % T) * 3 * 96 * 96, r3, s_in, tid);
  }
  __syncthreads();
  // tid = (c3*4 + kk3)*4 + kk2 : consecutive threads write consecutive 32-byte pixel slots of c1p
  const int kk2 = tid & 3, kk3 = (tid >> 2) & 3, c3 = tid >> 4;
  const int yl = 2 * (kk3 >> 1) + (kk2 >> 1);          // conv1 row inside the strip (0..3)
  const int x = 2 * (2 * c3 + (kk3 & 1)) + (kk2 & 1);  // conv1 column (0..47)
  float acc[16];
#pragma unroll
  for (int o = 0; o < 16; ++o) acc[o] = b1s[o];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float xv = s_in[(c * 8 + 2 * yl + (kk >> 1)) * 100 + 3 + 2 * x + (kk & 1)];
      const float4* wp = reinterpret_cast<const float4*>(w1s + (c * 4 + kk) * 16);
#pragma unroll
      for (int o4 = 0; o4 < 4; ++o4) {
        const float4 w = wp[o4];
        acc[o4 * 4 + 0] = fmaf(xv, w.x, acc[o4 * 4 + 0]);
        acc[o4 * 4 + 1] = fmaf(xv, w.y, acc[o4 * 4 + 1]);
        acc[o4 * 4 + 2] = fmaf(xv, w.z, acc[o4 * 4 + 2]);
        acc[o4 * 4 + 3] = fmaf(xv, w.w, acc[o4 * 4 + 3]);
      }
    }
#pragma unroll
  for (int o = 0; o < 16; ++o) acc[o] = fmaxf(acc[o], 0.f);
  bf16* dst = c1p + ((size_t)blockIdx.x * 192 + tid) * 16;  // (frame*144 + r3*12 + c3)*4*64 + kk3*64 + kk2*16
  store8(dst, acc);
  store8(dst + 8, acc + 8);
}

// dW1 (16,3,2,2) and db1 from d1 [M2][64] (gradient of the pre-ReLU conv1 output, same layout as c1p).
// A block walks (frame, r3) strips.  The kernel is bound by shared-memory reads, so every thread owns a 2 (o) x 4 (kh,kw)
// register tile of one input channel and one eighth of the strip's pixels: 1 LDS.64 + 4 LDS per 8 FMAs.
// thread = pixel group pg (8) x input channel c (3) x output pair op (8); partial tiles are folded through shared memory once,
// at the end of the kernel, then one atomicAdd per weight per block.
// `act` (optional): the conv1 output c1p; d1 is then the gradient of the POST-ReLU output and the ReLU mask (act > 0) is applied
// on load -- saves the separate relu_mask pass over d1 (read d1 + read c1p + write d1, ~175 us at 5120 frames).
__global__ void __launch_bounds__(192) enc_conv1_wgrad_kernel(const float* __restrict__ img, const bf16* __restrict__ d1, const bf16* __restrict__ act,
                                                              float* __restrict__ dw1, float* __restrict__ db1, int n_strips, int T, long long bstride) {
  __shared__ __align__(16) float s_in[24 * 100];
  constexpr int SDS = 193;                        // padded row stride of s_d: [16 channels][192 raster pixels]
  __shared__ __align__(16) float s_d[16 * SDS];
  const int tid = threadIdx.x;
  const int op = tid & 7, c = (tid >> 3) % 3, pg = tid / 24;   // outputs 2*op, 2*op+1; channel c; pixels pg, pg+8, ...
  const int kk2 = tid & 3, kk3 = (tid >> 2) & 3, c3 = tid >> 4;
  const int my_yl = 2 * (kk3 >> 1) + (kk2 >> 1), my_x = 2 * (2 * c3 + (kk3 & 1)) + (kk2 & 1);
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, accb[2] = {0.f, 0.f};
  for (int strip = blockIdx.x; strip < n_strips; strip += gridDim.x) {
    const int frame = strip / 12, r3 = strip % 12;
    __syncthreads();
    enc_stage_strip(img + (size_t)(frame / T) * bstride + (size_t)(frame % T) * 3 * 96 * 96, r3, s_in, tid);
    {
      float t8[8], a8[8];
      const bf16* src = d1 + ((size_t)strip * 192 + tid) * 16;
      const bf16* asrc = act ? act + ((size_t)strip * 192 + tid) * 16 : nullptr;
      float* dst = s_d + my_yl * 48 + my_x;   // channel-major: a warp's stores of one channel hit (nearly) distinct banks
      load8(src, t8);
      if (asrc) {
        load8(asrc, a8);
#pragma unroll
        for (int i = 0; i < 8; ++i) t8[i] = a8[i] > 0.f ? t8[i] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) dst[i * SDS] = t8[i];
      load8(src + 8, t8);
      if (asrc) {
        load8(asrc + 8, a8);
#pragma unroll
        for (int i = 0; i < 8; ++i) t8[i] = a8[i] > 0.f ? t8[i] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) dst[(8 + i) * SDS] = t8[i];
    }
    __syncthreads();
#pragma unroll
    for (int yl = 0; yl < 4; ++yl) {
      const float* r0 = s_in + (c * 8 + 2 * yl) * 100 + 3;
      const float* r1 = r0 + 100;
      const float* dd = s_d + (2 * op) * SDS + yl * 48;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int x = pg + 8 * j;
        const float2 d = make_float2(dd[x], dd[SDS + x]);
        const float i00 = r0[2 * x], i01 = r0[2 * x + 1], i10 = r1[2 * x], i11 = r1[2 * x + 1];
        acc[0][0] = fmaf(d.x, i00, acc[0][0]); acc[0][1] = fmaf(d.x, i01, acc[0][1]);
        acc[0][2] = fmaf(d.x, i10, acc[0][2]); acc[0][3] = fmaf(d.x, i11, acc[0][3]);
        acc[1][0] = fmaf(d.y, i00, acc[1][0]); acc[1][1] = fmaf(d.y, i01, acc[1][1]);
        acc[1][2] = fmaf(d.y, i10, acc[1][2]); acc[1][3] = fmaf(d.y, i11, acc[1][3]);
        accb[0] += d.x;
        accb[1] += d.y;
      }
    }
  }
  // fold the 8 pixel groups: s_d is free now
  __syncthreads();
  float* red = s_d;                       // [pg][16 o][12 ck] then [pg][16] bias partials at 8*192
#pragma unroll
  for (int e = 0; e < 2; ++e)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) red[pg * 192 + (2 * op + e) * 12 + c * 4 + kk] = acc[e][kk];
  if (c == 0) { red[8 * 192 + pg * 16 + 2 * op] = accb[0]; red[8 * 192 + pg * 16 + 2 * op + 1] = accb[1]; }
  __syncthreads();
  {
    float t = 0.f;
#pragma unroll
    for (int g8 = 0; g8 < 8; ++g8) t += red[g8 * 192 + tid];
    atomicAdd(dw1 + tid, t);              // (16,3,2,2) is [o][c*4+kk]
    if (tid < 16) {
      float tb = 0.f;
#pragma unroll
      for (int g8 = 0; g8 < 8; ++g8) tb += red[8 * 192 + g8 * 16 + tid];
      atomicAdd(db1 + tid, tb);
    }
  }
}

// The same gradient on the tensor cores (mma.sync m16n8k16, bf16 operands, fp32 accumulate): per strip
//   D[16 o][16 n] += A[16 o][192 pixels] * B[192 pixels][16 n],   n = c*4 + kh*2 + kw (12 columns), n = 12 a column of ones (the
// bias gradient), 13..15 zero.  A is the (masked) d1 strip as it lies in memory ([pixel][16 o] bf16, read with ldmatrix.trans),
// B is gathered from the fp32 input strip in shared memory (one scalar load per element, rounded to bf16).  The CUDA-core
// kernel above issues 144 shared-memory loads per warp and strip, this one 36; it is then bound by the HBM reads of the
// frames and of d1 / c1p.  A warp owns 2 of the 12 16-pixel blocks of every strip and keeps its accumulators for the whole
// kernel; the 6 warps are folded through shared memory at the end, then one atomicAdd per weight per block.
__global__ void __launch_bounds__(192) enc_conv1_wgrad_mma_kernel(const float* __restrict__ img, const bf16* __restrict__ d1, const bf16* __restrict__ act,
                                                                  float* __restrict__ dw1, float* __restrict__ db1, int n_strips, int T, long long bstride) {
  __shared__ __align__(16) float s_in[24 * 100];
  constexpr int DS = 24;                          // padded row stride (bf16) of s_d: 48-byte rows, conflict-free ldmatrix
  __shared__ __align__(16) bf16 s_d[192 * DS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  // B columns of this thread: n = g (tile 0) and 8 + g (tile 1); n -> (c, kh, kw) -> offset of the tap inside the strip
  const int n1 = 8 + g;
  const int off0 = ((g >> 2) * 8 + ((g >> 1) & 1)) * 100 + 3 + (g & 1);
  const int off1 = ((n1 >> 2) * 8 + ((n1 >> 1) & 1)) * 100 + 3 + (n1 & 1);   // used when n1 < 12
  const float fill1 = g == 4 ? 1.f : 0.f;
  for (int strip = blockIdx.x; strip < n_strips; strip += gridDim.x) {
    const int frame = strip / 12, r3 = strip % 12;
    __syncthreads();
    enc_stage_strip(img + (size_t)(frame / T) * bstride + (size_t)(frame % T) * 3 * 96 * 96, r3, s_in, tid);
    {
      const uint4* src = reinterpret_cast<const uint4*>(d1 + ((size_t)strip * 192 + tid) * 16);
      uint4 v[2] = {src[0], src[1]};
      if (act) {
        const uint4* as = reinterpret_cast<const uint4*>(act + ((size_t)strip * 192 + tid) * 16);
        const uint4 a[2] = {as[0], as[1]};
        const __nv_bfloat162 zero = __floats2bfloat162_rn(0.f, 0.f);
        __nv_bfloat162* vv = reinterpret_cast<__nv_bfloat162*>(v);
        const __nv_bfloat162* aa = reinterpret_cast<const __nv_bfloat162*>(a);
#pragma unroll
        for (int i = 0; i < 8; ++i) vv[i] = __hmul2(vv[i], __hgt2(aa[i], zero));   // ReLU mask: x 1.0 or x 0.0
      }
      *reinterpret_cast<uint4*>(s_d + tid * DS) = v[0];
      *reinterpret_cast<uint4*>(s_d + tid * DS + 8) = v[1];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const int k0 = (warp * 2 + kk) * 16;
      uint32_t a[4];
      const int mi = lane >> 3;
      ldsm_x4_trans(a, s_d + (k0 + 8 * (mi >> 1) + (lane & 7)) * DS + 8 * (mi & 1));
      float b0[4], b1[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = k0 + 2 * t + (e & 1) + 8 * (e >> 1);           // pixel of the strip, memory order of c1p / d1
        const int kk2 = j & 3, kk3 = (j >> 2) & 3, c3 = j >> 4;
        const int yl = 2 * (kk3 >> 1) + (kk2 >> 1), x = 2 * (2 * c3 + (kk3 & 1)) + (kk2 & 1);
        const int base = 2 * yl * 100 + 2 * x;
        b0[e] = s_in[off0 + base];
        b1[e] = g < 4 ? s_in[off1 + base] : fill1;
      }
      const uint32_t f0[2] = {pack_bf16x2(b0[0], b0[1]), pack_bf16x2(b0[2], b0[3])};
      const uint32_t f1[2] = {pack_bf16x2(b1[0], b1[1]), pack_bf16x2(b1[2], b1[3])};
      mma16816(acc[0], a, f0);
      mma16816(acc[1], a, f1);
    }
  }
  __syncthreads();
  float* red = s_in;                              // [6 warps][16 o][16 n]
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    red[(warp * 16 + g) * 16 + 8 * j + 2 * t] = acc[j][0];
    red[(warp * 16 + g) * 16 + 8 * j + 2 * t + 1] = acc[j][1];
    red[(warp * 16 + g + 8) * 16 + 8 * j + 2 * t] = acc[j][2];
    red[(warp * 16 + g + 8) * 16 + 8 * j + 2 * t + 1] = acc[j][3];
  }
  __syncthreads();
  for (int idx = tid; idx < 256; idx += 192) {
    float tsum = 0.f;
#pragma unroll
    for (int w = 0; w < 6; ++w) tsum += red[w * 256 + idx];
    const int o = idx >> 4, n = idx & 15;
    if (n < 12) atomicAdd(dw1 + o * 12 + n, tsum);   // (16,3,2,2) is [o][c*4 + kh*2 + kw]
    else if (n == 12) atomicAdd(db1 + o, tsum);
  }
}

// out = act > 0 ? d : 0   (ReLU backward) over [rows][64] bf16, in place allowed; optionally colsum[64] += column sums of the
// masked gradient (the conv bias gradient).  Grid-stride with a stride that keeps every thread on the same 8 columns.
__global__ void __launch_bounds__(256) relu_mask_kernel(const bf16* __restrict__ d, const bf16* __restrict__ act, bf16* __restrict__ out,
                                                        long long nvec, float* __restrict__ colsum) {
  __shared__ float red[256 * 8];
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long v = (long long)blockIdx.x * 256 + threadIdx.x; v < nvec; v += (long long)gridDim.x * 256) {
    float a[8], g[8];
    load8(act + v * 8, a);
    load8(d + v * 8, g);
#pragma unroll
    for (int i = 0; i < 8; ++i) { g[i] = a[i] > 0.f ? g[i] : 0.f; acc[i] += g[i]; }
    store8(out + v * 8, g);
  }
  if (!colsum) return;
#pragma unroll
  for (int i = 0; i < 8; ++i) red[threadIdx.x * 8 + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < 64) {  // column = (thread % 8) * 8 + i
    const int cg8 = threadIdx.x >> 3, i = threadIdx.x & 7;
    float t = 0.f;
    for (int r = 0; r < 32; ++r) t += red[(r * 8 + cg8) * 8 + i];
    atomicAdd(colsum + cg8 * 8 + i, t);
  }
}

// weight packs of the patch-GEMM encoder (see header of this section).  conv2 handles TWO conv1 patches per GEMM row with a
// block-diagonal weight, so that its 32 output channels fill the 64-wide tile without zero padding:
//   W2bd[pp*32 + o2][pp'*64 + kk2*16 + c1] = (pp == pp') * W2[o2][c1][kk2]     (64 x 128)
//   W3p[o][kk3*32 + c2] = W3[o][c2][kk3]                                         (64 x 128)
__global__ void enc_pack_w2_kernel(const float* __restrict__ w2, bf16* __restrict__ w2p, bf16* __restrict__ w2pT) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 128) return;
  const int ob = i / 128, kb = i % 128;
  const int pp = ob / 32, o2 = ob % 32, pq = kb / 64, k = kb % 64, kk2 = k / 16, c1 = k % 16;
  const float v = pp == pq ? w2[o2 * 64 + c1 * 4 + kk2] : 0.f;
  w2p[i] = __float2bfloat16_rn(v);
  w2pT[kb * 64 + ob] = __float2bfloat16_rn(v);
}
__global__ void enc_pack_w3_kernel(const float* __restrict__ w3, bf16* __restrict__ w3p, bf16* __restrict__ w3pT) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 128) return;
  const int o = i / 128, kb = i % 128, kk3 = kb / 32, c2 = kb % 32;
  const float v = w3[o * 128 + c2 * 4 + kk3];
  w3p[i] = __float2bfloat16_rn(v);
  w3pT[kb * 64 + o] = __float2bfloat16_rn(v);
}
__global__ void enc_pack_b2_kernel(const float* __restrict__ b2, float* __restrict__ b2p) {
  if (threadIdx.x < 64) b2p[threadIdx.x] = b2[threadIdx.x % 32];
}
// (128, 9216 chw) fp32 -> bf16 [9216 hwc][128]  (B operand of d feat = d enc_out @ Wl)
__global__ void enc_pack_linear_T16_kernel(const float* __restrict__ w, bf16* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 128LL * 9216) return;
  const int nrow = (int)(i % 128);
  const int k2 = (int)(i / 128);
  const int c = k2 % 64, p = k2 / 64;
  out[i] = __float2bfloat16_rn(w[(size_t)nrow * 9216 + c * 144 + p]);
}
// patch-GEMM weight gradients -> PyTorch layouts (accumulating): g2 = d W2bd (64 x 128, the two diagonal blocks are summed,
// the off-diagonal blocks are cross terms between neighbouring patches and are dropped), g3 = d W3p (64 x 128)
__global__ void enc_unpack_grads_kernel(const float* __restrict__ g2, const float* __restrict__ g3, const float* __restrict__ gb2 /*[64]*/,
                                        float* __restrict__ dw2, float* __restrict__ db2, float* __restrict__ dw3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 32 * 64) {   // dw2 (32,16,2,2): o2*64 + c1*4 + kk2
    const int o2 = i / 64, r = i % 64, c1 = r / 4, kk2 = r % 4;
    dw2[i] += g2[o2 * 128 + kk2 * 16 + c1] + g2[(32 + o2) * 128 + 64 + kk2 * 16 + c1];
  }
  if (i < 64 * 128) {  // dw3 (64,32,2,2): o*128 + c2*4 + kk3
    const int o = i / 128, r = i % 128, c2 = r / 4, kk3 = r % 4;
    dw3[i] += g3[o * 128 + kk3 * 32 + c2];
  }
  if (i < 32) db2[i] += gb2[i] + gb2[32 + i];
}
}  // namespace
void launch_enc_conv1_fwd(const float* img, const float* w1, const float* b1, bf16* c1p, int n, int T, long long bstride, cudaStream_t s) {
  enc_conv1_fwd_kernel<false><<<n * 12, 192, 0, s>>>(img, w1, b1, c1p, T, bstride);
  COUNT_LAUNCH();
}
void launch_enc_conv1_fwd_u8(const uint8_t* img_hwc, const float* w1, const float* b1, bf16* c1p, int n, cudaStream_t s) {
  enc_conv1_fwd_kernel<true><<<n * 12, 192, 0, s>>>(img_hwc, w1, b1, c1p, 1, 0);
  COUNT_LAUNCH();
}
void launch_enc_conv1_wgrad(const float* img, const bf16* d1, const bf16* act, float* dw1, float* db1, int n, int T, long long bstride, cudaStream_t s) {
  int grid = n * 12;
  if (grid > 148 * 8) grid = 148 * 8;
  static int simt = -1;   // SPDM_ENC_W1_SIMT=1: the CUDA-core kernel (A/B switch)
  if (simt < 0) { const char* e = getenv("SPDM_ENC_W1_SIMT"); simt = e ? atoi(e) : 0; }
  if (simt) enc_conv1_wgrad_kernel<<<grid, 192, 0, s>>>(img, d1, act, dw1, db1, n * 12, T, bstride);
  else enc_conv1_wgrad_mma_kernel<<<grid, 192, 0, s>>>(img, d1, act, dw1, db1, n * 12, T, bstride);
  COUNT_LAUNCH();
}
void launch_relu_mask(const bf16* d, const bf16* act, bf16* out, long long n, float* colsum64, cudaStream_t s) {
  int grid = cdiv(n / 8, 256 * 4);
  if (grid > 148 * 16) grid = 148 * 16;
  if (grid < 1) grid = 1;
  relu_mask_kernel<<<grid, 256, 0, s>>>(d, act, out, n / 8, colsum64);
  COUNT_LAUNCH();
}
void launch_enc_pack_w2(const float* w2, bf16* w2p, bf16* w2pT, cudaStream_t s) { enc_pack_w2_kernel<<<cdiv(64 * 128, 256), 256, 0, s>>>(w2, w2p, w2pT); }
void launch_enc_pack_w3(const float* w3, bf16* w3p, bf16* w3pT, cudaStream_t s) { enc_pack_w3_kernel<<<cdiv(64 * 128, 256), 256, 0, s>>>(w3, w3p, w3pT); }
void launch_enc_pack_b2(const float* b2, float* b2p, cudaStream_t s) { enc_pack_b2_kernel<<<1, 64, 0, s>>>(b2, b2p); }
void launch_enc_pack_linear_T16(const float* w, bf16* out, cudaStream_t s) {
  enc_pack_linear_T16_kernel<<<cdiv(128LL * 9216, 256), 256, 0, s>>>(w, out);
}
void launch_enc_unpack_grads(const float* g2, const float* g3, const float* gb2, float* dw2, float* db2, float* dw3, cudaStream_t s) {
  enc_unpack_grads_kernel<<<cdiv(64 * 128, 256), 256, 0, s>>>(g2, g3, gb2, dw2, db2, dw3);
  COUNT_LAUNCH();
}

// bf16 attention core forward on mma.sync (training forward); returns false when the shape does not fit shared memory
bool launch_sdpa_fwd_mma(const bf16* qkv, bf16* out, int B, int L, int C, int heads, cudaStream_t s) {
  const int hd = C / heads;
  bool done = false;
  if (hd == 16) done = launch_sdpa_fwd_mma_t<16>(qkv, out, B, L, C, heads, s);
  else if (hd == 32) done = launch_sdpa_fwd_mma_t<32>(qkv, out, B, L, C, heads, s);
  else if (hd == 64) done = launch_sdpa_fwd_mma_t<64>(qkv, out, B, L, C, heads, s);
  if (done) COUNT_LAUNCH();
  return done;
}
