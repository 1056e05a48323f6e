// kernels.cu — CUDA-core kernels of the denoising hot path (sm_100a).
//
// Everything that is not a dense bf16 contraction lives here: the fp32 reference-precision implicit
// GEMM (parity path and odd shapes), GroupNorm statistics / apply (+GELU, +time embedding, +FiLM),
// max-pool, bilinear upsample, LayerNorm, the small-L multi-head attention core, the first/last
// 1-channel convolutions, the observation encoder, and the fused posterior update.
// Reference lines are cited per kernel (paths relative to the reference repo root).
#include "common.cuh"

static long long g_launches = 0;
int g_spdm_pdl = 1;
long long kernels_launch_count() { return g_launches; }
void kernels_count_launch() { ++g_launches; }  // launches issued by sdpa_tc.cu / attn_tc.cu / attn_head.cu
#define COUNT_LAUNCH() (++g_launches)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// =================================================================================================
// fp32-accumulate implicit GEMM on CUDA cores.
//   out[r, n] = epi( sum_{tap, c} in[shift(r, tap), c] * w[tap][c][n] )
// nn.Conv2d(k=3, pad=1, bias=False)  (models/Unet_FiLmLayer.py:101,103) when taps == 9,
// nn.Linear / 1x1 when taps == 1.  64x64 tile, BK=16, 256 threads, 4x4 outputs per thread.
// =================================================================================================
namespace {
constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16;

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmSimtArgs a) {
  pdl_wait();
  pdl_trigger();
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Bs[SG_BK][SG_BN + 4];
  const TI* __restrict__ in = reinterpret_cast<const TI*>(a.in);
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * SG_BM, n0 = blockIdx.y * SG_BN;
  const int tx = tid & 15, ty = tid >> 4;

  // A-load assignment: one row, 4 consecutive k
  const int a_row = tid >> 2, a_k = (tid & 3) * 4;
  const int r = m0 + a_row;
  int rb = 0, rh = 0, rw = 0;
  const int HW = a.H * a.W;
  if (a.taps == 9) { rb = r / HW; const int rem = r - rb * HW; rh = rem / a.W; rw = rem - rh * a.W; }
  (void)rb;
  // B-load assignment: one k row, 4 consecutive n
  const int b_k = tid >> 4, b_n = (tid & 15) * 4;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const bool a_vec = (sizeof(TI) == 4) && (a.Cin % 4 == 0) && (a.ld_in % 4 == 0);
  const bool b_vec = (a.Cout % 4 == 0);

  for (int tap = 0; tap < a.taps; ++tap) {
    int dy = 0, dx = 0;
    if (a.taps == 9) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
    bool valid = r < a.M;
    long long src = r;
    if (a.taps == 9) {
      const int hh = rh + dy, ww = rw + dx;
      valid = valid && hh >= 0 && hh < a.H && ww >= 0 && ww < a.W;
      src = (long long)r + dy * a.W + dx;
    }
    if (a.taps == 9) {  // whole-block skip of structurally empty taps (W == 1 or H == 1)
      if ((a.W == 1 && dx != 0) || (a.H == 1 && dy != 0)) continue;
    }
    const float* __restrict__ wt = a.w + (size_t)tap * a.Cin * a.Cout;
    for (int c0 = 0; c0 < a.Cin; c0 += SG_BK) {
      // ---- load A tile (transposed into As[k][row]) ----
      float av[4] = {0.f, 0.f, 0.f, 0.f};
      if (valid) {
        const TI* p = in + src * a.ld_in + c0 + a_k;
        if (a_vec && c0 + a_k + 3 < a.Cin) {
          const float4 t = *reinterpret_cast<const float4*>(p);
          av[0] = t.x; av[1] = t.y; av[2] = t.z; av[3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (c0 + a_k + i < a.Cin) av[i] = to_f32<TI>(p[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) As[a_k + i][a_row] = av[i];
      // ---- load B tile ----
      float bv[4] = {0.f, 0.f, 0.f, 0.f};
      if (c0 + b_k < a.Cin) {
        const float* p = wt + (size_t)(c0 + b_k) * a.Cout + n0 + b_n;
        if (b_vec && n0 + b_n + 3 < a.Cout) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(p));
          bv[0] = t.x; bv[1] = t.y; bv[2] = t.z; bv[3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (n0 + b_n + i < a.Cout) bv[i] = __ldg(p + i);
        }
      }
      *reinterpret_cast<float4*>(&Bs[b_k][b_n]) = make_float4(bv[0], bv[1], bv[2], bv[3]);
      __syncthreads();
#pragma unroll
      for (int k = 0; k < SG_BK; ++k) {
        const float4 av4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 bv4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float ar[4] = {av4.x, av4.y, av4.z, av4.w};
        const float br[4] = {bv4.x, bv4.y, bv4.z, bv4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  // ---- epilogue ----
  TO* __restrict__ out = reinterpret_cast<TO*>(a.out);
  const TO* __restrict__ resid = reinterpret_cast<const TO*>(a.resid);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= a.Cout) continue;
      float v = acc[i][j];
      if (a.bias) v += __ldg(a.bias + n);
      if (a.act == ACT_GELU) v = gelu_exact(v);
      else if (a.act == ACT_RELU) v = fmaxf(v, 0.f);
      if (resid) v += to_f32<TO>(resid[(size_t)row * a.ld_res + n]);
      out[(size_t)row * a.ld_out + n] = from_f32<TO>(v);
    }
  }
}
}  // namespace

template <typename TI, typename TO> void launch_gemm_simt(const GemmSimtArgs& a, cudaStream_t s) {
  dim3 grid(cdiv(a.M, SG_BM), cdiv(a.Cout, SG_BN));
  launch_pdl(gemm_simt_kernel<TI, TO>, dim3(grid), dim3(256), 0, s, a);
  COUNT_LAUNCH();
}
template void launch_gemm_simt<float, float>(const GemmSimtArgs&, cudaStream_t);
template void launch_gemm_simt<bf16, bf16>(const GemmSimtArgs&, cudaStream_t);
template void launch_gemm_simt<float, bf16>(const GemmSimtArgs&, cudaStream_t);
template void launch_gemm_simt<bf16, float>(const GemmSimtArgs&, cudaStream_t);

// =================================================================================================
// GroupNorm(1, C) statistics (models/Unet_FiLmLayer.py:105): per-sample (sum, sumsq) over C*H*W.
// One block per sample, P = 1 partial.  (The tcgen05 conv writes its partials from its epilogue.)
// =================================================================================================
namespace {
template <typename T>
__global__ void __launch_bounds__(256) stats_kernel(const T* __restrict__ raw, float* __restrict__ stats, int HW, int C, int ld) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x;
  const int vec_per_row = C >> 3;
  const int nvec = HW * vec_per_row;
  float s = 0.f, q = 0.f;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    const int row = v / vec_per_row, c8 = (v - row * vec_per_row) << 3;
    float x[8];
    load8(raw + ((size_t)b * HW + row) * ld + c8, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += x[i]; q = fmaf(x[i], x[i], q); }
  }
  __shared__ float ss[8], sq[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  if ((threadIdx.x & 31) == 0) { ss[threadIdx.x >> 5] = s; sq[threadIdx.x >> 5] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ts = 0.f, tq = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { ts += ss[i]; tq += sq[i]; }
    stats[2 * b] = ts;
    stats[2 * b + 1] = tq;
  }
}
}  // namespace
template <typename T> void launch_stats(const T* raw, float* stats, int B, int HW, int C, int ld, cudaStream_t s) {
  launch_pdl(stats_kernel<T>, dim3(B), dim3(256), 0, s, raw, stats, HW, C, ld);
  COUNT_LAUNCH();
}
template void launch_stats<float>(const float*, float*, int, int, int, int, cudaStream_t);
template void launch_stats<bf16>(const bf16*, float*, int, int, int, int, cudaStream_t);

// =================================================================================================
// GroupNorm apply (+GELU) (+ time embedding) (+ FiLM):
//   y = (x - mean) * rstd * gamma[c] + beta[c]           models/Unet_FiLmLayer.py:112,115
//   y = gelu(y)                                           :113       (first conv of a DoubleConvolution)
//   y = y + temb[c]                                       :165-168   (end of a Down/Up stage)
//   y = scale[b,c] * y + bias[b,c]                        :171-177   (FiLM)
// grid (chunks, B); every block first folds the P partial sums of its sample (in double).
// =================================================================================================
namespace {
constexpr int APPLY_THREADS = 128;

// erf to ~1.5e-7 absolute (Abramowitz-Stegun 7.1.26): enough for the bf16 path, a third of erff's cost
__device__ __forceinline__ float erf_fast(float x) {
  const float ax = fabsf(x);
  const float t = __frcp_rn(fmaf(0.3275911f, ax, 1.0f));
  const float poly = t * (0.254829592f + t * (-0.284496736f + t * (1.421413741f + t * (-1.453152027f + t * 1.061405429f))));
  return copysignf(1.0f - poly * __expf(-ax * ax), x);
}

// GELU for the bf16 path: y * Phi(y) with Phi through the hardware tanh (one MUFU + 5 FMA-class instructions per element instead
// of the ~18 of the erf polynomial: the GELU launches of apply_kernel were issue-bound, ncu smsp__issue_active 59-69 %, twice the
// time of the GELU-free launches on the same bytes).  |gelu_tanh - gelu_erf| <= 5e-4 absolute, an eighth of the bf16 rounding step
// of the stored result; the fp32 parity path (EXACT) keeps erff.
__device__ __forceinline__ float gelu_tanh_fast(float y) {
  const float u = 0.7978845608028654f * fmaf(0.044715f * y * y, y, y);
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(u));
  return fmaf(0.5f * y, th, 0.5f * y);
}

template <typename TI, typename TO, bool EXACT, int APPLY_VEC_PER_THREAD>
__global__ void __launch_bounds__(APPLY_THREADS) apply_kernel(ApplyArgs a) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y;
  // every thread folds the P partial sums of its sample itself (broadcast loads, no barrier)
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

  const TI* __restrict__ raw = reinterpret_cast<const TI*>(a.raw);
  TO* __restrict__ out = reinterpret_cast<TO*>(a.out);
  const int vec_per_row = a.C >> 3;  // a power of two <= APPLY_THREADS: a thread keeps the same 8 channels
  const int nvec = a.HW * vec_per_row;
  const int c8 = (threadIdx.x & (vec_per_row - 1)) << 3;
  float g[8], be[8], te[8], fs[8], fb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { te[i] = 0.f; fs[i] = 1.f; fb[i] = 0.f; }
  load8(a.gamma + c8, g);
  load8(a.beta + c8, be);
  const bool has_temb = a.temb_mode != TEMB_NONE, has_film = a.film != nullptr;
  if (has_temb) {
    int trow = 0;
    if (a.temb_mode == TEMB_PER_SAMPLE) trow = b;
    else if (a.temb_mode == TEMB_STEP) trow = *a.step_ptr + a.step_off;
    load8(a.temb + (size_t)trow * SPDM_TEMB_WIDTH + a.temb_off + c8, te);
  }
  if (has_film) {
    const float* film = a.film + (size_t)b * SPDM_FILM_WIDTH + a.film_off;
    load8(film + c8, fs);
    load8(film + a.C + c8, fb);
  }
  if (!EXACT) {  // fold the normalisation into one FMA per element
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float ga = rstd * g[i]; be[i] = fmaf(-mean, ga, be[i]); g[i] = ga; }
  }
  const int v0 = blockIdx.x * (APPLY_THREADS * APPLY_VEC_PER_THREAD);
#pragma unroll
  for (int it = 0; it < APPLY_VEC_PER_THREAD; ++it) {
    const int v = v0 + it * APPLY_THREADS + threadIdx.x;
    if (v >= nvec) break;
    const int row = v / vec_per_row;
    float x[8];
    load8(raw + ((size_t)b * a.HW + row) * a.ld_in + c8, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y;
      if (EXACT) y = (x[i] - mean) * rstd * g[i] + be[i];
      else y = fmaf(x[i], g[i], be[i]);
      if (a.act == ACT_GELU) {
        if (EXACT) y = gelu_exact(y);
        else y = gelu_tanh_fast(y);
      }
      if (has_temb) y += te[i];
      if (has_film) y = fs[i] * y + fb[i];
      x[i] = y;
    }
    store8(out + ((size_t)b * a.HW + row) * a.ld_out + c8, x);
  }
}
}  // namespace
namespace {
// Split-K reduction + GroupNorm(1, C) + apply in one pass: a block owns one sample, whose HW*C values (<= 16384) stay
// in registers between the statistics pass and the normalisation pass.  Output bf16.
template <int EPT>  // float4 groups per thread: HW*C = 256 threads * EPT * 4
__global__ void __launch_bounds__(256) apply_partial_kernel(ApplyArgs a, const float* __restrict__ partial, int S, long long M) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = a.HW * a.C;
  const size_t sample0 = (size_t)b * n;              // partial tiles are [M][C] with ld == C: a sample is contiguous
  const size_t slice = (size_t)M * a.C;
  float4 x[EPT];
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int k = 0; k < EPT; ++k) {
    const int idx = (k * 256 + tid) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (idx < n) {
      for (int sp = 0; sp < S; ++sp) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(partial + sp * slice + sample0 + idx));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    x[k] = acc;
    s += (acc.x + acc.y) + (acc.z + acc.w);
    q = fmaf(acc.x, acc.x, q); q = fmaf(acc.y, acc.y, q); q = fmaf(acc.z, acc.z, q); q = fmaf(acc.w, acc.w, q);
  }
  __shared__ float ss[8], sq[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  if ((tid & 31) == 0) { ss[tid >> 5] = s; sq[tid >> 5] = q; }
  __syncthreads();
  float ts = 0.f, tq = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { ts += ss[i]; tq += sq[i]; }
  const float inv_n = 1.0f / (float)n;
  const float mean = ts * inv_n;
  const float rstd = rsqrtf(fmaxf(tq * inv_n - mean * mean, 0.f) + a.eps);
  const float* temb = nullptr;
  if (a.temb_mode != TEMB_NONE) {
    int trow = 0;
    if (a.temb_mode == TEMB_PER_SAMPLE) trow = b;
    else if (a.temb_mode == TEMB_STEP) trow = *a.step_ptr + a.step_off;
    temb = a.temb + (size_t)trow * SPDM_TEMB_WIDTH + a.temb_off;
  }
  const float* film = a.film ? a.film + (size_t)b * SPDM_FILM_WIDTH + a.film_off : nullptr;
  bf16* __restrict__ out = reinterpret_cast<bf16*>(a.out);
#pragma unroll
  for (int k = 0; k < EPT; ++k) {
    const int idx = (k * 256 + tid) * 4;
    if (idx >= n) break;
    const int row = idx / a.C, c = idx - row * a.C;
    const float4 g = __ldg(reinterpret_cast<const float4*>(a.gamma + c));
    const float4 be = __ldg(reinterpret_cast<const float4*>(a.beta + c));
    float y[4] = {(x[k].x - mean) * rstd * g.x + be.x, (x[k].y - mean) * rstd * g.y + be.y,
                  (x[k].z - mean) * rstd * g.z + be.z, (x[k].w - mean) * rstd * g.w + be.w};
    if (a.act == ACT_GELU) {
#pragma unroll
      for (int i = 0; i < 4; ++i) y[i] = gelu_tanh_fast(y[i]);
    }
    if (temb) {
      const float4 t4 = __ldg(reinterpret_cast<const float4*>(temb + c));
      y[0] += t4.x; y[1] += t4.y; y[2] += t4.z; y[3] += t4.w;
    }
    if (film) {
      const float4 f0 = __ldg(reinterpret_cast<const float4*>(film + c));
      const float4 f1 = __ldg(reinterpret_cast<const float4*>(film + a.C + c));
      y[0] = fmaf(f0.x, y[0], f1.x); y[1] = fmaf(f0.y, y[1], f1.y); y[2] = fmaf(f0.z, y[2], f1.z); y[3] = fmaf(f0.w, y[3], f1.w);
    }
    uint2 o;
    __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
    ho[0] = __floats2bfloat162_rn(y[0], y[1]);
    ho[1] = __floats2bfloat162_rn(y[2], y[3]);
    *reinterpret_cast<uint2*>(out + ((size_t)b * a.HW + row) * a.ld_out + c) = o;
  }
}
}  // namespace
void launch_apply_partial(const ApplyArgs& a, const float* partial, int S, long long M, int B, cudaStream_t s) {
  const int n = a.HW * a.C;
  const int ept = (n + 1023) / 1024;
  if (ept <= 2) launch_pdl(apply_partial_kernel<2>, dim3(B), dim3(256), 0, s, a, partial, S, M);
  else if (ept <= 4) launch_pdl(apply_partial_kernel<4>, dim3(B), dim3(256), 0, s, a, partial, S, M);
  else if (ept <= 8) launch_pdl(apply_partial_kernel<8>, dim3(B), dim3(256), 0, s, a, partial, S, M);
  else launch_pdl(apply_partial_kernel<16>, dim3(B), dim3(256), 0, s, a, partial, S, M);
  COUNT_LAUNCH();
}
template <typename TI, typename TO> void launch_apply(const ApplyArgs& a, int B, cudaStream_t s) {
  const int nvec = a.HW * (a.C >> 3);
  // the per-thread prologue (statistics fold, 40 per-channel constants) is amortised over more vectors once the
  // launch is large enough to fill the machine anyway
  const bool big = (long long)B * nvec >= 148LL * 8 * APPLY_THREADS * 16;
  const int vpt = big ? 16 : 4;
  dim3 grid(cdiv(nvec, APPLY_THREADS * vpt), B);
  if (sizeof(TI) == 4) launch_pdl(apply_kernel<TI, TO, true, 4>, dim3(cdiv(nvec, APPLY_THREADS * 4), B), dim3(APPLY_THREADS), 0, s, a);
  else if (big) launch_pdl(apply_kernel<TI, TO, false, 16>, grid, dim3(APPLY_THREADS), 0, s, a);
  else launch_pdl(apply_kernel<TI, TO, false, 4>, grid, dim3(APPLY_THREADS), 0, s, a);
  COUNT_LAUNCH();
}
template void launch_apply<float, float>(const ApplyArgs&, int, cudaStream_t);
template void launch_apply<bf16, bf16>(const ApplyArgs&, int, cudaStream_t);

// =================================================================================================
// MaxPool2d(2) (models/Unet_FiLmLayer.py:132) and bilinear x2 upsample, align_corners=True (:191).
// =================================================================================================
namespace {
template <typename T>
__global__ void pool_kernel(const T* __restrict__ in, int ld_in, T* __restrict__ out, int ld_out, long long total, int Ho, int Wo, int C) {
  pdl_wait();
  pdl_trigger();
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= total) return;
  const int vec_per_row = C >> 3;
  const long long orow = v / vec_per_row;
  const int c8 = (int)(v - orow * vec_per_row) << 3;
  const int wo = (int)(orow % Wo);
  const long long t = orow / Wo;
  const int ho = (int)(t % Ho);
  const long long b = t / Ho;
  const int Wi = Wo * 2, Hi = Ho * 2;
  const long long r00 = (b * Hi + 2 * ho) * Wi + 2 * wo;
  float m[8], x[8];
  load8(in + r00 * ld_in + c8, m);
  load8(in + (r00 + 1) * ld_in + c8, x);
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], x[i]);
  load8(in + (r00 + Wi) * ld_in + c8, x);
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], x[i]);
  load8(in + (r00 + Wi + 1) * ld_in + c8, x);
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], x[i]);
  store8(out + orow * ld_out + c8, m);
}

template <typename T>
__global__ void upsample_kernel(const T* __restrict__ in, int ld_in, T* __restrict__ out, int ld_out, long long total, int Hi, int Wi, int C) {
  pdl_wait();
  pdl_trigger();
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= total) return;
  const int vec_per_row = C >> 3;
  const long long orow = v / vec_per_row;
  const int c8 = (int)(v - orow * vec_per_row) << 3;
  const int Ho = Hi * 2, Wo = Wi * 2;
  const int wo = (int)(orow % Wo);
  const long long t = orow / Wo;
  const int ho = (int)(t % Ho);
  const long long b = t / Ho;
  // align_corners=True: src = dst * (in-1)/(out-1)
  const float sh = (Ho > 1) ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f;
  const float sw = (Wo > 1) ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
  const float fh = sh * ho, fw = sw * wo;
  const int h0 = (int)fh, w0 = (int)fw;
  const int h1 = h0 + ((h0 < Hi - 1) ? 1 : 0), w1 = w0 + ((w0 < Wi - 1) ? 1 : 0);
  const float lh1 = fh - h0, lh0 = 1.f - lh1, lw1 = fw - w0, lw0 = 1.f - lw1;
  float a00[8], a01[8], a10[8], a11[8], y[8];
  const long long base = b * Hi;
  load8(in + ((base + h0) * Wi + w0) * ld_in + c8, a00);
  load8(in + ((base + h0) * Wi + w1) * ld_in + c8, a01);
  load8(in + ((base + h1) * Wi + w0) * ld_in + c8, a10);
  load8(in + ((base + h1) * Wi + w1) * ld_in + c8, a11);
#pragma unroll
  for (int i = 0; i < 8; ++i) y[i] = lh0 * (lw0 * a00[i] + lw1 * a01[i]) + lh1 * (lw0 * a10[i] + lw1 * a11[i]);
  store8(out + orow * ld_out + c8, y);
}
}  // namespace
template <typename T> void launch_pool(const T* in, int ld_in, T* out, int ld_out, int B, int Ho, int Wo, int C, cudaStream_t s) {
  const long long total = (long long)B * Ho * Wo * (C >> 3);
  launch_pdl(pool_kernel<T>, dim3(cdiv(total, 256)), dim3(256), 0, s, in, ld_in, out, ld_out, total, Ho, Wo, C);
  COUNT_LAUNCH();
}
template <typename T> void launch_upsample(const T* in, int ld_in, T* out, int ld_out, int B, int Hi, int Wi, int C, cudaStream_t s) {
  const long long total = (long long)B * Hi * 2 * Wi * 2 * (C >> 3);
  launch_pdl(upsample_kernel<T>, dim3(cdiv(total, 256)), dim3(256), 0, s, in, ld_in, out, ld_out, total, Hi, Wi, C);
  COUNT_LAUNCH();
}
template void launch_pool<float>(const float*, int, float*, int, int, int, int, int, cudaStream_t);
template void launch_pool<bf16>(const bf16*, int, bf16*, int, int, int, int, int, cudaStream_t);
template void launch_upsample<float>(const float*, int, float*, int, int, int, int, int, cudaStream_t);
template void launch_upsample<bf16>(const bf16*, int, bf16*, int, int, int, int, int, cudaStream_t);

// =================================================================================================
// LayerNorm over C (models/Unet_FiLmLayer.py:51,53): one warp per token row, C in {64,128,256}.
// =================================================================================================
namespace {
template <typename T, int LANES>  // LANES = C / 8 lanes per token row (8 channels = one 16-byte vector per lane)
__global__ void __launch_bounds__(256) layernorm_kernel(const T* __restrict__ in, int ld_in, T* __restrict__ out, int ld_out,
                                                        const float* __restrict__ g, const float* __restrict__ bta, long long M) {
  pdl_wait();
  pdl_trigger();
  constexpr int ROWS_PER_WARP = 32 / LANES;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LANES, l = lane % LANES;
  const long long row = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * ROWS_PER_WARP + sub;
  const bool ok = row < M;
  const int c0 = l * 8;
  constexpr float inv_c = 1.0f / (float)(LANES * 8);
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = 0.f;
  if (ok) load8(in + row * ld_in + c0, x);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * inv_c;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float d = x[i] - mean; q = fmaf(d, d, q); }
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q * inv_c + 1e-5f);
  if (!ok) return;
  float gg[8], bb[8];
  load8(g + c0, gg);
  load8(bta + c0, bb);
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = (x[i] - mean) * rstd * gg[i] + bb[i];
  store8(out + row * ld_out + c0, x);
}
// wide rows (C = 512): one warp per row, 16 channels per lane
template <typename T>
__global__ void __launch_bounds__(256) layernorm_wide_kernel(const T* __restrict__ in, int ld_in, T* __restrict__ out, int ld_out,
                                                             const float* __restrict__ g, const float* __restrict__ bta, long long M, int C) {
  pdl_wait();
  pdl_trigger();
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += to_f32<T>(in[row * ld_in + c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float q = 0.f;
  for (int c = lane; c < C; c += 32) { const float d = to_f32<T>(in[row * ld_in + c]) - mean; q = fmaf(d, d, q); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)C + 1e-5f);
  for (int c = lane; c < C; c += 32)
    out[row * ld_out + c] = from_f32<T>((to_f32<T>(in[row * ld_in + c]) - mean) * rstd * __ldg(g + c) + __ldg(bta + c));
}
}  // namespace
template <typename T> void launch_layernorm(const T* in, int ld_in, T* out, int ld_out, const float* g, const float* b, long long M, int C, cudaStream_t s) {
  if (C == 64) launch_pdl(layernorm_kernel<T, 8>, dim3(cdiv(M, 8 * 4)), dim3(256), 0, s, in, ld_in, out, ld_out, g, b, M);
  else if (C == 128) launch_pdl(layernorm_kernel<T, 16>, dim3(cdiv(M, 8 * 2)), dim3(256), 0, s, in, ld_in, out, ld_out, g, b, M);
  else if (C == 256) launch_pdl(layernorm_kernel<T, 32>, dim3(cdiv(M, 8)), dim3(256), 0, s, in, ld_in, out, ld_out, g, b, M);
  else launch_pdl(layernorm_wide_kernel<T>, dim3(cdiv(M, 8)), dim3(256), 0, s, in, ld_in, out, ld_out, g, b, M, C);
  COUNT_LAUNCH();
}
template void launch_layernorm<float>(const float*, int, float*, int, const float*, const float*, long long, int, cudaStream_t);
template void launch_layernorm<bf16>(const bf16*, int, bf16*, int, const float*, const float*, long long, int, cudaStream_t);

// =================================================================================================
// Multi-head attention core (nn.MultiheadAttention, 4 heads, models/Unet_FiLmLayer.py:50,76):
//   o[b, i, h*hd:(h+1)*hd] = softmax_j(q_i . k_j / sqrt(hd)) v_j,   qkv rows = [q | k | v] (3C wide).
// L = H*W tokens is tiny (4..1024), head_dim 16/32/64: one thread per query row, K/V of the
// (sample, head) group broadcast from shared memory, online softmax in fp32.
// =================================================================================================
namespace {
template <typename T, int HD>
__global__ void __launch_bounds__(128) sdpa_kernel(const T* __restrict__ qkv, T* __restrict__ out, int n_groups, int L, int C, int heads,
                                                   int groups_per_block, int threads_per_group) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm[];  // [groups_per_block][2][L][HD]
  const int gl = threadIdx.x / threads_per_group;
  const int tl = threadIdx.x - gl * threads_per_group;
  const int grp = blockIdx.x * groups_per_block + gl;
  const bool active = gl < groups_per_block && grp < n_groups;
  const int b = active ? grp / heads : 0, h = active ? grp - b * heads : 0;
  float* Ks = sm + (size_t)gl * 2 * L * HD;
  float* Vs = Ks + (size_t)L * HD;
  const int ld = 3 * C;
  if (active) {
    const int vec_per_row = HD >> 3;
    for (int v = tl; v < L * vec_per_row; v += threads_per_group) {
      const int j = v / vec_per_row, d8 = (v - j * vec_per_row) << 3;
      const T* row = qkv + ((size_t)b * L + j) * ld + h * HD + d8;
      float k8[8], v8[8];
      load8(row + C, k8);
      load8(row + 2 * C, v8);
#pragma unroll
      for (int i = 0; i < 8; ++i) { Ks[j * HD + d8 + i] = k8[i]; Vs[j * HD + d8 + i] = v8[i]; }
    }
  }
  __syncthreads();
  if (!active) return;
  const float scale = rsqrtf((float)HD);
  for (int i = tl; i < L; i += threads_per_group) {
    float q[HD], acc[HD];
    const T* qrow = qkv + ((size_t)b * L + i) * ld + h * HD;
#pragma unroll
    for (int d = 0; d < HD; d += 8) {
      float t[8];
      load8(qrow + d, t);
#pragma unroll
      for (int e = 0; e < 8; ++e) { q[d + e] = t[e] * scale; acc[d + e] = 0.f; }
    }
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < L; ++j) {
      const float* kj = Ks + j * HD;
      float sdot = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) sdot = fmaf(q[d], kj[d], sdot);
      const float* vj = Vs + j * HD;
      if (sdot > m) {
        const float corr = __expf(m - sdot);
        l = l * corr + 1.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) acc[d] = fmaf(acc[d], corr, vj[d]);
        m = sdot;
      } else {
        const float p = __expf(sdot - m);
        l += p;
#pragma unroll
        for (int d = 0; d < HD; ++d) acc[d] = fmaf(p, vj[d], acc[d]);
      }
    }
    const float inv = 1.f / l;
    T* orow = out + ((size_t)b * L + i) * C + h * HD;
#pragma unroll
    for (int d = 0; d < HD; d += 8) {
      float t[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) t[e] = acc[d + e] * inv;
      store8(orow + d, t);
    }
  }
}
}  // namespace
template <typename T> void launch_sdpa(const T* qkv, T* out, int B, int L, int C, int heads, cudaStream_t s) {
  const int hd = C / heads;
  const int n_groups = B * heads;
  const int tpg = L >= 128 ? 128 : (L < 4 ? 4 : L);  // L is a power of two times small factors; groups share a block
  const int gpb = 128 / tpg;
  const size_t smem = (size_t)gpb * 2 * L * hd * sizeof(float);
  const int grid = cdiv(n_groups, gpb);
#define SDPA_CASE(HD)                                                                                         \
  {                                                                                                           \
    static bool attr_set = false;                                                                             \
    if (!attr_set) {                                                                                          \
      cudaFuncSetAttribute(sdpa_kernel<T, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);      \
      attr_set = true;                                                                                        \
    }                                                                                                         \
    launch_pdl(sdpa_kernel<T, HD>, dim3(grid), dim3(128), smem, s, qkv, out, n_groups, L, C, heads, gpb, tpg);                   \
  }
  if (hd == 16) SDPA_CASE(16)
  else if (hd == 32) SDPA_CASE(32)
  else if (hd == 64) SDPA_CASE(64)
#undef SDPA_CASE
  COUNT_LAUNCH();
}
template void launch_sdpa<float>(const float*, float*, int, int, int, int, cudaStream_t);
template void launch_sdpa<bf16>(const bf16*, bf16*, int, int, int, int, cudaStream_t);

// =================================================================================================
// inc.first: Conv2d(1 -> 64, 3x3, pad 1) straight from the unpadded sample, pad_to folded in
// (models/Unet_FiLmLayer.py:15-34,286,101).  outc: Conv2d(64 -> 1, 1x1) + bias + unpad (:264,310).
// =================================================================================================
namespace {
// One block per sample: the unpadded sample and the 9x64 weights sit in shared memory, every thread keeps 8 channels
// of a pixel, and the block also produces the sample's GroupNorm (sum, sumsq) (P = 1 partial) -- no separate pass.
template <typename T>
__global__ void __launch_bounds__(256) conv_in_kernel(const float* __restrict__ x, const float* __restrict__ w /*[9][64]*/, T* __restrict__ out,
                                                      float* __restrict__ stats, int H, int W, int rows, int dim, int lh, int lw) {
  extern __shared__ float csm[];  // [9*64] weights, then [rows*dim] sample
  float* sw = csm;
  float* sx = csm + 9 * 64;
  const int tid = threadIdx.x, b = blockIdx.x;
  for (int i = tid; i < 9 * 64; i += 256) sw[i] = __ldg(w + i);
  pdl_wait();
  pdl_trigger();
  for (int i = tid; i < rows * dim; i += 256) sx[i] = x[(size_t)b * rows * dim + i];
  __syncthreads();
  const int c8 = (tid & 7) << 3;
  float s = 0.f, q = 0.f;
  for (int px = tid >> 3; px < H * W; px += 32) {
    const int hh = px / W, ww = px - hh * W;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int sh = hh + tap / 3 - 1 - lh, sw_ = ww + tap % 3 - 1 - lw;  // coordinates in the unpadded sample
      if (sh < 0 || sh >= rows || sw_ < 0 || sw_ >= dim) continue;
      const float xv = sx[sh * dim + sw_];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(xv, sw[tap * 64 + c8 + i], acc[i]);
    }
    store8(out + ((size_t)b * H * W + px) * 64 + c8, acc);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float r = to_f32<T>(from_f32<T>(acc[i]));  // statistics of the values the next kernel will read
      s += r;
      q = fmaf(r, r, q);
    }
  }
  __shared__ float ss[8], sq[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  if ((tid & 31) == 0) { ss[tid >> 5] = s; sq[tid >> 5] = q; }
  __syncthreads();
  if (tid == 0) {
    float ts = 0.f, tq = 0.f;
    for (int i = 0; i < 8; ++i) { ts += ss[i]; tq += sq[i]; }
    stats[2 * b] = ts;
    stats[2 * b + 1] = tq;
  }
}

// inc.first + GroupNorm(1, 64) + GELU of the default 32x8 map in ONE kernel (bf16 path): the block already owns the whole sample,
// so its 256 x 64 outputs stay in registers (8 pixels x 8 channels per thread) between the statistics and the apply --
// conv_in_kernel + apply_kernel without the raw round trip (at batch 4096: 104 + 140 us of the step).
__global__ void __launch_bounds__(256) conv_in_gn_kernel(const float* __restrict__ x, const float* __restrict__ w /*[9][64]*/,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         bf16* __restrict__ out, int H, int W, int rows, int dim, int lh, int lw, float eps) {
  extern __shared__ float csm[];  // [9*64] weights, then [rows*dim] sample
  float* sw = csm;
  float* sx = csm + 9 * 64;
  const int tid = threadIdx.x, b = blockIdx.x;
  for (int i = tid; i < 9 * 64; i += 256) sw[i] = __ldg(w + i);
  const int c8 = (tid & 7) << 3;
  float g[8], be[8];
  load8(gamma + c8, g);
  load8(beta + c8, be);
  pdl_wait();
  pdl_trigger();
  for (int i = tid; i < rows * dim; i += 256) sx[i] = x[(size_t)b * rows * dim + i];
  __syncthreads();
  float acc[8][8];
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {                 // H * W == 256 pixels: 8 per thread
    const int px = (tid >> 3) + 32 * j;
    const int hh = px / W, ww = px - hh * W;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int sh = hh + tap / 3 - 1 - lh, sw_ = ww + tap % 3 - 1 - lw;
      if (sh < 0 || sh >= rows || sw_ < 0 || sw_ >= dim) continue;
      const float xv = sx[sh * dim + sw_];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(xv, sw[tap * 64 + c8 + i], acc[j][i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += acc[j][i]; q = fmaf(acc[j][i], acc[j][i], q); }
  }
  __shared__ float ss[8], sq[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  if ((tid & 31) == 0) { ss[tid >> 5] = s; sq[tid >> 5] = q; }
  __syncthreads();
  float ts = 0.f, tq = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { ts += ss[i]; tq += sq[i]; }
  const float inv_n = 1.0f / (256.0f * 64.0f);
  const float mean = ts * inv_n;
  const float rstd = rsqrtf(fmaxf(tq * inv_n - mean * mean, 0.f) + eps);
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float ga = rstd * g[i]; be[i] = fmaf(-mean, ga, be[i]); g[i] = ga; }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int px = (tid >> 3) + 32 * j;
    float y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = gelu_tanh_fast(fmaf(acc[j][i], g[i], be[i]));
    store8(out + ((size_t)b * 256 + px) * 64 + c8, y);
  }
}

template <typename T>
__global__ void outc_kernel(const T* __restrict__ x, int ld, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ eps,
                            long long total, int H, int W, int C, int rows, int dim, int lh, int lw) {
  pdl_wait();
  pdl_trigger();
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= total) return;
  const int d = (int)(v % dim);
  const long long t = v / dim;
  const int rr = (int)(t % rows);
  const long long b = t / rows;
  const long long r = (b * H + rr + lh) * W + d + lw;
  float acc = 0.f;
  for (int c = 0; c < C; c += 8) {
    float t8[8];
    load8(x + r * ld + c, t8);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc = fmaf(t8[i], __ldg(w + c + i), acc);
  }
  eps[v] = acc + __ldg(bias);
}

template <typename T>
__global__ void to_nchw_kernel(const T* __restrict__ in, int ld, float* __restrict__ out, long long total, int HW, int C) {
  pdl_wait();
  pdl_trigger();
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= total) return;
  const int p = (int)(v % HW);
  const long long t = v / HW;
  const int c = (int)(t % C);
  const long long b = t / C;
  out[v] = to_f32<T>(in[(b * HW + p) * ld + c]);
}
}  // namespace
template <typename T> void launch_conv_in(const float* x, const float* w, T* out, float* stats, int B, int H, int W, int rows, int dim, int lh, int lw, cudaStream_t s) {
  const size_t smem = (size_t)(9 * 64 + rows * dim) * sizeof(float);
  launch_pdl(conv_in_kernel<T>, dim3(B), dim3(256), smem, s, x, w, out, stats, H, W, rows, dim, lh, lw);
  COUNT_LAUNCH();
}
void launch_conv_in_gn(const float* x, const float* w, const float* gamma, const float* beta, bf16* out, int B, int H, int W, int rows,
                       int dim, int lh, int lw, cudaStream_t s) {
  const size_t smem = (size_t)(9 * 64 + rows * dim) * sizeof(float);
  launch_pdl(conv_in_gn_kernel, dim3(B), dim3(256), smem, s, x, w, gamma, beta, out, H, W, rows, dim, lh, lw, 1e-5f);
  COUNT_LAUNCH();
}
template <typename T> void launch_outc(const T* x, int ld, const float* w, const float* bias, float* eps, int B, int H, int W, int C, int rows, int dim, int lh, int lw, cudaStream_t s) {
  const long long total = (long long)B * rows * dim;
  launch_pdl(outc_kernel<T>, dim3(cdiv(total, 128)), dim3(128), 0, s, x, ld, w, bias, eps, total, H, W, C, rows, dim, lh, lw);
  COUNT_LAUNCH();
}
template <typename T> void launch_to_nchw(const T* in, int ld, float* out, int B, int HW, int C, cudaStream_t s) {
  const long long total = (long long)B * HW * C;
  launch_pdl(to_nchw_kernel<T>, dim3(cdiv(total, 256)), dim3(256), 0, s, in, ld, out, total, HW, C);
}
template void launch_conv_in<float>(const float*, const float*, float*, float*, int, int, int, int, int, int, int, cudaStream_t);
template void launch_conv_in<bf16>(const float*, const float*, bf16*, float*, int, int, int, int, int, int, int, cudaStream_t);
template void launch_outc<float>(const float*, int, const float*, const float*, float*, int, int, int, int, int, int, int, int, cudaStream_t);
template void launch_outc<bf16>(const bf16*, int, const float*, const float*, float*, int, int, int, int, int, int, int, int, cudaStream_t);
template void launch_to_nchw<float>(const float*, int, float*, int, int, int, cudaStream_t);
template void launch_to_nchw<bf16>(const bf16*, int, float*, int, int, int, cudaStream_t);

// =================================================================================================
// Posterior update + add_constraints, one fused elementwise kernel per denoising step:
//   x0     = (x - c0*eps) / c1                       DDPMScheduler/DDIMScheduler.step (diffusers 0.17.1;
//   x_prev = k_x0*x0 + k_x*x + k_eps*eps + k_n*z      call sites models/diffusion_ddpm.py:211,274, ddim.py:61,72)
//   x_prev[:, :inpaint] = inpaint                     models/diffusion_ddpm.py:216-219
// coef row = {c0, c1, k_x0, k_x, k_eps, k_n, clip, 0} (clip > 0: x0 is clamped to [-clip, clip] first -- the schedulers'
// clip_sample=True / clip_sample_range, diffusers' default); the row index comes from a device counter so the
// launch is CUDA-graph replayable.  z is injected (parity) or Philox4x32-10 + Box-Muller (throughput).
// =================================================================================================
namespace {
__device__ __forceinline__ uint32_t mulhilo(uint32_t a, uint32_t b, uint32_t* hi) {
  const unsigned long long p = (unsigned long long)a * b;
  *hi = (uint32_t)(p >> 32);
  return (uint32_t)p;
}
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, hi1;
    const uint32_t lo0 = mulhilo(0xD2511F53u, c[0], &hi0);
    const uint32_t lo1 = mulhilo(0xCD9E8D57u, c[2], &hi1);
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
__device__ __forceinline__ float philox_normal(unsigned long long seed, unsigned long long idx, uint32_t step) {
  uint32_t c[4] = {(uint32_t)(idx >> 1), (uint32_t)(idx >> 33), step, 0x5bd1e995u};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const float u1 = ((float)(c[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = ((float)(c[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float rad = sqrtf(-2.0f * __logf(u1));
  float sn, cs;
  __sincosf(6.283185307179586f * u2, &sn, &cs);
  return (idx & 1) ? rad * sn : rad * cs;
}

__global__ void step_kernel(StepArgs a) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)a.B_total * a.n;  // elements of the whole batch (noise / history strides)
  const long long li = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= (long long)a.B * a.n) return;
  const long long i = (long long)a.b0 * a.n + li;      // element inside the whole batch
  int step = a.step_host, use_philox = 0;
  const float* noise = a.noise;
  const float* inpaint = a.inpaint;
  float* history = nullptr;
  unsigned long long seed = 0;
  if (a.dyn) {
    step = a.dyn->step + a.step_off;
    noise = a.dyn->noise ? a.dyn->noise + (size_t)step * total : nullptr;
    inpaint = a.dyn->inpaint;
    history = a.dyn->history;
    seed = a.dyn->seed;
    use_philox = a.dyn->use_philox;
  }
  const float* cf = a.coef + (size_t)step * 8;
  const float c0 = cf[0], c1 = cf[1], kx0 = cf[2], kx = cf[3], keps = cf[4], kn = cf[5], clip = cf[6];
  const long long b = i / a.n;
  const int e = (int)(i - b * a.n);
  float r;
  if (inpaint && e < a.inpaint_elems) {
    r = inpaint[b * a.inpaint_elems + e];
  } else {
    const float x = a.x[i], ep = a.eps[i];
    float x0 = (x - c0 * ep) / c1;
    if (clip > 0.f) x0 = fminf(fmaxf(x0, -clip), clip);  // clip_sample=True: x0.clamp(-range, range)
    r = kx0 * x0 + kx * x;
    if (keps != 0.f) r += keps * ep;
    if (kn != 0.f) {
      float z = 0.f;
      if (noise) z = noise[i];
      else if (use_philox) z = philox_normal(seed, (unsigned long long)i, (uint32_t)step);
      r += kn * z;
    }
  }
  a.x_out[i] = r;
  if (history) history[(size_t)(step + 1) * total + i] = r;
}
// outc (Conv2d(64 -> 1, 1x1) + bias + unpad, models/Unet_FiLmLayer.py:264,310) fused with the posterior update: the
// noise estimate of an element is consumed by that element's update only, so it never has to exist in memory.
template <typename T>
__global__ void outc_step_kernel(StepArgs a, const T* __restrict__ act, int ld, const float* __restrict__ w, const float* __restrict__ bias,
                                 int H, int W, int C, int rows, int dim, int lh, int lw) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)a.B_total * a.n;
  const long long li = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= (long long)a.B * a.n) return;
  const long long i = (long long)a.b0 * a.n + li;
  const int step = a.dyn->step + a.step_off;
  const float* noise = a.dyn->noise ? a.dyn->noise + (size_t)step * total : nullptr;
  const float* inpaint = a.dyn->inpaint;
  float* history = a.dyn->history;
  const float* cf = a.coef + (size_t)step * 8;
  const float c0 = cf[0], c1 = cf[1], kx0 = cf[2], kx = cf[3], keps = cf[4], kn = cf[5], clip = cf[6];
  const long long b = i / a.n;
  const int e = (int)(i - b * a.n);
  float r;
  if (inpaint && e < a.inpaint_elems) {
    r = inpaint[b * a.inpaint_elems + e];
  } else {
    const int d = e % dim, rr = e / dim;
    const long long lb = li / a.n;  // sample inside this lane's activation buffer
    const T* arow = act + ((lb * H + rr + lh) * W + d + lw) * ld;
    float ep = 0.f;
    for (int c = 0; c < C; c += 8) {
      float t8[8];
      load8(arow + c, t8);
#pragma unroll
      for (int k = 0; k < 8; ++k) ep = fmaf(t8[k], __ldg(w + c + k), ep);
    }
    ep += __ldg(bias);
    const float x = a.x[i];
    float x0 = (x - c0 * ep) / c1;
    if (clip > 0.f) x0 = fminf(fmaxf(x0, -clip), clip);  // clip_sample=True: x0.clamp(-range, range)
    r = kx0 * x0 + kx * x;
    if (keps != 0.f) r += keps * ep;
    if (kn != 0.f) {
      float z = 0.f;
      if (noise) z = noise[i];
      else if (a.dyn->use_philox) z = philox_normal(a.dyn->seed, (unsigned long long)i, (uint32_t)step);
      r += kn * z;
    }
  }
  a.x_out[i] = r;
  if (history) history[(size_t)(step + 1) * total + i] = r;
}
__global__ void advance_kernel(int* p, int d) {
  pdl_wait();
  pdl_trigger(); *p += d; }
__global__ void delay_kernel(long long cycles) {
  pdl_wait();
  pdl_trigger();
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {}
}

// DDPMScheduler.add_noise + add_constraints (models/diffusion_ddpm.py:167-168)
__global__ void add_noise_kernel(const float* __restrict__ x0, const float* __restrict__ noise, const long long* __restrict__ t,
                                 const float* __restrict__ sa, const float* __restrict__ sb, const float* __restrict__ inpaint,
                                 float* __restrict__ out, int n, int inpaint_elems, long long total) {
  pdl_wait();
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long b = i / n;
  const int e = (int)(i - b * n);
  if (inpaint && e < inpaint_elems) { out[i] = inpaint[b * inpaint_elems + e]; return; }
  const long long tb = t[b];
  out[i] = sa[tb] * x0[i] + sb[tb] * noise[i];
}
}  // namespace
void launch_step(const StepArgs& a, cudaStream_t s) {
  const long long total = (long long)a.B * a.n;
  launch_pdl(step_kernel, dim3(cdiv(total, 256)), dim3(256), 0, s, a);
  COUNT_LAUNCH();
}
template <typename T>
void launch_outc_step(const StepArgs& a, const T* act, int ld, const float* w, const float* bias, int H, int W, int C, int rows, int dim,
                      int lh, int lw, cudaStream_t s) {
  const long long total = (long long)a.B * a.n;
  launch_pdl(outc_step_kernel<T>, dim3(cdiv(total, 128)), dim3(128), 0, s, a, act, ld, w, bias, H, W, C, rows, dim, lh, lw);
  COUNT_LAUNCH();
}
template void launch_outc_step<float>(const StepArgs&, const float*, int, const float*, const float*, int, int, int, int, int, int, int, cudaStream_t);
template void launch_outc_step<bf16>(const StepArgs&, const bf16*, int, const float*, const float*, int, int, int, int, int, int, int, cudaStream_t);
void launch_delay(long long cycles, cudaStream_t s) { launch_pdl(delay_kernel, dim3(1), dim3(1), 0, s, cycles); }
void launch_advance(int* step_ptr, int delta, cudaStream_t s) { launch_pdl(advance_kernel, dim3(1), dim3(1), 0, s, step_ptr, delta); COUNT_LAUNCH(); }
void launch_add_noise(const float* x0, const float* noise, const long long* t, const float* sa, const float* sb, const float* inpaint,
                      float* out, int n, int inpaint_elems, int B, cudaStream_t s) {
  const long long total = (long long)B * n;
  launch_pdl(add_noise_kernel, dim3(cdiv(total, 256)), dim3(256), 0, s, x0, noise, t, sa, sb, inpaint, out, n, inpaint_elems, total);
  COUNT_LAUNCH();
}

// =================================================================================================
// Time embedding: pos_encoding (models/Unet_FiLmLayer.py:266-274) -> SiLU -> Linear(256 -> C) for
// all six stages at once (:136-142): out[row][896] = silu(posenc(t_row)) @ w_cat[256][896] + b_cat.
// One block per row.
// =================================================================================================
namespace {
__global__ void __launch_bounds__(256) temb_kernel(const long long* __restrict__ t_dev, const float* __restrict__ inv_freq,
                                                   const float* __restrict__ w_cat, const float* __restrict__ b_cat, float* __restrict__ out, int time_dim) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float pe[];  // [time_dim]
  const int row = blockIdx.x;
  const float t = (float)t_dev[row];
  const int half = time_dim >> 1;
  for (int i = threadIdx.x; i < time_dim; i += blockDim.x) {
    const float arg = t * inv_freq[i < half ? i : i - half];
    const float v = i < half ? sinf(arg) : cosf(arg);
    pe[i] = v / (1.f + expf(-v));  // SiLU
  }
  __syncthreads();
  for (int n = threadIdx.x; n < SPDM_TEMB_WIDTH; n += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < time_dim; ++k) acc = fmaf(pe[k], __ldg(w_cat + (size_t)k * SPDM_TEMB_WIDTH + n), acc);
    out[(size_t)row * SPDM_TEMB_WIDTH + n] = acc + __ldg(b_cat + n);
  }
}
// Mish (models/Unet_FiLmLayer.py:150): x * tanh(softplus(x)), softplus threshold 20 as torch
__global__ void mish_kernel(const float* __restrict__ in, float* __restrict__ out, long long n) {
  pdl_wait();
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = in[i];
  const float sp = x > 20.f ? x : log1pf(expf(x));
  out[i] = x * tanhf(sp);
}
}  // namespace
void launch_temb(const long long* t_dev, int n_t, const float* inv_freq, const float* w_cat, const float* b_cat, float* out, int time_dim, cudaStream_t s) {
  launch_pdl(temb_kernel, dim3(n_t), dim3(256), time_dim * sizeof(float), s, t_dev, inv_freq, w_cat, b_cat, out, time_dim);
  COUNT_LAUNCH();
}
void launch_mish(const float* in, float* out, long long n, cudaStream_t s) {
  launch_pdl(mish_kernel, dim3(cdiv(n, 256)), dim3(256), 0, s, in, out, n);
  COUNT_LAUNCH();
}

// =================================================================================================
// Observation encoder, conv part (models/encoder/autoencoder.py:11-17):
//   Conv(3->16,k2,s2,p1)+ReLU -> Conv(16->32,k2,s2)+ReLU -> Conv(32->64,k2,s2)+ReLU  on 96x96 frames.
// k == stride == 2, so the three layers are non-overlapping patch contractions: final pixel (i,j) of
// the 12x12 map depends only on an 8x8 input patch at rows 8i-1..8i+6 (row/col -1 = zero padding; the
// 49th conv1 row/col is never consumed).  One block per (frame, final row i): the 8x96x3 input strip
// is staged in shared memory once and all three layers run out of shared memory.
// Output feat[frame][(i*12+j)*64 + c]  (HWC order; the Linear weight is repacked to match).
// =================================================================================================
namespace {
// One block per frame; the three layers of one 12-pixel output row run out of shared memory, weights staged once per
// block in [k][channel] order so that a warp (32 consecutive channels of one position) reads them conflict-free while
// the activation is a broadcast.  w2t = [16*4][32], w3t = [32*4][64] (transposed at load time), w1 = (16,3,2,2).
template <typename TO>
__global__ void __launch_bounds__(256) enc_convs_kernel(const float* __restrict__ img, const float* __restrict__ w1, const float* __restrict__ b1,
                                                        const float* __restrict__ w2t, const float* __restrict__ b2, const float* __restrict__ w3t,
                                                        const float* __restrict__ b3, TO* __restrict__ feat) {
  extern __shared__ float es[];
  float* w3s = es;                 // 128*64
  float* w2s = w3s + 128 * 64;     // 64*32
  float* w1s = w2s + 64 * 32;      // 12*16
  float* s_in = w1s + 12 * 16;     // [3][8][97]
  float* s_c1 = s_in + 3 * 8 * 97; // [16][4][48]
  float* s_c2 = s_c1 + 16 * 4 * 48;// [32][2][24]
  const int tid = threadIdx.x;
  // weights are constants: stage them before waiting on the previous kernel
  for (int e = tid; e < 128 * 64; e += 256) w3s[e] = __ldg(w3t + e);
  for (int e = tid; e < 64 * 32; e += 256) w2s[e] = __ldg(w2t + e);
  for (int e = tid; e < 12 * 16; e += 256) w1s[e] = __ldg(w1 + (e % 16) * 12 + e / 16);
  pdl_wait();
  pdl_trigger();
  const int frame = blockIdx.x;
  const float* im = img + (size_t)frame * 3 * 96 * 96;
  const int ch1 = tid & 15, ch2 = tid & 31, ch3 = tid & 63;
  const float bias1 = __ldg(b1 + ch1), bias2 = __ldg(b2 + ch2), bias3 = __ldg(b3 + ch3);
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
    __syncthreads();
    // conv1: (ch1, pos) pos = tid/16 + 16k, k < 12  -> rr = pos/48, cc = pos%48
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
    // conv2: (ch2, pos) pos = tid/32 + 8k, k < 6 -> rr = pos/24, cc = pos%24
    {
      float acc[6];
      int off[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int pos = (tid >> 5) + 8 * k;
        const int rr = pos / 24, cc = pos - rr * 24;
        acc[k] = bias2;
        off[k] = (2 * rr) * 48 + 2 * cc;
      }
      for (int c = 0; c < 16; ++c) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float w = w2s[(c * 4 + kk) * 32 + ch2];
          const int o = c * 4 * 48 + (kk >> 1) * 48 + (kk & 1);
#pragma unroll
          for (int k = 0; k < 6; ++k) acc[k] = fmaf(s_c1[o + off[k]], w, acc[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int pos = (tid >> 5) + 8 * k;
        const int rr = pos / 24, cc = pos - rr * 24;
        s_c2[(ch2 * 2 + rr) * 24 + cc] = fmaxf(acc[k], 0.f);
      }
    }
    __syncthreads();
    // conv3: (ch3, j) j = tid/64 + 4k, k < 3
    {
      float acc[3] = {bias3, bias3, bias3};
      for (int c = 0; c < 32; ++c) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float w = w3s[(c * 4 + kk) * 64 + ch3];
          const int o = (c * 2 + (kk >> 1)) * 24 + (kk & 1);
#pragma unroll
          for (int k = 0; k < 3; ++k) acc[k] = fmaf(s_c2[o + 2 * ((tid >> 6) + 4 * k)], w, acc[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int j = (tid >> 6) + 4 * k;
        feat[(size_t)frame * 9216 + (i * 12 + j) * 64 + ch3] = from_f32<TO>(fmaxf(acc[k], 0.f));
      }
    }
  }
}

__global__ void cast_f32_kernel(const bf16* __restrict__ in, float* __restrict__ out, long long n) {
  pdl_wait();
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __bfloat162float(in[i]);
}

// prepare_obs_cond_vectors (models/diffusion_ddpm.py:317-330): cat[pos(2), act(3), vel(2), img_feat(128)]
__global__ void build_cond_kernel(const float* __restrict__ pos, const float* __restrict__ act, const float* __restrict__ vel,
                                  const float* __restrict__ feat, float* __restrict__ cond, long long total, int cond_dim) {
  pdl_wait();
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int d = (int)(i % cond_dim);
  const long long bt = i / cond_dim;
  float v;
  if (d < 2) v = pos[bt * 2 + d];
  else if (d < 5) v = act[bt * 3 + d - 2];
  else if (d < 7) v = vel[bt * 2 + d - 5];
  else v = feat[bt * (cond_dim - 7) + d - 7];
  cond[i] = v;
}
}  // namespace
template <typename TO>
void launch_enc_convs(const float* img, const float* w1, const float* b1, const float* w2t, const float* b2, const float* w3t,
                      const float* b3, TO* feat, int n, cudaStream_t s) {
  constexpr size_t smem = (128 * 64 + 64 * 32 + 12 * 16 + 3 * 8 * 97 + 16 * 4 * 48 + 32 * 2 * 24) * sizeof(float);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(enc_convs_kernel<TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
  launch_pdl(enc_convs_kernel<TO>, dim3(n), dim3(256), smem, s, img, w1, b1, w2t, b2, w3t, b3, feat);
  COUNT_LAUNCH();
}
template void launch_enc_convs<float>(const float*, const float*, const float*, const float*, const float*, const float*, const float*, float*, int, cudaStream_t);
template void launch_enc_convs<bf16>(const float*, const float*, const float*, const float*, const float*, const float*, const float*, bf16*, int, cudaStream_t);
void launch_cast_f32(const bf16* in, float* out, long long n, cudaStream_t s) {
  launch_pdl(cast_f32_kernel, dim3(cdiv(n, 256)), dim3(256), 0, s, in, out, n);
  COUNT_LAUNCH();
}
void launch_build_cond(const float* pos, const float* act, const float* vel, const float* feat, float* cond, int B, int T, int cond_dim, cudaStream_t s) {
  const long long total = (long long)B * T * cond_dim;
  launch_pdl(build_cond_kernel, dim3(cdiv(total, 256)), dim3(256), 0, s, pos, act, vel, feat, cond, total, cond_dim);
  COUNT_LAUNCH();
}

// =================================================================================================
// Weight repacking (load_state_dict time; PyTorch layouts -> kernel layouts)
// =================================================================================================
namespace {
__global__ void pack_conv_f32_kernel(const float* __restrict__ oihw, float* __restrict__ out, int Cout, int Cin, int kk) {
  pdl_wait();
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)Cout * Cin * kk;
  if (i >= total) return;
  const int o = (int)(i % Cout);
  const long long t = i / Cout;
  const int c = (int)(t % Cin);
  const int tap = (int)(t / Cin);
  out[i] = oihw[((size_t)o * Cin + c) * kk + tap];
}
__global__ void pack_conv_bf16_kernel(const float* __restrict__ oihw, bf16* __restrict__ out, int Cout, int Cin, int kk) {
  pdl_wait();
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)Cout * Cin * kk;
  if (i >= total) return;
  const int c = (int)(i % Cin);
  const long long t = i / Cin;
  const int tap = (int)(t % kk);
  const int o = (int)(t / kk);
  out[i] = __float2bfloat16_rn(oihw[((size_t)o * Cin + c) * kk + tap]);
}
// 3x3 weights of a conv on a W = 2 map, folded so that the two pixels of an image row become extra input / output channels:
//   out[(wo*Cout + co)][tap (dy, dx = 0)][wi*Cin + ci] = w[co][ci][dy][(wi - wo) + 1]
// in the ordinary [Cout'][9][Cin'] layout with Cout' = 2 Cout, Cin' = 2 Cin (the dx != 0 taps of the folded conv stay zero and
// are never read: a W = 1 geometry skips them).  The folded conv is dense -- no zero-padding taps along W.
__global__ void pack_conv_fold2_bf16_kernel(const float* __restrict__ oihw, bf16* __restrict__ out, int Cout, int Cin) {
  pdl_wait();
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = 2LL * Cout * 3 * 2 * Cin;
  if (i >= total) return;
  const int cin2 = (int)(i % (2 * Cin));
  const long long t = i / (2 * Cin);
  const int ty = (int)(t % 3);
  const int co2 = (int)(t / 3);
  const int wi = cin2 / Cin, ci = cin2 - wi * Cin, wo = co2 / Cout, co = co2 - wo * Cout;
  const int tx = wi - wo + 1;
  out[((size_t)co2 * 9 + ty * 3 + 1) * (2 * Cin) + cin2] = __float2bfloat16_rn(oihw[((size_t)co * Cin + ci) * 9 + ty * 3 + tx]);
}
// Pair fold of a 64-output-channel 3x3 conv (conv_tc_swap_kernel, TcParams::fold): two horizontally adjacent pixels (a "pair",
// wp = w / 2) become one GEMM row, so that the swapped-operand MMA gets M = 128 real rows (wo, co) instead of 64 -- a 64-row
// tcgen05.mma takes as long as a 128-row one.  Per dy the K dimension is four Cin-wide blocks:
//   blk 0: pair wp - 1, its pixel wi = 1 (w = 2 wp - 1)     blk 1: pair wp, wi = 0 (w = 2 wp)
//   blk 2: pair wp, wi = 1 (w = 2 wp + 1)                   blk 3: pair wp + 1, its pixel wi = 0 (w = 2 wp + 2)
//   out[(wo*64 + co)][(dy*4 + blk)*Cin + ci] = w[co][ci][dy][dx + 1] with dx = w_in - (2 wp + wo), zero where |dx| > 1
// i.e. 8 of the 16 (wo, blk) sub-blocks per dy are dense, 6 of them... the two corner ones (blk 0 / wo 1, blk 3 / wo 0) are zero.
__global__ void pack_conv_pfold_bf16_kernel(const float* __restrict__ oihw, bf16* __restrict__ out, int Cin) {
  pdl_wait();
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = 128LL * 12 * Cin;
  if (i >= total) return;
  const int r = (int)(i % (12 * Cin));
  const int m = (int)(i / (12 * Cin));
  const int ci = r % Cin, blk = (r / Cin) & 3, dyi = r / (4 * Cin);
  const int wo = m >> 6, co = m & 63;
  const int w_in = blk == 0 ? -1 : (blk == 1 ? 0 : (blk == 2 ? 1 : 2));
  const int dx = w_in - wo;
  float v = 0.f;
  if (dx >= -1 && dx <= 1) v = oihw[((size_t)co * Cin + ci) * 9 + dyi * 3 + dx + 1];
  out[i] = __float2bfloat16_rn(v);
}
__global__ void pack_linear_f32_kernel(const float* __restrict__ nk, float* __restrict__ out, int N, int K, int ld_out, int col_off) {
  pdl_wait();
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * K) return;
  const int n = (int)(i % N);
  const int k = (int)(i / N);
  out[(size_t)k * ld_out + col_off + n] = nk[(size_t)n * K + k];
}
__global__ void cast_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, long long n) {
  pdl_wait();
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16_rn(in[i]);
}
__global__ void pack_enc_linear_kernel(const float* __restrict__ w, float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  // w (128, 9216) with k = c*144 + p  ->  out[(p*64 + c)][128]
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 128LL * 9216) return;
  const int n = (int)(i % 128);
  const int k2 = (int)(i / 128);
  const int c = k2 % 64, p = k2 / 64;
  out[i] = w[(size_t)n * 9216 + c * 144 + p];
}
}  // namespace
void launch_pack_conv_f32(const float* oihw, float* out, int Cout, int Cin, int k, cudaStream_t s) {
  const long long total = (long long)Cout * Cin * k * k;
  launch_pdl(pack_conv_f32_kernel, dim3(cdiv(total, 256)), dim3(256), 0, s, oihw, out, Cout, Cin, k * k);
}
namespace {
// 3x3 weights, tiled through shared memory (Cout, Cin multiples of 32): a block reads 32 output channels x 32 input channels x
// 9 taps as 1152-byte runs of the OIHW tensor and writes 64-byte runs of the forward operand [Cout][tap][Cin] and (training) of
// the dgrad operand [Cin][tap'][Cout] (tap' = 8 - tap) -- one pass over the fp32 weights for both.  The elementwise kernels
// above read with a 36-byte stride; every index split below is by a compile-time constant.
constexpr int PK_T = 32;
constexpr int PK_ROW = PK_T * 9;          // 288 floats of one output channel's 32 input channels
__global__ void __launch_bounds__(256) pack_conv3_tiled_kernel(const float* __restrict__ oihw, bf16* __restrict__ out_f, bf16* __restrict__ out_d,
                                                               int Cout, int Cin) {
  __shared__ float tile[PK_T][PK_ROW + 1];
  pdl_wait();
  pdl_trigger();
  const int ci0 = blockIdx.x * PK_T, co0 = blockIdx.y * PK_T;
  const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;   // 8 warps
#pragma unroll
  for (int rr = 0; rr < PK_T / 8; ++rr) {
    const int r = q + 8 * rr;
    const float* src = oihw + ((size_t)(co0 + r) * Cin + ci0) * 9;
#pragma unroll
    for (int k = lane; k < PK_ROW; k += 32) tile[r][k] = src[k];
  }
  __syncthreads();
  if (out_f) {   // lane = input channel: 32 x 2 B contiguous
#pragma unroll 4
    for (int idx = q; idx < PK_T * 9; idx += 8) {
      const int r = idx / 9, tap = idx - r * 9;
      out_f[((size_t)(co0 + r) * 9 + tap) * Cin + ci0 + lane] = __float2bfloat16_rn(tile[r][lane * 9 + tap]);
    }
  }
  if (out_d) {   // lane = output channel
#pragma unroll 4
    for (int idx = q; idx < PK_T * 9; idx += 8) {
      const int c = idx / 9, tp = idx - c * 9;
      out_d[((size_t)(ci0 + c) * 9 + tp) * Cout + co0 + lane] = __float2bfloat16_rn(tile[lane][c * 9 + 8 - tp]);
    }
  }
}
__global__ void __launch_bounds__(256) unpack_conv3_tiled_kernel(const float* __restrict__ packed, float* __restrict__ dst, int Cout, int Cin) {
  __shared__ float tile[PK_T][PK_ROW + 1];
  const int ci0 = blockIdx.x * PK_T, co0 = blockIdx.y * PK_T;
  const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
#pragma unroll
  for (int idx = q; idx < PK_T * 9; idx += 8) {   // packed [tap][Cout][Cin]: lane = input channel, 128 B contiguous (36 loads in flight)
    const int tap = idx >> 5, r = idx & 31;
    tile[r][lane * 9 + tap] = packed[((size_t)tap * Cout + co0 + r) * Cin + ci0 + lane];
  }
  __syncthreads();
#pragma unroll
  for (int rr = 0; rr < PK_T / 8; ++rr) {
    const int r = q + 8 * rr;
    float* d = dst + ((size_t)(co0 + r) * Cin + ci0) * 9;
#pragma unroll
    for (int k = lane; k < PK_ROW; k += 32) d[k] = tile[r][k];
  }
}
}  // namespace
void launch_pack_conv_bf16(const float* oihw, bf16* out, int Cout, int Cin, int k, cudaStream_t s) {
  if (k == 3 && Cout % PK_T == 0 && Cin % PK_T == 0) { launch_pack_conv3_bf16(oihw, out, nullptr, Cout, Cin, s); return; }
  const long long total = (long long)Cout * Cin * k * k;
  launch_pdl(pack_conv_bf16_kernel, dim3(cdiv(total, 256)), dim3(256), 0, s, oihw, out, Cout, Cin, k * k);
}
bool launch_pack_conv3_bf16(const float* oihw, bf16* out_fwd, bf16* out_dgrad, int Cout, int Cin, cudaStream_t s) {
  if (Cout % PK_T || Cin % PK_T) return false;
  launch_pdl(pack_conv3_tiled_kernel, dim3(Cin / PK_T, Cout / PK_T), dim3(256), 0, s, oihw, out_fwd, out_dgrad, Cout, Cin);
  return true;
}
bool launch_unpack_conv3_grad(const float* packed, float* dst, int Cout, int Cin, cudaStream_t s) {
  if (Cout % PK_T || Cin % PK_T) return false;
  unpack_conv3_tiled_kernel<<<dim3(Cin / PK_T, Cout / PK_T), 256, 0, s>>>(packed, dst, Cout, Cin);
  COUNT_LAUNCH();
  return true;
}
void launch_pack_conv_fold2_bf16(const float* oihw, bf16* out, int Cout, int Cin, cudaStream_t s) {
  const long long total = 2LL * Cout * 3 * 2 * Cin;
  launch_pdl(pack_conv_fold2_bf16_kernel, dim3(cdiv(total, 256)), dim3(256), 0, s, oihw, out, Cout, Cin);
}
namespace {
__global__ void pack_conv_tf32_kernel(const float* __restrict__ oihw, float* __restrict__ out, int Cout, int Cin) {
  pdl_wait();
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Cout * Cin * 9) return;
  const int ci = (int)(i % Cin);
  const long long t = i / Cin;
  const int tap = (int)(t % 9), co = (int)(t / 9);
  out[i] = oihw[((size_t)co * Cin + ci) * 9 + tap];
}
}  // namespace
void launch_pack_conv_tf32(const float* oihw, float* out, int Cout, int Cin, cudaStream_t s) {
  launch_pdl(pack_conv_tf32_kernel, dim3(cdiv((long long)Cout * Cin * 9, 256)), dim3(256), 0, s, oihw, out, Cout, Cin);
}
void launch_pack_conv_pfold_bf16(const float* oihw, bf16* out, int Cin, cudaStream_t s) {
  const long long total = 128LL * 12 * Cin;
  launch_pdl(pack_conv_pfold_bf16_kernel, dim3(cdiv(total, 256)), dim3(256), 0, s, oihw, out, Cin);
}
void launch_pack_linear_f32(const float* nk, float* out, int N, int K, int ld_out, int col_off, cudaStream_t s) {
  launch_pdl(pack_linear_f32_kernel, dim3(cdiv((long long)N * K, 256)), dim3(256), 0, s, nk, out, N, K, ld_out, col_off);
}
void launch_cast_bf16(const float* in, bf16* out, long long n, cudaStream_t s) {
  launch_pdl(cast_bf16_kernel, dim3(cdiv(n, 256)), dim3(256), 0, s, in, out, n);
}
namespace {
__global__ void pack_enc_linear_bf16_kernel(const float* __restrict__ w, bf16* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 128LL * 9216) return;
  const int k2 = (int)(i % 9216);
  const int n = (int)(i / 9216);
  const int c = k2 % 64, p = k2 / 64;
  out[i] = __float2bfloat16_rn(w[(size_t)n * 9216 + c * 144 + p]);
}
}  // namespace
void launch_pack_enc_linear_bf16(const float* w, bf16* out, cudaStream_t s) {
  launch_pdl(pack_enc_linear_bf16_kernel, dim3(cdiv(128LL * 9216, 256)), dim3(256), 0, s, w, out);
}
void launch_pack_enc_linear(const float* w, float* out, cudaStream_t s) {
  launch_pdl(pack_enc_linear_kernel, dim3(cdiv(128LL * 9216, 256)), dim3(256), 0, s, w, out);
}
