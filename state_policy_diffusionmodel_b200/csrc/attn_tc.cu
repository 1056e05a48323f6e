// attn_tc.cu — the tail of a SelfAttention block as ONE tcgen05 kernel (sm_100a).
//
//   a   = out_proj(att) + x                         reference models/Unet_FiLmLayer.py:76-77 (MHA out projection + residual)
//   out = Linear2(GELU(Linear1(LN(a)))) + a         :54-58,78 (ff_self + residual)
//
// Everything is local to a token row, so one CTA owns a 128-token tile and runs the three C x C GEMMs back to back:
// the A operand of each GEMM is produced in shared memory by the epilogue of the previous one (bf16, K-major,
// 128B swizzle), the weights stream in by TMA while that epilogue runs, the accumulator and the fp32 residual `a`
// live in tensor memory.  Replaces four launches (GEMM, LayerNorm, GEMM, GEMM) and three HBM round trips per block.
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

struct TailParams {
  const bf16* x;       // residual input [M][ld_x]
  int ld_x;
  bf16* out;           // [M][ld_out]
  int ld_out;
  const float *bo, *b1, *b2, *ln_g, *ln_b;
};

__device__ __forceinline__ float erf_fast_(float x) {
  const float ax = fabsf(x);
  const float t = __frcp_rn(fmaf(0.3275911f, ax, 1.0f));
  const float poly = t * (0.254829592f + t * (-0.284496736f + t * (1.421413741f + t * (-1.453152027f + t * 1.061405429f))));
  return copysignf(1.0f - poly * __expf(-ax * ax), x);
}

// GELU through the hardware tanh (see kernels.cu::gelu_tanh_fast): the epilogue runs on 4 warps and is issue-bound
__device__ __forceinline__ float gelu_tanh_fast_(float y) {
  const float u = 0.7978845608028654f * fmaf(0.044715f * y * y, y, y);
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(u));
  return fmaf(0.5f * y, th, 0.5f * y);
}

// 32 fp32 values of row r, columns [c, c+32) -> bf16 -> K-major 128B-swizzled A tile (k-block c/64)
__device__ __forceinline__ void store_a_chunk(uint8_t* sA, int r, int c, const float f[32]) {
  uint8_t* rowp = sA + (c >> 6) * 16384 + r * 128;
  const int j0 = (c & 63) >> 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 val;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&val);
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(f[8 * j + 2 * e], f[8 * j + 2 * e + 1]);
    *reinterpret_cast<uint4*>(rowp + (((j0 + j) ^ (r & 7)) << 4)) = val;
  }
}

// 256 threads: two threads per token row, each owns half of the row's C columns in every epilogue (the epilogues -- 128 threads x C
// columns each -- were the longest part of the C = 256 instances: ~20 us per launch on 8..32 CTAs); the LayerNorm statistics of a
// row are combined through shared memory and a 64-thread named barrier.
template <int C>
__global__ void __launch_bounds__(256)
attn_tail_kernel(const __grid_constant__ CUtensorMap map_att, const __grid_constant__ CUtensorMap map_wo,
                 const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w2, const TailParams p) {
  constexpr int KB = C / 64;
  constexpr int SA = KB * 16384;        // 128 rows x C
  constexpr int SW = C * C * 2;         // KB blocks of [C rows][64]
  constexpr int WBLK = C * 128;
  constexpr int TMEM_COLS = 2 * C < 32 ? 32 : 2 * C;  // D: [0,C)  R: [C,2C)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sW = smem + SA;
  __shared__ __align__(8) uint64_t bar_load;
  __shared__ __align__(8) uint64_t bar_mma;
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_bo[C], s_b1[C], s_b2[C], s_g[C], s_b[C];  // per-column constants (weights: no dependency)
  __shared__ float s_ls[2][128], s_lq[2][128];                                // LayerNorm partial sums of the two column halves

  const int tid = threadIdx.x, warp = tid >> 5;
  const int row0 = blockIdx.x * 128;
  for (int i = tid; i < C; i += 256) {
    s_bo[i] = __ldg(p.bo + i); s_b1[i] = __ldg(p.b1 + i); s_b2[i] = __ldg(p.b2 + i);
    s_g[i] = __ldg(p.ln_g + i); s_b[i] = __ldg(p.ln_b + i);
  }
  if (tid == 0) {
    tma_prefetch_desc(&map_att); tma_prefetch_desc(&map_wo); tma_prefetch_desc(&map_w1); tma_prefetch_desc(&map_w2);
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<TMEM_COLS>(&tmem_base_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_smem;
  auto load_w = [&](const CUtensorMap* m) {
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) tma_load_2d(sW + kb * WBLK, m, &bar_load, kb * 64, 0);
  };
  // the weights do not depend on the previous kernel: Wo is put in flight before the programmatic-dependency wait
  if (tid == 0) {
    mbar_expect_tx(&bar_load, SA + SW);
    load_w(&map_wo);
  }
  pdl_wait();
  pdl_trigger();

  auto issue_gemm = [&]() {  // D[128][C] = A[128][C] * W[C][C]^T
    const uint32_t idesc = make_idesc(C);
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
      const uint64_t da = make_smem_desc(smem_u32(sA + kb * 16384));
      const uint64_t dw = make_smem_desc(smem_u32(sW + kb * WBLK));
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma_bf16(tmem, da + (uint64_t)(2 * kk), dw + (uint64_t)(2 * kk), idesc, (kb > 0 || kk > 0) ? 1u : 0u);
    }
    umma_commit(&bar_mma);
  };

  uint32_t load_phase = 0, mma_phase = 0;
  const int half = warp >> 2;
  const int r = (warp & 3) * 32 + (tid & 31);
  const int c_lo = half * (C / 2), c_hi = c_lo + C / 2;   // this thread's columns
  const uint32_t pair_bar = 1 + (warp & 3);
  const long long row = (long long)row0 + r;
  const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);

  // ---------------- GEMM 1: att @ Wo^T ----------------
  if (tid == 0) {
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) tma_load_2d(sA + kb * 16384, &map_att, &bar_load, kb * 64, row0);
    mbar_wait(&bar_load, load_phase);
    tc_fence_after();
    issue_gemm();
  }
  load_phase ^= 1u;
  const bf16* xrow = p.x + row * p.ld_x;
  uint4 xn[4];   // residual row, prefetched one 32-column chunk ahead of the accumulator reads
#pragma unroll
  for (int j = 0; j < 4; ++j) xn[j] = *reinterpret_cast<const uint4*>(xrow + c_lo + 8 * j);
  mbar_wait(&bar_mma, mma_phase);
  mma_phase ^= 1u;
  tc_fence_after();
  if (tid == 0) {  // W1 streams in while the epilogue below runs (the MMAs that read Wo have retired)
    mbar_expect_tx(&bar_load, SW);
    load_w(&map_w1);
  }
  // epilogue 1: a = acc + bo + x  (kept fp32 in TMEM region R), LayerNorm statistics
  float s = 0.f, q = 0.f;
#pragma unroll 1
  for (int c = c_lo; c < c_hi; c += 32) {
    uint4 xc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) xc[j] = xn[j];
    if (c + 32 < c_hi) {
#pragma unroll
      for (int j = 0; j < 4; ++j) xn[j] = *reinterpret_cast<const uint4*>(xrow + c + 32 + 8 * j);
    }
    uint32_t v[32];
    tmem_ld_32x32(t_lane + (uint32_t)c, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&xc[j]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = 8 * j + 2 * e;
        const float a0 = __uint_as_float(v[i]) + s_bo[c + i] + __low2float(h2[e]);
        const float a1 = __uint_as_float(v[i + 1]) + s_bo[c + i + 1] + __high2float(h2[e]);
        s += a0 + a1;
        q = fmaf(a0, a0, q);
        q = fmaf(a1, a1, q);
        v[i] = __float_as_uint(a0);
        v[i + 1] = __float_as_uint(a1);
      }
    }
    tmem_st_32x32(t_lane + (uint32_t)(C + c), v);
  }
  tmem_st_wait();
  s_ls[half][r] = s;
  s_lq[half][r] = q;
  asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
  s = s_ls[0][r] + s_ls[1][r];
  q = s_lq[0][r] + s_lq[1][r];
  const float mean = s * (1.0f / C);
  const float rstd = rsqrtf(fmaxf(q * (1.0f / C) - mean * mean, 0.f) + 1e-5f);
#pragma unroll 1
  for (int c = c_lo; c < c_hi; c += 32) {
    uint32_t v[32];
    tmem_ld_32x32(t_lane + (uint32_t)(C + c), v);
    tmem_ld_wait();
    float f[32];
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 g4 = *reinterpret_cast<const float4*>(s_g + c + i);
      const float4 b4 = *reinterpret_cast<const float4*>(s_b + c + i);
      f[i] = (__uint_as_float(v[i]) - mean) * rstd * g4.x + b4.x;
      f[i + 1] = (__uint_as_float(v[i + 1]) - mean) * rstd * g4.y + b4.y;
      f[i + 2] = (__uint_as_float(v[i + 2]) - mean) * rstd * g4.z + b4.z;
      f[i + 3] = (__uint_as_float(v[i + 3]) - mean) * rstd * g4.w + b4.w;
    }
    store_a_chunk(sA, r, c, f);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();

  // ---------------- GEMM 2: LN(a) @ W1^T ----------------
  if (tid == 0) {
    mbar_wait(&bar_load, load_phase);
    tc_fence_after();
    issue_gemm();
  }
  load_phase ^= 1u;
  mbar_wait(&bar_mma, mma_phase);
  mma_phase ^= 1u;
  tc_fence_after();
  if (tid == 0) {
    mbar_expect_tx(&bar_load, SW);
    load_w(&map_w2);
  }
  // epilogue 2: GELU(acc + b1) -> A operand of GEMM 3
#pragma unroll 1
  for (int c = c_lo; c < c_hi; c += 32) {
    uint32_t v[32];
    tmem_ld_32x32(t_lane + (uint32_t)c, v);
    tmem_ld_wait();
    float f[32];
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(s_b1 + c + i);
      const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float y = __uint_as_float(v[i + e]) + bb[e];
        f[i + e] = gelu_tanh_fast_(y);
      }
    }
    store_a_chunk(sA, r, c, f);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();

  // ---------------- GEMM 3: h @ W2^T ----------------
  if (tid == 0) {
    mbar_wait(&bar_load, load_phase);
    tc_fence_after();
    issue_gemm();
  }
  load_phase ^= 1u;
  mbar_wait(&bar_mma, mma_phase);
  mma_phase ^= 1u;
  tc_fence_after();
  // epilogue 3: out = acc + b2 + a
  bf16* orow = p.out + row * p.ld_out;
#pragma unroll 1
  for (int c = c_lo; c < c_hi; c += 32) {
    uint32_t v[32], a[32];
    tmem_ld_32x32(t_lane + (uint32_t)c, v);
    tmem_ld_32x32(t_lane + (uint32_t)(C + c), a);
    tmem_ld_wait();
    float f[32];
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(s_b2 + c + i);
      f[i] = __uint_as_float(v[i]) + b4.x + __uint_as_float(a[i]);
      f[i + 1] = __uint_as_float(v[i + 1]) + b4.y + __uint_as_float(a[i + 1]);
      f[i + 2] = __uint_as_float(v[i + 2]) + b4.z + __uint_as_float(a[i + 2]);
      f[i + 3] = __uint_as_float(v[i + 3]) + b4.w + __uint_as_float(a[i + 3]);
    }
#pragma unroll
    for (int i = 0; i < 32; i += 8) store8(orow + c + i, f + i);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) return nullptr;
    fn = (EncodeTiledFn)p;
  }
  return fn;
}
bool encode_2d(EncodeTiledFn enc, CUtensorMap* m, const void* ptr, uint64_t cols, uint64_t rows, uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int C> void launch_tail(const struct AttnTail* g, const TailParams& p, long long M, cudaStream_t s);

}  // namespace

struct AttnTail {
  CUtensorMap map_att, map_wo, map_w1, map_w2;
  TailParams p;
  int C;
};

namespace {
template <int C> void launch_tail(const AttnTail* g, const TailParams& p, long long M, cudaStream_t s) {
  constexpr int smem = (C / 64) * 16384 + C * C * 2 + 1024;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(attn_tail_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
  launch_pdl(attn_tail_kernel<C>, dim3((unsigned)(M / 128)), dim3(256), smem, s, g->map_att, g->map_wo, g->map_w1, g->map_w2, p);
}
}  // namespace

bool attn_tail_supported(int C) { return C == 64 || C == 128 || C == 256; }

AttnTail* attn_tail_create(const bf16* att, long long Mcap, int C, const bf16* wo, const bf16* w1, const bf16* w2, const float* bo,
                           const float* b1, const float* b2, const float* ln_g, const float* ln_b) {
  EncodeTiledFn enc = get_encode();
  if (!enc || !attn_tail_supported(C)) return nullptr;
  AttnTail* g = new AttnTail();
  memset(g, 0, sizeof(*g));
  g->C = C;
  g->p.bo = bo; g->p.b1 = b1; g->p.b2 = b2; g->p.ln_g = ln_g; g->p.ln_b = ln_b;
  if (!encode_2d(enc, &g->map_att, att, C, Mcap, 128) || !encode_2d(enc, &g->map_wo, wo, C, C, C) ||
      !encode_2d(enc, &g->map_w1, w1, C, C, C) || !encode_2d(enc, &g->map_w2, w2, C, C, C)) {
    delete g;
    return nullptr;
  }
  return g;
}
void attn_tail_destroy(AttnTail* g) { delete g; }

void attn_tail_launch(const AttnTail* g, const bf16* x, int ld_x, bf16* out, int ld_out, long long M, cudaStream_t s) {
  kernels_count_launch();
  TailParams p = g->p;
  p.x = x; p.ld_x = ld_x; p.out = out; p.ld_out = ld_out;
  if (g->C == 64) launch_tail<64>(g, p, M, s);
  else if (g->C == 128) launch_tail<128>(g, p, M, s);
  else launch_tail<256>(g, p, M, s);
}
