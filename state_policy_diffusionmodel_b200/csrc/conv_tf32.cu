// conv_tf32.cu — tcgen05 kind::tf32 implicit-GEMM 3x3 convolution: the TF32 precision mode of the denoising path (sm_100a).
//
// north_star allows "the bf16/TF32 path" for the tensor-core contractions (final trajectory within rel 1e-2).  This is the TF32
// one: activations stay fp32 in HBM exactly as on the fp32 parity path (same elementwise / norm / attention kernels), only the
// 3x3 convolutions (91 % of the FLOPs) move from CUDA cores to the tensor cores:
//   out[r, n] = sum_{tap, c} in[shift(r, tap), c] * w[n][tap*Cin + c]        fp32 operands, read as TF32 (10-bit mantissa) by
//                                                                             tcgen05.mma.kind::tf32, fp32 accumulation in TMEM
//   A tile = one 4-D TMA box {32 ch, W, Hb, Bt} of fp32 (128 bytes per pixel row, 128B swizzle; conv zero padding = TMA OOB fill)
//   B tile = 2-D TMA box {32 k, BLOCK_N} of the fp32 weight matrix [Cout][9*Cin] (K-major)
//   MMA    = M 128, N = BLOCK_N, K = 8 per instruction (32 bytes): four per 32-channel k-step
// Same persistent warp-specialised pipeline as conv_tc_kernel (conv_tc.cu): warp 0 TMA, warp 1 MMA, warps 2..5 epilogue, double
// buffered accumulators.  The epilogue stores the raw fp32 conv output; GroupNorm statistics / apply are the fp32 path's kernels.
#include <cuda.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int TF_BLOCK_M = 128;
constexpr int TF_BLOCK_K = 32;                              // fp32 elements = 128 bytes = one swizzle row
constexpr int TF_UMMA_K = 8;
constexpr int TF_A_STAGE = TF_BLOCK_M * TF_BLOCK_K * 4;     // 16 KB
constexpr int TF_THREADS = 192;

struct TfParams {
  int H, W, Hb, Bt, Cin, Cout, kb_per_tap, n_tiles, m_tiles, total_tiles, ld_out;
  float* out;
};

// kind::tf32 instruction descriptor: D = f32, A = B = tf32, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t make_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(TF_THREADS, 1)
conv_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TfParams p) {
  constexpr int B_STAGE = BLOCK_N * TF_BLOCK_K * 4;
  constexpr int TMEM_COLS = 2 * BLOCK_N <= 128 ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * TF_A_STAGE;
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool skip_dx = p.W == 1, skip_dy = p.H == 1;
  const int ntx = skip_dx ? 1 : 3, nty = skip_dy ? 1 : 3;
  const int k_iters = ntx * nty * p.kb_per_tap;
  const int tiles_per_sample = p.H / p.Hb;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tmem_full_bar[0], 1); mbar_init(&tmem_full_bar[1], 1);
    mbar_init(&tmem_empty_bar[0], 4); mbar_init(&tmem_empty_bar[1], 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(&tmem_base_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();
  pdl_trigger();

  auto tile_coord = [&](int tile, int& n0, int& b0, int& h0) {
    const int n_tile = tile % p.n_tiles, m_tile = tile / p.n_tiles;
    n0 = n_tile * BLOCK_N;
    if (tiles_per_sample > 1) { b0 = m_tile / tiles_per_sample; h0 = (m_tile - b0 * tiles_per_sample) * p.Hb; }
    else { b0 = m_tile * p.Bt; h0 = 0; }
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t kit = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int n0, b0, h0;
        tile_coord(tile, n0, b0, h0);
        for (int it = 0; it < k_iters; ++it, ++kit) {
          const int s = kit % STAGES;
          const uint32_t ph = (kit / STAGES) & 1u;
          const int tap_i = it / p.kb_per_tap, kb = it - tap_i * p.kb_per_tap;
          const int ty = tap_i / ntx, tx = tap_i - ty * ntx;
          const int dy = skip_dy ? 0 : ty - 1, dx = skip_dx ? 0 : tx - 1;
          const int tap = (dy + 1) * 3 + (dx + 1);
          mbar_wait(&empty_bar[s], ph ^ 1u);
          mbar_expect_tx(&full_bar[s], TF_A_STAGE + B_STAGE);
          tma_load_2d(smem_b + s * B_STAGE, &map_b, &full_bar[s], tap * p.Cin + kb * TF_BLOCK_K, n0);
          tma_load_4d(smem_a + s * TF_A_STAGE, &map_a, &full_bar[s], kb * TF_BLOCK_K, dx, h0 + dy, b0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(BLOCK_N);
      uint32_t kit = 0;
      int lt = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
        const int acc = lt & 1;
        const uint32_t aph = (uint32_t)(lt >> 1) & 1u;
        mbar_wait(&tmem_empty_bar[acc], aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int it = 0; it < k_iters; ++it, ++kit) {
          const int s = kit % STAGES;
          const uint32_t ph = (kit / STAGES) & 1u;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint64_t da = make_smem_desc(smem_u32(smem_a + s * TF_A_STAGE));
          const uint64_t db = make_smem_desc(smem_u32(smem_b + s * B_STAGE));
#pragma unroll
          for (int k = 0; k < TF_BLOCK_K / TF_UMMA_K; ++k)   // 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
            umma_tf32(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (it > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else {
    const int q = warp & 3;
    const int r_t = q * 32 + lane;
    int lt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
      int n0, b0, h0;
      tile_coord(tile, n0, b0, h0);
      const int acc = lt & 1;
      const uint32_t aph = (uint32_t)(lt >> 1) & 1u;
      const long long row = ((long long)b0 * p.H + h0) * p.W + r_t;   // tile rows are contiguous in the [M, C] map
      float* orow = p.out + row * p.ld_out + n0;
      mbar_wait(&tmem_full_bar[acc], aph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_addr + (uint32_t)c, v);
        tmem_ld_wait();
        if (c + 32 == BLOCK_N) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<uint4*>(orow + c + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tf() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) return nullptr;
    fn = (EncodeTiledFn)p;
  }
  return fn;
}
int tf_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}
thread_local char g_tf_err[256] = "";

template <int BLOCK_N, int STAGES>
void tf_launch(const CUtensorMap& ma, const CUtensorMap& mb, const TfParams& p, cudaStream_t s) {
  constexpr int smem = STAGES * (TF_A_STAGE + BLOCK_N * TF_BLOCK_K * 4) + 1024;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(conv_tf32_kernel<BLOCK_N, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
  const int grid = p.total_tiles < tf_num_sms() ? p.total_tiles : tf_num_sms();
  launch_pdl(conv_tf32_kernel<BLOCK_N, STAGES>, dim3(grid), dim3(TF_THREADS), smem, s, ma, mb, p);
}

}  // namespace

struct TfGemm {
  CUtensorMap map_a, map_b;
  TfParams p;
  int block_n;
};

const char* tf32_last_error() { return g_tf_err; }

// in: fp32 [Bcap*H*W, Cin] (ld_in); w_packed: fp32 [Cout][9*Cin] (launch_pack_conv_tf32)
TfGemm* tf32_conv_create(const float* in, int ld_in, const float* w_packed, int Cin, int Cout, int H, int W, int Bcap) {
  EncodeTiledFn enc = get_encode_tf();
  if (!enc) { snprintf(g_tf_err, sizeof g_tf_err, "cuTensorMapEncodeTiled entry point not available"); return nullptr; }
  if (Cin % TF_BLOCK_K || Cout % 64 || ld_in % 4) { snprintf(g_tf_err, sizeof g_tf_err, "tf32 conv: unsupported shape Cin=%d Cout=%d ld=%d", Cin, Cout, ld_in); return nullptr; }
  if (TF_BLOCK_M % W || W > TF_BLOCK_M) { snprintf(g_tf_err, sizeof g_tf_err, "tf32 conv: W=%d does not divide 128", W); return nullptr; }
  int Hb = TF_BLOCK_M / W;
  if (Hb > H) Hb = H;
  if (H % Hb || TF_BLOCK_M % (Hb * W)) { snprintf(g_tf_err, sizeof g_tf_err, "tf32 conv: H=%d W=%d not tileable", H, W); return nullptr; }
  const int Bt = TF_BLOCK_M / (Hb * W);
  if (Bcap % Bt) { snprintf(g_tf_err, sizeof g_tf_err, "tf32 conv: Bcap=%d not a multiple of %d", Bcap, Bt); return nullptr; }
  TfGemm* g = new TfGemm();
  memset(g, 0, sizeof *g);
  g->block_n = Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : 64);
  TfParams& p = g->p;
  p.H = H; p.W = W; p.Hb = Hb; p.Bt = Bt; p.Cin = Cin; p.Cout = Cout; p.kb_per_tap = Cin / TF_BLOCK_K;
  p.n_tiles = Cout / g->block_n;
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Bcap};
    cuuint64_t strides[3] = {(cuuint64_t)ld_in * 4, (cuuint64_t)W * ld_in * 4, (cuuint64_t)H * W * ld_in * 4};
    cuuint32_t box[4] = {(cuuint32_t)TF_BLOCK_K, (cuuint32_t)W, (cuuint32_t)Hb, (cuuint32_t)Bt};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&g->map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(g_tf_err, sizeof g_tf_err, "cuTensorMapEncodeTiled(A, tf32) failed: %d", (int)r); delete g; return nullptr; }
  }
  {
    const cuuint64_t Ktot = (cuuint64_t)9 * Cin;
    cuuint64_t dims[2] = {Ktot, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {Ktot * 4};
    cuuint32_t box[2] = {(cuuint32_t)TF_BLOCK_K, (cuuint32_t)g->block_n};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&g->map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)w_packed, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(g_tf_err, sizeof g_tf_err, "cuTensorMapEncodeTiled(B, tf32) failed: %d", (int)r); delete g; return nullptr; }
  }
  return g;
}
void tf32_conv_destroy(TfGemm* g) { delete g; }

// out fp32 [B*H*W, Cout] (ld_out): the raw conv output.  B must be a multiple of tc_batch_multiple(H, W).
void tf32_conv_launch(const TfGemm* g, float* out, int ld_out, int B, cudaStream_t s) {
  TfParams p = g->p;
  p.out = out; p.ld_out = ld_out;
  p.m_tiles = (int)(((long long)B * p.H * p.W) / TF_BLOCK_M);
  p.total_tiles = p.m_tiles * p.n_tiles;
  if (g->block_n == 256) tf_launch<256, 4>(g->map_a, g->map_b, p, s);
  else if (g->block_n == 128) tf_launch<128, 6>(g->map_a, g->map_b, p, s);
  else tf_launch<64, 8>(g->map_a, g->map_b, p, s);
  kernels_count_launch();
}
