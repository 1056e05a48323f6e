// conv_tc.cu — tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a.
//
// Computes  out[r, n] = epi( sum_{tap, c} in[shift(r, tap), c] * w[n][tap*Cin + c] )  in bf16 with fp32
// accumulation in tensor memory, for the 3x3 pad-1 bias-free convolutions of the FiLM U-Net
// (reference models/Unet_FiLmLayer.py:101,103) and, with taps == 1, for the Linear layers of the
// SelfAttention blocks (:50,54-57).
//
//   M tile  = 128 output pixels = Bt samples x Hb rows x W columns of the channels-last map
//   A tile  = one 4-D TMA box {64 ch, W, Hb, Bt} at (c0, dx, h0+dy, b0): the shifted window of tap
//             (dy,dx); the convolution's zero padding is TMA out-of-bounds fill.  128B swizzle.
//   B tile  = 2-D TMA box {64 k, BLOCK_N} of the [Cout][taps*Cin] weight matrix (K-major).
//   MMA     = tcgen05.mma.cta_group::1.kind::f16, M=128, N=BLOCK_N, K=16, issued by one thread;
//             accumulator = BLOCK_N TMEM columns.
//   roles   = warp 0: TMA producer, warp 1: MMA issuer, warps 2..5: epilogue (tcgen05.ld -> bias /
//             GELU / residual / GroupNorm partial sums -> bf16 global stores).
#include <cuda.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"

static long long g_tc_launches = 0;
long long tc_launch_count() { return g_tc_launches; }
static thread_local char g_tc_err[256] = "";
const char* tc_last_error() { return g_tc_err; }

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // bf16 elements = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int NUM_THREADS = 192;

struct TcParams {
  int H, W, Hb, Bt;       // sample geometry and M-tile decomposition (Hb*W*Bt == 128)
  int Cin, Cout;
  int taps;               // 1 or 9
  int kb_per_tap;         // Cin / 64
  int n_tiles;
  int P;                  // stats partial slots per sample
  int ld_out, ld_res;
  int flags;
  bf16* out;
  float* stats;
  const float* bias;
  const bf16* resid;
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap (error surfaced to the host), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) { printf("spdm conv_tc: mbarrier timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x); __trap(); }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

template <int COLS> __device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS> __device__ __forceinline__ void tmem_dealloc(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T   (both operands K-major)
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 columns of fp32: thread i of the warp gets row (lane base + i), 32 consecutive columns
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t v[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row swizzle atoms 1024 bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);   // start address        bits [0,14)
  d |= (uint64_t)0 << 16;                        // leading byte offset  bits [16,30)  (unused: one atom along K)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset   bits [32,46)
  d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=n
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// The kernel.  grid = (m_tiles, n_tiles); dynamic smem = STAGES*(A+B) + 1024 (alignment slack).
// ---------------------------------------------------------------------------------------------
template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcParams p) {
  constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- tile coordinates ----
  const int m_tile = blockIdx.x, n_tile = blockIdx.y;
  const int tiles_per_sample = p.H / p.Hb;  // > 1 only when Bt == 1
  int b0, h0;
  if (tiles_per_sample > 1) { b0 = m_tile / tiles_per_sample; h0 = (m_tile - b0 * tiles_per_sample) * p.Hb; }
  else { b0 = m_tile * p.Bt; h0 = 0; }
  const int n0 = n_tile * BLOCK_N;

  // ---- K iteration space: valid taps x 64-channel blocks ----
  const bool skip_dx = (p.taps == 9 && p.W == 1), skip_dy = (p.taps == 9 && p.H == 1);
  const int ntx = (p.taps == 9) ? (skip_dx ? 1 : 3) : 1;
  const int nty = (p.taps == 9) ? (skip_dy ? 1 : 3) : 1;
  const int k_iters = ntx * nty * p.kb_per_tap;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<BLOCK_N>(&tmem_base_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      for (int it = 0; it < k_iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        const int tap_i = it / p.kb_per_tap, kb = it - tap_i * p.kb_per_tap;
        int dy = 0, dx = 0, tap = 0;
        if (p.taps == 9) {
          const int ty = tap_i / ntx, tx = tap_i - ty * ntx;
          dy = skip_dy ? 0 : ty - 1;
          dx = skip_dx ? 0 : tx - 1;
          tap = (dy + 1) * 3 + (dx + 1);
        }
        mbar_expect_tx(&full_bar[s], A_STAGE_BYTES + B_STAGE_BYTES);
        tma_load_4d(smem_a + s * A_STAGE_BYTES, &map_a, &full_bar[s], kb * BLOCK_K, dx, h0 + dy, b0);
        tma_load_2d(smem_b + s * B_STAGE_BYTES, &map_b, &full_bar[s], tap * p.Cin + kb * BLOCK_K, n0);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BLOCK_N);
      for (int it = 0; it < k_iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint64_t da = make_smem_desc(smem_u32(smem_a + s * A_STAGE_BYTES));
        const uint64_t db = make_smem_desc(smem_u32(smem_b + s * B_STAGE_BYTES));
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          // advance 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
          umma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem slot when these MMAs retire
      }
      umma_commit(&tmem_full_bar);   // accumulator complete
    }
  } else {
    // ================= epilogue: warps 2..5 -> TMEM lane quarters (warp % 4) =================
    const int q = warp & 3;
    const int r_t = q * 32 + lane;                 // tile row == TMEM lane
    const int rps = p.Hb * p.W;                    // rows per sample inside the tile
    const long long row0 = ((long long)b0 * p.H + h0) * p.W;  // tile rows are contiguous in the [M, C] map
    const long long row = row0 + r_t;
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    float rs = 0.f, rq = 0.f;
    bf16* orow = p.out + row * p.ld_out + n0;
    const bf16* rrow = (p.flags & EPI_RESID) ? p.resid + row * p.ld_res + n0 : nullptr;
#pragma unroll 1
    for (int c = 0; c < BLOCK_N; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      tmem_ld_wait();
      float f[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
      if (p.flags & EPI_BIAS) {
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] += __ldg(p.bias + n0 + c + i);
      }
      if (p.flags & EPI_GELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = gelu_exact(f[i]);
      }
      if (rrow) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          float t[8];
          load8(rrow + c + i, t);
#pragma unroll
          for (int e = 0; e < 8; ++e) f[i + e] += t[e];
        }
      }
      if (p.flags & EPI_STATS) {
#pragma unroll
        for (int i = 0; i < 32; ++i) { rs += f[i]; rq = fmaf(f[i], f[i], rq); }
      }
#pragma unroll
      for (int i = 0; i < 32; i += 8) store8(orow + c + i, f + i);
    }
    if (p.flags & EPI_STATS) {
      // reduce over the lanes that belong to the same sample, one deterministic slot per writer
      const int span = rps >= 32 ? 32 : rps;  // rps in {128, 64, 32, 16, 8, 4, ...}: power of two
      for (int o = span >> 1; o > 0; o >>= 1) { rs += __shfl_xor_sync(0xffffffffu, rs, o); rq += __shfl_xor_sync(0xffffffffu, rq, o); }
      if ((lane & (span - 1)) == 0) {
        const int b = b0 + r_t / rps;
        int slot;
        if (rps >= 32) {
          const int warps_per_sample_tile = (rps >= 128 ? 128 : rps) / 32;
          const int tile_in_sample = tiles_per_sample > 1 ? (m_tile % tiles_per_sample) : 0;
          const int warp_in_sample = (r_t % rps) / 32;
          slot = (tile_in_sample * warps_per_sample_tile + warp_in_sample) * p.n_tiles + n_tile;
        } else {
          slot = n_tile;
        }
        float* dst = p.stats + ((size_t)b * p.P + slot) * 2;
        dst[0] = rs;
        dst[1] = rq;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<BLOCK_N>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
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

template <int BLOCK_N, int STAGES> constexpr int smem_bytes() { return STAGES * (A_STAGE_BYTES + BLOCK_N * BLOCK_K * 2) + 1024; }

template <int BLOCK_N, int STAGES>
void launch_cfg(const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& p, int m_tiles, cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<BLOCK_N, STAGES>());
    attr = true;
  }
  dim3 grid(m_tiles, p.n_tiles);
  conv_tc_kernel<BLOCK_N, STAGES><<<grid, NUM_THREADS, smem_bytes<BLOCK_N, STAGES>(), s>>>(ma, mb, p);
  ++g_tc_launches;
}

}  // namespace

struct TcGemm {
  CUtensorMap map_a, map_b;
  TcParams p;
  int block_n;
  int Bcap;
};

int tc_batch_multiple(int H, int W) {
  const int hw = H * W;
  return hw >= BLOCK_M ? 1 : BLOCK_M / hw;
}

TcGemm* tc_gemm_create(const bf16* in, int ld_in, const bf16* w_packed, int Cin, int Cout, int taps, int H, int W, int Bcap) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { snprintf(g_tc_err, sizeof g_tc_err, "cuTensorMapEncodeTiled entry point not available"); return nullptr; }
  if (Cin % BLOCK_K || Cout % 64 || (taps != 1 && taps != 9) || ld_in % 8) {
    snprintf(g_tc_err, sizeof g_tc_err, "tc_gemm_create: unsupported shape Cin=%d Cout=%d taps=%d ld=%d", Cin, Cout, taps, ld_in);
    return nullptr;
  }
  if (BLOCK_M % W || W > BLOCK_M) { snprintf(g_tc_err, sizeof g_tc_err, "tc_gemm_create: W=%d does not divide 128", W); return nullptr; }
  int Hb = BLOCK_M / W;
  if (Hb > H) Hb = H;
  if (H % Hb || BLOCK_M % (Hb * W)) { snprintf(g_tc_err, sizeof g_tc_err, "tc_gemm_create: H=%d W=%d not tileable", H, W); return nullptr; }
  const int Bt = BLOCK_M / (Hb * W);
  if (Bcap % Bt) { snprintf(g_tc_err, sizeof g_tc_err, "tc_gemm_create: Bcap=%d not a multiple of %d", Bcap, Bt); return nullptr; }

  TcGemm* g = new TcGemm();
  memset(g, 0, sizeof(*g));
  g->Bcap = Bcap;
  g->block_n = (Cout % 128 == 0) ? 128 : 64;
  TcParams& p = g->p;
  p.H = H; p.W = W; p.Hb = Hb; p.Bt = Bt; p.Cin = Cin; p.Cout = Cout; p.taps = taps;
  p.kb_per_tap = Cin / BLOCK_K;
  p.n_tiles = Cout / g->block_n;
  const int rps = Hb * W;
  p.P = rps >= 32 ? (H / Hb) * ((rps >= 128 ? 128 : rps) / 32) * p.n_tiles : p.n_tiles;

  {  // A: 4-D (C, W, H, B) view of the channels-last activation
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Bcap};
    cuuint64_t strides[3] = {(cuuint64_t)ld_in * 2, (cuuint64_t)W * ld_in * 2, (cuuint64_t)H * W * ld_in * 2};
    cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)W, (cuuint32_t)Hb, (cuuint32_t)Bt};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&g->map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(g_tc_err, sizeof g_tc_err, "cuTensorMapEncodeTiled(A) failed: %d", (int)r); delete g; return nullptr; }
  }
  {  // B: 2-D (K, Cout) weights
    const cuuint64_t Ktot = (cuuint64_t)taps * Cin;
    cuuint64_t dims[2] = {Ktot, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {Ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)g->block_n};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&g->map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)w_packed, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(g_tc_err, sizeof g_tc_err, "cuTensorMapEncodeTiled(B) failed: %d", (int)r); delete g; return nullptr; }
  }
  return g;
}

void tc_gemm_destroy(TcGemm* g) { delete g; }
int tc_gemm_partials(const TcGemm* g) { return g->p.P; }

void tc_gemm_launch(const TcGemm* g, bf16* out, int ld_out, float* stats, const float* bias, const bf16* resid, int ld_res, int flags,
                    int B, cudaStream_t s) {
  TcParams p = g->p;
  p.out = out; p.ld_out = ld_out; p.stats = stats; p.bias = bias; p.resid = resid; p.ld_res = ld_res; p.flags = flags;
  const int m_tiles = (int)(((long long)B * p.H * p.W) / BLOCK_M);
  if (g->block_n == 128) launch_cfg<128, 6>(g->map_a, g->map_b, p, m_tiles, s);
  else launch_cfg<64, 8>(g->map_a, g->map_b, p, m_tiles, s);
}
