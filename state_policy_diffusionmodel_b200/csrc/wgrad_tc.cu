// wgrad_tc.cu — tcgen05 / TMEM / TMA weight-gradient GEMM for sm_100a.
//
//   dw[tap][co][ci] += sum_p  x[shift(p, tap)][ci] * dy[p][co]          (3x3 pad-1 conv, taps == 9; Linear, taps == 1)
//
// the backward-by-weights of the convolutions of models/Unet_FiLmLayer.py:101,103 and of the Linear layers of the
// SelfAttention blocks (:50,54-57).  The contraction runs over PIXELS, which is the slow (row) index of the channels-last
// activations, so both operands are MN-major for the tensor core (channels contiguous):
//   A = x^T  : M = 128 input channels (two 64-channel swizzle atoms), K = 64 pixels per stage
//   B = dy^T : N = 64..256 output channels,                           K = 64 pixels per stage
// Every 64-channel atom of a stage is one 4-D TMA box {64 ch, W, Hb, Bt} (Hb*W*Bt == 64 pixels) of the (C, W, H, B) view of
// the map; for A the box sits at (c0, dx, h0 + dy, b0), i.e. the tap-shifted window, with the convolution's zero padding
// supplied by TMA out-of-bounds fill (the same trick as the forward kernel, conv_tc.cu).  In shared memory a box is 64 rows
// (pixels) of 128 bytes under the 128-byte swizzle = the canonical MN-major SWIZZLE_128B layout
//   ((8,8,m),(8,k)) : ((1,8,LBO),(64,SBO))   [bf16 elements]   LBO = 8 KB (next 64-channel atom), SBO = 1 KB (next 8 pixels)
// Accumulator: 128 TMEM lanes (ci) x N columns (co), fp32.  Work unit = (tap, ci tile, co tile, pixel slice); one unit per CTA.
// Epilogue: tcgen05.ld -> red.global.add.f32 into dw[tap][co][ci]: a warp's 32 lanes hit 32 consecutive ci -> one 128-byte
// reduction per (warp, co).  warp 0: TMA producer, warp 1: MMA issuer, warps 2..5: epilogue.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <tuple>

#include "tc_ptx.cuh"
#include "train.cuh"

static thread_local char g_wg_err[256] = "";
const char* wgrad_tc_last_error() { return g_wg_err; }
long long wgrad_tc_launch_count_value = 0;

namespace {
constexpr int WG_KB = 64;                         // pixels per pipeline stage
constexpr int WG_ATOM_BYTES = WG_KB * 128;        // one {64 ch x 64 px} box
constexpr int WG_A_BYTES = 2 * WG_ATOM_BYTES;     // 128 input channels
constexpr int WG_THREADS = 192;

struct WgParams {
  int H, W, Hb, Bt;         // Hb*W*Bt == 64
  int Cin, Cout, taps;
  int n_tile;               // co per unit: 64, 128 or 256
  int ci_tiles, co_tiles;
  int ntx, nty;             // valid taps along x / y (1 when W == 1 / H == 1)
  int kblocks;              // 64-pixel blocks over the whole batch
  int ksplit;
  uint32_t lbo, sbo;        // descriptor fields (>> 4)
  float* dw;                // [taps][Cout][Cin] fp32, accumulated
};

// MN-major, 128B-swizzled operand: see file header
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t lbo16, uint32_t sbo16) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(lbo16 & 0x3FFF) << 16;
  d |= (uint64_t)(sbo16 & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16, D = f32, A = B = bf16, both MN-major (bits 15, 16), M = 128
__host__ __device__ constexpr uint32_t make_idesc_mn(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <int N_TILE, int STAGES>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy, const WgParams p) {
  constexpr int B_BYTES = (N_TILE / 64) * WG_ATOM_BYTES;
  constexpr int TMEM_COLS = N_TILE <= 64 ? 64 : (N_TILE <= 128 ? 128 : 256);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * WG_A_BYTES;
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t acc_bar;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- decode the work unit ----
  int u = blockIdx.x;
  const int ks = u % p.ksplit; u /= p.ksplit;
  const int co_t = u % p.co_tiles; u /= p.co_tiles;
  const int ci_t = u % p.ci_tiles; u /= p.ci_tiles;
  const int tap_i = u;
  int dy = 0, dx = 0, tap = 0;
  if (p.taps == 9) {
    const int ty = tap_i / p.ntx, tx = tap_i - ty * p.ntx;
    dy = p.nty == 1 ? 0 : ty - 1;
    dx = p.ntx == 1 ? 0 : tx - 1;
    tap = (dy + 1) * 3 + (dx + 1);
  }
  const int kb_begin = (int)((long long)ks * p.kblocks / p.ksplit), kb_end = (int)((long long)(ks + 1) * p.kblocks / p.ksplit);
  const int ci0 = ci_t * 128, co0 = co_t * N_TILE;
  const int blocks_per_sample = p.H / p.Hb;  // > 1 only when Bt == 1

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_dy);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(&tmem_base_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t kit = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb, ++kit) {
        const int s = kit % STAGES;
        const uint32_t ph = (kit / STAGES) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        int b0, h0;
        if (blocks_per_sample > 1) { b0 = kb / blocks_per_sample; h0 = (kb - b0 * blocks_per_sample) * p.Hb; }
        else { b0 = kb * p.Bt; h0 = 0; }
        mbar_expect_tx(&full_bar[s], WG_A_BYTES + B_BYTES);
        uint8_t* a = smem_a + s * WG_A_BYTES;
        uint8_t* b = smem_b + s * B_BYTES;
        // A: x^T, two 64-channel atoms at the tap-shifted window (channels past Cin are zero-filled)
        tma_load_4d(a, &map_x, &full_bar[s], ci0, dx, h0 + dy, b0);
        tma_load_4d(a + WG_ATOM_BYTES, &map_x, &full_bar[s], ci0 + 64, dx, h0 + dy, b0);
#pragma unroll
        for (int j = 0; j < N_TILE / 64; ++j) tma_load_4d(b + j * WG_ATOM_BYTES, &map_dy, &full_bar[s], co0 + 64 * j, 0, h0, b0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_mn(N_TILE);
      uint32_t kit = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb, ++kit) {
        const int s = kit % STAGES;
        const uint32_t ph = (kit / STAGES) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem_a + s * WG_A_BYTES), b_addr = smem_u32(smem_b + s * B_BYTES);
#pragma unroll
        for (int k = 0; k < WG_KB / 16; ++k) {  // 16 pixels per MMA = two 8-pixel groups = 2 KB
          const uint64_t da = make_desc_mn(a_addr + k * 2048, p.lbo, p.sbo);
          const uint64_t db = make_desc_mn(b_addr + k * 2048, p.lbo, p.sbo);
          umma_bf16(tmem_base, da, db, idesc, (kit > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&acc_bar);
    }
  } else {
    // ---- epilogue: lane quarter q of TMEM = input channels ci0 + 32q .. +31 ----
    const int q = warp & 3;
    const int ci = ci0 + q * 32 + lane;
    if (kb_end > kb_begin) {
      mbar_wait(&acc_bar, 0);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16);
      float* base = p.dw + (size_t)tap * p.Cout * p.Cin + ci;
#pragma unroll 1
      for (int c = 0; c < N_TILE; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_addr + (uint32_t)c, v);
        tmem_ld_wait();
        if (ci < p.Cin) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int co = co0 + c + i;
            if (co < p.Cout) atomicAdd(base + (size_t)co * p.Cin, __uint_as_float(v[i]));
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc<TMEM_COLS>(tmem_base); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

struct WgPlan { CUtensorMap map_x, map_dy; WgParams p; int stages; size_t smem; };
typedef std::tuple<const void*, int, const void*, int, long long, int, int, int, int, int> WgKey;
std::map<WgKey, WgPlan*>& wg_cache() { static std::map<WgKey, WgPlan*> c; return c; }

bool decompose(int H, int W, long long M, int taps, int* Hb, int* Bt, long long* B) {
  if (taps == 1) { *Hb = 1; *Bt = WG_KB; *B = M; return M % WG_KB == 0; }
  if (W > WG_KB || WG_KB % W) return false;
  int hb = WG_KB / W;
  if (hb > H) hb = H;
  if (H % hb || WG_KB % (hb * W)) return false;
  *Hb = hb; *Bt = WG_KB / (hb * W);
  *B = M / ((long long)H * W);
  return (*B) % (*Bt) == 0 && M % ((long long)H * W) == 0;
}

bool make_map(CUtensorMap* map, const bf16* ptr, int C, int ld, int H, int W, long long B, int Hb, int Bt) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)W, (cuuint32_t)Hb, (cuuint32_t)Bt};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

bool wgrad_tc_supported(int Cin, int Cout, int H, int W, int taps, int ld_x, int ld_dy, long long M) {
  int Hb, Bt;
  long long B;
  if (Cin % 64 || Cout % 64 || ld_x % 8 || ld_dy % 8 || (taps != 1 && taps != 9)) return false;
  return decompose(taps == 1 ? 1 : H, taps == 1 ? 1 : W, M, taps, &Hb, &Bt, &B);
}

int wgrad_tc_launch(const bf16* x, int ld_x, const bf16* dy, int ld_dy, long long M, int Cin, int Cout, int H, int W, int taps, float* dw,
                    cudaStream_t s) {
  if (taps == 1) { H = 1; W = 1; }
  const WgKey key(x, ld_x, dy, ld_dy, M, Cin, Cout, H, W, taps);
  WgPlan*& pl = wg_cache()[key];
  if (!pl) {
    if (!wgrad_tc_supported(Cin, Cout, H, W, taps, ld_x, ld_dy, M)) {
      snprintf(g_wg_err, sizeof g_wg_err, "wgrad_tc: unsupported shape Cin=%d Cout=%d H=%d W=%d taps=%d M=%lld", Cin, Cout, H, W, taps, M);
      return -1;
    }
    WgPlan* n = new WgPlan();
    memset(n, 0, sizeof(*n));
    WgParams& p = n->p;
    long long B;
    decompose(H, W, M, taps, &p.Hb, &p.Bt, &B);
    p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.taps = taps;
    p.n_tile = Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : 64);
    p.ci_tiles = (Cin + 127) / 128;
    p.co_tiles = Cout / p.n_tile;
    p.ntx = (taps == 9 && W > 1) ? 3 : 1;
    p.nty = (taps == 9 && H > 1) ? 3 : 1;
    p.kblocks = (int)(M / WG_KB);
    const int units = p.ntx * p.nty * p.ci_tiles * p.co_tiles;
    // one CTA per SM (192 KB of pipeline stages): keep the grid within one wave, every CTA does the same amount of work
    int sms = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    int ksplit = sms / units;
    if (ksplit > p.kblocks / 4) ksplit = p.kblocks / 4;   // at least four 64-pixel blocks per unit
    if (ksplit < 1) ksplit = 1;
    p.ksplit = ksplit;
    p.lbo = WG_ATOM_BYTES >> 4;
    p.sbo = 1024 >> 4;
    if (const char* e = getenv("SPDM_WGRAD_SWAP_LBO")) { if (atoi(e)) { const uint32_t t = p.lbo; p.lbo = p.sbo; p.sbo = t; } }
    if (!make_map(&n->map_x, x, Cin, ld_x, H, W, B, p.Hb, p.Bt) || !make_map(&n->map_dy, dy, Cout, ld_dy, H, W, B, p.Hb, p.Bt)) {
      snprintf(g_wg_err, sizeof g_wg_err, "wgrad_tc: cuTensorMapEncodeTiled failed");
      delete n;
      return -1;
    }
    const int b_bytes = (p.n_tile / 64) * WG_ATOM_BYTES;
    n->stages = 4;
    n->smem = (size_t)n->stages * (WG_A_BYTES + b_bytes) + 1024;
    pl = n;
  }
  WgParams p = pl->p;
  p.dw = dw;
  const int grid = p.ntx * p.nty * p.ci_tiles * p.co_tiles * p.ksplit;
#define WG_LAUNCH(NT)                                                                                               \
  {                                                                                                                 \
    static bool attr = false;                                                                                       \
    if (!attr) { cudaFuncSetAttribute(wgrad_tc_kernel<NT, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (WG_A_BYTES + (NT / 64) * WG_ATOM_BYTES) + 1024); attr = true; } \
    wgrad_tc_kernel<NT, 4><<<grid, WG_THREADS, pl->smem, s>>>(pl->map_x, pl->map_dy, p);                             \
  }
  if (p.n_tile == 256) WG_LAUNCH(256)
  else if (p.n_tile == 128) WG_LAUNCH(128)
  else WG_LAUNCH(64)
#undef WG_LAUNCH
  ++wgrad_tc_launch_count_value;
  return 0;
}
