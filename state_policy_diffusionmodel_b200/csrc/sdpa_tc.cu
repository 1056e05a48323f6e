// sdpa_tc.cu — attention core of the SelfAttention blocks on tcgen05 (sm_100a).
//
//   o[b, i, h] = softmax_j( q_i . k_j / sqrt(hd) ) v_j        nn.MultiheadAttention(channels, 4, batch_first=True),
//                                                              reference models/Unet_FiLmLayer.py:50,76
// The token count per sample is tiny (L = H*W in {4,16,64,256}), so a whole softmax row fits in tensor memory:
//   CTA        = one 128-token query tile x one 64-channel block (= 64/hd heads) of one attention tile
//   keys       = the LK in {128,256} tokens of the attention tile: 128/L whole samples when L <= 128 (block-diagonal
//                mask in the softmax), or the 256 tokens of the query's sample when L == 256
//   S = Q K^T  : tcgen05.mma M=128, N=LK, K=hd; Q and K tiles are plain 2-D TMA boxes of the qkv activation
//                ([tokens][3C], K-major, 128B swizzle); a head is a 2*hd-byte column offset inside the 128-byte row
//   softmax    : one thread per query row reads its S row from TMEM (two passes: max, then exp2/sum), writes P as
//                bf16 into shared memory in the canonical K-major 128B-swizzled layout
//   O = P V    : tcgen05.mma M=128, N=hd, K=LK; V^T tiles ([C][LK], K-major) are written by the in_proj GEMM
//                epilogue (EPI_VT), so both operands of both GEMMs are K-major
//   epilogue   : O row / rowsum -> bf16 -> att[token][C]
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

// 2^x through MUFU.EX2 directly (arguments are <= 0 here; exp2f() adds range-handling instructions the softmax loop is bound by)
// three-input maximum (one FMNMX3 on sm_100)
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

namespace {

struct SdpaParams {
  int C, L, hd, heads_per_blk;
  float scale_log2;  // log2(e) / sqrt(hd)
  bf16* out;         // [M][C]
};

// 256 threads: warps 0-3 and 4-7 both map onto the four TMEM lane quarters (warp % 4); the two groups split the key
// columns of every softmax row between them, which halves the per-thread instruction stream and gives each SM
// sub-partition two warps to hide the TMEM / MUFU latencies.
template <int LK, bool MASKED>
__global__ void __launch_bounds__(256)
sdpa_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_vt, const SdpaParams p) {
  constexpr int KB = LK / 64;  // key blocks of 64
  constexpr int SQ = 16384, SK = LK * 128, SVT = KB * 8192;
  constexpr int TMEM_COLS = (LK == 256) ? 512 : 256;  // S: LK columns, O: up to 64 columns
  constexpr int HALF = LK / 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + SQ;
  uint8_t* sVt = sK + SK;
  uint8_t* sP = sVt + SVT;
  __shared__ __align__(8) uint64_t bar_load;
  __shared__ __align__(8) uint64_t bar_mma;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_red[2][128];   // per-row partial max / partial sum of the two column halves

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = warp >> 2;             // which half of the key columns this thread handles
  const int r = (warp & 3) * 32 + lane;   // query row inside the tile == TMEM lane
  const int q_row0 = blockIdx.x * 128;
  const int cb = blockIdx.y;
  const int kv_tile = q_row0 / LK;
  const int key_row0 = kv_tile * LK;

  if (tid == 0) {
    tma_prefetch_desc(&map_qkv);
    tma_prefetch_desc(&map_vt);
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<TMEM_COLS>(&tmem_base_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_smem;
  pdl_wait();
  pdl_trigger();

  if (tid == 0) {
    mbar_expect_tx(&bar_load, SQ + SK + SVT);
    tma_load_2d(sQ, &map_qkv, &bar_load, cb * 64, q_row0);
#pragma unroll
    for (int i = 0; i < LK / 128; ++i) tma_load_2d(sK + i * 16384, &map_qkv, &bar_load, p.C + cb * 64, key_row0 + i * 128);
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) tma_load_2d(sVt + kb * 8192, &map_vt, &bar_load, kb * 64, kv_tile * p.C + cb * 64);
  }
  mbar_wait(&bar_load, 0);

  int k_lo = 0, k_hi = LK;
  if (MASKED) {  // several samples share the tile: a query only sees the keys of its own sample
    const int pos = (q_row0 - key_row0) + r;
    k_lo = (pos / p.L) * p.L;
    k_hi = k_lo + p.L;
  }
  const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t idesc_s = make_idesc(LK), idesc_o = make_idesc(p.hd);
  uint32_t mma_phase = 0;
  const int c_begin = half * HALF, c_end = c_begin + HALF;

  for (int h = 0; h < p.heads_per_blk; ++h) {
    // ---------------- S = Q_h K_h^T ----------------
    if (tid == 0) {
      tc_fence_after();
      const uint64_t head_off = (uint64_t)((h * p.hd * 2) >> 4);
      const uint64_t dq = make_smem_desc(smem_u32(sQ)) + head_off;
      const uint64_t dk = make_smem_desc(smem_u32(sK)) + head_off;
      for (int kk = 0; kk < p.hd / 16; ++kk) umma_bf16(tmem, dq + (uint64_t)(2 * kk), dk + (uint64_t)(2 * kk), idesc_s, kk > 0 ? 1u : 0u);
      umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, mma_phase);
    mma_phase ^= 1u;
    tc_fence_after();

    // ---------------- softmax over the row (this thread: columns [c_begin, c_end)) ----------------
    float m = -INFINITY;
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32(t_lane + (uint32_t)c, v);
      tmem_ld_wait();
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float x = __uint_as_float(v[i]);
        if (!MASKED || (c + i >= k_lo && c + i < k_hi)) m4[i & 3] = fmaxf(m4[i & 3], x);
      }
      m = fmaxf(m, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
    }
    s_red[half][r] = m;
    __syncthreads();
    m = fmaxf(s_red[0][r], s_red[1][r]);
    const float mscaled = m * p.scale_log2;
    float sum = 0.f;
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32(t_lane + (uint32_t)c, v);
      tmem_ld_wait();
      uint32_t packed[16];
      float s2[2] = {0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2, -mscaled));
        float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, -mscaled));
        if (MASKED) {
          if (!(c + i >= k_lo && c + i < k_hi)) p0 = 0.f;
          if (!(c + i + 1 >= k_lo && c + i + 1 < k_hi)) p1 = 0.f;
        }
        const __nv_bfloat162 b2 = __floats2bfloat162_rn(p0, p1);
        s2[0] += p0;   // fp32 exponentials: see attn_head.cu
        s2[1] += p1;
        packed[i >> 1] = *reinterpret_cast<const uint32_t*>(&b2);
      }
      sum += s2[0] + s2[1];
      // row r of key block kb = c/64, 16-byte chunks j0..j0+3, 128B swizzle: chunk j lives at (j ^ (r & 7))
      uint8_t* rowp = sP + (c >> 6) * 16384 + r * 128;
      const int j0 = (c & 63) >> 3;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 val = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
        *reinterpret_cast<uint4*>(rowp + (((j0 + j) ^ (r & 7)) << 4)) = val;
      }
    }
    __syncthreads();            // every thread has consumed the partial maxima
    s_red[half][r] = sum;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();

    // ---------------- O = P V_h ----------------
    if (tid == 0) {
      tc_fence_after();
#pragma unroll 1
      for (int kb = 0; kb < KB; ++kb) {
        const uint64_t dp = make_smem_desc(smem_u32(sP + kb * 16384));
        const uint64_t dv = make_smem_desc(smem_u32(sVt + kb * 8192 + h * p.hd * 128));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16(tmem + (uint32_t)LK, dp + (uint64_t)(2 * kk), dv + (uint64_t)(2 * kk), idesc_o, (kb > 0 || kk > 0) ? 1u : 0u);
      }
      umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, mma_phase);
    mma_phase ^= 1u;
    tc_fence_after();

    if (half == 0) {
      const float inv = 1.f / (s_red[0][r] + s_red[1][r]);
      bf16* orow = p.out + (size_t)(q_row0 + r) * p.C + cb * 64 + h * p.hd;
#pragma unroll 1
      for (int c = 0; c < p.hd; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_lane + (uint32_t)(LK + c), v);
        tmem_ld_wait();
        const int n = p.hd - c < 32 ? p.hd - c : 32;
        for (int i = 0; i < n; i += 8) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[i + e]) * inv;
          store8(orow + c + i, f);
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // S, P, O and s_red are reused by the next head
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem);
  }
}

// Second generation of the same tile: P never goes through shared memory.  The softmax threads write P (bf16 pairs) back
// into tensor memory, over the S columns they have already consumed, and O = P V is issued in the TS form of
// tcgen05.mma (A operand from TMEM).  In the SS form every K=16 step of the P V product costs the ~128-cycle shared-memory
// read of a 128-row A operand although N = hd is only 16..64 (16 steps per head at LK = 256: ~2 000 cycles, as long as
// the softmax itself); from TMEM the step costs ~N/2 cycles.  Without the 64 KB P buffer a CTA needs 48-80 KB of shared
// memory and 256 TMEM columns, so two CTAs share an SM and one's softmax runs under the other's MMA / barrier latencies.
//   TMEM columns, LK = 256: S [0,256); P keys 0..127 -> [0,64), keys 128..255 -> [128,192); O_h -> [64, 64+hd)
//                 LK = 128: S [0,128); P keys 0..63  -> [0,32), keys 64..127  -> [64,96);   O_h -> [128, 128+hd)
//   (a thread owns one row and one half of its key columns: its P columns lie inside its own, already-read S columns)
template <int LK, bool MASKED>
__global__ void __launch_bounds__(256, 2)
sdpa_tc2_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_vt, const SdpaParams p) {
  constexpr int KB = LK / 64;
  constexpr int SQ = 16384, SK = LK * 128, SVT = KB * 8192;
  constexpr int TMEM_COLS = 256;
  constexpr int HALF = LK / 2;
  constexpr uint32_t O_COL = (LK == 256) ? 64 : 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + SQ;
  uint8_t* sVt = sK + SK;
  __shared__ __align__(8) uint64_t bar_load;
  __shared__ __align__(8) uint64_t bar_mma;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_max[2][128];   // per-row partial maximum of the two column halves
  __shared__ float s_sum[2][128];   // per-row partial sum

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = warp >> 2;
  const int r = (warp & 3) * 32 + lane;   // query row inside the tile == TMEM lane
  const int q_row0 = blockIdx.x * 128;
  const int cb = blockIdx.y;
  const int kv_tile = q_row0 / LK;
  const int key_row0 = kv_tile * LK;

  if (tid == 0) {
    tma_prefetch_desc(&map_qkv);
    tma_prefetch_desc(&map_vt);
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<TMEM_COLS>(&tmem_base_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_smem;
  pdl_wait();
  pdl_trigger();

  if (tid == 0) {
    mbar_expect_tx(&bar_load, SQ + SK + SVT);
    tma_load_2d(sQ, &map_qkv, &bar_load, cb * 64, q_row0);
#pragma unroll
    for (int i = 0; i < LK / 128; ++i) tma_load_2d(sK + i * 16384, &map_qkv, &bar_load, p.C + cb * 64, key_row0 + i * 128);
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) tma_load_2d(sVt + kb * 8192, &map_vt, &bar_load, kb * 64, kv_tile * p.C + cb * 64);
  }
  mbar_wait(&bar_load, 0);

  int k_lo = 0, k_hi = LK;
  int wk_lo = 0, wk_hi = LK;   // union of the key ranges of this warp's 32 rows (warp-uniform): chunks outside it hold no key of
  if (MASKED) {                // any of them -- no TMEM read, no exp, P = 0 (at L = 16 that is 7 of every 8 columns)
    const int pos = (q_row0 - key_row0) + r;
    k_lo = (pos / p.L) * p.L;
    k_hi = k_lo + p.L;
    const int pos_w = (q_row0 - key_row0) + (warp & 3) * 32;
    wk_lo = (pos_w / p.L) * p.L;
    wk_hi = ((pos_w + 31) / p.L + 1) * p.L;
  }
  const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t idesc_s = make_idesc(LK), idesc_o = make_idesc(p.hd);
  uint32_t mma_phase = 0;
  const int c_begin = half * HALF, c_end = c_begin + HALF;
  const uint32_t pair_bar = 1 + (warp & 3);   // named barrier shared by the two warps that own the same 32 rows

  for (int h = 0; h < p.heads_per_blk; ++h) {
    // ---------------- S = Q_h K_h^T ----------------
    if (tid == 0) {
      tc_fence_after();
      const uint64_t head_off = (uint64_t)((h * p.hd * 2) >> 4);
      const uint64_t dq = make_smem_desc(smem_u32(sQ)) + head_off;
      const uint64_t dk = make_smem_desc(smem_u32(sK)) + head_off;
      for (int kk = 0; kk < p.hd / 16; ++kk) umma_bf16(tmem, dq + (uint64_t)(2 * kk), dk + (uint64_t)(2 * kk), idesc_s, kk > 0 ? 1u : 0u);
      umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, mma_phase);
    mma_phase ^= 1u;
    tc_fence_after();

    // ---------------- pass 1: row maximum over this thread's columns ----------------
    float m = -INFINITY;
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 32) {
      if (MASKED && (c + 32 <= wk_lo || c >= wk_hi)) continue;
      uint32_t v[32];
      tmem_ld_32x32(t_lane + (uint32_t)c, v);
      tmem_ld_wait();
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float x = __uint_as_float(v[i]);
        if (!MASKED || (c + i >= k_lo && c + i < k_hi)) m4[i & 3] = fmaxf(m4[i & 3], x);
      }
      m = fmaxf(m, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
    }
    s_max[half][r] = m;
    asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
    m = fmaxf(s_max[0][r], s_max[1][r]);
    // a fully masked half (MASKED, L < HALF) leaves -inf in one slot; the row maximum itself is always finite
    const float mscaled = m * p.scale_log2;

    // ---------------- pass 2: P = exp2(S - max) -> bf16 pairs -> TMEM (behind the read pointer) ----------------
    float sum = 0.f;
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 32) {
      uint32_t packed[16];
      if (MASKED && (c + 32 <= wk_lo || c >= wk_hi)) {
#pragma unroll
        for (int i = 0; i < 16; ++i) packed[i] = 0u;
        tmem_st_32x16(t_lane + (uint32_t)(c_begin + ((c - c_begin) >> 1)), packed);
        continue;
      }
      uint32_t v[32];
      tmem_ld_32x32(t_lane + (uint32_t)c, v);
      tmem_ld_wait();
      float s2[2] = {0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2, -mscaled));
        float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, -mscaled));
        if (MASKED) {
          if (!(c + i >= k_lo && c + i < k_hi)) p0 = 0.f;
          if (!(c + i + 1 >= k_lo && c + i + 1 < k_hi)) p1 = 0.f;
        }
        const __nv_bfloat162 b2 = __floats2bfloat162_rn(p0, p1);   // .x (low half) = the even key
        s2[0] += p0;   // fp32 exponentials: see attn_head.cu
        s2[1] += p1;
        packed[i >> 1] = *reinterpret_cast<const uint32_t*>(&b2);
      }
      sum += s2[0] + s2[1];
      tmem_st_32x16(t_lane + (uint32_t)(c_begin + ((c - c_begin) >> 1)), packed);
    }
    tmem_st_wait();
    s_sum[half][r] = sum;
    tc_fence_before();
    __syncthreads();   // every row of P is in tensor memory; S is dead

    // ---------------- O = P V_h (A = P from TMEM) ----------------
    if (tid == 0) {
      tc_fence_after();
#pragma unroll 1
      for (int k16 = 0; k16 < LK / 16; ++k16) {
        const int kb = k16 >> 2, kk = k16 & 3;
        const uint32_t a_col = (uint32_t)((k16 >= LK / 32 ? HALF : 0) + 8 * (k16 % (LK / 32)));
        const uint64_t dv = make_smem_desc(smem_u32(sVt + kb * 8192 + h * p.hd * 128)) + (uint64_t)(2 * kk);
        umma_bf16_ts(tmem + O_COL, tmem + a_col, dv, idesc_o, k16 > 0 ? 1u : 0u);
      }
      umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, mma_phase);
    mma_phase ^= 1u;
    tc_fence_after();

    if (half == 0) {
      const float inv = 1.f / (s_sum[0][r] + s_sum[1][r]);
      bf16* orow = p.out + (size_t)(q_row0 + r) * p.C + cb * 64 + h * p.hd;
#pragma unroll 1
      for (int c = 0; c < p.hd; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_lane + O_COL + (uint32_t)c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          if (c + i < p.hd) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[i + e]) * inv;
            store8(orow + c + i, f);
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // S / P / O columns and the exchange arrays are reused by the next head
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem);
  }
}

// Long sequences (L = 512 or 1024 tokens per sample: the 64x8 / 128x8 maps of the 2x / 4x prediction horizons): the
// keys of a sample are walked in blocks of 256.  Two passes, so that no accumulator ever has to be rescaled:
//   pass A  per key block, per head: S = Q K^T -> running row maximum
//   pass B  per key block, per head: S again -> P = exp2(S - max) -> smem -> O_h += P V_h   (O_h: own TMEM columns)
// K and V^T blocks are loaded once per pass and shared by the 64/hd heads of the channel block.
__global__ void __launch_bounds__(256)
sdpa_tc_long_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_vt, const SdpaParams p) {
  constexpr int LK = 256, KB = 4, HALF = 128;
  constexpr int SQ = 16384, SK = LK * 128, SVT = KB * 8192;
  constexpr int TMEM_COLS = 512;  // S: 256 columns, O: 64 columns (all heads of the channel block)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + SQ;
  uint8_t* sVt = sK + SK;
  uint8_t* sP = sVt + SVT;
  __shared__ __align__(8) uint64_t bar_load;
  __shared__ __align__(8) uint64_t bar_mma;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_red[2][128];
  __shared__ float s_sum[2][4][128];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = warp >> 2;
  const int r = (warp & 3) * 32 + lane;
  const int q_row0 = blockIdx.x * 128;
  const int cb = blockIdx.y;
  const int sample = q_row0 / p.L;
  const int key_row0 = sample * p.L;
  const int nkb = p.L / LK;
  const int nh = p.heads_per_blk;

  if (tid == 0) {
    tma_prefetch_desc(&map_qkv);
    tma_prefetch_desc(&map_vt);
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<TMEM_COLS>(&tmem_base_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_smem;
  pdl_wait();
  pdl_trigger();

  const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t idesc_s = make_idesc(LK), idesc_o = make_idesc(p.hd);
  uint32_t mma_phase = 0, load_phase = 0;
  const int c_begin = half * HALF, c_end = c_begin + HALF;
  float mrow[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  float srow[4] = {0.f, 0.f, 0.f, 0.f};

  auto issue_s = [&](int h) {  // S = Q_h K_blk,h^T
    tc_fence_after();
    const uint64_t head_off = (uint64_t)((h * p.hd * 2) >> 4);
    const uint64_t dq = make_smem_desc(smem_u32(sQ)) + head_off;
    const uint64_t dk = make_smem_desc(smem_u32(sK)) + head_off;
    for (int kk = 0; kk < p.hd / 16; ++kk) umma_bf16(tmem, dq + (uint64_t)(2 * kk), dk + (uint64_t)(2 * kk), idesc_s, kk > 0 ? 1u : 0u);
    umma_commit(&bar_mma);
  };

  // ---------------- pass A: row maxima ----------------
  for (int jb = 0; jb < nkb; ++jb) {
    if (tid == 0) {
      mbar_expect_tx(&bar_load, (jb == 0 ? SQ : 0) + SK);
      if (jb == 0) tma_load_2d(sQ, &map_qkv, &bar_load, cb * 64, q_row0);
      tma_load_2d(sK, &map_qkv, &bar_load, p.C + cb * 64, key_row0 + jb * LK);
      tma_load_2d(sK + 16384, &map_qkv, &bar_load, p.C + cb * 64, key_row0 + jb * LK + 128);
    }
    mbar_wait(&bar_load, load_phase);
    load_phase ^= 1u;
    for (int h = 0; h < nh; ++h) {
      if (tid == 0) issue_s(h);
      mbar_wait(&bar_mma, mma_phase);
      mma_phase ^= 1u;
      tc_fence_after();
      float m = -INFINITY;
#pragma unroll 1
      for (int c = c_begin; c < c_end; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_lane + (uint32_t)c, v);
        tmem_ld_wait();
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 32; i += 2) m4[(i >> 1) & 3] = fmax3(m4[(i >> 1) & 3], __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
        m = fmaxf(m, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
      }
      mrow[h] = fmaxf(mrow[h], m);
      tc_fence_before();
      __syncthreads();  // S and sK are reused by the next head / key block
    }
  }
  // combine the two column halves of every row
  for (int h = 0; h < nh; ++h) {
    s_red[half][r] = mrow[h];
    __syncthreads();
    mrow[h] = fmaxf(s_red[0][r], s_red[1][r]) * p.scale_log2;  // pre-scaled maximum
    __syncthreads();
  }

  // ---------------- pass B: P = exp2(S - max), O_h += P V_h ----------------
  for (int jb = 0; jb < nkb; ++jb) {
    if (tid == 0) {
      mbar_expect_tx(&bar_load, SK + SVT);
      tma_load_2d(sK, &map_qkv, &bar_load, p.C + cb * 64, key_row0 + jb * LK);
      tma_load_2d(sK + 16384, &map_qkv, &bar_load, p.C + cb * 64, key_row0 + jb * LK + 128);
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(sVt + kb * 8192, &map_vt, &bar_load, jb * LK + kb * 64, sample * p.C + cb * 64);
    }
    mbar_wait(&bar_load, load_phase);
    load_phase ^= 1u;
    for (int h = 0; h < nh; ++h) {
      if (tid == 0) issue_s(h);
      mbar_wait(&bar_mma, mma_phase);
      mma_phase ^= 1u;
      tc_fence_after();
      float sum = 0.f;
#pragma unroll 1
      for (int c = c_begin; c < c_end; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_lane + (uint32_t)c, v);
        tmem_ld_wait();
        uint32_t packed[16];
        float s2[2] = {0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2, -mrow[h]));
          const float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, -mrow[h]));
          const __nv_bfloat162 b2 = __floats2bfloat162_rn(p0, p1);
          s2[0] += p0;   // fp32 exponentials: see attn_head.cu (the bf16-rounded P differs by 2^-9 per element, unbiased)
          s2[1] += p1;
          packed[i >> 1] = *reinterpret_cast<const uint32_t*>(&b2);
        }
        sum += s2[0] + s2[1];
        uint8_t* rowp = sP + (c >> 6) * 16384 + r * 128;
        const int j0 = (c & 63) >> 3;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 val = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
          *reinterpret_cast<uint4*>(rowp + (((j0 + j) ^ (r & 7)) << 4)) = val;
        }
      }
      srow[h] += sum;
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
#pragma unroll 1
        for (int kb = 0; kb < KB; ++kb) {
          const uint64_t dp = make_smem_desc(smem_u32(sP + kb * 16384));
          const uint64_t dv = make_smem_desc(smem_u32(sVt + kb * 8192 + h * p.hd * 128));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16(tmem + (uint32_t)(LK + h * p.hd), dp + (uint64_t)(2 * kk), dv + (uint64_t)(2 * kk), idesc_o,
                      (jb > 0 || kb > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&bar_mma);
      }
      mbar_wait(&bar_mma, mma_phase);   // P, S and (after the last head) sK / sVt may be overwritten
      mma_phase ^= 1u;
      tc_fence_after();
    }
    tc_fence_before();
    __syncthreads();
  }

  // ---------------- epilogue: O_h / rowsum ----------------
  for (int h = 0; h < nh; ++h) s_sum[half][h][r] = srow[h];
  __syncthreads();
  if (half == 0) {
    for (int h = 0; h < nh; ++h) {
      const float inv = 1.f / (s_sum[0][h][r] + s_sum[1][h][r]);
      bf16* orow = p.out + (size_t)(q_row0 + r) * p.C + cb * 64 + h * p.hd;
#pragma unroll 1
      for (int c = 0; c < p.hd; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_lane + (uint32_t)(LK + h * p.hd + c), v);
        tmem_ld_wait();
        const int n = p.hd - c < 32 ? p.hd - c : 32;
        for (int i = 0; i < n; i += 8) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[i + e]) * inv;
          store8(orow + c + i, f);
        }
      }
    }
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

}  // namespace

struct SdpaTc {
  CUtensorMap map_qkv, map_vt;
  SdpaParams p;
  int LK;
};

bool sdpa_tc_supported(int L, int C, int heads) {
  if (heads <= 0 || C % heads || C % 64) return false;
  const int hd = C / heads;
  if (hd != 16 && hd != 32 && hd != 64) return false;
  return (L <= 128 && 128 % L == 0) || L == 256 || L == 512 || L == 1024;
}
int sdpa_tc_keys_per_tile(int L) { return L <= 128 ? 128 : L; }

SdpaTc* sdpa_tc_create(const bf16* qkv, const bf16* vt, int C, int L, int heads, long long Mcap) {
  EncodeTiledFn enc = get_encode();
  if (!enc || !sdpa_tc_supported(L, C, heads)) return nullptr;
  SdpaTc* g = new SdpaTc();
  memset(g, 0, sizeof(*g));
  g->LK = sdpa_tc_keys_per_tile(L);
  const int hd = C / heads;
  g->p.C = C; g->p.L = L; g->p.hd = hd; g->p.heads_per_blk = 64 / hd;
  g->p.scale_log2 = 1.4426950408889634f / sqrtf((float)hd);
  {
    cuuint64_t dims[2] = {(cuuint64_t)3 * C, (cuuint64_t)Mcap};
    cuuint64_t strides[1] = {(cuuint64_t)3 * C * 2};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&g->map_qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)qkv, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      delete g;
      return nullptr;
    }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)g->LK, (cuuint64_t)(Mcap / g->LK) * C};
    cuuint64_t strides[1] = {(cuuint64_t)g->LK * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&g->map_vt, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)vt, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      delete g;
      return nullptr;
    }
  }
  return g;
}
void sdpa_tc_destroy(SdpaTc* g) { delete g; }

template <int LK, bool MASKED>
static void sdpa_launch_cfg(const SdpaTc* g, const SdpaParams& p, dim3 grid, cudaStream_t s) {
  constexpr int smem = 16384 + LK * 128 + (LK / 64) * 8192 + (LK / 64) * 16384 + 1024;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(sdpa_tc_kernel<LK, MASKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
  launch_pdl(sdpa_tc_kernel<LK, MASKED>, grid, dim3(256), smem, s, g->map_qkv, g->map_vt, p);
}

template <int LK, bool MASKED>
static void sdpa2_launch_cfg(const SdpaTc* g, const SdpaParams& p, dim3 grid, cudaStream_t s) {
  constexpr int smem = 16384 + LK * 128 + (LK / 64) * 8192 + 1024;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(sdpa_tc2_kernel<LK, MASKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
  launch_pdl(sdpa_tc2_kernel<LK, MASKED>, grid, dim3(256), smem, s, g->map_qkv, g->map_vt, p);
}

void sdpa_tc_launch(const SdpaTc* g, bf16* out, long long M, cudaStream_t s) {
  kernels_count_launch();
  SdpaParams p = g->p;
  p.out = out;
  dim3 grid((unsigned)(M / 128), (unsigned)(p.C / 64));
  static int v1 = -1;  // SPDM_SDPA_V1=1: the first-generation kernel (P through shared memory), kept for A/B runs
  if (v1 < 0) { const char* e = getenv("SPDM_SDPA_V1"); v1 = e ? atoi(e) : 0; }
  if (p.L <= 256 && !v1) {
    if (g->LK == 256) sdpa2_launch_cfg<256, false>(g, p, grid, s);
    else if (p.L == 128) sdpa2_launch_cfg<128, false>(g, p, grid, s);
    else sdpa2_launch_cfg<128, true>(g, p, grid, s);
    return;
  }
  if (p.L > 256) {
    constexpr int smem = 16384 + 256 * 128 + 4 * 8192 + 4 * 16384 + 1024;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(sdpa_tc_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
    launch_pdl(sdpa_tc_long_kernel, grid, dim3(256), smem, s, g->map_qkv, g->map_vt, p);
  } else if (g->LK == 256) sdpa_launch_cfg<256, false>(g, p, grid, s);
  else if (p.L == 128) sdpa_launch_cfg<128, false>(g, p, grid, s);
  else sdpa_launch_cfg<128, true>(g, p, grid, s);
}
