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
//             accumulators = 2 x BLOCK_N TMEM columns (double-buffered).
//   roles   = warp 0: TMA producer, warp 1: MMA issuer, warps 2..5: epilogue (tcgen05.ld -> bias /
//             GELU / residual / GroupNorm partial sums -> bf16 global stores).
//   The kernel is persistent: one CTA per SM walks the (m_tile, n_tile) list; the smem ring keeps
//   streaming across tile boundaries and the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "tc_ptx.cuh"

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
  int n_tiles, m_tiles, total_tiles;
  int P;                  // stats partial slots per sample
  int ld_out, ld_res;
  int flags;
  int dbg;                // microbenchmark switches (tc_set_debug): 1 = no B loads, 2 = no A loads, 4 = no epilogue stores, 8 = centre tap only
  bf16* out;
  float* stats;
  const float* bias;
  const bf16* resid;
  int ksplit;             // > 1: the K loop of every tile is cut into ksplit pieces run by different CTAs (split-K)
  float* partial;         // ksplit > 1: fp32 partial tiles [ksplit][m_tiles*128][Cout], summed by apply_partial_kernel
  int cl_ks;              // conv_tc_cluster_kernel: K slices per output tile inside a cluster (cluster size = n_tiles * cl_ks)
  ApplyArgs ap;           // EPI_APPLY: GroupNorm apply (+GELU, +temb, +FiLM) fused behind the accumulator (raw/out/stats unused)
  bf16* vt;               // EPI_VT: columns >= vt_c0 go, transposed, to vt[row / vt_lk][col - vt_c0][row % vt_lk]
  int vt_c0, vt_C, vt_lk;
  // pair fold (conv_tc_swap_kernel, 64 real output channels): W is the number of PAIRS per image row, Cout = 128 = (wo, co), the
  // K loop walks 3 dy x 4 blocks (kernels.cu::pack_conv_pfold_bf16_kernel); fold_c1 = channel offset of a pair's second pixel
  // in the pair-row view of the input (= ld_in), fold_ldo = real leading dimension of the output (ld_out is the pair-row stride)
  int fold, fold_c1, fold_ldo;
  // halo mode (conv_tc_swap_kernel, W == 8, 32 image rows of one sample per tile): the pixel operand of a (k block, dx) pair is loaded
  // ONCE as a (32 + 2)-row box and serves the three dy taps through 1024-byte descriptor offsets
  int halo;
  int pix256;             // conv_tc_swap_kernel: map_a has a 256-row box (one pixel load per k-step instead of two)
};

// ---------------------------------------------------------------------------------------------
// The kernel.  grid = min(total_tiles, #SM); dynamic smem = STAGES*(A+B) + 1024 (alignment slack).
// ---------------------------------------------------------------------------------------------
struct TileCoord { int m_tile, n_tile, b0, h0, n0, ks; };

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }  // the 4 epilogue warps

// GELU through the hardware tanh (kernels.cu::gelu_tanh_fast): the fused GroupNorm-apply epilogue is issue-bound on 4 warps
__device__ __forceinline__ float gelu_fast_tc(float y) {
  const float u = 0.7978845608028654f * fmaf(0.044715f * y * y, y, y);
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(u));
  return fmaf(0.5f * y, th, 0.5f * y);
}

// row of the time-embedding table used by this launch (or null)
__device__ __forceinline__ const float* temb_row(const ApplyArgs& ap, int b) {
  if (ap.temb_mode == TEMB_NONE) return nullptr;
  int trow = 0;
  if (ap.temb_mode == TEMB_PER_SAMPLE) trow = b;
  else if (ap.temb_mode == TEMB_STEP) trow = *ap.step_ptr + ap.step_off;
  return ap.temb + (size_t)trow * SPDM_TEMB_WIDTH + ap.temb_off;
}

__device__ __forceinline__ const float* temb_row_off(const ApplyArgs& ap, int b, int step_off) {
  if (ap.temb_mode == TEMB_NONE) return nullptr;
  int trow = 0;
  if (ap.temb_mode == TEMB_PER_SAMPLE) trow = b;
  else if (ap.temb_mode == TEMB_STEP) trow = *ap.step_ptr + step_off;
  return ap.temb + (size_t)trow * SPDM_TEMB_WIDTH + ap.temb_off;
}

__device__ __forceinline__ TileCoord decode_tile(const TcParams& p, int tile, int block_n) {
  TileCoord t;
  t.ks = 0;
  if (p.ksplit > 1) { t.ks = tile % p.ksplit; tile /= p.ksplit; }
  t.n_tile = tile % p.n_tiles;  // n fastest: CTAs running side by side share the same A tile in L2
  t.m_tile = tile / p.n_tiles;
  const int tiles_per_sample = p.H / p.Hb;  // > 1 only when Bt == 1
  if (tiles_per_sample > 1) { t.b0 = t.m_tile / tiles_per_sample; t.h0 = (t.m_tile - t.b0 * tiles_per_sample) * p.Hb; }
  else { t.b0 = t.m_tile * p.Bt; t.h0 = 0; }
  t.n0 = t.n_tile * block_n;
  return t;
}

// PAIR: the CTA-pair form (cta_group::2, tc_ptx.cuh) for persistent launches of the wide layers: a cluster of two CTAs takes two
// neighbouring M tiles of the same N tile; each CTA stages its own 128 pixel rows and HALF of the 256 weight rows per k-step
// (32 KB instead of 48 KB: six pipeline stages instead of four in the same shared memory, a third less operand traffic), the
// leader issues one M = 256 MMA for both.  Everything behind the accumulator (epilogue warps, statistics, fused apply) is per CTA
// and unchanged.  map_b must be the 128-row-box weight map.  ksplit == 1, m_tiles even.
template <int BLOCK_N, int STAGES, bool PAIR = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcParams p) {
  constexpr int B_STAGE_BYTES = (PAIR ? BLOCK_N / 2 : BLOCK_N) * BLOCK_K * 2;
  static_assert(!PAIR || BLOCK_N == 256, "the pair form is built for 256-wide N tiles");
  constexpr int TMEM_COLS = 2 * BLOCK_N <= 128 ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512);  // power of two
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_part[2][4][2];   // EPI_APPLY: per-warp (sum, sumsq) of the tile's samples, double-buffered like TMEM

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  // tile walk: a CTA takes tiles blockIdx.x, + gridDim.x, ...; a pair takes "super tiles" (two neighbouring M tiles of one N tile)
  const int t_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int t_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int t_total = PAIR ? p.total_tiles / 2 : p.total_tiles;
  auto real_tile = [&](int st) {
    if (!PAIR) return st;
    const int j = st / p.n_tiles, n = st - j * p.n_tiles;
    return (2 * j + (int)crank) * p.n_tiles + n;
  };

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
    mbar_init(&tmem_full_bar[0], 1); mbar_init(&tmem_full_bar[1], 1);
    mbar_init(&tmem_empty_bar[0], PAIR ? 8 : 4); mbar_init(&tmem_empty_bar[1], PAIR ? 8 : 4);  // one arrival per epilogue warp (of both CTAs)
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) tmem_alloc_pair<TMEM_COLS>(&tmem_base_smem);
    else tmem_alloc<TMEM_COLS>(&tmem_base_smem);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();   // the peer's barriers exist before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  // Programmatic dependent launch: everything above overlapped with the previous kernel's tail.  The weights do not depend
  // on the previous kernel either, so the producer also puts the B tiles of its first STAGES k-steps in flight BEFORE the
  // dependency wait; activations (A tiles) and every global write come after it.
  if (!(warp == 0 && lane == 0)) { pdl_wait(); pdl_trigger(); }

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      auto tap_of = [&](int it, int& dy, int& dx, int& tap, int& kb) {
        const int tap_i = it / p.kb_per_tap;
        kb = it - tap_i * p.kb_per_tap;
        dy = 0; dx = 0; tap = 0;
        if (p.taps == 9) {
          const int ty = tap_i / ntx, tx = tap_i - ty * ntx;
          dy = skip_dy ? 0 : ty - 1;
          dx = skip_dx ? 0 : tx - 1;
          tap = (dy + 1) * 3 + (dx + 1);
        }
      };
      // pair: the leader's barrier counts the bytes of both CTAs (its own expect_tx is the only arrival)
      const uint32_t stage_tx = (((p.dbg & 2) ? 0 : A_STAGE_BYTES) + ((p.dbg & 1) ? 0 : B_STAGE_BYTES)) * (PAIR ? 2u : 1u);
      const int b_row_off = PAIR ? (int)crank * (BLOCK_N / 2) : 0;   // this CTA's half of the weight rows
      auto load_b = [&](int s, int tap, int kb, int n0) {
        if (p.dbg & 1) return;
        if constexpr (PAIR) tma_load_2d_pair(smem_b + s * B_STAGE_BYTES, &map_b, &full_bar[s], tap * p.Cin + kb * BLOCK_K, n0 + b_row_off);
        else tma_load_2d(smem_b + s * B_STAGE_BYTES, &map_b, &full_bar[s], tap * p.Cin + kb * BLOCK_K, n0);
      };
      uint32_t n_pre = 0;   // k-steps of the first tile whose expect_tx + weight load were issued before the wait
      if (!(p.dbg & 16) && t_first < t_total) {
        const TileCoord t = decode_tile(p, real_tile(t_first), BLOCK_N);
        const int it_begin = p.ksplit > 1 ? t.ks * k_iters / p.ksplit : 0;
        const int it_end = p.ksplit > 1 ? (t.ks + 1) * k_iters / p.ksplit : k_iters;
        for (int it = it_begin; it < it_end && n_pre < (uint32_t)STAGES; ++it, ++n_pre) {
          int dy, dx, tap, kb;
          tap_of(it, dy, dx, tap, kb);
          if (!PAIR || crank == 0) mbar_expect_tx(&full_bar[n_pre], stage_tx);
          load_b((int)n_pre, tap, kb, t.n0);
        }
      }
      pdl_wait();
      pdl_trigger();
      if (!(p.dbg & 16)) {
        uint32_t kit = 0;
        long long prod_wait = 0, prod_etx = 0, prod_b = 0, prod_a = 0;
        const long long prod_t0 = clock64();
        for (int st = t_first; st < t_total; st += t_step) {
          const TileCoord t = decode_tile(p, real_tile(st), BLOCK_N);
          const int it_begin = p.ksplit > 1 ? t.ks * k_iters / p.ksplit : 0;
          const int it_end = p.ksplit > 1 ? (t.ks + 1) * k_iters / p.ksplit : k_iters;
          for (int it = it_begin; it < it_end; ++it, ++kit) {
            const int s = kit % STAGES;
            const uint32_t ph = (kit / STAGES) & 1u;
            int dy, dx, tap, kb;
            tap_of(it, dy, dx, tap, kb);
            if (p.dbg & 8) { dx = 0; dy = 0; }
            if (kit >= n_pre) {
              const long long tw0 = (p.dbg & 4096) ? clock64() : 0;
              mbar_wait(&empty_bar[s], ph ^ 1u);   // (pair: the leader's MMA commit arrives on this barrier in both CTAs)
              if (p.dbg & 4096) prod_wait += clock64() - tw0;
              const long long te0 = (p.dbg & 4096) ? clock64() : 0;
              if (!PAIR || crank == 0) mbar_expect_tx(&full_bar[s], stage_tx);
              const long long te1 = (p.dbg & 4096) ? clock64() : 0;
              load_b(s, tap, kb, t.n0);
              if (p.dbg & 4096) { prod_etx += te1 - te0; prod_b += clock64() - te1; }
            }
            if (!(p.dbg & 2)) {
              const long long ta0 = (p.dbg & 4096) ? clock64() : 0;
              if constexpr (PAIR) tma_load_4d_pair(smem_a + s * A_STAGE_BYTES, &map_a, &full_bar[s], kb * BLOCK_K, dx, t.h0 + dy, t.b0);
              else tma_load_4d(smem_a + s * A_STAGE_BYTES, &map_a, &full_bar[s], kb * BLOCK_K, dx, t.h0 + dy, t.b0);
              if (p.dbg & 4096) prod_a += clock64() - ta0;
            }
          }
        }
        if ((p.dbg & 4096) && blockIdx.x == 0)
          printf("spdm conv timing (producer, block 0): %u k-steps, %lld cycles, %lld waiting for a free stage, %lld in expect_tx, %lld issuing weight loads, %lld issuing pixel loads\n",
                 kit, clock64() - prod_t0, prod_wait, prod_etx, prod_b, prod_a);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0 && (!PAIR || crank == 0)) {   // pair: the leader issues for both CTAs
      constexpr uint32_t idesc = PAIR ? make_idesc_m(256, BLOCK_N) : make_idesc(BLOCK_N);
      uint32_t kit = 0;
      int lt = 0;
      long long mma_wait = 0, acc_wait = 0;
      const long long mma_t0 = clock64();
      for (int st = t_first; st < t_total; st += t_step, ++lt) {
        const int tile = real_tile(st);
        const int acc = lt & 1;
        const uint32_t aph = (uint32_t)(lt >> 1) & 1u;
        const long long ta0 = (p.dbg & 4096) ? clock64() : 0;
        mbar_wait(&tmem_empty_bar[acc], aph ^ 1u);  // epilogue has drained this accumulator (pair: in both CTAs)
        if (p.dbg & 4096) acc_wait += clock64() - ta0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        int it_begin = 0, it_end = k_iters;
        if (p.ksplit > 1) { const int ks = tile % p.ksplit; it_begin = ks * k_iters / p.ksplit; it_end = (ks + 1) * k_iters / p.ksplit; }
        for (int it = it_begin; it < it_end; ++it, ++kit) {
          const int s = kit % STAGES;
          const uint32_t ph = (kit / STAGES) & 1u;
          const long long tw0 = (p.dbg & 4096) ? clock64() : 0;
          if (!(p.dbg & 16)) mbar_wait(&full_bar[s], ph);
          if (p.dbg & 4096) mma_wait += clock64() - tw0;
          tc_fence_after();
          const uint64_t da = make_smem_desc(smem_u32(smem_a + s * A_STAGE_BYTES));
          const uint64_t db = make_smem_desc(smem_u32(smem_b + s * B_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
            if constexpr (PAIR) umma_bf16_pair(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (it > it_begin || k > 0) ? 1u : 0u);
            else umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (it > it_begin || k > 0) ? 1u : 0u);
          }
          if (!(p.dbg & 16)) {   // frees the smem slot when these MMAs retire
            if constexpr (PAIR) umma_commit_pair(&empty_bar[s]);
            else umma_commit(&empty_bar[s]);
          }
        }
        if constexpr (PAIR) umma_commit_pair(&tmem_full_bar[acc]);  // accumulator complete
        else umma_commit(&tmem_full_bar[acc]);
      }
      if ((p.dbg & 4096) && blockIdx.x == 0)
        printf("spdm conv timing (MMA issuer, block 0): %u k-steps, %lld cycles, %lld waiting for operands, %lld waiting for a free accumulator\n", kit,
               clock64() - mma_t0, mma_wait, acc_wait);
    }
  } else {
    // ================= epilogue: warps 2..5 -> TMEM lane quarters (warp % 4) =================
    const int q = warp & 3;
    const int r_t = q * 32 + lane;                 // tile row == TMEM lane
    const int rps = p.Hb * p.W;                    // rows per sample inside the tile
    const int tiles_per_sample = p.H / p.Hb;
    // hand the accumulator back to the MMA issuer (pair: the leader's barrier collects the warps of both CTAs)
    auto release_acc = [&](int acc) {
      if constexpr (PAIR) mbar_arrive_cluster(&tmem_empty_bar[acc], 0u);
      else mbar_arrive(&tmem_empty_bar[acc]);
    };
    int lt = 0;
    for (int st = t_first; st < t_total; st += t_step, ++lt) {
      const int tile = real_tile(st);
      const TileCoord t = decode_tile(p, tile, BLOCK_N);
      const int acc = lt & 1;
      const uint32_t aph = (uint32_t)(lt >> 1) & 1u;
      const long long row0 = ((long long)t.b0 * p.H + t.h0) * p.W;  // tile rows are contiguous in the [M, C] map
      const long long row = row0 + r_t;
      bf16* orow = p.out + row * p.ld_out + t.n0;
      const bf16* rrow = (p.flags & (EPI_RESID | EPI_MASK)) ? p.resid + row * p.ld_res + t.n0 : nullptr;
      float rs = 0.f, rq = 0.f;
      mbar_wait(&tmem_full_bar[acc], aph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
      if (p.dbg & 32) {  // microbenchmark: hand the accumulator straight back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release_acc(acc);
        continue;
      }
      if (p.ksplit > 1) {  // split-K: this CTA owns one slice of the K loop; fp32 partial tile, reduced by apply_partial_kernel
        float* prow = p.partial + ((size_t)t.ks * ((size_t)p.m_tiles * BLOCK_M) + (size_t)row) * p.Cout + t.n0;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(t_addr + (uint32_t)c, v);
          tmem_ld_wait();
          if (c + 32 == BLOCK_N) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) release_acc(acc);
          }
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<uint4*>(prow + c + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
        continue;
      }
      if (p.flags & EPI_APPLY) {
        // The tile holds whole samples and all their channels (host guarantees n_tiles == 1, H == Hb), so GroupNorm(1, C)
        // is tile-local: statistics from the fp32 accumulator, then normalise + GELU / time embedding / FiLM on a second
        // pass over TMEM and store the activated map directly -- the raw conv output never touches HBM.
        float ps = 0.f, pq = 0.f;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(t_addr + (uint32_t)c, v);
          tmem_ld_wait();
          float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 32; ++i) { const float f = __uint_as_float(v[i]); s4[i & 3] += f; q4[i & 3] = fmaf(f, f, q4[i & 3]); }
          ps += (s4[0] + s4[1]) + (s4[2] + s4[3]);
          pq += (q4[0] + q4[1]) + (q4[2] + q4[3]);
        }
        const int span = rps >= 32 ? 32 : rps;
        for (int o = span >> 1; o > 0; o >>= 1) { ps += __shfl_xor_sync(0xffffffffu, ps, o); pq += __shfl_xor_sync(0xffffffffu, pq, o); }
        if (rps >= 32) {  // a sample spans rps/32 warps: combine their partials in a fixed order
          if (lane == 0) { s_part[acc][q][0] = ps; s_part[acc][q][1] = pq; }
          epi_bar_sync();
          const int wps = rps / 32, q0 = (q / wps) * wps;
          ps = 0.f; pq = 0.f;
          for (int j = 0; j < wps; ++j) { ps += s_part[acc][q0 + j][0]; pq += s_part[acc][q0 + j][1]; }
        }
        const float inv_n = 1.0f / ((float)rps * (float)p.Cout);
        const float mean = ps * inv_n;
        const float rstd = rsqrtf(fmaxf(pq * inv_n - mean * mean, 0.f) + p.ap.eps);
        const int b = t.b0 + r_t / rps;
        const float* te = temb_row(p.ap, b);
        const float* fi = p.ap.film ? p.ap.film + (size_t)b * SPDM_FILM_WIDTH + p.ap.film_off : nullptr;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(t_addr + (uint32_t)c, v);
          tmem_ld_wait();
          if (c + 32 == BLOCK_N) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) release_acc(acc);
          }
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.ap.gamma + c + i));
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.ap.beta + c + i));
            f[i] = (__uint_as_float(v[i]) - mean) * rstd * g4.x + b4.x;
            f[i + 1] = (__uint_as_float(v[i + 1]) - mean) * rstd * g4.y + b4.y;
            f[i + 2] = (__uint_as_float(v[i + 2]) - mean) * rstd * g4.z + b4.z;
            f[i + 3] = (__uint_as_float(v[i + 3]) - mean) * rstd * g4.w + b4.w;
          }
          if (p.ap.act == ACT_GELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = gelu_fast_tc(f[i]);
          }
          if (te) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 t4 = __ldg(reinterpret_cast<const float4*>(te + c + i));
              f[i] += t4.x; f[i + 1] += t4.y; f[i + 2] += t4.z; f[i + 3] += t4.w;
            }
          }
          if (fi) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 s4 = __ldg(reinterpret_cast<const float4*>(fi + c + i));
              const float4 o4 = __ldg(reinterpret_cast<const float4*>(fi + p.Cout + c + i));
              f[i] = fmaf(s4.x, f[i], o4.x); f[i + 1] = fmaf(s4.y, f[i + 1], o4.y);
              f[i + 2] = fmaf(s4.z, f[i + 2], o4.z); f[i + 3] = fmaf(s4.w, f[i + 3], o4.w);
            }
          }
#pragma unroll
          for (int i = 0; i < 32; i += 8) store8(orow + c + i, f + i);
        }
        continue;
      }
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_addr + (uint32_t)c, v);
        tmem_ld_wait();
        if (c + 32 == BLOCK_N) {  // accumulator fully read: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) release_acc(acc);
        }
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
        if (p.flags & EPI_BIAS) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + t.n0 + c + i));
            f[i] += b4.x; f[i + 1] += b4.y; f[i + 2] += b4.z; f[i + 3] += b4.w;
          }
        }
        if (p.flags & EPI_GELU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = gelu_exact(f[i]);
        }
        if (p.flags & EPI_RELU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
        }
        if (rrow) {
          const bool mask = (p.flags & EPI_MASK) != 0;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            float t8[8];
            load8(rrow + c + i, t8);
#pragma unroll
            for (int e = 0; e < 8; ++e) f[i + e] = mask ? (t8[e] > 0.f ? f[i + e] : 0.f) : f[i + e] + t8[e];
          }
        }
        if (p.flags & EPI_STATS) {
          float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 32; ++i) { s4[i & 3] += f[i]; q4[i & 3] = fmaf(f[i], f[i], q4[i & 3]); }
          rs += (s4[0] + s4[1]) + (s4[2] + s4[3]);
          rq += (q4[0] + q4[1]) + (q4[2] + q4[3]);
        }
        if ((p.flags & EPI_VT) && t.n0 + c >= p.vt_c0) {
          // V^T for the tcgen05 attention core: consecutive lanes are consecutive tokens -> 64-byte segments per column
          const long long vtile = row / p.vt_lk;
          const int pos = (int)(row - vtile * p.vt_lk);
          bf16* dst = p.vt + ((size_t)vtile * p.vt_C + (t.n0 + c - p.vt_c0)) * p.vt_lk + pos;
#pragma unroll
          for (int i = 0; i < 32; ++i) dst[(size_t)i * p.vt_lk] = __float2bfloat16_rn(f[i]);
        } else if (!(p.dbg & 4)) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) store8(orow + c + i, f + i);
        }
      }
      if (p.flags & EPI_STATS) {
        // reduce over the lanes that belong to the same sample, one deterministic slot per writer
        const int span = rps >= 32 ? 32 : rps;  // rps in {128, 64, 32, 16, 8, 4, ...}: power of two
        for (int o = span >> 1; o > 0; o >>= 1) { rs += __shfl_xor_sync(0xffffffffu, rs, o); rq += __shfl_xor_sync(0xffffffffu, rq, o); }
        if ((lane & (span - 1)) == 0) {
          const int b = t.b0 + r_t / rps;
          int slot;
          if (rps >= 32) {
            const int warps_per_sample_tile = (rps >= 128 ? 128 : rps) / 32;
            const int tile_in_sample = tiles_per_sample > 1 ? (t.m_tile % tiles_per_sample) : 0;
            const int warp_in_sample = (r_t % rps) / 32;
            slot = (tile_in_sample * warps_per_sample_tile + warp_in_sample) * p.n_tiles + t.n_tile;
          } else {
            slot = t.n_tile;
          }
          float* dst = p.stats + ((size_t)b * p.P + slot) * 2;
          dst[0] = rs;
          dst[1] = rq;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();   // no CTA of the pair leaves (or frees tensor memory) while the other may still signal it
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair<TMEM_COLS>(tmem_base);
    else tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// Cluster split-K with the GroupNorm apply fused behind it (the 8x2 / 4x1 levels at small batch).
//
// At batch 256 a deep-level conv has 8-32 output tiles and a K loop of up to 72 serial ~0.27 us steps, so it is cut along K
// to occupy the machine.  The first version of that wrote fp32 partial tiles to global memory and a second kernel
// (apply_partial_kernel) summed them and applied GroupNorm: two launches (~10-15 us + ~5.5 us on the critical path of
// the step, measured by leaving them out of the graph: tools/ablate.py).  Here the K slices of one tile -- and, for
// Cout = 512, both of its N tiles -- form ONE thread-block cluster:
//   1. every CTA runs its K slice (same TMA / tcgen05 pipeline as conv_tc_kernel) into its own TMEM accumulator,
//   2. dumps the fp32 accumulator into its shared memory (over the now idle operand ring), column-slice major,
//   3. cluster barrier; CTA j sums column slice j of all K slices through distributed shared memory (fixed order),
//      keeps the result in its own shared memory and pushes its per-sample (sum, sumsq) to every CTA of the cluster,
//   4. cluster barrier; GroupNorm(1, C) statistics of the tile's whole samples are now local: normalise (+GELU, +time
//      embedding, +FiLM) and store the bf16 activation.  Neither the raw conv output nor a partial tile touches HBM.
// Requires whole samples per 128-row tile (H == Hb) with 4..32 rows per sample; cluster size n_tiles * cl_ks <= 8.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mapa_u32(uint32_t laddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(laddr), "r"(rank));
  return r;
}
__device__ __forceinline__ float4 ld_dsmem_v4(uint32_t addr) {
  float4 v;
  // volatile keeps it ordered against the (volatile) cluster barriers; no "memory" clobber so a batch of loads stays in flight
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void st_dsmem_v2(uint32_t addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}

// 320 threads: warp 0 TMA, warp 1 MMA, warps 2..9 dump the accumulator (two warps per TMEM lane quarter, half of the columns each);
// the reduction and the apply phase are walked by all 320 threads.
constexpr int CL_THREADS = 320;

// Sum column slice `ks` of the tile over its KS K-slice partials (the own one from local shared memory, the others over DSMEM),
// leave the reduced slice in `blk`, and put per-warp-run (sum, sumsq) partials into s_part.  A remote load costs ~1000 cycles of
// latency plus ~230 cycles per 16 bytes x 320 threads (profiles/r02_chain_phase_cycles.txt), so every thread keeps KS * U = 16
// of them in flight whatever the split is (was 2 * KS: a 2-way split paid the latency three times per slice).
template <int KS, int U>
__device__ __forceinline__ void cl_reduce_slice(float* blk, const uint32_t* src, int ks, int n4, int lw4, int W4, int RS, int tid, int lane,
                                                float (*s_part)[2]) {
#pragma unroll 1
  for (int i0 = tid; i0 < n4; i0 += U * CL_THREADS) {
    float4 v[KS][U];
    int off[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * CL_THREADS;
      ok[u] = i < n4;                              // n4 is a multiple of 32: warp-uniform
      const int row = i >> lw4, c4 = i & (W4 - 1);
      off[u] = ok[u] ? row * RS + c4 * 4 : 0;
    }
#pragma unroll
    for (int sp = 0; sp < KS; ++sp) {
      if (sp == ks) {   // this CTA's own partial: a plain shared-memory load (ld.shared::cluster is slow even to oneself)
#pragma unroll
        for (int u = 0; u < U; ++u) v[sp][u] = ok[u] ? *reinterpret_cast<const float4*>(blk + off[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) v[sp][u] = ok[u] ? ld_dsmem_v4(src[sp] + (uint32_t)off[u] * 4u) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float4 acc = v[0][u];
#pragma unroll
      for (int sp = 1; sp < KS; ++sp) { acc.x += v[sp][u].x; acc.y += v[sp][u].y; acc.z += v[sp][u].z; acc.w += v[sp][u].w; }
      if (ok[u]) *reinterpret_cast<float4*>(blk + off[u]) = acc;
      float ps = (acc.x + acc.y) + (acc.z + acc.w);
      float pq = fmaf(acc.x, acc.x, fmaf(acc.y, acc.y, fmaf(acc.z, acc.z, acc.w * acc.w)));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { ps += __shfl_xor_sync(0xffffffffu, ps, o); pq += __shfl_xor_sync(0xffffffffu, pq, o); }
      // the 32 float4 of a warp are 128 / Wd whole rows of ONE sample (rps >= 4 rows, aligned): slot = first index / 32
      if (lane == 0 && ok[u]) { const int q = (i0 + u * CL_THREADS) >> 5; s_part[q][0] = ps; s_part[q][1] = pq; }
    }
  }
}
__device__ __forceinline__ void cl_reduce_dispatch(int cl_ks, float* blk, const uint32_t* src, int ks, int n4, int lw4, int W4, int RS, int tid,
                                                   int lane, float (*s_part)[2]) {
  switch (cl_ks) {
    case 1: cl_reduce_slice<1, 8>(blk, src, ks, n4, lw4, W4, RS, tid, lane, s_part); break;
    case 2: cl_reduce_slice<2, 8>(blk, src, ks, n4, lw4, W4, RS, tid, lane, s_part); break;
    case 4: cl_reduce_slice<4, 4>(blk, src, ks, n4, lw4, W4, RS, tid, lane, s_part); break;
    default: cl_reduce_slice<8, 2>(blk, src, ks, n4, lw4, W4, RS, tid, lane, s_part); break;
  }
}
template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(CL_THREADS, 1)
conv_tc_cluster_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcParams p) {
  constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
  constexpr int TMEM_COLS = BLOCK_N <= 128 ? 128 : 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(8) float s_stat[32][8][2];   // [sample of the tile][source CTA rank] (sum, sumsq), filled by remote stores
  __shared__ float s_part[128][2];                   // (sum, sumsq) of every warp-sized run of the reduced slice
  __shared__ float s_mr[32][2];                      // per sample: rstd, -mean * rstd
  __shared__ __align__(16) float s_par[3][128];      // gamma, beta, time embedding of this CTA's channel slice
  __shared__ __align__(16) float s_film[4096];       // FiLM (scale | bias) of the tile's samples for the slice: [sample][2][Wd]
  long long tstamp[7];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int CS = p.n_tiles * p.cl_ks;
  const int m_tile = blockIdx.x / CS;
  const int nt = (int)rank % p.n_tiles, ks = (int)rank / p.n_tiles;
  const int b0 = m_tile * p.Bt;                  // whole samples per tile: h0 == 0
  const int n0 = nt * BLOCK_N;

  const bool skip_dx = (p.W == 1), skip_dy = (p.H == 1);
  const int ntx = skip_dx ? 1 : 3, nty = skip_dy ? 1 : 3;
  const int k_iters = ntx * nty * p.kb_per_tap;
  const int it_begin = ks * k_iters / p.cl_ks, it_end = (ks + 1) * k_iters / p.cl_ks;   // host guarantees cl_ks <= k_iters

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(&tmem_base_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (!(warp == 0 && lane == 0)) { pdl_wait(); pdl_trigger(); }   // the producer first puts its weight tiles in flight (below)

  const int Wd = BLOCK_N / p.cl_ks;              // columns of the tile this CTA reduces and finishes
  const int RS = Wd + 4;                         // padded row stride (floats): 16-byte row-per-lane accesses stay conflict-free
  float* red = reinterpret_cast<float*>(smem);   // [cl_ks slices][128 rows][RS], over the operand ring once every MMA has retired
  const int r_t = (warp & 3) * 32 + lane;        // epilogue threads: tile row == TMEM lane

  if (warp == 0) {
    if (lane == 0) {
      uint32_t n_pre = 0;
      for (int it = it_begin; it < it_end && n_pre < (uint32_t)STAGES; ++it, ++n_pre) {   // weights only: before the dependency wait
        const int tap_i = it / p.kb_per_tap, kb = it - tap_i * p.kb_per_tap;
        const int ty = tap_i / ntx, tx = tap_i - ty * ntx;
        const int dy = skip_dy ? 0 : ty - 1, dx = skip_dx ? 0 : tx - 1;
        const int tap = (dy + 1) * 3 + (dx + 1);
        mbar_expect_tx(&full_bar[n_pre], A_STAGE_BYTES + B_STAGE_BYTES);
        tma_load_2d(smem_b + n_pre * B_STAGE_BYTES, &map_b, &full_bar[n_pre], tap * p.Cin + kb * BLOCK_K, n0);
      }
      pdl_wait();
      pdl_trigger();
      uint32_t kit = 0;
      for (int it = it_begin; it < it_end; ++it, ++kit) {
        const int s = kit % STAGES;
        const uint32_t ph = (kit / STAGES) & 1u;
        const int tap_i = it / p.kb_per_tap, kb = it - tap_i * p.kb_per_tap;
        const int ty = tap_i / ntx, tx = tap_i - ty * ntx;
        const int dy = skip_dy ? 0 : ty - 1, dx = skip_dx ? 0 : tx - 1;
        const int tap = (dy + 1) * 3 + (dx + 1);
        if (kit >= n_pre) {
          mbar_wait(&empty_bar[s], ph ^ 1u);
          mbar_expect_tx(&full_bar[s], A_STAGE_BYTES + B_STAGE_BYTES);
          tma_load_2d(smem_b + s * B_STAGE_BYTES, &map_b, &full_bar[s], tap * p.Cin + kb * BLOCK_K, n0);
        }
        tma_load_4d(smem_a + s * A_STAGE_BYTES, &map_a, &full_bar[s], kb * BLOCK_K, dx, dy, b0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BLOCK_N);
      uint32_t kit = 0;
      for (int it = it_begin; it < it_end; ++it, ++kit) {
        const int s = kit % STAGES;
        const uint32_t ph = (kit / STAGES) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint64_t da = make_smem_desc(smem_u32(smem_a + s * A_STAGE_BYTES));
        const uint64_t db = make_smem_desc(smem_u32(smem_b + s * B_STAGE_BYTES));
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
          umma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (it > it_begin || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&tmem_full_bar);
    }
    __syncwarp();
  } else {
    // ---- while the K slice runs: per-channel constants of this CTA's output slice -> shared memory ----
    tstamp[0] = clock64();
    {
      const int ch0 = n0 + ks * Wd;
      const bool te_cta = p.ap.temb_mode == TEMB_ROW0 || p.ap.temb_mode == TEMB_STEP;
      const float* te = te_cta ? temb_row(p.ap, 0) : nullptr;
      if (r_t < Wd) {
        s_par[0][r_t] = __ldg(p.ap.gamma + ch0 + r_t);
        s_par[1][r_t] = __ldg(p.ap.beta + ch0 + r_t);
        s_par[2][r_t] = te ? __ldg(te + ch0 + r_t) : 0.f;
      }
      const int ns = BLOCK_M / (p.Hb * p.W);
      if (p.ap.film && ns * 2 * Wd <= 4096) {
        for (int i = r_t; i < ns * 2 * Wd; i += 128) {
          const int sm = i / (2 * Wd), rem = i - sm * 2 * Wd, hf = rem / Wd, cc = rem - hf * Wd;
          s_film[i] = __ldg(p.ap.film + (size_t)(b0 + sm) * SPDM_FILM_WIDTH + p.ap.film_off + hf * p.Cout + ch0 + cc);
        }
      }
    }
    // ---- accumulator -> own shared memory, column-slice major ----
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    tstamp[1] = clock64();
    const uint32_t t_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int eh = (warp - 2) >> 2;
#pragma unroll 1
    for (int c = eh * (BLOCK_N / 2); c < (eh + 1) * (BLOCK_N / 2); c += 32) {
      uint32_t v[32];
      tmem_ld_32x32(t_addr + (uint32_t)c, v);
      tmem_ld_wait();
      const int j = c / Wd, col = c - j * Wd;      // Wd >= 32: a 32-column chunk lies inside one slice
      float* dst = red + ((size_t)j * BLOCK_M + r_t) * RS + col;
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        *reinterpret_cast<uint4*>(dst + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
    tc_fence_before();
    tstamp[2] = clock64();
  }
  cluster_sync_all();   // every partial tile of the cluster is in shared memory
  tstamp[3] = clock64();

  // ---- the slice block [128 rows][Wd] is walked as a flat list of float4 by all 320 threads: thread t takes float4
  //      t + 320 k, so a warp reads contiguous 512-byte runs of the remote tiles (the SM-to-SM network is the bound of this
  //      phase: ~12 B/clk/SM measured) and later writes contiguous bf16 runs of the output ----
  const int rps = p.Hb * p.W;                    // rows per sample (4..32, power of two)
  const int ns = BLOCK_M / rps;                  // samples of the tile
  const int W4 = Wd >> 2;                        // float4 per row of the slice (8, 16 or 32)
  const int lw4 = __ffs(W4) - 1, lrps = __ffs(rps) - 1;
  const int n4 = BLOCK_M * W4;                   // float4 of the slice
  const int tid = (int)threadIdx.x;
  float* blk = red + (size_t)ks * BLOCK_M * RS;  // this CTA's own partial of slice `ks`, then the reduced slice
  {
    // sum column slice `ks` over the K slices of this N tile (ranks sp * n_tiles + nt), fixed order
    const uint32_t blk_addr = smem_u32(blk);
    uint32_t src[8];
#pragma unroll
    for (int sp = 0; sp < 8; ++sp) src[sp] = sp < p.cl_ks ? mapa_u32(blk_addr, (uint32_t)(sp * p.n_tiles + nt)) : 0u;
    cl_reduce_dispatch(p.cl_ks, blk, src, ks, n4, lw4, W4, RS, tid, lane, s_part);
    __syncthreads();
    if (tid < ns) {   // sample tid: slots [tid * sps, (tid + 1) * sps), summed in a fixed order, pushed to every CTA of the cluster
      const int sps = (rps * W4) >> 5;
      float ts = 0.f, tq = 0.f;
      for (int q = tid * sps; q < (tid + 1) * sps; ++q) { ts += s_part[q][0]; tq += s_part[q][1]; }
      const uint32_t slot = smem_u32(&s_stat[tid][rank][0]);
      for (int d = 0; d < CS; ++d) st_dsmem_v2(mapa_u32(slot, (uint32_t)d), ts, tq);
    }
    tstamp[4] = clock64();
  }
  cluster_sync_all();   // all statistics partials have arrived; nothing remote is touched below
  tstamp[5] = clock64();

  {
    if (tid < ns) {
      float ts = 0.f, tq = 0.f;
      for (int d = 0; d < CS; ++d) { ts += s_stat[tid][d][0]; tq += s_stat[tid][d][1]; }
      const float inv_n = 1.0f / ((float)rps * (float)p.Cout);
      const float mean = ts * inv_n;
      const float rstd = rsqrtf(fmaxf(tq * inv_n - mean * mean, 0.f) + p.ap.eps);
      s_mr[tid][0] = rstd;
      s_mr[tid][1] = -mean * rstd;
    }
    __syncthreads();
    const int ch0 = n0 + ks * Wd;
    const bool film_smem = p.ap.film && ns * 2 * Wd <= 4096;
#pragma unroll 4
    for (int i = tid; i < n4; i += CL_THREADS) {
      const int row = i >> lw4, c = (i & (W4 - 1)) * 4;
      const int sm = row >> lrps;
      const float A = s_mr[sm][0], Bm = s_mr[sm][1];
      const float4 x = *reinterpret_cast<const float4*>(blk + row * RS + c);
      const float4 g = *reinterpret_cast<const float4*>(&s_par[0][c]);
      const float4 e = *reinterpret_cast<const float4*>(&s_par[1][c]);
      float4 t = *reinterpret_cast<const float4*>(&s_par[2][c]);
      if (p.ap.temb_mode == TEMB_PER_SAMPLE) t = __ldg(reinterpret_cast<const float4*>(temb_row(p.ap, b0 + sm) + ch0 + c));
      float f[4] = {fmaf(fmaf(x.x, A, Bm), g.x, e.x), fmaf(fmaf(x.y, A, Bm), g.y, e.y), fmaf(fmaf(x.z, A, Bm), g.z, e.z),
                    fmaf(fmaf(x.w, A, Bm), g.w, e.w)};
      if (p.ap.act == ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 4; ++j) f[j] = gelu_fast_tc(f[j]);
      }
      f[0] += t.x; f[1] += t.y; f[2] += t.z; f[3] += t.w;
      if (p.ap.film) {
        float4 fs, fb;
        if (film_smem) {
          fs = *reinterpret_cast<const float4*>(&s_film[sm * 2 * Wd + c]);
          fb = *reinterpret_cast<const float4*>(&s_film[sm * 2 * Wd + Wd + c]);
        } else {
          const float* fi = p.ap.film + (size_t)(b0 + sm) * SPDM_FILM_WIDTH + p.ap.film_off + ch0 + c;
          fs = __ldg(reinterpret_cast<const float4*>(fi));
          fb = __ldg(reinterpret_cast<const float4*>(fi + p.Cout));
        }
        f[0] = fmaf(fs.x, f[0], fb.x); f[1] = fmaf(fs.y, f[1], fb.y); f[2] = fmaf(fs.z, f[2], fb.z); f[3] = fmaf(fs.w, f[3], fb.w);
      }
      uint2 o;
      __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
      ho[0] = __floats2bfloat162_rn(f[0], f[1]);
      ho[1] = __floats2bfloat162_rn(f[2], f[3]);
      *reinterpret_cast<uint2*>(p.out + ((long long)m_tile * BLOCK_M + row) * p.ld_out + ch0 + c) = o;
    }
  }
  tstamp[6] = clock64();
  if ((p.dbg & 2048) && blockIdx.x == 0 && threadIdx.x == 64)
    printf("spdm cluster timing H%d Cin%d Cout%d ks%d nt%d: mainloop %lld dump %lld sync1 %lld reduce %lld sync2 %lld apply %lld total %lld cycles\n", p.H, p.Cin,
           p.Cout, p.cl_ks, p.n_tiles, tstamp[1] - tstamp[0], tstamp[2] - tstamp[1], tstamp[3] - tstamp[2], tstamp[4] - tstamp[3],
           tstamp[5] - tstamp[4], tstamp[6] - tstamp[5], tstamp[6] - tstamp[0]);
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}


// ---------------------------------------------------------------------------------------------
// Chain of cluster convs: a whole run of [3x3 conv + GroupNorm (+GELU / time embedding / FiLM)] layers of one U-Net level in
// ONE launch (the 8x2 / 4x1 levels at small batch: down2/down3, the bottleneck, up1 -- 18 of the 54 launches of a batch-256
// step, each ~13 us of which only ~7 are work).
//
// At these levels a 128-row M tile holds whole samples (H == Hb), every layer of the run has the same geometry and GroupNorm is
// per sample, so M tile m of layer l+1 depends on M tile m of layer l only.  The cluster that owns tile m therefore walks the
// layers by itself: per layer the phases of conv_tc_cluster_kernel (K slice -> TMEM, dump to shared memory, DSMEM
// reduce-scatter, statistics exchange, normalise + activate + store), then a cluster barrier instead of a kernel boundary --
// the activations it wrote (generic proxy, other CTAs of the cluster) become the next layer's TMA operand after
// fence.proxy.async + barrier.cluster (release / acquire).  No grid-wide synchronisation, so clusters need not be co-resident.
// The MaxPool2d(2) / bilinear x2 upsample (+ channel concat) in front of the run is done by the cluster for its own samples
// before the first layer.  Tensor maps and per-layer constants live in a device array (ChainLayer).
// ---------------------------------------------------------------------------------------------
struct alignas(128) ChainLayer {
  CUtensorMap map_a, map_b;
  int Cin, Cout, kb_per_tap;
  int bn, n_tiles, ks;        // N tile width, N tiles per M tile, K slices per tile: n_tiles * ks == cluster size
  int ld_out, pad;
  bf16* out;
  ApplyArgs ap;
};
struct ChainParams {
  int H, W, Bt;               // geometry of the level; Bt whole samples per 128-row tile (H * W * Bt == 128)
  int n_layers;
  const ChainLayer* layers;   // device memory
  int pre_kind;               // 0 none, 1 MaxPool2d(2) from a (2H x 2W) map, 2 bilinear x2 upsample (align_corners) from (H/2 x W/2)
  const bf16* pre_in; int pre_ld_in; bf16* pre_out; int pre_ld_out; int pre_C;
  int dbg;                    // 2048: thread 64 of block 0 prints per-phase cycle counts of every layer (SPDM_CL_TIMING=1)
  int step_off;               // replaces ApplyArgs::step_off of every layer (one chain object serves every step of a captured graph)
};
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

constexpr int CH_STAGES = 4;
constexpr int CH_B_SLOT = 256 * BLOCK_K * 2;                       // 32 KB: a 256-wide weight tile (128-wide layers use half)
constexpr int CH_SMEM = CH_STAGES * (A_STAGE_BYTES + CH_B_SLOT) + 1024;

__global__ void __launch_bounds__(CL_THREADS, 1) conv_chain_kernel(const ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + CH_STAGES * A_STAGE_BYTES;
  __shared__ __align__(8) uint64_t full_bar[CH_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[CH_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(8) float s_stat[32][8][2];   // [sample of the tile][source CTA rank] (sum, sumsq), filled by remote stores
  __shared__ float s_part[128][2];
  __shared__ float s_mr[32][2];
  __shared__ __align__(16) float s_par[3][128];
  __shared__ __align__(16) float s_film[4096];
  __shared__ long long ts_log[16][9];              // SPDM_CL_TIMING only

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tid = (int)threadIdx.x;
  const uint32_t rank = cluster_ctarank();
  uint32_t CS;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(CS));
  const int m_tile = blockIdx.x / (int)CS;
  const int b0 = m_tile * p.Bt;
  const bool skip_dx = (p.W == 1), skip_dy = (p.H == 1);
  const int ntx = skip_dx ? 1 : 3, nty = skip_dy ? 1 : 3;
  const int rps = p.H * p.W;                     // rows per sample (4..32, power of two)
  const int ns = BLOCK_M / rps;                  // samples of the tile
  const int lrps = __ffs(rps) - 1;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < CH_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<256>(&tmem_base_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();
  pdl_trigger();

  // ---- resample in front of the run: this cluster's samples only ----
  if (p.pre_kind) {
    const int vpr = p.pre_C >> 3;
    const int total = p.Bt * rps * vpr;
    for (int v = (int)rank * CL_THREADS + tid; v < total; v += (int)CS * CL_THREADS) {
      const int orow_l = v / vpr, c8 = (v - orow_l * vpr) << 3;
      const int wo = orow_l % p.W, t2 = orow_l / p.W, ho = t2 % p.H, bl = t2 / p.H;
      const long long b = b0 + bl;
      float y[8];
      if (p.pre_kind == 1) {
        const int Wi = p.W * 2, Hi = p.H * 2;
        const long long r00 = (b * Hi + 2 * ho) * Wi + 2 * wo;
        float x[8];
        load8(p.pre_in + r00 * p.pre_ld_in + c8, y);
        load8(p.pre_in + (r00 + 1) * p.pre_ld_in + c8, x);
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = fmaxf(y[i], x[i]);
        load8(p.pre_in + (r00 + Wi) * p.pre_ld_in + c8, x);
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = fmaxf(y[i], x[i]);
        load8(p.pre_in + (r00 + Wi + 1) * p.pre_ld_in + c8, x);
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = fmaxf(y[i], x[i]);
      } else {   // same arithmetic as upsample_kernel (kernels.cu)
        const int Hi = p.H / 2, Wi = p.W / 2;
        const float sh = (p.H > 1) ? (float)(Hi - 1) / (float)(p.H - 1) : 0.f;
        const float sw = (p.W > 1) ? (float)(Wi - 1) / (float)(p.W - 1) : 0.f;
        const float fh = sh * ho, fw = sw * wo;
        const int h0 = (int)fh, w0 = (int)fw;
        const int h1 = h0 + ((h0 < Hi - 1) ? 1 : 0), w1 = w0 + ((w0 < Wi - 1) ? 1 : 0);
        const float lh1 = fh - h0, lh0 = 1.f - lh1, lw1 = fw - w0, lw0 = 1.f - lw1;
        float a00[8], a01[8], a10[8], a11[8];
        const long long base = b * Hi;
        load8(p.pre_in + ((base + h0) * Wi + w0) * p.pre_ld_in + c8, a00);
        load8(p.pre_in + ((base + h0) * Wi + w1) * p.pre_ld_in + c8, a01);
        load8(p.pre_in + ((base + h1) * Wi + w0) * p.pre_ld_in + c8, a10);
        load8(p.pre_in + ((base + h1) * Wi + w1) * p.pre_ld_in + c8, a11);
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = lh0 * (lw0 * a00[i] + lw1 * a01[i]) + lh1 * (lw0 * a10[i] + lw1 * a11[i]);
      }
      store8(p.pre_out + ((long long)b0 * rps + orow_l) * p.pre_ld_out + c8, y);
    }
    fence_proxy_async_all();
    cluster_sync_all();
  }

  uint32_t kit = 0;                              // k-steps issued so far (producer and MMA warp count the same sequence)
  const int r_t = (warp & 3) * 32 + lane;        // epilogue threads: tile row == TMEM lane
  long long ts[8];
#pragma unroll 1
  for (int l = 0; l < p.n_layers; ++l) {
    ts[0] = clock64();
    const ChainLayer* L = p.layers + l;
    const int bn = L->bn, n_tiles = L->n_tiles, cl_ks = L->ks, kb_per_tap = L->kb_per_tap, Cin = L->Cin, Cout = L->Cout;
    const int nt = (int)rank % n_tiles, ks = (int)rank / n_tiles;
    const int n0 = nt * bn;
    const int k_iters = ntx * nty * kb_per_tap;
    const int it_begin = ks * k_iters / cl_ks, it_end = (ks + 1) * k_iters / cl_ks;
    const int Wd = bn / cl_ks;                   // columns of the tile this CTA reduces and finishes (>= 32)
    const int RS = Wd + 4;
    float* red = reinterpret_cast<float*>(smem);
    const uint32_t stage_tx = (uint32_t)(A_STAGE_BYTES + bn * BLOCK_K * 2);

    if (warp == 0) {
      if (lane == 0) {
        fence_proxy_async_all();                 // the ring was last touched through the generic proxy (reduction scratch)
        tma_prefetch_desc(&L->map_a);
        tma_prefetch_desc(&L->map_b);
        for (int it = it_begin; it < it_end; ++it, ++kit) {
          const int s = kit % CH_STAGES;
          const uint32_t ph = (kit / CH_STAGES) & 1u;
          const int tap_i = it / kb_per_tap, kb = it - tap_i * kb_per_tap;
          const int ty = tap_i / ntx, tx = tap_i - ty * ntx;
          const int dy = skip_dy ? 0 : ty - 1, dx = skip_dx ? 0 : tx - 1;
          const int tap = (dy + 1) * 3 + (dx + 1);
          mbar_wait(&empty_bar[s], ph ^ 1u);
          mbar_expect_tx(&full_bar[s], stage_tx);
          tma_load_2d(smem_b + s * CH_B_SLOT, &L->map_b, &full_bar[s], tap * Cin + kb * BLOCK_K, n0);
          tma_load_4d(smem_a + s * A_STAGE_BYTES, &L->map_a, &full_bar[s], kb * BLOCK_K, dx, dy, b0);
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = make_idesc(bn);
        tc_fence_after();
        for (int it = it_begin; it < it_end; ++it, ++kit) {
          const int s = kit % CH_STAGES;
          const uint32_t ph = (kit / CH_STAGES) & 1u;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint64_t da = make_smem_desc(smem_u32(smem_a + s * A_STAGE_BYTES));
          const uint64_t db = make_smem_desc(smem_u32(smem_b + s * CH_B_SLOT));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (it > it_begin || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full_bar);
      }
      __syncwarp();
    } else {
      // ---- while the K slice runs: per-channel constants of this CTA's output slice -> shared memory ----
      {
        const int ch0 = n0 + ks * Wd;
        const bool te_cta = L->ap.temb_mode == TEMB_ROW0 || L->ap.temb_mode == TEMB_STEP;
        const float* te = te_cta ? temb_row_off(L->ap, 0, p.step_off) : nullptr;
        const int et = tid - 64;                 // 0..255 over the epilogue warps
        if (et < Wd) {
          s_par[0][et] = __ldg(L->ap.gamma + ch0 + et);
          s_par[1][et] = __ldg(L->ap.beta + ch0 + et);
          s_par[2][et] = te ? __ldg(te + ch0 + et) : 0.f;
        }
        if (L->ap.film && ns * 2 * Wd <= 4096) {
          for (int i = et; i < ns * 2 * Wd; i += 256) {
            const int sm = i / (2 * Wd), rem = i - sm * 2 * Wd, hf = rem / Wd, cc = rem - hf * Wd;
            s_film[i] = __ldg(L->ap.film + (size_t)(b0 + sm) * SPDM_FILM_WIDTH + L->ap.film_off + hf * Cout + ch0 + cc);
          }
        }
      }
      // ---- accumulator -> own shared memory, column-slice major ----
      ts[1] = clock64();
      mbar_wait(&tmem_full_bar, (uint32_t)l & 1u);
      tc_fence_after();
      ts[2] = clock64();
      const uint32_t t_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
      const int eh = (warp - 2) >> 2;
#pragma unroll 1
      for (int c = eh * (bn / 2); c < (eh + 1) * (bn / 2); c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_addr + (uint32_t)c, v);
        tmem_ld_wait();
        const int j = c / Wd, col = c - j * Wd;    // Wd >= 32: a 32-column chunk lies inside one slice
        float* dst = red + ((size_t)j * BLOCK_M + r_t) * RS + col;
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<uint4*>(dst + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
      tc_fence_before();
    }
    ts[3] = clock64();
    cluster_sync_all();   // every partial tile of the cluster is in shared memory
    ts[4] = clock64();

    const int W4 = Wd >> 2;
    const int lw4 = __ffs(W4) - 1;
    const int n4 = BLOCK_M * W4;
    float* blk = red + (size_t)ks * BLOCK_M * RS;
    {
      const uint32_t blk_addr = smem_u32(blk);
      uint32_t src[8];
#pragma unroll
      for (int sp = 0; sp < 8; ++sp) src[sp] = sp < cl_ks ? mapa_u32(blk_addr, (uint32_t)(sp * n_tiles + nt)) : 0u;
      cl_reduce_dispatch(cl_ks, blk, src, ks, n4, lw4, W4, RS, tid, lane, s_part);
      __syncthreads();
      if (tid < ns) {
        const int sps = (rps * W4) >> 5;
        float tsum = 0.f, tq = 0.f;
        for (int q = tid * sps; q < (tid + 1) * sps; ++q) { tsum += s_part[q][0]; tq += s_part[q][1]; }
        const uint32_t slot = smem_u32(&s_stat[tid][rank][0]);
        for (int d = 0; d < (int)CS; ++d) st_dsmem_v2(mapa_u32(slot, (uint32_t)d), tsum, tq);
      }
    }
    ts[5] = clock64();
    cluster_sync_all();   // all statistics partials have arrived; nothing remote is touched below
    ts[6] = clock64();

    {
      if (tid < ns) {
        float tsum = 0.f, tq = 0.f;
        for (int d = 0; d < (int)CS; ++d) { tsum += s_stat[tid][d][0]; tq += s_stat[tid][d][1]; }
        const float inv_n = 1.0f / ((float)rps * (float)Cout);
        const float mean = tsum * inv_n;
        const float rstd = rsqrtf(fmaxf(tq * inv_n - mean * mean, 0.f) + L->ap.eps);
        s_mr[tid][0] = rstd;
        s_mr[tid][1] = -mean * rstd;
      }
      __syncthreads();
      const int ch0 = n0 + ks * Wd;
      const bool has_film = L->ap.film != nullptr;
      const bool film_smem = has_film && ns * 2 * Wd <= 4096;
      const int act = L->ap.act, temb_mode = L->ap.temb_mode;
      bf16* outp = L->out;
      const int ld_out = L->ld_out;
#pragma unroll 4
      for (int i = tid; i < n4; i += CL_THREADS) {
        const int row = i >> lw4, c = (i & (W4 - 1)) * 4;
        const int sm = row >> lrps;
        const float A = s_mr[sm][0], Bm = s_mr[sm][1];
        const float4 x = *reinterpret_cast<const float4*>(blk + row * RS + c);
        const float4 g = *reinterpret_cast<const float4*>(&s_par[0][c]);
        const float4 e = *reinterpret_cast<const float4*>(&s_par[1][c]);
        float4 t = *reinterpret_cast<const float4*>(&s_par[2][c]);
        if (temb_mode == TEMB_PER_SAMPLE) t = __ldg(reinterpret_cast<const float4*>(temb_row_off(L->ap, b0 + sm, p.step_off) + ch0 + c));
        float f[4] = {fmaf(fmaf(x.x, A, Bm), g.x, e.x), fmaf(fmaf(x.y, A, Bm), g.y, e.y), fmaf(fmaf(x.z, A, Bm), g.z, e.z),
                      fmaf(fmaf(x.w, A, Bm), g.w, e.w)};
        if (act == ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 4; ++j) f[j] = gelu_fast_tc(f[j]);
        }
        f[0] += t.x; f[1] += t.y; f[2] += t.z; f[3] += t.w;
        if (has_film) {
          float4 fs, fb;
          if (film_smem) {
            fs = *reinterpret_cast<const float4*>(&s_film[sm * 2 * Wd + c]);
            fb = *reinterpret_cast<const float4*>(&s_film[sm * 2 * Wd + Wd + c]);
          } else {
            const float* fi = L->ap.film + (size_t)(b0 + sm) * SPDM_FILM_WIDTH + L->ap.film_off + ch0 + c;
            fs = __ldg(reinterpret_cast<const float4*>(fi));
            fb = __ldg(reinterpret_cast<const float4*>(fi + Cout));
          }
          f[0] = fmaf(fs.x, f[0], fb.x); f[1] = fmaf(fs.y, f[1], fb.y); f[2] = fmaf(fs.z, f[2], fb.z); f[3] = fmaf(fs.w, f[3], fb.w);
        }
        uint2 o;
        __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
        ho[0] = __floats2bfloat162_rn(f[0], f[1]);
        ho[1] = __floats2bfloat162_rn(f[2], f[3]);
        *reinterpret_cast<uint2*>(outp + ((long long)m_tile * BLOCK_M + row) * ld_out + ch0 + c) = o;
      }
    }
    ts[7] = clock64();
    if (l + 1 < p.n_layers) {
      // this layer's activations (written by every CTA of the cluster through the generic proxy) are the next layer's TMA
      // operand; the reduction scratch is about to be overwritten by TMA as well
      fence_proxy_async_all();
      cluster_sync_all();
    }
    if ((p.dbg & 2048) && blockIdx.x == 0 && threadIdx.x == 64 && l < 16) {   // logged, printed after the last layer (a printf here would delay this warp into the next layer's phases)
#pragma unroll
      for (int i = 0; i < 8; ++i) ts_log[l][i] = ts[i];
      ts_log[l][8] = clock64();
    }
  }
  if ((p.dbg & 2048) && blockIdx.x == 0 && threadIdx.x == 64)
    for (int l = 0; l < p.n_layers && l < 16; ++l) {
      const ChainLayer* L = p.layers + l;
      const long long* t = ts_log[l];
      printf("spdm chain timing layer %d (%d->%d bn %d ks %d): const %lld mainloop-wait %lld dump %lld sync1 %lld reduce %lld sync2 %lld apply %lld endsync %lld total %lld\n",
             l, L->Cin, L->Cout, L->bn, L->ks, t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4], t[6] - t[5], t[7] - t[6], t[8] - t[7], t[8] - t[0]);
    }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// Swapped-operand variant for narrow outputs (Cout = 64 or 128).
//
// A tcgen05.mma in SS mode costs ~128 cycles per 128-row A operand regardless of N (measured: tests/perf_conv.py),
// so an N = 64/128 tile runs the tensor pipe at 25-50 %.  Here the roles are exchanged: D^T[Cout][pixels] =
// W[Cout][K] * X^T[K][pixels] with M = 128 weight rows (rows >= Cout are TMA out-of-bounds zeros) and N = 256 pixels
// (two stacked 128-row boxes of the activation).  The accumulator then holds channels in TMEM lanes and pixels in
// columns; the epilogue writes the channels-last output with 64-byte warp stores (one pixel per instruction).
// Only EPI_STATS epilogues (the 3x3 convolutions) take this path.
// ---------------------------------------------------------------------------------------------
// WM = weight rows per tile = M of the MMA: 128 (rows >= Cout are TMA zero fill) or 64 for the 64-channel layers, whose
// accumulator then sits in 16 lanes of each TMEM lane quarter (row i -> lane 32*(i/16) + i%16).
// 320 threads: warp 0 TMA, warp 1 MMA, warps 2..9 epilogue -- two warps per TMEM lane quarter (warp % 4), each owns half of the
// tile's 256 pixel columns.  The epilogue (one 2-byte store per channel lane and pixel, half the lanes idle when WM = 64) is what
// the narrow layers are bound by once the tile's MMAs take only a few microseconds.
constexpr int SWAP_THREADS = 320;
constexpr int SWAP_THREADS16 = 576;   // 16 epilogue warps: fused GroupNorm-apply epilogues of persistent launches (instruction-latency bound at 2 warps per scheduler)
template <int STAGES, int WM>
__global__ void __launch_bounds__(SWAP_THREADS16, 1)
conv_tc_swap_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_h,
                    const TcParams p) {
  constexpr int PIX_BYTES = 2 * A_STAGE_BYTES;           // 256 pixels x 64 ch
  constexpr int W_BYTES = WM * BLOCK_K * 2;              // WM weight rows x 64 k
  constexpr int NPIX = 256;
  constexpr int TMEM_COLS = 2 * NPIX;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_pix = smem;
  uint8_t* smem_w = smem + STAGES * PIX_BYTES;
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  // halo mode: a ring of HP_STAGES (32 + 2)-row pixel boxes and a ring of HW_STAGES weight tiles, each with its own barriers
  constexpr int HP_STAGES = 3, HW_STAGES = 5;
  constexpr int HALO_BYTES = 34 * 8 * BLOCK_K * 2;       // 34 image rows x 8 pixels x 64 channels
  constexpr int HALO_STRIDE = 35 * 1024;
  static_assert(HP_STAGES * HALO_STRIDE + HW_STAGES * W_BYTES <= STAGES * (PIX_BYTES + W_BYTES), "halo rings must fit the stage ring");
  __shared__ __align__(8) uint64_t hp_full[HP_STAGES], hp_empty[HP_STAGES], hw_full[HW_STAGES], hw_empty[HW_STAGES];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_part[2][16][4][4][2];  // EPI_APPLY: per-sample, per-warp (lane quarter, column segment) (sum, sumsq); double-buffered like TMEM
  __shared__ float s_mr[2][16][2];       // per-sample (mean, rstd)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool skip_dx = (p.W == 1), skip_dy = (p.H == 1);
  const int ntx = skip_dx ? 1 : 3, nty = skip_dy ? 1 : 3;
  const int k_iters = p.fold ? 12 * p.kb_per_tap : ntx * nty * p.kb_per_tap;
  const int c_tiles = p.n_tiles;             // tiles of 128 output channels
  const int total = p.total_tiles;           // (m_tiles / 2) * c_tiles

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_w);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
#pragma unroll
    for (int s = 0; s < HP_STAGES; ++s) { mbar_init(&hp_full[s], 1); mbar_init(&hp_empty[s], 1); }
#pragma unroll
    for (int s = 0; s < HW_STAGES; ++s) { mbar_init(&hw_full[s], 1); mbar_init(&hw_empty[s], 1); }
    mbar_init(&tmem_full_bar[0], 1); mbar_init(&tmem_full_bar[1], 1);
    const uint32_t n_epi_w = (blockDim.x >> 5) - 2;   // 4 or 8 epilogue warps (chosen at launch)
    mbar_init(&tmem_empty_bar[0], n_epi_w); mbar_init(&tmem_empty_bar[1], n_epi_w);   // one arrival per epilogue warp
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(&tmem_base_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  // as in conv_tc_kernel: the producer puts the weight tiles of its first STAGES k-steps in flight before the dependency wait
  if (!(warp == 0 && lane == 0)) { pdl_wait(); pdl_trigger(); }
  TcParams pm = p;       // decode_tile works on 128-row tiles with n_tiles == 1
  pm.n_tiles = 1;

  uint8_t* smem_hw = smem + HP_STAGES * HALO_STRIDE;     // halo mode: weight ring behind the pixel ring
  if (warp == 0 && p.halo) {
    if (lane == 0 && !(p.dbg & 16)) {   // dbg 16: MMA free-run microbenchmark (no loads, no operand waits)
      // groups (k block kb, dx); per group one pixel box and the three dy weight tiles, in the order the MMA warp consumes them
      const int groups = 3 * p.kb_per_tap;
      auto wcol_of = [&](int grp, int ty) { const int kb = grp / 3, tx = grp - kb * 3; return (ty * 3 + tx) * p.Cin + kb * BLOCK_K; };
      uint32_t n_pre = 0;
      if ((int)blockIdx.x < total) {
        const int c_tile = (int)blockIdx.x % c_tiles;
        for (int w = 0; w < 3 * groups && n_pre < (uint32_t)HW_STAGES; ++w, ++n_pre) {
          mbar_expect_tx(&hw_full[n_pre], W_BYTES);
          tma_load_2d(smem_hw + n_pre * W_BYTES, &map_w, &hw_full[n_pre], wcol_of(w / 3, w % 3), c_tile * WM);
        }
      }
      pdl_wait();
      pdl_trigger();
      uint32_t wkit = 0, pkit = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int c_tile = tile % c_tiles, pt = tile / c_tiles;
        const TileCoord t0 = decode_tile(pm, 2 * pt, 0);
        for (int grp = 0; grp < groups; ++grp, ++pkit) {
          const int kb = grp / 3, dx = grp - kb * 3 - 1;
          const int ps = pkit % HP_STAGES;
          mbar_wait(&hp_empty[ps], ((pkit / HP_STAGES) & 1u) ^ 1u);
          mbar_expect_tx(&hp_full[ps], HALO_BYTES);
          tma_load_4d(smem + ps * HALO_STRIDE, &map_h, &hp_full[ps], kb * BLOCK_K, dx, t0.h0 - 1, t0.b0);
          for (int ty = 0; ty < 3; ++ty, ++wkit) {
            if (wkit < n_pre) continue;
            const int ws = wkit % HW_STAGES;
            mbar_wait(&hw_empty[ws], ((wkit / HW_STAGES) & 1u) ^ 1u);
            mbar_expect_tx(&hw_full[ws], W_BYTES);
            tma_load_2d(smem_hw + ws * W_BYTES, &map_w, &hw_full[ws], wcol_of(grp, ty), c_tile * WM);
          }
        }
      }
    }
  } else if (warp == 1 && p.halo) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_m(WM, NPIX);
      const int groups = 3 * p.kb_per_tap;
      uint32_t wkit = 0, pkit = 0;
      int lt = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++lt) {
        const int acc = lt & 1;
        const uint32_t aph = (uint32_t)(lt >> 1) & 1u;
        mbar_wait(&tmem_empty_bar[acc], aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NPIX);
        for (int grp = 0; grp < groups; ++grp, ++pkit) {
          const int ps = pkit % HP_STAGES;
          if (!(p.dbg & 16)) mbar_wait(&hp_full[ps], (pkit / HP_STAGES) & 1u);
          for (int ty = 0; ty < 3; ++ty, ++wkit) {
            const int ws = wkit % HW_STAGES;
            if (!(p.dbg & 16)) mbar_wait(&hw_full[ws], (wkit / HW_STAGES) & 1u);
            tc_fence_after();
            const uint64_t dw = make_smem_desc(smem_u32(smem_hw + ws * W_BYTES));                  // "A": weight rows of tap (dy = ty - 1, dx)
            const uint64_t dp = make_smem_desc(smem_u32(smem + ps * HALO_STRIDE + ty * 1024));     // "B": 256 pixel rows from image row h0 - 1 + ty
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_bf16(d_tmem, dw + (uint64_t)(2 * k), dp + (uint64_t)(2 * k), idesc, (grp > 0 || ty > 0 || k > 0) ? 1u : 0u);
            if (!(p.dbg & 16)) umma_commit(&hw_empty[ws]);
          }
          if (!(p.dbg & 16)) umma_commit(&hp_empty[ps]);
        }
        umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else if (warp == 0) {
    if (lane == 0 && !(p.dbg & 16)) {
      // k-step -> weight column, channel offset of the pixel box, (dx, dy) shift
      auto decode = [&](int it, int& wcol, int& c0, int& dx, int& dy) {
        if (p.fold) {
          const int per_dy = 4 * p.kb_per_tap;
          const int dyi = it / per_dy, r = it - dyi * per_dy, blk = r / p.kb_per_tap, kb = r - blk * p.kb_per_tap;
          dy = dyi - 1;
          dx = blk == 0 ? -1 : (blk == 3 ? 1 : 0);
          c0 = ((blk == 0 || blk == 2) ? p.fold_c1 : 0) + kb * BLOCK_K;
          wcol = (dyi * 4 + blk) * p.Cin + kb * BLOCK_K;
        } else {
          const int tap_i = it / p.kb_per_tap, kb = it - tap_i * p.kb_per_tap;
          const int ty = tap_i / ntx, tx = tap_i - ty * ntx;
          dy = skip_dy ? 0 : ty - 1;
          dx = skip_dx ? 0 : tx - 1;
          c0 = kb * BLOCK_K;
          wcol = ((dy + 1) * 3 + (dx + 1)) * p.Cin + kb * BLOCK_K;
        }
      };
      uint32_t n_pre = 0;
      if ((int)blockIdx.x < total) {
        const int c_tile = (int)blockIdx.x % c_tiles;
        for (int it = 0; it < k_iters && n_pre < (uint32_t)STAGES; ++it, ++n_pre) {
          int wcol, c0, dx, dy;
          decode(it, wcol, c0, dx, dy);
          mbar_expect_tx(&full_bar[n_pre], PIX_BYTES + W_BYTES);
          tma_load_2d(smem_w + n_pre * W_BYTES, &map_w, &full_bar[n_pre], wcol, c_tile * WM);
        }
      }
      pdl_wait();
      pdl_trigger();
      uint32_t kit = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int c_tile = tile % c_tiles, pt = tile / c_tiles;
        const TileCoord t0 = decode_tile(pm, 2 * pt, 0), t1 = decode_tile(pm, 2 * pt + 1, 0);
        for (int it = 0; it < k_iters; ++it, ++kit) {
          const int s = kit % STAGES;
          const uint32_t ph = (kit / STAGES) & 1u;
          int wcol, c0, dx, dy;
          decode(it, wcol, c0, dx, dy);
          if (kit >= n_pre) {
            mbar_wait(&empty_bar[s], ph ^ 1u);
            mbar_expect_tx(&full_bar[s], PIX_BYTES + W_BYTES);
            tma_load_2d(smem_w + s * W_BYTES, &map_w, &full_bar[s], wcol, c_tile * WM);
          }
          tma_load_4d(smem_pix + s * PIX_BYTES, &map_a, &full_bar[s], c0, dx, t0.h0 + dy, t0.b0);   // pix256: the box holds both tiles
          if (!p.pix256) tma_load_4d(smem_pix + s * PIX_BYTES + A_STAGE_BYTES, &map_a, &full_bar[s], c0, dx, t1.h0 + dy, t1.b0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_m(WM, NPIX);
      uint32_t kit = 0;
      int lt = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++lt) {
        const int acc = lt & 1;
        const uint32_t aph = (uint32_t)(lt >> 1) & 1u;
        mbar_wait(&tmem_empty_bar[acc], aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NPIX);
        for (int it = 0; it < k_iters; ++it, ++kit) {
          const int s = kit % STAGES;
          const uint32_t ph = (kit / STAGES) & 1u;
          if (!(p.dbg & 16)) mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint64_t dw = make_smem_desc(smem_u32(smem_w + s * W_BYTES));      // "A": 128 weight rows
          const uint64_t dp = make_smem_desc(smem_u32(smem_pix + s * PIX_BYTES));  // "B": 256 pixel rows
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_bf16(d_tmem, dw + (uint64_t)(2 * k), dp + (uint64_t)(2 * k), idesc, (it > 0 || k > 0) ? 1u : 0u);
          if (!(p.dbg & 16)) umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else {
    const int q = warp & 3;
    // 8 epilogue warps (launched where a CTA has a single tile and the epilogue is exposed): two warps per lane quarter, each owns
    // half of the tile's pixel columns.  4 warps (persistent multi-tile launches, where the epilogue hides under the next tile's
    // MMAs and extra warps only compete for issue slots: measured -1.3 % at batch 4096): every warp walks all 256 columns.
    // 16 warps (fused-apply epilogues of persistent launches): four warps per lane quarter, 64 columns each.
    const int ncs = ((int)(blockDim.x >> 5) - 2) >> 2;   // column segments of the tile: 1, 2 or 4
    const bool split = ncs == 2;
    const int eh = (warp - 2) >> 2;
    const int c_lo = eh * (NPIX / ncs), c_hi = c_lo + NPIX / ncs;
    // TMEM lane -> output channel inside the tile
    const bool layout_b = (p.dbg & 1024) != 0;          // M = 64 alternative hypothesis: rows 0..63 in lanes 0..63
    const int ch_local = (WM == 128 || layout_b) ? q * 32 + lane : q * 16 + lane;
    const bool lane_ok = (WM == 128 || layout_b) ? true : lane < 16;
    const int pps = p.H * p.W;                          // pixels per sample
    const int tiles_per_sample = pps > NPIX ? pps / NPIX : 1;
    const int nw = WM == 128 ? ((p.Cout >= BLOCK_M) ? 4 : p.Cout / 32) : (layout_b ? 2 : 4);  // warps that own real channels
    int lt = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++lt) {
      const int c_tile = tile % c_tiles, pt = tile / c_tiles;
      const int acc = lt & 1;
      const uint32_t aph = (uint32_t)(lt >> 1) & 1u;
      const long long row0 = (long long)pt * NPIX;      // the 256 pixel rows of the tile are contiguous
      mbar_wait(&tmem_full_bar[acc], aph);
      tc_fence_after();
      if (p.dbg & 32) {  // microbenchmark: hand the accumulator straight back
        tc_fence_before();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        continue;
      }
      const bool active = q < nw;
      // `ch`: element offset of this lane's channel inside an output row (pair fold: the (wo, co) row of the accumulator is pixel
      // wo of the pair, channel co); `chp`: the channel whose GroupNorm / time-embedding / FiLM parameters apply
      const int ch = p.fold ? (ch_local >> 6) * p.fold_ldo + (ch_local & 63) : c_tile * WM + ch_local;
      const int chp = p.fold ? (ch_local & 63) : c_tile * WM + ch_local;
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NPIX);
      float rs = 0.f, rq = 0.f;
      if (p.flags & EPI_APPLY) {
        // The 256-pixel tile holds NPIX/pps whole samples and every channel (host guarantees Cout <= 128, pps | 256):
        // GroupNorm is tile-local.  Pass 1 over TMEM: per-sample statistics; pass 2: normalise (+GELU/temb/FiLM) and store.
        const int n_s = NPIX / pps;
        const int span2 = pps < c_hi - c_lo ? pps : c_hi - c_lo;   // columns of one sample inside this warp's column range
        const int n_epi_thr = 128 * ncs;
        if (active) {
#pragma unroll 1
          for (int c = c_lo; c < c_hi; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32(t_addr + (uint32_t)c, v);
            tmem_ld_wait();
            if (pps >= 32) {
              float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int i = 0; i < 32; ++i) { const float f = __uint_as_float(v[i]); s4[i & 3] += f; q4[i & 3] = fmaf(f, f, q4[i & 3]); }
              rs += (s4[0] + s4[1]) + (s4[2] + s4[3]);
              rq += (q4[0] + q4[1]) + (q4[2] + q4[3]);
              if ((c + 32) % span2 == 0) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { rs += __shfl_xor_sync(0xffffffffu, rs, o); rq += __shfl_xor_sync(0xffffffffu, rq, o); }
                if (lane == 0) { s_part[acc][c / pps][q][eh][0] = rs; s_part[acc][c / pps][q][eh][1] = rq; }
                rs = 0.f; rq = 0.f;
              }
            } else {  // pps == 16
              float sa = 0.f, qa = 0.f, sb = 0.f, qb = 0.f;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float fa = __uint_as_float(v[i]), fb = __uint_as_float(v[16 + i]);
                sa += fa; qa = fmaf(fa, fa, qa); sb += fb; qb = fmaf(fb, fb, qb);
              }
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                sa += __shfl_xor_sync(0xffffffffu, sa, o); qa += __shfl_xor_sync(0xffffffffu, qa, o);
                sb += __shfl_xor_sync(0xffffffffu, sb, o); qb += __shfl_xor_sync(0xffffffffu, qb, o);
              }
              if (lane == 0) {
                s_part[acc][c / 16][q][eh][0] = sa; s_part[acc][c / 16][q][eh][1] = qa;
                s_part[acc][c / 16 + 1][q][eh][0] = sb; s_part[acc][c / 16 + 1][q][eh][1] = qb;
              }
            }
          }
        }
        asm volatile("bar.sync 1, %0;" ::"r"(n_epi_thr) : "memory");
        {
          const int et = (int)threadIdx.x - 64;  // 0..127 (255) over the epilogue warps
          if (et >= 0 && et < n_s) {
            // column halves that hold pixels of sample et (one, or both when the sample spans the whole tile)
            const int h_lo = (et * pps) / (NPIX / ncs), h_hi = ((et + 1) * pps - 1) / (NPIX / ncs);
            float ts = 0.f, tq = 0.f;
            for (int j = 0; j < nw; ++j)
              for (int hh = h_lo; hh <= h_hi; ++hh) { ts += s_part[acc][et][j][hh][0]; tq += s_part[acc][et][j][hh][1]; }
            const float inv_n = 1.0f / ((float)pps * (float)p.Cout);
            const float mean = ts * inv_n;
            s_mr[acc][et][0] = mean;
            s_mr[acc][et][1] = rsqrtf(fmaxf(tq * inv_n - mean * mean, 0.f) + p.ap.eps);
          }
        }
        asm volatile("bar.sync 1, %0;" ::"r"(n_epi_thr) : "memory");
        if (active) {
          const float g = __ldg(p.ap.gamma + chp), be = __ldg(p.ap.beta + chp);
          const long long b_first = row0 / pps;
          float te = 0.f;
          const bool has_film = p.ap.film != nullptr;
          int cur_s = -1;
          float A = 0.f, Bc = 0.f, fs = 1.f, fb = 0.f;
          auto set_sample = [&](int sl) {
            if (sl == cur_s) return;
            cur_s = sl;
            const float mean = s_mr[acc][sl][0], rstd = s_mr[acc][sl][1];
            A = rstd * g;
            Bc = fmaf(-mean, A, be);
            const float* tr = temb_row(p.ap, (int)(b_first + sl));
            te = tr ? __ldg(tr + chp) : 0.f;
            if (has_film) {
              const float* fi = p.ap.film + (size_t)(b_first + sl) * SPDM_FILM_WIDTH + p.ap.film_off;
              fs = __ldg(fi + chp);
              fb = __ldg(fi + p.ap.C + chp);
            }
          };
#pragma unroll 1
          for (int c = c_lo; c < c_hi; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32(t_addr + (uint32_t)c, v);
            tmem_ld_wait();
            if (c + 32 == c_hi) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
            }
            bf16* optr = p.out + (row0 + c) * p.ld_out + ch;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (pps >= 32) { if (i == 0) set_sample(c / pps); }
              else if ((i & 15) == 0) set_sample((c + i) / pps);
              float y = fmaf(__uint_as_float(v[i]), A, Bc);
              if (p.ap.act == ACT_GELU) y = gelu_fast_tc(y);
              y += te;
              y = fmaf(fs, y, fb);
              optr[(size_t)i * p.ld_out] = __float2bfloat16_rn(y);
            }
          }
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
        continue;
      }
      if (active) {
#pragma unroll 1
        for (int c = c_lo; c < c_hi; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(t_addr + (uint32_t)c, v);
          tmem_ld_wait();
          if (c + 32 == c_hi) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
          }
          bf16* optr = p.out + (row0 + c) * p.ld_out + ch;
          if (pps >= 32) {
            float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float f = lane_ok ? __uint_as_float(v[i]) : 0.f;
              s4[i & 3] += f; q4[i & 3] = fmaf(f, f, q4[i & 3]);
              if (lane_ok) optr[(size_t)i * p.ld_out] = __float2bfloat16_rn(f);
            }
            rs += (s4[0] + s4[1]) + (s4[2] + s4[3]);
            rq += (q4[0] + q4[1]) + (q4[2] + q4[3]);
            // columns of one sample inside this warp's half of the tile; a sample of >= 256 pixels spans both halves and gets one
            // partial slot per half
            const bool both = split && pps >= NPIX;
            const int span = both ? NPIX / 2 : (pps < NPIX ? pps : NPIX);
            if ((c + 32) % span == 0) {                 // end of a sample's columns: publish this warp's partial
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) { rs += __shfl_xor_sync(0xffffffffu, rs, o); rq += __shfl_xor_sync(0xffffffffu, rq, o); }
              if (lane == 0) {
                const long long b = (row0 + c) / pps;
                const int tile_in_sample = (int)(pt % tiles_per_sample);
                const int slot = both ? ((tile_in_sample * c_tiles + c_tile) * nw + q) * 2 + eh : (tile_in_sample * c_tiles + c_tile) * nw + q;
                float* dst = p.stats + ((size_t)b * p.P + slot) * 2;
                dst[0] = rs; dst[1] = rq;
              }
              rs = 0.f; rq = 0.f;
            }
          } else {  // pps == 16: two samples per 32-column chunk
            float sa = 0.f, qa = 0.f, sb = 0.f, qb = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float f = lane_ok ? __uint_as_float(v[i]) : 0.f;
              if (i < 16) { sa += f; qa = fmaf(f, f, qa); } else { sb += f; qb = fmaf(f, f, qb); }
              if (lane_ok) optr[(size_t)i * p.ld_out] = __float2bfloat16_rn(f);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              sa += __shfl_xor_sync(0xffffffffu, sa, o); qa += __shfl_xor_sync(0xffffffffu, qa, o);
              sb += __shfl_xor_sync(0xffffffffu, sb, o); qb += __shfl_xor_sync(0xffffffffu, qb, o);
            }
            if (lane == 0) {
              const long long b = (row0 + c) / 16;
              const int slot = c_tile * nw + q;
              float* dst = p.stats + ((size_t)b * p.P + slot) * 2;
              dst[0] = sa; dst[1] = qa;
              dst[(size_t)p.P * 2] = sb; dst[(size_t)p.P * 2 + 1] = qb;
            }
          }
        }
      } else {  // this warp's TMEM lanes hold the zero rows of a 64-channel layer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
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

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BLOCK_N, int STAGES> constexpr int smem_bytes() { return STAGES * (A_STAGE_BYTES + BLOCK_N * BLOCK_K * 2) + 1024; }

template <int BLOCK_N, int STAGES>
void launch_cfg(const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& p, cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<BLOCK_N, STAGES>());
    attr = true;
  }
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  launch_pdl(conv_tc_kernel<BLOCK_N, STAGES>, dim3(grid), dim3(NUM_THREADS), smem_bytes<BLOCK_N, STAGES>(), s, ma, mb, p);
  ++g_tc_launches;
}

// CTA-pair launch (conv_tc_kernel<256, 6, true>): clusters of two CTAs, an even grid
bool launch_pair(const CUtensorMap& ma, const CUtensorMap& mb128, const TcParams& p, cudaStream_t s) {
  constexpr int STG = 6;
  constexpr int smem = STG * (A_STAGE_BYTES + 128 * BLOCK_K * 2) + 1024;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(conv_tc_kernel<256, STG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr = true;
  }
  int grid = (p.total_tiles < num_sms() ? p.total_tiles : num_sms()) & ~1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_spdm_pdl ? 2 : 1;
  if (cudaLaunchKernelEx(&cfg, conv_tc_kernel<256, STG, true>, ma, mb128, p) != cudaSuccess) { cudaGetLastError(); return false; }
  ++g_tc_launches;
  return true;
}

}  // namespace

struct TcGemm {
  CUtensorMap map_a, map_b, map_b256, map_wswap;
  CUtensorMap map_halo;   // (32 + 2)-row pixel box of the swapped kernel's halo mode (has_halo)
  bool has_halo;
  CUtensorMap map_a256;   // 256-row pixel box: the swapped kernel's whole N tile in one TMA operation (has_a256)
  bool has_a256;
  bool can_swap;    // 3x3, Cout in {64,128}, geometry allows 256-pixel tiles with per-sample statistics
  TcParams p;
  int block_n;      // 64 or 128
  bool has256;      // Cout % 256 == 0: a 256-wide N tile is available
  int Bcap;
};

static int g_tc_dbg = 0;
void tc_set_debug(int v) { g_tc_dbg = v; }

int tc_batch_multiple(int H, int W) {
  const int hw = H * W;
  return hw >= BLOCK_M ? 1 : BLOCK_M / hw;
}

static int partials_for(const TcParams& p, int n_tiles) {
  const int rps = p.Hb * p.W;
  return rps >= 32 ? (p.H / p.Hb) * ((rps >= 128 ? 128 : rps) / 32) * n_tiles : n_tiles;
}

// 256-row pixel box of the swapped kernel: two neighbouring 128-row tiles = 2 Hb rows of one sample (maps of >= 256 pixels) or
// 2 Bt whole samples.  dims / strides as in the 128-row map.
static bool encode_a256(EncodeTiledFn enc, CUtensorMap* out, const void* in, const cuuint64_t* dims, const cuuint64_t* strides, int W, int H, int Hb,
                        int Bt, int Bcap) {
  int Hb2 = Hb, Bt2 = Bt;
  if (H / Hb > 1) { if ((H / Hb) % 2) return false; Hb2 = 2 * Hb; } else { Bt2 = 2 * Bt; if (Bcap % Bt2) return false; }
  if (Hb2 > 256 || Bt2 > 256 || W > 256) return false;
  cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)W, (cuuint32_t)Hb2, (cuuint32_t)Bt2};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
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
  g->block_n = (Cout % 128 == 0) ? 128 : (Cout == 192 ? 192 : 64);  // in_proj of a 64-channel block: one 192-wide tile
  g->has256 = (Cout % 256 == 0);
  TcParams& p = g->p;
  p.H = H; p.W = W; p.Hb = Hb; p.Bt = Bt; p.Cin = Cin; p.Cout = Cout; p.taps = taps;
  p.kb_per_tap = Cin / BLOCK_K;

  {  // A: 4-D (C, W, H, B) view of the channels-last activation
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Bcap};
    cuuint64_t strides[3] = {(cuuint64_t)ld_in * 2, (cuuint64_t)W * ld_in * 2, (cuuint64_t)H * W * ld_in * 2};
    cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)W, (cuuint32_t)Hb, (cuuint32_t)Bt};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&g->map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(g_tc_err, sizeof g_tc_err, "cuTensorMapEncodeTiled(A) failed: %d", (int)r); delete g; return nullptr; }
  }
  {
    const int pps = H * W;
    g->can_swap = taps == 9 && (Cout == 64 || Cout == 128) && (pps == 16 || (pps >= 32 && pps % 32 == 0 && (pps >= 256 ? pps % 256 == 0 : 256 % pps == 0)));
  }
  if (g->can_swap) {
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Bcap};
    cuuint64_t strides[3] = {(cuuint64_t)ld_in * 2, (cuuint64_t)W * ld_in * 2, (cuuint64_t)H * W * ld_in * 2};
    g->has_a256 = encode_a256(enc, &g->map_a256, in, dims, strides, W, H, Hb, Bt, Bcap);
  }
  if (g->can_swap && W == 8 && H % 32 == 0) {   // a 256-pixel tile is 32 whole image rows of one sample: halo mode
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Bcap};
    cuuint64_t strides[3] = {(cuuint64_t)ld_in * 2, (cuuint64_t)W * ld_in * 2, (cuuint64_t)H * W * ld_in * 2};
    cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, 8u, 34u, 1u};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&g->map_halo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    g->has_halo = r == CUDA_SUCCESS;
  }
  for (int pass = 0; pass < 3; ++pass) {  // B: 2-D (K, Cout) weights, one map per tile height
    if (pass == 1 && !g->has256) continue;
    if (pass == 2 && !g->can_swap) continue;
    const int bn = pass == 0 ? g->block_n : (pass == 1 ? 256 : BLOCK_M);
    const cuuint64_t Ktot = (cuuint64_t)taps * Cin;
    cuuint64_t dims[2] = {Ktot, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {Ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)bn};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(pass == 0 ? &g->map_b : (pass == 1 ? &g->map_b256 : &g->map_wswap), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)w_packed, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(g_tc_err, sizeof g_tc_err, "cuTensorMapEncodeTiled(B) failed: %d", (int)r); delete g; return nullptr; }
  }
  return g;
}

// Pair-folded conv (TcParams::fold): geometry of the PAIR grid (W / 2 columns), 128 = (wo, co) output rows, swapped kernel only.
TcGemm* tc_gemm_create_pfold(const bf16* in, int ld_in, const bf16* w_pfold, int Cin, int H, int W, int Bcap) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { snprintf(g_tc_err, sizeof g_tc_err, "cuTensorMapEncodeTiled entry point not available"); return nullptr; }
  if (Cin % BLOCK_K || ld_in % 8 || W % 2) { snprintf(g_tc_err, sizeof g_tc_err, "tc_gemm_create_pfold: unsupported shape Cin=%d ld=%d W=%d", Cin, ld_in, W); return nullptr; }
  const int Wp = W / 2;
  if (BLOCK_M % Wp) { snprintf(g_tc_err, sizeof g_tc_err, "tc_gemm_create_pfold: W/2=%d does not divide 128", Wp); return nullptr; }
  int Hb = BLOCK_M / Wp;
  if (Hb > H) Hb = H;
  if (H % Hb || BLOCK_M % (Hb * Wp)) { snprintf(g_tc_err, sizeof g_tc_err, "tc_gemm_create_pfold: H=%d W/2=%d not tileable", H, Wp); return nullptr; }
  const int Bt = BLOCK_M / (Hb * Wp);
  const int pps = H * Wp;
  if (Bcap % Bt || !(pps == 16 || (pps >= 32 && pps % 32 == 0 && (pps >= 256 ? pps % 256 == 0 : 256 % pps == 0)))) {
    snprintf(g_tc_err, sizeof g_tc_err, "tc_gemm_create_pfold: geometry %dx%d / Bcap %d does not fit 256-pair tiles", H, W, Bcap);
    return nullptr;
  }
  TcGemm* g = new TcGemm();
  memset(g, 0, sizeof(*g));
  g->Bcap = Bcap;
  g->block_n = 128;
  g->has256 = false;
  g->can_swap = true;
  TcParams& p = g->p;
  p.H = H; p.W = Wp; p.Hb = Hb; p.Bt = Bt; p.Cin = Cin; p.Cout = 128; p.taps = 9;
  p.kb_per_tap = Cin / BLOCK_K;
  p.fold = 1; p.fold_c1 = ld_in;
  {  // pair-row view of the channels-last activation: (2 pixels x channels, W / 2, H, B); the second pixel's channels start at ld_in
    cuuint64_t dims[4] = {(cuuint64_t)(ld_in + Cin), (cuuint64_t)Wp, (cuuint64_t)H, (cuuint64_t)Bcap};
    cuuint64_t strides[3] = {(cuuint64_t)2 * ld_in * 2, (cuuint64_t)W * ld_in * 2, (cuuint64_t)H * W * ld_in * 2};
    cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)Wp, (cuuint32_t)Hb, (cuuint32_t)Bt};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&g->map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(g_tc_err, sizeof g_tc_err, "cuTensorMapEncodeTiled(A, pair fold) failed: %d", (int)r); delete g; return nullptr; }
    g->has_a256 = encode_a256(enc, &g->map_a256, in, dims, strides, Wp, H, Hb, Bt, Bcap);
  }
  {
    const cuuint64_t Ktot = (cuuint64_t)12 * Cin;
    cuuint64_t dims[2] = {Ktot, 128};
    cuuint64_t strides[1] = {Ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)BLOCK_M};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&g->map_wswap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)w_pfold, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(g_tc_err, sizeof g_tc_err, "cuTensorMapEncodeTiled(W, pair fold) failed: %d", (int)r); delete g; return nullptr; }
  }
  return g;
}
bool tc_gemm_is_pfold(const TcGemm* g) { return g && g->p.fold != 0; }

void tc_gemm_destroy(TcGemm* g) { delete g; }

static bool swap_taken(const TcGemm* g, int m_tiles, int flags) {
  return g->can_swap && (flags & ~EPI_APPLY) == EPI_STATS && m_tiles % 2 == 0 && !(g_tc_dbg & 64);
}

// Split-K factor for a 3x3 conv launch of B samples: > 1 when the 128x256 tiles alone would leave most SMs idle
// (the K loop of a tile is a serial chain of ~150-cycle MMAs, so at small batch the deep levels are latency-bound).
int tc_gemm_split(const TcGemm* g, int B) {
  if (g_tc_dbg & 512) return 1;
  const TcParams& p = g->p;
  if (p.taps != 9) return 1;
  const int m_tiles = (int)(((long long)B * p.H * p.W) / BLOCK_M);
  const int bn = g->has256 ? 256 : g->block_n;
  const int tiles = m_tiles * (p.Cout / bn);
  const bool skip_dx = p.W == 1, skip_dy = p.H == 1;
  const int k_iters = (skip_dx ? 1 : 3) * (skip_dy ? 1 : 3) * p.kb_per_tap;
  if (tiles * 2 > num_sms() || k_iters < 8) return 1;
  int s = num_sms() / tiles;
  if (s > k_iters / 4) s = k_iters / 4;
  if (s > 8) s = 8;
  return s < 2 ? 1 : s;
}

// ---- cluster split-K + fused GroupNorm apply (conv_tc_cluster_kernel) ----
template <int BLOCK_N, int STAGES>
static cudaError_t launch_cluster_cfg(const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& p, int cs, cudaStream_t s, int* max_clusters) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(conv_tc_cluster_kernel<BLOCK_N, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<BLOCK_N, STAGES>());
    attr = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(p.m_tiles * cs));
  cfg.blockDim = dim3(CL_THREADS);
  cfg.dynamicSmemBytes = smem_bytes<BLOCK_N, STAGES>();
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_spdm_pdl ? 2 : 1;
  if (max_clusters) {  // occupancy query only
    cfg.numAttrs = 1;
    return cudaOccupancyMaxActiveClusters(max_clusters, conv_tc_cluster_kernel<BLOCK_N, STAGES>, &cfg);
  }
  return cudaLaunchKernelEx(&cfg, conv_tc_cluster_kernel<BLOCK_N, STAGES>, ma, mb, p);
}

// Clusters of `cs` CTAs (one CTA per SM: ~193 KB of shared memory each) that can be resident at once; cached per (N, cs).
static int max_active_clusters(const TcGemm* g, int bn, int cs) {
  static int cache[2][9] = {};
  int& c = cache[bn == 256][cs];
  if (c == 0) {
    TcParams p = g->p;
    p.m_tiles = 1;
    int n = 0;
    cudaError_t e = bn == 256 ? launch_cluster_cfg<256, 4>(g->map_a, g->map_b256, p, cs, nullptr, &n)
                              : launch_cluster_cfg<128, 6>(g->map_a, g->map_b, p, cs, nullptr, &n);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    c = n > 0 ? n : -1;
  }
  return c > 0 ? c : 0;
}

// Shape of the cluster launch for a deep-level conv of m_tiles M tiles: tile width bn, K slices ks (0 = not applicable).
//   * SPDM_CL_NOSPLIT_MAX = n (default 0 = off): K loops of <= n k-steps are NOT cut -- 128-wide N tiles, the cluster = the
//     Cout / 128 N tiles of an M tile and only exchanges the GroupNorm statistics.  Measured slower (0.926 vs 0.820 ms per step
//     with n = 24): a k-step costs ~830 cycles in this kernel, not the 512 of its four MMAs, so the longer chain outweighs the
//     saved reduction; kept as an A/B switch;
//   * otherwise 256-wide tiles where Cout allows and ks = 2..8 K slices, except that a 256-channel layer with few M tiles is
//     cut into two 128-wide N tiles so that a cluster of 8 shares the work (half the DSMEM bytes per CTA).
static void cluster_config(const TcGemm* g, int m_tiles, int* bn_out, int* ks_out) {
  *bn_out = 0; *ks_out = 0;
  const TcParams& p = g->p;
  const int bn0 = g->has256 ? 256 : g->block_n;
  if (bn0 != 128 && bn0 != 256) return;
  const int k_iters = (p.W == 1 ? 1 : 3) * (p.H == 1 ? 1 : 3) * p.kb_per_tap;
  static int wide = -1, nosplit_max = -1, verbose = -1;
  if (wide < 0) { const char* e = getenv("SPDM_CL_BN256"); wide = e ? atoi(e) : 0; }   // 1: always 256-wide tiles (A/B switch)
  if (nosplit_max < 0) { const char* e = getenv("SPDM_CL_NOSPLIT_MAX"); nosplit_max = e ? atoi(e) : 0; }
  if (verbose < 0) { const char* e = getenv("SPDM_VERBOSE"); verbose = e ? atoi(e) : 0; }
  if (k_iters <= nosplit_max && p.Cout % 128 == 0 && p.Cout / 128 <= 4 && m_tiles <= max_active_clusters(g, 128, p.Cout / 128)) {
    *bn_out = 128; *ks_out = 1;
    if (verbose) fprintf(stderr, "spdm cluster conv %dx%d %d->%d: m_tiles %d k_iters %d -> bn 128, no K split, cluster %d\n", p.H, p.W, p.Cin, p.Cout, m_tiles, k_iters, p.Cout / 128);
    return;
  }
  int bn = bn0;
  if (bn == 256 && p.Cout == 256 && !wide && m_tiles <= max_active_clusters(g, 128, 8)) bn = 128;
  const int n_tiles = p.Cout / bn;
  if (n_tiles > 2) return;
  for (int ks = 8 / n_tiles; ks >= 2; ks >>= 1) {
    if (ks * 2 > k_iters || bn / ks < 32) continue;
    const int mac = max_active_clusters(g, bn, ks * n_tiles);
    if (verbose) fprintf(stderr, "spdm cluster conv %dx%d %d->%d: m_tiles %d bn %d n_tiles %d k_iters %d ks %d: max active clusters %d\n", p.H, p.W, p.Cin, p.Cout, m_tiles, bn, n_tiles, k_iters, ks, mac);
    if (m_tiles > mac) continue;   // one wave of co-scheduled clusters
    *bn_out = bn; *ks_out = ks;
    return;
  }
}

int tc_gemm_cluster_split(const TcGemm* g, int B) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("SPDM_NO_CLUSTER"); off = e ? atoi(e) : 0; }
  if (off || tc_gemm_split(g, B) <= 1) return 0;   // only where the global-memory split-K path would be taken
  const TcParams& p = g->p;
  if (p.taps != 9 || p.H != p.Hb) return 0;
  const int rps = p.Hb * p.W;
  if (rps < 4 || rps > 32) return 0;
  int bn, ks;
  cluster_config(g, (int)(((long long)B * p.H * p.W) / BLOCK_M), &bn, &ks);
  return ks;
}

int tc_gemm_launch_cluster(const TcGemm* g, bf16* out, int ld_out, const ApplyArgs* ap, int B, int ks, cudaStream_t s) {
  TcParams p = g->p;
  p.ap = *ap;
  p.flags = EPI_APPLY;
  static int timing = -1;  // SPDM_CL_TIMING=1: block 0 prints its per-phase cycle counts (diagnostics)
  if (timing < 0) { const char* e = getenv("SPDM_CL_TIMING"); timing = e ? atoi(e) : 0; }
  p.dbg = timing ? 2048 : 0;
  p.out = out; p.ld_out = ld_out;
  p.m_tiles = (int)(((long long)B * p.H * p.W) / BLOCK_M);
  int bn, ks_cfg;
  cluster_config(g, p.m_tiles, &bn, &ks_cfg);
  if (bn == 0 || ks_cfg != ks) { snprintf(g_tc_err, sizeof g_tc_err, "cluster conv: inconsistent configuration"); return -1; }
  p.n_tiles = p.Cout / bn;
  p.cl_ks = ks;
  p.total_tiles = p.m_tiles * p.n_tiles * ks;
  const int cs = p.n_tiles * ks;
  cudaError_t e = bn == 256 ? launch_cluster_cfg<256, 4>(g->map_a, g->map_b256, p, cs, s, nullptr)
                            : launch_cluster_cfg<128, 6>(g->map_a, g->map_b, p, cs, s, nullptr);
  if (e != cudaSuccess) { snprintf(g_tc_err, sizeof g_tc_err, "cluster conv launch failed: %s", cudaGetErrorString(e)); return -1; }
  ++g_tc_launches;
  return 0;
}


// ---- chain of cluster convs (conv_chain_kernel) ----
struct TcChain {
  ChainParams p;
  ChainLayer* dev_layers;
  int cs, m_tiles, n_layers;
};

static cudaError_t chain_launch_cfg(const ChainParams& p, int grid, int cs, cudaStream_t s, int* max_clusters) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM);
    attr = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(CL_THREADS);
  cfg.dynamicSmemBytes = CH_SMEM;
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_spdm_pdl ? 2 : 1;
  if (max_clusters) {
    cfg.numAttrs = 1;
    return cudaOccupancyMaxActiveClusters(max_clusters, conv_chain_kernel, &cfg);
  }
  return cudaLaunchKernelEx(&cfg, conv_chain_kernel, p);
}

// Per-layer tile shape for a chain run by clusters of `cs`: N tile width bn and K slices ks with (Cout / bn) * ks == cs.
static bool chain_layer_shape(const TcGemm* g, int cs, int* bn_out, int* ks_out) {
  const TcParams& p = g->p;
  const int k_iters = (p.W == 1 ? 1 : 3) * (p.H == 1 ? 1 : 3) * p.kb_per_tap;
  // prefer the shape that moves the fewest bytes through DSMEM per CTA: (ks - 1) / ks * 128 * bn * 4
  int best_bn = 0, best_ks = 0;
  long long best_cost = 0;
  for (int bn = 128; bn <= 256; bn += 128) {
    if (p.Cout % bn) continue;
    const int n_tiles = p.Cout / bn;
    if (cs % n_tiles) continue;
    const int ks = cs / n_tiles;
    if (ks < 1 || ks > 8 || ks > k_iters || bn / ks < 32 || (bn / ks) % 32) continue;
    const long long cost = (long long)(ks - 1) * 128 * bn * 4 / ks + (long long)((k_iters + ks - 1) / ks) * 9000;   // DSMEM bytes + the ~9 KB a k-step (~530 cycles) is worth at ~17 B/clk
    if (!best_bn || cost < best_cost) { best_bn = bn; best_ks = ks; best_cost = cost; }
  }
  *bn_out = best_bn; *ks_out = best_ks;
  return best_bn != 0;
}

TcChain* tc_chain_create(const TcChainLayerDesc* layers, int n, int B, int pre_kind, const bf16* pre_in, int pre_ld_in, bf16* pre_out,
                         int pre_ld_out, int pre_C) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("SPDM_NO_CHAIN"); off = e ? atoi(e) : 0; }
  if (off || n < 1 || n > 16) return nullptr;
  const TcParams& p0 = layers[0].g->p;
  const int rps = p0.Hb * p0.W;
  if (p0.taps != 9 || p0.H != p0.Hb || rps < 4 || rps > 32) return nullptr;
  const int m_tiles = (int)(((long long)B * p0.H * p0.W) / BLOCK_M);
  if (m_tiles < 1 || (long long)m_tiles * BLOCK_M != (long long)B * p0.H * p0.W) return nullptr;
  for (int i = 0; i < n; ++i) {
    const TcParams& q = layers[i].g->p;
    if (q.taps != 9 || q.H != p0.H || q.W != p0.W || q.Hb != p0.Hb) return nullptr;
    if (tc_gemm_split(layers[i].g, B) <= 1) return nullptr;   // only where the tiles alone would leave most SMs idle
  }
  // largest cluster size whose clusters all run in one wave and that every layer can be cut for
  int cs = 0;
  ChainParams probe{};
  for (int c = 8; c >= 2; c >>= 1) {
    static int mac_cache[9] = {};
    if (mac_cache[c] == 0) {
      int nmax = 0;
      if (chain_launch_cfg(probe, c, c, nullptr, &nmax) != cudaSuccess) { cudaGetLastError(); nmax = 0; }
      mac_cache[c] = nmax > 0 ? nmax : -1;
    }
    if (mac_cache[c] < m_tiles) continue;
    bool ok = true;
    for (int i = 0; i < n && ok; ++i) { int bn, ks; ok = chain_layer_shape(layers[i].g, c, &bn, &ks); }
    if (ok) { cs = c; break; }
  }
  if (!cs) return nullptr;
  std::vector<ChainLayer> host((size_t)n);
  for (int i = 0; i < n; ++i) {
    const TcGemm* g = layers[i].g;
    ChainLayer& L = host[i];
    memset(&L, 0, sizeof L);
    int bn, ks;
    chain_layer_shape(g, cs, &bn, &ks);
    L.map_a = g->map_a;
    if (bn == 256) L.map_b = g->map_b256;
    else if (g->block_n == 128) L.map_b = g->map_b;
    else return nullptr;
    L.Cin = g->p.Cin; L.Cout = g->p.Cout; L.kb_per_tap = g->p.kb_per_tap;
    L.bn = bn; L.n_tiles = g->p.Cout / bn; L.ks = ks;
    L.out = layers[i].out; L.ld_out = layers[i].ld_out; L.ap = layers[i].ap;
    static int verbose = -1;
    if (verbose < 0) { const char* e = getenv("SPDM_VERBOSE"); verbose = e ? atoi(e) : 0; }
    if (verbose) fprintf(stderr, "spdm chain layer %d: %dx%d %d->%d m_tiles %d cluster %d: bn %d n_tiles %d ks %d\n", i, g->p.H, g->p.W, L.Cin, L.Cout, m_tiles, cs, bn, L.n_tiles, ks);
  }
  TcChain* c = new TcChain();
  memset(c, 0, sizeof *c);
  if (cudaMalloc(&c->dev_layers, sizeof(ChainLayer) * n) != cudaSuccess) { delete c; return nullptr; }
  if (cudaMemcpy(c->dev_layers, host.data(), sizeof(ChainLayer) * n, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(c->dev_layers); delete c; return nullptr; }
  c->cs = cs; c->m_tiles = m_tiles; c->n_layers = n;
  c->p.H = p0.H; c->p.W = p0.W; c->p.Bt = p0.Bt; c->p.n_layers = n; c->p.layers = c->dev_layers;
  {
    static int timing = -1;
    if (timing < 0) { const char* e = getenv("SPDM_CL_TIMING"); timing = e ? atoi(e) : 0; }
    c->p.dbg = timing ? 2048 : 0;
  }
  c->p.pre_kind = pre_kind; c->p.pre_in = pre_in; c->p.pre_ld_in = pre_ld_in; c->p.pre_out = pre_out; c->p.pre_ld_out = pre_ld_out; c->p.pre_C = pre_C;
  return c;
}

void tc_chain_destroy(TcChain* c) {
  if (!c) return;
  cudaFree(c->dev_layers);
  delete c;
}

int tc_chain_launch(const TcChain* c, cudaStream_t s, int step_off) {
  ChainParams p = c->p;
  p.step_off = step_off;
  cudaError_t e = chain_launch_cfg(p, c->m_tiles * c->cs, c->cs, s, nullptr);
  if (e != cudaSuccess) { snprintf(g_tc_err, sizeof g_tc_err, "chain launch failed: %s", cudaGetErrorString(e)); return -1; }
  ++g_tc_launches;
  return 0;
}

bool tc_gemm_can_fuse_apply(const TcGemm* g, int B) {
  if (g_tc_dbg & 256) return false;
  const TcParams& p = g->p;
  const int m_tiles = (int)(((long long)B * p.H * p.W) / BLOCK_M);
  const int pps = p.H * p.W;
  if (p.taps != 9) return false;
  if (swap_taken(g, m_tiles, EPI_STATS)) return pps <= 256 && 256 % pps == 0 && pps >= 16;
  const int bn = g->has256 ? 256 : g->block_n;
  return p.Cout == bn && p.H == p.Hb;  // one N tile, whole samples per M tile
}

// Does fusing the GroupNorm apply into this conv's epilogue pay?  Measured per layer at batch 256 and 4096 (event-timed eager
// steps, profiles/r02_fused_apply_per_layer.md):
//   * swapped-operand kernel, launch of at most two tiles per CTA: its 8-warp epilogue is exposed anyway (batch 256: -1 % of the step);
//   * swapped-operand kernel, persistent launch: the epilogue hides under the next tile's MMAs when the tile's K loop is long
//     enough (>= 18 k-steps) or the map is large (>= 256 pixels per sample, where the separate apply kernel is the expensive one):
//     batch 4096: 128->128 32x8 second conv 279 + 127 us -> 301 us, 256->256 16x4 254 + 80 -> 271; short K loops at 16x4 lose
//     (64->64: 41 + 23 -> 76 us) and stay unfused;
//   * plain kernel (Cout = 256 in one N tile): the 4x1 level gains (19 + 15 -> 25 us) and so do the long K loops of the 16x4 level
//     (256->256: 254 + 80 -> 271 us); the 8x2 level does not (56 + 23 -> 90 us).
bool tc_gemm_fuse_apply_pays(const TcGemm* g, int B) {
  if (!tc_gemm_can_fuse_apply(g, B)) return false;
  const TcParams& p = g->p;
  const int m_tiles = (int)(((long long)B * p.H * p.W) / BLOCK_M);
  static int big = -1;   // SPDM_FUSE_BIG=0: never fuse in persistent (multi-tile) launches (A/B switch: round-1 behaviour)
  if (big < 0) { const char* e = getenv("SPDM_FUSE_BIG"); big = e ? atoi(e) : 1; }
  const int k_iters0 = (p.W == 1 ? 1 : 3) * (p.H == 1 ? 1 : 3) * p.kb_per_tap;
  if (!swap_taken(g, m_tiles, EPI_STATS)) return big && (p.H * p.W <= 8 || (p.H * p.W >= 64 && k_iters0 >= 36));
  const int total = (m_tiles / 2) * ((p.Cout + BLOCK_M - 1) / BLOCK_M);
  if (total <= 2 * num_sms()) return true;
  if (!big) return false;
  // With 8 / 16 epilogue warps on the fused launches the short-K layers of the 16x4 level gain as well (batch 4096: 64->64 39 + 23 -> 54 us,
  // 64->128 50 + 44 -> 64 us; they lost with 4 warps: 76 us) -- every persistent swapped launch fuses.  SPDM_FUSE_SHORT=0: the old rule.
  static int fshort = -1;
  if (fshort < 0) { const char* e = getenv("SPDM_FUSE_SHORT"); fshort = e ? atoi(e) : 1; }
  if (fshort) return true;
  const int k_iters = p.fold ? 12 * p.kb_per_tap : (p.W == 1 ? 1 : 3) * (p.H == 1 ? 1 : 3) * p.kb_per_tap;
  const int pps_real = p.H * p.W * (p.fold ? 2 : 1);
  return k_iters >= 18 || pps_real >= 256;
}

int tc_gemm_launch(const TcGemm* g, bf16* out, int ld_out, float* stats, const float* bias, const bf16* resid, int ld_res, int flags,
                   int B, cudaStream_t s, bf16* vt, int vt_lk, const ApplyArgs* fuse, int ksplit, float* partial) {
  TcParams p = g->p;
  if (fuse) { p.ap = *fuse; flags |= EPI_APPLY; }
  p.ksplit = ksplit > 1 ? ksplit : 1;
  p.partial = partial;
  p.dbg = g_tc_dbg;
  p.vt = vt; p.vt_lk = vt_lk; p.vt_C = p.Cout / 3; p.vt_c0 = 2 * (p.Cout / 3);
  p.out = out; p.ld_out = ld_out; p.stats = stats; p.bias = bias; p.resid = resid; p.ld_res = ld_res; p.flags = flags;
  p.m_tiles = (int)(((long long)B * p.H * p.W) / BLOCK_M);
  if (p.fold) {   // pair fold: swapped kernel only; the output row of a pair is two real rows
    if (p.ksplit != 1 || !swap_taken(g, p.m_tiles, flags)) {
      snprintf(g_tc_err, sizeof g_tc_err, "pair-folded conv: only EPI_STATS launches over an even number of 128-pair tiles (B=%d)", B);
      return -1;
    }
    p.fold_ldo = ld_out;
    p.ld_out = 2 * ld_out;
  }
  if (p.ksplit == 1 && swap_taken(g, p.m_tiles, flags)) {
    p.n_tiles = (p.Cout + BLOCK_M - 1) / BLOCK_M;
    p.total_tiles = (p.m_tiles / 2) * p.n_tiles;
    const int pps = p.H * p.W;
    const int nw = p.Cout >= BLOCK_M ? 4 : p.Cout / 32;
    // 8 epilogue warps also for persistent launches that carry the fused GroupNorm apply: its two passes over the accumulator
    // (statistics, then normalise + GELU + store) are twice the plain epilogue's work; batch 4096: 15 550 -> 16 170 trajectories/s
    // (SPDM_FUSE_EPI8=0: A/B switch)
    static int epi8 = -1;
    if (epi8 < 0) { const char* e = getenv("SPDM_FUSE_EPI8"); epi8 = e ? atoi(e) : 1; }
    static int epi16 = -1;   // SPDM_FUSE_EPI16=0: at most 8 epilogue warps (A/B switch)
    if (epi16 < 0) { const char* e = getenv("SPDM_FUSE_EPI16"); epi16 = e ? atoi(e) : 1; }
    // 8 epilogue warps only where the epilogue is exposed (<= 2 tiles per CTA); 16 for the fused apply of persistent launches, whose
    // two passes over the accumulator outlast the tile's MMAs once the operand traffic is cut (halo mode)
    const int threads = (fuse && epi16 && epi8 && p.total_tiles > 2 * num_sms() && pps >= 64) ? SWAP_THREADS16
                        : ((p.total_tiles <= 2 * num_sms() || (epi8 && fuse)) ? SWAP_THREADS : NUM_THREADS);
    const int halves = (threads == SWAP_THREADS && pps >= 256) ? 2 : 1;   // then a sample of >= 256 pixels gets one partial per column half
    p.P = (pps > 256 ? pps / 256 : 1) * p.n_tiles * nw * halves;
    constexpr int STG = 4;
    constexpr int smem = STG * (2 * A_STAGE_BYTES + BLOCK_M * BLOCK_K * 2) + 1024;
    static int m64_mode = -1;  // SPDM_M64: 0 = off, 1 = on (TMEM layout A), 2 = on (layout B)
    if (m64_mode < 0) { const char* e = getenv("SPDM_M64"); m64_mode = e ? atoi(e) : 1; }
    const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
    // halo mode: persistent launches only (the L2 -> SM operand traffic is what bounds them: profiles/r02 summary); SPDM_NO_HALO=1 switches it off
    static int no_halo = -1;
    if (no_halo < 0) { const char* e = getenv("SPDM_NO_HALO"); no_halo = e ? atoi(e) : 0; }
    p.halo = (g->has_halo && !p.fold && !no_halo && p.total_tiles > num_sms()) ? 1 : 0;
    const CUtensorMap& map_h = g->has_halo ? g->map_halo : g->map_a;
    static int no256 = -1;   // SPDM_NO_PIX256=1: two 128-row pixel boxes per k-step as before (A/B switch)
    if (no256 < 0) { const char* e = getenv("SPDM_NO_PIX256"); no256 = e ? atoi(e) : 0; }
    p.pix256 = (g->has_a256 && !no256) ? 1 : 0;
    const CUtensorMap& map_px = p.pix256 ? g->map_a256 : g->map_a;
    if (p.Cout == 64 && !fuse && m64_mode > 0) {  // 64-row MMA: no zero rows, half the operand-read time per instruction
      const int nw64 = m64_mode == 2 ? 2 : 4;
      p.P = (pps > 256 ? pps / 256 : 1) * p.n_tiles * nw64 * halves;
      if (m64_mode == 2) p.dbg |= 1024;
      static bool attr64 = false;
      if (!attr64) { cudaFuncSetAttribute(conv_tc_swap_kernel<STG, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr64 = true; }
      launch_pdl(conv_tc_swap_kernel<STG, 64>, dim3(grid), dim3(threads), smem, s, map_px, g->map_b, map_h, p);
      ++g_tc_launches;
      return p.P;
    }
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(conv_tc_swap_kernel<STG, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
    launch_pdl(conv_tc_swap_kernel<STG, 128>, dim3(grid), dim3(threads), smem, s, map_px, g->map_wswap, map_h, p);
    ++g_tc_launches;
    return p.P;
  }
  // the MMA issue cost does not depend on N (the 128-row A operand read dominates): the widest N tile wins
  const bool use256 = g->has256 && !(g_tc_dbg & 128);
  const int bn = use256 ? 256 : g->block_n;
  p.n_tiles = p.Cout / bn;
  p.total_tiles = p.m_tiles * p.n_tiles * p.ksplit;
  p.P = partials_for(p, p.n_tiles);
  // SPDM_PAIR=1: persistent launches of the 256-wide tiles as CTA pairs (cta_group::2).  Correct (parity tests pass with it) but
  // measured slower than the single-CTA form at batch 4096 (256->256 at 16x4: 250 vs 234 us; the whole step -0.5 %): the operand
  // waits it was meant to shorten do not come from the stage count or the bytes per k-step (profiles/r02_conv_pipeline_probe.md).
  static int pair = -1;
  if (pair < 0) { const char* e = getenv("SPDM_PAIR"); pair = e ? atoi(e) : 0; }
  if (bn == 256 && pair && g->block_n == 128 && p.ksplit == 1 && p.m_tiles % 2 == 0 && p.total_tiles >= 2 * num_sms() &&
      !(flags & (EPI_VT | EPI_RESID | EPI_MASK)) && launch_pair(g->map_a, g->map_b, p, s))
    return p.P;
  if (bn == 256) launch_cfg<256, 4>(g->map_a, g->map_b256, p, s);
  else if (bn == 192) launch_cfg<192, 4>(g->map_a, g->map_b, p, s);
  else if (bn == 128) launch_cfg<128, 6>(g->map_a, g->map_b, p, s);
  else launch_cfg<64, 8>(g->map_a, g->map_b, p, s);
  return p.P;
}
