// data_kernels.cu — dataset -> batch on the device, sm_100a.
//
// The reference builds every training / inference batch on the host: CarRacingDataset.__getitem__ slices a strided window
// out of the in-memory arrays, normalises the positions of the window (min-max, centred on its first point, halved) and
// the DataLoader stacks B such items (utils/load_data.py:128-144, utils/data_utils.py:18-21,58-62) -- at batch 512 that is
// 566 MB of fp32 frames per step through pageable host memory.  Here the whole dataset stays resident in HBM (frames as the
// uint8 HWC the simulator wrote, 27 KB each instead of 110 KB) and one launch pair gathers a batch:
//   gather_images_kernel   frames [start + t*step] of every window -> (B, T_img, 3, H, W) fp32, uint8 -> float (/255) and
//                          HWC -> CHW in flight; HBM-bound: 1 B read + 4 B written per pixel-channel
//   gather_state_kernel    position (window-normalised), velocity / action (dataset-normalised), translation vector
//   unnormalize_position_kernel   utils/data_utils.py:35-40, for sampled trajectories / histories
// Every fp32 operation is a single correctly-rounded IEEE operation in the reference's order (no FMA contraction), so the
// results are bit-identical to the numpy float32 arithmetic of the reference.
#include <stdint.h>

#include "../../include/spdm.h"
#include "common.cuh"

int spdm_data_fail(const char* msg);  // plan.cu: sets spdm_last_error, returns -1
void spdm_count_data_launch();

namespace {

__device__ __forceinline__ float norm_rn(float x, float mn, float mx) {
  // (x - min) / (max - min) * 2 - 1, one rounding per operation (utils/data_utils.py:18-21)
  return __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(x, mn), __fsub_rn(mx, mn)), 2.0f), 1.0f);
}

// one thread = 4 consecutive pixels of one image row, all 3 channels
__global__ void __launch_bounds__(256) gather_images_u8_kernel(const uint8_t* __restrict__ img, const long long* __restrict__ starts,
                                                                int T_img, int step, int H, int W, long long n_frames_out,
                                                                float* __restrict__ out) {
  const int w4 = W >> 2;
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_frame = (long long)H * w4;
  if (v >= n_frames_out * per_frame) return;
  const long long f = v / per_frame;
  const int rem = (int)(v - f * per_frame);
  const int y = rem / w4, x = (rem - y * w4) * 4;
  const long long b = f / T_img;
  const int t = (int)(f - b * T_img);
  const long long src_frame = starts ? starts[b] + (long long)t * step : f;   // starts == null: plain decode, frame f -> frame f
  const uint32_t* src = reinterpret_cast<const uint32_t*>(img + ((src_frame * H + y) * W + x) * 3);   // 12 bytes, 4-byte aligned
  const uint32_t w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
  const uint32_t bytes[12] = {w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u, w0 >> 24, w1 & 255u, (w1 >> 8) & 255u,
                              (w1 >> 16) & 255u, w1 >> 24, w2 & 255u, (w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24};
  const long long plane = (long long)H * W;
  float* dst = out + f * 3 * plane + (long long)y * W + x;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float4 o;
    o.x = __fdiv_rn((float)bytes[c], 255.0f);
    o.y = __fdiv_rn((float)bytes[3 + c], 255.0f);
    o.z = __fdiv_rn((float)bytes[6 + c], 255.0f);
    o.w = __fdiv_rn((float)bytes[9 + c], 255.0f);
    __stcs(reinterpret_cast<float4*>(dst + c * plane), o);   // streamed: the batch is consumed once by the encoder
  }
}

// frames already float CHW (what the reference keeps in host memory): a strided gather copy, 16 bytes per thread
__global__ void __launch_bounds__(256) gather_images_f32_kernel(const float* __restrict__ img, const long long* __restrict__ starts,
                                                                 int T_img, int step, long long frame_elems, long long n_frames_out,
                                                                 float* __restrict__ out) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_frame = frame_elems >> 2;
  if (v >= n_frames_out * per_frame) return;
  const long long f = v / per_frame, e = (v - f * per_frame) * 4;
  const long long b = f / T_img;
  const int t = (int)(f - b * T_img);
  const long long src_frame = starts[b] + (long long)t * step;
  __stcs(reinterpret_cast<float4*>(out + f * frame_elems + e), __ldg(reinterpret_cast<const float4*>(img + src_frame * frame_elems + e)));
}

__global__ void gather_state_kernel(const float* __restrict__ position, const float* __restrict__ velocity,
                                    const float* __restrict__ action, const long long* __restrict__ starts, int B, int T, int step,
                                    spdm_data_stats st, float* __restrict__ out_pos, float* __restrict__ out_vel,
                                    float* __restrict__ out_act, float* __restrict__ out_tr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * T) return;
  const int b = i / T, t = i - b * T;
  const long long s0 = starts[b], src = s0 + (long long)t * step;
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const float first = norm_rn(__ldg(position + s0 * 2 + d), st.pos_min, st.pos_max);        // translation (load_data.py:130)
    const float pn = norm_rn(__ldg(position + src * 2 + d), st.pos_min, st.pos_max);
    out_pos[(size_t)i * 2 + d] = __fdiv_rn(__fsub_rn(pn, first), 2.0f);
    if (t == 0) out_tr[b * 2 + d] = first;
    out_vel[(size_t)i * 2 + d] = norm_rn(__ldg(velocity + src * 2 + d), st.vel_min[d], st.vel_max[d]);
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) out_act[(size_t)i * 3 + d] = norm_rn(__ldg(action + src * 3 + d), st.act_min[d], st.act_max[d]);
}

__global__ void unnormalize_position_kernel(const float* __restrict__ npos, const float* __restrict__ tr, float pmin, float pmax,
                                            long long n_rows, int rows_per_sample, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one (row, coordinate)
  if (i >= n_rows * 2) return;
  const long long row = i >> 1;
  const int d = (int)(i & 1);
  const long long b = row / rows_per_sample;
  // (x * 2 + translation + 1) / 2 * (max - min) + min   (utils/data_utils.py:23-26,35-40)
  const float v = __fadd_rn(__fmul_rn(npos[i], 2.0f), tr[b * 2 + d]);
  out[i] = __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn(v, 1.0f), 2.0f), __fsub_rn(pmax, pmin)), pmin);
}

}  // namespace

// uint8 HWC frames -> fp32 CHW (x / 255), frame by frame: the decode half of gather_images_u8_kernel without the window gather
// (spdm_encode_cond_u8 on an fp32 plan)
void launch_decode_u8_hwc(const uint8_t* img, float* out, long long frames, int H, int W, cudaStream_t s) {
  const long long n = frames * H * (W / 4);
  gather_images_u8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(img, nullptr, 1, 1, H, W, frames, out);
  spdm_count_data_launch();
}

extern "C" int spdm_gather_windows(const void* images, int32_t image_kind, int32_t H, int32_t W, const float* position,
                                   const float* velocity, const float* action, const int64_t* starts, int32_t B, int32_t T, int32_t T_img,
                                   int32_t step, const spdm_data_stats* stats, float* out_image, float* out_position, float* out_velocity,
                                   float* out_action, float* out_translation, void* stream) {
  if (!position || !velocity || !action || !starts || !stats || !out_position || !out_velocity || !out_action || !out_translation)
    return spdm_data_fail("spdm_gather_windows: null argument");
  if (B <= 0 || T <= 0 || step <= 0 || T_img < 0 || T_img > T) return spdm_data_fail("spdm_gather_windows: bad B / T / T_img / step");
  if (T_img > 0 && (!images || !out_image)) return spdm_data_fail("spdm_gather_windows: images requested but a pointer is null");
  if (T_img > 0 && (W % 4 || H <= 0)) return spdm_data_fail("spdm_gather_windows: image width must be a multiple of 4");
  if (image_kind != SPDM_IMAGES_U8_HWC && image_kind != SPDM_IMAGES_F32_CHW) return spdm_data_fail("spdm_gather_windows: unknown image_kind");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long* st = reinterpret_cast<const long long*>(starts);
  if (T_img > 0) {
    const long long frames = (long long)B * T_img;
    if (image_kind == SPDM_IMAGES_U8_HWC) {
      const long long n = frames * H * (W / 4);
      gather_images_u8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(reinterpret_cast<const uint8_t*>(images), st, T_img, step, H, W,
                                                                         frames, out_image);
    } else {
      const long long fe = 3LL * H * W, n = frames * (fe / 4);
      gather_images_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(reinterpret_cast<const float*>(images), st, T_img, step, fe,
                                                                          frames, out_image);
    }
    spdm_count_data_launch();
  }
  gather_state_kernel<<<(B * T + 127) / 128, 128, 0, s>>>(position, velocity, action, st, B, T, step, *stats, out_position, out_velocity,
                                                          out_action, out_translation);
  spdm_count_data_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return spdm_data_fail(cudaGetErrorString(e));
  return 0;
}

extern "C" int spdm_unnormalize_position(const float* npos, const float* translation, float pos_min, float pos_max, int64_t n_samples,
                                         int32_t rows_per_sample, float* out, void* stream) {
  if (!npos || !translation || !out || n_samples <= 0 || rows_per_sample <= 0) return spdm_data_fail("spdm_unnormalize_position: bad argument");
  const long long n = n_samples * rows_per_sample * 2;
  unnormalize_position_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      npos, translation, pos_min, pos_max, n_samples * rows_per_sample, rows_per_sample, out);
  spdm_count_data_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return spdm_data_fail(cudaGetErrorString(e));
  return 0;
}
