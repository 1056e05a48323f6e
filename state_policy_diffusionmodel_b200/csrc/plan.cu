// plan.cu — the C ABI of libspdm.so (include/spdm.h): plan, weights, U-Net forward, sampling loop.
//
// A plan owns the repacked weights, the channels-last activation workspace and the CUDA graphs of
// one (variant, precision, geometry, batch_max) configuration on one device.  All tensor arguments
// of the ABI are caller-owned device memory; every call only enqueues work on the caller's stream.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "../../include/spdm.h"
#include "common.cuh"
#include "train.cuh"

// -------------------------------------------------------------------------------------------------
// errors
// -------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return -1;
}
struct SpdmError { std::string msg; };
#define CUDA_OK(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess) {                                                                            \
      char _b[512];                                                                                     \
      snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      throw SpdmError{_b};                                                                              \
    }                                                                                                   \
  } while (0)
#define REQUIRE(cond, ...)                              \
  do {                                                  \
    if (!(cond)) {                                      \
      char _b[512];                                     \
      snprintf(_b, sizeof _b, __VA_ARGS__);             \
      throw SpdmError{_b};                              \
    }                                                   \
  } while (0)

extern "C" const char* spdm_last_error(void) { return g_err; }
// data_kernels.cu (plan-free entry points) reports through the same thread-local message
int spdm_data_fail(const char* msg) { return fail("%s", msg); }
static long long g_data_launches = 0;
void spdm_count_data_launch() { ++g_data_launches; }
extern "C" int64_t spdm_data_launch_count(void) { return g_data_launches; }
extern "C" const char* spdm_version(void) { return "spdm-b200 0.1 (sm_100a)"; }

// -------------------------------------------------------------------------------------------------
// plan
// -------------------------------------------------------------------------------------------------
namespace {

struct GemmW {  // a 3x3 convolution or a Linear layer
  int Cin = 0, Cout = 0, taps = 1;
  float* w32 = nullptr;  // [taps][Cin][Cout]   (fp32 path)
  bf16* w16 = nullptr;   // [Cout][taps*Cin]    (bf16 tcgen05 path)
  bf16* w16_fold = nullptr;  // 3x3 convs of the W = 2 level, inference: [2 Cout][9][2 Cin], the two pixels of a row folded into channels
  float* wtf = nullptr;      // TF32 plan: 3x3 conv weights [Cout][9][Cin] fp32, K-major B operand of tcgen05.mma.kind::tf32
  bf16* w16_pfold = nullptr; // 3x3 convs with 64 output channels, inference: pair fold [128][3][4][Cin] (kernels.cu::pack_conv_pfold_bf16_kernel)
  float* bias = nullptr; // [Cout] or null
  GemmW* twin = nullptr; // training: the data-gradient GEMM (Cin/Cout exchanged, transposed / tap-flipped weights)
};
struct NormW { float* g = nullptr; float* b = nullptr; int C = 0; };
struct StageInfo { const char* name; int cin, cout; int temb_off, film_off; };

static const StageInfo kStages[6] = {
    {"down1", 64, 128, 0, 0},      {"down2", 128, 256, 128, 256}, {"down3", 256, 256, 384, 768},
    {"up1", 512, 128, 640, 1280},  {"up2", 256, 64, 768, 1536},   {"up3", 128, 64, 832, 1664}};

}  // namespace

struct spdm_plan {
  spdm_config cfg;
  bool attention = true, bf16_mode = false, sched_only = false;
  bool tf32_mode = false;              // SPDM_PRECISION_TF32: fp32 activations, 3x3 convs on tcgen05.mma.kind::tf32 (conv_tf32.cu)
  std::map<std::string, TfGemm*> tf_cache;
  bool enc_resnet = false;             // SPDM_FLAG_ENCODER_RESNET18: ResNet18-GroupNorm vision encoder (resnet.inl), 512 features per frame
  void *rn_col = nullptr, *rn_a0 = nullptr, *rn_buf[4] = {};
  float* rn_stats = nullptr;
  std::map<std::string, long long> rn_rows_cap;
  int feat_dim() const { return enc_resnet ? 512 : 128; }
  bool simple = false;                 // SPDM_VARIANT_SIMPLE_UNET: models/simple_Unet.py UNet on the fp32 path (simple_unet.inl)
  float* su_table = nullptr; int su_table_rows = 0;   // its PositionalEncoding buffer [max_len][time_dim]
  int H0 = 0, W0 = 0, lh = 0, lw = 0;  // padded geometry (pad_to 8) and low-side pads
  int Bcap = 0, bm = 1;
  int G = 0;  // global_cond_dim
  std::vector<void*> allocs;
  size_t bytes = 0;
  long long launches = 0;

  // weights
  std::map<std::string, GemmW> gemms;
  std::map<std::string, NormW> norms;
  float* w_in = nullptr;               // inc.first [9][64]
  float* w_outc = nullptr; float* b_outc = nullptr;
  float* temb_w = nullptr; float* temb_b = nullptr;  // [256][896], [896]
  float* film_w = nullptr; float* film_b = nullptr;  // [G][1792], [1792]
  float* inv_freq = nullptr;                          // [time_dim/2]
  float *enc_w1 = nullptr, *enc_b1 = nullptr, *enc_w2 = nullptr, *enc_b2 = nullptr, *enc_w3 = nullptr, *enc_b3 = nullptr;
  float *enc_wl = nullptr, *enc_bl = nullptr;  // [9216][128], [128]
  bf16* enc_wl16 = nullptr;                    // bf16 plan: [128][9216] K-major for the tcgen05 GEMM
  bf16* enc_feat16 = nullptr; bf16* enc_out16 = nullptr; TcGemm* enc_tc = nullptr;
  // bf16 plan: conv2 / conv3 of the encoder as patch GEMMs (layouts: bwd_kernels.cu, "Vision encoder on the tensor cores")
  bf16 *enc_w2p = nullptr, *enc_w2pT = nullptr, *enc_w3p = nullptr, *enc_w3pT = nullptr;
  float* enc_b2p = nullptr;
  bf16 *enc_c1p = nullptr, *enc_c2 = nullptr; TcGemm *enc_tc2 = nullptr, *enc_tc3 = nullptr;
  bool enc_simt_infer = false;  // SPDM_ENC_SIMT_INFER=1: fused CUDA-core conv stack for inference (A/B switch)
  std::map<std::string, std::function<void(const float*, const int64_t*, int, cudaStream_t)>> loaders;
  std::set<std::string> missing_unet, missing_enc;

  // workspace (T = float or bf16, chosen by precision)
  void *raw[4] = {}, *hbuf[4] = {}, *abuf[4] = {}, *bbuf[4] = {}, *cat[3] = {};
  void *a_ln[4] = {}, *a_qkv[4] = {}, *a_att[4] = {}, *a_res[4] = {}, *a_ff[4] = {}, *a_vt[4] = {};
  std::map<std::string, SdpaTc*> sdpa_cache;
  std::map<std::string, AttnTail*> tail_cache;
  std::map<std::string, AttnHead*> head_cache;
  bool no_head_tail = false;           // SPDM_NO_HEAD_TAIL=1: C = 64 blocks keep head and tail as two launches (A/B switch)
  bool no_head = false;                // SPDM_NO_ATTN_HEAD=1: LayerNorm, in_proj and the attention core as three launches (A/B switch)
  float* stats = nullptr;   // [Bcap][SPDM_MAX_PARTIALS][2]
  float* film = nullptr;    // [Bcap][1792]
  float* cond = nullptr;    // [Bcap][G]
  float* cond_mish = nullptr;
  float* temb_call = nullptr;   // [Bcap][896]  (unet_forward with explicit t)
  float* temb_table = nullptr;  // [K][896]
  float* coef = nullptr;        // [K][8]
  long long* timesteps = nullptr;
  int K = 0, sched_kind = 0;
  bool temb_table_dirty = true;
  bool have_cond = false;
  float* xt = nullptr; float* eps = nullptr;  // [Bcap][n]
  StepDyn* dyn = nullptr;
  // pinned staging of the per-call parameter block: a small ring, so that a call never has to wait for the previous call's
  // upload (a slot is reused DYN_SLOTS calls later; its event has long completed by then)
  enum { DYN_SLOTS = 8 };
  StepDyn* dyn_host = nullptr; cudaEvent_t dyn_ev[DYN_SLOTS] = {}; bool dyn_used[DYN_SLOTS] = {}; unsigned dyn_next = 0;
  StepDyn* dyn_slot() {
    const unsigned j = dyn_next++ % DYN_SLOTS;
    if (dyn_used[j]) cudaEventSynchronize(dyn_ev[j]);
    dyn_used[j] = true;
    dyn_cur = j;
    return dyn_host + j;
  }
  unsigned dyn_cur = 0;
  cudaStream_t own_stream = nullptr;  // graphs are captured and replayed here (the caller's stream may be the legacy stream)
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  int split = 1;                       // sub-batches run concurrently per denoising step
  std::map<int, float*> partial;       // split-K fp32 partial tiles, one buffer per concurrent lane (keyed by b0)
  std::map<int, size_t> partial_cap;
  bool no_splitk = false;              // SPDM_NO_SPLITK=1 (A/B switch)
  bool no_fold = false;                // SPDM_NO_FOLD=1: W = 2 convs as ordinary 9-tap implicit GEMMs (A/B switch)
  bool no_pfold = false;               // SPDM_NO_PFOLD=1: 64-channel convs on 64-row MMAs instead of the pair fold (A/B switch)
  bool no_fuse = false;                // SPDM_NO_FUSE_APPLY=1: keep GroupNorm apply as a separate kernel (A/B switch)
  std::vector<char> skip;              // SPDM_SKIP_IDX=i,j,...: launches of one forward (in timed() order) that are NOT issued --
  int timed_idx = 0;                   // timing ablation only (tools/ablate.py), results are garbage
  int fuse_mode = -1;                  // SPDM_FUSE_MODE: -1 auto (default: GroupNorm apply inside the swapped conv's 8-warp epilogue where the launch has
                                       // at most two tiles per CTA; batch 256: 0.7256 -> 0.7201 ms per step), 0 never, 1 / 2 wherever possible (slower)
  cudaStream_t lane_stream[7] = {};
  cudaEvent_t ev_fork = nullptr, ev_lane[7] = {};
  float* enc_feat = nullptr; int enc_chunk = 0;  // [enc_chunk][9216]
  float* enc_out = nullptr;                       // [Bcap*T][128]
  float* enc_u8_stage = nullptr; int enc_u8_cap = 0;  // decoded frames of spdm_encode_cond_u8 on plans whose encoder reads fp32

  std::map<std::string, TcGemm*> tc_cache;
  std::map<std::string, TcChain*> chain_cache;   // runs of deep-level convs as one launch (null = tried, not applicable)
  // graphs keyed by batch: [0] = graph_steps-step body, [1] = 1-step body
  struct GraphSet { cudaGraphExec_t multi = nullptr, single = nullptr; long long n_multi = 0, n_single = 0; };
  std::map<long long, GraphSet> graphs;

  // eager profiling (spdm_profile_step): CUDA events around every launch, summed per kernel class
  struct ProfRec { int cat; cudaEvent_t e0, e1; double flops, bytes; };
  bool prof_on = false;
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;

  struct TrainState* tr = nullptr;  // training step state (train_impl.inl), null until spdm_train_enable

  // debug tap
  std::string tap_name; float* tap_out = nullptr; long long tap_count = -1;

  template <typename T> T* alloc(size_t n) {
    void* p = nullptr;
    CUDA_OK(cudaMalloc(&p, n * sizeof(T)));
    CUDA_OK(cudaMemset(p, 0, n * sizeof(T)));
    allocs.push_back(p);
    bytes += n * sizeof(T);
    return reinterpret_cast<T*>(p);
  }
  int n_elems() const { return cfg.rows * cfg.dim; }
  int levelH(int l) const { return H0 >> l; }
  int levelW(int l) const { return W0 >> l; }
};

namespace {

long long total_launches() { return kernels_launch_count() + tc_launch_count() + bwd_launch_count() + wgrad_tc_launch_count_value; }

float* train_film_wT(spdm_plan* p);  // training-only transposed weight copies (null when training is not enabled)
void train_film_weight_loaded(spdm_plan* p, const float* src, int off, int C2, cudaStream_t s);
float* train_enc_wlT(spdm_plan* p);
void train_destroy(spdm_plan* p);

// ---- weight registration -------------------------------------------------------------------------
void check_shape(const std::string& name, const int64_t* shape, int ndim, std::initializer_list<int64_t> want) {
  bool ok = (int)want.size() == ndim;
  int i = 0;
  if (ok) for (auto w : want) ok = ok && shape[i++] == w;
  if (!ok) {
    std::string got, exp;
    for (int k = 0; k < ndim; ++k) got += std::to_string(shape[k]) + (k + 1 < ndim ? "," : "");
    for (auto w : want) exp += std::to_string(w) + ",";
    throw SpdmError{"weight " + name + ": shape (" + got + ") != expected (" + exp + ")"};
  }
}

void reg_vec(spdm_plan* p, const std::string& name, float* dst, int n, std::set<std::string>& missing) {
  missing.insert(name);
  p->loaders[name] = [=](const float* src, const int64_t* shape, int ndim, cudaStream_t s) {
    long long cnt = 1;
    for (int i = 0; i < ndim; ++i) cnt *= shape[i];
    REQUIRE(cnt == n, "weight %s: %lld elements, expected %d", name.c_str(), cnt, n);
    CUDA_OK(cudaMemcpyAsync(dst, src, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, s));
  };
}

GemmW& reg_conv3(spdm_plan* p, const std::string& name, int Cin, int Cout) {
  GemmW& g = p->gemms[name];
  g.Cin = Cin; g.Cout = Cout; g.taps = 9;
  if (p->bf16_mode) g.w16 = p->alloc<bf16>((size_t)9 * Cin * Cout);
  else g.w32 = p->alloc<float>((size_t)9 * Cin * Cout);
  if (p->tf32_mode && Cin % 32 == 0 && Cout % 64 == 0) g.wtf = p->alloc<float>((size_t)9 * Cin * Cout);
  // the third level of the U-Net is H/4 x 2: a third of the MMAs of a tile-wise 9-tap implicit GEMM there multiply the zero
  // padding left and right of the two columns.  Folding the column into the channels (K = 3 x 2 Cin, N = 2 Cout over (b, h) rows)
  // makes the conv dense along W (Fwd::fold_ok)
  const bool level2 = name.rfind("down2.", 0) == 0 || name.rfind("up1.", 0) == 0;
  if (p->bf16_mode && level2 && p->W0 == 8 && Cin % 32 == 0 && Cout % 32 == 0) g.w16_fold = p->alloc<bf16>((size_t)2 * Cout * 9 * 2 * Cin);
  // 64 output channels: a 64-row tcgen05.mma costs what a 128-row one does, so pairs of pixels are folded into the rows (Fwd::pfold_tc)
  if (p->bf16_mode && Cout == 64 && Cin % 64 == 0 && !p->no_pfold) g.w16_pfold = p->alloc<bf16>((size_t)128 * 12 * Cin);
  p->missing_unet.insert(name + ".weight");
  GemmW* gp = &g;
  bool bfm = p->bf16_mode;
  p->loaders[name + ".weight"] = [=](const float* src, const int64_t* shape, int ndim, cudaStream_t s) {
    check_shape(name + ".weight", shape, ndim, {Cout, Cin, 3, 3});
    bool twin_done = false;
    if (bfm) {
      twin_done = gp->twin && launch_pack_conv3_bf16(src, gp->w16, gp->twin->w16, Cout, Cin, s);   // forward and dgrad operands in one pass
      if (!twin_done) launch_pack_conv_bf16(src, gp->w16, Cout, Cin, 3, s);
    } else {
      launch_pack_conv_f32(src, gp->w32, Cout, Cin, 3, s);
    }
    if (gp->wtf) launch_pack_conv_tf32(src, gp->wtf, Cout, Cin, s);
    if (gp->w16_fold && !p->tr) launch_pack_conv_fold2_bf16(src, gp->w16_fold, Cout, Cin, s);  // inference only (Fwd::fold_ok)
    if (gp->w16_pfold && !p->tr) launch_pack_conv_pfold_bf16(src, gp->w16_pfold, Cin, s);        // inference only (Fwd::pfold_tc)
    if (gp->twin && !twin_done) {
      if (bfm) launch_pack_conv_dgrad_bf16(src, gp->twin->w16, Cout, Cin, s);
      else launch_pack_conv_dgrad_f32(src, gp->twin->w32, Cout, Cin, s);
    }
  };
  return g;
}

GemmW& reg_linear(spdm_plan* p, const std::string& wname, const std::string& bname, int K, int N) {
  GemmW& g = p->gemms[wname];
  g.Cin = K; g.Cout = N; g.taps = 1;
  if (p->bf16_mode) g.w16 = p->alloc<bf16>((size_t)K * N);
  else g.w32 = p->alloc<float>((size_t)K * N);
  g.bias = p->alloc<float>(N);
  p->missing_unet.insert(wname);
  GemmW* gp = &g;
  bool bfm = p->bf16_mode;
  p->loaders[wname] = [=](const float* src, const int64_t* shape, int ndim, cudaStream_t s) {
    check_shape(wname, shape, ndim, {N, K});
    if (bfm) launch_cast_bf16(src, gp->w16, (long long)K * N, s);  // (N,K) row-major is already the K-major B operand
    else launch_pack_linear_f32(src, gp->w32, N, K, N, 0, s);
    if (gp->twin) {  // d x = d y W: [Cin_g = N][Cout_g = K] fp32 is the PyTorch layout itself; bf16 wants it K-major = [K][N]
      if (bfm) launch_pack_linear_dgrad_bf16(src, gp->twin->w16, N, K, s);
      else CUDA_OK(cudaMemcpyAsync(gp->twin->w32, src, (size_t)N * K * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
  };
  reg_vec(p, bname, g.bias, N, p->missing_unet);
  return g;
}

NormW& reg_norm(spdm_plan* p, const std::string& name, int C) {
  NormW& n = p->norms[name];
  n.C = C;
  n.g = p->alloc<float>(C);
  n.b = p->alloc<float>(C);
  reg_vec(p, name + ".weight", n.g, C, p->missing_unet);
  reg_vec(p, name + ".bias", n.b, C, p->missing_unet);
  return n;
}

void reg_double_conv(spdm_plan* p, const std::string& name, int Cin, int Cout, bool first_is_input_layer = false) {
  if (!first_is_input_layer) reg_conv3(p, name + ".first", Cin, Cout);
  reg_conv3(p, name + ".second", Cout, Cout);
  reg_norm(p, name + ".norm", Cout);
}

void register_weights(spdm_plan* p) {
  const int TD = p->cfg.time_dim;
  // inc
  p->w_in = p->alloc<float>(9 * 64);
  p->missing_unet.insert("inc.first.weight");
  {
    float* dst = p->w_in;
    p->loaders["inc.first.weight"] = [=](const float* src, const int64_t* shape, int ndim, cudaStream_t s) {
      check_shape("inc.first.weight", shape, ndim, {64, 1, 3, 3});
      launch_pack_conv_f32(src, dst, 64, 1, 3, s);
    };
  }
  reg_double_conv(p, "inc", 1, 64, true);
  p->temb_w = p->alloc<float>((size_t)TD * SPDM_TEMB_WIDTH);
  p->temb_b = p->alloc<float>(SPDM_TEMB_WIDTH);
  if (p->G > 0) {
    p->film_w = p->alloc<float>((size_t)p->G * SPDM_FILM_WIDTH);
    p->film_b = p->alloc<float>(SPDM_FILM_WIDTH);
  }
  for (const StageInfo& st : kStages) {
    const std::string n = st.name;
    reg_double_conv(p, n + ".doubleConv1", st.cin, st.cin);
    reg_double_conv(p, n + ".doubleConv2", st.cin, st.cout);
    {
      const std::string wn = n + ".emb_layer.1.weight";
      p->missing_unet.insert(wn);
      float* dst = p->temb_w;
      const int C = st.cout, off = st.temb_off;
      p->loaders[wn] = [=](const float* src, const int64_t* shape, int ndim, cudaStream_t s) {
        check_shape(wn, shape, ndim, {C, TD});
        launch_pack_linear_f32(src, dst, C, TD, SPDM_TEMB_WIDTH, off, s);
      };
      reg_vec(p, n + ".emb_layer.1.bias", p->temb_b + off, C, p->missing_unet);
    }
    if (p->G > 0) {
      const std::string wn = n + ".cond_encoder.2.weight";
      p->missing_unet.insert(wn);
      float* dst = p->film_w;
      const int C2 = 2 * st.cout, off = st.film_off, G = p->G;
      p->loaders[wn] = [=](const float* src, const int64_t* shape, int ndim, cudaStream_t s) {
        check_shape(wn, shape, ndim, {C2, G});
        launch_pack_linear_f32(src, dst, C2, G, SPDM_FILM_WIDTH, off, s);
        train_film_weight_loaded(p, src, off, C2, s);  // training-only copies (transposed fp32 / bf16 tensor-core operands)
      };
      reg_vec(p, n + ".cond_encoder.2.bias", p->film_b + off, C2, p->missing_unet);
    }
  }
  reg_double_conv(p, "bot1", 256, 512);
  reg_double_conv(p, "bot2", 512, 512);
  reg_double_conv(p, "bot3", 512, 256);
  if (p->attention) {
    const struct { const char* n; int C; } sas[6] = {{"sa1", 128}, {"sa2", 256}, {"sa3", 256}, {"sa4", 128}, {"sa5", 64}, {"sa6", 64}};
    for (auto& sa : sas) {
      const std::string n = sa.n;
      const int C = sa.C;
      reg_linear(p, n + ".attention.in_proj_weight", n + ".attention.in_proj_bias", C, 3 * C);
      reg_linear(p, n + ".attention.out_proj.weight", n + ".attention.out_proj.bias", C, C);
      reg_norm(p, n + ".ln", C);
      reg_norm(p, n + ".ff_self.0", C);
      reg_linear(p, n + ".ff_self.1.weight", n + ".ff_self.1.bias", C, C);
      reg_linear(p, n + ".ff_self.3.weight", n + ".ff_self.3.bias", C, C);
    }
  }
  p->w_outc = p->alloc<float>(64);
  p->b_outc = p->alloc<float>(1);
  reg_vec(p, "outc.weight", p->w_outc, 64, p->missing_unet);
  reg_vec(p, "outc.bias", p->b_outc, 1, p->missing_unet);

  // sinusoidal frequencies (models/Unet_FiLmLayer.py:267-270); overridable with the exact torch values
  {
    std::vector<float> f(TD / 2);
    for (int i = 0; i < TD / 2; ++i) {
      const float e = (float)(2 * i) / (float)TD;
      f[i] = 1.0f / powf(10000.0f, e);
    }
    p->inv_freq = p->alloc<float>(TD / 2);
    CUDA_OK(cudaMemcpy(p->inv_freq, f.data(), f.size() * sizeof(float), cudaMemcpyHostToDevice));
    std::set<std::string> dummy;
    reg_vec(p, "pos_encoding.inv_freq", p->inv_freq, TD / 2, dummy);
  }
  if (p->enc_resnet) return;   // its weights are registered by register_weights_resnet (resnet.inl)
  // vision encoder (models/encoder/autoencoder.py:11-20)
  p->enc_w1 = p->alloc<float>(16 * 3 * 4);   p->enc_b1 = p->alloc<float>(16);
  p->enc_w2 = p->alloc<float>(32 * 16 * 4);  p->enc_b2 = p->alloc<float>(32);
  p->enc_w3 = p->alloc<float>(64 * 32 * 4);  p->enc_b3 = p->alloc<float>(64);
  p->enc_wl = p->alloc<float>((size_t)9216 * 128);  p->enc_bl = p->alloc<float>(128);
  reg_vec(p, "vision_encoder.0.weight", p->enc_w1, 16 * 3 * 4, p->missing_enc);
  reg_vec(p, "vision_encoder.0.bias", p->enc_b1, 16, p->missing_enc);
  if (p->bf16_mode) {
    p->enc_w2p = p->alloc<bf16>(64 * 128); p->enc_w2pT = p->alloc<bf16>(128 * 64);
    p->enc_w3p = p->alloc<bf16>(64 * 128); p->enc_w3pT = p->alloc<bf16>(128 * 64);
    p->enc_b2p = p->alloc<float>(64);
  }
  p->missing_enc.insert("vision_encoder.2.weight");
  {
    float* dst = p->enc_w2;  // stored transposed: [16*4][32]
    p->loaders["vision_encoder.2.weight"] = [=](const float* src, const int64_t* shape, int ndim, cudaStream_t s) {
      check_shape("vision_encoder.2.weight", shape, ndim, {32, 16, 2, 2});
      launch_pack_linear_f32(src, dst, 32, 64, 32, 0, s);
      if (p->enc_w2p) launch_enc_pack_w2(src, p->enc_w2p, p->enc_w2pT, s);
    };
  }
  p->missing_enc.insert("vision_encoder.2.bias");
  {
    float* dst = p->enc_b2;
    p->loaders["vision_encoder.2.bias"] = [=](const float* src, const int64_t* shape, int ndim, cudaStream_t s) {
      check_shape("vision_encoder.2.bias", shape, ndim, {32});
      CUDA_OK(cudaMemcpyAsync(dst, src, 32 * sizeof(float), cudaMemcpyDeviceToDevice, s));
      if (p->enc_b2p) launch_enc_pack_b2(src, p->enc_b2p, s);
    };
  }
  p->missing_enc.insert("vision_encoder.4.weight");
  {
    float* dst = p->enc_w3;  // stored transposed: [32*4][64]
    p->loaders["vision_encoder.4.weight"] = [=](const float* src, const int64_t* shape, int ndim, cudaStream_t s) {
      check_shape("vision_encoder.4.weight", shape, ndim, {64, 32, 2, 2});
      launch_pack_linear_f32(src, dst, 64, 128, 64, 0, s);
      if (p->enc_w3p) launch_enc_pack_w3(src, p->enc_w3p, p->enc_w3pT, s);
    };
  }
  reg_vec(p, "vision_encoder.4.bias", p->enc_b3, 64, p->missing_enc);
  p->missing_enc.insert("vision_encoder.7.weight");
  {
    float* dst = p->enc_wl;
    if (p->bf16_mode) p->enc_wl16 = p->alloc<bf16>((size_t)9216 * 128);
    bf16* dst16 = p->enc_wl16;
    p->loaders["vision_encoder.7.weight"] = [=](const float* src, const int64_t* shape, int ndim, cudaStream_t s) {
      check_shape("vision_encoder.7.weight", shape, ndim, {128, 9216});
      launch_pack_enc_linear(src, dst, s);
      if (dst16) launch_pack_enc_linear_bf16(src, dst16, s);
      if (float* wt = train_enc_wlT(p)) launch_pack_enc_linear_t(src, wt, s);
    };
  }
  reg_vec(p, "vision_encoder.7.bias", p->enc_bl, 128, p->missing_enc);
}

// ---- workspace -----------------------------------------------------------------------------------
template <typename T> void alloc_workspace(spdm_plan* p) {
  static const int cmax[4] = {128, 256, 512, 512};
  static const int catt[4] = {64, 128, 256, 256};
  const size_t Bc = p->Bcap;
  for (int l = 0; l < 4; ++l) {
    const size_t hw = (size_t)p->levelH(l) * p->levelW(l);
    p->raw[l] = p->alloc<T>(Bc * hw * cmax[l]);
    p->hbuf[l] = p->alloc<T>(Bc * hw * cmax[l]);
    p->abuf[l] = p->alloc<T>(Bc * hw * cmax[l]);
    p->bbuf[l] = p->alloc<T>(Bc * hw * cmax[l]);
    if (l < 3) p->cat[l] = p->alloc<T>(Bc * hw * cmax[l]);
    if (p->attention) {
      p->a_ln[l] = p->alloc<T>(Bc * hw * catt[l]);
      p->a_qkv[l] = p->alloc<T>(Bc * hw * catt[l] * 3);
      p->a_att[l] = p->alloc<T>(Bc * hw * catt[l]);
      p->a_res[l] = p->alloc<T>(Bc * hw * catt[l]);
      p->a_ff[l] = p->alloc<T>(Bc * hw * catt[l]);
      if (sizeof(T) == 2) p->a_vt[l] = p->alloc<T>(Bc * hw * catt[l]);
    }
  }
}

// ---- forward -------------------------------------------------------------------------------------
struct FwdCtx {
  const float* x;        // (B,1,rows,dim) fp32
  float* out;            // (B,1,rows,dim) fp32
  const float* temb;     // rows of SPDM_TEMB_WIDTH
  int temb_mode;
  const int* step_ptr;
  int step_off;          // this step's position inside a captured multi-step graph (ApplyArgs::step_off)
  const float* film;     // [B][1792] or null
  int B;
  const StepArgs* fuse_step;  // non-null: replace outc by outc + posterior update (the sampling loop)
  int b0;                // first sample of this sub-batch inside the plan's workspace (lanes run concurrently)
  cudaStream_t s;
};

enum : int { PC_CONV3 = 0, PC_GEMM1, PC_APPLY, PC_STATS, PC_RESAMPLE, PC_LN, PC_SDPA, PC_IO, PC_STEP, PC_CONV3_GN, PC_N };

template <typename F> void timed(spdm_plan* p, cudaStream_t s, int cat, double flops, double bytes, F&& f) {
  if (!p->skip.empty()) {
    const int idx = p->timed_idx++;
    if (idx < (int)p->skip.size() && p->skip[idx]) return;
  }
  if (!p->prof_on) { f(); return; }
  spdm_plan::ProfRec r{cat, nullptr, nullptr, flops, bytes};
  while (p->ev_pool.size() < p->ev_used + 2) {  // events are pooled: creation is far slower than a small kernel
    cudaEvent_t e;
    CUDA_OK(cudaEventCreate(&e));
    p->ev_pool.push_back(e);
  }
  r.e0 = p->ev_pool[p->ev_used++];
  r.e1 = p->ev_pool[p->ev_used++];
  CUDA_OK(cudaEventRecord(r.e0, s));
  f();
  CUDA_OK(cudaEventRecord(r.e1, s));
  p->prof.push_back(r);
}

template <typename T> struct Fwd {
  spdm_plan* p;
  FwdCtx c;
  int curP = 1;
  int Bpad;
  bf16* vt = nullptr;  // set around the in_proj GEMM of an attention block that feeds sdpa_tc (EPI_VT)
  int vt_lk = 0;
  float* stats_ov = nullptr;  // training: every conv keeps its own GroupNorm partial sums for the backward pass

  Fwd(spdm_plan* p_, const FwdCtx& c_) : p(p_), c(c_) { Bpad = ((c.B + p->bm - 1) / p->bm) * p->bm; }

  // workspace of this sub-batch: every buffer is [Bcap][hw(level)][capacity] -> offset by b0 whole samples
  T* act(void* v, int level) const {
    static const int cmax[4] = {128, 256, 512, 512};
    return reinterpret_cast<T*>(v) + (size_t)c.b0 * p->levelH(level) * p->levelW(level) * cmax[level];
  }
  T* att(void* v, int level, int mult = 1) const {
    static const int catt[4] = {64, 128, 256, 256};
    return reinterpret_cast<T*>(v) + (size_t)c.b0 * p->levelH(level) * p->levelW(level) * catt[level] * mult;
  }
  float* stats() const { return stats_ov ? stats_ov : p->stats + (size_t)c.b0 * SPDM_MAX_PARTIALS * 2; }

  void tap(const std::string& name, const T* ptr, int ld, int C, int level) {
    if (p->tap_out && p->tap_name == name) {
      const int hw = p->levelH(level) * p->levelW(level);
      launch_to_nchw<T>(ptr, ld, p->tap_out, c.B, hw, C, c.s);
      p->tap_count = (long long)c.B * hw * C;
    }
  }

  TcGemm* get_tc(const std::string& wname, GemmW& g, const T* in, int ld_in, int level) {
    char key[160];
    snprintf(key, sizeof key, "%s|%p|%d", wname.c_str(), (const void*)in, ld_in);
    TcGemm*& tc = p->tc_cache[key];
    if (!tc) {
      tc = tc_gemm_create(reinterpret_cast<const bf16*>(in), ld_in, g.w16, g.Cin, g.Cout, g.taps, p->levelH(level), p->levelW(level), p->Bcap);
      REQUIRE(tc != nullptr, "%s: %s", wname.c_str(), tc_last_error());
    }
    return tc;
  }

  // W = 2 level, inference: run the conv in its folded form (GemmW::w16_fold) -- same memory for input and output, viewed as
  // [B*H rows][2*C]; needs both to be dense in the channel dimension
  // Only at batches where the conv's tiles fill the machine anyway: folding halves the number of M tiles, so where the K loop is
  // split to occupy the SMs (small batch: split-K / cluster path) it would trade parallelism for work.
  bool fold_ok(const std::string& wname, GemmW& g, const T* in, int ld_in, int level, int ld_out) {
    if constexpr (sizeof(T) == 2) {
      if (!(g.taps == 9 && g.w16_fold && !p->tr && !p->no_fold && p->fuse_mode <= 0 && p->levelW(level) == 2 && ld_in == g.Cin &&
            ld_out == g.Cout)) return false;
      return p->no_splitk || tc_gemm_split(get_tc(wname, g, in, ld_in, level), Bpad) <= 1;
    }
    return false;
  }
  TcGemm* get_tc_fold(const std::string& wname, GemmW& g, const T* in, int level) {
    char key[160];
    snprintf(key, sizeof key, "%s#fold|%p", wname.c_str(), (const void*)in);
    TcGemm*& tc = p->tc_cache[key];
    if (!tc) {
      tc = tc_gemm_create(reinterpret_cast<const bf16*>(in), 2 * g.Cin, g.w16_fold, 2 * g.Cin, 2 * g.Cout, 9, p->levelH(level), 1, p->Bcap);
      REQUIRE(tc != nullptr, "%s (folded): %s", wname.c_str(), tc_last_error());
    }
    return tc;
  }

  // 64 output channels, inference: the pair-folded form of the conv (null where it does not apply at this geometry / batch)
  TcGemm* pfold_tc(const std::string& wname, GemmW& g, const T* in, int ld_in, int level) {
    if constexpr (sizeof(T) == 2) {
      if (!g.w16_pfold || p->tr || g.taps != 9) return nullptr;
      const int H = p->levelH(level), W = p->levelW(level);
      if (W % 2 || ((long long)Bpad * H * (W / 2)) % 256) return nullptr;
      char key[160];
      snprintf(key, sizeof key, "%s#pfold|%p|%d", wname.c_str(), (const void*)in, ld_in);
      auto it = p->tc_cache.find(key);
      if (it == p->tc_cache.end())
        it = p->tc_cache.emplace(key, tc_gemm_create_pfold(reinterpret_cast<const bf16*>(in), ld_in, g.w16_pfold, g.Cin, H, W, p->Bcap)).first;
      return it->second;
    }
    return nullptr;
  }

  // split-K factor for this conv at the current batch (1 = none); sizes the lane's partial buffer on first use
  int split_for(const std::string& wname, const T* in, int ld_in, int level, int Cout) {
    if constexpr (sizeof(T) == 2) {
      if (p->no_splitk) return 1;
      const int n = p->levelH(level) * p->levelW(level) * Cout;
      if (n > 16384 || n % 1024) return 1;
      GemmW& g = p->gemms[wname];
      const int S = tc_gemm_split(get_tc(wname, g, in, ld_in, level), Bpad);   // (a folded conv is only used where this is 1)
      if (S > 1) {
        const size_t need = (size_t)S * Bpad * n;
        float*& buf = p->partial[c.b0];
        size_t& cap = p->partial_cap[c.b0];
        if (cap < need) { buf = p->alloc<float>(need); cap = need; }  // first (eager) step only; graphs reuse it
      }
      return S;
    }
    return 1;
  }

  // Deep-level conv at small batch: K slices of a tile as one thread-block cluster, reduced through distributed shared
  // memory with GroupNorm apply behind it (conv_tc_cluster_kernel).  Returns false when not applicable here.
  bool gemm_cluster(const std::string& wname, const T* in, int ld_in, int level, T* out, int ld_out, const ApplyArgs& a) {
    if constexpr (sizeof(T) == 2) {
      if (p->no_splitk) return false;
      GemmW& g = p->gemms[wname];
      TcGemm* tc = get_tc(wname, g, in, ld_in, level);
      const int ks = tc_gemm_cluster_split(tc, Bpad);
      if (ks < 1) return false;
      const int H = p->levelH(level), W = p->levelW(level);
      const double flops = 2.0 * g.Cin * g.Cout * (3.0 * H - 2) * (3.0 * W - 2) * c.B;
      const double bytes = ((double)c.B * H * W * (g.Cin + g.Cout) + (double)g.taps * g.Cin * g.Cout) * 2.0;
      timed(p, c.s, PC_CONV3_GN, flops, bytes, [&] {
        const int rc = tc_gemm_launch_cluster(tc, reinterpret_cast<bf16*>(out), ld_out, &a, Bpad, ks, c.s);
        REQUIRE(rc == 0, "%s: %s", wname.c_str(), tc_last_error());
      });
      return true;
    }
    return false;
  }

  // can GroupNorm apply run inside this conv's epilogue (whole samples and all channels in one tile)?
  // `min_mode`: the SPDM_FUSE_MODE value from which this conv is fused unconditionally (1: GELU-free second conv, 2: first conv)
  bool can_fuse(const std::string& wname, const T* in, int ld_in, int level, int min_mode) {
    if constexpr (sizeof(T) == 2) {
      if (p->no_fuse || p->fuse_mode == 0 || p->tr) return false;
      GemmW& g = p->gemms[wname];
      TcGemm* pf = pfold_tc(wname, g, in, ld_in, level);
      TcGemm* tc = pf ? pf : get_tc(wname, g, in, ld_in, level);
      if (p->fuse_mode < 0) return tc_gemm_fuse_apply_pays(tc, Bpad);
      return p->fuse_mode >= min_mode && tc_gemm_can_fuse_apply(tc, Bpad);
    }
    return false;
  }

  // conv / linear: in [M,Cin] (ld_in) -> out [M,Cout] (ld_out)
  void gemm(const std::string& wname, const T* in, int ld_in, int level, T* out, int ld_out, int flags, const T* resid = nullptr,
            int ld_res = 0, const ApplyArgs* fuse = nullptr, int ksplit = 1, float* partial = nullptr) {
    auto it = p->gemms.find(wname);
    REQUIRE(it != p->gemms.end(), "internal: unknown gemm %s", wname.c_str());
    GemmW& g = it->second;
    const int H = p->levelH(level), W = p->levelW(level);
    if constexpr (sizeof(T) == 2) {
      const bool fold = !fuse && !resid && ksplit <= 1 && (flags & ~EPI_STATS) == 0 && fold_ok(wname, g, in, ld_in, level, ld_out);
      TcGemm* pf = (!resid && ksplit <= 1 && (flags & ~EPI_STATS) == 0 && (flags & EPI_STATS)) ? pfold_tc(wname, g, in, ld_in, level) : nullptr;
      TcGemm* tc = pf ? pf : (fold ? get_tc_fold(wname, g, in, level) : get_tc(wname, g, in, ld_in, level));
      if (fold) ld_out *= 2;
      // algorithmic work: taps that fall inside the image only, real batch rows only
      const double flops = g.taps == 9 ? 2.0 * g.Cin * g.Cout * (3.0 * H - 2) * (3.0 * W - 2) * c.B
                                       : 2.0 * g.Cin * g.Cout * (double)H * W * c.B;
      const double bytes = ((double)c.B * H * W * (g.Cin + g.Cout) + (double)g.taps * g.Cin * g.Cout) * 2.0;
      int P = 1;
      timed(p, c.s, g.taps == 9 ? PC_CONV3 : PC_GEMM1, flops, bytes, [&] {
        P = tc_gemm_launch(tc, reinterpret_cast<bf16*>(out), ld_out, stats(), (flags & EPI_BIAS) ? g.bias : nullptr,
                           reinterpret_cast<const bf16*>(resid), ld_res, flags, Bpad, c.s, vt, vt_lk, fuse, ksplit, partial);
        REQUIRE(P >= 0, "%s: %s", wname.c_str(), tc_last_error());
      });
      REQUIRE(fuse || !(flags & EPI_STATS) || P <= SPDM_MAX_PARTIALS, "%s: too many GroupNorm partials (%d)", wname.c_str(), P);
      curP = P;
    } else {
      if (p->tf32_mode && g.taps == 9 && g.wtf && !resid && (flags & ~EPI_STATS) == 0) {
        // TF32 precision mode: the 3x3 convs run on the tensor cores (tcgen05.mma.kind::tf32), fp32 in / fp32 out
        char key[160];
        snprintf(key, sizeof key, "%s|%p|%d", wname.c_str(), (const void*)in, ld_in);
        TfGemm*& tf = p->tf_cache[key];
        if (!tf) {
          tf = tf32_conv_create(reinterpret_cast<const float*>(in), ld_in, g.wtf, g.Cin, g.Cout, H, W, p->Bcap);
          REQUIRE(tf != nullptr, "%s: %s", wname.c_str(), tf32_last_error());
        }
        const double flops = 2.0 * g.Cin * g.Cout * (3.0 * H - 2) * (3.0 * W - 2) * c.B;
        const double bytes = ((double)c.B * H * W * (g.Cin + g.Cout) + 9.0 * g.Cin * g.Cout) * 4.0;
        timed(p, c.s, PC_CONV3, flops, bytes, [&] { tf32_conv_launch(tf, reinterpret_cast<float*>(out), ld_out, Bpad, c.s); });
        if (flags & EPI_STATS) {
          timed(p, c.s, PC_STATS, 0, (double)c.B * H * W * g.Cout * 4.0,
                [&] { launch_stats<T>(out, stats(), c.B, H * W, g.Cout, ld_out, c.s); });
          curP = 1;
        }
        return;
      }
      GemmSimtArgs a{};
      a.in = in; a.w = g.w32; a.bias = (flags & EPI_BIAS) ? g.bias : nullptr; a.resid = (flags & EPI_RESID) ? resid : nullptr;
      a.out = out; a.M = c.B * H * W; a.Cin = g.Cin; a.Cout = g.Cout; a.ld_in = ld_in; a.ld_out = ld_out; a.ld_res = ld_res;
      a.H = H; a.W = W; a.taps = g.taps; a.act = (flags & EPI_GELU) ? ACT_GELU : ACT_NONE;
      const double flops = g.taps == 9 ? 2.0 * g.Cin * g.Cout * (3.0 * H - 2) * (3.0 * W - 2) * c.B
                                       : 2.0 * g.Cin * g.Cout * (double)H * W * c.B;
      const double bytes = ((double)c.B * H * W * (g.Cin + g.Cout) + (double)g.taps * g.Cin * g.Cout) * 4.0;
      timed(p, c.s, g.taps == 9 ? PC_CONV3 : PC_GEMM1, flops, bytes, [&] { launch_gemm_simt<float, float>(a, c.s); });
      if (flags & EPI_STATS) {
        timed(p, c.s, PC_STATS, 0, (double)c.B * H * W * g.Cout * 4.0,
              [&] { launch_stats<T>(out, stats(), c.B, H * W, g.Cout, ld_out, c.s); });
        curP = 1;
      }
    }
  }

  ApplyArgs make_apply(const std::string& norm, int C, int level, int act, const StageInfo* st) {
    NormW& n = p->norms[norm];
    ApplyArgs a{};
    a.stats = stats(); a.P = curP; a.gamma = n.g; a.beta = n.b;
    a.temb = nullptr; a.temb_mode = TEMB_NONE; a.film = nullptr;
    if (st) {
      a.temb = c.temb; a.temb_mode = c.temb_mode; a.temb_off = st->temb_off; a.step_ptr = c.step_ptr; a.step_off = c.step_off;
      if (c.film) { a.film = c.film; a.film_off = st->film_off; }
    }
    a.HW = p->levelH(level) * p->levelW(level); a.C = C; a.act = act; a.eps = 1e-5f;
    return a;
  }

  void apply(const std::string& norm, const T* raw, int ld_in, int C, int level, T* out, int ld_out, int act, const StageInfo* st) {
    ApplyArgs a = make_apply(norm, C, level, act, st);
    a.raw = raw; a.out = out; a.ld_in = ld_in; a.ld_out = ld_out;
    timed(p, c.s, PC_APPLY, 0, 2.0 * c.B * a.HW * C * sizeof(T), [&] { launch_apply<T, T>(a, c.B, c.s); });
  }

  // DoubleConvolution (models/Unet_FiLmLayer.py:85-115); `first_done`: raw already holds conv1's output.
  // On the tensor-core path GroupNorm apply runs inside the conv epilogue whenever a tile holds whole samples.
  // first_done: 1 = raw already holds conv1's output (+ statistics), 2 = h already holds GELU(GN(conv1))
  void double_conv(const std::string& name, const T* in, int ld_in, int Cout, int level, T* out, int ld_out, const StageInfo* st,
                   int first_done = 0) {
    T* raw = act(p->raw[level], level);
    T* h = act(p->hbuf[level], level);
    const bool tap1 = p->tap_out && p->tap_name == name + ".first", tap2 = p->tap_out && p->tap_name == name + ".second";
    bool fused = first_done == 2;
    const long long Mpad = (long long)Bpad * p->levelH(level) * p->levelW(level);
    if (!first_done) {
      const int S = tap1 ? 1 : split_for(name + ".first", in, ld_in, level, Cout);
      if (S > 1 && gemm_cluster(name + ".first", in, ld_in, level, h, Cout, make_apply(name + ".norm", Cout, level, ACT_GELU, nullptr))) {
        fused = true;
      } else if (S > 1) {
        gemm(name + ".first", in, ld_in, level, raw, Cout, EPI_STATS, nullptr, 0, nullptr, S, p->partial[c.b0]);
        ApplyArgs a = make_apply(name + ".norm", Cout, level, ACT_GELU, nullptr);
        a.out = h; a.ld_out = Cout;
        timed(p, c.s, PC_APPLY, 0, (4.0 * S + 2.0) * c.B * a.HW * Cout, [&] { launch_apply_partial(a, p->partial[c.b0], S, Mpad, c.B, c.s); });
        fused = true;
      } else if (!tap1 && can_fuse(name + ".first", in, ld_in, level, 2)) {
        const ApplyArgs a = make_apply(name + ".norm", Cout, level, ACT_GELU, nullptr);
        gemm(name + ".first", in, ld_in, level, h, Cout, EPI_STATS, nullptr, 0, &a);
        fused = true;
      } else {
        gemm(name + ".first", in, ld_in, level, raw, Cout, EPI_STATS);
      }
    }
    if (!fused) {
      tap(name + ".first", raw, Cout, Cout, level);
      apply(name + ".norm", raw, Cout, Cout, level, h, Cout, ACT_GELU, nullptr);
    }
    const int S2 = tap2 ? 1 : split_for(name + ".second", h, Cout, level, Cout);
    if (S2 > 1 && gemm_cluster(name + ".second", h, Cout, level, out, ld_out, make_apply(name + ".norm", Cout, level, ACT_NONE, st))) {
    } else if (S2 > 1) {
      gemm(name + ".second", h, Cout, level, raw, Cout, EPI_STATS, nullptr, 0, nullptr, S2, p->partial[c.b0]);
      ApplyArgs a = make_apply(name + ".norm", Cout, level, ACT_NONE, st);
      a.out = out; a.ld_out = ld_out;
      timed(p, c.s, PC_APPLY, 0, (4.0 * S2 + 2.0) * c.B * a.HW * Cout, [&] { launch_apply_partial(a, p->partial[c.b0], S2, Mpad, c.B, c.s); });
    } else if (!tap2 && can_fuse(name + ".second", h, Cout, level, 1)) {
      const ApplyArgs a = make_apply(name + ".norm", Cout, level, ACT_NONE, st);
      gemm(name + ".second", h, Cout, level, out, ld_out, EPI_STATS, nullptr, 0, &a);
    } else {
      gemm(name + ".second", h, Cout, level, raw, Cout, EPI_STATS);
      tap(name + ".second", raw, Cout, Cout, level);
      apply(name + ".norm", raw, Cout, Cout, level, out, ld_out, ACT_NONE, st);
    }
  }

  // A run of DoubleConvolutions of one deep level as ONE launch (conv_chain_kernel): `dcs` = {name, Cin, Cout, stage-or-null, out} in
  // order; the first conv reads `in` (ld_in), the map between the two convs of a DoubleConvolution lives in hbuf, every
  // DoubleConvolution writes its own `out`.  pre_kind 1 / 2: the MaxPool2d(2) / bilinear upsample that produces `in` (channels
  // [0, pre_C)) from pre_src.  Returns false when the run is not a small-batch cluster case: the caller then issues the layers
  // one by one.
  struct ChainDC { std::string name; int Cin, Cout; const StageInfo* st; T* out; int ld_out; };
  bool chain(const std::string& key0, const std::vector<ChainDC>& dcs, const T* in, int ld_in, int level, int pre_kind, const T* pre_src,
             int pre_ld_src, int pre_C) {
    if constexpr (sizeof(T) == 2) {
      if (p->no_splitk || p->tr || p->tap_out || !p->skip.empty()) return false;
      char key[256];
      snprintf(key, sizeof key, "%s|%d|%d|%p|%d|%p|%p", key0.c_str(), c.b0, Bpad, (const void*)c.temb, c.temb_mode, (const void*)c.film,
               (const void*)c.step_ptr);
      auto it = p->chain_cache.find(key);
      const int H = p->levelH(level), W = p->levelW(level);
      double flops = 0.0, bytes = 0.0;
      if (it == p->chain_cache.end()) {
        std::vector<TcChainLayerDesc> descs;
        T* h = act(p->hbuf[level], level);
        const T* cur = in;
        int cur_ld = ld_in;
        for (const ChainDC& d : dcs) {
          GemmW& g1 = p->gemms[d.name + ".first"];
          GemmW& g2 = p->gemms[d.name + ".second"];
          TcChainLayerDesc a{}, b{};
          a.g = get_tc(d.name + ".first", g1, cur, cur_ld, level);
          a.out = reinterpret_cast<bf16*>(h); a.ld_out = d.Cout;
          curP = 1;
          a.ap = make_apply(d.name + ".norm", d.Cout, level, ACT_GELU, nullptr);
          b.g = get_tc(d.name + ".second", g2, h, d.Cout, level);
          b.out = reinterpret_cast<bf16*>(d.out); b.ld_out = d.ld_out;
          b.ap = make_apply(d.name + ".norm", d.Cout, level, ACT_NONE, d.st);
          descs.push_back(a);
          descs.push_back(b);
          cur = d.out; cur_ld = d.ld_out;
        }
        TcChain* ch = tc_chain_create(descs.data(), (int)descs.size(), Bpad, pre_kind, reinterpret_cast<const bf16*>(pre_src), pre_ld_src,
                                      reinterpret_cast<bf16*>(const_cast<T*>(in)), ld_in, pre_C);
        it = p->chain_cache.emplace(key, ch).first;
      }
      if (!it->second) return false;
      for (const ChainDC& d : dcs) {
        flops += 2.0 * (d.Cin + d.Cout) * d.Cout * (3.0 * H - 2) * (3.0 * W - 2) * c.B;
        bytes += ((double)c.B * H * W * (d.Cin + 3.0 * d.Cout) + 9.0 * (d.Cin + d.Cout) * d.Cout) * 2.0;
      }
      timed(p, c.s, PC_CONV3_GN, flops, bytes, [&] {
        const int rc = tc_chain_launch(it->second, c.s, c.step_off);
        REQUIRE(rc == 0, "%s: %s", key0.c_str(), tc_last_error());
      });
      return true;
    }
    return false;
  }

  // SelfAttention (models/Unet_FiLmLayer.py:44-82)
  void self_attention(const std::string& name, const T* x, int ld_x, int C, int level, T* out, int ld_out) {
    const int L = p->levelH(level) * p->levelW(level);
    const long long M = (long long)c.B * L;
    T* ln = att(p->a_ln[level], level); T* qkv = att(p->a_qkv[level], level, 3); T* attn = att(p->a_att[level], level);
    T* res = att(p->a_res[level], level); T* ff = att(p->a_ff[level], level);
    NormW& n1 = p->norms[name + ".ln"];
    NormW& n2 = p->norms[name + ".ff_self.0"];
    const double ln_bytes = 2.0 * M * C * sizeof(T);
    bool head_done = false;
    if constexpr (sizeof(T) == 2) {
      // L <= 128: LayerNorm + in_proj + attention core in one launch (attn_head.cu); the tail kernel follows
      if (!p->tr && !p->no_head && attn_head_supported(L, C, 4) && attn_tail_supported(C) && !getenv("SPDM_NO_ATTN_TAIL")) {
        // C = 64 blocks (one CTA owns all channels of a tile) can also run their tail inside the head kernel.  Measured: +0.8 % at
        // batch 4096 (att never leaves the chip), -0.8 % at batch 256 (half as many CTAs share the tail work) -> large batches only
        const bool want_tail = C == 64 && !p->no_head_tail && (long long)Bpad * L >= 128LL * 4 * 148;
        AttnHead*& hk = p->head_cache[name + (want_tail ? "|whole" : "")];
        if (!hk) {
          GemmW& wi = p->gemms[name + ".attention.in_proj_weight"];
          GemmW& wo = p->gemms[name + ".attention.out_proj.weight"];
          GemmW& w1 = p->gemms[name + ".ff_self.1.weight"];
          GemmW& w2 = p->gemms[name + ".ff_self.3.weight"];
          const AttnHeadTail tl{wo.w16, w1.w16, w2.w16, wo.bias, w1.bias, w2.bias, n2.g, n2.b};
          hk = attn_head_create(wi.w16, wi.bias, n1.g, n1.b, C, L, 4, want_tail ? &tl : nullptr);
          REQUIRE(hk != nullptr, "%s: attn_head_create failed", name.c_str());
        }
        const long long Mpad = (long long)Bpad * L;
        const bool whole = attn_head_merges_tail(hk);   // C = 64: head and tail in one launch
        timed(p, c.s, PC_SDPA, (whole ? 12.0 : 6.0) * M * C * C + 4.0 * M * L * C, ((whole ? 3.0 : 2.0) * M * C + 3.0 * C * C) * sizeof(T), [&] {
          attn_head_launch(hk, reinterpret_cast<const bf16*>(x), ld_x, reinterpret_cast<bf16*>(attn), Mpad, c.s,
                           reinterpret_cast<bf16*>(out), ld_out);
        });
        if (whole) return;
        head_done = true;
      }
    }
    if (!head_done) timed(p, c.s, PC_LN, 0, ln_bytes, [&] { launch_layernorm<T>(x, ld_x, ln, C, n1.g, n1.b, M, C, c.s); });
    bool tc_sdpa = false;
    if constexpr (sizeof(T) == 2) tc_sdpa = sdpa_tc_supported(L, C, 4);
    if (head_done) {
    } else if (tc_sdpa) {
      if constexpr (sizeof(T) == 2) {
        SdpaTc*& sd = p->sdpa_cache[name + "|" + std::to_string(c.b0)];
        if (!sd) {
          sd = sdpa_tc_create(reinterpret_cast<const bf16*>(qkv), reinterpret_cast<const bf16*>(att(p->a_vt[level], level)), C, L, 4,
                              (long long)p->Bcap * L);
          REQUIRE(sd != nullptr, "%s: sdpa_tc_create failed", name.c_str());
        }
        vt = reinterpret_cast<bf16*>(att(p->a_vt[level], level));
        vt_lk = sdpa_tc_keys_per_tile(L);
        gemm(name + ".attention.in_proj_weight", ln, C, level, qkv, 3 * C, EPI_BIAS | EPI_VT);
        vt = nullptr;
        vt_lk = 0;
        const long long Mpad = (long long)Bpad * L;
        timed(p, c.s, PC_SDPA, 4.0 * M * L * C, 4.0 * M * C * sizeof(T),
              [&] { sdpa_tc_launch(sd, reinterpret_cast<bf16*>(attn), Mpad, c.s); });
      }
    } else {
      gemm(name + ".attention.in_proj_weight", ln, C, level, qkv, 3 * C, EPI_BIAS);
      timed(p, c.s, PC_SDPA, 4.0 * M * L * C, 4.0 * M * C * sizeof(T), [&] { launch_sdpa<T>(qkv, attn, c.B, L, C, 4, c.s); });
    }
    if constexpr (sizeof(T) == 2) {
      if (attn_tail_supported(C) && !getenv("SPDM_NO_ATTN_TAIL")) {
        AttnTail*& tl = p->tail_cache[name + "|" + std::to_string(c.b0)];
        if (!tl) {
          GemmW& wo = p->gemms[name + ".attention.out_proj.weight"];
          GemmW& w1 = p->gemms[name + ".ff_self.1.weight"];
          GemmW& w2 = p->gemms[name + ".ff_self.3.weight"];
          tl = attn_tail_create(reinterpret_cast<const bf16*>(attn), (long long)p->Bcap * L, C, wo.w16, w1.w16, w2.w16, wo.bias, w1.bias,
                                w2.bias, n2.g, n2.b);
          REQUIRE(tl != nullptr, "%s: attn_tail_create failed", name.c_str());
        }
        const long long Mpad = (long long)Bpad * L;
        timed(p, c.s, PC_GEMM1, 6.0 * M * C * C, (4.0 * M * C + 3.0 * C * C) * 2.0, [&] {
          attn_tail_launch(tl, reinterpret_cast<const bf16*>(x), ld_x, reinterpret_cast<bf16*>(out), ld_out, Mpad, c.s);
        });
        return;
      }
    }
    gemm(name + ".attention.out_proj.weight", attn, C, level, res, C, EPI_BIAS | EPI_RESID, x, ld_x);
    timed(p, c.s, PC_LN, 0, ln_bytes, [&] { launch_layernorm<T>(res, C, ln, C, n2.g, n2.b, M, C, c.s); });
    gemm(name + ".ff_self.1.weight", ln, C, level, ff, C, EPI_BIAS | EPI_GELU);
    gemm(name + ".ff_self.3.weight", ff, C, level, out, ld_out, EPI_BIAS | EPI_RESID, res, C);
  }

  // UNet_Film.forward (models/Unet_FiLmLayer.py:277-312) / UNet_Film_noAttention.forward
  void run() {
    p->timed_idx = 0;
    const int rows = p->cfg.rows, dim = p->cfg.dim;
    T* cat3 = act(p->cat[0], 0); T* cat2 = act(p->cat[1], 1); T* cat1 = act(p->cat[2], 2);
    // ---- inc ----
    const double hw0 = (double)p->H0 * p->W0;
    bool inc_applied = false;
    if constexpr (sizeof(T) == 2) {
      // default horizon: the first conv, its GroupNorm and the GELU in one launch (the block owns the whole 32x8 sample)
      if (p->H0 * p->W0 == 256 && !p->tr && !(p->tap_out && p->tap_name == "inc.first") && !getenv("SPDM_NO_CONV_IN_GN")) {
        NormW& n0 = p->norms["inc.norm"];
        timed(p, c.s, PC_IO, 2.0 * 9 * 64 * hw0 * c.B, c.B * hw0 * 64 * sizeof(T), [&] {
          launch_conv_in_gn(c.x, p->w_in, n0.g, n0.b, reinterpret_cast<bf16*>(act(p->hbuf[0], 0)), c.B, p->H0, p->W0, rows, dim, p->lh, p->lw, c.s);
        });
        inc_applied = true;
      }
    }
    if (!inc_applied) {
      timed(p, c.s, PC_IO, 2.0 * 9 * 64 * hw0 * c.B, c.B * hw0 * 64 * sizeof(T), [&] {
        launch_conv_in<T>(c.x, p->w_in, act(p->raw[0], 0), stats(), c.B, p->H0, p->W0, rows, dim, p->lh, p->lw, c.s);
      });
    }
    curP = 1;
    double_conv("inc", nullptr, 0, 64, 0, cat3 + 64, 128, nullptr, inc_applied ? 2 : 1);  // x1 -> skip slot of up3
    tap("x1", cat3 + 64, 128, 64, 0);
    tap("inc", cat3 + 64, 128, 64, 0);

    // ---- down path ----
    struct DownCfg { int stage; const char* sa; const T* in; int ld_in; int level; T* dest; int ld_dest; const char* tapname; };
    const DownCfg downs[3] = {
        {0, "sa1", cat3 + 64, 128, 1, cat2 + 128, 256, "x2"},
        {1, "sa2", cat2 + 128, 256, 2, cat1 + 256, 512, "x3"},
        {2, "sa3", cat1 + 256, 512, 3, act(p->bbuf[3], 3), 256, "x4"}};
    for (const DownCfg& d : downs) {
      const StageInfo& st = kStages[d.stage];
      const int l = d.level;
      T* a = act(p->abuf[l], l); T* b = act(p->bbuf[l], l);
      {  // deep levels at small batch: pool + the four convs of the stage in one launch
        T* last = p->attention ? a : d.dest;
        const int ld_last = p->attention ? st.cout : d.ld_dest;
        if (l >= 2 && chain(st.name, {{std::string(st.name) + ".doubleConv1", st.cin, st.cin, nullptr, b, st.cin},
                                       {std::string(st.name) + ".doubleConv2", st.cin, st.cout, &st, last, ld_last}},
                            a, st.cin, l, 1, d.in, d.ld_in, st.cin)) {
          if (p->attention) self_attention(d.sa, a, st.cout, st.cout, l, d.dest, d.ld_dest);
          continue;
        }
      }
      timed(p, c.s, PC_RESAMPLE, 0, 5.0 * c.B * p->levelH(l) * p->levelW(l) * st.cin * sizeof(T),
            [&] { launch_pool<T>(d.in, d.ld_in, a, st.cin, c.B, p->levelH(l), p->levelW(l), st.cin, c.s); });
      double_conv(std::string(st.name) + ".doubleConv1", a, st.cin, st.cin, l, b, st.cin, nullptr);
      if (p->attention) {
        double_conv(std::string(st.name) + ".doubleConv2", b, st.cin, st.cout, l, a, st.cout, &st);
        tap(st.name, a, st.cout, st.cout, l);
        self_attention(d.sa, a, st.cout, st.cout, l, d.dest, d.ld_dest);
        tap(d.sa, d.dest, d.ld_dest, st.cout, l);
      } else {
        double_conv(std::string(st.name) + ".doubleConv2", b, st.cin, st.cout, l, d.dest, d.ld_dest, &st);
        tap(st.name, d.dest, d.ld_dest, st.cout, l);
      }
      tap(d.tapname, d.dest, d.ld_dest, st.cout, l);
    }
    // ---- bottleneck ----
    T* a3 = act(p->abuf[3], 3); T* b3 = act(p->bbuf[3], 3);
    if (!chain("bot", {{"bot1", 256, 512, nullptr, a3, 512}, {"bot2", 512, 512, nullptr, b3, 512}, {"bot3", 512, 256, nullptr, a3, 256}}, b3, 256, 3, 0,
               nullptr, 0, 0)) {
    double_conv("bot1", b3, 256, 512, 3, a3, 512, nullptr);
    tap("bot1", a3, 512, 512, 3);
    double_conv("bot2", a3, 512, 512, 3, b3, 512, nullptr);
    tap("bot2", b3, 512, 512, 3);
    double_conv("bot3", b3, 512, 256, 3, a3, 256, nullptr);
    tap("bot3", a3, 256, 256, 3);
    }
    tap("x5", a3, 256, 256, 3);

    // ---- up path ----
    struct UpCfg { int stage; const char* sa; const T* low; int c_low; int level; T* catbuf; const char* tapname; };
    const UpCfg ups[3] = {
        {3, "sa4", a3, 256, 2, cat1, "u1"}, {4, "sa5", act(p->bbuf[2], 2), 128, 1, cat2, "u2"}, {5, "sa6", act(p->bbuf[1], 1), 64, 0, cat3, "u3"}};
    for (const UpCfg& u : ups) {
      const StageInfo& st = kStages[u.stage];
      const int l = u.level;
      T* a = act(p->abuf[l], l); T* b = act(p->bbuf[l], l);
      {  // deep level at small batch: upsample (+ concat) + the four convs of the stage in one launch
        T* last = p->attention ? a : b;
        if (l >= 2 && chain(st.name, {{std::string(st.name) + ".doubleConv1", st.cin, st.cin, nullptr, a, st.cin},
                                       {std::string(st.name) + ".doubleConv2", st.cin, st.cout, &st, last, st.cout}},
                            u.catbuf, st.cin, l, 2, u.low, u.c_low, u.c_low)) {
          if (p->attention) self_attention(u.sa, a, st.cout, st.cout, l, b, st.cout);
          continue;
        }
      }
      timed(p, c.s, PC_RESAMPLE, 0, 1.25 * c.B * p->levelH(l) * p->levelW(l) * u.c_low * sizeof(T), [&] {
        launch_upsample<T>(u.low, u.c_low, u.catbuf, st.cin, c.B, p->levelH(l + 1), p->levelW(l + 1), u.c_low, c.s);
      });
      double_conv(std::string(st.name) + ".doubleConv1", u.catbuf, st.cin, st.cin, l, a, st.cin, nullptr);
      if (p->attention) {
        // doubleConv2 reads a, writes a (safe: the input is dead once its first conv has run)
        double_conv(std::string(st.name) + ".doubleConv2", a, st.cin, st.cout, l, a, st.cout, &st);
        tap(st.name, a, st.cout, st.cout, l);
        self_attention(u.sa, a, st.cout, st.cout, l, b, st.cout);
        tap(u.sa, b, st.cout, st.cout, l);
      } else {
        double_conv(std::string(st.name) + ".doubleConv2", a, st.cin, st.cout, l, b, st.cout, &st);
        tap(st.name, b, st.cout, st.cout, l);
      }
      tap(u.tapname, b, st.cout, st.cout, l);
    }
    // ---- outc + unpad ----
    if (c.fuse_step) {  // graphed loop: outc + posterior update + inpaint in one kernel, eps never stored
      timed(p, c.s, PC_STEP, 2.0 * 64 * rows * dim * c.B, c.B * hw0 * 64 * sizeof(T), [&] {
        launch_outc_step<T>(*c.fuse_step, act(p->bbuf[0], 0), 64, p->w_outc, p->b_outc, p->H0, p->W0, 64, rows, dim, p->lh, p->lw, c.s);
      });
      return;
    }
    timed(p, c.s, PC_IO, 2.0 * 64 * rows * dim * c.B, c.B * hw0 * 64 * sizeof(T), [&] {
      launch_outc<T>(act(p->bbuf[0], 0), 64, p->w_outc, p->b_outc, c.out, c.B, p->H0, p->W0, 64, rows, dim, p->lh, p->lw, c.s);
    });
  }
};

}  // namespace
#include "simple_unet.inl"
#include "resnet.inl"
namespace {

void run_forward(spdm_plan* p, const FwdCtx& c) {
  if (p->simple) { SimpleFwd f(p, c); f.run(); return; }
  if (p->bf16_mode) { Fwd<bf16> f(p, c); f.run(); }
  else { Fwd<float> f(p, c); f.run(); }
}

void train_wait_packs(spdm_plan* p);  // train_impl.inl: block until weight repacks still running on the side stream are done
void check_ready(spdm_plan* p, bool training_step = false) {
  if (p->sched_only) throw SpdmError{"this plan was created scheduler-only (SPDM_FLAG_SCHEDULER_ONLY)"};
  if (p->tr && !training_step) train_wait_packs(p);
  if (!p->missing_unet.empty()) throw SpdmError{"U-Net weights missing, first: " + *p->missing_unet.begin()};
}

void ensure_temb_table(spdm_plan* p, cudaStream_t s) {
  REQUIRE(p->K > 0, "no schedule set (spdm_plan_set_schedule)");
  if (!p->temb_table_dirty) return;
  if (p->simple) {
    REQUIRE(p->su_table != nullptr, "weight pos_encoding.pos_encoding missing");
    launch_su_temb(p->timesteps, p->K, p->su_table, p->su_table_rows, p->temb_w, p->temb_b, p->temb_table, p->cfg.time_dim, SU_TEMB_WIDTH, s);
    p->temb_table_dirty = false;
    return;
  }
  launch_temb(p->timesteps, p->K, p->inv_freq, p->temb_w, p->temb_b, p->temb_table, p->cfg.time_dim, s);
  p->temb_table_dirty = false;
}

void compute_film(spdm_plan* p, int B, cudaStream_t s) {
  if (p->simple) { compute_cond_emb_simple(p, B, s); return; }
  launch_mish(p->cond, p->cond_mish, (long long)B * p->G, s);
  GemmSimtArgs a{};
  a.in = p->cond_mish; a.w = p->film_w; a.bias = p->film_b; a.out = p->film; a.M = B; a.Cin = p->G; a.Cout = SPDM_FILM_WIDTH;
  a.ld_in = p->G; a.ld_out = SPDM_FILM_WIDTH; a.H = 1; a.W = 1; a.taps = 1; a.act = ACT_NONE;
  launch_gemm_simt<float, float>(a, s);
  p->have_cond = true;
}

void one_lane(spdm_plan* p, int b0, int Bsub, int B, bool use_film, cudaStream_t s, int step_off) {
  const size_t n = p->n_elems();
  FwdCtx c{};
  c.x = p->xt + (size_t)b0 * n; c.out = p->eps + (size_t)b0 * n; c.temb = p->temb_table; c.temb_mode = TEMB_STEP;
  c.step_ptr = &p->dyn->step; c.step_off = step_off;
  c.film = use_film ? p->film + (size_t)b0 * SPDM_FILM_WIDTH : nullptr; c.B = Bsub; c.b0 = b0; c.s = s;
  StepArgs a{};
  a.x = p->xt; a.eps = p->eps; a.x_out = p->xt; a.coef = p->coef; a.dyn = p->dyn; a.n = p->n_elems();
  a.inpaint_elems = p->cfg.inpaint_rows * p->cfg.dim; a.B = Bsub; a.b0 = b0; a.B_total = B; a.step_off = step_off;
  c.fuse_step = &a;
  run_forward(p, c);
}

// One denoising step for B trajectories.  With split > 1 the batch is cut into sub-batches ("lanes") that run the
// whole U-Net concurrently on separate streams (forked from and joined back into `s`, also under stream capture, where
// they become parallel graph branches): the many small-grid kernels of the deep levels then overlap instead of
// leaving most SMs idle.
// `step_off` / `advance`: inside a captured graph of several steps every launch carries its step's offset from the device counter,
// which then advances once per graph launch instead of once per step (one launch fewer per denoising step).
void one_step(spdm_plan* p, int B, bool use_film, cudaStream_t s, int step_off = 0, int advance = 1) {
  int split = p->split;
  int chunk = (B + split - 1) / split;
  chunk = ((chunk + p->bm - 1) / p->bm) * p->bm;
  if (split <= 1 || chunk >= B) {
    one_lane(p, 0, B, B, use_film, s, step_off);
  } else {
    CUDA_OK(cudaEventRecord(p->ev_fork, s));
    int lane = 0;
    for (int b0 = 0; b0 < B; b0 += chunk, ++lane) {
      const int Bsub = B - b0 < chunk ? B - b0 : chunk;
      cudaStream_t ls = lane == 0 ? s : p->lane_stream[lane - 1];
      if (lane > 0) CUDA_OK(cudaStreamWaitEvent(ls, p->ev_fork, 0));
      one_lane(p, b0, Bsub, B, use_film, ls, step_off);
      if (lane > 0) {
        CUDA_OK(cudaEventRecord(p->ev_lane[lane - 1], ls));
        CUDA_OK(cudaStreamWaitEvent(s, p->ev_lane[lane - 1], 0));
      }
    }
  }
  if (advance > 0) launch_advance(&p->dyn->step, advance, s);
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
#define API_BEGIN try {
#define API_END                                   \
  }                                               \
  catch (const SpdmError& e) { return fail("%s", e.msg.c_str()); } \
  catch (const std::exception& e) { return fail("%s", e.what()); }

static void check_async(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw SpdmError{std::string(what) + ": " + cudaGetErrorString(e)};
}

extern "C" int spdm_plan_create(spdm_plan** out, const spdm_config* cfg) {
  API_BEGIN
  REQUIRE(out && cfg, "null argument");
  if (const char* e = getenv("SPDM_PDL")) g_spdm_pdl = atoi(e) != 0;  // A/B switch for programmatic dependent launch
  REQUIRE(cfg->rows > 0 && cfg->dim > 0 && cfg->batch_max > 0, "rows/dim/batch_max must be positive");
  REQUIRE(cfg->time_dim > 0 && cfg->time_dim % 2 == 0 && cfg->time_dim <= 4096, "bad time_dim");
  REQUIRE(cfg->precision == SPDM_PRECISION_FP32 || cfg->precision == SPDM_PRECISION_BF16 || cfg->precision == SPDM_PRECISION_TF32, "bad precision");
  REQUIRE(cfg->inpaint_rows >= 0 && cfg->inpaint_rows <= cfg->rows, "bad inpaint_rows");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) throw SpdmError{"no CUDA device: libspdm has no CPU fallback"};
  CUDA_OK(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, cfg->device));
  REQUIRE(prop.major == 10, "libspdm is built for sm_100a only (device is sm_%d%d)", prop.major, prop.minor);
  spdm_plan* p = new spdm_plan();
  p->cfg = *cfg;
  REQUIRE(cfg->variant == SPDM_VARIANT_ATTENTION || cfg->variant == SPDM_VARIANT_NO_ATTENTION || cfg->variant == SPDM_VARIANT_SIMPLE_UNET,
          "bad variant");
  p->attention = cfg->variant == SPDM_VARIANT_ATTENTION;
  p->simple = cfg->variant == SPDM_VARIANT_SIMPLE_UNET;
  p->bf16_mode = cfg->precision == SPDM_PRECISION_BF16;
  p->tf32_mode = cfg->precision == SPDM_PRECISION_TF32;
  if (p->simple && !(cfg->flags & SPDM_FLAG_SCHEDULER_ONLY)) {
    // its channel counts (16, 160, 288, 448, 224, 96, 112) are not multiples of the 64-wide tcgen05 operand tiles
    if (p->bf16_mode || p->tf32_mode) { delete p; throw SpdmError{"the simple U-Net (models/simple_Unet.py) runs on the fp32 path only: create the plan with SPDM_PRECISION_FP32"}; }
    if (cfg->obs_horizon * cfg->cond_dim <= 0) { delete p; throw SpdmError{"the simple U-Net needs conditioning (cond_dim > 0): its stages concatenate a 32-channel cond_emb"}; }
  }
  if (const char* e = getenv("SPDM_NO_PFOLD")) p->no_pfold = atoi(e) != 0;
  // pad_to(x, 8): models/Unet_FiLmLayer.py:15-34
  auto up8 = [](int v) { return v % 8 ? v + 8 - v % 8 : v; };
  p->H0 = up8(cfg->rows); p->W0 = up8(cfg->dim);
  p->lh = (p->H0 - cfg->rows) / 2; p->lw = (p->W0 - cfg->dim) / 2;
  p->G = cfg->obs_horizon * cfg->cond_dim;
  try {
    if (p->bf16_mode || p->tf32_mode) {
      REQUIRE(p->W0 <= 128 && 128 % p->W0 == 0, "bf16 path: padded width %d must divide 128", p->W0);
      p->bm = tc_batch_multiple(p->levelH(3), p->levelW(3));
      for (int l = 0; l < 4; ++l) {
        const int hw = p->levelH(l) * p->levelW(l);
        REQUIRE(hw >= 128 ? hw % 128 == 0 || (128 % p->levelW(l) == 0 && p->levelH(l) % (128 / p->levelW(l)) == 0) : 128 % hw == 0,
                "bf16 path: level %d geometry %dx%d is not tileable into 128-row tiles", l, p->levelH(l), p->levelW(l));
      }
    }
    p->Bcap = ((cfg->batch_max + p->bm - 1) / p->bm) * p->bm;
    p->sched_only = (cfg->flags & SPDM_FLAG_SCHEDULER_ONLY) != 0;  // scheduler-only plan: spdm_step / spdm_add_noise, no U-Net
    if (p->sched_only) { *out = p; return 0; }
    p->enc_resnet = (cfg->flags & SPDM_FLAG_ENCODER_RESNET18) != 0;
    REQUIRE(!(p->enc_resnet && p->simple), "the ResNet18 encoder is wired to the FiLM U-Nets");
    if (p->simple) register_weights_simple(p); else register_weights(p);
    if (p->enc_resnet) register_weights_resnet(p);
    if (p->bf16_mode) alloc_workspace<bf16>(p); else alloc_workspace<float>(p);
    p->stats = p->alloc<float>((size_t)p->Bcap * SPDM_MAX_PARTIALS * 2);
    p->temb_call = p->alloc<float>((size_t)p->Bcap * SPDM_TEMB_WIDTH);
    if (p->G > 0) {
      p->film = p->alloc<float>((size_t)p->Bcap * SPDM_FILM_WIDTH);
      p->cond = p->alloc<float>((size_t)p->Bcap * p->G);
      p->cond_mish = p->alloc<float>((size_t)p->Bcap * p->G);
    }
    p->xt = p->alloc<float>((size_t)p->Bcap * p->n_elems());
    p->eps = p->alloc<float>((size_t)p->Bcap * p->n_elems());
    p->dyn = p->alloc<StepDyn>(1);
    CUDA_OK(cudaMallocHost((void**)&p->dyn_host, sizeof(StepDyn) * spdm_plan::DYN_SLOTS));
    for (int i = 0; i < spdm_plan::DYN_SLOTS; ++i) CUDA_OK(cudaEventCreateWithFlags(&p->dyn_ev[i], cudaEventDisableTiming));
    CUDA_OK(cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking));
    CUDA_OK(cudaEventCreateWithFlags(&p->ev_in, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&p->ev_out, cudaEventDisableTiming));
    p->split = (cfg->flags >> 8) & 0xF;
    if (const char* e = getenv("SPDM_SPLIT")) p->split = atoi(e);
    if (const char* e = getenv("SPDM_NO_FUSE_APPLY")) p->no_fuse = atoi(e) != 0;
    if (const char* e = getenv("SPDM_FUSE_MODE")) p->fuse_mode = atoi(e);
    if (const char* e = getenv("SPDM_NO_SPLITK")) p->no_splitk = atoi(e) != 0;
    if (const char* e = getenv("SPDM_NO_FOLD")) p->no_fold = atoi(e) != 0;
    if (const char* e = getenv("SPDM_NO_ATTN_HEAD")) p->no_head = atoi(e) != 0;
    if (const char* e = getenv("SPDM_NO_HEAD_TAIL")) p->no_head_tail = atoi(e) != 0;
    if (const char* e = getenv("SPDM_SKIP_IDX")) {
      p->skip.assign(256, 0);
      for (const char* q = e; *q;) {
        char* end = nullptr;
        const long v = strtol(q, &end, 10);
        if (end == q) break;
        if (v >= 0 && v < 256) p->skip[v] = 1;
        q = (*end == ',') ? end + 1 : end;
      }
    }
    if (const char* e = getenv("SPDM_ENC_SIMT_INFER")) p->enc_simt_infer = atoi(e) != 0;
    if (p->split < 1) p->split = 1;
    if (p->split > 8) p->split = 8;
    CUDA_OK(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
    for (int i = 0; i < 7; ++i) {
      CUDA_OK(cudaStreamCreateWithFlags(&p->lane_stream[i], cudaStreamNonBlocking));
      CUDA_OK(cudaEventCreateWithFlags(&p->ev_lane[i], cudaEventDisableTiming));
    }
  } catch (...) {
    for (void* q : p->allocs) cudaFree(q);
    delete p;
    throw;
  }
  *out = p;
  return 0;
  API_END
}

extern "C" int spdm_plan_destroy(spdm_plan* p) {
  if (!p) return 0;
  cudaDeviceSynchronize();
  for (auto& kv : p->graphs) {
    if (kv.second.multi) cudaGraphExecDestroy(kv.second.multi);
    if (kv.second.single) cudaGraphExecDestroy(kv.second.single);
  }
  for (auto& kv : p->tc_cache) tc_gemm_destroy(kv.second);
  for (auto& kv : p->chain_cache) tc_chain_destroy(kv.second);
  for (auto& kv : p->tf_cache) tf32_conv_destroy(kv.second);
  if (p->enc_tc) tc_gemm_destroy(p->enc_tc);
  if (p->enc_tc2) tc_gemm_destroy(p->enc_tc2);
  if (p->enc_tc3) tc_gemm_destroy(p->enc_tc3);
  for (auto& kv : p->sdpa_cache) sdpa_tc_destroy(kv.second);
  for (auto& kv : p->head_cache) attn_head_destroy(kv.second);
  for (cudaEvent_t e : p->ev_pool) cudaEventDestroy(e);
  for (auto& kv : p->tail_cache) attn_tail_destroy(kv.second);
  for (void* q : p->allocs) cudaFree(q);
  train_destroy(p);
  if (p->dyn_host) cudaFreeHost(p->dyn_host);
  for (int i = 0; i < spdm_plan::DYN_SLOTS; ++i) if (p->dyn_ev[i]) cudaEventDestroy(p->dyn_ev[i]);
  if (p->own_stream) cudaStreamDestroy(p->own_stream);
  if (p->ev_in) cudaEventDestroy(p->ev_in);
  if (p->ev_out) cudaEventDestroy(p->ev_out);
  if (p->ev_fork) cudaEventDestroy(p->ev_fork);
  for (int i = 0; i < 7; ++i) {
    if (p->lane_stream[i]) cudaStreamDestroy(p->lane_stream[i]);
    if (p->ev_lane[i]) cudaEventDestroy(p->ev_lane[i]);
  }
  delete p;
  return 0;
}

extern "C" int spdm_plan_load_weight(spdm_plan* p, const char* name, const float* src, const int64_t* shape, int32_t ndim, void* stream) {
  API_BEGIN
  REQUIRE(p && name && src && shape, "null argument");
  auto it = p->loaders.find(name);
  REQUIRE(it != p->loaders.end(), "unexpected weight name '%s'", name);
  it->second(src, shape, ndim, (cudaStream_t)stream);
  check_async("load_weight");
  p->missing_unet.erase(name);
  p->missing_enc.erase(name);
  p->temb_table_dirty = true;
  return 0;
  API_END
}

extern "C" int spdm_plan_missing_weights(spdm_plan* p) {
  if (!p) return fail("null plan");
  std::string s;
  for (auto& n : p->missing_unet) s += n + " ";
  for (auto& n : p->missing_enc) s += n + " ";
  snprintf(g_err, sizeof g_err, "%s", s.c_str());
  return (int)(p->missing_unet.size() + p->missing_enc.size());
}

extern "C" int spdm_plan_set_schedule(spdm_plan* p, int32_t kind, int32_t K, const float* coef, const int64_t* timesteps, void* stream) {
  API_BEGIN
  REQUIRE(p && coef && timesteps && K > 0, "bad argument");
  REQUIRE(kind == SPDM_SCHED_DDPM || kind == SPDM_SCHED_DDIM, "bad schedule kind");
  cudaStream_t s = (cudaStream_t)stream;
  if (K > p->K || !p->coef) {  // grow (old buffers stay in the plan's allocation list)
    p->coef = p->alloc<float>((size_t)K * 8);
    p->timesteps = p->alloc<long long>(K);
    p->temb_table = p->alloc<float>((size_t)K * SPDM_TEMB_WIDTH);
    for (auto& kv : p->graphs) {  // graphs bake the table pointers
      if (kv.second.multi) cudaGraphExecDestroy(kv.second.multi);
      if (kv.second.single) cudaGraphExecDestroy(kv.second.single);
    }
    p->graphs.clear();
  }
  CUDA_OK(cudaStreamSynchronize(s));
  CUDA_OK(cudaMemcpy(p->coef, coef, (size_t)K * 8 * sizeof(float), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(p->timesteps, timesteps, (size_t)K * sizeof(long long), cudaMemcpyHostToDevice));
  p->K = K; p->sched_kind = kind; p->temb_table_dirty = true;
  return 0;
  API_END
}

// Autoencoder.encoder on n frames: `images` fp32 (n,3,96,96), or -- images_u8 non-null -- uint8 HWC (n,96,96,3) decoded x / 255 in flight
static void encode_images_impl(spdm_plan* p, const float* images, const uint8_t* images_u8, float* out, int32_t n, cudaStream_t s) {
  REQUIRE(p && (images || images_u8) && out && n > 0, "bad argument");
  if (!p->missing_enc.empty()) throw SpdmError{"vision encoder weights missing, first: " + *p->missing_enc.begin()};
  if (images_u8 && (p->enc_resnet || !(p->bf16_mode && !p->enc_simt_infer))) {
    // fp32 (parity) plans and the A/B CUDA-core conv stack read fp32 frames: decode into a plan-owned staging buffer first
    if (p->enc_u8_cap < n) { p->enc_u8_stage = p->alloc<float>((size_t)n * 3 * 96 * 96); p->enc_u8_cap = n; }
    launch_decode_u8_hwc(images_u8, p->enc_u8_stage, n, 96, 96, s);
    images = p->enc_u8_stage;
    images_u8 = nullptr;
  }
  if (p->enc_resnet) {   // ResNet18-GroupNorm: (n, 3, 96, 96) -> (n, 512)
    if (p->bf16_mode) resnet_encode<bf16>(p, images, out, n, s); else resnet_encode<float>(p, images, out, n, s);
    check_async("encode_images (resnet18)");
    return;
  }
  if (p->bf16_mode) {
    // bf16 plan: conv stack writes bf16 features, Linear(9216 -> 128) runs on the tcgen05 GEMM (one 128-frame tile per CTA)
    if (!p->enc_feat16) {
      p->enc_chunk = 4096;
      p->enc_feat16 = p->alloc<bf16>((size_t)p->enc_chunk * 9216);
      p->enc_out16 = p->alloc<bf16>((size_t)p->enc_chunk * 128);
      p->enc_tc = tc_gemm_create(p->enc_feat16, 9216, p->enc_wl16, 9216, 128, 1, 1, 1, p->enc_chunk);
      REQUIRE(p->enc_tc != nullptr, "encoder GEMM: %s", tc_last_error());
    }
    if (!p->enc_simt_infer && !p->enc_c1p) {  // conv2 / conv3 as flat tcgen05 GEMMs over patch-major rows
      const long long rows2 = (long long)p->enc_chunk * 288, rows3 = (long long)p->enc_chunk * 144;
      p->enc_c1p = p->alloc<bf16>((size_t)p->enc_chunk * 576 * 64);
      p->enc_c2 = p->alloc<bf16>((size_t)rows2 * 64);
      p->enc_tc2 = tc_gemm_create(p->enc_c1p, 128, p->enc_w2p, 128, 64, 1, 1, 1, (int)rows2);
      p->enc_tc3 = tc_gemm_create(p->enc_c2, 128, p->enc_w3p, 128, 64, 1, 1, 1, (int)rows3);
      REQUIRE(p->enc_tc2 && p->enc_tc3, "encoder conv GEMMs: %s", tc_last_error());
    }
    for (int f0 = 0; f0 < n; f0 += p->enc_chunk) {
      const int m = n - f0 < p->enc_chunk ? n - f0 : p->enc_chunk;
      if (p->enc_simt_infer) {
        launch_enc_convs<bf16>(images + (size_t)f0 * 3 * 96 * 96, p->enc_w1, p->enc_b1, p->enc_w2, p->enc_b2, p->enc_w3, p->enc_b3,
                               p->enc_feat16, m, s);
      } else {
        const int m8 = (m + 7) / 8 * 8;  // whole 128-row tiles: 288 * 8 and 144 * 8 rows; rows past m are scratch nobody reads
        if (images_u8) launch_enc_conv1_fwd_u8(images_u8 + (size_t)f0 * 3 * 96 * 96, p->enc_w1, p->enc_b1, p->enc_c1p, m, s);
        else launch_enc_conv1_fwd(images + (size_t)f0 * 3 * 96 * 96, p->enc_w1, p->enc_b1, p->enc_c1p, m, 1, 3 * 96 * 96, s);
        tc_gemm_launch(p->enc_tc2, p->enc_c2, 64, nullptr, p->enc_b2p, nullptr, 0, EPI_BIAS | EPI_RELU, m8 * 288, s);
        tc_gemm_launch(p->enc_tc3, p->enc_feat16, 64, nullptr, p->enc_b3, nullptr, 0, EPI_BIAS | EPI_RELU, m8 * 144, s);
      }
      tc_gemm_launch(p->enc_tc, p->enc_out16, 128, nullptr, p->enc_bl, nullptr, 0, EPI_BIAS, ((m + 127) / 128) * 128, s);
      launch_cast_f32(p->enc_out16, out + (size_t)f0 * 128, (long long)m * 128, s);
    }
  } else {
  if (!p->enc_feat) { p->enc_chunk = 4096; p->enc_feat = p->alloc<float>((size_t)p->enc_chunk * 9216); }
  for (int f0 = 0; f0 < n; f0 += p->enc_chunk) {
    const int m = n - f0 < p->enc_chunk ? n - f0 : p->enc_chunk;
    launch_enc_convs<float>(images + (size_t)f0 * 3 * 96 * 96, p->enc_w1, p->enc_b1, p->enc_w2, p->enc_b2, p->enc_w3, p->enc_b3, p->enc_feat, m, s);
    GemmSimtArgs a{};
    a.in = p->enc_feat; a.w = p->enc_wl; a.bias = p->enc_bl; a.out = out + (size_t)f0 * 128; a.M = m; a.Cin = 9216; a.Cout = 128;
    a.ld_in = 9216; a.ld_out = 128; a.H = 1; a.W = 1; a.taps = 1; a.act = ACT_NONE;
    launch_gemm_simt<float, float>(a, s);
  }
  }
  check_async("encode_images");
}

extern "C" int spdm_encode_images(spdm_plan* p, const float* images, float* out, int32_t n, void* stream) {
  API_BEGIN
  encode_images_impl(p, images, nullptr, out, n, (cudaStream_t)stream);
  return 0;
  API_END
}

extern "C" int spdm_set_cond(spdm_plan* p, const float* obs_cond, int32_t B, void* stream) {
  API_BEGIN
  REQUIRE(p && obs_cond && B > 0 && B <= p->cfg.batch_max, "bad argument (B=%d, batch_max=%d)", B, p ? p->cfg.batch_max : 0);
  REQUIRE(p->G > 0, "plan was created unconditional (cond_dim == 0)");
  check_ready(p);
  cudaStream_t s = (cudaStream_t)stream;
  CUDA_OK(cudaMemcpyAsync(p->cond, obs_cond, (size_t)B * p->G * sizeof(float), cudaMemcpyDeviceToDevice, s));
  compute_film(p, B, s);
  check_async("set_cond");
  return 0;
  API_END
}

static void encode_cond_impl(spdm_plan* p, const float* images, const uint8_t* images_u8, const float* position, const float* action,
                             const float* velocity, int32_t B, cudaStream_t s) {
  REQUIRE(p && (images || images_u8) && position && action && velocity && B > 0 && B <= p->cfg.batch_max, "bad argument");
  REQUIRE(p->cfg.cond_dim == 7 + p->feat_dim(), "encode_cond needs cond_dim == %d (2 pos + 3 act + 2 vel + %d image features)", 7 + p->feat_dim(),
          p->feat_dim());
  check_ready(p);
  const int T = p->cfg.obs_horizon;
  if (!p->enc_out) p->enc_out = p->alloc<float>((size_t)p->Bcap * T * p->feat_dim());
  encode_images_impl(p, images, images_u8, p->enc_out, B * T, s);
  launch_build_cond(position, action, velocity, p->enc_out, p->cond, B, T, p->cfg.cond_dim, s);
  compute_film(p, B, s);
  check_async("encode_cond");
}

extern "C" int spdm_encode_cond(spdm_plan* p, const float* images, const float* position, const float* action, const float* velocity,
                                int32_t B, void* stream) {
  API_BEGIN
  REQUIRE(images, "bad argument");
  encode_cond_impl(p, images, nullptr, position, action, velocity, B, (cudaStream_t)stream);
  return 0;
  API_END
}

extern "C" int spdm_encode_cond_u8(spdm_plan* p, const uint8_t* images_hwc, const float* position, const float* action, const float* velocity,
                                   int32_t B, void* stream) {
  API_BEGIN
  REQUIRE(images_hwc, "bad argument");
  encode_cond_impl(p, nullptr, images_hwc, position, action, velocity, B, (cudaStream_t)stream);
  return 0;
  API_END
}

extern "C" int spdm_get_cond(spdm_plan* p, float* out, int32_t B, void* stream) {
  API_BEGIN
  REQUIRE(p && out && B > 0 && B <= p->cfg.batch_max && p->G > 0, "bad argument");
  CUDA_OK(cudaMemcpyAsync(out, p->cond, (size_t)B * p->G * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
  API_END
}

static int unet_forward_impl(spdm_plan* p, const float* x, const int64_t* t, int32_t t_count, const float* y, int32_t use_cached_cond,
                             float* out, int32_t B, void* stream) {
  REQUIRE(p && x && t && out, "null argument");
  REQUIRE(B > 0 && B <= p->cfg.batch_max, "B=%d outside 1..batch_max=%d", B, p->cfg.batch_max);
  REQUIRE(t_count == 1 || t_count == B, "t_count must be 1 or B");
  check_ready(p);
  cudaStream_t s = (cudaStream_t)stream;
  const float* film = nullptr;
  if (y) {
    REQUIRE(p->G > 0, "plan was created unconditional");
    CUDA_OK(cudaMemcpyAsync(p->cond, y, (size_t)B * p->G * sizeof(float), cudaMemcpyDeviceToDevice, s));
    compute_film(p, B, s);
    film = p->film;
  } else if (use_cached_cond) {
    REQUIRE(p->have_cond, "no cached conditioning (call spdm_set_cond / spdm_encode_cond)");
    film = p->film;
  }
  if (p->simple) {
    REQUIRE(p->su_table != nullptr, "weight pos_encoding.pos_encoding missing");
    launch_su_temb(reinterpret_cast<const long long*>(t), t_count, p->su_table, p->su_table_rows, p->temb_w, p->temb_b, p->temb_call,
                   p->cfg.time_dim, SU_TEMB_WIDTH, s);
  } else {
    launch_temb(reinterpret_cast<const long long*>(t), t_count, p->inv_freq, p->temb_w, p->temb_b, p->temb_call, p->cfg.time_dim, s);
  }
  FwdCtx c{};
  c.x = x; c.out = out; c.temb = p->temb_call; c.temb_mode = t_count == 1 ? TEMB_ROW0 : TEMB_PER_SAMPLE; c.step_ptr = nullptr;
  c.film = film; c.B = B; c.s = s;
  const long long before = total_launches();
  run_forward(p, c);
  p->launches += total_launches() - before + 1;
  check_async("unet_forward");
  return 0;
}

extern "C" int spdm_unet_forward(spdm_plan* p, const float* x, const int64_t* t, int32_t t_count, const float* y,
                                 int32_t use_cached_cond, float* out, int32_t B, void* stream) {
  API_BEGIN
  return unet_forward_impl(p, x, t, t_count, y, use_cached_cond, out, B, stream);
  API_END
}

extern "C" int64_t spdm_debug_forward(spdm_plan* p, const float* x, const int64_t* t, int32_t t_count, const float* y,
                                      int32_t use_cached_cond, float* out, int32_t B, const char* tap_name, float* tap_out, void* stream) {
  API_BEGIN
  REQUIRE(p && tap_name && tap_out, "null argument");
  p->tap_name = tap_name; p->tap_out = tap_out; p->tap_count = -1;
  int rc = unet_forward_impl(p, x, t, t_count, y, use_cached_cond, out, B, stream);
  p->tap_out = nullptr;
  if (rc) return rc;
  REQUIRE(p->tap_count >= 0, "unknown tap '%s'", tap_name);
  return p->tap_count;
  API_END
}

extern "C" int spdm_step(spdm_plan* p, const float* x, const float* eps, const float* noise, const float* inpaint, float* x_out,
                         int32_t step, int32_t B, void* stream) {
  API_BEGIN
  REQUIRE(p && x && eps && x_out, "null argument");
  REQUIRE(p->K > 0 && step >= 0 && step < p->K, "step %d outside schedule of %d steps", step, p->K);
  REQUIRE(B > 0, "bad B");
  StepArgs a{};
  a.x = x; a.eps = eps; a.x_out = x_out; a.coef = p->coef; a.dyn = nullptr; a.noise = noise;
  a.inpaint = p->cfg.inpaint_rows > 0 ? inpaint : nullptr; a.step_host = step; a.b0 = 0; a.B_total = B;
  a.n = p->n_elems(); a.inpaint_elems = p->cfg.inpaint_rows * p->cfg.dim; a.B = B;
  launch_step(a, (cudaStream_t)stream);
  check_async("step");
  return 0;
  API_END
}

extern "C" int spdm_sample(spdm_plan* p, const float* x_T, const float* noise, const float* inpaint, float* out, float* history,
                           uint64_t seed, int32_t B, void* stream) {
  API_BEGIN
  REQUIRE(p && x_T && out, "null argument");
  REQUIRE(B > 0 && B <= p->cfg.batch_max, "B=%d outside 1..batch_max=%d", B, p->cfg.batch_max);
  check_ready(p);
  const bool use_film = p->G > 0;
  if (use_film) REQUIRE(p->have_cond, "no cached conditioning (call spdm_set_cond / spdm_encode_cond)");
  cudaStream_t user = (cudaStream_t)stream;
  cudaStream_t s = user;
  const int gs = p->cfg.graph_steps;
  if (gs > 0) {  // fork onto the plan's stream: capture is not allowed on the legacy default stream
    s = p->own_stream;
    CUDA_OK(cudaEventRecord(p->ev_in, user));
    CUDA_OK(cudaStreamWaitEvent(s, p->ev_in, 0));
  }
  ensure_temb_table(p, s);
  const size_t nb = (size_t)B * p->n_elems() * sizeof(float);
  // per-call dynamic parameters live in device memory so that the captured graphs are call-independent; the call itself only
  // enqueues (no host synchronisation: the pinned staging block comes from a ring)
  StepDyn d{};
  d.step = 0; d.seed = seed; d.noise = noise; d.inpaint = p->cfg.inpaint_rows > 0 ? inpaint : nullptr; d.history = history;
  d.use_philox = noise ? 0 : 1;
  StepDyn* dh = p->dyn_slot();
  *dh = d;
  CUDA_OK(cudaMemcpyAsync(p->dyn, dh, sizeof(StepDyn), cudaMemcpyHostToDevice, s));

  const long long key = ((long long)B << 1) | (use_film ? 1 : 0);
  const int n_multi = gs > 0 ? p->K / gs : 0, n_single = gs > 0 ? p->K % gs : 0;
  if (gs > 0) {
    spdm_plan::GraphSet& g = p->graphs[key];
    if ((n_multi > 0 && !g.multi) || (n_single > 0 && !g.single)) {
      // One eager step first: builds the tensor maps and sets the kernel attributes outside of stream capture.
      const long long before = total_launches();
      one_step(p, B, use_film, s);
      p->launches += total_launches() - before;
      CUDA_OK(cudaMemcpyAsync(p->dyn, dh, sizeof(StepDyn), cudaMemcpyHostToDevice, s));
      auto capture = [&](int steps, cudaGraphExec_t* exec, long long* count) {
        cudaGraph_t graph;
        const long long c0 = total_launches();
        CUDA_OK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        try {
          for (int k = 0; k < steps; ++k) one_step(p, B, use_film, s, k, k + 1 == steps ? steps : 0);
        } catch (...) {
          cudaStreamEndCapture(s, &graph);
          throw;
        }
        CUDA_OK(cudaStreamEndCapture(s, &graph));
        *count = total_launches() - c0;
        CUDA_OK(cudaGraphInstantiate(exec, graph, 0));
        CUDA_OK(cudaGraphDestroy(graph));
      };
      if (n_multi > 0 && !g.multi) capture(gs, &g.multi, &g.n_multi);
      if (n_single > 0 && !g.single) capture(1, &g.single, &g.n_single);
    }
  }
  CUDA_OK(cudaMemcpyAsync(p->xt, x_T, nb, cudaMemcpyDeviceToDevice, s));
  if (history) CUDA_OK(cudaMemcpyAsync(history, x_T, nb, cudaMemcpyDeviceToDevice, s));
  if (gs <= 0) {
    const long long before = total_launches();
    for (int k = 0; k < p->K; ++k) one_step(p, B, use_film, s);
    p->launches += total_launches() - before;
  } else {
    spdm_plan::GraphSet& g = p->graphs[key];
    for (int i = 0; i < n_multi; ++i) CUDA_OK(cudaGraphLaunch(g.multi, s));
    for (int i = 0; i < n_single; ++i) CUDA_OK(cudaGraphLaunch(g.single, s));
    p->launches += n_multi * g.n_multi + n_single * g.n_single;
  }
  CUDA_OK(cudaMemcpyAsync(out, p->xt, nb, cudaMemcpyDeviceToDevice, s));
  CUDA_OK(cudaEventRecord(p->dyn_ev[p->dyn_cur], s));
  if (s != user) {  // join back
    CUDA_OK(cudaEventRecord(p->ev_out, s));
    CUDA_OK(cudaStreamWaitEvent(user, p->ev_out, 0));
  }
  check_async("sample");
  return 0;
  API_END
}

extern "C" int spdm_add_noise(spdm_plan* p, const float* x0, const float* noise, const int64_t* t, const float* sqrt_ab,
                              const float* sqrt_1mab, const float* inpaint, float* out, int32_t B, void* stream) {
  API_BEGIN
  REQUIRE(p && x0 && noise && t && sqrt_ab && sqrt_1mab && out && B > 0, "bad argument");
  launch_add_noise(x0, noise, reinterpret_cast<const long long*>(t), sqrt_ab, sqrt_1mab, p->cfg.inpaint_rows > 0 ? inpaint : nullptr, out,
                   p->n_elems(), p->cfg.inpaint_rows * p->cfg.dim, B, (cudaStream_t)stream);
  check_async("add_noise");
  return 0;
  API_END
}

static_assert(PC_N == SPDM_PROFILE_CLASSES, "profile class count");

extern "C" int spdm_profile_step(spdm_plan* p, int32_t B, int32_t reps, double* out, void* stream) {
  API_BEGIN
  REQUIRE(p && out && B > 0 && B <= p->cfg.batch_max && reps > 0, "bad argument");
  check_ready(p);
  const bool use_film = p->G > 0;
  if (use_film) REQUIRE(p->have_cond, "no cached conditioning");
  cudaStream_t s = (cudaStream_t)stream;
  ensure_temb_table(p, s);
  CUDA_OK(cudaStreamSynchronize(s));
  StepDyn d{};
  d.use_philox = 1;
  StepDyn* dh = p->dyn_slot();
  *dh = d;
  CUDA_OK(cudaMemcpyAsync(p->dyn, dh, sizeof(StepDyn), cudaMemcpyHostToDevice, s));
  one_step(p, B, use_film, s);  // warm-up (tensor maps, attributes)
  for (int i = 0; i < SPDM_PROFILE_CLASSES * 4; ++i) out[i] = 0.0;
  {  // size the event pool before timing
    p->prof_on = true; p->prof.clear(); p->ev_used = 0;
    try { one_step(p, B, use_film, s); } catch (...) { p->prof_on = false; throw; }
    p->prof_on = false;
    CUDA_OK(cudaStreamSynchronize(s));
  }
  for (int r = 0; r < reps; ++r) {
    CUDA_OK(cudaMemcpyAsync(p->dyn, dh, sizeof(StepDyn), cudaMemcpyHostToDevice, s));
    // a ~3 ms spin kernel lets the host enqueue the whole step ahead of the GPU, so the event intervals are
    // back-to-back kernel durations and not host launch latency
    launch_delay(6000000LL, s);
    p->prof_on = true;
    p->prof.clear();
    p->ev_used = 0;
    try { one_step(p, B, use_film, s); } catch (...) { p->prof_on = false; throw; }
    p->prof_on = false;
    CUDA_OK(cudaStreamSynchronize(s));
    int rec_i = 0;
    for (auto& rec : p->prof) {
      float ms = 0.f;
      CUDA_OK(cudaEventElapsedTime(&ms, rec.e0, rec.e1));
      if (r == reps - 1 && getenv("SPDM_PROF_DUMP"))   // per-launch list of the last repetition (diagnostics)
        fprintf(stderr, "spdm prof %3d class %d  %9.2f us  %8.1f GFLOP  %8.1f MB\n", rec_i, rec.cat, ms * 1e3, rec.flops / 1e9, rec.bytes / 1e6);
      ++rec_i;
      out[rec.cat * 4 + 0] += ms / reps;
      out[rec.cat * 4 + 1] += 1.0 / reps;
      out[rec.cat * 4 + 2] += rec.flops / reps;
      out[rec.cat * 4 + 3] += rec.bytes / reps;
    }
    p->prof.clear();
  }
  CUDA_OK(cudaEventRecord(p->dyn_ev[p->dyn_cur], s));
  check_async("profile_step");
  return 0;
  API_END
}

// Microbenchmark of one tcgen05 implicit-GEMM launch on zero-filled buffers (tests/perf_conv.py): average ms over
// `iters` back-to-back launches.  `dbg` = TcParams::dbg switches (0 = the real kernel).
extern "C" int spdm_microbench_conv(int32_t H, int32_t W, int32_t B, int32_t Cin, int32_t Cout, int32_t taps, int32_t dbg,
                                    int32_t iters, float* ms_out) {
  API_BEGIN
  REQUIRE(ms_out && iters > 0, "bad argument");
  const size_t M = (size_t)B * H * W;
  bf16 *in = nullptr, *w = nullptr, *out = nullptr;
  float* stats = nullptr;
  CUDA_OK(cudaMalloc(&in, M * Cin * 2));
  CUDA_OK(cudaMalloc(&w, (size_t)taps * Cin * Cout * 2));
  CUDA_OK(cudaMalloc(&out, M * Cout * 2));
  CUDA_OK(cudaMalloc(&stats, (size_t)B * SPDM_MAX_PARTIALS * 2 * 4));
  CUDA_OK(cudaMemset(in, 0, M * Cin * 2));
  CUDA_OK(cudaMemset(w, 0, (size_t)taps * Cin * Cout * 2));
  TcGemm* g = tc_gemm_create(in, Cin, w, Cin, Cout, taps, H, W, B);
  REQUIRE(g != nullptr, "%s", tc_last_error());
  tc_set_debug(dbg);
  cudaEvent_t e0, e1;
  CUDA_OK(cudaEventCreate(&e0));
  CUDA_OK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) tc_gemm_launch(g, out, Cout, stats, nullptr, nullptr, 0, EPI_STATS, B, 0);
  CUDA_OK(cudaEventRecord(e0, 0));
  for (int i = 0; i < iters; ++i) tc_gemm_launch(g, out, Cout, stats, nullptr, nullptr, 0, EPI_STATS, B, 0);
  CUDA_OK(cudaEventRecord(e1, 0));
  cudaError_t e = cudaEventSynchronize(e1);
  tc_set_debug(0);
  CUDA_OK(e);
  float ms = 0.f;
  CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
  *ms_out = ms / iters;
  tc_gemm_destroy(g);
  cudaFree(in); cudaFree(w); cudaFree(out); cudaFree(stats);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return 0;
  API_END
}

extern "C" int64_t spdm_plan_launch_count(spdm_plan* p) { return p ? p->launches : 0; }
extern "C" int32_t spdm_plan_batch_multiple(spdm_plan* p) { return p ? p->bm : 0; }
extern "C" int64_t spdm_plan_workspace_bytes(spdm_plan* p) { return p ? (int64_t)p->bytes : 0; }

#include "train_impl.inl"
