// train_impl.inl — the training step of libspdm (included at the end of plan.cu: it needs spdm_plan and Fwd<T>).
//
//   spdm_train_fwd_bwd  =  Diffusion_DDPM.process_single_batch + loss.backward()   (models/diffusion_ddpm.py:128-173)
//   spdm_adam_step      =  clip_grad_norm_(0.5) + torch.optim.Adam.step()          (ddpm:115-125, train.py:104-107)
//
// Forward: the same kernels as inference (tcgen05 implicit-GEMM convs / Linears, GroupNorm apply, attention core ...), but
// every intermediate the backward pass needs is kept in a per-step arena instead of the ping-pong inference workspace, and
// the SelfAttention block runs unfused so that its pre-activations exist.  Backward: data gradients of convs / Linears are
// forward GEMMs with tap-flipped, transposed weights (the "#d" twins, repacked with every weight upload); weight gradients
// are the pixel-contraction GEMM of wgrad_tc.cu (bf16) / wgrad_simt (fp32 parity path); everything else is bwd_kernels.cu.
// Parameter gradients land, fp32 and in PyTorch layout, in the caller's flat gradient buffer at the bound offsets.

struct TrainBind { long long off; std::vector<int64_t> shape; };
struct TrainState {
  float* params = nullptr; float* grads = nullptr; long long total = 0;
  std::map<std::string, TrainBind> bind;
  // arena: buffers are handed out in the same order every step, so pointers (and the TMA maps built on them) are stable
  std::vector<void*> bufs; std::vector<size_t> buf_bytes; size_t cur = 0; int arena_B = -1;
  float* film_wT = nullptr;   // [1792][G]
  bf16* film_wf16 = nullptr;  // bf16 path: [1792][GP] forward / [GP][1792] data-gradient operands of the six FiLM Linears as ONE GEMM each
  bf16* film_wb16 = nullptr; int GP = 0;
  bool film_simt = false;     // SPDM_FILM_SIMT=1: CUDA-core fp32 FiLM Linears on the bf16 path too (A/B switch)
  float* enc_wlT = nullptr;   // [128][9216 hwc]
  float* packed = nullptr; size_t packed_elems = 0;  // bf16 path: conv weight gradients as [tap][Cout][Cin]
  std::map<std::string, size_t> packed_off;
  float* loss_dev = nullptr;
  // bf16 path: vision encoder as patch GEMMs (layouts: bwd_kernels.cu, "Vision encoder on the tensor cores")
  bf16* enc_wlT16 = nullptr;  // conv2 / conv3 packs live in the plan (inference uses them too)
  bool enc_simt = false;      // SPDM_ENC_SIMT=1: CUDA-core encoder on the bf16 path too (A/B switch)
  // weight / bias gradients are off the critical path (nothing in the step reads them): they run on a side stream, forked
  // from the main stream after the gradient they consume exists and joined at the end of the step
  cudaStream_t side = nullptr;
  std::vector<cudaEvent_t> evs; size_t ev_cur = 0;
  cudaEvent_t ev_join = nullptr;
  cudaEvent_t ev_packs = nullptr, ev_packs_fork = nullptr;  // U-Net weight repacks run on the side stream, under the encoder forward
  bool packs_pending = false;
  long long image_bstride = 0;  // elements between samples of `images` (0 = contiguous); spdm_train_set_image_stride
  int valid = 0;                // spdm_train_set_valid: real samples at the head of the batch (0 = all of them)
  // gradient-completion phases for overlapped data-parallel all-reduces: 0 = up path + outc + sa4-6 done, 1 = everything of the
  // U-Net except the time-embedding / FiLM Linears done, 2 = all gradients done (same point as the end of the step)
  cudaEvent_t ev_phase[3] = {nullptr, nullptr, nullptr};
  bool use_side = true;       // SPDM_TRAIN_SIDE=0 switches it off (A/B)
  bool dgrad_swap = true;     // SPDM_DGRAD_SWAP=0: plain N = Cout tiles for the narrow data gradients (A/B)
  bool wgrad_simt = false;    // SPDM_WGRAD_SIMT=1: CUDA-core weight gradients on the bf16 path too (A/B switch)
};

namespace {
float* train_film_wT(spdm_plan* p) { return p->tr ? p->tr->film_wT : nullptr; }
void train_film_weight_loaded(spdm_plan* p, const float* src, int off, int C2, cudaStream_t s) {
  if (!p->tr) return;
  TrainState* tr = p->tr;
  if (tr->film_wT)  // [1792][G]: rows off.. = this stage's (2C, G) weight as stored by PyTorch
    CUDA_OK(cudaMemcpyAsync(tr->film_wT + (size_t)off * p->G, src, (size_t)C2 * p->G * sizeof(float), cudaMemcpyDeviceToDevice, s));
  if (tr->film_wf16) launch_film_pack16(src, tr->film_wf16, tr->film_wb16, C2, p->G, tr->GP, off, s);
}
float* train_enc_wlT(spdm_plan* p) { return p->tr ? p->tr->enc_wlT : nullptr; }
void train_destroy(spdm_plan* p) {
  if (!p->tr) return;
  for (void* q : p->tr->bufs) cudaFree(q);
  for (cudaEvent_t e : p->tr->evs) cudaEventDestroy(e);
  if (p->tr->ev_join) cudaEventDestroy(p->tr->ev_join);
  for (cudaEvent_t e : p->tr->ev_phase) if (e) cudaEventDestroy(e);
  if (p->tr->ev_packs) cudaEventDestroy(p->tr->ev_packs);
  if (p->tr->ev_packs_fork) cudaEventDestroy(p->tr->ev_packs_fork);
  if (p->tr->side) cudaStreamDestroy(p->tr->side);
  delete p->tr;
  p->tr = nullptr;
}

void train_wait_packs(spdm_plan* p) {
  if (p->tr && p->tr->packs_pending) {
    CUDA_OK(cudaEventSynchronize(p->tr->ev_packs));
    p->tr->packs_pending = false;
  }
}

void arena_reset(spdm_plan* p, int B) {
  TrainState* tr = p->tr;
  if (tr->arena_B != B) {  // new batch size: drop the arena (and everything keyed on its pointers)
    CUDA_OK(cudaDeviceSynchronize());
    for (void* q : tr->bufs) cudaFree(q);
    tr->bufs.clear(); tr->buf_bytes.clear();
    tr->arena_B = B;
  }
  tr->cur = 0;
}
void* arena_alloc(spdm_plan* p, size_t bytes) {
  TrainState* tr = p->tr;
  if (tr->cur < tr->bufs.size()) {
    REQUIRE(tr->buf_bytes[tr->cur] == bytes, "internal: training arena replay mismatch at buffer %zu", tr->cur);
    return tr->bufs[tr->cur++];
  }
  void* q = nullptr;
  CUDA_OK(cudaMalloc(&q, bytes));
  CUDA_OK(cudaMemset(q, 0, bytes));
  tr->bufs.push_back(q); tr->buf_bytes.push_back(bytes); tr->cur++;
  p->bytes += bytes;
  return q;
}

template <typename T> struct Train {
  spdm_plan* p;
  TrainState* tr;
  Fwd<T> f;
  int B;
  cudaStream_t s;
  float* d_temb = nullptr;  // [B][896]
  float* d_film = nullptr;  // [B][1792]
  float* dgrad_stats = nullptr;  // scratch GroupNorm partial sums written (and ignored) by swapped-operand dgrad launches

  struct DC {
    std::string name; const T* in; int ld_in, Cin, Cout, level;
    T* raw1; float* st1; int P1; T* h; T* raw2; float* st2; int P2; const StageInfo* st; bool first_is_in;
  };
  struct SA { std::string name; const T* x; int ld_x, C, level; T *h1, *qkv, *attn, *a, *f0, *f1, *f2; };

  Train(spdm_plan* p_, const FwdCtx& c) : p(p_), tr(p_->tr), f(p_, c), B(c.B), s(c.s) {}

  long long M(int level) const { return (long long)f.Bpad * p->levelH(level) * p->levelW(level); }
  T* A(int level, int C) { return reinterpret_cast<T*>(arena_alloc(p, (size_t)M(level) * C * sizeof(T))); }
  float* Fbuf(size_t n) { return reinterpret_cast<float*>(arena_alloc(p, n * sizeof(float))); }
  float* S() { return Fbuf((size_t)f.Bpad * SPDM_MAX_PARTIALS * 2); }
  float* G(const std::string& name) {
    auto it = tr->bind.find(name);
    REQUIRE(it != tr->bind.end(), "training: parameter '%s' is not bound (spdm_train_bind)", name.c_str());
    return tr->grads + it->second.off;
  }

  // stream for work nothing downstream in the step depends on (ordered after everything enqueued on the main stream so far)
  cudaStream_t fork_side() {
    if (!tr->use_side) return s;
    if (tr->ev_cur == tr->evs.size()) {
      cudaEvent_t e;
      CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      tr->evs.push_back(e);
    }
    cudaEvent_t e = tr->evs[tr->ev_cur++];
    CUDA_OK(cudaEventRecord(e, s));
    CUDA_OK(cudaStreamWaitEvent(tr->side, e, 0));
    return tr->side;
  }

  // every gradient of phase `ph` has been enqueued (main or side stream): record the event a communication stream can wait on
  void mark_phase(int ph) {
    cudaStream_t sd = fork_side();  // the side stream now also waits for the main stream's work up to here
    CUDA_OK(cudaEventRecord(tr->ev_phase[ph], sd));
  }

  // ---- weight / bias gradients ----
  void wgrad(const std::string& wname, const T* x, int ld_x, const T* dy, int ld_dy, int level) {
    cudaStream_t s = fork_side();
    GemmW& g = p->gemms[wname];
    const int H = p->levelH(level), W = p->levelW(level);
    const long long m = M(level);
    const std::string pname = g.taps == 9 ? wname + ".weight" : wname;  // convs are registered without the ".weight" suffix
    if constexpr (sizeof(T) == 2) {
      if (!tr->wgrad_simt) {
        float* dst = g.taps == 9 ? tr->packed + tr->packed_off[wname] : G(pname);
        const int rc = wgrad_tc_launch(reinterpret_cast<const bf16*>(x), ld_x, reinterpret_cast<const bf16*>(dy), ld_dy, m, g.Cin, g.Cout, H, W,
                                       g.taps, dst, s);
        REQUIRE(rc == 0, "%s: %s", wname.c_str(), wgrad_tc_last_error());
        if (g.taps == 9 && !launch_unpack_conv3_grad(dst, G(pname), g.Cout, g.Cin, s)) launch_unpack_conv_grad(dst, G(pname), g.Cout, g.Cin, s);  // [tap][Cout][Cin] -> PyTorch layout, same stream
        return;
      }
    }
    WgradArgs a{};
    a.x = x; a.ld_x = ld_x; a.dy = dy; a.ld_dy = ld_dy; a.M = m; a.Cin = g.Cin; a.Cout = g.Cout; a.H = H; a.W = W; a.taps = g.taps;
    a.dw = G(pname);
    launch_wgrad_simt<T, T>(a, s);
  }
  void bias_grad(const std::string& bname, const T* dy, int ld, int level, int N) {
    launch_colsum<T>(dy, ld, M(level), N, G(bname), tr->use_side ? tr->side : s);  // always right after the wgrad of the same dy
  }

  // ---- DoubleConvolution (models/Unet_FiLmLayer.py:85-115) ----
  DC dc_fwd(const std::string& name, const T* in, int ld_in, int Cin, int Cout, int level, T* out, int ld_out, const StageInfo* st,
            bool first_is_in = false, const float* x_noisy = nullptr) {
    DC d{name, in, ld_in, Cin, Cout, level, nullptr, nullptr, 1, nullptr, nullptr, nullptr, 1, st, first_is_in};
    d.raw1 = A(level, Cout); d.st1 = S();
    f.stats_ov = d.st1;
    if (first_is_in) {
      launch_conv_in<T>(x_noisy, p->w_in, d.raw1, d.st1, B, p->H0, p->W0, p->cfg.rows, p->cfg.dim, p->lh, p->lw, s);
      f.curP = 1;
    } else {
      f.gemm(name + ".first", in, ld_in, level, d.raw1, Cout, EPI_STATS);
    }
    d.P1 = f.curP;
    d.h = A(level, Cout);
    f.apply(name + ".norm", d.raw1, Cout, Cout, level, d.h, Cout, ACT_GELU, nullptr);
    d.raw2 = A(level, Cout); d.st2 = S();
    f.stats_ov = d.st2;
    f.gemm(name + ".second", d.h, Cout, level, d.raw2, Cout, EPI_STATS);
    d.P2 = f.curP;
    f.apply(name + ".norm", d.raw2, Cout, Cout, level, out, ld_out, ACT_NONE, st);
    f.stats_ov = nullptr;
    return d;
  }

  void gn_bwd(const DC& d, bool second, const T* dy, int ld_dy, T* dx) {
    NormW& n = p->norms[d.name + ".norm"];
    GnBwdArgs a{};
    a.dy = dy; a.ld_dy = ld_dy; a.raw = second ? d.raw2 : d.raw1; a.ld_raw = d.Cout; a.dx = dx; a.ld_dx = d.Cout;
    a.stats = second ? d.st2 : d.st1; a.P = second ? d.P2 : d.P1;
    a.gamma = n.g; a.beta = n.b; a.dgamma = G(d.name + ".norm.weight"); a.dbeta = G(d.name + ".norm.bias");
    a.act = second ? ACT_NONE : ACT_GELU;
    if (second && d.st) {
      a.temb = p->temb_call; a.temb_off = d.st->temb_off; a.d_temb = d_temb;
      if (f.c.film) { a.film = f.c.film; a.film_off = d.st->film_off; a.d_film = d_film; }
    }
    a.HW = p->levelH(d.level) * p->levelW(d.level); a.C = d.Cout; a.eps = 1e-5f;
    a.part = Fbuf((size_t)f.Bpad * 2 * d.Cout);
    if (launch_gn_bwd<T>(a, B, s))   // batch sum of the (d gamma, d beta) partials: a parameter gradient, off the critical path
      launch_gn_part_finalize(a.part, a.dgamma, a.dbeta, B, a.C, fork_side());
  }

  // Data gradient of a 3x3 conv.  The swapped-operand kernel (N = 256 pixels per MMA instead of N = Cout <= 128) only exists
  // with the GroupNorm-statistics epilogue, so narrow dgrads ask for statistics into a scratch slot and ignore them.
  void dgrad3(const std::string& wname, const T* dy, int ld_dy, int level, T* dx, int ld_dx) {
    int flags = 0;
    if constexpr (sizeof(T) == 2) {
      if (tr->dgrad_swap && p->gemms[wname].Cout <= 128) {
        if (!dgrad_stats) dgrad_stats = S();
        f.stats_ov = dgrad_stats;
        flags = EPI_STATS;
      }
    }
    f.gemm(wname, dy, ld_dy, level, dx, ld_dx, flags);
    f.stats_ov = nullptr;
  }

  // d_in: [M, Cin] (ld_din) or null when the block's input needs no gradient (inc)
  void dc_bwd(const DC& d, const T* d_out, int ld_dout, T* d_in, int ld_din, const float* x_noisy = nullptr) {
    T* d_raw2 = A(d.level, d.Cout);
    gn_bwd(d, true, d_out, ld_dout, d_raw2);
    wgrad(d.name + ".second", d.h, d.Cout, d_raw2, d.Cout, d.level);
    T* d_h = A(d.level, d.Cout);
    dgrad3(d.name + ".second#d", d_raw2, d.Cout, d.level, d_h, d.Cout);
    T* d_raw1 = A(d.level, d.Cout);
    gn_bwd(d, false, d_h, d.Cout, d_raw1);
    if (d.first_is_in) {
      launch_conv_in_wgrad<T>(x_noisy, d_raw1, G("inc.first.weight"), B, p->H0, p->W0, p->cfg.rows, p->cfg.dim, p->lh, p->lw, s);
    } else {
      wgrad(d.name + ".first", d.in, d.ld_in, d_raw1, d.Cout, d.level);
      if (d_in) dgrad3(d.name + ".first#d", d_raw1, d.Cout, d.level, d_in, ld_din);
    }
  }

  // ---- SelfAttention (models/Unet_FiLmLayer.py:44-82), unfused so that the backward has its intermediates ----
  SA sa_fwd(const std::string& name, const T* x, int ld_x, int C, int level, T* out, int ld_out) {
    SA a{name, x, ld_x, C, level};
    const int L = p->levelH(level) * p->levelW(level);
    const long long m = (long long)B * L;
    NormW& n1 = p->norms[name + ".ln"];
    NormW& n2 = p->norms[name + ".ff_self.0"];
    a.h1 = A(level, C);
    launch_layernorm<T>(x, ld_x, a.h1, C, n1.g, n1.b, m, C, s);
    a.qkv = A(level, 3 * C);
    f.gemm(name + ".attention.in_proj_weight", a.h1, C, level, a.qkv, 3 * C, EPI_BIAS);
    a.attn = A(level, C);
    bool mma_done = false;
    if constexpr (sizeof(T) == 2) {
      static int simt = -1;  // SPDM_SDPA_FWD_SIMT=1: CUDA-core forward core (A/B switch)
      if (simt < 0) { const char* e = getenv("SPDM_SDPA_FWD_SIMT"); simt = e ? atoi(e) : 0; }
      if (!simt) mma_done = launch_sdpa_fwd_mma(reinterpret_cast<const bf16*>(a.qkv), reinterpret_cast<bf16*>(a.attn), B, L, C, 4, s);
    }
    if (!mma_done) launch_sdpa<T>(a.qkv, a.attn, B, L, C, 4, s);
    a.a = A(level, C);
    f.gemm(name + ".attention.out_proj.weight", a.attn, C, level, a.a, C, EPI_BIAS | EPI_RESID, x, ld_x);
    a.f0 = A(level, C);
    launch_layernorm<T>(a.a, C, a.f0, C, n2.g, n2.b, m, C, s);
    a.f1 = A(level, C);
    f.gemm(name + ".ff_self.1.weight", a.f0, C, level, a.f1, C, EPI_BIAS);
    a.f2 = A(level, C);
    launch_gelu_fwd<T>(a.f1, a.f2, M(level) * C, s);
    f.gemm(name + ".ff_self.3.weight", a.f2, C, level, out, ld_out, EPI_BIAS | EPI_RESID, a.a, C);
    return a;
  }

  T* sa_bwd(const SA& a, const T* d_out, int ld_dout) {
    const int C = a.C, level = a.level;
    const int L = p->levelH(level) * p->levelW(level);
    const long long m = (long long)B * L;
    const std::string& n = a.name;
    NormW& n1 = p->norms[n + ".ln"];
    NormW& n2 = p->norms[n + ".ff_self.0"];
    // out = W2 f2 + b2 + a
    wgrad(n + ".ff_self.3.weight", a.f2, C, d_out, ld_dout, level);
    bias_grad(n + ".ff_self.3.bias", d_out, ld_dout, level, C);
    T* d_f2 = A(level, C);
    f.gemm(n + ".ff_self.3.weight#d", d_out, ld_dout, level, d_f2, C, 0);
    T* d_f1 = A(level, C);
    launch_gelu_bwd<T>(d_f2, a.f1, d_f1, M(level) * C, s);
    wgrad(n + ".ff_self.1.weight", a.f0, C, d_f1, C, level);
    bias_grad(n + ".ff_self.1.bias", d_f1, C, level, C);
    T* d_f0 = A(level, C);
    f.gemm(n + ".ff_self.1.weight#d", d_f1, C, level, d_f0, C, 0);
    T* d_a = A(level, C);  // through LN2, plus the residual branch
    launch_layernorm_bwd<T>(d_f0, C, a.a, C, n2.g, d_out, ld_dout, d_a, C, G(n + ".ff_self.0.weight"), G(n + ".ff_self.0.bias"), m, C, s);
    // a = Wo attn + bo + x
    wgrad(n + ".attention.out_proj.weight", a.attn, C, d_a, C, level);
    bias_grad(n + ".attention.out_proj.bias", d_a, C, level, C);
    T* d_attn = A(level, C);
    f.gemm(n + ".attention.out_proj.weight#d", d_a, C, level, d_attn, C, 0);
    T* d_qkv = A(level, 3 * C);
    launch_sdpa_bwd<T>(a.qkv, a.attn, d_attn, d_qkv, B, L, C, 4, s);
    wgrad(n + ".attention.in_proj_weight", a.h1, C, d_qkv, 3 * C, level);
    bias_grad(n + ".attention.in_proj_bias", d_qkv, 3 * C, level, 3 * C);
    T* d_h1 = A(level, C);
    f.gemm(n + ".attention.in_proj_weight#d", d_qkv, 3 * C, level, d_h1, C, 0);
    T* d_x = A(level, C);
    launch_layernorm_bwd<T>(d_h1, C, a.x, a.ld_x, n1.g, d_a, C, d_x, C, G(n + ".ln.weight"), G(n + ".ln.bias"), m, C, s);
    return d_x;
  }

  // ---- the whole step: forward with saved activations, loss, backward ----
  void run(const float* x_noisy, const float* noise, const float* silu_pe, const long long* t_dev) {
    (void)t_dev;
    const bool att = p->attention;
    const int rows = p->cfg.rows, dim = p->cfg.dim;
    d_temb = Fbuf((size_t)f.Bpad * SPDM_TEMB_WIDTH);
    d_film = Fbuf((size_t)f.Bpad * SPDM_FILM_WIDTH);
    CUDA_OK(cudaMemsetAsync(d_temb, 0, (size_t)f.Bpad * SPDM_TEMB_WIDTH * sizeof(float), s));  // accumulated by gn_bwd
    CUDA_OK(cudaMemsetAsync(d_film, 0, (size_t)f.Bpad * SPDM_FILM_WIDTH * sizeof(float), s));
    T* cat3 = A(0, 128); T* cat2 = A(1, 256); T* cat1 = A(2, 512);

    // ================= forward =================
    DC inc = dc_fwd("inc", nullptr, 0, 1, 64, 0, cat3 + 64, 128, nullptr, true, x_noisy);
    struct DownRec { DC dc1, dc2; SA sa; const T* pool_in; int ld_pool_in; T* pooled; };
    struct DownCfg { int stage; const char* sa; const T* in; int ld_in; int level; T* dest; int ld_dest; };
    T* x4 = A(3, 256);
    const DownCfg downs[3] = {{0, "sa1", cat3 + 64, 128, 1, cat2 + 128, 256}, {1, "sa2", cat2 + 128, 256, 2, cat1 + 256, 512},
                              {2, "sa3", cat1 + 256, 512, 3, x4, 256}};
    std::vector<DownRec> drec(3);
    for (int i = 0; i < 3; ++i) {
      const DownCfg& d = downs[i];
      const StageInfo& st = kStages[d.stage];
      const int l = d.level;
      DownRec& r = drec[i];
      r.pool_in = d.in; r.ld_pool_in = d.ld_in;
      r.pooled = A(l, st.cin);
      launch_pool<T>(d.in, d.ld_in, r.pooled, st.cin, B, p->levelH(l), p->levelW(l), st.cin, s);
      T* b = A(l, st.cin);
      r.dc1 = dc_fwd(std::string(st.name) + ".doubleConv1", r.pooled, st.cin, st.cin, st.cin, l, b, st.cin, nullptr);
      if (att) {
        T* cbuf = A(l, st.cout);
        r.dc2 = dc_fwd(std::string(st.name) + ".doubleConv2", b, st.cin, st.cin, st.cout, l, cbuf, st.cout, &st);
        r.sa = sa_fwd(d.sa, cbuf, st.cout, st.cout, l, d.dest, d.ld_dest);
      } else {
        r.dc2 = dc_fwd(std::string(st.name) + ".doubleConv2", b, st.cin, st.cin, st.cout, l, d.dest, d.ld_dest, &st);
      }
    }
    T* b1o = A(3, 512); T* b2o = A(3, 512); T* x5 = A(3, 256);
    DC bot1 = dc_fwd("bot1", x4, 256, 256, 512, 3, b1o, 512, nullptr);
    DC bot2 = dc_fwd("bot2", b1o, 512, 512, 512, 3, b2o, 512, nullptr);
    DC bot3 = dc_fwd("bot3", b2o, 512, 512, 256, 3, x5, 256, nullptr);

    struct UpRec { DC dc1, dc2; SA sa; T* out; };
    struct UpCfg { int stage; const char* sa; const T* low; int c_low; int level; T* catbuf; };
    std::vector<UpRec> urec(3);
    const T* low = x5;
    T* const cats[3] = {cat1, cat2, cat3};
    const int c_lows[3] = {256, 128, 64};
    const char* sas[3] = {"sa4", "sa5", "sa6"};
    for (int i = 0; i < 3; ++i) {
      const StageInfo& st = kStages[3 + i];
      const int l = 2 - i;
      UpRec& r = urec[i];
      launch_upsample<T>(low, c_lows[i], cats[i], st.cin, B, p->levelH(l + 1), p->levelW(l + 1), c_lows[i], s);
      T* a = A(l, st.cin);
      r.dc1 = dc_fwd(std::string(st.name) + ".doubleConv1", cats[i], st.cin, st.cin, st.cin, l, a, st.cin, nullptr);
      T* cbuf = A(l, st.cout);
      r.dc2 = dc_fwd(std::string(st.name) + ".doubleConv2", a, st.cin, st.cin, st.cout, l, cbuf, st.cout, &st);
      if (att) {
        r.out = A(l, st.cout);
        r.sa = sa_fwd(sas[i], cbuf, st.cout, st.cout, l, r.out, st.cout);
      } else {
        r.out = cbuf;
      }
      low = r.out;
    }
    T* u3 = urec[2].out;
    launch_outc<T>(u3, 64, p->w_outc, p->b_outc, p->eps, B, p->H0, p->W0, 64, rows, dim, p->lh, p->lw, s);

    // ================= loss + backward =================
    T* d_u = A(0, 64);
    launch_mse_outc_bwd<T>(p->eps, noise, u3, 64, p->w_outc, d_u, G("outc.weight"), G("outc.bias"), tr->loss_dev, B, p->H0, p->W0, 64, rows,
                           dim, p->lh, p->lw, s, tr->valid);
    T* d_cats[3] = {nullptr, nullptr, nullptr};  // indexed like cats: cat1, cat2, cat3
    const T* d_cur = d_u;
    for (int i = 2; i >= 0; --i) {
      const StageInfo& st = kStages[3 + i];
      const int l = 2 - i;
      UpRec& r = urec[i];
      const T* d_c = att ? sa_bwd(r.sa, d_cur, st.cout) : d_cur;
      T* d_a = A(l, st.cin);
      dc_bwd(r.dc2, d_c, st.cout, d_a, st.cin);
      d_cats[i] = A(l, st.cin);
      dc_bwd(r.dc1, d_a, st.cin, d_cats[i], st.cin);
      T* d_low = A(l + 1, c_lows[i]);
      launch_upsample_bwd<T>(d_cats[i], st.cin, d_low, c_lows[i], B, p->levelH(l + 1), p->levelW(l + 1), c_lows[i], s);
      d_cur = d_low;
    }
    mark_phase(0);
    // d_cur = d x5
    T* d_b2o = A(3, 512); T* d_b1o = A(3, 512); T* d_x4 = A(3, 256);
    dc_bwd(bot3, d_cur, 256, d_b2o, 512);
    dc_bwd(bot2, d_b2o, 512, d_b1o, 512);
    dc_bwd(bot1, d_b1o, 512, d_x4, 256);
    d_cur = d_x4;
    for (int i = 2; i >= 0; --i) {
      const DownCfg& d = downs[i];
      const StageInfo& st = kStages[d.stage];
      const int l = d.level;
      DownRec& r = drec[i];
      const T* d_c = att ? sa_bwd(r.sa, d_cur, st.cout) : d_cur;
      T* d_b = A(l, st.cin);
      dc_bwd(r.dc2, d_c, st.cout, d_b, st.cin);
      T* d_pooled = A(l, st.cin);
      dc_bwd(r.dc1, d_b, st.cin, d_pooled, st.cin);
      // the skip connection: x_{i+1} also fed the concat of the matching up stage (cat index 2 - i), channels [c_low, 2 c_low)
      T* d_skipcat = d_cats[2 - i];
      const int cat_ld = kStages[3 + (2 - i)].cin, c_low = c_lows[2 - i];
      T* d_in = A(l - 1, st.cin);
      launch_pool_bwd<T>(r.pool_in, r.ld_pool_in, d_pooled, st.cin, d_skipcat + c_low, cat_ld, d_in, st.cin, B, p->levelH(l), p->levelW(l),
                         st.cin, s);
      d_cur = d_in;
    }
    dc_bwd(inc, d_cur, 64, nullptr, 0, x_noisy);
    mark_phase(1);

    // ---- time-embedding Linears (Unet_FiLmLayer.py:136-142,165-168) ----
    cudaStream_t s = fork_side();
    for (const StageInfo& st : kStages) {
      WgradArgs a{};
      a.x = silu_pe; a.ld_x = p->cfg.time_dim; a.dy = d_temb + st.temb_off; a.ld_dy = SPDM_TEMB_WIDTH; a.M = B; a.Cin = p->cfg.time_dim;
      a.Cout = st.cout; a.H = 1; a.W = 1; a.taps = 1; a.dw = G(std::string(st.name) + ".emb_layer.1.weight");
      launch_wgrad_simt<float, float>(a, s);
      launch_colsum<float>(d_temb + st.temb_off, SPDM_TEMB_WIDTH, B, st.cout, G(std::string(st.name) + ".emb_layer.1.bias"), s);
    }
  }
};
}  // namespace

// =================================================================================================
// C ABI — training
// =================================================================================================
extern "C" int spdm_train_enable(spdm_plan* p) {
  API_BEGIN
  REQUIRE(p && !p->sched_only, "bad plan");
  REQUIRE(!p->enc_resnet, "the native training step covers the autoencoder vision encoder; the ResNet18-GroupNorm encoder is inference-only here");
  REQUIRE(!p->tf32_mode, "the native training step runs in fp32 (CUDA cores) or bf16 (tensor cores); create the plan with one of them");
  REQUIRE(!p->simple, "the native training step covers the FiLM U-Nets; the simple U-Net (models/simple_Unet.py) is inference-only here");
  if (p->tr) return 0;
  TrainState* tr = new TrainState();
  p->tr = tr;
  if (const char* e = getenv("SPDM_WGRAD_SIMT")) tr->wgrad_simt = atoi(e) != 0;
  std::vector<std::string> names;
  for (auto& kv : p->gemms) names.push_back(kv.first);
  for (const std::string& n : names) {
    GemmW& g = p->gemms[n];
    GemmW& d = p->gemms[n + "#d"];
    d.Cin = g.Cout; d.Cout = g.Cin; d.taps = g.taps;
    const size_t cnt = (size_t)g.taps * g.Cin * g.Cout;
    if (p->bf16_mode) d.w16 = p->alloc<bf16>(cnt); else d.w32 = p->alloc<float>(cnt);
    p->gemms[n].twin = &d;
    if (g.taps == 9) { tr->packed_off[n] = tr->packed_elems; tr->packed_elems += cnt; }
  }
  if (p->bf16_mode) tr->packed = p->alloc<float>(tr->packed_elems);
  if (p->G > 0) tr->film_wT = p->alloc<float>((size_t)SPDM_FILM_WIDTH * p->G);
  if (const char* e = getenv("SPDM_FILM_SIMT")) tr->film_simt = atoi(e) != 0;
  if (p->G > 0 && p->bf16_mode && !tr->film_simt) {
    tr->GP = ((p->G + 63) / 64) * 64;
    tr->film_wf16 = p->alloc<bf16>((size_t)SPDM_FILM_WIDTH * tr->GP);   // zero-filled: the padding columns / rows stay zero
    tr->film_wb16 = p->alloc<bf16>((size_t)tr->GP * SPDM_FILM_WIDTH);
  }
  tr->enc_wlT = p->alloc<float>((size_t)128 * 9216);
  tr->loss_dev = p->alloc<float>(1);
  if (const char* e = getenv("SPDM_ENC_SIMT")) tr->enc_simt = atoi(e) != 0;
  if (const char* e = getenv("SPDM_TRAIN_SIDE")) tr->use_side = atoi(e) != 0;
  if (const char* e = getenv("SPDM_DGRAD_SWAP")) tr->dgrad_swap = atoi(e) != 0;
  CUDA_OK(cudaStreamCreateWithFlags(&tr->side, cudaStreamNonBlocking));
  CUDA_OK(cudaEventCreateWithFlags(&tr->ev_join, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&tr->ev_packs, cudaEventDisableTiming));
  for (int i = 0; i < 3; ++i) CUDA_OK(cudaEventCreateWithFlags(&tr->ev_phase[i], cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&tr->ev_packs_fork, cudaEventDisableTiming));
  if (p->bf16_mode && !tr->enc_simt) {
    tr->enc_wlT16 = p->alloc<bf16>((size_t)9216 * 128);
  }
  // every weight has to be uploaded again so that the twins are filled
  for (auto& kv : p->loaders) {
    if (kv.first.rfind("vision_encoder.", 0) == 0) p->missing_enc.insert(kv.first);
    else if (kv.first != "pos_encoding.inv_freq") p->missing_unet.insert(kv.first);
  }
  return 0;
  API_END
}

extern "C" int spdm_train_bind(spdm_plan* p, const char* name, int64_t offset, const int64_t* shape, int32_t ndim) {
  API_BEGIN
  REQUIRE(p && p->tr && name && shape && ndim > 0 && offset >= 0, "bad argument (spdm_train_enable first)");
  REQUIRE(p->loaders.count(name), "unknown parameter '%s'", name);
  TrainBind b;
  b.off = offset;
  b.shape.assign(shape, shape + ndim);
  p->tr->bind[name] = b;
  return 0;
  API_END
}

extern "C" int spdm_train_set_buffers(spdm_plan* p, float* params, float* grads, int64_t total) {
  API_BEGIN
  REQUIRE(p && p->tr && params && grads && total > 0, "bad argument");
  p->tr->params = params; p->tr->grads = grads; p->tr->total = total;
  return 0;
  API_END
}

// load_state_dict from the flat parameter buffer: every bound tensor is repacked into the kernel layouts (forward + twins)
extern "C" int spdm_train_sync_weights(spdm_plan* p, void* stream) {
  API_BEGIN
  REQUIRE(p && p->tr && p->tr->params, "training buffers not set");
  // The encoder's (few, small) tensors are repacked on the caller's stream; the U-Net's ~150 repack launches go to the side
  // stream, ordered after everything enqueued so far (the optimizer step), and the next training step only waits for them
  // after its encoder forward.  Any other consumer (sampling, unet_forward) waits at once, below.
  cudaStream_t main_s = (cudaStream_t)stream;
  cudaStream_t pack_s = p->tr->use_side ? p->tr->side : main_s;
  if (p->tr->use_side) {
    CUDA_OK(cudaEventRecord(p->tr->ev_packs_fork, main_s));
    CUDA_OK(cudaStreamWaitEvent(pack_s, p->tr->ev_packs_fork, 0));
  }
  for (auto& kv : p->tr->bind) {
    const TrainBind& b = kv.second;
    const bool enc = kv.first.rfind("vision_encoder.", 0) == 0;
    p->loaders[kv.first](p->tr->params + b.off, b.shape.data(), (int)b.shape.size(), enc ? main_s : pack_s);
    p->missing_unet.erase(kv.first);
    p->missing_enc.erase(kv.first);
  }
  if (p->tr->use_side) {
    CUDA_OK(cudaEventRecord(p->tr->ev_packs, pack_s));
    p->tr->packs_pending = true;
  }
  if (p->tr->enc_wlT16) {
    auto it = p->tr->bind.find("vision_encoder.7.weight");
    REQUIRE(it != p->tr->bind.end(), "training: parameter 'vision_encoder.7.weight' is not bound");
    launch_enc_pack_linear_T16(p->tr->params + it->second.off, p->tr->enc_wlT16, (cudaStream_t)stream);
  }
  p->temb_table_dirty = true;
  check_async("train_sync_weights");
  return 0;
  API_END
}

extern "C" int spdm_train_fwd_bwd(spdm_plan* p, const float* images, const float* position, const float* action, const float* velocity,
                                  const float* x0, const float* noise, const int64_t* t, const float* sqrt_ab, const float* sqrt_1mab,
                                  const float* inpaint, float* loss_out, int32_t B, void* stream) {
  API_BEGIN
  REQUIRE(p && p->tr && images && position && action && velocity && x0 && noise && t && sqrt_ab && sqrt_1mab && loss_out, "null argument");
  REQUIRE(B > 0 && B <= p->cfg.batch_max, "B=%d outside 1..batch_max=%d", B, p->cfg.batch_max);
  REQUIRE(p->tr->grads, "training buffers not set (spdm_train_set_buffers)");
  REQUIRE(p->cfg.cond_dim == 135 && p->G > 0, "training needs the conditional model (cond_dim == 135)");
  REQUIRE(B % p->bm == 0, "bf16 training: the batch (%d) must be a multiple of %d (pixel tiles of the weight-gradient GEMM hold whole samples); "
          "pad it and declare the real count with spdm_train_set_valid", B, p->bm);
  REQUIRE(p->tr->valid <= B, "spdm_train_set_valid(%d) exceeds the batch (%d)", p->tr->valid, B);
  check_ready(p, true);
  if (!p->missing_enc.empty()) throw SpdmError{"vision encoder weights missing, first: " + *p->missing_enc.begin()};
  TrainState* tr = p->tr;
  cudaStream_t s = (cudaStream_t)stream;
  const long long before = total_launches();
  arena_reset(p, B);
  tr->ev_cur = 0;
  CUDA_OK(cudaMemsetAsync(tr->grads, 0, (size_t)tr->total * sizeof(float), s));
  CUDA_OK(cudaMemsetAsync(tr->loss_dev, 0, sizeof(float), s));
  if (tr->packed) CUDA_OK(cudaMemsetAsync(tr->packed, 0, tr->packed_elems * sizeof(float), s));
  const int T = p->cfg.obs_horizon, n_frames = B * T;
  auto F = [&](size_t n) { return reinterpret_cast<float*>(arena_alloc(p, n * sizeof(float))); };

  // ---- q-sample + inpaint (ddpm:158-168) ----
  float* x_noisy = F((size_t)B * p->n_elems());
  launch_add_noise(x0, noise, reinterpret_cast<const long long*>(t), sqrt_ab, sqrt_1mab, p->cfg.inpaint_rows > 0 ? inpaint : nullptr, x_noisy,
                   p->n_elems(), p->cfg.inpaint_rows * p->cfg.dim, B, s);
  // ---- conditioning: encoder -> obs_cond -> Mish -> the six FiLM Linears (ddpm:317-330, Unet_FiLmLayer.py:149-154) ----
  const bool enc_tc = tr->enc_wlT16 != nullptr;
  const long long img_bstride = tr->image_bstride > 0 ? tr->image_bstride : (long long)T * 3 * 96 * 96;
  REQUIRE(enc_tc || img_bstride == (long long)T * 3 * 96 * 96, "a strided observation window needs the bf16 (tensor-core) encoder path");
  const long long n_pad = ((long long)n_frames + 127) / 128 * 128, M2 = (long long)n_frames * 576, M3 = (long long)n_frames * 144;
  auto H16 = [&](size_t n) { return reinterpret_cast<bf16*>(arena_alloc(p, n * sizeof(bf16))); };
  // flat tensor-core GEMM (taps = 1) on `rows` rows: out = epi(in @ w^T)
  auto tc_flat = [&](const char* tag, const bf16* in, int ld_in, const bf16* w, int Cin, int Cout, long long rows, bf16* out, int ld_out,
                     const float* bias, int flags, const bf16* mask_act = nullptr, int ld_mask = 0) {
    char key[96];
    snprintf(key, sizeof key, "enc|%s|%p|%lld", tag, (const void*)in, rows);
    TcGemm*& g = p->tc_cache[key];
    if (!g) {
      g = tc_gemm_create(in, ld_in, w, Cin, Cout, 1, 1, 1, (int)rows);
      REQUIRE(g != nullptr, "encoder GEMM %s: %s", tag, tc_last_error());
    }
    tc_gemm_launch(g, out, ld_out, nullptr, bias, mask_act, ld_mask, flags, (int)rows, s);
  };
  float* feat = nullptr;
  float* enc_out = F((size_t)n_frames * 128);
  bf16 *c1p = nullptr, *c2 = nullptr, *feat16 = nullptr;
  if (enc_tc) {
    REQUIRE(n_frames % 8 == 0, "bf16 training: B * obs_horizon (%d) must be a multiple of 8", n_frames);
    c1p = H16((size_t)M2 * 64);
    c2 = H16((size_t)M2 * 32);           // [M2/2][64] == [M3][128]
    feat16 = H16((size_t)n_pad * 9216);
    bf16* enc_out16 = H16((size_t)n_pad * 128);
    launch_enc_conv1_fwd(images, p->enc_w1, p->enc_b1, c1p, n_frames, T, img_bstride, s);
    tc_flat("conv2", c1p, 128, p->enc_w2p, 128, 64, M2 / 2, c2, 64, p->enc_b2p, EPI_BIAS | EPI_RELU);
    tc_flat("conv3", c2, 128, p->enc_w3p, 128, 64, M3, feat16, 64, p->enc_b3, EPI_BIAS | EPI_RELU);
    tc_flat("linear", feat16, 9216, p->enc_wl16, 9216, 128, n_pad, enc_out16, 128, p->enc_bl, EPI_BIAS);
    launch_cast_f32(enc_out16, enc_out, (long long)n_frames * 128, s);
  } else {
    feat = F((size_t)n_frames * 9216);
    launch_enc_convs<float>(images, p->enc_w1, p->enc_b1, p->enc_w2, p->enc_b2, p->enc_w3, p->enc_b3, feat, n_frames, s);
    GemmSimtArgs a{};
    a.in = feat; a.w = p->enc_wl; a.bias = p->enc_bl; a.out = enc_out; a.M = n_frames; a.Cin = 9216; a.Cout = 128;
    a.ld_in = 9216; a.ld_out = 128; a.H = 1; a.W = 1; a.taps = 1; a.act = ACT_NONE;
    launch_gemm_simt<float, float>(a, s);
  }
  launch_build_cond(position, action, velocity, enc_out, p->cond, B, T, p->cfg.cond_dim, s);
  if (tr->packs_pending) {  // the U-Net weights repacked on the side stream (spdm_train_sync_weights) are needed from here on
    CUDA_OK(cudaStreamWaitEvent(s, tr->ev_packs, 0));
    tr->packs_pending = false;
  }
  const bool film_tc = tr->film_wf16 != nullptr;
  const int GP = tr->GP;
  const int Bp = ((B + 127) / 128) * 128;                    // rows of the FiLM GEMMs (whole 128-row tiles)
  const int Bf = ((B + p->bm - 1) / p->bm) * p->bm;          // rows of d_film (Fwd::Bpad)
  bf16* cond_mish16 = nullptr;
  if (film_tc) {
    // the six FiLM Linears as one tensor-core GEMM [Bp][GP] x [GP][1792] (fp32: 160 us on the CUDA cores at batch 512)
    cond_mish16 = H16((size_t)Bp * GP);
    bf16* film16 = H16((size_t)Bp * SPDM_FILM_WIDTH);
    launch_mish_pad_bf16(p->cond, cond_mish16, B, Bp, p->G, GP, s);
    tc_flat("film", cond_mish16, GP, tr->film_wf16, GP, SPDM_FILM_WIDTH, Bp, film16, SPDM_FILM_WIDTH, p->film_b, EPI_BIAS);
    launch_cast_f32(film16, p->film, (long long)B * SPDM_FILM_WIDTH, s);
    p->have_cond = true;
  } else {
    compute_film(p, B, s);
  }
  // ---- time embedding rows (per sample) ----
  launch_temb(reinterpret_cast<const long long*>(t), B, p->inv_freq, p->temb_w, p->temb_b, p->temb_call, p->cfg.time_dim, s);
  float* silu_pe = F((size_t)B * p->cfg.time_dim);
  launch_posenc_silu(reinterpret_cast<const long long*>(t), B, p->inv_freq, silu_pe, p->cfg.time_dim, s);

  FwdCtx c{};
  c.x = x_noisy; c.out = p->eps; c.temb = p->temb_call; c.temb_mode = TEMB_PER_SAMPLE; c.step_ptr = nullptr; c.film = p->film; c.B = B; c.s = s;
  float* d_film = nullptr;
  if (p->bf16_mode) { Train<bf16> tr_(p, c); tr_.run(x_noisy, noise, silu_pe, reinterpret_cast<const long long*>(t)); d_film = tr_.d_film; }
  else { Train<float> tr_(p, c); tr_.run(x_noisy, noise, silu_pe, reinterpret_cast<const long long*>(t)); d_film = tr_.d_film; }

  auto G = [&](const std::string& name) {
    auto it = tr->bind.find(name);
    REQUIRE(it != tr->bind.end(), "training: parameter '%s' is not bound (spdm_train_bind)", name.c_str());
    return tr->grads + it->second.off;
  };
  // ---- FiLM Linears, d obs_cond, vision encoder ----
  auto fork = [&]() -> cudaStream_t {  // same contract as Train::fork_side
    if (!tr->use_side) return s;
    if (tr->ev_cur == tr->evs.size()) {
      cudaEvent_t e;
      CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      tr->evs.push_back(e);
    }
    cudaEvent_t e = tr->evs[tr->ev_cur++];
    CUDA_OK(cudaEventRecord(e, s));
    CUDA_OK(cudaStreamWaitEvent(tr->side, e, 0));
    return tr->side;
  };
  float* d_cond = F((size_t)B * p->G);
  if (film_tc) {
    bf16* d_film16 = H16((size_t)Bp * SPDM_FILM_WIDTH);
    launch_cast_bf16(d_film, d_film16, (long long)Bf * SPDM_FILM_WIDTH, s);   // rows B..Bf are zero (memset, never written)
    if (Bp > Bf) CUDA_OK(cudaMemsetAsync(d_film16 + (size_t)Bf * SPDM_FILM_WIDTH, 0, (size_t)(Bp - Bf) * SPDM_FILM_WIDTH * sizeof(bf16), s));
    // weight gradients of all six Linears in one pixel-contraction GEMM: dW [1792][GP] = d_film^T cond_mish, then row blocks -> parameters
    float* dw_all = F((size_t)SPDM_FILM_WIDTH * GP);
    cudaStream_t ws = fork();   // parameter gradients: side stream
    CUDA_OK(cudaMemsetAsync(dw_all, 0, (size_t)SPDM_FILM_WIDTH * GP * sizeof(float), ws));
    const int rc = wgrad_tc_launch(cond_mish16, GP, d_film16, SPDM_FILM_WIDTH, Bp, GP, SPDM_FILM_WIDTH, 1, 1, 1, dw_all, ws);
    REQUIRE(rc == 0, "FiLM wgrad: %s", wgrad_tc_last_error());
    for (const StageInfo& st : kStages) {
      CUDA_OK(cudaMemcpy2DAsync(G(std::string(st.name) + ".cond_encoder.2.weight"), (size_t)p->G * sizeof(float), dw_all + (size_t)st.film_off * GP,
                                (size_t)GP * sizeof(float), (size_t)p->G * sizeof(float), (size_t)2 * st.cout, cudaMemcpyDeviceToDevice, ws));
      launch_colsum<float>(d_film + st.film_off, SPDM_FILM_WIDTH, B, 2 * st.cout, G(std::string(st.name) + ".cond_encoder.2.bias"), ws);
    }
    bf16* d_cm16 = H16((size_t)Bp * GP);
    tc_flat("d_film", d_film16, SPDM_FILM_WIDTH, tr->film_wb16, SPDM_FILM_WIDTH, GP, Bp, d_cm16, GP, nullptr, 0);
    launch_mish_bwd_bf16(d_cm16, GP, p->cond, d_cond, B, p->G, s);
  } else {
    for (const StageInfo& st : kStages) {
      WgradArgs a{};
      a.x = p->cond_mish; a.ld_x = p->G; a.dy = d_film + st.film_off; a.ld_dy = SPDM_FILM_WIDTH; a.M = B; a.Cin = p->G; a.Cout = 2 * st.cout;
      a.H = 1; a.W = 1; a.taps = 1; a.dw = G(std::string(st.name) + ".cond_encoder.2.weight");
      launch_wgrad_simt<float, float>(a, s);
      launch_colsum<float>(d_film + st.film_off, SPDM_FILM_WIDTH, B, 2 * st.cout, G(std::string(st.name) + ".cond_encoder.2.bias"), s);
    }
    float* d_cond_mish = F((size_t)B * p->G);
    {
      GemmSimtArgs a{};
      a.in = d_film; a.w = tr->film_wT; a.out = d_cond_mish; a.M = B; a.Cin = SPDM_FILM_WIDTH; a.Cout = p->G;
      a.ld_in = SPDM_FILM_WIDTH; a.ld_out = p->G; a.H = 1; a.W = 1; a.taps = 1; a.act = ACT_NONE;
      launch_gemm_simt<float, float>(a, s);
    }
    launch_mish_bwd(d_cond_mish, p->cond, d_cond, (long long)B * p->G, s);
  }
  float* d_enc_out = F((size_t)n_frames * 128);
  launch_gather_feat_grad(d_cond, d_enc_out, B, T, p->cfg.cond_dim, s);
  if (enc_tc) {
    auto wg = [&](const bf16* x, int ld_x, const bf16* dy, int ld_dy, long long M, int Cin, int Cout, float* dst) {
      const int rc = wgrad_tc_launch(x, ld_x, dy, ld_dy, M, Cin, Cout, 1, 1, 1, dst, fork());
      REQUIRE(rc == 0, "encoder wgrad: %s", wgrad_tc_last_error());
    };
    bf16* d_eo16 = H16((size_t)n_pad * 128);
    launch_cast_bf16(d_enc_out, d_eo16, (long long)n_frames * 128, s);
    float* tmp = F((size_t)128 * 9216 + 64 * 128 + 64 * 128 + 64);  // hwc-ordered Linear grad | conv3 | block-diagonal conv2 | paired b2
    float *g3 = tmp + (size_t)128 * 9216, *g2 = g3 + 64 * 128, *gb2 = g2 + 64 * 128;
    CUDA_OK(cudaMemsetAsync(tmp, 0, ((size_t)128 * 9216 + 64 * 128 + 64 * 128 + 64) * sizeof(float), s));
    wg(feat16, 9216, d_eo16, 128, n_pad, 9216, 128, tmp);
    launch_enc_linear_grad_permute(tmp, G("vision_encoder.7.weight"), tr->use_side ? tr->side : s);
    launch_colsum<float>(d_enc_out, 128, n_frames, 128, G("vision_encoder.7.bias"), tr->use_side ? tr->side : s);
    bf16* d3 = H16((size_t)n_pad * 9216);   // d feat, then masked in place = gradient of the pre-ReLU conv3 output [M3][64]
    tc_flat("d_feat", d_eo16, 128, tr->enc_wlT16, 128, 9216, n_pad, d3, 9216, nullptr, 0);
    launch_relu_mask(d3, feat16, d3, M3 * 64, G("vision_encoder.4.bias"), s);
    wg(c2, 128, d3, 64, M3, 128, 64, g3);
    bf16* d2 = H16((size_t)M2 * 32);        // [M3][128] == [M2/2][64]
    tc_flat("d_c2", d3, 64, p->enc_w3pT, 64, 128, M3, d2, 128, nullptr, 0);
    launch_relu_mask(d2, c2, d2, M2 * 32, gb2, s);
    wg(c1p, 128, d2, 64, M2 / 2, 128, 64, g2);
    bf16* d1 = H16((size_t)M2 * 64);
    tc_flat("d_c1", d2, 64, p->enc_w2pT, 64, 128, M2 / 2, d1, 128, nullptr, 0);
    // (ReLU' in the GEMM epilogue -- EPI_MASK -- was measured slower: 457 us vs 182 + 130 us for GEMM + a streaming mask kernel);
    // d1 has a single consumer, the conv1 weight gradient, which applies the mask (c1p > 0) as it loads d1
    launch_enc_conv1_wgrad(images, d1, c1p, G("vision_encoder.0.weight"), G("vision_encoder.0.bias"), n_frames, T, img_bstride, s);
    {  // gb2 was accumulated on the main stream (relu mask), g2 / g3 on the side stream: unpack there, after this point of main
      cudaStream_t us = fork();
      launch_enc_unpack_grads(g2, g3, gb2, G("vision_encoder.2.weight"), G("vision_encoder.2.bias"), G("vision_encoder.4.weight"), us);
    }
  } else {
  {  // Linear(9216 -> 128): weight gradient in the kernel's hwc column order first, then permuted into (128, 9216 chw)
    float* tmp = F((size_t)128 * 9216);
    CUDA_OK(cudaMemsetAsync(tmp, 0, (size_t)128 * 9216 * sizeof(float), s));
    WgradArgs a{};
    a.x = feat; a.ld_x = 9216; a.dy = d_enc_out; a.ld_dy = 128; a.M = n_frames; a.Cin = 9216; a.Cout = 128; a.H = 1; a.W = 1; a.taps = 1; a.dw = tmp;
    launch_wgrad_simt<float, float>(a, s);
    launch_enc_linear_grad_permute(tmp, G("vision_encoder.7.weight"), s);
    launch_colsum<float>(d_enc_out, 128, n_frames, 128, G("vision_encoder.7.bias"), s);
  }
  float* d_feat = F((size_t)n_frames * 9216);
  {
    GemmSimtArgs a{};
    a.in = d_enc_out; a.w = tr->enc_wlT; a.out = d_feat; a.M = n_frames; a.Cin = 128; a.Cout = 9216;
    a.ld_in = 128; a.ld_out = 9216; a.H = 1; a.W = 1; a.taps = 1; a.act = ACT_NONE;
    launch_gemm_simt<float, float>(a, s);
  }
  launch_enc_convs_bwd(images, p->enc_w1, p->enc_b1, p->enc_w2, p->enc_b2, p->enc_w3, feat, d_feat, G("vision_encoder.0.weight"),
                       G("vision_encoder.0.bias"), G("vision_encoder.2.weight"), G("vision_encoder.2.bias"), G("vision_encoder.4.weight"),
                       G("vision_encoder.4.bias"), n_frames, s);
  }
  if (tr->use_side) {  // join: every weight gradient is in place from here on
    CUDA_OK(cudaEventRecord(tr->ev_join, tr->side));
    CUDA_OK(cudaStreamWaitEvent(s, tr->ev_join, 0));
  }
  CUDA_OK(cudaEventRecord(tr->ev_phase[2], s));
  CUDA_OK(cudaMemcpyAsync(loss_out, tr->loss_dev, sizeof(float), cudaMemcpyDeviceToDevice, s));
  p->launches += total_launches() - before;
  check_async("train_fwd_bwd");
  return 0;
  API_END
}

// clip_grad_norm_(max_norm) + Adam on flat buffers.  scratch: one device float.  max_norm <= 0: no clipping.
extern "C" int spdm_adam_step(float* params, float* grads, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                              int32_t step, float max_norm, float grad_scale, float* scratch, void* stream) {
  API_BEGIN
  REQUIRE(params && grads && m && v && n > 0 && step >= 1, "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  const float* sumsq = nullptr;
  if (max_norm > 0.f) {
    REQUIRE(scratch, "clipping needs a scratch float");
    CUDA_OK(cudaMemsetAsync(scratch, 0, sizeof(float), s));
    launch_sumsq(grads, n, scratch, s);
    sumsq = scratch;
  }
  launch_adam(params, grads, m, v, n, lr, beta1, beta2, eps, step, sumsq, max_norm, grad_scale, s);
  check_async("adam_step");
  return 0;
  API_END
}

// Makes `stream` wait for gradient-completion phase `phase` (0, 1, 2) of the most recent spdm_train_fwd_bwd: a communication
// stream calls this before all-reducing the matching slice of the flat gradient buffer, so the all-reduce of the up-path
// gradients runs under the rest of the backward pass.  Phase membership of a tensor: see engine.py::_grad_phase.
extern "C" int spdm_train_wait_phase(spdm_plan* p, int32_t phase, void* stream) {
  API_BEGIN
  REQUIRE(p && p->tr && phase >= 0 && phase < 3, "bad argument");
  CUDA_OK(cudaStreamWaitEvent((cudaStream_t)stream, p->tr->ev_phase[phase], 0));
  return 0;
  API_END
}

// A ragged batch (the last one of an epoch: the reference's DataLoader has no drop_last, utils/load_data.py:174) is padded by the
// caller to the tile granularity (spdm_plan_batch_multiple); only the first `valid` samples of the following spdm_train_fwd_bwd
// calls then count: the loss is the mean over them and the padding samples contribute no gradient.  0 = every sample is real.
extern "C" int spdm_train_set_valid(spdm_plan* p, int32_t valid) {
  API_BEGIN
  REQUIRE(p && p->tr && valid >= 0, "bad argument");
  p->tr->valid = valid;
  return 0;
  API_END
}

// The observation window handed to spdm_train_fwd_bwd may be a view of a longer recording: frame (b, t) then sits at
// images + b * stride + t * 3*96*96 (stride in floats; 0 = contiguous (B, T, 3, 96, 96)).  Saves the 566 MB slice copy per step
// that `batch['image'][:, :obs_horizon]` (models/diffusion_ddpm.py:283-298) otherwise costs.  bf16 path only.
extern "C" int spdm_train_set_image_stride(spdm_plan* p, int64_t stride) {
  API_BEGIN
  REQUIRE(p && p->tr && stride >= 0, "bad argument");
  p->tr->image_bstride = stride;
  return 0;
  API_END
}
