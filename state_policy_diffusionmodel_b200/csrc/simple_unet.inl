// simple_unet.inl — the legacy `UNet` of the reference (models/simple_Unet.py:260-339; the `model='UNet'` default of
// Diffusion_DDPM, models/diffusion_ddpm.py:60-62) on the fp32 CUDA-core path.  Included by plan.cu.
//
//   x1 = input_conv(x)                      DoubleConvolution(1, 16)                                  32x8  (default horizon)
//   x2 = down1(x1)  16 -> 32 (+32 cond)     pool, DC(16,16,residual), DC(16,32), +temb, cat cond      16x4
//   x3 = down2(x2)  64 -> 128 (+32)                                                                    8x2
//   x4 = down3(x3)  160 -> 256 (+32)                                                                   4x1
//   u  = up1(x4, x3) 448 -> 128 (+32) ; up2(u, x2) 224 -> 64 (+32) ; up3(u, x1) 112 -> 32 (+32) ; outc 64 -> 1
// Every DoubleConvolution is conv, GroupNorm(1, C), GELU, conv, GroupNorm (same module), [+ input], GELU (:104-125).  Skip tensors
// are written straight into their slice of the consumer's concat buffer, as on the FiLM path.
namespace {

struct SuStage { const char* name; int cin, cout, level, temb_off, idx; bool up; };
static const SuStage kSuStages[6] = {{"down1", 16, 32, 1, 0, 0, false},   {"down2", 64, 128, 2, 32, 1, false}, {"down3", 160, 256, 3, 160, 2, false},
                                     {"up1", 448, 128, 2, 416, 3, true},  {"up2", 224, 64, 1, 544, 4, true},   {"up3", 112, 32, 0, 608, 5, true}};
constexpr int SU_TEMB_WIDTH = 640;   // 32 + 128 + 256 + 128 + 64 + 32
constexpr int SU_COND_WIDTH = 192;   // 6 stages x 32

void register_weights_simple(spdm_plan* p) {
  const int TD = p->cfg.time_dim;
  p->w_in = p->alloc<float>(9 * 16);
  p->missing_unet.insert("input_conv.first.weight");
  {
    float* dst = p->w_in;
    p->loaders["input_conv.first.weight"] = [=](const float* src, const int64_t* shape, int ndim, cudaStream_t s) {
      check_shape("input_conv.first.weight", shape, ndim, {16, 1, 3, 3});
      launch_pack_conv_f32(src, dst, 16, 1, 3, s);
    };
  }
  reg_conv3(p, "input_conv.second", 16, 16);
  reg_norm(p, "input_conv.norm", 16);
  p->temb_w = p->alloc<float>((size_t)TD * SU_TEMB_WIDTH);
  p->temb_b = p->alloc<float>(SU_TEMB_WIDTH);
  p->film_w = p->alloc<float>((size_t)p->G * SU_COND_WIDTH);
  p->film_b = p->alloc<float>(SU_COND_WIDTH);
  for (const SuStage& st : kSuStages) {
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
        launch_pack_linear_f32(src, dst, C, TD, SU_TEMB_WIDTH, off, s);
      };
      reg_vec(p, n + ".emb_layer.1.bias", p->temb_b + off, C, p->missing_unet);
    }
    {
      const std::string wn = n + ".cond_emb_layer.1.weight";
      p->missing_unet.insert(wn);
      float* dst = p->film_w;
      const int off = st.idx * 32, G = p->G;
      p->loaders[wn] = [=](const float* src, const int64_t* shape, int ndim, cudaStream_t s) {
        check_shape(wn, shape, ndim, {32, G});
        launch_pack_linear_f32(src, dst, 32, G, SU_COND_WIDTH, off, s);
      };
      reg_vec(p, n + ".cond_emb_layer.1.bias", p->film_b + off, 32, p->missing_unet);
    }
  }
  p->w_outc = p->alloc<float>(64);
  p->b_outc = p->alloc<float>(1);
  reg_vec(p, "outc.weight", p->w_outc, 64, p->missing_unet);
  reg_vec(p, "outc.bias", p->b_outc, 1, p->missing_unet);
  // PositionalEncoding buffer (max_len = noise_steps + 1 rows of time_dim, :226-236): part of the state_dict
  p->missing_unet.insert("pos_encoding.pos_encoding");
  p->loaders["pos_encoding.pos_encoding"] = [=](const float* src, const int64_t* shape, int ndim, cudaStream_t s) {
    REQUIRE(ndim == 2 && shape[1] == TD && shape[0] > 0, "weight pos_encoding.pos_encoding: expected (max_len, %d)", TD);
    if (p->su_table_rows < shape[0]) { p->su_table = p->alloc<float>((size_t)shape[0] * TD); }
    p->su_table_rows = (int)shape[0];
    CUDA_OK(cudaMemcpyAsync(p->su_table, src, (size_t)shape[0] * TD * sizeof(float), cudaMemcpyDeviceToDevice, s));
  };
  // vision encoder: same as the FiLM variants
  p->enc_w1 = p->alloc<float>(16 * 3 * 4);   p->enc_b1 = p->alloc<float>(16);
  p->enc_w2 = p->alloc<float>(32 * 16 * 4);  p->enc_b2 = p->alloc<float>(32);
  p->enc_w3 = p->alloc<float>(64 * 32 * 4);  p->enc_b3 = p->alloc<float>(64);
  p->enc_wl = p->alloc<float>((size_t)9216 * 128);  p->enc_bl = p->alloc<float>(128);
  reg_vec(p, "vision_encoder.0.weight", p->enc_w1, 16 * 3 * 4, p->missing_enc);
  reg_vec(p, "vision_encoder.0.bias", p->enc_b1, 16, p->missing_enc);
  p->missing_enc.insert("vision_encoder.2.weight");
  {
    float* dst = p->enc_w2;
    p->loaders["vision_encoder.2.weight"] = [=](const float* src, const int64_t* shape, int ndim, cudaStream_t s) {
      check_shape("vision_encoder.2.weight", shape, ndim, {32, 16, 2, 2});
      launch_pack_linear_f32(src, dst, 32, 64, 32, 0, s);
    };
  }
  reg_vec(p, "vision_encoder.2.bias", p->enc_b2, 32, p->missing_enc);
  p->missing_enc.insert("vision_encoder.4.weight");
  {
    float* dst = p->enc_w3;
    p->loaders["vision_encoder.4.weight"] = [=](const float* src, const int64_t* shape, int ndim, cudaStream_t s) {
      check_shape("vision_encoder.4.weight", shape, ndim, {64, 32, 2, 2});
      launch_pack_linear_f32(src, dst, 64, 128, 64, 0, s);
    };
  }
  reg_vec(p, "vision_encoder.4.bias", p->enc_b3, 64, p->missing_enc);
  p->missing_enc.insert("vision_encoder.7.weight");
  {
    float* dst = p->enc_wl;
    p->loaders["vision_encoder.7.weight"] = [=](const float* src, const int64_t* shape, int ndim, cudaStream_t s) {
      check_shape("vision_encoder.7.weight", shape, ndim, {128, 9216});
      launch_pack_enc_linear(src, dst, s);
    };
  }
  reg_vec(p, "vision_encoder.7.bias", p->enc_bl, 128, p->missing_enc);
}

// cond_emb of the six stages: Linear(SiLU(cond)) (:144-148), all at once -> p->film [B][192]
void compute_cond_emb_simple(spdm_plan* p, int B, cudaStream_t s) {
  launch_su_silu(p->cond, p->cond_mish, (long long)B * p->G, s);
  GemmSimtArgs a{};
  a.in = p->cond_mish; a.w = p->film_w; a.bias = p->film_b; a.out = p->film; a.M = B; a.Cin = p->G; a.Cout = SU_COND_WIDTH;
  a.ld_in = p->G; a.ld_out = SU_COND_WIDTH; a.H = 1; a.W = 1; a.taps = 1; a.act = ACT_NONE;
  launch_gemm_simt<float, float>(a, s);
  p->have_cond = true;
}

struct SimpleFwd {
  spdm_plan* p;
  FwdCtx c;
  SimpleFwd(spdm_plan* p_, const FwdCtx& c_) : p(p_), c(c_) {}
  float* buf(void* v) const { return reinterpret_cast<float*>(v); }
  int HW(int l) const { return p->levelH(l) * p->levelW(l); }

  void conv(const std::string& wname, const float* in, int ld_in, int level, float* out) {
    auto it = p->gemms.find(wname);
    REQUIRE(it != p->gemms.end(), "internal: unknown gemm %s", wname.c_str());
    GemmW& g = it->second;
    GemmSimtArgs a{};
    a.in = in; a.w = g.w32; a.out = out; a.M = c.B * HW(level); a.Cin = g.Cin; a.Cout = g.Cout; a.ld_in = ld_in; a.ld_out = g.Cout;
    a.H = p->levelH(level); a.W = p->levelW(level); a.taps = 9; a.act = ACT_NONE;
    const double H = a.H, W = a.W;
    timed(p, c.s, PC_CONV3, 2.0 * g.Cin * g.Cout * (3.0 * H - 2) * (3.0 * W - 2) * c.B, ((double)a.M * (g.Cin + g.Cout) + 9.0 * g.Cin * g.Cout) * 4.0,
          [&] { launch_gemm_simt<float, float>(a, c.s); });
    timed(p, c.s, PC_STATS, 0, (double)a.M * g.Cout * 4.0, [&] { launch_stats<float>(out, p->stats, c.B, HW(level), g.Cout, g.Cout, c.s); });
  }
  void apply(const std::string& norm, const float* raw, int C, int level, float* out, int ld_out, const float* resid, int ld_res, int temb_off) {
    NormW& n = p->norms[norm];
    const float* temb = temb_off >= 0 ? c.temb + temb_off : nullptr;
    timed(p, c.s, PC_APPLY, 0, 2.0 * c.B * HW(level) * C * 4.0, [&] {
      launch_su_apply(raw, C, out, ld_out, p->stats, n.g, n.b, resid, ld_res, temb, SU_TEMB_WIDTH, temb ? c.temb_mode : TEMB_NONE, c.step_ptr, c.step_off,
                      c.B, HW(level), C, c.s);
    });
  }
  // DoubleConvolution(in, out, residual) (:82-125); `first_raw`: raw already holds the first conv's output + statistics
  void dc(const std::string& name, const float* in, int ld_in, int Cout, int level, float* out, int ld_out, bool residual, int temb_off,
          bool first_raw = false) {
    float* raw = buf(p->raw[level]);
    float* h = buf(p->hbuf[level]);
    if (!first_raw) conv(name + ".first", in, ld_in, level, raw);
    apply(name + ".norm", raw, Cout, level, h, Cout, nullptr, 0, -1);
    conv(name + ".second", h, Cout, level, raw);
    apply(name + ".norm", raw, Cout, level, out, ld_out, residual ? in : nullptr, ld_in, temb_off);
  }
  void cond_slot(const SuStage& st, float* out, int ld_out) {   // channels [cout, cout + 32) <- cond_emb of the stage
    timed(p, c.s, PC_RESAMPLE, 0, (double)c.B * HW(st.level) * 32 * 4.0,
          [&] { launch_su_bcast(c.film + st.idx * 32, SU_COND_WIDTH, out + st.cout, ld_out, c.B, HW(st.level), c.s); });
  }

  void run() {
    REQUIRE(c.film != nullptr, "the simple U-Net needs conditioning: its channel counts include the 32-channel cond_emb of every stage "
                               "(models/simple_Unet.py:268-276)");
    REQUIRE(p->su_table != nullptr, "weight pos_encoding.pos_encoding missing");
    const int rows = p->cfg.rows, dim = p->cfg.dim;
    float* cat0 = buf(p->cat[0]); float* cat1 = buf(p->cat[1]); float* cat2 = buf(p->cat[2]);
    // ---- input_conv -> x1 (16 ch) in the skip slot of up3's concat buffer [96 up | 16 x1] ----
    timed(p, c.s, PC_IO, 2.0 * 9 * 16 * HW(0) * c.B, (double)c.B * HW(0) * 16 * 4.0,
          [&] { launch_su_conv_in(c.x, p->w_in, buf(p->raw[0]), c.B, p->H0, p->W0, rows, dim, p->lh, p->lw, c.s); });
    timed(p, c.s, PC_STATS, 0, (double)c.B * HW(0) * 16 * 4.0, [&] { launch_stats<float>(buf(p->raw[0]), p->stats, c.B, HW(0), 16, 16, c.s); });
    dc("input_conv", nullptr, 0, 16, 0, cat0 + 96, 112, false, -1, true);
    // ---- down path: x2 -> cat1 + 160 (ld 224), x3 -> cat2 + 288 (ld 448), x4 -> abuf[3] (ld 288) ----
    struct Down { const float* in; int ld_in; float* dest; int ld_dest; };
    const Down downs[3] = {{cat0 + 96, 112, cat1 + 160, 224}, {cat1 + 160, 224, cat2 + 288, 448}, {cat2 + 288, 448, buf(p->abuf[3]), 288}};
    for (int i = 0; i < 3; ++i) {
      const SuStage& st = kSuStages[i];
      const int l = st.level;
      float* a = buf(p->abuf[l]); float* b = buf(p->bbuf[l]);
      timed(p, c.s, PC_RESAMPLE, 0, 5.0 * c.B * HW(l) * st.cin * 4.0,
            [&] { launch_pool<float>(downs[i].in, downs[i].ld_in, a, st.cin, c.B, p->levelH(l), p->levelW(l), st.cin, c.s); });
      dc(std::string(st.name) + ".doubleConv1", a, st.cin, st.cin, l, b, st.cin, true, -1);
      // (down3 writes x4 back into abuf[3]: the pooled map there is dead once doubleConv1 has read it as its residual)
      dc(std::string(st.name) + ".doubleConv2", b, st.cin, st.cout, l, downs[i].dest, downs[i].ld_dest, false, st.temb_off);
      cond_slot(st, downs[i].dest, downs[i].ld_dest);
    }
    // ---- up path ----
    struct Up { const float* low; int c_low; float* catbuf; float* dest; int ld_dest; };
    const Up ups[3] = {{buf(p->abuf[3]), 288, cat2, buf(p->bbuf[2]), 160}, {buf(p->bbuf[2]), 160, cat1, buf(p->bbuf[1]), 96},
                       {buf(p->bbuf[1]), 96, cat0, buf(p->bbuf[0]), 64}};
    for (int i = 0; i < 3; ++i) {
      const SuStage& st = kSuStages[3 + i];
      const int l = st.level;
      timed(p, c.s, PC_RESAMPLE, 0, 1.25 * c.B * HW(l) * ups[i].c_low * 4.0, [&] {
        launch_upsample<float>(ups[i].low, ups[i].c_low, ups[i].catbuf, st.cin, c.B, p->levelH(l + 1), p->levelW(l + 1), ups[i].c_low, c.s);
      });
      float* a = buf(p->abuf[l]);
      dc(std::string(st.name) + ".doubleConv1", ups[i].catbuf, st.cin, st.cin, l, a, st.cin, true, -1);
      dc(std::string(st.name) + ".doubleConv2", a, st.cin, st.cout, l, ups[i].dest, ups[i].ld_dest, false, st.temb_off);
      cond_slot(st, ups[i].dest, ups[i].ld_dest);
    }
    // ---- outc + unpad ----
    if (c.fuse_step) {
      timed(p, c.s, PC_STEP, 2.0 * 64 * rows * dim * c.B, (double)c.B * HW(0) * 64 * 4.0, [&] {
        launch_outc_step<float>(*c.fuse_step, buf(p->bbuf[0]), 64, p->w_outc, p->b_outc, p->H0, p->W0, 64, rows, dim, p->lh, p->lw, c.s);
      });
      return;
    }
    timed(p, c.s, PC_IO, 2.0 * 64 * rows * dim * c.B, (double)c.B * HW(0) * 64 * 4.0, [&] {
      launch_outc<float>(buf(p->bbuf[0]), 64, p->w_outc, p->b_outc, c.out, c.B, p->H0, p->W0, 64, rows, dim, p->lh, p->lw, c.s);
    });
  }
};

}  // namespace
