// resnet.inl — the ResNet18-GroupNorm vision encoder behind `vision_encoder='resnet18'` (reference models/Unet_FiLmLayer.py:316-386:
// `VisionEncoder()` = torchvision resnet18, fc = Identity, BatchNorm2d -> GroupNorm(C / 16, C)).  Included by plan.cu; kernels in
// resnet.cu.  512 features per 96x96 frame => cond_dim = 2 + 3 + 2 + 512 = 519.  Inference only.
namespace {

struct RnBlock { const char* name; int cin, cout, stride; bool down; };
static const RnBlock kRnBlocks[8] = {{"layer1.0", 64, 64, 1, false},  {"layer1.1", 64, 64, 1, false},   {"layer2.0", 64, 128, 2, true},
                                     {"layer2.1", 128, 128, 1, false}, {"layer3.0", 128, 256, 2, true}, {"layer3.1", 256, 256, 1, false},
                                     {"layer4.0", 256, 512, 2, true},  {"layer4.1", 512, 512, 1, false}};
constexpr int RN_CHUNK = 256;   // frames per pass (multiple of 128: every patch matrix is then whole 128-row tiles)

void rn_reg_vec(spdm_plan* p, const std::string& name, float* dst, int n) { reg_vec(p, name, dst, n, p->missing_enc); }

void rn_reg_conv(spdm_plan* p, const std::string& name, int Cin, int Cout, int k) {
  GemmW& g = p->gemms[name];
  g.Cin = k * k * Cin; g.Cout = Cout; g.taps = 1;
  if (p->bf16_mode) g.w16 = p->alloc<bf16>((size_t)g.Cin * Cout); else g.w32 = p->alloc<float>((size_t)g.Cin * Cout);
  const std::string wn = name + ".weight";
  p->missing_enc.insert(wn);
  GemmW* gp = &g;
  const bool bfm = p->bf16_mode;
  p->loaders[wn] = [=](const float* src, const int64_t* shape, int ndim, cudaStream_t s) {
    check_shape(wn, shape, ndim, {Cout, Cin, k, k});
    if (bfm) launch_pack_conv_bf16(src, gp->w16, Cout, Cin, k, s);    // [Cout][tap][Cin]: K-major B operand, K = (kh, kw, c)
    else launch_pack_conv_f32(src, gp->w32, Cout, Cin, k, s);         // [tap][Cin][Cout]
  };
}
void rn_reg_norm(spdm_plan* p, const std::string& name, int C) {
  NormW& n = p->norms[name];
  n.C = C;
  n.g = p->alloc<float>(C);
  n.b = p->alloc<float>(C);
  rn_reg_vec(p, name + ".weight", n.g, C);
  rn_reg_vec(p, name + ".bias", n.b, C);
}

void register_weights_resnet(spdm_plan* p) {
  const std::string v = "vision_encoder.";
  {
    GemmW& g = p->gemms[v + "conv1"];
    g.Cin = 192; g.Cout = 64; g.taps = 1;                              // K = 7*7*3 = 147, zero-padded to 3 x 64
    if (p->bf16_mode) g.w16 = p->alloc<bf16>(64 * 192); else g.w32 = p->alloc<float>(64 * 192);
    const std::string wn = v + "conv1.weight";
    p->missing_enc.insert(wn);
    GemmW* gp = &g;
    p->loaders[wn] = [=](const float* src, const int64_t* shape, int ndim, cudaStream_t s) {
      check_shape(wn, shape, ndim, {64, 3, 7, 7});
      launch_rn_pack_conv1(src, gp->w16, gp->w32, s);
    };
  }
  rn_reg_norm(p, v + "bn1", 64);
  for (const RnBlock& b : kRnBlocks) {
    const std::string n = v + b.name;
    rn_reg_conv(p, n + ".conv1", b.cin, b.cout, 3);
    rn_reg_norm(p, n + ".bn1", b.cout);
    rn_reg_conv(p, n + ".conv2", b.cout, b.cout, 3);
    rn_reg_norm(p, n + ".bn2", b.cout);
    if (b.down) {
      rn_reg_conv(p, n + ".downsample.0", b.cin, b.cout, 1);
      rn_reg_norm(p, n + ".downsample.1", b.cout);
    }
  }
}

template <typename T> struct RnFwd {
  spdm_plan* p;
  cudaStream_t s;
  T *col, *a0, *x, *y, *z, *d;

  void gemm(const std::string& name, long long rows, T* out) {
    GemmW& g = p->gemms[name];
    if constexpr (sizeof(T) == 2) {
      TcGemm*& tc = p->tc_cache["rn|" + name];
      if (!tc) {
        // capacity = this layer's rows for a full chunk (rows shrink with the spatial size, the K-major patch matrix starts at col)
        tc = tc_gemm_create(reinterpret_cast<const bf16*>(col), g.Cin, g.w16, g.Cin, g.Cout, 1, 1, 1, (int)p->rn_rows_cap[name]);
        REQUIRE(tc != nullptr, "%s: %s", name.c_str(), tc_last_error());
      }
      const int P = tc_gemm_launch(tc, reinterpret_cast<bf16*>(out), g.Cout, nullptr, nullptr, nullptr, 0, 0, (int)rows, s);
      REQUIRE(P >= 0, "%s: %s", name.c_str(), tc_last_error());
    } else {
      GemmSimtArgs a{};
      a.in = col; a.w = g.w32; a.out = out; a.M = (int)rows; a.Cin = g.Cin; a.Cout = g.Cout; a.ld_in = g.Cin; a.ld_out = g.Cout;
      a.H = 1; a.W = 1; a.taps = 1; a.act = ACT_NONE;
      launch_gemm_simt<float, float>(a, s);
    }
  }

  // frames [f0, f0 + m) of `img` (fp32 NCHW) -> out[f][512]
  void run(const float* img, float* out, int m) {
    const std::string v = "vision_encoder.";
    const long long mp = ((long long)m + 127) / 128 * 128;
    launch_rn_im2col_img<T>(img, col, mp, m, s);
    gemm(v + "conv1", mp * 2304, a0);
    NormW& n1 = p->norms[v + "bn1"];
    launch_rn_gn<T>(a0, a0, p->rn_stats, n1.g, n1.b, nullptr, 1, mp, 2304, 64, s);
    launch_rn_maxpool<T>(a0, x, mp, 48, 48, 64, s);
    int H = 24;
    for (const RnBlock& b : kRnBlocks) {
      const std::string n = v + b.name;
      const int Ho = H / b.stride;
      const long long rows = mp * Ho * Ho;
      launch_rn_im2col<T>(x, col, mp, H, H, b.cin, 3, b.stride, 1, s);
      gemm(n + ".conv1", rows, y);
      NormW& g1 = p->norms[n + ".bn1"];
      launch_rn_gn<T>(y, y, p->rn_stats, g1.g, g1.b, nullptr, 1, mp, Ho * Ho, b.cout, s);
      launch_rn_im2col<T>(y, col, mp, Ho, Ho, b.cout, 3, 1, 1, s);
      gemm(n + ".conv2", rows, z);
      const T* resid = x;
      if (b.down) {
        launch_rn_im2col<T>(x, col, mp, H, H, b.cin, 1, b.stride, 0, s);
        gemm(n + ".downsample.0", rows, d);
        NormW& gd = p->norms[n + ".downsample.1"];
        launch_rn_gn<T>(d, d, p->rn_stats, gd.g, gd.b, nullptr, 0, mp, Ho * Ho, b.cout, s);
        resid = d;
      }
      NormW& g2 = p->norms[n + ".bn2"];
      launch_rn_gn<T>(z, y, p->rn_stats, g2.g, g2.b, resid, 1, mp, Ho * Ho, b.cout, s);   // relu(bn2(conv2) + identity) -> y
      T* t = x; x = y; y = t;
      H = Ho;
    }
    launch_rn_avgpool<T>(x, out, m, 9, 512, s);
  }
};

template <typename T> void resnet_encode(spdm_plan* p, const float* images, float* out, int n, cudaStream_t s) {
  if (!p->rn_col) {
    const size_t c = RN_CHUNK;
    p->rn_col = p->alloc<T>(c * 2304 * 192);
    p->rn_a0 = p->alloc<T>(c * 2304 * 64);
    for (int i = 0; i < 4; ++i) p->rn_buf[i] = p->alloc<T>(c * 576 * 64);
    p->rn_stats = p->alloc<float>(c * 32 * 2);
    p->rn_rows_cap["vision_encoder.conv1"] = (long long)c * 2304;
    int H = 24;
    for (const RnBlock& b : kRnBlocks) {
      const int Ho = H / b.stride;
      const std::string nm = std::string("vision_encoder.") + b.name;
      p->rn_rows_cap[nm + ".conv1"] = p->rn_rows_cap[nm + ".conv2"] = p->rn_rows_cap[nm + ".downsample.0"] = (long long)c * Ho * Ho;
      H = Ho;
    }
  }
  for (int f0 = 0; f0 < n; f0 += RN_CHUNK) {
    const int m = n - f0 < RN_CHUNK ? n - f0 : RN_CHUNK;
    RnFwd<T> f{p, s, reinterpret_cast<T*>(p->rn_col), reinterpret_cast<T*>(p->rn_a0), reinterpret_cast<T*>(p->rn_buf[0]),
               reinterpret_cast<T*>(p->rn_buf[1]), reinterpret_cast<T*>(p->rn_buf[2]), reinterpret_cast<T*>(p->rn_buf[3])};
    f.run(images + (size_t)f0 * 3 * 96 * 96, out + (size_t)f0 * 512, m);
  }
}

}  // namespace
