// extern "C" surface of libbirefnet_b200.so (include/birefnet_b200.h).  Nothing throws across this boundary.
#include <cmath>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "model.h"

using namespace brn;

static thread_local std::string g_err;

template <class F>
static brn_status guard(F&& f) {
  try {
    f();
    return BRN_OK;
  } catch (const Error& e) {
    g_err = e.what();
    return (brn_status)e.status;
  } catch (const std::exception& e) {
    g_err = std::string("internal error: ") + e.what();
    return BRN_ERR_INVALID;
  } catch (...) {
    g_err = "unknown internal error";
    return BRN_ERR_INVALID;
  }
}

struct brn_model {
  Model impl;
  brn_model(const brn_config& c, int dev) : impl(c, dev) {}
};
struct brn_sharded {          // defined in sharded.cpp; the ABI only needs the handle list
  std::vector<Model*> models;
  std::vector<int> devices;
  ~brn_sharded();
};

extern "C" {

const char* brn_last_error(void) { return g_err.c_str(); }
const char* brn_version(void) { return "birefnet_b200 0.1 (sm_100a)"; }

void brn_config_swin_l(brn_config* cfg) {
  if (!cfg) return;
  memset(cfg, 0, sizeof(*cfg));
  cfg->embed_dim = 192;
  const int d[4] = {2, 2, 18, 2}, h[4] = {6, 12, 24, 48};
  for (int i = 0; i < 4; ++i) { cfg->depths[i] = d[i]; cfg->num_heads[i] = h[i]; }
  cfg->window_size = 12; cfg->mlp_ratio = 4; cfg->patch_size = 4;
  cfg->precision = BRN_PREC_FP16; cfg->deform_mode = BRN_DEFORM_DEFORMABLE; cfg->micro_batch = 0;
}

void brn_config_swin_b(brn_config* cfg) {
  if (!cfg) return;
  brn_config_swin_l(cfg);
  cfg->embed_dim = 128;
  const int h[4] = {4, 8, 16, 32};
  for (int i = 0; i < 4; ++i) cfg->num_heads[i] = h[i];
}

// SwinConfig::swin_t / swin_s (src/swin.rs:27-52): embed 96, heads 3/6/12/24, window 7 (49-token windows)
void brn_config_swin_t(brn_config* cfg) {
  if (!cfg) return;
  brn_config_swin_l(cfg);
  cfg->embed_dim = 96;
  const int d[4] = {2, 2, 6, 2}, h[4] = {3, 6, 12, 24};
  for (int i = 0; i < 4; ++i) { cfg->depths[i] = d[i]; cfg->num_heads[i] = h[i]; }
  cfg->window_size = 7;
}
void brn_config_swin_s(brn_config* cfg) {
  if (!cfg) return;
  brn_config_swin_t(cfg);
  cfg->depths[2] = 18;
}

brn_status brn_model_create(const brn_config* cfg, int device, brn_model** out) {
  return guard([&] {
    BRN_CHECK(cfg && out, 1, "brn_model_create: null argument");
    *out = new brn_model(*cfg, device);
  });
}

void brn_model_destroy(brn_model* m) { delete m; }

brn_status brn_model_set_tensor(brn_model* m, const char* key, const void* data, int dtype, const int64_t* shape,
                                int rank) {
  return guard([&] {
    BRN_CHECK(m, 1, "null model");
    std::lock_guard<std::mutex> lk(m->impl.mu);
    m->impl.set_tensor(key, data, dtype, shape, rank);
  });
}

int32_t brn_model_num_tensors(const brn_model* m) { return m ? (int32_t)m->impl.keys.size() : 0; }

brn_status brn_model_tensor_info(const brn_model* m, int32_t i, const char** key, int64_t shape[4], int32_t* rank) {
  return guard([&] {
    BRN_CHECK(m && key && shape && rank, 1, "null argument");
    BRN_CHECK(i >= 0 && i < (int32_t)m->impl.keys.size(), 1, "tensor index out of range");
    *key = m->impl.keys[i].c_str();
    const auto& s = m->impl.tensors[i].shape;
    *rank = (int32_t)s.size();
    for (size_t d = 0; d < 4; ++d) shape[d] = d < s.size() ? s[d] : 1;
  });
}

brn_status brn_model_load_safetensors(brn_model* m, const char* path, int32_t* n_loaded) {
  return guard([&] {
    BRN_CHECK(m && path, 1, "null argument");
    std::lock_guard<std::mutex> lk(m->impl.mu);
    const int n = load_safetensors(m->impl, path);
    if (n_loaded) *n_loaded = n;
  });
}

brn_status brn_model_finalize(brn_model* m) {
  return guard([&] {
    BRN_CHECK(m, 1, "null model");
    std::lock_guard<std::mutex> lk(m->impl.mu);
    m->impl.finalize();
  });
}

brn_status brn_model_set_precision(brn_model* m, int precision) {
  return guard([&] {
    BRN_CHECK(m, 1, "null model");
    BRN_CHECK(precision == BRN_PREC_FP32 || precision == BRN_PREC_BF16 || precision == BRN_PREC_FP16, 1, "bad precision");
    std::lock_guard<std::mutex> lk(m->impl.mu);
    m->impl.cfg.precision = precision;
  });
}

brn_status brn_model_set_deform_mode(brn_model* m, int mode) {
  return guard([&] {
    BRN_CHECK(m, 1, "null model");
    BRN_CHECK(mode == BRN_DEFORM_CPU_FALLBACK || mode == BRN_DEFORM_DEFORMABLE, 1, "bad deform mode");
    std::lock_guard<std::mutex> lk(m->impl.mu);
    m->impl.cfg.deform_mode = mode;
  });
}

brn_status brn_model_set_cuda_graph(brn_model* m, int on) {
  return guard([&] {
    BRN_CHECK(m, 1, "null model");
    std::unique_lock<std::mutex> lk(m->impl.mu);
    m->impl.quiesce(lk);
    m->impl.use_graph = on ? 1 : 0;
    if (!on) m->impl.drop_graphs();
  });
}

brn_status brn_forward_logits(brn_model* m, const float* x, int32_t B, int32_t H, int32_t W, int x_is_device,
                              float* out, int out_is_device, void* stream) {
  return guard([&] {
    BRN_CHECK(m, 1, "null model");
    m->impl.forward(x, B, H, W, x_is_device != 0, out, out_is_device != 0, (cudaStream_t)stream, false);
  });
}

brn_status brn_forward(brn_model* m, const float* x, int32_t B, int32_t H, int32_t W, int x_is_device, float* out,
                       int out_is_device, void* stream) {
  return guard([&] {
    BRN_CHECK(m, 1, "null model");
    m->impl.forward(x, B, H, W, x_is_device != 0, out, out_is_device != 0, (cudaStream_t)stream, true);
  });
}

brn_status brn_backbone_forward(brn_model* m, const float* x, int32_t B, int32_t H, int32_t W, int x_is_device,
                                float* const outs[4], int out_is_device, void* stream) {
  return guard([&] {
    BRN_CHECK(m && x && outs, 1, "null argument");
    m->impl.backbone_api(x, B, H, W, x_is_device != 0, outs, out_is_device != 0, (cudaStream_t)stream);
  });
}

brn_status brn_features_forward(brn_model* m, const float* x, int32_t B, int32_t H, int32_t W, int x_is_device,
                                float* const outs[4], int out_is_device, void* stream) {
  return guard([&] {
    BRN_CHECK(m && x && outs, 1, "null argument");
    m->impl.features_api(x, B, H, W, x_is_device != 0, outs, out_is_device != 0, (cudaStream_t)stream);
  });
}

brn_status brn_decoder_forward(brn_model* m, const float* x, const float* x1, const float* x2, const float* x3,
                               const float* x4, int32_t B, int32_t H, int32_t W, int is_device, float* out,
                               void* stream) {
  return guard([&] {
    BRN_CHECK(m && x && x1 && x2 && x3 && x4 && out, 1, "null argument");
    m->impl.decoder_api(x, x1, x2, x3, x4, B, H, W, is_device != 0, out, (cudaStream_t)stream);
  });
}

// ---------------------------------------------------------------------------------------------------------------
// single-process image sharding: one handle + one host thread per GPU
// ---------------------------------------------------------------------------------------------------------------
brn_status brn_sharded_create(const brn_config* cfg, const int32_t* devices, int32_t n_devices, brn_sharded** out) {
  return guard([&] {
    BRN_CHECK(cfg && devices && out, 1, "brn_sharded_create: null argument");
    std::vector<int> d(devices, devices + (n_devices > 0 ? n_devices : 0));
    *out = sharded_create(*cfg, d.data(), (int)d.size());
  });
}
void brn_sharded_destroy(brn_sharded* s) { delete s; }
int32_t brn_sharded_num_devices(const brn_sharded* s) { return s ? (int32_t)s->models.size() : 0; }
brn_status brn_sharded_set_tensor(brn_sharded* s, const char* key, const void* data, int dtype, const int64_t* shape,
                                  int rank) {
  return guard([&] { BRN_CHECK(s, 1, "null handle"); sharded_set_tensor(s, key, data, dtype, shape, rank); });
}
brn_status brn_sharded_load_safetensors(brn_sharded* s, const char* path, int32_t* n_loaded) {
  return guard([&] {
    BRN_CHECK(s && path, 1, "null argument");
    const int n = sharded_load_safetensors(s, path);
    if (n_loaded) *n_loaded = n;
  });
}
brn_status brn_sharded_finalize(brn_sharded* s) {
  return guard([&] { BRN_CHECK(s, 1, "null handle"); sharded_finalize(s); });
}
brn_status brn_sharded_forward_logits(brn_sharded* s, const float* x, int32_t B, int32_t H, int32_t W, float* out) {
  return guard([&] { BRN_CHECK(s, 1, "null handle"); sharded_forward(s, x, B, H, W, out, false); });
}
brn_status brn_sharded_forward(brn_sharded* s, const float* x, int32_t B, int32_t H, int32_t W, float* out) {
  return guard([&] { BRN_CHECK(s, 1, "null handle"); sharded_forward(s, x, B, H, W, out, true); });
}
// pinned, portable host memory: buffers every GPU of a sharded handle can copy from / to at full PCIe rate
void* brn_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void brn_host_free(void* p) { if (p) cudaFreeHost(p); }

int64_t brn_launch_count(const brn_model* m) { return m ? m->impl.launches : 0; }
void brn_launch_count_reset(brn_model* m) { if (m) m->impl.launches = 0; }
void brn_profile_enable(brn_model* m, int on) { if (m) m->impl.prof_on = on; }
int32_t brn_kernel_class_times(const brn_model* m, float* ms, double* flops, double* bytes, int32_t* counts, int32_t cap) {
  if (!m) return 0;
  const int n = cap < (int)KC_COUNT ? cap : (int)KC_COUNT;
  for (int c = 0; c < n; ++c) {
    if (ms) ms[c] = m->impl.kc_ms[c];
    if (flops) flops[c] = m->impl.kc_flops[c];
    if (bytes) bytes[c] = m->impl.kc_bytes[c];
    if (counts) counts[c] = m->impl.kc_count[c];
  }
  return (int32_t)KC_COUNT;
}
int32_t brn_profile_get(const brn_model* m, const char*** names, const float** ms, const double** flops) {
  if (!m) return 0;
  if (names) *names = const_cast<const char**>(m->impl.prof_names.data());
  if (ms) *ms = m->impl.prof_ms.data();
  if (flops) *flops = m->impl.prof_flops.data();
  return (int32_t)m->impl.prof_names.size();
}

// ---------------------------------------------------------------------------------------------------------------
// operator level: host buffers in, host buffers out; device scratch is allocated per call
// ---------------------------------------------------------------------------------------------------------------
struct Scratch {
  std::vector<void*> ptrs;
  cudaStream_t stream = nullptr;
  int prev_dev = -1;
  explicit Scratch(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    BRN_CHECK(e == cudaSuccess && n > 0, 2, "no CUDA device available (this library has no CPU fallback)");
    BRN_CHECK(device >= 0 && device < n, 1, "device index out of range");
    if (cudaGetDevice(&prev_dev) != cudaSuccess || prev_dev == device) prev_dev = -1;
    BRN_CUDA(cudaSetDevice(device));
    BRN_CUDA(cudaStreamCreate(&stream));
  }
  ~Scratch() {
    for (void* p : ptrs) cudaFree(p);
    if (stream) cudaStreamDestroy(stream);
    if (prev_dev >= 0) cudaSetDevice(prev_dev);   // the caller's current device is left as it was
  }
  void* alloc(size_t bytes) {
    void* p = nullptr;
    BRN_CUDA(cudaMalloc(&p, bytes ? bytes : 16));
    ptrs.push_back(p);
    return p;
  }
  float* put(const float* h, size_t n) {
    float* d = (float*)alloc(n * 4);
    BRN_CUDA(cudaMemcpyAsync(d, h, n * 4, cudaMemcpyHostToDevice, stream));
    return d;
  }
};

static LaunchCtx make_ctx(Scratch& s, int precision) {
  LaunchCtx c; c.stream = s.stream; c.precision = precision; c.dry = false; c.launches = nullptr;
  const char* v = getenv("BRN_FORCE_SIMT");
  c.force_simt = v && v[0] && v[0] != '0';
  return c;
}

brn_status brn_linear(int device, int precision, const float* a, const float* w, const float* bias,
                      const float* residual, int32_t M, int32_t N, int32_t K, int32_t act, float* out) {
  return guard([&] {
    BRN_CHECK(a && w && out && M > 0 && N > 0 && K > 0, 1, "brn_linear: bad argument");
    Scratch s(device);
    LaunchCtx ctx = make_ctx(s, precision);
    const int AD = precision == BRN_PREC_FP16 ? F16 : precision == BRN_PREC_BF16 ? BF16 : F32;
    LayerW L = make_layer_standalone(N, K, 1, 1, w, bias, s.ptrs);
    View a32 = make_view(s.put(a, (size_t)M * K), F32, 1, 1, M, K);
    View ax = a32;
    if (AD != F32) { ax = make_view(s.alloc((size_t)M * K * 2), AD, 1, 1, M, K); glue_copy_cast(ctx, a32, ax); }
    View o = make_view(s.alloc((size_t)M * N * 4), F32, 1, 1, M, N);
    GemmArgs g; g.x = ax; g.w = &L; g.act = act; g.out = o;
    if (residual) g.res = make_view(s.put(residual, (size_t)M * N), F32, 1, 1, M, N);
    op_gemm(ctx, g);
    BRN_CUDA(cudaMemcpyAsync(out, o.p, (size_t)M * N * 4, cudaMemcpyDeviceToHost, s.stream));
    BRN_CUDA(cudaStreamSynchronize(s.stream));
  });
}

brn_status brn_ln_linear(int device, int precision, const float* x, const float* gamma, const float* beta,
                         const float* w, const float* bias, int32_t M, int32_t N, int32_t K, int32_t act, float* out) {
  return guard([&] {
    BRN_CHECK(x && gamma && beta && w && out && M > 0 && N > 0 && K > 0, 1, "brn_ln_linear: bad argument");
    BRN_CHECK(precision != BRN_PREC_FP32, 7, "brn_ln_linear: the LayerNorm fold is a tensor-core-path feature");
    Scratch s(device);
    LaunchCtx ctx = make_ctx(s, precision);
    const int AD = precision == BRN_PREC_FP16 ? F16 : BF16;
    // fold exactly as Model::make_folded does
    std::vector<float> wf((size_t)N * K), bf(N);
    for (int n = 0; n < N; ++n) {
      double acc = bias ? bias[n] : 0.0;
      for (int c = 0; c < K; ++c) {
        wf[(size_t)n * K + c] = (float)((double)w[(size_t)n * K + c] * gamma[c]);
        acc += (double)w[(size_t)n * K + c] * beta[c];
      }
      bf[n] = (float)acc;
    }
    LayerW L = make_layer_standalone(N, K, 1, 1, wf.data(), bf.data(), s.ptrs, true);
    View x32 = make_view(s.put(x, (size_t)M * K), F32, 1, 1, M, K);
    View x16 = make_view(s.alloc((size_t)M * K * 2), AD, 1, 1, M, K);
    float2* stats = (float2*)s.alloc((size_t)M * sizeof(float2));
    glue_ln_stats_cast(ctx, x32, x16, stats);
    float2* mr = (float2*)s.alloc((size_t)M * sizeof(float2));
    glue_ln_finalize(ctx, stats, 1, M, M, K, mr);
    View o16 = make_view(s.alloc((size_t)M * N * 2), AD, 1, 1, M, N);
    GemmArgs g; g.x = x16; g.w = &L; g.act = act; g.out = o16;
    g.lnf.mr = mr; g.lnf.C = K;
    BRN_CHECK(tc_gemm_supported(g), 5, "brn_ln_linear: K must be a multiple of 8");
    tc_gemm(ctx, g);
    View o32 = make_view(s.alloc((size_t)M * N * 4), F32, 1, 1, M, N);
    glue_copy_cast(ctx, o16, o32);
    BRN_CUDA(cudaMemcpyAsync(out, o32.p, (size_t)M * N * 4, cudaMemcpyDeviceToHost, s.stream));
    BRN_CUDA(cudaStreamSynchronize(s.stream));
  });
}

// Swin MLP sub-block (src/swin.rs:103-107 inside :407): out = x + fc2(gelu_erf(fc1(LayerNorm(x)))) on the tensor-core
// path with the LayerNorm folded.  fused: 1 = the single-kernel path (mlp_tcgen05.cu; C in {128, 192}, hidden = 4C),
// 0 = two GEMMs, -1 = whatever the model would run.  out_mean_rstd (optional, [M, 2]): the (mean, rstd) the epilogue's
// emitted statistics give for the rows of `out` -- what the NEXT block's folded norm1 would consume.
brn_status brn_swin_mlp(int device, int precision, const float* x, const float* gamma, const float* beta,
                        const float* w1, const float* b1, const float* w2, const float* b2, int32_t M, int32_t C,
                        int32_t hidden, int32_t fused, float* out, float* out_mean_rstd) {
  return guard([&] {
    BRN_CHECK(x && gamma && beta && w1 && w2 && out && M > 0 && C > 0 && hidden > 0, 1, "brn_swin_mlp: bad argument");
    BRN_CHECK(precision != BRN_PREC_FP32, 7, "brn_swin_mlp: the LayerNorm fold is a tensor-core-path feature");
    BRN_CHECK(C % 16 == 0 && hidden % 8 == 0, 5, "brn_swin_mlp: C % 16 == 0 and hidden % 8 == 0");
    Scratch s(device);
    LaunchCtx ctx = make_ctx(s, precision);
    const int AD = precision == BRN_PREC_FP16 ? F16 : BF16;
    std::vector<float> wf((size_t)hidden * C), bf(hidden);
    for (int n = 0; n < hidden; ++n) {
      double acc = b1 ? b1[n] : 0.0;
      for (int c = 0; c < C; ++c) {
        wf[(size_t)n * C + c] = (float)((double)w1[(size_t)n * C + c] * gamma[c]);
        acc += (double)w1[(size_t)n * C + c] * beta[c];
      }
      bf[n] = (float)acc;
    }
    LayerW L1 = make_layer_standalone(hidden, C, 1, 1, wf.data(), bf.data(), s.ptrs, true);
    LayerW L2 = make_layer_standalone(C, hidden, 1, 1, w2, b2, s.ptrs);
    View xt = make_view(s.put(x, (size_t)M * C), F32, 1, 1, M, C);
    View x16 = make_view(s.alloc((size_t)M * C * 2), AD, 1, 1, M, C);
    const int parts = tc_gemm_ln_parts(C);
    float2* stats = (float2*)s.alloc((size_t)parts * M * sizeof(float2));
    float2* mr = (float2*)s.alloc((size_t)M * sizeof(float2));
    glue_ln_stats_cast(ctx, xt, x16, stats);
    glue_ln_finalize(ctx, stats, 1, M, M, C, mr);
    MlpArgs ml; ml.x16 = x16; ml.mr = mr; ml.fc1 = &L1; ml.fc2 = &L2; ml.xt = xt;
    ml.lne.stats = stats; ml.lne.stride = M; ml.lne.x16 = x16.p; ml.lne.x16dt = AD; ml.lne.ldx16 = C; ml.mr_out = mr;
    const bool can = hidden == 4 * C && tc_mlp_supported(ml);
    BRN_CHECK(fused != 1 || can, 7, "brn_swin_mlp: the fused kernel needs C in {128, 192} and hidden = 4C");
    if (fused != 0 && can) {
      tc_mlp(ctx, ml);
    } else {
      View hd = make_view(s.alloc((size_t)M * hidden * 2), AD, 1, 1, M, hidden);
      GemmArgs g; g.x = x16; g.w = &L1; g.act = ACT_GELU; g.out = hd; g.lnf.mr = mr; g.lnf.C = C;
      BRN_CHECK(tc_gemm_supported(g), 5, "brn_swin_mlp: fc1 shape unsupported");
      tc_gemm(ctx, g);
      GemmArgs g2; g2.x = hd; g2.w = &L2; g2.out = xt; g2.res = xt; g2.lne = ml.lne;
      BRN_CHECK(tc_gemm_supported(g2), 5, "brn_swin_mlp: fc2 shape unsupported");
      tc_gemm(ctx, g2);
      glue_ln_finalize(ctx, stats, parts, M, M, C, mr);
    }
    BRN_CUDA(cudaMemcpyAsync(out, xt.p, (size_t)M * C * 4, cudaMemcpyDeviceToHost, s.stream));
    std::vector<float2> h;
    if (out_mean_rstd) {
      h.resize(M);
      BRN_CUDA(cudaMemcpyAsync(h.data(), mr, (size_t)M * sizeof(float2), cudaMemcpyDeviceToHost, s.stream));
    }
    BRN_CUDA(cudaStreamSynchronize(s.stream));
    if (out_mean_rstd)
      for (int m = 0; m < M; ++m) { out_mean_rstd[2 * m] = -h[m].x; out_mean_rstd[2 * m + 1] = h[m].y; }
  });
}

brn_status brn_conv2d(int device, int precision, const float* x, const float* weight, const float* bias, int32_t B,
                      int32_t C, int32_t H, int32_t W, int32_t O, int32_t k, int32_t act, float* out) {
  return guard([&] {
    BRN_CHECK(x && weight && out && B > 0 && C > 0 && H > 0 && W > 0 && O > 0 && (k & 1), 1, "brn_conv2d: bad argument");
    Scratch s(device);
    LaunchCtx ctx = make_ctx(s, precision);
    const int AD = precision == BRN_PREC_FP16 ? F16 : precision == BRN_PREC_BF16 ? BF16 : F32;
    LayerW L = make_layer_standalone(O, C, k, k, weight, bias, s.ptrs);
    float* dx = s.put(x, (size_t)B * C * H * W);
    View xv = make_view(s.alloc((size_t)B * C * H * W * dsize(AD)), AD, B, H, W, C);
    glue_nchw_to_nhwc(ctx, dx, B, C, H, W, xv);
    View o = make_view(s.alloc((size_t)B * O * H * W * 4), F32, B, H, W, O);
    GemmArgs g; g.x = xv; g.w = &L; g.pad = k / 2; g.act = act; g.out = o;
    op_gemm(ctx, g);
    float* on = (float*)s.alloc((size_t)B * O * H * W * 4);
    glue_nhwc_to_nchw(ctx, o, on);
    BRN_CUDA(cudaMemcpyAsync(out, on, (size_t)B * O * H * W * 4, cudaMemcpyDeviceToHost, s.stream));
    BRN_CUDA(cudaStreamSynchronize(s.stream));
  });
}

brn_status brn_deform_conv2d(int device, int precision, const float* x, const float* offset, const float* mask,
                             const float* weight, const float* bias, int32_t B, int32_t C, int32_t H, int32_t W,
                             int32_t O, int32_t k, int32_t stride, int32_t padding, float* out) {
  return guard([&] {
    BRN_CHECK(x && offset && mask && weight && out && k > 0 && stride > 0 && padding >= 0, 1, "brn_deform_conv2d: bad argument");
    const int Ho = (H + 2 * padding - k) / stride + 1, Wo = (W + 2 * padding - k) / stride + 1;
    BRN_CHECK(Ho > 0 && Wo > 0, 5, "brn_deform_conv2d: empty output");
    Scratch s(device);
    LaunchCtx ctx = make_ctx(s, precision);
    const int AD = precision == BRN_PREC_FP16 ? F16 : precision == BRN_PREC_BF16 ? BF16 : F32;
    const int taps = k * k;
    LayerW L = make_layer_standalone(O, C, k, k, weight, bias, s.ptrs);
    float* dx = s.put(x, (size_t)B * C * H * W);
    View xv = make_view(s.alloc((size_t)B * C * H * W * dsize(AD)), AD, B, H, W, C);
    glue_nchw_to_nhwc(ctx, dx, B, C, H, W, xv);
    // offsets (2*taps channels) and modulators (taps channels), at the OUTPUT resolution -> one NHWC fp32 tensor [.., 3*taps]
    View om = make_view(s.alloc((size_t)B * Ho * Wo * 3 * taps * 4), F32, B, Ho, Wo, 3 * taps);
    glue_nchw_to_nhwc(ctx, s.put(offset, (size_t)B * 2 * taps * Ho * Wo), B, 2 * taps, Ho, Wo, om.slice(0, 2 * taps));
    glue_nchw_to_nhwc(ctx, s.put(mask, (size_t)B * taps * Ho * Wo), B, taps, Ho, Wo, om.slice(2 * taps, taps));
    View o = make_view(s.alloc((size_t)B * O * Ho * Wo * 4), F32, B, Ho, Wo, O);
    DeformArgs d; d.x = xv; d.om = om; d.w = &L; d.out = o; d.stride = stride; d.pad = padding;
    op_deform(ctx, d);
    float* on = (float*)s.alloc((size_t)B * O * Ho * Wo * 4);
    glue_nhwc_to_nchw(ctx, o, on);
    BRN_CUDA(cudaMemcpyAsync(out, on, (size_t)B * O * Ho * Wo * 4, cudaMemcpyDeviceToHost, s.stream));
    BRN_CUDA(cudaStreamSynchronize(s.stream));
  });
}

// DeformableConv2d::new + forward (src/deform_conv.rs:29-99): the module owns offset_conv, modulator_conv and
// regular_conv (all k x k, same stride / padding, with bias); offset = offset_conv(x), modulator = 2 sigmoid(
// modulator_conv(x)) (:83-86); then the modulated deformable conv (Metal path, :102-215) or -- deform_mode
// CPU_FALLBACK, what candle computes on Device::Cpu (:95-98) -- plain regular_conv(x).
brn_status brn_deformable_conv2d(int device, int precision, int deform_mode, const float* x, int32_t B, int32_t C, int32_t H,
                                 int32_t W, const float* offset_w, const float* offset_b, const float* modulator_w,
                                 const float* modulator_b, const float* regular_w, const float* regular_b, int32_t O,
                                 int32_t k, int32_t stride, int32_t padding, float* out) {
  return guard([&] {
    BRN_CHECK(x && offset_w && offset_b && modulator_w && modulator_b && regular_w && out, 1, "brn_deformable_conv2d: null argument");
    BRN_CHECK(B > 0 && C > 0 && O > 0 && k > 0 && stride > 0 && padding >= 0, 1, "brn_deformable_conv2d: bad argument");
    BRN_CHECK(deform_mode == BRN_DEFORM_CPU_FALLBACK || deform_mode == BRN_DEFORM_DEFORMABLE, 1, "bad deform mode");
    const int Ho = (H + 2 * padding - k) / stride + 1, Wo = (W + 2 * padding - k) / stride + 1;
    BRN_CHECK(Ho > 0 && Wo > 0, 5, "brn_deformable_conv2d: empty output");
    Scratch s(device);
    LaunchCtx ctx = make_ctx(s, precision);
    const int AD = precision == BRN_PREC_FP16 ? F16 : precision == BRN_PREC_BF16 ? BF16 : F32;
    const int taps = k * k;
    float* dx = s.put(x, (size_t)B * C * H * W);
    View xv = make_view(s.alloc((size_t)B * C * H * W * dsize(AD)), AD, B, H, W, C);
    glue_nchw_to_nhwc(ctx, dx, B, C, H, W, xv);
    LayerW Lr = make_layer_standalone(O, C, k, k, regular_w, regular_b, s.ptrs);
    View o = make_view(s.alloc((size_t)B * O * Ho * Wo * 4), F32, B, Ho, Wo, O);
    if (deform_mode == BRN_DEFORM_CPU_FALLBACK) {
      GemmArgs g; g.x = xv; g.w = &Lr; g.pad = padding; g.stride = stride; g.out = o;
      op_gemm(ctx, g);
    } else {
      // offset_conv ++ modulator_conv share input and geometry: one conv with 3 k^2 outputs, 2 sigmoid on the tail
      std::vector<float> wom((size_t)3 * taps * C * taps), bom((size_t)3 * taps);
      memcpy(wom.data(), offset_w, (size_t)2 * taps * C * taps * 4);
      memcpy(wom.data() + (size_t)2 * taps * C * taps, modulator_w, (size_t)taps * C * taps * 4);
      memcpy(bom.data(), offset_b, (size_t)2 * taps * 4);
      memcpy(bom.data() + 2 * taps, modulator_b, (size_t)taps * 4);
      LayerW Lom = make_layer_standalone(3 * taps, C, k, k, wom.data(), bom.data(), s.ptrs);
      View om = make_view(s.alloc((size_t)B * Ho * Wo * 3 * taps * 4), F32, B, Ho, Wo, 3 * taps);
      GemmArgs g; g.x = xv; g.w = &Lom; g.pad = padding; g.stride = stride; g.act = ACT_2SIGMOID_TAIL; g.act_from = 2 * taps;
      g.out = om;
      op_gemm(ctx, g);
      DeformArgs d; d.x = xv; d.om = om; d.w = &Lr; d.out = o; d.stride = stride; d.pad = padding;
      op_deform(ctx, d);
    }
    float* on = (float*)s.alloc((size_t)B * O * Ho * Wo * 4);
    glue_nhwc_to_nchw(ctx, o, on);
    BRN_CUDA(cudaMemcpyAsync(out, on, (size_t)B * O * Ho * Wo * 4, cudaMemcpyDeviceToHost, s.stream));
    BRN_CUDA(cudaStreamSynchronize(s.stream));
  });
}

brn_status brn_window_attention(int device, int precision, const float* qkv, const float* bias, int32_t n_windows,
                                int32_t heads, int32_t window_size, int32_t hp, int32_t wp, int32_t shift, float* out) {
  return guard([&] {
    BRN_CHECK(qkv && bias && out && n_windows > 0 && heads > 0, 1, "brn_window_attention: bad argument");
    const int ws = window_size, wn = ws * ws;
    BRN_CHECK(ws == 12 || ws == 7, 7, "window_size must be 12 or 7");
    BRN_CHECK(hp % ws == 0 && wp % ws == 0 && (shift == 0 || shift == ws / 2), 5,
              "hp, wp must be multiples of window_size; shift 0 | window_size / 2");
    const int nw = (hp / ws) * (wp / ws);
    BRN_CHECK(n_windows % nw == 0, 5, "n_windows must be a multiple of (hp/ws)*(wp/ws)");
    Scratch s(device);
    LaunchCtx ctx = make_ctx(s, precision);
    const int AD = precision == BRN_PREC_FP16 ? F16 : precision == BRN_PREC_BF16 ? BF16 : F32;
    const int C = heads * 32;
    const size_t rows = (size_t)n_windows * wn;
    // fold the q scale the way finalize does (src/swin.rs:278)
    std::vector<float> hq(qkv, qkv + rows * 3 * C);
    const float sc = 0.17677669529663687f;
    for (size_t r = 0; r < rows; ++r)
      for (int c = 0; c < C; ++c) hq[r * 3 * C + c] *= sc;
    View q32 = make_view(s.put(hq.data(), hq.size()), F32, 1, 1, (int)rows, 3 * C);
    View qx = q32;
    if (AD != F32) { qx = make_view(s.alloc(rows * 3 * C * 2), AD, 1, 1, (int)rows, 3 * C); glue_copy_cast(ctx, q32, qx); }
    float* b32 = s.put(bias, (size_t)heads * wn * wn);
    // padded fp32 copy [heads][144][148] for the tcgen05 kernel
    std::vector<float> padded((size_t)heads * wn * (wn + 4), 0.f);
    for (size_t r = 0; r < (size_t)heads * wn; ++r) memcpy(&padded[r * (wn + 4)], &bias[r * wn], (size_t)wn * 4);
    float* b32p = s.put(padded.data(), padded.size());
    View o = make_view(s.alloc(rows * C * dsize(AD)), AD, 1, 1, (int)rows, C);
    AttnArgs a; a.qkv = qx; a.bias32 = b32; a.bias32p = b32p; a.n_windows = n_windows;
    a.heads = heads; a.nwh = hp / ws; a.nww = wp / ws; a.shift = shift; a.ws = ws; a.out = o;
    op_attention(ctx, a);
    View o32 = o;
    if (AD != F32) { o32 = make_view(s.alloc(rows * C * 4), F32, 1, 1, (int)rows, C); glue_copy_cast(ctx, o, o32); }
    BRN_CUDA(cudaMemcpyAsync(out, o32.p, rows * C * 4, cudaMemcpyDeviceToHost, s.stream));
    BRN_CUDA(cudaStreamSynchronize(s.stream));
  });
}

// ---------------------------------------------------------------------------------------------------------------
// pre / post-processing: the steps either side of forward_logits in examples/infer_image.rs
// ---------------------------------------------------------------------------------------------------------------
brn_status brn_preprocess_rgb8(int device, const uint8_t* rgb, int32_t B, int32_t h, int32_t w, int32_t H, int32_t W,
                               float* out) {
  return guard([&] {
    BRN_CHECK(rgb && out && B > 0 && h > 0 && w > 0 && H > 0 && W > 0, 1, "brn_preprocess_rgb8: bad argument");
    Scratch s(device);
    const size_t nin = (size_t)B * h * w * 3, nout = (size_t)B * 3 * H * W;
    uint8_t* din = (uint8_t*)s.alloc(nin);
    BRN_CUDA(cudaMemcpyAsync(din, rgb, nin, cudaMemcpyHostToDevice, s.stream));
    float* tmp = (float*)s.alloc((size_t)B * H * w * 3 * 4);
    float* dout = (float*)s.alloc(nout * 4);
    prepost_preprocess(s.stream, din, B, h, w, H, W, tmp, dout);
    BRN_CUDA(cudaMemcpyAsync(out, dout, nout * 4, cudaMemcpyDeviceToHost, s.stream));
    BRN_CUDA(cudaStreamSynchronize(s.stream));
  });
}

brn_status brn_postprocess_mask(int device, const float* logits, int32_t B, int32_t H, int32_t W, int32_t orig_h,
                                int32_t orig_w, uint8_t* out) {
  return guard([&] {
    BRN_CHECK(logits && out && B > 0 && H > 0 && W > 0 && orig_h > 0 && orig_w > 0, 1, "brn_postprocess_mask: bad argument");
    Scratch s(device);
    float* dl = s.put(logits, (size_t)B * H * W);
    uint8_t* m8 = (uint8_t*)s.alloc((size_t)B * H * W);
    float* tmp = (float*)s.alloc((size_t)B * orig_h * W * 4);
    uint8_t* dout = (uint8_t*)s.alloc((size_t)B * orig_h * orig_w);
    prepost_postprocess(s.stream, dl, 0, B, H, W, orig_h, orig_w, m8, tmp, dout);
    BRN_CUDA(cudaMemcpyAsync(out, dout, (size_t)B * orig_h * orig_w, cudaMemcpyDeviceToHost, s.stream));
    BRN_CUDA(cudaStreamSynchronize(s.stream));
  });
}

brn_status brn_infer_rgb8(brn_model* m, const uint8_t* rgb, int32_t B, int32_t h, int32_t w, int32_t H, int32_t W,
                          uint8_t* masks) {
  return guard([&] {
    BRN_CHECK(m && rgb && masks && B > 0 && h > 0 && w > 0, 1, "brn_infer_rgb8: bad argument");
    BRN_CHECK(H > 0 && W > 0 && H % 32 == 0 && W % 32 == 0, 5, "H and W must be positive multiples of 32");
    Scratch s(m->impl.device);
    const size_t nin = (size_t)B * h * w * 3;
    uint8_t* din = (uint8_t*)s.alloc(nin);
    BRN_CUDA(cudaMemcpyAsync(din, rgb, nin, cudaMemcpyHostToDevice, s.stream));
    const size_t tmp_floats = std::max((size_t)B * H * w * 3, (size_t)B * h * W);
    float* tmp = (float*)s.alloc(tmp_floats * 4);
    float* dx = (float*)s.alloc((size_t)B * 3 * H * W * 4);
    float* dl = (float*)s.alloc((size_t)B * H * W * 4);
    uint8_t* m8 = (uint8_t*)s.alloc((size_t)B * H * W);
    uint8_t* dout = (uint8_t*)s.alloc((size_t)B * h * w);
    prepost_preprocess(s.stream, din, B, h, w, H, W, tmp, dx);
    m->impl.forward(dx, B, H, W, true, dl, true, s.stream, false);     // enqueued on the same stream
    prepost_postprocess(s.stream, dl, 0, B, H, W, h, w, m8, tmp, dout);
    BRN_CUDA(cudaMemcpyAsync(masks, dout, (size_t)B * h * w, cudaMemcpyDeviceToHost, s.stream));
    BRN_CUDA(cudaStreamSynchronize(s.stream));
  });
}

// ---------------------------------------------------------------------------------------------------------------
// kernel micro-benchmarks on device-resident synthetic data (scripts/kernel_bench.py; ncu -k targets)
// ---------------------------------------------------------------------------------------------------------------
__global__ void fill_kernel(void* p, int dt, long long n, unsigned seed, float scale) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned h = (unsigned)(i * 2654435761u) ^ seed;
  h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
  float v = ((h & 0xffffff) / 8388608.0f - 1.0f) * scale;
  if (dt == F32) ((float*)p)[i] = v;
  else if (dt == BF16) ((__nv_bfloat16*)p)[i] = __float2bfloat16(v);
  else ((__half*)p)[i] = __float2half_rn(v);
}
static void fill(Scratch& s, void* p, int dt, long long n, unsigned seed, float scale) {
  fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s.stream>>>(p, dt, n, seed, scale);
}

// kind 0: conv/linear GEMM  x[B,H,W,C] * w[N,k,k,C] (act, optional fp32 residual, out dtype odt: 0 f32 / 1 = act dtype)
// kind 1: window attention   (B = n_windows, C = heads)
// kind 2: deformable conv    x[B,H,W,64] -> N, kernel k (offsets ~ N(0, 2 px))
// Returns the mean device time per launch in *ms_out (CUDA events around `iters` back-to-back launches).
brn_status brn_bench_op(int device, int precision, int kind, int32_t B, int32_t H, int32_t W, int32_t C, int32_t N,
                        int32_t k, int32_t act, int32_t with_res, int32_t out_f32, int32_t iters, float* ms_out) {
  return guard([&] {
    BRN_CHECK(ms_out && iters > 0, 1, "brn_bench_op: bad argument");
    Scratch s(device);
    LaunchCtx ctx = make_ctx(s, precision);
    const int AD = precision == BRN_PREC_FP16 ? F16 : precision == BRN_PREC_BF16 ? BF16 : F32;
    cudaEvent_t e0, e1;
    BRN_CUDA(cudaEventCreate(&e0)); BRN_CUDA(cudaEventCreate(&e1));
    std::function<void()> launch;
    LayerW L{}, L2{}; GemmArgs g; AttnArgs at; DeformArgs d; MlpArgs ml; View hd;
    const size_t px = (size_t)B * H * W;
    if (kind == 0 || kind == 2) {
      const int taps = k * k;
      std::vector<float> hw((size_t)N * C * taps), hb(N);
      for (size_t i = 0; i < hw.size(); ++i) hw[i] = (float)((int)((i * 2654435761u) >> 20 & 1023) - 512) / (512.0f * sqrtf((float)C * taps));
      for (int i = 0; i < N; ++i) hb[i] = 0.01f * (i % 7);
      L = make_layer_standalone(N, C, k, k, hw.data(), hb.data(), s.ptrs);
      View x = make_view(s.alloc(px * C * dsize(AD)), AD, B, H, W, C);
      fill(s, x.p, AD, (long long)px * C, 1u, 1.0f);
      const int OD = out_f32 ? F32 : AD;
      View o = make_view(s.alloc(px * N * dsize(OD)), OD, B, H, W, N);
      if (kind == 0) {
        g.x = x; g.w = &L; g.pad = k / 2; g.act = act; g.out = o;
        if (with_res == 2) {          // 16-bit residual, separate 16-bit output (decoder laterals)
          g.res = make_view(s.alloc(px * N * dsize(AD)), AD, B, H, W, N); fill(s, g.res.p, AD, (long long)px * N, 2u, 1.0f);
        } else if (with_res) {        // fp32 residual stream updated in place (backbone proj / fc2)
          g.res = make_view(s.alloc(px * N * 4), F32, B, H, W, N); fill(s, g.res.p, F32, (long long)px * N, 2u, 1.0f);
          g.out = make_view(g.res.p, F32, B, H, W, N);
        }
        launch = [&] { op_gemm(ctx, g); };
      } else {
        View om = make_view(s.alloc(px * 3 * taps * 4), F32, B, H, W, 3 * taps);
        fill(s, om.p, F32, (long long)px * 3 * taps, 3u, 2.0f);
        d.x = x; d.om = om; d.w = &L; d.act = act; d.out = o;
        // random offsets are layout-agnostic: read them tile-major, as the model does (tc_gemm out_tiled)
        d.om_tiled = (H % 8 == 0 && W % 16 == 0 && precision != BRN_PREC_FP32) ? 1 : 0;
        launch = [&] { op_deform(ctx, d); };
      }
    } else if (kind == 3) {
      // Swin MLP half-block on M = B*H*W rows of width C: with_res = 1 the fused kernel, 0 fc1 + fc2 through HBM
      const int hid = 4 * C;
      std::vector<float> hw((size_t)hid * C), hb(hid), hw2((size_t)C * hid), hb2(C);
      for (size_t i = 0; i < hw.size(); ++i) hw[i] = (float)((int)((i * 2654435761u) >> 20 & 1023) - 512) / (512.0f * sqrtf((float)C));
      for (size_t i = 0; i < hw2.size(); ++i) hw2[i] = (float)((int)((i * 2246822519u) >> 20 & 1023) - 512) / (512.0f * sqrtf((float)hid));
      for (int i = 0; i < hid; ++i) hb[i] = 0.01f * (i % 7);
      for (int i = 0; i < C; ++i) hb2[i] = 0.01f * (i % 5);
      L = make_layer_standalone(hid, C, 1, 1, hw.data(), hb.data(), s.ptrs, true);
      L2 = make_layer_standalone(C, hid, 1, 1, hw2.data(), hb2.data(), s.ptrs);
      ml.xt = make_view(s.alloc(px * C * 4), F32, 1, 1, (int)px, C);
      fill(s, ml.xt.p, F32, (long long)px * C, 2u, 1.0f);
      ml.x16 = make_view(s.alloc(px * C * 2), AD, 1, 1, (int)px, C);
      const int parts = tc_gemm_ln_parts(C);
      float2* stats = (float2*)s.alloc((size_t)parts * px * sizeof(float2));
      float2* mr = (float2*)s.alloc(px * sizeof(float2));
      glue_ln_stats_cast(ctx, ml.xt, ml.x16, stats);
      glue_ln_finalize(ctx, stats, 1, (long long)px, (long long)px, C, mr);
      ml.mr = mr; ml.fc1 = &L; ml.fc2 = &L2;
      ml.lne.stats = stats; ml.lne.stride = (long long)px; ml.lne.x16 = ml.x16.p; ml.lne.x16dt = AD; ml.lne.ldx16 = C;
      ml.mr_out = mr;
      hd = make_view(s.alloc(px * hid * 2), AD, 1, 1, (int)px, hid);
      // (the statistics are not refreshed between iterations: the values drift, the timing does not depend on them)
      if (with_res) launch = [&] { tc_mlp(ctx, ml); };
      else launch = [&] {
        GemmArgs g1; g1.x = ml.x16; g1.w = &L; g1.act = ACT_GELU; g1.out = hd; g1.lnf.mr = ml.mr; g1.lnf.C = C;
        tc_gemm(ctx, g1);
        GemmArgs g2; g2.x = hd; g2.w = &L2; g2.out = ml.xt; g2.res = ml.xt; g2.lne = ml.lne;
        tc_gemm(ctx, g2);
      };
    } else {
      const int heads = C, Cc = heads * 32, nwin = B;
      const size_t rows = (size_t)nwin * 144;
      View q = make_view(s.alloc(rows * 3 * Cc * dsize(AD)), AD, 1, 1, (int)rows, 3 * Cc);
      fill(s, q.p, AD, (long long)rows * 3 * Cc, 5u, 1.0f);
      float* b32p = (float*)s.alloc((size_t)heads * 144 * 148 * 4);
      fill(s, b32p, F32, (long long)heads * 144 * 148, 6u, 0.5f);
      float* b32 = (float*)s.alloc((size_t)heads * 144 * 144 * 4);
      fill(s, b32, F32, (long long)heads * 144 * 144, 7u, 0.5f);
      View o = make_view(s.alloc(rows * Cc * dsize(AD)), AD, 1, 1, (int)rows, Cc);
      at.qkv = q; at.bias32 = b32; at.bias32p = b32p; at.n_windows = nwin; at.heads = heads;
      at.nwh = H; at.nww = W; at.shift = k; at.out = o;
      if (with_res) {       // the model's configuration: token-order output + pad-row fix-up, N pad tokens per grid side
        at.h = H * 12 - N; at.w = W * 12 - N; at.token_out = 1;
        void* b16 = s.alloc((size_t)3 * Cc * 2);
        fill(s, b16, AD, (long long)3 * Cc, 8u, 0.5f);
        at.qkv_bias16 = b16;
        at.out = make_view(o.p, AD, 1, 1, (nwin / (H * W)) * at.h * at.w, Cc);
      }
      launch = [&] { op_attention(ctx, at); };
    }
    for (int i = 0; i < 3; ++i) launch();
    BRN_CUDA(cudaStreamSynchronize(s.stream));
    BRN_CUDA(cudaEventRecord(e0, s.stream));
    for (int i = 0; i < iters; ++i) launch();
    BRN_CUDA(cudaEventRecord(e1, s.stream));
    BRN_CUDA(cudaStreamSynchronize(s.stream));
    float ms = 0; BRN_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *ms_out = ms / iters;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  });
}

}  // extern "C"
