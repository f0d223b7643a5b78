// Model handle: schema (HF safetensors keys of the reference), finalize-time folding and
// upload, and the forward graph of BiRefNet::forward_logits (src/birefnet.rs:412-461) expressed over NHWC views.
#include "model.h"

#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>

#include "device_utils.cuh"

namespace brn {

static bool env_flag(const char* n);

static const int kIptIn[5] = {3, 48, 192, 768, 3072};     // image2patches channel counts 3*g*g
static const int kIptOut[5] = {48, 96, 192, 384, 384};    // src/birefnet.rs:180

// ------------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------------
void op_gemm(const LaunchCtx& ctx, const GemmArgs& a) {
  if (ctx.precision != BRN_PREC_FP32 && !ctx.force_simt && tc_gemm_supported(a)) tc_gemm(ctx, a);
  else simt_gemm(ctx, a);
}
void op_deform(const LaunchCtx& ctx, const DeformArgs& a) {
  const bool tc = ctx.precision != BRN_PREC_FP32 && !ctx.force_simt && tc_deform_supported(a);
  if (tc && a.scratch && a.w->taps() == 1) {
    // 1x1: sampling kernel + ordinary 1x1 implicit GEMM (same rounding points as the fused kernel)
    View smp = make_view(a.scratch, a.x.dt, a.x.B, a.x.H, a.x.W, 64);
    glue_deform_sample_k1(ctx, a.x, a.om, a.om_tiled, a.om_layer, smp);
    GemmArgs g; g.x = smp; g.w = a.w; g.bias = a.bias; g.act = a.act; g.out = a.out;
    BRN_CHECK(tc_gemm_supported(g), 5, "deform k=1: tcgen05 GEMM unavailable");
    tc_gemm(ctx, g);
    return;
  }
  if (tc) tc_deform(ctx, a);
  else simt_deform(ctx, a);
}
void op_attention(const LaunchCtx& ctx, const AttnArgs& a) {
  if (ctx.precision != BRN_PREC_FP32 && !ctx.force_simt && a.ws == 12 && (a.qkv.dt == BF16 || a.qkv.dt == F16) &&
      a.out.dt == a.qkv.dt)
    tc_attention(ctx, a);
  else simt_attention(ctx, a);
}

// ------------------------------------------------------------------------------------------------
// construction / schema
// ------------------------------------------------------------------------------------------------
Model::Model(const brn_config& c, int dev) : cfg(c), device(dev) {
  BRN_CHECK(c.window_size == 12 || c.window_size == 7, 7, "window_size must be 12 (swin_b / swin_l) or 7 (swin_t / swin_s)");
  BRN_CHECK(c.embed_dim > 0 && c.embed_dim % 32 == 0, 7, "embed_dim must be a multiple of 32");
  for (int i = 0; i < 4; ++i)
    BRN_CHECK(c.num_heads[i] * 32 == (c.embed_dim << i), 7, "head_dim must be 32 at every stage");
  BRN_CHECK(c.patch_size == 4 && c.mlp_ratio >= 1, 7, "patch_size must be 4");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  BRN_CHECK(e == cudaSuccess && ndev > 0, 2,
            std::string("no CUDA device available (this library has no CPU fallback): ") + cudaGetErrorString(e));
  BRN_CHECK(dev >= 0 && dev < ndev, 1, "device index out of range");
  DeviceGuard dg(dev);
  cudaDeviceProp prop;
  BRN_CUDA(cudaGetDeviceProperties(&prop, dev));
  BRN_CHECK(prop.major == 10, 2, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                     ", this library is built for sm_100a only");
  BRN_CUDA(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
  for (auto& l : lanes) {
    BRN_CUDA(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
    BRN_CUDA(cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming));
  }
  build_schema();
}

Model::~Model() {
  int prev_dev = -1;
  cudaGetDevice(&prev_dev);
  cudaSetDevice(device);
  drop_graphs();
  for (void* p : allocs) cudaFree(p);
  if (arena.base) cudaFree(arena.base);
  for (auto& l : lanes) {
    if (l.arena.base) cudaFree(l.arena.base);
    if (l.done) cudaEventDestroy(l.done);
    if (l.stream) cudaStreamDestroy(l.stream);
  }
  for (auto& pe : prof) { if (pe.e0) cudaEventDestroy(pe.e0); if (pe.e1) cudaEventDestroy(pe.e1); }
  for (auto e : ktimer.pool) cudaEventDestroy(e);
  if (own_stream) cudaStreamDestroy(own_stream);
  if (prev_dev >= 0 && prev_dev != device) cudaSetDevice(prev_dev);
}

void Model::build_schema() {
  auto add = [&](const std::string& k, std::vector<int64_t> shape) {
    index[k] = (int)keys.size();
    keys.push_back(k);
    HostTensor t; t.shape = std::move(shape);
    tensors.push_back(std::move(t));
  };
  auto lin = [&](const std::string& p, int o, int i, bool bias = true) {
    add(p + ".weight", {o, i});
    if (bias) add(p + ".bias", {o});
  };
  auto ln = [&](const std::string& p, int c) { add(p + ".weight", {c}); add(p + ".bias", {c}); };
  auto conv = [&](const std::string& p, int o, int i, int k, bool bias = true) {
    add(p + ".weight", {o, i, k, k});
    if (bias) add(p + ".bias", {o});
  };
  auto bn = [&](const std::string& p, int c) {
    add(p + ".running_mean", {c}); add(p + ".running_var", {c}); add(p + ".weight", {c}); add(p + ".bias", {c});
  };
  const int E = cfg.embed_dim;
  conv("bb.patch_embed.proj", E, 3, cfg.patch_size);
  ln("bb.patch_embed.norm", E);
  for (int i = 0; i < 4; ++i) {
    const int Ci = E << i;
    for (int j = 0; j < cfg.depths[i]; ++j) {
      std::string p = "bb.layers." + std::to_string(i) + ".blocks." + std::to_string(j);
      ln(p + ".norm1", Ci);
      lin(p + ".attn.qkv", 3 * Ci, Ci);
      lin(p + ".attn.proj", Ci, Ci);
      add(p + ".attn.relative_position_bias_table", {(2 * cfg.window_size - 1) * (2 * cfg.window_size - 1), cfg.num_heads[i]});
      ln(p + ".norm2", Ci);
      lin(p + ".mlp.fc1", cfg.mlp_ratio * Ci, Ci);
      lin(p + ".mlp.fc2", Ci, cfg.mlp_ratio * Ci);
    }
    if (i < 3) {
      ln("bb.layers." + std::to_string(i) + ".downsample.norm", 4 * Ci);
      lin("bb.layers." + std::to_string(i) + ".downsample.reduction", 2 * Ci, 4 * Ci, false);
    }
    ln("bb.norm" + std::to_string(i), Ci);
  }
  auto dec_blk = [&](const std::string& p, int cin, int cout) {
    conv(p + ".conv_in", 64, cin, 3);
    bn(p + ".bn_in", 64);
    const std::string a = p + ".dec_att";
    const int ks[4] = {1, 1, 3, 7};
    for (int b = 0; b < 4; ++b) {
      std::string bp = b == 0 ? a + ".aspp1" : a + ".aspp_deforms." + std::to_string(b - 1);
      int k = ks[b];
      conv(bp + ".atrous_conv.offset_conv", 2 * k * k, 64, k);
      conv(bp + ".atrous_conv.modulator_conv", k * k, 64, k);
      conv(bp + ".atrous_conv.regular_conv", 256, 64, k, false);
      bn(bp + ".bn", 256);
    }
    conv(a + ".global_avg_pool.1", 256, 64, 1, false);
    bn(a + ".global_avg_pool.2", 256);
    conv(a + ".conv1", 64, 1280, 1, false);
    bn(a + ".bn1", 64);
    conv(p + ".conv_out", cout, 64, 3);
    bn(p + ".bn_out", cout);
  };
  const int dec_out[4] = {lat(2), lat(1), lat(0), lat(0) / 2};
  const int dec_in[4] = {lat(3) + kIptOut[4], dec_out[0] + kIptOut[3], dec_out[1] + kIptOut[2], dec_out[2] + kIptOut[1]};
  dec_blk("squeeze_module.0", x4_channels(), lat(3));
  for (int n = 0; n < 5; ++n) {
    conv("decoder.ipt_blk" + std::to_string(n + 1) + ".conv1", 64, kIptIn[n], 3);
    conv("decoder.ipt_blk" + std::to_string(n + 1) + ".conv_out", kIptOut[n], 64, 3);
  }
  for (int d = 0; d < 4; ++d) dec_blk("decoder.decoder_block" + std::to_string(4 - d), dec_in[d], dec_out[d]);
  for (int d = 0; d < 3; ++d) conv("decoder.lateral_block" + std::to_string(4 - d) + ".conv", lat(2 - d), lat(2 - d), 1);
  for (int d = 0; d < 3; ++d) {
    std::string n = std::to_string(4 - d);
    conv("decoder.gdt_convs_" + n + ".0", 16, dec_out[d], 3);
    bn("decoder.gdt_convs_" + n + ".1", 16);
    conv("decoder.gdt_convs_attn_" + n + ".0", 1, 16, 1);
    conv("decoder.gdt_convs_pred_" + n + ".0", 1, 16, 1);  // loaded, unused in forward (src/birefnet.rs:230-232)
    conv("decoder.conv_ms_spvn_" + n, 1, dec_out[d], 1);   // loaded, unused in forward (src/birefnet.rs:241-243)
  }
  conv("decoder.conv_out1.0", 1, dec_out[3] + kIptOut[0], 1);
}

void Model::set_tensor(const char* key, const void* data, int dtype, const int64_t* shape, int rank) {
  BRN_CHECK(key && data && shape, 1, "set_tensor: null argument");
  BRN_CHECK(!finalized, 6, "set_tensor after finalize");
  auto it = index.find(key);
  BRN_CHECK(it != index.end(), 4, std::string("unknown tensor key: ") + key);
  HostTensor& t = tensors[it->second];
  bool ok = rank == (int)t.shape.size();
  for (int i = 0; ok && i < rank; ++i) ok = shape[i] == t.shape[i];
  BRN_CHECK(ok, 5, std::string("shape mismatch for ") + key);
  size_t n = t.numel();
  t.data.resize(n);
  if (dtype == BRN_F32) {
    memcpy(t.data.data(), data, n * sizeof(float));
  } else if (dtype == BRN_BF16) {
    const uint16_t* s = (const uint16_t*)data;
    for (size_t i = 0; i < n; ++i) { uint32_t u = (uint32_t)s[i] << 16; memcpy(&t.data[i], &u, 4); }
  } else if (dtype == BRN_F16) {
    const __half* s = (const __half*)data;
    for (size_t i = 0; i < n; ++i) t.data[i] = __half2float(s[i]);
  } else {
    throw Error(5, "set_tensor: unsupported dtype");
  }
  t.set = true;
}

const HostTensor& Model::T(const std::string& k) const {
  auto it = index.find(k);
  BRN_CHECK(it != index.end(), 4, "internal: key not in schema: " + k);
  return tensors[it->second];
}

float* Model::upload(const std::vector<float>& v) {
  float* d = nullptr;
  BRN_CUDA(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(float)));
  allocs.push_back(d);
  BRN_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
  return d;
}

static uint16_t f2bf(float f) {  // round-to-nearest-even
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

static uint16_t f2h(float f) {
  __half h = __float2half_rn(f);
  uint16_t u; memcpy(&u, &h, 2);
  return u;
}

LayerW make_layer_standalone(int N, int Cin, int kh, int kw, const float* w, const float* bias,
                             std::vector<void*>& allocs, bool folded_ln) {
  LayerW L;
  L.N = N; L.Cin = Cin; L.kh = kh; L.kw = kw;
  L.cin_pad = (Cin + 63) / 64 * 64;
  const int taps = kh * kw;
  std::vector<float> w32((size_t)N * taps * Cin);
  std::vector<uint16_t> wbf((size_t)N * taps * L.cin_pad, 0), wfp((size_t)N * taps * L.cin_pad, 0);
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < Cin; ++c)
      for (int t = 0; t < taps; ++t) {
        float v = w[((size_t)n * Cin + c) * taps + t];
        w32[((size_t)n * taps + t) * Cin + c] = v;
        wbf[((size_t)n * taps + t) * L.cin_pad + c] = f2bf(v);
        wfp[((size_t)n * taps + t) * L.cin_pad + c] = f2h(v);
      }
  if (!folded_ln) {      // a gamma-folded copy is only ever read by the tensor-core path
    BRN_CUDA(cudaMalloc(&L.w32, w32.size() * 4)); allocs.push_back(L.w32);
    BRN_CUDA(cudaMemcpy(L.w32, w32.data(), w32.size() * 4, cudaMemcpyHostToDevice));
  } else {
    // column sums of the weights AS ROUNDED: the LnFold epilogue computes acc - mean * colsum, and acc was
    // accumulated from the rounded values
    std::vector<float> cb(N), cf(N);
    for (int n = 0; n < N; ++n) {
      double sb = 0, sf = 0;
      for (size_t i = 0; i < (size_t)taps * L.cin_pad; ++i) {
        uint32_t ub = (uint32_t)wbf[(size_t)n * taps * L.cin_pad + i] << 16; float fb; memcpy(&fb, &ub, 4);
        __half hh; memcpy(&hh, &wfp[(size_t)n * taps * L.cin_pad + i], 2);
        sb += fb; sf += __half2float(hh);
      }
      cb[n] = (float)sb; cf[n] = (float)sf;
    }
    BRN_CUDA(cudaMalloc(&L.colsum_bf16, (size_t)N * 4)); allocs.push_back(L.colsum_bf16);
    BRN_CUDA(cudaMemcpy(L.colsum_bf16, cb.data(), (size_t)N * 4, cudaMemcpyHostToDevice));
    BRN_CUDA(cudaMalloc(&L.colsum_fp16, (size_t)N * 4)); allocs.push_back(L.colsum_fp16);
    BRN_CUDA(cudaMemcpy(L.colsum_fp16, cf.data(), (size_t)N * 4, cudaMemcpyHostToDevice));
    L.h_fold = std::make_shared<std::vector<float>>((size_t)3 * N, 0.f);
    for (int n = 0; n < N; ++n) {
      (*L.h_fold)[n] = cb[n]; (*L.h_fold)[(size_t)N + n] = cf[n];
      if (bias) (*L.h_fold)[(size_t)2 * N + n] = bias[n];
    }
  }
  BRN_CUDA(cudaMalloc(&L.w_bf16, wbf.size() * 2)); allocs.push_back(L.w_bf16);
  BRN_CUDA(cudaMemcpy(L.w_bf16, wbf.data(), wbf.size() * 2, cudaMemcpyHostToDevice));
  BRN_CUDA(cudaMalloc(&L.w_fp16, wfp.size() * 2)); allocs.push_back(L.w_fp16);
  BRN_CUDA(cudaMemcpy(L.w_fp16, wfp.data(), wfp.size() * 2, cudaMemcpyHostToDevice));
  if (bias) {
    BRN_CUDA(cudaMalloc(&L.bias, (size_t)N * 4)); allocs.push_back(L.bias);
    BRN_CUDA(cudaMemcpy(L.bias, bias, (size_t)N * 4, cudaMemcpyHostToDevice));
    if (N <= 1024) L.h_bias = std::make_shared<std::vector<float>>(bias, bias + N);
  }
  return L;
}

LayerW Model::make_layer(int N, int Cin, int kh, int kw, const std::vector<float>& w, const std::vector<float>* bias,
                         bool folded_ln) {
  BRN_CHECK(w.size() == (size_t)N * Cin * kh * kw, 5, "internal: make_layer size");
  return make_layer_standalone(N, Cin, kh, kw, w.data(), bias ? bias->data() : nullptr, allocs, folded_ln);
}

// LayerNorm folded into the linear layer that consumes it:
//   LN(x) W^T + b = rstd * (x (gamma .* W)^T - mean * colsum(gamma .* W)) + (W beta + b)
LayerW Model::make_folded(const std::vector<float>& w, const std::vector<float>& b, const std::vector<float>& gamma,
                          const std::vector<float>& beta, int N, int C) {
  std::vector<float> wf((size_t)N * C), bf(N);
  for (int n = 0; n < N; ++n) {
    double acc = b[n];
    for (int c = 0; c < C; ++c) {
      wf[(size_t)n * C + c] = (float)((double)w[(size_t)n * C + c] * gamma[c]);
      acc += (double)w[(size_t)n * C + c] * beta[c];
    }
    bf[n] = (float)acc;
  }
  return make_layer(N, C, 1, 1, wf, &bf, true);
}

// eval BatchNorm as (scale, shift): y = x*scale + shift  (candle batch_norm(C,1e-5).forward_t(x,false))
static void bn_affine(const HostTensor& mean, const HostTensor& var, const HostTensor& g, const HostTensor& b,
                      std::vector<double>& scale, std::vector<double>& shift) {
  size_t n = mean.data.size();
  scale.resize(n); shift.resize(n);
  for (size_t i = 0; i < n; ++i) {
    scale[i] = (double)g.data[i] / std::sqrt((double)var.data[i] + 1e-5);
    shift[i] = (double)b.data[i] - (double)mean.data[i] * scale[i];
  }
}

void Model::finalize() {
  BRN_CHECK(!finalized, 6, "finalize called twice");
  for (size_t i = 0; i < keys.size(); ++i)
    BRN_CHECK(tensors[i].set, 3, "missing tensor: " + keys[i]);
  DeviceGuard dg(device);

  // conv (+ optional BN fold) -> LayerW
  auto conv_bn = [&](const std::string& cp, bool has_bias, const std::string& bnp) -> LayerW {
    const HostTensor& w = T(cp + ".weight");
    int N = (int)w.shape[0], Cin = (int)w.shape[1], k = (int)w.shape[2];
    std::vector<float> wf = w.data;
    std::vector<float> bf(N, 0.f);
    if (has_bias) bf = T(cp + ".bias").data;
    bool any_bias = has_bias;
    if (!bnp.empty()) {
      std::vector<double> sc, sh;
      bn_affine(T(bnp + ".running_mean"), T(bnp + ".running_var"), T(bnp + ".weight"), T(bnp + ".bias"), sc, sh);
      size_t per = (size_t)Cin * k * k;
      for (int n = 0; n < N; ++n) {
        for (size_t i = 0; i < per; ++i) wf[n * per + i] = (float)((double)wf[n * per + i] * sc[n]);
        bf[n] = (float)((double)bf[n] * sc[n] + sh[n]);
      }
      any_bias = true;
    }
    return make_layer(N, Cin, k, k, wf, any_bias ? &bf : nullptr);
  };
  auto linear = [&](const std::string& p, bool bias) -> LayerW {
    const HostTensor& w = T(p + ".weight");
    return make_layer((int)w.shape[0], (int)w.shape[1], 1, 1, w.data, bias ? &T(p + ".bias").data : nullptr);
  };

  // ---- backbone ----
  {
    const HostTensor& w = T("bb.patch_embed.proj.weight");   // [E,3,4,4] -> linear [E, 48], k = c*16+ky*4+kx
    patch_embed = make_layer((int)w.shape[0], 3 * cfg.patch_size * cfg.patch_size, 1, 1, w.data,
                             &T("bb.patch_embed.proj.bias").data);
    pe_g = upload(T("bb.patch_embed.norm.weight").data);
    pe_b = upload(T("bb.patch_embed.norm.bias").data);
  }
  const double scale = 1.0 / std::sqrt(32.0);   // head_dim^-0.5 (src/swin.rs:134), folded into the q rows of qkv
  for (int i = 0; i < 4; ++i) {
    const int Ci = C(i), heads = cfg.num_heads[i];
    StageW& S = stages[i];
    for (int j = 0; j < cfg.depths[i]; ++j) {
      std::string p = "bb.layers." + std::to_string(i) + ".blocks." + std::to_string(j);
      BlockW B{};
      B.n1g = upload(T(p + ".norm1.weight").data); B.n1b = upload(T(p + ".norm1.bias").data);
      B.n2g = upload(T(p + ".norm2.weight").data); B.n2b = upload(T(p + ".norm2.bias").data);
      {
        std::vector<float> w = T(p + ".attn.qkv.weight").data, b = T(p + ".attn.qkv.bias").data;
        for (size_t r = 0; r < (size_t)Ci; ++r) {
          for (size_t c = 0; c < (size_t)Ci; ++c) w[r * Ci + c] = (float)((double)w[r * Ci + c] * scale);
          b[r] = (float)((double)b[r] * scale);
        }
        B.qkv = make_layer(3 * Ci, Ci, 1, 1, w, &b);
        B.qkv_f = make_folded(w, b, T(p + ".norm1.weight").data, T(p + ".norm1.bias").data, 3 * Ci, Ci);
        // qkv of a pad token = the (q-scaled) bias, in both operand types: [3C] bf16 then [3C] fp16
        std::vector<uint16_t> b16((size_t)6 * Ci);
        for (int n = 0; n < 3 * Ci; ++n) { b16[n] = f2bf(b[n]); b16[3 * Ci + n] = f2h(b[n]); }
        BRN_CUDA(cudaMalloc(&B.qkv_bias16, b16.size() * 2)); allocs.push_back(B.qkv_bias16);
        BRN_CUDA(cudaMemcpy(B.qkv_bias16, b16.data(), b16.size() * 2, cudaMemcpyHostToDevice));
      }
      B.proj = linear(p + ".attn.proj", true);
      B.fc1 = linear(p + ".mlp.fc1", true);
      B.fc1_f = make_folded(T(p + ".mlp.fc1.weight").data, T(p + ".mlp.fc1.bias").data, T(p + ".norm2.weight").data,
                            T(p + ".norm2.bias").data, cfg.mlp_ratio * Ci, Ci);
      B.fc2 = linear(p + ".mlp.fc2", true);
      {
        // WindowAttention::new (src/swin.rs:143-152): bias[h,q,k] = table[index[q,k], h],
        // index[(i,j),(k,l)] = (i-k+11)*23 + (j-l+11)  (src/swin.rs:182-184)
        const HostTensor& tb = T(p + ".attn.relative_position_bias_table");
        // generic window side ws: index = (qi - ki + ws - 1) * (2 ws - 1) + (qj - kj + ws - 1)
        const int ws = cfg.window_size, n = ws * ws, ldp = n + 4;
        std::vector<float> b32((size_t)heads * n * n), b32p((size_t)heads * n * ldp, 0.f);
        for (int h = 0; h < heads; ++h)
          for (int q = 0; q < n; ++q)
            for (int k = 0; k < n; ++k) {
              int idx = (q / ws - k / ws + ws - 1) * (2 * ws - 1) + (q % ws - k % ws + ws - 1);
              float v = tb.data[(size_t)idx * heads + h];
              b32[((size_t)h * n + q) * n + k] = v;
              b32p[((size_t)h * n + q) * ldp + k] = v;
            }
        B.bias32 = upload(b32);
        B.bias32p = upload(b32p);
      }
      S.blocks.push_back(B);
    }
    S.has_down = i < 3;
    if (S.has_down) {
      std::string p = "bb.layers." + std::to_string(i) + ".downsample";
      S.dng = upload(T(p + ".norm.weight").data); S.dnb = upload(T(p + ".norm.bias").data);
      S.red = linear(p + ".reduction", false);
    }
    S.ng = upload(T("bb.norm" + std::to_string(i) + ".weight").data);
    S.nb = upload(T("bb.norm" + std::to_string(i) + ".bias").data);
  }

  // ---- decoder blocks ----
  auto dec_blk = [&](const std::string& p) -> DecBlkW {
    DecBlkW D;
    D.conv_in = conv_bn(p + ".conv_in", true, p + ".bn_in");
    const std::string a = p + ".dec_att";
    const int ks[4] = {1, 1, 3, 7};
    for (int b = 0; b < 4; ++b) {
      std::string bp = b == 0 ? a + ".aspp1" : a + ".aspp_deforms." + std::to_string(b - 1);
      int k = ks[b];
      D.br[b].k = k;
      // offset_conv ++ modulator_conv share input and geometry -> one conv with 3k^2 outputs
      const HostTensor &ow = T(bp + ".atrous_conv.offset_conv.weight"), &ob = T(bp + ".atrous_conv.offset_conv.bias");
      const HostTensor &mw = T(bp + ".atrous_conv.modulator_conv.weight"), &mb = T(bp + ".atrous_conv.modulator_conv.bias");
      std::vector<float> w = ow.data; w.insert(w.end(), mw.data.begin(), mw.data.end());
      std::vector<float> bb = ob.data; bb.insert(bb.end(), mb.data.begin(), mb.data.end());
      D.br[b].om = make_layer(3 * k * k, 64, k, k, w, &bb);
      D.br[b].reg = conv_bn(bp + ".atrous_conv.regular_conv", false, bp + ".bn");
    }
    D.gap = conv_bn(a + ".global_avg_pool.1", false, a + ".global_avg_pool.2");
    {
      const HostTensor& w = T(a + ".conv1.weight");  // [64,1280,1,1]
      std::vector<double> sc, sh;
      bn_affine(T(a + ".bn1.running_mean"), T(a + ".bn1.running_var"), T(a + ".bn1.weight"), T(a + ".bn1.bias"), sc, sh);
      std::vector<float> head((size_t)64 * 1024), tail((size_t)64 * 256), shift(64);
      for (int o = 0; o < 64; ++o) {
        for (int c = 0; c < 1024; ++c) head[(size_t)o * 1024 + c] = (float)((double)w.data[(size_t)o * 1280 + c] * sc[o]);
        for (int c = 0; c < 256; ++c) tail[(size_t)o * 256 + c] = (float)((double)w.data[(size_t)o * 1280 + 1024 + c] * sc[o]);
        shift[o] = (float)sh[o];
      }
      D.conv1 = make_layer(64, 1024, 1, 1, head, nullptr);
      D.conv1_tail = upload(tail);
      D.bn1_shift = upload(shift);
    }
    D.conv_out = conv_bn(p + ".conv_out", true, p + ".bn_out");
    return D;
  };
  dw.squeeze = dec_blk("squeeze_module.0");
  for (int d = 0; d < 4; ++d) dw.dec[d] = dec_blk("decoder.decoder_block" + std::to_string(4 - d));
  for (int n = 1; n < 5; ++n) {
    dw.ipt_conv1[n] = conv_bn("decoder.ipt_blk" + std::to_string(n + 1) + ".conv1", true, "");
    dw.ipt_out[n] = conv_bn("decoder.ipt_blk" + std::to_string(n + 1) + ".conv_out", true, "");
  }
  for (int d = 0; d < 3; ++d) {
    std::string n = std::to_string(4 - d);
    dw.lat[d] = conv_bn("decoder.lateral_block" + n + ".conv", true, "");
    dw.gdt[d] = conv_bn("decoder.gdt_convs_" + n + ".0", true, "decoder.gdt_convs_" + n + ".1");
    dw.gdt_attn_w[d] = upload(T("decoder.gdt_convs_attn_" + n + ".0.weight").data);
    dw.gdt_attn_b[d] = T("decoder.gdt_convs_attn_" + n + ".0.bias").data[0];
  }
  {
    // final layer rewrite; done in double
    const HostTensor& wo = T("decoder.conv_out1.0.weight");   // [1, P + 48, 1, 1]
    const int P = lat(0) / 2;
    {
      // conv_out1 is linear in p1 and decoder_block1 ends in conv_out + BN with no activation (src/decoder.rs:137-140):
      // w_p . (BN(conv3x3(a))) is ONE 3x3 conv 64 -> 1 with weights sum_n w_p[n] * sc[n] * W[n], so the 192-channel
      // p1 is never materialised.
      const std::string bp = "decoder.decoder_block1";
      const HostTensor& cw = T(bp + ".conv_out.weight");          // [P, 64, 3, 3]
      const HostTensor& cb = T(bp + ".conv_out.bias");
      std::vector<double> sc, sh;
      bn_affine(T(bp + ".bn_out.running_mean"), T(bp + ".bn_out.running_var"), T(bp + ".bn_out.weight"), T(bp + ".bn_out.bias"), sc, sh);
      std::vector<float> wq((size_t)64 * 9), bq(1);
      for (size_t i = 0; i < wq.size(); ++i) {
        double acc = 0;
        for (int n = 0; n < P; ++n) acc += (double)wo.data[n] * sc[n] * (double)cw.data[(size_t)n * 64 * 9 + i];
        wq[i] = (float)acc;
      }
      double accb = 0;
      for (int n = 0; n < P; ++n) accb += (double)wo.data[n] * ((double)cb.data[n] * sc[n] + sh[n]);
      bq[0] = (float)accb;
      dw.out_q = make_layer(1, 64, 3, 3, wq, &bq);
    }
    const HostTensor& c1 = T("decoder.ipt_blk1.conv1.weight");     // [64,3,3,3]
    const HostTensor& co = T("decoder.ipt_blk1.conv_out.weight");  // [48,64,3,3]
    const HostTensor& cob = T("decoder.ipt_blk1.conv_out.bias");
    std::vector<double> wc((size_t)64 * 9);
    for (int c = 0; c < 64; ++c)
      for (int t = 0; t < 9; ++t) {
        double s = 0;
        for (int o = 0; o < 48; ++o) s += (double)wo.data[P + o] * (double)co.data[((size_t)o * 64 + c) * 9 + t];
        wc[(size_t)c * 9 + t] = s;
      }
    double bc = T("decoder.conv_out1.0.bias").data[0];
    for (int o = 0; o < 48; ++o) bc += (double)wo.data[P + o] * (double)cob.data[o];
    std::vector<float> tab(336);
    build_final_table(c1.data.data(), T("decoder.ipt_blk1.conv1.bias").data.data(), wc.data(), bc, tab.data());
    dw.fin_tab = upload(tab);
    dw.fin_tab_host = tab;
  }
  // host copies are no longer needed
  for (auto& t : tensors) { std::vector<float>().swap(t.data); }
  finalized = true;
}

// ------------------------------------------------------------------------------------------------
// forward graph
// ------------------------------------------------------------------------------------------------
void Model::drop_graphs(const void* arena_base) {
  for (size_t i = graphs.size(); i-- > 0;) {
    if (arena_base && graphs[i].key.arena_base != arena_base) continue;
    if (graphs[i].exec) cudaGraphExecDestroy(graphs[i].exec);
    graphs.erase(graphs.begin() + i);
  }
}

int Model::acquire_lane(std::unique_lock<std::mutex>& lk, cudaStream_t s) {
  // Preference: the lane this caller stream used last (stream order makes reuse free and its CUDA graphs match), then
  // a lane whose previous call has finished on the device, then any lane nobody is launching into (the new call is
  // ordered behind the old one with the lane's event).
  for (;;) {
    int pick = -1;
    if (s)
      for (int i = 0; i < kLanes && pick < 0; ++i) if (!lanes[i].busy && lanes[i].last == s) pick = i;
    for (int i = 0; i < kLanes && pick < 0; ++i)
      if (!lanes[i].busy) {
        if (cudaEventQuery(lanes[i].done) == cudaSuccess) pick = i;
        else cudaGetLastError();
      }
    for (int i = 0; i < kLanes && pick < 0; ++i) if (!lanes[i].busy) pick = i;
    if (pick >= 0) {
      lanes[pick].busy = true;
      std::swap(arena, lanes[pick].arena);
      return pick;
    }
    lane_cv.wait(lk);
  }
}

void Model::release_lane(int i) {
  lanes[i].busy = false;
  lane_cv.notify_all();   // two kinds of waiter share the condition variable (acquire_lane, quiesce)
}

void Model::quiesce(std::unique_lock<std::mutex>& lk) {
  lane_cv.wait(lk, [&] { for (auto& l : lanes) if (l.busy) return false; return true; });
}

void Model::ensure_arena(size_t bytes) {
  if (arena.cap >= bytes) return;
  if (arena.base) drop_graphs(arena.base);          // captured kernels hold arena addresses
  if (arena.base) { BRN_CUDA(cudaFree(arena.base)); arena.base = nullptr; arena.cap = 0; }
  size_t want = bytes + (bytes >> 4) + (1 << 20);
  BRN_CUDA(cudaMalloc(&arena.base, want));
  arena.cap = want;
}

int Model::micro_batch(int B, int H, int W) const {
  if (cfg.micro_batch > 0) return std::min(B, (int)cfg.micro_batch);
  // ~2.6 GB of workspace per 1024^2 image in bf16, twice that in fp32; keep well inside 180 GB
  double per = 3.0e9 * ((double)H * W / (1024.0 * 1024.0)) * (cfg.precision != BRN_PREC_FP32 ? 1.0 : 2.0);
  int mb = (int)std::max(1.0, std::floor(100.0e9 / per));
  return std::min(B, std::min(mb, 16));
}

void Model::prof_begin(LaunchCtx& ctx, const char* name) {
  if (!prof_on || ctx.dry) return;
  ProfEntry pe; pe.name = name;
  BRN_CUDA(cudaEventCreate(&pe.e0)); BRN_CUDA(cudaEventCreate(&pe.e1));
  BRN_CUDA(cudaEventRecord(pe.e0, ctx.stream));
  prof.push_back(pe);
}
void Model::prof_end(LaunchCtx& ctx) {
  if (!prof_on || ctx.dry) return;
  BRN_CUDA(cudaEventRecord(prof.back().e1, ctx.stream));
}

// SwinTransformer::forward (src/swin.rs:768-797).  feats[i]: destination NHWC views (dtype = activation dtype).
// Optional second input (img2 at H2 x W2, features to feats2): BiRefNet runs the SAME backbone on the image and on its
// half-resolution copy (src/birefnet.rs:416-426).  On the tensor-core path both token grids are concatenated along M,
// so every row-wise op (qkv / proj / fc1 / fc2 GEMMs, norm2, window attention) is ONE launch over 1.25x the rows
// instead of two launches of which the small one fills 2.6 waves of the GPU; only the ops that see the 2-D geometry
// (window gather, patch merging, stage norms into the concat buffers) run once per grid.  Row results do not depend on
// the M tiling, so the outputs are bit-identical to two separate passes.
void Model::run_backbone(LaunchCtx& ctx, const float* img, int B, int H, int W, View feats[4], const float* img2, int H2,
                         int W2, View* feats2) {
  const int AD = act_dtype();
  const size_t m0 = arena.mark();
  const int nseg = img2 ? 2 : 1;
  const float* imgs[2] = {img, img2};
  int h[2] = {H / 4, img2 ? H2 / 4 : 0}, w[2] = {W / 4, img2 ? W2 / 4 : 0};
  auto rows_of = [&](const int* hh, const int* ww, long long* r) { r[0] = (long long)B * hh[0] * ww[0]; r[1] = nseg > 1 ? (long long)B * hh[1] * ww[1] : 0; };
  long long T[2]; rows_of(h, w, T);
  // PatchEmbed (src/swin.rs:692-714): conv 4x4/4 as a [T,48]x[48,E] GEMM, then LN over C
  const long long Tt0 = T[0] + T[1];
  View a0 = make_view(arena.alloc((size_t)Tt0 * 48 * dsize(AD)), AD, 1, 1, (int)Tt0, 48);
  for (int s = 0; s < nseg; ++s) {
    View seg = make_view((char*)a0.p + (size_t)(s ? T[0] : 0) * 48 * dsize(AD), AD, 1, 1, (int)T[s], 48);
    glue_patch_im2col(ctx, imgs[s], B, s ? H2 : H, s ? W2 : W, 4, seg);
  }
  float* xbuf = (float*)arena.alloc((size_t)Tt0 * C(0) * 4);
  {
    GemmArgs g; g.x = a0; g.w = &patch_embed; g.out = make_view(xbuf, F32, 1, 1, (int)Tt0, C(0));
    op_gemm(ctx, g);
    LnArgs l; l.x = g.out; l.gamma = pe_g; l.beta = pe_b; l.out = g.out; l.mode = LN_PLAIN;
    glue_layernorm(ctx, l);
  }
  // LayerNorm folding needs the tcgen05 epilogues (16-column granules: C % 16 == 0 holds for head_dim 32)
  // ... and the tcgen05 attention kernel's pad fix-up / token-order store, which is built for 12x12 windows
  const int ws = cfg.window_size, wn = ws * ws;
  const bool fold = ctx.precision != BRN_PREC_FP32 && !ctx.force_simt && ws == 12 && !env_flag("BRN_LN_UNFUSED");
  bool have_stats = false;          // x16 / stats describe the current residual stream
  View next_x16{}; float2* next_stats = nullptr;
  for (int i = 0; i < 4; ++i) {
    const int Ci = C(i), heads = cfg.num_heads[i];
    int hp[2], wp[2]; long long Tp[2] = {0, 0};
    for (int s = 0; s < nseg; ++s) { hp[s] = (h[s] + ws - 1) / ws * ws; wp[s] = (w[s] + ws - 1) / ws * ws; Tp[s] = (long long)B * hp[s] * wp[s]; }
    rows_of(h, w, T);
    const long long Tt = T[0] + T[1], Tpt = Tp[0] + Tp[1];
    BRN_CHECK(Tpt < (1ll << 31), 5, "too many tokens for one pass: lower micro_batch");
    View xt = make_view(xbuf, F32, 1, 1, (int)Tt, Ci);          // token-matrix view of the residual stream (both grids)
    auto grid_view = [&](int s) { return make_view(xbuf + (size_t)(s ? T[0] : 0) * Ci, F32, B, h[s], w[s], Ci); };
    // Tensor-core path: LayerNorm is folded into the GEMM that consumes it (LnFold / LnEmit, brn_common.h).  The
    // epilogue that writes the fp32 residual stream (proj, fc2, the PatchMerging reduction) also writes its raw 16-bit
    // copy `x16` and per-row (sum, sumsq) partials `stats`; qkv and fc1 run on x16 with gamma folded into their weights
    // and apply mean / rstd in their own epilogues.  All four GEMMs of a block run in TOKEN order: the qkv epilogue
    // scatters rows into the window-ordered padded layout the attention kernel loads (pad rows are synthesised inside
    // that kernel), and attention stores straight back to token order -- norm1 -> pad -> roll -> partition and
    // window_reverse -> roll -> crop (src/swin.rs:355-401) never exist as passes.
    View x16{}; float2 *stats = nullptr, *mr = nullptr; int parts = 0;
    if (fold) {
      parts = tc_gemm_ln_parts(Ci);
      mr = (float2*)arena.alloc((size_t)Tt * sizeof(float2));
      if (have_stats) {     // the previous stage's reduction GEMM emitted into these buffers
        x16 = next_x16; stats = next_stats;
        glue_ln_finalize(ctx, stats, parts, Tt, Tt, Ci, mr);
      } else {
        x16 = make_view(arena.alloc((size_t)Tt * Ci * dsize(AD)), AD, 1, 1, (int)Tt, Ci);
        stats = (float2*)arena.alloc((size_t)parts * Tt * sizeof(float2));
      }
    }
    // producer: op_gemm with the emit fields set, then the partials -> (-mean, rstd) pass the consumers prefetch from
    auto emit_gemm = [&](GemmArgs& g) {
      if (fold) { g.lne.stats = stats; g.lne.stride = Tt; g.lne.x16 = x16.p; g.lne.x16dt = AD; g.lne.ldx16 = Ci; }
      op_gemm(ctx, g);
      if (fold) glue_ln_finalize(ctx, stats, parts, Tt, Tt, Ci, mr);
    };
    auto folded = [&](GemmArgs& g) { g.lnf.mr = mr; g.lnf.C = Ci; };
    for (size_t j = 0; j < stages[i].blocks.size(); ++j) {
      const BlockW& bw = stages[i].blocks[j];
      const int shift = (j % 2 == 0) ? 0 : ws / 2;     // src/swin.rs:548,552
      const size_t mb = arena.mark();
      View qkv = make_view(arena.alloc((size_t)Tpt * 3 * Ci * dsize(AD)), AD, 1, 1, (int)Tpt, 3 * Ci);
      View xw{};
      RowMap wmap;      // window-ordered padded rows <-> token rows of the (merged) grids
      wmap.h = h[0]; wmap.w = w[0]; wmap.hp = hp[0]; wmap.wp = wp[0]; wmap.shift = shift; wmap.ws = ws;
      if (nseg > 1) { wmap.split = Tp[0]; wmap.h2 = h[1]; wmap.w2 = w[1]; wmap.hp2 = hp[1]; wmap.wp2 = wp[1]; wmap.tok2 = T[0]; }
      if (fold && have_stats) {
        // norm1 folded: qkv = LN(x) Wqkv^T + b on the raw 16-bit stream, rows scattered to the window layout
        GemmArgs g; g.x = x16; g.w = &bw.qkv_f; g.out = qkv; folded(g);
        g.rowmap = wmap; g.rowmap.enabled = 2;
        BRN_CHECK(tc_gemm_supported(g), 5, "folded qkv: tcgen05 path unavailable");
        tc_gemm(ctx, g);
      } else {
        // norm1 -> pad -> roll -> partition (src/swin.rs:355-380) in one gather kernel per grid
        xw = make_view(arena.alloc((size_t)Tpt * Ci * dsize(AD)), AD, 1, 1, (int)Tpt, Ci);
        LnArgs l; l.x = grid_view(0); l.gamma = bw.n1g; l.beta = bw.n1b; l.out = xw;
        l.mode = LN_WINDOW; l.hp = hp[0]; l.wp = wp[0]; l.shift = shift; l.ws = ws;
        if (nseg > 1) { l.split = Tp[0]; l.tok2 = T[0]; l.h2 = h[1]; l.w2 = w[1]; l.hp2 = hp[1]; l.wp2 = wp[1]; }
        glue_layernorm(ctx, l);
        GemmArgs g; g.x = xw; g.w = &bw.qkv; g.out = qkv; op_gemm(ctx, g);
      }
      // attention output: token order on the tensor-core path, window order on the SIMT path
      View ao = fold ? make_view(arena.alloc((size_t)Tt * Ci * dsize(AD)), AD, 1, 1, (int)Tt, Ci)
                     : make_view(xw.p, AD, 1, 1, (int)Tpt, Ci);   // SIMT: reuse the xw buffer (qkv GEMM has consumed it)
      { AttnArgs a; a.qkv = qkv; a.bias32 = bw.bias32; a.bias32p = bw.bias32p; a.n_windows = (int)(Tpt / wn); a.ws = ws;
        a.heads = heads; a.nwh = hp[0] / ws; a.nww = wp[0] / ws; a.shift = shift; a.out = ao;
        if (nseg > 1) { a.split_win = (int)(Tp[0] / wn); a.nwh2 = hp[1] / ws; a.nww2 = wp[1] / ws; }
        if (fold) {
          a.h = h[0]; a.w = w[0]; a.h2 = h[1]; a.w2 = w[1]; a.tok2 = T[0]; a.token_out = 1;
          a.qkv_bias16 = bw.qkv_bias16 + (AD == F16 ? 3 * Ci : 0);
        }
        op_attention(ctx, a); }
      // proj (+ window_reverse + roll back + crop on the SIMT path) + residual (src/swin.rs:310,387-406)
      { GemmArgs g; g.x = ao; g.w = &bw.proj; g.out = xt; g.res = xt;
        if (!fold) { g.rowmap = wmap; g.rowmap.enabled = 1; }
        emit_gemm(g); }
      // x + fc2(gelu(fc1(norm2(x))))  (src/swin.rs:407)
      // Early stages (C <= 256, mlp_ratio 4): one fused kernel, the hidden activations stay on the SM (mlp_tcgen05.cu)
      MlpArgs ml; ml.x16 = x16; ml.mr = mr; ml.fc1 = &bw.fc1_f; ml.fc2 = &bw.fc2; ml.xt = xt;
      ml.lne.x16 = x16.p; ml.lne.x16dt = AD; ml.lne.ldx16 = Ci; ml.mr_out = mr;
      if (fold && cfg.mlp_ratio == 4 && !env_flag("BRN_MLP_UNFUSED") && tc_mlp_supported(ml)) {
        tc_mlp(ctx, ml);       // writes the next block's (-mean, rstd) itself
      } else {
        View hd = make_view(arena.alloc((size_t)Tt * cfg.mlp_ratio * Ci * dsize(AD)), AD, 1, 1, (int)Tt, cfg.mlp_ratio * Ci);
        if (fold) {
          GemmArgs g; g.x = x16; g.w = &bw.fc1_f; g.act = ACT_GELU; g.out = hd; folded(g);
          BRN_CHECK(tc_gemm_supported(g), 5, "folded fc1: tcgen05 path unavailable");
          tc_gemm(ctx, g);
        } else {
          View xn = make_view(arena.alloc((size_t)Tt * Ci * dsize(AD)), AD, 1, 1, (int)Tt, Ci);
          { LnArgs l; l.x = xt; l.gamma = bw.n2g; l.beta = bw.n2b; l.out = xn; l.mode = LN_PLAIN; glue_layernorm(ctx, l); }
          GemmArgs g; g.x = xn; g.w = &bw.fc1; g.act = ACT_GELU; g.out = hd; op_gemm(ctx, g);
        }
        { GemmArgs g; g.x = hd; g.w = &bw.fc2; g.out = xt; g.res = xt; emit_gemm(g); }
      }
      have_stats = fold;
      arena.release(mb);
    }
    // norm{i} -> NCHW view (src/swin.rs:784-788): written straight into the caller's NHWC slices
    for (int s = 0; s < nseg; ++s) {
      View* f = s ? feats2 : feats;
      LnArgs l; l.x = make_view(xbuf + (size_t)(s ? T[0] : 0) * Ci, F32, 1, 1, (int)T[s], Ci); l.gamma = stages[i].ng; l.beta = stages[i].nb;
      l.out = make_view(f[i].p, f[i].dt, 1, 1, (int)T[s], Ci, f[i].ld); l.mode = LN_PLAIN;
      glue_layernorm(ctx, l);
    }
    if (stages[i].has_down) {
      // PatchMerging (src/swin.rs:491-527): 2x2 gather + LN per grid, one reduction GEMM
      int h2[2], w2[2]; long long T2[2] = {0, 0};
      for (int s = 0; s < nseg; ++s) { h2[s] = (h[s] + 1) / 2; w2[s] = (w[s] + 1) / 2; T2[s] = (long long)B * h2[s] * w2[s]; }
      const long long T2t = T2[0] + T2[1];
      View xm = make_view(arena.alloc((size_t)T2t * 4 * Ci * dsize(AD)), AD, 1, 1, (int)T2t, 4 * Ci);
      for (int s = 0; s < nseg; ++s) {
        LnArgs l; l.x = grid_view(s); l.gamma = stages[i].dng; l.beta = stages[i].dnb;
        l.out = make_view((char*)xm.p + (size_t)(s ? T2[0] : 0) * 4 * Ci * dsize(AD), AD, 1, 1, (int)T2[s], 4 * Ci);
        l.mode = LN_MERGE;
        glue_layernorm(ctx, l);
      }
      float* xnew = (float*)arena.alloc((size_t)T2t * 2 * Ci * 4);
      { GemmArgs g; g.x = xm; g.w = &stages[i].red; g.out = make_view(xnew, F32, 1, 1, (int)T2t, 2 * Ci);
        if (fold) {     // the next stage's first norm1 is folded too: this epilogue produces its statistics
          next_x16 = make_view(arena.alloc((size_t)T2t * 2 * Ci * dsize(AD)), AD, 1, 1, (int)T2t, 2 * Ci);
          next_stats = (float2*)arena.alloc((size_t)tc_gemm_ln_parts(2 * Ci) * T2t * sizeof(float2));
          g.lne.stats = next_stats; g.lne.stride = T2t; g.lne.x16 = next_x16.p; g.lne.x16dt = AD; g.lne.ldx16 = 2 * Ci;
          have_stats = true;
        }
        op_gemm(ctx, g); }
      xbuf = xnew;
      for (int s = 0; s < nseg; ++s) { h[s] = h2[s]; w[s] = w2[s]; }
    }
  }
  arena.release(m0);
}

// BasicDecBlk::forward (src/decoder.rs:126-141) with ASPPDeformable (src/aspp.rs:303-333)
void Model::run_decblk(LaunchCtx& ctx, const DecBlkW& w, View in, View out, const LayerW* conv_out) {
  const int AD = dec_dtype();
  const size_t m0 = arena.mark();
  const int B = in.B, H = in.H, W = in.W;
  const size_t px = (size_t)B * H * W;
  View t = make_view(arena.alloc(px * 64 * dsize(AD)), AD, B, H, W, 64);
  { GemmArgs g; g.x = in; g.w = &w.conv_in; g.pad = 1; g.act = ACT_RELU; g.out = t; op_gemm(ctx, g); }
  View cat = make_view(arena.alloc(px * 1024 * dsize(AD)), AD, B, H, W, 1024);
  for (int b = 0; b < 4; ++b) {
    const int k = w.br[b].k;
    View dst = cat.slice(256 * b, 256);
    if (cfg.deform_mode == BRN_DEFORM_DEFORMABLE) {
      const size_t m1 = arena.mark();
      // tensor-core path: the offset/modulator conv writes tile-major [16x8-pixel tile][3k^2][128] so that the
      // gather producers of tc_deform read their per-tap (dy, dx, m) with coalesced loads
      GemmArgs g; g.x = t; g.w = &w.br[b].om; g.pad = k / 2; g.act = ACT_2SIGMOID_TAIL; g.act_from = 2 * k * k;
      DeformArgs d; d.x = t; d.w = &w.br[b].reg; d.act = ACT_RELU; d.out = dst;
      const bool tiled = ctx.precision != BRN_PREC_FP32 && !ctx.force_simt;
      size_t om_elems = px * 3 * k * k;
      if (tiled) om_elems = (size_t)B * ((H + 7) / 8) * ((W + 15) / 16) * 128 * 3 * k * k;
      View om = make_view(arena.alloc(om_elems * 4), F32, B, H, W, 3 * k * k);
      g.out = om;
      if (tiled) { g.tile_w = 16; g.out_tiled = 1; d.om_tiled = 1; BRN_CHECK(tc_gemm_supported(g), 5, "om conv: tcgen05 path unavailable"); }
      d.om = om;
      if (tiled && k == 1) {
        // 1x1 branch: the sampler computes its own three offset / modulator values (no om conv launch)
        d.scratch = arena.alloc(px * 64 * dsize(AD));
        d.om_layer = &w.br[b].om;
      } else {
        op_gemm(ctx, g);
      }
      if (tiled) BRN_CHECK(tc_deform_supported(d), 5, "deform: tcgen05 path unavailable");
      op_deform(ctx, d);
      arena.release(m1);
    } else {
      GemmArgs g; g.x = t; g.w = &w.br[b].reg; g.pad = k / 2; g.act = ACT_RELU; g.out = dst; op_gemm(ctx, g);
    }
  }
  float* sums = (float*)arena.alloc((size_t)B * glue_gap_blocks(H * W) * 64 * 4);
  float* pb = (float*)arena.alloc((size_t)B * 64 * 4);
  glue_gap_sum(ctx, t, sums);
  glue_aspp_pool_bias(ctx, sums, B, H * W, &w.gap, w.conv1_tail, w.bn1_shift, pb);
  View a = make_view(t.p, AD, B, H, W, 64);  // reuse t (all branches have consumed it)
  { GemmArgs g; g.x = cat; g.w = &w.conv1; g.bias = pb; g.bias_bstride = 64; g.act = ACT_RELU; g.out = a; op_gemm(ctx, g); }
  { GemmArgs g; g.x = a; g.w = conv_out ? conv_out : &w.conv_out; g.pad = 1; g.out = out; op_gemm(ctx, g); }
  arena.release(m0);
}

// BiRefNetDecoder::forward (src/birefnet.rs:278-376).  D4in[:, :lat3] already holds the squeezed x4.
void Model::run_decoder(LaunchCtx& ctx, const float* img, int B, int H, int W, View X1, View X2, View X3, View D4in,
                        float* out, bool apply_sigmoid) {
  const int AD = dec_dtype();
  const int dec_out[4] = {lat(2), lat(1), lat(0), lat(0) / 2};
  View lat_src[3] = {X3, X2, X1};
  View din = D4in;
  View p{};
  for (int d = 0; d < 4; ++d) {
    const int n = 4 - d;                       // ipt_blk{n+1} feeds decoder_block{n}
    const int hh = din.H, ww = din.W;
    const size_t px = (size_t)B * hh * ww;
    // ipt_blk{n+1}(image2patches(x)) (src/birefnet.rs:304-317), written into the tail channels of the block input;
    // the reference's later upsample_bilinear2d of it is a same-size identity (src/birefnet.rs:337,352,367).
    {
      const size_t m1 = arena.mark();
      View patches = make_view(arena.alloc(px * kIptIn[n] * dsize(AD)), AD, B, hh, ww, kIptIn[n]);
      glue_image2patches(ctx, img, B, H, W, hh, ww, patches);
      View mid = make_view(arena.alloc(px * 64 * dsize(AD)), AD, B, hh, ww, 64);
      { GemmArgs g; g.x = patches; g.w = &dw.ipt_conv1[n]; g.pad = 1; g.out = mid; op_gemm(ctx, g); }
      { GemmArgs g; g.x = mid; g.w = &dw.ipt_out[n]; g.pad = 1; g.out = din.slice(din.C - kIptOut[n], kIptOut[n]);
        op_gemm(ctx, g); }
      arena.release(m1);
    }
    if (d == 3) {
      // p1 only feeds the 1x1 output conv, which is folded into the block's last conv: q = w_p . p1 directly (fp32)
      p = make_view(arena.alloc(px * 4), F32, B, hh, ww, 1);
      run_decblk(ctx, dw.dec[d], din, p, &dw.out_q);
      break;
    }
    p = make_view(arena.alloc(px * dec_out[d] * dsize(AD)), AD, B, hh, ww, dec_out[d]);
    run_decblk(ctx, dw.dec[d], din, p);
    // GDT gate (src/birefnet.rs:327-329)
    {
      const size_t m1 = arena.mark();
      View g16 = make_view(arena.alloc(px * 16 * dsize(AD)), AD, B, hh, ww, 16);
      { GemmArgs g; g.x = p; g.w = &dw.gdt[d]; g.pad = 1; g.act = ACT_RELU; g.out = g16; op_gemm(ctx, g); }
      glue_gate(ctx, p, g16, dw.gdt_attn_w[d], dw.gdt_attn_b[d]);
      arena.release(m1);
    }
    // next block input: [ up(p) + lateral(x_k) | ipt ]  (src/birefnet.rs:332-338)
    View xs = lat_src[d];
    const int cn = dec_out[d] + kIptOut[n - 1];
    View dn = make_view(arena.alloc((size_t)B * xs.H * xs.W * cn * dsize(AD)), AD, B, xs.H, xs.W, cn);
    View head = dn.slice(0, dec_out[d]);
    glue_resize_nhwc(ctx, p, head);
    { GemmArgs g; g.x = xs; g.w = &dw.lat[d]; g.out = head; g.res = head; op_gemm(ctx, g); }
    din = dn;
  }
  // final: conv_out1(cat(up(p1), ipt_blk1(x))) (src/birefnet.rs:372-375), rewritten (Appendix F.9)
  glue_final(ctx, img, B, H, W, dw.fin_tab, dw.fin_tab_host.data(), (const float*)p.p, p.H, p.W, out, apply_sigmoid ? 1 : 0);
}

void Model::run_squeeze_decoder(LaunchCtx& ctx, const float* img, int B, int H, int W, View X1, View X2, View X3,
                                View X4cat, float* out, bool apply_sigmoid) {
  const int AD = dec_dtype();
  const int h4 = H / 32, w4 = W / 32;
  const int c4 = lat(3) + kIptOut[4];
  View D4in = make_view(arena.alloc((size_t)B * h4 * w4 * c4 * dsize(AD)), AD, B, h4, w4, c4);
  prof_begin(ctx, "squeeze");
  run_decblk(ctx, dw.squeeze, X4cat, D4in.slice(0, lat(3)));   // src/birefnet.rs:457
  prof_end(ctx);
  prof_begin(ctx, "decoder");
  run_decoder(ctx, img, B, H, W, X1, X2, X3, D4in, out, apply_sigmoid);
  prof_end(ctx);
}

static bool env_flag(const char* n);
int Model::dec_dtype() const {
  if (cfg.precision == BRN_PREC_BF16) {
    const char* v = getenv("BRN_BF16_DECODER");            // experiments: "bf16" | "fp16" overrides the handle's setting
    if (v && v[0]) return v[0] == 'f' ? F16 : BF16;
    return bf16_decoder_fp16 ? F16 : BF16;
  }
  return act_dtype();
}
static bool env_flag(const char* n) { const char* v = getenv(n); return v && v[0] && v[0] != '0'; }

// First half of BiRefNet::forward_logits (src/birefnet.rs:412-454): both backbone passes, the multi-scale concat
// (x_k = [bb(x)_k | up(bb(x_half)_k)]) and the cxt concat into X4cat.  X[0..2] and X4cat are caller-allocated.
void Model::run_features(LaunchCtx& ctx, const float* img, int B, int H, int W, View X[3], View X4cat) {
  const int AD = dec_dtype();
  const int off4 = lat(0) + lat(1) + lat(2);
  View feats[4] = {X[0].slice(0, C(0)), X[1].slice(0, C(1)), X[2].slice(0, C(2)), X4cat.slice(off4, C(3))};
  // (the SIMT attention kernel handles one token grid per launch, so window-7 models run the two passes separately)
  const bool merged = cfg.precision != BRN_PREC_FP32 && !ctx.force_simt && cfg.window_size == 12 && !env_flag("BRN_SPLIT_BACKBONE");
  {
    const size_t m1 = arena.mark();
    const int H2 = H / 2, W2 = W / 2;
    float* half = (float*)arena.alloc((size_t)B * 3 * H2 * W2 * 4);
    // token grids of the half-resolution pass: the backbone's own chain (PatchEmbed /4, then PatchMerging rounds
    // odd grids UP: src/swin.rs:496-503,586) -- not (H/4 >> i) / 2, which rounds down when H/32 is odd
    View fh[4];
    int hh = H2 / 4, wh = W2 / 4;
    for (int i = 0; i < 4; ++i) {
      fh[i] = make_view(arena.alloc((size_t)B * hh * wh * C(i) * dsize(AD)), AD, B, hh, wh, C(i));
      hh = (hh + 1) / 2; wh = (wh + 1) / 2;
    }
    if (merged) {
      prof_begin(ctx, "backbone_full+half");
      glue_resize_nchw(ctx, img, B, 3, H, W, half, H2, W2);                    // :425
      run_backbone(ctx, img, B, H, W, feats, half, H2, W2, fh);                // :416-420 and :426 in one pass
      prof_end(ctx);
      prof_begin(ctx, "half_upsample");
    } else {
      prof_begin(ctx, "backbone_full");
      run_backbone(ctx, img, B, H, W, feats);                                  // :416-420
      prof_end(ctx);
      prof_begin(ctx, "backbone_half");
      glue_resize_nchw(ctx, img, B, 3, H, W, half, H2, W2);                    // :425
      run_backbone(ctx, half, B, H2, W2, fh);                                  // :426
    }
    View dst[4] = {X[0].slice(C(0), C(0)), X[1].slice(C(1), C(1)), X[2].slice(C(2), C(2)), X4cat.slice(off4 + C(3), C(3))};
    for (int i = 0; i < 4; ++i) glue_resize_nhwc(ctx, fh[i], dst[i]);        // :435-443
    arena.release(m1);
    prof_end(ctx);
  }
  prof_begin(ctx, "cxt");
  glue_resize_nhwc(ctx, X[0], X4cat.slice(0, lat(0)));                       // :450-453
  glue_resize_nhwc(ctx, X[1], X4cat.slice(lat(0), lat(1)));
  glue_resize_nhwc(ctx, X[2], X4cat.slice(lat(0) + lat(1), lat(2)));
  prof_end(ctx);
}

// BiRefNet::forward_logits (src/birefnet.rs:412-461)
void Model::run_forward(LaunchCtx& ctx, const float* img, int B, int H, int W, float* out, bool apply_sigmoid) {
  const int AD = dec_dtype();
  const size_t m0 = arena.mark();
  // scratch for split-K partial sums (small batches: the 32x32-level decoder convs have 8 output tiles per image)
  ctx.splitk_bytes = tc_gemm_splitk_scratch_bytes(B);
  ctx.splitk = (float*)arena.alloc(ctx.splitk_bytes);
  int hs[4], ws[4];
  for (int i = 0; i < 4; ++i) { hs[i] = H / (4 << i); ws[i] = W / (4 << i); }
  View X[3];
  for (int i = 0; i < 3; ++i)
    X[i] = make_view(arena.alloc((size_t)B * hs[i] * ws[i] * lat(i) * dsize(AD)), AD, B, hs[i], ws[i], lat(i));
  const int c4 = x4_channels();
  View X4cat = make_view(arena.alloc((size_t)B * hs[3] * ws[3] * c4 * dsize(AD)), AD, B, hs[3], ws[3], c4);
  run_features(ctx, img, B, H, W, X, X4cat);
  run_squeeze_decoder(ctx, img, B, H, W, X[0], X[1], X[2], X4cat, out, apply_sigmoid);
  arena.release(m0);
}

void Model::forward(const float* x, int B, int H, int W, bool x_dev, float* out, bool out_dev, cudaStream_t s,
                    bool apply_sigmoid) {
  BRN_CHECK(finalized, 6, "forward before finalize");
  BRN_CHECK(x && out && B > 0, 1, "forward: bad argument");
  BRN_CHECK(H > 0 && W > 0 && H % 32 == 0 && W % 32 == 0, 5, "H and W must be positive multiples of 32");
  DeviceGuard dg(device);
  std::unique_lock<std::mutex> lk(mu);
  const int li = acquire_lane(lk, s);
  cudaStream_t st = s ? s : lanes[li].stream;
  lanes[li].last = st;
  const float* last_src = nullptr; float* last_dst = nullptr; size_t last_bytes = 0;
  try {
  // an earlier call may have used this lane's workspace on another stream
  BRN_CUDA(cudaStreamWaitEvent(st, lanes[li].done, 0));
  const int mb = micro_batch(B, H, W);
  // plan pass
  LaunchCtx ctx; ctx.stream = st; ctx.precision = cfg.precision; ctx.force_simt = env_flag("BRN_FORCE_SIMT");
  long long dummy = 0;
  arena.dry = true; arena.off = 0; arena.peak = 0;
  ctx.dry = true; ctx.launches = &dummy;
  const size_t in_bytes = (size_t)mb * 3 * H * W * 4, out_bytes = (size_t)mb * H * W * 4;
  float* din = x_dev ? nullptr : (float*)arena.alloc(in_bytes);
  float* dout = out_dev ? nullptr : (float*)arena.alloc(out_bytes);
  run_forward(ctx, (const float*)0x10000, mb, H, W, (float*)0x10000, apply_sigmoid);
  ensure_arena(arena.peak);
  arena.dry = false; ctx.dry = false; ctx.launches = &launches;
  if (prof_on >= 2) { ctx.kt = &ktimer; ktimer.reset(); }
  if (prof_on) {
    for (auto& pe : prof) { cudaEventDestroy(pe.e0); cudaEventDestroy(pe.e1); }
    prof.clear();
  }
  static const bool graph_env_off = [] { const char* v = getenv("BRN_CUDA_GRAPH"); return v && v[0] == '0'; }();
  for (int b0 = 0; b0 < B; b0 += mb) {
    const int nb = std::min(mb, B - b0);
    arena.off = 0;
    din = x_dev ? nullptr : (float*)arena.alloc(in_bytes);
    dout = out_dev ? nullptr : (float*)arena.alloc(out_bytes);
    const float* xi = x + (size_t)b0 * 3 * H * W;
    float* oi = out + (size_t)b0 * H * W;
    if (!x_dev) { BRN_CUDA(cudaMemcpyAsync(din, xi, (size_t)nb * 3 * H * W * 4, cudaMemcpyHostToDevice, st)); xi = din; }
    float* ko = out_dev ? oi : dout;
    bool done = false;
    // the legacy / per-thread default streams cannot be captured
    const bool capturable = st != nullptr && st != cudaStreamLegacy && st != cudaStreamPerThread;
    if (use_graph && !graph_env_off && !prof_on && !ctx.force_simt && capturable) {
      GraphKey key{xi, ko, arena.base, nb, H, W, (int)cfg.precision | (dec_dtype() << 8), (int)cfg.deform_mode, apply_sigmoid ? 1 : 0};
      GraphEntry* e = nullptr;
      for (auto& g : graphs) if (g.key == key) { e = &g; break; }
      if (!e) {
        if (graphs.size() >= 16) {     // evict the least recently used entry
          size_t lru = 0;
          for (size_t i = 1; i < graphs.size(); ++i) if (graphs[i].stamp < graphs[lru].stamp) lru = i;
          if (graphs[lru].exec) cudaGraphExecDestroy(graphs[lru].exec);
          graphs.erase(graphs.begin() + lru);
        }
        graphs.push_back(GraphEntry{});
        e = &graphs.back();
        e->key = key;
      }
      e->stamp = ++graph_clock;
      if (!e->exec && e->seen >= 1) {
        // second sighting: capture (nothing executes during capture), instantiate, then replay below
        long long counted = 0;
        LaunchCtx cctx = ctx; cctx.launches = &counted;
        cudaGraph_t graph = nullptr;
        BRN_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        try {
          run_forward(cctx, xi, nb, H, W, ko, apply_sigmoid);
        } catch (...) {
          cudaStreamEndCapture(st, &graph);
          if (graph) cudaGraphDestroy(graph);
          throw;
        }
        BRN_CUDA(cudaStreamEndCapture(st, &graph));
        cudaError_t ie = cudaGraphInstantiate(&e->exec, graph, 0);
        cudaGraphDestroy(graph);
        BRN_CHECK(ie == cudaSuccess, 2, std::string("cudaGraphInstantiate failed: ") + cudaGetErrorString(ie));
        e->launches = counted;
      }
      if (e->exec) {
        BRN_CUDA(cudaGraphLaunch(e->exec, st));
        launches += e->launches;
        done = true;
      }
      e->seen++;
    }
    if (!done) run_forward(ctx, xi, nb, H, W, ko, apply_sigmoid);
    if (!out_dev) {
      // the last read-back is issued after `mu` is dropped (a pageable destination makes the copy block the host)
      if (b0 + mb >= B && !prof_on) { last_src = dout; last_dst = oi; last_bytes = (size_t)nb * H * W * 4; }
      else BRN_CUDA(cudaMemcpyAsync(oi, dout, (size_t)nb * H * W * 4, cudaMemcpyDeviceToHost, st));
    }
  }
  if (prof_on) BRN_CUDA(cudaStreamSynchronize(st));
  collect_profile();
  } catch (...) {
    cudaStreamSynchronize(st);
    std::swap(arena, lanes[li].arena);
    release_lane(li);
    throw;
  }
  // host-side state is done with: hand the workspace back to the lane and wait for the device without the lock
  std::swap(arena, lanes[li].arena);
  cudaError_t werr = cudaSuccess;
  if (last_bytes || ((!x_dev || !out_dev) && !prof_on)) {
    lk.unlock();
    if (last_bytes) werr = cudaMemcpyAsync(last_dst, last_src, last_bytes, cudaMemcpyDeviceToHost, st);
    if (werr == cudaSuccess) werr = cudaStreamSynchronize(st);
    lk.lock();
  } else {
    werr = cudaEventRecord(lanes[li].done, st);
  }
  release_lane(li);
  BRN_CHECK(werr == cudaSuccess, 2, std::string("forward: ") + cudaGetErrorString(werr));
}

// after a profiled pass has completed on the device: per-class sums of the per-launch events, the optional per-launch
// CSV (BRN_KERNEL_LOG) and the per-stage times
void Model::collect_profile() {
  if (prof_on >= 2) {
    for (int c = 0; c < KC_COUNT; ++c) { kc_ms[c] = 0; kc_flops[c] = 0; kc_bytes[c] = 0; kc_count[c] = 0; }
    for (auto& r : ktimer.recs) {
      float ms = 0; cudaEventElapsedTime(&ms, r.e0, r.e1);
      kc_ms[r.cls] += ms; kc_flops[r.cls] += r.flops; kc_bytes[r.cls] += r.bytes; kc_count[r.cls]++;
    }
    if (const char* path = getenv("BRN_KERNEL_LOG")) {
      if (FILE* f = fopen(path, "w")) {
        fprintf(f, "class,ms,gflop,mbytes,desc\n");
        for (auto& r : ktimer.recs) {
          float ms = 0; cudaEventElapsedTime(&ms, r.e0, r.e1);
          fprintf(f, "%d,%.5f,%.4f,%.3f,%s\n", r.cls, ms, r.flops / 1e9, r.bytes / 1e6, r.desc.c_str());
        }
        fclose(f);
      }
    }
  }
  if (prof_on) {
    prof_names.clear(); prof_ms.clear(); prof_flops.clear();
    for (auto& pe : prof) {
      float ms = 0; cudaEventElapsedTime(&ms, pe.e0, pe.e1);
      prof_names.push_back(pe.name.c_str()); prof_ms.push_back(ms); prof_flops.push_back(pe.flops);
    }
  }
}
// profiled passes of the partial entry points (backbone / features / decoder): same event bookkeeping as forward()
void Model::begin_profile(LaunchCtx& ctx) {
  if (prof_on >= 2) { ctx.kt = &ktimer; ktimer.reset(); }
  if (prof_on) {
    for (auto& pe : prof) { cudaEventDestroy(pe.e0); cudaEventDestroy(pe.e1); }
    prof.clear();
  }
}

void Model::backbone_api(const float* x, int B, int H, int W, bool x_dev, float* const outs[4], bool out_dev,
                         cudaStream_t s) {
  BRN_CHECK(finalized, 6, "backbone_forward before finalize");
  BRN_CHECK(H > 0 && W > 0 && H % 32 == 0 && W % 32 == 0, 5, "H and W must be positive multiples of 32");
  DeviceGuard dg(device);
  std::lock_guard<std::mutex> lk(mu);
  cudaStream_t st = s ? s : own_stream;
  const int AD = act_dtype();
  LaunchCtx ctx; ctx.stream = st; ctx.precision = cfg.precision; ctx.force_simt = env_flag("BRN_FORCE_SIMT");
  long long dummy = 0;
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1) begin_profile(ctx);
    arena.dry = pass == 0; ctx.dry = pass == 0; ctx.launches = pass == 0 ? &dummy : &launches;
    arena.off = 0; if (pass == 0) arena.peak = 0;
    float* din = (float*)arena.alloc((size_t)B * 3 * H * W * 4);
    View f[4]; float* dn[4];
    for (int i = 0; i < 4; ++i) {
      int hh = H / (4 << i), ww = W / (4 << i);
      f[i] = make_view(arena.alloc((size_t)B * hh * ww * C(i) * dsize(AD)), AD, B, hh, ww, C(i));
      dn[i] = (float*)arena.alloc((size_t)B * hh * ww * C(i) * 4);
    }
    if (pass == 1) {
      if (x_dev) BRN_CUDA(cudaMemcpyAsync(din, x, (size_t)B * 3 * H * W * 4, cudaMemcpyDeviceToDevice, st));
      else BRN_CUDA(cudaMemcpyAsync(din, x, (size_t)B * 3 * H * W * 4, cudaMemcpyHostToDevice, st));
    }
    run_backbone(ctx, din, B, H, W, f);
    for (int i = 0; i < 4; ++i) {
      glue_nhwc_to_nchw(ctx, f[i], dn[i]);
      if (pass == 1)
        BRN_CUDA(cudaMemcpyAsync(outs[i], dn[i], (size_t)f[i].rows() * C(i) * 4,
                                 out_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    }
    if (pass == 0) ensure_arena(arena.peak);
  }
  BRN_CUDA(cudaStreamSynchronize(st));
  collect_profile();
}

// x1..x3 and the cxt-concatenated x4 (src/birefnet.rs:412-454) as NCHW fp32: the inputs brn_decoder_forward takes
void Model::features_api(const float* x, int B, int H, int W, bool x_dev, float* const outs[4], bool out_dev,
                         cudaStream_t s) {
  BRN_CHECK(finalized, 6, "features_forward before finalize");
  BRN_CHECK(x && outs && B > 0, 1, "features_forward: bad argument");
  BRN_CHECK(H > 0 && W > 0 && H % 32 == 0 && W % 32 == 0, 5, "H and W must be positive multiples of 32");
  DeviceGuard dg(device);
  std::lock_guard<std::mutex> lk(mu);
  cudaStream_t st = s ? s : own_stream;
  const int AD = dec_dtype();
  LaunchCtx ctx; ctx.stream = st; ctx.precision = cfg.precision; ctx.force_simt = env_flag("BRN_FORCE_SIMT");
  long long dummy = 0;
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1) begin_profile(ctx);
    arena.dry = pass == 0; ctx.dry = pass == 0; ctx.launches = pass == 0 ? &dummy : &launches;
    arena.off = 0; if (pass == 0) arena.peak = 0;
    float* din = (float*)arena.alloc((size_t)B * 3 * H * W * 4);
    View X[4]; float* dn[4];
    for (int i = 0; i < 4; ++i) {
      const int hh = H / (4 << i), ww = W / (4 << i), c = i < 3 ? lat(i) : x4_channels();
      X[i] = make_view(arena.alloc((size_t)B * hh * ww * c * dsize(AD)), AD, B, hh, ww, c);
      dn[i] = (float*)arena.alloc((size_t)B * hh * ww * c * 4);
    }
    if (pass == 1)
      BRN_CUDA(cudaMemcpyAsync(din, x, (size_t)B * 3 * H * W * 4, x_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    run_features(ctx, din, B, H, W, X, X[3]);
    for (int i = 0; i < 4; ++i) {
      glue_nhwc_to_nchw(ctx, X[i], dn[i]);
      if (pass == 1)
        BRN_CUDA(cudaMemcpyAsync(outs[i], dn[i], (size_t)X[i].rows() * X[i].C * 4,
                                 out_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    }
    if (pass == 0) ensure_arena(arena.peak);
  }
  BRN_CUDA(cudaStreamSynchronize(st));
  collect_profile();
}

void Model::decoder_api(const float* x, const float* x1, const float* x2, const float* x3, const float* x4, int B,
                        int H, int W, bool is_dev, float* out, cudaStream_t s) {
  BRN_CHECK(finalized, 6, "decoder_forward before finalize");
  BRN_CHECK(H > 0 && W > 0 && H % 32 == 0 && W % 32 == 0, 5, "H and W must be positive multiples of 32");
  DeviceGuard dg(device);
  std::lock_guard<std::mutex> lk(mu);
  cudaStream_t st = s ? s : own_stream;
  const int AD = dec_dtype();
  LaunchCtx ctx; ctx.stream = st; ctx.precision = cfg.precision; ctx.force_simt = env_flag("BRN_FORCE_SIMT");
  long long dummy = 0;
  const cudaMemcpyKind kin = is_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  const float* srcs[4] = {x1, x2, x3, x4};
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1) begin_profile(ctx);
    arena.dry = pass == 0; ctx.dry = pass == 0; ctx.launches = pass == 0 ? &dummy : &launches;
    arena.off = 0; if (pass == 0) arena.peak = 0;
    float* dimg = (float*)arena.alloc((size_t)B * 3 * H * W * 4);
    float* dout = (float*)arena.alloc((size_t)B * H * W * 4);
    if (pass == 1) BRN_CUDA(cudaMemcpyAsync(dimg, x, (size_t)B * 3 * H * W * 4, kin, st));
    View X[4];
    for (int i = 0; i < 4; ++i) {
      int hh = H / (4 << i), ww = W / (4 << i);
      int c = i < 3 ? lat(i) : x4_channels();
      X[i] = make_view(arena.alloc((size_t)B * hh * ww * c * dsize(AD)), AD, B, hh, ww, c);
      const size_t m1 = arena.mark();
      float* tmp = (float*)arena.alloc((size_t)B * hh * ww * c * 4);
      if (pass == 1) BRN_CUDA(cudaMemcpyAsync(tmp, srcs[i], (size_t)B * hh * ww * c * 4, kin, st));
      glue_nchw_to_nhwc(ctx, tmp, B, c, hh, ww, X[i]);
      arena.release(m1);
    }
    run_squeeze_decoder(ctx, dimg, B, H, W, X[0], X[1], X[2], X[3], dout, false);
    if (pass == 1)
      BRN_CUDA(cudaMemcpyAsync(out, dout, (size_t)B * H * W * 4, is_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    if (pass == 0) ensure_arena(arena.peak);
  }
  BRN_CUDA(cudaStreamSynchronize(st));
  collect_profile();
}

}  // namespace brn
