// Model handle: weight schema, finalize (folding + upload) and the forward graph.  Private to the library.
#pragma once
#include <condition_variable>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/birefnet_b200.h"
#include "brn_common.h"

namespace brn {

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
  bool set = false;
  size_t numel() const { size_t n = 1; for (auto d : shape) n *= (size_t)d; return n; }
};

// Bump allocator over one device buffer, reset per forward.  A dry "plan" pass measures the peak first.
struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0, peak = 0;
  bool dry = false;
  void* alloc(size_t bytes) {
    off = (off + 1023) & ~(size_t)1023;
    size_t o = off;
    off += bytes;
    if (off > peak) peak = off;
    if (dry) return (void*)(uintptr_t)(0x10000 + o);
    if (off > cap) throw Error(2, "arena overflow (plan pass under-estimated the workspace)");
    return base + o;
  }
  size_t mark() const { return off; }
  void release(size_t m) { off = m; }
};

struct BlockW {
  float *n1g, *n1b, *n2g, *n2b;
  LayerW qkv, proj, fc1, fc2;
  LayerW qkv_f, fc1_f;     // norm1 / norm2 folded in (gamma into the weights, W beta into the bias): tensor-core path
  uint16_t* qkv_bias16 = nullptr;   // [2][3C]: the q-scaled qkv bias as bf16, then as fp16 (qkv of a pad token)
  float* bias32;           // [heads][144][144]
  float* bias32p;          // [heads][144][148] (padded rows, tcgen05 attention kernel)
};
struct StageW {
  std::vector<BlockW> blocks;
  bool has_down = false;
  float *dng = nullptr, *dnb = nullptr;
  LayerW red;
  float *ng = nullptr, *nb = nullptr;
};
struct AsppBranchW {
  int k = 1;
  LayerW om;   // offset_conv ++ modulator_conv -> [3k^2, 64, k, k] (+bias)
  LayerW reg;  // regular_conv with BatchNorm folded -> [256, 64, k, k] (+bias = bn shift)
};
struct DecBlkW {
  LayerW conv_in;        // + bn_in folded
  AsppBranchW br[4];     // aspp1, aspp_deforms[0..2]
  LayerW gap;            // global_avg_pool.1 + .2 folded: [256][64]
  LayerW conv1;          // conv1[:, :1024] * bn1 scale, no bias
  float* conv1_tail = nullptr;  // [64][256] = conv1[:, 1024:1280] * bn1 scale
  float* bn1_shift = nullptr;   // [64]
  LayerW conv_out;       // + bn_out folded
};
struct DecoderW {
  LayerW ipt_conv1[5], ipt_out[5];   // index n-1 for ipt_blk{n}; [0] unused (fused final kernel)
  DecBlkW squeeze, dec[4];           // dec[0] = decoder_block4 ... dec[3] = decoder_block1
  LayerW lat[3];                     // lateral_block4,3,2
  LayerW gdt[3];                     // gdt_convs_{4,3,2}.0 + .1 folded
  float* gdt_attn_w[3] = {nullptr, nullptr, nullptr};
  float gdt_attn_b[3] = {0, 0, 0};
  LayerW out_q;                      // conv_out1 (p1 channels) folded into decoder_block1.conv_out + bn_out: 3x3, 64 -> 1
  float* fin_tab = nullptr;          // folded final-layer table (final_kernel.cu)
  std::vector<float> fin_tab_host;   // host copy: the 5x5 kernel is passed to final_kernel as a launch parameter
};

struct ProfEntry {
  std::string name;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  double flops = 0;
};

struct Model {
  brn_config cfg{};
  int device = 0;
  cudaStream_t own_stream = nullptr;
  std::mutex mu;
  std::vector<std::string> keys;
  std::unordered_map<std::string, int> index;
  std::vector<HostTensor> tensors;
  bool finalized = false;
  std::vector<void*> allocs;

  StageW stages[4];
  LayerW patch_embed;
  float *pe_g = nullptr, *pe_b = nullptr;
  DecoderW dw;

  // `arena` is the workspace the host-side code is currently planning / launching into (always accessed under `mu`).
  // forward() borrows one of two lanes (workspace + stream + completion event) for the duration of a call and
  // swaps the lane's arena in while it holds `mu`; it drops `mu` while it waits for the device, so two host threads
  // can keep two calls in flight: the host<->device copies of one overlap the kernels of the other.
  Arena arena;
  struct Lane { Arena arena; cudaStream_t stream = nullptr, last = nullptr; cudaEvent_t done = nullptr; bool busy = false; };
  static constexpr int kLanes = 2;
  Lane lanes[kLanes];
  std::condition_variable lane_cv;
  int acquire_lane(std::unique_lock<std::mutex>& lk, cudaStream_t s);
  void release_lane(int i);
  void quiesce(std::unique_lock<std::mutex>& lk);     // wait until no forward() is in flight
  long long launches = 0;

  // CUDA-graph replay of the forward: the launch sequence of run_forward for one (buffers, shape, mode) key is
  // captured the second time the key is seen (the first call runs eagerly: lazy module loading, attribute setup) and
  // replayed afterwards, which removes ~310 launch gaps per pass (batch-1 latency is launch-bound otherwise).
  struct GraphKey {
    const void* x; const void* out; void* arena_base; int B, H, W, precision, deform, sigmoid;
    bool operator==(const GraphKey& o) const {
      return x == o.x && out == o.out && arena_base == o.arena_base && B == o.B && H == o.H && W == o.W &&
             precision == o.precision && deform == o.deform && sigmoid == o.sigmoid;
    }
  };
  struct GraphEntry { GraphKey key; cudaGraphExec_t exec = nullptr; long long launches = 0; int seen = 0; long long stamp = 0; };
  std::vector<GraphEntry> graphs;
  long long graph_clock = 0;
  int use_graph = 1;
  void drop_graphs(const void* arena_base = nullptr);   // nullptr: all

  int prof_on = 0;          // 1: per-stage events, 2: + per-kernel-class events (KTimer)
  KTimer ktimer;
  float kc_ms[KC_COUNT] = {0};
  double kc_flops[KC_COUNT] = {0};
  double kc_bytes[KC_COUNT] = {0};
  int kc_count[KC_COUNT] = {0};
  std::vector<ProfEntry> prof;
  std::vector<const char*> prof_names;
  std::vector<float> prof_ms;
  std::vector<double> prof_flops;

  // derived
  int C(int i) const { return cfg.embed_dim << i; }
  int lat(int i) const { return 2 * C(i); }
  int x4_channels() const { return lat(0) + lat(1) + lat(2) + lat(3); }

  explicit Model(const brn_config& c, int dev);
  ~Model();
  void build_schema();
  void set_tensor(const char* key, const void* data, int dtype, const int64_t* shape, int rank);
  void finalize();

  // graph
  void forward(const float* x, int B, int H, int W, bool x_dev, float* out, bool out_dev, cudaStream_t s,
               bool apply_sigmoid);
  void backbone_api(const float* x, int B, int H, int W, bool x_dev, float* const outs[4], bool out_dev,
                    cudaStream_t s);
  void features_api(const float* x, int B, int H, int W, bool x_dev, float* const outs[4], bool out_dev,
                    cudaStream_t s);
  void decoder_api(const float* x, const float* x1, const float* x2, const float* x3, const float* x4, int B, int H,
                   int W, bool is_dev, float* out, cudaStream_t s);

 private:
  void run_features(LaunchCtx& ctx, const float* img, int B, int H, int W, View X[3], View X4cat);
  void run_forward(LaunchCtx& ctx, const float* img, int B, int H, int W, float* out, bool apply_sigmoid);
  void run_backbone(LaunchCtx& ctx, const float* img, int B, int H, int W, View feats[4], const float* img2 = nullptr,
                    int H2 = 0, int W2 = 0, View* feats2 = nullptr);
  void run_decblk(LaunchCtx& ctx, const DecBlkW& w, View in, View out, const LayerW* conv_out = nullptr);
  void run_decoder(LaunchCtx& ctx, const float* img, int B, int H, int W, View X1, View X2, View X3, View D4in,
                   float* out, bool apply_sigmoid);
  void run_squeeze_decoder(LaunchCtx& ctx, const float* img, int B, int H, int W, View X1, View X2, View X3,
                           View X4cat, float* out, bool apply_sigmoid);
  void ensure_arena(size_t bytes);
  int micro_batch(int B, int H, int W) const;
  // element type of stored activations / GEMM operands for the current precision (fp32 | bf16 | fp16)
  int act_dtype() const { return cfg.precision == BRN_PREC_BF16 ? BF16 : cfg.precision == BRN_PREC_FP16 ? F16 : F32; }
  // squeeze module + decoder operands.  precision = bf16 keeps bf16 in the backbone (fp32 residual stream, wide dynamic
  // range) and, with bf16_decoder_fp16 set, runs the BN-normalised O(1) decoder activations in fp16 (same tensor-core
  // rate, 3 more mantissa bits -- what IoU >= 0.999 on near-threshold logits needs)
  int bf16_decoder_fp16 = 1;
  int dec_dtype() const;
  void begin_profile(LaunchCtx& ctx);
  void collect_profile();
  void prof_begin(LaunchCtx& ctx, const char* name);
  void prof_end(LaunchCtx& ctx);

  const HostTensor& T(const std::string& k) const;
  float* upload(const std::vector<float>& v);
  LayerW make_layer(int N, int Cin, int kh, int kw, const std::vector<float>& w_oihw, const std::vector<float>* bias,
                    bool folded_ln = false);
  LayerW make_folded(const std::vector<float>& w, const std::vector<float>& b, const std::vector<float>& gamma,
                     const std::vector<float>& beta, int N, int C);
};

}  // namespace brn

// sharded.cpp: one handle + one host thread per GPU behind a single C-ABI object
struct brn_sharded;
namespace brn {
brn_sharded* sharded_create(const brn_config& cfg, const int* devices, int n);
void sharded_set_tensor(brn_sharded*, const char* key, const void* data, int dtype, const int64_t* shape, int rank);
int sharded_load_safetensors(brn_sharded*, const char* path);
void sharded_finalize(brn_sharded*);
void sharded_forward(brn_sharded*, const float* x, int B, int H, int W, float* out, bool apply_sigmoid);

// safetensors_loader.cpp: sets every tensor of `path` that the schema knows; returns how many were set
int load_safetensors(Model& m, const char* path);

// generic dispatchers (precision + support -> tcgen05 or SIMT)
void op_gemm(const LaunchCtx&, const GemmArgs&);
void op_deform(const LaunchCtx&, const DeformArgs&);
void op_attention(const LaunchCtx&, const AttnArgs&);

// standalone layer upload for the operator-level ABI
LayerW make_layer_standalone(int N, int Cin, int kh, int kw, const float* w_oihw, const float* bias,
                             std::vector<void*>& allocs, bool folded_ln = false);

}  // namespace brn
