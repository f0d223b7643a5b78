// Shared host-side declarations for libbirefnet_b200 (private; the public ABI is include/birefnet_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdlib.h>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace brn {

struct Error : std::runtime_error {
  int status;
  Error(int s, const std::string& m) : std::runtime_error(m), status(s) {}
};

#define BRN_CUDA(call)                                                                              \
  do {                                                                                              \
    cudaError_t e__ = (call);                                                                       \
    if (e__ != cudaSuccess)                                                                         \
      throw ::brn::Error(2, std::string(#call) + " failed: " + cudaGetErrorString(e__) + " at " +   \
                                __FILE__ + ":" + std::to_string(__LINE__));                         \
  } while (0)

#define BRN_CHECK(cond, status, msg)                        \
  do {                                                      \
    if (!(cond)) throw ::brn::Error((status), (msg));       \
  } while (0)

// cudaSetDevice for the duration of an entry point; the caller's current device is restored on exit
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) BRN_CUDA(cudaSetDevice(dev));
    else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Activation / operand element types.  The 16-bit tensor-core path uses bf16 in the backbone (fp32 residual stream,
// wide dynamic range) and fp16 in the squeeze module + decoder (BN-normalised, O(1) activations; 3 more mantissa
// bits are what the IoU >= 0.999 tolerance on near-zero logits needs, see DESIGN.md "precision").
enum DType { F32 = 0, BF16 = 1, F16 = 2 };
inline size_t dsize(int dt) { return dt == F32 ? 4 : 2; }

// NHWC activation view.  Element (b,y,x,c) lives at p + (((b*H+y)*W+x)*ld + c) elements.  Token matrices
// [rows, C] are views with B=1,H=1,W=rows.  `ld` >= C lets a producer write into a channel slice of a wider
// (concatenated) buffer, which is how every Tensor::cat of the reference disappears.
struct View {
  void* p = nullptr;
  int dt = F32;
  int B = 1, H = 1, W = 1, C = 0;
  int ld = 0;
  long long rows() const { return (long long)B * H * W; }
  View slice(int c0, int c) const {
    View v = *this;
    v.p = (char*)p + (size_t)c0 * dsize(dt);
    v.C = c;
    return v;
  }
};

inline View make_view(void* p, int dt, int B, int H, int W, int C, int ld = 0) {
  View v;
  v.p = p; v.dt = dt; v.B = B; v.H = H; v.W = W; v.C = C; v.ld = ld ? ld : C;
  return v;
}

// Weights of one conv / linear layer after finalize (BN folded, tap-major K order).
struct LayerW {
  int N = 0;        // output channels
  int Cin = 0;      // input channels
  int kh = 1, kw = 1;
  int cin_pad = 0;  // Cin rounded up to 64 (K pitch per tap of the bf16 matrix)
  float* w32 = nullptr;          // [N][kh*kw][Cin]      fp32 (SIMT path)
  void* w_bf16 = nullptr;        // [N][kh*kw][cin_pad]  bf16, zero padded (tcgen05 path, precision bf16)
  void* w_fp16 = nullptr;        // same, fp16 (precision fp16)
  const void* w16(int dt) const { return dt == F16 ? w_fp16 : w_bf16; }
  float* bias = nullptr;         // [N] fp32 or null
  // LayerNorm folded into this linear layer (LnFold): sum_c of the ROUNDED 16-bit weights per output row, one array per
  // operand type (the epilogue subtracts mean * colsum, which must cancel against what the MMA really multiplied)
  float* colsum_bf16 = nullptr;
  float* colsum_fp16 = nullptr;
  const float* colsum(int dt) const { return dt == F16 ? colsum_fp16 : colsum_bf16; }
  // host copies for the folded layers: [colsum_bf16 (N) | colsum_fp16 (N) | bias (N)] -- the fused MLP kernel takes its
  // per-column constants as kernel parameters (constant bank) instead of reading them through the shared-memory pipe
  std::shared_ptr<std::vector<float>> h_fold;
  std::shared_ptr<std::vector<float>> h_bias;     // host copy of `bias` (layers with N <= 1024)
  int taps() const { return kh * kw; }
};

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_2SIGMOID_TAIL = 3 };

// window-reverse row map of the proj GEMM epilogue (src/swin.rs:387-401): window-ordered padded row -> token row
// Two segments: rows [0, split) use geometry (h, w, hp, wp) and map to token rows [0, ...); rows [split, ...) belong
// to a second token grid (the half-resolution backbone pass runs merged with the full-resolution one: same weights,
// token matrices concatenated along M) with geometry (h2, w2, hp2, wp2), its tokens starting at row tok2.
struct RowMap {
  int enabled = 0;               // 1: window-ordered padded row -> token row (proj; pad rows are skipped)
                                 // 2: token row -> window-ordered padded row (qkv computed in token order and scattered
                                 //    into the layout the attention kernel loads; pad rows are never written)
  int h = 0, w = 0, hp = 0, wp = 0, shift = 0;
  int ws = 12;                   // window side (12: swin_b / swin_l, 7: swin_t / swin_s)
  long long split = 0;           // 0: single segment
  int h2 = 0, w2 = 0, hp2 = 0, wp2 = 0;
  long long tok2 = 0;
};

// LayerNorm folded into the GEMM that consumes it (src/swin.rs:355,407 norm1 -> qkv,
// norm2 -> fc1):  LN(x) W^T + b  =  rstd * (x (gamma .* W)^T  -  mean * colsum(gamma .* W))  +  (W beta + b).
// The GEMM runs on the RAW 16-bit copy of the residual stream with gamma folded into the weights; the per-row
// statistics come from the epilogue that produced the stream (LnEmit) as `parts` partial (sum, sum of squares)
// pairs per row, summed here in a fixed order (deterministic: results do not depend on the batch an image is in).
struct LnFold {
  const float2* mr = nullptr;      // [rows] (-mean, rstd) of the fp32 stream (glue_ln_finalize); null: off
  int C = 0;                       // row length the statistics cover (checked against the layer's K)
};
// Producer side: the epilogue that writes the fp32 residual stream also writes its raw 16-bit copy and the row
// statistics of the values it stores (one partial per (N tile, column part)).
struct LnEmit {
  float2* stats = nullptr;         // [parts][stride]; null: off
  long long stride = 0;
  void* x16 = nullptr; int x16dt = BF16; int ldx16 = 0;
};
int tc_gemm_ln_parts(int N);       // partials per row an LnEmit epilogue writes for an N-column output

// One implicit-GEMM problem: out[m, n] = act(sum_k A[m,k] W[n,k] + bias) (+ res[m,n]).
// A is the conv patch matrix of `x` (NHWC, stride 1, zero padding `pad`), K order (tap, channel).
struct GemmArgs {
  View x;                 // input activation
  const LayerW* w = nullptr;
  int pad = 0;
  int stride = 1;                // > 1 (or pad != k/2): SIMT kernels only; the model's convs are all stride 1, "same"
  const float* bias = nullptr;   // overrides w->bias when non-null
  int bias_bstride = 0;          // >0: bias is [B][N] (per-image bias; ASPP global-pool branch)
  int act = ACT_NONE;
  int act_from = 0;              // ACT_2SIGMOID_TAIL: columns >= act_from get 2*sigmoid (modulator, src/aspp.rs:174)
  View res;                      // optional residual (res.p == nullptr: none); indexed like out
  View out;                      // output view (dtype decides conversion)
  RowMap rowmap;                 // optional output row scatter
  int tile_w = 0;                // conv M tile = tile_w x (128 / tile_w) pixels (power of two); 0 = widest that fits W
  int out_tiled = 0;             // 1: fp32 out is [m_tile][N][128] (tile-major, row-in-tile fastest): the layout the
                                 //    deformable gather reads its offsets from with coalesced loads
  LnFold lnf;                    // consumer of a folded LayerNorm (tcgen05 path only)
  LnEmit lne;                    // producer of the statistics + raw 16-bit copy (tcgen05 path only)
  double flops = 0;
};

// Fused Swin MLP on the tensor-core path (mlp_tcgen05.cu): xt <- xt + fc2(gelu(fc1(LN(xt)))) with the LayerNorm folded
// (x16 = raw 16-bit copy of xt, mr = its row statistics).  The kernel writes the updated rows' raw copy to lne.x16 (may
// be x16 itself) and their (-mean, rstd) to mr_out (may be mr): a tile's rows are read before they are rewritten, and a
// warp holds whole rows, so there are no partials (lne.stats unused) and no finalize pass.  src/swin.rs:103-107,407.
struct MlpArgs {
  View x16;                      // [rows, C] 16-bit
  const float2* mr = nullptr;    // [rows] (-mean, rstd)
  const LayerW* fc1 = nullptr;   // gamma-folded, with column sums
  const LayerW* fc2 = nullptr;
  View xt;                       // [rows, C] fp32 residual stream, updated in place
  LnEmit lne;
  float2* mr_out = nullptr;
};

struct DeformArgs {
  View x;                 // [B,H,W,C] input
  View om;                // [B,H,W,3*taps] fp32: 2*taps offsets (dy,dx interleaved) then taps modulators
  int om_tiled = 0;       // 1: om.p is [m_tile][3*taps][128] over 16x8-pixel tiles (tc_gemm out_tiled, tile_w 16)
  void* scratch = nullptr; // k = 1 only: [B*H*W, 64] elements of x.dt; when set the conv runs as sample kernel + tc_gemm
  const LayerW* om_layer = nullptr;  // k = 1 with scratch: the sample kernel computes the offset conv itself (om unused)
  const LayerW* w = nullptr;
  const float* bias = nullptr;
  int act = ACT_NONE;
  int stride = 1, pad = -1;       // pad < 0: k/2; stride > 1 or pad != k/2: SIMT kernel only (DeformableConv2d, src/deform_conv.rs:29-99)
  View out;
};

struct AttnArgs {
  View qkv;               // [rows = nWin*144, 3C], window order, q pre-scaled
  const float* bias32 = nullptr;          // [heads][144][144] fp32
  const float* bias32p = nullptr;         // [heads][144][148] fp32: rows padded to 592 B (conflict-free row-per-thread
                                          // 16-byte shared-memory reads in tc_attn_kernel)
  int n_windows = 0;      // total windows (B * nW)
  int heads = 0;
  int nwh = 0, nww = 0;   // windows per image along h / w
  int shift = 0;          // 0: no mask at all (src/swin.rs:383)
  int ws = 12;            // window side; the tcgen05 kernel is built for 12 (144-token windows), 7 runs on the SIMT kernel
  int split_win = 0;      // > 0: windows [split_win, n_windows) belong to a second grid with nwh2 x nww2 windows per image
  int nwh2 = 0, nww2 = 0;
  // Token geometry (tcgen05 kernel).  h > 0: rows of a window that fall into the pad region of the [h, w] token grid
  // get q = k = v = qkv_bias (pad tokens are zeros AFTER norm1, src/swin.rs:355-366, so their qkv is the bias) written
  // into the staged tiles by the kernel itself -- the qkv matrix need not hold valid pad rows.
  int h = 0, w = 0, h2 = 0, w2 = 0;
  const void* qkv_bias16 = nullptr;   // [3C] in the operand type (q part pre-scaled); required when h > 0
  // token_out: out is the TOKEN-ordered [tokens, C] matrix (window_reverse + roll back + crop, src/swin.rs:387-401, done
  // by the store); second-grid tokens start at row tok2.  Otherwise out is window-ordered like qkv.
  int token_out = 0;
  long long tok2 = 0;
  View out;               // [rows, C]
};

// Per-kernel-class device timing (CUDA events on the launch stream around every launch of the class); used by
// bench.py for the roofline of the dominant kernel and the kernel-time shares.  Off on the timed throughput path.
enum KClass { KC_GEMM_TC = 0, KC_ATTN_TC, KC_DEFORM_TC, KC_GEMM_SIMT, KC_ATTN_SIMT, KC_LN, KC_GLUE, KC_COUNT };
struct KTimer {
  struct Rec { int cls; cudaEvent_t e0, e1; double flops; double bytes; std::string desc; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;   // events are created once and reused: no cudaEventCreate on the launch path
  size_t next = 0;
  cudaEvent_t get() {
    if (next == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
    return pool[next++];
  }
  void reset() { recs.clear(); next = 0; }
};

struct LaunchCtx {
  cudaStream_t stream = nullptr;
  int precision = 0;
  bool dry = false;        // plan pass: no launches
  long long* launches = nullptr;
  bool force_simt = false;
  KTimer* kt = nullptr;
  float* splitk = nullptr;     // scratch for split-K partial sums (tc_gemm decides per launch); null: never split
  size_t splitk_bytes = 0;
};

struct KScope {
  const LaunchCtx& c;
  cudaEvent_t e1 = nullptr;
  KScope(const LaunchCtx& ctx, int cls, double flops, double bytes = 0, const char* desc = nullptr) : c(ctx) {
    if (!c.kt || c.dry) return;
    KTimer::Rec r{cls, c.kt->get(), c.kt->get(), flops, bytes, desc ? desc : ""};
    cudaEventRecord(r.e0, c.stream);
    e1 = r.e1;
    c.kt->recs.push_back(r);
  }
  ~KScope() { if (e1) cudaEventRecord(e1, c.stream); }
};

// ---- SIMT (fp32 FMA) kernels: kernels_simt.cu ----
void simt_gemm(const LaunchCtx&, const GemmArgs&);
void simt_deform(const LaunchCtx&, const DeformArgs&);
void simt_attention(const LaunchCtx&, const AttnArgs&);

// Programmatic dependent launch (tc_ptx.cuh pdl_*): appends the stream-serialization attribute unless BRN_PDL=0
inline int pdl_attr(cudaLaunchAttribute* attr, int n) {
  static const bool off = [] { const char* v = getenv("BRN_PDL"); return v && v[0] == '0'; }();
  if (off) return n;
  attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[n].val.programmaticStreamSerializationAllowed = 1;
  return n + 1;
}

// ---- tcgen05 kernels ----
bool tc_gemm_supported(const GemmArgs&);
void tc_gemm(const LaunchCtx&, const GemmArgs&);
size_t tc_gemm_splitk_scratch_bytes(int batch);   // LaunchCtx::splitk size that every split-K launch of a batch fits in
void tc_attention(const LaunchCtx&, const AttnArgs&);
bool tc_mlp_supported(const MlpArgs&);
void tc_mlp(const LaunchCtx&, const MlpArgs&);
bool tc_deform_supported(const DeformArgs&);
void tc_deform(const LaunchCtx&, const DeformArgs&);

// ---- HBM-bound glue kernels: kernels_glue.cu ----
enum LnMode { LN_PLAIN = 0, LN_WINDOW = 1, LN_MERGE = 2 };
struct LnArgs {
  View x;                  // source rows [B,h,w,C] (fp32 stream) -- for LN_MERGE the pre-merge grid
  const float* gamma = nullptr;
  const float* beta = nullptr;
  View out;                // destination rows
  int mode = LN_PLAIN;
  int hp = 0, wp = 0, shift = 0;  // LN_WINDOW
  int ws = 12;                    // LN_WINDOW: window side
  // LN_WINDOW over two grids in one launch: output rows >= split gather from a second [B,h2,w2,C] grid whose tokens
  // start at row tok2 of x (the merged full + half resolution backbone pass)
  long long split = 0, tok2 = 0;
  int h2 = 0, w2 = 0, hp2 = 0, wp2 = 0;
};
void glue_layernorm(const LaunchCtx&, const LnArgs&);
void glue_patch_im2col(const LaunchCtx&, const float* x_nchw, int B, int H, int W, int patch, View out);
void glue_resize_nchw(const LaunchCtx&, const float* x, int B, int C, int H, int W, float* out, int Ho, int Wo);
void glue_resize_nhwc(const LaunchCtx&, View in, View out);                      // bilinear, align_corners=true
void glue_image2patches(const LaunchCtx&, const float* x_nchw, int B, int H, int W, int th, int tw, View out);
void glue_nchw_to_nhwc(const LaunchCtx&, const float* x, int B, int C, int H, int W, View out);
void glue_nhwc_to_nchw(const LaunchCtx&, View in, float* out);
int glue_gap_blocks(int HW);                                                  // partial-sum blocks per image
void glue_gap_sum(const LaunchCtx&, View x, float* part /*[B][glue_gap_blocks(H*W)][C] partial sums*/);
void glue_aspp_pool_bias(const LaunchCtx&, const float* part, int B, int HW, const LayerW* gap_conv,
                         const float* conv1_tail /*[64][256] fp32, bn1-scaled*/, const float* bn1_shift /*[64]*/,
                         float* out /*[B][64]*/);
void glue_gate(const LaunchCtx&, View p, View g16, const float* w16, float b0);
// 1x1 deformable conv, sampling half: out[px, 0:64] = m(px) * bilinear(x, px + (dy, dx)) (torchvision semantics)
void glue_deform_sample_k1(const LaunchCtx&, View x, View om, int om_tiled, const LayerW* om_layer /*fused 1x1 offset conv or null*/,
                           View out);
void glue_dot1(const LaunchCtx&, View p, const float* w, float* out /*[rows]*/);
void glue_final(const LaunchCtx&, const float* x_nchw, int B, int H, int W, const float* tab /*final_kernel.cu*/,
                const float* tab_host /*same table, host copy*/, const float* q, int qh, int qw, float* out,
                int apply_sigmoid);
void build_final_table(const float* w1 /*[64][27]*/, const float* b1 /*[64]*/, const double* wc /*[64][9]*/, double bc,
                       float* tab /*[336]*/);
void glue_copy_cast(const LaunchCtx&, View in, View out);
void glue_ln_stats_cast(const LaunchCtx&, View x, View x16, float2* stats /*[rows] (sum, sumsq)*/);
// (sum, sumsq) partials [parts][stride] -> (-mean, rstd) [rows], partials added in index order (deterministic); eps 1e-5
void glue_ln_finalize(const LaunchCtx&, const float2* stats, int parts, long long stride, long long rows, int C, float2* mr);
void glue_sigmoid(const LaunchCtx&, float* p, long long n);

// ---- pre / post-processing around the hot path: prepost.cu (examples/infer_image.rs:44-67, 85-105) ----
void prepost_preprocess(cudaStream_t st, const uint8_t* src_dev, int B, int h, int w, int H, int W, float* scratch_dev,
                        float* out_dev);
void prepost_postprocess(cudaStream_t st, const float* logits_dev, int already_prob, int B, int H, int W, int oh, int ow,
                         uint8_t* m8_dev, float* scratch_dev, uint8_t* out_dev);

}  // namespace brn
