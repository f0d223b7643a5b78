/*
 * birefnet_b200.h -- C ABI of libbirefnet_b200.so
 *
 * B200-native (sm_100a) replacement for the forward path of the Rust crate
 * imperatormk/candle-birefnet.  Every entry point names the reference
 * interface it replaces (paths relative to the reference repo).  No torch /
 * candle types cross this boundary: plain pointers, sizes and an opaque handle.
 *
 * Data layout at the boundary is the reference's: NCHW, fp32, contiguous
 * (`Tensor[B,3,H,W]` in, `Tensor[B,1,H,W]` out, src/birefnet.rs:412-461).
 * H and W must be multiples of 32 (the reference's image2patches uses integer
 * division, src/birefnet.rs:290-299, and fails in reshape otherwise; here it is
 * BRN_ERR_SHAPE).  Internal layouts (NHWC bf16, window-ordered tokens) are private.
 *
 * Errors: every function returns brn_status; brn_last_error() gives the
 * thread-local message (mirrors candle_core::Result / bail!, src/aspp.rs:101,117).
 * Nothing throws or aborts across the ABI.
 *
 * Threading: every entry point may be called from several host threads on one
 * handle (the reference's `&self` forward is immutable and callable
 * concurrently).  Host-side planning and launching is serialised by an
 * internal mutex; brn_forward_logits / brn_forward then wait for the device
 * WITHOUT the mutex, on one of two internal lanes (workspace + stream), so two
 * threads keep two calls in flight and the host<->device copies of one
 * overlap the kernels of the other.  Use one handle per GPU for image
 * sharding.
 */
#ifndef BIREFNET_B200_H
#define BIREFNET_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define BRN_API __attribute__((visibility("default")))
#else
#define BRN_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct brn_model brn_model;

typedef enum {
  BRN_OK = 0,
  BRN_ERR_INVALID = 1,        /* bad argument / null pointer                         */
  BRN_ERR_CUDA = 2,           /* CUDA runtime/driver error (message names the call)  */
  BRN_ERR_MISSING_TENSOR = 3, /* finalize: a key of the schema was never set         */
  BRN_ERR_UNKNOWN_TENSOR = 4, /* set_tensor: key not in the schema                   */
  BRN_ERR_SHAPE = 5,          /* shape / dtype mismatch, H or W not a multiple of 32 */
  BRN_ERR_STATE = 6,          /* call order (forward before finalize, ...)           */
  BRN_ERR_UNSUPPORTED = 7     /* config outside window 12 / head_dim 32              */
} brn_status;

typedef enum { BRN_F32 = 0, BRN_BF16 = 1, BRN_F16 = 2 } brn_dtype;

/* precision of the arithmetic (north_star): FP32 = SIMT fp32 FMA path (max |dlogit| <= 1e-3 vs the reference's
 * fp32 CPU forward); BF16 / FP16 = tcgen05 tensor-core path with bf16 / fp16 operands, fp32 accumulate, fp32
 * residual stream / LN / softmax (north-star tolerance max |dsigmoid| <= 1e-2 and IoU@0.5 >= 0.999; on random-init
 * Swin-L weights bf16 operands reach IoU ~0.9988, fp16 operands ~0.9997).  Switchable at
 * run time: finalize keeps fp32, bf16 and fp16 copies of the weights. */
typedef enum { BRN_PREC_FP32 = 0, BRN_PREC_BF16 = 1, BRN_PREC_FP16 = 2 } brn_precision;

/* What `DeformConvASPP::forward` computes (src/aspp.rs:168-187):
 *   CPU_FALLBACK: regular_conv(x), offsets/modulator discarded -- the reference's behaviour on Device::Cpu
 *                 (src/aspp.rs:183-185, src/deform_conv.rs:95-98);
 *   DEFORMABLE:   modulated deformable conv, torchvision DCNv2 semantics -- the reference's Metal path
 *                 (src/aspp.rs:58-165). */
typedef enum { BRN_DEFORM_CPU_FALLBACK = 0, BRN_DEFORM_DEFORMABLE = 1 } brn_deform_mode;

/* Mirror of SwinConfig (src/swin.rs:13-23) + the BiRefNetConfig fields that are read (src/birefnet.rs:13-30).
 * window_size must be 12 or 7 and embed_dim/num_heads[0] must be 32. */
typedef struct {
  int32_t embed_dim;      /* 192 for swin_l                      */
  int32_t depths[4];      /* {2,2,18,2}                          */
  int32_t num_heads[4];   /* {6,12,24,48}                        */
  int32_t window_size;    /* 12 (swin_b / swin_l) or 7 (swin_t/s) */
  int32_t mlp_ratio;      /* 4                                   */
  int32_t patch_size;     /* 4                                   */
  int32_t precision;      /* brn_precision (initial; changeable) */
  int32_t deform_mode;    /* brn_deform_mode (initial)           */
  int32_t micro_batch;    /* images per internal pass; 0 = auto  */
} brn_config;

/* BiRefNetConfig::swin_l() (src/birefnet.rs:64-66) + SwinConfig::swin_l() (src/swin.rs:69-80). */
BRN_API void brn_config_swin_l(brn_config* cfg);

/* SwinConfig::swin_b() (src/swin.rs:54-66): the other window-12 / head_dim-32 member of the family (embed 128,
 * heads 4/8/16/32).  The reference's BiRefNet::new always builds swin_l (src/birefnet.rs:390-391); the decoder
 * widths follow the backbone's channel counts, so the same path runs at this width. */
BRN_API void brn_config_swin_b(brn_config* cfg);

/* SwinConfig::swin_t() / swin_s() (src/swin.rs:27-52): embed 96, heads 3/6/12/24, depths 2/2/6/2 and 2/2/18/2, window 7
 * (49-token windows, shift 3).  Window-7 models run LayerNorm, the GEMMs, the decoder and the deformable convs on the
 * same kernels as swin_l; their attention runs on the SIMT kernel (the tcgen05 attention tile is 144 tokens). */
BRN_API void brn_config_swin_t(brn_config* cfg);
BRN_API void brn_config_swin_s(brn_config* cfg);

/* ---- model lifetime: replaces BiRefNet::new(config, vb) (src/birefnet.rs:389-409) ------------------------- */

/* Allocates a handle on CUDA device `device`.  Fails loudly (BRN_ERR_CUDA) when no sm_100 device is present:
 * there is no CPU fallback. */
BRN_API brn_status brn_model_create(const brn_config* cfg, int device, brn_model** out);

/* Replaces `vb.get(shape, key)` / candle_core::safetensors::load (examples/infer_image.rs:35-40).  `key` is the
 * HF safetensors name, `data` a HOST pointer to a contiguous tensor of `dtype`.
 * Unknown key -> BRN_ERR_UNKNOWN_TENSOR; wrong shape -> BRN_ERR_SHAPE (candle's `vb.get` shape check). */
BRN_API brn_status brn_model_set_tensor(brn_model* m, const char* key, const void* data, int dtype,
                                const int64_t* shape, int rank);

/* Replaces `candle_core::safetensors::load(path)` + `VarBuilder::from_tensors` (examples/infer_image.rs:35-40): reads a
 * .safetensors file (F32 / F16 / BF16 tensors) and sets every tensor the schema knows; like VarBuilder, tensors the
 * model never asks for (e.g. `num_batches_tracked`, cached index buffers of the HF checkpoint) are ignored, and a
 * tensor that is missing shows up at brn_model_finalize as BRN_ERR_MISSING_TENSOR.  *n_loaded = tensors set. */
BRN_API brn_status brn_model_load_safetensors(brn_model* m, const char* path, int32_t* n_loaded);

/* Number of tensors in the schema and the i-th key/shape (for loaders and tests). */
BRN_API int32_t brn_model_num_tensors(const brn_model* m);
BRN_API brn_status brn_model_tensor_info(const brn_model* m, int32_t index, const char** key, int64_t shape[4], int32_t* rank);

/* Folds eval-BatchNorm into the preceding convs, pre-gathers the relative-position bias to [heads,144,144]
 * (WindowAttention::new, src/swin.rs:143-152), re-lays conv weights out tap-major and uploads fp32 + bf16 copies.
 * Missing key -> BRN_ERR_MISSING_TENSOR (mirrors `vb.get` failing inside BiRefNet::new). */
BRN_API brn_status brn_model_finalize(brn_model* m);

BRN_API brn_status brn_model_set_precision(brn_model* m, int precision);
BRN_API brn_status brn_model_set_deform_mode(brn_model* m, int deform_mode);
/* CUDA-graph replay of the forward (default on): the second call with the same device buffers, shape and modes captures
 * the launch sequence (a few hundred kernels) of forward_logits into a CUDA graph; later calls replay it (batch-1 latency is
 * launch-bound otherwise).  Host-pointer calls use the handle's own staging buffers, so they replay as well.  Has no
 * counterpart in the reference (candle launches op by op). */
BRN_API brn_status brn_model_set_cuda_graph(brn_model* m, int on);

BRN_API void brn_model_destroy(brn_model* m);

/* ---- the hot path ------------------------------------------------------------------------------------------- */

/* BiRefNet::forward_logits (src/birefnet.rs:412-461).  x: [B,3,H,W] fp32 NCHW; out: [B,1,H,W] fp32.
 * x_is_device / out_is_device: 0 = host pointer (copied inside the call), 1 = device pointer on the handle's GPU.
 * `stream` is a cudaStream_t (NULL = the handle's own stream).  The call returns after the work is enqueued when
 * both pointers are device pointers, and after completion otherwise. */
BRN_API brn_status brn_forward_logits(brn_model* m, const float* x, int32_t B, int32_t H, int32_t W, int x_is_device,
                              float* out, int out_is_device, void* stream);

/* BiRefNet::forward / impl Module (src/birefnet.rs:466-476): sigmoid(forward_logits). */
BRN_API brn_status brn_forward(brn_model* m, const float* x, int32_t B, int32_t H, int32_t W, int x_is_device,
                       float* out, int out_is_device, void* stream);

/* SwinTransformer::forward (src/swin.rs:768-797): 4 NCHW fp32 maps [B,C_i,H/4>>i,W/4>>i] (BASELINE config 2). */
BRN_API brn_status brn_backbone_forward(brn_model* m, const float* x, int32_t B, int32_t H, int32_t W, int x_is_device,
                                float* const outs[4], int out_is_device, void* stream);

/* First half of BiRefNet::forward_logits (src/birefnet.rs:412-454): both backbone passes (full and half resolution),
 * the multi-scale concat and the cxt concat.  outs[0..2] = x1..x3 [B,2C_i,H/4>>i,W/4>>i], outs[3] = the
 * cxt-concatenated x4 [B,x4_channels,H/32,W/32] -- exactly the inputs brn_decoder_forward takes (the reference's
 * examples/bench_inference.rs:37-92 times these pieces separately).  NCHW fp32. */
BRN_API brn_status brn_features_forward(brn_model* m, const float* x, int32_t B, int32_t H, int32_t W, int x_is_device,
                                float* const outs[4], int out_is_device, void* stream);

/* SqueezeModule::forward + BiRefNetDecoder::forward (src/birefnet.rs:86-94, 278-376) on caller-provided
 * multi-scale features (BASELINE config 4).  x1..x3: [B,2C_i,...] NCHW fp32 as produced by forward_logits lines
 * 435-443; x4: the cxt-concatenated [B,x4_channels,H/32,W/32] (line 453), before the squeeze module. */
BRN_API brn_status brn_decoder_forward(brn_model* m, const float* x, const float* x1, const float* x2, const float* x3,
                               const float* x4, int32_t B, int32_t H, int32_t W, int is_device, float* out,
                               void* stream);

/* ---- image sharding inside one process ----------------------------------------------------
 * The reference is single-device; its forward is independent per image (eval-mode BatchNorm: src/decoder.rs:129,139,
 * src/aspp.rs:220,316,330), so a batch shards by images with no exchange step.  A brn_sharded owns one model handle
 * and one host thread per listed GPU (a device may be listed more than once), replicates the weights, splits a batch
 * into contiguous image ranges (the first B % n shards get one extra image) and runs the shards concurrently; the
 * caller's HOST buffers are read and written directly by every GPU (use brn_host_alloc for pinned memory).  Results are
 * bit-identical to a single-handle brn_forward_logits. */
typedef struct brn_sharded brn_sharded;
BRN_API brn_status brn_sharded_create(const brn_config* cfg, const int32_t* devices, int32_t n_devices, brn_sharded** out);
BRN_API brn_status brn_sharded_set_tensor(brn_sharded* s, const char* key, const void* data, int dtype, const int64_t* shape,
                                  int rank);
BRN_API brn_status brn_sharded_load_safetensors(brn_sharded* s, const char* path, int32_t* n_loaded);
BRN_API brn_status brn_sharded_finalize(brn_sharded* s);
BRN_API int32_t brn_sharded_num_devices(const brn_sharded* s);
/* x: HOST fp32 [B,3,H,W]; out: HOST fp32 [B,1,H,W] (forward_logits / forward = sigmoid, src/birefnet.rs:412-469). */
BRN_API brn_status brn_sharded_forward_logits(brn_sharded* s, const float* x, int32_t B, int32_t H, int32_t W, float* out);
BRN_API brn_status brn_sharded_forward(brn_sharded* s, const float* x, int32_t B, int32_t H, int32_t W, float* out);
BRN_API void brn_sharded_destroy(brn_sharded* s);
/* Pinned (page-locked, portable) host memory for input / output buffers; NULL on failure. */
BRN_API void* brn_host_alloc(size_t bytes);
BRN_API void brn_host_free(void* p);

/* ---- the steps either side of the path (examples/infer_image.rs) ---------------------------- */

/* examples/infer_image.rs:44-67: `img.resize_exact(W, H, FilterType::Triangle)` + ImageNet normalisation.
 * rgb: HOST u8 [B,h,w,3]; out: HOST fp32 NCHW [B,3,H,W].  The resampling restates the `image` crate 0.25.9
 * (`imageops::sample`: vertical pass in f32, horizontal pass, clamp, round half away from zero). */
BRN_API brn_status brn_preprocess_rgb8(int device, const uint8_t* rgb, int32_t B, int32_t h, int32_t w, int32_t H, int32_t W,
                               float* out);

/* examples/infer_image.rs:85-105: sigmoid -> `(v * 255.0).clamp(0, 255) as u8` -> `imageops::resize(orig_w, orig_h,
 * Lanczos3)`.  logits: HOST fp32 [B,H,W]; out: HOST u8 [B,orig_h,orig_w]. */
BRN_API brn_status brn_postprocess_mask(int device, const float* logits, int32_t B, int32_t H, int32_t W, int32_t orig_h,
                                int32_t orig_w, uint8_t* out);

/* examples/infer_image.rs:44-105 end to end on the device: RGB8 images in, u8 masks of the same size out; only the
 * 8-bit pixels cross PCIe (3 bytes in, 1 byte out per source pixel).  rgb: HOST u8 [B,h,w,3]; inference runs at
 * H x W (the CLI uses 1024 x 1024); masks: HOST u8 [B,h,w]. */
BRN_API brn_status brn_infer_rgb8(brn_model* m, const uint8_t* rgb, int32_t B, int32_t h, int32_t w, int32_t H, int32_t W,
                          uint8_t* masks);

/* ---- operator level (kernel parity tests; the reference's native boundaries) ---------------------------------- */

/* Plain-chain window attention: out = softmax(scale*q k^T + bias[h] (+ mask[w % nW])) v, the computation of
 * WindowAttention::forward_standard (src/swin.rs:266-311) == what flash_attention_with_[repeating_]bias replaces
 * (src/swin.rs:243,252; examples/test_flash_bias.rs:30-36).  qkv: HOST fp32 [n_windows,144,3*heads*32] (channel =
 * s*C + head*32 + d, src/swin.rs:218-223), bias: HOST fp32 [heads,144,144]; shift geometry (hp,wp in tokens, shift
 * 0 or window_size / 2) selects the analytic -100 mask of create_attention_mask (src/swin.rs:603-655).
 * window_size 12 (144-token windows: the tcgen05 kernel in the 16-bit precisions) or 7 (49 tokens, swin_t / swin_s:
 * SIMT kernel); with N = window_size^2 the shapes are qkv [n_windows,N,3*heads*32], bias [heads,N,N],
 * out: HOST fp32 [n_windows,N,heads*32]. */
BRN_API brn_status brn_window_attention(int device, int precision, const float* qkv, const float* bias, int32_t n_windows,
                                int32_t heads, int32_t window_size, int32_t hp, int32_t wp, int32_t shift, float* out);

/* Modulated deformable conv == call_deformable_im2col + weight matmul (src/aspp.rs:138-164,
 * src/deform_conv.rs:177-214), torchvision `deform_conv2d(x, offset, weight, bias, stride, padding, dilation=1, mask)`
 * semantics.  x: HOST fp32 NCHW [B,C,H,W]; with Ho = (H + 2 padding - k) / stride + 1: offset [B,2k^2,Ho,Wo] (dy,dx
 * interleaved per tap), mask [B,k^2,Ho,Wo], weight [O,C,k,k], bias [O] or NULL; dilation 1, one offset group.
 * out: HOST fp32 NCHW [B,O,Ho,Wo].  The tcgen05 kernel serves stride 1, padding k/2, C = 64 (the ASPP shapes); any other
 * geometry runs on the SIMT fp32 kernel. */
BRN_API brn_status brn_deform_conv2d(int device, int precision, const float* x, const float* offset, const float* mask,
                             const float* weight, const float* bias, int32_t B, int32_t C, int32_t H, int32_t W,
                             int32_t O, int32_t k, int32_t stride, int32_t padding, float* out);

/* DeformableConv2d::new(in, out, k, stride, padding, vb) + forward (src/deform_conv.rs:29-99), exported by the crate
 * root (src/lib.rs:13): the module owns offset_conv [2k^2,C,k,k], modulator_conv [k^2,C,k,k] and regular_conv
 * [O,C,k,k], all with bias and the same stride / padding.  forward: offset = offset_conv(x); modulator =
 * 2 sigmoid(modulator_conv(x)) (:83-86); deform_mode DEFORMABLE = the Metal path (:102-215), CPU_FALLBACK =
 * regular_conv(x), what candle computes on Device::Cpu (:95-98).  All HOST fp32; out [B,O,Ho,Wo]; regular_b may be
 * NULL. */
BRN_API brn_status brn_deformable_conv2d(int device, int precision, int deform_mode, const float* x, int32_t B, int32_t C,
                                 int32_t H, int32_t W, const float* offset_w, const float* offset_b,
                                 const float* modulator_w, const float* modulator_b, const float* regular_w,
                                 const float* regular_b, int32_t O, int32_t k, int32_t stride, int32_t padding,
                                 float* out);

/* D[M,N] = act(A[M,K] W[N,K]^T + bias[N]) (+ residual[M,N]); candle_nn::linear (src/swin.rs:98-107,130-131).
 * act: 0 none, 1 relu, 2 exact-erf gelu.  All HOST fp32 row-major. */
BRN_API brn_status brn_linear(int device, int precision, const float* a, const float* w, const float* bias,
                      const float* residual, int32_t M, int32_t N, int32_t K, int32_t act, float* out);

/* act(LayerNorm(x; gamma, beta, eps 1e-5) W^T + bias) with the LayerNorm FOLDED into the GEMM the way the Swin blocks
 * run it on the tensor-core path (norm1 -> qkv, norm2 -> fc1: src/swin.rs:355,217 and :407,104): the GEMM reads the raw
 * 16-bit copy of x with gamma folded into W, the epilogue applies rstd * (acc - mean * colsum) + (W beta + bias).
 * x: HOST fp32 [M,K]; out: HOST fp32 [M,N] (values rounded to the 16-bit operand type on the way).  precision must be
 * BRN_PREC_BF16 or BRN_PREC_FP16.  act: 0 none, 2 exact-erf gelu. */
BRN_API brn_status brn_ln_linear(int device, int precision, const float* x, const float* gamma, const float* beta,
                         const float* w, const float* bias, int32_t M, int32_t N, int32_t K, int32_t act, float* out);

/* The MLP half of a Swin block, out = x + fc2(gelu_erf(fc1(LayerNorm(x; gamma, beta, eps 1e-5)))): Mlp::forward
 * (src/swin.rs:103-107) as SwinTransformerBlock::forward calls it (src/swin.rs:407).  x, out: HOST fp32 [M, C];
 * w1 [hidden, C], b1 [hidden], w2 [C, hidden], b2 [C] (b1 / b2 may be NULL).  precision BRN_PREC_BF16 / BRN_PREC_FP16.
 * fused: 1 = the single-kernel path of the early stages (C in {128, 192}, hidden = 4C: hidden activations stay in
 * shared / tensor memory), 0 = two GEMMs through HBM, -1 = what the model itself would run for this shape.
 * out_mean_rstd (optional, HOST [M, 2]): per-row (mean, rstd) of `out` as emitted for the next block's folded norm1. */
BRN_API brn_status brn_swin_mlp(int device, int precision, const float* x, const float* gamma, const float* beta,
                        const float* w1, const float* b1, const float* w2, const float* b2, int32_t M, int32_t C,
                        int32_t hidden, int32_t fused, float* out, float* out_mean_rstd);

/* conv2d, stride 1, zero padding k/2, NCHW fp32 HOST tensors (candle_nn::conv2d; src/decoder.rs:44-45,104,113). */
BRN_API brn_status brn_conv2d(int device, int precision, const float* x, const float* weight, const float* bias, int32_t B,
                      int32_t C, int32_t H, int32_t W, int32_t O, int32_t k, int32_t act, float* out);

/* Kernel micro-benchmark on device-resident synthetic data: mean device ms per launch over `iters` back-to-back
 * launches.  kind 0: conv / linear implicit GEMM x[B,H,W,C] * w[N,k,k,C]; kind 1: window attention (B = windows,
 * C = heads, H,W = windows per image side, k = shift); kind 2: deformable conv (C must be 64);
 * kind 3: the MLP half of a Swin block on B*H*W rows of width C (with_res = 1: fused kernel, 0: fc1 + fc2). */
BRN_API brn_status brn_bench_op(int device, int precision, int kind, int32_t B, int32_t H, int32_t W, int32_t C, int32_t N,
                                int32_t k, int32_t act, int32_t with_res, int32_t out_f32, int32_t iters, float* ms_out);

/* ---- introspection --------------------------------------------------------------------------------------------- */

/* Kernels launched by this handle since creation / since the last reset (bench.py's `gpu_launches`). */
BRN_API int64_t brn_launch_count(const brn_model* m);
BRN_API void brn_launch_count_reset(brn_model* m);

/* Per-stage device time of the last forward, CUDA events on the launch stream (needs brn_profile_enable(m,1)).
 * names/ms arrays are owned by the handle; returns the number of entries. */
/* on = 1: per-stage events; on = 2: additionally CUDA events around EVERY kernel launch, summed per kernel class
 * (0 gemm_tcgen05, 1 attn_tcgen05, 2 deform_tcgen05, 3 gemm_simt, 4 attn_simt, 5 layernorm, 6 glue). */
BRN_API void brn_profile_enable(brn_model* m, int on);
/* Per-class device time (ms), algorithmic flops and bytes, and launch counts of the last forward (mode 2).
 * Arrays of capacity `cap`; returns the number of classes. */
BRN_API int32_t brn_kernel_class_times(const brn_model* m, float* ms, double* flops, double* bytes, int32_t* counts,
                                       int32_t cap);
BRN_API int32_t brn_profile_get(const brn_model* m, const char*** names, const float** ms, const double** flops);

BRN_API const char* brn_last_error(void);
BRN_API const char* brn_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BIREFNET_B200_H */
