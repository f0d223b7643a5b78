// The steps either side of the hot path (examples/infer_image.rs:44-67 and :85-105) as device
// kernels, so that an images/s pipeline never touches the pixels on the host:
//   preprocess : RGB8 [h,w,3] -> resize_exact(H, W, Triangle) -> (p/255 - mean) / std -> fp32 NCHW [3,H,W]
//   postprocess: logits [H,W] -> sigmoid -> (v * 255).clamp(0,255) as u8 -> resize(orig_w, orig_h, Lanczos3) -> u8
// The resampling follows the `image` crate 0.25.9 (`imageops::sample`, Cargo.lock:1102-1105; the crate itself is not
// in the reference tree): a vertical pass into an f32 image, then a horizontal pass with clamp + round-half-away to
// u8; per output sample the source window is [floor(c - s), ceil(c + s)) around c = (o + 0.5) * ratio with
// s = support * max(ratio, 1), weights kernel((i - (c - 0.5)) / max(ratio, 1)) normalised to sum 1, everything in f32
// and accumulated as t += p * w in tap order.  The weights are computed on the host with exactly those f32 operations
// and the kernels use __fmul_rn / __fadd_rn (no FMA contraction), so Triangle results are bit-identical to the
// restatement in oracle/imageops_ref.py; Lanczos3 weights go through sinf and may differ in the last ulp.
// HBM-bound: bytes in + bytes out per pass.
#include <cmath>
#include <vector>

#include "brn_common.h"

namespace brn {

struct ResampleTable {      // device copies
  int* left = nullptr;      // [n_out] first source index
  int* start = nullptr;     // [n_out + 1] offsets into w
  float* w = nullptr;       // normalised weights
  int n_out = 0;
};

static float tri_kernel(float x) { const float a = fabsf(x); return a < 1.0f ? 1.0f - a : 0.0f; }
static float sinc_f(float t) { const float a = t * (float)M_PI; return t == 0.0f ? 1.0f : sinf(a) / a; }
static float lanczos3_kernel(float x) { return fabsf(x) < 3.0f ? sinc_f(x) * sinc_f(x / 3.0f) : 0.0f; }

// image::imageops::sample: the window / weight loop shared by vertical_sample and horizontal_sample
static void build_weights(int n_in, int n_out, int filter, std::vector<int>& left, std::vector<int>& start,
                          std::vector<float>& w) {
  const float support = filter == 0 ? 1.0f : 3.0f;
  const float ratio = (float)n_in / (float)n_out;
  const float sratio = ratio < 1.0f ? 1.0f : ratio;
  const float src_support = support * sratio;
  left.resize(n_out); start.resize(n_out + 1); w.clear();
  for (int o = 0; o < n_out; ++o) {
    float c = ((float)o + 0.5f) * ratio;
    long long l = (long long)floorf(c - src_support);
    l = std::min<long long>(std::max<long long>(l, 0), n_in - 1);
    long long r = (long long)ceilf(c + src_support);
    r = std::min<long long>(std::max<long long>(r, l + 1), n_in);
    c = c - 0.5f;
    left[o] = (int)l; start[o] = (int)w.size();
    float sum = 0.0f;
    for (long long i = l; i < r; ++i) {
      const float x = ((float)i - c) / sratio;
      const float v = filter == 0 ? tri_kernel(x) : lanczos3_kernel(x);
      w.push_back(v);
      sum += v;
    }
    for (size_t k = start[o]; k < w.size(); ++k) w[k] /= sum;
  }
  start[n_out] = (int)w.size();
}

struct TableHolder {
  std::vector<void*> ptrs;
  ~TableHolder() { for (void* p : ptrs) cudaFree(p); }
  ResampleTable make(int n_in, int n_out, int filter, cudaStream_t st) {
    std::vector<int> left, start; std::vector<float> w;
    build_weights(n_in, n_out, filter, left, start, w);
    ResampleTable t; t.n_out = n_out;
    BRN_CUDA(cudaMalloc(&t.left, left.size() * 4)); ptrs.push_back(t.left);
    BRN_CUDA(cudaMalloc(&t.start, start.size() * 4)); ptrs.push_back(t.start);
    BRN_CUDA(cudaMalloc(&t.w, std::max<size_t>(w.size(), 1) * 4)); ptrs.push_back(t.w);
    BRN_CUDA(cudaMemcpyAsync(t.left, left.data(), left.size() * 4, cudaMemcpyHostToDevice, st));
    BRN_CUDA(cudaMemcpyAsync(t.start, start.data(), start.size() * 4, cudaMemcpyHostToDevice, st));
    BRN_CUDA(cudaMemcpyAsync(t.w, w.data(), w.size() * 4, cudaMemcpyHostToDevice, st));
    BRN_CUDA(cudaStreamSynchronize(st));      // the host vectors go out of scope
    return t;
  }
};

// vertical pass: src u8 [B,h,w,C] -> tmp f32 [B,H,w,C]; one thread per output element, consecutive threads along (x, c)
__global__ void __launch_bounds__(256) vsample_u8_kernel(const uint8_t* __restrict__ src, int B, int h, int wc /* w*C */,
                                                         ResampleTable t, float* __restrict__ tmp) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)B * t.n_out * wc;
  if (i >= total) return;
  const int xc = (int)(i % wc);
  const long long r = i / wc;
  const int oy = (int)(r % t.n_out), b = (int)(r / t.n_out);
  const int l = t.left[oy], k0 = t.start[oy], k1 = t.start[oy + 1];
  const uint8_t* s = src + ((long long)b * h + l) * wc + xc;
  float acc = 0.0f;
  for (int k = k0; k < k1; ++k, s += wc) acc = __fadd_rn(acc, __fmul_rn((float)*s, t.w[k]));
  tmp[i] = acc;
}

// horizontal pass: tmp f32 [B,H,w,C] -> clamp, round half away -> u8; MODE 0: u8 NHWC out [B,H,W,C];
// MODE 1 (C = 3): ImageNet-normalised fp32 NCHW out [B,3,H,W] (examples/infer_image.rs:54-67)
template <int MODE>
__global__ void __launch_bounds__(256) hsample_kernel(const float* __restrict__ tmp, int B, int H, int w, int C,
                                                      ResampleTable t, uint8_t* __restrict__ out8, float* __restrict__ outf) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  const int W = t.n_out;
  const long long total = (long long)B * H * W * C;
  if (i >= total) return;
  int c, ox; long long row;
  if (MODE == 1) { ox = (int)(i % W); const long long r = i / W; const int y = (int)(r % H); const long long bc = r / H;
                   c = (int)(bc % C); row = (bc / C) * H + y; }
  else { c = (int)(i % C); const long long r = i / C; ox = (int)(r % W); row = r / W; }
  const int l = t.left[ox], k0 = t.start[ox], k1 = t.start[ox + 1];
  const float* s = tmp + (row * w + l) * C + c;
  float acc = 0.0f;
  for (int k = k0; k < k1; ++k, s += C) acc = __fadd_rn(acc, __fmul_rn(*s, t.w[k]));
  acc = fminf(fmaxf(acc, 0.0f), 255.0f);
  const float q = roundf(acc);                                  // f32::round: half away from zero
  if (MODE == 0) { out8[i] = (uint8_t)q; return; }
  const float mean = c == 0 ? 0.485f : c == 1 ? 0.456f : 0.406f;
  const float sd = c == 0 ? 0.229f : c == 1 ? 0.224f : 0.225f;
  outf[i] = __fdiv_rn(__fsub_rn(__fdiv_rn(q, 255.0f), mean), sd);
}

// same-size request: the crate copies; normalise only
__global__ void __launch_bounds__(256) normalize_kernel(const uint8_t* __restrict__ src, int B, int H, int W,
                                                        float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)B * 3 * H * W;
  if (i >= total) return;
  const int x = (int)(i % W); const long long r = i / W; const int y = (int)(r % H); const long long bc = r / H;
  const int c = (int)(bc % 3); const long long b = bc / 3;
  const float q = (float)src[((b * H + y) * W + x) * 3 + c];
  const float mean = c == 0 ? 0.485f : c == 1 ? 0.456f : 0.406f;
  const float sd = c == 0 ? 0.229f : c == 1 ? 0.224f : 0.225f;
  out[i] = __fdiv_rn(__fsub_rn(__fdiv_rn(q, 255.0f), mean), sd);
}

// sigmoid -> (v * 255).clamp(0, 255) as u8 (truncation; examples/infer_image.rs:85-97)
__global__ void __launch_bounds__(256) mask_u8_kernel(const float* __restrict__ logits, long long n, int already_prob,
                                                      uint8_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float v = logits[i];
  if (!already_prob) v = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v)));
  v = fminf(fmaxf(__fmul_rn(v, 255.0f), 0.0f), 255.0f);
  out[i] = (uint8_t)v;
}

static unsigned blocks(long long n) { return (unsigned)((n + 255) / 256); }

// src: DEVICE u8 [B,h,w,3]; out: DEVICE fp32 [B,3,H,W].  scratch: DEVICE, >= B*H*w*3 floats.
void prepost_preprocess(cudaStream_t st, const uint8_t* src, int B, int h, int w, int H, int W, float* scratch, float* out) {
  if (h == H && w == W) {
    normalize_kernel<<<blocks((long long)B * 3 * H * W), 256, 0, st>>>(src, B, H, W, out);
    BRN_CUDA(cudaGetLastError());
    return;
  }
  TableHolder th;
  const ResampleTable tv = th.make(h, H, 0, st), thz = th.make(w, W, 0, st);
  vsample_u8_kernel<<<blocks((long long)B * H * w * 3), 256, 0, st>>>(src, B, h, w * 3, tv, scratch);
  hsample_kernel<1><<<blocks((long long)B * 3 * H * W), 256, 0, st>>>(scratch, B, H, w, 3, thz, nullptr, out);
  BRN_CUDA(cudaGetLastError());
  BRN_CUDA(cudaStreamSynchronize(st));     // the tables are freed on return
}

// logits: DEVICE fp32 [B,H,W]; out: DEVICE u8 [B,oh,ow].  m8: DEVICE u8 [B,H,W]; scratch: DEVICE >= B*oh*W floats.
void prepost_postprocess(cudaStream_t st, const float* logits, int already_prob, int B, int H, int W, int oh, int ow,
                         uint8_t* m8, float* scratch, uint8_t* out) {
  const long long n = (long long)B * H * W;
  mask_u8_kernel<<<blocks(n), 256, 0, st>>>(logits, n, already_prob, (oh == H && ow == W) ? out : m8);
  BRN_CUDA(cudaGetLastError());
  if (oh == H && ow == W) return;
  TableHolder th;
  const ResampleTable tv = th.make(H, oh, 1, st), thz = th.make(W, ow, 1, st);
  vsample_u8_kernel<<<blocks((long long)B * oh * W), 256, 0, st>>>(m8, B, H, W, tv, scratch);
  hsample_kernel<0><<<blocks((long long)B * oh * ow), 256, 0, st>>>(scratch, B, oh, W, 1, thz, out, nullptr);
  BRN_CUDA(cudaGetLastError());
  BRN_CUDA(cudaStreamSynchronize(st));
}

}  // namespace brn
