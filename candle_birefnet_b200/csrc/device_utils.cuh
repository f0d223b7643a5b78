// Small device helpers shared by every kernel file.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "brn_common.h"

namespace brn {

// Division by a run-time constant as multiply-high + shift (valid for numerators < 2^31): the epilogue warps derive
// the tile coordinates and the window row map per tile, and a hardware-less integer division is ~20 instructions.
struct FastDiv {
  uint32_t d = 1, mul = 0, shr = 0;
  FastDiv() = default;
  explicit FastDiv(uint32_t div) : d(div) {
    if (div > 1) {
      uint32_t lg = 0;
      while ((1ull << lg) < div) ++lg;
      const uint32_t pw = 31 + lg;
      mul = (uint32_t)(((1ull << pw) + div - 1) / div);
      shr = pw - 32;
    }
  }
  __host__ __device__ __forceinline__ uint32_t div_any(uint32_t n) const { return n / d; }
  __device__ __forceinline__ uint32_t div(uint32_t n) const { return d != 1 ? __umulhi(n, mul) >> shr : n; }
};

// Window-ordered padded row m (batch-major, window id, token ti*12+tj; src/swin.rs:446-459) -> token row of the
// un-shifted, un-padded [B,h,w] grid, or -1 for a pad position.  Inverse of pad -> roll(-shift) -> partition
// (src/swin.rs:359-380) == window_reverse -> roll(+shift) -> crop (src/swin.rs:387-401).
__host__ __device__ __forceinline__ long long window_row_to_token(long long m, int h, int w, int hp, int wp,
                                                                  int shift, int ws = 12) {
  const int n = ws * ws;
  const int nww = wp / ws;
  const int nw = (hp / ws) * nww;
  const long long b = m / ((long long)nw * n);
  const int rem = (int)(m - b * (long long)nw * n);
  const int wid = rem / n, t = rem - wid * n;
  const int wi = wid / nww, wj = wid - wi * nww;
  const int ti = t / ws, tj = t - ti * ws;
  int r = wi * ws + ti + shift, c = wj * ws + tj + shift;
  if (r >= hp) r -= hp;
  if (c >= wp) c -= wp;
  if (r >= h || c >= w) return -1;
  return (b * h + r) * (long long)w + c;
}

__host__ __device__ __forceinline__ long long rowmap_token(const RowMap& rm, long long m) {
  if (rm.split > 0 && m >= rm.split) {
    const long long t = window_row_to_token(m - rm.split, rm.h2, rm.w2, rm.hp2, rm.wp2, rm.shift, rm.ws);
    return t < 0 ? t : t + rm.tok2;
  }
  return window_row_to_token(m, rm.h, rm.w, rm.hp, rm.wp, rm.shift, rm.ws);
}

// element load / store by runtime dtype tag (F32 / BF16 / F16)
__device__ __forceinline__ float ld_elem(const void* p, int dt, long long i) {
  if (dt == F32) return ((const float*)p)[i];
  if (dt == BF16) return __bfloat162float(((const __nv_bfloat16*)p)[i]);
  return __half2float(((const __half*)p)[i]);
}
__device__ __forceinline__ void st_elem(void* p, int dt, long long i, float v) {
  if (dt == F32) ((float*)p)[i] = v;
  else if (dt == BF16) ((__nv_bfloat16*)p)[i] = __float2bfloat16(v);
  else ((__half*)p)[i] = __float2half_rn(v);
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ float apply_act(float v, int act, int n, int act_from) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_GELU) return gelu_erf(v);
  if (act == ACT_2SIGMOID_TAIL) return n >= act_from ? 2.f / (1.f + expf(-v)) : v;
  return v;
}

// sm_100 packed fp32 pairs (FFMA2 / FADD2: two lanes of math per issue slot; the epilogue is issue-bound)
__device__ __forceinline__ unsigned long long pk2(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(unsigned long long v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long fmul2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace brn
