// Shared tcgen05 epilogue: 16 accumulator columns of one output row -> bias / activation / residual -> store.
#pragma once
#include "brn_common.h"
#include "device_utils.cuh"

namespace brn {

struct EpiP {
  int N;
  const float* bias; int bias_bstride;
  int act, act_from;
  const void* res; int resdt; int ldres; int vec_res;
  void* out; int odt; int ldo; int vec_out;
};

// erf via Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7): 2 MUFU + ~10 FMA instead of erff's ~40 instructions.  The
// epilogue evaluates ~0.5 G GELUs per 1024^2 image, so this is what keeps fc1 MMA-paced rather than epilogue-paced.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  poly *= t;
  const float erf_abs = 1.f - poly * __expf(-z * z);
  return 0.5f * x * (1.f + copysignf(erf_abs, x));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// two 16-bit elements (bf16 or fp16 by tag) <-> two floats
__device__ __forceinline__ uint32_t pack16x2(float a, float b, int dt) { return dt == BF16 ? pack_bf16x2(a, b) : pack_f16x2(a, b); }
__device__ __forceinline__ float2 unpack16x2(uint32_t u, int dt) {
  if (dt == BF16) return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
  return __half22float2(*reinterpret_cast<const __half2*>(&u));
}

// v: 16 fp32 accumulators (as raw bits) of output row `orow`, columns [nb, nb+16).  `bias` already points at the
// row's image (per-image bias) or at the shared bias vector.
__device__ __forceinline__ void epilogue_store16(const EpiP& p, const uint32_t (&v)[16], int nb, long long orow,
                                                 const float* bias) {
  const int esz = p.odt == F32 ? 4 : 2;
  float f[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
  const bool full16 = nb + 16 <= p.N;
  if (bias) {
    if (full16) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        float4 bv = __ldg(reinterpret_cast<const float4*>(bias + nb + j));
        f[j] += bv.x; f[j + 1] += bv.y; f[j + 2] += bv.z; f[j + 3] += bv.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) if (nb + j < p.N) f[j] += __ldg(bias + nb + j);
    }
  }
  if (p.act == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
  } else if (p.act == ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = gelu_fast(f[j]);
  } else if (p.act == ACT_2SIGMOID_TAIL) {
#pragma unroll
    for (int j = 0; j < 16; ++j) if (nb + j >= p.act_from) f[j] = 2.f / (1.f + __expf(-f[j]));
  }
  if (p.res) {
    if (p.vec_res && full16) {
      if (p.resdt == F32) {
        const float4* rp = reinterpret_cast<const float4*>((const float*)p.res + orow * p.ldres + nb);
#pragma unroll
        for (int j = 0; j < 4; ++j) { float4 rv = rp[j]; f[4 * j] += rv.x; f[4 * j + 1] += rv.y; f[4 * j + 2] += rv.z; f[4 * j + 3] += rv.w; }
      } else {
        const uint4* rp = reinterpret_cast<const uint4*>((const __nv_bfloat16*)p.res + orow * p.ldres + nb);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint4 rv = rp[j];
          const uint32_t* h = reinterpret_cast<const uint32_t*>(&rv);
#pragma unroll
          for (int t = 0; t < 4; ++t) { float2 ff = unpack16x2(h[t], p.resdt); f[8 * j + 2 * t] += ff.x; f[8 * j + 2 * t + 1] += ff.y; }
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (nb + j < p.N)
          f[j] += ld_elem(p.res, p.resdt, orow * p.ldres + nb + j);
    }
  }
  char* op = (char*)p.out + (orow * p.ldo + nb) * esz;
  if (p.vec_out && full16) {
    if (p.odt == F32) {
      float4* o4 = reinterpret_cast<float4*>(op);
#pragma unroll
      for (int j = 0; j < 4; ++j) o4[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
    } else {
      uint4* o4 = reinterpret_cast<uint4*>(op);
#pragma unroll
      for (int j = 0; j < 2; ++j)
        o4[j] = make_uint4(pack16x2(f[8 * j], f[8 * j + 1], p.odt), pack16x2(f[8 * j + 2], f[8 * j + 3], p.odt),
                           pack16x2(f[8 * j + 4], f[8 * j + 5], p.odt), pack16x2(f[8 * j + 6], f[8 * j + 7], p.odt));
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (nb + j < p.N) {
        st_elem(op, p.odt, j, f[j]);
      }
  }
}

}  // namespace brn
