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

// Epilogue registers of one 16-column chunk: accumulators (raw TMEM bits), bias and residual, all fetched ahead of
// use so the TMEM load, the L1/L2 bias load and the global residual load of chunk c+1 overlap the math and the
// stores of chunk c (the epilogue warps have nothing else to hide latency with).
struct EpiRegs {
  uint32_t v[16];
  float bias[16];
  float res[16];
};

__device__ __forceinline__ void tmem_ld16_raw(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// tcgen05.wait::ld with the destination registers as in/out operands, so the compiler cannot schedule a use of the
// loaded values above the wait
__device__ __forceinline__ void tmem_wait_dep(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                 "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :
               : "memory");
}

// Issue the loads of chunk [nb, nb+16) of output row `orow`.  The TMEM load is warp-collective: every lane calls this.
__device__ __forceinline__ void epi_prefetch(const EpiP& p, EpiRegs& r, uint32_t taddr, int nb, long long orow,
                                             const float* bias, bool valid) {
  tmem_ld16_raw(taddr, r.v);
  const bool full16 = nb + 16 <= p.N;
  if (bias) {
    if (full16) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + nb + j));
        r.bias[j] = bv.x; r.bias[j + 1] = bv.y; r.bias[j + 2] = bv.z; r.bias[j + 3] = bv.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) r.bias[j] = nb + j < p.N ? __ldg(bias + nb + j) : 0.f;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) r.bias[j] = 0.f;
  }
  if (p.res && valid) {
    if (p.vec_res && full16) {
      if (p.resdt == F32) {
        const float4* rp = reinterpret_cast<const float4*>((const float*)p.res + orow * p.ldres + nb);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float4 rv = rp[j]; r.res[4 * j] = rv.x; r.res[4 * j + 1] = rv.y; r.res[4 * j + 2] = rv.z; r.res[4 * j + 3] = rv.w; }
      } else {
        const uint4* rp = reinterpret_cast<const uint4*>((const uint16_t*)p.res + orow * p.ldres + nb);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint4 rv = rp[j];
          const uint32_t* h = reinterpret_cast<const uint32_t*>(&rv);
#pragma unroll
          for (int t = 0; t < 4; ++t) { const float2 ff = unpack16x2(h[t], p.resdt); r.res[8 * j + 2 * t] = ff.x; r.res[8 * j + 2 * t + 1] = ff.y; }
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) r.res[j] = nb + j < p.N ? ld_elem(p.res, p.resdt, orow * p.ldres + nb + j) : 0.f;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) r.res[j] = 0.f;
  }
}

// bias / activation / residual / convert / store of a prefetched chunk (call after tcgen05.wait::ld)
__device__ __forceinline__ void epi_finish(const EpiP& p, const EpiRegs& r, int nb, long long orow) {
  const int esz = p.odt == F32 ? 4 : 2;
  const bool full16 = nb + 16 <= p.N;
  float f[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(r.v[j]) + r.bias[j];
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
#pragma unroll
  for (int j = 0; j < 16; ++j) f[j] += r.res[j];
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
      if (nb + j < p.N) st_elem(op, p.odt, j, f[j]);
  }
}

// One output row of an accumulator tile: chunks first, first+step, ... < nchunks (16 columns each), software
// pipelined over two register sets.  taddr = TMEM address of column 0 of the tile in this warp's lane quadrant.
__device__ __forceinline__ void epi_row(const EpiP& p, uint32_t taddr, int n0, int nchunks, int first, int step,
                                        long long orow, const float* bias, bool valid) {
  EpiRegs r0, r1;
  int c = first;
  if (c >= nchunks) return;
  epi_prefetch(p, r0, taddr + c * 16, n0 + c * 16, orow, bias, valid);
  while (true) {
    tmem_wait_dep(r0.v);
    const int c1 = c + step;
    if (c1 < nchunks) epi_prefetch(p, r1, taddr + c1 * 16, n0 + c1 * 16, orow, bias, valid);
    if (valid) epi_finish(p, r0, n0 + c * 16, orow);
    if (c1 >= nchunks) break;
    tmem_wait_dep(r1.v);
    const int c2 = c1 + step;
    if (c2 < nchunks) epi_prefetch(p, r0, taddr + c2 * 16, n0 + c2 * 16, orow, bias, valid);
    if (valid) epi_finish(p, r1, n0 + c1 * 16, orow);
    if (c2 >= nchunks) break;
    c = c2;
  }
}

}  // namespace brn
