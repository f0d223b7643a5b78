// Shared tcgen05 epilogue: accumulator tile (TMEM) -> bias / activation / residual -> coalesced global stores.
//
// tcgen05.ld 32x32b gives every thread ONE ROW of the tile, so a direct store makes each warp instruction touch 32
// different rows (32 half-used sectors, ~1.5 LSU wavefronts per lane).  Instead every epilogue warp owns a 2 KB
// shared-memory staging block and works in 64-byte output granules (32 x 16-bit or 16 x fp32 columns of its 32 rows):
//   phase A (thread = row)   : TMEM -> registers, + bias (broadcast LDS), activation, convert, STS.128 into the
//                              XOR-swizzled staging block (conflict free);
//   phase B (warp = 8 rows x 64 B per instruction): LDS.128, fp32 residual add (coalesced LDG.128), STG.128 -- every
//                              store instruction writes 8 full 64-byte row segments.
// The activation is a template parameter (no per-element branches); erf-GELU is 12 FMA-pipe + 2 MUFU instructions.
#pragma once
#include "brn_common.h"
#include "device_utils.cuh"
#include "tc_ptx.cuh"

namespace brn {

struct EpiP {
  int N;
  const float* bias; int bias_bstride;
  int act, act_from;
  const void* res; int resdt; int ldres;
  void* out; int odt; int ldo;
  int vec;       // out (and res) base pointers and row pitches are 16-byte aligned
  // LayerNorm fold, consumer side (LnFold): (-mean, rstd) per A row of the fp32 stream the operand was rounded from
  const float2* lnf_mr;
  const float* lnf_colsum;   // [N] column sums of the rounded gamma-folded weights
  // producer side (LnEmit): statistics + raw 16-bit copy of the rows this epilogue stores
  float2* lne_stats; long long lne_stride; void* x16; int x16dt; int ldx16;
};

constexpr int EPI_STAGE_BYTES = 32 * 64;   // per epilogue warp

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exact-erf GELU (candle `gelu_erf`, src/swin.rs:105):  gelu(x) = relu(x) - |x| * q(|x|),  q(a) = 0.5 * erfc(a / sqrt2).
// log2 q is smooth and nearly quadratic, so a degree-5 polynomial (weighted minimax fit on [0, 6.2], coefficients from
// scripts/fit_gelu.py) followed by ONE ex2.approx gives q directly: max |error| 6.4e-7 over fp32 inputs in [-60, 60].
// The leading coefficient is negative, so the polynomial keeps falling beyond the fit interval (q -> 0) and the
// argument needs no clamp.  8 instructions per element, 1 MUFU -- the fc1 epilogue evaluates ~0.5 G GELUs per 1024^2
// image and is issue-bound on this math (was: degree 6 + clamp, 10 instructions; before that Abramowitz-Stegun with
// rcp + ex2, 14 instructions and 2 MUFU).
__device__ __forceinline__ float gelu_fast(float x) {
  const float a = fabsf(x);
  float l = fmaf(a, -0.0004732935631182045f, 0.007084455341100693f);
  l = fmaf(a, l, -0.05182714760303497f);
  l = fmaf(a, l, -0.4599926769733429f);
  l = fmaf(a, l, -1.1507877111434937f);
  l = fmaf(a, l, -1.000037670135498f);
  return fmaf(-a, ex2_approx(l), fmaxf(x, 0.f));
}

__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a0, a1)), "l"(pk2(b0, b1)));
  upk2(r, a0, a1);
}
// gelu_fast on a pair: the polynomial runs in -|x| (odd coefficients negated; the sign trick folds into the FFMA2
// operand modifier), 6 FFMA2 + 2 FMNMX + 2 MUFU per two elements
__device__ __forceinline__ void gelu_fast2(float& x0, float& x1) {
  const unsigned long long na = pk2(-fabsf(x0), -fabsf(x1));
  unsigned long long l = fma2(na, pk2(0.0004732935631182045f, 0.0004732935631182045f),
                              pk2(0.007084455341100693f, 0.007084455341100693f));
  l = fma2(na, l, pk2(0.05182714760303497f, 0.05182714760303497f));
  l = fma2(na, l, pk2(-0.4599926769733429f, -0.4599926769733429f));
  l = fma2(na, l, pk2(1.1507877111434937f, 1.1507877111434937f));
  l = fma2(na, l, pk2(-1.000037670135498f, -1.000037670135498f));
  float l0, l1;
  upk2(l, l0, l1);
  const unsigned long long r = fma2(na, pk2(ex2_approx(l0), ex2_approx(l1)), pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
  upk2(r, x0, x1);
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

__device__ __forceinline__ void tmem_ld16_raw(uint32_t taddr, uint32_t* v) {
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
__device__ __forceinline__ void tmem_wait_dep(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                 "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]),
                 "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
                 "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}

// 16-byte global load that asks L2 to fetch the whole 256-byte neighbourhood on a miss: the fp32 residual is read in
// 64-byte row segments (8 rows per instruction), and the next three granules of the same rows follow shortly after
__device__ __forceinline__ uint4 ldg_l2_256(const void* p) {
  uint4 v;
  asm volatile("ld.global.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// packed add of two 16-bit pairs (bf16x2 or f16x2 by run-time type)
__device__ __forceinline__ uint32_t add16x2(uint32_t a, uint32_t b, int dt) {
  uint32_t r;
  if (dt == BF16) asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  else asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

template <int ACT>
__device__ __forceinline__ float epi_act(float f, int col, int act_from) {
  if (ACT == ACT_RELU) return fmaxf(f, 0.f);
  if (ACT == ACT_GELU) return gelu_fast(f);
  if (ACT == ACT_2SIGMOID_TAIL) return col >= act_from ? 2.f * rcp_approx(1.f + ex2_approx(-1.4426950408889634f * f)) : f;
  return f;
}

// One epilogue warp's share of an accumulator tile: its 32 rows (TMEM lane quadrant encoded in taddr) x tile columns
// [c0, c1) (multiples of 16; tile column 0 is output column n0).
//   orow  : this lane's output row (row index into out / res), < 0 for rows that must not be written
//   sbias : SHARED address of the tile's bias (index = tile column, zero padded to c1 rounded up to 32; zeros if none)
//   stage : SHARED address of this warp's EPI_STAGE_BYTES staging block (16-byte aligned)
// O32: fp32 output (16 columns per 64-byte granule), else 16-bit output by p.odt (32 columns per granule).
// Residual (out may alias res: the Swin residual stream is updated in place, so residual loads can never be hoisted
// above earlier stores by the compiler -- they are issued explicitly, two granules ahead, right after the stores):
//   mode 1: fp32 residual + fp32 output, added in phase B (coalesced LDG.128)
//   mode 2: 16-bit residual of a 16-bit output, added packed in phase B (coalesced LDG.128; one extra 16-bit rounding)
//   mode 3: anything else, scalar loads in phase A
// LNF : folded LayerNorm -- value = rstd * (acc - mean * colsum[col]) + bias[col]; `nmu` = -mean and `rstd` of THIS
//       lane's row (phase A geometry), `scs` = SHARED address of the staged column sums (indexed like sbias).
// EMIT: (fp32 output only) the stored rows' raw 16-bit copy goes to p.x16 and their (sum, sumsq) over this warp's
//       columns to p.lne_stats[part_idx * stride + row] (phase-B geometry: the values after the residual add).
// TS  : (16-bit output, no residual, plain row-major token matrix) phase B is ONE TMA store per granule: lane 0 hands
//       the staged 32 rows x 64 B block (its XOR swizzle is exactly CU_TENSOR_MAP_SWIZZLE_64B) to cp.async.bulk.tensor;
//       rows / columns past the matrix are clipped by the tensor map, so the granule needs no predicates, no LDS and no
//       STG.  `tmO` = output tensor map (box 32 elements x 32 rows), `trow0` = output row of this warp's lane 0.
template <int ACT, bool O32, int RM, bool PP = (RM == 0), bool LNF = false, bool EMIT = false, bool TS = false>
__device__ __forceinline__ void epi_warp(const EpiP& p, uint32_t taddr, int n0, int c0, int c1, long long orow,
                                         uint32_t sbias, uint32_t stage, int lane, float nmu = 0.f, float rstd = 1.f,
                                         uint32_t scs = 0, int part_idx = 0, const CUtensorMap* tmO = nullptr,
                                         int trow0 = 0) {
  static_assert(!EMIT || O32, "LnEmit needs the fp32 output path");
  static_assert(!TS || (!O32 && RM == 0), "TMA store: 16-bit output without residual");
  if (c0 >= c1) {
    if (EMIT) {   // an empty column part still owns a statistics slot: zeros
      if (orow >= 0) p.lne_stats[(long long)part_idx * p.lne_stride + orow] = make_float2(0.f, 0.f);
    }
    return;
  }
  constexpr int GC = O32 ? 16 : 32;       // columns per 64-byte granule
  constexpr int PER = O32 ? 4 : 8;        // elements per 16-byte chunk
  constexpr int ESZ = O32 ? 4 : 2;
  constexpr int rmode = RM;
  // phase-B geometry: instruction `it` covers rows it*8 + lane/4, 16-byte chunk lane%4
  const int kb = lane & 3;
  long long orow_b[4];
#pragma unroll
  for (int it = 0; it < 4; ++it) orow_b[it] = __shfl_sync(0xffffffffu, orow, it * 8 + (lane >> 2));
  const uint32_t my_row = stage + lane * 64;
  const int sw_a = (lane >> 1) & 3;
  // phase-B row base pointers (byte address of column n0 of each of this lane's 4 rows), once per tile
  char* orow_p[4];
#pragma unroll
  for (int it = 0; it < 4; ++it) orow_p[it] = (char*)p.out + (orow_b[it] * p.ldo + n0) * ESZ;

  // residual prefetch: slot (granule & 1) holds the residual of that granule
  // residual prefetch ring: slot (granule % RSLOTS) holds the residual of that granule.  Four granules (4 x 16-byte
  // loads x 32 lanes = 8 KB per warp, 64 KB per SM) in flight: with two the residual GEMMs sat at 3.7-3.9 TB/s whatever
  // their shape (r02 run C) -- Little's law on ~1 us of loaded DRAM latency.
  constexpr int RSLOTS = RM == 1 ? 4 : 2;                 // (the 16-bit residual variant spills at four)
  uint4 rq0[4], rq1[4], rq2[RSLOTS == 4 ? 4 : 1], rq3[RSLOTS == 4 ? 4 : 1];
  auto res_issue = [&](int c, uint4 (&q)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = make_uint4(0u, 0u, 0u, 0u);
    if (c >= c1) return;
    if (rmode == 1) {
      const int col = n0 + c + kb * 4;
      const int nl = min(p.N - col, min(GC, c1 - c) - kb * 4);
#pragma unroll
      for (int it = 0; it < 4; ++it)
        if (orow_b[it] >= 0 && nl >= 4)
          q[it] = ldg_l2_256((const float*)p.res + orow_b[it] * p.ldres + col);
    } else if (rmode == 2) {
      // 16-bit residual of a 16-bit output: fetched in the phase-B geometry (8 rows x 64 bytes per instruction)
      const int col = n0 + c + kb * 8;
      const int nl = min(p.N - col, min(GC, c1 - c) - kb * 8);
#pragma unroll
      for (int it = 0; it < 4; ++it)
        if (orow_b[it] >= 0 && nl >= 8)
          q[it] = ldg_l2_256((const uint16_t*)p.res + orow_b[it] * p.ldres + col);
    }
  };

  // PP: two TMEM destination register sets in ping-pong -- the load of granule g+1 is in flight while granule g is
  // worked on.  !PP (register-starved callers: the 80-register deformable kernel, the residual variants): one set,
  // which costs one register move per element to free it for the next load.
  uint32_t va[GC], vb[PP ? GC : 1];
#pragma unroll
  for (int j = 0; j < GC; ++j) { va[j] = 0u; if (PP) vb[j] = 0u; }
  tmem_ld16_raw(taddr + c0, va);
  if (!O32 && c0 + 16 < c1) tmem_ld16_raw(taddr + c0 + 16, va + (GC - 16));
  if (rmode == 1 || rmode == 2) {
    res_issue(c0, rq0); res_issue(c0 + GC, rq1);
    if constexpr (RSLOTS == 4) { res_issue(c0 + 2 * GC, rq2); res_issue(c0 + 3 * GC, rq3); }
  }
  // every row of this warp is a real output row: the interior store path needs no per-row predicate
  const bool rows_ok = __all_sync(0xffffffffu, orow >= 0);
  float es[EMIT ? 4 : 1], eq[EMIT ? 4 : 1];
  if (EMIT) {
#pragma unroll
    for (int it = 0; it < 4; ++it) { es[it] = 0.f; eq[it] = 0.f; }
  }
  const unsigned long long nmu2 = pk2(nmu, nmu), rstd2 = pk2(rstd, rstd);

  auto granule = [&](int c, uint4 (&rq)[4], uint32_t (&v)[GC], uint32_t* vn) {
    const int ncol = min(GC, c1 - c);
    tmem_wait_dep(v);
    // the next granule's TMEM load overlaps the math and the stores of this one
    auto load_next = [&] {
      if (c + GC < c1) {
        tmem_ld16_raw(taddr + c + GC, vn);
        if (!O32 && c + GC + 16 < c1) tmem_ld16_raw(taddr + c + GC + 16, vn + (GC - 16));
      }
    };
    if (PP) load_next();
    // ---- phase A: this thread's row, columns [c, c + ncol) ----
    float f[GC];
#pragma unroll
    for (int j = 0; j < GC; ++j) f[j] = __uint_as_float(v[j]);
    if (!PP) load_next();
    // the bias is always staged (zeros when the layer has none)
    if (LNF) {
#pragma unroll
      for (int j = 0; j < GC; j += 4) {
        const uint4 bv = ptx::lds128(sbias + (c + j) * 4), cv = ptx::lds128(scs + (c + j) * 4);
        unsigned long long t0 = fma2(nmu2, pk2(__uint_as_float(cv.x), __uint_as_float(cv.y)), pk2(f[j], f[j + 1]));
        unsigned long long t1 = fma2(nmu2, pk2(__uint_as_float(cv.z), __uint_as_float(cv.w)), pk2(f[j + 2], f[j + 3]));
        t0 = fma2(rstd2, t0, pk2(__uint_as_float(bv.x), __uint_as_float(bv.y)));
        t1 = fma2(rstd2, t1, pk2(__uint_as_float(bv.z), __uint_as_float(bv.w)));
        upk2(t0, f[j], f[j + 1]);
        upk2(t1, f[j + 2], f[j + 3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < GC; j += 4) {
        const uint4 bv = ptx::lds128(sbias + (c + j) * 4);
        add2(f[j], f[j + 1], __uint_as_float(bv.x), __uint_as_float(bv.y));
        add2(f[j + 2], f[j + 3], __uint_as_float(bv.z), __uint_as_float(bv.w));
      }
    }
    if (ACT == ACT_GELU) {
#pragma unroll
      for (int j = 0; j < GC; j += 2) gelu_fast2(f[j], f[j + 1]);
    } else {
#pragma unroll
      for (int j = 0; j < GC; ++j) f[j] = epi_act<ACT>(f[j], n0 + c + j, p.act_from);
    }
    if (rmode == 3 && orow >= 0) {
#pragma unroll
      for (int j = 0; j < GC; ++j)
        if (j < ncol && n0 + c + j < p.N) f[j] += ld_elem(p.res, p.resdt, orow * p.ldres + n0 + c + j);
    }
    if (TS) {       // the previous granule's TMA store must have read the staging block before it is overwritten
      if (lane == 0) ptx::tma_store_wait_read();
      __syncwarp();
    }
    if (O32) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        ptx::sts128(my_row + ((k ^ sw_a) << 4), make_uint4(__float_as_uint(f[4 * k]), __float_as_uint(f[4 * k + 1]),
                                                           __float_as_uint(f[4 * k + 2]), __float_as_uint(f[4 * k + 3])));
    } else if (p.odt == BF16) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        ptx::sts128(my_row + ((k ^ sw_a) << 4),
            make_uint4(pack_bf16x2(f[(8 * k) % GC], f[(8 * k + 1) % GC]), pack_bf16x2(f[(8 * k + 2) % GC], f[(8 * k + 3) % GC]),
                       pack_bf16x2(f[(8 * k + 4) % GC], f[(8 * k + 5) % GC]), pack_bf16x2(f[(8 * k + 6) % GC], f[(8 * k + 7) % GC])));
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        ptx::sts128(my_row + ((k ^ sw_a) << 4),
            make_uint4(pack_f16x2(f[(8 * k) % GC], f[(8 * k + 1) % GC]), pack_f16x2(f[(8 * k + 2) % GC], f[(8 * k + 3) % GC]),
                       pack_f16x2(f[(8 * k + 4) % GC], f[(8 * k + 5) % GC]), pack_f16x2(f[(8 * k + 6) % GC], f[(8 * k + 7) % GC])));
    }
    if (TS) {
      ptx::fence_proxy_async_smem();        // generic-proxy STS -> visible to the TMA engine
      __syncwarp();
      if (lane == 0) { ptx::tma_store_2d(tmO, stage, n0 + c, trow0); ptx::tma_store_commit(); }
      return;
    }
    __syncwarp();
    // ---- phase B: 8 rows x 64 bytes per instruction ----
    const int col = n0 + c + kb * PER;                   // first output column of this lane's 16-byte chunk
    const int nleft = min(p.N - col, ncol - kb * PER);   // valid elements in the chunk (<= 0: none)
    uint4 val[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int r = it * 8 + (lane >> 2);
      val[it] = ptx::lds128(stage + r * 64 + ((kb ^ ((r >> 1) & 3)) << 4));
    }
    if (O32 && rmode == 1) {
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        float4* fv = reinterpret_cast<float4*>(&val[it]);
        const float4* rv = reinterpret_cast<const float4*>(&rq[it]);
        add2(fv->x, fv->y, rv->x, rv->y); add2(fv->z, fv->w, rv->z, rv->w);
      }
    }
    if (!O32 && rmode == 2) {
      // packed 16-bit add of the already rounded value (one extra rounding of <= 0.5 ulp of the 16-bit type)
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        val[it].x = add16x2(val[it].x, rq[it].x, p.odt); val[it].y = add16x2(val[it].y, rq[it].y, p.odt);
        val[it].z = add16x2(val[it].z, rq[it].z, p.odt); val[it].w = add16x2(val[it].w, rq[it].w, p.odt);
      }
    }
    if (EMIT) {
      // LnEmit (N % 16 == 0 and vec are launch preconditions, so every granule is whole): statistics of the stored fp32
      // values and their raw 16-bit copy, 8 bytes per lane (32 contiguous bytes = one sector per row)
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const float4 fv = *reinterpret_cast<const float4*>(&val[it]);
        es[it] += (fv.x + fv.y) + (fv.z + fv.w);
        eq[it] = fmaf(fv.x, fv.x, fmaf(fv.y, fv.y, fmaf(fv.z, fv.z, fmaf(fv.w, fv.w, eq[it]))));
        if (orow_b[it] >= 0) {
          const uint2 h = p.x16dt == BF16 ? make_uint2(pack_bf16x2(fv.x, fv.y), pack_bf16x2(fv.z, fv.w))
                                          : make_uint2(pack_f16x2(fv.x, fv.y), pack_f16x2(fv.z, fv.w));
          *reinterpret_cast<uint2*>((uint16_t*)p.x16 + orow_b[it] * p.ldx16 + n0 + c + kb * 4) = h;
        }
      }
    }
    if (rows_ok && p.vec && ncol == GC && n0 + c + GC <= p.N) {
      // interior granule: four unpredicated 16-byte stores
#pragma unroll
      for (int it = 0; it < 4; ++it) *reinterpret_cast<uint4*>(orow_p[it] + (c + kb * PER) * ESZ) = val[it];
    } else
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const long long ro = orow_b[it];
      if (ro < 0 || nleft <= 0) continue;
      char* op = orow_p[it] + (c + kb * PER) * ESZ;
      if (p.vec && nleft >= PER) {
        *reinterpret_cast<uint4*>(op) = val[it];
      } else if (O32) {
        // unaligned / tail columns: element stores; a mode-1 residual of a partial chunk was not prefetched
        const float* fv = reinterpret_cast<const float*>(&val[it]);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (e < nleft) ((float*)op)[e] = fv[e] + (rmode == 1 ? ((const float*)p.res)[ro * p.ldres + col + e] : 0.f);
      } else {
        const uint16_t* hv = reinterpret_cast<const uint16_t*>(&val[it]);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (e < nleft) {
            uint16_t h = hv[e];
            // a mode-2 residual of a partial chunk was not prefetched
            if (rmode == 2) h = (uint16_t)add16x2(h, ((const uint16_t*)p.res)[ro * p.ldres + col + e], p.odt);
            ((uint16_t*)op)[e] = h;
          }
      }
    }
    if (rmode == 1 || rmode == 2) res_issue(c + RSLOTS * GC, rq);   // this slot's next granule, after the stores (aliasing)
    __syncwarp();
  };

  if constexpr (RSLOTS == 4) {
    static_assert(!PP, "the residual variants use one TMEM register set");
    for (int c = c0; c < c1; c += 4 * GC) {
      granule(c, rq0, va, va);
      if (c + GC < c1) granule(c + GC, rq1, va, va);
      if (c + 2 * GC < c1) granule(c + 2 * GC, rq2, va, va);
      if (c + 3 * GC < c1) granule(c + 3 * GC, rq3, va, va);
    }
  } else {
    for (int c = c0; c < c1; c += 2 * GC) {
      if constexpr (PP) {
        granule(c, rq0, va, vb);
        if (c + GC < c1) granule(c + GC, rq1, vb, va);
      } else {
        granule(c, rq0, va, va);
        if (c + GC < c1) granule(c + GC, rq1, va, va);
      }
    }
  }
  if (TS) {
    if (lane == 0) ptx::tma_store_wait_read();          // the next tile's first granule reuses the block
    __syncwarp();
  }
  if (EMIT) {
    // the four lanes of a row (16-byte chunks kb = 0..3) hold its partial sums: fixed-order butterfly, lane kb = 0 writes
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      es[it] += __shfl_xor_sync(0xffffffffu, es[it], 1); eq[it] += __shfl_xor_sync(0xffffffffu, eq[it], 1);
      es[it] += __shfl_xor_sync(0xffffffffu, es[it], 2); eq[it] += __shfl_xor_sync(0xffffffffu, eq[it], 2);
      if (kb == 0 && orow_b[it] >= 0)
        p.lne_stats[(long long)part_idx * p.lne_stride + orow_b[it]] = make_float2(es[it], eq[it]);
    }
  }
}

// Tile-major fp32 output (GemmArgs::out_tiled): element (row r of M tile t, column n) -> out[(t * N + n) * 128 + r].
// With lane == row the plain thread-per-row store is already perfectly coalesced (32 consecutive floats per
// instruction); this is the layout the deformable gather reads its offsets / modulators from.
template <int ACT>
__device__ __forceinline__ void epi_warp_tiled(const EpiP& p, uint32_t taddr, int n0, int c0, int c1, long long tile,
                                               int row, uint32_t sbias) {
  float* const base = (float*)p.out + tile * p.N * 128 + row;
  for (int c = c0; c < c1; c += 16) {
    uint32_t v[16];
    tmem_ld16_raw(taddr + c, v);
    tmem_wait_dep(v);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int col = n0 + c + j;
      if (col < p.N) {
        float f = __uint_as_float(v[j]) + (sbias ? ptx::lds32(sbias + (c + j) * 4) : 0.f);
        base[(long long)col * 128] = epi_act<ACT>(f, col, p.act_from);
      }
    }
  }
}
__device__ __forceinline__ void epi_warp_tiled_dyn(const EpiP& p, uint32_t taddr, int n0, int c0, int c1, long long tile,
                                                   int row, uint32_t sbias) {
  switch (p.act) {
    case ACT_RELU: epi_warp_tiled<ACT_RELU>(p, taddr, n0, c0, c1, tile, row, sbias); break;
    case ACT_2SIGMOID_TAIL: epi_warp_tiled<ACT_2SIGMOID_TAIL>(p, taddr, n0, c0, c1, tile, row, sbias); break;
    default: epi_warp_tiled<ACT_NONE>(p, taddr, n0, c0, c1, tile, row, sbias); break;
  }
}

// runtime activation / output type / residual mode -> template instance
template <bool O32, int RM>
__device__ __forceinline__ void epi_warp_act(const EpiP& p, uint32_t taddr, int n0, int c0, int c1, long long orow,
                                             uint32_t sbias, uint32_t stage, int lane) {
  switch (p.act) {
    case ACT_RELU: epi_warp<ACT_RELU, O32, RM>(p, taddr, n0, c0, c1, orow, sbias, stage, lane); break;
    case ACT_GELU: epi_warp<ACT_GELU, O32, RM>(p, taddr, n0, c0, c1, orow, sbias, stage, lane); break;
    case ACT_2SIGMOID_TAIL: epi_warp<ACT_2SIGMOID_TAIL, O32, RM>(p, taddr, n0, c0, c1, orow, sbias, stage, lane); break;
    default: epi_warp<ACT_NONE, O32, RM>(p, taddr, n0, c0, c1, orow, sbias, stage, lane); break;
  }
}
__device__ __forceinline__ void epi_warp_dyn(const EpiP& p, uint32_t taddr, int n0, int c0, int c1, long long orow,
                                             uint32_t sbias, uint32_t stage, int lane) {
  const bool o32 = p.odt == F32;
  if (!p.res) {
    if (o32) epi_warp_act<true, 0>(p, taddr, n0, c0, c1, orow, sbias, stage, lane);
    else epi_warp_act<false, 0>(p, taddr, n0, c0, c1, orow, sbias, stage, lane);
  } else if (p.act == ACT_NONE && o32 && p.resdt == F32 && p.vec) {
    epi_warp<ACT_NONE, true, 1>(p, taddr, n0, c0, c1, orow, sbias, stage, lane);
  } else if (p.act == ACT_NONE && !o32 && p.resdt == p.odt && p.vec && p.N % 8 == 0) {
    epi_warp<ACT_NONE, false, 2>(p, taddr, n0, c0, c1, orow, sbias, stage, lane);
  } else {
    if (o32) epi_warp_act<true, 3>(p, taddr, n0, c0, c1, orow, sbias, stage, lane);
    else epi_warp_act<false, 3>(p, taddr, n0, c0, c1, orow, sbias, stage, lane);
  }
}

}  // namespace brn
