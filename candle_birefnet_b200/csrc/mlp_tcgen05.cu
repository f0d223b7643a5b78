// Fused Swin MLP (sm_100a):  x <- x + fc2(gelu(fc1(LayerNorm(x))))   (src/swin.rs:103-107 Mlp::forward, :407 the block)
//
// for the early stages (C <= 192), where the two GEMMs are bound by their epilogues and by the 4C-wide hidden matrix
// they exchange through HBM (stage 0 of Swin-L at 1024^2, batch 16: 2.0 GB written by fc1 and read again by fc2 per
// block, for 0.77 TFLOP of work).  One persistent CTA owns a 128-row tile of the token matrix from the raw 16-bit copy
// of the residual stream to the updated fp32 stream; the hidden activations never leave the SM:
//
//   A tile [128 x C] (TMA, resident for the whole tile)
//   for each 128-wide chunk j of the hidden dimension:
//     G1: H (TMEM, 128 fp32 columns)       = A x W1'[chunk j]^T                     (tcgen05.mma, N = 128, K = C)
//     E1: H -> registers (H is free again) -> LayerNorm fold (rstd * (acc - mean * colsum) + bias') -> erf-GELU ->
//         16-bit -> Hs[j%2] (shared memory, the 128B-swizzled K-major layout TMA would have produced: an A operand)
//     G2: O[tile%2] (TMEM, C fp32 columns) += Hs[j%2] x W2[:, chunk j]^T            (tcgen05.mma, N = C, K = 128)
//   E2: O + bias2 + residual -> fp32 stream, its raw 16-bit copy, and (-mean, rstd) of the new rows for the next
//       block's folded norm1 (a warp holds whole rows, so the statistics need no partials and no finalize pass)
//
// Warp roles: warp 0 = TMA producer (A tile; W1 / W2 K blocks into per-block slots, each CTA of a 2-CTA cluster
// loading half of every weight block and multicasting it), warp 1 = MMA issuer, warps 2-9 = E1 (two per TMEM lane
// quadrant, 64 hidden columns each per chunk), warps 10-13 = E2 (one per lane quadrant).  The issuer runs G1 one chunk
// ahead of G2, so the tensor pipe works on chunk j+1 while the E1 warps run GELU on chunk j; O is double buffered, so E2
// of tile t overlaps E1 of tile t+1 on its own warps.  E2 is one latency chain per warp, so everything that can be
// asynchronous is: the fp32 residual of a 16-column granule (32 rows x 64 B) arrives by TMA in a three-deep ring, the
// updated rows and their 16-bit copy leave by TMA store from double-buffered blocks, and the thread = row layout of
// tcgen05.ld is kept end to end (the TMA engine does the transposition the staging round trip used to do; row
// statistics are per thread).  A/B r02: with register loads / STG.128 stores and one LDS round trip per granule the E2
// warps set the kernel's pace (E2 without its global traffic 960 -> 652 us; GELU math removed only 993 -> 960 us).
// Roofline: HBM (A + residual in + residual out + 16-bit copy = 12 C bytes per row) and epilogue issue (4C GELUs per
// row); tensor time is 16 C^2 flop per row.
#include <cuda.h>

#include <algorithm>
#include <cstdio>

#include "brn_common.h"
#include "device_utils.cuh"
#include "tc_ptx.cuh"
#include "tc_epilogue.cuh"

namespace brn {

CUtensorMap make_tmap_16(const void* base, int dt, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box, CUtensorMapSwizzle swz);
int device_sm_count();

constexpr int ML_BM = 128, ML_BK = 64, ML_HC = 128;
constexpr int ML_E1_WARPS = 8, ML_E2_WARPS = 4;
constexpr int ML_THREADS = 64 + 32 * (ML_E1_WARPS + ML_E2_WARPS);
constexpr int ML_AKB = ML_BM * ML_BK * 2;        // 16 KB: 128 rows x one 64-wide K block, 128B swizzle
constexpr int ML_HS_BYTES = (ML_HC / ML_BK) * ML_AKB;   // one hidden chunk as an A operand (two K blocks)
constexpr int ML_SMEM_MAX = 227 * 1024;
constexpr int ML_RES_DEPTH = 3;                  // residual granules in flight per E2 warp
// per E2 warp: residual ring (32 rows x 64 B each), two fp32 output blocks, two 16-bit output blocks (32 rows x 32 B)
constexpr int ML_E2_WARP_BYTES = ML_RES_DEPTH * 2048 + 2 * 2048 + 2 * 1024;
constexpr int ML_E2_BYTES = ML_E2_WARPS * ML_E2_WARP_BYTES;

template <int C>
struct MlpCfg {
  static_assert(C % 64 == 0 && C >= 128 && C <= 192, "fused MLP: C in {128, 192}");
  static constexpr int KB1 = C / ML_BK, HID = 4 * C, NCH = HID / ML_HC;
  // weight blocks in flight: one chunk's worth in DEDICATED slots -- KB1 slots of 16 KB for the K blocks of W1[chunk]
  // ([128 x 64]) and two of C x 128 B for those of W2[:, chunk] ([C x 64]).  Every slot is used once per chunk, so its
  // refill is requested a whole chunk period before the next use (a 4-deep uniform ring held less than one chunk: the
  // last W1 block of every chunk was requested only when the first W2 block of the previous one retired, ~1 us of
  // exposed TMA latency per chunk)
  static constexpr int W1_SLOT = ML_HC * ML_BK * 2, W2_SLOT = C * ML_BK * 2;
  static constexpr int NSLOT = KB1 + ML_HC / ML_BK;
  static constexpr int RING = KB1 * W1_SLOT + (ML_HC / ML_BK) * W2_SLOT;
  static constexpr int FIXED = KB1 * ML_AKB + ML_HS_BYTES + ML_E2_BYTES + 512;
  static constexpr int SMEM = 1024 + FIXED + RING;
  static constexpr int TMEM_O = ML_HC;            // accumulator columns: H, then O[0] and O[1] (C columns each)
  static_assert(SMEM <= ML_SMEM_MAX && TMEM_O + 2 * C <= 512 && NSLOT <= 6, "fused MLP: shared / tensor memory budget");
};

// fc1's per-column constants as a kernel parameter (constant bank): every thread needs every column's pair once per
// chunk, and as broadcast LDS.128 those reads were 57 % of the kernel's shared-memory wavefronts on a shared-memory pipe
// that the MMA operand reads and the TMA fills already keep busy (ncu r02: 77 % LSU data-pipe + 25 % tensor reads).
// cb[i] = (colsum[2i], colsum[2i+1], bias'[2i], bias'[2i+1]): column sums of the rounded gamma-folded W1; fc1 bias with
// W1 beta folded in.
template <int C>
struct MlpConst { float4 cb[2 * C]; float b2[C]; };     // + fc2 bias

struct MlpP {
  long long rows;
  int m_tiles;
  const float2* mr;       // [rows] (-mean, rstd) of the input rows
  const float* xt;        // fp32 residual stream (L2 prefetch only; E2 goes through the tensor maps)
  float2* mr_out;         // [rows] (-mean, rstd) of the updated rows (may alias mr: a tile's rows are read before E2 writes them)
};

template <bool BF>
__device__ __forceinline__ uint32_t mlp_pack(float a, float b) { return BF ? pack_bf16x2(a, b) : pack_f16x2(a, b); }

// Eight hidden columns of this thread's row: fold + bias + GELU -> 16-bit, one 16-byte chunk of the A operand -- in
// three steps, so that the caller can put the polynomial of the NEXT eight columns between the exponentials of a group
// and their first use (with two E1 warps per scheduler the MUFU latency was the largest single stall, ncu r02: a third
// of the E1 samples sat on the FFMA2 right after the four MUFU.EX2).
struct MlpG8 {
  unsigned long long na[4], l[4];     // -|x| pairs; log2 q(|x|) pairs (gelu_fast2, tc_epilogue.cuh)
  float relu[8];
};
__device__ __forceinline__ void mlp_g8_poly(MlpG8& g, const uint32_t* v, const float4* cb, unsigned long long nmu2,
                                            unsigned long long rstd2) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float4 ca = cb[2 * j], cc = cb[2 * j + 1];      // (colsum pair, bias pair) of columns 4j..4j+1, 4j+2..4j+3
    unsigned long long t0 = fma2(nmu2, pk2(ca.x, ca.y), pk2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])));
    unsigned long long t1 = fma2(nmu2, pk2(cc.x, cc.y), pk2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
    t0 = fma2(rstd2, t0, pk2(ca.z, ca.w));
    t1 = fma2(rstd2, t1, pk2(cc.z, cc.w));
    float x0, x1, x2, x3;
    upk2(t0, x0, x1);
    upk2(t1, x2, x3);
    g.na[2 * j] = pk2(-fabsf(x0), -fabsf(x1));
    g.na[2 * j + 1] = pk2(-fabsf(x2), -fabsf(x3));
    g.relu[4 * j] = fmaxf(x0, 0.f); g.relu[4 * j + 1] = fmaxf(x1, 0.f);
    g.relu[4 * j + 2] = fmaxf(x2, 0.f); g.relu[4 * j + 3] = fmaxf(x3, 0.f);
  }
#if defined(BRN_MLP_EXP) && BRN_MLP_EXP >= 1      // A/B experiment builds (scripts/history/r02_gpu_p.sh): no polynomial
#pragma unroll
  for (int i = 0; i < 4; ++i) g.l[i] = g.na[i];
  return;
#endif
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    unsigned long long l = fma2(g.na[i], pk2(0.0004732935631182045f, 0.0004732935631182045f),
                                pk2(0.007084455341100693f, 0.007084455341100693f));
    l = fma2(g.na[i], l, pk2(0.05182714760303497f, 0.05182714760303497f));
    l = fma2(g.na[i], l, pk2(-0.4599926769733429f, -0.4599926769733429f));
    l = fma2(g.na[i], l, pk2(1.1507877111434937f, 1.1507877111434937f));
    g.l[i] = fma2(g.na[i], l, pk2(-1.000037670135498f, -1.000037670135498f));
  }
}
__device__ __forceinline__ void mlp_g8_exp(const MlpG8& g, float (&e)[8]) {
#if defined(BRN_MLP_EXP) && BRN_MLP_EXP >= 1      // ... and no exponential
#pragma unroll
  for (int i = 0; i < 4; ++i) upk2(g.l[i], e[2 * i], e[2 * i + 1]);
  return;
#endif
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float l0, l1;
    upk2(g.l[i], l0, l1);
    e[2 * i] = ex2_approx(l0); e[2 * i + 1] = ex2_approx(l1);
  }
}
template <bool BF>
__device__ __forceinline__ uint4 mlp_g8_pack(const MlpG8& g, const float (&e)[8]) {
  uint32_t h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float r0, r1;
    upk2(fma2(g.na[i], pk2(e[2 * i], e[2 * i + 1]), pk2(g.relu[2 * i], g.relu[2 * i + 1])), r0, r1);
    h[i] = mlp_pack<BF>(r0, r1);
  }
  return make_uint4(h[0], h[1], h[2], h[3]);
}

template <int C, int CL, bool BF>
__global__ void __launch_bounds__(ML_THREADS, 1)
tc_mlp_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
              const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmX,
              const __grid_constant__ CUtensorMap tmX16, const __grid_constant__ MlpConst<C> cst, const MlpP p) {
  using K = MlpCfg<C>;
  constexpr int KB1 = K::KB1, NCH = K::NCH, NSLOT = K::NSLOT;
  ptx::pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sHs = sA + KB1 * ML_AKB;
  uint8_t* ring = sHs + ML_HS_BYTES;
  uint8_t* sE2 = ring + K::RING;
  uint64_t* full = (uint64_t*)(sE2 + ML_E2_BYTES);
  uint64_t* empty = full + NSLOT;
  uint64_t* afull = empty + NSLOT;
  uint64_t* aempty = afull + 1;
  uint64_t* hfull = aempty + 1;      // G1 of a chunk complete
  uint64_t* hfree = hfull + 1;       // every E1 thread holds its part of H in registers
  uint64_t* hsfull = hfree + 1;      // the 16-bit chunk is in shared memory
  uint64_t* hsempty = hsfull + 1;    // G2 has read it
  uint64_t* ofull = hsempty + 1;     // [2] all G2 of a tile complete
  uint64_t* ofree = ofull + 2;       // [2] E2 has read the accumulator
  uint64_t* rfull = ofree + 2;       // [E2 warps][ML_RES_DEPTH] residual granule landed
  uint32_t* tmem_slot = (uint32_t*)(rfull + ML_E2_WARPS * ML_RES_DEPTH);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = CL > 1 ? (int)ptx::cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSLOT; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], CL); }
    ptx::mbar_init(afull, 1); ptx::mbar_init(aempty, 1);
    ptx::mbar_init(hfull, 1); ptx::mbar_init(hfree, 32 * ML_E1_WARPS);
    ptx::mbar_init(hsfull, 32 * ML_E1_WARPS); ptx::mbar_init(hsempty, 1);
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&ofull[a], 1); ptx::mbar_init(&ofree[a], 32 * ML_E2_WARPS); }
    for (int a = 0; a < ML_E2_WARPS * ML_RES_DEPTH; ++a) ptx::mbar_init(&rfull[a], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA); ptx::prefetch_tmap(&tmW1); ptx::prefetch_tmap(&tmW2); ptx::prefetch_tmap(&tmX); ptx::prefetch_tmap(&tmX16);
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  if (CL > 1) ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_wait();

  // work: cluster `cid` owns tile pairs cid, cid + ncl, ...; this CTA's tile of a pair is pair * CL + rank.  Both CTAs of
  // a cluster run the same number of iterations (the weight ring is shared); a tile past the end loads zeros and stores nothing.
  const int pairs = (p.m_tiles + CL - 1) / CL;
  const int cid = blockIdx.x / CL, ncl = gridDim.x / CL;
  const int n_iter = cid < pairs ? (pairs - cid + ncl - 1) / ncl : 0;
  const int G = n_iter * NCH;
  // global row of tile-row `row` of this CTA's tile `it` (-1: past the end)
  auto grow_of = [&](int it, int row) -> long long {
    if (it >= n_iter) return -1;
    const long long r = (long long)((cid + it * ncl) * CL + rank) * ML_BM + row;
    return r < p.rows ? r : -1;
  };

  if (warp == 0) {
    if (ptx::elect_one()) {
      // ===== TMA producer: per global chunk g the K blocks of W1[chunk g], then those of W2[:, chunk g-1] =====
      constexpr int w1_rows = ML_HC / CL, w2_rows = C / CL;
      for (int g = 0; g <= G; ++g) {
        if (g < G) {
          const int it = g / NCH, j = g - it * NCH;
          if (j == 0) {
            const int m_tile = (cid + it * ncl) * CL + rank;
            ptx::mbar_wait(aempty, (it & 1) ^ 1);
            ptx::mbar_expect_tx(afull, KB1 * ML_AKB);
#pragma unroll
            for (int kb = 0; kb < KB1; ++kb) ptx::tma_load_2d(sA + kb * ML_AKB, &tmA, afull, kb * ML_BK, m_tile * ML_BM);
            // the tile's fp32 residual rows (contiguous) -> L2 now: E2 reads them a tile period later with two granules
            // in flight per warp, which covers an L2 hit but not an HBM miss (A/B r02: E2 without its global traffic
            // 960 -> 652 us, GELU math removed 993 -> 960 us)
            const long long r0 = (long long)m_tile * ML_BM, nr = min((long long)ML_BM, p.rows - r0);
            if (p.xt && nr > 0) ptx::bulk_prefetch_l2(p.xt + r0 * C, (uint32_t)(nr * C * 4));
          }
#pragma unroll
          for (int kb = 0; kb < KB1; ++kb) {
            ptx::mbar_wait(&empty[kb], (g & 1) ^ 1);
            ptx::mbar_expect_tx(&full[kb], K::W1_SLOT);
            uint8_t* dst = ring + kb * K::W1_SLOT + rank * w1_rows * (ML_BK * 2);
            if (CL > 1) ptx::tma_load_2d_mc(dst, &tmW1, &full[kb], kb * ML_BK, j * ML_HC + rank * w1_rows, (uint16_t)((1u << CL) - 1));
            else ptx::tma_load_2d(dst, &tmW1, &full[kb], kb * ML_BK, j * ML_HC);
          }
        }
        if (g >= 1) {
          const int jp = (g - 1) % NCH;
#pragma unroll
          for (int kb = 0; kb < ML_HC / ML_BK; ++kb) {
            const int slot = KB1 + kb;
            ptx::mbar_wait(&empty[slot], ((g - 1) & 1) ^ 1);
            ptx::mbar_expect_tx(&full[slot], K::W2_SLOT);
            uint8_t* dst = ring + KB1 * K::W1_SLOT + kb * K::W2_SLOT + rank * w2_rows * (ML_BK * 2);
            if (CL > 1) ptx::tma_load_2d_mc(dst, &tmW2, &full[slot], jp * ML_HC + kb * ML_BK, rank * w2_rows, (uint16_t)((1u << CL) - 1));
            else ptx::tma_load_2d(dst, &tmW2, &full[slot], jp * ML_HC + kb * ML_BK, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      // ===== MMA issuer =====
      const uint32_t idesc1 = ptx::make_idesc_16(ML_BM, ML_HC, 0, 0, BF ? 1 : 0);
      const uint32_t idesc2 = ptx::make_idesc_16(ML_BM, C, 0, 0, BF ? 1 : 0);
      auto release = [&](uint64_t* bar) {
        if (CL > 1) ptx::umma_commit_mc(bar, (uint16_t)((1u << CL) - 1)); else ptx::umma_commit(bar);
      };
      for (int g = 0; g <= G; ++g) {
        if (g < G) {
          // G1(g): H = A x W1[chunk]^T, as soon as the E1 warps hold chunk g-1 in registers
          const int it = g / NCH, j = g - it * NCH;
          if (j == 0) ptx::mbar_wait(afull, it & 1);
          ptx::mbar_wait(hfree, (g & 1) ^ 1);
          ptx::tc_fence_after();
#pragma unroll
          for (int kb = 0; kb < KB1; ++kb) {
            ptx::mbar_wait(&full[kb], g & 1);
            ptx::tc_fence_after();
            const uint64_t a_desc = ptx::make_smem_desc(ptx::smem_u32(sA + kb * ML_AKB), 16, 1024, ptx::SW_128B);
            const uint64_t b_desc = ptx::make_smem_desc(ptx::smem_u32(ring + kb * K::W1_SLOT), 16, 1024, ptx::SW_128B);
#pragma unroll
            for (int k = 0; k < ML_BK / 16; ++k) ptx::umma_f16_ss(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc1, (kb | k) != 0);
            release(&empty[kb]);
          }
          ptx::umma_commit(hfull);
          if (j == NCH - 1) ptx::umma_commit(aempty);       // the A tile may be overwritten by the next one
        }
        if (g >= 1) {
          // G2(g - 1): O[tile & 1] += Hs x W2[:, chunk]^T
          const int gp = g - 1, itp = gp / NCH, jp = gp - itp * NCH, ob = itp & 1;
          if (jp == 0) ptx::mbar_wait(&ofree[ob], ((itp >> 1) & 1) ^ 1);   // E2 of tile itp - 2 has read O[ob]
          ptx::mbar_wait(hsfull, gp & 1);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + K::TMEM_O + ob * C;
#pragma unroll
          for (int kb = 0; kb < ML_HC / ML_BK; ++kb) {
            const int slot = KB1 + kb;
            ptx::mbar_wait(&full[slot], gp & 1);
            ptx::tc_fence_after();
            const uint64_t a_desc = ptx::make_smem_desc(ptx::smem_u32(sHs + kb * ML_AKB), 16, 1024, ptx::SW_128B);
            const uint64_t b_desc = ptx::make_smem_desc(ptx::smem_u32(ring + KB1 * K::W1_SLOT + kb * K::W2_SLOT), 16, 1024, ptx::SW_128B);
#pragma unroll
            for (int k = 0; k < ML_BK / 16; ++k) ptx::umma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc2, (jp | kb | k) != 0);
            release(&empty[slot]);
          }
          ptx::umma_commit(hsempty);
          if (jp == NCH - 1) ptx::umma_commit(&ofull[ob]);
        }
      }
    }
  } else if (warp < 2 + ML_E1_WARPS) {
    // ===== E1 warps: warp % 4 = TMEM lane quadrant, (warp - 2) / 4 = half of the chunk's 128 hidden columns =====
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int sw = row & 7;
    // this thread's row inside K block `part` of a hidden chunk (128-byte rows, 8-row swizzle atoms of 1 KB)
    const uint32_t hs_row = ptx::smem_u32(sHs) + part * ML_AKB + (row >> 3) * 1024 + (row & 7) * 128;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + part * 64;
    float2 mr_next = make_float2(0.f, 1.f);
    { const long long r0 = grow_of(0, row); if (r0 >= 0) mr_next = __ldg(p.mr + r0); }
    for (int it = 0; it < n_iter; ++it) {
      const unsigned long long nmu2 = pk2(mr_next.x, mr_next.x), rstd2 = pk2(mr_next.y, mr_next.y);
      { const long long rn = grow_of(it + 1, row); if (rn >= 0) mr_next = __ldg(p.mr + rn); }
      for (int j = 0; j < NCH; ++j) {
        const int g = it * NCH + j;
        ptx::mbar_wait(hfull, g & 1);
        ptx::tc_fence_after();
        uint32_t va[32], vb[32];
        ptx::tmem_ld32(taddr, va);
        ptx::tmem_ld32(taddr + 32, vb);
        tmem_wait_dep(va);
        tmem_wait_dep(vb);
        ptx::tc_fence_before();
        ptx::mbar_arrive(hfree);                            // G1 of chunk g + 1 may overwrite H
        const float4* cb = cst.cb + (j * ML_HC + part * 64) / 2;     // warp-uniform index: constant-cache loads
        // eight groups of eight columns, software-pipelined: exp(k) issued, poly(k+1) computed, then pack(k).  The
        // packed chunk stays in registers until G2 of the PREVIOUS chunk has read the (single) hidden buffer -- that
        // MMA runs while this math does.
        uint4 out[8];
        MlpG8 ga, gb;
        mlp_g8_poly(ga, va, cb, nmu2, rstd2);
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
          float e[8];
          mlp_g8_exp(ga, e);
          mlp_g8_poly(gb, (c + 1 < 4 ? va : vb) + 8 * ((c + 1) & 3), cb + (c + 1) * 4, nmu2, rstd2);
          out[c] = mlp_g8_pack<BF>(ga, e);
          mlp_g8_exp(gb, e);
          if (c + 2 < 8) mlp_g8_poly(ga, (c + 2 < 4 ? va : vb) + 8 * ((c + 2) & 3), cb + (c + 2) * 4, nmu2, rstd2);
          out[c + 1] = mlp_g8_pack<BF>(gb, e);
        }
        ptx::mbar_wait(hsempty, (g & 1) ^ 1);               // G2 of chunk g - 1 has read Hs
#pragma unroll
        for (int c = 0; c < 8; ++c) ptx::sts128(hs_row + ((c ^ sw) << 4), out[c]);
        ptx::fence_proxy_async_smem();                      // generic-proxy stores -> visible to the MMA's operand reads
        ptx::mbar_arrive(hsfull);
      }
    }
  } else {
    // ===== E2 warps: one per TMEM lane quadrant, all C columns of its 32 rows in 16-column granules, thread = row =====
    const int q = warp & 3, ew = warp - 2 - ML_E1_WARPS;
    const int row = q * 32 + lane;
    const uint32_t base = ptx::smem_u32(sE2) + ew * ML_E2_WARP_BYTES;
    constexpr int NG = C / 16, D = ML_RES_DEPTH;
    // residual slot d: base + d * 2 KB; fp32 output block b: base + (D + b) * 2 KB; 16-bit block b: base + (D + 2) * 2 KB + b * 1 KB.
    // fp32 blocks are [32 rows][64 B] with CU_TENSOR_MAP_SWIZZLE_64B (16-byte chunk k of row r at (k ^ ((r >> 1) & 3)) * 16:
    // conflict-free for thread = row), 16-bit blocks [32 rows][32 B] unswizzled.
    const uint32_t sw_a = (lane >> 1) & 3;
    const uint32_t res_row = base + lane * 64, out_row = base + D * 2048 + lane * 64, x16_row = base + (D + 2) * 2048 + lane * 32;
    uint64_t* rf = rfull + ew * D;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + K::TMEM_O;
    auto tile_row0 = [&](int it) { return ((cid + it * ncl) * CL + rank) * ML_BM + q * 32; };   // past-the-end tiles: TMA zero-fills / clips
    const int NT = n_iter * NG;            // granules of this warp
    // residual of granule n -> slot n % D (lane 0)
    auto res_issue = [&](int n) {
#if defined(BRN_MLP_EXP) && BRN_MLP_EXP >= 2      // A/B experiment build: E2 without its HBM traffic
      return;
#endif
      if (n >= NT) return;
      const int it = n / NG, gran = n - it * NG, d = n % D;
      ptx::mbar_expect_tx(&rf[d], 2048);
      ptx::tma_load_2d((void*)(sE2 + ew * ML_E2_WARP_BYTES + d * 2048), &tmX, &rf[d], gran * 16, tile_row0(it));
    };
    if (lane == 0) {
#pragma unroll
      for (int d = 0; d < D; ++d) res_issue(d);
    }
    int n = 0;
    for (int it = 0; it < n_iter; ++it) {
      float es = 0.f, eq = 0.f;
      const uint32_t o_addr = lane_base + (it & 1) * C;
      const int row0 = tile_row0(it);
      ptx::mbar_wait(&ofull[it & 1], (it >> 1) & 1);
      ptx::tc_fence_after();
      uint32_t vo[16];
      tmem_ld16_raw(o_addr, vo);
#pragma unroll
      for (int gran = 0; gran < NG; ++gran, ++n) {
        const int c = gran * 16, d = n % D, b = n & 1;
        // the TMA stores issued two granules ago have read output blocks b
        if (lane == 0) ptx::tma_store_wait_read_n<1>();
        __syncwarp();
#if !(defined(BRN_MLP_EXP) && BRN_MLP_EXP >= 2)
        ptx::mbar_wait(&rf[d], (n / D) & 1);
#endif
        tmem_wait_dep(vo);
        float f[16];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint4 r = ptx::lds128(res_row + d * 2048 + ((k ^ sw_a) << 4));
          f[4 * k] = __uint_as_float(vo[4 * k]) + cst.b2[c + 4 * k] + __uint_as_float(r.x);
          f[4 * k + 1] = __uint_as_float(vo[4 * k + 1]) + cst.b2[c + 4 * k + 1] + __uint_as_float(r.y);
          f[4 * k + 2] = __uint_as_float(vo[4 * k + 2]) + cst.b2[c + 4 * k + 2] + __uint_as_float(r.z);
          f[4 * k + 3] = __uint_as_float(vo[4 * k + 3]) + cst.b2[c + 4 * k + 3] + __uint_as_float(r.w);
        }
        if (gran + 1 < NG) {
          tmem_ld16_raw(o_addr + c + 16, vo);
        } else {
          ptx::tc_fence_before();
          ptx::mbar_arrive(&ofree[it & 1]);      // the accumulator is in registers: G2 of tile it + 2 may start
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) { es += f[k]; eq = fmaf(f[k], f[k], eq); }
#pragma unroll
        for (int k = 0; k < 4; ++k)
          ptx::sts128(out_row + b * 2048 + ((k ^ sw_a) << 4), make_uint4(__float_as_uint(f[4 * k]), __float_as_uint(f[4 * k + 1]),
                                                                           __float_as_uint(f[4 * k + 2]), __float_as_uint(f[4 * k + 3])));
#pragma unroll
        for (int k = 0; k < 2; ++k)
          ptx::sts128(x16_row + b * 1024 + k * 16, make_uint4(mlp_pack<BF>(f[8 * k], f[8 * k + 1]), mlp_pack<BF>(f[8 * k + 2], f[8 * k + 3]),
                                                               mlp_pack<BF>(f[8 * k + 4], f[8 * k + 5]), mlp_pack<BF>(f[8 * k + 6], f[8 * k + 7])));
        ptx::fence_proxy_async_smem();       // staged blocks -> visible to the TMA engine; residual slot d: reads done
        __syncwarp();
        if (lane == 0) {
#if !(defined(BRN_MLP_EXP) && BRN_MLP_EXP >= 2)
          ptx::tma_store_2d(&tmX, base + (D + b) * 2048, c, row0);
          ptx::tma_store_2d(&tmX16, base + (D + 2) * 2048 + b * 1024, c, row0);
          ptx::tma_store_commit();
#endif
          res_issue(n + D);                  // slot d again, D granules ahead (rolls over into the next tile)
        }
      }
      // (-mean, rstd) of the updated row exactly as ln_finalize_kernel forms them
      const long long gr = grow_of(it, row);
      if (gr >= 0) {
        const float mu = es * (1.0f / (float)C);
        const float var = fmaxf(fmaf(-mu, mu, eq * (1.0f / (float)C)), 0.f);
        p.mr_out[gr] = make_float2(-mu, rsqrtf(var + 1e-5f));
      }
    }
    if (lane == 0) ptx::tma_store_wait_all();      // this warp's bulk stores are performed before the CTA retires
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CL > 1) ptx::cluster_sync_all();     // no CTA leaves while a peer may still multicast into it / arrive on its barriers
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
bool tc_mlp_supported(const MlpArgs& a) {
  if (!a.fc1 || !a.fc2 || !a.mr || !a.mr_out || !a.lne.x16) return false;
  const int C = a.x16.C;
  if (C != 128 && C != 192) return false;
  if (a.x16.dt != BF16 && a.x16.dt != F16) return false;
  const LayerW &w1 = *a.fc1, &w2 = *a.fc2;
  if (w1.taps() != 1 || w2.taps() != 1 || w1.Cin != C || w1.cin_pad != C || w1.N != 4 * C || w2.Cin != 4 * C ||
      w2.cin_pad != 4 * C || w2.N != C)
    return false;
  if (!w1.w16(a.x16.dt) || !w2.w16(a.x16.dt) || !w1.h_fold || w1.h_fold->size() != (size_t)3 * w1.N) return false;
  if (a.x16.rows() >= (1ll << 31) - 256) return false;
  if (a.x16.B != 1 || a.x16.H != 1 || a.x16.ld % 8 != 0 || ((uintptr_t)a.x16.p & 15)) return false;
  if (w2.bias && (!w2.h_bias || w2.h_bias->size() != (size_t)C)) return false;
  if (a.xt.dt != F32 || a.xt.rows() != a.x16.rows() || a.xt.C != C || a.xt.ld % 4 != 0 || ((uintptr_t)a.xt.p & 15)) return false;
  if (a.lne.x16dt != a.x16.dt || a.lne.ldx16 % 8 != 0 || ((uintptr_t)a.lne.x16 & 15)) return false;
  return true;
}

template <int C>
static void launch_mlp(const LaunchCtx& ctx, const MlpArgs& a, MlpP& p) {
  using K = MlpCfg<C>;
  const int dt = a.x16.dt;
  const int sms = device_sm_count();
  static const bool no_cluster = [] { const char* v = getenv("BRN_GEMM_CLUSTER"); return v && v[0] == '1'; }();
  const int CL = (!no_cluster && p.m_tiles >= sms) ? 2 : 1;
  uint64_t adims[2] = {(uint64_t)C, (uint64_t)p.rows};
  uint64_t astr[1] = {(uint64_t)a.x16.ld * 2};
  uint32_t abox[2] = {(uint32_t)ML_BK, (uint32_t)ML_BM};
  CUtensorMap tmA = make_tmap_16(a.x16.p, dt, 2, adims, astr, abox, CU_TENSOR_MAP_SWIZZLE_128B);
  uint64_t w1dims[2] = {(uint64_t)C, (uint64_t)K::HID};
  uint64_t w1str[1] = {(uint64_t)C * 2};
  uint32_t w1box[2] = {(uint32_t)ML_BK, (uint32_t)(ML_HC / CL)};
  CUtensorMap tmW1 = make_tmap_16(a.fc1->w16(dt), dt, 2, w1dims, w1str, w1box, CU_TENSOR_MAP_SWIZZLE_128B);
  uint64_t w2dims[2] = {(uint64_t)K::HID, (uint64_t)C};
  uint64_t w2str[1] = {(uint64_t)K::HID * 2};
  uint32_t w2box[2] = {(uint32_t)ML_BK, (uint32_t)(C / CL)};
  CUtensorMap tmW2 = make_tmap_16(a.fc2->w16(dt), dt, 2, w2dims, w2str, w2box, CU_TENSOR_MAP_SWIZZLE_128B);
  MlpConst<C> cst;
  {
    const std::vector<float>& hf = *a.fc1->h_fold;
    const float* cs = hf.data() + (dt == F16 ? K::HID : 0);
    const float* b1 = hf.data() + 2 * K::HID;
    for (int i = 0; i < 2 * C; ++i) cst.cb[i] = make_float4(cs[2 * i], cs[2 * i + 1], b1[2 * i], b1[2 * i + 1]);
    for (int i = 0; i < C; ++i) cst.b2[i] = a.fc2->bias ? (*a.fc2->h_bias)[i] : 0.f;
  }
  // E2: the fp32 stream (load + store) and the 16-bit copy (store) as 32-row x 16-column boxes
  uint64_t xdims[2] = {(uint64_t)C, (uint64_t)p.rows};
  uint64_t xstr[1] = {(uint64_t)a.xt.ld * 4};
  uint32_t xbox[2] = {16, 32};
  CUtensorMap tmX = make_tmap_16(a.xt.p, F32, 2, xdims, xstr, xbox, CU_TENSOR_MAP_SWIZZLE_64B);
  uint64_t x16str[1] = {(uint64_t)a.lne.ldx16 * 2};
  CUtensorMap tmX16 = make_tmap_16(a.lne.x16, dt, 2, xdims, x16str, xbox, CU_TENSOR_MAP_SWIZZLE_NONE);
  auto launch = [&](auto kern) {
    BRN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
    const int pairs = (p.m_tiles + CL - 1) / CL;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CL * std::min(pairs, sms / CL));
    cfg.blockDim = dim3(ML_THREADS);
    cfg.dynamicSmemBytes = K::SMEM;
    cfg.stream = ctx.stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_attr(attr, 1);
    BRN_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmW1, tmW2, tmX, tmX16, cst, p));
  };
  if (dt == BF16) { if (CL == 2) launch(tc_mlp_kernel<C, 2, true>); else launch(tc_mlp_kernel<C, 1, true>); }
  else { if (CL == 2) launch(tc_mlp_kernel<C, 2, false>); else launch(tc_mlp_kernel<C, 1, false>); }
  BRN_CUDA(cudaGetLastError());
}

void tc_mlp(const LaunchCtx& ctx, const MlpArgs& a) {
  if (ctx.launches) ++*ctx.launches;
  BRN_CHECK(tc_mlp_supported(a), 5, "tc_mlp: unsupported shapes / operands");
  if (ctx.dry) return;
  const int C = a.x16.C;
  MlpP p{};
  p.rows = a.x16.rows();
  p.m_tiles = (int)((p.rows + ML_BM - 1) / ML_BM);
  p.mr = a.mr;
  p.xt = a.xt.ld == C ? (const float*)a.xt.p : nullptr;
  p.mr_out = a.mr_out;
  const double rows = (double)p.rows;
  char desc[96] = "";
  if (ctx.kt) snprintf(desc, sizeof desc, "mlp M=%lld C=%d hid=%d tiles=%d", (long long)p.rows, C, 4 * C, p.m_tiles);
  KScope ks(ctx, KC_GEMM_TC, 2.0 * rows * C * 4 * C * 2, rows * C * (2 + 4 + 4 + 2) + 2.0 * 4 * C * C * 2, desc);
  switch (C) {
    case 128: launch_mlp<128>(ctx, a, p); break;
    default: launch_mlp<192>(ctx, a, p); break;
  }
}

}  // namespace brn
