// Fused Swin MLP (sm_100a):  x <- x + fc2(gelu(fc1(LayerNorm(x))))   (src/swin.rs:103-107 Mlp::forward, :407 the block)
//
// for the early stages (C <= 256), where the two GEMMs are bound by their epilogues and by the 4C-wide hidden matrix
// they exchange through HBM (stage 0 of Swin-L at 1024^2, batch 16: 2.0 GB written by fc1 and read again by fc2 per
// block, for 0.77 TFLOP of work).  One persistent CTA owns a 128-row tile of the token matrix from the raw 16-bit copy
// of the residual stream to the updated fp32 stream; the hidden activations never leave the SM:
//
//   A tile [128 x C] (TMA, resident for the whole tile)
//   for each 128-wide chunk j of the hidden dimension:
//     G1: H[j%2] (TMEM, 128 fp32 columns)  = A x W1'[chunk j]^T                     (tcgen05.mma, N = 128, K = C)
//     E1: H -> registers -> LayerNorm fold (rstd * (acc - mean * colsum) + bias') -> erf-GELU -> 16-bit ->
//         Hs[j%2] (shared memory, the 128B-swizzled K-major layout TMA would have produced: an A operand)
//     G2: O (TMEM, C fp32 columns)        += Hs[j%2] x W2[:, chunk j]^T             (tcgen05.mma, N = C, K = 128)
//   E2: O + bias2 + residual -> fp32 stream, raw 16-bit copy and LayerNorm partials for the next block
//       (the RES32 + LnEmit epilogue of tc_epilogue.cuh, unchanged)
//
// Warp roles: warp 0 = TMA producer (A tile; W1 / W2 K blocks through one mbarrier ring, each CTA of a 2-CTA cluster
// loading half of every weight block and multicasting it), warp 1 = MMA issuer, warps 2-9 = epilogue (two per TMEM lane
// quadrant, 64 hidden columns each per chunk).  The issuer runs G1 one chunk ahead of G2, so the tensor pipe works on
// chunk j+1 while the epilogue warps run GELU on chunk j.
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
constexpr int ML_EPI_WARPS = 8;
constexpr int ML_THREADS = 64 + 32 * ML_EPI_WARPS;
constexpr int ML_AKB = ML_BM * ML_BK * 2;        // 16 KB: 128 rows x one 64-wide K block, 128B swizzle
constexpr int ML_HS_BYTES = (ML_HC / ML_BK) * ML_AKB;   // one hidden chunk as an A operand (two K blocks)
constexpr int ML_SMEM_MAX = 227 * 1024;

template <int C>
struct MlpCfg {
  static_assert(C % 64 == 0 && C >= 128 && C <= 256, "fused MLP: C in {128, 192, 256}");
  static constexpr int KB1 = C / ML_BK, HID = 4 * C, NCH = HID / ML_HC;
  static constexpr int STAGE = C * ML_BK * 2;     // ring slot: a W2 K block [C x 64]; a W1 block [128 x 64] uses its first 16 KB
  static constexpr int FIXED = KB1 * ML_AKB + 2 * ML_HS_BYTES + (2 * HID + C) * 4 + 256;
  static constexpr int NST_FIT = (ML_SMEM_MAX - 1024 - FIXED) / STAGE;
  static constexpr int NST = NST_FIT > 6 ? 6 : NST_FIT;
  static constexpr int SMEM = 1024 + FIXED + NST * STAGE;
  static constexpr int TMEM_O = 2 * ML_HC;        // accumulator columns: H[0], H[1], then O (C columns)
  static_assert(NST >= 2 && TMEM_O + C <= 512 && NCH % 2 == 0, "fused MLP: shared / tensor memory budget");
};

struct MlpP {
  long long rows;
  int m_tiles;
  int in_bf16;
  const float* bias1;     // [4C] fc1 bias with W1 beta folded in
  const float* colsum1;   // [4C] column sums of the rounded gamma-folded W1
  const float* bias2;     // [C]
  const float2* mr;       // [rows] (-mean, rstd)
  EpiP epi;               // E2: fp32 residual in place + LnEmit
};

// eight hidden columns of this thread's row: fold + bias + GELU -> 16-bit, one 16-byte chunk of the A operand
__device__ __forceinline__ uint4 mlp_gelu8(const uint32_t* v, uint32_t sb, uint32_t scs, unsigned long long nmu2,
                                           unsigned long long rstd2, int bf16) {
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; j += 4) {
    const uint4 bv = ptx::lds128(sb + j * 4), cv = ptx::lds128(scs + j * 4);
    unsigned long long t0 = fma2(nmu2, pk2(__uint_as_float(cv.x), __uint_as_float(cv.y)),
                                 pk2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
    unsigned long long t1 = fma2(nmu2, pk2(__uint_as_float(cv.z), __uint_as_float(cv.w)),
                                 pk2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
    t0 = fma2(rstd2, t0, pk2(__uint_as_float(bv.x), __uint_as_float(bv.y)));
    t1 = fma2(rstd2, t1, pk2(__uint_as_float(bv.z), __uint_as_float(bv.w)));
    upk2(t0, f[j], f[j + 1]);
    upk2(t1, f[j + 2], f[j + 3]);
  }
#pragma unroll
  for (int j = 0; j < 8; j += 2) gelu_fast2(f[j], f[j + 1]);
  if (bf16)
    return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  return make_uint4(pack_f16x2(f[0], f[1]), pack_f16x2(f[2], f[3]), pack_f16x2(f[4], f[5]), pack_f16x2(f[6], f[7]));
}

template <int C, int CL>
__global__ void __launch_bounds__(ML_THREADS, 1)
tc_mlp_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
              const __grid_constant__ CUtensorMap tmW2, const MlpP p) {
  using K = MlpCfg<C>;
  constexpr int KB1 = K::KB1, HID = K::HID, NCH = K::NCH, NST = K::NST, STAGE = K::STAGE;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sHs = sA + KB1 * ML_AKB;
  uint8_t* ring = sHs + 2 * ML_HS_BYTES;
  float* sB1 = (float*)(ring + NST * STAGE);
  float* sCs1 = sB1 + HID;
  float* sB2 = sCs1 + HID;
  uint64_t* full = (uint64_t*)(sB2 + C);
  uint64_t* empty = full + NST;
  uint64_t* afull = empty + NST;
  uint64_t* aempty = afull + 1;
  uint64_t* hfull = aempty + 1;      // [2] G1 of a chunk complete
  uint64_t* hfree = hfull + 2;       // [2] every epilogue thread holds its part of H in registers
  uint64_t* hsfull = hfree + 2;      // [2] the 16-bit chunk is in shared memory
  uint64_t* hsempty = hsfull + 2;    // [2] G2 has read it
  uint64_t* ofull = hsempty + 2;
  uint64_t* ofree = ofull + 1;
  uint32_t* tmem_slot = (uint32_t*)(ofree + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = CL > 1 ? (int)ptx::cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], CL); }
    ptx::mbar_init(afull, 1); ptx::mbar_init(aempty, 1);
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&hfull[a], 1); ptx::mbar_init(&hfree[a], 32 * ML_EPI_WARPS);
      ptx::mbar_init(&hsfull[a], 32 * ML_EPI_WARPS); ptx::mbar_init(&hsempty[a], 1);
    }
    ptx::mbar_init(ofull, 1); ptx::mbar_init(ofree, 32 * ML_EPI_WARPS);
    ptx::fence_barrier_init();
  }
  for (int t = threadIdx.x; t < HID; t += ML_THREADS) { sB1[t] = __ldg(p.bias1 + t); sCs1[t] = __ldg(p.colsum1 + t); }
  for (int t = threadIdx.x; t < C; t += ML_THREADS) sB2[t] = p.bias2 ? __ldg(p.bias2 + t) : 0.f;
  if (warp == 0 && lane == 0) { ptx::prefetch_tmap(&tmA); ptx::prefetch_tmap(&tmW1); ptx::prefetch_tmap(&tmW2); }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  if (CL > 1) ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work: cluster `cid` owns tile pairs cid, cid + ncl, ...; this CTA's tile of a pair is pair * CL + rank.  Both CTAs of
  // a cluster run the same number of iterations (the weight ring is shared); a tile past the end loads zeros and stores nothing.
  const int pairs = (p.m_tiles + CL - 1) / CL;
  const int cid = blockIdx.x / CL, ncl = gridDim.x / CL;
  const int n_iter = cid < pairs ? (pairs - cid + ncl - 1) / ncl : 0;
  const int G = n_iter * NCH;

  if (warp == 0) {
    if (ptx::elect_one()) {
      // ===== TMA producer: per global chunk g the K blocks of W1[chunk g], then those of W2[:, chunk g-1] =====
      int stage = 0; uint32_t phase = 0;
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
          }
          for (int kb = 0; kb < KB1; ++kb) {
            ptx::mbar_wait(&empty[stage], phase ^ 1);
            ptx::mbar_expect_tx(&full[stage], ML_HC * ML_BK * 2);
            uint8_t* dst = ring + stage * STAGE + rank * w1_rows * (ML_BK * 2);
            if (CL > 1) ptx::tma_load_2d_mc(dst, &tmW1, &full[stage], kb * ML_BK, j * ML_HC + rank * w1_rows, (uint16_t)((1u << CL) - 1));
            else ptx::tma_load_2d(dst, &tmW1, &full[stage], kb * ML_BK, j * ML_HC);
            if (++stage == NST) { stage = 0; phase ^= 1; }
          }
        }
        if (g >= 1) {
          const int jp = (g - 1) % NCH;
          for (int kb = 0; kb < ML_HC / ML_BK; ++kb) {
            ptx::mbar_wait(&empty[stage], phase ^ 1);
            ptx::mbar_expect_tx(&full[stage], C * ML_BK * 2);
            uint8_t* dst = ring + stage * STAGE + rank * w2_rows * (ML_BK * 2);
            if (CL > 1) ptx::tma_load_2d_mc(dst, &tmW2, &full[stage], jp * ML_HC + kb * ML_BK, rank * w2_rows, (uint16_t)((1u << CL) - 1));
            else ptx::tma_load_2d(dst, &tmW2, &full[stage], jp * ML_HC + kb * ML_BK, 0);
            if (++stage == NST) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      // ===== MMA issuer =====
      const uint32_t idesc1 = ptx::make_idesc_16(ML_BM, ML_HC, 0, 0, p.in_bf16);
      const uint32_t idesc2 = ptx::make_idesc_16(ML_BM, C, 0, 0, p.in_bf16);
      int stage = 0; uint32_t phase = 0;
      auto next_stage = [&] { if (++stage == NST) { stage = 0; phase ^= 1; } };
      auto release = [&](uint64_t* bar) {
        if (CL > 1) ptx::umma_commit_mc(bar, (uint16_t)((1u << CL) - 1)); else ptx::umma_commit(bar);
      };
      for (int g = 0; g <= G; ++g) {
        if (g < G) {
          // G1(g): H[g & 1] = A x W1[chunk]^T
          const int it = g / NCH, j = g - it * NCH, hb = g & 1;
          const uint32_t ph = (g >> 1) & 1;
          if (j == 0) ptx::mbar_wait(afull, it & 1);
          ptx::mbar_wait(&hfree[hb], ph ^ 1);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + hb * ML_HC;
          for (int kb = 0; kb < KB1; ++kb) {
            ptx::mbar_wait(&full[stage], phase);
            ptx::tc_fence_after();
            const uint64_t a_desc = ptx::make_smem_desc(ptx::smem_u32(sA + kb * ML_AKB), 16, 1024, ptx::SW_128B);
            const uint64_t b_desc = ptx::make_smem_desc(ptx::smem_u32(ring + stage * STAGE), 16, 1024, ptx::SW_128B);
#pragma unroll
            for (int k = 0; k < ML_BK / 16; ++k) ptx::umma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc1, (kb | k) != 0);
            release(&empty[stage]);
            next_stage();
          }
          ptx::umma_commit(&hfull[hb]);
          if (j == NCH - 1) ptx::umma_commit(aempty);       // the A tile may be overwritten by the next one
        }
        if (g >= 1) {
          // G2(g - 1): O += Hs[(g - 1) & 1] x W2[:, chunk]^T
          const int gp = g - 1, itp = gp / NCH, jp = gp - itp * NCH, hb = gp & 1;
          const uint32_t ph = (gp >> 1) & 1;
          if (jp == 0) ptx::mbar_wait(ofree, (itp & 1) ^ 1);   // E2 of the previous tile has read O
          ptx::mbar_wait(&hsfull[hb], ph);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + K::TMEM_O;
          for (int kb = 0; kb < ML_HC / ML_BK; ++kb) {
            ptx::mbar_wait(&full[stage], phase);
            ptx::tc_fence_after();
            const uint64_t a_desc = ptx::make_smem_desc(ptx::smem_u32(sHs + hb * ML_HS_BYTES + kb * ML_AKB), 16, 1024, ptx::SW_128B);
            const uint64_t b_desc = ptx::make_smem_desc(ptx::smem_u32(ring + stage * STAGE), 16, 1024, ptx::SW_128B);
#pragma unroll
            for (int k = 0; k < ML_BK / 16; ++k) ptx::umma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc2, (jp | kb | k) != 0);
            release(&empty[stage]);
            next_stage();
          }
          ptx::umma_commit(&hsempty[hb]);
          if (jp == NCH - 1) ptx::umma_commit(ofull);
        }
      }
    }
  } else {
    // ===== epilogue warps: warp % 4 = TMEM lane quadrant, (warp - 2) / 4 = column half =====
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int sw = row & 7;
    // this thread's row inside K block `part` of a hidden chunk (128-byte rows, 8-row swizzle atoms of 1 KB)
    const uint32_t hs_row = ptx::smem_u32(sHs) + part * ML_AKB + (row >> 3) * 1024 + (row & 7) * 128;
    // E2 staging block: the first 2 KB of this warp's OWN 4 KB of Hs[0] (both hidden buffers are idle while O is drained,
    // and nobody else ever touches these rows)
    const uint32_t stage = ptx::smem_u32(sHs) + part * ML_AKB + q * 4096;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    constexpr int nch = C / 16, per = (nch + 1) / 2;
    const int c0 = min(part * per, nch) * 16, c1 = min((part + 1) * per, nch) * 16;
    auto grow_of = [&](int it) -> long long {
      if (it >= n_iter) return -1;
      const long long r = (long long)((cid + it * ncl) * CL + rank) * ML_BM + row;
      return r < p.rows ? r : -1;
    };
    float2 mr_next = make_float2(0.f, 1.f);
    { const long long r0 = grow_of(0); if (r0 >= 0) mr_next = __ldg(p.mr + r0); }
    for (int it = 0; it < n_iter; ++it) {
      const long long grow = grow_of(it);
      const unsigned long long nmu2 = pk2(mr_next.x, mr_next.x), rstd2 = pk2(mr_next.y, mr_next.y);
      { const long long rn = grow_of(it + 1); if (rn >= 0) mr_next = __ldg(p.mr + rn); }
      for (int j = 0; j < NCH; ++j) {
        const int g = it * NCH + j, hb = g & 1;
        const uint32_t ph = (g >> 1) & 1;
        ptx::mbar_wait(&hfull[hb], ph);
        ptx::tc_fence_after();
        uint32_t va[32], vb[32];
        const uint32_t taddr = lane_base + hb * ML_HC + part * 64;
        ptx::tmem_ld32(taddr, va);
        ptx::tmem_ld32(taddr + 32, vb);
        tmem_wait_dep(va);
        tmem_wait_dep(vb);
        ptx::tc_fence_before();
        ptx::mbar_arrive(&hfree[hb]);                   // G1 of chunk g + 2 may overwrite H[hb]
        ptx::mbar_wait(&hsempty[hb], ph ^ 1);           // G2 of chunk g - 2 has read Hs[hb]
        const uint32_t cb = (uint32_t)(j * ML_HC + part * 64) * 4;
        const uint32_t sb = ptx::smem_u32(sB1) + cb, scs = ptx::smem_u32(sCs1) + cb;
        const uint32_t dst = hs_row + hb * ML_HS_BYTES;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          ptx::sts128(dst + ((c ^ sw) << 4), mlp_gelu8(va + 8 * c, sb + c * 32, scs + c * 32, nmu2, rstd2, p.in_bf16));
#pragma unroll
        for (int c = 0; c < 4; ++c)
          ptx::sts128(dst + (((4 + c) ^ sw) << 4), mlp_gelu8(vb + 8 * c, sb + (4 + c) * 32, scs + (4 + c) * 32, nmu2, rstd2, p.in_bf16));
        ptx::fence_proxy_async_smem();                  // generic-proxy stores -> visible to the MMA's operand reads
        ptx::mbar_arrive(&hsfull[hb]);
      }
      // E2: O + bias2 + residual -> fp32 stream (+ 16-bit copy + LayerNorm partials)
      ptx::mbar_wait(ofull, it & 1);
      ptx::tc_fence_after();
      epi_warp<ACT_NONE, true, 1, false, false, true>(p.epi, lane_base + K::TMEM_O, 0, c0, c1, grow, ptx::smem_u32(sB2),
                                                      stage, lane, 0.f, 1.f, 0, part);
      ptx::tc_fence_before();
      ptx::mbar_arrive(ofree);
    }
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
  if (!a.fc1 || !a.fc2 || !a.mr || !a.lne.stats || !a.lne.x16) return false;
  const int C = a.x16.C;
  if (C != 128 && C != 192 && C != 256) return false;
  if (a.x16.dt != BF16 && a.x16.dt != F16) return false;
  const LayerW &w1 = *a.fc1, &w2 = *a.fc2;
  if (w1.taps() != 1 || w2.taps() != 1 || w1.Cin != C || w1.cin_pad != C || w1.N != 4 * C || w2.Cin != 4 * C ||
      w2.cin_pad != 4 * C || w2.N != C)
    return false;
  if (!w1.w16(a.x16.dt) || !w2.w16(a.x16.dt) || !w1.colsum(a.x16.dt) || !w1.bias) return false;
  if (a.x16.B != 1 || a.x16.H != 1 || a.x16.ld % 8 != 0 || ((uintptr_t)a.x16.p & 15)) return false;
  if (a.xt.dt != F32 || a.xt.rows() != a.x16.rows() || a.xt.C != C || a.xt.ld % 4 != 0 || ((uintptr_t)a.xt.p & 15)) return false;
  if (a.lne.x16dt != a.x16.dt || a.lne.ldx16 % 4 != 0 || ((uintptr_t)a.lne.x16 & 7)) return false;
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
  auto launch = [&](auto kern) {
    BRN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
    const int pairs = (p.m_tiles + CL - 1) / CL;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CL * std::min(pairs, sms / CL));
    cfg.blockDim = dim3(ML_THREADS);
    cfg.dynamicSmemBytes = K::SMEM;
    cfg.stream = ctx.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    BRN_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmW1, tmW2, p));
  };
  if (CL == 2) launch(tc_mlp_kernel<C, 2>); else launch(tc_mlp_kernel<C, 1>);
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
  p.in_bf16 = a.x16.dt == BF16 ? 1 : 0;
  p.bias1 = a.fc1->bias; p.colsum1 = a.fc1->colsum(a.x16.dt); p.bias2 = a.fc2->bias;
  p.mr = a.mr;
  EpiP e{};
  e.N = C; e.act = ACT_NONE;
  e.res = a.xt.p; e.resdt = F32; e.ldres = a.xt.ld;
  e.out = a.xt.p; e.odt = F32; e.ldo = a.xt.ld;
  e.vec = 1;
  e.lne_stats = a.lne.stats; e.lne_stride = a.lne.stride; e.x16 = a.lne.x16; e.x16dt = a.lne.x16dt; e.ldx16 = a.lne.ldx16;
  p.epi = e;
  const double rows = (double)p.rows;
  char desc[96] = "";
  if (ctx.kt) snprintf(desc, sizeof desc, "mlp M=%lld C=%d hid=%d tiles=%d", (long long)p.rows, C, 4 * C, p.m_tiles);
  KScope ks(ctx, KC_GEMM_TC, 2.0 * rows * C * 4 * C * 2, rows * C * (2 + 4 + 4 + 2) + 2.0 * 4 * C * C * 2, desc);
  switch (C) {
    case 128: launch_mlp<128>(ctx, a, p); break;
    case 192: launch_mlp<192>(ctx, a, p); break;
    default: launch_mlp<256>(ctx, a, p); break;
  }
}

}  // namespace brn
