// Modulated deformable convolution (DCNv2) as a gather producer feeding a tcgen05 implicit GEMM (sm_100a).
//
// Replaces `call_deformable_im2col` + `weight.matmul(columns)` of the reference's Metal path (src/aspp.rs:138-164,
// src/deform_conv.rs:177-214): the `columns[C*k*k, B*H*W]` matrix (822 MB fp32 per image for the 7x7 branch at
// 256^2) is never materialised.  Per 128-pixel M tile and per tap (one tap = one 64-channel K slab):
//   warps 0-7 : gather.  Two threads per pixel (32 channels each): read (dy, dx, modulator) of the tap, form the 4
//               bilinear corner weights (torchvision semantics: zero outside (-1,H)x(-1,W), per-corner validity),
//               fold the modulator into them, load the corners as 16-byte NHWC vectors, blend in fp32, round once
//               to bf16 and store into the 128B-swizzled K-major A tile (chunk ^ (row & 7)).
//   warp 8    : one elected thread TMA-loads the tap's weight slab W[N, tap, 0:64] (B operand) and issues
//               tcgen05.mma (M=128, N<=256, 4 x K=16) once producers + TMA have filled the stage.
//   warps 9-12: epilogue (shared with the GEMM kernel): bias (folded BatchNorm) + ReLU -> bf16 into the ASPP concat.
// 4-stage ring (16 KB A + 32 KB B per stage), two TMEM accumulator stages, persistent over M tiles.
// Bound: gather (L1/L2 bandwidth of 4 corners x 128 B per pixel-tap) vs tensor pipe 512 cycles per tap-tile.
#include <cuda.h>

#include <cstdio>

#include "brn_common.h"
#include "device_utils.cuh"
#include "tc_epilogue.cuh"
#include "tc_ptx.cuh"

namespace brn {

CUtensorMap make_tmap_16(const void* base, int dt, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box, CUtensorMapSwizzle swz);
int device_sm_count();
EpiP make_epi(int N, const float* bias, int bias_bstride, int act, int act_from, const View& res, const View& out);

constexpr int DF_STAGES = 2;                      // 2 x 48 KB: leaves ~100 KB of L1 for the gathered neighbourhood (84 KB at k=7)
constexpr int DF_PRODUCERS = 512;                 // 16 warps = 2 groups of 8: group q gathers the taps g with g % 2 == q
constexpr int DF_GROUP = DF_PRODUCERS / 2;
constexpr int DF_MMA_WARP = DF_PRODUCERS / 32;    // warp 16: weight TMA + MMA issue; warps 17-20: epilogue
constexpr int DF_THREADS = DF_PRODUCERS + 32 + 128;
constexpr int DF_A_BYTES = 128 * 128;             // 128 pixels x 64 ch bf16
constexpr int DF_B_BYTES = 256 * 128;
constexpr int DF_BIAS_LD = 288;
constexpr int DF_PARAM_BYTES = 2 * 2 * 128 * 32;  // per group: double-buffered sampling parameters, 128 pixels x 32 B
constexpr int DF_EPI_BYTES = 4 * EPI_STAGE_BYTES + DF_BIAS_LD * 4 + DF_PARAM_BYTES;
constexpr int DF_SMEM = DF_STAGES * (DF_A_BYTES + DF_B_BYTES) + DF_EPI_BYTES + 256 + 1024;
constexpr int DF_TW = 16, DF_TH = 8;              // an M tile is a 16 x 8 pixel patch of one image (2-D gather locality)

struct DeformP {
  const uint16_t* x; int ldx; int B, H, W;   // bf16 or fp16 elements (xdt)
  int xdt;
  const float* om; int ldom; int om_tiled;   // om_tiled: [m_tile][3*taps][128] (written by tc_gemm with out_tiled)
  int k, pad, taps;
  int tiles_x, tiles_y, m_tiles; int BN;
  EpiP epi;
};

__device__ __forceinline__ float ldg_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

template <int DT>
__device__ __forceinline__ void fma_16x8(float (&acc)[8], const uint4& v, float w) {
  const uint32_t* h = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    float2 f;
    if (DT == BF16) { f.x = __uint_as_float(h[t] << 16); f.y = __uint_as_float(h[t] & 0xffff0000u); }
    else f = __half22float2(*reinterpret_cast<const __half2*>(&h[t]));
    acc[2 * t] = fmaf(w, f.x, acc[2 * t]);
    acc[2 * t + 1] = fmaf(w, f.y, acc[2 * t + 1]);
  }
}
template <int DT>
__device__ __forceinline__ uint32_t pack2(float a, float b) { return DT == BF16 ? pack_bf16x2(a, b) : pack_f16x2(a, b); }

// base + 32-bit byte offset as one IMAD.WIDE.U32 (the plain C++ form compiles to an IMAD + two IADD3 per load: the
// 16 corner addresses per thread-tap were 20 % of the kernel's instructions)
__device__ __forceinline__ const uint4* addr_off(const void* base, uint32_t off_bytes) {
  unsigned long long r;
  asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(r) : "r"(off_bytes), "l"((unsigned long long)base));
  return reinterpret_cast<const uint4*>(r);
}
// fp16 operands: blend the 4 corners in packed half2 arithmetic (one HMUL2 + three HFMA2 per channel pair instead of
// 4 x (unpack + FFMA) + pack).  The weights arrive as duplicated half2; the result is the A-operand bit pattern.
__device__ __forceinline__ uint4 blend_h2(const uint4 (&v)[4], const uint32_t (&w)[4]) {
  uint4 r;
  uint32_t* ro = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    __half2 a = __hmul2(*reinterpret_cast<const __half2*>(&reinterpret_cast<const uint32_t*>(&v[0])[t]), *reinterpret_cast<const __half2*>(&w[0]));
#pragma unroll
    for (int c = 1; c < 4; ++c)
      a = __hfma2(*reinterpret_cast<const __half2*>(&reinterpret_cast<const uint32_t*>(&v[c])[t]), *reinterpret_cast<const __half2*>(&w[c]), a);
    ro[t] = *reinterpret_cast<uint32_t*>(&a);
  }
  return r;
}

// CL = CTAs per cluster.  With CL = 2 the two CTAs walk M tiles 2j and 2j+1 through the same tap sequence; each loads
// half of the tap's weight slab and TMA-multicasts it into both CTAs (the kernel is L2 -> SM bandwidth bound: 32 KB
// of weights + the L1 misses of the gather per tap-tile).  DT = operand element type (BF16 / F16).
template <int CL, int DT>
__global__ void __launch_bounds__(DF_THREADS, 1)
tc_deform_kernel(const __grid_constant__ CUtensorMap tmB, const DeformP p) {
  ptx::pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + DF_STAGES * DF_A_BYTES;
  uint8_t* sStage = sB + DF_STAGES * DF_B_BYTES;                 // epilogue staging, 2 KB per epilogue warp
  float* sBias = (float*)(sStage + 4 * EPI_STAGE_BYTES);
  uint8_t* sParams = (uint8_t*)sBias + DF_BIAS_LD * 4;
  uint64_t* full = (uint64_t*)(sParams + DF_PARAM_BYTES);
  uint64_t* empty = full + DF_STAGES;
  uint64_t* tfull = empty + DF_STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    // full: one arrival per producer warp (after its lanes' proxy fences) + the weight TMA's expect_tx arrival
    for (int s = 0; s < DF_STAGES; ++s) { ptx::mbar_init(&full[s], DF_GROUP / 32 + 1); ptx::mbar_init(&empty[s], CL); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull[a], 1); ptx::mbar_init(&tempty[a], 128); }
    ptx::fence_barrier_init();
  }
  if (warp == DF_MMA_WARP) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  if (CL > 1) ptx::cluster_sync_all();     // peers' barriers are initialised before any multicast can land
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_wait();
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  // work items: this CTA's M tile of item i is i * CL + rank (may be >= m_tiles for the odd tail: all rows masked)
  const int rank = CL > 1 ? (int)ptx::cluster_ctarank() : 0;
  const int n_items = (p.m_tiles + CL - 1) / CL;
  const int item0 = blockIdx.x / CL, item_step = gridDim.x / CL;

  if (warp < DF_MMA_WARP) {
    // ===== gather producers =====
    // Two groups of 8 warps; the taps of this CTA's tile sequence are numbered g = 0, 1, 2, ... (tile-major) and group
    // q owns the taps with g % 2 == q, i.e. ring stage q.  While one group waits for its corner loads the other one
    // blends -- the gather is latency bound (L2 hits of a 2-D neighbourhood), not bandwidth bound.
    // Sampling parameters (4 modulated corner weights + 4 clamped corner offsets per pixel) are computed once per
    // (pixel, tap) by one thread pair of the group and shared through shared memory; the gather maps 8 consecutive
    // lanes to the 8 16-byte chunks of ONE corner row (64 channels = 128 B), so a warp instruction touches 4 full
    // cache lines instead of 32 partial ones (the L1 tag rate is one line per cycle).
    static_assert(DF_STAGES == 2, "group q <-> ring stage q");
    const int q = warp >> 3, gtid = threadIdx.x & (DF_GROUP - 1);
    const int pr = gtid & 127, prow = gtid >> 7;      // parameter role: pixel pr, corner row prow (y0 or y0 + 1)
    const int grp = gtid >> 3, l8 = gtid & 7;         // gather role: pixels grp + 32 i, channels [8 l8, 8 l8 + 8)
    const int ncols = 3 * p.taps;
    const uint32_t sPar = ptx::smem_u32(sParams) + q * 8192;
    const uint32_t sa = ptx::smem_u32(sA) + q * DF_A_BYTES;
    const int n_my_items = item0 < n_items ? (n_items - 1 - item0) / item_step + 1 : 0;
    const long long img_elems = (long long)p.H * p.W * p.ldx;

    // position in the tap sequence: local item `it`, tap (ky, kx); per-pixel data of the parameter role
    int it = 0, tap = q, ky = 0, kx = 0, y = 0, x = 0, tile = 0, img = 0;
    bool row_ok = false;
    const float* o = p.om; long long os1 = 1;
    auto enter_tile = [&]() {
      while (tap >= p.taps) { tap -= p.taps; ++it; }
      if (it >= n_my_items) return;
      const int t = (item0 + it * item_step) * CL + rank;
      tile = min(t, p.m_tiles - 1);
      const int b = tile / tiles_per_img, t2 = tile - b * tiles_per_img;
      img = b;
      y = (t2 / p.tiles_x) * DF_TH + (pr >> 4); x = (t2 % p.tiles_x) * DF_TW + (pr & 15);
      row_ok = t < p.m_tiles && y < p.H && x < p.W;
      ky = tap / p.k; kx = tap - ky * p.k;
      // offsets / modulator of tap t: o[2t * os1], o[(2t + 1) * os1], o[(2 taps + t) * os1]
      if (p.om_tiled) { o = p.om + (long long)tile * ncols * 128 + pr; os1 = 128; }
      else { o = p.om + (((long long)b * p.H + min(y, p.H - 1)) * p.W + min(x, p.W - 1)) * p.ldom; os1 = 1; }
    };
    auto advance2 = [&]() {
      tap += 2; kx += 2;
      if (tap >= p.taps) { enter_tile(); return; }
      if (kx >= p.k) { kx -= p.k; ++ky; }
      if (kx >= p.k) { kx -= p.k; ++ky; }
    };
    auto params = [&](int buf, float dy, float dx, float mk) {
      // torchvision semantics: zero outside (-1,H)x(-1,W), per-corner validity; folded into the corner weights so
      // that every corner load is unconditional (clamped address, weight 0)
      const float py = (float)(y - p.pad + ky) + dy, px = (float)(x - p.pad + kx) + dx;
      const bool inb = row_ok && py > -1.f && py < (float)p.H && px > -1.f && px < (float)p.W;
      const float fy = floorf(py), fx = floorf(px);
      const int yy = (int)fy + prow, x0 = (int)fx;
      const float ly = py - fy, lx = px - fx;
      const float wy = (yy >= 0 && yy <= p.H - 1 && inb) ? (prow ? ly : 1.f - ly) * mk : 0.f;
      const float wx0 = (x0 >= 0) ? 1.f - lx : 0.f, wx1 = (x0 + 1 <= p.W - 1) ? lx : 0.f;
      const int yc = min(max(yy, 0), p.H - 1);
      const int xc0 = min(max(x0, 0), p.W - 1), xc1 = min(max(x0 + 1, 0), p.W - 1);
      const uint32_t o0 = (uint32_t)((yc * p.W + xc0) * p.ldx) * 2u, o1 = (uint32_t)((yc * p.W + xc1) * p.ldx) * 2u;   // bytes
      uint32_t w0, w1;
      if (DT == F16) {
        const __half2 h0 = __float2half2_rn(wy * wx0), h1 = __float2half2_rn(wy * wx1);
        w0 = *reinterpret_cast<const uint32_t*>(&h0); w1 = *reinterpret_cast<const uint32_t*>(&h1);
      } else { w0 = __float_as_uint(wy * wx0); w1 = __float_as_uint(wy * wx1); }
      ptx::sts128(sPar + buf * 4096 + pr * 32 + prow * 16, make_uint4(w0, w1, o0, o1));
    };

    enter_tile();
    if (it < n_my_items) params(0, ldg_stream(o + 2 * tap * os1), ldg_stream(o + (2 * tap + 1) * os1), ldg_stream(o + (2 * p.taps + tap) * os1));
    asm volatile("bar.sync %0, 256;" ::"r"(2 + q) : "memory");
    for (int k = 0; it < n_my_items; ++k) {
      const uint16_t* xb = p.x + (long long)img * img_elems + l8 * 8;   // image of the CURRENT tap's tile (no division per tap)
      // step to the group's next tap and fetch its offsets now; its parameters are computed below, while the second
      // half of this tap's corners is in flight
      advance2();
      const bool more = it < n_my_items;
      float ndy = 0.f, ndx = 0.f, nmk = 0.f;
      if (more) { ndy = ldg_stream(o + 2 * tap * os1); ndx = ldg_stream(o + (2 * tap + 1) * os1); nmk = ldg_stream(o + (2 * p.taps + tap) * os1); }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        // ---- gather: 2 pixels per half, 4 corners each ----
        uint4 v[2][4];
        uint32_t wgt[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const uint32_t pa = sPar + (k & 1) * 4096 + (grp + 32 * (2 * h + i)) * 32;
          const uint4 q0 = ptx::lds128(pa), q1 = ptx::lds128(pa + 16);
          wgt[i][0] = q0.x; wgt[i][1] = q0.y; wgt[i][2] = q1.x; wgt[i][3] = q1.y;
          v[i][0] = __ldg(addr_off(xb, q0.z));
          v[i][1] = __ldg(addr_off(xb, q0.w));
          v[i][2] = __ldg(addr_off(xb, q1.z));
          v[i][3] = __ldg(addr_off(xb, q1.w));
        }
        if (h == 1 && more) params((k + 1) & 1, ndy, ndx, nmk);
        uint4 res[2];
        if (DT == F16) {
#pragma unroll
          for (int i = 0; i < 2; ++i) res[i] = blend_h2(v[i], wgt[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            float acc[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
            for (int cnr = 0; cnr < 4; ++cnr) fma_16x8<DT>(acc, v[i][cnr], __uint_as_float(wgt[i][cnr]));
            res[i] = make_uint4(pack2<DT>(acc[0], acc[1]), pack2<DT>(acc[2], acc[3]), pack2<DT>(acc[4], acc[5]), pack2<DT>(acc[6], acc[7]));
          }
        }
        if (h == 0) ptx::mbar_wait_backoff(&empty[q], (uint32_t)((k & 1) ^ 1));
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int r = grp + 32 * (2 * h + i);
          ptx::sts128(sa + r * 128 + ((l8 ^ (r & 7)) << 4), res[i]);
        }
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&full[q]);
      // the next tap's parameters are visible to the whole group; everybody is done reading this tap's
      asm volatile("bar.sync %0, 256;" ::"r"(2 + q) : "memory");
    }
  } else if (warp == DF_MMA_WARP) {
    if (ptx::elect_one()) {
      // ===== weight TMA + MMA issuer =====
      ptx::prefetch_tmap(&tmB);
      const uint32_t idesc = ptx::make_idesc_16(128, p.BN, 0, 0, DT == BF16 ? 1u : 0u);
      const uint32_t b_bytes = p.BN * 128;
      const int b_rows = p.BN / CL;                            // rows of the slab this CTA loads (and multicasts)
      const int n_my_tiles = item0 < n_items ? (n_items - 1 - item0) / item_step + 1 : 0;
      const long long total = (long long)n_my_tiles * p.taps;
      // the B loads run DF_STAGES ahead of the MMAs (same ring, same stage order)
      long long issued = 0;
      int btap = 0;                                            // tap of the next B load (issued % taps without the 64-bit division)
      int lstage = 0; uint32_t lphase = 0;
      auto issue_b = [&]() {
        ptx::mbar_wait(&empty[lstage], lphase ^ 1);
        ptx::mbar_expect_tx(&full[lstage], b_bytes);
        const int tap = btap;
        if (++btap == p.taps) btap = 0;
        uint8_t* bdst = sB + lstage * DF_B_BYTES + rank * b_rows * 128;
        if (CL > 1) ptx::tma_load_2d_mc(bdst, &tmB, &full[lstage], tap * 64, rank * b_rows, (uint16_t)((1u << CL) - 1));
        else ptx::tma_load_2d(bdst, &tmB, &full[lstage], tap * 64, 0);
        ++issued;
        if (++lstage == DF_STAGES) { lstage = 0; lphase ^= 1; }
      };
      for (int i = 0; i < DF_STAGES && issued < total; ++i) issue_b();
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = 0; t < n_my_tiles; ++t) {
        ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int tap = 0; tap < p.taps; ++tap) {
          ptx::mbar_wait_backoff(&full[stage], phase);      // the gather is the pace setter: do not burn issue slots spinning
          ptx::tc_fence_after();
          const uint64_t a_desc = ptx::make_smem_desc(ptx::smem_u32(sA + stage * DF_A_BYTES), 16, 1024, ptx::SW_128B);
          const uint64_t b_desc = ptx::make_smem_desc(ptx::smem_u32(sB + stage * DF_B_BYTES), 16, 1024, ptx::SW_128B);
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (tap | k) != 0);
          // the stage is free once the MMAs of EVERY CTA that received the multicast have read it
          if (CL > 1) ptx::umma_commit_mc(&empty[stage], (uint16_t)((1u << CL) - 1));
          else ptx::umma_commit(&empty[stage]);
          if (++stage == DF_STAGES) { stage = 0; phase ^= 1; }
          if (issued < total) issue_b();
        }
        ptx::umma_commit(&tfull[acc]);
        acc ^= 1; if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===== epilogue (4 warps, one TMEM lane quadrant each, all BN columns) =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int eth = threadIdx.x - (DF_PRODUCERS + 32);
    const uint32_t stage = ptx::smem_u32(sStage) + (warp - DF_MMA_WARP - 1) * EPI_STAGE_BYTES;
    const uint32_t sb = ptx::smem_u32(sBias);
    {
      for (int t = eth; t < DF_BIAS_LD; t += 128)
        ptx::sts32(sb + t * 4, (p.epi.bias && t < p.epi.N) ? __ldg(p.epi.bias + t) : 0.f);
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    const int c1 = ((p.epi.N + 15) >> 4) * 16;
    int acc = 0; uint32_t acc_phase = 0;
    for (int item = item0; item < n_items; item += item_step) {
      const int tile = item * CL + rank;
      const int b = tile / tiles_per_img, t2 = tile - b * tiles_per_img;
      const int y = (t2 / p.tiles_x) * DF_TH + (row >> 4), x = (t2 % p.tiles_x) * DF_TW + (row & 15);
      const long long orow = (tile < p.m_tiles && y < p.H && x < p.W) ? ((long long)b * p.H + y) * p.W + x : -1;
      ptx::mbar_wait(&tfull[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256;
      // only the epilogues a deformable conv can have (no residual; none / ReLU): keeps the 80-register kernel small
      const uint32_t sbb = sb;
      if (p.epi.odt == F32) {
        if (p.epi.act == ACT_RELU) epi_warp<ACT_RELU, true, 0, false>(p.epi, taddr, 0, 0, c1, orow, sbb, stage, lane);
        else epi_warp<ACT_NONE, true, 0, false>(p.epi, taddr, 0, 0, c1, orow, sbb, stage, lane);
      } else {
        if (p.epi.act == ACT_RELU) epi_warp<ACT_RELU, false, 0, false>(p.epi, taddr, 0, 0, c1, orow, sbb, stage, lane);
        else epi_warp<ACT_NONE, false, 0, false>(p.epi, taddr, 0, 0, c1, orow, sbb, stage, lane);
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tempty[acc]);
      acc ^= 1; if (acc == 0) acc_phase ^= 1;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CL > 1) ptx::cluster_sync_all();     // no CTA leaves while a peer may still multicast into it / arrive on its barriers
  if (warp == DF_MMA_WARP) ptx::tmem_dealloc(tmem_base, 512);
}

bool tc_deform_supported(const DeformArgs& a) {
  return a.w && a.stride == 1 && (a.pad < 0 || a.pad == a.w->kh / 2) && (a.act == ACT_NONE || a.act == ACT_RELU) && (a.x.dt == BF16 || a.x.dt == F16) && a.w->w16(a.x.dt) && a.x.C == 64 && a.w->cin_pad == 64 && a.x.ld % 8 == 0 &&
         (((uintptr_t)a.x.p) & 15) == 0 && a.w->N <= 256 && a.om.dt == F32;
}

void tc_deform(const LaunchCtx& ctx, const DeformArgs& a) {
  if (ctx.launches) ++*ctx.launches;
  if (ctx.dry) return;
  const LayerW& w = *a.w;
  DeformP p{};
  p.x = (const uint16_t*)a.x.p; p.xdt = a.x.dt; p.ldx = a.x.ld; p.B = a.x.B; p.H = a.x.H; p.W = a.x.W;
  p.om = (const float*)a.om.p; p.ldom = a.om.ld; p.om_tiled = a.om_tiled;
  p.k = w.kh; p.pad = w.kh / 2; p.taps = w.taps();
  p.tiles_x = (p.W + DF_TW - 1) / DF_TW; p.tiles_y = (p.H + DF_TH - 1) / DF_TH;
  p.m_tiles = p.B * p.tiles_x * p.tiles_y;
  p.BN = (w.N + 15) / 16 * 16;
  View none{};
  p.epi = make_epi(w.N, a.bias ? a.bias : w.bias, 0, a.act, 0, none, a.out);
  const uint64_t ktot = (uint64_t)w.taps() * 64;
  uint64_t bdims[2] = {ktot, (uint64_t)w.N};
  uint64_t bstr[1] = {ktot * 2};
  const int sms = device_sm_count();
  const int CL = (p.BN % 16 == 0 && p.m_tiles >= 2 * sms) ? 2 : 1;
  uint32_t bbox[2] = {64, (uint32_t)(p.BN / CL)};
  CUtensorMap tmB = make_tmap_16(w.w16(a.x.dt), a.x.dt, 2, bdims, bstr, bbox, CU_TENSOR_MAP_SWIZZLE_128B);
  char desc[96] = "";
  const double M = (double)a.x.rows();
  // bytes = COMPULSORY HBM traffic (input + offsets/modulators + output + weights once); the gather's sampled bytes
  // (4 corners x 128 B per pixel and tap, served by L1 / L2) go into the description
  if (ctx.kt) snprintf(desc, sizeof desc, "M=%lld N=%d k=%d cl=%d sampled_mb=%.1f", (long long)a.x.rows(), w.N, p.k, CL,
                       M * w.taps() * 4 * 128 / 1e6);
  KScope ks(ctx, KC_DEFORM_TC, 2.0 * M * w.N * w.taps() * 64,
            M * (64.0 * dsize(a.x.dt) + 3.0 * w.taps() * 4 + (double)w.N * dsize(a.out.dt)) + (double)w.N * w.taps() * 64 * 2, desc);
  auto launch = [&](auto kern) {
    BRN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, DF_SMEM));
    // keep the shared-memory carve-out at what the kernel needs: the rest of the 228 KB is the gather's L1
    BRN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (DF_SMEM * 100 + 228 * 1024 - 1) / (228 * 1024)));
    const int items = (p.m_tiles + CL - 1) / CL;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CL * std::min(items, sms / CL));
    cfg.blockDim = dim3(DF_THREADS);
    cfg.dynamicSmemBytes = DF_SMEM;
    cfg.stream = ctx.stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_attr(attr, 1);
    BRN_CUDA(cudaLaunchKernelEx(&cfg, kern, tmB, p));
  };
  if (CL == 2) { if (a.x.dt == BF16) launch(tc_deform_kernel<2, BF16>); else launch(tc_deform_kernel<2, F16>); }
  else { if (a.x.dt == BF16) launch(tc_deform_kernel<1, BF16>); else launch(tc_deform_kernel<1, F16>); }
  BRN_CUDA(cudaGetLastError());
}

}  // namespace brn
