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

constexpr int DF_STAGES = 4;
constexpr int DF_PRODUCERS = 256;                 // 8 warps
constexpr int DF_THREADS = DF_PRODUCERS + 32 + 128;
constexpr int DF_A_BYTES = 128 * 128;             // 128 pixels x 64 ch bf16
constexpr int DF_B_BYTES = 256 * 128;
constexpr int DF_SMEM = DF_STAGES * (DF_A_BYTES + DF_B_BYTES) + 256 + 1024;

struct DeformP {
  const uint16_t* x; int ldx; int B, H, W;   // bf16 or fp16 elements (xdt)
  int xdt;
  const float* om; int ldom;
  int k, pad, taps;
  long long M; int m_tiles; int BN;
  EpiP epi;
};

__device__ __forceinline__ void fma_16x8(float (&acc)[8], const uint4& v, float w, int dt) {
  const uint32_t* h = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    float2 f = unpack16x2(h[t], dt);
    acc[2 * t] = fmaf(w, f.x, acc[2 * t]);
    acc[2 * t + 1] = fmaf(w, f.y, acc[2 * t + 1]);
  }
}

__global__ void __launch_bounds__(DF_THREADS, 1)
tc_deform_kernel(const __grid_constant__ CUtensorMap tmB, const DeformP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + DF_STAGES * DF_A_BYTES;
  uint64_t* full = (uint64_t*)(sB + DF_STAGES * DF_B_BYTES);
  uint64_t* empty = full + DF_STAGES;
  uint64_t* tfull = empty + DF_STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < DF_STAGES; ++s) { ptx::mbar_init(&full[s], DF_PRODUCERS + 1); ptx::mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull[a], 1); ptx::mbar_init(&tempty[a], 128); }
    ptx::fence_barrier_init();
  }
  if (warp == 8) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // ===== gather producers =====
    const int r = threadIdx.x & 127, half = threadIdx.x >> 7;
    const int HW = p.H * p.W;
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
      const long long m = (long long)tile * 128 + r;
      const bool row_ok = m < p.M;
      int b = 0, y = 0, x = 0;
      if (row_ok) { b = (int)(m / HW); int rem = (int)(m - (long long)b * HW); y = rem / p.W; x = rem - y * p.W; }
      const uint16_t* xb = p.x + (long long)b * HW * p.ldx + half * 32;
      const float* o = p.om + m * p.ldom;
      for (int tap = 0; tap < p.taps; ++tap) {
        float acc[4][8];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[j][e] = 0.f;
        if (row_ok) {
          const int ky = tap / p.k, kx = tap - ky * p.k;
          const float py = (float)(y - p.pad + ky) + __ldg(o + 2 * tap);
          const float px = (float)(x - p.pad + kx) + __ldg(o + 2 * tap + 1);
          const float mk = __ldg(o + 2 * p.taps + tap);
          if (py > -1.f && py < (float)p.H && px > -1.f && px < (float)p.W) {
            const int y0 = (int)floorf(py), x0 = (int)floorf(px);
            const float ly = py - y0, lx = px - x0, hy = 1.f - ly, hx = 1.f - lx;
            const bool y0ok = y0 >= 0, y1ok = y0 + 1 <= p.H - 1, x0ok = x0 >= 0, x1ok = x0 + 1 <= p.W - 1;
            const float wgt[4] = {hy * hx * mk, hy * lx * mk, ly * hx * mk, ly * lx * mk};
            const bool ok[4] = {y0ok && x0ok, y0ok && x1ok, y1ok && x0ok, y1ok && x1ok};
            const int cy[4] = {y0, y0, y0 + 1, y0 + 1}, cx[4] = {x0, x0 + 1, x0, x0 + 1};
#pragma unroll
            for (int cnr = 0; cnr < 4; ++cnr) {
              if (!ok[cnr]) continue;
              const uint4* src = reinterpret_cast<const uint4*>(xb + ((long long)cy[cnr] * p.W + cx[cnr]) * p.ldx);
              uint4 v[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) v[j] = __ldg(src + j);
#pragma unroll
              for (int j = 0; j < 4; ++j) fma_16x8(acc[j], v[j], wgt[cnr], p.xdt);
            }
          }
        }
        ptx::mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* rowp = sA + stage * DF_A_BYTES + r * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int chunk = (half * 4 + j) ^ (r & 7);
          *reinterpret_cast<uint4*>(rowp + chunk * 16) =
              make_uint4(pack16x2(acc[j][0], acc[j][1], p.xdt), pack16x2(acc[j][2], acc[j][3], p.xdt),
                         pack16x2(acc[j][4], acc[j][5], p.xdt), pack16x2(acc[j][6], acc[j][7], p.xdt));
        }
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&full[stage]);
        if (++stage == DF_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 8) {
    if (ptx::elect_one()) {
      // ===== weight TMA + MMA issuer =====
      ptx::prefetch_tmap(&tmB);
      const uint32_t idesc = ptx::make_idesc_16(128, p.BN, 0, 0, p.xdt == BF16 ? 1u : 0u);
      const uint32_t b_bytes = p.BN * 128;
      const int n_my_tiles = blockIdx.x < p.m_tiles ? (p.m_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
      const long long total = (long long)n_my_tiles * p.taps;
      // the B loads run DF_STAGES ahead of the MMAs (same ring, same stage order)
      long long issued = 0;
      int lstage = 0; uint32_t lphase = 0;
      auto issue_b = [&]() {
        ptx::mbar_wait(&empty[lstage], lphase ^ 1);
        ptx::mbar_expect_tx(&full[lstage], b_bytes);
        const int tap = (int)(issued % p.taps);
        ptx::tma_load_2d(sB + lstage * DF_B_BYTES, &tmB, &full[lstage], tap * 64, 0);
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
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          const uint64_t a_desc = ptx::make_smem_desc(ptx::smem_u32(sA + stage * DF_A_BYTES), 16, 1024, ptx::SW_128B);
          const uint64_t b_desc = ptx::make_smem_desc(ptx::smem_u32(sB + stage * DF_B_BYTES), 16, 1024, ptx::SW_128B);
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (tap | k) != 0);
          ptx::umma_commit(&empty[stage]);
          if (++stage == DF_STAGES) { stage = 0; phase ^= 1; }
          if (issued < total) issue_b();
        }
        ptx::umma_commit(&tfull[acc]);
        acc ^= 1; if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===== epilogue =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
      const long long orow = (long long)tile * 128 + row;
      const bool valid = orow < p.M;
      ptx::mbar_wait(&tfull[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256;
      epi_row(p.epi, taddr, 0, (p.epi.N + 15) >> 4, 0, 1, valid ? orow : 0, p.epi.bias, valid);
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tempty[acc]);
      acc ^= 1; if (acc == 0) acc_phase ^= 1;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) ptx::tmem_dealloc(tmem_base, 512);
}

bool tc_deform_supported(const DeformArgs& a) {
  return a.w && (a.x.dt == BF16 || a.x.dt == F16) && a.w->w16(a.x.dt) && a.x.C == 64 && a.w->cin_pad == 64 && a.x.ld % 8 == 0 &&
         (((uintptr_t)a.x.p) & 15) == 0 && a.w->N <= 256 && a.om.dt == F32;
}

void tc_deform(const LaunchCtx& ctx, const DeformArgs& a) {
  if (ctx.launches) ++*ctx.launches;
  if (ctx.dry) return;
  const LayerW& w = *a.w;
  DeformP p{};
  p.x = (const uint16_t*)a.x.p; p.xdt = a.x.dt; p.ldx = a.x.ld; p.B = a.x.B; p.H = a.x.H; p.W = a.x.W;
  p.om = (const float*)a.om.p; p.ldom = a.om.ld;
  p.k = w.kh; p.pad = w.kh / 2; p.taps = w.taps();
  p.M = a.x.rows(); p.m_tiles = (int)((p.M + 127) / 128);
  p.BN = (w.N + 15) / 16 * 16;
  View none{};
  p.epi = make_epi(w.N, a.bias ? a.bias : w.bias, 0, a.act, 0, none, a.out);
  const uint64_t ktot = (uint64_t)w.taps() * 64;
  uint64_t bdims[2] = {ktot, (uint64_t)w.N};
  uint64_t bstr[1] = {ktot * 2};
  uint32_t bbox[2] = {64, (uint32_t)p.BN};
  CUtensorMap tmB = make_tmap_16(w.w16(a.x.dt), a.x.dt, 2, bdims, bstr, bbox, CU_TENSOR_MAP_SWIZZLE_128B);
  cudaFuncSetAttribute(tc_deform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DF_SMEM);
  const int grid = std::min(p.m_tiles, device_sm_count());
  char desc[96] = "";
  if (ctx.kt) snprintf(desc, sizeof desc, "M=%lld N=%d k=%d", p.M, w.N, p.k);
  KScope ks(ctx, KC_DEFORM_TC, 2.0 * (double)p.M * w.N * w.taps() * 64, (double)p.M * w.taps() * 4 * 128, desc);
  tc_deform_kernel<<<grid, DF_THREADS, DF_SMEM, ctx.stream>>>(tmB, p);
  BRN_CUDA(cudaGetLastError());
}

}  // namespace brn
