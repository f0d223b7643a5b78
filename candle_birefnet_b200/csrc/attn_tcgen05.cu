// Fused shifted-window attention for 12x12 windows, head_dim 32 (sm_100a, tcgen05 + TMEM + TMA).
//
// Replaces WindowAttention::forward_standard (src/swin.rs:266-311) == the Metal flash_attention_with_[repeating_]bias
// calls (src/swin.rs:243,252): per (window, head)   O = softmax(q k^T + bias[h] + mask) v   with q pre-scaled, the
// relative-position bias of WindowAttention::new (src/swin.rs:143-152) and the analytic -100 region mask of
// BasicLayer::create_attention_mask (src/swin.rs:603-655).  The [nW,heads,144,144] score tensor never leaves the SM.
//
// Work unit = (window, head).  A persistent CTA owns ONE head (its 144x144 bias stays resident in shared memory as
// bf16, rows padded to 304 B so row-per-thread 16-byte reads are bank-conflict free) and walks the windows.
//   warp 5   : one elected thread issues TMA (Q,K,V tiles of the window-ordered qkv matrix, 64B-swizzled, 2 stages)
//              and all tcgen05.mma:  S = Q K^T (two M=128 tiles for the 144 queries, N=144, K=32) into TMEM,
//              then O = P V (M=128 x2, N=32, K=144; V is the MN-major B operand straight from the TMA tile).
//   warps 0-3: softmax for query rows 0..127 (one row per thread: TMEM lane == query), warp 4: rows 128..143.
//              Two passes over the S row (tcgen05.ld 32 columns at a time): max, then exp2 / sum / bf16 P written
//              to shared memory in the 32B-swizzled K-major layout the P V MMA reads.  1/sum is applied to O.
// Bound: MUFU (144*144 exp per unit) -- see DESIGN.md; tensor work per unit is ~600 cycles vs ~1300 of exp.
#include <cuda.h>

#include <cstdio>

#include "brn_common.h"
#include "device_utils.cuh"
#include "tc_ptx.cuh"
#include "tc_epilogue.cuh"

namespace brn {

CUtensorMap make_tmap_16(const void* base, int dt, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box, CUtensorMapSwizzle swz);
int device_sm_count();

constexpr int AT_THREADS = 192;
constexpr int AT_BIAS_LD = 152;                          // bf16 elements per bias row (304 B)
constexpr int AT_BIAS_BYTES = 144 * AT_BIAS_LD * 2;      // 43,776
constexpr int AT_BIAS_REGION = 44 * 1024;                // 45,056 (1 KB aligned)
constexpr int AT_TILE_BYTES = 144 * 64;                  // one of Q/K/V: 9,216
constexpr int AT_STAGE_BYTES = 3 * AT_TILE_BYTES;        // 27,648 = 27 KB
constexpr int AT_P_BLOCK = 144 * 32;                     // one K=16 step of P: 4,608
constexpr int AT_P_REGION = 44 * 1024;                   // 9 blocks (41,472) + over-read slack of the 16-row tile
constexpr int AT_SMEM = AT_BIAS_REGION + 2 * AT_STAGE_BYTES + AT_P_REGION + 256 + 1024;
constexpr uint32_t AT_COL_S0 = 0, AT_COL_S1 = 144, AT_COL_O0 = 288, AT_COL_O1 = 320;

struct AttnP {
  const uint16_t* bias16;        // [heads][144][152], bf16 or fp16 (dt)
  int dt;                        // operand / output element type: BF16 or F16
  int n_windows, heads, C;
  int nwh, nww, shift;
  uint16_t* out; int ldo;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(AT_THREADS, 1)
tc_attn_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sBias = smem;
  uint8_t* sQKV = smem + AT_BIAS_REGION;
  uint8_t* sP = sQKV + 2 * AT_STAGE_BYTES;
  uint64_t* bars = (uint64_t*)(sP + AT_P_REGION);
  uint64_t* qkv_full = bars;        // [2]
  uint64_t* qkv_empty = bars + 2;   // [2]
  uint64_t* s_full = bars + 4;
  uint64_t* p_full = bars + 5;
  uint64_t* o_full = bars + 6;
  uint64_t* bias_bar = bars + 7;
  uint32_t* tmem_slot = (uint32_t*)(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.x % p.heads;
  const int w_first = blockIdx.x / p.heads, w_step = gridDim.x / p.heads;

  if (threadIdx.x == 160) {
    ptx::mbar_init(&qkv_full[0], 1); ptx::mbar_init(&qkv_full[1], 1);
    ptx::mbar_init(&qkv_empty[0], 1); ptx::mbar_init(&qkv_empty[1], 1);
    ptx::mbar_init(s_full, 1); ptx::mbar_init(p_full, 160); ptx::mbar_init(o_full, 1);
    ptx::mbar_init(bias_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 5) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 5) {
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&tmQKV);
      ptx::mbar_expect_tx(bias_bar, AT_BIAS_BYTES);
      ptx::bulk_load(sBias, p.bias16 + (size_t)head * 144 * AT_BIAS_LD, AT_BIAS_BYTES, bias_bar);
      auto load_unit = [&](int win, int s) {
        uint8_t* st = sQKV + s * AT_STAGE_BYTES;
        ptx::mbar_expect_tx(&qkv_full[s], AT_STAGE_BYTES);
        ptx::tma_load_2d(st, &tmQKV, &qkv_full[s], head * 32, win * 144);
        ptx::tma_load_2d(st + AT_TILE_BYTES, &tmQKV, &qkv_full[s], p.C + head * 32, win * 144);
        ptx::tma_load_2d(st + 2 * AT_TILE_BYTES, &tmQKV, &qkv_full[s], 2 * p.C + head * 32, win * 144);
      };
      const uint32_t idesc_s = ptx::make_idesc_16(128, 144, 0, 0, p.dt == BF16 ? 1u : 0u);   // S = Q K^T : both K-major
      const uint32_t idesc_o = ptx::make_idesc_16(128, 32, 0, 1, p.dt == BF16 ? 1u : 0u);    // O = P V   : V is MN-major
      if (w_first < p.n_windows) load_unit(w_first, 0);
      uint32_t full_ph[2] = {0, 0}, empty_ph[2] = {0, 0}, pf_ph = 0;
      int i = 0;
      for (int win = w_first; win < p.n_windows; win += w_step, ++i) {
        const int s = i & 1;
        const uint32_t q_addr = ptx::smem_u32(sQKV + s * AT_STAGE_BYTES);
        const uint32_t k_addr = q_addr + AT_TILE_BYTES, v_addr = q_addr + 2 * AT_TILE_BYTES;
        ptx::mbar_wait(&qkv_full[s], full_ph[s]); full_ph[s] ^= 1;
        ptx::tc_fence_after();
        // S tiles: K-major, 64B swizzle: 8-row groups 512 B apart; +32 B per K=16 step
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const uint64_t a = ptx::make_smem_desc(q_addr + t * 128 * 64, 16, 512, ptx::SW_64B);
          const uint64_t b = ptx::make_smem_desc(k_addr, 16, 512, ptx::SW_64B);
#pragma unroll
          for (int k = 0; k < 2; ++k)
            ptx::umma_f16_ss(tmem_base + (t ? AT_COL_S1 : AT_COL_S0), a + 2 * k, b + 2 * k, idesc_s, k);
        }
        ptx::umma_commit(s_full);
        // prefetch the next window's tiles into the other stage
        if (win + w_step < p.n_windows) {
          if (i >= 1) { ptx::mbar_wait(&qkv_empty[s ^ 1], empty_ph[s ^ 1]); empty_ph[s ^ 1] ^= 1; }
          load_unit(win + w_step, s ^ 1);
        }
        // O tiles once P is in shared memory
        ptx::mbar_wait(p_full, pf_ph); pf_ph ^= 1;
        ptx::tc_fence_after();
        const uint32_t p_addr = ptx::smem_u32(sP);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
#pragma unroll
          for (int j = 0; j < 9; ++j) {
            // A = P: K-major, 32B swizzle, rows contiguous (8-row groups 256 B apart), one 4,608 B block per K step
            const uint64_t a = ptx::make_smem_desc(p_addr + j * AT_P_BLOCK + t * 128 * 32, 16, 256, ptx::SW_32B);
            // B = V: MN-major, 64B swizzle: 32 dims contiguous per key row, 8-key groups 512 B apart, 16 keys per step
            const uint64_t b = ptx::make_smem_desc(v_addr + j * 16 * 64, 16, 512, ptx::SW_64B);
            ptx::umma_f16_ss(tmem_base + (t ? AT_COL_O1 : AT_COL_O0), a, b, idesc_o, j);
          }
        }
        ptx::umma_commit(o_full);
        ptx::umma_commit(&qkv_empty[s]);
      }
    }
  } else {
    // ===== softmax + epilogue: warps 0-3 -> rows 0..127, warp 4 -> rows 128..143 =====
    const int tile = warp == 4 ? 1 : 0;
    const int r = tile ? 128 + lane : warp * 32 + lane;      // query row (>= 144 for the idle lanes of warp 4)
    const bool row_ok = r < 144;
    const int rr = row_ok ? r : 143;
    const uint32_t lane_base = tmem_base + ((uint32_t)((tile ? 0 : warp) * 32) << 16);
    const uint32_t s_col = tile ? AT_COL_S1 : AT_COL_S0, o_col = tile ? AT_COL_O1 : AT_COL_O0;
    const int qi = rr / 12, qj = rr % 12;
    const float LOG2E = 1.4426950408889634f;
    ptx::mbar_wait(bias_bar, 0);
    const uint8_t* brow = sBias + (size_t)rr * AT_BIAS_LD * 2;
    uint32_t ph = 0;
    const int nw = p.nwh * p.nww;
    for (int win = w_first; win < p.n_windows; win += w_step) {
      // analytic shift mask (src/swin.rs:603-655): only the last window row / column mixes regions
      const int wl = win % nw, wi = wl / p.nww, wj = wl - wi * p.nww;
      const bool last_r = p.shift > 0 && wi == p.nwh - 1, last_c = p.shift > 0 && wj == p.nww - 1;
      float mk[4];   // index = (ki>=6)*2 + (kj>=6)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const bool kih = c >> 1, kjh = c & 1;
        mk[c] = ((last_r && (kih != (qi >= 6))) || (last_c && (kjh != (qj >= 6)))) ? -100.0f : 0.0f;
      }
      ptx::mbar_wait(s_full, ph);
      ptx::tc_fence_after();
      // ---- pass 1: row max of s + bias + mask ----
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        uint32_t v[32];
        if (c < 4) ptx::tmem_ld32(lane_base + s_col + c * 32, v);
        else { uint32_t u[16]; ptx::tmem_ld16(lane_base + s_col + 128, u);
#pragma unroll
               for (int j = 0; j < 16; ++j) v[j] = u[j]; }
        ptx::tmem_ld_wait();
        const int ncol = c < 4 ? 32 : 16;
#pragma unroll
        for (int j8 = 0; j8 < ncol / 8; ++j8) {
          uint4 bq = *reinterpret_cast<const uint4*>(brow + (c * 32 + j8 * 8) * 2);
          const uint32_t* bh = reinterpret_cast<const uint32_t*>(&bq);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            float2 bf = unpack16x2(bh[t], p.dt);
            const int k0 = c * 32 + j8 * 8 + 2 * t, k1 = k0 + 1;
            float s0 = __uint_as_float(v[j8 * 8 + 2 * t]) + bf.x + mk[((k0 / 12 >= 6) ? 2 : 0) + ((k0 % 12 >= 6) ? 1 : 0)];
            float s1 = __uint_as_float(v[j8 * 8 + 2 * t + 1]) + bf.y + mk[((k1 / 12 >= 6) ? 2 : 0) + ((k1 % 12 >= 6) ? 1 : 0)];
            mx = fmaxf(mx, fmaxf(s0, s1));
          }
        }
      }
      const float moff = mx * LOG2E;
      // ---- pass 2: p = exp2((s - max) * log2e), row sum, bf16 P -> shared memory (K-major, 32B swizzle) ----
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        uint32_t v[32];
        if (c < 4) ptx::tmem_ld32(lane_base + s_col + c * 32, v);
        else { uint32_t u[16]; ptx::tmem_ld16(lane_base + s_col + 128, u);
#pragma unroll
               for (int j = 0; j < 16; ++j) v[j] = u[j]; }
        ptx::tmem_ld_wait();
        const int ncol = c < 4 ? 32 : 16;
#pragma unroll
        for (int j8 = 0; j8 < ncol / 8; ++j8) {
          uint4 bq = *reinterpret_cast<const uint4*>(brow + (c * 32 + j8 * 8) * 2);
          const uint32_t* bh = reinterpret_cast<const uint32_t*>(&bq);
          uint32_t packed[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            float2 bf = unpack16x2(bh[t], p.dt);
            const int k0 = c * 32 + j8 * 8 + 2 * t, k1 = k0 + 1;
            float s0 = __uint_as_float(v[j8 * 8 + 2 * t]) + bf.x + mk[((k0 / 12 >= 6) ? 2 : 0) + ((k0 % 12 >= 6) ? 1 : 0)];
            float s1 = __uint_as_float(v[j8 * 8 + 2 * t + 1]) + bf.y + mk[((k1 / 12 >= 6) ? 2 : 0) + ((k1 % 12 >= 6) ? 1 : 0)];
            float p0 = ex2(fmaf(s0, LOG2E, -moff)), p1 = ex2(fmaf(s1, LOG2E, -moff));
            sum += p0 + p1;
            packed[t] = pack16x2(p0, p1, p.dt);
          }
          if (row_ok) {
            // keys [8*g, 8*g+8): K step j16 = g/2, 16-byte chunk (g&1) of the row's 32 B, XOR-swizzled with row bit 2
            const int g = c * 4 + j8, j16 = g >> 1, ch = (g & 1) ^ ((r >> 2) & 1);
            *reinterpret_cast<uint4*>(sP + j16 * AT_P_BLOCK + r * 32 + ch * 16) =
                make_uint4(packed[0], packed[1], packed[2], packed[3]);
          }
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(p_full);
      // ---- epilogue: O / sum -> bf16, head-major channel (src/swin.rs:306-307) ----
      ptx::mbar_wait(o_full, ph);
      ptx::tc_fence_after();
      {
        uint32_t v[32];
        ptx::tmem_ld32(lane_base + o_col, v);
        ptx::tmem_ld_wait();
        if (row_ok) {
          const float inv = 1.f / sum;
          uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t)win * 144 + r) * p.ldo + head * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t w4[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              w4[t] = pack16x2(__uint_as_float(v[j * 8 + 2 * t]) * inv, __uint_as_float(v[j * 8 + 2 * t + 1]) * inv, p.dt);
            }
            dst[j] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
          }
        }
      }
      ph ^= 1;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 5) ptx::tmem_dealloc(tmem_base, 512);
}

void tc_attention(const LaunchCtx& ctx, const AttnArgs& a) {
  if (ctx.launches) ++*ctx.launches;
  if (ctx.dry) return;
  BRN_CHECK((a.qkv.dt == BF16 || a.qkv.dt == F16) && a.out.dt == a.qkv.dt, 5, "tc_attention: bf16 or fp16 only");
  BRN_CHECK(a.qkv.ld % 8 == 0 && a.out.ld % 8 == 0 && (((uintptr_t)a.qkv.p | (uintptr_t)a.out.p) & 15) == 0, 5,
            "tc_attention: alignment");
  AttnP p{};
  p.bias16 = (const uint16_t*)a.bias16; p.dt = a.qkv.dt; p.n_windows = a.n_windows; p.heads = a.heads; p.C = a.heads * 32;
  p.nwh = a.nwh; p.nww = a.nww; p.shift = a.shift;
  p.out = (uint16_t*)a.out.p; p.ldo = a.out.ld;
  const uint64_t rows = (uint64_t)a.n_windows * 144;
  uint64_t dims[2] = {(uint64_t)3 * p.C, rows};
  uint64_t str[1] = {(uint64_t)a.qkv.ld * 2};
  uint32_t box[2] = {32, 144};
  CUtensorMap tm = make_tmap_16(a.qkv.p, a.qkv.dt, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B);
  cudaFuncSetAttribute(tc_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
  const int sms = device_sm_count();
  int per_head = std::max(1, sms / a.heads);
  per_head = std::min(per_head, a.n_windows);
  char desc[96] = "";
  if (ctx.kt) snprintf(desc, sizeof desc, "windows=%d heads=%d shift=%d grid=%d", a.n_windows, a.heads, a.shift, a.heads * per_head);
  KScope ks(ctx, KC_ATTN_TC, 4.0 * 144 * 144 * 32 * (double)a.n_windows * a.heads, 0, desc);
  tc_attn_kernel<<<a.heads * per_head, AT_THREADS, AT_SMEM, ctx.stream>>>(tm, p);
  BRN_CUDA(cudaGetLastError());
}

}  // namespace brn
