// Fused shifted-window attention for 12x12 windows, head_dim 32 (sm_100a, tcgen05 + TMEM + TMA).
//
// Replaces WindowAttention::forward_standard (src/swin.rs:266-311) == the Metal flash_attention_with_[repeating_]bias
// calls (src/swin.rs:243,252): per (window, head)   O = softmax(q k^T + bias[h] + mask) v   with q pre-scaled, the
// relative-position bias of WindowAttention::new (src/swin.rs:143-152) and the analytic -100 region mask of
// BasicLayer::create_attention_mask (src/swin.rs:603-655).  The [nW,heads,144,144] score tensor never leaves the SM.
//
// Work unit = (window, head).  A persistent CTA owns ONE head (its 144x144 fp32 bias stays resident in shared memory,
// rows padded to 592 B so row-per-thread 16-byte reads are bank-conflict free) and walks the windows.
//   warp 10  : one elected thread issues TMA (Q,K,V tiles of the window-ordered qkv matrix, 64B-swizzled, 3 stages)
//              and all tcgen05.mma:  S = Q K^T (M=128 tiles over the 144 queries, N=144, K=32) into TMEM, then
//              O = P V (M=128 x2, N=32, K=144; V is the MN-major B operand straight from the TMA tile).
//   warps 0-14: softmax.  Every query row is split between THREE threads (keys 0-47 / 48-95 / 96-143), so a thread
//              keeps its 48 scores in registers: S is read from TMEM once and released immediately, which lets the MMA
//              warp compute S of the NEXT window while this one is in its exp phase.  warps 0-3 / 4-7 / 8-11: rows
//              0-127 (TMEM lane quadrant = warp % 4), warps 12 / 13 / 14: rows 128-143.  The kernel is bound by the
//              dependent chain of a softmax thread (TMEM load -> bias -> max -> exchange -> exp -> pack -> store), not
//              by issue slots or the MUFU pipe: 15 warps of 48 scores instead of 10 of 72 shorten the chain by a third
//              and put four warps on every SMSP.  Row max and row sum are combined through shared memory (a named
//              barrier per warp trio); the -100 shift mask is applied only in the border windows; P (bf16/fp16) goes
//              to shared memory in the 32B-swizzled K-major layout the P V MMA reads; 1/sum is applied to O in the
//              epilogue, which is deferred into the next unit's softmax so the P V latency is hidden.
//   TMEM   : S0 [0,144) rows 0-127 | S1 [144,288): rows 128-143 against key third t in columns [144+48t, 192+48t),
//            lanes 0-15 of lane quadrant t (A tile started at query row 128-32t; for warp 12+t) | O0 [288,320) |
//            O1 [320,352).
// Bound: MUFU (144*144 exp per unit = 1296 clk at 16/clk/SM) vs ~650 clk of tensor work per unit.
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

constexpr int AT_SOFT_WARPS = 15;                        // 12 for query rows 0-127 (3 per TMEM lane quadrant) + 3 for rows 128-143
constexpr int AT_SOFT_THREADS = 32 * AT_SOFT_WARPS;      // 480
constexpr int AT_THREADS = AT_SOFT_THREADS + 32;         // + control warp
constexpr int AT_BIAS_LD = 148;                          // fp32 elements per bias row (592 B: conflict-free 16-byte row reads)
constexpr int AT_BIAS_BYTES = 144 * AT_BIAS_LD * 4;      // 85,248
constexpr int AT_BIAS_REGION = 84 * 1024;
constexpr int AT_TILE_BYTES = 144 * 64;                  // one of Q/K/V: 9,216
constexpr int AT_STAGE_BYTES = 3 * AT_TILE_BYTES;        // 27,648
constexpr int AT_STAGES = 3;
constexpr int AT_P_BLOCK = 144 * 32;                     // one K=16 step of P: 4,608
constexpr int AT_P_REGION = 44 * 1024;                   // 9 blocks (41,472) + over-read slack of the 16-row tile
constexpr int AT_STAT_BYTES = 2 * 2 * 3 * 144 * 4;       // {max,sum} x parity x third x row
constexpr int AT_B16_BYTES = 3 * 64;                     // this head's q / k / v bias rows (the qkv of a pad token)
constexpr int AT_SMEM = AT_BIAS_REGION + AT_STAGES * AT_STAGE_BYTES + AT_P_REGION + AT_STAT_BYTES + AT_B16_BYTES + 256 + 1024;
constexpr uint32_t AT_COL_S0 = 0, AT_COL_S1 = 144, AT_COL_O0 = 288, AT_COL_O1 = 320;

struct AttnP {
  const float* bias32p;          // [heads][144][148] fp32
  int dt;                        // operand / output element type: BF16 or F16
  int n_windows, heads, C;
  int nwh, nww, shift;
  int split_win, nwh2, nww2;     // windows >= split_win (if > 0) belong to a second grid (merged half-resolution pass)
  uint16_t* out; int ldo;
  // token geometry: h > 0 enables the pad-row fix-up (q = k = v = bias for window rows in the pad region) and, with
  // token_out, the store straight to token order (window_reverse + roll back + crop, src/swin.rs:387-401)
  int h, w, h2, w2;
  int token_out; long long tok2;
  const uint16_t* bias16;        // [3C] qkv bias in the operand type
  FastDiv fd_nw1, fd_nww1, fd_nw2, fd_nww2;   // windows per image / per window row, both grids
};

// padded-grid coordinate of window-local index t (0..11) of window index wi: (12 wi + t + shift) mod hp
__device__ __forceinline__ int at_src_coord(int wi, int t, int shift, int hp) {
  int r = wi * 12 + t + shift;
  return r >= hp ? r - hp : r;
}

// (grid, image, window row, window column) of a window index; the divisions are multiply-high (FastDiv)
struct AtGeo {
  int g2, b, wi, wj;
};
__device__ __forceinline__ AtGeo at_geo(const AttnP& p, int win) {
  AtGeo g;
  g.g2 = (p.split_win > 0 && win >= p.split_win) ? 1 : 0;
  if (g.g2) win -= p.split_win;
  const FastDiv& fnw = g.g2 ? p.fd_nw2 : p.fd_nw1;
  const FastDiv& fnww = g.g2 ? p.fd_nww2 : p.fd_nww1;
  g.b = (int)fnw.div((uint32_t)win);
  const int wl = win - g.b * (int)fnw.d;
  g.wi = (int)fnww.div((uint32_t)wl);
  g.wj = wl - g.wi * (int)fnww.d;
  return g;
}
// does the window hold any pad position of its [h, w] token grid?  Its rows cover padded coordinates
// [12 wi + shift, 12 wi + shift + 12) mod hp; the pad region is [h, hp).
__device__ __forceinline__ bool at_has_pad(const AttnP& p, const AtGeo& g) {
  const int nwh = g.g2 ? p.nwh2 : p.nwh, nww = g.g2 ? p.nww2 : p.nww, h = g.g2 ? p.h2 : p.h, w = g.g2 ? p.w2 : p.w;
  const int hp = nwh * 12, wp = nww * 12;
  const int a = g.wi * 12 + p.shift, c = g.wj * 12 + p.shift;
  const bool rp = hp > h && (a + 12 > hp || a + 11 >= h);
  const bool cp = wp > w && (c + 12 > wp || c + 11 >= w);
  return rp || cp;
}
// Everything a softmax thread needs to know about one unit, computed once per unit with the grid's constants taken
// straight from the parameter bank (one warp-uniform branch on the grid instead of a select per constant):
//   orow  : output row of this thread's query row (token row with token_out, else window row; -1: pad position)
//   flags : 1 = last window row of a shifted block, 2 = last window column (the analytic mask applies), 4 = the window
//           holds pad positions (its staged tiles need the bias rows)
struct AtUnit {
  int orow;
  uint32_t flags;
};
__device__ __forceinline__ AtUnit at_unit_grid(const AttnP& p, int win_abs, int win, int r, int qi, int qj, int h, int w,
                                               int nwh, int nww, long long t0, const FastDiv& fnw, const FastDiv& fnww) {
  const int b = (int)fnw.div((uint32_t)win);
  const int wl = win - b * (int)fnw.d;
  const int wi = (int)fnww.div((uint32_t)wl), wj = wl - wi * (int)fnww.d;
  const int hp = nwh * 12, wp = nww * 12;
  const int a = wi * 12 + p.shift, c = wj * 12 + p.shift;
  AtUnit u;
  u.flags = (p.shift > 0 && wi == nwh - 1 ? 1u : 0u) | (p.shift > 0 && wj == nww - 1 ? 2u : 0u) |
            (((hp > h && (a + 12 > hp || a + 11 >= h)) || (wp > w && (c + 12 > wp || c + 11 >= w))) ? 4u : 0u);
  if (!p.token_out) {
    u.orow = win_abs * 144 + r;
  } else {
    int pr = a + qi, pc = c + qj;
    if (pr >= hp) pr -= hp;
    if (pc >= wp) pc -= wp;
    u.orow = (pr < h && pc < w) ? (int)t0 + (b * h + pr) * w + pc : -1;
  }
  return u;
}
__device__ __forceinline__ AtUnit at_unit(const AttnP& p, int win, int r, int qi, int qj) {
  if (p.split_win > 0 && win >= p.split_win)
    return at_unit_grid(p, win, win - p.split_win, r, qi, qj, p.h2, p.w2, p.nwh2, p.nww2, p.tok2, p.fd_nw2, p.fd_nww2);
  return at_unit_grid(p, win, win, r, qi, qj, p.h, p.w, p.nwh, p.nww, 0, p.fd_nw1, p.fd_nww1);
}
__device__ __forceinline__ bool at_row_is_pad(const AttnP& p, const AtGeo& g, int ti, int tj) {
  const int nwh = g.g2 ? p.nwh2 : p.nwh, nww = g.g2 ? p.nww2 : p.nww, h = g.g2 ? p.h2 : p.h, w = g.g2 ? p.w2 : p.w;
  return at_src_coord(g.wi, ti, p.shift, nwh * 12) >= h || at_src_coord(g.wj, tj, p.shift, nww * 12) >= w;
}



__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// predicated 16-byte shared store (no branch around it)
__device__ __forceinline__ void sts128_pred(uint32_t addr, const uint4& v, bool pred) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %5, 0;\n"
      "@p st.shared.v4.b32 [%0], {%1,%2,%3,%4};\n"
      "}\n" ::"r"(addr),
      "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"((uint32_t)pred)
      : "memory");
}

// three-input maximum (one FMNMX3 on sm_100)
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
// tcgen05.wait::ld tied to the destination registers (keeps uses below the wait)
__device__ __forceinline__ void tmem_wait32(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                 "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                 "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_wait8(uint32_t (&v)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7])
               :
               : "memory");
}
__device__ __forceinline__ void soft_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(AT_SOFT_THREADS) : "memory"); }
// the two warps that share a set of query rows (key halves 0 / 1) exchange their partial row max through shared memory
__device__ __forceinline__ void pair_bar_sync(int pair) { asm volatile("bar.sync %0, 96;" ::"r"(2 + pair) : "memory"); }

// exp2 on the FMA pipe for a pair of non-positive arguments (Cody-Waite split + degree-3 minimax on [-0.5, 0.5], max
// relative error 7.5e-5 -- the 16-bit rounding of P is 4.9e-4 / 3.9e-3): the kernel's binding pipe is XU (48 MUFU.EX2 +
// 24 F2FP per thread and unit at 4 lanes/clk/SMSP), so AT_POLY_PAIRS of every 4 score pairs take this path instead.
//   t = a + 1.5*2^23 (round to nearest integer n in the low mantissa bits), f = a - n, p = poly(f), bits += n << 23
#ifndef BRN_ATTN_POLY_PAIRS
#define BRN_ATTN_POLY_PAIRS 1      // A/B on B200 (r02 run D): 0 -> 62.0, 1 -> 65.8, 2 -> 63.2, 3 -> 58.2 units/us (7744 windows x 6 heads)
#endif
constexpr int AT_POLY_PAIRS = BRN_ATTN_POLY_PAIRS;
__device__ __forceinline__ void ex2_poly2(unsigned long long a2, float& p0, float& p1) {
  float a0, a1;
  upk2(a2, a0, a1);
  a0 = fmaxf(a0, -125.f); a1 = fmaxf(a1, -125.f);          // masked scores (-100 -> -144 in base 2): clamp, result ~ 2^-125
  const unsigned long long ac = pk2(a0, a1);
  const unsigned long long magic = pk2(12582912.f, 12582912.f), nmagic = pk2(-12582912.f, -12582912.f);
  const unsigned long long t2 = fadd2(ac, magic);
  const unsigned long long n2 = fadd2(t2, nmagic);
  const unsigned long long f2 = fma2(n2, pk2(-1.f, -1.f), ac);
  unsigned long long q = fma2(f2, pk2(0.05517164617776871f, 0.05517164617776871f), pk2(0.2426111251115799f, 0.2426111251115799f));
  q = fma2(f2, q, pk2(0.6932609677314758f, 0.6932609677314758f));
  q = fma2(f2, q, pk2(0.9999280571937561f, 0.9999280571937561f));
  float q0, q1, t0, t1;
  upk2(q, q0, q1);
  upk2(t2, t0, t1);
  // (t_bits << 23) == n * 2^23 mod 2^32 (the magic constant's low 9 bits are zero): one IMAD per element
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

template <int DT>
__device__ __forceinline__ uint32_t at_pack(float a, float b) { return DT == BF16 ? pack_bf16x2(a, b) : pack_f16x2(a, b); }
template <int DT>
__global__ void __launch_bounds__(AT_THREADS, 1)
tc_attn_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnP p) {
  ptx::pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sBias = smem;
  uint8_t* sQKV = smem + AT_BIAS_REGION;
  uint8_t* sP = sQKV + AT_STAGES * AT_STAGE_BYTES;
  float* sStat = (float*)(sP + AT_P_REGION);                 // smax[par][third][row], then ssum[par][third][row]
  uint8_t* sB16 = (uint8_t*)sStat + AT_STAT_BYTES;           // [3][64 B]
  uint64_t* bars = (uint64_t*)(sB16 + AT_B16_BYTES);
  uint64_t* qkv_full = bars;        // [3]
  uint64_t* qkv_empty = bars + 3;   // [3]
  uint64_t* s_full = bars + 6;
  uint64_t* s_empty = bars + 7;
  uint64_t* p_full = bars + 8;
  uint64_t* o_full = bars + 9;
  uint64_t* bias_bar = bars + 10;
  uint32_t* tmem_slot = (uint32_t*)(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.x % p.heads;
  const int w_first = blockIdx.x / p.heads, w_step = gridDim.x / p.heads;
  const int n_units = w_first < p.n_windows ? (p.n_windows - w_first + w_step - 1) / w_step : 0;

  if (threadIdx.x == AT_SOFT_THREADS) {
    for (int s = 0; s < AT_STAGES; ++s) { ptx::mbar_init(&qkv_full[s], 1); ptx::mbar_init(&qkv_empty[s], 1); }
    ptx::mbar_init(s_full, 1); ptx::mbar_init(s_empty, AT_SOFT_THREADS);
    ptx::mbar_init(p_full, AT_SOFT_THREADS); ptx::mbar_init(o_full, 1);
    ptx::mbar_init(bias_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == AT_SOFT_WARPS) ptx::tmem_alloc(tmem_slot, 512);
  if (warp == 0 && lane < 12 && p.h > 0)
    reinterpret_cast<uint4*>(sB16)[lane] =
        __ldg(reinterpret_cast<const uint4*>(p.bias16 + (size_t)(lane >> 2) * p.C + head * 32) + (lane & 3));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_wait();       // above: barriers, TMEM, the head's qkv bias (a weight)

  if (warp == AT_SOFT_WARPS) {
    if (n_units > 0) {
      // ===== control warp: lane 0 issues TMA and every tcgen05.mma; all 32 lanes patch pad rows of staged tiles =====
      const bool leader = lane == 0;
      const uint32_t is_bf = DT == BF16 ? 1u : 0u;
      const uint32_t idesc_s = ptx::make_idesc_16(128, 144, 0, 0, is_bf);   // S = Q K^T : both K-major
      const uint32_t idesc_s1 = ptx::make_idesc_16(128, 48, 0, 0, is_bf);   // rows 128-143 against one key third
      const uint32_t idesc_o = ptx::make_idesc_16(128, 32, 0, 1, is_bf);    // O = P V   : V is MN-major
      if (leader) {
        ptx::prefetch_tmap(&tmQKV);
        ptx::mbar_expect_tx(bias_bar, AT_BIAS_BYTES);
        ptx::bulk_load(sBias, p.bias32p + (size_t)head * 144 * AT_BIAS_LD, AT_BIAS_BYTES, bias_bar);
      }
      // this lane's 16-byte chunk (lane & 3) of the head's q / k / v bias rows: the qkv of a pad token
      uint4 bq[3];
      if (p.h > 0) {
#pragma unroll
        for (int t = 0; t < 3; ++t) bq[t] = ptx::lds128(ptx::smem_u32(sB16) + t * 64 + (lane & 3) * 16);
      }
      auto load_unit = [&](int i) {
        const int s = i % AT_STAGES, win = w_first + i * w_step;
        uint8_t* st = sQKV + s * AT_STAGE_BYTES;
        ptx::mbar_expect_tx(&qkv_full[s], AT_STAGE_BYTES);
        ptx::tma_load_2d(st, &tmQKV, &qkv_full[s], head * 32, win * 144);
        ptx::tma_load_2d(st + AT_TILE_BYTES, &tmQKV, &qkv_full[s], p.C + head * 32, win * 144);
        ptx::tma_load_2d(st + 2 * AT_TILE_BYTES, &tmQKV, &qkv_full[s], 2 * p.C + head * 32, win * 144);
      };
      // Pad tokens are zeros AFTER norm1 (src/swin.rs:355-366), so their q, k, v equal the qkv bias and they take part
      // as keys.  The token-order qkv GEMM never computes those rows: they are written into the staged tiles (64B
      // swizzle: 16-byte chunk c of row r lives at chunk c ^ ((r >> 1) & 3)) -- for unit 0 by this warp, for every later
      // unit by the softmax warps just before they release S (480 threads: a few stores each; this warp alone would
      // need a whole unit time for the loop and stall the MMA issue).
      auto fix_pads0 = [&]() {
        if (p.h <= 0) return;
        const AtGeo g = at_geo(p, w_first);
        if (!at_has_pad(p, g)) return;
        const uint32_t base = ptx::smem_u32(sQKV);
        const int ch = lane & 3;
        for (int r = lane >> 2; r < 144; r += 8) {
          const int ti = r / 12, tj = r - ti * 12;
          if (at_row_is_pad(p, g, ti, tj)) {
            const uint32_t a = base + r * 64 + ((ch ^ ((r >> 1) & 3)) << 4);
#pragma unroll
            for (int t = 0; t < 3; ++t) ptx::sts128(a + t * AT_TILE_BYTES, bq[t]);
          }
        }
        ptx::fence_proxy_async_smem();
      };
      auto issue_s = [&](int i) {
        const uint32_t q_addr = ptx::smem_u32(sQKV + (i % AT_STAGES) * AT_STAGE_BYTES), k_addr = q_addr + AT_TILE_BYTES;
        // K-major, 64B swizzle: 8-row groups 512 B apart; +32 B per K=16 step.
        // S0: query rows 0-127 x all 144 keys.  S1 (rows 128-143), once per key third t: the A tile starts at query
        // row 128 - 32 t, so the 16 rows land in lanes 0-15 of TMEM lane quadrant t (the quadrant of softmax warp
        // 12 + t), against keys [48 t, 48 t + 48) only (N = 48).
        const uint64_t b = ptx::make_smem_desc(k_addr, 16, 512, ptx::SW_64B);
        {
          const uint64_t a = ptx::make_smem_desc(q_addr, 16, 512, ptx::SW_64B);
#pragma unroll
          for (int k = 0; k < 2; ++k) ptx::umma_f16_ss(tmem_base + AT_COL_S0, a + 2 * k, b + 2 * k, idesc_s, k);
        }
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const uint64_t a = ptx::make_smem_desc(q_addr + (128 - 32 * t) * 64, 16, 512, ptx::SW_64B);
          const uint64_t bt = ptx::make_smem_desc(k_addr + t * 48 * 64, 16, 512, ptx::SW_64B);
#pragma unroll
          for (int k = 0; k < 2; ++k) ptx::umma_f16_ss(tmem_base + AT_COL_S1 + 48 * t, a + 2 * k, bt + 2 * k, idesc_s1, k);
        }
        ptx::umma_commit(s_full);
      };
      auto issue_pv = [&](int i) {
        const int s = i % AT_STAGES;
        const uint32_t v_addr = ptx::smem_u32(sQKV + s * AT_STAGE_BYTES) + 2 * AT_TILE_BYTES;
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
      };
      if (leader) {
        load_unit(0);
        if (n_units > 1) load_unit(1);
      }
      ptx::mbar_wait(&qkv_full[0], 0);
      fix_pads0();
      __syncwarp();
      if (leader) {
        ptx::tc_fence_after();
        issue_s(0);
        for (int i = 0; i < n_units; ++i) {
          if (i + 1 < n_units) {      // S of the next unit as soon as this unit's scores sit in registers (and its pad rows are patched)
            ptx::mbar_wait_backoff(&qkv_full[(i + 1) % AT_STAGES], ((i + 1) / AT_STAGES) & 1);
            ptx::mbar_wait_backoff(s_empty, i & 1);
            ptx::tc_fence_after();
            issue_s(i + 1);
          }
          if (i + 2 < n_units) {      // prefetch two units ahead; that stage held unit i-1
            if (i >= 1) ptx::mbar_wait_backoff(&qkv_empty[(i + 2) % AT_STAGES], ((i - 1) / AT_STAGES) & 1);
            load_unit(i + 2);
          }
          ptx::mbar_wait_backoff(p_full, i & 1);
          ptx::tc_fence_after();
          issue_pv(i);
        }
      }
      __syncwarp();
    }
  } else if (n_units > 0) {
    // ===== softmax + epilogue =====
    const int tile = warp >= 12 ? 1 : 0;
    const int third = tile ? warp - 12 : warp >> 2;            // keys [48*third, 48*third + 48)
    const int quad = warp & 3;
    const int pair = tile ? 4 : quad;                          // warps (q, q+4, q+8) and (12, 13, 14) share query rows
    const int r = tile ? 128 + lane : quad * 32 + lane;        // query row (>= 144 for the idle lanes of warps 12-14)
    const bool row_ok = r < 144;
    const int rr = row_ok ? r : 143;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t s_col = (tile ? AT_COL_S1 : AT_COL_S0) + third * 48;
    const int qi = rr / 12, qj = rr % 12;
    const float LOG2E = 1.4426950408889634f;
    // all shared-memory traffic of this role goes through explicit shared-space instructions
    const uint32_t smax = ptx::smem_u32(sStat);                // [par][third][144]
    const uint32_t ssum = smax + 2 * 3 * 144 * 4;
    const uint32_t sP_a = ptx::smem_u32(sP);
    ptx::mbar_wait(bias_bar, 0);
    const uint32_t brow = ptx::smem_u32(sBias) + rr * AT_BIAS_LD * 4 + third * 192;

    // output row of this thread's query row for a unit: the window-ordered row, or with token_out the token row
    // (-1: pad position, nothing is stored)
    // pad rows of the NEXT unit's staged q / k / v tiles (see fix_pads0): thread t < 432 owns row t % 144 of tile t / 144
    // Only the K and V tiles are patched (a pad QUERY row only produces an output row that is never stored), by
    // threads 0-287 (row t % 144 of tile 1 + t / 144); only the threads that actually write wait for the TMA and fence.
    auto fix_pads_next = [&](int i1, const AtGeo& g) {
      const int t = threadIdx.x;
      if (t < 288) {
        const int tq = 1 + t / 144, row = t - (tq - 1) * 144;
        const int ti = row / 12, tj = row - ti * 12;
        if (at_row_is_pad(p, g, ti, tj)) {
          ptx::mbar_wait(&qkv_full[i1 % AT_STAGES], (i1 / AT_STAGES) & 1);
          const uint32_t bsrc = ptx::smem_u32(sB16) + tq * 64;
          const uint32_t a = ptx::smem_u32(sQKV + (i1 % AT_STAGES) * AT_STAGE_BYTES) + tq * AT_TILE_BYTES + row * 64;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) ptx::sts128(a + ((ch ^ ((row >> 1) & 3)) << 4), ptx::lds128(bsrc + ch * 16));
          ptx::fence_proxy_async_smem();
        }
      }
    };
    AtUnit cur = at_unit(p, w_first, rr, qi, qj);
    uint32_t pbase[2];
    pbase[0] = sP_a + 3 * third * AT_P_BLOCK + rr * 32 + ((rr >> 2) & 1) * 16;
    pbase[1] = sP_a + 3 * third * AT_P_BLOCK + rr * 32 + (((rr >> 2) & 1) ^ 1) * 16;
    int orow_prev = -1;             // output row of the unit whose epilogue is still pending

    auto epilogue = [&](int j) {   // O(j) / sum(j) -> 16-bit, head-major channel (src/swin.rs:306-307)
      const int par = j & 1;
      ptx::mbar_wait(o_full, par);
      ptx::tc_fence_after();
      const float inv = 1.f / (ptx::lds32(ssum + ((par * 3 + 0) * 144 + rr) * 4) + ptx::lds32(ssum + ((par * 3 + 1) * 144 + rr) * 4) +
                               ptx::lds32(ssum + ((par * 3 + 2) * 144 + rr) * 4));
      const unsigned long long inv2 = pk2(inv, inv);
      auto pack_scaled = [&](uint32_t a, uint32_t b) {
        float x, y;
        upk2(fmul2(pk2(__uint_as_float(a), __uint_as_float(b)), inv2), x, y);
        return at_pack<DT>(x, y);
      };
      if (!tile) {
        if (third == 2) return;       // the 32 output dims of a row are written by its first two warps, 16 each
        uint32_t v[16];
        ptx::tmem_ld16(lane_base + AT_COL_O0 + third * 16, v);
        tmem_wait_dep(v);
        uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t)orow_prev * p.ldo + head * 32 + third * 16);
        if (orow_prev >= 0)
#pragma unroll
        for (int g = 0; g < 2; ++g)
          dst[g] = make_uint4(pack_scaled(v[8 * g], v[8 * g + 1]),
                              pack_scaled(v[8 * g + 2], v[8 * g + 3]),
                              pack_scaled(v[8 * g + 4], v[8 * g + 5]),
                              pack_scaled(v[8 * g + 6], v[8 * g + 7]));
      } else if (third == 0) {      // rows 128-143 live in lanes 0-15 of quadrant 0: warp 12 writes all 32 dims
        uint32_t v[32];
        ptx::tmem_ld32(lane_base + AT_COL_O1, v);
        tmem_wait32(v);
        if (row_ok && orow_prev >= 0) {
          uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t)orow_prev * p.ldo + head * 32);
#pragma unroll
          for (int g = 0; g < 4; ++g)
            dst[g] = make_uint4(pack_scaled(v[8 * g], v[8 * g + 1]),
                                pack_scaled(v[8 * g + 2], v[8 * g + 3]),
                                pack_scaled(v[8 * g + 4], v[8 * g + 5]),
                                pack_scaled(v[8 * g + 6], v[8 * g + 7]));
        }
      }
    };

    for (int i = 0; i < n_units; ++i) {
      const int par = i & 1, win = w_first + i * w_step;
      // analytic shift mask (src/swin.rs:603-655): only the last window row / column mixes regions.  Key c of this
      // third sits at ki = 4*third + c/12, kj = c%12; interior windows take the mask-free path.
      const bool last_r = (cur.flags & 1u) != 0, last_c = (cur.flags & 2u) != 0;

      // ---- scores: TMEM -> registers once, then hand the S region back to the MMA warp ----
      uint32_t v0[32], v1[16];
      ptx::mbar_wait(s_full, par);
      ptx::tc_fence_after();
      ptx::tmem_ld32(lane_base + s_col, v0);
      ptx::tmem_ld16(lane_base + s_col + 32, v1);
      tmem_wait32(v0); tmem_wait_dep(v1);
      // geometry of the next unit; its pad rows are patched before S is released (the MMA warp issues S(i+1) then)
      AtUnit nxt = cur;
      if (i + 1 < n_units) {
        nxt = at_unit(p, win + w_step, rr, qi, qj);
        if (p.h > 0 && (nxt.flags & 4u)) fix_pads_next(i + 1, at_geo(p, win + w_step));     // rare: border windows of padded grids
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(s_empty);

      // ---- pass 1: s + bias (+ mask) kept in registers as packed fp32 pairs (FADD2), partial row max ----
      // The masked and the plain version are two COMPLETE loops that both define sc[] (a mask applied in place after
      // the bias add made the compiler copy all 48 score registers in front of the branch: ~40 moves per warp and unit).
      unsigned long long sc[24];
      float mx = -INFINITY;
      auto sv = [&](int cc) { return __uint_as_float(cc < 32 ? v0[cc & 31] : v1[cc & 15]); };
      if (last_r || last_c) {      // warp-uniform: border windows of a shifted block only
#pragma unroll
        for (int kr = 0; kr < 4; ++kr) {                   // the four key rows of this third (3 bias chunks = 12 keys each)
          const bool rmask = last_r && ((4 * third + kr >= 6) != (qi >= 6));
          const float m0 = (rmask || (last_c && (qj >= 6))) ? -100.0f : 0.0f;    // kj < 6
          const float m1 = (rmask || (last_c && (qj < 6))) ? -100.0f : 0.0f;     // kj >= 6
          const unsigned long long mk0 = pk2(m0, m0), mk1 = pk2(m1, m1);
#pragma unroll
          for (int gg = 0; gg < 3; ++gg) {
            const int g = kr * 3 + gg, c = g * 4;
            const uint4 bq = ptx::lds128(brow + g * 16);
            const unsigned long long b0 = fadd2(pk2(__uint_as_float(bq.x), __uint_as_float(bq.y)), (2 * gg) >= 3 ? mk1 : mk0);
            const unsigned long long b1 = fadd2(pk2(__uint_as_float(bq.z), __uint_as_float(bq.w)), (2 * gg + 1) >= 3 ? mk1 : mk0);
            sc[2 * g] = fadd2(pk2(sv(c), sv(c + 1)), b0);
            sc[2 * g + 1] = fadd2(pk2(sv(c + 2), sv(c + 3)), b1);
          }
        }
      } else {
#pragma unroll
        for (int g = 0; g < 12; ++g) {
          const uint4 bq = ptx::lds128(brow + g * 16);       // 4 fp32 bias values
          const int c = g * 4;                               // key column within this third; 48 = 4 * 12 so kj = c % 12
          sc[2 * g] = fadd2(pk2(sv(c), sv(c + 1)), pk2(__uint_as_float(bq.x), __uint_as_float(bq.y)));
          sc[2 * g + 1] = fadd2(pk2(sv(c + 2), sv(c + 3)), pk2(__uint_as_float(bq.z), __uint_as_float(bq.w)));
        }
      }
#pragma unroll
      for (int c2 = 0; c2 < 24; ++c2) {
        float lo, hi;
        upk2(sc[c2], lo, hi);
        mx = fmax3(mx, lo, hi);
      }
      if (row_ok) ptx::sts32(smax + ((par * 3 + third) * 144 + r) * 4, mx);
      pair_bar_sync(pair);
      const float m = fmax3(ptx::lds32(smax + ((par * 3 + 0) * 144 + rr) * 4), ptx::lds32(smax + ((par * 3 + 1) * 144 + rr) * 4),
                            ptx::lds32(smax + ((par * 3 + 2) * 144 + rr) * 4));
      const float moff = m * LOG2E;

      // ---- deferred epilogue of the previous unit (its P V finished long ago); also frees the P buffer ----
      if (i > 0) epilogue(i - 1);
      orow_prev = row_ok ? cur.orow : -1;
      cur = nxt;

      // ---- pass 2: p = exp2((s - max) * log2e), partial row sum, 16-bit P -> shared (K-major, 32B swizzle) ----
      unsigned long long sum2 = pk2(0.f, 0.f);
      const unsigned long long l2e2 = pk2(LOG2E, LOG2E), noff2 = pk2(-moff, -moff);
#pragma unroll
      for (int g = 0; g < 6; ++g) {
        uint32_t packed[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const unsigned long long a2 = fma2(sc[g * 4 + t], l2e2, noff2);
          float p0, p1;
          if (t < AT_POLY_PAIRS) {
            ex2_poly2(a2, p0, p1);
          } else {
            float a0, a1;
            upk2(a2, a0, a1);
            p0 = ex2(a0); p1 = ex2(a1);
          }
          sum2 = fadd2(sum2, pk2(p0, p1));
          packed[t] = at_pack<DT>(p0, p1);
        }
        // keys [8*G, 8*G+8), G = 6*third + g: K step G/2 = 3*third + g/2, 16-byte chunk (g&1) of the row's 32 B XOR row
        // bit 2 -- two per-thread bases (pbase[g&1]) plus a compile-time offset; predicated, no branch
        sts128_pred(pbase[g & 1] + (g >> 1) * AT_P_BLOCK, make_uint4(packed[0], packed[1], packed[2], packed[3]), row_ok);
      }
      float sum;
      {
        float s0, s1;
        upk2(sum2, s0, s1);
        sum = s0 + s1;
      }
      if (row_ok) ptx::sts32(ssum + ((par * 3 + third) * 144 + r) * 4, sum);
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(p_full);
    }
    soft_bar_sync();               // every thread's sum of the last unit is visible
    epilogue(n_units - 1);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == AT_SOFT_WARPS) ptx::tmem_dealloc(tmem_base, 512);
}

void tc_attention(const LaunchCtx& ctx, const AttnArgs& a) {
  if (ctx.launches) ++*ctx.launches;
  if (ctx.dry) return;
  BRN_CHECK((a.qkv.dt == BF16 || a.qkv.dt == F16) && a.out.dt == a.qkv.dt, 5, "tc_attention: bf16 or fp16 only");
  BRN_CHECK(a.qkv.ld % 8 == 0 && a.out.ld % 8 == 0 && (((uintptr_t)a.qkv.p | (uintptr_t)a.out.p) & 15) == 0, 5,
            "tc_attention: alignment");
  AttnP p{};
  BRN_CHECK(a.bias32p != nullptr, 1, "tc_attention: padded fp32 bias missing");
  p.bias32p = a.bias32p; p.dt = a.qkv.dt; p.n_windows = a.n_windows; p.heads = a.heads; p.C = a.heads * 32;
  p.nwh = a.nwh; p.nww = a.nww; p.shift = a.shift;
  p.split_win = a.split_win; p.nwh2 = a.nwh2; p.nww2 = a.nww2;
  p.out = (uint16_t*)a.out.p; p.ldo = a.out.ld;
  p.h = a.h; p.w = a.w; p.h2 = a.h2; p.w2 = a.w2; p.token_out = a.token_out; p.tok2 = a.tok2;
  p.bias16 = (const uint16_t*)a.qkv_bias16;
  p.fd_nw1 = FastDiv((uint32_t)(a.nwh * a.nww)); p.fd_nww1 = FastDiv((uint32_t)a.nww);
  if (a.split_win > 0) { p.fd_nw2 = FastDiv((uint32_t)(a.nwh2 * a.nww2)); p.fd_nww2 = FastDiv((uint32_t)a.nww2); }
  BRN_CHECK(a.h <= 0 || (a.qkv_bias16 && (((uintptr_t)a.qkv_bias16) & 15) == 0 && a.w > 0), 1,
            "tc_attention: token geometry needs the 16-bit qkv bias");
  BRN_CHECK(!a.token_out || a.h > 0, 1, "tc_attention: token-order output needs the token geometry");
  const uint64_t rows = (uint64_t)a.n_windows * 144;
  uint64_t dims[2] = {(uint64_t)3 * p.C, rows};
  uint64_t str[1] = {(uint64_t)a.qkv.ld * 2};
  uint32_t box[2] = {32, 144};
  CUtensorMap tm = make_tmap_16(a.qkv.p, a.qkv.dt, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B);
  if (a.qkv.dt == BF16) cudaFuncSetAttribute(tc_attn_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
  else cudaFuncSetAttribute(tc_attn_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
  const int sms = device_sm_count();
  int per_head = std::max(1, sms / a.heads);
  per_head = std::min(per_head, a.n_windows);
  char desc[96] = "";
  if (ctx.kt) snprintf(desc, sizeof desc, "windows=%d heads=%d shift=%d grid=%d", a.n_windows, a.heads, a.shift, a.heads * per_head);
  KScope ks(ctx, KC_ATTN_TC, 4.0 * 144 * 144 * 32 * (double)a.n_windows * a.heads,
            (double)a.n_windows * a.heads * 144 * 32 * 4 * dsize(a.qkv.dt), desc);   // q, k, v in + o out
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(a.heads * per_head);
  cfg.blockDim = dim3(AT_THREADS);
  cfg.dynamicSmemBytes = AT_SMEM;
  cfg.stream = ctx.stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr; cfg.numAttrs = pdl_attr(attr, 0);
  if (a.qkv.dt == BF16) BRN_CUDA(cudaLaunchKernelEx(&cfg, tc_attn_kernel<BF16>, tm, p));
  else BRN_CUDA(cudaLaunchKernelEx(&cfg, tc_attn_kernel<F16>, tm, p));
  BRN_CUDA(cudaGetLastError());
}

}  // namespace brn
