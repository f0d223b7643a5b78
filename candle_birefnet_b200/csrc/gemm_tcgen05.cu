// tcgen05 implicit-GEMM kernel (sm_100a): every linear layer and every stride-1 convolution of the model.
//
//   out[m, n] = act( sum_{tap, c} X[pixel(m) + tap, c] * W[n, tap, c] + bias ) (+ residual)
//
// * A operand: NHWC 16-bit (bf16 / fp16) activations through a 4-D TMA tensor map {C, W, H, B}; one M tile is a TH x TW pixel
//   patch (TH*TW = 128) of one image, so a conv tap is just a shifted box and TMA's out-of-bounds zero fill is the
//   conv's zero padding (and the channel tail when C % 64 != 0).  Linear layers are the 1-tap case on {K, rows,1,1}.
// * B operand: weights [N][taps][cin_pad] 16-bit (K-major) through a 2-D map, box {64, BN}; in 2-CTA clusters each
//   CTA loads half of the tile and multicasts it.
// * 128B-swizzled K-major smem tiles, 4-stage mbarrier ring, warp-specialised: warp 0 = TMA producer,
//   warp 1 = MMA issuer (one elected thread, tcgen05.mma cta_group::1 kind::f16, M=128, N=BN<=256, K=16),
//   warps 2-9 = epilogue (tc_epilogue.cuh: tcgen05.ld 32x32b, bias / ReLU / erf-GELU / 2*sigmoid / residual /
//   window-reverse row scatter, smem-staged coalesced 16-bit or fp32 stores).  Two TMEM accumulator stages (2 x 256
//   columns) overlap the epilogue of tile i with the MMAs of tile i+1; the kernel is persistent (grid = min(tiles, #SM))
//   and instantiated once per hot epilogue variant (EpiKind).
// Roofline: tensor pipe (2*M*N*K flop per launch) for K >= 384 without GELU; HBM / epilogue issue below that.
#include <cuda.h>

#include <cstdio>
#include <mutex>

#include "brn_common.h"
#include "device_utils.cuh"
#include "tc_ptx.cuh"
#include "tc_epilogue.cuh"

namespace brn {

constexpr int TC_BM = 128, TC_BK = 64, TC_STAGES = 4;
#ifndef BRN_EPI_WARPS
#define BRN_EPI_WARPS 8
#endif
constexpr int TC_EPI_WARPS = BRN_EPI_WARPS;                               // two per TMEM lane quadrant (contiguous column parts; 12 warps measured slower)
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;            // warp 0 TMA, warp 1 MMA, warps 2.. epilogue
constexpr int TC_A_BYTES = (TC_BM + 16) * TC_BK * 2;   // 18 KB: a 16x8-pixel tile plus one halo row above and below (tap groups)
constexpr int TC_B_BYTES = 256 * TC_BK * 2;     // 32 KB (BN <= 256)
constexpr int TC_BIAS_LD = 288;                                // floats per accumulator stage (BN <= 256, padded to 32)
constexpr int TC_EPI_BYTES = TC_EPI_WARPS * EPI_STAGE_BYTES + 4 * TC_BIAS_LD * 4;   // bias + LnFold column sums, x2 accumulator stages
constexpr int TC_SMEM = TC_STAGES * (TC_A_BYTES + TC_B_BYTES) + TC_EPI_BYTES + 256 + 1024;
#ifndef BRN_GEMM_U2_STAGES
#define BRN_GEMM_U2_STAGES 6
#endif
constexpr int TC_STAGES_U2 = BRN_GEMM_U2_STAGES;               // CTA-pair instances (half a B tile per CTA and stage)
constexpr int TC_SMEM_U2 = TC_STAGES_U2 * (TC_A_BYTES + TC_B_BYTES / 2) + TC_EPI_BYTES + 256 + 1024;
static_assert(TC_SMEM <= 232448 && TC_SMEM_U2 <= 232448, "shared memory per CTA");

struct TcGemmP {
  int B, H, W;
  int tw_log2, tiles_x, tiles_y;
  int m_tiles, n_tiles, BN;
  int taps, kw, pad, cblocks, cin_pad;
  int in_bf16;   // operand format: 1 = bf16, 0 = fp16
  int ksplit, kb_per;   // split-K (tiny M, long K): K blocks [ks*kb_per, (ks+1)*kb_per) per work item; fp32 partials
  long long rows_total; //   of split ks go to out rows [ks*rows_total, ...); a reduce kernel adds bias / activation
  int tg;        // 3x3 tap groups: one A load of (16+2) x 8 pixels serves the three vertical taps of a kernel column
  int out_tiled; // fp32 out is [m_tile][N][128]
  int tma_store; // 16-bit row-major token-matrix output without residual / row map: the epilogue stores through tmO
  int u2;        // CL = 2 only: the pair runs ONE tcgen05.mma.cta_group::2 (M = 256) per K step instead of two M = 128
                 // MMAs on multicast copies of B: each CTA keeps its 128 rows of A and HALF of the B tile, which cuts the
                 // shared-memory traffic per K block from 48 KB written + 48 KB read to 32 + 32 (the limiter of the
                 // K >= 768 GEMMs: 64 % tensor-pipe active with the multicast scheme, ncu r02)
  EpiP epi;
  RowMap rm;
  FastDiv fd_nt, fd_tpi, fd_tx;        // n_tiles, tiles_x * tiles_y, tiles_x
  FastDiv fd_ks;                       // ksplit
  FastDiv fd_rows1, fd_nww1;           // row map, first grid: rows per image (windows * 144), windows per row
  FastDiv fd_rows2, fd_nww2;           // second grid (merged two-resolution pass)
  FastDiv fd_hw1, fd_w1, fd_hw2, fd_w2;   // row map mode 2 (token -> window row): tokens per image, tokens per row
};

// rowmap_token (device_utils.cuh) with the run-time divisions replaced; rows < 2^31
__device__ __forceinline__ long long window_row_to_token_fd(uint32_t m, int h, int w, int hp, int wp, int shift,
                                                           const FastDiv& rows, const FastDiv& nww, uint32_t ws) {
  const uint32_t b = rows.div(m);
  const uint32_t rem = m - b * rows.d;
  uint32_t wid, t, ti, tj;
  if (ws == 12u) { wid = rem / 144u; t = rem - wid * 144u; ti = t / 12u; tj = t - ti * 12u; }     // constant divisors
  else { wid = rem / (ws * ws); t = rem - wid * ws * ws; ti = t / ws; tj = t - ti * ws; }
  const uint32_t wi = nww.div(wid), wj = wid - wi * nww.d;
  int r = (int)(wi * ws + ti) + shift, c = (int)(wj * ws + tj) + shift;
  if (r >= hp) r -= hp;
  if (c >= wp) c -= wp;
  if (r >= h || c >= w) return -1;
  return ((long long)b * h + r) * (long long)w + c;
}
// inverse map (RowMap mode 2): token row of the [B,h,w] grid -> its window-ordered padded row (pad -> roll(-shift) ->
// window_partition, src/swin.rs:359-380)
__device__ __forceinline__ long long token_to_window_row_fd(uint32_t m, int h, int w, int hp, int wp, int shift,
                                                           const FastDiv& hw, const FastDiv& fw) {
  const uint32_t b = hw.div(m);
  const uint32_t rem = m - b * hw.d;
  const uint32_t r = fw.div(rem), c = rem - r * fw.d;
  int pr = (int)r - shift, pc = (int)c - shift;
  if (pr < 0) pr += hp;
  if (pc < 0) pc += wp;
  const uint32_t wi = (uint32_t)pr / 12u, ti = (uint32_t)pr - wi * 12u;
  const uint32_t wj = (uint32_t)pc / 12u, tj = (uint32_t)pc - wj * 12u;
  const uint32_t nww = (uint32_t)wp / 12u, nw = ((uint32_t)hp / 12u) * nww;
  return (long long)(b * nw + wi * nww + wj) * 144 + ti * 12u + tj;
}
__device__ __forceinline__ long long rowmap_token_fd(const TcGemmP& p, long long m) {
  const RowMap& rm = p.rm;
  if (rm.enabled == 2) {
    if (rm.split > 0 && m >= rm.tok2)
      return rm.split + token_to_window_row_fd((uint32_t)(m - rm.tok2), rm.h2, rm.w2, rm.hp2, rm.wp2, rm.shift, p.fd_hw2, p.fd_w2);
    return token_to_window_row_fd((uint32_t)m, rm.h, rm.w, rm.hp, rm.wp, rm.shift, p.fd_hw1, p.fd_w1);
  }
  if (rm.split > 0 && m >= rm.split) {
    const long long t = window_row_to_token_fd((uint32_t)(m - rm.split), rm.h2, rm.w2, rm.hp2, rm.wp2, rm.shift,
                                               p.fd_rows2, p.fd_nww2, (uint32_t)rm.ws);
    return t < 0 ? t : t + rm.tok2;
  }
  return window_row_to_token_fd((uint32_t)m, rm.h, rm.w, rm.hp, rm.wp, rm.shift, p.fd_rows1, p.fd_nww1, (uint32_t)rm.ws);
}

// CL = CTAs per cluster (1 or 2).  With CL = 2 the two CTAs own M tiles 2j and 2j+1 of the same N tile: each loads
// half of the shared B (weight) tile and TMA-multicasts it into both CTAs' shared memory, so the L2 -> SMEM bytes per
// 128x256x64 MMA block drop from 48 KB to 32 KB (the kernel is L2-bandwidth bound at 48 KB: ~42 B/clk/SM of L2 vs
// 94 B/clk/SM needed to keep the tensor pipe busy).
// EPI = epilogue variant compiled into this instance (one kernel per hot epilogue instead of a run-time switch over all
// 18 variants inside one 27k-instruction kernel: per-variant register allocation, hot loop within the instruction cache).
enum EpiKind { EK_GENERIC = 0, EK_NONE16, EK_RELU16, EK_GELU16, EK_NONE32, EK_SIG32_TILED, EK_RES32, EK_RES16,
               EK_LNF_NONE16, EK_LNF_GELU16,      // folded LayerNorm consumers (qkv, fc1)
               EK_RES32_EMIT, EK_NONE32_EMIT };   // fp32 stream producers that also emit statistics + the raw 16-bit copy

// U2: CTA-pair instance -- every tcgen05 alloc / mma / commit / dealloc of the kernel carries cta_group::2 (a kernel
// must not mix the two CTA-group forms: the run-time switch of the first version worked in isolation and hung as soon
// as two forwards were in flight on two streams, r02 run H).
template <int CL, int EPI, bool U2 = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, const TcGemmP p) {
  ptx::pdl_launch_dependents();
  // CTA pairs keep HALF of the B tile per CTA: 34 KB per stage instead of 50 KB, so the ring is six stages deep (2,560
  // instead of 1,536 tensor-pipe cycles of look-ahead for the TMA round trip)
  constexpr int NST = U2 ? TC_STAGES_U2 : TC_STAGES;
  constexpr int BBYTES = U2 ? TC_B_BYTES / 2 : TC_B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + NST * TC_A_BYTES;
  uint8_t* sStage = sB + NST * BBYTES;                 // epilogue staging, 2 KB per epilogue warp
  float* sBias = (float*)(sStage + TC_EPI_WARPS * EPI_STAGE_BYTES);
  float* sCs = sBias + 2 * TC_BIAS_LD;                           // LnFold column sums, staged like the bias
  uint64_t* full = (uint64_t*)((uint8_t*)sBias + 4 * TC_BIAS_LD * 4);
  uint64_t* empty = full + NST;
  uint64_t* tfull = empty + NST;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = CL > 1 ? (int)ptx::cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    constexpr bool u2i = U2;
    for (int s = 0; s < NST; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], u2i ? 1 : CL); }
    // pair mode: one elected lane per epilogue warp of BOTH CTAs arrives on the leader's accumulator-empty barrier
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull[a], 1); ptx::mbar_init(&tempty[a], u2i ? 2 * TC_EPI_WARPS : 32 * TC_EPI_WARPS); }
    ptx::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { ptx::prefetch_tmap(&tmA); ptx::prefetch_tmap(&tmB); if (p.tma_store) ptx::prefetch_tmap(&tmO); }
  constexpr bool u2 = U2;
  static_assert(!U2 || CL == 2, "CTA pairs are 2-CTA clusters");
  if (warp == 1) { if (u2) ptx::tmem_alloc_2sm(tmem_slot, 512); else ptx::tmem_alloc(tmem_slot, 512); }
  ptx::tc_fence_before();
  __syncthreads();
  if (CL > 1) ptx::cluster_sync_all();     // peers' barriers are initialised before any multicast can land
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_wait();       // everything above touched only this CTA's shared / tensor memory and kernel parameters

  // work items: (m_group, n_tile); this CTA's M tile is m_group * CL + rank
  const int m_groups = (p.m_tiles + CL - 1) / CL;
  const int num_items = m_groups * p.n_tiles * p.ksplit;
  const int item0 = blockIdx.x / CL, item_step = gridDim.x / CL;
  const int kblocks = (p.tg ? 3 : p.taps) * p.cblocks;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int TW = 1 << p.tw_log2, TH = TC_BM >> p.tw_log2;

  if (warp == 0) {
    if (ptx::elect_one()) {
      // ===== TMA producer =====
      int stage = 0; uint32_t phase = 0;
      const uint32_t tx_bytes = p.tg ? TC_A_BYTES + 3 * p.BN * TC_BK * 2 : TC_BM * TC_BK * 2 + p.BN * TC_BK * 2;
      const int b_rows = p.BN / CL;                           // rows of B this CTA loads (and multicasts)
      for (int item = item0; item < num_items; item += item_step) {
        // work-item decode by multiply-high, K-block decode by counters: the producer sits on the refill path of every
        // stage (it wakes when a stage is released), so an integer division here is tensor-pipe idle time
        const int it2 = (int)p.fd_ks.div((uint32_t)item), ks = item - it2 * p.ksplit;
        const int kb0 = ks * p.kb_per, kb1 = min(kblocks, kb0 + p.kb_per);
        const int mg = (int)p.fd_nt.div((uint32_t)it2);
        const int m_tile = mg * CL + rank, n_tile = it2 - mg * p.n_tiles;
        const int b = (int)p.fd_tpi.div((uint32_t)m_tile), r = m_tile - b * tiles_per_img;   // b >= B for the odd tail: TMA zero-fills
        const int ty = (int)p.fd_tx.div((uint32_t)r);
        const int y0 = ty * TH, x0 = (r - ty * p.tiles_x) * TW;
        int tap = 0, cb = kb0, ky = 0, kx = 0;
        if (kb0 != 0) { tap = kb0 / p.cblocks; cb = kb0 - tap * p.cblocks; if (!p.tg) { ky = tap / p.kw; kx = tap - ky * p.kw; } }
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait_backoff(&empty[stage], phase ^ 1);
          if (!u2) ptx::mbar_expect_tx(&full[stage], tx_bytes);
          else if (rank == 0) ptx::mbar_expect_tx(&full[stage], 2 * (TC_BM * TC_BK * 2 + b_rows * TC_BK * 2));
          // (tap, cb, ky, kx) of this K block; `next_kb` advances them
          auto next_kb = [&] {
            if (++cb == p.cblocks) { cb = 0; ++tap; if (++kx == p.kw) { kx = 0; ++ky; } }
            if (++stage == NST) { stage = 0; phase ^= 1; }
          };
          if (p.tg) {
            // `tap` is the kernel column kx: the (TH+2) x TW box holds the input rows of all three vertical taps
            ptx::tma_load_4d(sA + stage * TC_A_BYTES, &tmA, &full[stage], cb * TC_BK, x0 + tap - 1, y0 - 1, b);
            const int bn = n_tile * p.BN + rank * b_rows;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              uint8_t* bdst = sB + stage * BBYTES + ky * p.BN * (TC_BK * 2) + rank * b_rows * (TC_BK * 2);
              const int bk = (ky * 3 + tap) * p.cin_pad + cb * TC_BK;
              if (CL > 1) ptx::tma_load_2d_mc(bdst, &tmB, &full[stage], bk, bn, (uint16_t)((1u << CL) - 1));
              else ptx::tma_load_2d(bdst, &tmB, &full[stage], bk, bn);
            }
            next_kb();
            continue;
          }
          if (u2) {
            // CTA pair: both CTAs' boxes are counted on the LEADER's barrier (it expects the bytes of both); B is not
            // multicast -- this CTA keeps rows [rank * BN/2, (rank + 1) * BN/2) of the tile at offset 0 of its B stage
            // (the expect_tx above went to this CTA's own barrier: only the leader's is ever waited on, so the leader
            // expects twice the per-CTA bytes and the follower's own barrier is simply not used in this mode)
            const uint32_t lbar = ptx::mapa_shared(ptx::smem_u32(&full[stage]), 0);
            ptx::tma_load_4d_2sm(sA + stage * TC_A_BYTES, &tmA, lbar, cb * TC_BK, x0 + kx - p.pad, y0 + ky - p.pad, b);
            ptx::tma_load_2d_2sm(sB + stage * BBYTES, &tmB, lbar, tap * p.cin_pad + cb * TC_BK, n_tile * p.BN + rank * b_rows);
            next_kb();
            continue;
          }
          ptx::tma_load_4d(sA + stage * TC_A_BYTES, &tmA, &full[stage], cb * TC_BK, x0 + kx - p.pad, y0 + ky - p.pad, b);
          uint8_t* bdst = sB + stage * BBYTES + rank * b_rows * (TC_BK * 2);
          const int bk = tap * p.cin_pad + cb * TC_BK, bn = n_tile * p.BN + rank * b_rows;
          if (CL > 1) ptx::tma_load_2d_mc(bdst, &tmB, &full[stage], bk, bn, (uint16_t)((1u << CL) - 1));
          else ptx::tma_load_2d(bdst, &tmB, &full[stage], bk, bn);
          next_kb();
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one() && !(u2 && rank != 0)) {      // pair mode: the leader CTA issues for both
      // ===== MMA issuer =====
      const uint32_t idesc = ptx::make_idesc_16(u2 ? 2 * TC_BM : TC_BM, p.BN, 0, 0, p.in_bf16);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
#ifdef BRN_GEMM_TIMING
      long long tm_empty = 0, tm_full = 0, tm_t0 = clock64(), tm_a;
#endif
      for (int item = item0; item < num_items; item += item_step) {
#ifdef BRN_GEMM_TIMING
        tm_a = clock64();
#endif
        ptx::mbar_wait_backoff(&tempty[acc], acc_phase ^ 1);
#ifdef BRN_GEMM_TIMING
        tm_empty += clock64() - tm_a;
#endif
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        const int ks = item - (int)p.fd_ks.div((uint32_t)item) * p.ksplit;
        const int kbn = min(kblocks, (ks + 1) * p.kb_per) - ks * p.kb_per;     // K blocks of this work item
        for (int kb = 0; kb < kbn; ++kb) {
#ifdef BRN_GEMM_TIMING
          tm_a = clock64();
#endif
          ptx::mbar_wait(&full[stage], phase);
#ifdef BRN_GEMM_TIMING
          tm_full += clock64() - tm_a;
#endif
          ptx::tc_fence_after();
          const uint64_t a_desc = ptx::make_smem_desc(ptx::smem_u32(sA + stage * TC_A_BYTES), 16, 1024, ptx::SW_128B);
          const uint64_t b_desc = ptx::make_smem_desc(ptx::smem_u32(sB + stage * BBYTES), 16, 1024, ptx::SW_128B);
          if (p.tg) {
            // vertical tap ky = the same box read one pixel row (8 pixels = one 1024-byte swizzle atom) further down
            const uint32_t b_step = (uint32_t)(p.BN * TC_BK * 2) >> 4;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
              for (int k = 0; k < TC_BK / 16; ++k)
                ptx::umma_f16_ss(d_tmem, a_desc + ky * (1024 >> 4) + 2 * k, b_desc + ky * b_step + 2 * k, idesc, (kb | ky | k) != 0);
          } else if (u2) {
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k)
              ptx::umma_f16_ss_2sm(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          } else {
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k)   // +32 B (= 2 in the >>4 address field) per K=16 step inside the 128B atom
              ptx::umma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          }
          // the stage is free once the MMAs of EVERY CTA that received the multicast have read it
          if (u2) ptx::umma_commit_2sm(&empty[stage], (uint16_t)3);
          else if (CL > 1) ptx::umma_commit_mc(&empty[stage], (uint16_t)((1u << CL) - 1));
          else ptx::umma_commit(&empty[stage]);
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
        if (u2) ptx::umma_commit_2sm(&tfull[acc], (uint16_t)3);     // both CTAs' epilogues
        else ptx::umma_commit(&tfull[acc]);
        acc ^= 1; if (acc == 0) acc_phase ^= 1;
      }
#ifdef BRN_GEMM_TIMING
      if (blockIdx.x == 0 || blockIdx.x == 77)
        printf("[gemm timing] cta %d EPI %d items %d: total %lld clk, mma waits: acc-empty %lld, smem-full %lld\n", (int)blockIdx.x, EPI,
               (num_items - item0 + item_step - 1) / item_step, clock64() - tm_t0, tm_empty, tm_full);
#endif
    }
  } else {
    // ===== epilogue: warp % 4 selects the TMEM lane quadrant, (warp - 2) / 4 the column part =====
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;
    constexpr int PER_Q = TC_EPI_WARPS / 4;
    const int row = q * 32 + lane;
    const int eth = threadIdx.x - 64;
    const uint32_t stage = ptx::smem_u32(sStage) + (warp - 2) * EPI_STAGE_BYTES;
    const bool o32 = p.epi.odt == F32;
    int acc = 0; uint32_t acc_phase = 0;
    // A-matrix row of this thread's tile row for a work item (-1: past the end / outside the image)
    auto a_row_of = [&](int item) -> long long {
      if (item >= num_items) return -1;
      const int it2 = (int)p.fd_ks.div((uint32_t)item);
      const int mg = (int)p.fd_nt.div((uint32_t)it2);
      const int m_tile = mg * CL + rank;
      const int b = (int)p.fd_tpi.div((uint32_t)m_tile), r = m_tile - b * tiles_per_img;
      const int ty = (int)p.fd_tx.div((uint32_t)r), tx = r - ty * p.tiles_x;
      const int y = ty * TH + (row >> p.tw_log2), x = tx * TW + (row & (TW - 1));
      if (!(m_tile < p.m_tiles && y < p.H && x < p.W)) return -1;
      return ((long long)b * p.H + y) * p.W + x;
    };
#ifdef BRN_GEMM_TIMING
    long long tm_ewait = 0; const long long tm_es = clock64();
#endif
    // Bias (and LnFold column sums) of ALL N columns staged once per kernel when they fit the staging area (N <= 576 with
    // LnFold, N <= 1152 without: the column-sum half of the area is free then; one bias vector for every image).  The
    // per-tile staging below is an L2 round trip + a barrier of the eight epilogue warps in front of every tile, which
    // the epilogue-bound short-K GEMMs (200 tiles per SM at stage 0) cannot hide.
    constexpr bool kLnfK = EPI == EK_LNF_NONE16 || EPI == EK_LNF_GELU16;
    constexpr int kOnceCap = (kLnfK ? 2 : 4) * TC_BIAS_LD;
    const bool bias_once = p.epi.bias_bstride == 0 && p.epi.N <= kOnceCap;
    if (bias_once) {
      for (int t = eth; t < kOnceCap; t += 32 * TC_EPI_WARPS) {
        ptx::sts32(ptx::smem_u32(sBias) + t * 4, (p.epi.bias && t < p.epi.N) ? __ldg(p.epi.bias + t) : 0.f);
        if (kLnfK) ptx::sts32(ptx::smem_u32(sCs) + t * 4, t < p.epi.N ? __ldg(p.epi.lnf_colsum + t) : 0.f);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
    }
    // otherwise: per tile, double buffered with the accumulator stage -- but fetched into registers one tile ahead, so
    // the L2 latency of the loads overlaps the previous tile's epilogue (thread t holds elements t and t + 256)
    float pb0 = 0.f, pb1 = 0.f, pc0 = 0.f, pc1 = 0.f;
    auto bias_fetch = [&](int item) {
      if (item >= num_items) return;
      const int it2 = (int)p.fd_ks.div((uint32_t)item);
      const int mg = (int)p.fd_nt.div((uint32_t)it2);
      const int m_tile = mg * CL + rank, n0 = (it2 - mg * p.n_tiles) * p.BN;
      const int b = (int)p.fd_tpi.div((uint32_t)m_tile);
      const float* bias = p.epi.bias ? p.epi.bias + (long long)(m_tile < p.m_tiles ? b : 0) * p.epi.bias_bstride : nullptr;
      const int t1 = eth + 32 * TC_EPI_WARPS;
      pb0 = (bias && eth < p.BN && n0 + eth < p.epi.N) ? __ldg(bias + n0 + eth) : 0.f;
      pb1 = (bias && t1 < p.BN && n0 + t1 < p.epi.N) ? __ldg(bias + n0 + t1) : 0.f;
      if (kLnfK) {
        pc0 = (eth < p.BN && n0 + eth < p.epi.N) ? __ldg(p.epi.lnf_colsum + n0 + eth) : 0.f;
        pc1 = (t1 < p.BN && n0 + t1 < p.epi.N) ? __ldg(p.epi.lnf_colsum + n0 + t1) : 0.f;
      }
    };
    static_assert(TC_BIAS_LD <= 64 * TC_EPI_WARPS, "two staged elements per epilogue thread");
    if (!bias_once) bias_fetch(item0);
    float2 mr_next = make_float2(0.f, 1.f);
    if (EPI == EK_LNF_NONE16 || EPI == EK_LNF_GELU16) {
      const long long a0 = a_row_of(item0);
      if (a0 >= 0) mr_next = __ldg(p.epi.lnf_mr + a0);
    }
    for (int item = item0; item < num_items; item += item_step) {
      const int it2 = (int)p.fd_ks.div((uint32_t)item), ks = item - it2 * p.ksplit;
      const int mg = (int)p.fd_nt.div((uint32_t)it2);
      const int m_tile = mg * CL + rank, n_tile = it2 - mg * p.n_tiles;
      const int b = (int)p.fd_tpi.div((uint32_t)m_tile), r = m_tile - b * tiles_per_img;
      const int ty = (int)p.fd_tx.div((uint32_t)r), tx = r - ty * p.tiles_x;
      const int y = ty * TH + (row >> p.tw_log2), x = tx * TW + (row & (TW - 1));
      const bool valid = m_tile < p.m_tiles && y < p.H && x < p.W;
      long long orow = valid ? ((long long)b * p.H + y) * p.W + x : -1;
      constexpr bool kLnf = EPI == EK_LNF_NONE16 || EPI == EK_LNF_GELU16;
      // LnFold: this row's (-mean, rstd) was fetched during the previous tile's epilogue; fetch the next tile's now
      float nmu = 0.f, rstd = 1.f;
      if (kLnf) {
        nmu = mr_next.x; rstd = mr_next.y;
        const long long an = a_row_of(item + item_step);
        if (an >= 0) mr_next = __ldg(p.epi.lnf_mr + an);
      }
      if (valid && p.rm.enabled) orow = rowmap_token_fd(p, orow);
      if (valid) orow += ks * p.rows_total;
      const int n0 = n_tile * p.BN;
      const int nch = min(p.BN, p.epi.N - n0 + 15) >> 4;     // 16-column chunks of this tile that hold real columns
      const int gsz = o32 ? 1 : 2;                           // 16-bit output: keep the split on 32-column granules
      const int per = ((nch + gsz - 1) / gsz + PER_Q - 1) / PER_Q * gsz;
      const int c0 = min(part * per, nch) * 16, c1 = min((part + 1) * per, nch) * 16;
      const uint32_t sb = ptx::smem_u32(sBias) + (bias_once ? n0 : acc * TC_BIAS_LD) * 4;
      const uint32_t scs_t = ptx::smem_u32(sCs) + (bias_once ? n0 : acc * TC_BIAS_LD) * 4;
      if (!bias_once) {
        // this tile's bias (fetched during the previous tile) -> shared; an M tile lies inside one image.  Staged
        // unconditionally (zeros without a bias): an optional add costs a register move per element.
        const int t1 = eth + 32 * TC_EPI_WARPS;
        ptx::sts32(sb + eth * 4, pb0);
        if (t1 < TC_BIAS_LD) ptx::sts32(sb + t1 * 4, pb1);
        if (kLnf) {
          ptx::sts32(scs_t + eth * 4, pc0);
          if (t1 < TC_BIAS_LD) ptx::sts32(scs_t + t1 * 4, pc1);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
        bias_fetch(item + item_step);
      }
#ifdef BRN_GEMM_TIMING
      const long long tm_e0 = clock64();
#endif
      ptx::mbar_wait(&tfull[acc], acc_phase);
#ifdef BRN_GEMM_TIMING
      tm_ewait += clock64() - tm_e0;
#endif
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256;
      const uint32_t sbb = sb;
      const int trow0 = m_tile * TC_BM + q * 32;       // TMA-store path: token matrices only (tiles = 128 consecutive rows)
      if (EPI == EK_NONE16 && p.tma_store)
        epi_warp<ACT_NONE, false, 0, true, false, false, true>(p.epi, taddr, n0, c0, c1, orow, sbb, stage, lane, 0.f, 1.f, 0, 0, &tmO, trow0);
      else if (EPI == EK_GELU16 && p.tma_store)
        epi_warp<ACT_GELU, false, 0, true, false, false, true>(p.epi, taddr, n0, c0, c1, orow, sbb, stage, lane, 0.f, 1.f, 0, 0, &tmO, trow0);
      else if (EPI == EK_LNF_GELU16 && p.tma_store)
        epi_warp<ACT_GELU, false, 0, true, true, false, true>(p.epi, taddr, n0, c0, c1, orow, sbb, stage, lane, nmu, rstd,
                                                              scs_t, 0, &tmO, trow0);
      else if (EPI == EK_NONE16) epi_warp<ACT_NONE, false, 0>(p.epi, taddr, n0, c0, c1, orow, sbb, stage, lane);
      else if (EPI == EK_RELU16) epi_warp<ACT_RELU, false, 0>(p.epi, taddr, n0, c0, c1, orow, sbb, stage, lane);
      else if (EPI == EK_GELU16) epi_warp<ACT_GELU, false, 0>(p.epi, taddr, n0, c0, c1, orow, sbb, stage, lane);
      else if (EPI == EK_NONE32) epi_warp<ACT_NONE, true, 0>(p.epi, taddr, n0, c0, c1, orow, sbb, stage, lane);
      else if (EPI == EK_RES32) epi_warp<ACT_NONE, true, 1>(p.epi, taddr, n0, c0, c1, orow, sbb, stage, lane);
      else if (EPI == EK_RES16) epi_warp<ACT_NONE, false, 2>(p.epi, taddr, n0, c0, c1, orow, sbb, stage, lane);
      else if (EPI == EK_LNF_NONE16)
        epi_warp<ACT_NONE, false, 0, true, true, false>(p.epi, taddr, n0, c0, c1, orow, sbb, stage, lane, nmu, rstd,
                                                        scs_t);
      else if (EPI == EK_LNF_GELU16)
        epi_warp<ACT_GELU, false, 0, true, true, false>(p.epi, taddr, n0, c0, c1, orow, sbb, stage, lane, nmu, rstd,
                                                        scs_t);
      else if (EPI == EK_RES32_EMIT)
        epi_warp<ACT_NONE, true, 1, false, false, true>(p.epi, taddr, n0, c0, c1, orow, sbb, stage, lane, 0.f, 1.f, 0,
                                                        n_tile * PER_Q + part);
      else if (EPI == EK_NONE32_EMIT)
        epi_warp<ACT_NONE, true, 0, true, false, true>(p.epi, taddr, n0, c0, c1, orow, sbb, stage, lane, 0.f, 1.f, 0,
                                                       n_tile * PER_Q + part);
      else if (EPI == EK_SIG32_TILED) {
        if (m_tile < p.m_tiles) epi_warp_tiled<ACT_2SIGMOID_TAIL>(p.epi, taddr, n0, c0, c1, m_tile, row, sbb);
      } else if (p.out_tiled) {
        if (m_tile < p.m_tiles) epi_warp_tiled_dyn(p.epi, taddr, n0, c0, c1, m_tile, row, sbb);
      } else {
        epi_warp_dyn(p.epi, taddr, n0, c0, c1, orow, sbb, stage, lane);
      }
      ptx::tc_fence_before();
      if (u2) {
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa_shared(ptx::smem_u32(&tempty[acc]), 0));
      } else {
        ptx::mbar_arrive(&tempty[acc]);
      }
      acc ^= 1; if (acc == 0) acc_phase ^= 1;
    }
#ifdef BRN_GEMM_TIMING
    if ((blockIdx.x == 0 || blockIdx.x == 77) && lane == 0 && (warp == 2 || warp == 9))
      printf("[gemm timing] cta %d epilogue warp %d: total %lld clk, waits on acc-full %lld\n", (int)blockIdx.x, warp, clock64() - tm_es, tm_ewait);
#endif
    if ((EPI == EK_NONE16 || EPI == EK_GELU16 || EPI == EK_LNF_GELU16) && p.tma_store && lane == 0)
      ptx::tma_store_wait_all();           // this warp's bulk stores are performed before the CTA retires
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CL > 1) ptx::cluster_sync_all();     // no CTA leaves while a peer may still multicast into it / arrive on its barriers
  if (warp == 1) { if (u2) ptx::tmem_dealloc_2sm(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)f;
  });
  BRN_CHECK(fn != nullptr, 2, "cuTensorMapEncodeTiled is not available from the driver");
  return fn;
}

// rank-`rank` bf16 / fp16 / fp32 tensor map, zero OOB fill.  dims/strides innermost first; strides[0] is implicit.
CUtensorMap make_tmap_16(const void* base, int dt, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box, CUtensorMapSwizzle swz) {
  CUtensorMap m;
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  const CUtensorMapDataType cdt = dt == F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : dt == F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = get_encode()(&m, cdt, rank, const_cast<void*>(base), gd, gs, bx, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BRN_CHECK(r == CUDA_SUCCESS, 2, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return m;
}

int device_sm_count() {
  static int n = 0;
  if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); }
  return n;
}

bool tc_gemm_supported(const GemmArgs& a) {
  if (!a.w) return false;
  if ((a.x.dt != BF16 && a.x.dt != F16) || !a.w->w16(a.x.dt)) return false;
  if (a.x.C % 8 != 0 || a.x.ld % 8 != 0 || ((uintptr_t)a.x.p & 15)) return false;
  if (a.stride != 1 || a.pad != (a.w->kh - 1) / 2 || a.w->kh != a.w->kw) return false;   // stride 1, "same" padding only
  if (a.rowmap.enabled && !(a.x.B == 1 && a.x.H == 1)) return false;
  if (a.bias_bstride && a.x.H * a.x.W < 1) return false;
  return true;
}

EpiP make_epi(int N, const float* bias, int bias_bstride, int act, int act_from, const View& res, const View& out) {
  EpiP e{};
  e.N = N; e.bias = bias; e.bias_bstride = bias_bstride; e.act = act; e.act_from = act_from;
  e.res = res.p; e.resdt = res.dt; e.ldres = res.ld;
  e.out = out.p; e.odt = out.dt; e.ldo = out.ld;
  e.vec = (((uintptr_t)out.p & 15) == 0) && ((out.ld * dsize(out.dt)) % 16 == 0) &&
          (!res.p || ((((uintptr_t)res.p & 15) == 0) && ((res.ld * dsize(res.dt)) % 16 == 0)));
  return e;
}

static int pick_bn(int N);
int tc_gemm_ln_parts(int N) {
  const int bn = pick_bn(N);
  return ((N + bn - 1) / bn) * (TC_EPI_WARPS / 4);
}

static int pick_bn(int N) {
  if (N <= 256) return (N + 15) / 16 * 16;
  int best = 256, best_pad = (N + 255) / 256 * 256;
  const int cands[2] = {192, 128};
  for (int c : cands) {
    int padded = (N + c - 1) / c * c;
    if (padded < best_pad) { best = c; best_pad = padded; }
  }
  return best;
}

// split-K tail: out[m, n] = act(sum_s part[s][m][n] + bias[n]) in a fixed order (deterministic), 4 columns per thread
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ part, int S, long long rows, int N,
                                                            const float* __restrict__ bias, int act, void* out, int odt,
                                                            int ldo) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  const int n4 = N >> 2;
  if (i >= rows * n4) return;
  const long long m = i / n4; const int n = (int)(i - m * n4) * 4;
  float4 acc = bias ? __ldg(reinterpret_cast<const float4*>(bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < S; ++s) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(part + ((long long)s * rows + m) * N + n));
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  if (act == ACT_RELU) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
  if (odt == F32) {
    float* o = (float*)out + m * ldo + n;
    o[0] = acc.x; o[1] = acc.y; o[2] = acc.z; o[3] = acc.w;
  } else {
    uint16_t* o = (uint16_t*)out + m * ldo + n;
    const uint32_t lo = odt == BF16 ? pack_bf16x2(acc.x, acc.y) : pack_f16x2(acc.x, acc.y);
    const uint32_t hi = odt == BF16 ? pack_bf16x2(acc.z, acc.w) : pack_f16x2(acc.z, acc.w);
    o[0] = (uint16_t)lo; o[1] = (uint16_t)(lo >> 16); o[2] = (uint16_t)hi; o[3] = (uint16_t)(hi >> 16);
  }
}

// Split-K plan: with very few output tiles per image (decoder convs at 32x32: 8 tiles, 270 K stages each) a handful of
// SMs walks a long K loop while the rest idle at small batch.  The decision and the split depend on the PER-IMAGE
// geometry only, never on the batch size, so an image's result does not depend on what it is batched with.
// Returns the number of K splits (1 = none) and the K blocks per split.
constexpr int SPLITK_MAX = 16, SPLITK_MAX_TILES = 16, SPLITK_MAX_N = 64;
size_t tc_gemm_splitk_scratch_bytes(int batch) {
  return (size_t)batch * SPLITK_MAX * SPLITK_MAX_TILES * TC_BM * SPLITK_MAX_N * 4;
}
static int plan_splitk(const LaunchCtx& ctx, const GemmArgs& a, int tiles_per_img, int n_tiles, int kblocks, int* kb_per) {
  static const bool off = [] { const char* v = getenv("BRN_GEMM_SPLITK"); return v && v[0] == '0'; }();
  *kb_per = kblocks;
  const LayerW& w = *a.w;
  if (off || !ctx.splitk || a.out_tiled || a.rowmap.enabled || a.res.p || a.bias_bstride != 0 || w.N % 4 != 0 ||
      w.N > SPLITK_MAX_N || (a.act != ACT_NONE && a.act != ACT_RELU) || tiles_per_img * n_tiles > SPLITK_MAX_TILES ||
      kblocks < 16)
    return 1;
  int S = std::min(SPLITK_MAX, kblocks / 4);
  const int per = (kblocks + S - 1) / S;
  S = (kblocks + per - 1) / per;
  if (S < 2) return 1;
  BRN_CHECK((size_t)S * (size_t)a.x.rows() * w.N * 4 <= ctx.splitk_bytes, 2, "tc_gemm: split-K scratch too small");
  *kb_per = per;
  return S;
}

void tc_gemm(const LaunchCtx& ctx, const GemmArgs& a) {
  const LayerW& w = *a.w;
  if (ctx.launches) ++*ctx.launches;
  BRN_CHECK(w.Cin == a.x.C, 5, "tc_gemm: weight/input channel mismatch");
  TcGemmP p{};
  p.B = a.x.B; p.H = a.x.H; p.W = a.x.W;
  // 3x3 convs with few output channels are L2->SMEM bound when every tap re-loads its 128-pixel box (9 loads per
  // channel block for <= 80 MMA columns): with 16x8-pixel tiles one (16+2) x 8 box serves the three vertical taps.
  static const bool no_tg = [] { const char* v = getenv("BRN_GEMM_TG"); return v && v[0] == '0'; }();
  const int bn0 = pick_bn(w.N);
  const bool tg = !no_tg && w.kh == 3 && w.kw == 3 && a.pad == 1 && a.tile_w <= 0 && w.N <= bn0 &&
                  3 * bn0 * TC_BK * 2 <= TC_B_BYTES && bn0 % 8 == 0 && a.x.W >= 8 && a.x.H >= 16;
  int tw = 1, lg = 0;
  if (tg) { tw = 8; lg = 3; }
  else if (a.tile_w > 0) { while (tw < std::min(a.tile_w, TC_BM)) { tw <<= 1; ++lg; } }
  else { while (tw < std::min(a.x.W, TC_BM)) { tw <<= 1; ++lg; } }
  p.tg = tg ? 1 : 0;
  p.tw_log2 = lg;
  p.out_tiled = a.out_tiled;
  BRN_CHECK(!a.out_tiled || (a.out.dt == F32 && !a.res.p && !a.rowmap.enabled && w.N <= 256), 5,
            "tc_gemm: tile-major output is fp32, single N tile, no residual");
  const int TH = TC_BM / tw;
  p.tiles_x = (a.x.W + tw - 1) / tw;
  p.tiles_y = (a.x.H + TH - 1) / TH;
  p.m_tiles = a.x.B * p.tiles_x * p.tiles_y;
  p.BN = pick_bn(w.N);
  p.n_tiles = (w.N + p.BN - 1) / p.BN;
  p.taps = w.taps(); p.kw = w.kw; p.pad = a.pad; p.cin_pad = w.cin_pad; p.cblocks = w.cin_pad / TC_BK;
  // split-K: decided from shapes only, so the plan pass (which stops here) counts the reduce launch as well
  const int kblocks = (tg ? 3 : w.taps()) * p.cblocks;
  int kb_per = kblocks;
  const int S = plan_splitk(ctx, a, p.tiles_x * p.tiles_y, p.n_tiles, kblocks, &kb_per);
  if (S > 1 && ctx.launches) ++*ctx.launches;
  if (ctx.dry) return;
  p.ksplit = S; p.kb_per = kb_per; p.rows_total = a.x.rows(); p.fd_ks = FastDiv((uint32_t)S);
  const View part = S > 1 ? make_view(ctx.splitk, F32, 1, 1, (int)(S * a.x.rows()), w.N) : View{};
  if (S > 1) p.epi = make_epi(w.N, nullptr, 0, ACT_NONE, 0, View{}, part);
  else p.epi = make_epi(w.N, a.bias ? a.bias : w.bias, a.bias_bstride, a.act, a.act_from, a.res, a.out);
  p.rm = a.rowmap;
  p.fd_nt = FastDiv((uint32_t)p.n_tiles);
  p.fd_tpi = FastDiv((uint32_t)(p.tiles_x * p.tiles_y));
  p.fd_tx = FastDiv((uint32_t)p.tiles_x);
  if (p.rm.enabled == 2) {
    p.fd_hw1 = FastDiv((uint32_t)(p.rm.h * p.rm.w)); p.fd_w1 = FastDiv((uint32_t)p.rm.w);
    if (p.rm.split > 0) { p.fd_hw2 = FastDiv((uint32_t)(p.rm.h2 * p.rm.w2)); p.fd_w2 = FastDiv((uint32_t)p.rm.w2); }
  }
  if (a.lnf.mr) {
    BRN_CHECK(S == 1 && !a.out_tiled && !a.res.p && a.out.dt != F32 && w.taps() == 1 && w.colsum(a.x.dt) &&
              (a.act == ACT_NONE || a.act == ACT_GELU) && a.lnf.C == w.Cin, 5,
              "tc_gemm: LnFold needs a 1x1 layer with column sums, a 16-bit output and act none|gelu");
    p.epi.lnf_mr = a.lnf.mr; p.epi.lnf_colsum = w.colsum(a.x.dt);
  }
  if (a.lne.stats) {
    BRN_CHECK(S == 1 && !a.out_tiled && a.out.dt == F32 && a.act == ACT_NONE && w.N % 16 == 0 && p.epi.vec &&
              (!a.res.p || a.res.dt == F32) && a.lne.x16 && a.lne.ldx16 % 4 == 0 && (((uintptr_t)a.lne.x16) & 7) == 0, 5,
              "tc_gemm: LnEmit needs an fp32 output with N % 16 == 0 and aligned rows");
    p.epi.lne_stats = a.lne.stats; p.epi.lne_stride = a.lne.stride;
    p.epi.x16 = a.lne.x16; p.epi.x16dt = a.lne.x16dt; p.epi.ldx16 = a.lne.ldx16;
  }
  if (p.rm.enabled) {
    BRN_CHECK((long long)a.x.rows() < (1ll << 31), 5, "tc_gemm: row map needs fewer than 2^31 rows");
    const int ws = p.rm.ws;
    BRN_CHECK(p.rm.enabled != 2 || ws == 12, 7, "tc_gemm: the token -> window scatter is built for 12x12 windows");
    p.fd_rows1 = FastDiv((uint32_t)((p.rm.hp / ws) * (p.rm.wp / ws) * ws * ws));
    p.fd_nww1 = FastDiv((uint32_t)(p.rm.wp / ws));
    if (p.rm.split > 0) {
      p.fd_rows2 = FastDiv((uint32_t)((p.rm.hp2 / ws) * (p.rm.wp2 / ws) * ws * ws));
      p.fd_nww2 = FastDiv((uint32_t)(p.rm.wp2 / ws));
    }
  }
  p.in_bf16 = a.x.dt == BF16 ? 1 : 0;

  const uint64_t ld2 = (uint64_t)a.x.ld * 2;
  uint64_t adims[4] = {(uint64_t)a.x.C, (uint64_t)a.x.W, (uint64_t)a.x.H, (uint64_t)a.x.B};
  uint64_t astr[3] = {ld2, ld2 * a.x.W, ld2 * a.x.W * a.x.H};
  uint32_t abox[4] = {(uint32_t)TC_BK, (uint32_t)tw, (uint32_t)(tg ? TH + 2 : TH), 1};
  CUtensorMap tmA = make_tmap_16(a.x.p, a.x.dt, 4, adims, astr, abox, CU_TENSOR_MAP_SWIZZLE_128B);
  const uint64_t ktot = (uint64_t)w.taps() * w.cin_pad;
  uint64_t bdims[2] = {ktot, (uint64_t)w.N};
  uint64_t bstr[1] = {ktot * 2};
  // 2-CTA clusters (B multicast) whenever there is enough work for every SM pair and BN splits into two swizzle-
  // aligned halves; BRN_GEMM_CLUSTER=1 forces single-CTA launches (A/B testing)
  static const bool no_cluster = [] { const char* v = getenv("BRN_GEMM_CLUSTER"); return v && v[0] == '1'; }();
  const int sms = device_sm_count();
  const int CL = (!no_cluster && p.BN % 32 == 0 && p.m_tiles >= 2 && (long long)p.m_tiles * p.n_tiles >= sms) ? 2 : 1;
  uint32_t bbox[2] = {(uint32_t)TC_BK, (uint32_t)(p.BN / CL)};
  CUtensorMap tmB = make_tmap_16(w.w16(a.x.dt), a.x.dt, 2, bdims, bstr, bbox, CU_TENSOR_MAP_SWIZZLE_128B);

  const double rows = (double)a.x.rows();
  char desc[128] = "";
  if (ctx.kt) snprintf(desc, sizeof desc, "M=%lld N=%d K=%dx%d BN=%d tiles=%d act=%d res=%d odt=%d cl=%d tg=%d", (long long)a.x.rows(),
                       w.N, w.taps(), w.Cin, p.BN, p.m_tiles * p.n_tiles, a.act, a.res.p ? 1 : 0, a.out.dt, CL, p.tg);
  KScope ks(ctx, KC_GEMM_TC, 2.0 * rows * w.N * w.taps() * w.Cin,
            rows * a.x.C * 2 + rows * w.N * dsize(a.out.dt) + (double)w.N * w.taps() * w.cin_pad * 2, desc);
  // epilogue variant (the conditions mirror epi_warp_dyn's dispatch)
  const bool o32 = S > 1 || a.out.dt == F32;
  int ek = EK_GENERIC;
  if (S > 1) ek = EK_NONE32;
  else if (a.out_tiled) ek = (a.act == ACT_2SIGMOID_TAIL) ? EK_SIG32_TILED : EK_GENERIC;
  else if (!a.res.p) {
    if (!o32) ek = a.act == ACT_NONE ? EK_NONE16 : a.act == ACT_RELU ? EK_RELU16 : a.act == ACT_GELU ? EK_GELU16 : EK_GENERIC;
    else ek = a.act == ACT_NONE ? EK_NONE32 : EK_GENERIC;
  } else if (a.act == ACT_NONE && o32 && a.res.dt == F32 && p.epi.vec) ek = EK_RES32;
  else if (a.act == ACT_NONE && !o32 && a.res.dt == a.out.dt && p.epi.vec && w.N % 8 == 0) ek = EK_RES16;
  if (a.lnf.mr) ek = a.act == ACT_GELU ? EK_LNF_GELU16 : EK_LNF_NONE16;
  if (a.lne.stats) {
    BRN_CHECK(ek == EK_RES32 || ek == EK_NONE32, 5, "tc_gemm: LnEmit on an unsupported epilogue variant");
    ek = ek == EK_RES32 ? EK_RES32_EMIT : EK_NONE32_EMIT;
  }

  // TMA store of the staged 16-bit tile: plain row-major token matrices (a tile = 128 consecutive output rows), no
  // residual, no row map.  BRN_GEMM_TMA_STORE=0 keeps the LDS + STG phase B (A/B testing).
  static const bool no_ts = [] { const char* v = getenv("BRN_GEMM_TMA_STORE"); return v && v[0] == '0'; }();
  // CTA-pair UMMA (cta_group::2) for every 2-CTA launch that is not a tap-group conv; BRN_GEMM_UMMA2=0: multicast scheme
  static const bool no_u2 = [] { const char* v = getenv("BRN_GEMM_UMMA2"); return v && v[0] == '0'; }();
  // (K >= 1536 only: the pair couples the two CTAs' epilogues, which costs 20-40 % on the epilogue-bound short-K shapes
  //  and gains 1-2 % on the long-K ones -- kernel_bench A/B, r02 run G)
  // K threshold (in 64-element K blocks) from which the pair form is used, separately for the LayerNorm-fold consumers
  // (qkv, fc1) and the fp32-stream producers (proj, fc2, merge).  Isolated, the pair form wins from K = 1536 (24 blocks)
  // and loses ~1 % at K = 768; in the power-capped full-model step 24 / 24, 12 / 24 and 12 / 12 are within the run-to-run
  // noise of each other (271.4-273.4 images/s, same box, alternating: r02 runs D2 / E2)
  static const int u2_minkb = [] { const char* v = getenv("BRN_GEMM_U2_MINKB"); return v ? atoi(v) : 24; }();
  static const int u2_minkb_res = [] { const char* v = getenv("BRN_GEMM_U2_MINKB_RES"); return v ? atoi(v) : 24; }();
  const bool ek_lnf = ek == EK_LNF_GELU16 || ek == EK_LNF_NONE16, ek_emit = ek == EK_RES32_EMIT || ek == EK_NONE32_EMIT;
  p.u2 = (!no_u2 && CL == 2 && !tg && S == 1 && p.BN == 256 &&
          ((ek_lnf && kblocks >= u2_minkb) || (ek_emit && kblocks >= u2_minkb_res))) ? 1 : 0;
  CUtensorMap tmO = tmA;                 // placeholder when unused (never dereferenced)
  p.tma_store = 0;
  if (!no_ts && (ek == EK_NONE16 || ek == EK_GELU16 || ek == EK_LNF_GELU16) && a.x.B == 1 && a.x.H == 1 && !a.rowmap.enabled &&
      p.epi.vec && a.out.B == 1 && a.out.H == 1 && tw == TC_BM) {
    uint64_t odims[2] = {(uint64_t)w.N, (uint64_t)a.out.rows()};
    uint64_t ostr[1] = {(uint64_t)a.out.ld * 2};
    uint32_t obox[2] = {32, 32};
    tmO = make_tmap_16(a.out.p, a.out.dt, 2, odims, ostr, obox, CU_TENSOR_MAP_SWIZZLE_64B);
    p.tma_store = 1;
  }
  auto launch = [&](auto kern) {
    const int TC_SMEM = p.u2 ? brn::TC_SMEM_U2 : brn::TC_SMEM;
    BRN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    const int items = ((p.m_tiles + CL - 1) / CL) * p.n_tiles * p.ksplit;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CL * std::min(items, sms / CL));
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = TC_SMEM;
    cfg.stream = ctx.stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_attr(attr, 1);
    BRN_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmO, p));
  };
#define TC_EK_CASE(E) case E: if (CL == 2) launch(tc_gemm_kernel<2, E>); else launch(tc_gemm_kernel<1, E>); break
  if (p.u2) {      // CTA-pair instances exist for the epilogues of the long-K backbone GEMMs only
    switch (ek) {
      case EK_RES32_EMIT: launch(tc_gemm_kernel<2, EK_RES32_EMIT, true>); break;
      case EK_NONE32_EMIT: launch(tc_gemm_kernel<2, EK_NONE32_EMIT, true>); break;
      case EK_LNF_GELU16: launch(tc_gemm_kernel<2, EK_LNF_GELU16, true>); break;
      case EK_LNF_NONE16: launch(tc_gemm_kernel<2, EK_LNF_NONE16, true>); break;
      default: BRN_CHECK(false, 2, "internal: no CTA-pair instance for this epilogue");
    }
  } else
  switch (ek) {
    TC_EK_CASE(EK_NONE16); TC_EK_CASE(EK_RELU16); TC_EK_CASE(EK_GELU16); TC_EK_CASE(EK_NONE32);
    TC_EK_CASE(EK_SIG32_TILED); TC_EK_CASE(EK_RES32); TC_EK_CASE(EK_RES16);
    TC_EK_CASE(EK_LNF_NONE16); TC_EK_CASE(EK_LNF_GELU16); TC_EK_CASE(EK_RES32_EMIT); TC_EK_CASE(EK_NONE32_EMIT);
    default: if (CL == 2) launch(tc_gemm_kernel<2, EK_GENERIC>); else launch(tc_gemm_kernel<1, EK_GENERIC>); break;
  }
#undef TC_EK_CASE
  BRN_CUDA(cudaGetLastError());
  if (S > 1) {
    const long long rows = a.x.rows(), work = rows * (w.N / 4);
    splitk_reduce_kernel<<<(unsigned)((work + 255) / 256), 256, 0, ctx.stream>>>(
        (const float*)ctx.splitk, S, rows, w.N, a.bias ? a.bias : w.bias, a.act, a.out.p, a.out.dt, a.out.ld);
    BRN_CUDA(cudaGetLastError());
  }
}

}  // namespace brn
