// LayerNorm kernels (candle_nn::layer_norm, eps 1e-5, biased variance), HBM-bound: one warp per destination row, the
// row held in registers as float4 vectors (one global read of the fp32 residual stream, one 16-bit write).
//  LN_PLAIN : row m <- row m                                     (norm2, patch_embed.norm, norm{i})
//  LN_WINDOW: window-ordered padded row m <- token row, zeros for pad rows: norm1 -> pad -> roll -> window_partition
//             (src/swin.rs:355-380) as one gather; pad rows are zeros AFTER the norm
//  LN_MERGE : row (b,i,j) <- [x(2i,2j) | x(2i+1,2j) | x(2i,2j+1) | x(2i+1,2j+1)], LN over 4C (src/swin.rs:505-525)
// Algorithmic bytes per row: 4*n read + esize*n written (n = C or 4C).
#include "brn_common.h"
#include "device_utils.cuh"
#include "tc_ptx.cuh"

namespace brn {

int device_sm_count();

struct LnP {
  const void* x; int xdt; int ldx; int B, h, w, C;
  const float* gamma; const float* beta;
  void* out; int odt; int ldo;
  int mode, hp, wp, shift, ws;
  long long rows;
  // LN_WINDOW over two token grids in one launch (merged two-resolution pass): window rows >= split belong to grid 2
  long long split, tok2;
  int h2, w2, hp2, wp2;
};

__device__ __forceinline__ long long ln_window_token(const LnP& p, long long m) {
  if (p.split > 0 && m >= p.split) {
    const long long t = window_row_to_token(m - p.split, p.h2, p.w2, p.hp2, p.wp2, p.shift, p.ws);
    return t < 0 ? t : t + p.tok2;
  }
  return window_row_to_token(m, p.h, p.w, p.hp, p.wp, p.shift, p.ws);
}

__device__ __forceinline__ void ln_store4(void* out, int odt, long long idx, float4 v) {
  if (odt == F32) {
    *reinterpret_cast<float4*>((float*)out + idx) = v;
  } else if (odt == BF16) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    *reinterpret_cast<uint2*>((__nv_bfloat16*)out + idx) = u;
  } else {
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 u = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    *reinterpret_cast<uint2*>((__half*)out + idx) = u;
  }
}

// fp32 source, n % 4 == 0, NV = ceil(n / 128) float4 per lane
template <int NV>
__global__ void __launch_bounds__(256) ln_vec_kernel(LnP p) {
  const int lane = threadIdx.x & 31;
  const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= p.rows) return;
  const int n = p.mode == LN_MERGE ? 4 * p.C : p.C;
  const int n4 = n >> 2;
  const float* xs = (const float*)p.x;
  long long src[4] = {-1, -1, -1, -1};
  if (p.mode == LN_PLAIN) {
    src[0] = m * p.ldx;
  } else if (p.mode == LN_WINDOW) {
    const long long tok = ln_window_token(p, m);
    if (tok < 0) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int e4 = lane + 32 * i;
        if (e4 < n4) ln_store4(p.out, p.odt, m * p.ldo + 4 * e4, make_float4(0.f, 0.f, 0.f, 0.f));
      }
      return;
    }
    src[0] = tok * p.ldx;
  } else {
    const int h2 = (p.h + 1) / 2, w2 = (p.w + 1) / 2;
    const long long b = m / ((long long)h2 * w2);
    const int r = (int)(m - b * (long long)h2 * w2);
    const int i = r / w2, j = r - i * w2;
#pragma unroll
    for (int part = 0; part < 4; ++part) {
      const int y = 2 * i + (part & 1), xx = 2 * j + (part >> 1);
      if (y < p.h && xx < p.w) src[part] = ((b * p.h + y) * (long long)p.w + xx) * p.ldx;
    }
  }
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int e4 = lane + 32 * i;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e4 < n4) {
      if (p.mode == LN_MERGE) {
        const int e = 4 * e4, part = e / p.C, c = e - part * p.C;
        if (src[part] >= 0) v[i] = __ldg(reinterpret_cast<const float4*>(xs + src[part] + c));
      } else {
        v[i] = __ldg(reinterpret_cast<const float4*>(xs + src[0] + 4 * e4));
      }
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(s) / n;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (lane + 32 * i < n4) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / n + 1e-5f);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int e4 = lane + 32 * i;
    if (e4 < n4) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma) + e4);
      const float4 bb = __ldg(reinterpret_cast<const float4*>(p.beta) + e4);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + bb.x;
      o.y = (v[i].y - mean) * rstd * g.y + bb.y;
      o.z = (v[i].z - mean) * rstd * g.z + bb.z;
      o.w = (v[i].w - mean) * rstd * g.w + bb.w;
      ln_store4(p.out, p.odt, m * p.ldo + 4 * e4, o);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Bulk-copy pipelined LayerNorm (LN_PLAIN / LN_WINDOW, fp32 source): the register-staged kernel above keeps only
// ~3 KB per warp in flight and reaches ~55 % of the HBM roofline.  Here a producer warp streams chunks of R
// destination rows into a 4-stage shared-memory ring with cp.async.bulk (one copy per run of contiguous source rows:
// a whole chunk for LN_PLAIN, the 12-token window rows for LN_WINDOW, nothing for pad rows), so 72 KB per CTA and
// ~200 KB per SM are in flight with no register cost; 8 consumer warps normalise one row each from shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int LNB_STAGES = 4;
constexpr int LNB_CHUNK_FLOATS = 4608;     // R * C <= 4608 floats = 18 KB per stage
constexpr int LNB_THREADS = 288;           // 8 consumer warps + 1 producer warp

// NV float4 per lane, LPR lanes per row (32, or 16 for short rows: two rows per warp, no idle lanes at C = 192)
// ODT = output element type; EXACT: n / 4 == NV * LPR (no column predicates).
template <int NV, int LPR, int ODT, bool EXACT>
__global__ void __launch_bounds__(LNB_THREADS) ln_bulk_kernel(LnP p, int R, long long n_chunks) {
  constexpr int RPW = 32 / LPR;                                     // rows per warp pass
  constexpr bool GREG = NV <= 6;                                    // gamma / beta of this lane's columns live in registers
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  const int n = p.mode == LN_MERGE ? 4 * p.C : p.C, n4 = n >> 2;   // normalised row length
  float* ring = (float*)smem;                                      // [LNB_STAGES][R * n]
  float* sg = ring + LNB_STAGES * LNB_CHUNK_FLOATS;                // gamma [n]
  float* sb = sg + n;                                              // beta  [n]
  uint64_t* full = (uint64_t*)(sb + n);
  uint64_t* empty = full + LNB_STAGES;
  int* sflag = (int*)(empty + LNB_STAGES);                         // [LNB_STAGES][32]: 1 = real row, 0 = pad / past the end
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < LNB_STAGES; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 8); }
    ptx::fence_barrier_init();
  }
  for (int i = threadIdx.x; i < n; i += LNB_THREADS) { sg[i] = p.gamma[i]; sb[i] = p.beta[i]; }
  __syncthreads();
  const float* xs = (const float*)p.x;
  const uint32_t row_bytes = (uint32_t)n * 4;

  if (warp == 8) {
    // ===== producer warp: lane j owns row j of the chunk =====
    int stage = 0; uint32_t phase = 0;
    for (long long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
      const long long m = ch * R + lane;
      if (p.mode == LN_MERGE) {
        // row (b, i, j) of the merged grid <- [x(2i,2j) | x(2i+1,2j) | x(2i,2j+1) | x(2i+1,2j+1)] (src/swin.rs:505-515):
        // four bulk copies of C floats per row (h and w are even here, so no part is padding)
        const bool valid = lane < R && m < p.rows;
        const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
        ptx::mbar_wait_backoff(&empty[stage], phase ^ 1);
        sflag[stage * 32 + lane] = valid ? 1 : 0;
        __syncwarp();
        if (lane == 0) ptx::mbar_expect_tx(&full[stage], (uint32_t)__popc(vmask) * row_bytes);
        __syncwarp();
        if (valid) {
          const int h2 = p.h >> 1, w2 = p.w >> 1, C = p.C;
          const long long b = m / ((long long)h2 * w2);
          const int r = (int)(m - b * (long long)h2 * w2);
          const int i = r / w2, j = r - i * w2;
#pragma unroll
          for (int part = 0; part < 4; ++part) {
            const long long srow = (b * p.h + 2 * i + (part & 1)) * (long long)p.w + 2 * j + (part >> 1);
            ptx::bulk_load(ring + stage * LNB_CHUNK_FLOATS + lane * n + part * C, xs + srow * p.ldx, (uint32_t)C * 4, &full[stage]);
          }
        }
        if (++stage == LNB_STAGES) { stage = 0; phase ^= 1; }
        continue;
      }
      long long tok = -1;
      if (lane < R && m < p.rows) tok = p.mode == LN_WINDOW ? ln_window_token(p, m) : m;
      const long long prev = __shfl_up_sync(0xffffffffu, tok, 1);
      const bool valid = tok >= 0;
      const bool start = valid && (lane == 0 || prev < 0 || tok != prev + 1 || p.ldx != n);
      const uint32_t vmask = __ballot_sync(0xffffffffu, valid), smask = __ballot_sync(0xffffffffu, start);
      ptx::mbar_wait_backoff(&empty[stage], phase ^ 1);
      sflag[stage * 32 + lane] = valid ? 1 : 0;
      __syncwarp();
      if (lane == 0) ptx::mbar_expect_tx(&full[stage], (uint32_t)__popc(vmask) * row_bytes);
      __syncwarp();
      if (start) {
        // the run ends before the next run start or the first invalid row after this lane
        const uint32_t above = ~((2u << lane) - 1u);                 // lanes > this one
        const uint32_t stop = (smask | ~vmask) & above;
        const int end = stop ? __ffs(stop) - 1 : 32;
        ptx::bulk_load(ring + stage * LNB_CHUNK_FLOATS + lane * n, xs + tok * p.ldx, (uint32_t)(end - lane) * row_bytes,
                       &full[stage]);
      }
      if (++stage == LNB_STAGES) { stage = 0; phase ^= 1; }
    }
  } else {
    // ===== consumers: row groups (RPW adjacent rows) go round-robin over the 8 warps across chunks =====
    // The first version of this loop issued ~22 instructions per element (64-bit index math per store, run-time
    // output-type switch, gamma / beta re-read per row, column predicates) and ran at 63 % issue-active / 52-64 % of
    // DRAM peak (ncu, profiles/r01_kernel_notes.md): everything row- or lane-invariant is hoisted now.
    const int sub = lane / LPR, l = lane % LPR;
    const int gpc = (R + RPW - 1) / RPW;                            // row groups per chunk
    const uint32_t ring_a = ptx::smem_u32(ring), sg_a = ptx::smem_u32(sg), sb_a = ptx::smem_u32(sb);
    const float inv_n = 1.f / (float)n;
    constexpr int OSZ = ODT == F32 ? 4 : 2;
    float4 gr[GREG ? NV : 1], br[GREG ? NV : 1];
    if (GREG) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int e4 = l + LPR * i;
        gr[i] = br[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (EXACT || e4 < n4) {
          const uint4 g4 = ptx::lds128(sg_a + e4 * 16), b4 = ptx::lds128(sb_a + e4 * 16);
          gr[i] = make_float4(__uint_as_float(g4.x), __uint_as_float(g4.y), __uint_as_float(g4.z), __uint_as_float(g4.w));
          br[i] = make_float4(__uint_as_float(b4.x), __uint_as_float(b4.y), __uint_as_float(b4.z), __uint_as_float(b4.w));
        }
      }
    }
    int stage = 0; uint32_t phase = 0;
    long long seq = 0;
    for (long long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x, ++seq) {
      ptx::mbar_wait(&full[stage], phase);
      const uint32_t src = ring_a + stage * (LNB_CHUNK_FLOATS * 4);
      for (int gj = (int)((8 + warp - (seq * gpc) % 8) % 8); gj < gpc; gj += 8) {
        const int j = gj * RPW + sub;
        const long long m = ch * R + j;
        const bool in_range = j < R && m < p.rows;
        const bool real = in_range && sflag[stage * 32 + j] != 0;
        const uint32_t srow = src + (uint32_t)(j * n + 4 * l) * 4;
        char* const orow = (char*)p.out + (m * p.ldo + 4 * l) * OSZ;
        float4 v[NV];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if ((EXACT || l + LPR * i < n4) && real) {
            const uint4 t = ptx::lds128(srow + i * (LPR * 16));
            v[i] = make_float4(__uint_as_float(t.x), __uint_as_float(t.y), __uint_as_float(t.z), __uint_as_float(t.w));
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
          }
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * inv_n;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          if (EXACT || l + LPR * i < n4) {
            v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
            q = fmaf(v[i].x, v[i].x, q); q = fmaf(v[i].y, v[i].y, q); q = fmaf(v[i].z, v[i].z, q); q = fmaf(v[i].w, v[i].w, q);
          }
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q * inv_n + 1e-5f);
        if (!in_range) continue;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          if (EXACT || l + LPR * i < n4) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);            // pad rows of LN_WINDOW are zeros AFTER the norm
            if (real) {
              float4 g, bb;
              if (GREG) { g = gr[i]; bb = br[i]; }
              else {
                const uint4 g4 = ptx::lds128(sg_a + (l + LPR * i) * 16), b4 = ptx::lds128(sb_a + (l + LPR * i) * 16);
                g = make_float4(__uint_as_float(g4.x), __uint_as_float(g4.y), __uint_as_float(g4.z), __uint_as_float(g4.w));
                bb = make_float4(__uint_as_float(b4.x), __uint_as_float(b4.y), __uint_as_float(b4.z), __uint_as_float(b4.w));
              }
              o.x = fmaf(v[i].x * rstd, g.x, bb.x);
              o.y = fmaf(v[i].y * rstd, g.y, bb.y);
              o.z = fmaf(v[i].z * rstd, g.z, bb.z);
              o.w = fmaf(v[i].w * rstd, g.w, bb.w);
            }
            char* op = orow + i * (LPR * 4 * OSZ);
            if (ODT == F32) {
              *reinterpret_cast<float4*>(op) = o;
            } else if (ODT == BF16) {
              __nv_bfloat162 a = __floats2bfloat162_rn(o.x, o.y), b2 = __floats2bfloat162_rn(o.z, o.w);
              *reinterpret_cast<uint2*>(op) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b2));
            } else {
              __half2 a = __floats2half2_rn(o.x, o.y), b2 = __floats2half2_rn(o.z, o.w);
              *reinterpret_cast<uint2*>(op) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b2));
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&empty[stage]);
      if (++stage == LNB_STAGES) { stage = 0; phase ^= 1; }
    }
  }
}

// generic fallback (any dtype / alignment): three passes over the row
__global__ void __launch_bounds__(256) ln_generic_kernel(LnP p) {
  const int lane = threadIdx.x & 31;
  const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= p.rows) return;
  const int n = p.mode == LN_MERGE ? 4 * p.C : p.C;
  long long src[4] = {-1, -1, -1, -1};
  if (p.mode == LN_PLAIN) {
    src[0] = m * p.ldx;
  } else if (p.mode == LN_WINDOW) {
    long long tok = ln_window_token(p, m);
    if (tok < 0) {
      for (int e = lane; e < n; e += 32) st_elem(p.out, p.odt, m * p.ldo + e, 0.f);
      return;
    }
    src[0] = tok * p.ldx;
  } else {
    const int h2 = (p.h + 1) / 2, w2 = (p.w + 1) / 2;
    long long b = m / ((long long)h2 * w2);
    int r = (int)(m - b * (long long)h2 * w2);
    int i = r / w2, j = r - i * w2;
#pragma unroll
    for (int part = 0; part < 4; ++part) {
      int y = 2 * i + (part & 1), xx = 2 * j + (part >> 1);
      if (y < p.h && xx < p.w) src[part] = ((b * p.h + y) * (long long)p.w + xx) * p.ldx;
    }
  }
  auto load = [&](int e) -> float {
    if (p.mode == LN_MERGE) {
      int part = e / p.C, c = e - part * p.C;
      return src[part] < 0 ? 0.f : ld_elem(p.x, p.xdt, src[part] + c);
    }
    return ld_elem(p.x, p.xdt, src[0] + e);
  };
  float s = 0.f;
  for (int e = lane; e < n; e += 32) s += load(e);
  const float mean = warp_sum(s) / n;
  float v = 0.f;
  for (int e = lane; e < n; e += 32) { float d = load(e) - mean; v += d * d; }
  const float rstd = rsqrtf(warp_sum(v) / n + 1e-5f);
  for (int e = lane; e < n; e += 32)
    st_elem(p.out, p.odt, m * p.ldo + e, (load(e) - mean) * rstd * p.gamma[e] + p.beta[e]);
}

void glue_layernorm(const LaunchCtx& ctx, const LnArgs& a) {
  if (ctx.launches) ++*ctx.launches;
  if (ctx.dry) return;
  LnP p{};
  p.x = a.x.p; p.xdt = a.x.dt; p.ldx = a.x.ld; p.B = a.x.B; p.h = a.x.H; p.w = a.x.W; p.C = a.x.C;
  p.gamma = a.gamma; p.beta = a.beta;
  p.out = a.out.p; p.odt = a.out.dt; p.ldo = a.out.ld;
  p.mode = a.mode; p.hp = a.hp; p.wp = a.wp; p.shift = a.shift; p.ws = a.ws;
  p.rows = a.out.rows();
  p.split = a.split; p.tok2 = a.tok2; p.h2 = a.h2; p.w2 = a.w2; p.hp2 = a.hp2; p.wp2 = a.wp2;
  const int n = a.mode == LN_MERGE ? 4 * a.x.C : a.x.C;
  KScope ks(ctx, KC_LN, 0.0, (double)p.rows * n * (4 + dsize(a.out.dt)),
            a.mode == LN_WINDOW ? "ln_window" : a.mode == LN_MERGE ? "ln_merge" : "ln_plain");
  const unsigned grid = (unsigned)((p.rows + 7) / 8);
  const bool vec = a.x.dt == F32 && a.x.C % 4 == 0 && a.x.ld % 4 == 0 && (((uintptr_t)a.x.p) & 15) == 0 &&
                   (a.out.ld * dsize(a.out.dt)) % (4 * dsize(a.out.dt)) == 0 && a.out.ld % 4 == 0 &&
                   (((uintptr_t)a.out.p) & (4 * dsize(a.out.dt) - 1)) == 0 && (((uintptr_t)a.gamma | (uintptr_t)a.beta) & 15) == 0;
  const int nv = (n + 127) / 128;
  static const bool no_bulk = [] { const char* v = getenv("BRN_LN_BULK"); return v && v[0] == '0'; }();
  const bool merge_ok = a.mode != LN_MERGE || (a.x.H % 2 == 0 && a.x.W % 2 == 0 && a.x.p != a.out.p);
  if (vec && !no_bulk && merge_ok && n <= LNB_CHUNK_FLOATS && nv <= 24 && a.x.p != nullptr &&
      !(a.x.p == a.out.p && a.mode == LN_WINDOW)) {
    int R = std::min(32, LNB_CHUNK_FLOATS / n);
    if (a.mode == LN_WINDOW) {      // chunks of whole window rows (runs of ws contiguous tokens), or divisors of one
      const int cands12[6] = {24, 12, 6, 4, 3, 2}, cands7[3] = {28, 14, 7};
      int r = 1;
      if (a.ws == 12) { for (int c : cands12) if (c <= R) { r = c; break; } }
      else if (a.ws == 7) { for (int c : cands7) if (c <= R) { r = c; break; } }
      R = r;
    }
    const long long n_chunks = (p.rows + R - 1) / R;
    const int smem = LNB_STAGES * LNB_CHUNK_FLOATS * 4 + 2 * n * 4 + 2 * LNB_STAGES * 8 + LNB_STAGES * 32 * 4 + 128;
#define LNB_LAUNCH(NV, LPR, ODT, EX)                                                                      \
    do {                                                                                                  \
      auto kern = ln_bulk_kernel<NV, LPR, ODT, EX>;                                                       \
      BRN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));            \
      int occ = 1;                                                                                        \
      BRN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, LNB_THREADS, smem));             \
      const int grid = (int)std::min<long long>(n_chunks, (long long)device_sm_count() * std::max(occ, 1)); \
      kern<<<grid, LNB_THREADS, smem, ctx.stream>>>(p, R, n_chunks);                                      \
    } while (0)
#define LNB_CASE(NV, LPR)                                                                                 \
    do {                                                                                                  \
      const bool ex = (n / 4) == NV * LPR;                                                                \
      if (a.out.dt == F32) { if (ex) LNB_LAUNCH(NV, LPR, F32, true); else LNB_LAUNCH(NV, LPR, F32, false); } \
      else if (a.out.dt == BF16) { if (ex) LNB_LAUNCH(NV, LPR, BF16, true); else LNB_LAUNCH(NV, LPR, BF16, false); } \
      else { if (ex) LNB_LAUNCH(NV, LPR, F16, true); else LNB_LAUNCH(NV, LPR, F16, false); }                \
    } while (0)
    const int n4 = n / 4;
    if (n4 <= 16) LNB_CASE(1, 16); else if (n4 <= 32) LNB_CASE(2, 16); else if (n4 <= 48) LNB_CASE(3, 16);
    else if (nv <= 2) LNB_CASE(2, 32); else if (nv <= 3) LNB_CASE(3, 32); else if (nv <= 4) LNB_CASE(4, 32);
    else if (nv <= 6) LNB_CASE(6, 32); else if (nv <= 8) LNB_CASE(8, 32); else if (nv <= 12) LNB_CASE(12, 32);
    else if (nv <= 16) LNB_CASE(16, 32); else LNB_CASE(24, 32);
#undef LNB_LAUNCH
#undef LNB_CASE
    BRN_CUDA(cudaGetLastError());
    return;
  }
  if (vec && nv <= 24) {
#define LN_CASE(NV) ln_vec_kernel<NV><<<grid, 256, 0, ctx.stream>>>(p)
    if (nv <= 1) LN_CASE(1); else if (nv <= 2) LN_CASE(2); else if (nv <= 3) LN_CASE(3); else if (nv <= 4) LN_CASE(4);
    else if (nv <= 6) LN_CASE(6); else if (nv <= 8) LN_CASE(8); else if (nv <= 12) LN_CASE(12);
    else if (nv <= 16) LN_CASE(16); else LN_CASE(24);
#undef LN_CASE
  } else {
    ln_generic_kernel<<<grid, 256, 0, ctx.stream>>>(p);
  }
  BRN_CUDA(cudaGetLastError());
}

}  // namespace brn
