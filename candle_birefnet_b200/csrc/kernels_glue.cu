// HBM-bound glue kernels: bilinear resampling, image2patches, layout converts, global average pool, GDT gate
// (LayerNorm lives in ln_kernels.cu, the fused final layer in final_kernel.cu).  All are bandwidth-bound; threads map to the contiguous
// (channel) dimension of NHWC so every warp access is a run of consecutive addresses.
#include "brn_common.h"
#include "device_utils.cuh"

namespace brn {

__device__ __forceinline__ float g_ld(const void* p, int dt, long long i) { return ld_elem(p, dt, i); }
__device__ __forceinline__ void g_st(void* p, int dt, long long i, float v) { st_elem(p, dt, i, v); }

// `bytes` = compulsory HBM traffic of the launch (every input element read once + every output element written once):
// the figure bench.py divides by the measured time for the glue class's achieved GB/s
#define GLUE_LAUNCH_PROLOGUE(ctx, bytes)  \
  if ((ctx).launches) ++*(ctx).launches; \
  if ((ctx).dry) return;                 \
  KScope ks__((ctx), KC_GLUE, 0.0, (double)(bytes), __func__);

// ------------------------------------------------------------------------------------------------
// PatchEmbed im2col (src/swin.rs:692-704): row (b,py,px), k = c*P*P + ky*P + kx  <-  x[b,c,py*P+ky,px*P+kx]
// ------------------------------------------------------------------------------------------------
__global__ void patch_im2col_kernel(const float* x, int B, int H, int W, int P, void* out, int odt, int ldo,
                                    long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int K = 3 * P * P;
  long long row = i / K; int k = (int)(i - row * K);
  int c = k / (P * P), r = k - c * P * P, ky = r / P, kx = r - ky * P;
  int Wp = W / P, Hp = H / P;
  long long b = row / ((long long)Hp * Wp); int rr = (int)(row - b * (long long)Hp * Wp);
  int py = rr / Wp, px = rr - py * Wp;
  float v = x[((b * 3 + c) * H + (py * P + ky)) * (long long)W + px * P + kx];
  g_st(out, odt, row * ldo + k, v);
}

// P = 4, 16-bit dense output rows (48 elements = 96 B): a warp owns 32 consecutive patches of one patch row.  Loads are
// float4 per (c, ky) with lane = patch (512 contiguous bytes per instruction); the 32 x 96 B block is transposed
// through shared memory and leaves as six fully coalesced 512-byte stores (32 consecutive rows are contiguous).
__global__ void __launch_bounds__(256) patch_im2col4_kernel(const float* __restrict__ x, int H, int W, uint16_t* out, int odt,
                                                            long long n_groups) {
  __shared__ __align__(16) uint8_t sm[8][32 * 96];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long g = (long long)blockIdx.x * 8 + warp;     // group = 32 consecutive patches of one patch row
  if (g >= n_groups) return;
  const int Wp = W >> 2, Hp = H >> 2, gpr = Wp >> 5;         // groups per patch row (Wp % 32 == 0)
  const int px = (int)(g % gpr) * 32 + lane;
  const long long t = g / gpr;
  const int py = (int)(t % Hp); const long long b = t / Hp;
  uint8_t* mine = sm[warp];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + ((b * 3 + c) * H + (py * 4 + ky)) * (long long)W + px * 4));
      uint2 u;
      if (odt == BF16) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), d = __floats2bfloat162_rn(v.z, v.w);
        u = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&d));
      } else {
        __half2 a = __floats2half2_rn(v.x, v.y), d = __floats2half2_rn(v.z, v.w);
        u = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&d));
      }
      *reinterpret_cast<uint2*>(mine + lane * 96 + (c * 16 + ky * 4) * 2) = u;   // 96-byte pitch: 8-byte stores, 2-way conflicts at most
    }
  __syncwarp();
  uint4* dst = reinterpret_cast<uint4*>(out + (((b * Hp + py) * (long long)Wp + (px - lane)) * 48));
#pragma unroll
  for (int i = 0; i < 6; ++i) dst[i * 32 + lane] = *reinterpret_cast<const uint4*>(mine + (i * 32 + lane) * 16);
}

void glue_patch_im2col(const LaunchCtx& ctx, const float* x, int B, int H, int W, int P, View out) {
  GLUE_LAUNCH_PROLOGUE(ctx, (double)B * 3 * H * W * (4 + dsize(out.dt)));
  if (P == 4 && out.dt != F32 && out.ld == 48 && (W / 4) % 32 == 0 && ((uintptr_t)out.p & 15) == 0 && ((uintptr_t)x & 15) == 0) {
    const long long groups = (long long)B * (H / 4) * (W / 4 / 32);
    patch_im2col4_kernel<<<(unsigned)((groups + 7) / 8), 256, 0, ctx.stream>>>(x, H, W, (uint16_t*)out.p, out.dt, groups);
    BRN_CUDA(cudaGetLastError());
    return;
  }
  long long total = (long long)B * (H / P) * (W / P) * 3 * P * P;
  patch_im2col_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx.stream>>>(x, B, H, W, P, out.p, out.dt, out.ld,
                                                                              total);
  BRN_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// Bilinear resample, align_corners = true (Tensor::upsample_bilinear2d(h,w,true), src/birefnet.rs x16; also used
// to down-sample, point-sampled, no antialias).  src = dst * (in-1)/(out-1).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bilin_coord(int dst, int in, int out, int& i0, int& i1, float& l) {
  float scale = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
  float s = scale * dst;
  i0 = min((int)s, in - 1);
  i1 = min(i0 + 1, in - 1);
  l = s - (float)i0;
}

__global__ void resize_nchw_kernel(const float* x, int C, int H, int W, float* out, int Ho, int Wo, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int ox, oy; long long bc;
  if (total <= 0x7fffffffll) {          // 32-bit index decode (the 64-bit divisions dominated the instruction count)
    const uint32_t i32 = (uint32_t)i, t32 = i32 / (uint32_t)Wo, b32 = t32 / (uint32_t)Ho;
    ox = (int)(i32 - t32 * (uint32_t)Wo); oy = (int)(t32 - b32 * (uint32_t)Ho); bc = b32;
  } else {
    ox = (int)(i % Wo); long long t = i / Wo; oy = (int)(t % Ho); bc = t / Ho;
  }
  int y0, y1, x0, x1; float ly, lx;
  bilin_coord(oy, H, Ho, y0, y1, ly);
  bilin_coord(ox, W, Wo, x0, x1, lx);
  const float* s = x + bc * (long long)H * W;
  float v = (1.f - ly) * ((1.f - lx) * s[(long long)y0 * W + x0] + lx * s[(long long)y0 * W + x1]) +
            ly * ((1.f - lx) * s[(long long)y1 * W + x0] + lx * s[(long long)y1 * W + x1]);
  out[i] = v;
}

void glue_resize_nchw(const LaunchCtx& ctx, const float* x, int B, int C, int H, int W, float* out, int Ho, int Wo) {
  GLUE_LAUNCH_PROLOGUE(ctx, (double)B * C * 4 * ((double)H * W + (double)Ho * Wo));
  long long total = (long long)B * C * Ho * Wo;
  resize_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx.stream>>>(x, C, H, W, out, Ho, Wo, total);
  BRN_CUDA(cudaGetLastError());
}

struct ResizeP {
  const void* x; int xdt; int ldx; int H, W, C;
  void* out; int odt; int ldo; int Ho, Wo;
  long long total;
};

__global__ void resize_nhwc_kernel(ResizeP p) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.total) return;
  int c = (int)(i % p.C); long long t = i / p.C;
  int ox = (int)(t % p.Wo); t /= p.Wo; int oy = (int)(t % p.Ho); long long b = t / p.Ho;
  int y0, y1, x0, x1; float ly, lx;
  bilin_coord(oy, p.H, p.Ho, y0, y1, ly);
  bilin_coord(ox, p.W, p.Wo, x0, x1, lx);
  long long base = b * p.H * p.W;
  float v00 = g_ld(p.x, p.xdt, (base + (long long)y0 * p.W + x0) * p.ldx + c);
  float v01 = g_ld(p.x, p.xdt, (base + (long long)y0 * p.W + x1) * p.ldx + c);
  float v10 = g_ld(p.x, p.xdt, (base + (long long)y1 * p.W + x0) * p.ldx + c);
  float v11 = g_ld(p.x, p.xdt, (base + (long long)y1 * p.W + x1) * p.ldx + c);
  float v = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
  g_st(p.out, p.odt, ((b * p.Ho + oy) * (long long)p.Wo + ox) * p.ldo + c, v);
}

// vector variant: 16-bit in == 16-bit out, 8 channels (16 bytes) per thread
__device__ __forceinline__ float2 up16(uint32_t u, int dt) {
  if (dt == BF16) return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
  return __half22float2(*reinterpret_cast<const __half2*>(&u));
}
__device__ __forceinline__ uint32_t pk16(float a, float b, int dt) {
  if (dt == BF16) { __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
  __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h);
}

// vector variant: 16-bit in == 16-bit out, 8 channels (16 bytes) per thread.  blockIdx.y = output row, blockIdx.z =
// image: the row interpolation is block-uniform and the only per-thread division is t / C8 (the flat index math of
// the scalar kernel was ~40 % of this kernel's instructions).
constexpr int RS_ROWS = 8;
template <int DT>
__global__ void __launch_bounds__(256) resize_nhwc_vec8_kernel(ResizeP p) {
  const int C8 = p.C >> 3;
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= p.Wo * C8) return;
  const int ox = t / C8, c8 = t - ox * C8;
  const long long b = blockIdx.z;
  int x0, x1; float lx;
  bilin_coord(ox, p.W, p.Wo, x0, x1, lx);
  const uint16_t* xs = (const uint16_t*)p.x + b * p.H * p.W * p.ldx + c8 * 8;
  const float hx = 1.f - lx;
  const unsigned long long hx2 = pk2(hx, hx), lx2 = pk2(lx, lx);
  // RS_ROWS consecutive output rows per thread.  The horizontally interpolated input row
  //   T(y) = (1-lx) * v(y, x0) + lx * v(y, x1)                      (8 channels, packed fp32 pairs)
  // is computed ONCE per input row and kept in registers: neighbouring output rows of an up-sampling share their two
  // input rows, so 8 output rows cost ~5 row loads instead of 16 (the kernel was issue / L1 bound at 4 loads per output).
  // Same association as the scalar kernel / F.interpolate: out = (1-ly) * T(y0) + ly * T(y1)  -- bit-identical.
  auto load_row = [&](int y, unsigned long long (&T)[4]) {
    const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(xs + ((long long)y * p.W + x0) * p.ldx));
    const uint4 a1 = __ldg(reinterpret_cast<const uint4*>(xs + ((long long)y * p.W + x1) * p.ldx));
    const uint32_t *q0 = reinterpret_cast<const uint32_t*>(&a0), *q1 = reinterpret_cast<const uint32_t*>(&a1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f0 = up16(q0[j], DT), f1 = up16(q1[j], DT);
      T[j] = fma2(lx2, pk2(f1.x, f1.y), fmul2(hx2, pk2(f0.x, f0.y)));
    }
  };
  unsigned long long TA[4], TB[4];
  int ya = -1, yb = -1;                     // input rows held in TA / TB (block-uniform: no divergence below)
  const int oy_end = min((int)(blockIdx.y + 1) * RS_ROWS, p.Ho);
  for (int oy = blockIdx.y * RS_ROWS; oy < oy_end; ++oy) {
    int y0, y1; float ly;
    bilin_coord(oy, p.H, p.Ho, y0, y1, ly);
    if (y0 != ya) {
      if (y0 == yb) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { const unsigned long long tmp = TA[j]; TA[j] = TB[j]; TB[j] = tmp; }
        yb = ya;
      } else {
        load_row(y0, TA);
      }
      ya = y0;
    }
    if (y1 != yb) {
      if (y1 == ya) {
#pragma unroll
        for (int j = 0; j < 4; ++j) TB[j] = TA[j];
      } else {
        load_row(y1, TB);
      }
      yb = y1;
    }
    const float hy = 1.f - ly;
    const unsigned long long hy2 = pk2(hy, hy), ly2 = pk2(ly, ly);
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float ax, ay;
      upk2(fma2(ly2, TB[j], fmul2(hy2, TA[j])), ax, ay);
      o[j] = pk16(ax, ay, DT);
    }
    *reinterpret_cast<uint4*>((uint16_t*)p.out + ((b * p.Ho + oy) * (long long)p.Wo + ox) * p.ldo + c8 * 8) =
        make_uint4(o[0], o[1], o[2], o[3]);
  }
}

void glue_resize_nhwc(const LaunchCtx& ctx, View in, View out) {
  GLUE_LAUNCH_PROLOGUE(ctx, (double)in.rows() * in.C * dsize(in.dt) + (double)out.rows() * out.C * dsize(out.dt));
  ResizeP p{in.p, in.dt, in.ld, in.H, in.W, in.C, out.p, out.dt, out.ld, out.H, out.W, 0};
  const bool vec = in.dt != F32 && in.dt == out.dt && in.C % 8 == 0 && in.ld % 8 == 0 && out.ld % 8 == 0 &&
                   (((uintptr_t)in.p | (uintptr_t)out.p) & 15) == 0;
  if (vec && out.H <= 65535 && in.B <= 65535) {
    dim3 grid((out.W * (in.C / 8) + 255) / 256, (out.H + RS_ROWS - 1) / RS_ROWS, in.B);
    if (in.dt == BF16) resize_nhwc_vec8_kernel<BF16><<<grid, 256, 0, ctx.stream>>>(p);
    else resize_nhwc_vec8_kernel<F16><<<grid, 256, 0, ctx.stream>>>(p);
    BRN_CUDA(cudaGetLastError());
    return;
  }
  p.total = (long long)in.B * out.H * out.W * in.C;
  resize_nhwc_kernel<<<(unsigned)((p.total + 255) / 256), 256, 0, ctx.stream>>>(p);
  BRN_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// image2patches (src/birefnet.rs:288-300): out[b,ty,tx, c*gh*gw + gy*gw + gx] = x[b,c,gy*th+ty,gx*tw+tx]
// ------------------------------------------------------------------------------------------------
__global__ void image2patches_kernel(const float* x, int H, int W, int th, int tw, void* out, int odt, int ldo,
                                     long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int gh = H / th, gw = W / tw, Cn = 3 * gh * gw;
  int ch = (int)(i % Cn); long long t = i / Cn;
  int tx = (int)(t % tw); t /= tw; int ty = (int)(t % th); long long b = t / th;
  int c = ch / (gh * gw), r = ch - c * gh * gw, gy = r / gw, gx = r - gy * gw;
  float v = x[((b * 3 + c) * H + (gy * th + ty)) * (long long)W + gx * tw + tx];
  g_st(out, odt, ((b * th + ty) * (long long)tw + tx) * ldo + ch, v);
}

// Tiled variant (16-bit output, tw % 32 == 0): a block owns 32 consecutive output pixels (b, ty, tx0..tx0+31) and walks
// the channels in chunks of CH.  For a channel (c, gy, gx) those 32 pixels are 32 CONSECUTIVE floats of one image row
// (one coalesced 128-byte read per warp); the [32 px][CH] chunk is transposed through shared memory and written as
// runs of CH contiguous 16-bit channels per pixel.
template <int CH>
__global__ void __launch_bounds__(256) image2patches_tiled_kernel(const float* __restrict__ x, int H, int W, int th, int tw,
                                                                  uint16_t* out, int odt, int ldo) {
  constexpr int PITCH = CH + 2;                       // 16-bit elements per shared row (odd word count: conflict-free columns)
  __shared__ __align__(16) uint16_t sm[32 * PITCH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = H / th, gw = W / tw, Cn = 3 * g * gw;
  const int lg = (g == gw && (g & (g - 1)) == 0) ? 31 - __clz(g) : -1;
  // one CH-channel chunk per block (blockIdx.z = image * chunks + chunk): at batch 1 the 32x32 level has only 32
  // pixel strips, far too few blocks when each of them walks all 3072 channels
  const int nchunks = Cn / CH;
  const int tx0 = blockIdx.x * 32, ty = blockIdx.y; const long long b = blockIdx.z / nchunks;
  {
    const int ch0 = (int)(blockIdx.z % nchunks) * CH;
    // four independent loads in flight per warp (a rolled loop had one: the kernel ran at 1.4 TB/s of reads)
    auto src_of = [&](int ch) -> const float* {
      int c, gy, gx;
      if (lg >= 0) {            // the grid is 2^lg x 2^lg (always, for H, W multiples of 32): shifts instead of divisions
        c = ch >> (2 * lg); gy = (ch >> lg) & (gw - 1); gx = ch & (gw - 1);
      } else {
        c = ch / (g * gw); const int r = ch - c * g * gw; gy = r / gw; gx = r - gy * gw;
      }
      return x + ((b * 3 + c) * H + (gy * th + ty)) * (long long)W + gx * tw + tx0 + lane;
    };
    auto cvt = [&](float v) -> uint16_t {
      if (odt == BF16) { __nv_bfloat16 t = __float2bfloat16(v); return *reinterpret_cast<uint16_t*>(&t); }
      __half t = __float2half_rn(v); return *reinterpret_cast<uint16_t*>(&t);
    };
    static_assert(CH % 32 == 0 || CH == 48, "channel chunk");
    constexpr int U = CH % 32 == 0 ? 4 : 2;            // CH / 8 iterations per warp: 32, 24 or 6
    for (int j0 = warp; j0 < CH; j0 += 8 * U) {
      float v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = (j0 + 8 * u < CH) ? __ldg(src_of(ch0 + j0 + 8 * u)) : 0.f;
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (j0 + 8 * u < CH) sm[lane * PITCH + j0 + 8 * u] = cvt(v[u]);
    }
    __syncthreads();
    for (int px = warp; px < 32; px += 8) {
      uint32_t* dst = reinterpret_cast<uint32_t*>(out + ((b * th + ty) * (long long)tw + tx0 + px) * ldo + ch0);
      const uint32_t* src = reinterpret_cast<const uint32_t*>(sm + px * PITCH);
      for (int wd = lane; wd < CH / 2; wd += 32) dst[wd] = src[wd];
    }
    __syncthreads();
  }
}

void glue_image2patches(const LaunchCtx& ctx, const float* x, int B, int H, int W, int th, int tw, View out) {
  GLUE_LAUNCH_PROLOGUE(ctx, (double)B * 3 * H * W * (4 + dsize(out.dt)));
  const int Cn = 3 * (H / th) * (W / tw);
  if (out.dt != F32 && tw % 32 == 0 && out.ld % 2 == 0 && ((uintptr_t)out.p & 3) == 0 && th <= 65535 && (long long)B * (Cn / 48) <= 65535) {
    auto grid_for = [&](int ch) { return dim3(tw / 32, th, B * (Cn / ch)); };
    if (Cn % 256 == 0) {
      image2patches_tiled_kernel<256><<<grid_for(256), 256, 0, ctx.stream>>>(x, H, W, th, tw, (uint16_t*)out.p, out.dt, out.ld);
      BRN_CUDA(cudaGetLastError());
      return;
    }
    if (Cn % 192 == 0) {
      image2patches_tiled_kernel<192><<<grid_for(192), 256, 0, ctx.stream>>>(x, H, W, th, tw, (uint16_t*)out.p, out.dt, out.ld);
      BRN_CUDA(cudaGetLastError());
      return;
    }
    if (Cn == 48) {
      image2patches_tiled_kernel<48><<<grid_for(48), 256, 0, ctx.stream>>>(x, H, W, th, tw, (uint16_t*)out.p, out.dt, out.ld);
      BRN_CUDA(cudaGetLastError());
      return;
    }
  }
  long long total = (long long)B * 3 * H * W;
  image2patches_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx.stream>>>(x, H, W, th, tw, out.p, out.dt,
                                                                               out.ld, total);
  BRN_CUDA(cudaGetLastError());
}

__global__ void nchw_to_nhwc_kernel(const float* x, int C, int H, int W, void* out, int odt, int ldo, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % C); long long t = i / C;
  int xx = (int)(t % W); t /= W; int y = (int)(t % H); long long b = t / H;
  g_st(out, odt, ((b * H + y) * (long long)W + xx) * ldo + c, x[((b * C + c) * H + y) * (long long)W + xx]);
}
void glue_nchw_to_nhwc(const LaunchCtx& ctx, const float* x, int B, int C, int H, int W, View out) {
  GLUE_LAUNCH_PROLOGUE(ctx, (double)B * C * H * W * (4 + dsize(out.dt)));
  long long total = (long long)B * C * H * W;
  nchw_to_nhwc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx.stream>>>(x, C, H, W, out.p, out.dt, out.ld, total);
  BRN_CUDA(cudaGetLastError());
}

__global__ void nhwc_to_nchw_kernel(const void* x, int xdt, int ldx, int C, int H, int W, float* out, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int xx = (int)(i % W); long long t = i / W;
  int y = (int)(t % H); t /= H; int c = (int)(t % C); long long b = t / C;
  out[i] = g_ld(x, xdt, ((b * H + y) * (long long)W + xx) * ldx + c);
}
void glue_nhwc_to_nchw(const LaunchCtx& ctx, View in, float* out) {
  GLUE_LAUNCH_PROLOGUE(ctx, (double)in.rows() * in.C * (4 + dsize(in.dt)));
  long long total = (long long)in.B * in.C * in.H * in.W;
  nhwc_to_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx.stream>>>(in.p, in.dt, in.ld, in.C, in.H, in.W,
                                                                              out, total);
  BRN_CUDA(cudaGetLastError());
}

__global__ void copy_cast_kernel(const void* x, int xdt, int ldx, void* out, int odt, int ldo, int C, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  long long row = i / C; int c = (int)(i - row * C);
  g_st(out, odt, row * ldo + c, g_ld(x, xdt, row * ldx + c));
}
void glue_copy_cast(const LaunchCtx& ctx, View in, View out) {
  GLUE_LAUNCH_PROLOGUE(ctx, (double)in.rows() * in.C * (dsize(in.dt) + dsize(out.dt)));
  long long total = in.rows() * in.C;
  copy_cast_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx.stream>>>(in.p, in.dt, in.ld, out.p, out.dt, out.ld,
                                                                           in.C, total);
  BRN_CUDA(cudaGetLastError());
}

// Row statistics + raw 16-bit copy of an fp32 matrix: what an LnEmit epilogue produces (one partial per row), for
// streams that no GEMM epilogue wrote (operator-level entry brn_ln_linear).  One warp per row.
__global__ void __launch_bounds__(256) ln_stats_cast_kernel(const float* __restrict__ x, int ldx, long long rows, int C,
                                                            void* x16, int dt, int ld16, float2* stats) {
  const int lane = threadIdx.x & 31;
  const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= rows) return;
  float s = 0.f, q = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float v = x[m * ldx + c];
    s += v; q = fmaf(v, v, q);
    g_st(x16, dt, m * ld16 + c, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  if (lane == 0) stats[m] = make_float2(s, q);
}
void glue_ln_stats_cast(const LaunchCtx& ctx, View x, View x16, float2* stats) {
  GLUE_LAUNCH_PROLOGUE(ctx, (double)x.rows() * (x.C * 6.0 + 8.0));
  BRN_CHECK(x.dt == F32 && x16.dt != F32, 5, "ln_stats_cast: fp32 in, 16-bit out");
  ln_stats_cast_kernel<<<(unsigned)((x.rows() + 7) / 8), 256, 0, ctx.stream>>>((const float*)x.p, x.ld, x.rows(), x.C, x16.p,
                                                                             x16.dt, x16.ld, stats);
  BRN_CUDA(cudaGetLastError());
}

// The consumer GEMM's epilogue needs one (-mean, rstd) pair per row, one tile ahead of its math: a compact array it can
// prefetch with a single 8-byte load (summing the partials inside the epilogue put an L2 round trip on the critical
// path of every tile: +30 % on the epilogue-bound stage-0 qkv GEMM, r02 run B).
__global__ void __launch_bounds__(256) ln_finalize_kernel(const float2* __restrict__ stats, int parts, long long stride,
                                                          long long rows, float invC, float2* __restrict__ mr) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // programmatic dependent launch, tc_ptx.cuh pdl_*
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const long long m = (long long)blockIdx.x * 256 + threadIdx.x;
  if (m >= rows) return;
  float s = 0.f, q = 0.f;
  for (int i = 0; i < parts; ++i) {
    const float2 t = __ldg(stats + (long long)i * stride + m);
    s += t.x; q += t.y;
  }
  const float mu = s * invC;
  const float var = fmaxf(fmaf(-mu, mu, q * invC), 0.f);
  mr[m] = make_float2(-mu, rsqrtf(var + 1e-5f));
}
void glue_ln_finalize(const LaunchCtx& ctx, const float2* stats, int parts, long long stride, long long rows, int C,
                      float2* mr) {
  if (ctx.launches) ++*ctx.launches;
  if (ctx.dry) return;
  KScope ks(ctx, KC_LN, 0.0, (double)rows * (parts + 1) * 8, "ln_finalize");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((rows + 255) / 256));
  cfg.blockDim = dim3(256);
  cfg.stream = ctx.stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr; cfg.numAttrs = pdl_attr(attr, 0);
  BRN_CUDA(cudaLaunchKernelEx(&cfg, ln_finalize_kernel, stats, parts, stride, rows, 1.0f / (float)C, mr));
  BRN_CUDA(cudaGetLastError());
}

__global__ void sigmoid_kernel(float* p, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 1.f / (1.f + expf(-p[i]));
}
void glue_sigmoid(const LaunchCtx& ctx, float* p, long long n) {
  GLUE_LAUNCH_PROLOGUE(ctx, 8.0 * (double)n);
  sigmoid_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx.stream>>>(p, n);
  BRN_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// Global average pool partial sums (src/aspp.rs:314) and the pooled branch folded to a per-image bias of conv1:
//   x5 = relu(bn(conv1x1(mean)))  (src/aspp.rs:315-317), broadcast over H,W (:318), then its 256 channels times
//   conv1.weight[:, 1024:1280] (:327-329) is a constant 64-vector per image.
// ------------------------------------------------------------------------------------------------
// Deterministic two-level reduction (no atomics: a batch must equal its per-image results bit for bit):
// block (blk, b) writes the partial sums of its pixel chunk to part[b][blk][C]; aspp_pool_bias_kernel adds the
// partials in block order.
constexpr int GAP_CHUNK = 1024;
__global__ void __launch_bounds__(256) gap_sum_kernel(const void* x, int xdt, int ldx, int C, int HW, float* part) {
  __shared__ float sh[256];
  const int b = blockIdx.y;
  const int lanes = 256 / C;   // pixel lanes (C divides 256: C = 64)
  const int c = threadIdx.x % C, pl = threadIdx.x / C;
  const int p0 = blockIdx.x * GAP_CHUNK, p1 = min(p0 + GAP_CHUNK, HW);
  float s = 0.f;
  if (pl < lanes)
    for (int px = p0 + pl; px < p1; px += lanes) s += g_ld(x, xdt, ((long long)b * HW + px) * ldx + c);
  sh[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x < C) {
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += sh[l * C + threadIdx.x];
    part[((long long)b * gridDim.x + blockIdx.x) * C + threadIdx.x] = t;
  }
}
// C = 64, 16-bit, 16-byte aligned rows: 8 lanes x 16 bytes cover one pixel, 32 pixel lanes per block, 4 independent
// accumulation chains per thread (same partial-sum layout and block -> chunk mapping as the scalar kernel)
template <int DT>
__global__ void __launch_bounds__(256) gap_sum64_kernel(const uint16_t* __restrict__ x, int ldx, int HW, float* part) {
  __shared__ float sh[32][65];
  const int b = blockIdx.y;
  const int c8 = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int p0 = blockIdx.x * GAP_CHUNK, p1 = min(p0 + GAP_CHUNK, HW);
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  for (int px = p0 + pl; px < p1; px += 32) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + ((long long)b * HW + px) * ldx + c8 * 8));
    const uint32_t* h = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
    for (int t = 0; t < 4; ++t) { const float2 f = up16(h[t], DT); acc[2 * t] += f.x; acc[2 * t + 1] += f.y; }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) sh[pl][c8 * 8 + e] = acc[e];
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
    for (int l = 0; l < 32; ++l) t += sh[l][threadIdx.x];
    part[((long long)b * gridDim.x + blockIdx.x) * 64 + threadIdx.x] = t;
  }
}

int glue_gap_blocks(int HW) { return (HW + GAP_CHUNK - 1) / GAP_CHUNK; }

void glue_gap_sum(const LaunchCtx& ctx, View x, float* part) {
  GLUE_LAUNCH_PROLOGUE(ctx, (double)x.rows() * x.C * dsize(x.dt));
  BRN_CHECK(x.C <= 256 && 256 % x.C == 0, 5, "gap_sum: C must divide 256");
  const int HW = x.H * x.W;
  dim3 grid(glue_gap_blocks(HW), x.B);
  if (x.C == 64 && x.dt != F32 && x.ld % 8 == 0 && ((uintptr_t)x.p & 15) == 0) {
    if (x.dt == BF16) gap_sum64_kernel<BF16><<<grid, 256, 0, ctx.stream>>>((const uint16_t*)x.p, x.ld, HW, part);
    else gap_sum64_kernel<F16><<<grid, 256, 0, ctx.stream>>>((const uint16_t*)x.p, x.ld, HW, part);
    BRN_CUDA(cudaGetLastError());
    return;
  }
  gap_sum_kernel<<<grid, 256, 0, ctx.stream>>>(x.p, x.dt, x.ld, x.C, HW, part);
  BRN_CUDA(cudaGetLastError());
}

__global__ void __launch_bounds__(256) aspp_pool_bias_kernel(const float* part, int nblk, int HW, const float* wg /*[256][64]*/,
                                                             const float* bg /*[256]*/, const float* tail /*[64][256]*/,
                                                             const float* shift /*[64]*/, float* out) {
  __shared__ float mean[64];
  __shared__ float x5[256];
  const int b = blockIdx.x, t = threadIdx.x;
  if (t < 64) {
    float s = 0.f;
    for (int k = 0; k < nblk; ++k) s += part[((long long)b * nblk + k) * 64 + t];
    mean[t] = s / (float)HW;
  }
  __syncthreads();
  // both mat-vecs with coalesced weight rows and a fixed shuffle tree per output (one block per image: at batch 1 this
  // kernel is pure latency, and a thread-per-output loop walks 64 / 256 strided loads in sequence)
  const int warp = t >> 5, lane = t & 31;
  const float m0 = mean[lane], m1 = mean[lane + 32];
#pragma unroll 8
  for (int o = warp * 32; o < warp * 32 + 32; ++o) {
    float s = fmaf(wg[o * 64 + lane], m0, wg[o * 64 + lane + 32] * m1);
    s = warp_sum(s);
    if (lane == 0) x5[o] = fmaxf(s + bg[o], 0.f);
  }
  __syncthreads();
  float xv[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) xv[k] = x5[lane + 32 * k];
#pragma unroll
  for (int o = warp * 8; o < warp * 8 + 8; ++o) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s = fmaf(tail[o * 256 + lane + 32 * k], xv[k], s);
    s = warp_sum(s);
    if (lane == 0) out[b * 64 + o] = s + shift[o];
  }
}

void glue_aspp_pool_bias(const LaunchCtx& ctx, const float* part, int B, int HW, const LayerW* gap_conv,
                         const float* conv1_tail, const float* bn1_shift, float* out) {
  GLUE_LAUNCH_PROLOGUE(ctx, (double)B * (glue_gap_blocks(HW) * 64 + 64) * 4.0 + (256.0 * 64 + 64 * 256) * 4.0);
  aspp_pool_bias_kernel<<<B, 256, 0, ctx.stream>>>(part, glue_gap_blocks(HW), HW, gap_conv->w32, gap_conv->bias, conv1_tail,
                                                   bn1_shift, out);
  BRN_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// GDT gate (src/birefnet.rs:327-329): p *= sigmoid(conv1x1_{16->1}(g) + b)
// ------------------------------------------------------------------------------------------------
__global__ void gate_kernel(void* p, int pdt, int ldp, int C, const void* g, int gdt, int ldg, const float* w, float b0,
                            long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  long long row = i / C; int c = (int)(i - row * C);
  float s = b0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s = fmaf(w[j], g_ld(g, gdt, row * ldg + j), s);
  float gate = 1.f / (1.f + expf(-s));
  g_st(p, pdt, row * ldp + c, g_ld(p, pdt, row * ldp + c) * gate);
}
__global__ void __launch_bounds__(256) gate_vec8_kernel(uint16_t* p, int dt, int ldp, int C8, const uint16_t* g, int ldg,
                                                        const float* w, float b0, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long row = i / C8; const int c8 = (int)(i - row * C8);
  const uint4 g0 = __ldg(reinterpret_cast<const uint4*>(g + row * ldg)), g1 = __ldg(reinterpret_cast<const uint4*>(g + row * ldg + 8));
  const uint32_t* gp0 = reinterpret_cast<const uint32_t*>(&g0); const uint32_t* gp1 = reinterpret_cast<const uint32_t*>(&g1);
  float s = b0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 a = up16(gp0[j], dt), b = up16(gp1[j], dt);
    s = fmaf(__ldg(w + 2 * j), a.x, s); s = fmaf(__ldg(w + 2 * j + 1), a.y, s);
    s = fmaf(__ldg(w + 8 + 2 * j), b.x, s); s = fmaf(__ldg(w + 8 + 2 * j + 1), b.y, s);
  }
  const float gate = 1.f / (1.f + expf(-s));
  uint4* pp = reinterpret_cast<uint4*>(p + row * ldp + c8 * 8);
  uint4 v = *pp;
  uint32_t* vp = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) { const float2 f = up16(vp[j], dt); vp[j] = pk16(f.x * gate, f.y * gate, dt); }
  *pp = v;
}

void glue_gate(const LaunchCtx& ctx, View p, View g16, const float* w16, float b0) {
  GLUE_LAUNCH_PROLOGUE(ctx, (double)p.rows() * (2.0 * p.C * dsize(p.dt) + 16.0 * dsize(g16.dt)));
  if (p.dt != F32 && g16.dt == p.dt && p.C % 8 == 0 && p.ld % 8 == 0 && g16.ld % 8 == 0 && g16.C == 16 &&
      (((uintptr_t)p.p | (uintptr_t)g16.p) & 15) == 0) {
    const long long tot = p.rows() * (p.C / 8);
    gate_vec8_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx.stream>>>((uint16_t*)p.p, p.dt, p.ld, p.C / 8,
                                                                           (const uint16_t*)g16.p, g16.ld, w16, b0, tot);
    BRN_CUDA(cudaGetLastError());
    return;
  }
  long long total = p.rows() * p.C;
  gate_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx.stream>>>(p.p, p.dt, p.ld, p.C, g16.p, g16.dt, g16.ld, w16,
                                                                      b0, total);
  BRN_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// 1x1 modulated deformable conv = per-pixel bilinear resample with a learned offset, times the modulator, followed by
// an ordinary 1x1 conv (src/deform_conv.rs:101-215 with k = 1).  With a single tap the gather -> MMA pipeline of
// tc_deform_kernel has nothing to overlap with (its 4 epilogue warps become the pace setter), so the sampling runs
// here (8 lanes x 16 bytes per pixel, same arithmetic as the fused kernel) and the contraction goes to tc_gemm.
// ------------------------------------------------------------------------------------------------
// FUSED: the 1x1 offset / modulator conv (3 outputs: dy, dx, 2*sigmoid(m); src/aspp.rs:58-99) is computed here from the
// pixel's own 64 channels (8 lanes x 8 channels, 16-bit weights, fp32 accumulation, 3-step shuffle reduction) instead of
// a 3-column GEMM launch plus a round trip of its output through HBM.
template <int DT, bool FUSED>
__global__ void __launch_bounds__(256) deform_sample_k1_kernel(const uint16_t* __restrict__ x, int ldx, int B, int H, int W,
                                                               const float* __restrict__ om, int ldom, int om_tiled,
                                                               const uint16_t* __restrict__ omw, int omw_ld,
                                                               const float* __restrict__ omb,
                                                               uint16_t* out, int ldo, long long total) {
  const long long i0 = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long i = FUSED ? min(i0, total - 1) : i0;      // FUSED: every lane takes part in the shuffles
  if (!FUSED && i >= total) return;
  const int l8 = (int)(i & 7); const long long px = i >> 3;
  // pixel -> (b, y, x) in 32-bit arithmetic when the pixel count allows it (always, in the model): the four 64-bit
  // divisions were the larger part of this kernel's instructions
  int xx, y; long long b;
  if (total <= 0x7fffffffll * 8) {
    const uint32_t p32 = (uint32_t)px, t32 = p32 / (uint32_t)W;
    xx = (int)(p32 - t32 * (uint32_t)W);
    const uint32_t b32 = t32 / (uint32_t)H;
    y = (int)(t32 - b32 * (uint32_t)H); b = b32;
  } else {
    xx = (int)(px % W); const long long t = px / W; y = (int)(t % H); b = t / H;
  }
  float dy, dx, mk;
  if (FUSED) {
    const uint4 xv = __ldg(reinterpret_cast<const uint4*>(x + px * ldx + l8 * 8));
    const uint32_t* xw = reinterpret_cast<const uint32_t*>(&xv);
    float acc[3];
#pragma unroll
    for (int n = 0; n < 3; ++n) {
      const uint4 wv = __ldg(reinterpret_cast<const uint4*>(omw + n * omw_ld + l8 * 8));
      const uint32_t* ww = reinterpret_cast<const uint32_t*>(&wv);
      float a = 0.f;
#pragma unroll
      for (int tt = 0; tt < 4; ++tt) {
        const float2 xf = up16(xw[tt], DT), wf = up16(ww[tt], DT);
        a = fmaf(xf.x, wf.x, a); a = fmaf(xf.y, wf.y, a);
      }
      a += __shfl_xor_sync(0xffffffffu, a, 1); a += __shfl_xor_sync(0xffffffffu, a, 2); a += __shfl_xor_sync(0xffffffffu, a, 4);
      acc[n] = a + __ldg(omb + n);
    }
    dy = acc[0]; dx = acc[1]; mk = 2.f / (1.f + __expf(-acc[2]));
    if (i0 >= total) return;
  } else if (om_tiled) {
    const int tiles_x = (W + 15) / 16, tiles_y = (H + 7) / 8;
    const long long tile = (b * tiles_y + (y >> 3)) * tiles_x + (xx >> 4);
    const float* o = om + tile * 3 * 128 + ((y & 7) * 16 + (xx & 15));
    dy = __ldg(o); dx = __ldg(o + 128); mk = __ldg(o + 256);
  } else {
    const float* o = om + px * ldom;
    dy = __ldg(o); dx = __ldg(o + 1); mk = __ldg(o + 2);
  }
  const float py = (float)y + dy, pxf = (float)xx + dx;
  const bool inb = py > -1.f && py < (float)H && pxf > -1.f && pxf < (float)W;
  const float fy = floorf(py), fx = floorf(pxf);
  const int y0 = (int)fy, x0 = (int)fx;
  const float ly = py - fy, lx = pxf - fx;
  const float m0 = inb ? mk : 0.f;
  const float wy0 = (y0 >= 0 && y0 <= H - 1) ? (1.f - ly) * m0 : 0.f, wy1 = (y0 + 1 >= 0 && y0 + 1 <= H - 1) ? ly * m0 : 0.f;
  const float wx0 = (x0 >= 0) ? 1.f - lx : 0.f, wx1 = (x0 + 1 <= W - 1) ? lx : 0.f;
  const int yc0 = min(max(y0, 0), H - 1), yc1 = min(max(y0 + 1, 0), H - 1);
  const int xc0 = min(max(x0, 0), W - 1), xc1 = min(max(x0 + 1, 0), W - 1);
  const uint16_t* xb = x + b * H * W * ldx + l8 * 8;
  const uint4 v00 = __ldg(reinterpret_cast<const uint4*>(xb + ((long long)yc0 * W + xc0) * ldx));
  const uint4 v01 = __ldg(reinterpret_cast<const uint4*>(xb + ((long long)yc0 * W + xc1) * ldx));
  const uint4 v10 = __ldg(reinterpret_cast<const uint4*>(xb + ((long long)yc1 * W + xc0) * ldx));
  const uint4 v11 = __ldg(reinterpret_cast<const uint4*>(xb + ((long long)yc1 * W + xc1) * ldx));
  const float w[4] = {wy0 * wx0, wy0 * wx1, wy1 * wx0, wy1 * wx1};
  const uint4* v[4] = {&v00, &v01, &v10, &v11};
  uint4 r;
  uint32_t* ro = reinterpret_cast<uint32_t*>(&r);
  if (DT == F16) {      // same packed half2 blend as tc_deform_kernel (HMUL2 + 3 HFMA2 per channel pair)
    __half2 wh[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) wh[c] = __float2half2_rn(w[c]);
#pragma unroll
    for (int tt = 0; tt < 4; ++tt) {
      __half2 a = __hmul2(*reinterpret_cast<const __half2*>(&reinterpret_cast<const uint32_t*>(v[0])[tt]), wh[0]);
#pragma unroll
      for (int c = 1; c < 4; ++c) a = __hfma2(*reinterpret_cast<const __half2*>(&reinterpret_cast<const uint32_t*>(v[c])[tt]), wh[c], a);
      ro[tt] = *reinterpret_cast<uint32_t*>(&a);
    }
  } else {
#pragma unroll
    for (int tt = 0; tt < 4; ++tt) {
      float ax = 0.f, ay = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) { const float2 f = up16(reinterpret_cast<const uint32_t*>(v[c])[tt], DT); ax = fmaf(w[c], f.x, ax); ay = fmaf(w[c], f.y, ay); }
      ro[tt] = pk16(ax, ay, DT);
    }
  }
  *reinterpret_cast<uint4*>(out + px * ldo + l8 * 8) = r;
}

void glue_deform_sample_k1(const LaunchCtx& ctx, View x, View om, int om_tiled, const LayerW* om_layer, View out) {
  GLUE_LAUNCH_PROLOGUE(ctx, (double)x.rows() * 64 * (dsize(x.dt) + dsize(out.dt)) + (om_layer ? 0.0 : (double)x.rows() * 12));
  BRN_CHECK(x.C == 64 && x.dt != F32 && out.dt == x.dt && x.ld % 8 == 0 && out.ld % 8 == 0, 5, "deform_sample_k1: layout");
  const long long total = x.rows() * 8;
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (om_layer) {
    BRN_CHECK(om_layer->N == 3 && om_layer->Cin == 64 && om_layer->taps() == 1 && om_layer->bias && om_layer->w16(x.dt), 5,
              "deform_sample_k1: fused offset conv is 64 -> 3, 1x1, with bias");
    const uint16_t* w = (const uint16_t*)om_layer->w16(x.dt);
    if (x.dt == BF16)
      deform_sample_k1_kernel<BF16, true><<<grid, 256, 0, ctx.stream>>>((const uint16_t*)x.p, x.ld, x.B, x.H, x.W, nullptr, 0, 0,
          w, om_layer->cin_pad, om_layer->bias, (uint16_t*)out.p, out.ld, total);
    else
      deform_sample_k1_kernel<F16, true><<<grid, 256, 0, ctx.stream>>>((const uint16_t*)x.p, x.ld, x.B, x.H, x.W, nullptr, 0, 0,
          w, om_layer->cin_pad, om_layer->bias, (uint16_t*)out.p, out.ld, total);
  } else if (x.dt == BF16) {
    deform_sample_k1_kernel<BF16, false><<<grid, 256, 0, ctx.stream>>>((const uint16_t*)x.p, x.ld, x.B, x.H, x.W,
        (const float*)om.p, om.ld, om_tiled, nullptr, 0, nullptr, (uint16_t*)out.p, out.ld, total);
  } else {
    deform_sample_k1_kernel<F16, false><<<grid, 256, 0, ctx.stream>>>((const uint16_t*)x.p, x.ld, x.B, x.H, x.W,
        (const float*)om.p, om.ld, om_tiled, nullptr, 0, nullptr, (uint16_t*)out.p, out.ld, total);
  }
  BRN_CUDA(cudaGetLastError());
}

// out[row] = sum_c w[c] * p[row, c]   (the p1 half of conv_out1)
__global__ void __launch_bounds__(256) dot1_kernel(const void* p, int pdt, int ldp, int C, const float* w, float* out,
                                                   long long rows) {
  const int lane = threadIdx.x & 31;
  const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= rows) return;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s = fmaf(w[c], g_ld(p, pdt, m * ldp + c), s);
  s = warp_sum(s);
  if (lane == 0) out[m] = s;
}
void glue_dot1(const LaunchCtx& ctx, View p, const float* w, float* out) {
  GLUE_LAUNCH_PROLOGUE(ctx, (double)p.rows() * (p.C * dsize(p.dt) + 4.0));
  long long rows = p.rows();
  dot1_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, ctx.stream>>>(p.p, p.dt, p.ld, p.C, w, out, rows);
  BRN_CUDA(cudaGetLastError());
}

}  // namespace brn
