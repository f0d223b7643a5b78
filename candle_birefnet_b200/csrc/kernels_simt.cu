// SIMT fp32-FMA kernels: the "fp32 path" of the north star (max |dlogit| <= 1e-3) and the on-device
// cross-check for every tcgen05 kernel.  Plain tiled implicit GEMM, one-row-per-thread window attention.
#include "brn_common.h"
#include "device_utils.cuh"

namespace brn {

// ------------------------------------------------------------------------------------------------
// Implicit GEMM: M = B*H*W output pixels (stride 1, "same" padding), N = Cout, K = taps*Cin (tap-major)
// ------------------------------------------------------------------------------------------------
struct SimtGemmP {
  const void* x; int xdt; int B, H, W, Cin, ldx;
  int kh, kw, pad;
  int stride, Ho, Wo;    // output grid (Ho = (H + 2 pad - kh) / stride + 1); the model itself only uses stride 1, "same"
  const float* w;        // [N][taps][Cin]
  const float* bias; int bias_bstride;
  int N; int act; int act_from;
  const void* res; int resdt; int ldres;
  void* out; int odt; int ldo;
  RowMap rm;
  // deformable sampling (A element = modulator * bilinear(x))
  const float* om; int ldom; int deform;
  long long M;
};

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16;

__device__ __forceinline__ float ld_act(const void* p, int dt, long long i) { return ld_elem(p, dt, i); }
__device__ __forceinline__ void st_act(void* p, int dt, long long i, float v) { st_elem(p, dt, i, v); }

// torchvision deform_conv2d bilinear_interpolate semantics (zero outside (-1,H)x(-1,W))
__device__ __forceinline__ float deform_sample(const SimtGemmP& p, long long base_b, float py, float px, int c) {
  if (py <= -1.f || py >= (float)p.H || px <= -1.f || px >= (float)p.W) return 0.f;
  int y0 = (int)floorf(py), x0 = (int)floorf(px);
  int y1 = y0 + 1, x1 = x0 + 1;
  float ly = py - y0, lx = px - x0, hy = 1.f - ly, hx = 1.f - lx;
  float v1 = 0, v2 = 0, v3 = 0, v4 = 0;
  if (y0 >= 0 && x0 >= 0) v1 = ld_act(p.x, p.xdt, base_b + ((long long)y0 * p.W + x0) * p.ldx + c);
  if (y0 >= 0 && x1 <= p.W - 1) v2 = ld_act(p.x, p.xdt, base_b + ((long long)y0 * p.W + x1) * p.ldx + c);
  if (y1 <= p.H - 1 && x0 >= 0) v3 = ld_act(p.x, p.xdt, base_b + ((long long)y1 * p.W + x0) * p.ldx + c);
  if (y1 <= p.H - 1 && x1 <= p.W - 1) v4 = ld_act(p.x, p.xdt, base_b + ((long long)y1 * p.W + x1) * p.ldx + c);
  return hy * hx * v1 + hy * lx * v2 + ly * hx * v3 + ly * lx * v4;
}

__global__ void __launch_bounds__(256) simt_gemm_kernel(SimtGemmP p) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Ws[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const long long m0 = (long long)blockIdx.x * SG_BM;
  const int n0 = blockIdx.y * SG_BN;
  const int K = p.kh * p.kw * p.Cin;

  // loader coordinates: thread loads 4 consecutive k of one row (A) / one n (W)
  const int lrow = tid / 4, lk = (tid % 4) * 4;
  const long long am = m0 + lrow;
  const bool arow_ok = am < p.M;
  int ab = 0, ay = 0, ax = 0;
  if (arow_ok) {
    long long hw = (long long)p.Ho * p.Wo;
    ab = (int)(am / hw);
    int r = (int)(am % hw);
    ay = (r / p.Wo) * p.stride; ax = (r % p.Wo) * p.stride;      // top-left input coordinate before padding
  }
  const long long abase = (long long)ab * p.H * p.W * p.ldx;
  const int wn = n0 + lrow;

  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += SG_BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int k = k0 + lk + i;
      float av = 0.f, wv = 0.f;
      if (k < K) {
        int tap = k / p.Cin, c = k - tap * p.Cin;
        if (arow_ok) {
          int ky = tap / p.kw, kx = tap - ky * p.kw;
          if (p.deform) {
            const float* o = p.om + am * p.ldom;
            float py = (float)(ay - p.pad + ky) + o[2 * tap];
            float px = (float)(ax - p.pad + kx) + o[2 * tap + 1];
            av = o[2 * p.kh * p.kw + tap] * deform_sample(p, abase, py, px, c);
          } else {
            int iy = ay - p.pad + ky, ix = ax - p.pad + kx;
            if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
              av = ld_act(p.x, p.xdt, abase + ((long long)iy * p.W + ix) * p.ldx + c);
          }
        }
        if (wn < p.N) wv = p.w[(long long)wn * K + k];
      }
      As[lk + i][lrow] = av;
      Ws[lk + i][lrow] = wv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Ws[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    long long orow = m;
    if (p.rm.enabled) {
      orow = rowmap_token(p.rm, m);
      if (orow < 0) continue;
    }
    int b = p.bias_bstride ? (int)(m / ((long long)p.Ho * p.Wo)) : 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[(long long)b * p.bias_bstride + n];
      v = apply_act(v, p.act, n, p.act_from);
      if (p.res) v += ld_act(p.res, p.resdt, orow * p.ldres + n);
      st_act(p.out, p.odt, orow * p.ldo + n, v);
    }
  }
}

static void launch_simt_gemm(const LaunchCtx& ctx, SimtGemmP& p) {
  if (ctx.launches) ++*ctx.launches;
  if (ctx.dry) return;
  KScope ks(ctx, KC_GEMM_SIMT, 2.0 * (double)p.M * p.N * p.kh * p.kw * p.Cin);
  dim3 grid((unsigned)((p.M + SG_BM - 1) / SG_BM), (unsigned)((p.N + SG_BN - 1) / SG_BN));
  simt_gemm_kernel<<<grid, 256, 0, ctx.stream>>>(p);
  BRN_CUDA(cudaGetLastError());
}

void simt_gemm(const LaunchCtx& ctx, const GemmArgs& a) {
  SimtGemmP p{};
  p.x = a.x.p; p.xdt = a.x.dt; p.B = a.x.B; p.H = a.x.H; p.W = a.x.W; p.Cin = a.x.C; p.ldx = a.x.ld;
  BRN_CHECK(a.w && a.w->Cin == a.x.C, 5, "simt_gemm: weight/input channel mismatch");
  BRN_CHECK(a.rowmap.enabled != 2 && !a.lnf.mr && !a.lne.stats, 7,
            "simt_gemm: the folded-LayerNorm epilogues and the token->window row map are tensor-core-path features");
  BRN_CHECK(a.w->w32 != nullptr, 7, "simt_gemm: layer has no fp32 weights (a folded tensor-core-only copy)");
  p.kh = a.w->kh; p.kw = a.w->kw; p.pad = a.pad;
  p.stride = a.stride > 0 ? a.stride : 1;
  p.Ho = (a.x.H + 2 * a.pad - a.w->kh) / p.stride + 1; p.Wo = (a.x.W + 2 * a.pad - a.w->kw) / p.stride + 1;
  BRN_CHECK(p.Ho > 0 && p.Wo > 0 && (a.rowmap.enabled || (long long)a.x.B * p.Ho * p.Wo == a.out.rows()), 5,
            "simt_gemm: output grid mismatch");
  p.w = a.w->w32;
  p.bias = a.bias ? a.bias : a.w->bias; p.bias_bstride = a.bias_bstride;
  p.N = a.w->N; p.act = a.act; p.act_from = a.act_from;
  p.res = a.res.p; p.resdt = a.res.dt; p.ldres = a.res.ld;
  p.out = a.out.p; p.odt = a.out.dt; p.ldo = a.out.ld;
  p.rm = a.rowmap;
  p.om = nullptr; p.ldom = 0; p.deform = 0;
  p.M = (long long)a.x.B * p.Ho * p.Wo;
  launch_simt_gemm(ctx, p);
}

void simt_deform(const LaunchCtx& ctx, const DeformArgs& a) {
  SimtGemmP p{};
  p.x = a.x.p; p.xdt = a.x.dt; p.B = a.x.B; p.H = a.x.H; p.W = a.x.W; p.Cin = a.x.C; p.ldx = a.x.ld;
  p.kh = a.w->kh; p.kw = a.w->kw; p.pad = a.pad >= 0 ? a.pad : a.w->kh / 2;
  p.stride = a.stride > 0 ? a.stride : 1;
  p.Ho = (a.x.H + 2 * p.pad - a.w->kh) / p.stride + 1; p.Wo = (a.x.W + 2 * p.pad - a.w->kw) / p.stride + 1;
  BRN_CHECK(p.Ho > 0 && p.Wo > 0 && (long long)a.x.B * p.Ho * p.Wo == a.out.rows(), 5, "simt_deform: output grid mismatch");
  BRN_CHECK(a.w->w32 != nullptr, 7, "simt_deform: layer has no fp32 weights");
  p.w = a.w->w32;
  p.bias = a.bias ? a.bias : a.w->bias; p.bias_bstride = 0;
  p.N = a.w->N; p.act = a.act; p.act_from = 0;
  p.res = nullptr; p.resdt = 0; p.ldres = 0;
  p.out = a.out.p; p.odt = a.out.dt; p.ldo = a.out.ld;
  p.om = (const float*)a.om.p; p.ldom = a.om.ld; p.deform = 1;
  BRN_CHECK(a.om.dt == F32 && !a.om_tiled, 5, "simt_deform: offsets must be fp32, pixel-major");
  p.M = a.out.rows();
  launch_simt_gemm(ctx, p);
}

// ------------------------------------------------------------------------------------------------
// Window attention, one (window, head) per CTA, one query row per thread, online softmax in fp32.
// Follows WindowAttention::forward_standard (src/swin.rs:266-311): S = q k^T (q pre-scaled) + bias (+ -100 mask).
// ------------------------------------------------------------------------------------------------
struct SimtAttnP {
  const void* qkv; int dt; int ldq; int C;
  const float* bias;
  int heads, nwh, nww, shift;
  void* out; int odt; int ldo;
};

template <int WS>
__global__ void __launch_bounds__(160) simt_attn_kernel(SimtAttnP p) {
  constexpr int N = WS * WS;                  // tokens per window: 144 (swin_b / swin_l) or 49 (swin_t / swin_s)
  constexpr int HALF = WS - WS / 2;           // first local row / column of the shifted-in region (src/swin.rs:608-629)
  __shared__ float Ks[N][33];
  __shared__ float Vs[N][33];
  const int win = blockIdx.x, head = blockIdx.y;
  const int tid = threadIdx.x;
  const long long row0 = (long long)win * N;
  for (int i = tid; i < N * 32; i += blockDim.x) {
    int r = i / 32, d = i % 32;
    Ks[r][d] = ld_act(p.qkv, p.dt, (row0 + r) * p.ldq + p.C + head * 32 + d);
    Vs[r][d] = ld_act(p.qkv, p.dt, (row0 + r) * p.ldq + 2 * p.C + head * 32 + d);
  }
  __syncthreads();
  if (tid >= N) return;
  float q[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) q[d] = ld_act(p.qkv, p.dt, (row0 + tid) * p.ldq + head * 32 + d);
  const int nw = p.nwh * p.nww;
  const int wi = (win % nw) / p.nww, wj = (win % nw) % p.nww;
  const bool last_r = p.shift > 0 && wi == p.nwh - 1, last_c = p.shift > 0 && wj == p.nww - 1;
  const int qi = tid / WS, qj = tid % WS;
  const float* brow = p.bias + ((long long)head * N + tid) * N;
  float mx = -INFINITY, sum = 0.f, o[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) o[d] = 0.f;
  for (int k = 0; k < N; ++k) {
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) s = fmaf(q[d], Ks[k][d], s);
    s += brow[k];
    int ki = k / WS, kj = k % WS;
    if ((last_r && ((ki >= HALF) != (qi >= HALF))) || (last_c && ((kj >= HALF) != (qj >= HALF)))) s += -100.0f;
    float nm = fmaxf(mx, s);
    float corr = __expf(mx - nm), e = __expf(s - nm);
    sum = sum * corr + e;
#pragma unroll
    for (int d = 0; d < 32; ++d) o[d] = fmaf(o[d], corr, e * Vs[k][d]);
    mx = nm;
  }
  float inv = 1.f / sum;
#pragma unroll
  for (int d = 0; d < 32; ++d) st_act(p.out, p.odt, (row0 + tid) * p.ldo + head * 32 + d, o[d] * inv);
}

void simt_attention(const LaunchCtx& ctx, const AttnArgs& a) {
  if (ctx.launches) ++*ctx.launches;
  if (ctx.dry) return;
  BRN_CHECK(a.split_win == 0, 7, "simt_attention: merged two-grid passes are a tensor-core-path feature");
  SimtAttnP p{};
  p.qkv = a.qkv.p; p.dt = a.qkv.dt; p.ldq = a.qkv.ld; p.C = a.heads * 32;
  p.bias = a.bias32; p.heads = a.heads; p.nwh = a.nwh; p.nww = a.nww; p.shift = a.shift;
  p.out = a.out.p; p.odt = a.out.dt; p.ldo = a.out.ld;
  dim3 grid(a.n_windows, a.heads);
  BRN_CHECK(a.ws == 12 || a.ws == 7, 7, "simt_attention: window side must be 7 or 12");
  BRN_CHECK(!a.token_out, 7, "simt_attention: token-order output is a tensor-core-path feature");
  const double n = (double)a.ws * a.ws;
  KScope ks(ctx, KC_ATTN_SIMT, 4.0 * n * n * 32 * (double)a.n_windows * a.heads);
  if (a.ws == 12) simt_attn_kernel<12><<<grid, 160, 0, ctx.stream>>>(p);
  else simt_attn_kernel<7><<<grid, 160, 0, ctx.stream>>>(p);
  BRN_CUDA(cudaGetLastError());
}

}  // namespace brn
