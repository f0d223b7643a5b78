// Fused final layer of BiRefNetDecoder::forward (src/birefnet.rs:320,372-375 + SimpleConvs, src/decoder.rs:50-56):
//
//   logits = conv_out1( cat( up(p1), ipt_blk1(x) ) )
//          = up( w_p . p1 )  +  sum_o w_i[o] * conv_out_o( conv1(x) )  +  b          (bilinear up-sampling is linear)
//
// ipt_blk1 is two stacked 3x3 convs with NO activation in between, so for a pixel whose 3x3 neighbourhood lies inside
// the image the second term collapses exactly to one 5x5 convolution of the 3-channel input (75 MACs instead of
// 9*64*27 + 64*9 = 16,128).  On the 1-pixel image border the intermediate 64-channel map is ZERO-padded (it is not
// the conv of a padded input), so there the sum runs over the valid intermediate taps only:
//   out(y,x) = b + up(q)(y,x) + sum_{(i,j) valid} ( Bt[i,j] + sum_{ci,ky,kx} M[i,j][ci,ky,kx] x~[ci, y+i+ky-2, x+j+kx-2] )
// with M[i,j] = sum_c Wc[c,i,j] w1[c], Bt[i,j] = sum_c Wc[c,i,j] b1[c], Wc = sum_o w_i[o] conv_out.weight[o] -- all
// folded on the host in double at finalize.  The 240-channel full-resolution tensor (960 MiB fp32 per image) and the
// 64-channel intermediate never exist.  HBM-bound: 12 B in + 4 B out per pixel.
#include "brn_common.h"
#include "device_utils.cuh"

namespace brn {

// table layout (floats): [0,75) K5[ci][u][v] | [75,318) M[ij][ci*9+ky*3+kx] | [318,327) Bt[ij] | 327 b | 328 b + sum Bt
constexpr int FIN_TAB = 336;
constexpr int FT_W = 128, FT_H = 8, FT_PX = 4;          // block = 128 x 8 outputs, FT_PX horizontally adjacent outputs per thread
constexpr int FT_THREADS = (FT_W / FT_PX) * FT_H;
constexpr int FT_LD = FT_W + 8;                          // 136 floats: shared column j holds image column tx0 - 4 + j, so the
                                                         // tile is filled with aligned 16-byte global loads (halo 2, loaded as 4)
constexpr int FT_OFF = 2;                                // shared column of image column (tx0 - 2): first tap of output 0

// the 5x5 kernel and its bias travel in the kernel parameter (constant bank): the unrolled FFMAs read them as
// immediate-offset constant operands instead of one shared-memory load per MAC
struct FinK5 { float k[76]; };

__device__ __forceinline__ float fin_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }
__device__ __forceinline__ void fin_bilin(int dst, int in, float scale, int& i0, int& i1, float& l) {
  float s = scale * dst;
  i0 = min((int)s, in - 1);
  i1 = min(i0 + 1, in - 1);
  l = s - (float)i0;
}

__global__ void __launch_bounds__(FT_THREADS) final_kernel(const float* __restrict__ x, int H, int W, const FinK5 k5,
                                                           const float* __restrict__ tab, const float* __restrict__ q,
                                                           int qh, int qw, float* __restrict__ out, int apply_sigmoid) {
  __shared__ __align__(16) float xin[3][FT_H + 4][FT_LD];
  __shared__ float st[FIN_TAB];
  const int b = blockIdx.z, ty0 = blockIdx.y * FT_H, tx0 = blockIdx.x * FT_W, tid = threadIdx.x;
  for (int i = tid; i < FIN_TAB; i += FT_THREADS) st[i] = tab[i];
  if ((W & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    // 16-byte loads: a float4 at image column tx0 - 4 + 4 k is either completely inside or completely outside the row
    constexpr int V = FT_LD / 4;
    for (int i = tid; i < 3 * (FT_H + 4) * V; i += FT_THREADS) {
      const int c = i / ((FT_H + 4) * V), r = i % ((FT_H + 4) * V), yy = r / V, k = r % V;
      const int gy = ty0 + yy - 2, gx = tx0 - 4 + 4 * k;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(reinterpret_cast<const float4*>(x + ((long long)(b * 3 + c) * H + gy) * W + gx));
      *reinterpret_cast<float4*>(&xin[c][yy][4 * k]) = v;
    }
  } else {
    for (int i = tid; i < 3 * (FT_H + 4) * FT_LD; i += FT_THREADS) {
      const int c = i / ((FT_H + 4) * FT_LD), r = i % ((FT_H + 4) * FT_LD), yy = r / FT_LD, xx = r % FT_LD;
      const int gy = ty0 + yy - 2, gx = tx0 + xx - 4;
      xin[c][yy][xx] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(x + ((long long)(b * 3 + c) * H + gy) * W + gx) : 0.f;
    }
  }
  __syncthreads();
  const int ly = tid / (FT_W / FT_PX), lx = (tid % (FT_W / FT_PX)) * FT_PX;
  const int gy = ty0 + ly, gx0 = tx0 + lx;
  if (gy >= H || gx0 >= W) return;
  int y0, y1; float fy;
  const float sc_y = fin_scale(qh, H), sc_x = fin_scale(qw, W);     // one division each per thread, not per output
  fin_bilin(gy, qh, sc_y, y0, y1, fy);
  const float* qb = q + (long long)b * qh * qw;
  auto finish = [&](int gx, float a) {       // + up(q), optional sigmoid
    int x0, x1; float fx;
    fin_bilin(gx, qw, sc_x, x0, x1, fx);
    const float up = (1.f - fy) * ((1.f - fx) * __ldg(qb + y0 * qw + x0) + fx * __ldg(qb + y0 * qw + x1)) +
                     fy * ((1.f - fx) * __ldg(qb + y1 * qw + x0) + fx * __ldg(qb + y1 * qw + x1));
    float v = a + up;
    if (apply_sigmoid) v = 1.f / (1.f + expf(-v));
    return v;
  };
  float* op = out + ((long long)b * H + gy) * W + gx0;
  if (gy >= 1 && gy <= H - 2 && gx0 >= 1 && gx0 + FT_PX - 1 <= W - 2 && (W & 3) == 0) {
    // interior: one 5x5x3 convolution per output; a thread's FT_PX outputs share each 8-float input row segment
    float acc[FT_PX];
#pragma unroll
    for (int e = 0; e < FT_PX; ++e) acc[e] = k5.k[75];
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        // image columns gx0 - 2 .. gx0 + 5 = shared columns lx + 2 .. lx + 9
        const float4 r0 = *reinterpret_cast<const float4*>(&xin[ci][ly + u][lx]);
        const float4 r1 = *reinterpret_cast<const float4*>(&xin[ci][ly + u][lx + 4]);
        const float4 r2 = *reinterpret_cast<const float4*>(&xin[ci][ly + u][lx + 8]);
        const float row[8] = {r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y};
#pragma unroll
        for (int v = 0; v < 5; ++v)
#pragma unroll
          for (int e = 0; e < FT_PX; ++e) acc[e] = fmaf(k5.k[ci * 25 + u * 5 + v], row[e + v], acc[e]);
      }
    *reinterpret_cast<float4*>(op) = make_float4(finish(gx0, acc[0]), finish(gx0 + 1, acc[1]), finish(gx0 + 2, acc[2]),
                                                 finish(gx0 + 3, acc[3]));
    return;
  }
  // image border (and the odd-width fallback): per-output path with the valid intermediate taps only
  for (int e = 0; e < FT_PX; ++e) {
    const int gx = gx0 + e;
    if (gx >= W) break;
    float a;
    if (gy >= 1 && gy <= H - 2 && gx >= 1 && gx <= W - 2) {
      a = st[328];
      for (int ci = 0; ci < 3; ++ci)
        for (int u = 0; u < 5; ++u)
          for (int v = 0; v < 5; ++v) a = fmaf(st[ci * 25 + u * 5 + v], xin[ci][ly + u][lx + FT_OFF + e + v], a);
    } else {
      a = st[327];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          const int py = gy + i - 1, px = gx + j - 1;
          if (py < 0 || py >= H || px < 0 || px >= W) continue;
          const float* m = &st[75 + (i * 3 + j) * 27];
          float s = st[318 + i * 3 + j];
          for (int ci = 0; ci < 3; ++ci)
            for (int ky = 0; ky < 3; ++ky)
              for (int kx = 0; kx < 3; ++kx) s = fmaf(m[ci * 9 + ky * 3 + kx], xin[ci][ly + i + ky][lx + FT_OFF + e + j + kx], s);
          a += s;
        }
    }
    op[e] = finish(gx, a);
  }
}

void glue_final(const LaunchCtx& ctx, const float* x, int B, int H, int W, const float* tab, const float* tab_host,
                const float* q, int qh, int qw, float* out, int apply_sigmoid) {
  if (ctx.launches) ++*ctx.launches;
  if (ctx.dry) return;
  KScope ks(ctx, KC_GLUE, 2.0 * 75 * (double)B * H * W, 16.0 * B * H * W, "final");
  FinK5 k5;
  for (int i = 0; i < 75; ++i) k5.k[i] = tab_host[i];
  k5.k[75] = tab_host[328];
  dim3 grid((W + FT_W - 1) / FT_W, (H + FT_H - 1) / FT_H, B);
  final_kernel<<<grid, FT_THREADS, 0, ctx.stream>>>(x, H, W, k5, tab, q, qh, qw, out, apply_sigmoid);
  BRN_CUDA(cudaGetLastError());
}

// host-side fold (double): w1 [64][27], b1 [64], wc [64][9] (c; i*3+j), bc -> table
void build_final_table(const float* w1, const float* b1, const double* wc, double bc, float* tab) {
  double M[9][27] = {}, Bt[9] = {}, K5[3][5][5] = {};
  for (int ij = 0; ij < 9; ++ij) {
    for (int c = 0; c < 64; ++c) {
      Bt[ij] += wc[c * 9 + ij] * (double)b1[c];
      for (int t = 0; t < 27; ++t) M[ij][t] += wc[c * 9 + ij] * (double)w1[c * 27 + t];
    }
  }
  double bsum = bc;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      bsum += Bt[i * 3 + j];
      for (int ci = 0; ci < 3; ++ci)
        for (int ky = 0; ky < 3; ++ky)
          for (int kx = 0; kx < 3; ++kx) K5[ci][i + ky][j + kx] += M[i * 3 + j][ci * 9 + ky * 3 + kx];
    }
  for (int i = 0; i < FIN_TAB; ++i) tab[i] = 0.f;
  for (int ci = 0; ci < 3; ++ci)
    for (int u = 0; u < 5; ++u)
      for (int v = 0; v < 5; ++v) tab[ci * 25 + u * 5 + v] = (float)K5[ci][u][v];
  for (int ij = 0; ij < 9; ++ij) {
    for (int t = 0; t < 27; ++t) tab[75 + ij * 27 + t] = (float)M[ij][t];
    tab[318 + ij] = (float)Bt[ij];
  }
  tab[327] = (float)bc;
  tab[328] = (float)bsum;
}

}  // namespace brn
