// Micro-benchmark: does the ORDER in which a persistent CTA touches a 128 x 256 fp32 tile of a [M, 768] residual stream
// matter for the HBM throughput of the read-modify-write (+ 16-bit copy) the proj / fc2 GEMM epilogues perform?
//   P0  linear streaming (grid-stride float4): the ceiling for this read / write mix
//   P1  epilogue order: warp = 32 rows x 128 columns, walked in 16-column granules (8 rows x 64 B per instruction),
//       four granules of loads in flight -- what tc_epilogue.cuh does
//   P2  same tiles, same CTA / warp count, but a warp instruction covers ONE row x 512 B (row-contiguous order)
//   P3  like P1 with 32-column granules (8 rows x 128 B per instruction: full lines)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o rmw_pattern rmw_pattern.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int N = 768, BM = 128, BN = 256;

__global__ void __launch_bounds__(256) p0_linear(float4* x, uint2* x16, long long n4) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    float4 v = x[i];
    v.x += 1.f; v.y += 1.f; v.z += 1.f; v.w += 1.f;
    x[i] = v;
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    x16[i] = make_uint2(*(uint32_t*)&a, *(uint32_t*)&b);
  }
}

template <int GC>   // fp32 columns per granule: 16 (64 B per row) or 32 (128 B per row)
__global__ void __launch_bounds__(256) p1_granules(float* x, __half* x16, int m_tiles) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, part = warp >> 2;
  constexpr int LPR = GC / 4;            // lanes per row
  constexpr int RPI = 32 / LPR;          // rows per instruction
  constexpr int NI = 32 / RPI;           // instructions per granule
  constexpr int NG = 128 / GC;           // granules per warp and tile
  const int n_tiles = N / BN, items = m_tiles * n_tiles;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int mt = item / n_tiles, nt = item - mt * n_tiles;
    const long long row0 = (long long)mt * BM + q * 32;
    const int col0 = nt * BN + part * 128 + (lane % LPR) * 4;
    float4 r[4][NI];
    auto issue = [&](int g, float4 (&d)[NI]) {
#pragma unroll
      for (int it = 0; it < NI; ++it)
        d[it] = *reinterpret_cast<const float4*>(x + (row0 + it * RPI + lane / LPR) * N + col0 + g * GC);
    };
    auto finish = [&](int g, float4 (&d)[NI]) {
#pragma unroll
      for (int it = 0; it < NI; ++it) {
        float4 v = d[it];
        v.x += 1.f; v.y += 1.f; v.z += 1.f; v.w += 1.f;
        const long long o = (row0 + it * RPI + lane / LPR) * N + col0 + g * GC;
        *reinterpret_cast<float4*>(x + o) = v;
        __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
        *reinterpret_cast<uint2*>(x16 + o) = make_uint2(*(uint32_t*)&a, *(uint32_t*)&b);
      }
    };
    constexpr int D = NG < 4 ? NG : 4;
#pragma unroll
    for (int g = 0; g < D; ++g) issue(g, r[g]);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      finish(g, r[g % D]);
      if (g + D < NG) issue(g + D, r[g % D]);
    }
  }
}

__global__ void __launch_bounds__(256) p2_rows(float* x, __half* x16, int m_tiles) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, part = warp >> 2;
  const int n_tiles = N / BN, items = m_tiles * n_tiles;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int mt = item / n_tiles, nt = item - mt * n_tiles;
    const long long row0 = (long long)mt * BM + q * 32;
    const int col0 = nt * BN + part * 128 + lane * 4;      // 32 lanes x 16 B = the warp's 512 B of one row
    float4 r[4][4];
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int it = 0; it < 4; ++it) r[g][it] = *reinterpret_cast<const float4*>(x + (row0 + g * 4 + it) * N + col0);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        float4 v = r[g % 4][it];
        v.x += 1.f; v.y += 1.f; v.z += 1.f; v.w += 1.f;
        const long long o = (row0 + g * 4 + it) * N + col0;
        *reinterpret_cast<float4*>(x + o) = v;
        __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
        *reinterpret_cast<uint2*>(x16 + o) = make_uint2(*(uint32_t*)&a, *(uint32_t*)&b);
      }
      if (g + 4 < 8)
#pragma unroll
        for (int it = 0; it < 4; ++it) r[g % 4][it] = *reinterpret_cast<const float4*>(x + (row0 + (g + 4) * 4 + it) * N + col0);
    }
  }
}

int main() {
  const int M = 81920, m_tiles = M / BM;
  float* x; __half* x16;
  cudaMalloc(&x, (size_t)M * N * 4); cudaMalloc(&x16, (size_t)M * N * 2);
  cudaMemset(x, 0, (size_t)M * N * 4);
  float* flush; cudaMalloc(&flush, 512u << 20);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double bytes = (double)M * N * (4 + 4 + 2);
  auto run = [&](const char* name, auto launch) {
    float best = 1e9f, sum = 0.f;
    for (int i = 0; i < 12; ++i) {
      cudaMemsetAsync(flush, i, 512u << 20);       // evict the stream from L2
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (i >= 2) { best = ms < best ? ms : best; sum += ms; }
    }
    printf("%-28s best %7.1f us  mean %7.1f us  %6.0f GB/s (mean)\n", name, best * 1e3, sum / 10 * 1e3, bytes / (sum / 10) / 1e6);
  };
  run("P0 linear x16 CTAs/SM", [&] { p0_linear<<<148 * 16, 256>>>((float4*)x, (uint2*)x16, (long long)M * N / 4); });
  run("P0 linear x1 CTA/SM", [&] { p0_linear<<<148, 256>>>((float4*)x, (uint2*)x16, (long long)M * N / 4); });
  run("P1 granules 16 col (64 B)", [&] { p1_granules<16><<<148, 256>>>(x, x16, m_tiles); });
  run("P3 granules 32 col (128 B)", [&] { p1_granules<32><<<148, 256>>>(x, x16, m_tiles); });
  run("P2 row-contiguous 512 B", [&] { p2_rows<<<148, 256>>>(x, x16, m_tiles); });
  run("P1 x2 CTAs/SM", [&] { p1_granules<16><<<296, 256>>>(x, x16, m_tiles); });
  run("P2 x2 CTAs/SM", [&] { p2_rows<<<296, 256>>>(x, x16, m_tiles); });
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
