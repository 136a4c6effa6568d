// tc_gather.cuh -- the producer-side gather shared by the tensor-core kernels: a quarter-warp (8 lanes x
// 16 B = one 128-byte row segment) walks one list of neighbour rows and reduces NKC consecutive
// 32-channel panels of those rows in registers.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace c3p {
namespace tc {

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// acc[kc] = sum_m w_m * src[list[m], col0 + kc*32 + l8*4 .. +4]   (w_m = 1 when !WEIGHTED).
// Must be called by all 32 lanes (n is uniform inside a quarter-warp, may differ between the four
// quarter-warps): the row ids are fetched 8 at a time, one per lane, and broadcast with shuffles, four
// rows x NKC panels of loads are in flight per lane.
template <int NKC, bool WEIGHTED>
__device__ __forceinline__ void gather_rows(float4 (&acc)[NKC], const float* __restrict__ src, int Csrc,
                                            int col0, const int* __restrict__ rows,
                                            const float* __restrict__ weights, size_t lbase, int n, int l8) {
#pragma unroll
  for (int kc = 0; kc < NKC; ++kc) acc[kc] = make_float4(0.f, 0.f, 0.f, 0.f);
  int nmax = max(n, __shfl_xor_sync(C3P_FULL_MASK, n, 8));
  nmax = max(nmax, __shfl_xor_sync(C3P_FULL_MASK, nmax, 16));
  const int* list = rows + lbase;
  for (int m0 = 0; m0 < nmax; m0 += 8) {
    const int nr = max(0, min(8, n - m0));
    const int nrmax = min(8, nmax - m0);
    const int my_id = l8 < nr ? __ldg(list + m0 + l8) : 0;
    float my_w = 1.f;
    if (WEIGHTED) my_w = l8 < nr ? __ldg(weights + lbase + m0 + l8) : 0.f;
    for (int mb = 0; mb < nrmax; mb += 4) {
      float4 v[4][NKC];
      float wv[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int id = __shfl_sync(C3P_FULL_MASK, my_id, mb + m, 8);
        wv[m] = WEIGHTED ? __shfl_sync(C3P_FULL_MASK, my_w, mb + m, 8) : 1.f;
        const float* p = src + (size_t)id * Csrc + col0 + l8 * 4;
#pragma unroll
        for (int kc = 0; kc < NKC; ++kc)
          v[m][kc] = (mb + m < nr) ? ldg_f4(p + kc * PANEL_K) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int kc = 0; kc < NKC; ++kc) {
          if (WEIGHTED) {
            acc[kc].x = fmaf(wv[m], v[m][kc].x, acc[kc].x); acc[kc].y = fmaf(wv[m], v[m][kc].y, acc[kc].y);
            acc[kc].z = fmaf(wv[m], v[m][kc].z, acc[kc].z); acc[kc].w = fmaf(wv[m], v[m][kc].w, acc[kc].w);
          } else {
            acc[kc].x += v[m][kc].x; acc[kc].y += v[m][kc].y;
            acc[kc].z += v[m][kc].z; acc[kc].w += v[m][kc].w;
          }
        }
    }
  }
}

// Writes the TF32 hi part of v at `dst` and the lo part `lo_offset` bytes further (16-byte chunk).
__device__ __forceinline__ void store_split(unsigned char* dst, uint32_t lo_offset, const float4& v) {
  const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
  const float4 l = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
  *reinterpret_cast<float4*>(dst) = h;
  *reinterpret_cast<float4*>(dst + lo_offset) = l;
}

}  // namespace tc
}  // namespace c3p
