// tc_gather.cuh -- the producer-side gather shared by the tensor-core kernels: a quarter-warp (8 lanes x
// 16 B = one 128-byte row segment) walks one list of neighbour rows and reduces NKC consecutive
// 32-channel panels of those rows in registers.  The walk is split in two so callers can software-pipeline
// it: fetch_slot() issues the (unconditional) loads of the first eight list entries -- nothing depends on
// them until gather_slot() runs, so it is a true prefetch -- and gather_slot() streams the rows.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace c3p {
namespace tc {

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

struct GatherSlot {
  int n;        // members of the cell (0: empty)
  uint32_t lb;  // position of the cell's first entry in the list arrays
  int ids;      // lane l8: row id of entry l8 (lanes past the end re-read entry 0)
  float w;      // lane l8: weight of entry l8 (WEIGHTED lists only)
};

template <bool WEIGHTED>
__device__ __forceinline__ GatherSlot fetch_slot(int n, uint32_t lb, const int* __restrict__ rows,
                                                 const float* __restrict__ weights, int l8) {
  GatherSlot d;
  d.n = n;
  d.lb = lb;
  // select on the ADDRESS: the loaded values have no consumer until the slot is gathered
  const uint32_t at = n > 0 ? lb + (uint32_t)min(l8, n - 1) : 0u;
  d.ids = __ldg(rows + at);
  d.w = 1.f;
  if (WEIGHTED) d.w = __ldg(weights + at);
  return d;
}

template <int NKC, bool WEIGHTED>
__device__ __forceinline__ void load4(float4 (&v)[4][NKC], float (&wv)[4], int ids, float w, int m0, int n,
                                      const float* __restrict__ src, int Csrc, int col, int l8) {
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int id = __shfl_sync(C3P_FULL_MASK, ids, (m0 + m) & 7, 8);
    wv[m] = WEIGHTED ? __shfl_sync(C3P_FULL_MASK, w, (m0 + m) & 7, 8) : 1.f;
    const float* p = src + (size_t)id * Csrc + col + l8 * 4;
#pragma unroll
    for (int kc = 0; kc < NKC; ++kc)
      v[m][kc] = (m0 + m < n) ? ldg_f4(p + kc * PANEL_K) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <int NKC, bool WEIGHTED>
__device__ __forceinline__ void add4(float4 (&acc)[NKC], const float4 (&v)[4][NKC], const float (&wv)[4]) {
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

// acc[kc] = sum_m w_m * src[list[m], col + kc*32 + l8*4 .. +4]; unweighted lists return the MEAN.
// Must be called by all 32 lanes (n is uniform inside a quarter-warp, may differ between the four
// quarter-warps of the warp: trip counts are made warp-uniform because the id broadcast is a shuffle).
template <int NKC, bool WEIGHTED>
__device__ __forceinline__ void gather_slot(float4 (&acc)[NKC], const GatherSlot& d,
                                            const float* __restrict__ src, int Csrc, int col,
                                            const int* __restrict__ rows, const float* __restrict__ weights,
                                            int l8) {
#pragma unroll
  for (int kc = 0; kc < NKC; ++kc) acc[kc] = make_float4(0.f, 0.f, 0.f, 0.f);
  int nmax = max(d.n, __shfl_xor_sync(C3P_FULL_MASK, d.n, 8));
  nmax = max(nmax, __shfl_xor_sync(C3P_FULL_MASK, nmax, 16));
  if (nmax == 0) return;
  float4 v[4][NKC];
  float wv[4];
  load4<NKC, WEIGHTED>(v, wv, d.ids, d.w, 0, d.n, src, Csrc, col, l8);
  add4<NKC, WEIGHTED>(acc, v, wv);
  if (nmax > 4) {
    load4<NKC, WEIGHTED>(v, wv, d.ids, d.w, 4, d.n, src, Csrc, col, l8);
    add4<NKC, WEIGHTED>(acc, v, wv);
  }
  for (int m0 = 8; m0 < nmax; m0 += 8) {  // long lists: further rounds of eight
    const uint32_t at = d.lb + (uint32_t)min(m0 + l8, max(d.n, 1) - 1);
    const int idr = __ldg(rows + at);
    float wr = 1.f;
    if (WEIGHTED) wr = __ldg(weights + at);
    load4<NKC, WEIGHTED>(v, wv, idr, wr, m0, d.n, src, Csrc, col, l8);
    add4<NKC, WEIGHTED>(acc, v, wv);
    if (m0 + 4 < nmax) {
      load4<NKC, WEIGHTED>(v, wv, idr, wr, m0 + 4, d.n, src, Csrc, col, l8);
      add4<NKC, WEIGHTED>(acc, v, wv);
    }
  }
  if (!WEIGHTED && d.n > 1) {
    const float inv = __fdiv_rn(1.f, (float)d.n);  // one divide per cell, not one per channel
#pragma unroll
    for (int kc = 0; kc < NKC; ++kc) {
      acc[kc].x *= inv; acc[kc].y *= inv; acc[kc].z *= inv; acc[kc].w *= inv;
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
