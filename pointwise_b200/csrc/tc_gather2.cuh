// tc_gather2.cuh -- producer-side helpers of the second-generation tensor-core kernels (gather_mma2.cu,
// backward_filter2.cu): a quarter-warp (8 lanes x 16 B = one 128-byte row segment) serves one work item of a
// k_group_items list -- it walks the item's list of neighbour rows and reduces NKC consecutive 32-channel panels
// of those rows in registers.  The gather selects on the ADDRESS and the WEIGHT, never on the loaded value:
// absent members re-read a valid row with weight 0 and every accumulation is a packed FFMA2.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace c3p {
namespace tc {

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// acc += w * v on both halves of a float4 (two FFMA2)
__device__ __forceinline__ void fma4(float4& acc, float w, const float4& v) {
  const float2 ww = make_float2(w, w);
  const float2 lo = __ffma2_rn(ww, make_float2(v.x, v.y), make_float2(acc.x, acc.y));
  const float2 hi = __ffma2_rn(ww, make_float2(v.z, v.w), make_float2(acc.z, acc.w));
  acc = make_float4(lo.x, lo.y, hi.x, hi.y);
}

struct G2Item {
  uint32_t pos;   // position of the cell's first entry in the list arrays
  int p;          // row of the sub-tile (0..127)
  int n;          // members (0: store zeros)
  float inv;      // 1 / n (unweighted lists; MUFU reciprocal, within 1 ulp)
  int ids;        // lane l8: row id of entry l8 of the list (prefetched)
  float w;        // lane l8: its weight (WEIGHTED)
};

template <bool WEIGHTED>
__device__ __forceinline__ void g2_prefetch(G2Item& it, const int* __restrict__ rows,
                                            const float* __restrict__ weights, int m0, int l8,
                                            unsigned max_row) {
  // select on the ADDRESS: nothing consumes the loaded values until the item is gathered
  const uint32_t at = it.n > 0 ? it.pos + (uint32_t)min(m0 + l8, it.n - 1) : 0u;
  it.ids = (int)min((unsigned)__ldg(rows + at), max_row);
  if (WEIGHTED) it.w = __ldg(weights + at);
}

// acc[kc] = sum_m w_m * src[list[m], col + kc*32 + l8*4 .. +4], M members per round.  Called by all 32 lanes; n
// is uniform inside a quarter-warp and the trip count is made warp-uniform (the id broadcast is a shuffle).
template <int NKC, int M, bool WEIGHTED>
__device__ __forceinline__ void g2_gather(float4 (&acc)[NKC], G2Item& it, int nmax,
                                          const float* __restrict__ src, int Csrc, int col,
                                          const int* __restrict__ rows, const float* __restrict__ weights,
                                          int l8, unsigned max_row) {
#pragma unroll
  for (int kc = 0; kc < NKC; ++kc) acc[kc] = make_float4(0.f, 0.f, 0.f, 0.f);
  // row address = base + id * (row bytes): one IMAD.WIDE.U32 per member
  const char* base = reinterpret_cast<const char*>(src + col + l8 * 4);
  const unsigned row_bytes = (unsigned)Csrc * 4u;
  for (int m0 = 0; m0 < nmax; m0 += M) {
    if (m0 && !(m0 & 7)) g2_prefetch<WEIGHTED>(it, rows, weights, m0, l8, max_row);
    float4 v[M][NKC];
    float wv[M];
    // (Predicated loads into zeroed registers instead of weight-0 re-reads were measured slower: 0.91 against
    // 0.87 ms forward -- the extra zeroing instructions cost more than the saved L1 wavefronts.)
#pragma unroll
    for (int m = 0; m < M; ++m) {
      const int id = __shfl_sync(C3P_FULL_MASK, it.ids, (m0 + m) & 7, 8);
      float w = it.inv;
      if (WEIGHTED) w = __shfl_sync(C3P_FULL_MASK, it.w, (m0 + m) & 7, 8);
      wv[m] = (m0 + m < it.n) ? w : 0.f;
      const float* p = reinterpret_cast<const float*>(base + (size_t)(unsigned)id * row_bytes);
#pragma unroll
      for (int kc = 0; kc < NKC; ++kc) v[m][kc] = ldg4(p + kc * PANEL_K);
    }
#pragma unroll
    for (int m = 0; m < M; ++m)
#pragma unroll
      for (int kc = 0; kc < NKC; ++kc) fma4(acc[kc], wv[m], v[m][kc]);
  }
}

// Same on a 32-bit shared-window address.
__device__ __forceinline__ void g2_store_split(uint32_t dst, uint32_t lo_offset, const float4& v) {
  const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
  const float2 l0 = __fadd2_rn(make_float2(v.x, v.y), make_float2(-h.x, -h.y));
  const float2 l1 = __fadd2_rn(make_float2(v.z, v.w), make_float2(-h.z, -h.w));
  sts128(dst, h);
  sts128(dst + lo_offset, make_float4(l0.x, l0.y, l1.x, l1.y));
}

// TF32 main panel + BF16 correction panel (tc_common.cuh, "BF16 correction products").  A lane owns 4 channels of its
// row: hi = tf32_rn(v) goes to the fp32 panel as one 16-byte chunk, the BF16 pieces of lo and hi are 8 bytes each.
// (A lane-pair exchange that turns the two 8-byte stores into one conflict-free 16-byte store per lane was measured:
// grad_input 1.52 -> 1.62 ms -- the shuffles and selects cost the producers more than the bank conflicts.)
// hi = tf32_rn(v) as one fp32 chunk at `hi_dst`; the BF16 pairs of lo = v - hi and of hi (8 bytes each) at `lo16_dst`
// and `hi16_dst`.
__device__ __forceinline__ void g2_store_split16(uint32_t hi_dst, uint32_t lo16_dst, uint32_t hi16_dst, const float4& v) {
  const float4 h = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
  const float2 l0 = __fadd2_rn(make_float2(v.x, v.y), make_float2(-h.x, -h.y));
  const float2 l1 = __fadd2_rn(make_float2(v.z, v.w), make_float2(-h.z, -h.w));
  sts128(hi_dst, h);
  sts64(lo16_dst, pack_bf16x2(l0.x, l0.y), pack_bf16x2(l1.x, l1.y));
  sts64(hi16_dst, pack_bf16x2(h.x, h.y), pack_bf16x2(h.z, h.w));
}
// Interleaved K order (tc_common.cuh, panel_offset16i): chunk = [lo (4 bf16) | hi (4 bf16)], one 16-byte store.
__device__ __forceinline__ void g2_store_split16i(uint32_t hi_dst, uint32_t c_dst, const float4& v) {
  const float4 h = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
  const float2 l0 = __fadd2_rn(make_float2(v.x, v.y), make_float2(-h.x, -h.y));
  const float2 l1 = __fadd2_rn(make_float2(v.z, v.w), make_float2(-h.z, -h.w));
  sts128(hi_dst, h);
  sts128(c_dst, make_float4(__uint_as_float(pack_bf16x2(l0.x, l0.y)), __uint_as_float(pack_bf16x2(l1.x, l1.y)),
                            __uint_as_float(pack_bf16x2(h.x, h.y)), __uint_as_float(pack_bf16x2(h.z, h.w))));
}
__device__ __forceinline__ void g2_store_split16(unsigned char* hi_dst, unsigned char* lo16_dst, unsigned char* hi16_dst,
                                                 const float4& v) {
  const float4 h = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
  const float2 l0 = __fadd2_rn(make_float2(v.x, v.y), make_float2(-h.x, -h.y));
  const float2 l1 = __fadd2_rn(make_float2(v.z, v.w), make_float2(-h.z, -h.w));
  *reinterpret_cast<float4*>(hi_dst) = h;
  *reinterpret_cast<uint2*>(lo16_dst) = make_uint2(pack_bf16x2(l0.x, l0.y), pack_bf16x2(l1.x, l1.y));
  *reinterpret_cast<uint2*>(hi16_dst) = make_uint2(pack_bf16x2(h.x, h.y), pack_bf16x2(h.z, h.w));
}

// 1 / x for x >= 1 (member counts): one MUFU.RCP, within 1 ulp.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Writes the TF32 hi part of v at `dst` and the lo part `lo_offset` bytes further (16-byte chunk).
__device__ __forceinline__ void g2_store_split(unsigned char* dst, uint32_t lo_offset, const float4& v) {
  const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
  const float2 l0 = __fadd2_rn(make_float2(v.x, v.y), make_float2(-h.x, -h.y));
  const float2 l1 = __fadd2_rn(make_float2(v.z, v.w), make_float2(-h.z, -h.w));
  *reinterpret_cast<float4*>(dst) = h;
  *reinterpret_cast<float4*>(dst + lo_offset) = make_float4(l0.x, l0.y, l1.x, l1.y);
}

}  // namespace tc
}  // namespace c3p
