// tc_selftest.cu -- a one-CTA 3xTF32 GEMM D[128 x N] = A[128 x K] * B[N x K]^T on tcgen05/TMEM.
// It exercises exactly the plumbing the fused Conv3p tensor-core kernels rely on (operand panels
// written by ordinary threads in the 128B-swizzled K-major layout, descriptors, kind::tf32 MMA with
// the hi/lo split, tcgen05.commit -> mbarrier, tcgen05.ld) so that a layout or descriptor mistake is
// caught by a direct comparison with a float64 product (tests/test_gpu_tc.py).
#include "common.cuh"
#include "tc_common.cuh"

namespace c3p {

using namespace tc;

__global__ void __launch_bounds__(128, 1)
k_tc_selftest(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int N, int K,
              int split) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* a_hi = smem;
  unsigned char* a_lo = a_hi + 128 * PANEL_ROW_BYTES;
  unsigned char* b_hi = a_lo + 128 * PANEL_ROW_BYTES;
  unsigned char* b_lo = b_hi + (size_t)N * PANEL_ROW_BYTES;
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const bool bf16c = (split & 8) != 0;       // the production split: TF32 main product + BF16 correction products
  const uint32_t idesc = make_idesc_tf32(128, N);
  split &= 1;

  uint32_t phase = 0;
  for (int kp = 0; kp < K / PANEL_K; ++kp) {
    // fill the panels: 16-byte chunks, thread per chunk
    for (int e = tid; e < 128 * 8; e += 128) {
      const int r = e >> 3, c = e & 7;
      const float4 v = *reinterpret_cast<const float4*>(A + (size_t)r * K + kp * PANEL_K + c * 4);
      float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
      if (bf16c) h = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
      float4 l = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
      if (!split) h = v;
      const uint32_t o = panel_chunk_offset(r, c);
      *reinterpret_cast<float4*>(a_hi + o) = h;
      if (bf16c) {   // A_c = [A_lo | A_hi] as bf16
        *reinterpret_cast<uint2*>(a_lo + panel_offset16(r, 4 * c)) = make_uint2(pack_bf16x2(l.x, l.y), pack_bf16x2(l.z, l.w));
        *reinterpret_cast<uint2*>(a_lo + panel_offset16(r, 32 + 4 * c)) = make_uint2(pack_bf16x2(h.x, h.y), pack_bf16x2(h.z, h.w));
      } else {
        *reinterpret_cast<float4*>(a_lo + o) = l;
      }
    }
    for (int e = tid; e < N * 8; e += 128) {
      const int r = e >> 3, c = e & 7;
      const float4 v = *reinterpret_cast<const float4*>(B + (size_t)r * K + kp * PANEL_K + c * 4);
      float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
      if (bf16c) h = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
      float4 l = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
      if (!split) h = v;
      const uint32_t o = panel_chunk_offset(r, c);
      *reinterpret_cast<float4*>(b_hi + o) = h;
      if (bf16c) {   // B_c = [B_hi | B_lo] as bf16
        *reinterpret_cast<uint2*>(b_lo + panel_offset16(r, 4 * c)) = make_uint2(pack_bf16x2(h.x, h.y), pack_bf16x2(h.z, h.w));
        *reinterpret_cast<uint2*>(b_lo + panel_offset16(r, 32 + 4 * c)) = make_uint2(pack_bf16x2(l.x, l.y), pack_bf16x2(l.z, l.w));
      } else {
        *reinterpret_cast<float4*>(b_lo + o) = l;
      }
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after_sync();
      const uint64_t dah = make_smem_desc(smem_u32(a_hi)), dal = make_smem_desc(smem_u32(a_lo));
      const uint64_t dbh = make_smem_desc(smem_u32(b_hi)), dbl = make_smem_desc(smem_u32(b_lo));
#pragma unroll
      for (int ks = 0; ks < PANEL_K / UMMA_K; ++ks) {
        const uint64_t adv = (uint64_t)((ks * UMMA_K * 4) >> 4);
        mma_tf32(tmem, dah + adv, dbh + adv, idesc, (kp | ks) ? 1u : 0u);
        if (split && !bf16c) {
          mma_tf32(tmem, dal + adv, dbh + adv, idesc, 1u);
          mma_tf32(tmem, dah + adv, dbl + adv, idesc, 1u);
        }
      }
      if (bf16c) {   // 64 bf16 of K in 4 steps of 16 (32 bytes)
        const uint32_t idesc16 = make_idesc_bf16(128, N);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma_bf16(tmem, dal + (uint64_t)(ks * 2), dbl + (uint64_t)(ks * 2), idesc16, 1u);
      }
      mma_commit(&done_bar);
    }
    mbar_wait(&done_bar, phase);
    phase ^= 1;
  }
  tc_fence_after_sync();
  // epilogue: warp w owns TMEM lanes [32w, 32w+32) == rows
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (c0 + j < N) D[(size_t)row * N + c0 + j] = v[j];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// MN-major variant: D[128 x N] = A^T * B with A given as [K x 128] and B as [K x N] (contraction index
// outermost in memory, as the point index is in the weight-gradient kernel).  K <= 64, N % 32 == 0.
__global__ void __launch_bounds__(128, 1)
k_tc_selftest_mn(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int N, int K,
                 int split) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t panel = (uint32_t)K * PANEL_ROW_BYTES;   // one panel: K rows x 32 fp32
  const bool bf16c = (split & 8) != 0;                    // TF32 main product + BF16 correction products
  split &= 1;
  unsigned char* a_hi = smem;                             // 4 panels (M = 128)
  unsigned char* a_lo = a_hi + 4 * panel;                 // (bf16c: 2 panels of 2K rows x 64 bf16: rows [0,K) lo, [K,2K) hi)
  unsigned char* b_hi = a_lo + 4 * panel;                 // N/32 panels
  unsigned char* b_lo = b_hi + (N / 32) * panel;          // (bf16c: ceil(N/64) panels of 2K rows: rows [0,K) hi, [K,2K) lo)
  const uint32_t panel16 = 2u * panel;
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  for (int e = tid; e < K * 32; e += 128) {      // A: K rows x 32 chunks of 16 B
    const int r = e >> 5, c = e & 31;
    const float4 v = *reinterpret_cast<const float4*>(A + (size_t)r * 128 + c * 4);
    float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
    if (bf16c) h = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
    const float4 l = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
    if (!split) h = v;
    const uint32_t o = (uint32_t)(c >> 3) * panel + panel_chunk_offset_mn(r, c & 7);
    *reinterpret_cast<float4*>(a_hi + o) = h;
    if (bf16c) {   // channel 4c: MN block (4c) >> 6, element (4c) & 63
      unsigned char* pc = a_lo + (uint32_t)((4 * c) >> 6) * panel16;
      *reinterpret_cast<uint2*>(pc + panel_offset16(r, (4 * c) & 63)) = make_uint2(pack_bf16x2(l.x, l.y), pack_bf16x2(l.z, l.w));
      *reinterpret_cast<uint2*>(pc + panel_offset16(K + r, (4 * c) & 63)) = make_uint2(pack_bf16x2(h.x, h.y), pack_bf16x2(h.z, h.w));
    } else {
      *reinterpret_cast<float4*>(a_lo + o) = l;
    }
  }
  for (int e = tid; e < K * (N / 4); e += 128) {
    const int r = e / (N / 4), c = e % (N / 4);
    const float4 v = *reinterpret_cast<const float4*>(B + (size_t)r * N + c * 4);
    float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
    if (bf16c) h = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
    const float4 l = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
    if (!split) h = v;
    const uint32_t o = (uint32_t)(c >> 3) * panel + panel_chunk_offset_mn(r, c & 7);
    *reinterpret_cast<float4*>(b_hi + o) = h;
    if (bf16c) {
      unsigned char* pc = b_lo + (uint32_t)((4 * c) >> 6) * panel16;
      *reinterpret_cast<uint2*>(pc + panel_offset16(r, (4 * c) & 63)) = make_uint2(pack_bf16x2(h.x, h.y), pack_bf16x2(h.z, h.w));
      *reinterpret_cast<uint2*>(pc + panel_offset16(K + r, (4 * c) & 63)) = make_uint2(pack_bf16x2(l.x, l.y), pack_bf16x2(l.z, l.w));
    } else {
      *reinterpret_cast<float4*>(b_lo + o) = l;
    }
  }
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after_sync();
    const uint32_t idesc = make_idesc_tf32_mn(128, N);
    for (int j = 0; j < K / 8; ++j) {
      const uint32_t adv = (uint32_t)j * 1024u;
      const uint64_t dah = make_smem_desc_mn(smem_u32(a_hi) + adv, panel);
      const uint64_t dal = make_smem_desc_mn(smem_u32(a_lo) + adv, panel);
      const uint64_t dbh = make_smem_desc_mn(smem_u32(b_hi) + adv, panel);
      const uint64_t dbl = make_smem_desc_mn(smem_u32(b_lo) + adv, panel);
      mma_tf32(tmem, dah, dbh, idesc, j ? 1u : 0u);
      if (split && !bf16c) {
        mma_tf32(tmem, dal, dbh, idesc, 1u);
        mma_tf32(tmem, dah, dbl, idesc, 1u);
      }
    }
    if (bf16c) {   // [lo ; hi] x [hi ; lo]: 2K rows in steps of 16
      const uint32_t idesc16 = make_idesc_bf16_mn(128, N);
      for (int j = 0; j < 2 * K / 16; ++j) {
        const uint32_t adv = (uint32_t)j * 2048u;
        mma_bf16(tmem, make_smem_desc_mn16(smem_u32(a_lo) + adv, panel16), make_smem_desc_mn16(smem_u32(b_lo) + adv, panel16),
                 idesc16, 1u);
      }
    }
    mma_commit(&done_bar);
  }
  mbar_wait(&done_bar, 0);
  tc_fence_after_sync();
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) D[(size_t)row * N + c0 + j] = v[j];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// Issue-rate probe: one CTA issues `reps` rounds of 8 tcgen05.mma (M = 128, N columns, operands at fixed shared-memory
// addresses: contents irrelevant) and measures the cycles until the last one has completed.  mode bit 0: MN-major
// operands (else K-major); bit 1: kind::f16 BF16 (K = 16 per instruction) instead of kind::tf32 (K = 8).
__global__ void __launch_bounds__(128, 1) k_mma_rate(int N, int mode, int reps, unsigned long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 * 1024) / 16; i += 128) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid == 0) {
    mbar_init(&done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const bool mn = mode & 1, bf = mode & 2;
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 32 * 1024;          // A: 4 panels of 64 rows, B: up to 8 panels of 32 rows
    const uint32_t idesc = bf ? (mn ? make_idesc_bf16_mn(128, N) : make_idesc_bf16(128, N))
                              : (mn ? make_idesc_tf32_mn(128, N) : make_idesc_tf32(128, N));
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint64_t da, db;
        if (!mn) {          // K-major: panels of [rows x 128 B], 32 bytes of K per step (4 steps per panel)
          da = make_smem_desc(a0 + (uint32_t)(j >> 2) * 16384u) + (uint64_t)((j & 3) * 2);
          db = make_smem_desc(b0 + (uint32_t)(j >> 2) * (uint32_t)N * 128u) + (uint64_t)((j & 3) * 2);
        } else if (!bf) {   // MN-major fp32: 8 rows (1 KB) per step, 32-wide MN blocks 8 KB (A) / 4 KB (B) apart
          da = make_smem_desc_mn(a0 + (uint32_t)j * 1024u, 8192u);
          db = make_smem_desc_mn(b0 + (uint32_t)(j & 3) * 1024u, 4096u);
        } else {            // MN-major bf16: 16 rows (2 KB) per step, 64-wide MN blocks 16 KB (A) / 8 KB (B) apart
          da = make_smem_desc_mn16(a0 + (uint32_t)j * 2048u, 16384u);
          db = make_smem_desc_mn16(b0 + (uint32_t)(j & 3) * 2048u, 8192u);
        }
        if (bf) mma_bf16(tmem, da, db, idesc, (r | j) ? 1u : 0u);
        else mma_tf32(tmem, da, db, idesc, (r | j) ? 1u : 0u);
      }
    }
    mma_commit(&done_bar);
    mbar_wait(&done_bar, 0);
    out[0] = (unsigned long long)(clock64() - t0);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace c3p

extern "C" int conv3p_debug_mma_rate(int N, int mode, int reps, unsigned long long* cycles_device,
                                     conv3p_stream_t stream) {
  using namespace c3p;
  if (N < 16 || N > 256 || N % 16 || reps < 1 || !cycles_device) return CONV3P_ERR_INVALID_ARGUMENT;
  const size_t smem = 128 * 1024;
  { const int st_ = ensure_dynamic_smem(k_mma_rate, smem); if (st_) return st_; }
  k_mma_rate<<<1, 128, smem, stream>>>(N, mode, reps, cycles_device);
  C3P_LAUNCH_CHECK("k_mma_rate");
  return CONV3P_OK;
}

extern "C" int conv3p_selftest_tc_mn(const float* A, const float* B, float* D, int N, int K, int split,
                                     conv3p_stream_t stream) {
  using namespace c3p;
  if (N < 32 || N > 256 || N % 32 || K < 8 || K > 64 || K % 8) return CONV3P_ERR_INVALID_ARGUMENT;
  const size_t smem = 2 * (size_t)(4 + N / 32 + 1) * K * tc::PANEL_ROW_BYTES;
  { const int st_ = ensure_dynamic_smem(k_tc_selftest_mn, smem); if (st_) return st_; }
  k_tc_selftest_mn<<<1, 128, smem, stream>>>(A, B, D, N, K, split);
  C3P_LAUNCH_CHECK("k_tc_selftest_mn");
  return CONV3P_OK;
}

extern "C" int conv3p_selftest_tc(const float* A, const float* B, float* D, int N, int K, int split,
                                  conv3p_stream_t stream) {
  using namespace c3p;
  if (N < 16 || N > 256 || N % 16 || K < 32 || K % 32) return CONV3P_ERR_INVALID_ARGUMENT;
  const size_t smem = 2 * 128 * tc::PANEL_ROW_BYTES + 2 * (size_t)N * tc::PANEL_ROW_BYTES + 1024;
  { const int st_ = ensure_dynamic_smem(k_tc_selftest, smem); if (st_) return st_; }
  k_tc_selftest<<<1, 128, smem, stream>>>(A, B, D, N, K, split);
  C3P_LAUNCH_CHECK("k_tc_selftest");
  return CONV3P_OK;
}
