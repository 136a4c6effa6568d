// tc_common.cuh -- sm_100a tensor-core plumbing written as inline PTX: tcgen05.mma (kind::tf32),
// TMEM allocation / loads, mbarriers, bulk async copies, and the shared-memory operand layout.
//
// Operand layout ("K-major, 128-byte swizzle"): an operand panel is [rows x 32 fp32] = rows x 128 B,
// row r at byte r*128, and inside a row the 16-byte chunk c (0..7) is stored at chunk position
// c ^ (r & 7).  Panels are 1024-byte aligned (8-row swizzle atoms, stride-byte-offset 1024).  One
// tcgen05.mma of kind::tf32 consumes K = 8 fp32 (32 B) per row; successive K steps inside a panel
// advance the descriptor start address by 32 B.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace c3p {
namespace tc {

constexpr int PANEL_K = 32;           // fp32 per panel row (128 B)
constexpr int PANEL_ROW_BYTES = 128;
constexpr int UMMA_K = 8;             // fp32 per tcgen05.mma.kind::tf32 along K

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Byte offset of fp32 element (row, k) of a panel, k in [0,32).
__host__ __device__ __forceinline__ uint32_t panel_offset(int row, int k) {
  return (uint32_t)row * PANEL_ROW_BYTES + ((((uint32_t)k >> 2) ^ ((uint32_t)row & 7u)) << 4) +
         (((uint32_t)k & 3u) << 2);
}
// Byte offset of 16-byte chunk `chunk` (0..7) of a panel row.
__host__ __device__ __forceinline__ uint32_t panel_chunk_offset(int row, int chunk) {
  return (uint32_t)row * PANEL_ROW_BYTES + ((((uint32_t)chunk) ^ ((uint32_t)row & 7u)) << 4);
}

// Shared-memory matrix descriptor: K-major, SWIZZLE_128B, SBO = 1024 B, version 1 (Blackwell).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                        // descriptor version
  d |= (uint64_t)2 << 61;                        // layout type SWIZZLE_128B
  return d;
}

// MN-major fp32 operands (contraction index = panel row, as the point index is in the weight-gradient
// kernel).  For 32-bit MN-major operands the only swizzled layout is "128B swizzle with a 32-byte base"
// (CUTLASS UMMA::Layout_MN_SW128_32B_Atom): a panel is again [rows x 128 B], row r at r*128, but the
// swizzle permutes the four 32-byte units of a row, unit u stored at u ^ (r & 3); atoms are 4 rows
// (512 B).  MN runs along the row (32 fp32 per panel; further MN blocks live in further panels
// `mn_block_stride` bytes apart = leading byte offset), K runs over rows (stride byte offset 512 B
// between 4-row atoms; one kind::tf32 MMA consumes 8 rows).
__host__ __device__ __forceinline__ uint32_t panel_chunk_offset_mn(int row, int chunk) {
  const uint32_t unit = (((uint32_t)chunk >> 1) ^ ((uint32_t)row & 3u));
  return (uint32_t)row * PANEL_ROW_BYTES + (((unit << 1) | ((uint32_t)chunk & 1u)) << 4);
}
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr, uint32_t mn_block_stride) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((mn_block_stride >> 4) & 0x3FFFu) << 16;  // leading byte offset
  d |= (uint64_t)(512 >> 4) << 32;                          // stride byte offset (4-row atoms)
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                                   // layout type SWIZZLE_128B_BASE32B
  return d;
}
__host__ __device__ __forceinline__ uint32_t make_idesc_tf32_mn(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// Instruction descriptor: D fp32 (+)= A tf32 * B tf32, both K-major, shape M x N (x K=8).
__host__ __device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- BF16 correction products ------------------------------------------------------------------------------
// The two correction terms of the split product, A_lo*B_hi + A_hi*B_lo, are ~2^-11 of the main term: they only need
// ~8 significant bits, so they run as ONE kind::f16 MMA chain on BF16 operands with the two terms concatenated
// along K -- A_c = [A_lo | A_hi], B_c = [B_hi | B_lo] -- at twice the K per instruction of kind::tf32 (16 against 8):
// a third fewer tensor-core cycles and a third fewer shared-memory operand reads than three TF32 products, at
// |error| <= ~2^-18 per product (hi = fp32 ROUNDED to an 11-bit significand, so |lo| <= 2^-11 |x|; BF16 rounding of
// either factor of a correction term is 2^-9 relative).
// A BF16 panel row is 128 B = 64 elements, 128B-swizzled like the fp32 panels (16-byte chunk c at c ^ (row & 7)).
__host__ __device__ __forceinline__ uint32_t panel_offset16(int row, int j) {   // byte offset of bf16 element j (0..63)
  return (uint32_t)row * PANEL_ROW_BYTES + ((((uint32_t)j >> 3) ^ ((uint32_t)row & 7u)) << 4) + (((uint32_t)j & 7u) << 1);
}
// The FORWARD gather kernel's K-major correction panels use an INTERLEAVED K order (any order works as long as both
// operands agree): the 16-byte chunk of channels 4g..4g+3 holds [part 0 (4 bf16) | part 1 (4 bf16)], part 0 = lo on the
// A side / hi on the B side, part 1 the other one.  A producer lane owns exactly those 4 channels of its row, so it
// writes both parts with ONE conflict-free 16-byte store instead of two 8-byte ones.  Measured (ncu A/B, one box):
// forward 0.703 -> 0.689 ms; the weighted (grad_input) variant got SLOWER with it, 1.51 -> 1.59 ms (its producers wait
// for gather latency, not for the shared-memory pipe, and the store pattern changed their overlap), so it keeps the
// [lo (32 ch) | hi (32 ch)] order of panel_offset16.
__host__ __device__ __forceinline__ uint32_t panel_offset16i(int row, int ch, int part) {
  return (uint32_t)row * PANEL_ROW_BYTES + ((((uint32_t)ch >> 2) ^ ((uint32_t)row & 7u)) << 4) + ((uint32_t)part << 3) +
         (((uint32_t)ch & 3u) << 1);
}
__host__ __device__ __forceinline__ uint32_t make_idesc_bf16(int M, int N) {      // D fp32 += A bf16 * B bf16, K-major
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ __forceinline__ uint32_t make_idesc_bf16_mn(int M, int N) {   // both operands MN-major
  return make_idesc_bf16(M, N) | (1u << 15) | (1u << 16);
}
// MN-major 16-bit operands: standard 128B swizzle, 8-row atoms of 1024 B along K (stride byte offset), 64-element
// MN blocks `mn_block_stride` bytes apart (leading byte offset); one kind::f16 MMA consumes 16 rows.
__device__ __forceinline__ uint64_t make_smem_desc_mn16(uint32_t smem_addr, uint32_t mn_block_stride) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((mn_block_stride >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                                   // layout type SWIZZLE_128B
  return d;
}
// fp32 rounded to nearest into a 10-bit mantissa (the value kind::tf32 sees exactly); x - tf32_rn(x) is exact in fp32.
__device__ __forceinline__ float tf32_rn(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
// {bf16(lo_elem) in bits [0,16), bf16(hi_elem) in bits [16,32)}: lo_elem is the one at the lower address
__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
  return d;
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}

// ---- mbarrier ---------------------------------------------------------------------------------------
// try_wait with a suspend-time hint: the thread may sleep until the phase completes (measured: same speed without)
#define C3P_TRY_WAIT "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x4E20;\n\t"
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      C3P_TRY_WAIT
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}

// The same operations on 32-bit shared-window addresses.  smem_u32() of a generic pointer costs an S2R (cluster
// CTA id) + LEA every time it is evaluated, and the compiler does not hoist it across the volatile asm of a
// producer loop; hot loops therefore convert their base addresses ONCE and do address arithmetic in registers.
// smem_u32() whose result the compiler must keep in a register (an opaque move: it cannot be rematerialised).
__device__ __forceinline__ uint32_t smem_u32_once(const void* p) {
  uint32_t x = smem_u32(p), y;
  asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      C3P_TRY_WAIT
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr)
               : "memory");
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ int lds32(uint32_t addr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
      "l"(src_gmem), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// Generic-proxy writes (st.shared) -> visible to the async proxy (tcgen05.mma operand reads).
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- bulk async copy global -> shared, completion on an mbarrier (SASS: UBLKCP) ---------------------
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                              uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- TMEM ---------------------------------------------------------------------------------------------
// Called by one full warp.  Writes the TMEM base address to *slot (shared memory).
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.  accumulate == 0 overwrites D.
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The same with the descriptors as (low word, high word) pairs.  One thread issues every MMA of a CTA; measured with
// tools/mma_rate.py, an M=128 x N=64 instruction occupies the tensor core for ~43 cycles, so the issuing thread has
// only ~10 dependent instructions per MMA before IT becomes the bottleneck (64-bit descriptor arithmetic per MMA cost
// ~90 cycles each in the weight-gradient kernel).  The address field is the low 14 bits of the low word (bytes >> 4);
// stepping through a panel or a ring is a 32-bit add on that word that can never carry out of the field (shared
// memory addresses are < 256 KB), the high word is constant per operand.
struct Desc32 {
  uint32_t lo, hi;
};
__device__ __forceinline__ Desc32 split_desc(uint64_t d) { return Desc32{(uint32_t)d, (uint32_t)(d >> 32)}; }
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All previously issued tcgen05.mma of this thread arrive on `bar` when they complete (implies
// tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 columns of fp32: thread t of the warp receives TMEM lane (lane_base + t), columns
// [col, col+32).  The warp may only touch lanes [32*(warp%4), 32*(warp%4)+32).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- 3xTF32 operand split -------------------------------------------------------------------------------
// hi = the value the tensor core sees (fp32 truncated to a 10-bit mantissa), lo = x - hi (exact in fp32).
// D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo keeps ~21 mantissa bits of every product.
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ float tf32_lo(float x) { return x - tf32_hi(x); }

}  // namespace tc
}  // namespace c3p
