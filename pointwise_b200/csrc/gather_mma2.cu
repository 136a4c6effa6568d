// gather_mma2.cu -- fused gather + 3xTF32 tcgen05 kernel (forward and grad_input).
//
//   out[p, n] = sum_f sum_k A_f[p, k] * Wpanel_f[n, k],   A_f = per-cell mean (forward, tf_conv3p_atrous.cpp:
//   480-494) or weighted sum over the backward lists (grad_input, :682-692).
// For T sub-tiles of 128 voxel-sorted points the (27*Csrc) x Nout contraction runs as T accumulators of 128 x Nout
// fp32 in TMEM; K is walked as (kernel cell f, K batch, sub-tile t).  The aggregated operand A_f is never written
// to HBM: producer warps gather neighbour rows with 16-byte loads (quarter-warp per point = one 128-byte row
// segment per instruction), reduce them in registers, split the result into TF32 hi/lo parts and store them
// straight into the 128B-swizzled K-major operand panels of a shared-memory ring; one thread issues tcgen05.mma
// (kind::tf32, M=128, N=Nout, K=8): D += A_hi*W_hi + A_lo*W_hi + A_hi*W_lo (3xTF32, fp32 accumulation in TMEM).
// The CUDA-core side of the fusion bounds the kernel (profiles/r1_summary.md), so it is organised around it:
//
//  * A PRE-PASS (k_group_items, one warp per 128-point sub-tile) turns the count table into, for every (sub-tile,
//    cell) group, a list of 128 work items (list position, row, members), 1 KB per group, in scratch memory; the
//    main kernel streams the lists of its groups into a small shared-memory ring with bulk async copies.
//    Items are compacted by population class (> 8 members, 5..8, 1..4, empty), so the producer quarter-warps
//    of one warp see similar list lengths, the second repetition of a group is usually all-empty (58 % of the
//    (row, cell) slots are empty on the S3DIS-like bench cloud) and is then a plain zero fill, and the
//    producers no longer keep per-point prefix tables (28 KB) in shared memory.  (An in-kernel scheduler warp
//    doing the same ranking was measured first: it needed ~4500 cycles per group and paced the whole CTA.)
//  * The gather itself selects on the ADDRESS and the WEIGHT, never on the loaded value: absent members
//    re-read a valid row with weight 0 and every accumulation is a packed FFMA2, which removes the
//    zero-select instructions (one third of the old inner loop).
//  * The freed shared memory deepens the operand ring from 3 to 4-5 stages (two full groups in flight), and
//    weight panels are streamed in hi / lo units so one extra unit is enough to prefetch the next cell.
//
// Warp roles: warps [0, NPW) producers (also the epilogue), NPW = MMA issuer, NPW+1 = weight loader + TMEM
// allocator, NPW+2 = item-list loader.
#include <cstdio>
#include <cstdlib>
#include <utility>

#include "common.cuh"
#include "tc_common.cuh"
#include "tc_gather2.cuh"

namespace c3p {

using namespace tc;

constexpr int G2_NPW = 16;
constexpr int G2_WARPS = G2_NPW + 3;
constexpr int G2_THREADS = G2_WARPS * 32;
constexpr int G2_NIS = 4;                              // item-list slots (groups the scheduler may run ahead)
constexpr int G2_A_STAGE = 2 * 128 * PANEL_ROW_BYTES;  // hi + lo panels of 128 rows x 32 channels
constexpr int G2_MAX_NAS = 6, G2_MAX_NWU = 8;
constexpr int G2_END = -1;

struct G2Args {
  const float* src;         // gathered rows [B*N, Csrc]
  const unsigned char* wp;  // weight panel images [27][Csrc/32][hi,lo][Nout][128 B] (k_prep_weight_panels)
  float* out;               // [B*N, Nout]
  const int* rows;          // list entries: row ids
  const float* weights;     // per-entry weights (WEIGHTED) or nullptr (per-cell mean)
  const uint2* g_items;     // [subtiles][27][128] work items (k_group_items)
  const int* g_nnz;         // [subtiles][27] non-empty rows of the group
  const int* g_rowid;       // [subtiles*128] output row of every sorted position (-1 pad, -2-row poisoned)
  unsigned* counter;        // chunk counter of the dynamic tile scheduler (zeroed by k_group_items)
  long long total_points, subtiles;
  int N, Csrc, Nout, nkb, T, NAS, NWU;
  int debug;  // bit 32: accumulate the phase timers below
  int res_big, res_one;  // scheduler: sub-tiles kept back for chunks smaller than T / for single sub-tiles, in units of G/2
  long long src_stride, out_stride;  // floats between gathered rows / output rows (multiples of 4)
  int activation;                    // epilogue: CONV3P_ACT_*
  // WEIGHTED only, optional: the aggregated rows G_f[j, :] of every non-empty (point, cell) are also written to
  // g_store[(sorted position * 27 + f) * Csrc ...], where the weight-gradient kernel picks them up instead of
  // gathering the same lists a second time (backward_filter2.cu).
  float* g_store;
};

// cycles summed over CTAs (warp 0, lane 0): prologue | producer loop | wait for the last MMA | epilogue |
// in the loop: item fetch | first gather | wait for a free ring stage | stores + second repetition + arrive
__device__ unsigned long long g2_phase_cycles[8];
// TIMED build: sum and max over CTAs of a CTA's total cycles (how uneven the dynamic schedule ends)
__device__ unsigned long long g2_cta_cycles[2];

// ---- pre-pass: compacted work-item lists per (sub-tile, cell) group ---------------------------------------------
// One CTA per sub-tile of ROWS voxel-sorted points, one thread per row.  Within a group the rows are ranked by
// population class (0: more than 8 members, 1: 5..8, 2: 1..4, 3: empty): warp ballots give the rank inside the
// warp, per-warp class counts (one byte each) go through shared memory, so item index = class base + rank.
// The 27 lists of the sub-tile are staged in shared memory and written out with coalesced 16-byte stores.
// HALF (ROWS == 128 only): the same pass also emits the lists of the two 64-row tiles of the sub-tile (the weight-gradient
// kernel's tiling) -- the count table is read once for both gradient kernels instead of once per kernel.
struct HalfLists {
  uint2* items;      // [2 * subtiles][27][64]
  int* nnz;          // [2 * subtiles][27]
  int* rowid;        // [2 * subtiles * 64]
  unsigned* mask;    // [2 * subtiles]
  long long tiles;   // number of 64-row tiles (the second half of the last sub-tile may not exist)
};

template <int ROWS, bool HALF>
__global__ void __launch_bounds__(ROWS, 512 / ROWS)   // at least 512 threads per SM: caps the registers at 128
k_group_items(const int* __restrict__ cnt, const long long* __restrict__ begin, const int* __restrict__ len,
              const float4* __restrict__ sorted_xyzi, long long total_points, long long capacity, int N,
              uint2* __restrict__ g_items, int* __restrict__ g_nnz, int* __restrict__ g_rowid,
              unsigned* __restrict__ g_mask, unsigned* __restrict__ g_counter, HalfLists half) {
  constexpr int NW = ROWS / 32;
  __shared__ uint2 stage[C3P_NCELL * ROWS];
  __shared__ uint32_t wcls[C3P_NCELL][NW];  // per warp: rows of class 0..3, one byte each
  const long long sub = blockIdx.x;
  const int r = threadIdx.x, lane = r & 31, warp = r >> 5;
  const unsigned lt = lanemask_lt();
  const long long s = sub * ROWS + r;
  int row = -1;
  uint32_t pos = 0;
  bool ok = false;
  const int* crow = cnt;
  if (s < total_points) {
    const int b = (int)(s / N);
    row = b * N + __float_as_int(sorted_xyzi[s].w);
    const long long bg = begin[row];
    ok = bg + len[row] <= capacity;
    pos = (uint32_t)bg;
    crow = cnt + (size_t)row * C3P_NCELL;
    if (!ok) row = -2 - row;  // incomplete list: the consumer poisons this point's output
  }
  g_rowid[s] = row;
  int c[C3P_NCELL];
#pragma unroll
  for (int f = 0; f < C3P_NCELL; ++f) c[f] = ok ? __ldg(crow + f) : 0;
  uint32_t rank[C3P_NCELL];  // class | rank inside the warp << 8
#pragma unroll
  for (int f = 0; f < C3P_NCELL; ++f) {
    const int n = c[f];
    const int cls = n > 8 ? 0 : (n > 4 ? 1 : (n > 0 ? 2 : 3));
    uint32_t packed = 0, mine = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned m = __ballot_sync(C3P_FULL_MASK, cls == k);
      packed |= (uint32_t)__popc(m) << (8 * k);   // at most 32 per warp and class
      if (cls == k) mine = (uint32_t)__popc(m & lt);
    }
    rank[f] = (uint32_t)cls | (mine << 8);
    if (lane == 0) wcls[f][warp] = packed;
  }
  __syncthreads();
  unsigned mask = 0;
#pragma unroll
  for (int f = 0; f < C3P_NCELL; ++f) {
    const int cls = (int)(rank[f] & 255u);
    int idx = (int)(rank[f] >> 8), nonempty = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const uint32_t pk = wcls[f][w];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int v = (int)((pk >> (8 * k)) & 255u);
        if (k < cls || (k == cls && w < warp)) idx += v;
        if (k < 3) nonempty += v;
      }
    }
    stage[f * ROWS + idx] = make_uint2(pos, (uint32_t)r | ((uint32_t)c[f] << 8));
    pos += (uint32_t)c[f];
    if (nonempty) mask |= 1u << f;
    if (r == f) g_nnz[sub * C3P_NCELL + f] = nonempty;
  }
  if (r == 0) g_mask[sub] = mask;
  if (r == 0 && sub == 0) *g_counter = 0u;
  __syncthreads();
  const uint4* src = reinterpret_cast<const uint4*>(stage);
  uint4* dst = reinterpret_cast<uint4*>(g_items + sub * C3P_NCELL * ROWS);
  for (int e = r; e < C3P_NCELL * ROWS / 2; e += ROWS) dst[e] = src[e];
  if (HALF) {
    // the two 64-row tiles: rank among the tile's two warps; stage[f][tile][idx] has the layout of two consecutive tiles
    static_assert(!HALF || ROWS == 128, "half lists are the two 64-row tiles of a 128-row sub-tile");
    __syncthreads();
    const int tile = warp >> 1, r64 = r & 63;
    const bool tile_ok = sub * 2 + tile < half.tiles;
    if (tile_ok) half.rowid[s] = row;
    uint32_t p2 = 0;
    {   // list position of the row's first entry (recomputed: pos was advanced past the last cell above)
      int tot = 0;
#pragma unroll
      for (int f = 0; f < C3P_NCELL; ++f) tot += c[f];
      p2 = pos - (uint32_t)tot;
    }
    unsigned hmask = 0;
#pragma unroll
    for (int f = 0; f < C3P_NCELL; ++f) {
      const int cls = (int)(rank[f] & 255u);
      int idx = (int)(rank[f] >> 8), nonempty = 0;
#pragma unroll
      for (int w2 = 0; w2 < 2; ++w2) {
        const int w = tile * 2 + w2;
        const uint32_t pk = wcls[f][w];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int v = (int)((pk >> (8 * k)) & 255u);
          if (k < cls || (k == cls && w < warp)) idx += v;
          if (k < 3) nonempty += v;
        }
      }
      stage[(tile * C3P_NCELL + f) * 64 + idx] = make_uint2(p2, (uint32_t)r64 | ((uint32_t)c[f] << 8));
      p2 += (uint32_t)c[f];
      if (nonempty) hmask |= 1u << f;
      if (r64 == f && tile_ok) half.nnz[(sub * 2 + tile) * C3P_NCELL + f] = nonempty;
    }
    if (r64 == 0 && tile_ok) half.mask[sub * 2 + tile] = hmask;
    __syncthreads();
    uint4* dst2 = reinterpret_cast<uint4*>(half.items + sub * 2 * C3P_NCELL * 64);
    const int n16 = (sub * 2 + 1 < half.tiles ? 2 : 1) * C3P_NCELL * 64 / 2;   // 16-byte units of the existing tiles
    for (int e = r; e < n16; e += ROWS) dst2[e] = src[e];
  }
}

// Barrier slots of k_gather_mma2 in one shared array, addressed as bars + 8 * index.
enum {
  G2B_A_FULL = 0, G2B_A_EMPTY = G2B_A_FULL + G2_MAX_NAS, G2B_W_FULL = G2B_A_EMPTY + G2_MAX_NAS,
  G2B_W_EMPTY = G2B_W_FULL + G2_MAX_NWU, G2B_IT_FULL = G2B_W_EMPTY + G2_MAX_NWU, G2B_IT_EMPTY = G2B_IT_FULL + G2_NIS,
  G2B_ACC_FULL = G2B_IT_EMPTY + G2_NIS, G2B_COUNT
};

// Epilogue of one chunk for one producer warp (kept out of line: inlined, its 32 accumulator registers changed
// the register allocation of the producer loop and slowed the weighted variant by 5 %).
__device__ __noinline__ void g2_epilogue(float* out, long long out_stride, int activation, int Nout, uint32_t tmem,
                                         uint32_t tile, const unsigned* active, const int* rowid, int T, int warp,
                                         int lane) {
  const float qnan = __int_as_float(0x7fc00000);
  for (int task = warp; task < 4 * T; task += G2_NPW) {
    const int t = task >> 2, sub = task & 3;
    bool any = false;
    for (int f = 0; f < C3P_NCELL; ++f) any |= ((active[f] >> t) & 1u) != 0;
    const int pt0 = t * 128 + sub * 32;
    const bool poison = rowid[pt0 + lane] < -1;
    for (int c0_ = 0; c0_ < Nout; c0_ += 32) {
      float v[32];
      if (any) {
        tmem_ld_32x32(tmem + ((uint32_t)(sub * 32) << 16) + (uint32_t)(t * Nout + c0_), v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if (activation) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = apply_activation(v[j], activation);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 w4 = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        if (poison) w4 = make_float4(qnan, qnan, qnan, qnan);
        sts128(tile + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) << 4), w4);
      }
      __syncwarp();
      const int c = lane & 7;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + (lane >> 3);
        int row = rowid[pt0 + r];
        if (row < -1) row = -2 - row;
        const float4 w4 = lds128(tile + (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4));
        if (row >= 0 && c0_ + c * 4 < Nout)
          *reinterpret_cast<float4*>(out + (size_t)row * out_stride + c0_ + c * 4) = w4;
      }
      __syncwarp();
    }
  }
}

// BF16C: the two correction products of the hi/lo split run as one BF16 MMA chain (tc_common.cuh); false = three TF32
// products (3xTF32, engine flag 512: A/B timing and the accuracy reference of the tests).
template <int NKC, bool WEIGHTED, bool TIMED, bool BF16C>
__global__ void __launch_bounds__(G2_THREADS, 1) k_gather_mma2(const G2Args a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int Nout = a.Nout, NAS = a.NAS, NWU = a.NWU;
  const uint32_t unit_bytes = (uint32_t)Nout * PANEL_ROW_BYTES;       // hi (or lo) half of a weight panel
  __shared__ uint64_t bars[G2B_COUNT];
  __shared__ uint32_t tmem_slot;
  __shared__ unsigned active[C3P_NCELL];  // bit t: sub-tile t of the chunk has members in cell f
  __shared__ int hdr[G2_NIS];             // group in the slot: K batch | sub-tile << 8 | cell << 16, or G2_END
  // Shared-window addresses, converted once (see tc_common.cuh): operand ring | weight units | item lists | rows
  const uint32_t s_a = smem_u32_once(smem);                              // NAS stages
  const uint32_t s_w = s_a + (uint32_t)NAS * G2_A_STAGE;              // NWU units
  const uint32_t s_items = s_w + (uint32_t)NWU * unit_bytes;          // [NIS][128] uint2
  const uint32_t s_bars = smem_u32_once(bars), s_hdr = smem_u32_once(hdr);
  int* rowid = reinterpret_cast<int*>(smem + (size_t)NAS * G2_A_STAGE + (size_t)NWU * unit_bytes +
                                      G2_NIS * 128 * sizeof(uint2));  // [Tmax * 128]
  auto bar = [&](int which, int i) -> uint32_t { return s_bars + 8u * (uint32_t)(which + i); };

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool timed = TIMED && tid == 0;
  long long tk = timed ? clock64() : 0;
  const long long t_begin = tk;
  unsigned ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define G2_PHASE(i) do { if (TIMED && timed) { const long long t_ = clock64(); ph[i] += (unsigned)(t_ - tk); tk = t_; } } while (0)

  if (warp == G2_NPW && lane == 0) {
    for (int i = 0; i < NAS; ++i) {
      mbar_init(&bars[G2B_A_FULL + i], G2_NPW);
      mbar_init(&bars[G2B_A_EMPTY + i], 1);
    }
    for (int i = 0; i < NWU; ++i) {
      mbar_init(&bars[G2B_W_FULL + i], 1);
      mbar_init(&bars[G2B_W_EMPTY + i], 1);
    }
    for (int i = 0; i < G2_NIS; ++i) {
      mbar_init(&bars[G2B_IT_FULL + i], 1);
      mbar_init(&bars[G2B_IT_EMPTY + i], G2_NPW);
    }
    mbar_init(&bars[G2B_ACC_FULL], 1);
    mbar_fence_init();
  }
  if (warp == G2_NPW + 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const unsigned max_row = (unsigned)(a.total_points - 1);

  // Persistent CTA with a dynamic scheduler: chunks of sub-tiles (one TMEM accumulator per sub-tile) are claimed
  // from a global counter.  Chunk k maps to a fixed range: chunks of Tmax sub-tiles first, then -- for the last
  // few sub-tiles per CTA -- chunks of 2 and of 1, so that the tail of the launch is short although the work per
  // sub-tile varies a lot (dense floors and walls against sparse clutter).
  __shared__ long long s_sub0;
  __shared__ int s_T;
  const long long G = gridDim.x;
  const int Tmax = a.T;
  const long long keep_big = a.res_big * G / 2, keep_one = a.res_one * G / 2;
  const long long n_big = a.subtiles > keep_big ? (a.subtiles - keep_big) / Tmax : 0;   // chunks of Tmax
  const long long r1 = a.subtiles - n_big * Tmax;
  const long long n_mid = (Tmax >= 2 && r1 > keep_one) ? (r1 - keep_one) / 2 : 0;                      // chunks of 2
  const long long n_one = r1 - n_mid * 2;                                               // chunks of 1
  auto claim = [&]() {
    const long long k = (long long)atomicAdd(a.counter, 1u);
    long long sub = 0;
    int T = 0;
    if (k < n_big) { sub = k * Tmax; T = Tmax; }
    else if (k < n_big + n_mid) { sub = n_big * Tmax + (k - n_big) * 2; T = 2; }
    else if (k < n_big + n_mid + n_one) { sub = n_big * Tmax + n_mid * 2 + (k - n_big - n_mid); T = 1; }
    s_sub0 = sub;
    s_T = T;
  };
  if (tid == 0) claim();
  __syncthreads();

  // role state that lives across chunks
  const int q = warp * 4 + (lane >> 3);
  const int l8 = lane & 7;
  // (one register set shared by the roles: a thread only ever plays one of them)
  int st_a = 0, st_b = 0;
  uint32_t st_c = 0, st_d = 0;
  int& p_g = st_a; int& p_aslot = st_b; uint32_t& p_awrap = st_c;                          // producers
  int& m_aslot = st_a; int& m_wslot = st_b; uint32_t& m_aphase = st_c; uint32_t& m_wphase = st_d;  // MMA issuer
  int& l_slot = st_a; uint32_t& l_wrap = st_c;                                             // weight loader
  int& i_g = st_a;                                                                         // item-list loader
  int chunk = 0;

  for (;; ++chunk) {
    const long long sub0 = s_sub0;
    const int T = s_T;
    if (T == 0) break;
    const int PT = T * 128;
    // ---- output rows of the chunk and which (cell, sub-tile) groups have members at all ----------------------------
    for (int p = tid; p < PT; p += G2_THREADS) rowid[p] = __ldg(a.g_rowid + sub0 * 128 + p);
    if (tid < C3P_NCELL) {
      unsigned m = 0;
      for (int t = 0; t < T; ++t)
        if (__ldg(a.g_nnz + (sub0 + t) * C3P_NCELL + tid) > 0) m |= 1u << t;
      active[tid] = m;
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    G2_PHASE(0);
    if (tid == 0) claim();  // next chunk; everybody has read this chunk's range, the answer is read after the end barrier

    if (warp < G2_NPW) {
      // =========================== producers: gather -> mean -> hi/lo -> operand panels ===================
      // Quarter-warp q serves items e = (q + 4g) mod 64 and 127 - e of group g; lane l8 owns one 16-byte chunk of
      // the row segment.  Software-pipelined: the item and the list ids of the next group are fetched before
      // the rows of the current group are gathered.
      auto read_items = [&](int g_, G2Item& i0, G2Item& i1) -> int {
        const int slot = g_ & (G2_NIS - 1);
        mbar_wait(bar(G2B_IT_FULL, slot), (uint32_t)((g_ / G2_NIS) & 1));
        const int h = lds32(s_hdr + 4u * (uint32_t)slot);
        const int e = (q + 4 * g_) & 63;
        // the list is sorted by population: pairing item e with item 127 - e gives every warp one heavy and one
        // light (or empty) block per group, so the warps of a group finish closer together
        const uint2 u0 = lds64(s_items + (uint32_t)(slot * 128 + e) * 8u);
        const uint2 u1 = lds64(s_items + (uint32_t)(slot * 128 + 127 - e) * 8u);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(G2B_IT_EMPTY, slot));
        const bool live = h != G2_END;
        i0.pos = u0.x; i0.p = (int)(u0.y & 255u); i0.n = live ? (int)(u0.y >> 8) : 0;
        i1.pos = u1.x; i1.p = (int)(u1.y & 255u); i1.n = live ? (int)(u1.y >> 8) : 0;
        i0.inv = WEIGHTED ? 0.f : rcp_approx((float)i0.n);
        i1.inv = WEIGHTED ? 0.f : rcp_approx((float)i1.n);
        i0.w = 0.f; i1.w = 0.f;
        g2_prefetch<WEIGHTED>(i0, a.rows, a.weights, 0, l8, max_row);
        g2_prefetch<WEIGHTED>(i1, a.rows, a.weights, 0, l8, max_row);
        return h;
      };
      auto warp_max = [&](int n) -> int {
        n = max(n, __shfl_xor_sync(C3P_FULL_MASK, n, 8));
        return max(n, __shfl_xor_sync(C3P_FULL_MASK, n, 16));
      };
      G2Item c0, c1;
      int hc = read_items(p_g, c0, c1);
      long long tl = timed ? clock64() : 0;
      if (timed) tk = tl;
      while (hc != G2_END) {
        G2Item n0, n1;
        const int hn = read_items(p_g + 1, n0, n1);
        G2_PHASE(4);
        const int col = (hc & 255) * NKC * PANEL_K;
        float4 acc[NKC];
        g2_gather<NKC, 4, WEIGHTED>(acc, c0, warp_max(c0.n), a.src, (int)a.src_stride, col, a.rows, a.weights, l8, max_row);
        // row of G_f in the store: sorted position * 27 + cell
        float* gs = nullptr;
        if (WEIGHTED && a.g_store)
          gs = a.g_store + (((size_t)(sub0 + ((hc >> 8) & 255)) * 128) * C3P_NCELL + (size_t)(hc >> 16)) * a.Csrc +
               col + l8 * 4;
        if (WEIGHTED && gs && c0.n > 0) {
#pragma unroll
          for (int kc = 0; kc < NKC; ++kc)
            *reinterpret_cast<float4*>(gs + (size_t)c0.p * C3P_NCELL * a.Csrc + kc * PANEL_K) = acc[kc];
        }
        G2_PHASE(5);
        // the ring stages of this group must have been drained by the tensor core
        uint32_t stage[NKC];
        {
          int sl = p_aslot;
          uint32_t wr = p_awrap;
#pragma unroll
          for (int kc = 0; kc < NKC; ++kc) {
            if (wr >= 1) mbar_wait(bar(G2B_A_EMPTY, sl), (wr - 1) & 1u);
            stage[kc] = s_a + (uint32_t)sl * G2_A_STAGE;
            if (++sl == NAS) { sl = 0; ++wr; }
          }
        }
        G2_PHASE(6);
        // operand stage = [hi panel (fp32, TF32 values) | second panel]; second panel = lo (fp32) for 3xTF32, or the
        // BF16 correction panel [lo (32 ch) | hi (32 ch)] per row
        auto store_row = [&](int p, const float4 (&v)[NKC]) {
          const uint32_t o = panel_chunk_offset(p, l8);
          if (BF16C && !WEIGHTED) {
            // forward: interleaved correction panel, [lo | hi] of the lane's 4 channels at the hi panel's chunk position
#pragma unroll
            for (int kc = 0; kc < NKC; ++kc)
              g2_store_split16i(stage[kc] + o, stage[kc] + o + 128 * PANEL_ROW_BYTES, v[kc]);
          } else if (BF16C) {
            // correction panel row = [lo (32 ch) | hi (32 ch)]
            const uint32_t oc = 128 * PANEL_ROW_BYTES + (uint32_t)p * PANEL_ROW_BYTES + ((uint32_t)(l8 & 1) << 3);
            const uint32_t c_lo = oc + ((((uint32_t)l8 >> 1) ^ ((uint32_t)p & 7u)) << 4);
            const uint32_t c_hi = oc + (((4u + ((uint32_t)l8 >> 1)) ^ ((uint32_t)p & 7u)) << 4);
#pragma unroll
            for (int kc = 0; kc < NKC; ++kc) g2_store_split16(stage[kc] + o, stage[kc] + c_lo, stage[kc] + c_hi, v[kc]);
          } else {
#pragma unroll
            for (int kc = 0; kc < NKC; ++kc) g2_store_split(stage[kc] + o, 128 * PANEL_ROW_BYTES, v[kc]);
          }
        };
        store_row(c0.p, acc);
        const int nmax1 = warp_max(c1.n);
        if (nmax1 > 0) {
          g2_gather<NKC, 4, WEIGHTED>(acc, c1, nmax1, a.src, (int)a.src_stride, col, a.rows, a.weights, l8, max_row);
          if (WEIGHTED && gs && c1.n > 0) {
#pragma unroll
            for (int kc = 0; kc < NKC; ++kc)
              *reinterpret_cast<float4*>(gs + (size_t)c1.p * C3P_NCELL * a.Csrc + kc * PANEL_K) = acc[kc];
          }
        }
        if (nmax1 > 0) {
          store_row(c1.p, acc);
        } else {   // all-empty second repetition: zero the row's chunk in both panels (any layout: 2 x 16 bytes).
                   // (Remembering per ring stage which rows already hold zeros and skipping their stores -- 58 % of the
                   // (point, cell) slots are empty -- was measured slower: grad_input 1.52 -> 1.72 ms; the flag traffic
                   // and branches cost the producers more than the stores, and the epilogue reuses the ring anyway.)
          const uint32_t o = panel_chunk_offset(c1.p, l8);
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int kc = 0; kc < NKC; ++kc) {
            sts128(stage[kc] + o, z);
            sts128(stage[kc] + o + 128 * PANEL_ROW_BYTES, z);
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          int sl = p_aslot;
#pragma unroll
          for (int kc = 0; kc < NKC; ++kc) {
            mbar_arrive(bar(G2B_A_FULL, sl));
            if (++sl == NAS) sl = 0;
          }
        }
#pragma unroll
        for (int kc = 0; kc < NKC; ++kc)
          if (++p_aslot == NAS) { p_aslot = 0; ++p_awrap; }
        ++p_g;
        c0 = n0; c1 = n1; hc = hn;
        G2_PHASE(7);
      }
      ++p_g;  // the END slot
      if (timed) { tk = clock64(); ph[1] += (unsigned)(tk - tl); }
      // =========================== epilogue: TMEM -> registers -> shared -> global ======================
      // tcgen05.ld hands every lane one accumulator ROW; stored from there, a warp store touches 32 different
      // 128-byte lines with 16 bytes each (32 wavefronts per instruction -- measured 6 % of the kernel).  The
      // 32 x 32 block is therefore transposed through a 4 KB tile of the (now idle) operand ring, so that eight
      // lanes write one full 128-byte row segment and a store instruction covers four lines.
      mbar_wait(bar(G2B_ACC_FULL, 0), (uint32_t)(chunk & 1));
      tc_fence_after_sync();
      G2_PHASE(2);
      g2_epilogue(a.out, a.out_stride, a.activation, Nout, tmem, s_a + (uint32_t)warp * 4096u, active, rowid, T, warp,
                  lane);
      G2_PHASE(3);
    } else if (warp == G2_NPW) {
      // =========================== MMA issuer (one thread) ============================================
      // Per 32-channel panel: 4 K steps of (A_hi W_hi, A_lo W_hi), then 4 K steps of A_hi W_lo, so the lo half of
      // a weight panel is needed 512 MMA cycles after its hi half and the hi unit is released before the lo unit.
      if (lane == 0) {
        const uint32_t idesc = make_idesc_tf32(128, Nout), idesc16 = make_idesc_bf16(128, Nout);
        // descriptors of ring stage 0 / weight unit 0 (tc_common.cuh, Desc32): everything else is a 32-bit add
        const Desc32 da = split_desc(make_smem_desc(s_a)), dw = split_desc(make_smem_desc(s_w));
        const uint32_t a_lo_off = (uint32_t)((128 * PANEL_ROW_BYTES) >> 4);
        const uint32_t a_step = (uint32_t)(G2_A_STAGE >> 4), w_step = unit_bytes >> 4;
        unsigned started = 0;
        for (int f = 0; f < C3P_NCELL; ++f) {
          const unsigned act = active[f];
          if (!act) continue;
          const int t_first = __ffs(act) - 1, t_last = 31 - __clz(act);
          for (int kb = 0; kb < a.nkb; ++kb) {
            int us[2 * NKC];
            uint32_t up[2 * NKC];
#pragma unroll
            for (int i = 0; i < 2 * NKC; ++i) {
              us[i] = m_wslot; up[i] = m_wphase;
              if (++m_wslot == NWU) { m_wslot = 0; m_wphase ^= 1u; }
            }
            for (int t = 0; t < T; ++t) {
              if (!((act >> t) & 1u)) continue;
              const uint32_t d = tmem + (uint32_t)(t * Nout);
              uint32_t acc_flag = (started >> t) & 1u;
              started |= 1u << t;
#pragma unroll
              for (int kc = 0; kc < NKC; ++kc) {
                if (t == t_first) mbar_wait(bar(G2B_W_FULL, us[2 * kc]), up[2 * kc]);
                mbar_wait(bar(G2B_A_FULL, m_aslot), m_aphase);
                tc_fence_after_sync();
                const uint32_t ah = da.lo + (uint32_t)m_aslot * a_step, al = ah + a_lo_off;
                const uint32_t wh = dw.lo + (uint32_t)us[2 * kc] * w_step, wl = dw.lo + (uint32_t)us[2 * kc + 1] * w_step;
#pragma unroll
                for (int ks = 0; ks < PANEL_K / UMMA_K; ++ks) {
                  const uint32_t adv = (uint32_t)((ks * UMMA_K * 4) >> 4);
                  mma_tf32(d, ah + adv, da.hi, wh + adv, dw.hi, idesc, acc_flag);
                  if (!BF16C) mma_tf32(d, al + adv, da.hi, wh + adv, dw.hi, idesc, 1u);
                  acc_flag = 1u;
                }
                if (t == t_last) mma_commit(bar(G2B_W_EMPTY, us[2 * kc]));
                if (t == t_first) {
                  mbar_wait(bar(G2B_W_FULL, us[2 * kc + 1]), up[2 * kc + 1]);
                  tc_fence_after_sync();
                }
#pragma unroll
                for (int ks = 0; ks < PANEL_K / UMMA_K; ++ks) {
                  const uint32_t adv = (uint32_t)((ks * UMMA_K * 4) >> 4);
                  // BF16C: [A_lo | A_hi] x [W_hi | W_lo], 64 bf16 of K in 4 steps of 16 (32 bytes each, like TF32's 8 x 4)
                  if (BF16C) mma_bf16(d, al + adv, da.hi, wl + adv, dw.hi, idesc16, 1u);
                  else mma_tf32(d, ah + adv, da.hi, wl + adv, dw.hi, idesc, 1u);
                }
                mma_commit(bar(G2B_A_EMPTY, m_aslot));
                if (t == t_last) mma_commit(bar(G2B_W_EMPTY, us[2 * kc + 1]));
                if (++m_aslot == NAS) { m_aslot = 0; m_aphase ^= 1u; }
              }
            }
          }
        }
        mma_commit(bar(G2B_ACC_FULL, 0));
      }
    } else if (warp == G2_NPW + 1) {
      // =========================== weight loader (one thread) =========================================
      if (lane == 0) {
        const int units_per_cell = a.nkb * NKC * 2;
        for (int f = 0; f < C3P_NCELL; ++f) {
          if (!active[f]) continue;
          for (int u = 0; u < units_per_cell; ++u) {
            if (l_wrap >= 1) mbar_wait(bar(G2B_W_EMPTY, l_slot), (l_wrap - 1) & 1u);
            mbar_arrive_expect_tx(bar(G2B_W_FULL, l_slot), unit_bytes);
            bulk_copy_g2s(s_w + (uint32_t)l_slot * unit_bytes,
                          a.wp + ((size_t)f * units_per_cell + u) * unit_bytes, unit_bytes, bar(G2B_W_FULL, l_slot));
            if (++l_slot == NWU) { l_slot = 0; ++l_wrap; }
          }
        }
      }
    } else {
      // =========================== item-list loader (one thread) ======================================
      if (lane == 0) {
        auto acquire = [&]() -> int {
          const int slot = i_g & (G2_NIS - 1), use = i_g / G2_NIS;
          if (use >= 1) mbar_wait(bar(G2B_IT_EMPTY, slot), (uint32_t)((use - 1) & 1));
          ++i_g;
          return slot;
        };
        for (int f = 0; f < C3P_NCELL; ++f) {
          const unsigned act = active[f];
          if (!act) continue;
          for (int kb = 0; kb < a.nkb; ++kb) {
            for (int t = 0; t < T; ++t) {
              if (!((act >> t) & 1u)) continue;
              const int slot = acquire();
              hdr[slot] = kb | (t << 8) | (f << 16);
              mbar_arrive_expect_tx(bar(G2B_IT_FULL, slot), 128 * sizeof(uint2));
              bulk_copy_g2s(s_items + (uint32_t)slot * 128 * 8, a.g_items + ((sub0 + t) * C3P_NCELL + f) * 128,
                            128 * sizeof(uint2), bar(G2B_IT_FULL, slot));
            }
          }
        }
        const int slot = acquire();
        hdr[slot] = G2_END;
        mbar_arrive(bar(G2B_IT_FULL, slot));
      }
    }
    // the next chunk overwrites the row table, the group masks and (through the MMAs) the accumulators
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
  }
  if (TIMED && timed) {
    for (int i = 0; i < 8; ++i) atomicAdd(&g2_phase_cycles[i], (unsigned long long)ph[i]);
    const unsigned long long total = (unsigned long long)(clock64() - t_begin);
    atomicAdd(&g2_cta_cycles[0], total);
    atomicMax(&g2_cta_cycles[1], total);
  }
  if (warp == G2_NPW + 1) tmem_dealloc(tmem, 512);
}

struct G2Config {
  int NKC, nkb, T, NAS, NWU;
  size_t smem;
};

// Csrc = contraction width per cell (Cin forward, Cout backward), Nout = output width.
// Shared memory = operand ring (NAS stages of 32 KB: hi + lo panel of 128 rows x 32 channels) + weight units (hi or
// lo half of one [Nout x 32] panel each) + tables.  Two 32-channel panels per group (NKC = 2: ids, weights and
// predicates paid once per 256 bytes of row) when that leaves a ring of at least three stages, else one (e.g.
// 256 -> 256, where a weight unit alone is 32 KB).  At least one weight unit beyond the 2 * NKC of the panel in use,
// so the next panel's hi half is in flight while the tensor core still reads this one's lo half.
static bool g2_config(int N, long long capacity, int Csrc, int Nout, G2Config* c) {
  if (N > 65535) return false;                // members per cell are packed in 24 bits, rows in 8
  if (capacity >= (1LL << 32)) return false;  // list positions are 32-bit
  if (Csrc % 32 || Nout % 16 || Csrc < 32 || Nout < 16 || Nout > 256) return false;
  c->T = 512 / Nout >= 4 ? 4 : (512 / Nout);
  const size_t budget = 227 * 1024 - 2048;  // static shared memory (barriers, masks) + alignment slack
  const size_t unit = (size_t)Nout * PANEL_ROW_BYTES;
  const size_t tables = (size_t)c->T * 128 * 4 + (size_t)G2_NIS * 128 * 8;
  for (int nkc = (Csrc % 64 == 0) ? 2 : 1; nkc >= 1; --nkc) {
    c->NKC = nkc;
    c->nkb = Csrc / (32 * nkc);
    c->NWU = 2 * nkc;
    if (tables + c->NWU * unit > budget) continue;
    size_t room = budget - tables - c->NWU * unit;
    c->NAS = (int)(room / G2_A_STAGE);
    if (c->NAS > G2_MAX_NAS) c->NAS = G2_MAX_NAS;
    if (c->NAS < nkc + 1) continue;
    room -= (size_t)c->NAS * G2_A_STAGE;
    int extra = (int)(room / unit);
    if (extra == 0 && c->NAS >= nkc + 2 && room + G2_A_STAGE >= unit) {   // trade a ring stage for the spare unit
      c->NAS -= 1;
      room += G2_A_STAGE;
      extra = (int)(room / unit);
    }
    if (extra > 2) extra = 2;
    c->NWU += extra;
    c->smem = (size_t)c->NAS * G2_A_STAGE + (size_t)c->NWU * unit + tables;
    return true;
  }
  return false;
}

bool gather_mma2_supported(int N, long long capacity, int Csrc, int Nout) {
  G2Config c;
  return g2_config(N, capacity, Csrc, Nout, &c);
}

// scratch of one item-list set: [items | nnz | rowid | mask]
size_t group_items_bytes(long long pts, int rows) {
  const long long subtiles = (pts + rows - 1) / rows;
  return align_up((size_t)subtiles * C3P_NCELL * rows * sizeof(uint2)) +
         align_up((size_t)subtiles * C3P_NCELL * sizeof(int)) + align_up((size_t)subtiles * rows * sizeof(int)) +
         align_up((size_t)subtiles * sizeof(unsigned)) + 256;
}

GroupItems carve_group_items(void* scratch, long long pts, int rows) {
  const long long subtiles = (pts + rows - 1) / rows;
  char* sp = static_cast<char*>(scratch);
  GroupItems gi;
  gi.subtiles = subtiles;
  gi.items = reinterpret_cast<uint2*>(sp); sp += align_up((size_t)subtiles * C3P_NCELL * rows * sizeof(uint2));
  gi.nnz = reinterpret_cast<int*>(sp); sp += align_up((size_t)subtiles * C3P_NCELL * sizeof(int));
  gi.rowid = reinterpret_cast<int*>(sp); sp += align_up((size_t)subtiles * rows * sizeof(int));
  gi.mask = reinterpret_cast<unsigned*>(sp); sp += align_up((size_t)subtiles * sizeof(unsigned));
  gi.counter = reinterpret_cast<unsigned*>(sp);
  return gi;
}

int launch_group_items(const conv3p_geom_t* g, const PlanView& v, bool backward_lists, int rows,
                       const GroupItems& gi, cudaStream_t stream, const GroupItems* half) {
  const long long pts = (long long)g->B * g->N;
  if (pts == 0) return CONV3P_OK;
  const int* cnt = backward_lists ? v.bwd_count : v.count_table;
  HalfLists h{};
  if (half) { h.items = half->items; h.nnz = half->nnz; h.rowid = half->rowid; h.mask = half->mask; h.tiles = half->subtiles; }
  if (half && rows != 128) return CONV3P_ERR_INVALID_ARGUMENT;
  {
    LaunchTimer timer_("k_group_items", stream);
#define C3P_GI_ARGS cnt, v.pair_begin, v.pair_len, v.sorted_xyzi, pts, g->pair_capacity, g->N, gi.items, gi.nnz, gi.rowid, \
                    gi.mask, gi.counter, h
    if (rows == 128 && half) k_group_items<128, true><<<(unsigned)gi.subtiles, 128, 0, stream>>>(C3P_GI_ARGS);
    else if (rows == 128) k_group_items<128, false><<<(unsigned)gi.subtiles, 128, 0, stream>>>(C3P_GI_ARGS);
    else if (rows == 64) k_group_items<64, false><<<(unsigned)gi.subtiles, 64, 0, stream>>>(C3P_GI_ARGS);
    else k_group_items<32, false><<<(unsigned)gi.subtiles, 32, 0, stream>>>(C3P_GI_ARGS);
#undef C3P_GI_ARGS
  }
  C3P_LAUNCH_CHECK("k_group_items");
  return CONV3P_OK;
}

size_t gather_mma2_scratch_bytes(const conv3p_geom_t* g) { return group_items_bytes((long long)g->B * g->N, 128); }

int launch_gather_mma2(const conv3p_geom_t* g, const PlanView& v, const float* src, const void* wp, int Csrc,
                       int Nout, float* out, bool weighted, void* scratch, size_t scratch_bytes, const char* name,
                       cudaStream_t stream, float* g_store, const RowIO& io, const GroupItems* half_items) {
  G2Config c;
  const long long pts = (long long)g->B * g->N;
  if (!g2_config(g->N, g->pair_capacity, Csrc, Nout, &c)) return CONV3P_ERR_UNSUPPORTED;
  if (pts == 0) return CONV3P_OK;
  if (!scratch || scratch_bytes < gather_mma2_scratch_bytes(g)) return CONV3P_ERR_BUFFER_TOO_SMALL;
  const GroupItems gi = carve_group_items(scratch, pts, 128);
  const long long subtiles = gi.subtiles;
  {
    const int st = launch_group_items(g, v, weighted, 128, gi, stream, half_items);
    if (st) return st;
  }
  G2Args a{};
  a.src = src; a.wp = static_cast<const unsigned char*>(wp); a.out = out;
  a.rows = weighted ? v.bwd_row : v.pair_row;
  a.weights = weighted ? v.bwd_weight : nullptr;
  a.g_items = gi.items; a.g_nnz = gi.nnz; a.g_rowid = gi.rowid; a.counter = gi.counter;
  a.total_points = pts; a.subtiles = subtiles;
  a.N = g->N; a.Csrc = Csrc; a.Nout = Nout;
  a.nkb = c.nkb; a.T = c.T; a.NAS = c.NAS; a.NWU = c.NWU;
  a.debug = engine_flag(32) ? 32 : 0;   // phase timers (tools/engine_timing.py)
  a.g_store = weighted ? g_store : nullptr;
  // scheduler reserves: 6 G and 2 G sub-tiles (tuned at 2048 sub-tiles; CONV3P_SCHED="big,one", read once per
  // process, overrides them for sweeps)
  static const std::pair<int, int> sched = [] {
    std::pair<int, int> r(12, 4);
    if (const char* e = getenv("CONV3P_SCHED")) {
      int x = 0, y = 0;
      if (sscanf(e, "%d,%d", &x, &y) == 2 && x >= y && y >= 0) r = std::make_pair(x, y);
    }
    return r;
  }();
  a.res_big = sched.first; a.res_one = sched.second;
  a.src_stride = io.src_stride ? io.src_stride : Csrc;
  a.out_stride = io.out_stride ? io.out_stride : Nout;
  a.activation = io.activation;
  // 16-byte vector accesses: rows must start on 16-byte boundaries
  if (a.src_stride % 4 || a.out_stride % 4 || reinterpret_cast<uintptr_t>(src) % 16 ||
      reinterpret_cast<uintptr_t>(out) % 16 || a.src_stride >= (1LL << 29))
    return CONV3P_ERR_UNSUPPORTED;
  const int sms = sm_count();
  const long long tiles = subtiles < sms ? subtiles : sms;  // persistent CTAs, one per SM
  auto launch = [&](auto kern) -> int {
    const int st_ = ensure_dynamic_smem(kern, 227 * 1024);   // once per (kernel, device); granted minus static shared memory
    if (st_) return st_;
    {
      LaunchTimer timer_(name, stream);
      kern<<<(unsigned)tiles, G2_THREADS, c.smem, stream>>>(a);
    }
    C3P_LAUNCH_CHECK(name);
    return CONV3P_OK;
  };
  const bool tf32x3 = engine_flag(512);   // three TF32 products instead of TF32 + BF16 corrections (A/B, accuracy reference)
  if (a.debug & 32) {   // phase timers compiled in (tools/engine_timing.py); production split only
    if (c.NKC == 1) return weighted ? launch(k_gather_mma2<1, true, true, true>) : launch(k_gather_mma2<1, false, true, true>);
    return weighted ? launch(k_gather_mma2<2, true, true, true>) : launch(k_gather_mma2<2, false, true, true>);
  }
  if (tf32x3) {
    if (c.NKC == 1) return weighted ? launch(k_gather_mma2<1, true, false, false>) : launch(k_gather_mma2<1, false, false, false>);
    return weighted ? launch(k_gather_mma2<2, true, false, false>) : launch(k_gather_mma2<2, false, false, false>);
  }
  if (c.NKC == 1) return weighted ? launch(k_gather_mma2<1, true, false, true>) : launch(k_gather_mma2<1, false, false, true>);
  return weighted ? launch(k_gather_mma2<2, true, false, true>) : launch(k_gather_mma2<2, false, false, true>);
}

}  // namespace c3p

// profiling helper (tools/engine_timing.py): read and clear the phase timers of k_gather_mma2
extern "C" int conv3p_debug_cta_cycles(unsigned long long* host2) {
  unsigned long long zero[2] = {0, 0};
  if (cudaMemcpyFromSymbol(host2, c3p::g2_cta_cycles, sizeof(zero)) != cudaSuccess) return CONV3P_ERR_CUDA;
  if (cudaMemcpyToSymbol(c3p::g2_cta_cycles, zero, sizeof(zero)) != cudaSuccess) return CONV3P_ERR_CUDA;
  return CONV3P_OK;
}

extern "C" int conv3p_debug_phase_cycles(unsigned long long* host8) {
  unsigned long long zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (cudaMemcpyFromSymbol(host8, c3p::g2_phase_cycles, sizeof(zero)) != cudaSuccess) return CONV3P_ERR_CUDA;
  if (cudaMemcpyToSymbol(c3p::g2_phase_cycles, zero, sizeof(zero)) != cudaSuccess) return CONV3P_ERR_CUDA;
  return CONV3P_OK;
}
