// backward_filter2.cu -- weight gradient on the 5th-gen tensor cores.
//
//   grad_filter[f, k, c] = sum_j input[j, k] * G_f[j, c],    G_f[j, :] = sum_{(ii,w) in cell f of j} w * grad_out[ii, :]
//
// (tf_conv3p_atrous.cpp:694-696 regrouped by (j, f').)  The contraction runs over POINTS: for a tile of PTS
// voxel-sorted points and one "virtual accumulator" the tensor core computes D[m, k] += A^T X with MN-major 3xTF32
// operands (point index = panel row), M = 128 TMEM lanes, N = Cin, K = 8 points per instruction:
//   * Cout >= 128: the 128 lanes are one 128-channel block of one cell (Cout = 256: two blocks per cell);
//   * Cout <  128: the lanes stack 128 / Cout CELLS of Cout channels each -- different cells share the B operand X
//     (the tile's input rows), so stacking keeps M = 128 without an M = 64 instruction or padded lanes.
// 512 / Cin accumulators live in TMEM at once (a "pass" over the CTA's tiles); persistent CTA per SM, per-CTA
// partials, deterministic ordered reduce (launch_reduce_partials).  Work items come from k_group_items (PTS-row
// lists of the backward lists, compacted by population class) through a bulk-copy ring; a quarter-warp gathers all
// the channels of its row (packed FFMA2, weight-0 padding).  When grad_input is computed in the same call by
// k_gather_mma2, that kernel has already aggregated exactly these G_f rows and left them in the G store
// ([sorted position][27][Cout], only non-empty (point, cell) slots are written): the producers then read ONE row
// per non-empty item (prefetched one stage ahead) instead of walking its list (template FROM_STORE).
//
// Warp roles: warps [0, NPW) producers (also the flush), NPW = MMA issuer + TMEM allocator, NPW+1 = item-list loader;
// NPW = PTS / 4 (one quarter-warp per point of the tile).
//
// What bounds it (phase timers of C3P_W2_TIMED builds, tools/mma_rate.py, ncu; profiles/r2_summary.md): an M=128 x
// N=64 tcgen05.mma occupies the tensor core for ~75 cycles whatever the operand layout or type (~43 cycles fixed +
// N/2), so the 16 instructions of a 64-point stage take 1200 cycles -- the floor of this formulation, 0.46 ms at the
// headline shape; on top of that the L1/shared-memory data pipe carries the operand reads of those MMAs (96 KB per
// stage), the producers' panel stores (64 KB) and the row loads, ~1700 wavefront cycles per stage, and the producers
// spend 40 % of their time queueing behind it when they issue the next stage's loads.  Measured and dropped: two
// stages of register look-ahead (0.99 -> 1.08 ms), a ninth warp prefetching the G-store rows into L2 eight stages
// ahead with cp.async.bulk.prefetch.L2 (0.87 -> 1.15 ms: the stall is not DRAM latency), 32-point tiles with a 4- or
// 6-stage ring (1.56 ms), mbarrier waits without a suspend-time hint (no change).
#include <cstdio>
#include <cstdlib>
#include <utility>

#include "common.cuh"
#include "tc_common.cuh"
#include "tc_gather2.cuh"

namespace c3p {

using namespace tc;

// Compile-time switch (tools/build_variants.py): C3P_W2_TIMED = 1 compiles in the phase timers of producer warp 0
// (conv3p_debug_w2_cycles, tools/ab_backward.py).
#ifndef C3P_W2_TIMED
#define C3P_W2_TIMED 0
#endif

// cycles of producer warp 0, lane 0, summed over CTAs: items wait | first use of the prefetched rows | ring-slot wait |
// split + stores + fence + arrive | input panels | flush | total
// [8..15]: item loader: wait for a free item slot | total;  MMA issuer: wait input panels | wait G stage | wait flush | total
__device__ unsigned long long w2_phase_cycles[16];

constexpr int W2_NIS = 8;      // item-list slots
constexpr int W2_MASKS = 1024; // cell masks of the CTA's tiles kept in shared memory (more tiles: read through L2)
constexpr int W2_MAX_NGS = 6;  // G ring stages
constexpr int W2_END = -1;

struct W2Args {
  const float* grad_out;   // [B*N, Cout]
  const float* input;      // [B*N, Cin]
  const int* rows;         // backward lists: rows ii
  const float* weights;    // backward lists: 1 / count(ii, f')
  const uint2* g_items;    // [tiles][27][PTS]
  const int* g_rowid;      // [tiles*PTS]
  const unsigned* g_mask;  // [tiles] bit f: some point of the tile has members in cell f
  float* partial;          // [gridDim.x][27*Cin*Cout]
  const float* g_store;    // FROM_STORE: G_f rows [tiles*PTS][27][Cout] written by k_gather_mma2
  long long total_points, tiles;
  int Cin, Cout;
  int FG;    // accumulators per pass (512 / Cin)
  int NGS;   // G ring stages
  int NXB;   // X buffers (1 or 2)
};

// GP = 32-channel panels of G per cell inside one accumulator (Cout = 32, 64: GP = 1, 2 and 4 / GP cells stacked;
// Cout = 128, 256: GP = 4, one or two 128-channel blocks per cell).
template <int GP>
struct W2Map {
  static constexpr int CS = 4 / GP;   // cells stacked in one accumulator
  // number of virtual accumulators; mbs = log2(128-channel blocks per cell) (GP == 4 only)
  __host__ __device__ static int count(int mbs) { return GP == 4 ? C3P_NCELL << mbs : (C3P_NCELL + CS - 1) / CS; }
  // sub-mask (CS bits) of the accumulator's cells that have members in a tile with cell mask m
  __host__ __device__ static unsigned cells(unsigned m, int va, int mbs) {
    if (GP == 4) return (m >> (va >> mbs)) & 1u;
    return (m >> (va * CS)) & ((1u << CS) - 1u);
  }
  // bit i: accumulator va0 + i (i < n) has members in a tile with cell mask m
  __host__ __device__ static unsigned pass_mask(unsigned m, int va0, int n, int mbs) {
    if (GP == 4 && mbs == 0) return (m >> va0) & ((n >= 32) ? ~0u : ((1u << n) - 1u));
    unsigned r = 0;
    for (int i = 0; i < n; ++i)
      if (cells(m, va0 + i, mbs)) r |= 1u << i;
    return r;
  }
};

// Bytes of one input (X) buffer: the TF32 hi panels + the second part (fp32 lo panels, or the BF16 correction panels
// [hi ; lo] stacked along the point index: ceil(Cin / 64) panels of 2 * PTS rows) -- sized for either split.
__host__ __device__ inline uint32_t w2_x_half(int Cin, int pts) { return (uint32_t)(Cin / 32) * pts * PANEL_ROW_BYTES; }
__host__ __device__ inline uint32_t w2_x_second(int Cin, int pts) { return (uint32_t)((Cin + 63) / 64) * 2u * pts * PANEL_ROW_BYTES; }

// BF16C: TF32 main product + BF16 correction products (tc_common.cuh); false = three TF32 products (engine flag 512).
template <bool FROM_STORE, int GP, int PTS, bool BF16C>
__global__ void __launch_bounds__((PTS / 4 + 2) * 32, 1) k_backward_filter2(const W2Args a) {
  constexpr int NPW = PTS / 4;                         // producer warps: one quarter-warp per row of the tile
  constexpr int CS = W2Map<GP>::CS;
  constexpr uint32_t PANEL = (uint32_t)PTS * PANEL_ROW_BYTES;   // PTS rows x 32 fp32
  constexpr uint32_t PANEL16 = 2u * PANEL;             // BF16 correction panel: 2 * PTS rows x 64 bf16 ([lo ; hi] or [hi ; lo])
  constexpr uint32_t g_half = 4u * PANEL;              // hi part of a G stage: 4 panels = 128 lanes; the second part
                                                       // (fp32 lo panels, or 2 BF16 correction panels) has the same size
  extern __shared__ __align__(1024) unsigned char smem[];
  const int Cin = a.Cin, Cout = a.Cout, FG = a.FG, NGS = a.NGS, NXB = a.NXB;
  const int xp = Cin / 32;                             // panels of the input rows
  const int mbs = (GP == 4 && Cout == 256) ? 1 : 0;    // log2(128-channel blocks per cell)
  const int NVA = W2Map<GP>::count(mbs);
  const uint32_t x_half = w2_x_half(Cin, PTS);
  const uint32_t x_buf = x_half + w2_x_second(Cin, PTS);
  unsigned char* g_base = smem;                                     // NGS stages x (hi, second)
  unsigned char* x_base = g_base + (size_t)NGS * 2 * g_half;        // NXB buffers x (hi, second)
  uint2* items = reinterpret_cast<uint2*>(x_base + (size_t)NXB * x_buf);  // [NIS][CS][PTS]
  __shared__ uint64_t g_full[W2_MAX_NGS], g_empty[W2_MAX_NGS], x_full[2], x_empty[2], it_full[W2_NIS],
      it_empty[W2_NIS], acc_full, acc_empty;
  __shared__ uint32_t tmem_slot;
  __shared__ int hdr[W2_NIS];   // accumulator | cell sub-mask << 8 | tile (CTA-local) << 12, or W2_END
  // Cell masks of this CTA's tiles: every role walks them once per pass, the MMA issuer and the item-list loader
  // with nothing to hide a global-load latency behind (measured: 36 % of the producers' time was spent waiting for
  // item lists the loader had not requested yet).
  __shared__ unsigned tile_mask[W2_MASKS];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long per_cta = (a.tiles + gridDim.x - 1) / gridDim.x;
  const long long tile_lo = (long long)blockIdx.x * per_cta;
  const long long tile_hi = min(a.tiles, tile_lo + per_cta);
  const int npass = (NVA + FG - 1) / FG;
  const unsigned max_row = (unsigned)(a.total_points - 1);

  if (warp == NPW) {
    if (lane == 0) {
      for (int i = 0; i < NGS; ++i) {
        mbar_init(&g_full[i], NPW);
        mbar_init(&g_empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&x_full[i], NPW);
        mbar_init(&x_empty[i], 1);
      }
      for (int i = 0; i < W2_NIS; ++i) {
        mbar_init(&it_full[i], 1);
        mbar_init(&it_empty[i], NPW);
      }
      mbar_init(&acc_full, 1);
      mbar_init(&acc_empty, NPW);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(&tmem_slot, 512);
  }
  for (long long t = tile_lo + tid; t < tile_hi && t - tile_lo < W2_MASKS; t += blockDim.x)
    tile_mask[t - tile_lo] = __ldg(a.g_mask + t);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  auto cell_mask = [&](long long tile) -> unsigned {
    const long long i = tile - tile_lo;
    return i < W2_MASKS ? tile_mask[i] : __ldg(a.g_mask + tile);
  };

  // mask of the pass's accumulators that have members in a tile with cell mask m (bit = accumulator - va0)
  auto pass_vas = [&](unsigned m, int va0, int va1) -> unsigned { return W2Map<GP>::pass_mask(m, va0, va1 - va0, mbs); };

  if (warp < NPW) {
    // =========================== producers ===========================================================
    // Quarter-warp q serves row q of the tile for the input panels and item (q + 4g) mod PTS of every cell of
    // stage g.
    const int q = warp * 4 + (lane >> 3), l8 = lane & 7;
#if C3P_W2_TIMED
    const bool timed = tid == 0;
    long long tk = clock64();
    const long long t_begin = tk;
    unsigned long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define W2_PHASE(i) do { if (timed) { const long long t_ = clock64(); ph[i] += (unsigned long long)(t_ - tk); tk = t_; } } while (0)
#else
#define W2_PHASE(i) do { } while (0)
#endif
    struct Stage {
      G2Item it[CS];
      int h;   // header of the stage
    };
    auto read_stage = [&](int g_, Stage& s) {
      const int slot = g_ & (W2_NIS - 1);
      W2_PHASE(0);
      mbar_wait(&it_full[slot], (uint32_t)((g_ / W2_NIS) & 1));
      W2_PHASE(7);
      s.h = hdr[slot];
      const unsigned sub = s.h == W2_END ? 0u : ((unsigned)s.h >> 8) & 15u;
#pragma unroll
      for (int cs = 0; cs < CS; ++cs) {
        const uint2 u = items[(slot * CS + cs) * PTS + ((q + 4 * g_) & (PTS - 1))];
        G2Item& it = s.it[cs];
        it.pos = u.x; it.p = (int)(u.y & 255u);
        it.n = ((sub >> cs) & 1u) ? (int)(u.y >> 8) : 0;
        if (!((sub >> cs) & 1u)) { it.pos = 0u; it.p = (q + 4 * g_) & (PTS - 1); }   // no list was loaded: own row, zeros
        it.inv = 0.f; it.w = 0.f; it.ids = 0;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&it_empty[slot]);
      if (!FROM_STORE) {
#pragma unroll
        for (int cs = 0; cs < CS; ++cs) g2_prefetch<true>(s.it[cs], a.rows, a.weights, 0, l8, max_row);
      }
    };
    // FROM_STORE: the aggregated rows of the stage's items (index cs * GP + kc)
    auto fetch_rows = [&](const Stage& s, float4 (&v)[4]) {
      const int va = s.h & 255;
      const long long tile = tile_lo + (s.h >> 12);
#pragma unroll
      for (int cs = 0; cs < CS; ++cs) {
        const int f = GP == 4 ? va >> mbs : va * CS + cs;
        const int ch0 = GP == 4 ? (va & ((1 << mbs) - 1)) * 128 : 0;
        const float* src = a.g_store + (((size_t)tile * PTS + s.it[cs].p) * C3P_NCELL + f) * Cout + ch0 + l8 * 4;
#pragma unroll
        for (int kc = 0; kc < GP; ++kc)
          v[cs * GP + kc] = s.it[cs].n > 0 ? ldg4(src + kc * PANEL_K) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    int g_slot = 0, x_slot = 0;
    uint32_t g_wrap = 0, x_wrap = 0;   // completed trips around the G ring / the X buffers
    // One stage of look-ahead in two STATIC register buffers (A for even stages, B for odd ones): the items of stage
    // g + 1, with their first list ids (gather mode) or their G rows (store mode), are requested at the start of stage
    // g.  Measured with the phase timers: the rows have arrived when they are needed (1 % of the producer time); two
    // stages of look-ahead were slower (0.99 -> 1.08 ms), and so was a rotating register queue (moves of loaded values
    // stall like uses).
    Stage stA, stB;
    float4 rowsA[4], rowsB[4];
    bool odd_stage = false;
    int n_read = 0;
    bool ended = false;
    auto pull = [&](Stage& s, float4 (&r)[4]) {
      if (ended) {                 // nothing is published after the END marker
        s.h = W2_END;
#pragma unroll
        for (int cs = 0; cs < CS; ++cs) { s.it[cs].n = 0; s.it[cs].p = 0; s.it[cs].pos = 0u; }
        return;
      }
      read_stage(n_read++, s);
      ended = s.h == W2_END;
      if (FROM_STORE && !ended) fetch_rows(s, r);
    };
    pull(stA, rowsA);
    W2_PHASE(0);
    for (int pass = 0; pass < npass; ++pass) {
      const int va0 = pass * FG, va1 = min(NVA, va0 + FG);
      unsigned pass_mask = 0;
      // Input rows are requested one tile ahead and their row ids two tiles ahead (measured: loading them at the
      // point of use cost a quarter of the producers' time, one exposed DRAM latency per tile visit and pass).
      constexpr int XPRE = PTS == 64 ? 2 : 4;   // prefetched 32-channel panels of the row (the rest is loaded in place)
      auto rid_at = [&](long long t) -> int { return t < tile_hi ? __ldg(a.g_rowid + t * PTS + q) : -1; };
      float4 xv[XPRE];
      auto x_request = [&](int rid) {
        const float* xr = a.input + (size_t)(rid >= 0 ? rid : 0) * Cin + l8 * 4;
#pragma unroll
        for (int pnl = 0; pnl < XPRE; ++pnl)
          xv[pnl] = (rid >= 0 && pnl < xp) ? ldg4(xr + pnl * PANEL_K) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      int rid0 = rid_at(tile_lo), rid1 = rid_at(tile_lo + 1);
      x_request(rid0);
      unsigned mask_next = tile_lo < tile_hi ? pass_vas(cell_mask(tile_lo), va0, va1) : 0u;
      for (long long tile = tile_lo; tile < tile_hi; ++tile) {
        const unsigned mask = mask_next;
        const int row = rid0;
        float4 xc[XPRE];
#pragma unroll
        for (int pnl = 0; pnl < XPRE; ++pnl) xc[pnl] = xv[pnl];
        rid0 = rid1;
        rid1 = rid_at(tile + 2);
        x_request(rid0);
        if (tile + 1 < tile_hi) mask_next = pass_vas(cell_mask(tile + 1), va0, va1);
        if (!mask) continue;
        pass_mask |= mask;
        // ---- input rows of the tile -> X panels (hi/lo) ---------------------------------------------------
        {
          const int xb = x_slot;
          if (x_wrap >= 1) mbar_wait(&x_empty[xb], (x_wrap - 1) & 1u);
          if (++x_slot == NXB) { x_slot = 0; ++x_wrap; }
          unsigned char* xs = x_base + (size_t)xb * x_buf;
          const float* xr = a.input + (size_t)(row >= 0 ? row : 0) * Cin + l8 * 4;
          auto x_store = [&](int pnl, const float4& v) {
            unsigned char* hi_dst = xs + (size_t)pnl * PANEL + panel_chunk_offset_mn(q, l8);
            if (BF16C) {   // B side of the correction product: rows [0, PTS) hi, [PTS, 2 PTS) lo (see the G stores)
              unsigned char* c = xs + x_half + (size_t)(pnl >> 1) * PANEL16;
              const int j = (pnl & 1) * 32 + 4 * l8;
              g2_store_split16(hi_dst, c + panel_offset16(PTS + q, j), c + panel_offset16(q, j), v);
            } else {
              g2_store_split(hi_dst, x_half, v);
            }
          };
#pragma unroll
          for (int pnl = 0; pnl < XPRE; ++pnl)
            if (pnl < xp) x_store(pnl, xc[pnl]);
          for (int pnl = XPRE; pnl < xp; ++pnl)
            x_store(pnl, row >= 0 ? ldg4(xr + pnl * PANEL_K) : make_float4(0.f, 0.f, 0.f, 0.f));
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&x_full[xb]);
          W2_PHASE(4);
        }
        // ---- one G stage per active accumulator -----------------------------------------------------------
        auto do_stage = [&](Stage& st_buf, float4 (&row_buf)[4], Stage& st_other, float4 (&row_other)[4]) {
          const Stage cur = st_buf;
          pull(st_other, row_other);     // stage g + 1 into the other buffer, before this one's rows are touched
          W2_PHASE(0);
          float4 acc[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] = row_buf[i];
#if C3P_W2_TIMED
          if (FROM_STORE) {   // force the arrival of the prefetched rows here so the wait is attributed to phase 1
#pragma unroll
            for (int i = 0; i < 4; ++i) asm volatile("" ::"f"(acc[i].x), "f"(acc[i].y), "f"(acc[i].z), "f"(acc[i].w));
          }
          W2_PHASE(1);
#endif
          const int slot = g_slot;
          const uint32_t use = g_wrap;
          if (++g_slot == NGS) { g_slot = 0; ++g_wrap; }
          unsigned char* stage = g_base + (size_t)slot * 2 * g_half;
          bool waited = false;
#pragma unroll
          for (int cs = 0; cs < CS; ++cs) {
            G2Item item = cur.it[cs];
            if (!FROM_STORE) {
              int nmax = max(item.n, __shfl_xor_sync(C3P_FULL_MASK, item.n, 8));
              nmax = max(nmax, __shfl_xor_sync(C3P_FULL_MASK, nmax, 16));
              const int ch0 = GP == 4 ? ((cur.h & 255) & ((1 << mbs) - 1)) * 128 : 0;
              float4 part[GP];
              g2_gather<GP, 8 / GP, true>(part, item, nmax, a.grad_out, Cout, ch0, a.rows, a.weights, l8, max_row);
#pragma unroll
              for (int kc = 0; kc < GP; ++kc) acc[cs * GP + kc] = part[kc];
            }
            if (!waited) {
              if (use >= 1) mbar_wait(&g_empty[slot], (use - 1) & 1u);
              waited = true;
              W2_PHASE(2);
            }
            const int p = item.p;
            unsigned char* dst = stage + panel_chunk_offset_mn(p, l8);
#pragma unroll
            for (int kc = 0; kc < GP; ++kc) {
              const int pi = cs * GP + kc;                  // 32-lane panel of the accumulator's M = 128
              if (BF16C) {
                // A side of the correction product: rows [0, PTS) lo, [PTS, 2 PTS) hi
                unsigned char* c = stage + g_half + (size_t)(pi >> 1) * PANEL16;
                const int j = (pi & 1) * 32 + 4 * l8;
                g2_store_split16(dst + (size_t)pi * PANEL, c + panel_offset16(p, j), c + panel_offset16(PTS + p, j), acc[pi]);
              } else {
                g2_store_split(dst + (size_t)pi * PANEL, g_half, acc[pi]);
              }
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&g_full[slot]);
          W2_PHASE(3);
        };
        for (unsigned todo = mask; todo; todo &= todo - 1) {
          if (odd_stage) do_stage(stB, rowsB, stA, rowsA); else do_stage(stA, rowsA, stB, rowsB);
          odd_stage = !odd_stage;
        }
      }
      // ---- flush this pass's accumulators: partial[cta][f][k][c] = D[lane(c), k] ----------------------------
      mbar_wait(&acc_full, (uint32_t)(pass & 1));
      tc_fence_after_sync();
      {
        const int sub = warp & 3;
        const int L = sub * 32 + lane;                     // TMEM lane
        for (int ai = warp >> 2; ai < va1 - va0; ai += NPW / 4) {
          const int va = va0 + ai;
          const int f = GP == 4 ? va >> mbs : va * CS + L / (GP * 32);
          const int c = GP == 4 ? (va & ((1 << mbs) - 1)) * 128 + L : L % (GP * 32);
          const bool live = (pass_mask >> ai) & 1u;
          float* dst = a.partial + ((size_t)blockIdx.x * C3P_NCELL + (f < C3P_NCELL ? f : 0)) * Cin * Cout;
          for (int k0 = 0; k0 < Cin; k0 += 32) {
            float v[32];
            if (live) {
              tmem_ld_32x32(tmem + ((uint32_t)(sub * 32) << 16) + (uint32_t)(ai * Cin + k0), v);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
            if (f < C3P_NCELL) {
#pragma unroll
              for (int j = 0; j < 32; ++j) dst[(size_t)(k0 + j) * Cout + c] = v[j];
            }
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty);
      W2_PHASE(5);
    }
#if C3P_W2_TIMED
    if (timed) {
      ph[6] = (unsigned long long)(clock64() - t_begin);
      for (int i = 0; i < 8; ++i) atomicAdd(&w2_phase_cycles[i], ph[i]);
    }
#endif
  } else if (warp == NPW) {
    // =========================== MMA issuer (one thread) ===============================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32_mn(128, Cin), idesc16 = make_idesc_bf16_mn(128, Cin);
      // descriptors of ring stage 0 / input buffer 0 (tc_common.cuh, Desc32): everything else is a 32-bit add
      const Desc32 dg = split_desc(make_smem_desc_mn(smem_u32(g_base), PANEL));
      const Desc32 dx = split_desc(make_smem_desc_mn(smem_u32(x_base), PANEL));
      const Desc32 dg16 = split_desc(make_smem_desc_mn16(smem_u32(g_base) + g_half, PANEL16));
      const Desc32 dx16 = split_desc(make_smem_desc_mn16(smem_u32(x_base) + x_half, PANEL16));
      const uint32_t g_step = (2u * g_half) >> 4, x_step = x_buf >> 4, g_lo_off = g_half >> 4, x_lo_off = x_half >> 4;
      int g_slot = 0, x_slot = 0;
      uint32_t g_phase = 0, x_phase = 0;
#if C3P_W2_TIMED
      long long mk = clock64();
      const long long m_begin = mk;
      unsigned long long mph[3] = {0, 0, 0};
#define W2_MPHASE(i) do { const long long t_ = clock64(); mph[i] += (unsigned long long)(t_ - mk); mk = t_; } while (0)
#define W2_MTICK() do { mk = clock64(); } while (0)
#else
#define W2_MPHASE(i) do { } while (0)
#define W2_MTICK() do { } while (0)
#endif
      for (int pass = 0; pass < npass; ++pass) {
        const int va0 = pass * FG, va1 = min(NVA, va0 + FG);
        unsigned started = 0;
        if (pass > 0) {
          W2_MTICK();
          mbar_wait(&acc_empty, (uint32_t)((pass - 1) & 1));
          W2_MPHASE(2);
          tc_fence_after_sync();
        }
        for (long long tile = tile_lo; tile < tile_hi; ++tile) {
          const unsigned mask = pass_vas(cell_mask(tile), va0, va1);
          if (!mask) continue;
          const int xb = x_slot;
          W2_MTICK();
          mbar_wait(&x_full[xb], x_phase);
          W2_MPHASE(0);
          if (++x_slot == NXB) { x_slot = 0; x_phase ^= 1u; }
          const uint32_t xw = dx.lo + (uint32_t)xb * x_step, xw16 = dx16.lo + (uint32_t)xb * x_step;
          const int ai_last = 31 - __clz(mask);
          for (unsigned todo = mask; todo; todo &= todo - 1) {
            const int ai = __ffs(todo) - 1;
            const int slot = g_slot;
            W2_MTICK();
            mbar_wait(&g_full[slot], g_phase);
            W2_MPHASE(1);
            if (++g_slot == NGS) { g_slot = 0; g_phase ^= 1u; }
            tc_fence_after_sync();
            const uint32_t gw = dg.lo + (uint32_t)slot * g_step, gw16 = dg16.lo + (uint32_t)slot * g_step;
            const uint32_t d = tmem + (uint32_t)(ai * Cin);
            const uint32_t first = (started >> ai) & 1u;
#pragma unroll
            for (int j = 0; j < PTS / 8; ++j) {          // 8 points (1 KB of every panel) per TF32 instruction
              const uint32_t adv = (uint32_t)j * 64u;
              mma_tf32(d, gw + adv, dg.hi, xw + adv, dx.hi, idesc, j ? 1u : first);
              if (!BF16C) {
                mma_tf32(d, gw + g_lo_off + adv, dg.hi, xw + adv, dx.hi, idesc, 1u);
                mma_tf32(d, gw + adv, dg.hi, xw + x_lo_off + adv, dx.hi, idesc, 1u);
              }
            }
            if (BF16C) {   // [G_lo ; G_hi]^T x [X_hi ; X_lo]: 2 * PTS rows in steps of 16 (2 KB of every panel)
#pragma unroll
              for (int j = 0; j < PTS / 8; ++j) {
                const uint32_t adv = (uint32_t)j * 128u;
                mma_bf16(d, gw16 + adv, dg16.hi, xw16 + adv, dx16.hi, idesc16, 1u);
              }
            }
            started |= 1u << ai;
            mma_commit(&g_empty[slot]);
            if (ai == ai_last) mma_commit(&x_empty[xb]);
          }
        }
        mma_commit(&acc_full);
      }
#if C3P_W2_TIMED
      atomicAdd(&w2_phase_cycles[10], mph[0]);
      atomicAdd(&w2_phase_cycles[11], mph[1]);
      atomicAdd(&w2_phase_cycles[12], mph[2]);
      atomicAdd(&w2_phase_cycles[13], (unsigned long long)(clock64() - m_begin));
#endif
    }
  } else {
    // =========================== item-list loader (one thread) ==========================================
    if (lane == 0) {
      int g = 0;
#if C3P_W2_TIMED
      const long long l_begin = clock64();
      unsigned long long l_wait = 0;
#endif
      for (int pass = 0; pass < npass; ++pass) {
        const int va0 = pass * FG, va1 = min(NVA, va0 + FG);
        for (long long tile = tile_lo; tile < tile_hi; ++tile) {
          const unsigned cellmask = cell_mask(tile);
          for (unsigned todo = pass_vas(cellmask, va0, va1); todo; todo &= todo - 1) {
            const int va = va0 + __ffs(todo) - 1;
            const unsigned sub = W2Map<GP>::cells(cellmask, va, mbs);
            const int slot = g & (W2_NIS - 1), use = g / W2_NIS;
#if C3P_W2_TIMED
            const long long lw = clock64();
#endif
            if (use >= 1) mbar_wait(&it_empty[slot], (uint32_t)((use - 1) & 1));
#if C3P_W2_TIMED
            l_wait += (unsigned long long)(clock64() - lw);
#endif
            hdr[slot] = va | ((int)sub << 8) | ((int)(tile - tile_lo) << 12);
            mbar_arrive_expect_tx(&it_full[slot], (uint32_t)__popc(sub) * PTS * (uint32_t)sizeof(uint2));
#pragma unroll
            for (int cs = 0; cs < CS; ++cs) {
              if (!((sub >> cs) & 1u)) continue;
              const int f = GP == 4 ? va >> mbs : va * CS + cs;
              bulk_copy_g2s(items + (slot * CS + cs) * PTS, a.g_items + (tile * C3P_NCELL + f) * PTS,
                            PTS * sizeof(uint2), &it_full[slot]);
            }
            ++g;
          }
        }
      }
      const int slot = g & (W2_NIS - 1), use = g / W2_NIS;
      if (use >= 1) mbar_wait(&it_empty[slot], (uint32_t)((use - 1) & 1));
      hdr[slot] = W2_END;
      mbar_arrive(&it_full[slot]);
#if C3P_W2_TIMED
      atomicAdd(&w2_phase_cycles[8], l_wait);
      atomicAdd(&w2_phase_cycles[9], (unsigned long long)(clock64() - l_begin));
#endif
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == NPW) tmem_dealloc(tmem, 512);
}

struct W2Config {
  int GP, PTS, FG, NGS, NXB;
  size_t smem;
};

// Tile of 64 points when two G stages and the input panels fit next to each other, else 32; with the smaller tile
// the ring takes as many stages as fit (up to W2_MAX_NGS).  CONV3P_W2="pts,stages" (read once) overrides for sweeps.
static bool w2_config(int N, long long capacity, int Cin, int Cout, W2Config* c) {
  if (N > 65535 || capacity >= (1LL << 32)) return false;
  if (Cout != 32 && Cout != 64 && Cout != 128 && Cout != 256) return false;   // M = 128 lanes: see the header comment
  if (Cin % 32 || Cin < 32 || Cin > 256) return false;                         // N of the MMA, in 32-wide MN-major panels
  c->GP = Cout >= 128 ? 4 : Cout / 32;
  c->FG = 512 / Cin;
  static const std::pair<int, int> forced = [] {
    std::pair<int, int> r(0, 0);
    if (const char* e = getenv("CONV3P_W2")) {
      int x = 0, y = 0;
      if (sscanf(e, "%d,%d", &x, &y) == 2 && (x == 32 || x == 64) && y >= 2 && y <= W2_MAX_NGS) r = std::make_pair(x, y);
    }
    return r;
  }();
  const size_t budget = 227 * 1024 - 1024 - W2_MASKS * sizeof(unsigned);   // static: barriers, headers, tile masks
  const int CS = 4 / c->GP;
  // preference: 64-point tiles with two input buffers; else 32-point tiles (two input buffers, deeper ring); else a
  // single input buffer
  const int order[4][2] = {{64, 2}, {32, 2}, {64, 1}, {32, 1}};
  for (int o = 0; o < 4; ++o) {
    const int pts = order[o][0], nxb = order[o][1];
    if (forced.first && pts != forced.first) continue;
    const size_t panel = (size_t)pts * PANEL_ROW_BYTES;
    const size_t g_stage = 2 * 4 * panel, x_buf = (size_t)w2_x_half(Cin, pts) + w2_x_second(Cin, pts);
    const size_t items = (size_t)W2_NIS * CS * pts * sizeof(uint2);
    if (items + nxb * x_buf + 2 * g_stage > budget) continue;
    int ngs = (int)((budget - items - nxb * x_buf) / g_stage);
    if (ngs > W2_MAX_NGS) ngs = W2_MAX_NGS;
    if (forced.second && forced.second <= ngs) ngs = forced.second;
    c->PTS = pts; c->NGS = ngs; c->NXB = nxb;
    c->smem = (size_t)ngs * g_stage + nxb * x_buf + items;
    return true;
  }
  return false;
}

bool backward_filter2_supported(int N, long long capacity, int Cin, int Cout) {
  W2Config c;
  return w2_config(N, capacity, Cin, Cout, &c);
}

int backward_filter2_tile_rows(int N, long long capacity, int Cin, int Cout) {
  W2Config c;
  return w2_config(N, capacity, Cin, Cout, &c) ? c.PTS : 0;
}

static int w2_grid(long long tiles) {
  const int sms = sm_count();
  return (int)(tiles < sms ? (tiles < 1 ? 1 : tiles) : sms);
}

// scratch: [PTS-row work-item lists | per-CTA partials (sized for up to 256 SMs so the query needs no device)]
size_t backward_filter2_scratch_bytes(const conv3p_geom_t* g, int Cin, int Cout) {
  W2Config c;
  if (!w2_config(g->N, g->pair_capacity, Cin, Cout, &c)) return 0;
  const long long pts = (long long)g->B * g->N;
  const long long tiles = (pts + c.PTS - 1) / c.PTS;
  const long long ctas = tiles < 256 ? (tiles < 1 ? 1 : tiles) : 256;
  return group_items_bytes(pts, c.PTS) + align_up(sizeof(float) * (size_t)ctas * C3P_NCELL * Cin * Cout);
}

template <bool FROM_STORE, int GP, int PTS, bool BF16C>
static int w2_launch(const W2Args& a, int grid, size_t smem, cudaStream_t stream) {
  auto kern = k_backward_filter2<FROM_STORE, GP, PTS, BF16C>;
  const int st = ensure_dynamic_smem(kern, 227 * 1024);   // granted: 227 KB minus the kernel's static shared memory
  if (st) return st;
  {
    LaunchTimer timer_("k_backward_filter_tc", stream);
    kern<<<grid, (PTS / 4 + 2) * 32, smem, stream>>>(a);
  }
  C3P_LAUNCH_CHECK("k_backward_filter_tc");
  return CONV3P_OK;
}

template <bool FROM_STORE, bool BF16C>
static int w2_dispatch(const W2Config& c, const W2Args& a, int grid, cudaStream_t stream) {
  if (c.PTS == 64) {
    if (c.GP == 4) return w2_launch<FROM_STORE, 4, 64, BF16C>(a, grid, c.smem, stream);
    if (c.GP == 2) return w2_launch<FROM_STORE, 2, 64, BF16C>(a, grid, c.smem, stream);
    return w2_launch<FROM_STORE, 1, 64, BF16C>(a, grid, c.smem, stream);
  }
  if (c.GP == 4) return w2_launch<FROM_STORE, 4, 32, BF16C>(a, grid, c.smem, stream);
  if (c.GP == 2) return w2_launch<FROM_STORE, 2, 32, BF16C>(a, grid, c.smem, stream);
  return w2_launch<FROM_STORE, 1, 32, BF16C>(a, grid, c.smem, stream);
}

int launch_backward_filter2(const conv3p_geom_t* g, const PlanView& v, const float* grad_out, const float* input,
                            int Cin, int Cout, float* grad_filter, void* scratch, size_t scratch_bytes,
                            cudaStream_t stream, const float* g_store, bool items_ready) {
  const long long nW = (long long)C3P_NCELL * Cin * Cout;
  const long long pts = (long long)g->B * g->N;
  if (pts == 0) {
    C3P_CUDA(cudaMemsetAsync(grad_filter, 0, sizeof(float) * nW, stream));
    return CONV3P_OK;
  }
  W2Config c;
  if (!w2_config(g->N, g->pair_capacity, Cin, Cout, &c)) return CONV3P_ERR_UNSUPPORTED;
  if (!scratch || scratch_bytes < backward_filter2_scratch_bytes(g, Cin, Cout)) return CONV3P_ERR_BUFFER_TOO_SMALL;
  const GroupItems gi = carve_group_items(scratch, pts, c.PTS);
  float* partial = reinterpret_cast<float*>(static_cast<char*>(scratch) + group_items_bytes(pts, c.PTS));
  int st = items_ready ? CONV3P_OK : launch_group_items(g, v, true, c.PTS, gi, stream);   // (else built with grad_input's)
  if (st) return st;
  const int grid = w2_grid(gi.subtiles);
  if (grid > 256) return CONV3P_ERR_UNSUPPORTED;
  W2Args a{};
  a.grad_out = grad_out; a.input = input; a.rows = v.bwd_row; a.weights = v.bwd_weight;
  a.g_items = gi.items; a.g_rowid = gi.rowid; a.g_mask = gi.mask; a.partial = partial; a.g_store = g_store;
  a.total_points = pts; a.tiles = gi.subtiles; a.Cin = Cin; a.Cout = Cout;
  a.FG = c.FG; a.NGS = c.NGS; a.NXB = c.NXB;
  if (engine_flag(512))   // three TF32 products (A/B timing, accuracy reference)
    st = g_store ? w2_dispatch<true, false>(c, a, grid, stream) : w2_dispatch<false, false>(c, a, grid, stream);
  else
    st = g_store ? w2_dispatch<true, true>(c, a, grid, stream) : w2_dispatch<false, true>(c, a, grid, stream);
  if (st) return st;
  return launch_reduce_partials(partial, grid, nW, grad_filter, v.header, stream);
}

}  // namespace c3p

// profiling helper (C3P_W2_TIMED builds): read and clear the phase timers of k_backward_filter2
extern "C" int conv3p_debug_w2_cycles(unsigned long long* host16) {
  unsigned long long* host8 = host16;
  unsigned long long zero[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (cudaMemcpyFromSymbol(host8, c3p::w2_phase_cycles, sizeof(zero)) != cudaSuccess) return CONV3P_ERR_CUDA;
  if (cudaMemcpyToSymbol(c3p::w2_phase_cycles, zero, sizeof(zero)) != cudaSuccess) return CONV3P_ERR_CUDA;
  return CONV3P_OK;
}
