// backward_filter2.cu -- second-generation weight gradient on the 5th-gen tensor cores.
//
//   grad_filter[f, k, c] = sum_j input[j, k] * G_f[j, c],    G_f[j, :] = sum_{(ii,w) in cell f of j} w * grad_out[ii, :]
//
// (tf_conv3p_atrous.cpp:694-696 regrouped by (j, f'); same maths and the same MN-major 3xTF32 MMA as
// backward_filter_tc.cu: D_f[c, k] += G_f^T X, M = Cout = 128, N = Cin, K = 8 points per instruction, persistent CTA
// per SM, up to 512/Cin per-cell accumulators in TMEM per pass, deterministic partial reduce.)  What changed is the
// CUDA-core side, which bounded the first version:
//  * work items come from k_group_items (64-point sub-tiles of the backward lists, compacted by population
//    class) through a bulk-copy ring instead of per-visit prefix tables built by 64 threads with 29 dependent
//    global loads each;
//  * a quarter-warp gathers ALL 128 channels of its row (four 32-channel panels, two members per round), so
//    list ids, weights and predicates are paid once per 512-byte row rather than once per 256 bytes, and the
//    accumulation is packed FFMA2 with weight-0 padding instead of zero-selects;
//  * when grad_input is computed in the same call by k_gather_mma2, that kernel has already aggregated exactly these
//    G_f rows and left them in the G store ([sorted position][27][Cout], only non-empty (point, cell) slots are
//    written): the producers then read ONE 512-byte row per non-empty item (prefetched one group ahead) instead of
//    walking its list -- 4.2x fewer row reads on the bench cloud and no ids or weights (template FROM_STORE).
//
// Warp roles: warps [0, 16) producers (also the flush), 16 = MMA issuer + TMEM allocator, 17 = item-list loader.
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_gather2.cuh"

namespace c3p {

using namespace tc;

constexpr int W2_NPW = 16;
constexpr int W2_THREADS = (W2_NPW + 2) * 32;
constexpr int W2_PTS = 64;                          // points per stage (contraction length of a stage)
constexpr int W2_PANEL = W2_PTS * PANEL_ROW_BYTES;  // 8 KB: 64 rows x 32 fp32
constexpr int W2_NIS = 4;                           // item-list slots
constexpr int W2_GP = 4;                            // Cout == 128: four 32-channel panels of G
constexpr int W2_END = -1;

struct W2Args {
  const float* grad_out;   // [B*N, Cout]
  const float* input;      // [B*N, Cin]
  const int* rows;         // backward lists: rows ii
  const float* weights;    // backward lists: 1 / count(ii, f')
  const uint2* g_items;    // [tiles][27][64]
  const int* g_rowid;      // [tiles*64]
  const unsigned* g_mask;  // [tiles] bit f: some point of the tile has members in cell f
  float* partial;          // [gridDim.x][27*Cin*Cout]
  const float* g_store;    // FROM_STORE: G_f rows [tiles*64][27][Cout] written by k_gather_mma2
  long long total_points, tiles;
  int Cin, Cout, FG;       // FG = accumulators (cells) per pass
};

template <bool FROM_STORE>
__global__ void __launch_bounds__(W2_THREADS, 1) k_backward_filter2(const W2Args a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int Cin = a.Cin, Cout = a.Cout, FG = a.FG;
  const int xp = Cin / 32;                              // panels of the input rows
  constexpr uint32_t g_half = (uint32_t)W2_GP * W2_PANEL;   // hi (or lo) part of a G stage
  const uint32_t x_half = (uint32_t)xp * W2_PANEL;
  unsigned char* g_base = smem;                          // 2 stages x (hi, lo)
  unsigned char* x_base = g_base + 4 * (size_t)g_half;   // 2 buffers x (hi, lo)
  uint2* items = reinterpret_cast<uint2*>(x_base + 4 * (size_t)x_half);  // [NIS][64]
  __shared__ uint64_t g_full[2], g_empty[2], x_full[2], x_empty[2], it_full[W2_NIS], it_empty[W2_NIS], acc_full,
      acc_empty;
  __shared__ uint32_t tmem_slot;
  __shared__ int hdr[W2_NIS];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long per_cta = (a.tiles + gridDim.x - 1) / gridDim.x;
  const long long tile_lo = (long long)blockIdx.x * per_cta;
  const long long tile_hi = min(a.tiles, tile_lo + per_cta);
  const int npass = (C3P_NCELL + FG - 1) / FG;
  const unsigned max_row = (unsigned)(a.total_points - 1);

  if (warp == W2_NPW) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&g_full[i], W2_NPW);
        mbar_init(&g_empty[i], 1);
        mbar_init(&x_full[i], W2_NPW);
        mbar_init(&x_empty[i], 1);
      }
      for (int i = 0; i < W2_NIS; ++i) {
        mbar_init(&it_full[i], 1);
        mbar_init(&it_empty[i], W2_NPW);
      }
      mbar_init(&acc_full, 1);
      mbar_init(&acc_empty, W2_NPW);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(&tmem_slot, 512);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp < W2_NPW) {
    // =========================== producers ===========================================================
    // Quarter-warp q serves row q of the tile for the input panels and item (q + 4g) mod 64 of group g.
    const int q = warp * 4 + (lane >> 3), l8 = lane & 7;
    auto read_item = [&](int g_, G2Item& it) -> int {
      const int slot = g_ & (W2_NIS - 1);
      mbar_wait(&it_full[slot], (uint32_t)((g_ / W2_NIS) & 1));
      const int h = hdr[slot];
      const uint2 u = items[slot * W2_PTS + ((q + 4 * g_) & 63)];
      __syncwarp();
      if (lane == 0) mbar_arrive(&it_empty[slot]);
      it.pos = u.x; it.p = (int)(u.y & 255u); it.n = h != W2_END ? (int)(u.y >> 8) : 0;
      it.inv = 0.f; it.w = 0.f;
      if (!FROM_STORE) g2_prefetch<true>(it, a.rows, a.weights, 0, l8, max_row);
      return h;
    };
    // FROM_STORE: the item's aggregated row (hdr = tile-local index << 8 | cell)
    auto fetch_row = [&](const G2Item& it, int h, float4 (&v)[W2_GP]) {
      const float* src = a.g_store + (((size_t)(tile_lo + (h >> 8)) * W2_PTS + it.p) * C3P_NCELL + (h & 255)) * Cout +
                         l8 * 4;
#pragma unroll
      for (int kc = 0; kc < W2_GP; ++kc)
        v[kc] = it.n > 0 ? ldg4(src + kc * PANEL_K) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    int visit = 0, g = 0;
    // Look-ahead queue of work items: the next item, with its first list ids (gather mode) or its G row (store mode)
    // in flight.  (Three rows of look-ahead were measured slower: the kernel is bound by shared-memory bandwidth --
    // operand reads of the MMAs plus the panel stores -- not by the latency of the row loads, and the extra registers
    // spill.)
    constexpr int LOOK = 1;
    G2Item ahead[LOOK];
    float4 ahead_row[LOOK][W2_GP];
    int n_read = 0;
    bool ended = false;
    auto pull = [&](G2Item& it, float4 (&r)[W2_GP]) {
      if (ended) {                 // nothing is published after the END marker
        it.n = 0; it.p = 0; it.pos = 0u;
        return;
      }
      const int h = read_item(n_read++, it);
      ended = h == W2_END;
      if (FROM_STORE) fetch_row(it, ended ? 0 : h, r);
    };
#pragma unroll
    for (int d = 0; d < LOOK; ++d) pull(ahead[d], ahead_row[d]);
    for (int pass = 0; pass < npass; ++pass) {
      const int f0 = pass * FG, f1 = min(C3P_NCELL, f0 + FG);
      const unsigned pass_bits = ((f1 - f0) == 32 ? ~0u : ((1u << (f1 - f0)) - 1u)) << f0;
      unsigned pass_mask = 0;
      int xrow = tile_lo < tile_hi ? __ldg(a.g_rowid + tile_lo * W2_PTS + q) : -1;
      unsigned mask_next = tile_lo < tile_hi ? __ldg(a.g_mask + tile_lo) & pass_bits : 0u;
      for (long long tile = tile_lo; tile < tile_hi; ++tile) {
        const unsigned mask = mask_next;
        const int row = xrow;
        if (tile + 1 < tile_hi) {  // prefetch the next visit's row id and mask
          xrow = __ldg(a.g_rowid + (tile + 1) * W2_PTS + q);
          mask_next = __ldg(a.g_mask + tile + 1) & pass_bits;
        }
        if (!mask) continue;
        pass_mask |= mask;
        // ---- input rows of the tile -> X panels (hi/lo) ---------------------------------------------------
        {
          const int xb = visit & 1, use = visit >> 1;
          if (use >= 1) mbar_wait(&x_empty[xb], (uint32_t)((use - 1) & 1));
          unsigned char* xs = x_base + (size_t)xb * 2 * x_half;
          const float* xr = a.input + (size_t)(row >= 0 ? row : 0) * Cin + l8 * 4;
          for (int pnl = 0; pnl < xp; ++pnl) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row >= 0) v = ldg4(xr + pnl * PANEL_K);
            g2_store_split(xs + (size_t)pnl * W2_PANEL + panel_chunk_offset_mn(q, l8), x_half, v);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&x_full[xb]);
        }
        // ---- one G stage per active cell -----------------------------------------------------------------
        for (unsigned todo = mask; todo; todo &= todo - 1) {
          G2Item cur = ahead[0];
          float4 acc[W2_GP];
#pragma unroll
          for (int kc = 0; kc < W2_GP; ++kc) acc[kc] = ahead_row[0][kc];
#pragma unroll
          for (int d = 0; d + 1 < LOOK; ++d) {
            ahead[d] = ahead[d + 1];
#pragma unroll
            for (int kc = 0; kc < W2_GP; ++kc) ahead_row[d][kc] = ahead_row[d + 1][kc];
          }
          pull(ahead[LOOK - 1], ahead_row[LOOK - 1]);
          if (!FROM_STORE) {
            int nmax = max(cur.n, __shfl_xor_sync(C3P_FULL_MASK, cur.n, 8));
            nmax = max(nmax, __shfl_xor_sync(C3P_FULL_MASK, nmax, 16));
            g2_gather<W2_GP, 2, true>(acc, cur, nmax, a.grad_out, Cout, 0, a.rows, a.weights, l8, max_row);
          }
          const int slot = g & 1, use = g >> 1;
          if (use >= 1) mbar_wait(&g_empty[slot], (uint32_t)((use - 1) & 1));
          unsigned char* stage = g_base + (size_t)slot * 2 * g_half + panel_chunk_offset_mn(cur.p, l8);
#pragma unroll
          for (int kc = 0; kc < W2_GP; ++kc) g2_store_split(stage + (size_t)kc * W2_PANEL, g_half, acc[kc]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&g_full[slot]);
          ++g;
        }
        ++visit;
      }
      // ---- flush this pass's accumulators: partial[cta][f][k][c] = D_f[c][k] ------------------------------
      mbar_wait(&acc_full, (uint32_t)(pass & 1));
      tc_fence_after_sync();
      {
        const int sub = warp & 3;
        for (int ai = warp >> 2; ai < f1 - f0; ai += 4) {
          const int f = f0 + ai;
          const int c = sub * 32 + lane;                // TMEM lane == output channel c
          const bool live = (pass_mask >> f) & 1u;
          float* dst = a.partial + ((size_t)blockIdx.x * C3P_NCELL + f) * Cin * Cout;
          for (int k0 = 0; k0 < Cin; k0 += 32) {
            float v[32];
            if (live) {
              tmem_ld_32x32(tmem + ((uint32_t)(sub * 32) << 16) + (uint32_t)(ai * Cin + k0), v);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
            if (c < Cout) {
#pragma unroll
              for (int j = 0; j < 32; ++j) dst[(size_t)(k0 + j) * Cout + c] = v[j];
            }
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty);
    }
  } else if (warp == W2_NPW) {
    // =========================== MMA issuer (one thread) ===============================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32_mn(Cout, Cin);
      int visit = 0, gs = 0;
      for (int pass = 0; pass < npass; ++pass) {
        const int f0 = pass * FG, f1 = min(C3P_NCELL, f0 + FG);
        const unsigned pass_bits = ((f1 - f0) == 32 ? ~0u : ((1u << (f1 - f0)) - 1u)) << f0;
        unsigned started = 0;
        if (pass > 0) {
          mbar_wait(&acc_empty, (uint32_t)((pass - 1) & 1));
          tc_fence_after_sync();
        }
        for (long long tile = tile_lo; tile < tile_hi; ++tile) {
          const unsigned mask = __ldg(a.g_mask + tile) & pass_bits;
          if (!mask) continue;
          const int xb = visit & 1;
          mbar_wait(&x_full[xb], (uint32_t)((visit >> 1) & 1));
          const uint32_t x_hi = smem_u32(x_base + (size_t)xb * 2 * x_half), x_lo = x_hi + x_half;
          const int f_last = 31 - __clz(mask);
          for (int f = f0; f < f1; ++f) {
            if (!((mask >> f) & 1u)) continue;
            const int slot = gs & 1;
            mbar_wait(&g_full[slot], (uint32_t)((gs >> 1) & 1));
            tc_fence_after_sync();
            const uint32_t g_hi = smem_u32(g_base + (size_t)slot * 2 * g_half), g_lo = g_hi + g_half;
            const uint32_t d = tmem + (uint32_t)((f - f0) * Cin);
#pragma unroll
            for (int j = 0; j < W2_PTS / 8; ++j) {
              const uint32_t adv = (uint32_t)j * 1024u;  // 8 points further down the panels
              const uint64_t dgh = make_smem_desc_mn(g_hi + adv, W2_PANEL), dgl = make_smem_desc_mn(g_lo + adv, W2_PANEL);
              const uint64_t dxh = make_smem_desc_mn(x_hi + adv, W2_PANEL), dxl = make_smem_desc_mn(x_lo + adv, W2_PANEL);
              mma_tf32(d, dgh, dxh, idesc, (((started >> f) & 1u) | (unsigned)j) ? 1u : 0u);
              mma_tf32(d, dgl, dxh, idesc, 1u);
              mma_tf32(d, dgh, dxl, idesc, 1u);
            }
            started |= 1u << f;
            mma_commit(&g_empty[slot]);
            if (f == f_last) mma_commit(&x_empty[xb]);
            ++gs;
          }
          ++visit;
        }
        mma_commit(&acc_full);
      }
    }
  } else {
    // =========================== item-list loader (one thread) ==========================================
    if (lane == 0) {
      int g = 0;
      for (int pass = 0; pass < npass; ++pass) {
        const int f0 = pass * FG, f1 = min(C3P_NCELL, f0 + FG);
        const unsigned pass_bits = ((f1 - f0) == 32 ? ~0u : ((1u << (f1 - f0)) - 1u)) << f0;
        for (long long tile = tile_lo; tile < tile_hi; ++tile) {
          for (unsigned todo = __ldg(a.g_mask + tile) & pass_bits; todo; todo &= todo - 1) {
            const int f = __ffs(todo) - 1;
            const int slot = g & (W2_NIS - 1), use = g / W2_NIS;
            if (use >= 1) mbar_wait(&it_empty[slot], (uint32_t)((use - 1) & 1));
            hdr[slot] = f | ((int)(tile - tile_lo) << 8);
            mbar_arrive_expect_tx(&it_full[slot], W2_PTS * sizeof(uint2));
            bulk_copy_g2s(items + slot * W2_PTS, a.g_items + (tile * C3P_NCELL + f) * W2_PTS,
                          W2_PTS * sizeof(uint2), &it_full[slot]);
            ++g;
          }
        }
      }
      const int slot = g & (W2_NIS - 1), use = g / W2_NIS;
      if (use >= 1) mbar_wait(&it_empty[slot], (uint32_t)((use - 1) & 1));
      hdr[slot] = W2_END;
      mbar_arrive(&it_full[slot]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == W2_NPW) tmem_dealloc(tmem, 512);
}

__global__ void k_reduce_partials2(const float* __restrict__ partial, int S, long long nW,
                                   float* __restrict__ out) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nW) return;
  float s = 0.f;
  for (int i = 0; i < S; ++i) s += partial[(size_t)i * nW + w];  // fixed order: deterministic
  out[w] = s;
}

static int w2_grid(long long tiles) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  (void)cudaGetLastError();
  if (sms < 1) sms = 148;
  return (int)(tiles < sms ? (tiles < 1 ? 1 : tiles) : sms);
}

static size_t w2_smem_bytes(int Cin) {
  return 4 * (size_t)W2_GP * W2_PANEL + 4 * (size_t)(Cin / 32) * W2_PANEL + W2_NIS * W2_PTS * sizeof(uint2);
}

bool backward_filter2_supported(int N, long long capacity, int Cin, int Cout) {
  if (N > 65535 || capacity >= (1LL << 32)) return false;
  if (Cout != 128) return false;                         // M of the MMA (TMEM lane == channel)
  if (Cin % 32 || Cin < 32 || Cin > 256) return false;   // N of the MMA, in 32-wide MN-major panels
  return w2_smem_bytes(Cin) <= 227 * 1024 - 1024;
}

// scratch: [64-row work-item lists | per-CTA partials (sized for up to 256 SMs so the query needs no device)]
size_t backward_filter2_scratch_bytes(const conv3p_geom_t* g, int Cin, int Cout) {
  const long long pts = (long long)g->B * g->N;
  const long long tiles = (pts + W2_PTS - 1) / W2_PTS;
  const long long ctas = tiles < 256 ? (tiles < 1 ? 1 : tiles) : 256;
  return group_items_bytes(pts, W2_PTS) + align_up(sizeof(float) * (size_t)ctas * C3P_NCELL * Cin * Cout);
}

int launch_backward_filter2(const conv3p_geom_t* g, const PlanView& v, const float* grad_out, const float* input,
                            int Cin, int Cout, float* grad_filter, void* scratch, size_t scratch_bytes,
                            cudaStream_t stream, const float* g_store) {
  const long long nW = (long long)C3P_NCELL * Cin * Cout;
  const long long pts = (long long)g->B * g->N;
  if (pts == 0) {
    C3P_CUDA(cudaMemsetAsync(grad_filter, 0, sizeof(float) * nW, stream));
    return CONV3P_OK;
  }
  if (!scratch || scratch_bytes < backward_filter2_scratch_bytes(g, Cin, Cout)) return CONV3P_ERR_BUFFER_TOO_SMALL;
  const GroupItems gi = carve_group_items(scratch, pts, W2_PTS);
  float* partial = reinterpret_cast<float*>(static_cast<char*>(scratch) + group_items_bytes(pts, W2_PTS));
  int st = launch_group_items(g, v, true, W2_PTS, gi, stream);
  if (st) return st;
  const int grid = w2_grid(gi.subtiles);
  if (grid > 256) return CONV3P_ERR_UNSUPPORTED;
  W2Args a{};
  a.grad_out = grad_out; a.input = input; a.rows = v.bwd_row; a.weights = v.bwd_weight;
  a.g_items = gi.items; a.g_rowid = gi.rowid; a.g_mask = gi.mask; a.partial = partial; a.g_store = g_store;
  a.total_points = pts; a.tiles = gi.subtiles; a.Cin = Cin; a.Cout = Cout;
  a.FG = 512 / Cin > 8 ? 8 : 512 / Cin;
  const size_t smem = w2_smem_bytes(Cin);
  if (g_store) {
    C3P_CUDA(cudaFuncSetAttribute(k_backward_filter2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LaunchTimer timer_("k_backward_filter_tc", stream);
    k_backward_filter2<true><<<grid, W2_THREADS, smem, stream>>>(a);
  } else {
    C3P_CUDA(cudaFuncSetAttribute(k_backward_filter2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LaunchTimer timer_("k_backward_filter_tc", stream);
    k_backward_filter2<false><<<grid, W2_THREADS, smem, stream>>>(a);
  }
  C3P_LAUNCH_CHECK("k_backward_filter_tc");
  {
    LaunchTimer timer_("k_reduce_partials", stream);
    k_reduce_partials2<<<(unsigned)((nW + 255) / 256), 256, 0, stream>>>(partial, grid, nW, grad_filter);
  }
  C3P_LAUNCH_CHECK("k_reduce_partials");
  return CONV3P_OK;
}

}  // namespace c3p
