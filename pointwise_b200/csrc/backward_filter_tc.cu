// backward_filter_tc.cu -- the weight gradient on the 5th-gen tensor cores.
//
//   grad_filter[f, k, c] = sum_j input[j, k] * G_f[j, c],    G_f[j, :] = sum_{(ii,w) in cell f of j} w * grad_out[ii, :]
//
// (tf_conv3p_atrous.cpp:694-696 regrouped by (j, f'), see backward_filter.cu).  The contraction index is
// the POINT, so both MMA operands are "MN-major": a stage holds 64 points; G_f is gathered by the
// producer warps exactly like the forward aggregate (quarter-warp per 128-byte row segment, weighted
// sum in registers, TF32 hi/lo split) into point-major panels, the matching input rows are split into
// panels once per tile visit, and one thread issues  D_f[c, k] += G_f^T * X  as 3xTF32 tcgen05.mma
// (M = Cout, N = Cin, K = 8 points per instruction).  A persistent CTA owns a contiguous range of
// 64-point tiles and keeps up to 512/Cin per-cell accumulators in TMEM, so the 27 cells are covered in
// ceil(27*Cin/512) passes over its tiles; each pass ends with one TMEM -> global flush of the CTA's
// partial sums, and k_reduce_partials adds the partials of all CTAs in a fixed order (deterministic;
// the reference uses per-thread copies on the CPU and global atomicAdd on the GPU,
// tf_conv3p_atrous.cpp:611-621, tf_conv3p_atrous.cu:494).
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_gather.cuh"

namespace c3p {

using namespace tc;

constexpr int WG_NPW = 16;                    // producer warps (also the epilogue): 64 quarter-warps
constexpr int WG_NQ = WG_NPW * 4;
constexpr int WG_THREADS = (WG_NPW + 1) * 32; // + MMA issuer / TMEM allocator warp
constexpr int WG_PTS = 64;                    // points per stage (contraction length of a stage)
constexpr int WG_PANEL = WG_PTS * PANEL_ROW_BYTES;  // 8 KB: 64 rows x 32 fp32

struct WGArgs {
  const float* grad_out;   // [B*N, Cout]
  const float* input;      // [B*N, Cin]
  const int* cnt;          // bwd_count [B*N, 27]
  const long long* begin;
  const int* len;
  const int* rows;
  const float* weights;
  const float4* sorted_xyzi;
  const unsigned* tile_mask;  // [tiles] bit f: some point of the 64-point tile has members in cell f
  float* partial;             // [gridDim.x][27*Cin*Cout]
  long long total_points, capacity, tiles;
  int N, Cin, Cout, FG;       // FG = accumulators (cells) per pass
};

__global__ void k_tile_masks(const int* __restrict__ cnt, const long long* __restrict__ begin,
                             const int* __restrict__ len, const float4* __restrict__ sorted_xyzi,
                             long long total_points, long long capacity, int N, long long tiles,
                             unsigned* __restrict__ tile_mask) {
  const long long tile = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tile >= tiles) return;
  const int lane = threadIdx.x & 31;
  unsigned m = 0;
  for (int r = lane; r < WG_PTS; r += 32) {
    const long long s = tile * WG_PTS + r;
    if (s >= total_points) continue;
    const int b = (int)(s / N);
    const int row = b * N + __float_as_int(sorted_xyzi[s].w);
    if (begin[row] + len[row] > capacity) continue;
    for (int f = 0; f < C3P_NCELL; ++f)
      if (__ldg(cnt + (size_t)row * C3P_NCELL + f) > 0) m |= 1u << f;
  }
  for (int o = 16; o > 0; o >>= 1) m |= __shfl_xor_sync(C3P_FULL_MASK, m, o);
  if (lane == 0) tile_mask[tile] = m;
}

__global__ void __launch_bounds__(WG_THREADS, 1) k_backward_filter_tc(const WGArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int Cin = a.Cin, Cout = a.Cout, FG = a.FG;
  constexpr int gp = 4;                                 // Cout == 128: four 32-channel panels of G
  const int xp = Cin / 32;                              // panels of the input rows
  const uint32_t g_half = (uint32_t)gp * WG_PANEL;      // hi (or lo) part of a G stage
  const uint32_t x_half = (uint32_t)xp * WG_PANEL;
  unsigned char* g_base = smem;                          // 2 stages x (hi, lo)
  unsigned char* x_base = g_base + 4 * (size_t)g_half;   // 2 buffers x (hi, lo)
  unsigned char* tab = x_base + 4 * (size_t)x_half;
  uint16_t* pre16 = reinterpret_cast<uint16_t*>(tab);                 // [2][64][28]
  uint32_t* beg = reinterpret_cast<uint32_t*>(pre16 + 2 * WG_PTS * 28);  // [2][64]
  int* rowid = reinterpret_cast<int*>(beg + 2 * WG_PTS);              // [2][64]
  __shared__ uint64_t g_full[2], g_empty[2], x_full[2], x_empty[2], acc_full, acc_empty;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long per_cta = (a.tiles + gridDim.x - 1) / gridDim.x;
  const long long tile_lo = (long long)blockIdx.x * per_cta;
  const long long tile_hi = min(a.tiles, tile_lo + per_cta);
  const int npass = (C3P_NCELL + FG - 1) / FG;

  if (warp == WG_NPW) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&g_full[i], WG_NPW);
        mbar_init(&g_empty[i], 1);
        mbar_init(&x_full[i], WG_NPW);
        mbar_init(&x_empty[i], 1);
      }
      mbar_init(&acc_full, 1);
      mbar_init(&acc_empty, WG_NPW);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(&tmem_slot, 512);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp < WG_NPW) {
    // =========================== producers ===========================================================
    const int q = warp * 4 + (lane >> 3), l8 = lane & 7;
    int visit = 0, gs = 0;
    for (int pass = 0; pass < npass; ++pass) {
      const int f0 = pass * FG, f1 = min(C3P_NCELL, f0 + FG);
      const unsigned pass_bits = ((f1 - f0) == 32 ? ~0u : ((1u << (f1 - f0)) - 1u)) << f0;
      unsigned pass_mask = 0;
      for (long long tile = tile_lo; tile < tile_hi; ++tile) {
        const unsigned mask = __ldg(a.tile_mask + tile) & pass_bits;
        if (!mask) continue;
        pass_mask |= mask;
        const int tb = visit & 1;
        // ---- per-tile tables (safe to overwrite: every producer finished visit-2, see named barrier) ----
        if (tid < WG_PTS) {
          const long long s = tile * WG_PTS + tid;
          int row = -1;
          long long bg = 0;
          bool ok = false;
          if (s < a.total_points) {
            const int b = (int)(s / a.N);
            row = b * a.N + __float_as_int(a.sorted_xyzi[s].w);
            bg = a.begin[row];
            ok = bg + a.len[row] <= a.capacity;
          }
          int run = 0;
          uint16_t* pr = pre16 + (tb * WG_PTS + tid) * 28;
          for (int f = 0; f < C3P_NCELL; ++f) {
            pr[f] = (uint16_t)run;
            run += ok ? __ldg(a.cnt + (size_t)row * C3P_NCELL + f) : 0;
          }
          pr[27] = (uint16_t)run;
          beg[tb * WG_PTS + tid] = (uint32_t)bg;
          rowid[tb * WG_PTS + tid] = row;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(WG_NPW * 32) : "memory");
        // ---- input rows of the tile -> X panels (hi/lo) ---------------------------------------------------
        {
          const int xb = visit & 1, use = visit >> 1;
          if (use >= 1) mbar_wait(&x_empty[xb], (uint32_t)((use - 1) & 1));
          unsigned char* xs = x_base + (size_t)xb * 2 * x_half;
          for (int e = q; e < WG_PTS * xp; e += WG_NQ) {
            const int r = e / xp, pnl = e - r * xp;
            const int row = rowid[tb * WG_PTS + r];
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row >= 0) v = ldg_f4(a.input + (size_t)row * Cin + pnl * PANEL_K + l8 * 4);
            store_split(xs + (size_t)pnl * WG_PANEL + panel_chunk_offset_mn(r, l8), x_half, v);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&x_full[xb]);
        }
        // ---- one G stage per active cell -----------------------------------------------------------------
        // Work item = (point, pair of 32-channel panels); quarter-warp q serves items (q + rot) mod NQ (+ NQ).
        // Software-pipelined over (cell, repetition) slots of this visit: the list ids of the next slot are
        // fetched while the rows of the current one are in flight.
        const int items = WG_PTS * (gp / 2);
        auto fetch = [&](int f_, int gs_, int rep, int& e_out) -> GatherSlot {
          const int e = rep * WG_NQ + (q + gs_ * 32) % WG_NQ;
          const bool valid = e < items;
          e_out = valid ? e : -1;
          const int r = valid ? e / (gp / 2) : 0;
          const int off = pre16[(tb * WG_PTS + r) * 28 + f_];
          const int n = valid ? (int)pre16[(tb * WG_PTS + r) * 28 + f_ + 1] - off : 0;
          return fetch_slot<true>(n, beg[tb * WG_PTS + r] + (uint32_t)off, a.rows, a.weights, l8);
        };
        const int nrep = (items + WG_NQ - 1) / WG_NQ;
        unsigned todo = mask;
        int f = __ffs(todo) - 1;
        int e_cur = -1;
        GatherSlot d = fetch(f, gs, 0, e_cur);
        while (todo) {
          todo &= todo - 1;
          const int f_next = todo ? __ffs(todo) - 1 : -1;
          const int slot = gs & 1, use = gs >> 1;
          unsigned char* stage = g_base + (size_t)slot * 2 * g_half;
          for (int rep = 0; rep < nrep; ++rep) {
            int e_next = -1;
            GatherSlot dn;
            dn.n = 0; dn.lb = 0; dn.ids = 0; dn.w = 0.f;
            if (rep + 1 < nrep) dn = fetch(f, gs, rep + 1, e_next);
            else if (f_next >= 0) dn = fetch(f_next, gs + 1, 0, e_next);
            const int kb = e_cur >= 0 ? e_cur % (gp / 2) : 0, r = e_cur >= 0 ? e_cur / (gp / 2) : 0;
            float4 acc[2];
            gather_slot<2, true>(acc, d, a.grad_out, Cout, kb * 2 * PANEL_K, a.rows, a.weights, l8);
            if (rep == 0 && use >= 1) mbar_wait(&g_empty[slot], (uint32_t)((use - 1) & 1));
            if (e_cur >= 0) {
#pragma unroll
              for (int kc = 0; kc < 2; ++kc)
                store_split(stage + (size_t)(kb * 2 + kc) * WG_PANEL + panel_chunk_offset_mn(r, l8), g_half,
                            acc[kc]);
            }
            d = dn;
            e_cur = e_next;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&g_full[slot]);
          ++gs;
          f = f_next;
        }
        ++visit;
      }
      // ---- flush this pass's accumulators: partial[cta][f][k][c] = D_f[c][k] ------------------------------
      mbar_wait(&acc_full, (uint32_t)(pass & 1));
      tc_fence_after_sync();
      if (warp < 16) {
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
  } else {
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
          const unsigned mask = __ldg(a.tile_mask + tile) & pass_bits;
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
            for (int j = 0; j < WG_PTS / 8; ++j) {
              const uint32_t adv = (uint32_t)j * 1024u;  // 8 points further down the panels
              const uint64_t dgh = make_smem_desc_mn(g_hi + adv, WG_PANEL), dgl = make_smem_desc_mn(g_lo + adv, WG_PANEL);
              const uint64_t dxh = make_smem_desc_mn(x_hi + adv, WG_PANEL), dxl = make_smem_desc_mn(x_lo + adv, WG_PANEL);
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
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == WG_NPW) tmem_dealloc(tmem, 512);
}

__global__ void k_reduce_partials_tc(const float* __restrict__ partial, int S, long long nW,
                                     float* __restrict__ out) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nW) return;
  float s = 0.f;
  for (int i = 0; i < S; ++i) s += partial[(size_t)i * nW + w];  // fixed order: deterministic
  out[w] = s;
}

static int wg_grid(long long tiles) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  (void)cudaGetLastError();
  if (sms < 1) sms = 148;
  return (int)(tiles < sms ? (tiles < 1 ? 1 : tiles) : sms);
}

static size_t wg_smem_bytes(int Cin, int Cout) {
  return 4 * (size_t)(Cout / 32) * WG_PANEL + 4 * (size_t)(Cin / 32) * WG_PANEL +
         2 * WG_PTS * (56 + 4 + 4);
}

bool backward_filter_tc_supported(int N, long long capacity, int Cin, int Cout) {
  if (N > 65535 || capacity >= (1LL << 32)) return false;
  if (Cout != 128) return false;                         // M of the MMA (TMEM lane == channel)
  if (Cin % 32 || Cin < 32 || Cin > 256) return false;   // N of the MMA, in 32-wide MN-major panels
  return wg_smem_bytes(Cin, Cout) <= 227 * 1024 - 1024;
}

size_t backward_filter_tc_scratch_bytes(const conv3p_geom_t* g, int Cin, int Cout) {
  const long long pts = (long long)g->B * g->N;
  const long long tiles = (pts + WG_PTS - 1) / WG_PTS;
  // mask array + per-CTA partials (sized for up to 256 SMs so the query needs no device)
  const long long ctas = tiles < 256 ? (tiles < 1 ? 1 : tiles) : 256;
  return align_up(sizeof(unsigned) * (size_t)(tiles + 1)) +
         align_up(sizeof(float) * (size_t)ctas * C3P_NCELL * Cin * Cout);
}

int launch_backward_filter_tc(const conv3p_geom_t* g, const PlanView& v, const float* grad_out,
                              const float* input, int Cin, int Cout, float* grad_filter, void* scratch,
                              size_t scratch_bytes, cudaStream_t stream) {
  const long long nW = (long long)C3P_NCELL * Cin * Cout;
  const long long pts = (long long)g->B * g->N;
  if (pts == 0) {
    C3P_CUDA(cudaMemsetAsync(grad_filter, 0, sizeof(float) * nW, stream));
    return CONV3P_OK;
  }
  if (!scratch || scratch_bytes < backward_filter_tc_scratch_bytes(g, Cin, Cout))
    return CONV3P_ERR_BUFFER_TOO_SMALL;
  const long long tiles = (pts + WG_PTS - 1) / WG_PTS;
  const int grid = wg_grid(tiles);
  if (grid > 256) return CONV3P_ERR_UNSUPPORTED;
  unsigned* masks = static_cast<unsigned*>(scratch);
  float* partial = reinterpret_cast<float*>(static_cast<char*>(scratch) +
                                            align_up(sizeof(unsigned) * (size_t)(tiles + 1)));
  {
    LaunchTimer timer_("k_tile_masks", stream);
    k_tile_masks<<<(unsigned)((tiles + 7) / 8), 256, 0, stream>>>(v.bwd_count, v.pair_begin, v.pair_len,
                                                                  v.sorted_xyzi, pts, g->pair_capacity, g->N,
                                                                  tiles, masks);
  }
  C3P_LAUNCH_CHECK("k_tile_masks");
  WGArgs a{};
  a.grad_out = grad_out; a.input = input; a.cnt = v.bwd_count; a.begin = v.pair_begin; a.len = v.pair_len;
  a.rows = v.bwd_row; a.weights = v.bwd_weight; a.sorted_xyzi = v.sorted_xyzi; a.tile_mask = masks;
  a.partial = partial; a.total_points = pts; a.capacity = g->pair_capacity; a.tiles = tiles;
  a.N = g->N; a.Cin = Cin; a.Cout = Cout;
  a.FG = 512 / Cin > 8 ? 8 : 512 / Cin;
  const size_t smem = wg_smem_bytes(Cin, Cout);
  C3P_CUDA(cudaFuncSetAttribute(k_backward_filter_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {
    LaunchTimer timer_("k_backward_filter_tc", stream);
    k_backward_filter_tc<<<grid, WG_THREADS, smem, stream>>>(a);
  }
  C3P_LAUNCH_CHECK("k_backward_filter_tc");
  {
    LaunchTimer timer_("k_reduce_partials", stream);
    k_reduce_partials_tc<<<(unsigned)((nW + 255) / 256), 256, 0, stream>>>(partial, grid, nW, grad_filter);
  }
  C3P_LAUNCH_CHECK("k_reduce_partials");
  return CONV3P_OK;
}

}  // namespace c3p
