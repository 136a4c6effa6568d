// neighbors.cu -- stage 2 of the neighbour plan: per-point windowed search over the voxel-sorted
// cloud with the reference's EXACT predicate, emitting the count table [B*N,27] and cell-grouped
// neighbour lists; and stage 3, the backward lists.
//
// One warp per query point (in voxel-sorted order, so neighbouring warps touch the same candidate
// rows).  Candidate generation is free to over-cover: per axis, the window is the set of grid cells
// overlapped by the three dilated taps of the box (tf_conv3p_atrous.cpp:235-245, :280-288) widened
// by a rounding slop; along x each window range is one contiguous key range of the sorted cloud,
// located by binary search.  Every candidate then goes through the reference predicate -- closed
// box test (:277), fp32 subtract / IEEE divide / truncate / clamp (:280-282), hole test (:285) --
// so the accepted (j, f) sets equal Grid::neighbor's (:232-301) bit for bit.
#include "common.cuh"

namespace c3p {

constexpr int NB_WARPS = 8;
constexpr int NB_THREADS = NB_WARPS * 32;
constexpr int STASH_CAP = 512;   // pairs per point kept in shared memory; longer lists re-sweep
constexpr uint32_t J_MASK = (1u << 27) - 1;

struct AxisWin {
  int lo[3], hi[3];  // inclusive, ascending, disjoint cell ranges
  int n;             // number of ranges
  int cells;         // total cells
};

// Grid cells that can hold a point whose tap along this axis is 0, 1 or 2.
__device__ __forceinline__ void axis_window(float blo, float bhi, int stride, float voxel, float vmin,
                                            int dim, AxisWin& w) {
  const float mag = fmaxf(fmaxf(fabsf(blo), fabsf(bhi)), fabsf(vmin));
  const float slop = voxel * (1.0f / 1024.0f) + mag * 1e-6f;  // >> any fp32 rounding in the predicate
  w.n = 0;
  w.cells = 0;
  int prev_hi = -1;
  const int ntap = (stride == 1) ? 1 : 3;  // stride 1: the taps tile the box, one range
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    if (t < ntap) {
      float a = (stride == 1) ? blo - slop : blo + (float)(t * stride) * voxel - slop;
      float b = (stride == 1 || t == 2) ? bhi + slop : blo + (float)(t * stride + 1) * voxel + slop;
      int c0 = grid_coord(a, vmin, voxel, dim);
      int c1 = grid_coord(b, vmin, voxel, dim);
      c0 = max(c0, prev_hi + 1);
      if (c0 <= c1) {
        if (w.n > 0 && c0 == prev_hi + 1) {
          w.hi[w.n - 1] = c1;
        } else {
          w.lo[w.n] = c0;
          w.hi[w.n] = c1;
          w.n++;
        }
        w.cells += c1 - c0 + 1;
        prev_hi = c1;
      }
    }
  }
}

__device__ __forceinline__ int window_cell(const AxisWin& w, int i) {
  int l0 = w.hi[0] - w.lo[0] + 1;
  if (i < l0 || w.n == 1) return w.lo[0] + i;
  i -= l0;
  int l1 = w.hi[1] - w.lo[1] + 1;
  if (i < l1 || w.n == 2) return w.lo[1] + i;
  return w.lo[2] + (i - l1);
}

__device__ __forceinline__ int lower_bound_key(const uint32_t* __restrict__ keys, int n,
                                               uint32_t target) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(keys + mid) < target) lo = mid + 1; else hi = mid;
  }
  return lo;
}

struct Query {
  float lo[3], hi[3];
  int full[3], stride[3];
  float voxel, inv_voxel;
  AxisWin wx, wy, wz;
  int dimx, dimy;
};

// Visits every candidate of the query's window; calls fn(f, j) warp-collectively with f = kernel
// cell 0..26 of an accepted neighbour (or -1) and j its index inside the cloud.
// `bins` = the cloud's bin-offset table (k_cloud_sort: first sorted position of every grid cell, nullptr for a cloud
// with more cells than the table holds): a run of cells of one grid row is then two table reads instead of two
// bisections of the sorted keys.
template <typename Fn>
__device__ __forceinline__ void sweep(const Query& q, const uint32_t* __restrict__ keys,
                                      const uint32_t* __restrict__ bins, const float4* __restrict__ cand, int N,
                                      int lane, Fn fn) {
  const int nrows = q.wz.cells * q.wy.cells;
  const int nseg = nrows * q.wx.n;
  for (int s0 = 0; s0 < nseg; s0 += 32) {
    const int e = s0 + lane;
    int start = 0, len = 0;
    if (e < nseg) {
      int xr = 0, r = e;
      if (q.wx.n > 1) {  // dilated windows only: up to three x ranges per row
        r = e / q.wx.n;
        xr = e - r * q.wx.n;
      }
      const int rz = r / q.wy.cells;
      const int cy = window_cell(q.wy, r - rz * q.wy.cells);
      const int cz = window_cell(q.wz, rz);
      const uint32_t rowkey = (uint32_t)((cz * q.dimy + cy) * q.dimx);
      if (bins) {
        start = (int)__ldg(bins + rowkey + (uint32_t)q.wx.lo[xr]);
        len = (int)__ldg(bins + rowkey + (uint32_t)q.wx.hi[xr] + 1u) - start;
      } else {
        start = lower_bound_key(keys, N, rowkey + (uint32_t)q.wx.lo[xr]);
        len = lower_bound_key(keys, N, rowkey + (uint32_t)q.wx.hi[xr] + 1u) - start;
      }
    }
    int incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(C3P_FULL_MASK, incl, o);
      if (lane >= o) incl += u;
    }
    const int total = __shfl_sync(C3P_FULL_MASK, incl, 31);
    for (int t0 = 0; t0 < total; t0 += 32) {
      const int t = t0 + lane;
      const bool valid = t < total;
      const int tt = valid ? t : 0;
      int pos = 0;
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) {
        int vv = __shfl_sync(C3P_FULL_MASK, incl, pos + s - 1);
        if (vv <= tt) pos += s;
      }
      const int sstart = __shfl_sync(C3P_FULL_MASK, start, pos);
      const int sincl = __shfl_sync(C3P_FULL_MASK, incl, pos);
      const int slen = __shfl_sync(C3P_FULL_MASK, len, pos);
      int f = -1, j = 0;
      if (valid) {
        const float4 c = __ldg(cand + sstart + (tt - (sincl - slen)));
        j = __float_as_int(c.w);
        // closed box, tf_conv3p_atrous.cpp:277
        if (!(c.x < q.lo[0] || c.x > q.hi[0] || c.y < q.lo[1] || c.y > q.hi[1] ||
              c.z < q.lo[2] || c.z > q.hi[2])) {
          // (tap_of_fast: bit-identical to the IEEE divide of :280-282, see common.cuh)
          int tx = tap_of_fast(c.x, q.lo[0], q.voxel, q.inv_voxel, q.full[0], q.stride[0]);
          int ty = tap_of_fast(c.y, q.lo[1], q.voxel, q.inv_voxel, q.full[1], q.stride[1]);
          int tz = tap_of_fast(c.z, q.lo[2], q.voxel, q.inv_voxel, q.full[2], q.stride[2]);
          if ((tx | ty | tz) >= 0) f = (tz * 3 + ty) * 3 + tx;  // :290
        }
      }
      fn(f, j);
    }
  }
}

// Position of an accepted pair inside its point's list: lists are grouped by ascending cell, members
// of a cell keep visiting order.  pre[f] = first slot of cell f, run[f] = members placed so far.
__device__ __forceinline__ int place(bool valid, int f, int* pre, int* run) {
  const unsigned peers = __match_any_sync(C3P_FULL_MASK, valid ? f : C3P_NCELL);
  const int rank = __popc(peers & lanemask_lt());
  int slot = -1;
  if (valid) slot = pre[f] + run[f] + rank;
  __syncwarp();
  if (valid && rank == 0) run[f] += __popc(peers);
  __syncwarp();
  return slot;
}

// S = compile-time isotropic stride 1..4 (every reference layer: the tap arithmetic then has no runtime
// integer division), 0 = per-axis runtime strides.
template <int S>
__global__ void __launch_bounds__(NB_THREADS)
k_neighbor_search(int B, int N, int sx_, int sy_, int sz_, float voxel, long long capacity, PlanView v) {
  const int sx = S ? S : sx_, sy = S ? S : sy_, sz = S ? S : sz_;
  __shared__ uint32_t stash[NB_WARPS][STASH_CAP];
  __shared__ uint16_t stash_rank[NB_WARPS][STASH_CAP];   // position of the pair among its cell's members (visiting order)
  __shared__ int wcnt[NB_WARPS][32];
  __shared__ int wpre[NB_WARPS][32];
  __shared__ int wrun[NB_WARPS][32];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long qpos = (long long)blockIdx.x * NB_WARPS + warp;
  if (qpos >= (long long)B * N) return;
  const int b = (int)(qpos / N);
  const float* meta = v.cloud_meta + 8 * b;
  const float vmin[3] = {meta[0], meta[1], meta[2]};
  const int dim[3] = {__float_as_int(meta[4]), __float_as_int(meta[5]), __float_as_int(meta[6])};
  const uint32_t* keys = v.sorted_key + (size_t)b * N;
  const uint32_t* bins = (__float_as_int(meta[7]) & 256) ? v.cell_start + (size_t)b * ((size_t)v.cell_cap + 1) : nullptr;
  const float4* cand = v.sorted_xyzi + (size_t)b * N;
  const float4 me = v.sorted_xyzi[qpos];
  const size_t row = (size_t)b * N + __float_as_int(me.w);

  Query q;
  q.voxel = voxel;
  q.inv_voxel = __fdiv_rn(1.0f, voxel);
  q.stride[0] = sx; q.stride[1] = sy; q.stride[2] = sz;
  const float centre[3] = {me.x, me.y, me.z};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    q.full[a] = 2 * q.stride[a] + 1;  // (3-1)*stride+1, :235-237
    q.lo[a] = box_lo(centre[a], q.full[a], voxel);
    q.hi[a] = box_hi(centre[a], q.full[a], voxel);
  }
  axis_window(q.lo[0], q.hi[0], sx, voxel, vmin[0], dim[0], q.wx);
  axis_window(q.lo[1], q.hi[1], sy, voxel, vmin[1], dim[1], q.wy);
  axis_window(q.lo[2], q.hi[2], sz, voxel, vmin[2], dim[2], q.wz);
  q.dimx = dim[0];
  q.dimy = dim[1];

  int* cnt = wcnt[warp];
  int* pre = wpre[warp];
  int* run = wrun[warp];
  uint32_t* st = stash[warp];
  uint16_t* sr = stash_rank[warp];
  cnt[lane] = 0;
  run[lane] = 0;
  __syncwarp();

  // pass 1: count per cell, keep the first STASH_CAP pairs in shared memory together with their rank inside their
  // cell (members of a cell keep visiting order): one match per chunk instead of a shared-memory atomic per pair
  // here and a second match per pair when the list is written
  int found = 0;
  sweep(q, keys, bins, cand, N, lane, [&](int f, int j) {
    const bool hit = f >= 0;
    const unsigned hits = __ballot_sync(C3P_FULL_MASK, hit);
    const unsigned peers = __match_any_sync(C3P_FULL_MASK, hit ? f : C3P_NCELL);
    const unsigned before = peers & lanemask_lt();
    int rank = 0;
    if (hit) rank = cnt[f] + __popc(before);
    __syncwarp();
    if (hit && before == 0u) cnt[f] += __popc(peers);
    __syncwarp();
    if (hit) {
      const int slot = found + __popc(hits & lanemask_lt());
      if (slot < STASH_CAP) {
        st[slot] = (uint32_t)j | ((uint32_t)f << 27);
        sr[slot] = (uint16_t)rank;
      }
    }
    found += __popc(hits);
  });
  __syncwarp();
  const int K = found;
  const int mine = lane < C3P_NCELL ? cnt[lane] : 0;
  if (lane < C3P_NCELL) v.count_table[row * C3P_NCELL + lane] = mine;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int u = __shfl_up_sync(C3P_FULL_MASK, incl, o);
    if (lane >= o) incl += u;
  }
  pre[lane] = incl - mine;
  long long begin = 0;
  if (lane == 0) begin = (long long)atomicAdd((unsigned long long*)&v.header[H_CURSOR],
                                              (unsigned long long)K);
  begin = __shfl_sync(C3P_FULL_MASK, begin, 0);
  if (lane == 0) {
    v.pair_begin[row] = begin;
    v.pair_len[row] = K;
  }
  if (begin + K > capacity) {  // lists incomplete: flagged, caller rebuilds with a larger capacity
    if (lane == 0) v.header[H_OVERFLOW] = 1;
    return;
  }
  __syncwarp();
  int* out = v.pair_row + begin;
  const int rowbase = b * N;
  if (K <= STASH_CAP) {
    for (int c = lane; c < K; c += 32) {
      const uint32_t e = st[c];
      out[pre[e >> 27] + (int)sr[c]] = rowbase + (int)(e & J_MASK);
    }
  } else {  // pass 2 for very dense neighbourhoods: same visiting order, placed directly
    sweep(q, keys, bins, cand, N, lane, [&](int f, int j) {
      const int slot = place(f >= 0, f, pre, run);
      if (f >= 0) out[slot] = rowbase + j;
    });
  }
}

// Backward lists: for j and every ii in N(j) (j's forward list), the cell of j in ii's frame with NO
// box test; dropped if it is a hole or count(ii, f') == 0 -- tf_conv3p_atrous.cpp:654-679.  Stored at
// the same offsets as the forward lists (a backward list is never longer), grouped by f'.
template <int S>
__global__ void __launch_bounds__(NB_THREADS)
k_backward_lists(int B, int N, int sx_, int sy_, int sz_, float voxel, long long capacity,
                 const float* __restrict__ points, PlanView v) {
  const int sx = S ? S : sx_, sy = S ? S : sy_, sz = S ? S : sz_;
  __shared__ int wcnt[NB_WARPS][32];
  __shared__ int wpre[NB_WARPS][32];
  __shared__ int wrun[NB_WARPS][32];
  __shared__ uint32_t bstash[NB_WARPS][STASH_CAP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long qpos = (long long)blockIdx.x * NB_WARPS + warp;
  if (qpos >= (long long)B * N) return;
  const int b = (int)(qpos / N);
  const float4 me = v.sorted_xyzi[qpos];
  const size_t row = (size_t)b * N + __float_as_int(me.w);
  const int K = v.pair_len[row];
  const long long begin = v.pair_begin[row];
  int* cnt = wcnt[warp];
  int* pre = wpre[warp];
  int* run = wrun[warp];
  cnt[lane] = 0;
  run[lane] = 0;
  __syncwarp();
  if (begin + K > capacity) {
    if (lane < C3P_NCELL) v.bwd_count[row * C3P_NCELL + lane] = 0;
    return;
  }
  const int full[3] = {2 * sx + 1, 2 * sy + 1, 2 * sz + 1};
  const float inv_voxel = __fdiv_rn(1.0f, voxel);
  const int* fwd = v.pair_row + begin;

  auto rebin = [&](int m, int& ii, int& members) -> int {
    ii = __ldg(fwd + m);
    const float kx = __ldg(points + 3 * (size_t)ii), ky = __ldg(points + 3 * (size_t)ii + 1),
                kz = __ldg(points + 3 * (size_t)ii + 2);
    int tx = tap_of_fast(me.x, box_lo(kx, full[0], voxel), voxel, inv_voxel, full[0], sx);  // :658-669
    int ty = tap_of_fast(me.y, box_lo(ky, full[1], voxel), voxel, inv_voxel, full[1], sy);
    int tz = tap_of_fast(me.z, box_lo(kz, full[2], voxel), voxel, inv_voxel, full[2], sz);
    if ((tx | ty | tz) < 0) return -1;                                      // :672
    const int f = (tz * 3 + ty) * 3 + tx;                                   // :677
    members = __ldg(v.count_table + (size_t)ii * C3P_NCELL + f);
    return members == 0 ? -1 : f;                                           // :679
  };

  // pass 1: cell of every entry, its rank inside the cell (visiting order) and the cell's member count, kept in shared
  // memory for lists of up to STASH_CAP entries (longer lists re-bin a second time)
  uint32_t* st = bstash[warp];
  for (int c = 0; c < K; c += 32) {
    int ii = 0, members = 1, f = -1;
    if (c + lane < K) f = rebin(c + lane, ii, members);
    const bool hit = f >= 0;
    const unsigned peers = __match_any_sync(C3P_FULL_MASK, hit ? f : C3P_NCELL);
    const unsigned before = peers & lanemask_lt();
    int rank = 0;
    if (hit) rank = cnt[f] + __popc(before);
    __syncwarp();
    if (hit && before == 0u) cnt[f] += __popc(peers);
    __syncwarp();
    if (c + lane < K && c + lane < STASH_CAP)   // f (5 bits, 31 = dropped) | rank (9 bits) | members (18 bits, saturating)
      st[c + lane] = (uint32_t)(hit ? f : 31) | ((uint32_t)rank << 5) | ((uint32_t)min(members, (1 << 18) - 1) << 14);
  }
  __syncwarp();
  const int mine = lane < C3P_NCELL ? cnt[lane] : 0;
  if (lane < C3P_NCELL) v.bwd_count[row * C3P_NCELL + lane] = mine;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int u = __shfl_up_sync(C3P_FULL_MASK, incl, o);
    if (lane >= o) incl += u;
  }
  pre[lane] = incl - mine;
  const int Kb = __shfl_sync(C3P_FULL_MASK, incl, 31);
  if (lane == 0) atomicAdd((unsigned long long*)&v.header[H_BWD_PAIRS], (unsigned long long)Kb);
  __syncwarp();
  if (K <= STASH_CAP) {
    for (int m = lane; m < K; m += 32) {
      const uint32_t e = st[m];
      const int f = (int)(e & 31u);
      if (f == 31) continue;
      int members = (int)(e >> 14);
      const long long slot = begin + pre[f] + (int)((e >> 5) & 511u);
      const int ii = __ldg(fwd + m);
      if (members == (1 << 18) - 1) members = __ldg(v.count_table + (size_t)ii * C3P_NCELL + f);   // saturated: re-read
      v.bwd_row[slot] = ii;
      v.bwd_weight[slot] = __fdiv_rn(1.0f, (float)members);
    }
  } else {
    for (int c = 0; c < K; c += 32) {
      int ii = 0, members = 1, f = -1;
      if (c + lane < K) f = rebin(c + lane, ii, members);
      const int slot = place(f >= 0, f, pre, run);
      if (f >= 0) {
        v.bwd_row[begin + slot] = ii;
        v.bwd_weight[begin + slot] = __fdiv_rn(1.0f, (float)members);
      }
    }
  }
}

int launch_neighbor_search(const conv3p_geom_t* g, const PlanView& v, cudaStream_t stream) {
  const long long pts = (long long)g->B * g->N;
  C3P_CUDA(cudaMemsetAsync(v.header, 0, sizeof(long long) * H_SLOTS, stream));
  if (pts == 0) return CONV3P_OK;
  const unsigned grid = (unsigned)((pts + NB_WARPS - 1) / NB_WARPS);
  {
    LaunchTimer timer_("k_neighbor_search", stream);
    const int iso = (g->stride[0] == g->stride[1] && g->stride[1] == g->stride[2] && g->stride[0] <= 4)
                        ? g->stride[0] : 0;
    auto kern = iso == 1 ? k_neighbor_search<1> : iso == 2 ? k_neighbor_search<2>
              : iso == 3 ? k_neighbor_search<3> : iso == 4 ? k_neighbor_search<4> : k_neighbor_search<0>;
    kern<<<grid, NB_THREADS, 0, stream>>>(g->B, g->N, g->stride[0], g->stride[1], g->stride[2],
                                          g->voxel_size, g->pair_capacity, v);
  }
  C3P_LAUNCH_CHECK("k_neighbor_search");
  return CONV3P_OK;
}

int launch_backward_lists(const conv3p_geom_t* g, const float* points, const PlanView& v,
                          cudaStream_t stream) {
  const long long pts = (long long)g->B * g->N;
  C3P_CUDA(cudaMemsetAsync(v.header + H_BWD_PAIRS, 0, sizeof(long long), stream));
  if (pts > 0) {
    const unsigned grid = (unsigned)((pts + NB_WARPS - 1) / NB_WARPS);
    {
    LaunchTimer timer_("k_backward_lists", stream);
    const int iso = (g->stride[0] == g->stride[1] && g->stride[1] == g->stride[2] && g->stride[0] <= 4)
                        ? g->stride[0] : 0;
    auto kern = iso == 1 ? k_backward_lists<1> : iso == 2 ? k_backward_lists<2>
              : iso == 3 ? k_backward_lists<3> : iso == 4 ? k_backward_lists<4> : k_backward_lists<0>;
    kern<<<grid, NB_THREADS, 0, stream>>>(g->B, g->N, g->stride[0], g->stride[1], g->stride[2],
                                          g->voxel_size, g->pair_capacity, points, v);
  }
    C3P_LAUNCH_CHECK("k_backward_lists");
  }
  return CONV3P_OK;
}

}  // namespace c3p
