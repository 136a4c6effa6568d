// generic_filter.cu -- Conv3p / Conv3pGrad for filter shapes other than 3x3x3.
//
// The reference operator reads the filter dimensions from the tensor (tf_conv3p_atrous.cpp:425-427) and is generic in
// them, although every model of the repository uses 3x3x3.  The tuned engines of this library are specialised to 27
// cells; this file is the general path: any fz x fy x fx with at most 512 cells, any per-axis stride, fp32 SIMT.  It
// follows the reference step by step -- the box of (filter-1)*stride+1 voxels (:235-245), the closed box test (:277),
// the voxel index with IEEE division and clamp (:280-282), the dilation holes (:285-288), cell
// f = (fz*filter_y + fy)*filter_x + fx (:290), the per-cell mean in forward (:486-494) and the j-centred re-binning
// without box test plus the count == 0 skip in backward (:654-696) -- on top of the same voxel sort and bin-offset
// table as the 3x3x3 plan.  grad_filter is accumulated per cloud in a fixed order and reduced deterministically.
//
// Workspace: [standard plan buffer (sort, lists) | count table [B*N][cells] | forward cell of every pair |
//             backward cell of every pair | per-cloud grad_filter partials].
#include "common.cuh"

namespace c3p {

constexpr int GEN_WARPS = 8;
constexpr int GEN_THREADS = GEN_WARPS * 32;
constexpr int GEN_MAX_CELLS = 512;

struct GenGeom {
  int fx, fy, fz;      // filter dims
  int sx, sy, sz;      // strides
  int full[3];         // (filter - 1) * stride + 1 per axis (x, y, z)
  int cells;
  float voxel;
};

struct GenView {
  int* count;          // [B*N][cells]
  int* pair_f;         // [capacity] forward cell of every pair (pairs in visiting order)
  int* bwd_f;          // [capacity] backward cell f' of every kept backward pair
  float* partial;      // [B][cells*Cin*Cout]
};

static GenGeom make_gen(const conv3p_geom_t* g, const int dims_zyx[3]) {
  GenGeom q;
  q.fz = dims_zyx[0]; q.fy = dims_zyx[1]; q.fx = dims_zyx[2];
  q.sx = g->stride[0]; q.sy = g->stride[1]; q.sz = g->stride[2];
  q.full[0] = (q.fx - 1) * q.sx + 1;
  q.full[1] = (q.fy - 1) * q.sy + 1;
  q.full[2] = (q.fz - 1) * q.sz + 1;
  q.cells = q.fx * q.fy * q.fz;
  q.voxel = g->voxel_size;
  return q;
}

// Visits every neighbour of the query point `me` (cloud b): fn(j, f) with f the kernel cell, warp-collectively (lanes
// that have no accepted candidate in a round pass f = -1).  Candidates: the grid cells overlapped by the box (bin
// table), or the whole cloud when the cloud has no table.
template <typename Fn>
__device__ __forceinline__ void gen_sweep(const GenGeom& q, const PlanView& v, int b, int N, const float4& me, int lane, Fn fn) {
  const float* meta = v.cloud_meta + 8 * b;
  const int dimx = __float_as_int(meta[4]), dimy = __float_as_int(meta[5]), dimz = __float_as_int(meta[6]);
  const bool table = (__float_as_int(meta[7]) & 256) != 0;
  const float4* cand = v.sorted_xyzi + (size_t)b * N;
  float lo[3], hi[3];
  const float c3[3] = {me.x, me.y, me.z};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    lo[a] = box_lo(c3[a], q.full[a], q.voxel);
    hi[a] = box_hi(c3[a], q.full[a], q.voxel);
  }
  // The reference only visits grid cells within n = (int)((full + 1) * 0.5) of the query's cell per axis (:247-267).
  // For odd boxes (every 3-tap filter: half-extent s + 0.5 voxels, n = s + 1) that window always covers the box, which
  // is why the 3x3x3 engines need no such test; for even boxes (half-extent == n voxels exactly) a point on the closed
  // edge whose cell index rounds one further out is NOT visited by the reference -- reproduced here with the
  // reference's own cell arithmetic, (int)((x - min) / r) (:192-194).
  auto cell_of = [&](float x, int a) -> int { return __float2int_rz(__fdiv_rn(__fsub_rn(x, meta[a]), q.voxel)); };
  const int ci[3] = {cell_of(me.x, 0), cell_of(me.y, 1), cell_of(me.z, 2)};
  const int nwin[3] = {(q.full[0] + 1) / 2, (q.full[1] + 1) / 2, (q.full[2] + 1) / 2};
  auto test = [&](const float4& c) -> int {
    if (abs(cell_of(c.x, 0) - ci[0]) > nwin[0] || abs(cell_of(c.y, 1) - ci[1]) > nwin[1] ||
        abs(cell_of(c.z, 2) - ci[2]) > nwin[2])
      return -1;
    if (c.x < lo[0] || c.x > hi[0] || c.y < lo[1] || c.y > hi[1] || c.z < lo[2] || c.z > hi[2]) return -1;  // :277
    const int tx = tap_of(c.x, lo[0], q.voxel, q.full[0], q.sx);
    const int ty = tap_of(c.y, lo[1], q.voxel, q.full[1], q.sy);
    const int tz = tap_of(c.z, lo[2], q.voxel, q.full[2], q.sz);
    if ((tx | ty | tz) < 0) return -1;
    return (tz * q.fy + ty) * q.fx + tx;                                                                     // :290
  };
  if (!table) {
    for (int s0 = 0; s0 < N; s0 += 32) {
      const int s = s0 + lane;
      int f = -1, j = 0;
      if (s < N) {
        const float4 c = __ldg(cand + s);
        j = __float_as_int(c.w);
        f = test(c);
      }
      fn(j, f);
    }
    return;
  }
  const uint32_t* bins = v.cell_start + (size_t)b * ((size_t)v.cell_cap + 1);
  const float slop = q.voxel * (1.0f / 1024.0f) + fmaxf(fmaxf(fabsf(me.x), fabsf(me.y)), fabsf(me.z)) * 1e-6f;
  const int x0 = grid_coord(lo[0] - slop, meta[0], q.voxel, dimx), x1 = grid_coord(hi[0] + slop, meta[0], q.voxel, dimx);
  const int y0 = grid_coord(lo[1] - slop, meta[1], q.voxel, dimy), y1 = grid_coord(hi[1] + slop, meta[1], q.voxel, dimy);
  const int z0 = grid_coord(lo[2] - slop, meta[2], q.voxel, dimz), z1 = grid_coord(hi[2] + slop, meta[2], q.voxel, dimz);
  for (int cz = z0; cz <= z1; ++cz)
    for (int cy = y0; cy <= y1; ++cy) {
      const uint32_t rowkey = (uint32_t)((cz * dimy + cy) * dimx);
      const int start = (int)__ldg(bins + rowkey + (uint32_t)x0), end = (int)__ldg(bins + rowkey + (uint32_t)x1 + 1u);
      for (int s0 = start; s0 < end; s0 += 32) {
        const int s = s0 + lane;
        int f = -1, j = 0;
        if (s < end) {
          const float4 c = __ldg(cand + s);
          j = __float_as_int(c.w);
          f = test(c);
        }
        fn(j, f);
      }
    }
}

// Count table + pair lists.  One warp per point (sorted order).  Two sweeps: counts, then the list.
__global__ void __launch_bounds__(GEN_THREADS)
k_generic_search(GenGeom q, int B, int N, long long capacity, PlanView v, GenView gv) {
  extern __shared__ int gen_cnt[];   // [GEN_WARPS][cells]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long qpos = (long long)blockIdx.x * GEN_WARPS + warp;
  if (qpos >= (long long)B * N) return;
  const int b = (int)(qpos / N);
  const float4 me = v.sorted_xyzi[qpos];
  const size_t row = (size_t)b * N + __float_as_int(me.w);
  int* cnt = gen_cnt + warp * q.cells;
  for (int f = lane; f < q.cells; f += 32) cnt[f] = 0;
  __syncwarp();
  int found = 0;
  gen_sweep(q, v, b, N, me, lane, [&](int j, int f) {
    if (f >= 0) atomicAdd(&cnt[f], 1);
    found += __popc(__ballot_sync(C3P_FULL_MASK, f >= 0));
  });
  __syncwarp();
  for (int f = lane; f < q.cells; f += 32) gv.count[row * q.cells + f] = cnt[f];
  long long begin = 0;
  if (lane == 0) begin = (long long)atomicAdd((unsigned long long*)&v.header[H_CURSOR], (unsigned long long)found);
  begin = __shfl_sync(C3P_FULL_MASK, begin, 0);
  if (lane == 0) {
    v.pair_begin[row] = begin;
    v.pair_len[row] = found;
  }
  if (begin + found > capacity) {
    if (lane == 0) v.header[H_OVERFLOW] = 1;
    return;
  }
  int at = 0;
  gen_sweep(q, v, b, N, me, lane, [&](int j, int f) {
    const unsigned hits = __ballot_sync(C3P_FULL_MASK, f >= 0);
    if (f >= 0) {
      const long long slot = begin + at + __popc(hits & lanemask_lt());
      v.pair_row[slot] = b * N + j;
      gv.pair_f[slot] = f;
    }
    at += __popc(hits);
  });
}

// out[i, c] = sum over pairs (j, f) of W[f, k, c] * in[j, k] / count(i, f)   (:486-494).  One warp per point, lanes over c.
__global__ void __launch_bounds__(GEN_THREADS)
k_generic_forward(GenGeom q, long long pts, long long capacity, int Cin, int Cout, PlanView v, GenView gv,
                  const float* __restrict__ input, const float* __restrict__ filter, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * GEN_WARPS + (threadIdx.x >> 5);
  if (i >= pts) return;
  const long long begin = v.pair_begin[i];
  const int K = v.pair_len[i];
  const bool ok = begin + K <= capacity;
  for (int c0 = 0; c0 < Cout; c0 += 32) {
    const int c = c0 + lane;
    float acc = 0.f;
    if (ok && c < Cout) {
      for (int m = 0; m < K; ++m) {
        const int j = __ldg(v.pair_row + begin + m), f = __ldg(gv.pair_f + begin + m);
        const float inv = 1.0f / (float)__ldg(gv.count + (size_t)i * q.cells + f);
        const float* w = filter + (size_t)f * Cin * Cout + c;
        const float* x = input + (size_t)j * Cin;
        for (int k = 0; k < Cin; ++k) acc = fmaf(__ldg(w + (size_t)k * Cout), __ldg(x + k) * inv, acc);
      }
    }
    if (c < Cout) out[(size_t)i * Cout + c] = ok ? acc : __int_as_float(0x7fc00000);
  }
}

// Backward lists + grad_input.  One warp per j: for ii in N(j), f' = cell of j in ii's frame (no box test), dropped when
// it is a hole or count(ii, f') == 0 (:654-679); grad_in[j, k] += g[ii, c] * W[f', k, c] / count(ii, f')   (:692).
__global__ void __launch_bounds__(GEN_THREADS)
k_generic_backward_input(GenGeom q, long long pts, long long capacity, int Cin, int Cout, PlanView v, GenView gv,
                         const float* __restrict__ points, const float* __restrict__ grad_out,
                         const float* __restrict__ filter, float* __restrict__ grad_in) {
  const int lane = threadIdx.x & 31;
  const long long j = (long long)blockIdx.x * GEN_WARPS + (threadIdx.x >> 5);
  if (j >= pts) return;
  const long long begin = v.pair_begin[j];
  const int K = v.pair_len[j];
  const bool ok = begin + K <= capacity;
  const float px = points[3 * j], py = points[3 * j + 1], pz = points[3 * j + 2];
  // pass 1 (lanes over pairs): the kept backward pairs, compacted in place
  int kept = 0;
  if (ok) {
    for (int m0 = 0; m0 < K; m0 += 32) {
      const int m = m0 + lane;
      int ii = 0, f = -1, members = 1;
      if (m < K) {
        ii = __ldg(v.pair_row + begin + m);
        const int tx = tap_of(px, box_lo(__ldg(points + 3 * (size_t)ii), q.full[0], q.voxel), q.voxel, q.full[0], q.sx);
        const int ty = tap_of(py, box_lo(__ldg(points + 3 * (size_t)ii + 1), q.full[1], q.voxel), q.voxel, q.full[1], q.sy);
        const int tz = tap_of(pz, box_lo(__ldg(points + 3 * (size_t)ii + 2), q.full[2], q.voxel), q.voxel, q.full[2], q.sz);
        if ((tx | ty | tz) >= 0) {
          f = (tz * q.fy + ty) * q.fx + tx;
          members = __ldg(gv.count + (size_t)ii * q.cells + f);
          if (members == 0) f = -1;                                                              // :679
        }
      }
      const unsigned keep = __ballot_sync(C3P_FULL_MASK, f >= 0);
      if (f >= 0) {
        const long long slot = begin + kept + __popc(keep & lanemask_lt());
        v.bwd_row[slot] = ii;
        v.bwd_weight[slot] = __fdiv_rn(1.0f, (float)members);
        gv.bwd_f[slot] = f;
      }
      kept += __popc(keep);
    }
  }
  __syncwarp();
  if (lane == 0) v.bwd_count[j] = kept;   // (reused as the backward list length of j)
  // pass 2 (lanes over k)
  for (int k0 = 0; k0 < Cin; k0 += 32) {
    const int k = k0 + lane;
    float acc = 0.f;
    if (ok && k < Cin) {
      for (int m = 0; m < kept; ++m) {
        const int ii = v.bwd_row[begin + m], f = gv.bwd_f[begin + m];
        const float wgt = v.bwd_weight[begin + m];
        const float* w = filter + ((size_t)f * Cin + k) * Cout;
        const float* g = grad_out + (size_t)ii * Cout;
        float s = 0.f;
        for (int c = 0; c < Cout; ++c) s = fmaf(__ldg(g + c), __ldg(w + c), s);
        acc = fmaf(s, wgt, acc);
      }
    }
    if (k < Cin && grad_in) grad_in[(size_t)j * Cin + k] = ok ? acc : __int_as_float(0x7fc00000);
  }
}

// grad_filter[f', k, c] += g[ii, c] * in[j, k] / count(ii, f')   (:696).  One CTA per cloud walks its points and their
// backward pairs in order; thread e owns elements e, e + blockDim, ... of the cloud's partial (fixed order of additions).
__global__ void __launch_bounds__(GEN_THREADS)
k_generic_backward_filter(GenGeom q, int N, long long capacity, int Cin, int Cout, PlanView v, GenView gv,
                          const float* __restrict__ grad_out, const float* __restrict__ input) {
  const int b = blockIdx.x;
  const int KC = Cin * Cout;
  float* part = gv.partial + (size_t)b * q.cells * KC;
  for (int e = threadIdx.x; e < q.cells * KC; e += GEN_THREADS) part[e] = 0.f;
  __syncthreads();
  for (int jj = 0; jj < N; ++jj) {
    const size_t j = (size_t)b * N + jj;
    const long long begin = v.pair_begin[j];
    if (begin + v.pair_len[j] > capacity) continue;
    const int kept = v.bwd_count[j];
    for (int m = 0; m < kept; ++m) {
      const int ii = v.bwd_row[begin + m], f = gv.bwd_f[begin + m];
      const float wgt = v.bwd_weight[begin + m];
      float* pf = part + (size_t)f * KC;
      for (int e = threadIdx.x; e < KC; e += GEN_THREADS) {
        const int k = e / Cout, c = e - k * Cout;
        pf[e] = fmaf(__ldg(grad_out + (size_t)ii * Cout + c) * wgt, __ldg(input + j * Cin + k), pf[e]);
      }
    }
  }
}

bool generic_filter_supported(const int dims_zyx[3]) {
  if (!dims_zyx) return false;
  for (int a = 0; a < 3; ++a)
    if (dims_zyx[a] < 1 || dims_zyx[a] > 64) return false;
  return (long long)dims_zyx[0] * dims_zyx[1] * dims_zyx[2] <= GEN_MAX_CELLS;
}

static size_t gen_extra_bytes(const conv3p_geom_t* g, int cells, int Cin, int Cout) {
  const size_t pts = (size_t)g->B * g->N, cap = (size_t)g->pair_capacity;
  return align_up(sizeof(int) * pts * cells) + 2 * align_up(sizeof(int) * cap) +
         align_up(sizeof(float) * (size_t)g->B * cells * Cin * Cout) + 256;
}

size_t generic_workspace_bytes(const conv3p_geom_t* g, const int dims_zyx[3], int Cin, int Cout) {
  if (check_geom(g) || !generic_filter_supported(dims_zyx)) return 0;
  const size_t plan = conv3p_plan_bytes(g);
  if (!plan) return 0;
  return plan + gen_extra_bytes(g, dims_zyx[0] * dims_zyx[1] * dims_zyx[2], Cin, Cout);
}

static GenView carve_gen(const conv3p_geom_t* g, int cells, int Cin, int Cout, void* base) {
  const size_t pts = (size_t)g->B * g->N, cap = (size_t)g->pair_capacity;
  char* p = static_cast<char*>(base);
  GenView gv;
  gv.count = reinterpret_cast<int*>(p); p += align_up(sizeof(int) * pts * cells);
  gv.pair_f = reinterpret_cast<int*>(p); p += align_up(sizeof(int) * cap);
  gv.bwd_f = reinterpret_cast<int*>(p); p += align_up(sizeof(int) * cap);
  gv.partial = reinterpret_cast<float*>(p);
  return gv;
}

// Sort + search into the workspace; returns the views.
static int gen_plan(const conv3p_geom_t* g, const GenGeom& q, const float* points, int Cin, int Cout, void* ws,
                    size_t ws_bytes, cudaStream_t stream, PlanView* v, GenView* gv) {
  const size_t plan = conv3p_plan_bytes(g);
  if (!ws || ws_bytes < plan + gen_extra_bytes(g, q.cells, Cin, Cout)) return CONV3P_ERR_BUFFER_TOO_SMALL;
  int st = make_view(g, ws, plan, v);
  if (st) return st;
  *gv = carve_gen(g, q.cells, Cin, Cout, static_cast<char*>(ws) + plan);
  C3P_CUDA(cudaMemsetAsync(v->header, 0, sizeof(long long) * H_SLOTS, stream));
  const long long pts = (long long)g->B * g->N;
  if (pts == 0) return CONV3P_OK;
  st = launch_cloud_sort(g, points, *v, stream);
  if (st) return st;
  {
    LaunchTimer timer_("k_generic_search", stream);
    k_generic_search<<<(unsigned)((pts + GEN_WARPS - 1) / GEN_WARPS), GEN_THREADS, sizeof(int) * GEN_WARPS * q.cells, stream>>>(
        q, g->B, g->N, g->pair_capacity, *v, *gv);
  }
  C3P_LAUNCH_CHECK("k_generic_search");
  return CONV3P_OK;
}

int generic_forward(const conv3p_geom_t* g, const int dims_zyx[3], const float* points, const float* input,
                    const float* filter, int Cin, int Cout, float* output, void* ws, size_t ws_bytes,
                    cudaStream_t stream) {
  if (!generic_filter_supported(dims_zyx)) return CONV3P_ERR_UNSUPPORTED;
  const GenGeom q = make_gen(g, dims_zyx);
  PlanView v;
  GenView gv;
  int st = gen_plan(g, q, points, Cin, Cout, ws, ws_bytes, stream, &v, &gv);
  if (st) return st;
  const long long pts = (long long)g->B * g->N;
  if (pts == 0) return CONV3P_OK;
  if (!points || !input || !filter || !output) return CONV3P_ERR_INVALID_ARGUMENT;
  {
    LaunchTimer timer_("k_generic_forward", stream);
    k_generic_forward<<<(unsigned)((pts + GEN_WARPS - 1) / GEN_WARPS), GEN_THREADS, 0, stream>>>(
        q, pts, g->pair_capacity, Cin, Cout, v, gv, input, filter, output);
  }
  C3P_LAUNCH_CHECK("k_generic_forward");
  return CONV3P_OK;
}

int generic_backward(const conv3p_geom_t* g, const int dims_zyx[3], const float* grad_out, const float* points,
                     const float* input, const float* filter, int Cin, int Cout, float* grad_input,
                     float* grad_filter, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (!generic_filter_supported(dims_zyx)) return CONV3P_ERR_UNSUPPORTED;
  const GenGeom q = make_gen(g, dims_zyx);
  PlanView v;
  GenView gv;
  int st = gen_plan(g, q, points, Cin, Cout, ws, ws_bytes, stream, &v, &gv);
  if (st) return st;
  const long long pts = (long long)g->B * g->N;
  const long long nW = (long long)q.cells * Cin * Cout;
  if (pts == 0) {
    if (grad_filter) C3P_CUDA(cudaMemsetAsync(grad_filter, 0, sizeof(float) * nW, stream));
    return CONV3P_OK;
  }
  if (!points || !grad_out || !input || !filter) return CONV3P_ERR_INVALID_ARGUMENT;
  // grad_input kernel: also builds the backward lists the weight gradient walks (grad_input itself may be skipped)
  float* gi = grad_input;
  {
    LaunchTimer timer_("k_generic_backward_input", stream);
    k_generic_backward_input<<<(unsigned)((pts + GEN_WARPS - 1) / GEN_WARPS), GEN_THREADS, 0, stream>>>(
        q, pts, g->pair_capacity, Cin, Cout, v, gv, points, grad_out, filter, gi);
  }
  C3P_LAUNCH_CHECK("k_generic_backward_input");
  if (grad_filter) {
    {
      LaunchTimer timer_("k_generic_backward_filter", stream);
      k_generic_backward_filter<<<g->B, GEN_THREADS, 0, stream>>>(q, g->N, g->pair_capacity, Cin, Cout, v, gv, grad_out, input);
    }
    C3P_LAUNCH_CHECK("k_generic_backward_filter");
    st = launch_reduce_partials(gv.partial, g->B, nW, grad_filter, v.header, stream);
    if (st) return st;
  }
  return CONV3P_OK;
}

}  // namespace c3p
