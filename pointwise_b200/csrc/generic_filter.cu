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
// T = double: the reference registers both ops for double as well (register_op.cpp:45, 64; CPU kernels
// tf_conv3p_atrous.cpp:516, 727), with points, features, filter and voxel size all in double and the whole predicate
// evaluated in double.  The same kernels, instantiated for double, serve conv3p_op_{forward,backward}_f64 (every filter
// shape, 3x3x3 included): the voxel sort and the bin table are built from the points rounded to float and only
// GENERATE candidates (the window is widened by more than the rounding), every test and every sum is done in double
// on the caller's double data.
//
// Workspace: [standard plan buffer (sort, lists) | count table [B*N][cells] | forward cell of every pair |
//             backward cell of every pair | grad_filter partials per (cloud, chunk) (T) | double only: points as float,
//             per-cloud minimum in double].
#include <algorithm>

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
  double voxel_d;      // T = double
};

struct GenView {
  int* count;          // [B*N][cells]
  int* pair_f;         // [capacity] forward cell of every pair (pairs in visiting order)
  int* bwd_f;          // [capacity] backward cell f' of every kept backward pair
  void* partial;       // [B][cells*Cin*Cout] of T
  float* points_f32;   // T = double: the points rounded to float (candidate generation only)
  double* dmin;        // T = double: per-cloud minimum [B][3] (the reference's vmin, :156-171)
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
  q.voxel_d = (double)g->voxel_size;
  return q;
}

// ---- the predicate's arithmetic in T.  float: the fp32 helpers of the tuned engines (common.cuh).  double: every
// step an explicit IEEE double operation (no FMA contraction), as the reference's x86 object code evaluates it. ---
template <typename T> struct GenAr;
template <> struct GenAr<float> {
  static __device__ __forceinline__ float voxel(const GenGeom& q) { return q.voxel; }
  static __device__ __forceinline__ float lo(float c, int full, float vox) { return box_lo(c, full, vox); }
  static __device__ __forceinline__ float hi(float c, int full, float vox) { return box_hi(c, full, vox); }
  static __device__ __forceinline__ int tap(float v, float lo_, float vox, int full, int stride) { return tap_of(v, lo_, vox, full, stride); }
  static __device__ __forceinline__ int cell(float x, float vmin, float vox) { return __float2int_rz(__fdiv_rn(__fsub_rn(x, vmin), vox)); }
  static __device__ __forceinline__ float rcp(int n) { return __fdiv_rn(1.0f, (float)n); }
};
template <> struct GenAr<double> {
  static __device__ __forceinline__ double voxel(const GenGeom& q) { return q.voxel_d; }
  static __device__ __forceinline__ double lo(double c, int full, double vox) { return __dsub_rn(c, __dmul_rn((double)full * 0.5, vox)); }   // :239
  static __device__ __forceinline__ double hi(double c, int full, double vox) { return __dadd_rn(c, __dmul_rn((double)full * 0.5, vox)); }   // :240
  static __device__ __forceinline__ int tap(double v, double lo_, double vox, int full, int stride) {                                          // :280-288
    int c = __double2int_rz(__ddiv_rn(__dsub_rn(v, lo_), vox));
    c = min(c, full - 1);
    if (c < 0) return -1;
    const int t = c / stride;
    return (t * stride == c) ? t : -1;
  }
  static __device__ __forceinline__ int cell(double x, double vmin, double vox) { return __double2int_rz(__ddiv_rn(__dsub_rn(x, vmin), vox)); }
  static __device__ __forceinline__ double rcp(int n) { return __ddiv_rn(1.0, (double)n); }
};
template <typename T> __device__ __forceinline__ T gen_ldg(const T* p) { return __ldg(p); }
template <typename T> __device__ __forceinline__ T gen_nan();
template <> __device__ __forceinline__ float gen_nan<float>() { return __int_as_float(0x7fc00000); }
template <> __device__ __forceinline__ double gen_nan<double>() { return __longlong_as_double(0x7ff8000000000000LL); }
template <typename T> __device__ __forceinline__ T gen_fma(T a, T b, T c);
template <> __device__ __forceinline__ float gen_fma<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> __device__ __forceinline__ double gen_fma<double>(double a, double b, double c) { return fma(a, b, c); }

// Visits every neighbour of the query point (mx, my, mz) of cloud b: fn(j, f) with f the kernel cell, warp-collectively
// (lanes that have no accepted candidate in a round pass f = -1).  Candidates: the grid cells overlapped by the box
// (bin table), or the whole cloud when the cloud has no table.  T = double: `pts` is the cloud's double coordinates
// (the sorted float copies only name the candidates) and `vmin` the cloud's double minimum.
template <typename T, typename Fn>
__device__ __forceinline__ void gen_sweep(const GenGeom& q, const PlanView& v, int b, int N, T mx, T my, T mz,
                                          const T* __restrict__ pts, const double* __restrict__ dmin, int lane, Fn fn) {
  using A = GenAr<T>;
  constexpr bool F64 = sizeof(T) == 8;
  const float* meta = v.cloud_meta + 8 * b;
  const int dimx = __float_as_int(meta[4]), dimy = __float_as_int(meta[5]), dimz = __float_as_int(meta[6]);
  const bool table = (__float_as_int(meta[7]) & 256) != 0;
  const float4* cand = v.sorted_xyzi + (size_t)b * N;
  const T vox = A::voxel(q);
  T lo[3], hi[3], vmin[3];
  const T c3[3] = {mx, my, mz};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    lo[a] = A::lo(c3[a], q.full[a], vox);
    hi[a] = A::hi(c3[a], q.full[a], vox);
    vmin[a] = F64 ? (T)dmin[3 * b + a] : (T)meta[a];
  }
  // The reference only visits grid cells within n = (int)((full + 1) * 0.5) of the query's cell per axis (:247-267).
  // For odd boxes (every 3-tap filter: half-extent s + 0.5 voxels, n = s + 1) that window always covers the box, which
  // is why the 3x3x3 engines need no such test; for even boxes (half-extent == n voxels exactly) a point on the closed
  // edge whose cell index rounds one further out is NOT visited by the reference -- reproduced here with the
  // reference's own cell arithmetic, (int)((x - min) / r) (:192-194).
  const int ci[3] = {A::cell(mx, vmin[0], vox), A::cell(my, vmin[1], vox), A::cell(mz, vmin[2], vox)};
  const int nwin[3] = {(q.full[0] + 1) / 2, (q.full[1] + 1) / 2, (q.full[2] + 1) / 2};
  auto test = [&](const float4& c) -> int {
    T cx = (T)c.x, cy = (T)c.y, cz = (T)c.z;
    if (F64) {
      const T* pj = pts + 3 * (size_t)__float_as_int(c.w);
      cx = gen_ldg(pj); cy = gen_ldg(pj + 1); cz = gen_ldg(pj + 2);
    }
    if (abs(A::cell(cx, vmin[0], vox) - ci[0]) > nwin[0] || abs(A::cell(cy, vmin[1], vox) - ci[1]) > nwin[1] ||
        abs(A::cell(cz, vmin[2], vox) - ci[2]) > nwin[2])
      return -1;
    if (cx < lo[0] || cx > hi[0] || cy < lo[1] || cy > hi[1] || cz < lo[2] || cz > hi[2]) return -1;         // :277
    const int tx = A::tap(cx, lo[0], vox, q.full[0], q.sx);
    const int ty = A::tap(cy, lo[1], vox, q.full[1], q.sy);
    const int tz = A::tap(cz, lo[2], vox, q.full[2], q.sz);
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
  // Candidate window in the float grid, widened by a slop that covers the rounding of the box bounds and (T = double)
  // of the candidates' float copies (6e-8 relative, against 1e-6 here).
  const uint32_t* bins = v.cell_start + (size_t)b * ((size_t)v.cell_cap + 1);
  const float fx_ = (float)mx, fy_ = (float)my, fz_ = (float)mz;
  const float slop = q.voxel * (1.0f / 1024.0f) + fmaxf(fmaxf(fabsf(fx_), fabsf(fy_)), fabsf(fz_)) * 1e-6f;
  const int x0 = grid_coord((float)lo[0] - slop, meta[0], q.voxel, dimx), x1 = grid_coord((float)hi[0] + slop, meta[0], q.voxel, dimx);
  const int y0 = grid_coord((float)lo[1] - slop, meta[1], q.voxel, dimy), y1 = grid_coord((float)hi[1] + slop, meta[1], q.voxel, dimy);
  const int z0 = grid_coord((float)lo[2] - slop, meta[2], q.voxel, dimz), z1 = grid_coord((float)hi[2] + slop, meta[2], q.voxel, dimz);
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

// T = double, before the sort: the points rounded to float, and the per-cloud minimum in double (:156-171; the
// reference starts from 1e6f).
__global__ void k_generic_prepare_f64(const double* __restrict__ points, int N, float* __restrict__ points_f32,
                                      double* __restrict__ dmin) {
  const int b = blockIdx.x;
  const double* p = points + (size_t)b * N * 3;
  float* o = points_f32 + (size_t)b * N * 3;
  double m[3] = {1e6, 1e6, 1e6};
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double x = p[3 * (size_t)i + a];
      o[3 * (size_t)i + a] = (float)x;
      m[a] = fmin(m[a], x);
    }
  }
  __shared__ double red[3][GEN_WARPS];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    for (int o2 = 16; o2; o2 >>= 1) m[a] = fmin(m[a], __shfl_xor_sync(C3P_FULL_MASK, m[a], o2));
    if ((threadIdx.x & 31) == 0) red[a][threadIdx.x >> 5] = m[a];
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double r = red[threadIdx.x][0];
    for (int w = 1; w < GEN_WARPS; ++w) r = fmin(r, red[threadIdx.x][w]);
    dmin[3 * b + threadIdx.x] = r;
  }
}

// Count table + pair lists.  One warp per point (sorted order).  Two sweeps: counts, then the list.
template <typename T>
__global__ void __launch_bounds__(GEN_THREADS)
k_generic_search(GenGeom q, int B, int N, long long capacity, PlanView v, GenView gv, const T* __restrict__ points) {
  extern __shared__ int gen_cnt[];   // [GEN_WARPS][cells]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long qpos = (long long)blockIdx.x * GEN_WARPS + warp;
  if (qpos >= (long long)B * N) return;
  const int b = (int)(qpos / N);
  const float4 me = v.sorted_xyzi[qpos];
  const size_t row = (size_t)b * N + __float_as_int(me.w);
  const T* pts = points + (size_t)b * N * 3;
  T mx = (T)me.x, my = (T)me.y, mz = (T)me.z;
  if (sizeof(T) == 8) { mx = points[3 * row]; my = points[3 * row + 1]; mz = points[3 * row + 2]; }
  int* cnt = gen_cnt + warp * q.cells;
  for (int f = lane; f < q.cells; f += 32) cnt[f] = 0;
  __syncwarp();
  int found = 0;
  gen_sweep<T>(q, v, b, N, mx, my, mz, pts, gv.dmin, lane, [&](int j, int f) {
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
  gen_sweep<T>(q, v, b, N, mx, my, mz, pts, gv.dmin, lane, [&](int j, int f) {
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
template <typename T>
__global__ void __launch_bounds__(GEN_THREADS)
k_generic_forward(GenGeom q, long long pts, long long capacity, int Cin, int Cout, PlanView v, GenView gv,
                  const T* __restrict__ input, const T* __restrict__ filter, T* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * GEN_WARPS + (threadIdx.x >> 5);
  if (i >= pts) return;
  const long long begin = v.pair_begin[i];
  const int K = v.pair_len[i];
  const bool ok = begin + K <= capacity;
  for (int c0 = 0; c0 < Cout; c0 += 32) {
    const int c = c0 + lane;
    T acc = (T)0;
    if (ok && c < Cout) {
      for (int m = 0; m < K; ++m) {
        const int j = __ldg(v.pair_row + begin + m), f = __ldg(gv.pair_f + begin + m);
        const T inv = (T)1 / (T)__ldg(gv.count + (size_t)i * q.cells + f);
        const T* w = filter + (size_t)f * Cin * Cout + c;
        const T* x = input + (size_t)j * Cin;
        for (int k = 0; k < Cin; ++k) acc = gen_fma<T>(gen_ldg(w + (size_t)k * Cout), gen_ldg(x + k) * inv, acc);
      }
    }
    if (c < Cout) out[(size_t)i * Cout + c] = ok ? acc : gen_nan<T>();
  }
}

// Backward lists + grad_input.  One warp per j: for ii in N(j), f' = cell of j in ii's frame (no box test), dropped when
// it is a hole or count(ii, f') == 0 (:654-679); grad_in[j, k] += g[ii, c] * W[f', k, c] / count(ii, f')   (:692).
// The list keeps count(ii, f') itself (an int in the plan's weight slots); 1 / count is formed in T where it is used.
template <typename T>
__global__ void __launch_bounds__(GEN_THREADS)
k_generic_backward_input(GenGeom q, long long pts, long long capacity, int Cin, int Cout, PlanView v, GenView gv,
                         const T* __restrict__ points, const T* __restrict__ grad_out,
                         const T* __restrict__ filter, T* __restrict__ grad_in) {
  using A = GenAr<T>;
  const int lane = threadIdx.x & 31;
  const long long j = (long long)blockIdx.x * GEN_WARPS + (threadIdx.x >> 5);
  if (j >= pts) return;
  const long long begin = v.pair_begin[j];
  const int K = v.pair_len[j];
  const bool ok = begin + K <= capacity;
  const T vox = A::voxel(q);
  const T px = points[3 * j], py = points[3 * j + 1], pz = points[3 * j + 2];
  int* members_of = reinterpret_cast<int*>(v.bwd_weight);
  // pass 1 (lanes over pairs): the kept backward pairs, compacted in place
  int kept = 0;
  if (ok) {
    for (int m0 = 0; m0 < K; m0 += 32) {
      const int m = m0 + lane;
      int ii = 0, f = -1, members = 1;
      if (m < K) {
        ii = __ldg(v.pair_row + begin + m);
        const int tx = A::tap(px, A::lo(gen_ldg(points + 3 * (size_t)ii), q.full[0], vox), vox, q.full[0], q.sx);
        const int ty = A::tap(py, A::lo(gen_ldg(points + 3 * (size_t)ii + 1), q.full[1], vox), vox, q.full[1], q.sy);
        const int tz = A::tap(pz, A::lo(gen_ldg(points + 3 * (size_t)ii + 2), q.full[2], vox), vox, q.full[2], q.sz);
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
        members_of[slot] = members;
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
    T acc = (T)0;
    if (ok && k < Cin) {
      for (int m = 0; m < kept; ++m) {
        const int ii = v.bwd_row[begin + m], f = gv.bwd_f[begin + m];
        const T wgt = A::rcp(members_of[begin + m]);
        const T* w = filter + ((size_t)f * Cin + k) * Cout;
        const T* g = grad_out + (size_t)ii * Cout;
        T s = (T)0;
        for (int c = 0; c < Cout; ++c) s = gen_fma<T>(gen_ldg(g + c), gen_ldg(w + c), s);
        acc = gen_fma<T>(s, wgt, acc);
      }
    }
    if (k < Cin && grad_in) grad_in[(size_t)j * Cin + k] = ok ? acc : gen_nan<T>();
  }
}

// grad_filter[f', k, c] += g[ii, c] * in[j, k] / count(ii, f')   (:696).  CTA (b, s) walks chunk s of cloud b's points and
// their backward pairs in order; thread e owns elements e, e + blockDim, ... of the CTA's partial (fixed order of
// additions; the partials are summed in a fixed order afterwards).
template <typename T>
__global__ void __launch_bounds__(GEN_THREADS)
k_generic_backward_filter(GenGeom q, int N, int chunks, long long capacity, int Cin, int Cout, PlanView v, GenView gv,
                          const T* __restrict__ grad_out, const T* __restrict__ input) {
  using A = GenAr<T>;
  const int b = blockIdx.x / chunks, sc = blockIdx.x - b * chunks;
  const int KC = Cin * Cout;
  T* part = static_cast<T*>(gv.partial) + (size_t)blockIdx.x * q.cells * KC;
  const int* members_of = reinterpret_cast<const int*>(v.bwd_weight);
  for (int e = threadIdx.x; e < q.cells * KC; e += GEN_THREADS) part[e] = (T)0;
  __syncthreads();
  const int j_lo = (int)((long long)N * sc / chunks), j_hi = (int)((long long)N * (sc + 1) / chunks);
  for (int jj = j_lo; jj < j_hi; ++jj) {
    const size_t j = (size_t)b * N + jj;
    const long long begin = v.pair_begin[j];
    if (begin + v.pair_len[j] > capacity) continue;
    const int kept = v.bwd_count[j];
    for (int m = 0; m < kept; ++m) {
      const int ii = v.bwd_row[begin + m], f = gv.bwd_f[begin + m];
      const T wgt = A::rcp(members_of[begin + m]);
      T* pf = part + (size_t)f * KC;
      for (int e = threadIdx.x; e < KC; e += GEN_THREADS) {
        const int k = e / Cout, c = e - k * Cout;
        pf[e] = gen_fma<T>(gen_ldg(grad_out + (size_t)ii * Cout + c) * wgt, gen_ldg(input + j * Cin + k), pf[e]);
      }
    }
  }
}

// ordered sum of the per-cloud partials (double; the float path uses the library's k_reduce_partials)
__global__ void k_generic_reduce_f64(const double* __restrict__ partial, int S, long long nW, double* __restrict__ out,
                                     const long long* __restrict__ header) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nW) return;
  double s = 0.0;
  for (int i = 0; i < S; ++i) s += partial[(size_t)i * nW + w];
  if (header[H_OVERFLOW] != 0) s = gen_nan<double>();
  out[w] = s;
}

bool generic_filter_supported(const int dims_zyx[3]) {
  if (!dims_zyx) return false;
  for (int a = 0; a < 3; ++a)
    if (dims_zyx[a] < 1 || dims_zyx[a] > 64) return false;
  return (long long)dims_zyx[0] * dims_zyx[1] * dims_zyx[2] <= GEN_MAX_CELLS;
}

// Chunks per cloud of the weight-gradient kernel: about two CTAs per SM in total, at least 32 points per chunk, and at
// most 512 MB of partials.
static int gen_chunks(const conv3p_geom_t* g, int cells, int Cin, int Cout, size_t elem) {
  if (g->B <= 0 || g->N <= 0) return 1;
  long long s = (2LL * sm_count() + g->B - 1) / g->B;
  s = std::min<long long>(s, std::max(1, g->N / 32));
  const size_t one = elem * (size_t)g->B * cells * Cin * Cout;
  if (one > 0) s = std::min<long long>(s, (long long)((512ull << 20) / one));
  return (int)std::max<long long>(1, s);
}

static size_t gen_extra_bytes(const conv3p_geom_t* g, int cells, int Cin, int Cout, size_t elem) {
  const size_t pts = (size_t)g->B * g->N, cap = (size_t)g->pair_capacity;
  const size_t parts = (size_t)g->B * gen_chunks(g, cells, Cin, Cout, elem);
  size_t n = align_up(sizeof(int) * pts * cells) + 2 * align_up(sizeof(int) * cap) +
             align_up(elem * parts * cells * Cin * Cout) + 256;
  if (elem == 8) n += align_up(sizeof(float) * pts * 3) + align_up(sizeof(double) * (size_t)g->B * 3);
  return n;
}

size_t generic_workspace_bytes(const conv3p_geom_t* g, const int dims_zyx[3], int Cin, int Cout, int elem_bytes) {
  if (check_geom(g) || !generic_filter_supported(dims_zyx) || (elem_bytes != 4 && elem_bytes != 8)) return 0;
  const size_t plan = conv3p_plan_bytes(g);
  if (!plan) return 0;
  return plan + gen_extra_bytes(g, dims_zyx[0] * dims_zyx[1] * dims_zyx[2], Cin, Cout, (size_t)elem_bytes);
}

static GenView carve_gen(const conv3p_geom_t* g, int cells, int Cin, int Cout, void* base, size_t elem) {
  const size_t pts = (size_t)g->B * g->N, cap = (size_t)g->pair_capacity;
  char* p = static_cast<char*>(base);
  GenView gv;
  gv.count = reinterpret_cast<int*>(p); p += align_up(sizeof(int) * pts * cells);
  gv.pair_f = reinterpret_cast<int*>(p); p += align_up(sizeof(int) * cap);
  gv.bwd_f = reinterpret_cast<int*>(p); p += align_up(sizeof(int) * cap);
  gv.partial = p; p += align_up(elem * (size_t)g->B * gen_chunks(g, cells, Cin, Cout, elem) * cells * Cin * Cout);
  gv.points_f32 = nullptr;
  gv.dmin = nullptr;
  if (elem == 8) {
    gv.points_f32 = reinterpret_cast<float*>(p); p += align_up(sizeof(float) * pts * 3);
    gv.dmin = reinterpret_cast<double*>(p);
  }
  return gv;
}

// Sort + search into the workspace; returns the views.
template <typename T>
static int gen_plan(const conv3p_geom_t* g, const GenGeom& q, const T* points, int Cin, int Cout, void* ws,
                    size_t ws_bytes, cudaStream_t stream, PlanView* v, GenView* gv) {
  const size_t plan = conv3p_plan_bytes(g);
  if (!ws || ws_bytes < plan + gen_extra_bytes(g, q.cells, Cin, Cout, sizeof(T))) return CONV3P_ERR_BUFFER_TOO_SMALL;
  int st = make_view(g, ws, plan, v);
  if (st) return st;
  *gv = carve_gen(g, q.cells, Cin, Cout, static_cast<char*>(ws) + plan, sizeof(T));
  C3P_CUDA(cudaMemsetAsync(v->header, 0, sizeof(long long) * H_SLOTS, stream));
  const long long pts = (long long)g->B * g->N;
  if (pts == 0) return CONV3P_OK;
  if (!points) return CONV3P_ERR_INVALID_ARGUMENT;
  const float* sort_points = reinterpret_cast<const float*>(points);
  if (sizeof(T) == 8) {
    {
      LaunchTimer timer_("k_generic_prepare_f64", stream);
      k_generic_prepare_f64<<<g->B, GEN_THREADS, 0, stream>>>(reinterpret_cast<const double*>(points), g->N,
                                                               gv->points_f32, gv->dmin);
    }
    C3P_LAUNCH_CHECK("k_generic_prepare_f64");
    sort_points = gv->points_f32;
  }
  st = launch_cloud_sort(g, sort_points, *v, stream);
  if (st) return st;
  {
    LaunchTimer timer_("k_generic_search", stream);
    k_generic_search<T><<<(unsigned)((pts + GEN_WARPS - 1) / GEN_WARPS), GEN_THREADS, sizeof(int) * GEN_WARPS * q.cells, stream>>>(
        q, g->B, g->N, g->pair_capacity, *v, *gv, points);
  }
  C3P_LAUNCH_CHECK("k_generic_search");
  return CONV3P_OK;
}

template <typename T>
static int generic_forward_t(const conv3p_geom_t* g, const int dims_zyx[3], double voxel, const T* points, const T* input,
                             const T* filter, int Cin, int Cout, T* output, void* ws, size_t ws_bytes,
                             cudaStream_t stream) {
  if (!generic_filter_supported(dims_zyx)) return CONV3P_ERR_UNSUPPORTED;
  GenGeom q = make_gen(g, dims_zyx);
  if (sizeof(T) == 8) q.voxel_d = voxel;
  PlanView v;
  GenView gv;
  int st = gen_plan<T>(g, q, points, Cin, Cout, ws, ws_bytes, stream, &v, &gv);
  if (st) return st;
  const long long pts = (long long)g->B * g->N;
  if (pts == 0) return CONV3P_OK;
  if (!points || !input || !filter || !output) return CONV3P_ERR_INVALID_ARGUMENT;
  {
    LaunchTimer timer_("k_generic_forward", stream);
    k_generic_forward<T><<<(unsigned)((pts + GEN_WARPS - 1) / GEN_WARPS), GEN_THREADS, 0, stream>>>(
        q, pts, g->pair_capacity, Cin, Cout, v, gv, input, filter, output);
  }
  C3P_LAUNCH_CHECK("k_generic_forward");
  return CONV3P_OK;
}

template <typename T>
static int generic_backward_t(const conv3p_geom_t* g, const int dims_zyx[3], double voxel, const T* grad_out,
                              const T* points, const T* input, const T* filter, int Cin, int Cout, T* grad_input,
                              T* grad_filter, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (!generic_filter_supported(dims_zyx)) return CONV3P_ERR_UNSUPPORTED;
  GenGeom q = make_gen(g, dims_zyx);
  if (sizeof(T) == 8) q.voxel_d = voxel;
  PlanView v;
  GenView gv;
  int st = gen_plan<T>(g, q, points, Cin, Cout, ws, ws_bytes, stream, &v, &gv);
  if (st) return st;
  const long long pts = (long long)g->B * g->N;
  const long long nW = (long long)q.cells * Cin * Cout;
  if (pts == 0) {
    if (grad_filter) C3P_CUDA(cudaMemsetAsync(grad_filter, 0, sizeof(T) * nW, stream));
    return CONV3P_OK;
  }
  if (!points || !grad_out || !input || !filter) return CONV3P_ERR_INVALID_ARGUMENT;
  // grad_input kernel: also builds the backward lists the weight gradient walks (grad_input itself may be skipped)
  {
    LaunchTimer timer_("k_generic_backward_input", stream);
    k_generic_backward_input<T><<<(unsigned)((pts + GEN_WARPS - 1) / GEN_WARPS), GEN_THREADS, 0, stream>>>(
        q, pts, g->pair_capacity, Cin, Cout, v, gv, points, grad_out, filter, grad_input);
  }
  C3P_LAUNCH_CHECK("k_generic_backward_input");
  if (grad_filter) {
    const int chunks = gen_chunks(g, q.cells, Cin, Cout, sizeof(T));
    {
      LaunchTimer timer_("k_generic_backward_filter", stream);
      k_generic_backward_filter<T><<<g->B * chunks, GEN_THREADS, 0, stream>>>(q, g->N, chunks, g->pair_capacity, Cin, Cout,
                                                                               v, gv, grad_out, input);
    }
    C3P_LAUNCH_CHECK("k_generic_backward_filter");
    if (sizeof(T) == 8) {
      {
        LaunchTimer timer_("k_reduce_partials", stream);
        k_generic_reduce_f64<<<(unsigned)((nW + 255) / 256), 256, 0, stream>>>(
            static_cast<const double*>(gv.partial), g->B * chunks, nW, reinterpret_cast<double*>(grad_filter), v.header);
      }
      C3P_LAUNCH_CHECK("k_reduce_partials");
    } else {
      st = launch_reduce_partials(static_cast<const float*>(gv.partial), g->B * chunks, nW, reinterpret_cast<float*>(grad_filter),
                                  v.header, stream);
      if (st) return st;
    }
  }
  return CONV3P_OK;
}

int generic_forward(const conv3p_geom_t* g, const int dims_zyx[3], const float* points, const float* input,
                    const float* filter, int Cin, int Cout, float* output, void* ws, size_t ws_bytes,
                    cudaStream_t stream) {
  return generic_forward_t<float>(g, dims_zyx, 0.0, points, input, filter, Cin, Cout, output, ws, ws_bytes, stream);
}

int generic_backward(const conv3p_geom_t* g, const int dims_zyx[3], const float* grad_out, const float* points,
                     const float* input, const float* filter, int Cin, int Cout, float* grad_input,
                     float* grad_filter, void* ws, size_t ws_bytes, cudaStream_t stream) {
  return generic_backward_t<float>(g, dims_zyx, 0.0, grad_out, points, input, filter, Cin, Cout, grad_input, grad_filter,
                                   ws, ws_bytes, stream);
}

int generic_forward_f64(const conv3p_geom_t* g, const int dims_zyx[3], double voxel, const double* points,
                        const double* input, const double* filter, int Cin, int Cout, double* output, void* ws,
                        size_t ws_bytes, cudaStream_t stream) {
  return generic_forward_t<double>(g, dims_zyx, voxel, points, input, filter, Cin, Cout, output, ws, ws_bytes, stream);
}

int generic_backward_f64(const conv3p_geom_t* g, const int dims_zyx[3], double voxel, const double* grad_out,
                         const double* points, const double* input, const double* filter, int Cin, int Cout,
                         double* grad_input, double* grad_filter, void* ws, size_t ws_bytes, cudaStream_t stream) {
  return generic_backward_t<double>(g, dims_zyx, voxel, grad_out, points, input, filter, Cin, Cout, grad_input,
                                    grad_filter, ws, ws_bytes, stream);
}

}  // namespace c3p
