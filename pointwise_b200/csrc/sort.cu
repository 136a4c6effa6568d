// sort.cu -- stage 1 of the neighbour plan: per-cloud bounding box, voxel keys and a stable LSD
// radix sort of (key, point index), one CTA per cloud, digit histograms and warp-rank tables staged
// in shared memory.  Replaces the reference's host-side counting sort (Grid::Grid,
// tf_conv3p_atrous.cpp:157-227) and the O(N^2) sweeps of its GPU op (tf_conv3p_atrous.cu:252-327).
//
// The grid is the reference's: cell size = voxel size, cell = (int)((x - vmin) / voxel) per axis,
// dims = (int)((vmax - vmin) / voxel) + 2.  Any monotone cell function would do -- the search applies
// the exact predicate to a superset of candidates -- so dims are clamped to 1024 per axis (a 30-bit
// key); clouds wider than 1023 voxels merely share their last cell.
#include "common.cuh"
#include "radix_sort.cuh"

namespace c3p {

constexpr int MAX_DIM = 1024;

__global__ void __launch_bounds__(SORT_THREADS)
k_cloud_sort(const float* __restrict__ points, int N, float voxel, float* __restrict__ cloud_meta,
             uint32_t* __restrict__ sorted_key, float4* __restrict__ sorted_xyzi,
             uint32_t* __restrict__ tmp, uint32_t* __restrict__ cell_start, int cell_cap) {
  __shared__ float red[6][SORT_WARPS];
  __shared__ float box[6];
  __shared__ uint32_t base[256];
  __shared__ uint32_t wsum[8];
  __shared__ uint16_t wcount[SORT_WARPS][256];
  __shared__ uint16_t wpre[SORT_WARPS][256];

  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* P = points + (size_t)b * N * 3;
  if (N == 0) return;

  // ---- bounding box (tf_conv3p_atrous.cpp:163-177: starts at +-1e6, std::min / std::max) --------
  float mn[3] = {1e6f, 1e6f, 1e6f}, mx[3] = {-1e6f, -1e6f, -1e6f};
  for (int i = tid; i < N; i += SORT_THREADS) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float v = P[3 * i + a];
      mn[a] = v < mn[a] ? v : mn[a];
      mx[a] = mx[a] < v ? v : mx[a];
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    for (int o = 16; o > 0; o >>= 1) {
      float u = __shfl_xor_sync(C3P_FULL_MASK, mn[a], o);
      float w = __shfl_xor_sync(C3P_FULL_MASK, mx[a], o);
      mn[a] = u < mn[a] ? u : mn[a];
      mx[a] = mx[a] < w ? w : mx[a];
    }
    if (lane == 0) {
      red[a][warp] = mn[a];
      red[3 + a][warp] = mx[a];
    }
  }
  __syncthreads();
  if (tid < 6) {
    float r = red[tid][0];
    for (int w = 1; w < SORT_WARPS; ++w) {
      float u = red[tid][w];
      r = (tid < 3) ? (u < r ? u : r) : (r < u ? u : r);
    }
    box[tid] = r;
  }
  __syncthreads();
  const float vminx = box[0], vminy = box[1], vminz = box[2];
  int dim[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {  // :179-181
    int d = __float2int_rz(__fdiv_rn(__fsub_rn(box[3 + a], box[a]), voxel)) + 2;
    dim[a] = max(1, min(d, MAX_DIM));
  }
  const uint32_t cells = (uint32_t)dim[0] * dim[1] * dim[2];
  const int nbits = cells > 1 ? 32 - __clz(cells - 1) : 0;
  const bool has_table = cells <= (uint32_t)cell_cap;   // bin-offset table below; else the search bisects
  if (tid == 0) {
    float* m = cloud_meta + 8 * b;
    m[0] = vminx; m[1] = vminy; m[2] = vminz; m[3] = voxel;
    m[4] = __int_as_float(dim[0]); m[5] = __int_as_float(dim[1]); m[6] = __int_as_float(dim[2]);
    m[7] = __int_as_float(nbits | (has_table ? 256 : 0));
  }

  // ---- voxel keys ---------------------------------------------------------------------------------
  uint32_t* kin = tmp + (size_t)b * 4 * N;
  uint32_t* iin = kin + N;
  uint32_t* kout = kin + 2 * (size_t)N;
  uint32_t* iout = kin + 3 * (size_t)N;
  for (int i = tid; i < N; i += SORT_THREADS) {
    int cx = grid_coord(P[3 * i + 0], vminx, voxel, dim[0]);
    int cy = grid_coord(P[3 * i + 1], vminy, voxel, dim[1]);
    int cz = grid_coord(P[3 * i + 2], vminz, voxel, dim[2]);
    kin[i] = (uint32_t)((cz * dim[1] + cy) * dim[0] + cx);
    iin[i] = (uint32_t)i;
  }
  for (int e = tid; e < SORT_WARPS * 256; e += SORT_THREADS) (&wcount[0][0])[e] = 0;
  __syncthreads();

  // ---- stable LSD radix sort, 8 bits per pass -----------------------------------------------------
  RadixTables tb{base, wsum, wcount, wpre};
  radix_sort_pairs(kin, iin, kout, iout, N, 0, nbits, tb);

  // ---- bin offsets: start[c] = first sorted position whose key is >= c, for c in [0, cells] ---------------
  // (the reference's cell_start scan, tf_conv3p_atrous.cpp:203-214, as a lower-bound table: a run of grid cells
  // [c0, c1] of one grid row is the contiguous range [start[c0], start[c1 + 1]) of the sorted cloud)
  if (has_table) {
    uint32_t* start = cell_start + (size_t)b * ((size_t)cell_cap + 1);
    for (uint32_t c = tid; c <= cells; c += SORT_THREADS) start[c] = (uint32_t)N;
    __syncthreads();
    for (int s = tid; s < N; s += SORT_THREADS) {
      const uint32_t k = kin[s];
      if (s == 0 || kin[s - 1] != k) start[k] = (uint32_t)s;   // first position of every occupied cell
    }
    __syncthreads();
    // empty cells take the start of the next occupied one: a suffix minimum (occupied starts increase with c)
    const uint32_t per = (cells + 1 + SORT_THREADS - 1) / SORT_THREADS;
    const uint32_t c_lo = min((uint32_t)tid * per, cells + 1), c_hi = min(c_lo + per, cells + 1);
    uint32_t m = (uint32_t)N;
    for (uint32_t c = c_hi; c > c_lo; --c) m = min(m, start[c - 1]);
    // suffix minimum over the threads' chunk minima (reuses the digit table of the sort)
    uint32_t v = m;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_down_sync(C3P_FULL_MASK, v, o);
      if (lane + o < 32) v = min(v, u);
    }
    if (lane == 0) base[warp] = v;          // minimum of the warp's chunks
    __syncthreads();
    uint32_t later = (uint32_t)N;           // minimum over all chunks after this thread's
    for (int w = warp + 1; w < SORT_WARPS; ++w) later = min(later, base[w]);
    const uint32_t nxt = __shfl_down_sync(C3P_FULL_MASK, v, 1);   // suffix minimum starting at the next lane
    if (lane < 31) later = min(later, nxt);
    uint32_t run = later;
    for (uint32_t c = c_hi; c > c_lo; --c) {
      run = min(run, start[c - 1]);
      start[c - 1] = run;
    }
  }

  // ---- emit sorted keys and (x, y, z, index) ------------------------------------------------------
  for (int s = tid; s < N; s += SORT_THREADS) {
    uint32_t idx = iin[s];
    sorted_key[(size_t)b * N + s] = kin[s];
    sorted_xyzi[(size_t)b * N + s] =
        make_float4(P[3 * idx], P[3 * idx + 1], P[3 * idx + 2], __int_as_float((int)idx));
  }
}

int launch_cloud_sort(const conv3p_geom_t* g, const float* points, const PlanView& v,
                      cudaStream_t stream) {
  if (g->B == 0 || g->N == 0) return CONV3P_OK;
  {
    LaunchTimer timer_("k_cloud_sort", stream);
    k_cloud_sort<<<g->B, SORT_THREADS, 0, stream>>>(points, g->N, g->voxel_size, v.cloud_meta,
                                                 v.sorted_key, v.sorted_xyzi, v.sort_tmp, v.cell_start, v.cell_cap);
  }
  C3P_LAUNCH_CHECK("k_cloud_sort");
  return CONV3P_OK;
}

}  // namespace c3p
