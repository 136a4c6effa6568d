// small_channels.cu -- fp32 SIMT engine for the reference models' own layer shapes (3->9, 9->9, 36->13:
// pointcnn2_acsd.py:48-67, scene_seg/pointcnn_scene_seg_acsd.py:51-57), where a cell's weight matrix is far
// too small for an MMA tile.  One warp per point (voxel-sorted order), the whole filter resident in shared
// memory, persistent CTAs:
//
//   forward / grad_input : for every non-empty cell of the point, lanes = channels gather the cell's rows
//                          (mean, or weighted sum over the backward list) and lanes = outputs accumulate
//                          out[n] += sum_k A[k] * W_f[k, n]   (tf_conv3p_atrous.cpp:480-494 / :682-692)
//   grad_filter          : the gathered G_f[j, :] is multiplied with input[j, :] and added into a per-warp-
//                          private copy of grad_filter in shared memory (no atomics, fixed order); the
//                          copies and then the CTA partials are reduced in a fixed order
//                          (tf_conv3p_atrous.cpp:694-716).  Used when 8 copies fit in shared memory.
#include "common.cuh"
#include "tc_common.cuh"

namespace c3p {

constexpr int SC_WARPS = 8;
constexpr int SC_THREADS = SC_WARPS * 32;
constexpr int SC_MAXC = 64;  // widest channel count served (two values per lane)

struct SCArgs {
  const float* src;        // gathered rows [B*N, Csrc]
  const float* filter;     // [27, Cin, Cout]
  float* out;              // [B*N, Nout]
  const int* cnt;          // [B*N, 27]
  const long long* begin;
  const int* len;
  const int* rows;
  const float* weights;    // nullptr: per-cell mean
  const float4* sorted_xyzi;
  long long total_points, capacity;
  int N, Cin, Cout, Csrc, Nout;
  long long src_stride, out_stride;  // floats between rows
  int activation;
  int transposed;          // 0: B[k][n] = W[f][k][n]; 1: B[k][n] = W[f][n][k]
};

// out[p, n] = sum_f sum_k A_f[p, k] * B_f[k, n]
__global__ void __launch_bounds__(SC_THREADS) k_small_gather_contract(const SCArgs a) {
  extern __shared__ float sc_w[];  // [27][Csrc][Nout] (B_f, contraction index major)
  const int Csrc = a.Csrc, Nout = a.Nout;
  for (int e = threadIdx.x; e < C3P_NCELL * Csrc * Nout; e += SC_THREADS) {
    const int n = e % Nout, k = (e / Nout) % Csrc, f = e / (Nout * Csrc);
    sc_w[e] = a.transposed ? a.filter[((size_t)f * a.Cin + n) * a.Cout + k]
                           : a.filter[((size_t)f * a.Cin + k) * a.Cout + n];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long stride = (long long)gridDim.x * SC_WARPS;
  for (long long s = (long long)blockIdx.x * SC_WARPS + warp; s < a.total_points; s += stride) {
    const int b = (int)(s / a.N);
    const int row = b * a.N + __float_as_int(a.sorted_xyzi[s].w);
    const long long bg = a.begin[row];
    const bool ok = bg + a.len[row] <= a.capacity;
    const int mine = (ok && lane < C3P_NCELL) ? __ldg(a.cnt + (size_t)row * C3P_NCELL + lane) : 0;
    float acc0 = 0.f, acc1 = 0.f;  // outputs n = lane, lane + 32
    long long at = bg;
    int excl = mine;               // list position of cell f = exclusive prefix of the counts
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(C3P_FULL_MASK, excl, o);
      if (lane >= o) excl += u;
    }
    excl -= mine;
    unsigned cells = __ballot_sync(C3P_FULL_MASK, mine > 0);
    while (cells) {
      const int f = __ffs(cells) - 1;
      cells &= cells - 1;
      const int before = __shfl_sync(C3P_FULL_MASK, excl, f);
      const int n = __shfl_sync(C3P_FULL_MASK, mine, f);
      const int* r = a.rows + at + before;
      float a0 = 0.f, a1 = 0.f;  // aggregate channels k = lane, lane + 32
      int m = 0;
      for (; m + 2 <= n; m += 2) {  // two members in flight
        const int j0 = __ldg(r + m), j1 = __ldg(r + m + 1);
        const float w0 = a.weights ? __ldg(a.weights + at + before + m) : 1.f;
        const float w1 = a.weights ? __ldg(a.weights + at + before + m + 1) : 1.f;
        const float* p0 = a.src + (size_t)j0 * a.src_stride;
        const float* p1 = a.src + (size_t)j1 * a.src_stride;
        const float x00 = lane < Csrc ? __ldg(p0 + lane) : 0.f, x10 = lane < Csrc ? __ldg(p1 + lane) : 0.f;
        const float x01 = lane + 32 < Csrc ? __ldg(p0 + lane + 32) : 0.f;
        const float x11 = lane + 32 < Csrc ? __ldg(p1 + lane + 32) : 0.f;
        a0 = fmaf(w0, x00, a0); a0 = fmaf(w1, x10, a0);
        a1 = fmaf(w0, x01, a1); a1 = fmaf(w1, x11, a1);
      }
      if (m < n) {
        const int j0 = __ldg(r + m);
        const float w0 = a.weights ? __ldg(a.weights + at + before + m) : 1.f;
        const float* p0 = a.src + (size_t)j0 * a.src_stride;
        a0 = fmaf(w0, lane < Csrc ? __ldg(p0 + lane) : 0.f, a0);
        a1 = fmaf(w0, lane + 32 < Csrc ? __ldg(p0 + lane + 32) : 0.f, a1);
      }
      if (!a.weights && n > 1) {
        const float inv = __fdiv_rn(1.f, (float)n);
        a0 *= inv;
        a1 *= inv;
      }
      const float* wf = sc_w + (size_t)f * Csrc * Nout;
      for (int k = 0; k < Csrc; ++k) {
        const float ak = __shfl_sync(C3P_FULL_MASK, k < 32 ? a0 : a1, k & 31);
        if (lane < Nout) acc0 = fmaf(ak, wf[k * Nout + lane], acc0);
        if (lane + 32 < Nout) acc1 = fmaf(ak, wf[k * Nout + lane + 32], acc1);
      }
    }
    const float nanv = __int_as_float(0x7fc00000);
    if (lane < Nout) a.out[(size_t)row * a.out_stride + lane] = ok ? apply_activation(acc0, a.activation) : nanv;
    if (lane + 32 < Nout)
      a.out[(size_t)row * a.out_stride + lane + 32] = ok ? apply_activation(acc1, a.activation) : nanv;
  }
}

// Second version of the same operator for Csrc, Nout <= 16 -- the shapes of every 9-channel layer of the two
// reference networks.  The first version walks a point's list cell by cell with lanes = channels (9 of 32 lanes
// busy) and two dependent loads (id, row) per pair of members: ~22 exposed L2 round trips per point.  Here
//   * lanes = (member slot, channel): G = 32 / Csrc slots walk the point's WHOLE list (all cells, it is grouped by
//     ascending cell) in strides of G, four members in flight per slot; the cell of a member is found by a pointer
//     that only moves forward; a slot's sum for a cell is written once to the slot's own copy A_g[cell][channel]
//     in shared memory (no atomics, fixed order of additions);
//   * the contraction runs over the non-empty cells only with lanes = (k group, output): KG = 32 / Nout groups share
//     the (cell, k) pairs and are summed by shuffles in a fixed order.
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// Lists of one point and the rows they gather (forward lists + input, or backward lists + grad_out).
struct SC2Lists {
  const float* src;
  const int* cnt;
  const long long* begin;
  const int* len;
  const int* rows;
  const float* weights;    // nullptr: per-cell mean
  long long src_stride, capacity;
};

// Phase 1 of the second-version kernels, by one warp for the point in `row`: A[f][ch] = sum over the members of
// cell f of w * src[member][ch] (w = 1 / members, or the list's weights) for the cells f0 <= f < f1, in copy 0 of
// the warp's A area.  Csrc <= 32: G = 32 / Csrc member slots, one channel per lane; Csrc <= 64: one slot, two
// channels per lane.  Returns the mask of non-empty cells of the range; `ok` = the point's list is complete.
__device__ __forceinline__ unsigned sc2_aggregate(const SC2Lists& p, int row, int lane, int Csrc, int G, int f0, int f1,
                                                  uint32_t s_A, uint32_t s_pre, uint32_t s_inv, bool& ok) {
  const int cell_words = C3P_NCELL * Csrc;
  const int g = lane / Csrc, ch = lane - g * Csrc;     // (Csrc > 32: g == 0, ch == lane)
  const bool two = Csrc > 32, second = two && lane + 32 < Csrc;
  const long long bg = p.begin[row];
  ok = bg + p.len[row] <= p.capacity;
  const int mine = (ok && lane < C3P_NCELL) ? __ldg(p.cnt + (size_t)row * C3P_NCELL + lane) : 0;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(C3P_FULL_MASK, incl, o);
    if (lane >= o) incl += u;
  }
  const int total = __shfl_sync(C3P_FULL_MASK, incl, 31);
  const unsigned range = (f1 >= 32 ? ~0u : ((1u << f1) - 1u)) & ~((1u << f0) - 1u);
  const unsigned cells = __ballot_sync(C3P_FULL_MASK, mine > 0) & range;
  __syncwarp();
  if (lane < C3P_NCELL) {
    sts_f32(s_pre + 4u * lane, __int_as_float(incl - mine));
    sts_f32(s_inv + 4u * lane, mine > 0 ? __fdiv_rn(1.f, (float)mine) : 0.f);
  }
  if (lane == 31) sts_f32(s_pre + 4u * C3P_NCELL, __int_as_float(total));
  for (int gg = 0; gg < G; ++gg)   // the range's part of every slot's copy
    for (int e = f0 * Csrc + lane; e < f1 * Csrc; e += 32) sts_f32(s_A + 4u * (uint32_t)(gg * cell_words + e), 0.f);
  __syncwarp();
  if (!cells) return 0u;
  const int Kbeg = __float_as_int(lds_f32(s_pre + 4u * (uint32_t)f0)), K = __float_as_int(lds_f32(s_pre + 4u * (uint32_t)f1));
  if (g < G) {
    const uint32_t Ag = s_A + 4u * (uint32_t)(g * cell_words + ch);
    const int* rlist = p.rows + bg;
    const float* wlist = p.weights ? p.weights + bg : nullptr;
    int cf = f0;                                         // cell of the member being accumulated
    int next_start = __float_as_int(lds_f32(s_pre + 4u * (uint32_t)(f0 + 1)));   // pre[cf + 1]
    float wc = lds_f32(s_inv + 4u * (uint32_t)f0);       // 1 / members of cell cf
    float acc = 0.f, acc2 = 0.f;
    for (int e0 = Kbeg + g; e0 < K; e0 += 4 * G) {
      int j[4];
      float w[4], x[4], x2[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * G;
        j[u] = e < K ? __ldg(rlist + e) : 0;
        w[u] = (wlist && e < K) ? __ldg(wlist + e) : 1.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool in = e0 + u * G < K;
        const float* r = p.src + (size_t)j[u] * p.src_stride + ch;
        x[u] = in ? __ldg(r) : 0.f;
        x2[u] = (in && second) ? __ldg(r + 32) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * G;
        if (e < K) {
          if (e >= next_start) {                         // the member starts a later cell: close the current one
            sts_f32(Ag + 4u * (uint32_t)(cf * Csrc), acc);
            if (second) sts_f32(Ag + 4u * (uint32_t)(cf * Csrc + 32), acc2);
            acc = 0.f;
            acc2 = 0.f;
            do {
              ++cf;
              next_start = __float_as_int(lds_f32(s_pre + 4u * (uint32_t)(cf + 1)));
            } while (e >= next_start);
            wc = lds_f32(s_inv + 4u * (uint32_t)cf);
          }
          const float ww = wlist ? w[u] : wc;
          acc = fmaf(ww, x[u], acc);
          acc2 = fmaf(ww, x2[u], acc2);
        }
      }
    }
    if (Kbeg + g < K) {
      sts_f32(Ag + 4u * (uint32_t)(cf * Csrc), acc);
      if (second) sts_f32(Ag + 4u * (uint32_t)(cf * Csrc + 32), acc2);
    }
  }
  __syncwarp();
  // fold the G copies into copy 0, fixed order
  if (G > 1) {
    for (int e = f0 * Csrc + lane; e < f1 * Csrc; e += 32) {
      float v = lds_f32(s_A + 4u * e);
      for (int gg = 1; gg < G; ++gg) v += lds_f32(s_A + 4u * (uint32_t)(gg * cell_words + e));
      sts_f32(s_A + 4u * e, v);
    }
    __syncwarp();
  }
  return cells;
}

constexpr int SC2_MAXC = 64;                               // widest channel count of the second version
// per warp: A copies (G * 27 * Csrc <= 27 * max(32, Csrc) floats) | cell starts (32 ints) | 1 / members (32 floats)
__host__ __device__ inline int sc2_a_words(int Csrc) { return C3P_NCELL * (Csrc > 32 ? Csrc : 32); }
__host__ __device__ inline int sc2_warp_words(int Csrc) { return sc2_a_words(Csrc) + 64; }

__device__ __forceinline__ int sc2_slots(int C) { return C > 32 ? 1 : 32 / C; }

__global__ void __launch_bounds__(SC_THREADS) k_small_gather_contract2(const SCArgs a) {
  extern __shared__ float sc2[];
  const int Csrc = a.Csrc, Nout = a.Nout;
  const int nWf = C3P_NCELL * Csrc * Nout;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int G = sc2_slots(Csrc), KG = sc2_slots(Nout);
  // shared-window addresses, converted once (tc_common.cuh): filter | per warp: A copies, cell starts, 1 / members
  const uint32_t s_w = tc::smem_u32_once(sc2);                                   // [27][Csrc][Nout]
  const uint32_t s_A = s_w + 4u * (uint32_t)(((nWf + 3) & ~3) + warp * sc2_warp_words(Csrc));   // [G][27][Csrc]
  const uint32_t s_pre = s_A + 4u * (uint32_t)sc2_a_words(Csrc);                 // [28] ints, pre[27] = K
  const uint32_t s_inv = s_pre + 4u * 32;                                        // [27]
  for (int e = threadIdx.x; e < nWf; e += SC_THREADS) {
    const int n = e % Nout, k = (e / Nout) % Csrc, f = e / (Nout * Csrc);
    sc2[e] = a.transposed ? a.filter[((size_t)f * a.Cin + n) * a.Cout + k]
                          : a.filter[((size_t)f * a.Cin + k) * a.Cout + n];
  }
  __syncthreads();
  // outputs: Nout <= 32: lanes = (k group, n); Nout <= 64: lanes = n and n + 32, one k group
  const int kg = lane / Nout, nn = lane - kg * Nout;
  const bool out_on = kg < KG;
  const bool out2 = Nout > 32 && lane + 32 < Nout;
  SC2Lists L;
  L.src = a.src; L.cnt = a.cnt; L.begin = a.begin; L.len = a.len; L.rows = a.rows; L.weights = a.weights;
  L.src_stride = a.src_stride; L.capacity = a.capacity;
  const long long stride = (long long)gridDim.x * SC_WARPS;
  for (long long s = (long long)blockIdx.x * SC_WARPS + warp; s < a.total_points; s += stride) {
    const int b = (int)(s / a.N);
    const int row = b * a.N + __float_as_int(a.sorted_xyzi[s].w);
    bool ok;
    const unsigned cells = sc2_aggregate(L, row, lane, Csrc, G, 0, C3P_NCELL, s_A, s_pre, s_inv, ok);
    // ---- phase 2: out[n] = sum over non-empty cells f, channels k of A[f][k] * W_f[k][n] ------------------------------
    float o = 0.f, o2 = 0.f;
    if (out_on) {
      for (unsigned todo = cells; todo; todo &= todo - 1) {
        const int f = __ffs(todo) - 1;
        const uint32_t af = s_A + 4u * (uint32_t)(f * Csrc), wf = s_w + 4u * (uint32_t)(f * Csrc * Nout + nn);
        for (int k = kg; k < Csrc; k += KG) {
          const float av = lds_f32(af + 4u * k);
          o = fmaf(av, lds_f32(wf + 4u * (uint32_t)(k * Nout)), o);
          if (out2) o2 = fmaf(av, lds_f32(wf + 4u * (uint32_t)(k * Nout + 32)), o2);
        }
      }
    }
    // sum the KG partial sums in a fixed order (lanes nn, nn + Nout, nn + 2 Nout, ...)
    float total = o;
    if (KG > 1) {
      total = 0.f;
      for (int q = 0; q < KG; ++q) total += __shfl_sync(C3P_FULL_MASK, o, (nn + q * Nout) & 31);
    }
    const float nanv = __int_as_float(0x7fc00000);
    if (lane < Nout) a.out[(size_t)row * a.out_stride + lane] = ok ? apply_activation(total, a.activation) : nanv;
    if (out2) a.out[(size_t)row * a.out_stride + lane + 32] = ok ? apply_activation(o2, a.activation) : nanv;
  }
}

struct SFArgs {
  const float* grad_out;
  const float* input;
  const int* cnt;
  const long long* begin;
  const int* len;
  const int* rows;
  const float* weights;
  const float4* sorted_xyzi;
  float* partial;  // [gridDim.x][27*Cin*Cout]
  long long total_points, capacity;
  int N, Cin, Cout;
};

__global__ void __launch_bounds__(SC_THREADS) k_small_backward_filter(const SFArgs a) {
  extern __shared__ float sf_gw[];  // [warps][27][Cin][Cout]: one private copy per warp
  const int Cin = a.Cin, Cout = a.Cout;
  const int nW = C3P_NCELL * Cin * Cout;
  const int copies = SC_WARPS;
  for (int e = threadIdx.x; e < copies * nW; e += SC_THREADS) sf_gw[e] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* gw = sf_gw + (size_t)warp * nW;
  const int slots = Cout <= 32 ? 32 / Cout : 1;
  const int slot = Cout <= 32 ? lane / Cout : 0, kk = Cout <= 32 ? lane - slot * Cout : lane;
  const bool slot_ok = slot < slots;
  const long long stride = (long long)gridDim.x * SC_WARPS;
  for (long long s = (long long)blockIdx.x * SC_WARPS + warp; s < a.total_points; s += stride) {
    const int b = (int)(s / a.N);
    const int row = b * a.N + __float_as_int(a.sorted_xyzi[s].w);
    const long long bg = a.begin[row];
    if (bg + a.len[row] > a.capacity) continue;
    const int mine = lane < C3P_NCELL ? __ldg(a.cnt + (size_t)row * C3P_NCELL + lane) : 0;
    const float x0 = lane < Cin ? __ldg(a.input + (size_t)row * Cin + lane) : 0.f;
    const float x1 = lane + 32 < Cin ? __ldg(a.input + (size_t)row * Cin + lane + 32) : 0.f;
    int excl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(C3P_FULL_MASK, excl, o);
      if (lane >= o) excl += u;
    }
    excl -= mine;
    unsigned cells = __ballot_sync(C3P_FULL_MASK, mine > 0);
    while (cells) {
      const int f = __ffs(cells) - 1;
      cells &= cells - 1;
      const int before = __shfl_sync(C3P_FULL_MASK, excl, f);
      const int n = __shfl_sync(C3P_FULL_MASK, mine, f);
      float g0 = 0.f, g1 = 0.f;  // G_f[j, c], c = lane, lane + 32 (after the slot reduction)
      const int* r = a.rows + bg + before;
      const float* wl = a.weights + bg + before;
      for (int m = 0; m < n; m += 2 * slots) {
        const int m0 = m + slot, m1 = m + slots + slot;
        const bool v0 = slot_ok && m0 < n, v1 = slot_ok && m1 < n;
        const int j0 = v0 ? __ldg(r + m0) : 0, j1 = v1 ? __ldg(r + m1) : 0;
        const float w0 = v0 ? __ldg(wl + m0) : 0.f, w1 = v1 ? __ldg(wl + m1) : 0.f;
        const float* p0 = a.grad_out + (size_t)j0 * Cout;
        const float* p1 = a.grad_out + (size_t)j1 * Cout;
        g0 = fmaf(w0, v0 ? __ldg(p0 + kk) : 0.f, g0);
        g0 = fmaf(w1, v1 ? __ldg(p1 + kk) : 0.f, g0);
        if (Cout > 32) {
          g1 = fmaf(w0, (v0 && lane + 32 < Cout) ? __ldg(p0 + lane + 32) : 0.f, g1);
          g1 = fmaf(w1, (v1 && lane + 32 < Cout) ? __ldg(p1 + lane + 32) : 0.f, g1);
        }
      }
      for (int sl = 1; sl < slots; ++sl) {
        const float o = __shfl_sync(C3P_FULL_MASK, g0, (lane + sl * Cout) & 31);
        if (lane < Cout) g0 += o;
      }
      float* gf = gw + (size_t)f * Cin * Cout;
      for (int k = 0; k < Cin; ++k) {
        const float xk = __shfl_sync(C3P_FULL_MASK, k < 32 ? x0 : x1, k & 31);
        // the warp owns this copy: plain read-modify-write, fixed order
        if (lane < Cout) gf[k * Cout + lane] = fmaf(xk, g0, gf[k * Cout + lane]);
        if (lane + 32 < Cout) gf[k * Cout + lane + 32] = fmaf(xk, g1, gf[k * Cout + lane + 32]);
      }
    }
  }
  __syncthreads();
  float* dst = a.partial + (size_t)blockIdx.x * nW;
  for (int e = threadIdx.x; e < nW; e += SC_THREADS) {
    float sum = 0.f;
    for (int c = 0; c < copies; ++c) sum += sf_gw[(size_t)c * nW + e];  // fixed order over the warps
    dst[e] = sum;
  }
}

// Second version of the weight gradient: the aggregate G_f[j, :] of every non-empty cell comes from sc2_aggregate
// (whole list in flight, see above), then lanes = (k, c) pairs add the rank-1 update x[j, k] * G_f[j, c] into the
// warp's private copy of grad_filter (plain read-modify-write, fixed order).  When 8 private copies of all 27 cells do
// not fit in shared memory (36 x 13: 50 KB each) the cells are covered in passes of `cells_per_pass`; a pass walks
// only its own cells' part of every list, so nothing is gathered twice.
__global__ void __launch_bounds__(SC_THREADS) k_small_backward_filter2(const SFArgs a, int cells_per_pass) {
  extern __shared__ float sf2[];   // [warps][cells_per_pass][Cin][Cout] private copies | per warp: A copies, cell starts, 1 / members
  const int Cin = a.Cin, Cout = a.Cout;
  const int pairs = Cin * Cout;
  const int nWp = cells_per_pass * pairs;              // one private copy
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int G = sc2_slots(Cout);
  const uint32_t s_all = tc::smem_u32_once(sf2);
  const uint32_t s_gw = s_all + 4u * (uint32_t)(warp * nWp);
  const uint32_t s_A = s_all + 4u * (uint32_t)(((SC_WARPS * nWp + 3) & ~3) + warp * sc2_warp_words(Cout));
  const uint32_t s_pre = s_A + 4u * (uint32_t)sc2_a_words(Cout), s_inv = s_pre + 4u * 32;
  SC2Lists L;
  L.src = a.grad_out; L.cnt = a.cnt; L.begin = a.begin; L.len = a.len; L.rows = a.rows; L.weights = a.weights;
  L.src_stride = Cout; L.capacity = a.capacity;
  const long long stride = (long long)gridDim.x * SC_WARPS;
  float* dst = a.partial + (size_t)blockIdx.x * C3P_NCELL * pairs;
  for (int f0 = 0; f0 < C3P_NCELL; f0 += cells_per_pass) {
    const int f1 = min(C3P_NCELL, f0 + cells_per_pass);
    for (int e = threadIdx.x; e < SC_WARPS * nWp; e += SC_THREADS) sf2[e] = 0.f;
    __syncthreads();
    for (long long s = (long long)blockIdx.x * SC_WARPS + warp; s < a.total_points; s += stride) {
      const int b = (int)(s / a.N);
      const int row = b * a.N + __float_as_int(a.sorted_xyzi[s].w);
      const float x0 = lane < Cin ? __ldg(a.input + (size_t)row * Cin + lane) : 0.f;
      const float x1 = lane + 32 < Cin ? __ldg(a.input + (size_t)row * Cin + lane + 32) : 0.f;
      bool ok;
      const unsigned cells = sc2_aggregate(L, row, lane, Cout, G, f0, f1, s_A, s_pre, s_inv, ok);
      if (!cells) continue;   // (an incomplete list has no cells either)
      for (int p0 = 0; p0 < pairs; p0 += 32) {
        const int p = p0 + lane;
        const int k = p / Cout, c = p - k * Cout;
        const float xa = __shfl_sync(C3P_FULL_MASK, x0, k & 31), xb = __shfl_sync(C3P_FULL_MASK, x1, k & 31);
        const float xk = k < 32 ? xa : xb;
        if (p < pairs) {
          for (unsigned todo = cells; todo; todo &= todo - 1) {
            const int f = __ffs(todo) - 1;
            const uint32_t at = s_gw + 4u * (uint32_t)((f - f0) * pairs + p);
            sts_f32(at, fmaf(xk, lds_f32(s_A + 4u * (uint32_t)(f * Cout + c)), lds_f32(at)));
          }
        }
      }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < (f1 - f0) * pairs; e += SC_THREADS) {
      float sum = 0.f;
      for (int c = 0; c < SC_WARPS; ++c) sum += sf2[(size_t)c * nWp + e];  // fixed order over the warps
      dst[(size_t)f0 * pairs + e] = sum;
    }
    __syncthreads();
  }
}

static int sc_sms() { return sm_count(); }

// Measured on B200 (profiles/r1_summary.md): the warp-per-point gather-contract wins up to ~16 channels (2x at
// ModelNet40 densities), the tile engine from 36->13 up; the private-copy grad_filter kernel wins whenever it fits.
// second version: both channel counts <= 64 and the filter + the warps' aggregate areas in shared memory; measured
// against the tile engine it wins for the reference models' shapes (3->9, 9->9, 36->13); kept to Cin * Cout <= 512
static bool sc2_shape(int Csrc, int Nout) {
  return Csrc <= SC2_MAXC && Nout <= SC2_MAXC && Csrc * Nout <= 512;
}
// forward / grad_input: with more than 32 outputs (two per lane, one k group) the tile engine is faster
// (13->36 grad_input at 16 x 4096 points: 0.39 ms against 0.41 ms)
static bool sc2_gather_shape(int Csrc, int Nout) { return sc2_shape(Csrc, Nout) && Nout <= 32; }
bool small_channels_supported(int Cin, int Cout) {
  if (engine_flag(1024)) return Cin <= 16 && Cout <= 16;   // first version only (A/B timing)
  return sc2_gather_shape(Cin, Cout) && sc2_gather_shape(Cout, Cin);   // forward and grad_input both
}
bool small_forward_supported(int Cin, int Cout) {
  if (engine_flag(1024)) return Cin <= 16 && Cout <= 16;
  return sc2_gather_shape(Cin, Cout);
}
bool small_backward_input_supported(int Cin, int Cout) {
  if (engine_flag(1024)) return Cin <= 16 && Cout <= 16;
  return sc2_gather_shape(Cout, Cin);
}
bool small_backward_filter_supported(int Cin, int Cout) {
  if (!(engine_flag(1024)) && sc2_shape(Cin, Cout)) return true;
  return Cin <= 40 && Cout <= 40 && (size_t)SC_WARPS * C3P_NCELL * Cin * Cout * 4 <= 96 * 1024;
}

static int launch_small_gc(SCArgs& a, const char* name, cudaStream_t stream) {
  if (a.total_points == 0) return CONV3P_OK;
  if (sc2_gather_shape(a.Csrc, a.Nout) && !(engine_flag(1024))) {   // engine bit 1024: first version (A/B timing)
    const size_t smem2 = sizeof(float) * (((size_t)C3P_NCELL * a.Csrc * a.Nout + 3) / 4 * 4 +
                                          (size_t)SC_WARPS * sc2_warp_words(a.Csrc));
    if (smem2 > 40 * 1024)
      { const int st_ = ensure_dynamic_smem(k_small_gather_contract2, smem2); if (st_) return st_; }
    long long grid2 = (long long)sc_sms() * (smem2 > 100 * 1024 ? 1 : (smem2 > 48 * 1024 ? 2 : 4));
    const long long need2 = (a.total_points + SC_WARPS - 1) / SC_WARPS;
    if (grid2 > need2) grid2 = need2;
    {
      LaunchTimer timer_(name, stream);
      k_small_gather_contract2<<<(unsigned)grid2, SC_THREADS, smem2, stream>>>(a);
    }
    C3P_LAUNCH_CHECK(name);
    return CONV3P_OK;
  }
  const size_t smem = sizeof(float) * C3P_NCELL * a.Csrc * a.Nout;
  if (smem > 40 * 1024)  // (static shared memory counts against the 48 KB default too)
    { const int st_ = ensure_dynamic_smem(k_small_gather_contract, smem); if (st_) return st_; }
  const int per_sm = smem > 100 * 1024 ? 1 : (smem > 48 * 1024 ? 2 : 4);
  long long grid = (long long)sc_sms() * per_sm;
  const long long need = (a.total_points + SC_WARPS - 1) / SC_WARPS;
  if (grid > need) grid = need;
  {
    LaunchTimer timer_(name, stream);
    k_small_gather_contract<<<(unsigned)grid, SC_THREADS, smem, stream>>>(a);
  }
  C3P_LAUNCH_CHECK(name);
  return CONV3P_OK;
}

int launch_forward_small(const conv3p_geom_t* g, const PlanView& v, const float* input, const float* filter,
                         int Cin, int Cout, float* output, cudaStream_t stream, const RowIO& io) {
  SCArgs a{};
  a.src = input; a.filter = filter; a.out = output; a.cnt = v.count_table; a.begin = v.pair_begin;
  a.len = v.pair_len; a.rows = v.pair_row; a.weights = nullptr; a.sorted_xyzi = v.sorted_xyzi;
  a.total_points = (long long)g->B * g->N; a.capacity = g->pair_capacity; a.N = g->N;
  a.Cin = Cin; a.Cout = Cout; a.Csrc = Cin; a.Nout = Cout; a.transposed = 0;
  a.src_stride = io.src_stride ? io.src_stride : Cin; a.out_stride = io.out_stride ? io.out_stride : Cout;
  a.activation = io.activation;
  return launch_small_gc(a, "k_small_forward", stream);
}

int launch_backward_input_small(const conv3p_geom_t* g, const PlanView& v, const float* grad_out,
                                const float* filter, int Cin, int Cout, float* grad_input,
                                cudaStream_t stream) {
  SCArgs a{};
  a.src = grad_out; a.filter = filter; a.out = grad_input; a.cnt = v.bwd_count; a.begin = v.pair_begin;
  a.len = v.pair_len; a.rows = v.bwd_row; a.weights = v.bwd_weight; a.sorted_xyzi = v.sorted_xyzi;
  a.total_points = (long long)g->B * g->N; a.capacity = g->pair_capacity; a.N = g->N;
  a.Cin = Cin; a.Cout = Cout; a.Csrc = Cout; a.Nout = Cin; a.transposed = 1;
  a.src_stride = Cout; a.out_stride = Cin; a.activation = 0;
  return launch_small_gc(a, "k_small_backward_input", stream);
}

size_t backward_filter_small_scratch_bytes(int Cin, int Cout) {
  return align_up(sizeof(float) * (size_t)4 * 256 * C3P_NCELL * Cin * Cout);  // up to 4 CTAs on up to 256 SMs
}

int launch_backward_filter_small(const conv3p_geom_t* g, const PlanView& v, const float* grad_out,
                                 const float* input, int Cin, int Cout, float* grad_filter, void* scratch,
                                 size_t scratch_bytes, cudaStream_t stream) {
  const int nW = C3P_NCELL * Cin * Cout;
  const long long pts = (long long)g->B * g->N;
  if (pts == 0) {
    C3P_CUDA(cudaMemsetAsync(grad_filter, 0, sizeof(float) * nW, stream));
    return CONV3P_OK;
  }
  if (!scratch || scratch_bytes < backward_filter_small_scratch_bytes(Cin, Cout)) return CONV3P_ERR_BUFFER_TOO_SMALL;
  SFArgs a{};
  a.grad_out = grad_out; a.input = input; a.cnt = v.bwd_count; a.begin = v.pair_begin; a.len = v.pair_len;
  a.rows = v.bwd_row; a.weights = v.bwd_weight; a.sorted_xyzi = v.sorted_xyzi;
  a.partial = static_cast<float*>(scratch);
  a.total_points = pts; a.capacity = g->pair_capacity; a.N = g->N; a.Cin = Cin; a.Cout = Cout;
  const bool v2 = sc2_shape(Cin, Cout) && !(engine_flag(1024));   // engine bit 1024: first version (A/B timing)
  // second version: as many cells per pass as 8 private copies fit next to the aggregate areas with two CTAs per SM
  int cells_per_pass = C3P_NCELL;
  if (v2) {
    const size_t per_cell = sizeof(float) * (size_t)SC_WARPS * Cin * Cout;
    const size_t room = 112 * 1024 - sizeof(float) * (size_t)SC_WARPS * sc2_warp_words(Cout);
    if (per_cell * C3P_NCELL > room) cells_per_pass = (int)(room / per_cell);
    if (cells_per_pass < 1) cells_per_pass = 1;
  }
  const size_t smem = v2 ? sizeof(float) * (((size_t)SC_WARPS * cells_per_pass * Cin * Cout + 3) / 4 * 4 +
                                            (size_t)SC_WARPS * sc2_warp_words(Cout))
                         : sizeof(float) * (size_t)SC_WARPS * nW;
  if (smem > 40 * 1024) {  // (static shared memory counts against the 48 KB default too)
    if (v2)
      { const int st_ = ensure_dynamic_smem(k_small_backward_filter2, smem); if (st_) return st_; }
    else
      { const int st_ = ensure_dynamic_smem(k_small_backward_filter, smem); if (st_) return st_; }
  }
  const int per_sm = smem > 113 * 1024 ? 1 : 2;
  int sms = sc_sms();
  if (sms > 256) sms = 256;
  long long grid = (long long)sms * per_sm;
  const long long need = (pts + SC_WARPS - 1) / SC_WARPS;
  if (grid > need) grid = need;
  {
    LaunchTimer timer_("k_small_backward_filter", stream);
    if (v2)
      k_small_backward_filter2<<<(unsigned)grid, SC_THREADS, smem, stream>>>(a, cells_per_pass);
    else
      k_small_backward_filter<<<(unsigned)grid, SC_THREADS, smem, stream>>>(a);
  }
  C3P_LAUNCH_CHECK("k_small_backward_filter");
  return launch_reduce_partials(a.partial, (int)grid, nW, grad_filter, v.header, stream);
}

}  // namespace c3p
