// backward_filter.cu -- fp32 SIMT engine for the weight gradient.
//
//   grad_filter[f, k, c] = sum_j input[j, k] * G_f[j, c],   G_f[j, c] = sum_{(ii,w) in cell f of j} w * grad_out[ii, c]
//
// which is the reference's accumulation `grad_filter[f',k,c] += g[ii,c] * in[j,k] / count`
// (tf_conv3p_atrous.cpp:694-696) regrouped by (j, f').  The reference reduces with per-thread copies
// (:611-621, :709-716) on the CPU and with one global atomicAdd per (pair, k, c) on the GPU
// (tf_conv3p_atrous.cu:494); here the reduction over points is a split-K contraction: CTA
// (f, split, output block) keeps its [KB x CB] block of grad_filter[f] in registers while it walks
// its share of the point tiles, writes one partial, and a second kernel sums the partials in a fixed
// order -- deterministic, no atomics.
#include "common.cuh"

namespace c3p {

constexpr int BF_THREADS = 256;
constexpr int BF_P = 32;  // points per tile

struct BFArgs {
  const float* grad_out;
  const float* input;
  const int* cnt;          // bwd_count [B*N,27]
  const long long* begin;
  const int* len;
  const int* rows;
  const float* weights;
  const float4* sorted_xyzi;
  float* partial;          // [S][27*Cin*Cout]
  long long total_points, capacity;
  int N, Cin, Cout;
  int nTx, nTy, KB, CB, nCB;
  int tiles_per_split;
};

template <int TK, int TC>
__global__ void __launch_bounds__(BF_THREADS) k_backward_filter(const BFArgs a) {
  extern __shared__ __align__(16) float bf_smem[];
  float* Xsm = bf_smem;                        // [BF_P][KB]
  float* Gsm = Xsm + (size_t)BF_P * a.KB;      // [BF_P][CB]
  __shared__ long long at[BF_P];
  __shared__ int members[BF_P];
  __shared__ int rowid[BF_P];

  const int f = blockIdx.x, split = blockIdx.y;
  const int kb = blockIdx.z / a.nCB, cb = blockIdx.z % a.nCB;
  const int k0 = kb * a.KB, c0 = cb * a.CB;
  const int kw = min(a.KB, a.Cin - k0), cw = min(a.CB, a.Cout - c0);
  const int tid = threadIdx.x;
  const int tx = tid % a.nTx, ty = tid / a.nTx;
  const bool active = ty < a.nTy;

  float acc[TK][TC];
#pragma unroll
  for (int i = 0; i < TK; ++i)
#pragma unroll
    for (int j = 0; j < TC; ++j) acc[i][j] = 0.f;

  const long long tiles = (a.total_points + BF_P - 1) / BF_P;
  const long long t_begin = (long long)split * a.tiles_per_split;
  const long long t_end = min(tiles, t_begin + a.tiles_per_split);
  for (long long t = t_begin; t < t_end; ++t) {
    int n = 0;
    if (tid < BF_P) {
      const long long s = t * BF_P + tid;
      int row = -1;
      long long where = 0;
      if (s < a.total_points) {
        const int b = (int)(s / a.N);
        row = b * a.N + __float_as_int(a.sorted_xyzi[s].w);
        where = a.begin[row];
        if (where + a.len[row] <= a.capacity) {
          const int* c = a.cnt + (size_t)row * C3P_NCELL;
          for (int ff = 0; ff < f; ++ff) where += __ldg(c + ff);
          n = __ldg(c + f);
        }
      }
      at[tid] = where;
      members[tid] = n;
      rowid[tid] = row;
    }
    if (!__syncthreads_or(n > 0)) continue;  // cell f empty for the whole tile

    for (int e = tid; e < BF_P * a.KB; e += BF_THREADS) {
      const int p = e / a.KB, k = e - p * a.KB;
      float x = 0.f;
      if (k < kw && members[p] > 0) x = __ldg(a.input + (size_t)rowid[p] * a.Cin + k0 + k);
      Xsm[e] = x;
    }
    for (int e = tid; e < BF_P * a.CB; e += BF_THREADS) {
      const int p = e / a.CB, c = e - p * a.CB;
      float s = 0.f;
      const int m_n = members[p];
      if (c < cw && m_n > 0) {
        const int* r = a.rows + at[p];
        const float* w = a.weights + at[p];
        for (int m = 0; m < m_n; ++m)
          s = fmaf(__ldg(w + m), __ldg(a.grad_out + (size_t)__ldg(r + m) * a.Cout + c0 + c), s);
      }
      Gsm[e] = s;
    }
    __syncthreads();
    if (active) {
      const float* X = Xsm + ty * TK;
      const float* G = Gsm + tx * TC;
#pragma unroll 4
      for (int p = 0; p < BF_P; ++p) {
        float xv[TK], gv[TC];
#pragma unroll
        for (int i = 0; i < TK; ++i) xv[i] = X[(size_t)p * a.KB + i];
#pragma unroll
        for (int j = 0; j < TC; ++j) gv[j] = G[(size_t)p * a.CB + j];
#pragma unroll
        for (int i = 0; i < TK; ++i)
#pragma unroll
          for (int j = 0; j < TC; ++j) acc[i][j] = fmaf(xv[i], gv[j], acc[i][j]);
      }
    }
    __syncthreads();
  }

  if (active) {
    float* out = a.partial + ((size_t)split * C3P_NCELL + f) * a.Cin * a.Cout;
#pragma unroll
    for (int i = 0; i < TK; ++i) {
      const int k = ty * TK + i;
      if (k >= kw) continue;
#pragma unroll
      for (int j = 0; j < TC; ++j) {
        const int c = tx * TC + j;
        if (c < cw) out[(size_t)(k0 + k) * a.Cout + c0 + c] = acc[i][j];
      }
    }
  }
}

struct BFConfig {
  int TK, TC, nTx, nTy, KB, CB, nKB, nCB, S, tiles_per_split;
};

static BFConfig bf_config(const conv3p_geom_t* g, int Cin, int Cout) {
  BFConfig c;
  const bool big = (long long)Cin * Cout >= 2048 && Cout >= 16;
  c.TK = big ? 4 : 1;
  c.TC = big ? 8 : 1;
  const int maxTx = big ? 16 : 32;
  c.nTx = (Cout + c.TC - 1) / c.TC;
  if (c.nTx > maxTx) c.nTx = maxTx;
  c.nTy = BF_THREADS / c.nTx;
  const int needTy = (Cin + c.TK - 1) / c.TK;
  if (c.nTy > needTy) c.nTy = needTy;
  c.KB = c.nTy * c.TK;
  c.CB = c.nTx * c.TC;
  c.nKB = (Cin + c.KB - 1) / c.KB;
  c.nCB = (Cout + c.CB - 1) / c.CB;
  const long long pts = (long long)g->B * g->N;
  const long long tiles = (pts + BF_P - 1) / BF_P;
  long long S = (4 * 148 + C3P_NCELL * c.nKB * c.nCB - 1) / (C3P_NCELL * c.nKB * c.nCB);
  if (S > tiles) S = tiles;
  if (S < 1) S = 1;
  c.tiles_per_split = (int)((tiles + S - 1) / S);
  if (c.tiles_per_split < 1) c.tiles_per_split = 1;
  c.S = (int)((tiles + c.tiles_per_split - 1) / c.tiles_per_split);
  if (c.S < 1) c.S = 1;
  return c;
}

size_t backward_filter_scratch_bytes(const conv3p_geom_t* g, int Cin, int Cout) {
  const BFConfig c = bf_config(g, Cin, Cout);
  return align_up(sizeof(float) * (size_t)c.S * C3P_NCELL * Cin * Cout);
}

int launch_backward_filter_simt(const conv3p_geom_t* g, const PlanView& v, const float* grad_out,
                                const float* input, int Cin, int Cout, float* grad_filter,
                                void* scratch, size_t scratch_bytes, cudaStream_t stream) {
  const long long nW = (long long)C3P_NCELL * Cin * Cout;
  const long long pts = (long long)g->B * g->N;
  if (pts == 0) {
    C3P_CUDA(cudaMemsetAsync(grad_filter, 0, sizeof(float) * nW, stream));
    return CONV3P_OK;
  }
  const BFConfig c = bf_config(g, Cin, Cout);
  if (scratch_bytes < sizeof(float) * (size_t)c.S * nW || scratch == nullptr)
    return CONV3P_ERR_BUFFER_TOO_SMALL;
  BFArgs a{};
  a.grad_out = grad_out; a.input = input; a.cnt = v.bwd_count; a.begin = v.pair_begin;
  a.len = v.pair_len; a.rows = v.bwd_row; a.weights = v.bwd_weight; a.sorted_xyzi = v.sorted_xyzi;
  a.partial = static_cast<float*>(scratch);
  a.total_points = pts; a.capacity = g->pair_capacity; a.N = g->N; a.Cin = Cin; a.Cout = Cout;
  a.nTx = c.nTx; a.nTy = c.nTy; a.KB = c.KB; a.CB = c.CB; a.nCB = c.nCB;
  a.tiles_per_split = c.tiles_per_split;
  const size_t smem = sizeof(float) * (size_t)BF_P * (c.KB + c.CB);
  dim3 grid(C3P_NCELL, c.S, c.nKB * c.nCB);
  if (c.TK == 4) {
    if (smem > 40 * 1024) {  // (static shared memory counts against the 48 KB default too)
      const int st = ensure_dynamic_smem(k_backward_filter<4, 8>, smem);
      if (st) return st;
    }
    {
    LaunchTimer timer_("k_backward_filter", stream);
    k_backward_filter<4, 8><<<grid, BF_THREADS, smem, stream>>>(a);
  }
  } else {
    {
    LaunchTimer timer_("k_backward_filter", stream);
    k_backward_filter<1, 1><<<grid, BF_THREADS, smem, stream>>>(a);
  }
  }
  C3P_LAUNCH_CHECK("k_backward_filter");
  return launch_reduce_partials(a.partial, c.S, nW, grad_filter, v.header, stream);
}

}  // namespace c3p
