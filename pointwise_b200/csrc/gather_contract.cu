// gather_contract.cu -- fp32 SIMT engine for the Conv3p forward pass and the input gradient.
//
// Both are the same shape of work: for a tile of P points and each of the 27
// kernel cells f, gather the rows of the cell's members into a per-point aggregate
//     forward : A_f[p, k] = (1/cnt(p,f)) * sum_{j in cell f of p} input[j, k]        (per-cell mean)
//     backward: A_f[p, c] = sum_{(ii,w) in cell f' of p} w * grad_out[ii, c]           (w = 1/cnt(ii,f'))
// and contract it with the cell's weight matrix
//     forward : out[p, c]     += sum_k A_f[p,k] * W[f,k,c]     (tf_conv3p_atrous.cpp:480-494)
//     backward: grad_in[p, k] += sum_c A_f[p,c] * W[f,k,c]     (tf_conv3p_atrous.cpp:682-692)
// The aggregate lives only in shared memory (it is 27*C floats per point -- writing it to HBM would
// be ~9x the compulsory traffic), the weights are staged per (cell, K-chunk), accumulators stay in
// registers, and the output row is written exactly once (no pre-zeroing, no read-modify-write, in
// contrast to tf_conv3p_atrous.cu:366-375).  Tiles follow the voxel-sorted order of the plan so the
// members gathered by one CTA share cache lines.  Cells empty for the whole tile are skipped.
#include "common.cuh"

namespace c3p {

constexpr int GC_THREADS = 256;

struct GCArgs {
  const float* src;        // gathered rows [B*N, Csrc]
  const float* filter;     // [27, Cin, Cout]
  float* out;              // [B*N, Nout]
  const int* cnt;          // [B*N, 27] group sizes of the lists
  const long long* begin;  // [B*N]
  const int* len;          // [B*N] forward list length (capacity guard)
  const int* rows;         // list entries (global rows)
  const float* weights;    // per-entry weights, or nullptr for the per-cell mean
  const float4* sorted_xyzi;
  long long total_points;
  long long capacity;
  int N, Cin, Cout;
  int Csrc, Nout;
  long long src_stride, out_stride;  // floats between rows
  int activation;
  int transposed;          // 0: B[kk][n] = W[f][kk][n] (forward); 1: B[kk][n] = W[f][n][kk]
  int P, nTx, nTy, KC, NoutPad;
};

template <int TP, int TC>
__global__ void __launch_bounds__(GC_THREADS) k_gather_contract(const GCArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int P = a.P, KC = a.KC, NoutPad = a.NoutPad, lda = KC + 1;
  float* Bsm = reinterpret_cast<float*>(smem_raw);                 // [KC][NoutPad]
  float* Asm = Bsm + (size_t)KC * NoutPad;                         // [P][KC+1]
  long long* beg = reinterpret_cast<long long*>(Asm + (((size_t)P * lda + 1) & ~(size_t)1));  // [P]
  int* pref = reinterpret_cast<int*>(beg + P);                     // [P][28] exclusive prefix of cnt
  int* rowid = pref + (size_t)P * 28;                              // [P]
  __shared__ unsigned tile_mask;

  const int tid = threadIdx.x;
  const long long s0 = (long long)blockIdx.x * P;
  const int n0 = blockIdx.y * NoutPad;  // first output column of this CTA
  if (tid == 0) tile_mask = 0;
  __syncthreads();
  for (int p = tid; p < P; p += GC_THREADS) {
    const long long s = s0 + p;
    int row = -1;
    long long bg = 0;
    unsigned mask = 0;
    int run = 0;
    if (s < a.total_points) {
      const int b = (int)(s / a.N);
      row = b * a.N + __float_as_int(a.sorted_xyzi[s].w);
      bg = a.begin[row];
      if (bg + a.len[row] > a.capacity) row = -2 - row;  // list incomplete: poison this point
    }
    for (int f = 0; f < C3P_NCELL; ++f) {
      pref[p * 28 + f] = run;
      const int c = row >= 0 ? a.cnt[(size_t)row * C3P_NCELL + f] : 0;
      run += c;
      if (c) mask |= 1u << f;
    }
    pref[p * 28 + 27] = run;
    rowid[p] = row;
    beg[p] = bg;
    if (mask) atomicOr(&tile_mask, mask);
  }
  __syncthreads();
  const unsigned cells = tile_mask;

  const int tx = tid % a.nTx, ty = tid / a.nTx;
  const bool active = ty < a.nTy;
  float acc[TP][TC];
#pragma unroll
  for (int i = 0; i < TP; ++i)
#pragma unroll
    for (int j = 0; j < TC; ++j) acc[i][j] = 0.f;

  for (int f = 0; f < C3P_NCELL; ++f) {
    if (!((cells >> f) & 1u)) continue;
    for (int kc = 0; kc < a.Csrc; kc += KC) {
      const int kw = min(KC, a.Csrc - kc);
      // ---- gather + aggregate: thread per (point, channel), channel fastest -> coalesced rows ----
      for (int e = tid; e < P * KC; e += GC_THREADS) {
        const int p = e / KC, k = e - p * KC;
        float s = 0.f;
        const int o = pref[p * 28 + f];
        const int n = pref[p * 28 + f + 1] - o;
        if (k < kw && n > 0) {
          const long long at = beg[p] + o;
          const int* r = a.rows + at;
          if (a.weights) {
            const float* w = a.weights + at;
            for (int m = 0; m < n; ++m)
              s = fmaf(__ldg(w + m), __ldg(a.src + (size_t)__ldg(r + m) * a.src_stride + kc + k), s);
          } else {
            for (int m = 0; m < n; ++m) s += __ldg(a.src + (size_t)__ldg(r + m) * a.src_stride + kc + k);
            s = __fdiv_rn(s, (float)n);
          }
        }
        Asm[p * lda + k] = s;
      }
      // ---- stage the cell's weights ------------------------------------------------------------------
      if (!a.transposed) {
        for (int e = tid; e < KC * NoutPad; e += GC_THREADS) {
          const int kk = e / NoutPad, n = e - kk * NoutPad;
          float w = 0.f;
          if (kk < kw && n0 + n < a.Nout)
            w = __ldg(a.filter + ((size_t)f * a.Cin + kc + kk) * a.Cout + n0 + n);
          Bsm[e] = w;
        }
      } else {
        for (int e = tid; e < KC * NoutPad; e += GC_THREADS) {
          const int n = e / KC, kk = e - n * KC;  // kk (= c) fastest: coalesced global reads
          float w = 0.f;
          if (kk < kw && n0 + n < a.Nout)
            w = __ldg(a.filter + ((size_t)f * a.Cin + n0 + n) * a.Cout + kc + kk);
          Bsm[kk * NoutPad + n] = w;
        }
      }
      __syncthreads();
      // ---- contract: TP x TC register tile per thread --------------------------------------------------
      if (active) {
        const float* Arow = Asm + (size_t)(ty * TP) * lda;
        const float* Bcol = Bsm + tx * TC;
#pragma unroll 4
        for (int kk = 0; kk < kw; ++kk) {
          float av[TP], bv[TC];
#pragma unroll
          for (int i = 0; i < TP; ++i) av[i] = Arow[i * lda + kk];
          if (TC % 4 == 0) {
#pragma unroll
            for (int j = 0; j < TC; j += 4) {
              const float4 t = *reinterpret_cast<const float4*>(Bcol + (size_t)kk * NoutPad + j);
              bv[j] = t.x; bv[j + 1] = t.y; bv[j + 2] = t.z; bv[j + 3] = t.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < TC; ++j) bv[j] = Bcol[(size_t)kk * NoutPad + j];
          }
#pragma unroll
          for (int i = 0; i < TP; ++i)
#pragma unroll
            for (int j = 0; j < TC; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
      }
      __syncthreads();
    }
  }

  if (active) {
#pragma unroll
    for (int i = 0; i < TP; ++i) {
      const int p = ty * TP + i;
      int row = rowid[p];
      if (row == -1) continue;
      const bool poison = row < -1;
      if (poison) row = -2 - row;
      float* o = a.out + (size_t)row * a.out_stride + n0 + tx * TC;
#pragma unroll
      for (int j = 0; j < TC; ++j)
        if (n0 + tx * TC + j < a.Nout)
          o[j] = poison ? __int_as_float(0x7fc00000) : apply_activation(acc[i][j], a.activation);
    }
  }
}

template <int TP, int TC>
static int launch_cfg(GCArgs& a, cudaStream_t stream) {
  const int maxTx = 32;
  int nTx = (a.Nout + TC - 1) / TC;
  if (nTx > maxTx) nTx = maxTx;
  a.nTx = nTx;
  a.nTy = GC_THREADS / nTx;
  a.P = a.nTy * TP;
  a.NoutPad = nTx * TC;
  a.KC = a.Csrc < 32 ? a.Csrc : 32;
  const int ncol = (a.Nout + a.NoutPad - 1) / a.NoutPad;
  const size_t smem = sizeof(float) * ((size_t)a.KC * a.NoutPad + (((size_t)a.P * (a.KC + 1) + 1) & ~(size_t)1)) +
                      sizeof(long long) * a.P + sizeof(int) * ((size_t)a.P * 28 + a.P);
  auto kern = k_gather_contract<TP, TC>;
  if (smem > 40 * 1024)  // (static shared memory counts against the 48 KB default too)
    { const int st_ = ensure_dynamic_smem(kern, smem); if (st_) return st_; }
  const long long tiles = (a.total_points + a.P - 1) / a.P;
  dim3 grid((unsigned)tiles, (unsigned)ncol);
  {
    LaunchTimer timer_(a.transposed ? "k_gather_contract_bwd_input" : "k_gather_contract_fwd", stream);
    kern<<<grid, GC_THREADS, smem, stream>>>(a);
  }
  C3P_LAUNCH_CHECK("k_gather_contract");
  return CONV3P_OK;
}

static int launch_gc(GCArgs& a, cudaStream_t stream) {
  if (a.total_points == 0) return CONV3P_OK;
  if (a.Nout >= 32) return launch_cfg<4, 8>(a, stream);
  return launch_cfg<2, 4>(a, stream);
}

int launch_forward_simt(const conv3p_geom_t* g, const PlanView& v, const float* input,
                        const float* filter, int Cin, int Cout, float* output, cudaStream_t stream,
                        const RowIO& io) {
  GCArgs a{};
  a.src = input; a.filter = filter; a.out = output;
  a.cnt = v.count_table; a.begin = v.pair_begin; a.len = v.pair_len; a.rows = v.pair_row;
  a.weights = nullptr; a.sorted_xyzi = v.sorted_xyzi;
  a.total_points = (long long)g->B * g->N; a.capacity = g->pair_capacity;
  a.N = g->N; a.Cin = Cin; a.Cout = Cout; a.Csrc = Cin; a.Nout = Cout; a.transposed = 0;
  a.src_stride = io.src_stride ? io.src_stride : Cin; a.out_stride = io.out_stride ? io.out_stride : Cout;
  a.activation = io.activation;
  return launch_gc(a, stream);
}

int launch_backward_input_simt(const conv3p_geom_t* g, const PlanView& v, const float* grad_out,
                               const float* filter, int Cin, int Cout, float* grad_input,
                               cudaStream_t stream) {
  GCArgs a{};
  a.src = grad_out; a.filter = filter; a.out = grad_input;
  a.cnt = v.bwd_count; a.begin = v.pair_begin; a.len = v.pair_len; a.rows = v.bwd_row;
  a.weights = v.bwd_weight; a.sorted_xyzi = v.sorted_xyzi;
  a.total_points = (long long)g->B * g->N; a.capacity = g->pair_capacity;
  a.N = g->N; a.Cin = Cin; a.Cout = Cout; a.Csrc = Cout; a.Nout = Cin; a.transposed = 1;
  a.src_stride = Cout; a.out_stride = Cin; a.activation = 0;
  return launch_gc(a, stream);
}

}  // namespace c3p
