// tc_launch.cu -- host side of the tensor-core (tcgen05, 3xTF32) forward / grad_input path: which shapes take it,
// the weight panel images the fused gather + MMA kernel (gather_mma2.cu) streams, and the two launch wrappers.
//
//   forward     out[p, :]        = sum_f mean_f(input rows)[p, :]    * W_f        (tf_conv3p_atrous.cpp:480-494)
//   grad_input  grad_input[j, :] = sum_f' G_f'[j, :]                 * W_f'^T     (tf_conv3p_atrous.cpp:682-692)
//
// Both are the same contraction out[p, n] = sum_f sum_k A_f[p, k] * Wpanel_f[n, k]; only the gathered rows, the
// lists and the orientation of the weight panels differ.  The weights are pre-split (TF32 hi / lo), pre-transposed
// and pre-swizzled ONCE per call into panel images, so one thread of the main kernel streams them with bulk async
// copies (UBLKCP) signalled on mbarriers.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace c3p {

using namespace tc;

// weights [27][Cin][Cout] -> panel images: for (f, kc, hl) a [R rows x 32 k] K-major 128B-swizzled panel
// Second panel of a pair: the lo parts as fp32 (3xTF32, bf16c == 0) or the BF16 correction panel [W_hi | W_lo]
// (64 bf16 per row) that multiplies the A-side panel [A_lo | A_hi] (tc_common.cuh).
__global__ void k_prep_weight_panels(const float* __restrict__ filter, unsigned char* __restrict__ wp,
                                     int Cin, int Cout, int transposed_out, int bf16c) {
  // transposed_out == 0: rows = Cout (n = c), K = Cin (forward B operand, W^T)
  // transposed_out == 1: rows = Cin  (n = k), K = Cout (input-gradient B operand, W)
  const int R = transposed_out ? Cin : Cout;   // panel rows
  const int KD = transposed_out ? Cout : Cin;  // contraction length
  const int nkc = KD / PANEL_K;
  const long long total = (long long)C3P_NCELL * nkc * R * PANEL_K;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e % PANEL_K);
    const int r = (int)((e / PANEL_K) % R);
    const int kc = (int)((e / ((long long)PANEL_K * R)) % nkc);
    const int f = (int)(e / ((long long)PANEL_K * R * nkc));
    const int kk = kc * PANEL_K + k;
    const float w = transposed_out ? filter[((size_t)f * Cin + r) * Cout + kk]
                                   : filter[((size_t)f * Cin + kk) * Cout + r];
    const float h = bf16c ? tf32_rn(w) : tf32_hi(w);
    unsigned char* base = wp + ((size_t)(f * nkc + kc) * 2) * R * PANEL_ROW_BYTES;
    *reinterpret_cast<float*>(base + panel_offset(r, k)) = h;
    unsigned char* second = base + (size_t)R * PANEL_ROW_BYTES;
    if (bf16c && !transposed_out) {   // forward: interleaved K order (tc_common.cuh, panel_offset16i)
      *reinterpret_cast<__nv_bfloat16*>(second + panel_offset16i(r, k, 0)) = __float2bfloat16_rn(h);
      *reinterpret_cast<__nv_bfloat16*>(second + panel_offset16i(r, k, 1)) = __float2bfloat16_rn(w - h);
    } else if (bf16c) {
      *reinterpret_cast<__nv_bfloat16*>(second + panel_offset16(r, k)) = __float2bfloat16_rn(h);
      *reinterpret_cast<__nv_bfloat16*>(second + panel_offset16(r, 32 + k)) = __float2bfloat16_rn(w - h);
    } else {
      *reinterpret_cast<float*>(second + panel_offset(r, k)) = w - h;
    }
  }
}

// Shapes the tensor-core path takes (a real dense GEMM per cell: channel counts in multiples of 32 / 16 whose
// operand ring fits in shared memory); everything else stays on the fp32 SIMT engines.
bool forward_tc_supported(int N, long long capacity, int Cin, int Cout) {
  return gather_mma2_supported(N, capacity, Cin, Cout);
}
bool backward_input_tc_supported(int N, long long capacity, int Cin, int Cout) {
  return gather_mma2_supported(N, capacity, Cout, Cin);
}

size_t weight_panel_bytes(int Cin, int Cout) { return align_up((size_t)2 * C3P_NCELL * Cin * Cout * 4); }

size_t tc_items_bytes(const conv3p_geom_t* g, int Cin, int Cout) {
  if (gather_mma2_supported(g->N, g->pair_capacity, Cin, Cout) ||
      gather_mma2_supported(g->N, g->pair_capacity, Cout, Cin))
    return gather_mma2_scratch_bytes(g);
  return 0;
}

int launch_prep_weight_panels(const float* filter, void* wp, int Cin, int Cout, int transposed_out,
                              cudaStream_t stream) {
  const long long total = (long long)C3P_NCELL * Cin * Cout;
  const int threads = 256;
  const int blocks = (int)((total + threads - 1) / threads < 2048 ? (total + threads - 1) / threads : 2048);
  {
    LaunchTimer timer_("k_prep_weight_panels", stream);
    k_prep_weight_panels<<<blocks, threads, 0, stream>>>(filter, static_cast<unsigned char*>(wp), Cin, Cout,
                                                         transposed_out, engine_flag(512) ? 0 : 1);
  }
  C3P_LAUNCH_CHECK("k_prep_weight_panels");
  return CONV3P_OK;
}

int launch_forward_tc(const conv3p_geom_t* g, const PlanView& v, const float* input, const float* filter,
                      int Cin, int Cout, float* output, void* scratch, size_t scratch_bytes,
                      cudaStream_t stream, const RowIO& io) {
  if (!forward_tc_supported(g->N, g->pair_capacity, Cin, Cout)) return CONV3P_ERR_UNSUPPORTED;
  const size_t wpb = weight_panel_bytes(Cin, Cout);
  if (!scratch || scratch_bytes < wpb) return CONV3P_ERR_BUFFER_TOO_SMALL;
  int st = launch_prep_weight_panels(filter, scratch, Cin, Cout, 0, stream);
  if (st) return st;
  return launch_gather_mma2(g, v, input, scratch, Cin, Cout, output, false, static_cast<char*>(scratch) + wpb,
                            scratch_bytes - wpb, "k_forward_tc", stream, nullptr, io);
}

int launch_backward_input_tc(const conv3p_geom_t* g, const PlanView& v, const float* grad_out,
                             const float* filter, int Cin, int Cout, float* grad_input, void* scratch,
                             size_t scratch_bytes, cudaStream_t stream, float* g_store, const GroupItems* half_items) {
  if (!backward_input_tc_supported(g->N, g->pair_capacity, Cin, Cout)) return CONV3P_ERR_UNSUPPORTED;
  const size_t wpb = weight_panel_bytes(Cin, Cout);
  if (!scratch || scratch_bytes < wpb) return CONV3P_ERR_BUFFER_TOO_SMALL;
  int st = launch_prep_weight_panels(filter, scratch, Cin, Cout, 1, stream);
  if (st) return st;
  return launch_gather_mma2(g, v, grad_out, scratch, Cout, Cin, grad_input, true, static_cast<char*>(scratch) + wpb,
                            scratch_bytes - wpb, "k_backward_input_tc", stream, g_store, RowIO(), half_items);
}

}  // namespace c3p
