// common.cuh -- shared declarations of libconv3p_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/conv3p_b200.h"

#define C3P_NCELL 27
#define C3P_FULL_MASK 0xffffffffu

// Header slots (int64 each) at plan_layout.header.
enum { H_CURSOR = 0, H_OVERFLOW = 1, H_BWD_PAIRS = 2, H_HAS_BWD = 3, H_SLOTS = 16 };

struct PlanView {
  long long* header;
  float* cloud_meta;   // [B][8]
  uint32_t* sorted_key;
  float4* sorted_xyzi;
  int* count_table;
  long long* pair_begin;
  int* pair_len;
  int* pair_row;
  int* bwd_count;
  int* bwd_row;
  float* bwd_weight;
  uint32_t* sort_tmp;
  uint32_t* cell_start;   // [B][cell_cap + 1] bin offsets of the voxel grid (k_cloud_sort), see cell_cap()
  int cell_cap;
};

namespace c3p {

// Row layout and epilogue of a forward call (conv3p_forward_ex_f32): rows of `input` / `output` may live inside wider
// buffers (row strides in floats; 0 = dense), and the output may pass through SELU before it is stored --
// the activation that follows every Conv3p of the reference's networks (scene_seg/pointcnn_scene_seg_acsd.py:35-36)
// and the concat of their outputs (:56) then cost no extra pass over memory.
struct RowIO {
  long long src_stride = 0;   // floats between consecutive gathered rows
  long long out_stride = 0;   // floats between consecutive output rows
  int activation = 0;         // CONV3P_ACT_NONE / CONV3P_ACT_SELU
};

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }
// Cells of the per-cloud bin-offset table: enough for the clouds of the reference's data at voxel 0.1 (a [-1,1]^3
// shape has 22^3 = 10,648 cells, an S3DIS block 12 x 12 x 32 = 4,608); a cloud with more cells falls back to bisection.
inline int cell_cap(int N) {
  const long long c = 16LL * N;
  return (int)(c < 4096 ? 4096 : (c > 65536 ? 65536 : c));
}

int compute_layout(const conv3p_geom_t* g, conv3p_plan_layout_t* L);
int make_view(const conv3p_geom_t* g, const void* plan, size_t plan_bytes, PlanView* v);
int check_geom(const conv3p_geom_t* g);

int cuda_fail(cudaError_t e, const char* what);  // records the text, returns CONV3P_ERR_CUDA

// Brackets one kernel launch with CUDA events on its stream while conv3p_profile_enable(1) is in
// effect (bench.py's live per-kernel timing); a no-op otherwise.
struct LaunchTimer {
  LaunchTimer(const char* name, cudaStream_t stream);
  ~LaunchTimer();
  int slot;
  cudaStream_t stream;
};
void count_launch(int n = 1);
// conv3p_set_engine value: low three bits = contraction engine (0 auto, 1 fp32 SIMT, 2 tensor cores where the shape
// allows, 3 generic fp32 tile kernels only); higher bits = ablation flags for A/B timing (see the header).
int engine();
inline int engine_kind() { return engine() & 7; }
inline bool engine_flag(int bit) { return (engine() & bit) != 0; }
inline bool engine_allows_tc() { return engine_kind() == 0 || engine_kind() == 2; }
inline bool engine_allows_small() { return engine_kind() != 3; }

// Per-device facts and per-kernel attributes, queried / set ONCE (not on every launch).
int sm_count();                                   // multiprocessors of the current device (cached per device)
int ensure_dynamic_smem(const void* kernel, size_t bytes);  // opt-in dynamic shared memory, once per (kernel, device)
template <typename K>
inline int ensure_dynamic_smem(K kernel, size_t bytes) {
  return ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), bytes);
}

// out[w] = sum_i partial[i][w] in a fixed order (deterministic).  When the plan's overflow flag is set (neighbour
// lists incomplete because pair_capacity was too small) the weight gradient would silently miss terms: it is
// poisoned with NaN instead, like the affected rows of output / grad_input.
int launch_reduce_partials(const float* partial, int S, long long nW, float* out, const long long* plan_header,
                           cudaStream_t stream);

// stages (each enqueues kernels on `stream` and returns a status)
int launch_cloud_sort(const conv3p_geom_t* g, const float* points, const PlanView& v,
                      cudaStream_t stream);
int launch_neighbor_search(const conv3p_geom_t* g, const PlanView& v, cudaStream_t stream);
int launch_backward_lists(const conv3p_geom_t* g, const float* points, const PlanView& v,
                          cudaStream_t stream);
int launch_forward_simt(const conv3p_geom_t* g, const PlanView& v, const float* input,
                        const float* filter, int Cin, int Cout, float* output, cudaStream_t stream,
                        const RowIO& io = RowIO());
int launch_backward_input_simt(const conv3p_geom_t* g, const PlanView& v, const float* grad_out,
                               const float* filter, int Cin, int Cout, float* grad_input,
                               cudaStream_t stream);
int launch_backward_filter_simt(const conv3p_geom_t* g, const PlanView& v, const float* grad_out,
                                const float* input, int Cin, int Cout, float* grad_filter,
                                void* scratch, size_t scratch_bytes, cudaStream_t stream);
size_t backward_filter_scratch_bytes(const conv3p_geom_t* g, int Cin, int Cout);

struct GroupItems;   // work-item lists of the tensor-core kernels (below)

// tensor-core (tcgen05, 3xTF32) engine
bool forward_tc_supported(int N, long long capacity, int Cin, int Cout);
bool backward_input_tc_supported(int N, long long capacity, int Cin, int Cout);
int launch_backward_input_tc(const conv3p_geom_t* g, const PlanView& v, const float* grad_out,
                             const float* filter, int Cin, int Cout, float* grad_input, void* scratch,
                             size_t scratch_bytes, cudaStream_t stream, float* g_store = nullptr,
                             const GroupItems* half_items = nullptr);
size_t weight_panel_bytes(int Cin, int Cout);
int launch_prep_weight_panels(const float* filter, void* wp, int Cin, int Cout, int transposed_out,
                              cudaStream_t stream);
int launch_forward_tc(const conv3p_geom_t* g, const PlanView& v, const float* input, const float* filter,
                      int Cin, int Cout, float* output, void* scratch, size_t scratch_bytes,
                      cudaStream_t stream, const RowIO& io = RowIO());

// fused gather + MMA kernel (gather_mma2.cu)
bool gather_mma2_supported(int N, long long capacity, int Csrc, int Nout);
size_t gather_mma2_scratch_bytes(const conv3p_geom_t* g);  // work-item lists of one launch
// Work-item lists of the tensor-core kernels (k_group_items): per sub-tile of `rows` voxel-sorted points and per
// kernel cell, `rows` items (list position, row | members << 8) compacted by population class.
struct GroupItems {
  uint2* items;     // [subtiles][27][rows]
  int* nnz;         // [subtiles][27] non-empty rows of the group
  int* rowid;       // [subtiles*rows] tensor row of every sorted position (-1 pad, -2-row: incomplete list)
  unsigned* mask;   // [subtiles] bit f: group f has members
  unsigned* counter;  // one word, zeroed by k_group_items: chunk counter of the consumer's tile scheduler
  long long subtiles;
};
size_t group_items_bytes(long long pts, int rows);
GroupItems carve_group_items(void* scratch, long long pts, int rows);
// `half` (rows == 128 only): also emit the lists of the 64-row tiles (two per sub-tile) into *half
int launch_group_items(const conv3p_geom_t* g, const PlanView& v, bool backward_lists, int rows,
                       const GroupItems& gi, cudaStream_t stream, const GroupItems* half = nullptr);
int launch_gather_mma2(const conv3p_geom_t* g, const PlanView& v, const float* src, const void* wp, int Csrc,
                       int Nout, float* out, bool weighted, void* scratch, size_t scratch_bytes, const char* name,
                       cudaStream_t stream, float* g_store = nullptr, const RowIO& io = RowIO(),
                       const GroupItems* half_items = nullptr);
// scratch layout of one forward / backward call: [weight panel images | work-item lists | grad_filter partials]
size_t tc_items_bytes(const conv3p_geom_t* g, int Cin, int Cout);

// weight-gradient kernel on tensor cores (backward_filter2.cu)
bool backward_filter2_supported(int N, long long capacity, int Cin, int Cout);
size_t backward_filter2_scratch_bytes(const conv3p_geom_t* g, int Cin, int Cout);
int launch_backward_filter2(const conv3p_geom_t* g, const PlanView& v, const float* grad_out, const float* input,
                            int Cin, int Cout, float* grad_filter, void* scratch, size_t scratch_bytes,
                            cudaStream_t stream, const float* g_store = nullptr, bool items_ready = false);
int backward_filter2_tile_rows(int N, long long capacity, int Cin, int Cout);   // 64 or 32 (0: shape not supported)

// general filter shapes (generic_filter.cu): any fz x fy x fx with at most 512 cells, SIMT, one-shot calls; the same
// kernels instantiated for double serve the reference's T = double registration (every shape, 3x3x3 included)
bool generic_filter_supported(const int dims_zyx[3]);
size_t generic_workspace_bytes(const conv3p_geom_t* g, const int dims_zyx[3], int Cin, int Cout, int elem_bytes = 4);
int generic_forward_f64(const conv3p_geom_t* g, const int dims_zyx[3], double voxel, const double* points,
                        const double* input, const double* filter, int Cin, int Cout, double* output, void* ws,
                        size_t ws_bytes, cudaStream_t stream);
int generic_backward_f64(const conv3p_geom_t* g, const int dims_zyx[3], double voxel, const double* grad_out,
                         const double* points, const double* input, const double* filter, int Cin, int Cout,
                         double* grad_input, double* grad_filter, void* ws, size_t ws_bytes, cudaStream_t stream);
int generic_forward(const conv3p_geom_t* g, const int dims_zyx[3], const float* points, const float* input,
                    const float* filter, int Cin, int Cout, float* output, void* ws, size_t ws_bytes,
                    cudaStream_t stream);
int generic_backward(const conv3p_geom_t* g, const int dims_zyx[3], const float* grad_out, const float* points,
                     const float* input, const float* filter, int Cin, int Cout, float* grad_input,
                     float* grad_filter, void* ws, size_t ws_bytes, cudaStream_t stream);

// warp-per-point fp32 engine for the reference models' small channel counts (3, 9, 13, 36)
bool small_channels_supported(int Cin, int Cout);
bool small_forward_supported(int Cin, int Cout);         // per direction (36->13 forward yes, 13->36 grad_input no)
bool small_backward_input_supported(int Cin, int Cout);
bool small_backward_filter_supported(int Cin, int Cout);
size_t backward_filter_small_scratch_bytes(int Cin, int Cout);
int launch_forward_small(const conv3p_geom_t* g, const PlanView& v, const float* input, const float* filter,
                         int Cin, int Cout, float* output, cudaStream_t stream, const RowIO& io = RowIO());
int launch_backward_input_small(const conv3p_geom_t* g, const PlanView& v, const float* grad_out,
                                const float* filter, int Cin, int Cout, float* grad_input, cudaStream_t stream);
int launch_backward_filter_small(const conv3p_geom_t* g, const PlanView& v, const float* grad_out,
                                 const float* input, int Cin, int Cout, float* grad_filter, void* scratch,
                                 size_t scratch_bytes, cudaStream_t stream);

}  // namespace c3p

#define C3P_CUDA(expr)                                        \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return c3p::cuda_fail(_e, #expr);  \
  } while (0)

#define C3P_LAUNCH_CHECK(name)                                   \
  do {                                                           \
    cudaError_t _e = cudaGetLastError();                         \
    if (_e != cudaSuccess) return c3p::cuda_fail(_e, name);      \
    c3p::count_launch();                                         \
  } while (0)

// ------------------------------------------------------------------------------------------------
// Device helpers: the reference's exact neighbour predicate (tf_conv3p_atrous.cpp:235-290), kept in
// one place so the search, the backward re-binning and the tests of both agree bit for bit.
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
namespace c3p {

// Box bound: formed in double, rounded to float on assignment -- tf_conv3p_atrous.cpp:240-245
// (`T xmin = x - filter_full_x * 0.5 * voxel_size;`, int*double*float).  full*0.5*voxel is exact in
// double (<= 29 significant bits), so FMA contraction cannot change the result.
__device__ __forceinline__ float box_lo(float centre, int full, float voxel) {
  return __double2float_rn((double)centre - ((double)full * 0.5) * (double)voxel);
}
__device__ __forceinline__ float box_hi(float centre, int full, float voxel) {
  return __double2float_rn((double)centre + ((double)full * 0.5) * (double)voxel);
}

// Tap (0..2) of coordinate v in a box starting at lo, or -1 for a dilation hole.  fp32 subtract,
// IEEE fp32 divide (never a reciprocal multiply, never fast-math), truncation, clamp, hole test --
// tf_conv3p_atrous.cpp:280-288.  Negative voxel indices (only reachable in the backward re-binning,
// where the reference would index out of bounds) are reported as holes.
__device__ __forceinline__ int tap_of(float v, float lo, float voxel, int full, int stride) {
  int c = __float2int_rz(__fdiv_rn(__fsub_rn(v, lo), voxel));
  c = min(c, full - 1);
  if (c < 0) return -1;
  int t = c / stride;
  return (t * stride == c) ? t : -1;
}

// Same result, cheaper on average: the voxel index is first formed with a reciprocal multiply; only when that
// quotient lies within a guard band of an integer (where it could truncate differently from the IEEE
// quotient: |q_approx - q_ieee| <= 1.8e-7 * q) is the exact division evaluated.  Bit-identical to tap_of.
__device__ __forceinline__ int tap_of_fast(float v, float lo, float voxel, float inv_voxel, int full,
                                           int stride) {
  const float d = __fsub_rn(v, lo);
  const float qa = d * inv_voxel;
  int c = __float2int_rz(qa);
  const float fr = qa - (float)c;
  const float guard = 1e-4f + fabsf(qa) * 1e-6f;
  if (!(qa >= 0.f) || fr < guard || fr > 1.f - guard) c = __float2int_rz(__fdiv_rn(d, voxel));
  c = min(c, full - 1);
  if (c < 0) return -1;
  int t = c / stride;
  return (t * stride == c) ? t : -1;
}

// Uniform-grid coordinate of v: (int)((v - vmin) / cell) -- tf_conv3p_atrous.cpp:192-194.  Monotone
// non-decreasing in v, which is all the candidate windows rely on.
__device__ __forceinline__ int grid_coord(float v, float vmin, float cell, int dim) {
  int c = __float2int_rz(__fdiv_rn(__fsub_rn(v, vmin), cell));
  return max(0, min(c, dim - 1));
}

// SELU as the reference defines it (selu.py:22-26): scale * (x >= 0 ? x : alpha * (exp(x) - 1)).
__device__ __forceinline__ float selu_f(float x) {
  const float alpha = 1.6732632423543772848170429916717f, scale = 1.0507009873554804934193349852946f;
  return scale * (x >= 0.f ? x : alpha * expm1f(x));
}
__device__ __forceinline__ float apply_activation(float x, int activation) {
  return activation == CONV3P_ACT_SELU ? selu_f(x) : x;
}

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

}  // namespace c3p
#endif
