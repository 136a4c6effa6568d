/* conv3p_b200.h -- C ABI of libconv3p_b200.so: the Conv3p pointwise-convolution operator
 * (forward + Conv3pGrad) for NVIDIA B200 (sm_100a).
 *
 * Drop-in boundary.  These entry points are what a binding of the reference operator would call
 * instead of the reference's TensorFlow kernels:
 *
 *   reference interface (hkust-vgd/pointwise)                      replaced by
 *   -------------------------------------------------------------  ---------------------------
 *   REGISTER_OP("Conv3p")      tf_ops/conv3p/register_op.cpp:44-61  conv3p_op_forward_f32
 *   REGISTER_OP("Conv3pGrad")  tf_ops/conv3p/register_op.cpp:63-75  conv3p_op_backward_f32
 *   Conv3pOp<GPU>::Compute     tf_conv3p_atrous.cu:541-641          conv3p_plan_build_f32 + conv3p_forward_f32
 *   Conv3pGradOp<GPU>::Compute tf_conv3p_atrous.cu:659-774          conv3p_plan_build_backward + conv3p_backward_f32
 *   Grid::build_neighbor_count tf_conv3p_atrous.cu:288-327          conv3p_plan_build_f32 (count table inside the plan)
 *   session.run feed/fetch     train_modelnet40_acsd.py:132         conv3p_host_forward_f32 / conv3p_host_backward_f32
 *
 * Conventions (identical to the reference, tf_conv3p_atrous.cpp:409-444, :490):
 *   points  [B,N,3]   float32 row-major           input   [B,N,Cin]  float32
 *   filter  [3,3,3,Cin,Cout] float32, dims ordered z,y,x; weight index (f*Cin+k)*Cout+c,
 *           f = (fz*3+fy)*3+fx                    output  [B,N,Cout] float32
 *   stride  int[3] in x,y,z order                 voxel_size  float
 * float32 3x3x3 filters (every reference model) run on the tuned engines through the plan calls below; other filter
 * shapes (the reference reads fz, fy, fx from the tensor, tf_conv3p_atrous.cpp:425-427) are served by the one-shot
 * conv3p_op_* calls on a general path (up to 512 cells), and so is the reference's T = double registration
 * (register_op.cpp:45, 64): conv3p_op_*_f64, every tensor in double.
 *
 * Memory: the library never allocates device memory.  The caller owns inputs, outputs, the plan
 * buffer and the scratch buffer (sizes from the *_bytes functions).  Outputs are fully overwritten.
 * Streams: everything is enqueued on the caller's stream; no entry point synchronises the host
 * except conv3p_plan_stats and the conv3p_host_* convenience calls (which say so).
 * Errors: every entry point returns a status code (0 = OK); the library never exits the process
 * (contrast tf_conv3p_atrous.cu:46-54).  All pointers except where noted are DEVICE pointers.
 */
#ifndef CONV3P_B200_H_
#define CONV3P_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* conv3p_stream_t; /* == cudaStream_t */

enum {
  CONV3P_OK = 0,
  CONV3P_ERR_INVALID_ARGUMENT = 1, /* shapes / strides / voxel size rejected */
  CONV3P_ERR_BUFFER_TOO_SMALL = 2, /* plan or scratch buffer smaller than *_bytes() */
  CONV3P_ERR_CUDA = 3,             /* a CUDA runtime call failed; see conv3p_last_cuda_error */
  CONV3P_ERR_UNSUPPORTED = 4,      /* e.g. a filter with more than 512 cells */
  CONV3P_ERR_PAIR_OVERFLOW = 5,    /* reported by conv3p_plan_stats: pair_capacity too small */
  CONV3P_ERR_NO_BACKWARD_LISTS = 6 /* conv3p_backward_f32 on a plan that conv3p_plan_build_f32 built in this process
                                      and conv3p_plan_build_backward has not completed (tracked per plan address) */
};

/* Geometry of one call: B clouds of N points, the dilation stride per axis (x,y,z), the voxel size
 * and the capacity of the neighbour-pair lists.  pair_capacity bounds the TOTAL number of
 * (point, neighbour) pairs over the whole batch; conv3p_plan_stats reports the number needed. */
typedef struct {
  int B;
  int N;
  int stride[3];
  float voxel_size;
  long long pair_capacity;
} conv3p_geom_t;

/* Result of conv3p_plan_stats (host memory). */
typedef struct {
  long long total_pairs;    /* forward (i, j) pairs found == sum of the count table */
  long long backward_pairs; /* pairs kept by the backward rule (0 until build_backward ran) */
  int overflow;             /* 1 if total_pairs > pair_capacity: lists are incomplete, rebuild */
  int has_backward;
} conv3p_plan_stats_t;

/* Byte offsets of the plan's arrays inside the plan buffer (for parity tests and debugging).
 * count_table is exactly the reference's neighbor_count table [B*N, 27] int32
 * (tf_conv3p_atrous.cpp:369-379). */
typedef struct {
  size_t header;       /* int64[16] device-side counters                                   */
  size_t cloud_meta;   /* float[B][8]: vmin xyz, voxel, then dims x,y,z and key bits as int */
  size_t sorted_key;   /* uint32[B*N] voxel keys in sorted order                            */
  size_t sorted_xyzi;  /* float4[B*N]  (x, y, z, bit-cast original index) in sorted order   */
  size_t count_table;  /* int32[B*N][27] by original point                                  */
  size_t pair_begin;   /* int64[B*N] start of each point's list                             */
  size_t pair_len;     /* int32[B*N] forward list length K_i                                */
  size_t pair_row;     /* int32[capacity] neighbour rows b*N+j, grouped by ascending cell f */
  size_t bwd_count;    /* int32[B*N][27] backward list cell counts                          */
  size_t bwd_row;      /* int32[capacity] rows b*N+ii                                       */
  size_t bwd_weight;   /* float[capacity] 1/count(ii, f')                                   */
  size_t sort_tmp;     /* scratch of the radix sort                                         */
  size_t cell_start;   /* uint32[B][cell_cap+1] bin offsets: first sorted position of every voxel-grid cell (lower
                          bound), cell_cap = clamp(16*N, 4096, 65536); clouds with more cells search by bisection */
  size_t total_bytes;
} conv3p_plan_layout_t;

/* ---- neighbour plan: voxel-key radix sort + windowed exact-predicate search ------------------ */

size_t conv3p_plan_bytes(const conv3p_geom_t* geom);
int conv3p_plan_layout(const conv3p_geom_t* geom, conv3p_plan_layout_t* out /* host */);

/* Builds the forward neighbour structure of `points` into `plan` (device, >= conv3p_plan_bytes):
 * per-cloud bounding box, voxel keys, radix sort, per-point windowed search with the reference's
 * exact predicate (tf_conv3p_atrous.cpp:232-301), count table and cell-grouped neighbour lists. */
int conv3p_plan_build_f32(const conv3p_geom_t* geom, const float* points, void* plan,
                          size_t plan_bytes, conv3p_stream_t stream);

/* Adds the backward lists: for every j and every ii in N(j), the cell f' of j in ii's frame without
 * a box test, dropped when it is a hole or count(ii, f') == 0 (tf_conv3p_atrous.cpp:654-679). */
int conv3p_plan_build_backward(const conv3p_geom_t* geom, const float* points, void* plan,
                               size_t plan_bytes, conv3p_stream_t stream);

/* Copies the plan's counters to the host.  SYNCHRONISES `stream`. */
int conv3p_plan_stats(const conv3p_geom_t* geom, const void* plan, conv3p_plan_stats_t* out /* host */,
                      conv3p_stream_t stream);

/* Asynchronous variant: enqueues a one-warp kernel that stores the plan's 16 counters (int64: [0] total pairs,
 * [1] overflow flag, [2] backward pairs, [3] has_backward) into `host_mapped16`, which must be page-locked host
 * memory that the device can address (cudaHostAlloc / cudaMallocHost under unified addressing).  No copy engine, no
 * synchronisation: the caller waits on an event recorded after this call before reading the values. */
int conv3p_plan_publish_stats(const conv3p_geom_t* geom, const void* plan, long long* host_mapped16,
                              conv3p_stream_t stream);

/* ---- the operator on a built plan ------------------------------------------------------------- */

size_t conv3p_scratch_bytes(const conv3p_geom_t* geom, int Cin, int Cout);
/* >= conv3p_scratch_bytes.  A backward call given this much scratch lets the grad_input kernel leave its
 * per-(point, cell) aggregates of grad_output in the scratch tail (27*Cout floats per point) for the grad_filter
 * kernel, which then does not walk the backward lists a second time; with only conv3p_scratch_bytes the result
 * is the same, the weight gradient just gathers on its own. */
size_t conv3p_backward_scratch_bytes(const conv3p_geom_t* geom, int Cin, int Cout);

/* output[B,N,Cout] = Conv3p(points, input, filter); follows tf_conv3p_atrous.cpp:456-504. */
int conv3p_forward_f32(const conv3p_geom_t* geom, const void* plan, const float* input,
                       const float* filter, int Cin, int Cout, float* output, void* scratch,
                       size_t scratch_bytes, conv3p_stream_t stream);

/* Forward with the epilogue the reference's networks apply right after every Conv3p, fused:
 *   - `activation` = CONV3P_ACT_SELU stores selu(Conv3p(...)) -- `selu(conv3p(...))` at
 *     scene_seg/pointcnn_scene_seg_acsd.py:35-36, selu.py:22-26 -- instead of a second pass over the output;
 *   - rows of `input` and `output` may sit inside wider row-major buffers (`*_row_stride` floats from one row to
 *     the next, >= the channel count; 0 = dense), so the four 9-channel layers can write straight into the
 *     [B,N,36] concat buffer (:56) and read their predecessor's slice of it.
 * The tensor-core engine needs 16-byte aligned rows (strides % 4 == 0, aligned base pointers); other layouts run on
 * the fp32 engines.  Everything else as conv3p_forward_f32. */
enum { CONV3P_ACT_NONE = 0, CONV3P_ACT_SELU = 1 };
int conv3p_forward_ex_f32(const conv3p_geom_t* geom, const void* plan, const float* input,
                          long long input_row_stride, const float* filter, int Cin, int Cout, float* output,
                          long long output_row_stride, int activation, void* scratch, size_t scratch_bytes,
                          conv3p_stream_t stream);

/* out[r, c] = grad[r, c] * selu'(x) expressed through the ACTIVATED value y = selu(x) the forward stored:
 * scale for y >= 0 (the reference's SELU takes the linear branch for x >= 0, selu.py:25), y + scale*alpha otherwise.  y and grad rows may be strided (floats; 0 = dense); out is dense
 * [rows, C].  The backward of the fused epilogue: feed `out` to conv3p_backward_f32 as grad_output. */
int conv3p_selu_backward_f32(const float* y, long long y_row_stride, const float* grad, long long grad_row_stride,
                             float* out, long long rows, int C, conv3p_stream_t stream);

/* grad_input[B,N,Cin], grad_filter[27,Cin,Cout]; follows tf_conv3p_atrous.cpp:622-716.
 * Either output pointer may be NULL to skip it.  grad_filter is reduced in a fixed order
 * (deterministic; the reference's OpenMP/atomicAdd reductions are not). */
int conv3p_backward_f32(const conv3p_geom_t* geom, const void* plan, const float* grad_output,
                        const float* input, const float* filter, int Cin, int Cout,
                        float* grad_input, float* grad_filter, void* scratch, size_t scratch_bytes,
                        conv3p_stream_t stream);

/* ---- input pipeline on the GPU (reference: modelnet_provider.py:23-75, util.py:55-109) -------- */
/* out[b,i,:] = float32(clip(sigma * noise[b,i,:], -clip, clip) + float32(data[b,i,:] @ R_y(angles[b]))) --
 * rotate_point_cloud followed by jitter_point_cloud on [B,N,3] clouds.  The random draws are inputs (float64, as
 * numpy produces them): one angle per cloud, one standard-normal sample per coordinate; either may be NULL to skip
 * that stage.  Arithmetic in double, rounded to float32 once per stage like the provider's float32 arrays. */
int conv3p_augment_rotate_jitter_f32(const float* data, const double* angles, const double* noise, double sigma,
                                     double clip, int B, int N, float* out, conv3p_stream_t stream);

/* sort_point_cloud_xyz / sort_point_cloud_xyz2: every cloud's rows ordered by x, then y, then z (the first three of
 * K channels), ties by original position.  order[B,N] receives the source row of every output row; sorted_data
 * [B,N,K] and sorted_attributes [B,N,M] (optional, may be NULL) the permuted rows. */
size_t conv3p_xyz_sort_workspace_bytes(int B, int N);
int conv3p_xyz_sort_f32(const float* data, int K, const float* attributes, int M, int B, int N, int* order,
                        float* sorted_data, float* sorted_attributes, void* workspace, size_t workspace_bytes,
                        conv3p_stream_t stream);

/* ---- one-shot operator calls with the reference's Compute() shape ----------------------------- */
/* workspace >= conv3p_op_workspace_bytes(); holds plan + scratch.  filter_dims = {fz,fy,fx}. */
size_t conv3p_op_workspace_bytes(const conv3p_geom_t* geom, int Cin, int Cout);
/* >= conv3p_op_workspace_bytes: with this much workspace conv3p_op_backward_f32 (and conv3p_host_backward_f32 for
 * its device part) shares one gather between the two gradients, see conv3p_backward_scratch_bytes. */
size_t conv3p_op_backward_workspace_bytes(const conv3p_geom_t* geom, int Cin, int Cout);

/* Workspace of a one-shot call for ANY supported filter shape (filter_dims = {fz,fy,fx}); backward != 0 sizes it for
 * conv3p_op_backward_f32 (3x3x3: with the shared gather).  0 = unsupported shape / invalid geometry. */
size_t conv3p_op_workspace_bytes_ex(const conv3p_geom_t* geom, const int filter_dims[3], int Cin, int Cout,
                                    int backward);

int conv3p_op_forward_f32(const float* points, const float* input, const float* filter,
                          const int filter_dims[3], const int stride_xyz[3], float voxel_size, int B,
                          int N, int Cin, int Cout, long long pair_capacity, float* output,
                          void* workspace, size_t workspace_bytes, conv3p_stream_t stream);

int conv3p_op_backward_f32(const float* grad_output, const float* points, const float* input,
                           const float* filter, const int filter_dims[3], const int stride_xyz[3],
                           float voxel_size, int B, int N, int Cin, int Cout,
                           long long pair_capacity, float* grad_input, float* grad_filter,
                           void* workspace, size_t workspace_bytes, conv3p_stream_t stream);

/* ---- T = double ----------------------------------------------------------------------------------
 * The reference registers Conv3p / Conv3pGrad for double as well (register_op.cpp:45, 64; CPU kernels
 * tf_conv3p_atrous.cpp:516, 727): points, input, filter, voxel size and the outputs in double, the whole neighbour
 * predicate evaluated in double (:239-290, :654-679).  One-shot calls, every filter shape up to 512 cells (3x3x3
 * included), fp64 SIMT; same status codes and NaN poisoning on pair overflow as the float calls.  `geom` of the
 * workspace query carries (float)voxel_size. */
size_t conv3p_op_workspace_bytes_f64(const conv3p_geom_t* geom, const int filter_dims[3], int Cin, int Cout);
int conv3p_op_forward_f64(const double* points, const double* input, const double* filter, const int filter_dims[3],
                          const int stride_xyz[3], double voxel_size, int B, int N, int Cin, int Cout,
                          long long pair_capacity, double* output, void* workspace, size_t workspace_bytes,
                          conv3p_stream_t stream);
int conv3p_op_backward_f64(const double* grad_output, const double* points, const double* input, const double* filter,
                           const int filter_dims[3], const int stride_xyz[3], double voxel_size, int B, int N, int Cin,
                           int Cout, long long pair_capacity, double* grad_input, double* grad_filter, void* workspace,
                           size_t workspace_bytes, conv3p_stream_t stream);

/* ---- host-buffer calls (the reference's feed/fetch shape) -------------------------------------- */
/* All tensor pointers are HOST pointers (pinned memory gives asynchronous copies); `workspace` is
 * DEVICE memory >= conv3p_host_workspace_bytes().  Copies inputs host->device, runs the operator,
 * copies results device->host and SYNCHRONISES `stream` before returning. */
size_t conv3p_host_workspace_bytes(const conv3p_geom_t* geom, int Cin, int Cout);

int conv3p_host_forward_f32(const float* h_points, const float* h_input, const float* h_filter,
                            const int stride_xyz[3], float voxel_size, int B, int N, int Cin,
                            int Cout, long long pair_capacity, float* h_output, void* workspace,
                            size_t workspace_bytes, conv3p_stream_t stream);

int conv3p_host_backward_f32(const float* h_grad_output, const float* h_points,
                             const float* h_input, const float* h_filter, const int stride_xyz[3],
                             float voxel_size, int B, int N, int Cin, int Cout,
                             long long pair_capacity, float* h_grad_input, float* h_grad_filter,
                             void* workspace, size_t workspace_bytes, conv3p_stream_t stream);

/* ---- diagnostics -------------------------------------------------------------------------------- */
const char* conv3p_status_string(int status);
const char* conv3p_last_cuda_error(void); /* thread-local text of the last CONV3P_ERR_CUDA */
int conv3p_abi_version(void);
/* Number of kernels this library launched (all threads of the process) since the last reset. */
long long conv3p_launch_count(int reset);
/* Selects the contraction engine.  Low three bits: 0 = auto (tensor cores where the shape is a real dense GEMM, else
 * the fp32 engines), 1 = fp32 SIMT only (warp-per-point or tile kernels by channel count), 2 = tensor cores (3xTF32)
 * where supported, 3 = generic fp32 tile kernels only (no tensor cores, no warp-per-point kernels).  Higher bits are
 * ablation flags for A/B timing (tools/engine_timing.py, tools/ab_backward.py): 32 = phase timers of the gather+MMA
 * kernel, 256 = no G store shared between the two gradient kernels, 512 = three TF32 products (3xTF32) instead of
 * one TF32 product + BF16 correction products in the tensor-core kernels, 1024 = first version of the small-channel
 * kernels, 2048 = no channel padding onto the tensor-core kernels (36->13 and the like stay on the fp32 engines).  Returns the previous value.  Process-wide (a test/benchmark knob, not part of the operator's state). */
int conv3p_set_engine(int engine);

/* Profiling helpers of tools/engine_timing.py: read and clear the device-side cycle counters the gather+MMA kernel
 * accumulates while engine flag 32 is set (8 phase sums; sum and max of a CTA's total cycles).  HOST pointers.
 * Synchronise the device. */
int conv3p_debug_phase_cycles(unsigned long long* host8);
int conv3p_debug_cta_cycles(unsigned long long* host2);
/* Same for the weight-gradient kernel (library variants built with -DC3P_W2_TIMED=1; zeros otherwise): 16 counters
 * summed over all CTAs -- [0..7] phases of producer warp 0, [8..9] item loader, [10..13] MMA issuer. */
int conv3p_debug_w2_cycles(unsigned long long* host16);

/* Per-kernel timing for benchmarks: while enabled, every kernel launch is bracketed by CUDA events
 * on its stream.  conv3p_profile_enable(on) clears the records and returns the previous state.
 * conv3p_profile_read writes "kernel_name launches total_ms\n" lines into buf (after the caller
 * synchronised) and returns the number of launches recorded. */
int conv3p_profile_enable(int on);
long long conv3p_profile_read(char* buf, size_t cap);

/* Self-test of the tensor-core plumbing: D[128,N] = A[128,K] * B[N,K]^T on one CTA with tcgen05.  split: 0 = plain
 * TF32, 1 = 3xTF32 (hi/lo split, three TF32 products), 1|8 = the production split (TF32 product of the rounded hi
 * parts + one BF16 chain for the two correction terms).  N % 16 == 0, 16 <= N <= 256, K % 32 == 0.  Device pointers. */
int conv3p_selftest_tc(const float* A, const float* B, float* D, int N, int K, int split,
                       conv3p_stream_t stream);
/* Same with the contraction index outermost in memory (MN-major operands, as in the weight-gradient
 * kernel): D[128,N] = A^T * B, A[K,128], B[K,N]; N % 32 == 0, K % 8 == 0, K <= 64. */
int conv3p_selftest_tc_mn(const float* A, const float* B, float* D, int N, int K, int split,
                          conv3p_stream_t stream);

/* Issue-rate probe of the tensor core (tools/mma_rate.py): one CTA issues reps x 8 tcgen05.mma with M = 128 and N
 * columns and writes the elapsed SM cycles to cycles_device[0].  mode bit 0: MN-major operands (else K-major), bit 1:
 * BF16 kind::f16 with K = 16 per instruction (else kind::tf32, K = 8). */
int conv3p_debug_mma_rate(int N, int mode, int reps, unsigned long long* cycles_device, conv3p_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CONV3P_B200_H_ */
