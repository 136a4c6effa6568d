// oracle/ref_runner_gpu.cu -- TEST / BENCHMARK INFRASTRUCTURE, not product code.
//
// C-ABI runner around the reference's own GPU op: /root/reference/tf_ops/conv3p/tf_conv3p_atrous.cu is pulled in
// UNMODIFIED through REF_SRC_GPU (oracle/Makefile, target refgpu) and compiled for sm_100a with the reference's
// own nvcc flags (tf_conv3p_compile.sh:33, incl. -use_fast_math) against oracle/tf_shim.  It is the "reference GPU
// code on the same box" number of BASELINE.md section 2 -- a brute-force O(N^2) neighbour sweep with one thread per
// point (tf_conv3p_atrous.cu:335-533).  Because of -use_fast_math it is NOT a parity reference (hard part 1):
// bench.py times it, nothing is checked against it.  All pointers are DEVICE pointers (stride and voxel too: the op
// copies them back itself, :577, :586).
#ifndef REF_SRC_GPU
#error "REF_SRC_GPU must be the quoted path of the reference tf_conv3p_atrous.cu"
#endif
#include REF_SRC_GPU
#undef min
#undef max

#include <memory>
#include <string>

namespace {
std::string g_last_error;
template <typename T>
tensorflow::Tensor wrap(const T* p, std::initializer_list<tensorflow::int64> dims) {
  return tensorflow::Tensor(tensorflow::TensorShape(dims), const_cast<T*>(p));
}
int finish(const tensorflow::OpKernelContext& ctx) {
  if (ctx.status.ok()) return cudaDeviceSynchronize() == cudaSuccess ? 0 : 3;
  g_last_error = ctx.status.error_message();
  return 1;
}
}  // namespace

extern "C" {

const char* refgpu_last_error() { return g_last_error.c_str(); }

int refgpu_conv3p_forward_f32(const float* points, const float* input, const float* filter, const int* stride,
                              const float* voxel, int B, int N, int Cin, int Cout, float* output) {
  using namespace tensorflow;
  std::unique_ptr<OpKernel> k(CreateKernel<float>("Conv3p", DEVICE_GPU));
  if (!k) return 2;
  OpKernelContext ctx;
  ctx.elem_bytes = sizeof(float);
  ctx.inputs.push_back(wrap(points, {B, N, 3}));
  ctx.inputs.push_back(wrap(input, {B, N, Cin}));
  ctx.inputs.push_back(wrap(filter, {3, 3, 3, Cin, Cout}));
  ctx.inputs.push_back(wrap(stride, {3}));
  ctx.inputs.push_back(wrap(voxel, {1}));
  ctx.out_buffers.push_back(output);
  k->Compute(&ctx);
  return finish(ctx);
}

int refgpu_conv3p_backward_f32(const float* grad_out, const float* points, const float* input, const float* filter,
                               const int* stride, const float* voxel, int B, int N, int Cin, int Cout,
                               float* grad_input, float* grad_filter) {
  using namespace tensorflow;
  std::unique_ptr<OpKernel> k(CreateKernel<float>("Conv3pGrad", DEVICE_GPU));
  if (!k) return 2;
  OpKernelContext ctx;
  ctx.elem_bytes = sizeof(float);
  ctx.inputs.push_back(wrap(grad_out, {B, N, Cout}));
  ctx.inputs.push_back(wrap(points, {B, N, 3}));
  ctx.inputs.push_back(wrap(input, {B, N, Cin}));
  ctx.inputs.push_back(wrap(filter, {3, 3, 3, Cin, Cout}));
  ctx.inputs.push_back(wrap(stride, {3}));
  ctx.inputs.push_back(wrap(voxel, {1}));
  ctx.out_buffers.push_back(grad_input);
  ctx.out_buffers.push_back(grad_filter);
  k->Compute(&ctx);
  return finish(ctx);
}

}  // extern "C"
