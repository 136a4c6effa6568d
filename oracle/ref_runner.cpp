// oracle/ref_runner.cpp -- TEST INFRASTRUCTURE, not product code.
//
// C-ABI runner around the reference's own CPU op.  The reference translation unit
// (/root/reference/tf_ops/conv3p/tf_conv3p_atrous.cpp) is pulled in UNMODIFIED through
// REF_SRC (set by oracle/Makefile) and compiled against oracle/tf_shim, so Conv3pOp<CPU>
// (tf_conv3p_atrous.cpp:395-517), Conv3pGradOp<CPU> (:520-728) and Grid<CpuAlloc,T>
// (:138-388) below are the reference's object code.  Nothing here restates arithmetic.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
// may load the resulting oracle/_ref/*.so.
#ifndef REF_SRC
#error "REF_SRC must be the quoted path of the reference tf_conv3p_atrous.cpp"
#endif
#include REF_SRC

#include <string>

namespace {
std::string g_last_error;

template <typename T>
tensorflow::Tensor wrap(const T* p, std::initializer_list<tensorflow::int64> dims) {
  return tensorflow::Tensor(tensorflow::TensorShape(dims), const_cast<T*>(p));
}

int finish(const tensorflow::OpKernelContext& ctx) {
  if (ctx.status.ok()) return 0;
  g_last_error = ctx.status.error_message();
  return 1;
}
}  // namespace

extern "C" {

const char* ref_last_error() { return g_last_error.c_str(); }

int ref_num_threads() {
#ifdef CONV_OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int ref_hardware_concurrency() { return (int)std::thread::hardware_concurrency(); }

// Inputs in schema order (register_op.cpp:46-50).  `points_rank3` = 0 lets tests provoke the
// rank check (tf_conv3p_atrous.cpp:410).  n_stride / n_voxel let tests provoke :437 / :443.
int ref_conv3p_forward_f32(const float* points, const float* input, const float* filter,
                           const int* stride, int n_stride, const float* voxel, int n_voxel,
                           int B, int N, int Cin, int Cout, int fz, int fy, int fx,
                           int input_B, int input_N, int filter_Cin, int points_rank3,
                           float* output) {
  using namespace tensorflow;
  std::unique_ptr<OpKernel> k(CreateKernel<float>("Conv3p", DEVICE_CPU));
  if (!k) {
    g_last_error = "Conv3p/CPU/float not registered";
    return 2;
  }
  OpKernelContext ctx;
  ctx.elem_bytes = sizeof(float);
  if (points_rank3)
    ctx.inputs.push_back(wrap(points, {B, N, 3}));
  else
    ctx.inputs.push_back(wrap(points, {(int64)B * N, 3}));
  ctx.inputs.push_back(wrap(input, {input_B, input_N, Cin}));
  ctx.inputs.push_back(wrap(filter, {fz, fy, fx, filter_Cin, Cout}));
  ctx.inputs.push_back(wrap(stride, {n_stride}));
  ctx.inputs.push_back(wrap(voxel, {n_voxel}));
  ctx.out_buffers.push_back(output);
  k->Compute(&ctx);
  return finish(ctx);
}

// Inputs in schema order (register_op.cpp:65-70).
int ref_conv3p_backward_f32(const float* grad_out, const float* points, const float* input,
                            const float* filter, const int* stride, const float* voxel, int B,
                            int N, int Cin, int Cout, int fz, int fy, int fx, int grad_B,
                            int grad_N, int grad_C, float* grad_input, float* grad_filter) {
  using namespace tensorflow;
  std::unique_ptr<OpKernel> k(CreateKernel<float>("Conv3pGrad", DEVICE_CPU));
  if (!k) {
    g_last_error = "Conv3pGrad/CPU/float not registered";
    return 2;
  }
  OpKernelContext ctx;
  ctx.elem_bytes = sizeof(float);
  ctx.inputs.push_back(wrap(grad_out, {grad_B, grad_N, grad_C}));
  ctx.inputs.push_back(wrap(points, {B, N, 3}));
  ctx.inputs.push_back(wrap(input, {B, N, Cin}));
  ctx.inputs.push_back(wrap(filter, {fz, fy, fx, Cin, Cout}));
  ctx.inputs.push_back(wrap(stride, {3}));
  ctx.inputs.push_back(wrap(voxel, {1}));
  ctx.out_buffers.push_back(grad_input);
  ctx.out_buffers.push_back(grad_filter);
  k->Compute(&ctx);
  return finish(ctx);
}

// Count table [N, fz*fy*fx] of ONE cloud through Grid::neighbor_count (tf_conv3p_atrous.cpp:369-379).
void ref_neighbor_count_f32(const float* points, int N, int fz, int fy, int fx, const int* stride,
                            float voxel, int* count) {
  Grid<CpuAlloc, float> grid(Array<CpuAlloc, float>(const_cast<float*>(points), N), voxel);
  Array<CpuAlloc, int> cnt(count, N * fz * fy * fx);
  grid.neighbor_count(fx, fy, fz, stride[0], stride[1], stride[2], voxel, cnt);
}

// Per-point (j, f) lists of ONE cloud through Grid::neighbor (tf_conv3p_atrous.cpp:232-301), in the
// reference's own emission order.  off has N+1 entries.  Returns the total number of pairs; pairs
// beyond `capacity` are counted but not stored.
long long ref_neighbors_f32(const float* points, int N, int fz, int fy, int fx, const int* stride,
                            float voxel, long long* off, int* nbr_j, int* nbr_f,
                            long long capacity) {
  Grid<CpuAlloc, float> grid(Array<CpuAlloc, float>(const_cast<float*>(points), N), voxel);
  Array<CpuAlloc, int> point_index, filter_cell, filter_cell_count;
  point_index.alloc(N);
  filter_cell.alloc(N);
  filter_cell_count.resize(fz * fy * fx);
  long long total = 0;
  for (int i = 0; i < N; ++i) {
    off[i] = total;
    grid.neighbor(points[3 * i], points[3 * i + 1], points[3 * i + 2], fx, fy, fz, stride[0],
                  stride[1], stride[2], voxel, point_index, filter_cell, filter_cell_count);
    for (int k = 0; k < point_index.size; ++k) {
      if (total < capacity) {
        nbr_j[total] = point_index[k];
        nbr_f[total] = filter_cell[k];
      }
      ++total;
    }
  }
  off[N] = total;
  point_index.free();
  filter_cell.free();
  filter_cell_count.free();
  return total;
}

// ---- T = double: the reference registers Conv3p / Conv3pGrad for double as well (register_op.cpp:45, 64;
// tf_conv3p_atrous.cpp:516, 727).  Same schema order, every floating-point tensor in double. ----------------------
int ref_conv3p_forward_f64(const double* points, const double* input, const double* filter, const int* stride,
                           const double* voxel, int B, int N, int Cin, int Cout, int fz, int fy, int fx,
                           double* output) {
  using namespace tensorflow;
  std::unique_ptr<OpKernel> k(CreateKernel<double>("Conv3p", DEVICE_CPU));
  if (!k) {
    g_last_error = "Conv3p/CPU/double not registered";
    return 2;
  }
  OpKernelContext ctx;
  ctx.elem_bytes = sizeof(double);
  ctx.inputs.push_back(wrap(points, {B, N, 3}));
  ctx.inputs.push_back(wrap(input, {B, N, Cin}));
  ctx.inputs.push_back(wrap(filter, {fz, fy, fx, Cin, Cout}));
  ctx.inputs.push_back(wrap(stride, {3}));
  ctx.inputs.push_back(wrap(voxel, {1}));
  ctx.out_buffers.push_back(output);
  k->Compute(&ctx);
  return finish(ctx);
}

int ref_conv3p_backward_f64(const double* grad_out, const double* points, const double* input, const double* filter,
                            const int* stride, const double* voxel, int B, int N, int Cin, int Cout, int fz, int fy,
                            int fx, double* grad_input, double* grad_filter) {
  using namespace tensorflow;
  std::unique_ptr<OpKernel> k(CreateKernel<double>("Conv3pGrad", DEVICE_CPU));
  if (!k) {
    g_last_error = "Conv3pGrad/CPU/double not registered";
    return 2;
  }
  OpKernelContext ctx;
  ctx.elem_bytes = sizeof(double);
  ctx.inputs.push_back(wrap(grad_out, {B, N, Cout}));
  ctx.inputs.push_back(wrap(points, {B, N, 3}));
  ctx.inputs.push_back(wrap(input, {B, N, Cin}));
  ctx.inputs.push_back(wrap(filter, {fz, fy, fx, Cin, Cout}));
  ctx.inputs.push_back(wrap(stride, {3}));
  ctx.inputs.push_back(wrap(voxel, {1}));
  ctx.out_buffers.push_back(grad_input);
  ctx.out_buffers.push_back(grad_filter);
  k->Compute(&ctx);
  return finish(ctx);
}

// Count table of ONE cloud through Grid<CpuAlloc, double>::neighbor_count.
void ref_neighbor_count_f64(const double* points, int N, int fz, int fy, int fx, const int* stride, double voxel,
                            int* count) {
  Grid<CpuAlloc, double> grid(Array<CpuAlloc, double>(const_cast<double*>(points), N), voxel);
  Array<CpuAlloc, int> cnt(count, N * fz * fy * fx);
  grid.neighbor_count(fx, fy, fz, stride[0], stride[1], stride[2], voxel, cnt);
}

}  // extern "C"
