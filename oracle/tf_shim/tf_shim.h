// oracle/tf_shim/tf_shim.h -- TEST INFRASTRUCTURE, not product code.
//
// A minimal stand-in for the TensorFlow 1.x framework headers, just large enough to
// compile the reference CPU op (tf_ops/conv3p/tf_conv3p_atrous.cpp) UNMODIFIED where it
// lies under /root/reference.  It supplies tensor plumbing only -- every arithmetic
// instruction of the oracle is the reference's own object code.
//
// TensorFlow surface used by the reference (tf_conv3p_atrous.cpp):
//   context->input(i), Tensor::dims/shape().dim_size/flat<T>()          :409-444
//   context->allocate_output(idx, shape, &Tensor*)                       :448, :578, :588
//   OP_REQUIRES / OP_REQUIRES_OK / errors::InvalidArgument               :410-443, :583-585
//   OpKernel, OpKernelConstruction, REGISTER_KERNEL_BUILDER, TF_CALL_*   :396-401, :511-517
#pragma once
#ifdef TF_SHIM_GPU
#include <cuda_runtime.h>
#endif
#include <cassert>
#include <cstdint>
#include <cstring>
#include <initializer_list>
#include <map>
#include <memory>
#include <string>
#include <typeinfo>
#include <vector>

namespace Eigen {
struct ThreadPoolDevice {};
struct GpuDevice {};
}  // namespace Eigen

namespace tensorflow {

typedef long long int64;

class Status {
 public:
  Status() : ok_(true) {}
  explicit Status(const std::string& m) : ok_(false), msg_(m) {}
  bool ok() const { return ok_; }
  const std::string& error_message() const { return msg_; }
  static Status OK() { return Status(); }

 private:
  bool ok_;
  std::string msg_;
};

namespace errors {
inline Status InvalidArgument(const char* m) { return Status(std::string(m)); }
inline Status InvalidArgument(const std::string& m) { return Status(m); }
}  // namespace errors

class TensorShape {
 public:
  TensorShape() {}
  TensorShape(std::initializer_list<int64> d) : d_(d) {}
  explicit TensorShape(const std::vector<int64>& d) : d_(d) {}
  int dims() const { return (int)d_.size(); }
  int64 dim_size(int i) const { return d_[i]; }
  int64 num_elements() const {
    int64 n = 1;
    for (size_t i = 0; i < d_.size(); ++i) n *= d_[i];
    return n;
  }

 private:
  std::vector<int64> d_;
};

template <typename T>
struct Flat {
  T* p;
  int64 n;
  T& operator()(int64 i) const { return p[i]; }
  int64 size() const { return n; }
};

// A tensor either wraps caller memory (inputs, and outputs handed in by the runner) or owns it.
class Tensor {
 public:
  Tensor() : data_(nullptr) {}
  Tensor(const TensorShape& s, void* external) : shape_(s), data_(external) {}
  Tensor(const TensorShape& s, size_t elem_bytes) : shape_(s) {
    own_.reset(new std::vector<unsigned char>((size_t)s.num_elements() * elem_bytes + 16));
    data_ = own_->data();
  }
  int dims() const { return shape_.dims(); }
  const TensorShape& shape() const { return shape_; }
  template <typename T>
  Flat<T> flat() const {
    Flat<T> f;
    f.p = reinterpret_cast<T*>(data_);
    f.n = shape_.num_elements();
    return f;
  }
  void* raw() const { return data_; }

 private:
  TensorShape shape_;
  void* data_;
  std::shared_ptr<std::vector<unsigned char> > own_;
};

class OpKernelConstruction {};

// The reference's GPU op (tf_conv3p_atrous.cu:97-106) takes its scratch memory from
// context->allocate_temp(DataTypeToEnum<double>::value, shape, &tensor); the shim serves it with cudaMalloc
// (TF_SHIM_GPU builds only, compiled by nvcc) and frees it when the context dies.
enum DataType { DT_FLOAT = 1, DT_DOUBLE = 2 };
template <typename T>
struct DataTypeToEnum {
  static const DataType value = DT_DOUBLE;
};

class OpKernelContext {
 public:
  std::vector<Tensor> inputs;
  // Output i is written into out_buffers[i] when the runner provides one (so results land in
  // caller memory); otherwise the shim allocates.  elem_bytes is sizeof(T) of the op instance.
  std::vector<void*> out_buffers;
  std::vector<std::unique_ptr<Tensor> > outputs;
  size_t elem_bytes;
  Status status;

  OpKernelContext() : elem_bytes(4) {}
  const Tensor& input(int i) const { return inputs[i]; }
  Status allocate_output(int idx, const TensorShape& shape, Tensor** out) {
    if ((int)outputs.size() <= idx) outputs.resize(idx + 1);
    if (idx < (int)out_buffers.size() && out_buffers[idx])
      outputs[idx].reset(new Tensor(shape, out_buffers[idx]));
    else
      outputs[idx].reset(new Tensor(shape, elem_bytes));
    *out = outputs[idx].get();
    return Status::OK();
  }
  void SetStatus(const Status& s) { status = s; }
#ifdef TF_SHIM_GPU
  std::vector<void*> temps;
  Status allocate_temp(DataType, const TensorShape& shape, Tensor* out) {
    void* p = nullptr;
    if (cudaMalloc(&p, (size_t)shape.num_elements() * 8 + 16) != cudaSuccess) return Status("cudaMalloc failed");
    temps.push_back(p);
    *out = Tensor(shape, p);
    return Status::OK();
  }
  ~OpKernelContext() {
    for (size_t i = 0; i < temps.size(); ++i) cudaFree(temps[i]);
  }
#endif
};

class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction*) {}
  virtual ~OpKernel() {}
  virtual void Compute(OpKernelContext* context) = 0;
};

#define OP_REQUIRES(CTX, EXP, STATUS) \
  do {                                \
    if (!(EXP)) {                     \
      (CTX)->SetStatus(STATUS);       \
      return;                         \
    }                                 \
  } while (0)

#define OP_REQUIRES_OK(CTX, ...)             \
  do {                                       \
    ::tensorflow::Status _s(__VA_ARGS__);    \
    if (!_s.ok()) {                          \
      (CTX)->SetStatus(_s);                  \
      return;                                \
    }                                        \
  } while (0)

// ---- kernel registry -------------------------------------------------------------------
static const char* const DEVICE_CPU = "CPU";
static const char* const DEVICE_GPU = "GPU";

struct KernelDef {
  std::string op, device, type;
};

class Name {
 public:
  explicit Name(const char* op) { def_.op = op; }
  Name& Device(const char* d) {
    def_.device = d;
    return *this;
  }
  template <typename T>
  Name& TypeConstraint(const char*) {
    def_.type = typeid(T).name();
    return *this;
  }
  const KernelDef& def() const { return def_; }

 private:
  KernelDef def_;
};

typedef OpKernel* (*KernelFactory)(OpKernelConstruction*);

inline std::map<std::string, KernelFactory>& KernelRegistry() {
  static std::map<std::string, KernelFactory> r;
  return r;
}

struct KernelRegistrar {
  KernelRegistrar(const Name& n, KernelFactory f) {
    const KernelDef& d = n.def();
    KernelRegistry()[d.op + "/" + d.device + "/" + d.type] = f;
  }
};

template <typename T>
inline OpKernel* CreateKernel(const char* op, const char* device) {
  std::string key = std::string(op) + "/" + device + "/" + typeid(T).name();
  std::map<std::string, KernelFactory>::iterator it = KernelRegistry().find(key);
  if (it == KernelRegistry().end()) return nullptr;
  OpKernelConstruction c;
  return it->second(&c);
}

#define TF_SHIM_CAT2(a, b) a##b
#define TF_SHIM_CAT(a, b) TF_SHIM_CAT2(a, b)
// The class name may contain a comma (Conv3pOp<CPUDevice, T>), hence variadic.
#define REGISTER_KERNEL_BUILDER(BUILDER, ...)                                            \
  static ::tensorflow::KernelRegistrar TF_SHIM_CAT(tf_shim_registrar_, __COUNTER__)(     \
      ::tensorflow::BUILDER,                                                             \
      [](::tensorflow::OpKernelConstruction* c) -> ::tensorflow::OpKernel* {             \
        return new __VA_ARGS__(c);                                                       \
      })

#define TF_CALL_float(m) m(float)
#define TF_CALL_double(m) m(double)

}  // namespace tensorflow
