// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Minimal stand-in for the slice of the TensorFlow 1.2 C++ op API that
// /root/reference/rendering_layer/ops_src/render_depth_op.{h,cc} touches, so that
// the UNMODIFIED reference translation unit compiles where it lies (TensorFlow and
// Eigen are not installable offline).  Nothing here restates reference logic: it only
// provides row-major tensor views, an op/kernel registry and a context object, so the
// reference's own RenderDepthOp<CPUDevice>::Compute / RenderDepthOpGrad<CPUDevice>::Compute
// (render_depth_op.cc:378-458, 470-528) and its shape functions (:535-589) can be driven
// from oracle/ref_harness.cc.
#ifndef FR_ORACLE_TF_MIN_H_
#define FR_ORACLE_TF_MIN_H_

// Pull in every std header the reference TU (or this shim) needs BEFORE the reference
// header defines its function-like min/max macros (render_depth_op.h:15-16).
#include <cmath>
#include <math.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <initializer_list>
#include <map>
#include <memory>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace Eigen {
struct ThreadPoolDevice {};
struct GpuDevice {};
}  // namespace Eigen

namespace tensorflow {

typedef long long int64;

// ---------------------------------------------------------------- status / errors
class Status {
 public:
  Status() : ok_(true) {}
  explicit Status(const std::string& msg) : ok_(false), msg_(msg) {}
  static Status OK() { return Status(); }
  bool ok() const { return ok_; }
  const std::string& error_message() const { return msg_; }

 private:
  bool ok_;
  std::string msg_;
};

namespace errors {
inline Status InvalidArgument(const std::string& m) { return Status(m); }
}  // namespace errors

// ---------------------------------------------------------------- tensor views
// Row-major N-d view with Eigen::TensorMap-style call-operator indexing.
template <typename T, int N>
class RowMajorView {
 public:
  RowMajorView() : data_(nullptr) { for (int i = 0; i < N; ++i) dims_[i] = 0; }
  RowMajorView(T* data, const int64* dims) : data_(data) {
    for (int i = 0; i < N; ++i) dims_[i] = dims[i];
  }
  int64 dimension(int i) const { return dims_[i]; }
  T* data() const { return data_; }

  template <typename... Ix>
  T& operator()(Ix... ix) const {
    static_assert(sizeof...(Ix) == N, "rank mismatch");
    const int64 idx[N] = {static_cast<int64>(ix)...};
    int64 off = 0;
    for (int i = 0; i < N; ++i) off = off * dims_[i] + idx[i];
    return data_[off];
  }

 private:
  T* data_;
  int64 dims_[N];
};

template <typename T, int N>
struct TTypes {
  typedef RowMajorView<T, N> Tensor;
  typedef RowMajorView<const T, N> ConstTensor;
};

class TensorShape {
 public:
  TensorShape() {}
  TensorShape(std::initializer_list<int64> d) : dims_(d) {}
  explicit TensorShape(const std::vector<int64>& d) : dims_(d) {}
  int dims() const { return static_cast<int>(dims_.size()); }
  int64 dim_size(int i) const { return dims_[i]; }
  int64 num_elements() const {
    int64 n = 1;
    for (int64 d : dims_) n *= d;
    return n;
  }
  const std::vector<int64>& vec() const { return dims_; }
  bool operator==(const TensorShape& o) const { return dims_ == o.dims_; }

 private:
  std::vector<int64> dims_;
};

// Float-only tensor: either borrows caller memory or owns a std::vector.
class Tensor {
 public:
  Tensor() : ptr_(nullptr) {}
  Tensor(const TensorShape& s, float* borrowed) : shape_(s), ptr_(borrowed) {}
  explicit Tensor(const TensorShape& s) : shape_(s), own_(static_cast<size_t>(s.num_elements())) {
    ptr_ = own_.data();
  }
  const TensorShape& shape() const { return shape_; }
  float* raw() const { return ptr_; }

  template <typename T, int N>
  typename TTypes<T, N>::Tensor tensor() {
    check_rank(N);
    return typename TTypes<T, N>::Tensor(ptr_, shape_.vec().data());
  }
  template <typename T, int N>
  typename TTypes<T, N>::ConstTensor tensor() const {
    check_rank(N);
    return typename TTypes<T, N>::ConstTensor(ptr_, shape_.vec().data());
  }

 private:
  void check_rank(int n) const {
    if (shape_.dims() != n) {
      std::fprintf(stderr, "tf_min: tensor<%d>() on a rank-%d tensor\n", n, shape_.dims());
      std::abort();
    }
  }
  TensorShape shape_;
  float* ptr_;
  std::vector<float> own_;
};

// ---------------------------------------------------------------- kernels
class OpKernelConstruction {};

class OpKernelContext {
 public:
  std::vector<const Tensor*> inputs;
  // Outputs are bound by the harness to caller memory; allocate_output checks the shape
  // the reference asked for against the bound buffer.
  std::vector<Tensor> outputs;
  std::vector<TensorShape> requested;
  Status status;
  Eigen::ThreadPoolDevice cpu;
  Eigen::GpuDevice gpu;

  const Tensor& input(int i) const { return *inputs[static_cast<size_t>(i)]; }

  Status allocate_output(int i, const TensorShape& shape, Tensor** out) {
    if (static_cast<size_t>(i) >= outputs.size()) return Status("allocate_output: index out of range");
    if (requested.size() < outputs.size()) requested.resize(outputs.size());
    requested[static_cast<size_t>(i)] = shape;
    if (outputs[static_cast<size_t>(i)].raw() == nullptr) {
      outputs[static_cast<size_t>(i)] = Tensor(shape);
    } else if (outputs[static_cast<size_t>(i)].shape().num_elements() != shape.num_elements()) {
      return Status("allocate_output: bound buffer has the wrong size");
    } else {
      outputs[static_cast<size_t>(i)] = Tensor(shape, outputs[static_cast<size_t>(i)].raw());
    }
    *out = &outputs[static_cast<size_t>(i)];
    return Status::OK();
  }

  template <typename Device>
  const Device& eigen_device() const;

  void SetStatus(const Status& s) { status = s; }
};
template <>
inline const Eigen::ThreadPoolDevice& OpKernelContext::eigen_device<Eigen::ThreadPoolDevice>() const { return cpu; }
template <>
inline const Eigen::GpuDevice& OpKernelContext::eigen_device<Eigen::GpuDevice>() const { return gpu; }

class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction*) {}
  virtual ~OpKernel() {}
  virtual void Compute(OpKernelContext* context) = 0;
};

#define OP_REQUIRES(CTX, EXP, STATUS) \
  do {                                \
    if (!(EXP)) {                     \
      (CTX)->SetStatus((STATUS));     \
      return;                         \
    }                                 \
  } while (0)

#define OP_REQUIRES_OK(CTX, ...)                 \
  do {                                           \
    ::tensorflow::Status _s(__VA_ARGS__);        \
    if (!_s.ok()) {                              \
      (CTX)->SetStatus(_s);                      \
      return;                                    \
    }                                            \
  } while (0)

// ---------------------------------------------------------------- shape inference
namespace shape_inference {
struct DimensionHandle {
  int64 v;
  DimensionHandle() : v(-1) {}
  explicit DimensionHandle(int64 x) : v(x) {}
};
struct DimensionOrConstant {
  int64 v;
  DimensionOrConstant(DimensionHandle d) : v(d.v) {}  // NOLINT
  DimensionOrConstant(int64 x) : v(x) {}              // NOLINT
  DimensionOrConstant(int x) : v(x) {}                // NOLINT
};
struct ShapeHandle {
  std::vector<int64> dims;
};
class InferenceContext {
 public:
  std::vector<ShapeHandle> in;
  std::vector<ShapeHandle> out;
  ShapeHandle input(int i) const { return in[static_cast<size_t>(i)]; }
  DimensionHandle Dim(const ShapeHandle& s, int i) const { return DimensionHandle(s.dims[static_cast<size_t>(i)]); }
  ShapeHandle MakeShape(std::initializer_list<DimensionOrConstant> d) const {
    ShapeHandle s;
    for (const DimensionOrConstant& x : d) s.dims.push_back(x.v);
    return s;
  }
  void set_output(int i, const ShapeHandle& s) {
    if (out.size() <= static_cast<size_t>(i)) out.resize(static_cast<size_t>(i) + 1);
    out[static_cast<size_t>(i)] = s;
  }
};
}  // namespace shape_inference

// ---------------------------------------------------------------- registries
typedef std::function<Status(shape_inference::InferenceContext*)> ShapeFn;

struct OpDef {
  std::string name;
  std::vector<std::string> inputs, outputs;
  ShapeFn shape_fn;
};

struct Registry {
  std::map<std::string, OpDef> ops;
  // key: "<OpName>/<device>"
  std::map<std::string, std::function<OpKernel*(OpKernelConstruction*)>> kernels;
  static Registry& get() {
    static Registry r;
    return r;
  }
};

class OpDefBuilder {
 public:
  explicit OpDefBuilder(const char* name) { def_.name = name; }
  OpDefBuilder& Input(const char* s) { def_.inputs.push_back(s); return *this; }
  OpDefBuilder& Output(const char* s) { def_.outputs.push_back(s); return *this; }
  OpDefBuilder& Attr(const char*) { return *this; }
  OpDefBuilder& SetShapeFn(ShapeFn f) { def_.shape_fn = f; return *this; }
  const OpDef& def() const { return def_; }

 private:
  OpDef def_;
};
struct OpRegistrar {
  OpRegistrar(const OpDefBuilder& b) { Registry::get().ops[b.def().name] = b.def(); }  // NOLINT
};

static const char* const DEVICE_CPU = "CPU";
static const char* const DEVICE_GPU = "GPU";

class Name {
 public:
  explicit Name(const char* op) : op_(op) {}
  Name& Device(const char* d) { dev_ = d; return *this; }
  std::string key() const { return op_ + "/" + dev_; }

 private:
  std::string op_, dev_;
};
struct KernelRegistrar {
  KernelRegistrar(const Name& n, std::function<OpKernel*(OpKernelConstruction*)> f) {
    Registry::get().kernels[n.key()] = f;
  }
};

#define TFMIN_CAT_(a, b) a##b
#define TFMIN_CAT(a, b) TFMIN_CAT_(a, b)

#define REGISTER_OP(NAME) \
  static ::tensorflow::OpRegistrar TFMIN_CAT(tfmin_op_reg_, __COUNTER__) = ::tensorflow::OpDefBuilder(NAME)

#define REGISTER_KERNEL_BUILDER(NAME_EXPR, ...)                                         \
  static ::tensorflow::KernelRegistrar TFMIN_CAT(tfmin_kernel_reg_, __COUNTER__)(       \
      ::tensorflow::NAME_EXPR,                                                          \
      [](::tensorflow::OpKernelConstruction* c) -> ::tensorflow::OpKernel* { return new __VA_ARGS__(c); })

}  // namespace tensorflow

#endif  // FR_ORACLE_TF_MIN_H_
