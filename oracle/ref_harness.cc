// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// C-ABI driver for the UNMODIFIED reference CPU op.  This TU is linked with
// /root/reference/rendering_layer/ops_src/render_depth_op.cc (compiled where it lies,
// against oracle/tf_shim) into oracle/_ref/libref_render_depth.so.  It looks the
// reference's kernels up in the registry they registered themselves in
// (render_depth_op.cc:592-593) and calls their Compute() methods
// (render_depth_op.cc:378-458 forward, :470-528 backward), so shape validation
// (:405-418, :497-502), output shapes and the CPU functors (:132-322, :325-368) all run
// exactly as written upstream.
#include "tf_shim/tf_min.h"

#include <algorithm>
#include <string>

using tensorflow::int64;
using tensorflow::OpKernel;
using tensorflow::OpKernelConstruction;
using tensorflow::OpKernelContext;
using tensorflow::Registry;
using tensorflow::Tensor;
using tensorflow::TensorShape;

namespace {

void put_err(char* err, int errlen, const std::string& m) {
  if (err == nullptr || errlen <= 0) return;
  std::snprintf(err, static_cast<size_t>(errlen), "%s", m.c_str());
}

TensorShape shape_of(const int64* d, int rank) { return TensorShape(std::vector<int64>(d, d + rank)); }

int run_kernel(const char* key, OpKernelContext* ctx, char* err, int errlen) {
  auto it = Registry::get().kernels.find(key);
  if (it == Registry::get().kernels.end()) {
    put_err(err, errlen, std::string("kernel not registered: ") + key);
    return 2;
  }
  OpKernelConstruction cons;
  std::unique_ptr<OpKernel> k(it->second(&cons));
  k->Compute(ctx);
  if (!ctx->status.ok()) {
    put_err(err, errlen, ctx->status.error_message());
    return 1;
  }
  return 0;
}

}  // namespace

extern "C" {

// Forward.  *_dims are row-major shapes: vertex[3], tri[2], texture[3], image[4].
// Outputs are caller buffers sized from image dims: depth[B,H,W,1], texture_image[B,H,W,ch],
// normal[B,H,W,3], tri_ind[B,H,W,1].  `image` values are never read by the reference
// (render_depth_op.cc:397-403 only takes its dims), so no image data pointer is taken.
// Returns 0 ok, 1 = the reference raised InvalidArgument (message in err), 2 = harness error.
int ref_render_depth_op(const float* vertex, const int64* vertex_dims, const float* tri, const int64* tri_dims,
                        const float* texture, const int64* texture_dims, const int64* image_dims, float* depth,
                        float* texture_image, float* normal, float* tri_ind, char* err, int errlen) {
  Tensor t_vertex(shape_of(vertex_dims, 3), const_cast<float*>(vertex));
  Tensor t_tri(shape_of(tri_dims, 2), const_cast<float*>(tri));
  Tensor t_texture(shape_of(texture_dims, 3), const_cast<float*>(texture));
  Tensor t_image(shape_of(image_dims, 4), nullptr);
  const int64 B = image_dims[0], H = image_dims[1], W = image_dims[2], ch = texture_dims[1];
  OpKernelContext ctx;
  ctx.inputs = {&t_vertex, &t_tri, &t_texture, &t_image};
  ctx.outputs.resize(4);
  ctx.outputs[0] = Tensor(TensorShape({B, H, W, 1}), depth);
  ctx.outputs[1] = Tensor(TensorShape({B, H, W, ch}), texture_image);
  ctx.outputs[2] = Tensor(TensorShape({B, H, W, 3}), normal);
  ctx.outputs[3] = Tensor(TensorShape({B, H, W, 1}), tri_ind);
  return run_kernel("RenderDepth/CPU", &ctx, err, errlen);
}

// Backward.  vertex_grad is a caller buffer of vertex's shape; the reference does NOT
// zero it (render_depth_op.cc:514-516) and dereferences tri(k, tri_ind) for every pixel
// including tri_ind == -1 (:349-353), so callers zero vertex_grad and sanitise background
// pixels first (see oracle/__init__.py).  `depth` is passed through but never read (:329).
int ref_render_depth_grad_op(const float* depth_grad, const int64* depth_grad_dims, const float* vertex,
                             const int64* vertex_dims, const float* tri, const int64* tri_dims,
                             const float* depth, const float* tri_ind, const int64* image_dims,
                             float* vertex_grad, char* err, int errlen) {
  Tensor t_dg(shape_of(depth_grad_dims, 4), const_cast<float*>(depth_grad));
  Tensor t_vertex(shape_of(vertex_dims, 3), const_cast<float*>(vertex));
  Tensor t_tri(shape_of(tri_dims, 2), const_cast<float*>(tri));
  Tensor t_depth(shape_of(depth_grad_dims, 4), const_cast<float*>(depth));
  Tensor t_tri_ind(shape_of(depth_grad_dims, 4), const_cast<float*>(tri_ind));
  Tensor t_image(shape_of(image_dims, 4), nullptr);
  OpKernelContext ctx;
  ctx.inputs = {&t_dg, &t_vertex, &t_tri, &t_depth, &t_tri_ind, &t_image};
  ctx.outputs.resize(1);
  ctx.outputs[0] = Tensor(shape_of(vertex_dims, 3), vertex_grad);
  return run_kernel("RenderDepthGrad/CPU", &ctx, err, errlen);
}

// Runs the reference's registered shape function (render_depth_op.cc:544-569, :586-589).
// in_dims is the concatenation of all input shapes; out_* receive up to 4 outputs x 4 dims.
// Returns the number of outputs, or -1 when the op is unknown.
int ref_infer_shapes(const char* op, int n_in, const int* in_ranks, const int64* in_dims, int* out_ranks,
                     int64* out_dims) {
  auto it = Registry::get().ops.find(op);
  if (it == Registry::get().ops.end()) return -1;
  tensorflow::shape_inference::InferenceContext c;
  const int64* p = in_dims;
  for (int i = 0; i < n_in; ++i) {
    tensorflow::shape_inference::ShapeHandle s;
    s.dims.assign(p, p + in_ranks[i]);
    p += in_ranks[i];
    c.in.push_back(s);
  }
  tensorflow::Status st = it->second.shape_fn(&c);
  if (!st.ok()) return -1;
  const int n_out = static_cast<int>(std::min<size_t>(c.out.size(), 4));
  for (int i = 0; i < n_out; ++i) {
    out_ranks[i] = static_cast<int>(c.out[static_cast<size_t>(i)].dims.size());
    for (int j = 0; j < out_ranks[i] && j < 4; ++j) out_dims[4 * i + j] = c.out[static_cast<size_t>(i)].dims[static_cast<size_t>(j)];
  }
  return n_out;
}

// Registered op signature, e.g. "vertex: float|tri: float|...->depth: float|...".
int ref_op_signature(const char* op, char* buf, int buflen) {
  auto it = Registry::get().ops.find(op);
  if (it == Registry::get().ops.end()) return -1;
  std::string s;
  for (size_t i = 0; i < it->second.inputs.size(); ++i) s += (i ? "|" : "") + it->second.inputs[i];
  s += "->";
  for (size_t i = 0; i < it->second.outputs.size(); ++i) s += (i ? "|" : "") + it->second.outputs[i];
  std::snprintf(buf, static_cast<size_t>(buflen), "%s", s.c_str());
  return static_cast<int>(s.size());
}

}  // extern "C"
