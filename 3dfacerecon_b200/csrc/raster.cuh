// Z-buffer rasterizer kernels (forward visibility + resolve, backward scatter) for sm_100a.
//
// Replaces the reference's CUDA op (render_depth_op.cu.cc:35-381, four kernels + 13 doubles of scratch
// per triangle per face) and reproduces its CPU op (render_depth_op.cc:132-368) bit for bit:
//   1. raster_keys_kernel    one thread per (triangle, face group): bbox + cull + FP64 inside test,
//                            visibility resolved with a packed (depth, ~index) u64 atomicMax -- order
//                            independent, so no race (the reference's kernel 3 has one, .cu.cc:217-231).
//   2. raster_resolve_kernel one thread per pixel: decode the winning triangle and recompute its depth,
//                            normal and mean texture from the vertices (no per-triangle scratch).
//   3. render_backward_kernel one thread per pixel: (g*1.0f)/3.0f to the z of the triangle's 3 vertices,
//                            warp-aggregated when lanes share a triangle.
#ifndef FR_RASTER_CUH_
#define FR_RASTER_CUH_

#include "fr_common.cuh"
#include "raster_core.h"

namespace fr {

constexpr int kRasterThreads = 256;

// float triangle index -> int the way the reference does ((int)tri(k,i), render_depth_op.cc:204-206),
// rejecting anything that would index outside [0, nver).
__device__ __forceinline__ bool tri_vertex_index(float f, int nver, int* out) {
  if (!(f > -1.0f && f < (float)nver)) return false;
  *out = (int)f;
  return true;
}

// Phase A: one thread per (triangle, FPT faces): gather xy, exact bounding-box cull in float arithmetic
// (fr_tri_bbox_fast).  ~70 % of the sub-pixel BFM triangles contain no pixel centre and stop here; the survivors are
// compacted into a shared-memory queue so that
// Phase B runs the FP64 edge setup + inside tests + atomicMax with every lane busy.
template <int FPT>
__global__ void __launch_bounds__(kRasterThreads)
raster_keys_kernel(const float* __restrict__ vertex, const float* __restrict__ tri, unsigned long long* __restrict__ keys,
                   int batch, int nver, int ntri, int height, int width) {
  __shared__ float4 q_a[kRasterThreads * FPT];      // x1 y1 x2 y2
  __shared__ float2 q_b[kRasterThreads * FPT];      // x3 y3
  __shared__ unsigned short q_id[kRasterThreads * FPT];  // (local triangle << 3) | face slot
  __shared__ int s_idx[3][kRasterThreads];
  __shared__ int q_count;
  static_assert(FPT <= 8, "face slot is packed into 3 bits");

  const int tid = threadIdx.x;
  const unsigned lane = tid & 31u;
  if (tid == 0) q_count = 0;
  __syncthreads();

  const int t = blockIdx.x * kRasterThreads + tid;
  const int b0 = blockIdx.y * FPT;
  int p1 = 0, p2 = 0, p3 = 0;
  bool valid = t < ntri;
  if (valid)
    valid = tri_vertex_index(__ldg(tri + t), nver, &p1) && tri_vertex_index(__ldg(tri + ntri + t), nver, &p2) &&
            tri_vertex_index(__ldg(tri + 2 * (size_t)ntri + t), nver, &p3);
  s_idx[0][tid] = p1;
  s_idx[1][tid] = p2;
  s_idx[2][tid] = p3;

  float x1[FPT], y1[FPT], x2[FPT], y2[FPT], x3[FPT], y3[FPT];
#pragma unroll
  for (int f = 0; f < FPT; ++f) {  // all gathers in flight before the first use
    const int b = min(b0 + f, batch - 1);
    const float* vx = vertex + (size_t)b * 3 * nver;
    const float* vy = vx + nver;
    x1[f] = __ldg(vx + p1);
    x2[f] = __ldg(vx + p2);
    x3[f] = __ldg(vx + p3);
    y1[f] = __ldg(vy + p1);
    y2[f] = __ldg(vy + p2);
    y3[f] = __ldg(vy + p3);
  }
#pragma unroll
  for (int f = 0; f < FPT; ++f) {
    FrBBox bb;
    const bool keep = valid && (b0 + f < batch) &&
                      fr_tri_bbox_fast(x1[f], y1[f], x2[f], y2[f], x3[f], y3[f], width, height, &bb);
    const unsigned m = __ballot_sync(0xFFFFFFFFu, keep);
    if (m != 0u) {
      int base = 0;
      if (lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(&q_count, __popc(m));
      base = __shfl_sync(0xFFFFFFFFu, base, __ffs(m) - 1);
      if (keep) {
        const int pos = base + __popc(m & ((1u << lane) - 1u));
        q_a[pos] = make_float4(x1[f], y1[f], x2[f], y2[f]);
        q_b[pos] = make_float2(x3[f], y3[f]);
        q_id[pos] = (unsigned short)((tid << 3) | f);
      }
    }
  }
  __syncthreads();

  const int n = q_count;
  const size_t npix = (size_t)height * width;
  for (int i = tid; i < n; i += kRasterThreads) {
    const float4 a = q_a[i];
    const float2 c = q_b[i];
    const int id = q_id[i];
    const int tl = id >> 3;
    const int b = b0 + (id & 7);
    const float* vz = vertex + ((size_t)b * 3 + 2) * nver;
    const float h = fr_tri_depth(__ldg(vz + s_idx[0][tl]), __ldg(vz + s_idx[1][tl]), __ldg(vz + s_idx[2][tl]));
    if (!fr_depth_draws(h)) continue;
    FrBBox bb;
    fr_tri_bbox_fast(a.x, a.y, a.z, a.w, c.x, c.y, width, height, &bb);
    FrTriEdge e;
    fr_tri_edge_setup(a.x, a.y, a.z, a.w, c.x, c.y, &e);
    const unsigned long long key = fr_pack_key(h, blockIdx.x * kRasterThreads + tl);
    unsigned long long* kb = keys + (size_t)b * npix;
    for (int y = bb.y_min; y <= bb.y_max; ++y)
      for (int x = bb.x_min; x <= bb.x_max; ++x)
        if (fr_point_in_tri(&e, x, y)) atomicMax(kb + (size_t)y * width + x, key);
  }
}

// One thread per pixel.  Depth and triangle index are decoded straight from the key (a pure streaming pass); the
// vertex gathers only happen when normals / texture are requested or the decoded depth is a signed-zero tie.
template <bool kAttributes>
__global__ void __launch_bounds__(kRasterThreads)
raster_resolve_kernel(const unsigned long long* __restrict__ keys, const float* __restrict__ vertex,
                      const float* __restrict__ tri, const float* __restrict__ texture, long long texture_batch_stride,
                      float* __restrict__ depth, float* __restrict__ texture_image, float* __restrict__ normal,
                      float* __restrict__ tri_ind, int nver, int ntri, int npix) {
  const int p = blockIdx.x * kRasterThreads + threadIdx.x;
  if (p >= npix) return;
  const int b = blockIdx.y;
  const size_t o = (size_t)b * npix + p;
  const unsigned long long key = keys[o];
  float d = __uint_as_float(FR_BACKGROUND_DEPTH_BITS);  // render_depth_op.cc:187
  float ti = -1.0f;                                     // :192
  float n[3] = {0.0f, 0.0f, 0.0f};                      // :189-191
  float tx[3] = {0.0f, 0.0f, 0.0f};                     // :258-260
  if (key != 0ull) {
    const int t = fr_key_triangle(key);
    ti = (float)t;
    bool ambiguous;
    d = fr_key_depth(key, &ambiguous);
    if (kAttributes || ambiguous) {
      const int p1 = (int)__ldg(tri + t), p2 = (int)__ldg(tri + ntri + t), p3 = (int)__ldg(tri + 2 * (size_t)ntri + t);
      const float* vx = vertex + (size_t)b * 3 * nver;
      const float* vy = vx + nver;
      const float* vz = vy + nver;
      const float z1 = __ldg(vz + p1), z2 = __ldg(vz + p2), z3 = __ldg(vz + p3);
      d = fr_tri_depth(z1, z2, z3);  // exact bits of the winner's depth (keeps a -0.0 the key folded away)
      if (kAttributes) {
        if (normal != nullptr)
          fr_tri_normal(__ldg(vx + p1), __ldg(vy + p1), z1, __ldg(vx + p2), __ldg(vy + p2), z2, __ldg(vx + p3),
                        __ldg(vy + p3), z3, n);
        if (texture_image != nullptr) {
          const float* tex = texture + (size_t)b * texture_batch_stride;
#pragma unroll
          for (int c = 0; c < 3; ++c)
            tx[c] = fr_tri_mean(__ldg(tex + (size_t)c * nver + p1), __ldg(tex + (size_t)c * nver + p2),
                                __ldg(tex + (size_t)c * nver + p3));
        }
      }
    }
  }
  depth[o] = d;
  tri_ind[o] = ti;
  if (kAttributes) {
    if (normal != nullptr) {
      normal[3 * o + 0] = n[0];
      normal[3 * o + 1] = n[1];
      normal[3 * o + 2] = n[2];
    }
    if (texture_image != nullptr) {
      texture_image[3 * o + 0] = tx[0];
      texture_image[3 * o + 1] = tx[1];
      texture_image[3 * o + 2] = tx[2];
    }
  }
}

// Backward (render_depth_op.cc:325-368).  vertex_grad must be zero on entry (the API memsets it).
__global__ void __launch_bounds__(kRasterThreads)
render_backward_kernel(const float* __restrict__ depth_grad, const float* __restrict__ tri,
                       const float* __restrict__ tri_ind, float* __restrict__ vertex_grad, int nver, int ntri, int npix) {
  const int p = blockIdx.x * kRasterThreads + threadIdx.x;
  const int b = blockIdx.y;
  const unsigned lane = threadIdx.x & 31u;
  int t = -1;
  float share = 0.0f;
  if (p < npix) {
    const size_t o = (size_t)b * npix + p;
    const float tf = __ldg(tri_ind + o);
    if (tf >= 0.0f && tf < (float)ntri) {
      t = (int)tf;
      share = __fdiv_rn(__fmul_rn(__ldg(depth_grad + o), 1.0f), 3.0f);  // (g * 1.0f) / 3.0f, :361
    }
  }
  // warp aggregation: lanes that hit the same triangle add their shares once (lane order => deterministic
  // within the warp); skipped when every lane has its own triangle, the common case for sub-pixel meshes.
  const unsigned peers = __match_any_sync(0xFFFFFFFFu, t);
  const bool leader = (peers & ((1u << lane) - 1u)) == 0u;
  if (__any_sync(0xFFFFFFFFu, t >= 0 && peers != (1u << lane))) {
    float sum = 0.0f;
    for (int src = 0; src < 32; ++src) {
      const float v = __shfl_sync(0xFFFFFFFFu, share, src);
      if ((peers >> src) & 1u) sum += v;
    }
    share = sum;
  }
  if (t >= 0 && leader) {
    int p1, p2, p3;
    if (tri_vertex_index(__ldg(tri + t), nver, &p1) && tri_vertex_index(__ldg(tri + ntri + t), nver, &p2) &&
        tri_vertex_index(__ldg(tri + 2 * (size_t)ntri + t), nver, &p3)) {
      float* gz = vertex_grad + ((size_t)b * 3 + 2) * nver;
      atomicAdd(gz + p1, share);
      atomicAdd(gz + p2, share);
      atomicAdd(gz + p3, share);
    }
  }
}

}  // namespace fr
#endif  // FR_RASTER_CUH_
