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

template <int FPT>  // faces handled by one thread (index loads and address math amortised over them)
__global__ void __launch_bounds__(kRasterThreads)
raster_keys_kernel(const float* __restrict__ vertex, const float* __restrict__ tri, unsigned long long* __restrict__ keys,
                   int batch, int nver, int ntri, int height, int width) {
  const int t = blockIdx.x * kRasterThreads + threadIdx.x;
  if (t >= ntri) return;
  int p1, p2, p3;
  if (!tri_vertex_index(__ldg(tri + t), nver, &p1) || !tri_vertex_index(__ldg(tri + ntri + t), nver, &p2) ||
      !tri_vertex_index(__ldg(tri + 2 * (size_t)ntri + t), nver, &p3))
    return;
  const int b0 = blockIdx.y * FPT;
  const size_t npix = (size_t)height * width;

  // issue every xy gather of this thread's faces before using any of them
  float x1[FPT], y1[FPT], x2[FPT], y2[FPT], x3[FPT], y3[FPT];
#pragma unroll
  for (int f = 0; f < FPT; ++f) {
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
    const int b = b0 + f;
    if (b >= batch) break;
    FrBBox bb;
    if (!fr_tri_bbox(x1[f], y1[f], x2[f], y2[f], x3[f], y3[f], width, height, &bb)) continue;
    const float* vz = vertex + ((size_t)b * 3 + 2) * nver;
    const float h = fr_tri_depth(__ldg(vz + p1), __ldg(vz + p2), __ldg(vz + p3));
    if (!fr_depth_draws(h)) continue;
    FrTriEdge e;
    fr_tri_edge_setup(x1[f], y1[f], x2[f], y2[f], x3[f], y3[f], &e);
    const unsigned long long key = fr_pack_key(h, t);
    unsigned long long* kb = keys + (size_t)b * npix;
    for (int y = bb.y_min; y <= bb.y_max; ++y)
      for (int x = bb.x_min; x <= bb.x_max; ++x)
        if (fr_point_in_tri(&e, x, y)) atomicMax(kb + (size_t)y * width + x, key);
  }
}

// One thread per pixel.  texture_image / normal may be null (skipped).
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
    const int p1 = (int)__ldg(tri + t), p2 = (int)__ldg(tri + ntri + t), p3 = (int)__ldg(tri + 2 * (size_t)ntri + t);
    const float* vx = vertex + (size_t)b * 3 * nver;
    const float* vy = vx + nver;
    const float* vz = vy + nver;
    const float z1 = __ldg(vz + p1), z2 = __ldg(vz + p2), z3 = __ldg(vz + p3);
    d = fr_tri_depth(z1, z2, z3);  // exact bits of the winner's depth (keeps a -0.0 the key folded away)
    ti = (float)t;
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
  depth[o] = d;
  tri_ind[o] = ti;
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
