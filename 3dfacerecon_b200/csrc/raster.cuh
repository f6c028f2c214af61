// Z-buffer rasterizer kernels (forward visibility + resolve, backward scatter) for sm_100a.
//
// Replaces the reference's CUDA op (render_depth_op.cu.cc:35-381, four kernels + 13 doubles of scratch
// per triangle per face) and reproduces its CPU op (render_depth_op.cc:132-368) bit for bit:
//   0. raster_snap_kernel    one thread per (vertex, face): biased ceil/floor pixel coordinates, 8 bytes.
//   1. raster_keys_kernel    one thread per (triangle, face group): exact integer bbox cull on the snapped
//                            vertices, survivors compacted in shared memory, then FP64 inside tests;
//                            visibility resolved with a packed (depth, ~index) u64 atomicMax -- order
//                            independent, so no race (the reference's kernel 3 has one, .cu.cc:217-231).
//   2. raster_resolve_kernel one thread per pixel: depth and index decoded from the key; normal and mean
//                            texture recomputed from the winner's vertices (no per-triangle scratch).
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

// Snap every vertex of every face to its biased ceil/floor pixel coordinates once (raster_core.h "per-vertex pixel
// snapping"): 8 bytes per vertex replace the 6 float gathers + min/max/ceil/floor per (triangle, face) of the cull.
// The same pass clears the face's visibility keys (saves a separate memset pass over the key buffer).  Each thread
// handles kSnapPerThread vertices, all loads issued before the first use: with one vertex per thread the kernel is
// bound by CTA turnover (one DRAM/L2 latency per 256 vertices), not by bandwidth.
constexpr int kSnapPerThread = 4;
__global__ void __launch_bounds__(kRasterThreads)
raster_snap_kernel(const float* __restrict__ vertex, uint2* __restrict__ snap, unsigned long long* __restrict__ keys,
                   int nver, int npix, int width, int height) {
  const int b = blockIdx.y;
  const int base = blockIdx.x * (kRasterThreads * kSnapPerThread) + threadIdx.x;
  const float* vx = vertex + (size_t)b * 3 * nver;
  float x[kSnapPerThread], y[kSnapPerThread];
#pragma unroll
  for (int j = 0; j < kSnapPerThread; ++j) {
    const int n = base + j * kRasterThreads;
    x[j] = (n < nver) ? __ldg(vx + n) : 0.0f;
    y[j] = (n < nver) ? __ldg(vx + nver + n) : 0.0f;
  }
  unsigned long long* kb = keys + (size_t)b * npix;
  for (int p = base; p < npix; p += gridDim.x * (kRasterThreads * kSnapPerThread)) {
#pragma unroll
    for (int j = 0; j < kSnapPerThread; ++j)
      if (p + j * kRasterThreads < npix) kb[p + j * kRasterThreads] = 0ull;
  }
#pragma unroll
  for (int j = 0; j < kSnapPerThread; ++j) {
    const int n = base + j * kRasterThreads;
    if (n < nver) {
      const FrSnap s = fr_snap_vertex(x[j], y[j], width, height);
      snap[(size_t)b * nver + n] = make_uint2(s.lo, s.hi);
    }
  }
}

// Phase A: one thread per (triangle, FPT faces): three 8-byte gathers of snapped vertices and a handful of packed
// integer ops decide the reference's bounding-box cull exactly.  ~70 % of the sub-pixel BFM triangles contain no
// pixel centre and stop here; the survivors are compacted into a shared-memory queue so that
// Phase B gathers the float vertices and runs the FP64 edge setup + inside tests + atomicMax with every lane busy.
template <int FPT>
__global__ void __launch_bounds__(kRasterThreads)
raster_keys_kernel(const float* __restrict__ vertex, const uint2* __restrict__ snap, const float* __restrict__ tri,
                   unsigned long long* __restrict__ keys, int batch, int nver, int ntri, int height, int width) {
  __shared__ uint2 q_box[kRasterThreads * FPT];           // biased bbox (lo_min, hi_max)
  __shared__ unsigned short q_id[kRasterThreads * FPT];   // (local triangle << 3) | face slot
  __shared__ int s_idx[3][kRasterThreads];
  __shared__ unsigned q_count;                           // single-pixel survivors | multi-pixel survivors << 16
  static_assert(FPT <= 8, "face slot is packed into 3 bits");

  const int tid = threadIdx.x;
  if (tid == 0) q_count = 0u;
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

  const uint32_t limit = (uint32_t)width | ((uint32_t)height << 16);
  uint2 sa[FPT], sb[FPT], sc[FPT];
#pragma unroll
  for (int f = 0; f < FPT; ++f) {  // all gathers in flight before the first use
    const uint2* sp = snap + (size_t)min(b0 + f, batch - 1) * nver;
    sa[f] = __ldg(sp + p1);
    sb[f] = __ldg(sp + p2);
    sc[f] = __ldg(sp + p3);
  }
  uint2 box[FPT];
  unsigned keepmask = 0u;
#pragma unroll
  for (int f = 0; f < FPT; ++f) {
    FrSnap a, b, c;
    a.lo = sa[f].x; a.hi = sa[f].y;
    b.lo = sb[f].x; b.hi = sb[f].y;
    c.lo = sc[f].x; c.hi = sc[f].y;
    const bool keep = fr_snap_keep(a, b, c, limit, &box[f].x, &box[f].y) && valid && (b0 + f < batch);
    keepmask |= (keep ? 1u : 0u) << f;
  }
  // Survivors whose bbox is a single pixel (about 60 % on the BFM mesh) are queued from the front, the others from the
  // back: phase B then runs warps of single-pixel work without a pixel loop and confines the divergent bbox loops to the
  // multi-pixel warps.  One packed shared atomic reserves both ranges.
  unsigned singlemask = 0u;
#pragma unroll
  for (int f = 0; f < FPT; ++f) singlemask |= ((box[f].x == box[f].y) ? 1u : 0u) << f;
  singlemask &= keepmask;
  const unsigned multimask = keepmask & ~singlemask;
  if (keepmask != 0u) {
    const unsigned got = atomicAdd(&q_count, (unsigned)__popc(singlemask) | ((unsigned)__popc(multimask) << 16));
    int ps = (int)(got & 0xFFFFu), pm = kRasterThreads * FPT - 1 - (int)(got >> 16);
#pragma unroll
    for (int f = 0; f < FPT; ++f) {
      if ((keepmask >> f) & 1u) {
        const int pos = ((singlemask >> f) & 1u) ? ps++ : pm--;
        q_box[pos] = box[f];
        q_id[pos] = (unsigned short)((tid << 3) | f);
      }
    }
  }
  __syncthreads();

  const unsigned counts = q_count;
  const int n_single = (int)(counts & 0xFFFFu), n_multi = (int)(counts >> 16);
  const size_t npix = (size_t)height * width;
  // ---- phase B1: one pixel per survivor, two survivors per thread and trip so that 18 gathers are in flight at once
  for (int i0 = tid; i0 < n_single; i0 += 2 * kRasterThreads) {
    const int i1e = i0 + kRasterThreads;
    const bool two = i1e < n_single;
    const int ia = i0, ib = two ? i1e : i0;
    const uint2 bxa = q_box[ia], bxb = q_box[ib];
    const int ida = q_id[ia], idb = q_id[ib];
    const int tla = ida >> 3, tlb = idb >> 3;
    const int ba = b0 + (ida & 7), bb_ = b0 + (idb & 7);
    const float* vxa = vertex + (size_t)ba * 3 * nver;
    const float* vxb = vertex + (size_t)bb_ * 3 * nver;
    const int a1 = s_idx[0][tla], a2 = s_idx[1][tla], a3 = s_idx[2][tla];
    const int c1 = s_idx[0][tlb], c2 = s_idx[1][tlb], c3 = s_idx[2][tlb];
    float xa[3], ya[3], za[3], xb[3], yb[3], zb[3];
    xa[0] = __ldg(vxa + a1); xa[1] = __ldg(vxa + a2); xa[2] = __ldg(vxa + a3);
    ya[0] = __ldg(vxa + nver + a1); ya[1] = __ldg(vxa + nver + a2); ya[2] = __ldg(vxa + nver + a3);
    za[0] = __ldg(vxa + 2 * (size_t)nver + a1); za[1] = __ldg(vxa + 2 * (size_t)nver + a2); za[2] = __ldg(vxa + 2 * (size_t)nver + a3);
    xb[0] = __ldg(vxb + c1); xb[1] = __ldg(vxb + c2); xb[2] = __ldg(vxb + c3);
    yb[0] = __ldg(vxb + nver + c1); yb[1] = __ldg(vxb + nver + c2); yb[2] = __ldg(vxb + nver + c3);
    zb[0] = __ldg(vxb + 2 * (size_t)nver + c1); zb[1] = __ldg(vxb + 2 * (size_t)nver + c2); zb[2] = __ldg(vxb + 2 * (size_t)nver + c3);
    {
      const float h = fr_tri_depth(za[0], za[1], za[2]);
      if (fr_depth_draws(h)) {
        FrTriEdge e;
        fr_tri_edge_setup(xa[0], ya[0], xa[1], ya[1], xa[2], ya[2], &e);
        const int x = (int)(bxa.x & 0xFFFFu) - 1, y = (int)(bxa.x >> 16) - 1;
        if (fr_point_in_tri(&e, x, y))
          atomicMax(keys + ((size_t)ba * npix + (size_t)y * width + x), fr_pack_key(h, blockIdx.x * kRasterThreads + tla));
      }
    }
    if (two) {
      const float h = fr_tri_depth(zb[0], zb[1], zb[2]);
      if (fr_depth_draws(h)) {
        FrTriEdge e;
        fr_tri_edge_setup(xb[0], yb[0], xb[1], yb[1], xb[2], yb[2], &e);
        const int x = (int)(bxb.x & 0xFFFFu) - 1, y = (int)(bxb.x >> 16) - 1;
        if (fr_point_in_tri(&e, x, y))
          atomicMax(keys + ((size_t)bb_ * npix + (size_t)y * width + x), fr_pack_key(h, blockIdx.x * kRasterThreads + tlb));
      }
    }
  }
  // ---- phase B2: survivors with several candidate pixels
  for (int j = tid; j < n_multi; j += kRasterThreads) {
    const int i = kRasterThreads * FPT - 1 - j;
    const uint2 bx = q_box[i];
    const int id = q_id[i];
    const int tl = id >> 3;
    const int b = b0 + (id & 7);
    const int i1 = s_idx[0][tl], i2 = s_idx[1][tl], i3 = s_idx[2][tl];
    const float* vx = vertex + (size_t)b * 3 * nver;
    const float* vy = vx + nver;
    const float* vz = vy + nver;
    const float x1 = __ldg(vx + i1), x2 = __ldg(vx + i2), x3 = __ldg(vx + i3);
    const float y1 = __ldg(vy + i1), y2 = __ldg(vy + i2), y3 = __ldg(vy + i3);
    const float h = fr_tri_depth(__ldg(vz + i1), __ldg(vz + i2), __ldg(vz + i3));
    if (!fr_depth_draws(h)) continue;
    FrBBox bb;
    fr_snap_bbox(bx.x, bx.y, &bb);
    FrTriEdge e;
    fr_tri_edge_setup(x1, y1, x2, y2, x3, y3, &e);
    const unsigned long long key = fr_pack_key(h, blockIdx.x * kRasterThreads + tl);
    unsigned long long* kb = keys + (size_t)b * npix;
    for (int y = bb.y_min; y <= bb.y_max; ++y)
      for (int x = bb.x_min; x <= bb.x_max; ++x)
        if (fr_point_in_tri(&e, x, y)) atomicMax(kb + (size_t)y * width + x, key);
  }
}

// kResolvePerThread pixels per thread (keys loaded up front, coalesced).  Depth and triangle index are decoded straight
// from the key (a pure streaming pass); the vertex gathers only happen when normals / texture are requested or the
// decoded depth is a signed-zero tie.
constexpr int kResolvePerThread = 4;
template <bool kAttributes>
__global__ void __launch_bounds__(kRasterThreads)
raster_resolve_kernel(const unsigned long long* __restrict__ keys, const float* __restrict__ vertex,
                      const float* __restrict__ tri, const float* __restrict__ texture, long long texture_batch_stride,
                      float* __restrict__ depth, float* __restrict__ texture_image, float* __restrict__ normal,
                      float* __restrict__ tri_ind, int nver, int ntri, int npix) {
  const int b = blockIdx.y;
  const int base = blockIdx.x * (kRasterThreads * kResolvePerThread) + threadIdx.x;
  const size_t fo = (size_t)b * npix;
  unsigned long long k[kResolvePerThread];
#pragma unroll
  for (int j = 0; j < kResolvePerThread; ++j) {
    const int p = base + j * kRasterThreads;
    k[j] = (p < npix) ? keys[fo + p] : 0ull;
  }
#pragma unroll
  for (int j = 0; j < kResolvePerThread; ++j) {
    const int p = base + j * kRasterThreads;
    if (p >= npix) continue;
    const size_t o = fo + p;
    const unsigned long long key = k[j];
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
