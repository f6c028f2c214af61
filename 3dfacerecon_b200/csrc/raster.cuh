// Z-buffer rasterizer kernels (forward visibility + resolve, backward scatter) for sm_100a.
//
// Replaces the reference's CUDA op (render_depth_op.cu.cc:35-381, four kernels + 13 doubles of scratch
// per triangle per face) and reproduces its CPU op (render_depth_op.cc:132-368) bit for bit:
//   0. raster_pack_kernel    (vertex, face) -> 16-byte record {x, y, z, snap code}; clears the visibility keys.
//   1. raster_keys_kernel    one thread per (triangle, face group): exact integer bbox cull on the snap
//                            codes, survivors compacted in shared memory, then FP64 inside tests;
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
#ifndef FR_KEYS_THREADS
#define FR_KEYS_THREADS 128   // A/B on B200 at B = 64: 64 -> 55.5 us, 128 -> 54.5 us, 256 -> 59 us, 512 -> 66 us (smaller blocks: less time lost at the phase barrier)
#endif
constexpr int kKeysThreads = FR_KEYS_THREADS;   // threads (= triangles) per block of raster_keys_kernel

// float triangle index -> int the way the reference does ((int)tri(k,i), render_depth_op.cc:204-206),
// rejecting anything that would index outside [0, nver).
__device__ __forceinline__ bool tri_vertex_index(float f, int nver, int* out) {
  if (!(f > -1.0f && f < (float)nver)) return false;
  *out = (int)f;
  return true;
}

// Vertex records: every (vertex, face) is repacked ONCE into 16 bytes  { x, y, z, snap code }  (raster_core.h "one-word
// snap code").  The cull of a (triangle, face) pair then needs three 4-byte gathers of the code words, and a surviving
// pair gets everything else with three 16-byte gathers of the very sectors the cull just pulled into L1 -- instead of
// nine 4-byte gathers from the three coordinate planes.  The same pass clears the face's visibility keys.  Each thread
// handles kSnapPerThread vertices with all loads issued before the first use (one vertex per thread is bound by CTA
// turnover, not by bandwidth).
constexpr int kSnapPerThread = 4;
__global__ void __launch_bounds__(kRasterThreads)
raster_pack_kernel(const float* __restrict__ vertex, float4* __restrict__ rec, unsigned long long* __restrict__ keys,
                   const int32_t* __restrict__ vert_rank, int nver, int npix, int width, int height) {
  pdl_trigger();
  const int b = blockIdx.y;
  const int base = blockIdx.x * (kRasterThreads * kSnapPerThread) + threadIdx.x;
  const float* vx = vertex + (size_t)b * 3 * nver;
  float x[kSnapPerThread], y[kSnapPerThread], z[kSnapPerThread];
#pragma unroll
  for (int j = 0; j < kSnapPerThread; ++j) {
    const int n = base + j * kRasterThreads;
    const bool ok = n < nver;
    x[j] = ok ? __ldg(vx + n) : 0.0f;
    y[j] = ok ? __ldg(vx + nver + n) : 0.0f;
    z[j] = ok ? __ldg(vx + 2 * (size_t)nver + n) : 0.0f;
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
    if (n < nver)   // records are stored by vertex rank (mesh_table.h) when a mesh table is in play
      rec[(size_t)b * nver + (vert_rank != nullptr ? __ldg(vert_rank + n) : n)] =
          make_float4(x[j], y[j], z[j], __uint_as_float(fr_snap_code(x[j], y[j], width, height)));
  }
}

// Phase B body: depth, FP64 edge setup, inside tests over the bbox (one flat loop: a warp runs as many trips as its largest
// box has pixels, not rows x columns of the lane-wise maxima), packed (depth, ~index) atomicMax.
__device__ __forceinline__ void raster_draw(const float4& r1, const float4& r2, const float4& r3, uint2 box, int tri_index,
                                            unsigned long long* __restrict__ kb, int width) {
  const float h = fr_tri_depth(r1.z, r2.z, r3.z);
  if (!fr_depth_draws(h)) return;
  FrTriEdge e;
  fr_tri_edge_setup(r1.x, r1.y, r2.x, r2.y, r3.x, r3.y, &e);
  const unsigned long long key = fr_pack_key(h, tri_index);
  const int x0 = (int)(box.x & 0xFFFFu) - 1, y0 = (int)(box.x >> 16) - 1;
  const int x1 = (int)(box.y & 0xFFFFu) - 1, y1 = (int)(box.y >> 16) - 1;
  int x = x0, y = y0;    // (nested row / column loops measured 53.4 us against 50.6 us for this flat walk)
  while (y <= y1) {
    if (fr_point_in_tri(&e, x, y)) atomicMax(kb + (y * width + x), key);
    if (++x > x1) {
      x = x0;
      ++y;
    }
  }
}

// Phase A: one thread per (triangle, FPT faces): three 4-byte gathers of snap codes and a handful of packed integer ops
// decide the reference's bounding-box cull exactly.  ~55 % of the sub-pixel BFM triangles contain no pixel centre and
// stop here; the survivors are compacted into a shared-memory queue (single-pixel boxes from the front, the others from
// the back) so that Phase B gathers the records and runs the FP64 edge setup + inside tests + atomicMax with every lane
// busy and without a divergent pixel loop in the single-pixel warps.
// All record indices are 32-bit: the API guarantees batch * 3 * nver < 2^31.
#ifndef FR_KEYS_MINB
#define FR_KEYS_MINB 1
#endif

// kTable: the triangles come from a mesh table (mesh_table.h) instead of the reference's float index tensor: one 16-byte
// load of pre-validated integer vertex ids per triangle, in CLUSTER order -- the triangles of a block touch one compact
// patch of vertices whatever the numbering of the mesh (the three float loads + conversions + range checks of the generic
// flavour, and its dependence on the generator's triangle order, go away).  tri_vid[s] = {p1, p2, p3, original index}.
template <int FPT, bool kTable>
__global__ void __launch_bounds__(kKeysThreads, FR_KEYS_MINB)
raster_keys_kernel(const float4* __restrict__ rec, const float* __restrict__ tri, const uint4* __restrict__ tri_vid,
                   unsigned long long* __restrict__ keys, int batch, int nver, int ntri, int height, int width) {
  __shared__ uint2 q_box[kKeysThreads * FPT];           // biased bbox (lo_min, hi_max)
  __shared__ unsigned short q_id[kKeysThreads * FPT];   // (local triangle << 3) | face slot
  __shared__ int s_idx[4][kKeysThreads];                // three vertex ids + the triangle's index in the reference's order
  __shared__ unsigned q_count;                            // single-pixel survivors | multi-pixel survivors << 16
  static_assert(FPT <= 8, "face slot is packed into 3 bits");

  const int tid = threadIdx.x;
  if (tid == 0) q_count = 0u;
  pdl_trigger();   // the resolve pass may become resident once every block of this grid has started
  pdl_wait();      // records and cleared keys of the producing kernel (pack pass or reconstruction epilogue) are complete
  __syncthreads();

  const int t = blockIdx.x * kKeysThreads + tid;          // ntri = triangles (generic) / table entries (kTable)
  const int b0 = blockIdx.y * FPT;
  int p1 = 0, p2 = 0, p3 = 0, torig = t;
  bool valid = t < ntri;
  if (valid) {
    if (kTable) {
      const uint4 e = __ldg(tri_vid + t);
      p1 = (int)e.x;
      p2 = (int)e.y;
      p3 = (int)e.z;
      torig = (int)e.w;
    } else {
      valid = tri_vertex_index(__ldg(tri + t), nver, &p1) && tri_vertex_index(__ldg(tri + ntri + t), nver, &p2) &&
              tri_vertex_index(__ldg(tri + 2 * (size_t)ntri + t), nver, &p3);
    }
  }
  s_idx[0][tid] = p1;
  s_idx[1][tid] = p2;
  s_idx[2][tid] = p3;
  s_idx[3][tid] = torig;

  const uint32_t limit = (uint32_t)width | ((uint32_t)height << 16);
  uint32_t e1[FPT], e2[FPT], e3[FPT];
#pragma unroll
  for (int f = 0; f < FPT; ++f) {  // all gathers in flight before the first use
    const unsigned fb = (unsigned)min(b0 + f, batch - 1) * (unsigned)nver;
    e1[f] = __float_as_uint(__ldg(&rec[fb + (unsigned)p1].w));
    e2[f] = __float_as_uint(__ldg(&rec[fb + (unsigned)p2].w));
    e3[f] = __float_as_uint(__ldg(&rec[fb + (unsigned)p3].w));
  }
  uint2 box[FPT];
  unsigned keepmask = 0u, singlemask = 0u;
#pragma unroll
  for (int f = 0; f < FPT; ++f) {
    const bool keep = fr_code_keep(e1[f], e2[f], e3[f], limit, &box[f].x, &box[f].y) && valid && (b0 + f < batch);
    keepmask |= (keep ? 1u : 0u) << f;
    singlemask |= ((box[f].x == box[f].y) ? 1u : 0u) << f;
  }
  singlemask &= keepmask;
  const unsigned multimask = keepmask & ~singlemask;
  if (keepmask != 0u) {     // one packed shared atomic reserves both queue ranges
    const unsigned got = atomicAdd(&q_count, (unsigned)__popc(singlemask) | ((unsigned)__popc(multimask) << 16));
    int ps = (int)(got & 0xFFFFu), pm = kKeysThreads * FPT - 1 - (int)(got >> 16);
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
  const int npix = height * width;
  // ---- phase B: one work list (one-pixel survivors first, then the others), one survivor per thread and trip: the lanes left
  // over when the one-pixel class runs out start on the multi-pixel class instead of idling through a partial trip
  // (measured against two separate loops with two one-pixel survivors per thread: 55.1 -> 53.9 us)
  for (int i = tid; i < n_single + n_multi; i += kKeysThreads) {
    const bool single = i < n_single;
    const int q = single ? i : kKeysThreads * FPT - 1 - (i - n_single);
    const uint2 bx = q_box[q];
    const int id = q_id[q];
    const int tl = id >> 3;
    const int b = b0 + (id & 7);
    const unsigned fb = (unsigned)b * (unsigned)nver;
    const float4 r1 = __ldg(rec + (fb + (unsigned)s_idx[0][tl])), r2 = __ldg(rec + (fb + (unsigned)s_idx[1][tl])),
                 r3 = __ldg(rec + (fb + (unsigned)s_idx[2][tl]));
    raster_draw(r1, r2, r3, bx, s_idx[3][tl], keys + (size_t)b * npix, width);
  }
}

// A texture shared by all faces (texture_batch_stride == 0: [3, nver] planar, the reference's tiled vertex colours) repacked once
// per call as one float4 per vertex RANK, so that the resolve pass fetches a winner's three texture vectors with three 16-byte
// gathers instead of nine 4-byte ones.
__global__ void __launch_bounds__(kRasterThreads)
raster_pack_texture_kernel(const float* __restrict__ texture, const int32_t* __restrict__ vert_rank, float4* __restrict__ tex4, int nver) {
  const int n = blockIdx.x * kRasterThreads + threadIdx.x;
  if (n >= nver) return;
  tex4[vert_rank != nullptr ? __ldg(vert_rank + n) : n] =
      make_float4(__ldg(texture + n), __ldg(texture + nver + n), __ldg(texture + 2 * (size_t)nver + n), 0.0f);
}

// kResolvePerThread pixels per thread (keys loaded up front, coalesced).  Depth and triangle index are decoded straight
// from the key (a pure streaming pass that never touches the vertices); the vertex gathers only happen when normals /
// texture are requested.
constexpr int kResolvePerThread = 4;
// Extra outputs of FaceRecNet.rendering_layer (nets/network.py:184-199) when the post-processing is fused into the resolve
// pass (fr_rendering_layer_forward): `texture_image` then receives pncc = clip(tex, 1e-6, 1), `normal` the normals flipped to
// +z and normalised, `depth` max(depth, 1e-6); maskimg = clip(depth, 1e-6, 1) * im_gray; raw_depth keeps the op's depth for
// the gradient gates.  All pointers null = the plain op.
struct LayerOut {
  float* maskimg;
  const float* im_gray;   // [B,H,W,1] or null (mask only)
  float* raw_depth;
  bool enabled;
};

template <bool kAttributes>
__global__ void __launch_bounds__(kRasterThreads)
raster_resolve_kernel(const unsigned long long* __restrict__ keys, const float* __restrict__ vertex, const float4* __restrict__ rec,
                      const int32_t* __restrict__ vert_rank, const uint4* __restrict__ tri_rank4, const float4* __restrict__ tex4,
                      const float* __restrict__ tri, const float* __restrict__ texture, long long texture_batch_stride,
                      float* __restrict__ depth, float* __restrict__ texture_image, float* __restrict__ normal,
                      float* __restrict__ tri_ind, int nver, int ntri, int npix, LayerOut layer) {
  __shared__ __align__(16) float s_attr[kAttributes ? 2 : 1][kAttributes ? 3 * kRasterThreads * kResolvePerThread : 4];
  pdl_wait();      // every atomicMax of the keys kernel has landed
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
      d = fr_key_depth(key);           // exact bits of the winner's depth, including the sign of a zero (raster_core.h)
      if (kAttributes) {               // the winner's vertices are only needed for normals / texture
        if (tri_rank4 != nullptr && rec != nullptr) {
          // mesh table: ONE 16-byte gather gives the three vertex ranks; records and the packed texture are indexed by rank
          const uint4 tr = __ldg(tri_rank4 + t);
          if (normal != nullptr) {
            const float4* rb = rec + (size_t)b * nver;
            const float4 r1 = __ldg(rb + tr.x), r2 = __ldg(rb + tr.y), r3 = __ldg(rb + tr.z);
            fr_tri_normal(r1.x, r1.y, r1.z, r2.x, r2.y, r2.z, r3.x, r3.y, r3.z, n);
          }
          if (texture_image != nullptr) {
            if (tex4 != nullptr) {
              const float4 a = __ldg(tex4 + tr.x), bb = __ldg(tex4 + tr.y), c = __ldg(tex4 + tr.z);
              tx[0] = fr_tri_mean(a.x, bb.x, c.x);
              tx[1] = fr_tri_mean(a.y, bb.y, c.y);
              tx[2] = fr_tri_mean(a.z, bb.z, c.z);
            } else {                   // per-face textures: planar gathers by vertex id
              const int p1 = (int)__ldg(tri + t), p2 = (int)__ldg(tri + ntri + t), p3 = (int)__ldg(tri + 2 * (size_t)ntri + t);
              const float* tex = texture + (size_t)b * texture_batch_stride;
#pragma unroll
              for (int c = 0; c < 3; ++c)
                tx[c] = fr_tri_mean(__ldg(tex + (size_t)c * nver + p1), __ldg(tex + (size_t)c * nver + p2),
                                    __ldg(tex + (size_t)c * nver + p3));
            }
          }
        } else {
        const int p1 = (int)__ldg(tri + t), p2 = (int)__ldg(tri + ntri + t), p3 = (int)__ldg(tri + 2 * (size_t)ntri + t);
        if (normal != nullptr) {
          if (rec != nullptr) {        // the rasterizer's 16-byte records (by rank with a mesh table): three gathers, not nine
            const float4* rb = rec + (size_t)b * nver;
            const int q1 = vert_rank ? __ldg(vert_rank + p1) : p1, q2 = vert_rank ? __ldg(vert_rank + p2) : p2,
                      q3 = vert_rank ? __ldg(vert_rank + p3) : p3;
            const float4 r1 = __ldg(rb + q1), r2 = __ldg(rb + q2), r3 = __ldg(rb + q3);
            fr_tri_normal(r1.x, r1.y, r1.z, r2.x, r2.y, r2.z, r3.x, r3.y, r3.z, n);
          } else {
            const float* vx = vertex + (size_t)b * 3 * nver;
            const float* vy = vx + nver;
            const float* vz = vy + nver;
            fr_tri_normal(__ldg(vx + p1), __ldg(vy + p1), __ldg(vz + p1), __ldg(vx + p2), __ldg(vy + p2), __ldg(vz + p2),
                          __ldg(vx + p3), __ldg(vy + p3), __ldg(vz + p3), n);
          }
        }
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
    if (kAttributes && layer.enabled) {
      // nets/network.py:185-199, same float operations in the same order as the torch / TF elementwise passes
      if (layer.raw_depth != nullptr) layer.raw_depth[o] = d;
#pragma unroll
      for (int c = 0; c < 3; ++c) tx[c] = fminf(fmaxf(tx[c], 1e-6f), 1.0f);                     // :185 pncc
      if (n[2] < 0.0f) {                                                                          // :188-189
        n[0] = __fmul_rn(-1.0f, n[0]);
        n[1] = __fmul_rn(-1.0f, n[1]);
        n[2] = __fmul_rn(-1.0f, n[2]);
      }
      float mag = __fadd_rn(__fadd_rn(__fmul_rn(n[0], n[0]), __fmul_rn(n[1], n[1])), __fmul_rn(n[2], n[2]));   // :190
      mag = (mag > 1e-6f) ? mag : 1.0f;                                                          // :191
      const float den = __fadd_rn(__fsqrt_rn(mag), 1e-6f);                                        // :192
      n[0] = __fdiv_rn(n[0], den);
      n[1] = __fdiv_rn(n[1], den);
      n[2] = __fdiv_rn(n[2], den);
      const float m = fminf(fmaxf(d, 1e-6f), 1.0f);                                               // :195
      layer.maskimg[o] = (layer.im_gray != nullptr) ? __fmul_rn(m, __ldg(layer.im_gray + o)) : m; // :196
      d = fmaxf(d, 1e-6f);                                                                        // :199
    }
    depth[o] = d;
    tri_ind[o] = ti;
    if (kAttributes) {   // 3-channel outputs go through shared memory: a 12-byte-stride store per channel triples the L2 write traffic
      const int pl = p - blockIdx.x * (kRasterThreads * kResolvePerThread);
      s_attr[0][3 * pl + 0] = n[0];
      s_attr[0][3 * pl + 1] = n[1];
      s_attr[0][3 * pl + 2] = n[2];
      s_attr[1][3 * pl + 0] = tx[0];
      s_attr[1][3 * pl + 1] = tx[1];
      s_attr[1][3 * pl + 2] = tx[2];
    }
  }
  if (kAttributes) {
    __syncthreads();
    constexpr int kBlockPix = kRasterThreads * kResolvePerThread;
    const int p0 = blockIdx.x * kBlockPix;
    const int cnt = min(kBlockPix, npix - p0) * 3;                         // floats of this block per attribute
    float* outs[2] = {normal, texture_image};
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      if (outs[a] == nullptr) continue;
      float* dst = outs[a] + 3 * (fo + p0);
      if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        for (int i = threadIdx.x * 4; i < cnt; i += kRasterThreads * 4) {
          if (i + 4 <= cnt) *reinterpret_cast<float4*>(dst + i) = *reinterpret_cast<const float4*>(&s_attr[a][i]);
          else for (int e = i; e < cnt; ++e) dst[e] = s_attr[a][e];
        }
      } else {
        for (int i = threadIdx.x; i < cnt; i += kRasterThreads) dst[i] = s_attr[a][i];
      }
    }
  }
}

// Depth + index only (the fused params -> depth-map call): a pure streaming pass over the keys of ALL faces as one flat array,
// four consecutive pixels per thread -- two 16-byte key loads, one 16-byte store per output (render_depth_op.cc:187,192 for
// the background values).  The caller guarantees 16-byte aligned pointers and n4 = n / 4 groups; the last (n % 4) pixels are
// done by the first threads of block 0.
__global__ void __launch_bounds__(kRasterThreads)
raster_resolve_depth_kernel(const unsigned long long* __restrict__ keys, float* __restrict__ depth, float* __restrict__ tri_ind,
                            unsigned long long n) {
  FR_MARK_MIN(6);
  pdl_trigger();   // the next call's prep kernel (a dependent launch that waits for this grid before it touches anything) may become resident
  pdl_wait();      // every atomicMax of the visibility pass has landed
  FR_MARK_MIN(4);
  const unsigned long long i = (unsigned long long)blockIdx.x * kRasterThreads + threadIdx.x;
  const unsigned long long n4 = n >> 2;
  auto decode = [](unsigned long long key, float* d, float* t) {
    *d = key != 0ull ? fr_key_depth(key) : __uint_as_float(FR_BACKGROUND_DEPTH_BITS);
    *t = key != 0ull ? (float)fr_key_triangle(key) : -1.0f;
  };
  if (i < n4) {
    const ulonglong2 a = reinterpret_cast<const ulonglong2*>(keys)[2 * i], b = reinterpret_cast<const ulonglong2*>(keys)[2 * i + 1];
    float4 d, t;
    decode(a.x, &d.x, &t.x);
    decode(a.y, &d.y, &t.y);
    decode(b.x, &d.z, &t.z);
    decode(b.y, &d.w, &t.w);
    reinterpret_cast<float4*>(depth)[i] = d;
    reinterpret_cast<float4*>(tri_ind)[i] = t;
  }
  if (i < (n & 3ull)) decode(keys[4 * n4 + i], depth + 4 * n4 + i, tri_ind + 4 * n4 + i);
  FR_MARK_MAX(5);
}

// Backward (render_depth_op.cc:325-368).  vertex_grad must be zero on entry (the API memsets it).
__global__ void __launch_bounds__(kRasterThreads)
render_backward_kernel(const float* __restrict__ depth_grad, const float* __restrict__ tri,
                       const float* __restrict__ tri_ind, float* __restrict__ vertex_grad, int nver, int ntri, int npix,
                       const float* __restrict__ mask_grad, const float* __restrict__ im_gray,
                       const float* __restrict__ raw_depth) {
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
      float g;
      if (raw_depth == nullptr) {
        g = __ldg(depth_grad + o);
      } else {
        // fused rendering layer: the op's depth_grad is what autodiff would have summed from its two consumers,
        // max(depth, 1e-6) (passes where depth >= 1e-6) and clip(depth, 1e-6, 1) * im_gray (passes inside the clip range)
        const float dr = __ldg(raw_depth + o);
        g = 0.0f;
        if (depth_grad != nullptr && dr >= 1e-6f) g = __ldg(depth_grad + o);
        if (mask_grad != nullptr && dr >= 1e-6f && dr <= 1.0f) {
          const float gm = __ldg(mask_grad + o);
          g = __fadd_rn(g, (im_gray != nullptr) ? __fmul_rn(gm, __ldg(im_gray + o)) : gm);
        }
      }
      share = __fdiv_rn(__fmul_rn(g, 1.0f), 3.0f);  // (g * 1.0f) / 3.0f, :361
    }
  }
  // warp aggregation: lanes that hit the same triangle add their shares once (lane order => deterministic
  // within the warp); skipped when every lane has its own triangle, the common case for sub-pixel meshes.
  // Pixels of one large triangle sit next to each other in a row: only when some lane shares its triangle with its right-hand
  // neighbour is the (slow) match_any / shuffle aggregation worth running; sub-pixel meshes skip it (batch-256 backward:
  // 212 -> 190 us without it).
  const int t_next = __shfl_down_sync(0xFFFFFFFFu, t, 1);       // (every lane takes part: no short-circuit in front of it)
  const bool share_next = t >= 0 && t_next == t && lane != 31u;
  unsigned peers = 1u << lane;
  const bool aggregate = __any_sync(0xFFFFFFFFu, share_next);
  if (aggregate) peers = __match_any_sync(0xFFFFFFFFu, t);
  const bool leader = (peers & ((1u << lane) - 1u)) == 0u;
  if (aggregate) {
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
