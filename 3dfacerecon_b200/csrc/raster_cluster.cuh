// Cluster rasterizer: the visibility pass of the z-buffer renderer with a cluster's vertices staged in shared memory.
//
// Replaces the reference's kernels 2 and 3 (render_depth_op.cu.cc:68-125 per-triangle setup into 13 doubles of global
// scratch, :176-237 racy per-triangle raster loop) and reproduces the CPU op's semantics (render_depth_op.cc:263-316)
// bit for bit; the arithmetic is raster_core.h's, untouched.
//
// Unit of work: one CLUSTER of the mesh table (mesh_table.h: <= 128 vertices, <= 256 triangles with 8-bit local
// indices) x up to 32 faces.  The projected vertices of the cluster are written ONCE per face into a shared-memory
// stage -- structure of arrays [face][vertex slot]: x, y, z and the one-word snap code -- either by the tensor-core
// reconstruction epilogue that has just computed them (recon_f16.cuh, the fused params -> depth-map path: the
// vertices never travel through global memory) or by one gather from the planar vertex tensor
// (raster_cluster_kernel below, the stand-alone render_depth op).  Then every warp walks (32 triangles) x (4 faces)
// items with lane = triangle: three conflict-free shared loads of snap codes and a handful of packed integer
// operations decide the reference's bounding-box cull (:276-282); survivors are appended to a warp-private ring
// (ballot + popc, no atomics), and whenever 32 of them are queued the warp drains them with all lanes busy: nine
// shared loads, flat depth, FP64 edge setup, FP64 inside tests over the bounding box, one 64-bit atomicMax of the
// packed (depth, index) key per covered pixel.
#ifndef FR_RASTER_CLUSTER_CUH_
#define FR_RASTER_CLUSTER_CUH_

#include "fr_common.cuh"
#include "mesh_table.h"
#include "raster_core.h"

namespace fr {
namespace rc {

constexpr int kItemFaces = 4;      // faces a warp culls per item (32 triangles x 4 faces)
constexpr int kQueue = 64;         // ring entries per warp (<= 31 left over + <= 32 appended)

template <int NF>                  // faces per stage (<= 32): 2 KB per face
struct Stage {
  static constexpr int kFaces = NF;
  float x[NF][kClusterVerts];
  float y[NF][kClusterVerts];
  float z[NF][kClusterVerts];
  uint32_t code[NF][kClusterVerts];
};
struct TriList {                   // 2 KB
  uint32_t local[kClusterTris];    // l1 | l2 << 8 | l3 << 16
  uint32_t id[kClusterTris];       // original triangle index
};
struct WarpQueue {                 // 640 B per warp
  uint2 box[kQueue];               // biased bounding box (fr_code_keep)
  unsigned short id[kQueue];       // local triangle << 5 | face of the stage
};

struct TableView {                 // device view of a mesh table blob
  const int32_t* cluster_vert;     // [nclusters][128]
  const int32_t* tri_begin;        // [nclusters + 1]
  const uint2* tri_entry;          // [ntri_slots]
  int nclusters;
};
__device__ __forceinline__ TableView table_view(const unsigned char* table) {
  const MeshTableHeader* h = reinterpret_cast<const MeshTableHeader*>(table);
  TableView v;
  v.cluster_vert = reinterpret_cast<const int32_t*>(table + h->off_vert);
  v.tri_begin = reinterpret_cast<const int32_t*>(table + h->off_tri_begin);
  v.tri_entry = reinterpret_cast<const uint2*>(table + h->off_tri);
  v.nclusters = h->nclusters;
  return v;
}

// The cluster's triangle entries -> shared memory (coalesced 8-byte loads), by `nthreads` threads.
__device__ __forceinline__ void load_tri_list(TriList& tl, const uint2* __restrict__ entries, int ntri_c, int tid, int nthreads) {
  for (int i = tid; i < ntri_c; i += nthreads) {
    const uint2 e = __ldg(entries + i);
    tl.local[i] = e.x;
    tl.id[i] = e.y;
  }
}

// One survivor: depth, FP64 edge setup, inside tests in a flat walk over the bounding box, packed-key atomicMax.
__device__ __forceinline__ void draw_one(float x1, float y1, float z1, float x2, float y2, float z2, float x3, float y3, float z3,
                                         uint2 box, int tri_index, unsigned long long* __restrict__ kb, int width) {
  const float h = fr_tri_depth(z1, z2, z3);
  if (!fr_depth_draws(h)) return;
  FrTriEdge e;
  fr_tri_edge_setup(x1, y1, x2, y2, x3, y3, &e);
  const unsigned long long key = fr_pack_key(h, tri_index);
  const int x0 = (int)(box.x & 0xFFFFu) - 1, y0 = (int)(box.x >> 16) - 1;
  const int xe = (int)(box.y & 0xFFFFu) - 1, ye = (int)(box.y >> 16) - 1;
  int x = x0, y = y0;
  while (y <= ye) {
    if (fr_point_in_tri(&e, x, y)) atomicMax(kb + (y * width + x), key);
    if (++x > xe) {
      x = x0;
      ++y;
    }
  }
}

// Drains `n` (<= 32) queued survivors starting at ring position `head`, one per lane.
template <int NF>
__device__ __forceinline__ void drain(const Stage<NF>& st, const TriList& tl, const WarpQueue& q, unsigned head, int n, int lane,
                                      unsigned long long* __restrict__ keys0, int npix, int width) {
  if (lane < n) {
    const unsigned pos = (head + (unsigned)lane) & (kQueue - 1);
    const unsigned id = q.id[pos];
    const uint2 box = q.box[pos];
    const unsigned t = id >> 5, f = id & 31u;
    const uint32_t w = tl.local[t];
    const unsigned l1 = w & 0xFFu, l2 = (w >> 8) & 0xFFu, l3 = (w >> 16) & 0xFFu;
    const float* fx = st.x[f];
    const float* fy = st.y[f];
    const float* fz = st.z[f];
    draw_one(fx[l1], fy[l1], fz[l1], fx[l2], fy[l2], fz[l2], fx[l3], fy[l3], fz[l3], box, (int)tl.id[t],
             keys0 + (size_t)f * npix, width);
  }
  __syncwarp();   // the ring slots just read may be overwritten by the next appends
}

// Visibility pass of one staged cluster: `ntri_c` triangles x `nfaces` (<= NF) faces, by `nwarps` warps of which this
// is number `warp`.  keys0 = visibility keys of the stage's face 0 (faces are consecutive, npix apart).
template <int NF>
__device__ __forceinline__ void raster_stage(const Stage<NF>& st, const TriList& tl, WarpQueue& q, int ntri_c, int nfaces, int warp,
                                             int nwarps, int lane, unsigned long long* __restrict__ keys0, int npix, int width,
                                             int height) {
  const uint32_t limit = (uint32_t)width | ((uint32_t)height << 16);
  const int nchunks = (ntri_c + 31) >> 5;
  const int nitems = nchunks * ((nfaces + kItemFaces - 1) / kItemFaces);
  const unsigned lt_mask = (1u << lane) - 1u;
  unsigned head = 0, tail = 0;      // ring positions (warp-uniform)
  for (int item = warp; item < nitems; item += nwarps) {
    const int chunk = item % nchunks, fq = item / nchunks;
    const int t = (chunk << 5) + lane;
    const bool valid = t < ntri_c;
    const uint32_t w = valid ? tl.local[t] : 0u;
    const unsigned l1 = w & 0xFFu, l2 = (w >> 8) & 0xFFu, l3 = (w >> 16) & 0xFFu;
    const int f0 = fq * kItemFaces;
    const uint32_t* c0 = st.code[f0];
    uint32_t e1[kItemFaces], e2[kItemFaces], e3[kItemFaces];
#pragma unroll
    for (int j = 0; j < kItemFaces; ++j) {     // all loads in flight before the first use
      e1[j] = c0[j * kClusterVerts + l1];
      e2[j] = c0[j * kClusterVerts + l2];
      e3[j] = c0[j * kClusterVerts + l3];
    }
#pragma unroll
    for (int j = 0; j < kItemFaces; ++j) {
      uint2 box;
      const bool keep = fr_code_keep(e1[j], e2[j], e3[j], limit, &box.x, &box.y) && valid && (f0 + j < nfaces);
      const unsigned m = __ballot_sync(0xFFFFFFFFu, keep);
      if (m != 0u) {                          // warp-uniform
        if (keep) {
          const unsigned pos = (tail + (unsigned)__popc(m & lt_mask)) & (kQueue - 1);
          q.box[pos] = box;
          q.id[pos] = (unsigned short)((t << 5) | (f0 + j));
        }
        tail += (unsigned)__popc(m);
        __syncwarp();
        if (tail - head >= 32u) {
          drain(st, tl, q, head, 32, lane, keys0, npix, width);
          head += 32u;
        }
      }
    }
  }
  if (tail != head) drain(st, tl, q, head, (int)(tail - head), lane, keys0, npix, width);
}

// ---------------------------------------------------------------------------------------------- stand-alone op
// render_depth on a planar vertex tensor [B,3,N] with a mesh table: grid (cluster stride loop, 32-face group).
// Block = kWarps warps; every thread stages vertex slot (tid % 128) for the faces (tid / 128), (tid / 128) + W/4, ...
constexpr int kWarps = 16;
constexpr int kThreads = kWarps * 32;
constexpr int kStageFaces = 32;
struct KernelSmem {
  Stage<kStageFaces> stage;
  TriList tris;
  WarpQueue queue[kWarps];
};

__global__ void __launch_bounds__(kThreads, 2)
raster_cluster_kernel(const float* __restrict__ vertex, const unsigned char* __restrict__ table, unsigned long long* __restrict__ keys,
                      int batch, int nver, int height, int width) {
  extern __shared__ __align__(16) unsigned char rc_smem[];
  KernelSmem& s = *reinterpret_cast<KernelSmem*>(rc_smem);
  const TableView tv = table_view(table);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b0 = blockIdx.y * kStageFaces;
  const int nfaces = min(kStageFaces, batch - b0);
  const int npix = height * width;
  const int v = tid & (kClusterVerts - 1), fsub = tid >> 7;        // 4 faces in flight per pass of the block
  pdl_trigger();
  pdl_wait();       // the visibility keys are cleared / the vertex tensor is written by the preceding work
  for (int c = blockIdx.x; c < tv.nclusters; c += gridDim.x) {
    const int tb = __ldg(tv.tri_begin + c), ntri_c = __ldg(tv.tri_begin + c + 1) - tb;
    if (ntri_c == 0) continue;                                      // a cluster of loose vertices: nothing to draw
    load_tri_list(s.tris, tv.tri_entry + tb, ntri_c, tid, kThreads);
    const int vid_raw = __ldg(tv.cluster_vert + (size_t)c * kClusterVerts + v);
    const int vid = vid_raw < 0 ? -1 : (int)((uint32_t)vid_raw & kVertIdMask);
    constexpr int kPass = kThreads / kClusterVerts;                 // faces per pass
    constexpr int kUnroll = kStageFaces / kPass;
    float x[kUnroll], y[kUnroll], z[kUnroll];
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      const int f = fsub + j * kPass;
      const bool ok = vid >= 0 && f < nfaces;
      const float* vb = vertex + (size_t)(b0 + (ok ? f : 0)) * 3 * nver + (ok ? vid : 0);
      x[j] = ok ? __ldg(vb) : 0.0f;
      y[j] = ok ? __ldg(vb + nver) : 0.0f;
      z[j] = ok ? __ldg(vb + 2 * (size_t)nver) : 0.0f;
    }
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      const int f = fsub + j * kPass;
      s.stage.x[f][v] = x[j];
      s.stage.y[f][v] = y[j];
      s.stage.z[f][v] = z[j];
      s.stage.code[f][v] = fr_snap_code(x[j], y[j], width, height);
    }
    __syncthreads();
    raster_stage(s.stage, s.tris, s.queue[warp], ntri_c, nfaces, warp, kWarps, lane, keys + (size_t)b0 * npix, npix, width, height);
    __syncthreads();
  }
}

}  // namespace rc
}  // namespace fr
#endif  // FR_RASTER_CLUSTER_CUH_
