// Cluster rasterizer: the visibility pass of the z-buffer renderer with a cluster's vertices staged in shared memory.
//
// Replaces the reference's kernels 2 and 3 (render_depth_op.cu.cc:68-125 per-triangle setup into 13 doubles of global
// scratch, :176-237 racy per-triangle raster loop) and reproduces the CPU op's semantics (render_depth_op.cc:263-316)
// bit for bit; the arithmetic is raster_core.h's, untouched.
//
// Unit of work: one CLUSTER of the mesh table (mesh_table.h: <= 128 vertices, <= 256 triangles with 8-bit local
// indices) x up to 32 faces.  The projected vertices of the cluster are written ONCE per face into a shared-memory
// stage -- structure of arrays [face][vertex slot]: x, y, z and the one-word snap code -- by the tensor-core
// reconstruction epilogue that has just computed them (recon_f16.cuh, the FR_CLUSTER_TILES flavour of the fused call: the
// vertices never travel through global memory).  A stage is then rasterized in two block-wide phases:
//   cull   every warp walks (32 triangles) x (8 faces) items with lane = triangle: three conflict-free shared loads of
//          snap codes per face and a handful of packed integer operations decide the reference's bounding-box cull
//          (:276-282); the survivors' 16-bit ids (local triangle, face) go to ONE block-wide list, one-pixel boxes from
//          the front and larger boxes from the back (a warp scan + one shared atomic per item reserves the slots);
//   draw   the dense list is drained one survivor per thread and trip with every lane busy: box from the codes again,
//          nine shared loads, flat depth, FP64 edge setup, FP64 inside tests over the box, one 64-bit atomicMax of the
//          packed (depth, index) key per covered pixel.  One-pixel survivors come first, so whole warps run one trip.
#ifndef FR_RASTER_CLUSTER_CUH_
#define FR_RASTER_CLUSTER_CUH_

#include "fr_common.cuh"
#include "mesh_table.h"
#include "raster_core.h"

namespace fr {
namespace rc {

#ifndef FR_ITEM_FACES
#define FR_ITEM_FACES 8
#endif
#ifndef FR_DRAW_DYNAMIC
#define FR_DRAW_DYNAMIC 0
#endif
constexpr int kItemFaces = FR_ITEM_FACES;      // faces a warp culls per item (32 triangles x 8 faces)
constexpr int kQueueCap = 8192;    // kClusterTris x 32 faces: every (triangle, face) pair of a stage fits -- no overflow path

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
// Block-wide survivor list of one stage: ids (local triangle << 5 | face), one-pixel boxes from the front, the others
// from the back; count = one-pixel survivors | others << 16 (reserved with one shared atomic per warp and item).
struct StageQueue {                // 16 KB
  unsigned short id[kQueueCap];
  unsigned count;
  unsigned next;                   // draw phase: first list position not yet claimed by a warp
  unsigned pad[2];
};

// shared-memory accesses through 32-bit shared-window addresses with immediate offsets (keeps the address arithmetic of
// the hot loops to one register per stream)
template <int OFF>
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
  return v;
}
template <int OFF>
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFF));
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ uint32_t shared_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// snap codes of a triangle's three vertices for kItemFaces consecutive faces (one 512-byte face row apart)
template <int J>
__device__ __forceinline__ void load_codes(uint32_t a1, uint32_t a2, uint32_t a3, uint32_t* e1, uint32_t* e2, uint32_t* e3) {
  if constexpr (J < kItemFaces) {
    e1[J] = lds_u32<J * kClusterVerts * 4>(a1);
    e2[J] = lds_u32<J * kClusterVerts * 4>(a2);
    e3[J] = lds_u32<J * kClusterVerts * 4>(a3);
    load_codes<J + 1>(a1, a2, a3, e1, e2, e3);
  }
}

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

// Phase "cull" of one staged cluster (all warps, then a block barrier): ntri_c triangles x nfaces (<= NF) faces.
// q.count must be 0 on entry.
template <int NF>
__device__ __forceinline__ void cull_stage(const Stage<NF>& st, const TriList& tl, StageQueue& q, int ntri_c, int nfaces, int warp,
                                           int nwarps, int lane, int width, int height) {
  static_assert(NF % kItemFaces == 0, "a stage holds whole items");
  const uint32_t limit = (uint32_t)width | ((uint32_t)height << 16);
  const uint32_t a_code = shared_addr(&st.code[0][0]), a_local = shared_addr(tl.local), a_ids = shared_addr(q.id);
  const int nchunks = (ntri_c + 31) >> 5;
  const int nfq = (nfaces + kItemFaces - 1) / kItemFaces;
  int chunk = warp, fq = 0;                     // items dealt round-robin, chunk index fastest
  while (chunk >= nchunks && fq < nfq) {
    chunk -= nchunks;
    ++fq;
  }
  while (fq < nfq) {
    const int t = (chunk << 5) + lane;
    const bool valid = t < ntri_c;
    const uint32_t w = valid ? lds_u32<0>(a_local + 4u * (uint32_t)t) : 0u;
    const int f0 = fq * kItemFaces;
    const uint32_t fbase = a_code + (uint32_t)f0 * (kClusterVerts * 4u);
    const uint32_t a1 = fbase + 4u * (w & 0xFFu), a2 = fbase + 4u * ((w >> 8) & 0xFFu), a3 = fbase + 4u * ((w >> 16) & 0xFFu);
    uint32_t e1[kItemFaces], e2[kItemFaces], e3[kItemFaces];
    load_codes<0>(a1, a2, a3, e1, e2, e3);      // all loads in flight before the first use
    // faces beyond nfaces / lanes beyond ntri_c: one mask instead of a test per face
    const unsigned live = valid ? ((nfaces - f0 >= kItemFaces) ? ((1u << kItemFaces) - 1u) : ((1u << (nfaces - f0)) - 1u)) : 0u;
    unsigned kmask = 0u, smask = 0u;            // kept / kept with a one-pixel box
#pragma unroll
    for (int j = 0; j < kItemFaces; ++j) {
      uint32_t lo, hi;
      if (fr_code_keep(e1[j], e2[j], e3[j], limit, &lo, &hi)) kmask |= 1u << j;
      if (lo == hi) smask |= 1u << j;
    }
    kmask &= live;
    smask &= kmask;
    const unsigned mmask = kmask ^ smask;
    // slots: inclusive warp scan of (one-pixel count | other count << 16), one shared atomic per warp
    const unsigned mine = (unsigned)__popc(smask) | ((unsigned)__popc(mmask) << 16);
    unsigned incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= d) incl += up;
    }
    const unsigned total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    if (total != 0u) {                          // warp-uniform
      unsigned base = 0u;
      if (lane == 0) base = atomicAdd(&q.count, total);
      base = __shfl_sync(0xFFFFFFFFu, base, 0) + (incl - mine);
      uint32_t ps = a_ids + 2u * (base & 0xFFFFu);
      uint32_t pm = a_ids + 2u * ((uint32_t)kQueueCap - 1u - (base >> 16));
      const uint32_t idt = ((uint32_t)t << 5) | (uint32_t)f0;
#pragma unroll
      for (int j = 0; j < kItemFaces; ++j) {
        if ((smask >> j) & 1u) {
          sts_u16(ps, idt + j);
          ps += 2u;
        }
        if ((mmask >> j) & 1u) {
          sts_u16(pm, idt + j);
          pm -= 2u;
        }
      }
    }
    chunk += nwarps;
    while (chunk >= nchunks && fq < nfq) {
      chunk -= nchunks;
      ++fq;
    }
  }
}

// Phase "draw" (after the block barrier that follows cull_stage): the survivor list, one survivor per thread and trip.
// (FR_DRAW_DYNAMIC=1 lets the warps claim 32 list positions at a time instead of the static split -- measured slower on
// B200: 85 -> 92.6 us for the stand-alone kernel; the shared atomic per trip costs more than the imbalance it removes.)
// keys0 = visibility keys of the stage's face 0 (faces are consecutive, npix apart).  q.next must be 0 on entry.
template <int NF>
__device__ __forceinline__ void draw_stage(const Stage<NF>& st, const TriList& tl, StageQueue& q, int tid, int nthreads, int lane,
                                           unsigned long long* __restrict__ keys0, int npix, int width, int height) {
  const uint32_t a_x = shared_addr(&st.x[0][0]), a_local = shared_addr(tl.local), a_ids = shared_addr(q.id);
  constexpr int kPlane = NF * kClusterVerts * 4;       // bytes between the x, y, z and code arrays of a stage
  const unsigned count = q.count;
  const int n_single = (int)(count & 0xFFFFu), n_total = n_single + (int)(count >> 16);
#if FR_DRAW_DYNAMIC
  for (;;) {
    int i = 0;
    if (lane == 0) i = (int)atomicAdd(&q.next, 32u);
    i = __shfl_sync(0xFFFFFFFFu, i, 0) + lane;
    if (i - lane >= n_total) break;               // warp-uniform
    if (i >= n_total) continue;
#else
  for (int i = tid; i < n_total; i += nthreads) {
#endif
    const int pos = (i < n_single) ? i : kQueueCap - 1 - (i - n_single);
    const uint32_t id = lds_u16(a_ids + 2u * (uint32_t)pos);
    const uint32_t t = id >> 5, f = id & 31u;
    const uint32_t w = lds_u32<0>(a_local + 4u * t);
    const uint32_t tri_index = lds_u32<kClusterTris * 4>(a_local + 4u * t);      // TriList::id follows TriList::local
    const uint32_t fb = a_x + f * (kClusterVerts * 4u);
    const uint32_t o1 = fb + 4u * (w & 0xFFu), o2 = fb + 4u * ((w >> 8) & 0xFFu), o3 = fb + 4u * ((w >> 16) & 0xFFu);
    const uint32_t c1 = lds_u32<3 * kPlane>(o1), c2 = lds_u32<3 * kPlane>(o2), c3 = lds_u32<3 * kPlane>(o3);
    const float z1 = lds_f32<2 * kPlane>(o1), z2 = lds_f32<2 * kPlane>(o2), z3 = lds_f32<2 * kPlane>(o3);
    const float x1 = lds_f32<0>(o1), x2 = lds_f32<0>(o2), x3 = lds_f32<0>(o3);
    const float y1 = lds_f32<kPlane>(o1), y2 = lds_f32<kPlane>(o2), y3 = lds_f32<kPlane>(o3);
    const float h = fr_tri_depth(z1, z2, z3);
    if (!fr_depth_draws(h)) continue;
    uint32_t lo, hi;
    fr_code_box(c1, c2, c3, &lo, &hi);                           // (kept in the cull phase: only the box is needed)
    FrTriEdge e;
    fr_tri_edge_setup(x1, y1, x2, y2, x3, y3, &e);
    const unsigned long long key = fr_pack_key(h, (int)tri_index);
    unsigned long long* kb = keys0 + (size_t)f * npix;
    const int x0 = (int)(lo & 0xFFFFu) - 1, y0 = (int)(lo >> 16) - 1;
    const int xe = (int)(hi & 0xFFFFu) - 1, ye = (int)(hi >> 16) - 1;
    int x = x0, y = y0;
    while (y <= ye) {                                            // flat walk over the box (one trip for the one-pixel class)
      if (fr_point_in_tri(&e, x, y)) atomicMax(kb + (y * width + x), key);
      if (++x > xe) {
        x = x0;
        ++y;
      }
    }
  }
}

// Block barrier over `nthreads` threads (a multiple of 32) on hardware barrier `id`.
__device__ __forceinline__ void block_barrier(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Visibility pass of one staged cluster by the `nwarps` warps that share barrier `bar` (all of them must call this).
// The stage, the triangle list and q.count == 0 must be in place before (a barrier separates them from the callers'
// writes); on return every thread is past the last read of stage / list and q.count is 0 again.
template <int NF>
__device__ __forceinline__ void raster_stage(const Stage<NF>& st, const TriList& tl, StageQueue& q, int ntri_c, int nfaces, int warp,
                                             int nwarps, int lane, int bar, unsigned long long* __restrict__ keys0, int npix, int width,
                                             int height) {
  cull_stage(st, tl, q, ntri_c, nfaces, warp, nwarps, lane, width, height);
  block_barrier(bar, nwarps * 32);
  draw_stage(st, tl, q, warp * 32 + lane, nwarps * 32, lane, keys0, npix, width, height);
  block_barrier(bar, nwarps * 32);
  if (warp == 0 && lane == 0) {                  // visible to the next stage's cull through the caller's stage barrier
    q.count = 0u;
    q.next = 0u;
  }
}

}  // namespace rc
}  // namespace fr
#endif  // FR_RASTER_CLUSTER_CUH_
