// Tile rasterizer: the visibility pass of the z-buffer renderer on shared-memory tiles of the mesh.
//
// Replaces the reference's kernels 2 and 3 (render_depth_op.cu.cc:68-125 per-triangle setup into 13 doubles of global
// scratch, :176-237 racy per-triangle raster loop) and reproduces the CPU op's semantics (render_depth_op.cc:263-316)
// bit for bit.
//
// Unit of work: one CLUSTER of the mesh table (mesh_table.h: <= 128 vertices, <= 256 triangles with 8-bit local vertex
// slots) x NF faces per thread block.
//   stage  the cluster's 16-byte vertex records {x, y, z, snap code} of NF faces are copied ONCE from global memory
//          (coalesced: records are stored by rank, i.e. in cluster order) into shared memory, plus the snap codes again as
//          one uint4 per (vertex, 4 faces) so that the cull fetches four faces with one shared load per vertex;
//   cull   thread = triangle: per face two packed min3 / max3 on the codes and five integer operations decide whether the
//          reference's integer bounding box holds a pixel at all (raster_core.h: fr_code_nonempty).  ~53 % of the
//          sub-pixel BFM triangles stop here.  Each thread keeps two bit masks over the faces (kept / kept with a
//          one-pixel box); a warp scan and ONE shared atomic per warp reserve the survivors' slots in a block-wide list
//          of 16-bit ids (triangle, face): one-pixel boxes from the front, larger boxes from the back;
//   draw   the dense list is drained one survivor per thread and trip with every lane busy: three 16-byte shared loads,
//          image-range check (the other half of the reference's cull, :282), flat depth, the CERTIFIED FAST inside test
//          (raster_core.h: fr_fast_classify -- three cross products, no division; falls back to the literal PointInTri
//          whenever the reference's rounding could matter) and one 64-bit atomicMax of the packed (depth, index) key per
//          covered pixel.  One-pixel survivors come first, so whole warps run the straight-line single-pixel path.
// No per-triangle global gathers, no float -> int index conversion or validation (done once, in the table), no division.
#ifndef FR_RASTER_TILE_CUH_
#define FR_RASTER_TILE_CUH_

#include "fr_common.cuh"
#include "mesh_table.h"
#include "raster_core.h"

namespace fr {
namespace rt {

constexpr int kTileThreads = kClusterTris;       // one thread per triangle slot of a cluster
constexpr int kMaxExtent = 16000;                // code fields stay below 2^15 (fr_code_nonempty)

template <int NF>
struct alignas(16) TileSmem {
  static_assert(NF == 8 || NF == 16, "faces per tile");
  float4 rec[NF][kClusterVerts + 1];             // vertex records of the staged faces; one record of padding per face row:
                                                 // the draw phase reads one slot of MANY faces at once (bank spread)
  uint4 code4[NF / 4][kClusterVerts];            // snap codes, four consecutive faces per vertex slot
  uint4 tri[kClusterTris];                       // { 16 * slot of vertex 1, 2, 3, low word of the visibility key }
  unsigned short queue[kClusterTris * NF];       // survivor ids: (triangle << log2 NF) | face
  unsigned count;                                // one-pixel survivors | others << 16
  unsigned pad[3];
};

struct TableView {                               // device view of a mesh table blob
  const int32_t* cluster_rank;                   // [nclusters][128] record index (rank) of every slot, -1 = unused
  const int32_t* tri_begin;                      // [nclusters + 1]
  const uint2* tri_entry;                        // [ntri_slots]
};
__device__ __forceinline__ TableView table_view(const unsigned char* table) {
  const MeshTableHeader* h = reinterpret_cast<const MeshTableHeader*>(table);
  TableView v;
  v.cluster_rank = reinterpret_cast<const int32_t*>(table + mesh_off_cluster_rank(*h));
  v.tri_begin = reinterpret_cast<const int32_t*>(table + h->off_tri_begin);
  v.tri_entry = reinterpret_cast<const uint2*>(table + h->off_tri);
  return v;
}

// shared-memory accesses through 32-bit shared-window addresses (one register per stream in the hot loops)
__device__ __forceinline__ uint32_t shared_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}

// Survivor ids of one triangle into the block-wide list: face J's id goes to *ps (one-pixel box) or *pm (larger box) when
// the face is kept, and the pointer it used moves on (ps up, pm down; byte addresses in the shared window).  Spelled in PTX
// so that every face costs three predicate tests, one select, one store and two pointer bumps -- the compiler's own
// lowering of the equivalent C++ rebuilds the addresses for every face.
template <int NFACES, int J>
__device__ __forceinline__ void emit_faces(unsigned kmask, unsigned smask, unsigned mmask, uint32_t& ps, uint32_t& pm, unsigned idt) {
  if constexpr (J < NFACES) {
    asm volatile(
        "{\n\t.reg .pred k, s, m;\n\t.reg .b32 t, a;\n\t.reg .b16 v;\n\t"
        "and.b32 t, %2, %5;\n\tsetp.ne.b32 k, t, 0;\n\t"
        "and.b32 t, %3, %5;\n\tsetp.ne.b32 s, t, 0;\n\t"
        "and.b32 t, %4, %5;\n\tsetp.ne.b32 m, t, 0;\n\t"
        "selp.b32 a, %0, %1, s;\n\t"
        "add.u32 t, %6, %7;\n\tcvt.u16.u32 v, t;\n\t"
        "@k st.shared.u16 [a], v;\n\t"
        "@s add.u32 %0, %0, 2;\n\t"
        "@m sub.u32 %1, %1, 2;\n\t}"
        : "+r"(ps), "+r"(pm)
        : "r"(kmask), "r"(smask), "r"(mmask), "n"(1u << J), "r"(idt), "n"(J)
        : "memory");
    emit_faces<NFACES, J + 1>(kmask, smask, mmask, ps, pm, idt);
  }
}

// keys[off] = max(keys[off], key) on a global-window base address and a 32-bit element offset: one 64-bit multiply-add
// and one fire-and-forget reduction.
__device__ __forceinline__ void red_max_key(unsigned long long keys_global, unsigned off, unsigned long long key) {
  asm volatile("{\n\t.reg .u64 a;\n\tmad.wide.u32 a, %1, 8, %0;\n\tred.global.max.u64 [a], %2;\n\t}" ::"l"(keys_global), "r"(off), "l"(key)
               : "memory");
}

// One survivor: everything after the cull.  keys0 = visibility keys of the tile's face 0.  Returns false when some pixel
// of the box could not be certified by the fast inside test: the caller then redraws the survivor with draw_literal
// (the packed-key maximum is idempotent, so pixels already drawn may be drawn again).
template <int NF, bool kSingle>
__device__ __forceinline__ bool draw_fast(const TileSmem<NF>& s, uint32_t a_rec, unsigned entry, unsigned long long keys0, int npix,
                                          int width, uint32_t limit) {
  constexpr int kFaceBits = NF == 16 ? 4 : 3;
  const uint4 te = s.tri[entry >> kFaceBits];
  const uint32_t fb = a_rec + (entry & (unsigned)(NF - 1)) * (uint32_t)((kClusterVerts + 1) * 16);   // the face's row of records
  const float4 r1 = lds_f4(fb + te.x), r2 = lds_f4(fb + te.y), r3 = lds_f4(fb + te.z);
  const uint32_t c1 = __float_as_uint(r1.w), c2 = __float_as_uint(r2.w), c3 = __float_as_uint(r3.w);
  const uint32_t lo = fr_code_lo(__vimin3_u16x2(c1, c2, c3));
  const uint32_t hi = kSingle ? lo : fr_code_hi(__vimax3_u16x2(c1, c2, c3));
  if (!fr_box_in_image(lo, hi, limit)) return true;               // render_depth_op.cc:282, image-range half
  const float h = fr_tri_depth(r1.z, r2.z, r3.z);
  if (!fr_depth_draws(h)) return true;
  const unsigned long long key =
      ((unsigned long long)fr_float_order_bits(h) << 32) | (unsigned long long)(te.w | (__float_as_uint(h) == 0x80000000u ? 1u : 0u));
  const int xb = (int)(lo & 0xFFFFu), yb = (int)(lo >> 16);       // biased by +1
  // pixel (x, y) of face f: keys0[f * npix + y * width + x], a 32-bit offset (the API bounds batch * npix)
  const unsigned face_off = (entry & (unsigned)(NF - 1)) * (unsigned)npix - (unsigned)(width + 1);
  if (kSingle) {
    FrTriFast ff;
    fr_fast_setup(r1.x, r1.y, r2.x, r2.y, r3.x, r3.y, fr_fast_tol(1), &ff);
    const int in = fr_fast_classify(&ff, xb - 1, yb - 1);
    if (in > 0) red_max_key(keys0, face_off + (unsigned)(yb * width + xb), key);
    return in >= 0;
  } else {
    // larger box: the three numerators as plane equations around the box origin, two fused multiply-adds each per pixel
    const int w = (int)(hi & 0xFFFFu) - xb, h = (int)(hi >> 16) - yb;   // width - 1, height - 1
    FrTriPlanes pl;
    fr_planes_setup(r1.x, r1.y, r2.x, r2.y, r3.x, r3.y, xb - 1, yb - 1, fr_fast_tol(max(w, h) + 1), &pl);
    bool certified = true;
    unsigned off = face_off + (unsigned)(yb * width + xb);
    const float fw = (float)w;
    float dx = 0.0f, dy = 0.0f;
    for (int left = (w + 1) * (h + 1); left > 0; --left) {        // flat walk over the box
      const int in = fr_planes_classify(&pl, dx, dy);
      if (in > 0) red_max_key(keys0, off, key);
      certified = certified && in >= 0;
      const bool wrap = dx >= fw;
      off += wrap ? (unsigned)(width - w) : 1u;
      dy += wrap ? 1.0f : 0.0f;
      dx = wrap ? 0.0f : dx + 1.0f;
    }
    return certified;
  }
}

// The same survivor with the reference's literal double-precision inside test (raster_core.h: fr_point_in_tri).
template <int NF>
__device__ __forceinline__ void draw_literal(const TileSmem<NF>& s, unsigned entry, unsigned long long keys0, int npix, int width,
                                             uint32_t limit) {
  constexpr int kFaceBits = NF == 16 ? 4 : 3;
  const uint4 te = s.tri[entry >> kFaceBits];
  const unsigned f = entry & (unsigned)(NF - 1);
  const float4 r1 = s.rec[f][te.x >> 4], r2 = s.rec[f][te.y >> 4], r3 = s.rec[f][te.z >> 4];
  const uint32_t c1 = __float_as_uint(r1.w), c2 = __float_as_uint(r2.w), c3 = __float_as_uint(r3.w);
  const uint32_t lo = fr_code_lo(__vimin3_u16x2(c1, c2, c3)), hi = fr_code_hi(__vimax3_u16x2(c1, c2, c3));
  if (!fr_box_in_image(lo, hi, limit)) return;
  const float h = fr_tri_depth(r1.z, r2.z, r3.z);
  if (!fr_depth_draws(h)) return;
  const unsigned long long key =
      ((unsigned long long)fr_float_order_bits(h) << 32) | (unsigned long long)(te.w | (__float_as_uint(h) == 0x80000000u ? 1u : 0u));
  FrTriEdge e;
  fr_tri_edge_setup(r1.x, r1.y, r2.x, r2.y, r3.x, r3.y, &e);
  const unsigned face_off = f * (unsigned)npix;
  const int x0 = (int)(lo & 0xFFFFu) - 1, y0 = (int)(lo >> 16) - 1, x1 = (int)(hi & 0xFFFFu) - 1, y1 = (int)(hi >> 16) - 1;
  for (int y = y0; y <= y1; ++y)
    for (int x = x0; x <= x1; ++x)
      if (fr_point_in_tri(&e, x, y)) red_max_key(keys0, face_off + (unsigned)(y * width + x), key);
}

// ---- building blocks shared by the stand-alone kernel below and the fused reconstruction epilogue (recon_f16.cuh) --------
// The cluster's triangle list -> s.tri, by `nthreads` threads.  Returns the number of triangles (0: loose vertices only).
template <int NF>
__device__ __forceinline__ int load_tris(TileSmem<NF>& s, const TableView& tv, int cluster, int tid, int nthreads) {
  const int tb = __ldg(tv.tri_begin + cluster);
  const int ntri_c = __ldg(tv.tri_begin + cluster + 1) - tb;
  for (int i = tid; i < ntri_c; i += nthreads) {
    const uint2 e = __ldg(tv.tri_entry + tb + i);
    // byte offsets of the three vertex slots within a face's records (and within a code4 row); fr_pack_key's low word
    // without the zero-sign bit
    s.tri[i] = make_uint4((e.x & 0xFFu) << 4, ((e.x >> 8) & 0xFFu) << 4, ((e.x >> 16) & 0xFFu) << 4, (0x7FFFFFFFu - e.y) << 1);
  }
  return ntri_c;
}

// Cull of one triangle per lane (whole warp, converged) over the NQ face quads starting at quad q0; `live` = the faces of
// those quads that exist (bit j = face 4 q0 + j) -- 0 for a lane without a triangle.  Survivors go to the block-wide list.
template <int NF, int NQ>
__device__ __forceinline__ void cull_triangle(TileSmem<NF>& s, int t, int q0, unsigned live, int lane) {
  constexpr int kFaceBits = NF == 16 ? 4 : 3;
  constexpr int kCap = kClusterTris * NF;
  const uint4 te = s.tri[live != 0u ? t : 0];
  const unsigned char* cbase = reinterpret_cast<const unsigned char*>(&s.code4[0][0]) + q0 * (kClusterVerts * 16);
  unsigned kmask = 0u, smask = 0u;                                // kept / kept with a one-pixel box
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const uint4 a = *reinterpret_cast<const uint4*>(cbase + q * (kClusterVerts * 16) + te.x);
    const uint4 b = *reinterpret_cast<const uint4*>(cbase + q * (kClusterVerts * 16) + te.y);
    const uint4 c = *reinterpret_cast<const uint4*>(cbase + q * (kClusterVerts * 16) + te.z);
    const uint32_t e1[4] = {a.x, a.y, a.z, a.w}, e2[4] = {b.x, b.y, b.z, b.w}, e3[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t mn = __vimin3_u16x2(e1[j], e2[j], e3[j]), mx = __vimax3_u16x2(e1[j], e2[j], e3[j]);
      if (fr_code_nonempty(mn, mx)) kmask |= 1u << (4 * q + j);
      if (fr_code_single(mn, mx)) smask |= 1u << (4 * q + j);
    }
  }
  kmask &= live;
  smask &= kmask;
  const unsigned mine = (unsigned)__popc(smask) | ((unsigned)__popc(kmask ^ smask) << 16);
  unsigned incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= d) incl += up;
  }
  const unsigned total = __shfl_sync(0xFFFFFFFFu, incl, 31);
  if (total != 0u) {                                              // warp-uniform
    unsigned base = 0u;
    if (lane == 31) base = atomicAdd(&s.count, total);
    base = __shfl_sync(0xFFFFFFFFu, base, 31) + (incl - mine);
    // one-pixel survivors fill the list from the front, the others from the back: byte addresses, one predicated store
    // and one predicated pointer bump per kept face
    const uint32_t a_q = shared_addr(s.queue);
    uint32_t ps = a_q + 2u * (base & 0xFFFFu), pm = a_q + 2u * ((unsigned)kCap - 1u - (base >> 16));
    emit_faces<4 * NQ, 0>(kmask, smask, kmask ^ smask, ps, pm, ((unsigned)t << kFaceBits) | (unsigned)(4 * q0));
  }
}

// Drains the block-wide survivor list with `nthreads` threads (one-pixel survivors first), one survivor per thread and
// trip; keys = visibility keys of the tile's face 0.
template <int NF>
__device__ __forceinline__ void draw_list(const TileSmem<NF>& s, int tid, int nthreads, unsigned long long* __restrict__ keys, int npix,
                                          int width, int height) {
  constexpr int kCap = kClusterTris * NF;
  const unsigned cnt = s.count;
  const int n_single = (int)(cnt & 0xFFFFu), n_multi = (int)(cnt >> 16), n_total = n_single + n_multi;
  const uint32_t limit = (uint32_t)width | ((uint32_t)height << 16);
  const uint32_t a_rec = shared_addr(&s.rec[0][0]);
  const unsigned long long keys0 = (unsigned long long)__cvta_generic_to_global(keys);
  unsigned redo = 0u;                                             // trips whose survivor needs the literal inside test
  int trip = 0;
  // the larger boxes first (the long jobs), then the one-pixel boxes: the warps finish closer together
  for (int i = tid; i < n_total; i += nthreads, ++trip) {
    bool certified;
    if (i < n_multi) certified = draw_fast<NF, false>(s, a_rec, s.queue[kCap - 1 - i], keys0, npix, width, limit);
    else certified = draw_fast<NF, true>(s, a_rec, s.queue[i - n_multi], keys0, npix, width, limit);
    if (!certified) redo |= 1u << trip;
  }
  while (redo != 0u) {                                            // rare (pixel centres on an edge, slivers, degenerate triangles)
    const int k = __ffs((int)redo) - 1;
    redo &= redo - 1u;
    const int i = tid + k * nthreads;
    draw_literal<NF>(s, s.queue[i < n_multi ? kCap - 1 - i : i - n_multi], keys0, npix, width, limit);
  }
}

// ---- stand-alone visibility pass -----------------------------------------------------------------------------------------
// grid = (clusters, ceil(batch / NF)); rec = vertex records by rank [batch][nver]; keys cleared by a predecessor.
#ifndef FR_TILE_MINB16
#define FR_TILE_MINB16 4
#endif
#ifndef FR_TILE_MINB8
#define FR_TILE_MINB8 5
#endif
template <int NF>
__global__ void __launch_bounds__(kTileThreads, NF == 16 ? FR_TILE_MINB16 : FR_TILE_MINB8)
raster_tile_keys_kernel(const float4* __restrict__ rec, const unsigned char* __restrict__ table, unsigned long long* __restrict__ keys,
                        int batch, int nver, int height, int width) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<NF>& s = *reinterpret_cast<TileSmem<NF>*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31;
  const int cluster = blockIdx.x, b0 = blockIdx.y * NF;
  const TableView tv = table_view(table);
  const int ntri_c = load_tris(s, tv, cluster, tid, kTileThreads);   // static data: no need to wait
  if (ntri_c == 0) return;                                        // a cluster of loose vertices only (counts as triggered)
  if (tid == 0) s.count = 0u;
  pdl_wait();      // records and cleared keys of the producing kernels are complete

  // ---- stage: thread = (vertex slot, half of the faces); all loads in flight before the first use
  {
    constexpr int kFH = NF / 2;                                   // faces per thread
    const int slot = tid & (kClusterVerts - 1), half = tid >> 7;
    const int r = __ldg(tv.cluster_rank + (size_t)cluster * kClusterVerts + slot);
    if (r >= 0) {
      float4 v[kFH];
#pragma unroll
      for (int j = 0; j < kFH; ++j) {
        const unsigned b = (unsigned)min(b0 + half * kFH + j, batch - 1);
        v[j] = __ldg(rec + (b * (unsigned)nver + (unsigned)r));
      }
#pragma unroll
      for (int j = 0; j < kFH; ++j) s.rec[half * kFH + j][slot] = v[j];
#pragma unroll
      for (int q = 0; q < kFH / 4; ++q)
        s.code4[half * (kFH / 4) + q][slot] = make_uint4(__float_as_uint(v[4 * q].w), __float_as_uint(v[4 * q + 1].w),
                                                         __float_as_uint(v[4 * q + 2].w), __float_as_uint(v[4 * q + 3].w));
    }
  }
  __syncthreads();

  // ---- cull: thread = triangle, all NF faces (warps entirely beyond the cluster's list skip to the barrier)
  if ((tid & ~31) < ntri_c) {
    const int nlive = min(NF, batch - b0);                        // faces beyond the batch / lanes beyond the list: not live
    cull_triangle<NF, NF / 4>(s, tid, 0, tid < ntri_c ? ((1u << nlive) - 1u) : 0u, lane);
  }
  __syncthreads();

  // ---- draw
  draw_list(s, tid, kTileThreads, keys + (size_t)b0 * (size_t)(height * width), height * width, width, height);
  // The resolve pass may become resident when every block has come this far: triggering at the start would park its blocks
  // (24 KB of shared memory each when normals / textures are requested) on the SMs this kernel needs four 53 KB blocks of
  // -- measured 1.38 ms vs 1.12 ms per batch-256 forward + backward.
  pdl_trigger();
}

}  // namespace rt
}  // namespace fr
#endif  // FR_RASTER_TILE_CUH_
