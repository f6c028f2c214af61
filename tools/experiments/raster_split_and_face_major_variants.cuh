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
// nine 4-byte gathers from the three coordinate planes.  (A separate dense code array was measured: the cull does not
// get faster -- it is issue-bound, not L1-wavefront-bound -- and the survivors' record gathers lose their L1 hits.)  The same pass clears the face's visibility keys.  Each thread
// handles kSnapPerThread vertices with all loads issued before the first use (one vertex per thread is bound by CTA
// turnover, not by bandwidth).
constexpr int kSnapPerThread = 4;
__global__ void __launch_bounds__(kRasterThreads)
raster_pack_kernel(const float* __restrict__ vertex, float4* __restrict__ rec, unsigned long long* __restrict__ keys,
                   unsigned* __restrict__ counters, int nver, int npix, int width, int height) {
  const int b = blockIdx.y;
  if (blockIdx.x == 0 && b == 0 && threadIdx.x < 4) counters[threadIdx.x] = 0u;   // survivor-list lengths of the split pipeline
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
    if (n < nver)
      rec[(size_t)b * nver + n] = make_float4(x[j], y[j], z[j], __uint_as_float(fr_snap_code(x[j], y[j], width, height)));
  }
}

// Per-survivor state of phase B: decoded bounding box, depth key and the pixel-independent part of PointInTri.
struct RasterTri {
  FrTriEdge e;
  unsigned long long key;
  int x0, y0, x1, y1;
  bool draws;
};
__device__ __forceinline__ void raster_tri_setup(const float4& r1, const float4& r2, const float4& r3, uint32_t limit,
                                                 int tri_index, RasterTri* t) {
  uint32_t lo, hi;
  fr_code_keep(__float_as_uint(r1.w), __float_as_uint(r2.w), __float_as_uint(r3.w), limit, &lo, &hi);   // the box again, from the codes
  t->x0 = (int)(lo & 0xFFFFu) - 1;
  t->y0 = (int)(lo >> 16) - 1;
  t->x1 = (int)(hi & 0xFFFFu) - 1;
  t->y1 = (int)(hi >> 16) - 1;
  const float h = fr_tri_depth(r1.z, r2.z, r3.z);
  t->draws = fr_depth_draws(h);
  t->key = fr_pack_key(h, tri_index);
  fr_tri_edge_setup(r1.x, r1.y, r2.x, r2.y, r3.x, r3.y, &t->e);
}
// (A float pre-filter in front of the FP64 test -- raster_core.h fr_filter_pixel, exact whenever it answers and it answers
// 99.9 % of the time -- was measured twice on B200 and does not pay: phase B is bound by the dependent
// queue -> index -> record gather chain, not by FP64 issue.)
__device__ __forceinline__ bool raster_inside(const RasterTri& t, int px, int py) { return fr_point_in_tri_flat(&t.e, px, py); }

// Phase A: one thread per (triangle, FPT faces): three 4-byte gathers of snap codes and a handful of packed integer ops
// decide the reference's bounding-box cull exactly.  ~55 % of the sub-pixel BFM triangles contain no pixel centre and
// stop here.  The survivors are compacted into shared-memory queues by the number of candidate pixels in their box --
// one (about 60 %), two, more -- so that Phase B runs each class with every lane busy and in straight-line code:
//   B1  one pixel:  two survivors per thread and trip (six 16-byte record gathers in flight)
//   B2  two pixels: one survivor per thread, both inside tests interleaved
//   B3  the rest:   one survivor per thread, a flat loop over the box two pixels at a time
// Queue entries are 8 bits (the local triangle index; the face is the queue's slot): the box is recomputed from the
// records' code words.
// All record indices are 32-bit: the API guarantees batch * 3 * nver < 2^31.
#ifndef FR_KEYS_DBG
#define FR_KEYS_DBG 0    // timing experiments only (tools/ab_bench.sh): 1..5 = return after successive stages
#endif
#ifndef FR_KEYS_MINB
#define FR_KEYS_MINB 4   // resident CTAs per SM the register allocation is held to (A/B-tested on B200)
#endif
template <int FPT>
__global__ void __launch_bounds__(kRasterThreads, FR_KEYS_MINB)
raster_keys_kernel(const float4* __restrict__ rec, const float* __restrict__ tri, unsigned long long* __restrict__ keys,
                   int batch, int nver, int ntri, int height, int width) {
  // Queues are FACE-MAJOR: slot f owns kRasterThreads entries of q_a (one-pixel survivors from the front, two-pixel
  // survivors from the back) and of q_g.  Phase B then walks the concatenation of the per-face runs, so the 32 lanes of
  // a warp hold neighbouring triangles of the SAME face and their record gathers fall into a handful of cache lines
  // (with a triangle-major queue every lane of a gather hits a different line, and phase B is bound by L1 wavefronts).
  __shared__ unsigned char q_a[FPT][kRasterThreads];      // local triangle index
  __shared__ unsigned char q_g[FPT][kRasterThreads];
  __shared__ int s_idx[3][kRasterThreads];
  __shared__ unsigned q_count_a[FPT];                     // one-pixel | two-pixel << 16
  __shared__ unsigned q_count_g[FPT];
  static_assert(kRasterThreads == 256, "queue entries are 8-bit local triangle indices");

  const int tid = threadIdx.x;
  const unsigned lane = tid & 31u;
  if (tid < FPT) {
    q_count_a[tid] = 0u;
    q_count_g[tid] = 0u;
  }
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

#if FR_KEYS_DBG == 1
  return;
#endif
  const uint32_t limit = (uint32_t)width | ((uint32_t)height << 16);
  uint32_t e1[FPT], e2[FPT], e3[FPT];
#pragma unroll
  for (int f = 0; f < FPT; ++f) {  // all gathers in flight before the first use
    const unsigned fb = (unsigned)min(b0 + f, batch - 1) * (unsigned)nver;
    e1[f] = __float_as_uint(__ldg(&rec[fb + (unsigned)p1].w));
    e2[f] = __float_as_uint(__ldg(&rec[fb + (unsigned)p2].w));
    e3[f] = __float_as_uint(__ldg(&rec[fb + (unsigned)p3].w));
  }
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int f = 0; f < FPT; ++f) {
    uint32_t lo, hi;
    const bool keep = fr_code_keep(e1[f], e2[f], e3[f], limit, &lo, &hi) && valid && (b0 + f < batch);
    const uint32_t d = hi - lo;           // per-field extent - 1 (no borrow when the box is kept)
    const bool one = keep && d == 0u, two = keep && (d == 1u || d == 0x10000u), more = keep && !one && !two;
    // warp-aggregated append: one shared atomic per warp, face and queue
    const unsigned m1 = __ballot_sync(0xFFFFFFFFu, one), m2 = __ballot_sync(0xFFFFFFFFu, two);
    if ((m1 | m2) != 0u) {
      unsigned base = 0u;
      if (lane == 0) base = atomicAdd(&q_count_a[f], (unsigned)__popc(m1) | ((unsigned)__popc(m2) << 16));
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      if (one) q_a[f][(base & 0xFFFFu) + __popc(m1 & lt)] = (unsigned char)tid;
      if (two) q_a[f][kRasterThreads - 1 - ((base >> 16) + __popc(m2 & lt))] = (unsigned char)tid;
    }
    const unsigned m3 = __ballot_sync(0xFFFFFFFFu, more);
    if (m3 != 0u) {
      unsigned base = 0u;
      if (lane == 0) base = atomicAdd(&q_count_g[f], (unsigned)__popc(m3));
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      if (more) q_g[f][base + __popc(m3 & lt)] = (unsigned char)tid;
    }
  }
  __syncthreads();
#if FR_KEYS_DBG == 3
  return;
#endif

  // per-face run lengths -> exclusive prefix sums (FPT is tiny: every thread keeps them in registers)
  int pre1[FPT + 1], pre2[FPT + 1], pre3[FPT + 1];
  pre1[0] = pre2[0] = pre3[0] = 0;
#pragma unroll
  for (int f = 0; f < FPT; ++f) {
    const unsigned ca = q_count_a[f];
    pre1[f + 1] = pre1[f] + (int)(ca & 0xFFFFu);
    pre2[f + 1] = pre2[f] + (int)(ca >> 16);
    pre3[f + 1] = pre3[f] + (int)q_count_g[f];
  }
  // item index -> (face slot, position in the face's run)
  auto locate = [&](const int (&pre)[FPT + 1], int i, int* f, int* k) {
    int ff = 0;
#pragma unroll
    for (int g = 1; g < FPT; ++g) ff += (i >= pre[g]) ? 1 : 0;
    *f = ff;
    int base = 0;
#pragma unroll
    for (int g = 1; g < FPT; ++g) base = (i >= pre[g]) ? pre[g] : base;
    *k = i - base;
  };
  const int n_one = pre1[FPT], n_two = pre2[FPT], n_more = pre3[FPT];
  const int npix = height * width;
  const int tri0 = blockIdx.x * kRasterThreads;
  // ---- phase B1: one candidate pixel
  for (int i0 = tid; i0 < n_one; i0 += 2 * kRasterThreads) {
    const int i1 = i0 + kRasterThreads;
    const bool two = i1 < n_one;
    int fa, ka, fbq, kb;
    locate(pre1, i0, &fa, &ka);
    locate(pre1, two ? i1 : i0, &fbq, &kb);
    const int tla = q_a[fa][ka], tlb = q_a[fbq][kb];
    const int ba = b0 + fa, bb = b0 + fbq;
    const unsigned ra = (unsigned)ba * (unsigned)nver, rb = (unsigned)bb * (unsigned)nver;
    const float4 a1 = __ldg(rec + (ra + (unsigned)s_idx[0][tla])), a2 = __ldg(rec + (ra + (unsigned)s_idx[1][tla])),
                 a3 = __ldg(rec + (ra + (unsigned)s_idx[2][tla]));
    const float4 c1 = __ldg(rec + (rb + (unsigned)s_idx[0][tlb])), c2 = __ldg(rec + (rb + (unsigned)s_idx[1][tlb])),
                 c3 = __ldg(rec + (rb + (unsigned)s_idx[2][tlb]));
    RasterTri ta, tb;
    raster_tri_setup(a1, a2, a3, limit, tri0 + tla, &ta);
    raster_tri_setup(c1, c2, c3, limit, tri0 + tlb, &tb);
    const bool ina = raster_inside(ta, ta.x0, ta.y0) & ta.draws;
    const bool inb = raster_inside(tb, tb.x0, tb.y0) & tb.draws & two;
    if (ina) atomicMax(keys + (size_t)ba * npix + (ta.y0 * width + ta.x0), ta.key);
    if (inb) atomicMax(keys + (size_t)bb * npix + (tb.y0 * width + tb.x0), tb.key);
  }
#if FR_KEYS_DBG == 4
  return;
#endif
  // ---- phase B2: two candidate pixels: (x0, y0) and (x1, y1)
  for (int j = tid; j < n_two; j += kRasterThreads) {
    int f, k;
    locate(pre2, j, &f, &k);
    const int tl = q_a[f][kRasterThreads - 1 - k];
    const int b = b0 + f;
    const unsigned fb = (unsigned)b * (unsigned)nver;
    const float4 r1 = __ldg(rec + (fb + (unsigned)s_idx[0][tl])), r2 = __ldg(rec + (fb + (unsigned)s_idx[1][tl])),
                 r3 = __ldg(rec + (fb + (unsigned)s_idx[2][tl]));
    RasterTri tt;
    raster_tri_setup(r1, r2, r3, limit, tri0 + tl, &tt);
    const bool ina = raster_inside(tt, tt.x0, tt.y0) & tt.draws;
    const bool inb = raster_inside(tt, tt.x1, tt.y1) & tt.draws;
    unsigned long long* kb = keys + (size_t)b * npix;
    if (ina) atomicMax(kb + (tt.y0 * width + tt.x0), tt.key);
    if (inb) atomicMax(kb + (tt.y1 * width + tt.x1), tt.key);
  }
#if FR_KEYS_DBG == 5
  return;
#endif
  // ---- phase B3: larger boxes, row-major over the box, two pixels per trip
  for (int j = tid; j < n_more; j += kRasterThreads) {
    int f, k;
    locate(pre3, j, &f, &k);
    const int tl = q_g[f][k];
    const int b = b0 + f;
    const unsigned fb = (unsigned)b * (unsigned)nver;
    const float4 r1 = __ldg(rec + (fb + (unsigned)s_idx[0][tl])), r2 = __ldg(rec + (fb + (unsigned)s_idx[1][tl])),
                 r3 = __ldg(rec + (fb + (unsigned)s_idx[2][tl]));
    RasterTri tt;
    raster_tri_setup(r1, r2, r3, limit, tri0 + tl, &tt);
    if (!tt.draws) continue;
    unsigned long long* kb = keys + (size_t)b * npix;
    int x = tt.x0, y = tt.y0;
    while (y <= tt.y1) {
      const int xa = x, ya = y;
      if (++x > tt.x1) { x = tt.x0; ++y; }
      const int xb = x, yb = y;
      const bool second = yb <= tt.y1;
      if (++x > tt.x1) { x = tt.x0; ++y; }
      const bool ina = raster_inside(tt, xa, ya);
      const bool inb = raster_inside(tt, xb, yb) & second;
      if (ina) atomicMax(kb + (ya * width + xa), tt.key);
      if (inb) atomicMax(kb + (yb * width + xb), tt.key);
    }
  }
}

// ---------------------------------------------------------------------------------------------- split pipeline
// The same work as raster_keys_kernel in two kernels, for meshes with fewer than 2^20 triangles and at most 4096 faces per
// launch: raster_cull_kernel runs phase A alone (few registers, full occupancy) and appends its survivors -- one 32-bit
// word (face << 20 | triangle) each -- to three global lists by class (one / two / more candidate pixels);
// raster_draw_kernel then walks the dense lists with every lane busy and two survivors in flight per thread, whatever the
// per-block survivor counts were.  Lists: `la` holds one-pixel survivors from the front and two-pixel survivors from
// the back, `lg` the rest; capacity batch * ntri words each (the worst case); counters[0..2] = list lengths.
constexpr int kPairFaceShift = 20;
constexpr int kDrawThreads = 256;

template <int FPT>
__global__ void __launch_bounds__(kRasterThreads)
raster_cull_kernel(const float4* __restrict__ rec, const float* __restrict__ tri, uint32_t* __restrict__ la,
                   uint32_t* __restrict__ lg, unsigned* __restrict__ counters, unsigned list_cap, int batch, int nver,
                   int ntri, int height, int width) {
  constexpr int kQ = kRasterThreads * FPT;
  __shared__ unsigned short q_a[kQ];
  __shared__ unsigned short q_g[kQ];
  __shared__ unsigned q_count_a, q_count_g;
  __shared__ unsigned g_base[3];
  static_assert(FPT <= 8, "face slot is packed into 3 bits");

  const int tid = threadIdx.x;
  if (tid == 0) {
    q_count_a = 0u;
    q_count_g = 0u;
  }
  __syncthreads();

  const int t = blockIdx.x * kRasterThreads + tid;
  const int b0 = blockIdx.y * FPT;
  int p1 = 0, p2 = 0, p3 = 0;
  bool valid = t < ntri;
  if (valid)
    valid = tri_vertex_index(__ldg(tri + t), nver, &p1) && tri_vertex_index(__ldg(tri + ntri + t), nver, &p2) &&
            tri_vertex_index(__ldg(tri + 2 * (size_t)ntri + t), nver, &p3);
  const uint32_t limit = (uint32_t)width | ((uint32_t)height << 16);
  uint32_t e1[FPT], e2[FPT], e3[FPT];
#pragma unroll
  for (int f = 0; f < FPT; ++f) {
    const unsigned fb = (unsigned)min(b0 + f, batch - 1) * (unsigned)nver;
    e1[f] = __float_as_uint(__ldg(&rec[fb + (unsigned)p1].w));
    e2[f] = __float_as_uint(__ldg(&rec[fb + (unsigned)p2].w));
    e3[f] = __float_as_uint(__ldg(&rec[fb + (unsigned)p3].w));
  }
  unsigned keepmask = 0u, onemask = 0u, twomask = 0u;
#pragma unroll
  for (int f = 0; f < FPT; ++f) {
    uint32_t lo, hi;
    const bool keep = fr_code_keep(e1[f], e2[f], e3[f], limit, &lo, &hi) && valid && (b0 + f < batch);
    const uint32_t d = hi - lo;
    keepmask |= (keep ? 1u : 0u) << f;
    onemask |= ((d == 0u) ? 1u : 0u) << f;
    twomask |= ((d == 1u || d == 0x10000u) ? 1u : 0u) << f;
  }
  onemask &= keepmask;
  twomask &= keepmask;
  const unsigned moremask = keepmask & ~(onemask | twomask);
  if ((onemask | twomask) != 0u) {
    const unsigned got = atomicAdd(&q_count_a, (unsigned)__popc(onemask) | ((unsigned)__popc(twomask) << 16));
    int pf = (int)(got & 0xFFFFu), pb = kQ - 1 - (int)(got >> 16);
#pragma unroll
    for (int f = 0; f < FPT; ++f)
      if (((onemask | twomask) >> f) & 1u) q_a[((onemask >> f) & 1u) ? pf++ : pb--] = (unsigned short)((tid << 3) | f);
  }
  if (moremask != 0u) {
    int pg = (int)atomicAdd(&q_count_g, (unsigned)__popc(moremask));
#pragma unroll
    for (int f = 0; f < FPT; ++f)
      if ((moremask >> f) & 1u) q_g[pg++] = (unsigned short)((tid << 3) | f);
  }
  __syncthreads();
  const unsigned counts = q_count_a;
  const int n_one = (int)(counts & 0xFFFFu), n_two = (int)(counts >> 16), n_more = (int)q_count_g;
  if (tid < 3) {
    const int n = tid == 0 ? n_one : (tid == 1 ? n_two : n_more);
    g_base[tid] = n > 0 ? atomicAdd(counters + tid, (unsigned)n) : 0u;
  }
  __syncthreads();
  const uint32_t tri0 = (uint32_t)blockIdx.x * kRasterThreads;
  auto word = [&](unsigned id) { return ((uint32_t)(b0 + (int)(id & 7u)) << kPairFaceShift) | (tri0 + (id >> 3)); };
  for (int i = tid; i < n_one; i += kRasterThreads) la[g_base[0] + i] = word(q_a[i]);
  for (int i = tid; i < n_two; i += kRasterThreads) la[list_cap - 1u - (g_base[1] + i)] = word(q_a[kQ - 1 - i]);
  for (int i = tid; i < n_more; i += kRasterThreads) lg[g_base[2] + i] = word(q_g[i]);
}

// One survivor: list word -> triangle's vertex indices -> records -> per-triangle setup.
__device__ __forceinline__ void raster_fetch(uint32_t w, const float4* __restrict__ rec, const float* __restrict__ tri, int nver,
                                             int ntri, float4* r1, float4* r2, float4* r3, int* b, int* t) {
  *b = (int)(w >> kPairFaceShift);
  *t = (int)(w & ((1u << kPairFaceShift) - 1u));
  const int p1 = (int)__ldg(tri + *t), p2 = (int)__ldg(tri + ntri + *t), p3 = (int)__ldg(tri + 2 * (size_t)ntri + *t);   // validated by the cull
  const unsigned fb = (unsigned)*b * (unsigned)nver;
  *r1 = __ldg(rec + (fb + (unsigned)p1));
  *r2 = __ldg(rec + (fb + (unsigned)p2));
  *r3 = __ldg(rec + (fb + (unsigned)p3));
}

__global__ void __launch_bounds__(kDrawThreads, 4)
raster_draw_kernel(const float4* __restrict__ rec, const float* __restrict__ tri, const uint32_t* __restrict__ la,
                   const uint32_t* __restrict__ lg, const unsigned* __restrict__ counters, unsigned list_cap,
                   unsigned long long* __restrict__ keys, int nver, int ntri, int height, int width) {
  const unsigned n_one = counters[0], n_two = counters[1], n_more = counters[2];
  const unsigned gtid = blockIdx.x * kDrawThreads + threadIdx.x, gsize = gridDim.x * kDrawThreads;
  const uint32_t limit = (uint32_t)width | ((uint32_t)height << 16);
  const int npix = height * width;
  // ---- one candidate pixel: two survivors per thread and trip
  for (unsigned i = gtid; i < n_one; i += 2u * gsize) {
    const unsigned i2 = i + gsize;
    const bool two = i2 < n_one;
    const uint32_t wa = __ldg(la + i), wb = __ldg(la + (two ? i2 : i));
    float4 a1, a2, a3, c1, c2, c3;
    int ba, ta_, bb, tb_;
    raster_fetch(wa, rec, tri, nver, ntri, &a1, &a2, &a3, &ba, &ta_);
    raster_fetch(wb, rec, tri, nver, ntri, &c1, &c2, &c3, &bb, &tb_);
    RasterTri ta, tb;
    raster_tri_setup(a1, a2, a3, limit, ta_, &ta);
    raster_tri_setup(c1, c2, c3, limit, tb_, &tb);
    const bool ina = raster_inside(ta, ta.x0, ta.y0) & ta.draws;
    const bool inb = raster_inside(tb, tb.x0, tb.y0) & tb.draws & two;
    if (ina) atomicMax(keys + (size_t)ba * npix + (ta.y0 * width + ta.x0), ta.key);
    if (inb) atomicMax(keys + (size_t)bb * npix + (tb.y0 * width + tb.x0), tb.key);
  }
  // ---- two candidate pixels: (x0, y0) and (x1, y1)
  for (unsigned j = gtid; j < n_two; j += gsize) {
    const uint32_t w = __ldg(la + (list_cap - 1u - j));
    float4 r1, r2, r3;
    int b, t;
    raster_fetch(w, rec, tri, nver, ntri, &r1, &r2, &r3, &b, &t);
    RasterTri tt;
    raster_tri_setup(r1, r2, r3, limit, t, &tt);
    const bool ina = raster_inside(tt, tt.x0, tt.y0) & tt.draws;
    const bool inb = raster_inside(tt, tt.x1, tt.y1) & tt.draws;
    unsigned long long* kb = keys + (size_t)b * npix;
    if (ina) atomicMax(kb + (tt.y0 * width + tt.x0), tt.key);
    if (inb) atomicMax(kb + (tt.y1 * width + tt.x1), tt.key);
  }
  // ---- larger boxes, row-major over the box, two pixels per trip
  for (unsigned j = gtid; j < n_more; j += gsize) {
    const uint32_t w = __ldg(lg + j);
    float4 r1, r2, r3;
    int b, t;
    raster_fetch(w, rec, tri, nver, ntri, &r1, &r2, &r3, &b, &t);
    RasterTri tt;
    raster_tri_setup(r1, r2, r3, limit, t, &tt);
    if (!tt.draws) continue;
    unsigned long long* kb = keys + (size_t)b * npix;
    int x = tt.x0, y = tt.y0;
    while (y <= tt.y1) {
      const int xa = x, ya = y;
      if (++x > tt.x1) { x = tt.x0; ++y; }
      const int xb = x, yb = y;
      const bool second = yb <= tt.y1;
      if (++x > tt.x1) { x = tt.x0; ++y; }
      const bool ina = raster_inside(tt, xa, ya);
      const bool inb = raster_inside(tt, xb, yb) & second;
      if (ina) atomicMax(kb + (ya * width + xa), tt.key);
      if (inb) atomicMax(kb + (yb * width + xb), tt.key);
    }
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
