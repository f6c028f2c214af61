// Tensor-core reconstruction + projection forward, second generation: the basis is split ONCE, at pack time, into an
// fp16 hi/lo pair per element (22 significant bits after a per-column power-of-two scale), laid out as ready-made
// tcgen05 operand tiles.  The kernel then has no conversion stage at all:
//
//   V[b, (c,n)] = sum_k P[(c,n), k] * coef[b, k]
//              ~= 2^-t_b * sum_k (Phi + Plo)[(c,n), k] * (b0 + b1)[b, k]          (dropping Plo*b1, 2^-22 relative)
//   P 2^s_k = Phi + Plo (fp16 each),  coef 2^-s_k 2^t_b = b0 + b1 (fp16 each),  products accumulated in fp32 (TMEM)
//
// Per CTA (persistent, one per SM):
//   producer    1 warp   cp.async.bulk (TMA engine): the two resident coefficient operands of this 64-face batch tile
//                        once, then 24 KB stages (3 chunks of 16 k: hi tile + lo tile each) of the basis into a
//                        shared-memory ring -- a tile's 3 coordinate rows are one contiguous run of the packed file
//   MMA issuer  1 warp   whole warp converged, one elected lane: per chunk  D += Phi.b0 + Plo.b0 + Phi.b1
//                        (tcgen05.mma kind::f16, M128 N64 K16, BOTH operands from shared memory); a tcgen05.commit per
//                        stage hands the shared-memory stage back to the producer when its MMAs have retired
//   epilogue    8 / 24 warps: tcgen05.ld of the three 128x64 accumulators (x, y, z of the same vertices),
//                        (2^-t f.R).v + t, y flip; double-buffered against the next tile's MMAs.  Two flavours:
//       planar  (8 warps)   coalesced stores of vertex_proj [B,3,N]                  (fr_recon_project_forward)
//       raster  (24 warps)  the row tile is a CLUSTER of the mesh table (mesh_table.h): three independent groups of 8 warps
//                           take the projected vertices, 8 faces (an octet) at a time, into their shared-memory tiles of the
//                           tile rasterizer (raster_tile.cuh) and cull and draw the cluster's triangles from there while
//                           the tensor pipe works on the next cluster -- the vertices of the fused params -> depth-map
//                           call never touch global memory.  At 64 faces the rasterization, not the basis stream, bounds
//                           the kernel (12 us per cluster and SM against 7.5 us for its tile), so the work is dealt
//                           dynamically: octets within a CTA (Barriers::oct_taken), the clusters of the last round between
//                           the CTAs (ItemWalk)                           (fr_recon_render_forward with FR_CLUSTER_TILES)
// Measured on B200 (tools/mma_bench3.cu): an SS-form M128 N64 MMA takes 48 cycles (operand reads at 128 B/clk), K8 tf32
// and K16 f16 alike, so a 16-k chunk costs 144 tensor cycles here against 192 for the 3xTF32 TS-form kernel
// (tools/experiments/recon_tc_3xtf32.cuh) -- and that kernel also needs 8 converter warps and a TMEM operand ring only 4 chunks deep.
#ifndef FR_RECON_F16_CUH_
#define FR_RECON_F16_CUH_

#include <cuda_fp16.h>

#include "raster_tile.cuh"
#include "recon.cuh"
#include "schedule.h"
#include "tcgen05_common.cuh"

#ifndef FR_BASIS_EVICT_FIRST
#define FR_BASIS_EVICT_FIRST 1   // A/B on B200: 102.8 -> 101.6 us per step (the records / keys stay in L2 for the rasterizer)
#endif

namespace fr {
namespace f16 {

#ifdef FR_TIMELINE   // developer build (tools/timeline.py): per-CTA clock64 stamps of the forward kernel, per-cluster raster time
__device__ long long g_timeline[160][16];
__device__ float g_cluster_cost[1024];       // cycles per octet group 0 spent on the cluster (accumulators complete -> last draw)
#define FR_TL(slot) do { if (blockIdx.y == 0) g_timeline[blockIdx.x][slot] = clock64(); } while (0)
#define FR_TLG(slot) do { if (blockIdx.y == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_timeline[blockIdx.x][slot] = (long long)t_; } } while (0)
#else
#define FR_TL(slot) do {} while (0)
#define FR_TLG(slot) do {} while (0)
#endif

using tc::bulk_load;
using tc::elect_one;
using tc::mbar_arrive;
using tc::mbar_arrive_expect_tx;
using tc::mbar_init;
using tc::mbar_wait;
using tc::smem_u32;
using tc::tc_commit;
using tc::tc_fence_after;
using tc::tc_fence_before;
using tc::tmem_ld16;

constexpr int kN = 64;                 // faces per batch tile (MMA N)
constexpr int kChunkK = 16;            // k columns per chunk == one f16 MMA K step
constexpr uint32_t kHalfBytes = kTileVerts * kChunkK * 2;   // one operand tile (hi or lo): 128 rows x 16 k fp16 = 4 KB
constexpr uint32_t kChunkBytes = 2 * kHalfBytes;            // hi tile + lo tile
constexpr int kStageChunks = 3;        // chunks per bulk copy: a tile has 3 * nch16 chunks, always a multiple of 3
constexpr uint32_t kStageBytes = kStageChunks * kChunkBytes;
constexpr int kMaxStages = 4;          // ring depth of the planar flavour: 96 KB of basis in flight per SM
constexpr int kDBufs = 2;
constexpr int kDCols = 3 * kN;
constexpr int kTmemCols = 512;
constexpr int kPose16Stride = 16;      // floats per face: 2^-t f.R [9] | t [3] | pad

// Flavour of the forward kernel (see the header comment).
template <bool kRaster>
struct Cfg {
  // Epilogue warps; every one of them reads accumulators (its TMEM lane quarter = warp % 4).
  //   planar: 8 warps, 16 consecutive faces per warp and 32-face step.
  //   raster: 24 warps in kGroups = 3 independent GROUPS of 8 (two warps per lane quarter).  A group stages 8 faces
  //           (an "octet") of the cluster in its own shared-memory tile, culls (thread = triangle) and draws them, and
  //           moves on to its next octet; the groups only meet at the accumulator hand-over, so while one group waits
  //           at its own barrier or runs its short cull phase the others keep the issue slots busy.
  static constexpr int kGroups = kRaster ? 3 : 1;
#ifndef FR_FUSED_GROUP_WARPS
#define FR_FUSED_GROUP_WARPS 8      // 10: two more warps per group that only draw (A/B)
#endif
  static constexpr int kGroupWarps = kRaster ? FR_FUSED_GROUP_WARPS : 8;
  static constexpr int kEpiWarps = kGroups * kGroupWarps;
  static constexpr int kProducerWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1;
  static constexpr int kThreads = (kEpiWarps + 2) * 32;
  static constexpr int kStages = kRaster ? 3 : kMaxStages;       // the raster flavour's three tiles need 86 KB: 224 KB in all
  static constexpr int kStepFaces = kRaster ? 8 : 32;            // faces per epilogue step (raster: one staged octet)
  static constexpr int kFacesPerWarp = kRaster ? 4 : 16;         // consecutive faces a warp reads per step
};

// instruction descriptor (cute::UMMA::InstrDescriptor): c = F32 at [4,6), a = b = F16 (0) at [7,10) / [10,13), K-major
// A and B, n_dim = N >> 3 at [17,23), m_dim = M >> 4 at [24,29)
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kN >> 3) << 17) | ((128u >> 4) << 24);

struct Barriers {
  uint64_t raw_full[kMaxStages];
  uint64_t raw_empty[kMaxStages];
  uint64_t d_full[kDBufs];
  uint64_t d_empty[kDBufs];
  uint64_t b_full;
  uint32_t tmem_base;
  uint32_t pad;
  uint32_t oct_taken[4];   // raster flavour: octets of tile t handed out beyond the three static ones, slot t % 4
  uint64_t item_full[8];   // dynamic pool (ItemWalk): entry d % 8 of item_ring is valid
  int32_t item_ring[8];    // ... the d-th item this CTA took from the pool, -1 = the pool is empty
  uint32_t raster_pos;     // items whose rasterization has begun (the fastest group's count): throttles the draws from the pool
  uint32_t pad2;
};

using RasterSmem = rt::TileSmem<Cfg<true>::kStepFaces>;   // raster flavour only

struct SmemLayout {
  uint32_t b0, b1, raw, pose, bars, raster, total;
  uint32_t sbo;  // bytes between 8-face groups of a coefficient operand
};
template <bool kRaster>
__host__ __device__ inline SmemLayout smem_layout(int nch16) {
  SmemLayout L;
  L.sbo = (uint32_t)nch16 * 2u * 128u;            // 2 core matrices (8 faces x 8 k fp16 = 128 B) per chunk
  const uint32_t bsz = (kN / 8) * L.sbo;
  L.b0 = 0;
  L.b1 = bsz;
  L.raw = (2 * bsz + 1023u) / 1024u * 1024u;
  L.pose = L.raw + Cfg<kRaster>::kStages * kStageBytes;
  L.bars = L.pose + kN * kPose16Stride * 4;
  L.raster = (L.bars + (uint32_t)sizeof(Barriers) + 15u) / 16u * 16u;
  L.total = L.raster + (kRaster ? (uint32_t)(Cfg<true>::kGroups * sizeof(RasterSmem)) : 0u);
  return L;
}

// D[tmem_d] (+)= A[smem desc] . B[smem desc]   (M128 N64 K16, fp16 inputs, fp32 accumulate)
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(kIdesc), "r"((uint32_t)accumulate)
      : "memory");
}

// ---------------------------------------------------------------------------------------------- packing
// Column scale: s_k with max_n |P[n,k]| 2^s_k in [2^14, 2^15) (fp16 overflows at 2^16); the scale region first collects
// the column maxima as float bit patterns (non-negative floats order like unsigned integers), then holds 2^-s_k.
__global__ void __launch_bounds__(256)
basis_colmax_kernel(const float* __restrict__ mu, const float* __restrict__ pc_shape, const float* __restrict__ pc_exp,
                    int nver, int ks, int ke, unsigned* __restrict__ colmax_bits) {
  const int kreal = ks + ke + 1;
  const size_t rows = (size_t)3 * nver;
  const size_t r0 = (size_t)blockIdx.x * 512, r1 = min(rows, r0 + 512);
  for (int k = threadIdx.x; k < kreal; k += 256) {
    float m = 0.0f;
    for (size_t r = r0; r < r1; ++r) {
      const float v = (k < ks) ? pc_shape[r * ks + k] : (k < ks + ke ? pc_exp[r * ke + (k - ks)] : mu[r]);
      m = fmaxf(m, fabsf(v));          // NaN entries are ignored here and stay NaN in the packed tiles
    }
    atomicMax(colmax_bits + k, __float_as_uint(m));
  }
}

__global__ void basis_colscale_kernel(float* __restrict__ scale, int kreal, int kpad16) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= kpad16) return;
  float inv = 1.0f;
  if (k < kreal) {
    const float m = __uint_as_float(reinterpret_cast<unsigned*>(scale)[k]);
    if (m > 0.0f && m < 3.0e38f) {
      int e;
      frexpf(m, &e);                   // m = f 2^e, f in [0.5, 1)  =>  m 2^(15-e) in [2^14, 2^15)
      const int s = max(-100, min(100, 15 - e));
      inv = ldexpf(1.0f, -s);
    }
  }
  scale[k] = inv;
}

// fp16 operand tiles: for every (tile = cluster, coordinate) row, nch16 chunks of [hi tile | lo tile]; a tile is the canonical
// K-major no-swizzle layout of a 128 x 16 fp16 operand: [k / 8][row][k % 8], i.e. 8-row x 16-byte core matrices, 128 B
// between 8-row groups (SBO) and 2048 B between the two K halves (LBO).  One thread writes one 16-byte row piece.
__global__ void __launch_bounds__(256)
pack_basis_f16_kernel(const float* __restrict__ mu, const float* __restrict__ pc_shape, const float* __restrict__ pc_exp,
                      const float* __restrict__ inv_scale, int nver, int ks, int ke, int nch16, int ntiles, unsigned flags,
                      const int32_t* __restrict__ cluster_vert, uint4* __restrict__ tiles) {
  const size_t total = (size_t)ntiles * 3 * nch16 * 2 * kTileVerts;      // (tile, c, chunk, k half, row)
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int v = (int)(idx % kTileVerts);
  const int kh = (int)((idx / kTileVerts) % 2);
  const int ci = (int)((idx / (2 * kTileVerts)) % nch16);
  const int c = (int)((idx / ((size_t)2 * kTileVerts * nch16)) % 3);
  const int tile = (int)(idx / ((size_t)2 * kTileVerts * nch16 * 3));
  // row v of tile `tile`: the cluster's vertex in slot v (mesh_table.h), or the tile's v-th consecutive vertex
  int n = tile * kTileVerts + v;
  if (cluster_vert != nullptr) {
    const int32_t raw = cluster_vert[(size_t)tile * kTileVerts + v];
    n = raw < 0 ? nver : (int)((uint32_t)raw & kVertIdMask);
  }
  __half hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = ci * kChunkK + kh * 8 + j;
    float x = 0.0f;
    if (n < nver) {
      const size_t row_b = (flags & FR_BASIS_INTERLEAVED) ? (size_t)3 * n + c : (size_t)c * nver + n;
      const size_t row_m = (flags & FR_MEAN_INTERLEAVED) ? (size_t)3 * n + c : (size_t)c * nver + n;
      if (k < ks) x = pc_shape[row_b * ks + k];
      else if (k < ks + ke) x = pc_exp[row_b * ke + (k - ks)];
      else if (k == ks + ke) x = mu[row_m];
    }
    const float xs = x * (1.0f / inv_scale[k]);            // power of two: exact
    hi[j] = __float2half_rn(xs);
    lo[j] = __float2half_rn(xs - __half2float(hi[j]));     // the difference is exact in fp32
  }
  uint4 whi, wlo;
  memcpy(&whi, hi, 16);
  memcpy(&wlo, lo, 16);
  const size_t chunk = ((size_t)(tile * 3 + c) * nch16 + ci) * (kChunkBytes / 16);
  const size_t piece = (size_t)kh * kTileVerts + v;
  tiles[chunk + piece] = whi;
  tiles[chunk + kHalfBytes / 16 + piece] = wlo;
}

// ---------------------------------------------------------------------------------------------- prep
// One CTA per (padded) face (+ optional key-clearing CTAs behind them): coefficients with the column scale undone, a per-face power-of-two scale 2^t that brings the
// largest of them into [2^13, 2^14), split into fp16 b0 + b1 and stored per 64-face batch tile in the canonical K-major
// layout [b0|b1][8-face group][k / 8][face % 8][k % 8] the kernel bulk-copies; pose16 = 2^-t f.R | t3d.
__global__ void __launch_bounds__(256)
recon_prep_f16_kernel(const float* __restrict__ params, const float* __restrict__ inv_scale, int dparam, int batch, int ks,
                      int ke, int kpad16, unsigned flags, float im_size, unsigned char* __restrict__ bsplit,
                      float* __restrict__ pose16, int bpad, unsigned long long* __restrict__ keys, size_t key_vecs,
                      unsigned* __restrict__ counters) {
  __shared__ float red[8];
  __shared__ float s_pose[kPoseStride];
  __shared__ double s_sc[6];
  pdl_trigger();                                   // the reconstruction kernel may become resident (it waits before reading our output)
  // This grid is itself launched as a programmatic dependent of whatever precedes it in the stream -- in a serving loop the
  // previous call's resolve kernel, which triggers at its start: then these blocks and, behind them, the reconstruction
  // kernel's CTAs (TMEM allocation, first basis stages: nothing that depends on the stream's past) are already resident
  // when that kernel ends.  Everything below reads or overwrites what earlier work in the stream produced or still uses.
  pdl_wait();
  FR_MARK_MIN(0);
  const int tid = threadIdx.x;
  // fused call: the CTAs behind the (padded) faces only clear the visibility keys for the rasterizer that follows, 16 bytes
  // per store (the consumer kernels wait for this whole grid before their first atomicMax)
  if ((int)blockIdx.x >= bpad) {
    uint4* kv = reinterpret_cast<uint4*>(keys);
    const size_t stride = (size_t)(gridDim.x - bpad) * 256;
    for (size_t i = (size_t)((int)blockIdx.x - bpad) * 256 + tid; i < key_vecs; i += stride) kv[i] = make_uint4(0u, 0u, 0u, 0u);
    FR_MARK_MAX(1);
    return;
  }
  const int b = blockIdx.x;
  const bool live = b < batch;
  if (b == 0 && tid < 4 && counters != nullptr) counters[tid] = 0u;     // the forward kernel's pool counter (ItemWalk)
  float cmax = 0.0f;
  for (int k = tid; k < kpad16; k += 256) {
    float v = 0.0f;
    if (live) {
      if (k < ks + ke) v = read_param(params + (size_t)b * dparam, FR_NDIM_POSE + k, flags, ks, im_size);
      else if (k == ks + ke) v = 1.0f;
    }
    cmax = fmaxf(cmax, fabsf(v * inv_scale[k]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cmax = fmaxf(cmax, __shfl_xor_sync(0xFFFFFFFFu, cmax, o));
  if ((tid & 31) == 0) red[tid >> 5] = cmax;
  // the three float64 sincos are the longest dependency chain of this kernel: one angle per warp (lanes 0 of warps 1..3)
  if (live && (tid & 31) == 0 && tid >= 32 && tid < 128) {
    const int a = (tid >> 5) - 1;
    sincos((double)read_param(params + (size_t)b * dparam, a, flags, ks, im_size), &s_sc[2 * a], &s_sc[2 * a + 1]);
  }
  __syncthreads();
  if (tid == 0 && live) {
    float p7[FR_NDIM_POSE];
    read_pose_params(params + (size_t)b * dparam, flags, ks, im_size, p7);
    pose_matrices_sc(s_sc[0], s_sc[1], s_sc[2], s_sc[3], s_sc[4], s_sc[5], p7, flags, s_pose);
  }
  cmax = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) cmax = fmaxf(cmax, red[w]);
  int t = 0;
  if (cmax > 0.0f && cmax < 3.0e38f) {
    int e;
    frexpf(cmax, &e);                  // cmax 2^(14-e) in [2^13, 2^14)
    t = max(-60, min(60, 14 - e));
  }
  const float up = ldexpf(1.0f, t), down = ldexpf(1.0f, -t);
  const uint32_t sbo = (uint32_t)(kpad16 / 8) * 128u, half = (kN / 8) * sbo;
  const int n = b % kN;
  unsigned char* tile = bsplit + (size_t)(b / kN) * 2 * half;
  for (int k = tid; k < kpad16; k += 256) {
    float v = 0.0f;
    if (live) {
      if (k < ks + ke) v = read_param(params + (size_t)b * dparam, FR_NDIM_POSE + k, flags, ks, im_size);
      else if (k == ks + ke) v = 1.0f;
    }
    const float c = v * inv_scale[k] * up;
    const __half b0 = __float2half_rn(c);
    const __half b1 = __float2half_rn(c - __half2float(b0));
    const uint32_t off = (uint32_t)(n >> 3) * sbo + (uint32_t)(k >> 3) * 128u + (uint32_t)(n & 7) * 16u + (uint32_t)(k & 7) * 2u;
    *reinterpret_cast<__half*>(tile + off) = b0;
    *reinterpret_cast<__half*>(tile + half + off) = b1;
  }
  __syncthreads();                     // s_pose (written by thread 0 above)
  if (tid < kPose16Stride) {
    float v = 0.0f;
    if (live) {
      if (tid < 9) v = s_pose[tid] * down;
      else if (tid < 12) v = s_pose[tid];
    }
    pose16[(size_t)b * kPose16Stride + tid] = v;
  }
  // the operands written above are read by the next kernel's bulk-copy engine (async proxy) behind a programmatic dependency
  // wait instead of an ordinary kernel boundary: order this thread's generic-proxy stores against that proxy on the writer's
  // side as well (the reader fences after its wait)
  asm volatile("fence.proxy.async;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------- forward
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[N]);
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, float (&v)[16]) { tmem_ld16(taddr, v); }
template <>
__device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
template <>
__device__ __forceinline__ void tmem_ld<4>(uint32_t taddr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

// The CTA's sequence of work items (cluster, epilogue steps): the static schedule (TileWalk), then -- single-tile launches of
// the raster flavour -- items from a POOL shared by all CTAs of the launch.  Per-cluster raster times differ by +-16 % and
// every CTA only sees three or four clusters, so with a purely static deal the slowest CTA ran ~10 us behind the median of
// a 70 us kernel (tools/timeline.py).  The clusters of the last round are therefore cut into FR_POOL_PARTS step ranges
// and handed out through a global counter (workspace, zeroed by the prep kernel): a CTA that is early takes more of them.
// The producer warp draws (atomicAdd) and publishes the item in a shared-memory ring guarded by mbarriers; the MMA warp and
// the epilogue warps read it there.  -1 ends the kernel.  The draw for the CTA's item number T waits until the rasterizing
// warps have begun item T - FR_POOL_LOOKAHEAD (Barriers::raster_pos): left to the pipeline's own back-pressure (three stages,
// two accumulator sets) the producers run two to three items ahead and empty the pool in a round-robin long before the
// CTAs' raster times have diverged (measured: no gain); one item of lookahead is what the tile's stream and MMAs need.
#ifndef FR_POOL_LOOKAHEAD
#define FR_POOL_LOOKAHEAD 1
#endif

struct ItemWalk {
  TileWalk tw;
  Barriers* bars;
  unsigned* counter;               // null: static schedule only
  uint32_t d;                      // pool items seen so far
  uint32_t nstatic;                // items of the static schedule seen so far
  bool static_done;
  int tile, step0, step1;
  __device__ ItemWalk(int nclusters, int nsteps, bool share_first, unsigned* counter_, Barriers* bars_)
      : tw(nclusters, nsteps, share_first, counter_ != nullptr, (int)gridDim.x, (int)blockIdx.x), bars(bars_), counter(tw.npool > 0 ? counter_ : nullptr), d(0),
        nstatic(0), static_done(false), tile(0), step0(0), step1(0) {}
  __device__ bool take_static() {
    if (!static_done && tw.next()) {
      tile = tw.tile, step0 = tw.step0, step1 = tw.step1;
      ++nstatic;
      return true;
    }
    static_done = true;
    return false;
  }
  __device__ bool decode(int item) {
    if (item < 0) return false;
    pool_decode(tw, item, &tile, &step0, &step1);
    return true;
  }
  // producer (one thread): draws from the pool and publishes
  __device__ bool next_producer() {
    if (take_static()) return true;
    if (counter == nullptr) return false;
    if (d == 0) pdl_wait();                                      // the counter was zeroed by the prep kernel
    const uint32_t need = nstatic + d + 1u > (uint32_t)FR_POOL_LOOKAHEAD ? nstatic + d + 1u - (uint32_t)FR_POOL_LOOKAHEAD : 0u;
    // (bounded like mbar_wait: a protocol bug traps -- a launch error the API reports -- instead of hanging the GPU)
    uint32_t spin = 0;
    while (*reinterpret_cast<volatile uint32_t*>(&bars->raster_pos) < need) {
      __nanosleep(64);
      if (++spin > (1u << 25)) __trap();
    }
    const unsigned idx = atomicAdd(counter, 1u);
    const int item = idx < (unsigned)pool_items(tw) ? (int)idx : -1;
    bars->item_ring[d & 7u] = item;
    mbar_arrive(&bars->item_full[d & 7u]);                        // (release: the ring entry is visible to the waiters)
    ++d;
    return decode(item);
  }
  // MMA warp / epilogue warps (every thread)
  __device__ bool next_consumer() {
    if (take_static()) return true;
    if (counter == nullptr) return false;
    mbar_wait(&bars->item_full[d & 7u], (d >> 3) & 1u);
    const int item = *reinterpret_cast<volatile int32_t*>(&bars->item_ring[d & 7u]);
    ++d;
    return decode(item);
  }
};

__device__ __forceinline__ void tmem_ld1(uint32_t taddr, float* v) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  *v = __uint_as_float(r);
}

// What the raster flavour draws into (fused params -> depth-map call).
struct RasterTarget {
  const unsigned char* table;        // mesh table blob (device)
  unsigned long long* keys;          // visibility keys [B][height*width], cleared by the prep kernel
  int width, height;
  unsigned* pool_counter;            // ItemWalk's pool counter (16 bytes, zeroed by the prep kernel), or null
};

// tiles: fp16 operand tiles; cluster_vert: the vertex (id | owner flag, -1 = none) behind every row of every tile -- the mesh
// table's cluster lists (FR_CLUSTER_TILES) or its rank order (one tile = 128 consecutive ranks); null = consecutive vertex ids.
// kRecords (raster flavour only): the epilogue also writes the 16-byte vertex records (out.rec, by rank) for a resolve pass with
// normals / texture (fr_recon_render_forward_all).  A separate instantiation: the depth-only kernel runs at its register limit
// and the extra live values (record index, the branch) cost 4 us per 64 faces when they are merely optional at run time.
template <bool kRaster, bool kRecords = false>
__global__ void __launch_bounds__(Cfg<kRaster>::kThreads, 1)
recon_fwd_f16_kernel(const unsigned char* __restrict__ tiles, const unsigned char* __restrict__ bsplit,
                     const float* __restrict__ pose16, const int32_t* __restrict__ cluster_vert, ReconOut out,
                     RasterTarget target, int batch, int nver, int nch16, int nclusters, float im_size, unsigned flags, int bt0) {
  using C = Cfg<kRaster>;
  constexpr int kStages = C::kStages, kEpiWarps = C::kEpiWarps;
  constexpr int kProducerWarp = C::kProducerWarp, kMmaWarp = C::kMmaWarp;
  extern __shared__ __align__(1024) unsigned char smem[];
  const SmemLayout L = smem_layout<kRaster>(nch16);
  Barriers* bars = reinterpret_cast<Barriers*>(smem + L.bars);
  float* s_pose = reinterpret_cast<float*>(smem + L.pose);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int bt = bt0 + (int)blockIdx.y;                           // batch tile of this CTA (a launch covers tiles bt0 ...)
  const int b0 = bt * kN;
  const size_t tile_bytes = (size_t)3 * nch16 * kChunkBytes;
  constexpr int kStepFaces = C::kStepFaces;                        // faces the epilogue warps read per step (raster: one stage)
  const int nsteps = (min(kN, batch - b0) + kStepFaces - 1) / kStepFaces;
  unsigned* const pool_counter = (kRaster && gridDim.y == 1) ? target.pool_counter : nullptr;   // (ItemWalk)
  if (threadIdx.x == 0) { FR_TL(0); FR_TLG(15); }
  FR_MARK_MIN(2);

  // ---- one-time setup: barriers, TMEM, poses
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&bars->raw_full[i], 1);
      mbar_init(&bars->raw_empty[i], 1);         // released by a tcgen05.commit
    }
    for (int i = 0; i < kDBufs; ++i) {
      mbar_init(&bars->d_full[i], 1);
      mbar_init(&bars->d_empty[i], kEpiWarps);
    }
    mbar_init(&bars->b_full, 1);
    for (int i = 0; i < 4; ++i) bars->oct_taken[i] = 0u;
    for (int i = 0; i < 8; ++i) mbar_init(&bars->item_full[i], 1);
    bars->raster_pos = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kProducerWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_trigger();                                   // the next kernel's blocks may become resident (they wait before reading our output)
  // No pdl_wait() here: only the coefficient operands, the poses and the cleared keys come from the prep kernel.  The
  // basis stream and the TMEM allocation start at once and overlap the prep kernel; the producer waits before it loads
  // the coefficient operands, the epilogue warps before they read the poses / touch the keys.
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  if (threadIdx.x == 0) FR_TL(1);

  if (warp == kProducerWarp) {
    // ================================================================== producer (TMA engine)
    if (lane == 0) {
      bool b_loaded = false;
      auto load_b = [&]() {                                    // the prep kernel's output: wait for it (programmatic dependent launch)
        pdl_wait();
        // the operands were written by the prep kernel's ordinary (generic-proxy) stores and are read here by the bulk-copy
        // engine (async proxy): order the two proxies explicitly instead of relying on the dependency wait alone
        asm volatile("fence.proxy.async;" ::: "memory");
        const uint32_t bbytes = (kN / 8) * L.sbo;
        mbar_arrive_expect_tx(&bars->b_full, 2u * bbytes);
        bulk_load(smem + L.b0, bsplit + (size_t)bt * 2 * bbytes, bbytes, &bars->b_full);
        bulk_load(smem + L.b1, bsplit + (size_t)bt * 2 * bbytes + bbytes, bbytes, &bars->b_full);
        b_loaded = true;
      };
      uint32_t it = 0;
#if FR_BASIS_EVICT_FIRST
      const uint64_t stream_policy = tc::l2_evict_first_policy();
#endif
      for (ItemWalk tw(nclusters, nsteps, kRaster && gridDim.y == 1, pool_counter, bars); tw.next_producer();) {
        const unsigned char* src = tiles + (size_t)tw.tile * tile_bytes;
        for (int sg = 0; sg < nch16; ++sg, ++it) {             // nch16 stages of 3 chunks per tile
          const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
          if (it == (uint32_t)kStages) load_b();               // the ring is full of basis stages: the MMAs need the operands now
          mbar_wait(&bars->raw_empty[s], ph ^ 1u);
          mbar_arrive_expect_tx(&bars->raw_full[s], kStageBytes);
#if FR_BASIS_EVICT_FIRST
          tc::bulk_load_hint(smem + L.raw + s * kStageBytes, src + (size_t)sg * kStageBytes, kStageBytes, &bars->raw_full[s], stream_policy);
#else
          bulk_load(smem + L.raw + s * kStageBytes, src + (size_t)sg * kStageBytes, kStageBytes, &bars->raw_full[s]);
#endif
        }
      }
      if (!b_loaded) load_b();                                 // fewer stages than the ring holds
    }
  } else if (warp == kMmaWarp) {
    // ================================================================== MMA issuer (whole warp converged, one elected lane issues)
    mbar_wait(&bars->b_full, 0);
    tc_fence_after();
    if (lane == 0) FR_TL(5);
    const uint64_t db0 = tc::make_smem_desc(smem_u32(smem + L.b0), 128u, L.sbo);
    const uint64_t db1 = tc::make_smem_desc(smem_u32(smem + L.b1), 128u, L.sbo);
    const uint64_t da0 = tc::make_smem_desc(smem_u32(smem + L.raw), 2048u, 128u);   // same descriptor format for A
    uint32_t it = 0, tcount = 0;
    for (ItemWalk tw(nclusters, nsteps, kRaster && gridDim.y == 1, pool_counter, bars); tw.next_consumer(); ++tcount) {
      const uint32_t dbuf = tcount % kDBufs;
      mbar_wait(&bars->d_empty[dbuf], ((tcount / kDBufs) & 1u) ^ 1u);          // epilogue has drained this accumulator set
      tc_fence_after();
      uint32_t c = 0, ci = 0;                                                  // coordinate and chunk-in-row of the next chunk
#pragma unroll 1
      for (int sg = 0; sg < nch16; ++sg, ++it) {
        const uint32_t s = it % kStages;
        mbar_wait(&bars->raw_full[s], (it / kStages) & 1u);
        tc_fence_after();
        if (lane == 0 && it == 0) FR_TL(6);
        uint32_t cj[kStageChunks], cij[kStageChunks];                          // (coordinate, chunk-in-row) of the stage's chunks
#pragma unroll
        for (int j = 0; j < kStageChunks; ++j) {
          cj[j] = c;
          cij[j] = ci;
          if (++ci == (uint32_t)nch16) { ci = 0; ++c; }
        }
        if (elect_one()) {
          const uint64_t da = da0 + (uint64_t)((s * kStageBytes) >> 4);
#pragma unroll
          for (int j = 0; j < kStageChunks; ++j) {
            const uint32_t d_addr = tmem + dbuf * kDCols + cj[j] * kN;
            const uint64_t a_hi = da + (uint64_t)((j * kChunkBytes) >> 4), a_lo = a_hi + (uint64_t)(kHalfBytes >> 4);
            const uint64_t kb = (uint64_t)(cij[j] * (256u >> 4));              // two 128-byte core matrices per chunk
            mma_f16_ss(d_addr, a_hi, db0 + kb, cij[j] != 0u);
            mma_f16_ss(d_addr, a_lo, db0 + kb, true);
            mma_f16_ss(d_addr, a_hi, db1 + kb, true);
          }
          tc_commit(&bars->raw_empty[s]);                                      // stage reusable once these MMAs have retired
          if (sg == nch16 - 1) tc_commit(&bars->d_full[dbuf]);                 // all three accumulators of the tile complete
          if (sg == nch16 - 1 && tcount < 4) FR_TL(10 + tcount);
        }
        __syncwarp();
      }
    }
  } else {
    // ================================================================== epilogue (warps 0 .. kEpiWarps-1)
    constexpr int kFPW = C::kFacesPerWarp;
    const int qd = warp & 3;                                      // TMEM lane quarter == warp % 4
    const int v = qd * 32 + lane;                                 // row of the tile == vertex slot of the cluster
    const uint32_t lane_field = (uint32_t)(qd * 32) << 16;
    pdl_wait();                                                   // the prep kernel's poses (and the cleared keys)
    if (threadIdx.x == 0) FR_TL(4);
    for (int i = threadIdx.x; i < kN * kPose16Stride; i += kEpiWarps * 32) s_pose[i] = pose16[(size_t)b0 * kPose16Stride + i];
    if constexpr (!kRaster) {
      const int wq = warp >> 2;                                   // position within the quarter: faces wq * 16 ... of a step
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");   // the epilogue warps only
      uint32_t tcount = 0;
      for (ItemWalk tw(nclusters, nsteps, false, nullptr, bars); tw.next_consumer(); ++tcount) {
        const int tile = tw.tile;
        const uint32_t dbuf = tcount % kDBufs, dph = (tcount / kDBufs) & 1u;
        // vertex of this row, and whether this tile is the one that writes it
        int n = tile * kTileVerts + v;
        bool owner = n < nver;
        if (cluster_vert != nullptr) {
          const int32_t raw = __ldg(cluster_vert + (size_t)tile * kTileVerts + v);
          n = (int)((uint32_t)raw & kVertIdMask);
          owner = raw >= 0 && ((uint32_t)raw & kVertOwner) != 0u;
        }
        mbar_wait(&bars->d_full[dbuf], dph);
        tc_fence_after();
        const uint32_t d_addr = tmem + lane_field + dbuf * kDCols;
#pragma unroll 1
        for (int step = tw.step0; step < tw.step1; ++step) {
          float x[kFPW], y[kFPW], z[kFPW];                        // kFPW consecutive faces of the step per warp
          const int jb = step * kStepFaces + wq * kFPW;
          tmem_ld<kFPW>(d_addr + 0 * kN + jb, x);
          tmem_ld<kFPW>(d_addr + 1 * kN + jb, y);
          tmem_ld<kFPW>(d_addr + 2 * kN + jb, z);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (step == tw.step1 - 1) {                             // accumulators drained: the tensor pipe may reuse them
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->d_empty[dbuf]);
          }
#pragma unroll
          for (int j = 0; j < kFPW; ++j) {
            const int jf = jb + j;                                // face of the batch tile
            const float4* pp = reinterpret_cast<const float4*>(s_pose + jf * kPose16Stride);
            const float4 p0 = pp[0], p1 = pp[1], p2 = pp[2];
            const float P[12] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w, p2.x, p2.y, p2.z, p2.w};
            float X, Y, Z;
            project_vertex(P, x[j], y[j], z[j], im_size, flags, &X, &Y, &Z);
            if (owner && b0 + jf < batch) {
              if (out.planar != nullptr) store_planar(out.planar, b0 + jf, nver, n, X, Y, Z);
              // records go by rank: with a row map in rank order that is the row itself
              if (out.rec != nullptr) store_record(out, b0 + jf, nver, cluster_vert != nullptr ? tile * kTileVerts + v : n, X, Y, Z);
            }
          }
        }
      }
    } else {
      // ---- raster flavour: three groups of 8 warps, each with its own shared-memory tile (raster_tile.cuh)
      constexpr int kGW = C::kGroupWarps, kGT = kGW * 32;
      const int grp = warp / kGW, gw = warp % kGW;                // group; warp within the group
      const int fh = gw >> 2;                                     // which 4 faces of an octet this warp stages
      const int gtid = gw * 32 + lane;                            // thread within the group == triangle slot in the cull
      const int gbar = 1 + grp;                                   // the group's hardware barrier
      RasterSmem& ts = reinterpret_cast<RasterSmem*>(smem + L.raster)[grp];
      const rt::TableView tv = rt::table_view(target.table);
      const int npix = target.width * target.height;
      asm volatile("bar.sync 4, %0;" ::"n"(kEpiWarps * 32) : "memory");   // poses in place (all epilogue warps)
      uint32_t tcount = 0;
      for (ItemWalk tw(nclusters, nsteps, gridDim.y == 1, pool_counter, bars); tw.next_consumer(); ++tcount) {
        const int tile = tw.tile;
        const uint32_t dbuf = tcount % kDBufs, dph = (tcount / kDBufs) & 1u;
        // Octets of this tile: every group starts with one static octet (rotated from tile to tile), the rest are handed out
        // on demand through a shared-memory counter (FR_DYNAMIC_OCTETS): a group that drew a cheap octet -- faces differ a
        // lot with their poses -- takes the next one instead of idling until the others finish their fixed share.
        // Counter protocol: slot t % 4 serves tile t; a group only touches it between the tile's d_full and its own
        // d_empty arrivals; slot (t + 2) % 4 is reset (by each group, before its arrivals for tile t): tile t + 2 cannot be
        // complete before all those arrivals, and tile t - 2, which used that slot last, was finished by every group
        // before tile t could be computed.
        const int first = tw.step0 + (int)((grp + tcount) % 3u);
        uint32_t* const taken = &bars->oct_taken[tcount & 3u];
        // the accumulators are only handed back by warps that have seen them complete: a warp without work in this tile
        // must not run ahead and arrive for a later tile in this one's phase
        mbar_wait(&bars->d_full[dbuf], dph);
        if (threadIdx.x == 0 && tcount == 0) FR_TL(8);
        if (gtid == 0 && pool_counter != nullptr) atomicMax(&bars->raster_pos, tcount + 1u);
        if (first >= tw.step1) {
          if (gtid == 0) bars->oct_taken[(tcount + 2u) & 3u] = 0u;
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->d_empty[dbuf]);
          continue;
        }
        tc_fence_after();
        // vertex of this row and whether this cluster owns it (optional vertex_proj output only)
        int n = 0;
        [[maybe_unused]] int rank = 0;
        bool owner = false;
        if (kRecords || out.planar != nullptr) {
          const int32_t raw = __ldg(cluster_vert + (size_t)tile * kTileVerts + v);
          n = (int)((uint32_t)raw & kVertIdMask);
          owner = raw >= 0 && ((uint32_t)raw & kVertOwner) != 0u;
          if constexpr (kRecords) rank = __ldg(tv.cluster_rank + (size_t)tile * kTileVerts + v);   // record index of this slot
        }
        // the cluster's triangle list: this thread's entry travels in registers until the group's tile is free
        const int tb = __ldg(tv.tri_begin + tile);
        const int ntri_c = __ldg(tv.tri_begin + tile + 1) - tb;
        uint2 te = make_uint2(0u, 0u);
        if (gtid < ntri_c) te = __ldg(tv.tri_entry + tb + gtid);
        const uint32_t d_addr = tmem + lane_field + dbuf * kDCols;
#ifdef FR_TIMELINE
        const long long tl_t0 = clock64();
        int tl_octets = 0;
#endif
        int nxt = -1;
#pragma unroll 1
        for (int o = first; o >= 0; o = nxt) {
          // ---- this warp's share of the octet (faces fh * 4 ...) from the accumulators, projected, in registers: overlaps
          // the tail of the group's previous draw phase
          const bool stager = gw < 8;                             // (warps beyond the first eight of a group only draw)
          float x[kFPW], y[kFPW], z[kFPW];
          const int jb = o * kStepFaces + (fh & 1) * kFPW;        // first of the four faces within the batch tile
          if (stager) {
            tmem_ld<kFPW>(d_addr + 0 * kN + jb, x);
            tmem_ld<kFPW>(d_addr + 1 * kN + jb, y);
            tmem_ld<kFPW>(d_addr + 2 * kN + jb, z);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          }
          const int fb = b0 + o * kStepFaces;                     // first face of the octet
          const int nlive = min(kStepFaces, batch - fb);
          float4 r[kFPW];
          if (stager) {
#pragma unroll
            for (int j = 0; j < kFPW; ++j) {
              const float4* pp = reinterpret_cast<const float4*>(s_pose + (jb + j) * kPose16Stride);
              const float4 p0 = pp[0], p1 = pp[1], p2 = pp[2];
              const float P[12] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w, p2.x, p2.y, p2.z, p2.w};
              project_vertex(P, x[j], y[j], z[j], im_size, flags, &r[j].x, &r[j].y, &r[j].z);
              r[j].w = __uint_as_float(fr_snap_code(r[j].x, r[j].y, target.width, target.height));
              if constexpr (kRecords) {                               // vertices + the records the resolve pass gathers normals from
                if (owner && fh * kFPW + j < nlive) {
                  if (out.planar != nullptr) store_planar(out.planar, fb + fh * kFPW + j, nver, n, r[j].x, r[j].y, r[j].z);
                  out.rec[(size_t)(fb + fh * kFPW + j) * nver + rank] = r[j];
                }
              } else {
                if (owner && fh * kFPW + j < nlive) store_planar(out.planar, fb + fh * kFPW + j, nver, n, r[j].x, r[j].y, r[j].z);
              }
            }
          }
          asm volatile("bar.sync %0, %1;" ::"r"(gbar), "n"(kGT) : "memory");   // the group's previous draw is complete
          if (o == first && gtid < ntri_c)                        // (same packing as rt::load_tris; o == first only once: grabs are > first)
            ts.tri[gtid] = make_uint4((te.x & 0xFFu) << 4, ((te.x >> 8) & 0xFFu) << 4, ((te.x >> 16) & 0xFFu) << 4, (0x7FFFFFFFu - te.y) << 1);
          if (gtid == 0) {
            ts.count = 0u;
            const int want = tw.step0 + 3 + (int)atomicAdd(taken, 1u);          // the group's next octet of this tile, if any is left
            ts.pad[0] = (unsigned)(want < tw.step1 ? want : -1);
          }
          if (stager) {
#pragma unroll
            for (int j = 0; j < kFPW; ++j) ts.rec[fh * kFPW + j][v] = r[j];              // the staged records ...
            ts.code4[fh][v] = make_uint4(__float_as_uint(r[0].w), __float_as_uint(r[1].w), __float_as_uint(r[2].w),
                                         __float_as_uint(r[3].w));                       // ... and their code words again
          }
          asm volatile("bar.sync %0, %1;" ::"r"(gbar), "n"(kGT) : "memory");   // staged octet (and triangle list) complete
          nxt = (int)ts.pad[0];
          if (nxt < 0) {                                          // the group's last octet of this tile: accumulators drained
            if (gtid == 0) bars->oct_taken[(tcount + 2u) & 3u] = 0u;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->d_empty[dbuf]);
          }
          if ((gw << 5) < ntri_c)                                 // cull: thread = triangle, the octet's 8 faces
            rt::cull_triangle<kStepFaces, kStepFaces / 4>(ts, gtid, 0, (gtid < ntri_c) ? ((1u << nlive) - 1u) : 0u, lane);
          asm volatile("bar.sync %0, %1;" ::"r"(gbar), "n"(kGT) : "memory");   // survivor list complete
          rt::draw_list(ts, gtid, kGT, target.keys + (size_t)fb * npix, npix, target.width, target.height);
#ifdef FR_TIMELINE
          ++tl_octets;
#endif
        }
#ifdef FR_TIMELINE
        asm volatile("bar.sync %0, %1;" ::"r"(gbar), "n"(kGT) : "memory");
        if (grp == 0 && gtid == 0 && blockIdx.y == 0) g_cluster_cost[tile] = (float)(clock64() - tl_t0) / (float)max(tl_octets, 1);
#endif
      }
    }
  }

  // ---- teardown
  if (threadIdx.x == 0) FR_TL(9);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) FR_TL(14);
  FR_MARK_MAX(3);
  if (warp == kProducerWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)kTmemCols) : "memory");
  }
}

}  // namespace f16

// fp16 coefficient operands of every 64-face batch tile + pose16
inline size_t recon_f16_bsplit_bytes(int batch, const BasisGeom& g) {
  return (size_t)ceil_div(batch, f16::kN) * 2 * (f16::kN / 8) * (size_t)(g.kpad16 / 8) * 128;
}
inline size_t recon_f16_pose_bytes(int batch) { return sizeof(float) * (size_t)batch_padded(batch) * f16::kPose16Stride; }

inline bool recon_f16_fits(const BasisGeom& g, bool raster) {
  return (raster ? f16::smem_layout<true>(g.nch16).total : f16::smem_layout<false>(g.nch16).total) <= 227u * 1024u;
}

// Prep + tensor-core forward.  target == nullptr: the planar flavour (out.planar and / or out.rec); else the raster flavour
// (the epilogue rasterizes; out optional; needs cluster tiles).  clear_keys / clear_bytes: visibility keys the prep kernel
// clears on the side (bytes must be a multiple of 16), or nullptr.
// cluster_vert: device pointer to the mesh table's vertex lists the basis was packed with, or null (consecutive tiles).
inline int launch_recon_fwd_f16(const float* params, const float* packed, void* bsplit, float* pose16, ReconOut out,
                                const f16::RasterTarget* target, const int32_t* cluster_vert, unsigned long long* clear_keys,
                                size_t clear_bytes, int batch, int nver, const BasisGeom& g, float im_size, unsigned flags, int nsm,
                                cudaStream_t st) {
  static_assert(f16::kN == kBatchPad, "batch tiles are kBatchPad faces");
  const unsigned char* base = reinterpret_cast<const unsigned char*>(packed);
  const float* inv_scale = reinterpret_cast<const float*>(base + g.scale_offset());
  const int bpad = batch_padded(batch);
  const int dparam = FR_NDIM_POSE + g.ks + g.ke;
  const int nclear = clear_keys ? 2 * nsm : 0;
  FR_CUDA(launch_pdl(f16::recon_prep_f16_kernel, dim3(bpad + nclear), dim3(256), 0, st, pdl_enabled(), params, inv_scale, dparam, batch,
                     g.ks, g.ke, g.kpad16, flags, im_size, static_cast<unsigned char*>(bsplit), pose16, bpad, clear_keys,
                     clear_bytes / 16, target != nullptr ? target->pool_counter : static_cast<unsigned*>(nullptr)));   // (waits for everything before it in the stream at its first instruction)
  FR_LAUNCHED("recon_prep_f16_kernel");
  const int nbt = ceil_div(batch, f16::kN);
  const int ntiles_f = target != nullptr ? g.nclusters : g.ntiles;             // row tiles of the section this flavour streams
  const f16::RasterTarget none = {nullptr, nullptr, 0, 0};
  const f16::SmemLayout L = target != nullptr ? f16::smem_layout<true>(g.nch16) : f16::smem_layout<false>(g.nch16);
  const bool records = target != nullptr && out.rec != nullptr;
  if (records)
    FR_CUDA(cudaFuncSetAttribute(f16::recon_fwd_f16_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  else if (target != nullptr)
    FR_CUDA(cudaFuncSetAttribute(f16::recon_fwd_f16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  else
    FR_CUDA(cudaFuncSetAttribute(f16::recon_fwd_f16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  // One persistent CTA per SM; a batch tile (64 faces) is shared by `ctas` CTAs.  When nsm / nbt leaves many SMs without a
  // CTA (64 batch tiles: 2 x 64 = 128 of 148), the batch tiles go out in several launches that each fill the GPU: the
  // first groups with one CTA more per batch tile (49 tiles x 3 CTAs, then 15 x 9: 227 cluster passes per SM instead of 255).
  for (int bt0 = 0; bt0 < nbt;) {
    const int rem = nbt - bt0;
    int ctas = nsm / rem, nb = rem;
    if (ctas < 1) ctas = 1;
    if (ctas < ntiles_f && rem * ctas * 10 < nsm * 9 && nsm / (ctas + 1) >= 1) {   // < 90 % of the SMs: one CTA more, fewer tiles
      ++ctas;
      nb = nsm / ctas;
    }
    if (ctas > ntiles_f) ctas = ntiles_f;
    if (records) {
      FR_CUDA(launch_pdl(f16::recon_fwd_f16_kernel<true, true>, dim3(ctas, nb), dim3(f16::Cfg<true>::kThreads), L.total, st, pdl_enabled(),
                         base + g.f16c_offset(), static_cast<const unsigned char*>(bsplit), static_cast<const float*>(pose16), cluster_vert,
                         out, *target, batch, nver, g.nch16, g.nclusters, im_size, flags, bt0));
    } else if (target != nullptr) {
      FR_CUDA(launch_pdl(f16::recon_fwd_f16_kernel<true>, dim3(ctas, nb), dim3(f16::Cfg<true>::kThreads), L.total, st, pdl_enabled(),
                         base + g.f16c_offset(), static_cast<const unsigned char*>(bsplit), static_cast<const float*>(pose16), cluster_vert,
                         out, *target, batch, nver, g.nch16, g.nclusters, im_size, flags, bt0));
    } else {
      FR_CUDA(launch_pdl(f16::recon_fwd_f16_kernel<false>, dim3(ctas, nb), dim3(f16::Cfg<false>::kThreads), L.total, st, pdl_enabled(),
                         base + g.f16_offset(), static_cast<const unsigned char*>(bsplit), static_cast<const float*>(pose16), cluster_vert,
                         out, none, batch, nver, g.nch16, g.ntiles, im_size, flags, bt0));
    }
    FR_LAUNCHED("recon_fwd_f16_kernel");
    bt0 += nb;
  }
  return FR_OK;
}

}  // namespace fr

#endif  // FR_RECON_F16_CUH_
