// Tensor-core backward of the reconstruction:  G[b,k] = sum_{c,n} P[(c,n),k] * dv[b,c,n],  dv = R^T g~  (SURVEY App. A.4;
// recon.cuh recon_bwd_simt_kernel is the FFMA flavour and defines the semantics).  The contraction runs over the VERTICES,
// so both operands are re-tiled with the vertex index as the MMA K dimension:
//   A (rows = k, 128 per M tile):  the packed basis' fourth section -- the forward's fp16 hi/lo pairs of the column-scaled
//      basis, transposed once at pack time: per (tile, coordinate, 16-vertex chunk)  [hi m0 | hi m1 | lo m0 | lo m1],  each a
//      128 x 16 canonical K-major no-swizzle operand tile (4 KB)
//   B (rows = faces, up to 256 per batch tile):  dv 2^u_b split into fp16 hi + lo by recon_bwd_pack_grad_kernel, per
//      (tile, coordinate, chunk)  [hi | lo],  each NB x 16 in the same canonical layout
//   D[m] (TMEM, fp32, 128 lanes x NB columns per M tile)  +=  A_hi.B_hi + A_lo.B_hi + A_hi.B_lo   (tcgen05.mma kind::f16, K16)
// One persistent CTA per SM accumulates its share of the vertex tiles in TMEM and adds  D 2^-s_k 2^-u_b  to G with one
// coalesced fp32 reduction per accumulator at the end.  HBM traffic: the basis section once per 256 faces + dv once.
#ifndef FR_RECON_BWD_F16_CUH_
#define FR_RECON_BWD_F16_CUH_

#include "recon_f16.cuh"

namespace fr {
namespace b16 {

using tc::bulk_load;
using tc::elect_one;
using tc::mbar_arrive_expect_tx;
using tc::mbar_init;
using tc::mbar_wait;
using tc::smem_u32;
using tc::tc_commit;
using tc::tc_fence_after;
using tc::tc_fence_before;
using tc::tmem_ld16;

constexpr int kChunkVerts = 16;                       // vertices per stage == one f16 MMA K step
constexpr int kChunksPerTile = kTileVerts / kChunkVerts;
constexpr int kMaxFaces = 256;                        // faces per batch tile (MMA N)
constexpr int kStages = 6;
constexpr int kTmemCols = 512;
constexpr int kWarps = 10, kProducerWarp = 8, kMmaWarp = 9;
constexpr int kThreads = kWarps * 32;

__host__ __device__ inline int faces_per_tile(int batch) { return batch >= kMaxFaces ? kMaxFaces : (batch + 63) / 64 * 64; }
__host__ __device__ inline uint32_t a_stage_bytes(int mtiles) { return 2u * (uint32_t)mtiles * 4096u; }
__host__ __device__ inline uint32_t b_stage_bytes(int nb) { return 2u * (uint32_t)nb * 32u; }

struct Barriers {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t d_full;
  uint32_t tmem_base;
  uint32_t pad;
};

// dv operand tiles of one call: [batch tile][tile][c][chunk][hi | lo][face / 8][vertex half][face % 8][vertex % 8] fp16
inline size_t grad_tiles_bytes(int batch, const BasisGeom& g) {
  const int nb = faces_per_tile(batch);
  const int nbt = (batch + nb - 1) / nb;
  return (size_t)nbt * g.ntiles * 3 * kChunksPerTile * b_stage_bytes(nb);
}

// ---------------------------------------------------------------------------------------------- packing (once per model)
// One thread writes one 16-byte piece: 8 consecutive vertices of one k row.
__global__ void __launch_bounds__(256)
pack_basis_bwd_kernel(const float* __restrict__ pc_shape, const float* __restrict__ pc_exp,
                      const float* __restrict__ inv_scale, int nver, int ks, int ke, int mtiles, int ntiles, unsigned flags,
                      uint4* __restrict__ tiles) {
  const size_t per_chunk = (size_t)mtiles * 2 * kTileVerts;               // (m, vertex half, row) pieces of one hi (or lo) block
  const size_t total = (size_t)ntiles * 3 * kChunksPerTile * per_chunk;
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int row = (int)(idx % kTileVerts);
  const int vh = (int)((idx / kTileVerts) % 2);
  const int m = (int)((idx / (2 * kTileVerts)) % mtiles);
  const int j = (int)((idx / per_chunk) % kChunksPerTile);
  const int c = (int)((idx / (per_chunk * kChunksPerTile)) % 3);
  const int tile = (int)(idx / (per_chunk * kChunksPerTile * 3));
  const int k = m * 128 + row;
  const float up = (k <= ks + ke) ? 1.0f / inv_scale[k] : 0.0f;          // power of two: exact
  __half hi[8], lo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n = tile * kTileVerts + j * kChunkVerts + vh * 8 + i;
    float x = 0.0f;
    if (n < nver && k < ks + ke) {      // the mean row stays zero here: its contraction is done in fp32 (see below)
      const size_t row_b = (flags & FR_BASIS_INTERLEAVED) ? (size_t)3 * n + c : (size_t)c * nver + n;
      if (k < ks) x = pc_shape[row_b * ks + k];
      else x = pc_exp[row_b * ke + (k - ks)];
    }
    const float xs = x * up;
    hi[i] = __float2half_rn(xs);
    lo[i] = __float2half_rn(xs - __half2float(hi[i]));
  }
  uint4 whi, wlo;
  memcpy(&whi, hi, 16);
  memcpy(&wlo, lo, 16);
  const size_t stage = ((size_t)(tile * 3 + c) * kChunksPerTile + j) * (2 * per_chunk);   // in 16-byte pieces
  const size_t piece = (size_t)m * (2 * kTileVerts) + (size_t)vh * kTileVerts + row;
  tiles[stage + piece] = whi;
  tiles[stage + per_chunk + piece] = wlo;
}

// ---------------------------------------------------------------------------------------------- dv operand (per call)
// block (64 faces, 4 vertex groups of 8), kGroupsPerThread groups per thread; grid (ceil(ntiles*16 / (4*kGroupsPerThread)),
// ceil(batch / 64)).  gmax[b*4 + r] = max |g[b][r][:]|.
// The MEAN column of the contraction, G[b][kmean] = sum mu[(c,n)] dv[b,c,n], is accumulated here in fp32 (mu from the fp32
// section of the packed basis) and not on the tensor cores: it is orders of magnitude larger than the other columns,
// d f = sum_k coef_k G_k cancels against it, and the tensor core's truncating accumulation (a relative bias of a few 1e-6
// over the ~200 accumulation steps of a CTA) would eat the 1e-4 tolerance of d f.
constexpr int kGroupsPerThread = 8;
__global__ void __launch_bounds__(256)
recon_bwd_pack_grad_kernel(const float* __restrict__ vertex_grad, const float* __restrict__ pose, const float* __restrict__ gmax,
                           const float4* __restrict__ packed32, int kg, int kmean, int kpad, int batch, int nver, int ntiles,
                           int nb, unsigned flags, unsigned char* __restrict__ gtiles, float* __restrict__ gscale,
                           float* __restrict__ G) {
  __shared__ float red[4][64];
  const int fl64 = threadIdx.x & 63, vq = threadIdx.x >> 6;
  const int b = blockIdx.y * 64 + fl64;
  const bool live = b < batch;
  // per-face power-of-two scale: |dv| <= sqrt(3) max|g| < 2 max|g|;  2 max|g| 2^u in [2^13, 2^14)
  float up = 1.0f;
  if (live) {
    const float gm = 2.0f * fmaxf(fmaxf(gmax[b * 4 + 0], gmax[b * 4 + 1]), gmax[b * 4 + 2]);
    int u = 0;
    if (gm > 0.0f && gm < 3.0e38f) {
      int e;
      frexpf(gm, &e);
      u = max(-100, min(100, 14 - e));
    }
    up = ldexpf(1.0f, u);
    if (blockIdx.x == 0 && vq == 0) gscale[b] = ldexpf(1.0f, -u);
  }
  const float ysign = (flags & FR_YFLIP_NONE) ? 1.0f : -1.0f;
  float R[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (live) {
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = pose[(size_t)b * kPoseStride + 12 + i];
  }
  const int bt = b / nb, fl = b % nb;
  const size_t stage_b = b_stage_bytes(nb);
  float gmean = 0.0f;
  for (int r = 0; r < kGroupsPerThread; ++r) {
    const int vg = (blockIdx.x * kGroupsPerThread + r) * 4 + vq;           // global group of 8 vertices
    if (vg >= ntiles * (kTileVerts / 8)) break;
    const int tile = vg / (kTileVerts / 8), j = (vg % (kTileVerts / 8)) / 2, vh = vg & 1;
    __half hi[3][8], lo[3][8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = vg * 8 + i;
      float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f;
      if (live && n < nver) {
        const float* gp = vertex_grad + (size_t)b * 3 * nver + n;
        const float gx = gp[0], gy = ysign * gp[nver], gz = gp[2 * (size_t)nver];
        d0 = fmaf(R[6], gz, fmaf(R[3], gy, R[0] * gx));                     // same expression as recon_bwd_simt_kernel
        d1 = fmaf(R[7], gz, fmaf(R[4], gy, R[1] * gx));
        d2 = fmaf(R[8], gz, fmaf(R[5], gy, R[2] * gx));
        const float4* mp = packed32 + ((size_t)(tile * 3) * kg + (kmean >> 2)) * kTileVerts + (n - tile * kTileVerts);
        gmean = fmaf(f4_get(mp[0], kmean & 3), d0, gmean);
        gmean = fmaf(f4_get(mp[(size_t)kg * kTileVerts], kmean & 3), d1, gmean);
        gmean = fmaf(f4_get(mp[(size_t)2 * kg * kTileVerts], kmean & 3), d2, gmean);
      }
      const float d[3] = {d0 * up, d1 * up, d2 * up};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        hi[c][i] = __float2half_rn(d[c]);
        lo[c][i] = __float2half_rn(d[c] - __half2float(hi[c][i]));
      }
    }
    unsigned char* base = gtiles + ((size_t)bt * ntiles + tile) * 3 * kChunksPerTile * stage_b;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      unsigned char* st = base + ((size_t)c * kChunksPerTile + j) * stage_b + (size_t)(fl >> 3) * 256 + vh * 128 + (fl & 7) * 16;
      uint4 whi, wlo;
      memcpy(&whi, hi[c], 16);
      memcpy(&wlo, lo[c], 16);
      *reinterpret_cast<uint4*>(st) = whi;
      *reinterpret_cast<uint4*>(st + stage_b / 2) = wlo;
    }
  }
  red[vq][fl64] = gmean;
  __syncthreads();
  if (vq == 0 && live) atomicAdd(G + (size_t)b * kpad + kmean, (red[0][fl64] + red[1][fl64]) + (red[2][fl64] + red[3][fl64]));
}

// ---------------------------------------------------------------------------------------------- the contraction
__device__ __forceinline__ void mma_f16_ss_n(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}

// grid (ctas, batch tiles).  G [bpad][kpad] must be zero on entry.
__global__ void __launch_bounds__(kThreads, 1)
recon_bwd_f16_kernel(const unsigned char* __restrict__ atiles, const unsigned char* __restrict__ gtiles,
                     const float* __restrict__ inv_scale, const float* __restrict__ gscale, float* __restrict__ G, int batch,
                     int nb, int mtiles, int ntiles, int kpad) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t a_bytes = a_stage_bytes(mtiles), b_bytes = b_stage_bytes(nb), stage_bytes = a_bytes + b_bytes;
  Barriers* bars = reinterpret_cast<Barriers*>(smem + kStages * stage_bytes);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int b0 = blockIdx.y * nb;
  const uint32_t my_tiles = (blockIdx.x < (unsigned)ntiles) ? ((uint32_t)(ntiles - 1 - blockIdx.x) / gridDim.x + 1u) : 0u;
  const uint32_t per_tile = 3u * kChunksPerTile;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&bars->full[i], 1);
      mbar_init(&bars->empty[i], 1);
    }
    mbar_init(&bars->d_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kProducerWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == kProducerWarp) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const unsigned char* asrc = atiles + (size_t)tile * per_tile * a_bytes;
        const unsigned char* bsrc = gtiles + ((size_t)blockIdx.y * ntiles + tile) * per_tile * b_bytes;
        for (uint32_t q = 0; q < per_tile; ++q, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
          mbar_wait(&bars->empty[s], ph ^ 1u);
          mbar_arrive_expect_tx(&bars->full[s], stage_bytes);
          bulk_load(smem + s * stage_bytes, asrc + (size_t)q * a_bytes, a_bytes, &bars->full[s]);
          bulk_load(smem + s * stage_bytes + a_bytes, bsrc + (size_t)q * b_bytes, b_bytes, &bars->full[s]);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(nb >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t total = my_tiles * per_tile;
#pragma unroll 1
    for (uint32_t it = 0; it < total; ++it) {
      const uint32_t s = it % kStages;
      mbar_wait(&bars->full[s], (it / kStages) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + s * stage_bytes);
        const uint64_t da = tc::make_b_desc(sa, 2048u, 128u);
        const uint64_t db_hi = tc::make_b_desc(sa + a_bytes, 128u, 256u), db_lo = db_hi + (uint64_t)((b_bytes / 2) >> 4);
        for (int m = 0; m < mtiles; ++m) {
          const uint32_t d = tmem + (uint32_t)(m * nb);
          const uint64_t a_hi = da + (uint64_t)((m * 4096) >> 4), a_lo = a_hi + (uint64_t)((mtiles * 4096) >> 4);
          mma_f16_ss_n(d, a_hi, db_hi, idesc, it != 0u);
          mma_f16_ss_n(d, a_lo, db_hi, idesc, true);
          mma_f16_ss_n(d, a_hi, db_lo, idesc, true);
        }
        tc_commit(&bars->empty[s]);
        if (it == total - 1u) tc_commit(&bars->d_full);
      }
      __syncwarp();
    }
  }

  // ---- epilogue: every CTA that accumulated something adds its partial sums to G
  if (warp < 8 && my_tiles > 0u) {
    mbar_wait(&bars->d_full, 0);
    tc_fence_after();
    const int q = warp & 3;                                        // TMEM lane quarter
    const uint32_t lane_field = (uint32_t)(q * 32) << 16;
    const int jb0 = (warp >> 2) * (nb / 2), jb1 = jb0 + nb / 2;    // this warp's faces of the batch tile
    for (int m = 0; m < mtiles; ++m) {
      const int k = m * 128 + q * 32 + lane;
      const float sk = (k < kpad) ? inv_scale[min(k, kpad - 1)] : 0.0f;   // kpad <= kpad16: scales exist for every k < kpad
      for (int jb = jb0; jb < jb1; jb += 16) {
        float v[16];
        tmem_ld16(tmem + lane_field + (uint32_t)(m * nb + jb), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (k < kpad) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int b = b0 + jb + i;
            if (b < batch) atomicAdd(G + (size_t)b * kpad + k, v[i] * sk * gscale[b]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kProducerWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)kTmemCols) : "memory");
  }
}

inline uint32_t smem_bytes(int mtiles, int nb) { return kStages * (a_stage_bytes(mtiles) + b_stage_bytes(nb)) + (uint32_t)sizeof(Barriers); }

}  // namespace b16

inline bool recon_bwd_f16_fits(const BasisGeom& g) { return g.mtiles() * b16::kMaxFaces <= b16::kTmemCols; }

}  // namespace fr
#endif  // FR_RECON_BWD_F16_CUH_
