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

// plain fp32 copy of the mean, [3][ntiles*128] (zero beyond nver)
__global__ void __launch_bounds__(256)
pack_mean_kernel(const float* __restrict__ mu, int nver, int ntiles, unsigned flags, float* __restrict__ mean32) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const int npad = ntiles * kTileVerts;
  if (idx >= 3 * npad) return;
  const int c = idx / npad, n = idx - c * npad;
  mean32[idx] = (n < nver) ? mu[(flags & FR_MEAN_INTERLEAVED) ? (size_t)3 * n + c : (size_t)c * nver + n] : 0.0f;
}

// ---------------------------------------------------------------------------------------------- dv operand (per call)
// block = 8 faces (one per warp) x 256 vertices.  Each warp stages its face's three gradient rows in shared memory with
// coalesced loads (skewed so that a lane can then read ITS 8 consecutive vertices without bank conflicts), rotates, scales
// and splits them, and the fp16 pieces go back through shared memory so that every global store covers a full 128-byte
// core matrix (8 faces x 8 vertices).  grid (ceil(ntiles*128 / 256), ceil(batch / 8)).  gmax[b*4 + r] = max |g[b][r][:]|.
// The MEAN column of the contraction, G[b][kmean] = sum mu[(c,n)] dv[b,c,n], is accumulated here in fp32 and not on the
// tensor cores: it is orders of magnitude larger than the other columns, d f = sum_k coef_k G_k cancels against it, and the
// tensor core's truncating accumulation (a relative bias of a few 1e-6 over the ~200 accumulation steps of a CTA, always
// towards zero) would eat the 1e-4 tolerance of d f.
constexpr int kPgVerts = 256, kPgRow = kPgVerts + kPgVerts / 32;            // skew: one pad float per 32
__device__ __forceinline__ int pg_skew(int v) { return v + (v >> 5); }
__global__ void __launch_bounds__(256)
recon_bwd_pack_grad_kernel(const float* __restrict__ vertex_grad, const float* __restrict__ pose, const float* __restrict__ gmax,
                           const float* __restrict__ mean32, int batch, int nver, int ntiles, int nb,
                           unsigned flags, unsigned char* __restrict__ gtiles, float* __restrict__ gscale,
                           double* __restrict__ gmean64) {
  // the staged gradient rows and the outgoing fp16 pieces share one buffer (a barrier separates the two uses)
  __shared__ __align__(16) unsigned char s_buf[sizeof(float) * 8 * 3 * kPgRow];
  __shared__ float s_mu[3][kPgRow];
  float (*s_g)[3][kPgRow] = reinterpret_cast<float (*)[3][kPgRow]>(s_buf);
  uint4 (*pieces)[2][32][8] = reinterpret_cast<uint4 (*)[2][32][8]>(s_buf);   // [c][hi|lo][vertex group][face]
  static_assert(sizeof(uint4) * 3 * 2 * 32 * 8 <= sizeof(float) * 8 * 3 * kPgRow, "pieces must fit in the staging buffer");
  const int lane = threadIdx.x & 31, f8 = threadIdx.x >> 5;
  const int b = blockIdx.y * 8 + f8;
  const int n0 = blockIdx.x * kPgVerts, npad = ntiles * kTileVerts;
  const bool live = b < batch;
  for (int i = threadIdx.x; i < 3 * kPgVerts; i += 256) {
    const int c = i / kPgVerts, v = i - c * kPgVerts;
    s_mu[c][pg_skew(v)] = (n0 + v < npad) ? mean32[(size_t)c * npad + n0 + v] : 0.0f;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* row = vertex_grad + ((size_t)b * 3 + c) * nver;
#pragma unroll
    for (int t = 0; t < kPgVerts / 32; ++t) {
      const int v = t * 32 + lane;
      s_g[f8][c][pg_skew(v)] = (live && n0 + v < nver) ? row[n0 + v] : 0.0f;
    }
  }
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
    if (blockIdx.x == 0 && lane == 0) gscale[b] = ldexpf(1.0f, -u);
  }
  const float ysign = (flags & FR_YFLIP_NONE) ? 1.0f : -1.0f;
  float R[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (live) {
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = pose[(size_t)b * kPoseStride + 12 + i];
  }
  __syncthreads();
  float gmean = 0.0f;
  __half hi[3][8], lo[3][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int sv = pg_skew(lane * 8 + i);
    const float gx = s_g[f8][0][sv], gy = ysign * s_g[f8][1][sv], gz = s_g[f8][2][sv];
    const float d0 = fmaf(R[6], gz, fmaf(R[3], gy, R[0] * gx));             // same expression as recon_bwd_simt_kernel
    const float d1 = fmaf(R[7], gz, fmaf(R[4], gy, R[1] * gx));
    const float d2 = fmaf(R[8], gz, fmaf(R[5], gy, R[2] * gx));
    gmean = fmaf(s_mu[0][sv], d0, gmean);
    gmean = fmaf(s_mu[1][sv], d1, gmean);
    gmean = fmaf(s_mu[2][sv], d2, gmean);
    const float d[3] = {d0 * up, d1 * up, d2 * up};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      hi[c][i] = __float2half_rn(d[c]);
      lo[c][i] = __float2half_rn(d[c] - __half2float(hi[c][i]));
    }
  }
  __syncthreads();                                                         // every thread is done with s_g
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    memcpy(&pieces[c][0][lane][f8], hi[c], 16);
    memcpy(&pieces[c][1][lane][f8], lo[c], 16);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gmean += __shfl_xor_sync(0xFFFFFFFFu, gmean, o);
  if (lane == 0 && live) atomicAdd(gmean64 + b, (double)gmean);           // 200+ partial sums of ~1e7 per face: keep them exact
  __syncthreads();
  // write out: one 16-byte piece per thread and trip, 8 consecutive threads = one 128-byte core matrix
  const int b8 = blockIdx.y * 8;                                           // first face of the block (multiple of 8)
  const int bt = b8 / nb, fg = (b8 % nb) >> 3;
  const size_t stage_b = b_stage_bytes(nb);
  for (int p = threadIdx.x; p < 3 * 2 * 32 * 8; p += 256) {
    const int ff = p & 7, vgl = (p >> 3) & 31, half = (p >> 8) & 1, c = p >> 9;
    const int vgg = blockIdx.x * 32 + vgl;
    if (vgg >= ntiles * (kTileVerts / 8)) continue;
    const int t2 = vgg / (kTileVerts / 8), j = (vgg % (kTileVerts / 8)) / 2, vh = vgg & 1;
    unsigned char* dst = gtiles + (((size_t)bt * ntiles + t2) * 3 + c) * kChunksPerTile * stage_b + (size_t)j * stage_b +
                         (size_t)half * (stage_b / 2) + (size_t)fg * 256 + vh * 128 + ff * 16;
    *reinterpret_cast<uint4*>(dst) = pieces[c][half][vgl][ff];
  }
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
        const uint64_t da = tc::make_smem_desc(sa, 2048u, 128u);
        const uint64_t db_hi = tc::make_smem_desc(sa + a_bytes, 128u, 256u), db_lo = db_hi + (uint64_t)((b_bytes / 2) >> 4);
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
