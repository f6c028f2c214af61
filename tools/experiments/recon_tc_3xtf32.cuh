// Tensor-core flavour of the reconstruction + projection forward pass: tcgen05 (5th-gen tensor cores) with TMEM
// accumulators, 3xTF32 operand splitting for fp32-level accuracy, bulk-async (TMA engine) streaming of the basis.
//
//   V[b, (c,n)] = sum_k P[(c,n), k] * coef[b, k]        M = 128 vertices per tile (x3 coordinates), N = 64 faces, K = kpad
//
// Per CTA (persistent, one per SM, 18 warps):
//   warp 16  producer   cp.async.bulk (TMA engine): the resident B operand (pre-split coefficients of this batch tile) once,
//                       then 8 KB basis chunks (16 k-columns x 128 rows, contiguous in the packed layout) into an
//                       8-stage shared-memory ring, completion on mbarriers
//   warps 0-7 converter two groups of 4 warps alternate chunks: shared memory -> registers, split every fp32 value into
//                       hi = top 19 bits (exact tf32) and lo = x - hi, tcgen05.st both into a TMEM ring as the A operand
//   warp 17  MMA issuer one thread: per k8 step  D += A_hi.B_hi + A_lo.B_hi + A_hi.B_lo   (tcgen05.mma kind::tf32, A from
//                       TMEM, B = pre-split coefficients resident in shared memory in the canonical K-major layout)
//   warps 8-15 epilogue tcgen05.ld of the three 128x64 accumulators (x, y, z of the same vertices), (f.R).v + t, y flip,
//                       coalesced stores of vertex_proj; double-buffered against the next tile's MMAs
// The hi/lo split follows the 3xTF32 scheme (drop lo.lo): relative error ~2^-21 per product instead of 2^-11.
#ifndef FR_RECON_TC_CUH_
#define FR_RECON_TC_CUH_

#include <cstdlib>

#include "fr_common.cuh"
#include "recon.cuh"

namespace fr {
namespace tc {

constexpr int kN = 64;               // faces per batch tile (MMA N)
constexpr int kChunkGroups = 4;      // float4 k-groups per chunk  -> 16 k columns, 8 KB of basis
constexpr int kChunkK = kChunkGroups * 4;
constexpr int kRawStages = 8;        // shared-memory ring of raw basis chunks (8 x 8 KB)
constexpr int kDCols = 3 * kN;       // one accumulator set: x, y, z
constexpr int kTmemCols = 512;       // [0, DBUFS*192): accumulator sets; the rest: ring of split A chunks (32 columns each)
constexpr int kMaxAStages = 10;
constexpr int kEpiWarps = 8;         // two per TMEM lane quarter, each draining half of the tile's faces
constexpr int kConvWarps = 8, kEpiWarp0 = 8, kProducerWarp = kEpiWarp0 + kEpiWarps, kMmaWarp = kProducerWarp + 1;
constexpr int kThreads = (kMmaWarp + 1) * 32;
constexpr uint32_t kChunkBytes = kChunkGroups * kTileVerts * 16;

// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6), a=TF32 [7,10), b=TF32 [10,13), K-major A/B,
// n_dim = N>>3 at [17,23), m_dim = M>>4 at [24,29)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kN >> 3) << 17) | ((128u >> 4) << 24);

struct Barriers {
  uint64_t raw_full[kRawStages];
  uint64_t raw_empty[kRawStages];
  uint64_t a_full[2];
  uint64_t a_empty[2];
  uint64_t d_full[2];
  uint64_t d_empty[2];
  uint64_t b_full;
  uint32_t tmem_base;
  uint32_t pad;
};

// Optional per-role cycle accounting (developer diagnostics, FR_TC_DEBUG=1): [role][slot] accumulated clock64 deltas of
// block 0.  role 0 = converter warp 0, 1 = MMA issuer, 2 = epilogue warp 8, 3 = producer.
__device__ unsigned long long g_tc_dbg[4][8];
#ifdef FR_TC_INSTRUMENT   // compile with -DFR_TC_INSTRUMENT for the per-role cycle counters (costs issue slots in every role)
#define TC_T(var) const long long var = dbg ? clock64() : 0
#define TC_ACC(role, slot, t0, t1) do { if (dbg) g_tc_dbg[role][slot] += (unsigned long long)((t1) - (t0)); } while (0)
#else
#define TC_T(var) do { } while (0)
#define TC_ACC(role, slot, t0, t1) do { } while (0)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded spin: a protocol bug traps (-> launch error the API reports) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// one lane of a fully converged warp (keeps tcgen05.mma / commit on the uniform datapath: a lane-0 branch makes the
// compiler serialise every uniform-register operand through per-thread loops and triples the issue cost)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem_d] (+)= A[tmem_a] . B[smem desc]   (A: 128 lanes x 8 columns of tf32 in TMEM)
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(kIdesc), "r"((uint32_t)accumulate)
      : "memory");
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), K-major, no swizzle: 8-row x 16-byte core matrices,
// LBO = bytes between the two 16-byte K chunks of one k8 step, SBO = bytes between 8-row groups; version 1 (sm_100)
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory carve-up (dynamic, 128-byte aligned base)
struct SmemLayout {
  uint32_t b_hi, b_lo, raw, pose, bars, total;
  uint32_t sbo;  // bytes between 8-face groups of the coefficient operand
};
__host__ __device__ inline SmemLayout smem_layout(int kg) {
  SmemLayout L;
  L.sbo = (uint32_t)kg * 128u;                       // kg core matrices (8 faces x 4 k) per 8-face group
  const uint32_t bsz = (kN / 8) * L.sbo;
  L.b_hi = 0;
  L.b_lo = bsz;
  L.raw = 2 * bsz;
  L.pose = L.raw + kRawStages * kChunkBytes;
  L.bars = L.pose + kN * kPoseStride * 4;
  L.total = L.bars + (uint32_t)sizeof(Barriers);
  return L;
}

template <int DBUFS>
__global__ void __launch_bounds__(kThreads, 1)
recon_fwd_tc_kernel(const float4* __restrict__ packed, const unsigned char* __restrict__ bsplit, const float* __restrict__ pose,
                    ReconOut out, int batch, int nver, int kg, int ntiles, float im_size,
                    unsigned flags, int debug) {
  constexpr int kACol0 = DBUFS * kDCols;
  constexpr int kAStages = (kTmemCols - kACol0) / (2 * kChunkK);
  static_assert(kAStages % 2 == 0 && kAStages <= kMaxAStages, "the A ring is handed over in two halves");
  static_assert(kChunkGroups == 4, "the issuer spells out the two k8 steps of a chunk");
  // Hand-over granularity between the converters and the MMA issuer is HALF of the TMEM ring (kABatch chunks), in both
  // directions.  Measured on B200 (tools/mma_bench2.cu): a satisfied mbarrier wait costs the issuing thread ~130 cycles
  // and a tcgen05.commit ~200, while the six MMAs of one chunk execute in 192 cycles and the tensor pipe's queue is
  // shallow -- per-chunk waits/commits leave the pipe idle most of the time.
  constexpr int kABatch = kAStages / 2;
  extern __shared__ __align__(128) unsigned char smem[];
  const SmemLayout L = smem_layout(kg);
  Barriers* bars = reinterpret_cast<Barriers*>(smem + L.bars);
  float* s_pose = reinterpret_cast<float*>(smem + L.pose);
  // broadcast through a shuffle so the compiler treats the warp index (and every role branch on it) as warp-uniform:
  // the MMA issuer's addresses and descriptors then live in uniform registers instead of being moved there per use
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int b0 = blockIdx.y * kN;
  const int nchunks = (kg + kChunkGroups - 1) / kChunkGroups;
  const uint32_t per_tile = 3u * (uint32_t)nchunks;
  const uint32_t my_tiles = (blockIdx.x < (unsigned)ntiles) ? ((uint32_t)(ntiles - 1 - blockIdx.x) / gridDim.x + 1u) : 0u;
  const uint32_t total = my_tiles * per_tile;                              // chunks this CTA processes
  const uint32_t total_padded = (total + kABatch - 1) / kABatch * kABatch;
#ifdef FR_TC_INSTRUMENT
  const bool dbg = debug != 0 && blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0;
#else
  (void)debug;
#endif

  // ---- one-time setup: barriers, TMEM, poses (the resident B operand arrives by bulk copy)
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRawStages; ++i) {
      mbar_init(&bars->raw_full[i], 1);
      mbar_init(&bars->raw_empty[i], 4);       // one arrival per converter warp of the owning group
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->a_full[i], 4 * kABatch);  // 4 converter warps per chunk
      mbar_init(&bars->a_empty[i], 1);
      mbar_init(&bars->d_full[i], 1);
      mbar_init(&bars->d_empty[i], kEpiWarps);
    }
    mbar_init(&bars->b_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kProducerWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < kN * kPoseStride; i += kThreads) s_pose[i] = pose[(size_t)b0 * kPoseStride + i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == kProducerWarp) {
    // ================================================================== producer (TMA engine)
    if (lane == 0) {
      // resident B operand: the pre-split coefficients of this batch tile, already in the canonical K-major layout
      const uint32_t bbytes = 2u * (kN / 8) * L.sbo;
      mbar_arrive_expect_tx(&bars->b_full, bbytes);
      bulk_load(smem + L.b_hi, bsplit + (size_t)blockIdx.y * bbytes, bbytes / 2, &bars->b_full);
      bulk_load(smem + L.b_lo, bsplit + (size_t)blockIdx.y * bbytes + bbytes / 2, bbytes / 2, &bars->b_full);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int c = 0; c < 3; ++c) {
          const float4* src = packed + ((size_t)(tile * 3 + c) * kg) * kTileVerts;
          for (int ci = 0; ci < nchunks; ++ci, ++it) {
            const uint32_t s = it % kRawStages, ph = (it / kRawStages) & 1u;
            const int ng = min(kChunkGroups, kg - ci * kChunkGroups);
            const uint32_t bytes = (uint32_t)ng * kTileVerts * 16u;
            mbar_wait(&bars->raw_empty[s], ph ^ 1u);
            mbar_arrive_expect_tx(&bars->raw_full[s], bytes);
            bulk_load(smem + L.raw + s * kChunkBytes, src + (size_t)ci * kChunkGroups * kTileVerts, bytes, &bars->raw_full[s]);
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ================================================================== MMA issuer (whole warp converged, one elected lane issues)
    // The tensor pipe's queue is shallow and an M128 N64 K8 MMA executes in 32 cycles, so every instruction this warp
    // spends between two MMAs shows up as an idle pipe: the loop nest is the static tile / coordinate / chunk order, the
    // only running state is the chunk counter `it` (ring stage = it % kAStages), and the operands are plain sums of
    // warp-uniform values.
    mbar_wait(&bars->b_full, 0);
    const uint64_t dhi0 = make_b_desc(smem_u32(smem + L.b_hi), 128u, L.sbo);
    const uint64_t dlo0 = make_b_desc(smem_u32(smem + L.b_lo), 128u, L.sbo);
    const int last_nk8 = (kg - (nchunks - 1) * kChunkGroups) / 2;
    const uint32_t a_ring = tmem + kACol0;
    uint32_t it = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      const uint32_t dbuf = tcount % DBUFS;
      TC_T(t0);
      mbar_wait(&bars->d_empty[dbuf], ((tcount / DBUFS) & 1u) ^ 1u);          // epilogue has drained this accumulator set
      TC_T(t1);
      TC_ACC(1, 0, t0, t1);
      tc_fence_after();
#pragma unroll 1
      for (uint32_t c = 0; c < 3u; ++c) {
        const uint32_t d_addr = tmem + dbuf * kDCols + c * kN;
#pragma unroll 1
        for (uint32_t ci = 0; ci < (uint32_t)nchunks; ++ci, ++it) {
          const uint32_t s = it % kAStages, h = s / kABatch;
          if (s % kABatch == 0u) {
            TC_T(t2);
            mbar_wait(&bars->a_full[h], (it / kAStages) & 1u);                 // the converters have filled this half of the A ring
            TC_T(t3);
            TC_ACC(1, 1, t2, t3);
            tc_fence_after();
          }
          if (elect_one()) {
            const uint32_t a_hi = a_ring + s * (2 * kChunkK), a_lo = a_hi + kChunkK;
            const uint64_t koff = (uint64_t)(ci * ((kChunkGroups / 2) * 16u));  // 256 bytes per k8 step, in 16-byte units
            const uint64_t dhi = dhi0 + koff, dlo = dlo0 + koff;
            mma_tf32_ts(d_addr, a_lo, dhi, ci != 0u);
            mma_tf32_ts(d_addr, a_hi, dlo, true);
            mma_tf32_ts(d_addr, a_hi, dhi, true);
            if (ci != (uint32_t)nchunks - 1u || last_nk8 == 2) {
              mma_tf32_ts(d_addr, a_lo + 8, dhi + 16u, true);
              mma_tf32_ts(d_addr, a_hi + 8, dlo + 16u, true);
              mma_tf32_ts(d_addr, a_hi + 8, dhi + 16u, true);
            }
            if (s % kABatch == kABatch - 1u) tc_commit(&bars->a_empty[h]);     // this half of the A ring is reusable
            if (c == 2u && ci == (uint32_t)nchunks - 1u) tc_commit(&bars->d_full[dbuf]);   // all three accumulators complete
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < kConvWarps) {
    // ================================================================== converters (two groups alternate chunks)
    const int group = warp >> 2;
    const int row = (warp & 3) * 32 + lane;                       // TMEM lane == vertex row of the tile
    const uint32_t lane_field = (uint32_t)((warp & 3) * 32) << 16;
    for (uint32_t it = (uint32_t)group; it < total_padded; it += 2u) {
      const uint32_t as = it % kAStages, half = as / kABatch, aph = (it / kAStages) & 1u;
      if (it >= total) {                                          // pad the last half-ring batch so the issuer's wait completes
        mbar_wait(&bars->a_empty[half], aph ^ 1u);                // (in order: never ahead of the previous round's phase)
        if (lane == 0) mbar_arrive(&bars->a_full[half]);
        continue;
      }
      const uint32_t s = it % kRawStages, ph = (it / kRawStages) & 1u;
      const uint32_t ci = (it % per_tile) % (uint32_t)nchunks;
      const int ng = min(kChunkGroups, kg - (int)ci * kChunkGroups);
      TC_T(c0);
      mbar_wait(&bars->raw_full[s], ph);
      TC_T(c1);
      const float4* src = reinterpret_cast<const float4*>(smem + L.raw + s * kChunkBytes) + row;
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int g = 0; g < kChunkGroups; ++g) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g < ng) v = src[g * kTileVerts];
        const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t hbits = __float_as_uint(f[e]) & 0xFFFFE000u;
          hi[4 * g + e] = hbits;
          lo[4 * g + e] = __float_as_uint(f[e] - __uint_as_float(hbits));
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->raw_empty[s]);            // this warp's rows are in registers
      TC_T(c2);
      mbar_wait(&bars->a_empty[half], aph ^ 1u);                  // MMAs that read this half of the TMEM ring have retired
      TC_T(c3);
      tc_fence_after();
      const uint32_t a_addr = tmem + lane_field + kACol0 + as * (2 * kChunkK);
      tmem_st16(a_addr, hi);
      tmem_st16(a_addr + kChunkK, lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      TC_T(c4);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->a_full[half]);
      if (warp == 0) {
        TC_ACC(0, 0, c0, c1);   // wait raw_full
        TC_ACC(0, 1, c1, c2);   // LDS + split
        TC_ACC(0, 2, c2, c3);   // wait a_empty
        TC_ACC(0, 3, c3, c4);   // STTM + wait::st
#ifdef FR_TC_INSTRUMENT
        if (dbg) g_tc_dbg[0][7] += 1;
#endif
      }
    }
  } else {
    // ================================================================== epilogue (warps 8..11)
    const int qd = (warp - kEpiWarp0) & 3;                        // TMEM lane quarter == warp % 4
    const int jb0 = ((warp - kEpiWarp0) >> 2) * (kN / (kEpiWarps / 4)), jb1 = jb0 + kN / (kEpiWarps / 4);   // this warp's faces
    const int v = qd * 32 + lane;
    const uint32_t lane_field = (uint32_t)(qd * 32) << 16;
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      const uint32_t dbuf = tcount % DBUFS, dph = (tcount / DBUFS) & 1u;
      const int n = tile * kTileVerts + v;
      TC_T(e0);
      mbar_wait(&bars->d_full[dbuf], dph);
      TC_T(e1);
      tc_fence_after();
      const uint32_t d_addr = tmem + lane_field + dbuf * kDCols;
#pragma unroll 1
      for (int jb = jb0; jb < jb1; jb += 16) {
        float x[16], y[16], z[16];
        tmem_ld16(d_addr + 0 * kN + jb, x);
        tmem_ld16(d_addr + 1 * kN + jb, y);
        tmem_ld16(d_addr + 2 * kN + jb, z);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (n < nver) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int b = b0 + jb + j;
            if (b < batch) {
              const float4* pp = reinterpret_cast<const float4*>(s_pose + (jb + j) * kPoseStride);
              const float4 p0 = pp[0], p1 = pp[1], p2 = pp[2];
              const float P[12] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w, p2.x, p2.y, p2.z, p2.w};
              project_store(P, x[j], y[j], z[j], im_size, flags, out, b, nver, n);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->d_empty[dbuf]);
      if (warp == kEpiWarp0) {
        TC_T(e2);
        TC_ACC(2, 0, e0, e1);   // wait d_full
        TC_ACC(2, 1, e1, e2);   // drain + project + store
      }
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == kProducerWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)kTmemCols) : "memory");
  }
}

}  // namespace tc

// split coefficients of every 64-face batch tile in the canonical layout: [tile][hi|lo][8 groups][kg][8 faces][4 k]
inline size_t recon_tc_workspace_bytes(int batch, const BasisGeom& g) {
  return (size_t)ceil_div(batch, tc::kN) * 2 * (tc::kN / 8) * (size_t)g.kg * 128;
}

inline int env_int(const char* name, int dflt) {
  const char* e = std::getenv(name);
  return e ? std::atoi(e) : dflt;
}
// FR_PDL=0 launches every kernel with full stream serialisation (A/B switch for the programmatic dependent launches)
inline bool pdl_enabled() {
  static const int v = env_int("FR_PDL", 1);
  return v != 0;
}

// FR_RECON_PATH = simt | tf32 | f16 overrides the forward dispatch (debugging / A-B comparisons): 1, 2, 3; 0 = default.
inline int recon_path_override() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = std::getenv("FR_RECON_PATH");
    cached = (e == nullptr) ? 0 : (e[0] == 's' ? 1 : (e[0] == 't' ? 2 : (e[0] == 'f' ? 3 : 0)));
  }
  return cached;
}

inline bool recon_tc_applicable(int, const BasisGeom& g, unsigned) {
  return tc::smem_layout(g.kg).total <= 227u * 1024u;   // K too large for a resident coefficient operand otherwise
}

inline int launch_recon_fwd_tc(const float* packed, const float* coefT, const float* pose, void* tc_ws, ReconOut out,
                               int batch, int nver, const BasisGeom& g, float im_size, unsigned flags, int nsm,
                               cudaStream_t st) {
  static const int dbufs = env_int("FR_TC_DBUFS", 2);
  static const int debug = env_int("FR_TC_DEBUG", 0);
  const tc::SmemLayout L = tc::smem_layout(g.kg);
  unsigned char* bsplit = static_cast<unsigned char*>(tc_ws);   // filled by recon_prep_kernel
  static_assert(tc::kN == kBatchPad, "the prep kernel lays the split coefficients out per kBatchPad faces");
  const int nbt = ceil_div(batch, tc::kN);
  int ctas = nsm / nbt;
  if (ctas < 1) ctas = 1;
  if (ctas > g.ntiles) ctas = g.ntiles;
  if (dbufs == 1) {
    FR_CUDA(cudaFuncSetAttribute(tc::recon_fwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    tc::recon_fwd_tc_kernel<1><<<dim3(ctas, nbt), tc::kThreads, L.total, st>>>(
        reinterpret_cast<const float4*>(packed), bsplit, pose, out, batch, nver, g.kg, g.ntiles, im_size, flags, debug);
  } else {
    FR_CUDA(cudaFuncSetAttribute(tc::recon_fwd_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    tc::recon_fwd_tc_kernel<2><<<dim3(ctas, nbt), tc::kThreads, L.total, st>>>(
        reinterpret_cast<const float4*>(packed), bsplit, pose, out, batch, nver, g.kg, g.ntiles, im_size, flags, debug);
  }
  FR_LAUNCHED("recon_fwd_tc_kernel");
  return FR_OK;
}

}  // namespace fr

#endif  // FR_RECON_TC_CUH_
