// Developer microbenchmark 2 (GPU): tcgen05.mma kind::tf32 TS execution floor with a lean, warp-converged issuer
// (whole warp in the role, elect.sync picks the issuing lane, operands precomputed, 16 MMAs unrolled per loop trip).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 20); ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void mma(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// NACC: number of independent accumulators cycled through; per_commit in units of 16-MMA groups (0 = only at the end)
template <int N, int NACC>
__global__ void __launch_bounds__(128, 1) k(int groups, int groups_per_commit, long long* out, int sbo, int pattern, int bg) {
  __shared__ volatile int stop_flag;
  if (threadIdx.x == 0) stop_flag = 0;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tbase;
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  if (warp == 1) {  // whole warp converged in the role
    const uint32_t b_smem = smem_u32(smem);
    const uint64_t bdesc = (uint64_t)((b_smem & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
    const uint64_t bdesc2 = bdesc + (uint64_t)(pattern ? (59392 >> 4) : 0);
    const long long t0 = clock64();
    uint32_t phase = 0, commits = 0;
    for (int g = 0; g < groups; ++g) {
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          mma(tmem + (uint32_t)((j % NACC) * N), tmem + 448u + 8u * (j & 7), ((j & 1) ? bdesc2 : bdesc) + (uint64_t)(16 * (j & 3)), idesc, (g | (j / NACC)) != 0);
      }
      __syncwarp();
      if (groups_per_commit > 0 && (g + 1) % groups_per_commit == 0 && g + 1 < groups) {
        if (elect_one())
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        __syncwarp();
        ++commits;
      }
    }
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    __syncwarp();
    phase = commits & 1u;
    mbar_wait(&bar, phase);
    if (threadIdx.x == 32) { out[0] = clock64() - t0; stop_flag = 1; }
  } else if (bg && warp >= 2) {
    // background traffic like the recon kernel's other warps: bg&1 = TMEM stores into the A ring, bg&2 = TMEM loads of D,
    // bg&4 = shared-memory reads
    uint32_t v[16];
    for (int i = 0; i < 16; ++i) v[i] = threadIdx.x + i;
    const uint32_t lanef = ((uint32_t)(warp & 3) * 32u) << 16;
    float acc = 0.f;
    while (!stop_flag) {
      if (bg & 1) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(tmem + lanef + 384u),
          "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
          "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      if (bg & 2) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
            "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(tmem + lanef + 256u) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        v[0] ^= r[3];
      }
      if (bg & 4) {
        for (int i = 0; i < 8; ++i) acc += reinterpret_cast<volatile float*>(smem)[(threadIdx.x * 4 + i * 512) & 16383];
      }
    }
    if (acc == 123.f) out[1] = v[0];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

__global__ void __launch_bounds__(128, 1) kloop(int chunks, int variant, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar, done_bar;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&done_bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tbase;
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((128u >> 4) << 24);
  if (warp == 1) {
    const uint32_t b_smem = smem_u32(smem);
    const uint64_t dhi0 = (uint64_t)((b_smem & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(7424 >> 4) << 32) | (1ull << 46);
    const uint64_t dlo0 = dhi0 + (59392 >> 4);
    const long long t0 = clock64();
    for (int it = 0; it < chunks; ++it) {
      const uint32_t as = it % 4;
      if (variant & 1) mbar_wait(&done_bar, 1);                  // already-complete phase: the cost of a satisfied wait
      if (variant & 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t a_hi = tmem + 384u + as * 32u, a_lo = a_hi + 16u;
        const uint32_t d_addr = tmem + (uint32_t)((it / 15) % 3) * 64u;
        const uint64_t koff = (uint64_t)((it % 15) * 2) * 16u;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint64_t dhi = dhi0 + koff + 16u * j, dlo = dlo0 + koff + 16u * j;
          mma(d_addr, a_lo + 8 * j, dhi, idesc, (it | j) != 0);
          mma(d_addr, a_hi + 8 * j, dlo, idesc, 1);
          mma(d_addr, a_hi + 8 * j, dhi, idesc, 1);
        }
        if ((variant & 4) && (as & 1)) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
      __syncwarp();
    }
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done_bar)) : "memory");
    __syncwarp();
    mbar_wait(&done_bar, 0);
    if (threadIdx.x == 32) out[0] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

void runloop(int variant) {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(kloop, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) kloop<<<1, 128, 200 * 1024>>>(1350, variant, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("recon-style issue loop, variant %d (1=satisfied wait, 2=fence, 4=commit every 2 chunks): %7.1f cycles per chunk of 6 MMAs (%s)\n",
         variant, (double)h / 1350.0, cudaGetErrorString(e));
  fflush(stdout);
  cudaFree(d);
}

template <int N, int NACC>
void run(int groups, int gpc, int sbo = 256, int pattern = 0, int bg = 0) {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k<N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) k<N, NACC><<<1, 128, 200 * 1024>>>(groups, gpc, d, sbo, pattern, bg);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("bg %d N=%3d  %d acc  sbo %5d pattern %d commit every %3d MMAs: %7.1f cycles per MMA  (%s)\n", bg, N, NACC, sbo, pattern, gpc * 16,
         (double)h / ((double)groups * 16), cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int v : {0, 1, 2, 3, 4, 7}) runloop(v);
  for (int bg : {0, 1, 2, 4, 7}) run<64, 3>(400, 0, 7424, 1, bg);
  run<64, 1>(200, 0, 7424, 0);
  run<64, 1>(200, 0, 7424, 1);
  run<64, 1>(200, 0, 1024, 0);
  run<64, 1>(200, 0, 512, 0);
  run<64, 1>(200, 0, 7424 + 128, 0);
  run<64, 1>(200, 0, 7424 - 128, 0);
  run<128, 1>(200, 0, 7424, 0);
  run<64, 1>(200, 0);
  run<64, 2>(200, 0);
  run<64, 3>(200, 0);
  run<64, 4>(200, 0);
  run<128, 1>(200, 0);
  run<128, 2>(200, 0);
  run<128, 3>(200, 0);
  run<256, 1>(200, 0);
  run<64, 1>(200, 1);
  run<64, 3>(200, 1);
  run<64, 3>(200, 2);
  run<128, 3>(200, 1);
  run<128, 3>(200, 2);
  return 0;
}
