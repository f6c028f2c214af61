// Developer microbenchmark 3 (GPU): tcgen05.mma with BOTH operands from shared memory (SS form), M128 N64, for
// kind::tf32 (K8) and kind::f16 (K16, fp16 inputs), lean warp-converged issuer, 16 MMAs unrolled per elect.
// Question answered: does the A-operand read (4 KB per MMA) hold the 32-cycle floor at N = 64, alone and with
// concurrent LDS / bulk-copy traffic into the same shared memory?
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
  for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}
template <int KIND>  // 0 = tf32, 1 = f16
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

template <int KIND, int N>
__global__ void __launch_bounds__(256, 1) k(int groups, long long* out, int bg, const float4* gsrc) {
  __shared__ volatile int stop_flag;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar, cpbar;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    stop_flag = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&cpbar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 128 * 1024 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0u;   // zeros: valid in every format
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tbase;
  // idesc: c=F32 (1<<4); tf32: a,b format 2 at [7,10),[10,13); f16: format 0 (fp16); N>>3 at 17, M>>4 at 24
  constexpr uint32_t idesc = (1u << 4) | ((KIND == 0 ? 2u : 0u) << 7) | ((KIND == 0 ? 2u : 0u) << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  if (warp == 1) {
    // A: 8 KB chunks [k-group][128 rows][16 B] (LBO = 2048 between K core matrices, SBO = 128 between 8-row groups) at smem+0 .. 64 KB
    // B: resident [8-row groups][K core matrices] (LBO = 128, SBO = 7424) at smem + 64 KB
    const uint32_t a_smem = smem_u32(smem), b_smem = smem_u32(smem + 64 * 1024);
    const uint64_t adesc = desc(a_smem, 2048, 128), bdesc = desc(b_smem, 128, 7424);
    const long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 16; ++j)   // walk over 8 raw chunks (2 K steps each) and the matching B columns
          mma_ss<KIND>(tmem + (uint32_t)((j % 3) * N), adesc + (uint64_t)((j >> 1) * (8192 >> 4) + (j & 1) * (4096 >> 4)),
                       bdesc + (uint64_t)(16 * (j & 7)), idesc, (g | (j / 3)) != 0);
      }
      __syncwarp();
    }
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    __syncwarp();
    mbar_wait(&bar, 0);
    if (threadIdx.x == 32) { out[0] = clock64() - t0; stop_flag = 1; }
  } else if ((bg & 4) && warp >= 4) {          // LDS traffic: 4 warps reading 16 B per lane, like the converters
    float acc = 0.f;
    while (!stop_flag) {
      for (int i = 0; i < 8; ++i) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "r"(smem_u32(smem) + (uint32_t)(((threadIdx.x + i * 128) & 4095) * 16)) : "memory");
        acc += v.x + v.w;
      }
    }
    if (acc == 123.f) out[1] = 1;
  } else if ((bg & 8) && warp == 2) {          // bulk-copy traffic: 8 KB copies global -> smem (upper 32 KB of the A region is not read by the MMAs here)
    uint32_t ph = 0;
    while (!stop_flag) {
      if (elect_one()) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&cpbar)), "r"(8192u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + 96 * 1024)),
                     "l"(gsrc + (size_t)((ph * 512u) & 0xFFFFFu)), "r"(8192u), "r"(smem_u32(&cpbar)) : "memory");
      }
      __syncwarp();
      mbar_wait(&cpbar, ph & 1u);
      ++ph;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int KIND, int N>
void run(int bg, const float4* gsrc) {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k<KIND, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int groups = 400;
  for (int rep = 0; rep < 2; ++rep) k<KIND, N><<<1, 256, 200 * 1024>>>(groups, d, bg, gsrc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("SS %s N=%3d bg=%d (4 = LDS traffic, 8 = bulk copies): %7.1f cycles per MMA  (%s)\n", KIND == 0 ? "tf32 K8 " : "f16  K16", N, bg,
         (double)h / ((double)groups * 16), cudaGetErrorString(e));
  fflush(stdout);
  cudaFree(d);
}

int main() {
  float4* g;
  cudaMalloc(&g, 32 << 20);
  cudaMemset(g, 0, 32 << 20);
  for (int bg : {0, 4, 8, 12}) run<0, 64>(bg, g);
  for (int bg : {0, 4, 8, 12}) run<1, 64>(bg, g);
  run<0, 128>(0, g);
  run<1, 128>(0, g);
  return 0;
}
