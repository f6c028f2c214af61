// Developer microbenchmark (GPU): streaming bandwidth of cp.async.bulk (1-D TMA) global -> shared with an mbarrier ring,
// one producer lane + 4 consumer warps per CTA (consumers only touch the data lightly), 148 CTAs, cold data (> L2).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}
template <int STAGES, int CHUNK>
__global__ void __launch_bounds__(576, 1) k(const unsigned char* src, size_t bytes_per_cta, float* sink, int prefetch, int spinners, int backoff, int hold = 0) {
  __shared__ uint64_t never;
  __shared__ volatile int stop;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[STAGES], empty[STAGES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[i])), "r"(1) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[i])), "r"(4) : "memory");
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&never)), "r"(1) : "memory");
    stop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp >= 5) {   // extra warps polling a barrier that completes only at the end (like idle roles of the recon kernel)
    if (warp - 5 < spinners) {
      uint32_t done = 0;
      while (!done && !stop) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&never)), "r"(0u) : "memory");
        if (backoff) __nanosleep(backoff);
      }
    }
    return;
  }
  const unsigned char* base = src + (size_t)blockIdx.x * bytes_per_cta;
  const uint32_t n = (uint32_t)(bytes_per_cta / CHUNK);
  if (warp == 4) {
    if (lane == 0) {
      for (uint32_t it = 0; it < n; ++it) {
        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
        if (prefetch && it + prefetch < n)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + (size_t)(it + prefetch) * CHUNK), "r"((uint32_t)CHUNK) : "memory");
        mbar_wait(&empty[s], ph ^ 1u);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"((uint32_t)CHUNK) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + s * CHUNK)),
                     "l"(base + (size_t)it * CHUNK), "r"((uint32_t)CHUNK), "r"(smem_u32(&full[s])) : "memory");
      }
    }
  } else {
    float acc = 0.f;
    for (uint32_t it = 0; it < n; ++it) {
      const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
      mbar_wait(&full[s], ph);
      acc += reinterpret_cast<const float*>(smem + s * CHUNK)[threadIdx.x];
      if (hold) { const long long t = clock64(); while (clock64() - t < hold) {} }   // consumer keeps the stage for a while
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
    }
    if (acc == 12345.f) sink[0] = acc;
    __syncwarp();
    if (warp == 0 && lane == 0) stop = 1;
  }
}
template <int STAGES, int CHUNK>
void run(const unsigned char* d, size_t total, float* sink, int prefetch, int spinners = 0, int backoff = 0, int hold = 0) {
  const int ctas = 148;
  const size_t per = total / ctas / CHUNK * CHUNK;
  cudaFuncSetAttribute(k<STAGES, CHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * CHUNK);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e9f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a);
    k<STAGES, CHUNK><<<ctas, 160 + 32 * spinners, STAGES * CHUNK>>>(d, per, sink, prefetch, spinners, backoff, hold);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  printf("hold %4d spinners %2d backoff %4d stages %2d x %5d B  L2 prefetch %2d: %7.1f us  %6.0f GB/s   (%s)\n", hold, spinners, backoff, STAGES, CHUNK, prefetch, best * 1e3,
         (double)per * ctas / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
  fflush(stdout);
}
int main() {
  const size_t total = (size_t)1200 << 20;  // 1.2 GB >> L2
  unsigned char* d;
  float* sink;
  cudaMalloc(&d, total);
  cudaMalloc(&sink, 4);
  cudaMemset(d, 1, total);
  for (int hold : {0, 300, 600, 1000, 1500}) run<6, 16384>(d, (size_t)148 << 20, sink, 0, 0, 0, hold);
  for (int hold : {300, 600, 1000}) run<12, 8192>(d, (size_t)148 << 20, sink, 0, 0, 0, hold);
  for (size_t mb : {148}) {
    printf("-- %zu MB total\n", mb);
    run<6, 16384>(d, mb << 20, sink, 0, 0);
    run<3, 32768>(d, mb << 20, sink, 0, 0);
    run<12, 8192>(d, mb << 20, sink, 0, 0);
  }
  run<6, 16384>(d, total, sink, 0, 4);
  run<6, 16384>(d, total, sink, 0, 13);
  run<6, 16384>(d, total, sink, 0, 13, 200);
  run<6, 16384>(d, total, sink, 0, 13, 1000);
  run<8, 8192>(d, total, sink, 0);
  run<12, 8192>(d, total, sink, 0);
  run<16, 8192>(d, total, sink, 0);
  run<24, 8192>(d, total, sink, 0);
  run<12, 8192>(d, total, sink, 24);
  run<12, 8192>(d, total, sink, 48);
  run<6, 16384>(d, total, sink, 0);
  run<12, 16384>(d, total, sink, 0);
  run<24, 4096>(d, total, sink, 0);
  return 0;
}
