// Developer microbenchmark (GPU): issue-to-completion cost of tcgen05.mma kind::tf32 / kind::f16 variants on one CTA,
// to size the tensor-core reconstruction kernel (operand sources, smem layouts, N).  Data is garbage; only timing matters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_bench tools/mma_bench.cu && tools/mma_bench
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}

// mode bits: 0 = A from TMEM (TS) else smem (SS); 1 = B swizzle-128B else none; 2 = kind::f16 (bf16) else tf32
template <int MODE>
__global__ void __launch_bounds__(128, 1) k(int n, int iters, int per_commit, long long* out, int nowait = 0, int sttm = 0, int nacc = 1) {
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
  constexpr bool kTS = (MODE & 1) != 0, kSW = (MODE & 2) != 0, kF16 = (MODE & 4) != 0;
  const uint32_t idesc = (1u << 4) | ((kF16 ? 1u : 2u) << 7) | ((kF16 ? 1u : 2u) << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
  if (threadIdx.x == 0) {
    const uint32_t a_smem = smem_u32(smem), b_smem = smem_u32(smem + 32768);
    uint64_t bdesc, adesc;
    if (kSW) {  // 128-byte rows, 8-row groups 1024 bytes apart, SWIZZLE_128B (layout_type 2)
      bdesc = (uint64_t)((b_smem & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
      adesc = (uint64_t)((a_smem & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
    } else {    // no swizzle: core matrices 128 B apart along K, 8-row groups 7424 B apart (what the recon kernel uses)
      bdesc = (uint64_t)((b_smem & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
      adesc = (uint64_t)((a_smem & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
    }
    const long long t0 = clock64();
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      for (int j = 0; j < per_commit; ++j) {
        const uint32_t acc = (it | j) != 0;
        if (kTS) {
          if (kF16)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                         ::"r"(tmem), "r"(tmem + 256u + 8u * (j & 7)), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
          else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                         ::"r"(tmem + (uint32_t)((j % nacc) * 64)), "r"(tmem + 256u + 8u * (j & 7)), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
        } else {
          if (kF16)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
          else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      if (!nowait || it == iters - 1) mbar_wait(&bar, phase);
      phase ^= 1u;
    }
    out[0] = clock64() - t0;
    *(volatile int*)(&tbase) = 0x7fffffff;   // tell the STTM warps to stop
  } else if (sttm && warp >= 1) {
    // background TMEM store traffic like the converter warps of the recon kernel (columns 384.., own lane quarter)
    uint32_t v[16];
    for (int i = 0; i < 16; ++i) v[i] = threadIdx.x + i;
    const uint32_t addr = tmem + (((uint32_t)(warp & 3) * 32u) << 16) + 384u;
    while (*(volatile int*)(&tbase) != 0x7fffffff) {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(addr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  const uint32_t tmem_keep = tmem;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_keep), "r"(512u) : "memory");
}

template <int MODE>
void run(const char* name, int n, int iters, int per_commit, int nowait = 0, int sttm = 0, int nacc = 1) {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int rep = 0; rep < 2; ++rep) k<MODE><<<1, 128, 64 * 1024>>>(n, iters, per_commit, d, nowait, sttm, nacc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-44s N=%3d  %4d x %3d MMAs/commit: %8.1f cycles per MMA   (%s)\n", name, n, iters, per_commit,
         (double)h / ((double)iters * per_commit), cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int n : {64, 128, 256}) {
    run<1>("tf32 TS (A tmem), B no-swizzle", n, 50, 64);
    run<3>("tf32 TS (A tmem), B swizzle128", n, 50, 64);
    run<0>("tf32 SS (A smem), A/B no-swizzle", n, 50, 64);
    run<2>("tf32 SS (A smem), A/B swizzle128", n, 50, 64);
    run<5>("bf16 TS (A tmem), B no-swizzle", n, 50, 64);
    run<7>("bf16 TS (A tmem), B swizzle128", n, 50, 64);
  }
  run<1>("tf32 TS no-swizzle, commit+wait every 6", 64, 500, 6);
  for (int pc : {1, 2, 3, 4, 6, 12, 24}) run<1>("tf32 TS no-swizzle, commit (no wait)", 64, 3000 / pc, pc, 1);
  for (int pc : {4, 8}) run<1>("tf32 TS no-swizzle N=128, commit (no wait)", 128, 3000 / pc, pc, 1);
  for (int na : {1, 2, 3, 4}) {
    char nm[64];
    snprintf(nm, 64, "tf32 TS N=64, %d independent accumulators", na);
    run<1>(nm, 64, 50, 60, 0, 0, na);
    snprintf(nm, 64, "  ... commit (no wait) every 6, %d acc", na);
    run<1>(nm, 64, 500, 6, 1, 0, na);
    snprintf(nm, 64, "  ... commit (no wait) every 12, %d acc", na);
    run<1>(nm, 64, 250, 12, 1, 0, na);
  }
  run<1>("tf32 TS N=64 + background STTM (3 warps)", 64, 50, 64, 0, 1);
  run<1>("tf32 TS N=128 + background STTM (3 warps)", 128, 50, 64, 0, 1);
  return 0;
}
