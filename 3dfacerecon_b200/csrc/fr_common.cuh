// Shared host-side plumbing for the C ABI (error reporting, launch accounting) and small device helpers.
#ifndef FR_COMMON_CUH_
#define FR_COMMON_CUH_

#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/facerecon_b200.h"

namespace fr {

// ---- errors: status code + thread-local message (replaces the reference's printf-and-return,
// render_depth_op.cc:161-172 / render_depth_op.cu.cc:290-295)
inline char* error_buffer() {
  static thread_local char buf[512] = "";
  return buf;
}
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

inline std::atomic<unsigned long long>& launch_counter() {
  static std::atomic<unsigned long long> n(0);
  return n;
}

#define FR_REQUIRE(cond, ...) \
  do {                        \
    if (!(cond)) return ::fr::fail(FR_ERR_INVALID_ARGUMENT, __VA_ARGS__); \
  } while (0)

#define FR_CUDA(call)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::fr::fail(FR_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// Call after every kernel launch: counts it and surfaces launch-configuration errors.
#define FR_LAUNCHED(name)                                                                          \
  do {                                                                                             \
    ::fr::launch_counter().fetch_add(1, std::memory_order_relaxed);                                \
    cudaError_t _e = cudaGetLastError();                                                           \
    if (_e != cudaSuccess) return ::fr::fail(FR_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- packed-basis geometry (see DESIGN.md "Packed basis")
constexpr int kTileVerts = 128;  // vertices per tile == TMEM lanes == threads of one SIMT row block

struct BasisGeom {
  int nver, ks, ke;
  int kreal;   // ks + ke + 1 (the extra column is the mean, coefficient 1)
  int kpad;    // kreal rounded up to 8 (fp32 section: float4 groups)
  int kg;      // kpad / 4 float4 groups
  int kpad16;  // kreal rounded up to 16 (fp16-pair sections: one f16 MMA K step per chunk)
  int nch16;   // kpad16 / 16 chunks per (tile, coordinate) row
  int ntiles;  // ceil(nver / 128): tiles of consecutive vertices (fp32 section, backward section, mean)
  int nclusters;  // > 0: the packed basis carries a SECOND tensor-core forward section whose row tiles are the mesh table's
                  // clusters (FR_CLUSTER_TILES; the fused call rasterizes from it); 0: none
  // Sections of the packed basis (DESIGN.md "Packed basis"), in this order; only the last one depends on the mesh table:
  //   [0, f32_bytes)                  fp32, float4-tiled: SIMT forward (small batches) and SIMT backward
  //   [scale_offset, +4 kpad16)       2^-s_k per column (what a coefficient is multiplied by to undo the column scale)
  //   [bwd_offset, +bwd_bytes)        fp16 hi/lo pairs of the column-scaled basis, transposed for the backward contraction over
  //                                   the vertices: per (tile, coordinate, 16-vertex chunk)  [hi m0 | hi m1 | lo m0 | lo m1]  with 128 k rows each
  //   [mean_offset, +mean_bytes)      the mean once more as plain fp32 [3][ntiles*128] (the backward contracts it in fp32)
  //   [f16_offset, +f16_bytes)        fp16 hi/lo pairs in tcgen05 operand tiles of 128 consecutive vertices (vertex RANKS when
  //                                   packed with a mesh table): tensor-core forward writing vertex_proj / vertex records
  //   [f16c_offset, +f16c_bytes)      the same operand tiles once more, one 128-row tile per CLUSTER of the mesh table (border
  //                                   vertices repeated in every member cluster): tensor-core forward with the tile
  //                                   rasterizer in its epilogue (only with FR_CLUSTER_TILES)
  size_t tile_floats() const { return (size_t)3 * kg * kTileVerts * 4; }
  size_t f32_bytes() const { return (size_t)ntiles * tile_floats() * sizeof(float); }
  size_t scale_offset() const { return (f32_bytes() + 1023) / 1024 * 1024; }
  int mtiles() const { return (kpad16 + 127) / 128; }
  size_t bwd_offset() const { return scale_offset() + ((size_t)kpad16 * sizeof(float) + 1023) / 1024 * 1024; }
  size_t bwd_bytes() const { return (size_t)ntiles * 3 * (kTileVerts / 16) * 2 * mtiles() * 4096; }
  size_t mean_offset() const { return bwd_offset() + bwd_bytes(); }
  size_t mean_bytes() const { return (size_t)3 * ntiles * kTileVerts * sizeof(float); }
  size_t f16_offset() const { return (mean_offset() + mean_bytes() + 1023) / 1024 * 1024; }
  size_t f16_bytes() const { return (size_t)ntiles * 3 * nch16 * 8192; }
  size_t f16c_offset() const { return f16_offset() + f16_bytes(); }
  size_t f16c_bytes() const { return (size_t)nclusters * 3 * nch16 * 8192; }
  size_t bytes() const { return f16c_offset() + f16c_bytes(); }
};
inline BasisGeom basis_geom(int nver, int ks, int ke, int nclusters = 0) {
  BasisGeom g;
  g.nver = nver;
  g.ks = ks;
  g.ke = ke;
  g.kreal = ks + ke + 1;
  g.kpad = (g.kreal + 7) / 8 * 8;
  g.kg = g.kpad / 4;
  g.kpad16 = (g.kreal + 15) / 16 * 16;
  g.nch16 = g.kpad16 / 16;
  g.ntiles = (nver + kTileVerts - 1) / kTileVerts;
  g.nclusters = nclusters > 0 ? nclusters : 0;
  return g;
}

#ifdef __CUDACC__
#ifdef FR_TIMELINE   // developer build (tools/timeline.py): globaltimer marks of the kernels of one fused step
// [0] prep first start  [1] prep last end  [2] forward first CTA entry  [3] forward last CTA done  [4] resolve first CTA past its
// dependency wait  [5] resolve last end  [6] resolve first CTA entry
__device__ unsigned long long g_marks[8];
__device__ __forceinline__ unsigned long long fr_gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define FR_MARK_MIN(i) do { if (threadIdx.x == 0) atomicMin(&g_marks[i], fr_gtime()); } while (0)
#define FR_MARK_MAX(i) do { if (threadIdx.x == 0) atomicMax(&g_marks[i], fr_gtime()); } while (0)
#else
#define FR_MARK_MIN(i) do {} while (0)
#define FR_MARK_MAX(i) do {} while (0)
#endif
// Programmatic dependent launch (PDL): a kernel launched with launch_pdl() may become resident while its predecessor in
// the stream is still running; it must call pdl_wait() before touching anything the predecessor writes (and before
// writing anything the predecessor reads).  The predecessor calls pdl_trigger() once its blocks are resident.  Both are
// no-ops for normally launched kernels.  This hides the launch / drain gap between the kernels of one step.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool dependent,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = dependent ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// 128-bit streaming load through the read-only path without polluting L1 (basis is read once per CTA).
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float f4_get(const float4& v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); }
#endif

}  // namespace fr
#endif  // FR_COMMON_CUH_
