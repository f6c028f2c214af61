// 3DMM reconstruction + pose projection kernels (SIMT flavour) for sm_100a.
//
// Replaces FaceRecNet.vertices_transform (nets/network.py:140-171): two cuBLAS SGEMMs + three
// materialised transposes + a host round trip for the rotation matrices (tf.py_func, :150) become
//   recon_prep_kernel        params -> transposed coefficient matrix + per-face pose (f.R, t, R, f)
//   recon_fwd_simt_kernel    one streaming pass over the packed basis, projection + y-flip in the epilogue
// and its autodiff gradient becomes
//   recon_bwd_dt_kernel      d t3d
//   recon_bwd_simt_kernel    G[b,k] = sum_{c,n} P[(c,n),k] * (R_b^T g'_b)[c,n]   (second streaming pass)
//   recon_bwd_finalize_kernel d alpha = f.G,  d f = sum_k coef[k].G[k]
// The tensor-core (tcgen05) flavours live in recon_f16.cuh (forward) and recon_bwd_f16.cuh (backward).
#ifndef FR_RECON_CUH_
#define FR_RECON_CUH_

#include "fr_common.cuh"
#include "raster_core.h"

namespace fr {

constexpr int kPoseStride = 24;  // floats per face: M = f.R [9] | t [3] | R [9] | f | pad[2]
constexpr int kBatchPad = 64;    // coefficient matrix columns are padded to a multiple of this (== faces per tensor-core batch tile)

__host__ __device__ inline int batch_padded(int batch) { return (batch + kBatchPad - 1) / kBatchPad * kBatchPad; }

// ---------------------------------------------------------------------------------------------- packing
// packed float4 index = ((tile*3 + c)*kg + g)*128 + v  holds columns 4g..4g+3 of basis row (c, n = tile*128+v);
// column order [pc_shape | pc_exp | mu | 0...]; rows n >= nver are zero.
__global__ void __launch_bounds__(256)
pack_basis_kernel(const float* __restrict__ mu, const float* __restrict__ pc_shape, const float* __restrict__ pc_exp,
                  int nver, int ks, int ke, int kg, int ntiles, unsigned flags, float4* __restrict__ packed) {
  const size_t total = (size_t)ntiles * 3 * kg * kTileVerts;
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int v = (int)(idx % kTileVerts);
  const int g = (int)((idx / kTileVerts) % kg);
  const int c = (int)((idx / ((size_t)kTileVerts * kg)) % 3);
  const int tile = (int)(idx / ((size_t)kTileVerts * kg * 3));
  const int n = tile * kTileVerts + v;
  float out[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  if (n < nver) {
    const size_t row_b = (flags & FR_BASIS_INTERLEAVED) ? (size_t)3 * n + c : (size_t)c * nver + n;
    const size_t row_m = (flags & FR_MEAN_INTERLEAVED) ? (size_t)3 * n + c : (size_t)c * nver + n;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = 4 * g + j;
      if (k < ks) out[j] = pc_shape[row_b * ks + k];
      else if (k < ks + ke) out[j] = pc_exp[row_b * ke + (k - ks)];
      else if (k == ks + ke) out[j] = mu[row_m];
    }
  }
  packed[idx] = make_float4(out[0], out[1], out[2], out[3]);
}

// ---------------------------------------------------------------------------------------------- prep
// Rotation as the reference builds it: float64 sin/cos, float64 3x3 products, cast to float32
// (nets/network.py:277-290 / rendering_layer/sample_test.py:59-72); then M = f (*) R in float32 (network.py:165).
__device__ inline void pose_matrices_sc(double sp, double cp, double sg, double cg, double st, double ct, const float* __restrict__ p,
                                        unsigned flags, float* __restrict__ out);
__device__ inline void pose_matrices(const float* __restrict__ p, unsigned flags, float* __restrict__ out) {
  double sp, cp, sg, cg, st, ct;
  sincos((double)p[0], &sp, &cp);  // phi   : pitch
  sincos((double)p[1], &sg, &cg);  // gamma : yaw
  sincos((double)p[2], &st, &ct);  // theta : roll
  pose_matrices_sc(sp, cp, sg, cg, st, ct, p, flags, out);
}
// ... from the six sines / cosines (the tensor-core prep kernel evaluates the three float64 sincos on three threads)
__device__ inline void pose_matrices_sc(double sp, double cp, double sg, double cg, double st, double ct, const float* __restrict__ p,
                                        unsigned flags, float* __restrict__ out) {
  const double rx[9] = {1, 0, 0, 0, cp, sp, 0, -sp, cp};
  const double ry[9] = {cg, 0, -sg, 0, 1, 0, sg, 0, cg};
  const double rz[9] = {ct, st, 0, -st, ct, 0, 0, 0, 1};
  double tmp[9], r[9];
  const double* a1 = (flags & FR_ROT_ZYX) ? ry : rx;  // zyx: Rz.(Ry.Rx)   xyz: (Rx.Ry).Rz
  const double* b1 = (flags & FR_ROT_ZYX) ? rx : ry;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      tmp[3 * i + j] = a1[3 * i] * b1[j] + a1[3 * i + 1] * b1[3 + j] + a1[3 * i + 2] * b1[6 + j];
  const double* a2 = (flags & FR_ROT_ZYX) ? rz : tmp;
  const double* b2 = (flags & FR_ROT_ZYX) ? tmp : rz;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      r[3 * i + j] = a2[3 * i] * b2[j] + a2[3 * i + 1] * b2[3 + j] + a2[3 * i + 2] * b2[6 + j];
  const float f = p[6];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const float rf = (float)r[i];
    out[i] = __fmul_rn(f, rf);
    out[12 + i] = rf;
  }
  out[9] = p[3];
  out[10] = p[4];
  out[11] = p[5];
  out[21] = f;
  out[22] = 0.0f;
  out[23] = 0.0f;
}

// FaceRecNet.set_constraints (nets/network.py:204-218) for entry i of the parameter vector: sigmoid, then the per-block
// affine map, with the reference's separate multiply / subtract roundings.  d(constrained)/d(raw) = scale * s (1 - s).
__device__ __forceinline__ float param_scale(int i, int ks, float im_size) {
  return i < 3 ? 3.0f : (i < 5 ? im_size : (i == 5 ? 0.0f : (i == 6 ? 1e-3f : (i < FR_NDIM_POSE + ks ? 1e4f : 3.0f))));
}
__device__ __forceinline__ float param_sigmoid(float raw) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-raw))); }
__device__ __forceinline__ float read_param(const float* __restrict__ row, int i, unsigned flags, int ks, float im_size) {
  const float v = row[i];
  if (!(flags & FR_PARAMS_RAW)) return v;
  const float y = __fmul_rn(param_sigmoid(v), param_scale(i, ks, im_size));
  return (i < 3 || i >= FR_NDIM_POSE + ks) ? __fsub_rn(y, 1.5f) : y;
}
__device__ __forceinline__ void read_pose_params(const float* __restrict__ row, unsigned flags, int ks, float im_size, float* p7) {
#pragma unroll
  for (int i = 0; i < FR_NDIM_POSE; ++i) p7[i] = read_param(row, i, flags, ks, im_size);
}

// coefT [kpad][bpad]: row k = coefficient k of every face (k == ks+ke: the constant 1 that multiplies the mean
// column; beyond: 0; padded faces: 0).  pose [bpad][24].
__global__ void __launch_bounds__(256)
recon_prep_kernel(const float* __restrict__ params, int dparam, int batch, int bpad, int ks, int ke, int kpad,
                  unsigned flags, float im_size, float* __restrict__ coefT, float* __restrict__ pose) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx < kpad * bpad) {
    const int k = idx / bpad, b = idx - k * bpad;
    float v = 0.0f;
    if (b < batch) {
      if (k < ks + ke) v = read_param(params + (size_t)b * dparam, FR_NDIM_POSE + k, flags, ks, im_size);
      else if (k == ks + ke) v = 1.0f;
    }
    coefT[idx] = v;
  }
  if (idx < bpad) {
    if (idx < batch) {
      float p7[FR_NDIM_POSE];
      read_pose_params(params + (size_t)idx * dparam, flags, ks, im_size, p7);
      pose_matrices(p7, flags, pose + (size_t)idx * kPoseStride);
    } else {
      for (int i = 0; i < kPoseStride; ++i) pose[(size_t)idx * kPoseStride + i] = 0.0f;
    }
  }
}

// Where the forward kernels put a projected vertex.  `planar` is the reference's vertex_proj [B,3,N] tensor
// (nets/network.py:171); `rec` is the rasterizer's 16-byte vertex record array [B][N] {x, y, z, snap code}
// (raster.cuh): the fused params -> depth-map call has the reconstruction epilogue write the records directly, which
// saves the rasterizer's repack pass over the vertex tensor.  Either may be null.
struct ReconOut {
  float* planar;
  float4* rec;
  int width, height;   // image size the snap codes refer to (rec != nullptr)
  const int32_t* vert_rank;   // records are stored by vertex RANK (mesh_table.h) when a mesh table is in play, else by vertex id
};

// Projection + y flip of one reconstructed vertex (nets/network.py:163-169).
__device__ __forceinline__ void project_vertex(const float* __restrict__ P, float x, float y, float z, float im_size,
                                               unsigned flags, float* X, float* Y, float* Z) {
  *X = fmaf(P[2], z, fmaf(P[1], y, P[0] * x)) + P[9];
  float yy = fmaf(P[5], z, fmaf(P[4], y, P[3] * x)) + P[10];
  *Z = fmaf(P[8], z, fmaf(P[7], y, P[6] * x)) + P[11];
  if (!(flags & FR_YFLIP_NONE)) {
    yy = __fsub_rn(im_size, yy);                        // S - y
    if (!(flags & FR_YFLIP_S_Y)) yy = __fsub_rn(yy, 1.0f);  // ... - 1
  }
  *Y = yy;
}
__device__ __forceinline__ void store_planar(float* __restrict__ planar, int b, int nver, int n, float X, float Y, float Z) {
  float* o = planar + (size_t)b * 3 * nver;
  o[n] = X;
  o[(size_t)nver + n] = Y;
  o[2 * (size_t)nver + n] = Z;
}
__device__ __forceinline__ void store_record(const ReconOut& out, int b, int nver, int slot, float X, float Y, float Z) {
  out.rec[(size_t)b * nver + slot] = make_float4(X, Y, Z, __uint_as_float(fr_snap_code(X, Y, out.width, out.height)));
}
__device__ __forceinline__ void store_vertex(const ReconOut& out, int b, int nver, int n, float X, float Y, float Z) {
  if (out.planar != nullptr) store_planar(out.planar, b, nver, n, X, Y, Z);
  if (out.rec != nullptr) store_record(out, b, nver, out.vert_rank != nullptr ? __ldg(out.vert_rank + n) : n, X, Y, Z);
}
__device__ __forceinline__ void project_store(const float* __restrict__ P, float x, float y, float z, float im_size,
                                              unsigned flags, const ReconOut& out, int b, int nver, int n) {
  float X, Y, Z;
  project_vertex(P, x, y, z, im_size, flags, &X, &Y, &Z);
  store_vertex(out, b, nver, n, X, Y, Z);
}

// ---------------------------------------------------------------------------------------------- forward (SIMT)
// grid (ntiles, ceil(batch / (FB*blockDim.y))), block (128, GY).  Thread (v, gy) owns vertex tile*128+v for the
// FB faces b0 + gy*FB ...; it streams that vertex's three basis rows (coalesced float4 across the warp) and keeps
// 3*FB accumulators in registers; coefficients come from shared memory as warp-wide broadcasts.
template <int FB>
__global__ void __launch_bounds__(512)
recon_fwd_simt_kernel(const float4* __restrict__ packed, const float* __restrict__ coefT, const float* __restrict__ pose,
                      ReconOut out, int batch, int bpad, int nver, int kg, float im_size,
                      unsigned flags) {
  extern __shared__ __align__(16) float fr_smem[];
  const int gy = blockDim.y;
  const int fbt = FB * gy;
  const int kpad = kg * 4;
  const int b0 = blockIdx.y * fbt;
  const int tid = threadIdx.y * kTileVerts + threadIdx.x;
  const int nthr = kTileVerts * gy;
  {
    const int f4_per_row = fbt / 4;
    float4* dst = reinterpret_cast<float4*>(fr_smem);
    for (int i = tid; i < kpad * f4_per_row; i += nthr) {
      const int k = i / f4_per_row, q = i - k * f4_per_row;
      dst[i] = *reinterpret_cast<const float4*>(coefT + (size_t)k * bpad + b0 + 4 * q);
    }
  }
  __syncthreads();

  const int v = threadIdx.x;
  const int tile = blockIdx.x;
  const float4* bp = packed + (size_t)tile * 3 * kg * kTileVerts + v;
  const size_t cstride = (size_t)kg * kTileVerts;  // float4s between coordinates
  const float* cf = fr_smem + threadIdx.y * FB;

  float acc[3][FB];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int j = 0; j < FB; ++j) acc[c][j] = 0.0f;

  float4 a[3], nx[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) a[c] = ld_stream_f4(bp + c * cstride);
  for (int g = 0; g < kg; ++g) {
    if (g + 1 < kg) {
#pragma unroll
      for (int c = 0; c < 3; ++c) nx[c] = ld_stream_f4(bp + c * cstride + (size_t)(g + 1) * kTileVerts);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4* cr = reinterpret_cast<const float4*>(cf + (size_t)(4 * g + j) * fbt);
      const float a0 = f4_get(a[0], j), a1 = f4_get(a[1], j), a2 = f4_get(a[2], j);
#pragma unroll
      for (int q = 0; q < FB / 4; ++q) {
        const float4 w = cr[q];
        acc[0][4 * q + 0] = fmaf(a0, w.x, acc[0][4 * q + 0]);
        acc[0][4 * q + 1] = fmaf(a0, w.y, acc[0][4 * q + 1]);
        acc[0][4 * q + 2] = fmaf(a0, w.z, acc[0][4 * q + 2]);
        acc[0][4 * q + 3] = fmaf(a0, w.w, acc[0][4 * q + 3]);
        acc[1][4 * q + 0] = fmaf(a1, w.x, acc[1][4 * q + 0]);
        acc[1][4 * q + 1] = fmaf(a1, w.y, acc[1][4 * q + 1]);
        acc[1][4 * q + 2] = fmaf(a1, w.z, acc[1][4 * q + 2]);
        acc[1][4 * q + 3] = fmaf(a1, w.w, acc[1][4 * q + 3]);
        acc[2][4 * q + 0] = fmaf(a2, w.x, acc[2][4 * q + 0]);
        acc[2][4 * q + 1] = fmaf(a2, w.y, acc[2][4 * q + 1]);
        acc[2][4 * q + 2] = fmaf(a2, w.z, acc[2][4 * q + 2]);
        acc[2][4 * q + 3] = fmaf(a2, w.w, acc[2][4 * q + 3]);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) a[c] = nx[c];
  }

  const int n = tile * kTileVerts + v;
  if (n < nver) {
#pragma unroll
    for (int j = 0; j < FB; ++j) {
      const int b = b0 + threadIdx.y * FB + j;
      if (b < batch)
        project_store(pose + (size_t)b * kPoseStride, acc[0][j], acc[1][j], acc[2][j], im_size, flags, out, b, nver, n);
    }
  }
}

// ---------------------------------------------------------------------------------------------- backward
// d t3d[b][r] = sum_n g'[b][r][n]  (g' = upstream gradient with the y row negated when a flip is active); optionally also
// gmax[b*4 + r] = max_n |g[b][r][n]| (the tensor-core backward scales its fp16 operand by it).
__global__ void __launch_bounds__(256)
recon_bwd_dt_kernel(const float* __restrict__ vertex_grad, int nver, unsigned flags, float* __restrict__ dt,
                    float* __restrict__ gmax) {
  const int b = blockIdx.x, r = blockIdx.y;
  const float* g = vertex_grad + ((size_t)b * 3 + r) * nver;
  float s = 0.0f, m = 0.0f;
  for (int n = threadIdx.x; n < nver; n += 256) {
    const float v = g[n];
    s += v;
    m = fmaxf(m, fabsf(v));
  }
  __shared__ float red[8], redm[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
  }
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5] = s;
    redm[threadIdx.x >> 5] = m;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.0f, mx = 0.0f;
    for (int w = 0; w < 8; ++w) {
      tot += red[w];
      mx = fmaxf(mx, redm[w]);
    }
    if (r == 1 && !(flags & FR_YFLIP_NONE)) tot = -tot;
    dt[b * 4 + r] = tot;
    if (gmax != nullptr) gmax[b * 4 + r] = mx;
  }
}

constexpr int kBwdVC = 16;          // vertices staged per chunk
constexpr int kBwdKG = 64;          // k-groups (float4 columns) per CTA == blockDim.x
constexpr int kBwdFB = 16;          // faces per thread
constexpr int kBwdRow = kBwdVC + 1; // padded float4 row: conflict-free transposed reads

// grid (ctas, ceil(batch/(16*blockDim.y)), ceil(kg/64)), block (64, GY).  Thread (gx, gy) accumulates
// G[b][4g..4g+3] for g = blockIdx.z*64+gx and the 16 faces b0+gy*16..; CTAs stride over vertex tiles and finish with
// one atomicAdd per accumulator into G [bpad][kpad] (zeroed by the API).
__global__ void __launch_bounds__(256)
recon_bwd_simt_kernel(const float4* __restrict__ packed, const float* __restrict__ pose,
                      const float* __restrict__ vertex_grad, float* __restrict__ G, int batch, int nver, int kg,
                      int ntiles, unsigned flags) {
  extern __shared__ __align__(16) float fr_smem[];
  const int gy = blockDim.y;
  const int fbt = kBwdFB * gy;
  const int dstride = fbt + 4;                                   // floats per (c, v) row of dv
  float4* bs = reinterpret_cast<float4*>(fr_smem);               // [3][64][17] float4
  float* dvs = fr_smem + 3 * kBwdKG * kBwdRow * 4;               // [3][16][fbt+4]
  const int tid = threadIdx.y * kBwdKG + threadIdx.x;
  const int nthr = kBwdKG * gy;
  const int b0 = blockIdx.y * fbt;
  const int g0 = blockIdx.z * kBwdKG;
  const int kgc = min(kBwdKG, kg - g0);
  const int kpad = kg * 4;
  const float ysign = (flags & FR_YFLIP_NONE) ? 1.0f : -1.0f;

  float acc[4][kBwdFB];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int f = 0; f < kBwdFB; ++f) acc[j][f] = 0.0f;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    for (int vc = 0; vc < kTileVerts / kBwdVC; ++vc) {
      const int nbase = tile * kTileVerts + vc * kBwdVC;
      if (nbase >= nver) break;
      __syncthreads();  // previous chunk fully consumed
      // stage basis chunk: bs[c][g][vv]
      for (int i = tid; i < 3 * kgc * kBwdVC; i += nthr) {
        const int vv = i % kBwdVC;
        const int g = (i / kBwdVC) % kgc;
        const int c = i / (kBwdVC * kgc);
        bs[(c * kBwdKG + g) * kBwdRow + vv] =
            ld_stream_f4(packed + ((size_t)(tile * 3 + c) * kg + g0 + g) * kTileVerts + vc * kBwdVC + vv);
      }
      // stage dv[c][vv][f] = sum_r R_b[r][c] * g'[b][r][n]
      for (int i = tid; i < fbt * kBwdVC; i += nthr) {
        const int vv = i % kBwdVC;
        const int f = i / kBwdVC;
        const int b = b0 + f;
        const int n = nbase + vv;
        float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f;
        if (b < batch && n < nver) {
          const float* gp = vertex_grad + (size_t)b * 3 * nver + n;
          const float gx = gp[0], gyv = ysign * gp[nver], gz = gp[2 * (size_t)nver];
          const float* R = pose + (size_t)b * kPoseStride + 12;
          d0 = fmaf(R[6], gz, fmaf(R[3], gyv, R[0] * gx));
          d1 = fmaf(R[7], gz, fmaf(R[4], gyv, R[1] * gx));
          d2 = fmaf(R[8], gz, fmaf(R[5], gyv, R[2] * gx));
        }
        dvs[(0 * kBwdVC + vv) * dstride + f] = d0;
        dvs[(1 * kBwdVC + vv) * dstride + f] = d1;
        dvs[(2 * kBwdVC + vv) * dstride + f] = d2;
      }
      __syncthreads();
      if ((int)threadIdx.x < kgc) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
#pragma unroll 4
          for (int vv = 0; vv < kBwdVC; ++vv) {
            const float4 a = bs[(c * kBwdKG + threadIdx.x) * kBwdRow + vv];
            const float4* dr = reinterpret_cast<const float4*>(dvs + (c * kBwdVC + vv) * dstride + threadIdx.y * kBwdFB);
#pragma unroll
            for (int q = 0; q < kBwdFB / 4; ++q) {
              const float4 w = dr[q];
              const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                acc[0][4 * q + e] = fmaf(a.x, wv[e], acc[0][4 * q + e]);
                acc[1][4 * q + e] = fmaf(a.y, wv[e], acc[1][4 * q + e]);
                acc[2][4 * q + e] = fmaf(a.z, wv[e], acc[2][4 * q + e]);
                acc[3][4 * q + e] = fmaf(a.w, wv[e], acc[3][4 * q + e]);
              }
            }
          }
        }
      }
    }
  }
  if ((int)threadIdx.x < kgc) {
    const int k0 = 4 * (g0 + threadIdx.x);
#pragma unroll
    for (int f = 0; f < kBwdFB; ++f) {
      const int b = b0 + threadIdx.y * kBwdFB + f;
      if (b < batch) {
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(G + (size_t)b * kpad + k0 + j, acc[j][f]);
      }
    }
  }
}

// params_grad[b] = [0,0,0 | dt | df | f.G[:ks+ke]]   with df = sum_{k<=ks+ke} coef[k][b] * G[b][k]
__global__ void __launch_bounds__(256)
// gmean64 (nullable): the mean column's contraction accumulated separately in float64 (tensor-core backward).
// d f is a sum of ~230 terms of either sign around 1e7: it is accumulated in float64 so that the result is as good as its terms.
recon_bwd_finalize_kernel(const float* __restrict__ G, const float* __restrict__ coefT, const float* __restrict__ pose,
                          const float* __restrict__ dt, int bpad, int ks, int ke, int kpad, int dparam,
                          float* __restrict__ params_grad, const double* __restrict__ gmean64, const float* __restrict__ params,
                          unsigned flags, float im_size) {
  const int b = blockIdx.x;
  // FR_PARAMS_RAW: chain through set_constraints, d raw_i = d p_i * scale_i * s (1 - s)
  auto chain = [&](int i, float g) {
    if (!(flags & FR_PARAMS_RAW)) return g;
    const float s = param_sigmoid(params[(size_t)b * dparam + i]);
    return g * (param_scale(i, ks, im_size) * (s * (1.0f - s)));
  };
  const float f = pose[(size_t)b * kPoseStride + 21];
  const float* Gb = G + (size_t)b * kpad;
  float* out = params_grad + (size_t)b * dparam;
  double s = 0.0;
  for (int k = threadIdx.x; k <= ks + ke; k += 256) {
    const float gk = Gb[k];
    s += (double)coefT[(size_t)k * bpad + b] * (double)gk;
    if (k == ks + ke && gmean64 != nullptr) s += gmean64[b];
    if (k < ks + ke) out[FR_NDIM_POSE + k] = chain(FR_NDIM_POSE + k, f * gk);
  }
  __shared__ double red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot64 = 0.0;
    for (int w = 0; w < 8; ++w) tot64 += red[w];
    const float tot = (float)tot64;
    out[0] = 0.0f;  // tf.py_func has no gradient (nets/network.py:150)
    out[1] = 0.0f;
    out[2] = 0.0f;
    out[3] = chain(3, dt[b * 4 + 0]);
    out[4] = chain(4, dt[b * 4 + 1]);
    out[5] = chain(5, dt[b * 4 + 2]);
    out[6] = chain(6, tot);
  }
}

}  // namespace fr
#endif  // FR_RECON_CUH_
