// C ABI of facerecon_b200 (include/facerecon_b200.h): argument validation, workspace carving and kernel
// launches.  No CPU fallback anywhere: every compute entry point launches sm_100a kernels or fails.
#include <algorithm>
#include <cstring>
#include <exception>
#include <mutex>
#include <new>
#include <vector>

#include "fr_common.cuh"
#include "mesh_table.h"
#include "raster.cuh"
#include "raster_tile.cuh"
#include "recon.cuh"
#include "recon_f16.cuh"
#include "recon_bwd_f16.cuh"

// Host handle of a mesh table (mesh_table.h): the blob on the host and, when created for a device, its copy there.
struct fr_mesh_table {
  std::vector<unsigned char> host;
  unsigned char* dev;
  int device;
  fr::MeshTableHeader hdr;
};

using namespace fr;

namespace {

constexpr size_t kAlign = 256;

struct ReconWorkspace {
  float* coefT;   // [kpad][bpad]
  float* pose;    // [bpad][24]
  float* G;       // [bpad][kpad]
  double* gmean64;  // [bpad] the mean column of G in float64 (tensor-core backward); directly after G: one memset clears both
  float* dt;      // [bpad][4]
  void* bsplit16; // tensor-core forward: fp16 coefficient operands per 64-face batch tile
  float* pose16;  // ... and [bpad][16] scaled pose
  unsigned char* gtiles;  // tensor-core backward: fp16 operand tiles of the rotated vertex gradient
  float* gmax;            // ... [bpad][4] per-face, per-row gradient maxima
  float* gscale;          // ... [bpad] 2^-u_b
  size_t bytes;
};

ReconWorkspace carve_recon(void* base, int batch, const BasisGeom& g) {
  const int bpad = batch_padded(batch);
  ReconWorkspace w;
  size_t off = 0;
  auto take = [&](size_t n) {
    void* p = base ? static_cast<char*>(base) + off : nullptr;
    off += align_up(n, kAlign);
    return p;
  };
  w.coefT = static_cast<float*>(take(sizeof(float) * (size_t)g.kpad * bpad));
  w.pose = static_cast<float*>(take(sizeof(float) * (size_t)bpad * kPoseStride));
  w.G = static_cast<float*>(take(sizeof(float) * (size_t)bpad * g.kpad));
  w.gmean64 = static_cast<double*>(take(sizeof(double) * (size_t)bpad));
  w.dt = static_cast<float*>(take(sizeof(float) * (size_t)bpad * 4));
  w.bsplit16 = take(recon_f16_bsplit_bytes(batch, g));
  w.pose16 = static_cast<float*>(take(recon_f16_pose_bytes(batch)));
  w.gtiles = static_cast<unsigned char*>(take(recon_bwd_f16_fits(g) ? b16::grad_tiles_bytes(batch, g) : 0));
  w.gmax = static_cast<float*>(take(sizeof(float) * (size_t)bpad * 4));
  w.gscale = static_cast<float*>(take(sizeof(float) * (size_t)bpad));
  w.bytes = off;
  return w;
}

int check_model_dims(int batch, int nver, int ks, int ke) {
  FR_REQUIRE(batch >= 0, "batch must be >= 0 (got %d)", batch);
  FR_REQUIRE(nver > 0, "nver must be > 0 (got %d)", nver);
  FR_REQUIRE(ks >= 0 && ke >= 0 && ks + ke > 0, "ndim_shape/ndim_exp must be >= 0 and not both 0 (got %d, %d)", ks, ke);
  return FR_OK;
}

int check_workspace(void* ws, size_t have, size_t need) {
  if (need == 0) return FR_OK;
  if (ws == nullptr) return fail(FR_ERR_WORKSPACE, "workspace is null (%zu bytes required)", need);
  if (reinterpret_cast<uintptr_t>(ws) % kAlign != 0) return fail(FR_ERR_WORKSPACE, "workspace must be %zu-byte aligned", kAlign);
  if (have < need) return fail(FR_ERR_WORKSPACE, "workspace too small: %zu < %zu bytes", have, need);
  return FR_OK;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

template <int FB>
int launch_recon_fwd_simt(const float* packed, const ReconWorkspace& w, ReconOut out, int batch, int nver,
                          const BasisGeom& g, float im_size, unsigned flags, int gy, cudaStream_t st) {
  const int fbt = FB * gy;
  const size_t smem = sizeof(float) * (size_t)g.kpad * fbt;
  FR_CUDA(cudaFuncSetAttribute(recon_fwd_simt_kernel<FB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(g.ntiles, ceil_div(batch, fbt)), block(kTileVerts, gy);
  recon_fwd_simt_kernel<FB><<<grid, block, smem, st>>>(reinterpret_cast<const float4*>(packed), w.coefT, w.pose,
                                                       out, batch, batch_padded(batch), nver, g.kg, im_size, flags);
  FR_LAUNCHED("recon_fwd_simt_kernel");
  return FR_OK;
}

int check_mesh(const fr_mesh_table* mesh, int nver, int ntri) {
  if (mesh == nullptr) return FR_OK;
  FR_REQUIRE(mesh->dev != nullptr, "the mesh table was created without a device copy (device < 0)");
  FR_REQUIRE(mesh->hdr.nver == nver, "mesh table was built for nver=%d, called with nver=%d", mesh->hdr.nver, nver);
  FR_REQUIRE(ntri < 0 || mesh->hdr.ntri == ntri, "mesh table was built for ntri=%d, called with ntri=%d", mesh->hdr.ntri, ntri);
  return FR_OK;
}
const int32_t* mesh_cluster_vert(const fr_mesh_table* mesh) {
  return mesh ? reinterpret_cast<const int32_t*>(mesh->dev + mesh->hdr.off_vert) : nullptr;
}
const int32_t* mesh_rank_vert(const fr_mesh_table* mesh) {
  return mesh ? reinterpret_cast<const int32_t*>(mesh->dev + mesh->hdr.off_rank_vert) : nullptr;
}
const int32_t* mesh_vert_rank(const fr_mesh_table* mesh) {
  return mesh ? reinterpret_cast<const int32_t*>(mesh->dev + mesh->hdr.off_vert_rank) : nullptr;
}

// Tensor-core path for this batch?  (FR_RECON_PATH = simt | f16 overrides the dispatch, for A/B comparisons.)
bool use_f16_forward(const BasisGeom& g, int batch, bool raster) {
  const int ov = recon_path_override();
  return recon_f16_fits(g, raster) && (ov == 3 || (ov == 0 && batch > 8));
}

// Row tiles of the packed basis: the mesh table's clusters when the caller says so (FR_CLUSTER_TILES), else consecutive vertices.
int tile_clusters(const fr_mesh_table* mesh, unsigned flags) { return (mesh != nullptr && (flags & FR_CLUSTER_TILES)) ? mesh->hdr.nclusters : 0; }

// fr_recon_project_forward with a choice of outputs (out.planar / out.rec).  With `target` the tensor-core kernel also
// rasterizes (cluster tiles required, use_f16_forward(g, batch, true) must hold) and `out` becomes optional.
// clear_keys / clear_bytes: visibility keys to be cleared on the side (by the prep kernel on the tensor-core path).
int recon_project_forward_impl(const float* params, const float* packed, const fr_mesh_table* mesh, const ReconOut& out,
                               const f16::RasterTarget* target, unsigned long long* clear_keys, size_t clear_bytes, int batch, int nver,
                               int ndim_shape, int ndim_exp, float im_size, unsigned flags, void* workspace, size_t workspace_bytes,
                               void* stream) {
  if (int rc = check_model_dims(batch, nver, ndim_shape, ndim_exp)) return rc;
  if (int rc = check_mesh(mesh, nver, -1)) return rc;
  FR_REQUIRE(!(flags & FR_CLUSTER_TILES) || mesh != nullptr, "FR_CLUSTER_TILES needs the mesh table the basis was packed with");
  if (batch == 0) return FR_OK;
  FR_REQUIRE(params && packed && (out.planar || out.rec || target), "null pointer argument");
  const BasisGeom g = basis_geom(nver, ndim_shape, ndim_exp, tile_clusters(mesh, flags));
  const ReconWorkspace w = carve_recon(workspace, batch, g);
  if (int rc = check_workspace(workspace, workspace_bytes, w.bytes)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int bpad = batch_padded(batch);
  const int dparam = FR_NDIM_POSE + ndim_shape + ndim_exp;

  if (target != nullptr || use_f16_forward(g, batch, false)) {
    const bool fold = clear_keys != nullptr && clear_bytes % 16 == 0;     // the prep kernel clears the keys itself, 16 bytes at a time
    if (clear_keys != nullptr && !fold) FR_CUDA(cudaMemsetAsync(clear_keys, 0, clear_bytes, st));
    // row map of the operand tiles: the cluster lists for the raster flavour, else rank order (null without a table)
    const int32_t* row_vert = target != nullptr ? mesh_cluster_vert(mesh) : mesh_rank_vert(mesh);
    return launch_recon_fwd_f16(params, packed, w.bsplit16, w.pose16, out, target, row_vert, fold ? clear_keys : nullptr,
                                fold ? clear_bytes : 0, batch, nver, g, im_size, flags, sm_count(), st);
  }
  if (clear_keys != nullptr) FR_CUDA(cudaMemsetAsync(clear_keys, 0, clear_bytes, st));
  recon_prep_kernel<<<ceil_div(g.kpad * bpad, 256), 256, 0, st>>>(params, dparam, batch, bpad, ndim_shape, ndim_exp,
                                                                 g.kpad, flags, im_size, w.coefT, w.pose);
  FR_LAUNCHED("recon_prep_kernel");
  if (batch <= 4) return launch_recon_fwd_simt<4>(packed, w, out, batch, nver, g, im_size, flags, 1, st);
  if (batch <= 8) return launch_recon_fwd_simt<8>(packed, w, out, batch, nver, g, im_size, flags, 1, st);
  const int gy = batch <= 16 ? 1 : (batch <= 32 ? 2 : 4);
  return launch_recon_fwd_simt<16>(packed, w, out, batch, nver, g, im_size, flags, gy, st);
}

size_t key_bytes(int batch, int height, int width) { return align_up(sizeof(unsigned long long) * (size_t)batch * height * width, kAlign); }

int check_render_dims(int batch, int nver, int ntri, int height, int width) {
  FR_REQUIRE(batch >= 0 && nver > 0 && ntri >= 0 && height > 0 && width > 0,
             "bad dimensions batch=%d nver=%d ntri=%d height=%d width=%d", batch, nver, ntri, height, width);
  // render_depth_op.cc:161-166: the reference refuses ntri >= 10M (its static scratch); nver < 2^24 keeps float indices exact
  FR_REQUIRE(ntri < 10 * 1000 * 1000, "Too many triangular %d >= %d", ntri, 10 * 1000 * 1000);
  FR_REQUIRE(nver <= (1 << 24), "nver %d exceeds the exact range of float triangle indices", nver);
  FR_REQUIRE(height <= 32000 && width <= 32000 && (long long)height * width < (1ll << 31), "image too large");
  FR_REQUIRE((long long)batch * 3 * nver < (1ll << 31), "batch * 3 * nver must stay below 2^31: split the batch");
  FR_REQUIRE(batch <= 65535, "batch %d exceeds 65535 faces per call: split the batch", batch);   // faces ride in gridDim.y
  return FR_OK;
}

// Smallest batch that goes to the tile rasterizer (FR_TILE_KEYS = 0 disables it, n > 0 sets the threshold).
int tile_keys_min_batch() {
  static const int v = env_int("FR_TILE_KEYS", 8);
  return v <= 0 ? (1 << 30) : v;
}

// Resolve pass: keys -> depth / tri_ind (+ normals, texture, rendering-layer post-processing).
int launch_resolve(const unsigned long long* keys, const float* vertex, const float4* rec, const int32_t* vert_rank,
                   const uint4* tri_rank4, const float4* tex4, const float* tri, const float* texture,
                   long long texture_batch_stride, float* depth, float* texture_image, float* normal, float* tri_ind, int batch,
                   int nver, int ntri, int npix, const LayerOut& layer, bool dependent, cudaStream_t st) {
  const dim3 rgrid(ceil_div(npix, kRasterThreads * kResolvePerThread), batch);
  const bool plain = texture_image == nullptr && normal == nullptr && !layer.enabled;
  if (plain && ((reinterpret_cast<uintptr_t>(keys) | reinterpret_cast<uintptr_t>(depth) | reinterpret_cast<uintptr_t>(tri_ind)) & 15u) == 0) {
    const unsigned long long n = (unsigned long long)batch * (unsigned long long)npix;      // all faces as one flat array
    const unsigned long long groups = (n >> 2) > 0 ? (n >> 2) : 1ull;
    FR_CUDA(launch_pdl(raster_resolve_depth_kernel, dim3((unsigned)((groups + kRasterThreads - 1) / kRasterThreads)), dim3(kRasterThreads), 0, st,
                       dependent, keys, depth, tri_ind, n));
    FR_LAUNCHED("raster_resolve_depth_kernel");
    return FR_OK;
  }
  if (texture_image != nullptr || normal != nullptr)
    FR_CUDA(launch_pdl(raster_resolve_kernel<true>, rgrid, dim3(kRasterThreads), 0, st, dependent, keys, vertex, rec, vert_rank, tri_rank4, tex4, tri, texture,
                       texture_batch_stride, depth, texture_image, normal, tri_ind, nver, ntri, npix, layer));
  else
    FR_CUDA(launch_pdl(raster_resolve_kernel<false>, rgrid, dim3(kRasterThreads), 0, st, dependent, keys, vertex, rec, vert_rank, tri_rank4, tex4, tri, texture,
                       texture_batch_stride, depth, texture_image, normal, tri_ind, nver, ntri, npix, layer));
  FR_LAUNCHED("raster_resolve_kernel");
  return FR_OK;
}

// Visibility pass on 16-byte vertex records: one thread per (triangle, face group); with a mesh table the triangles come
// from it (integer ids, cluster order), else from the reference's float index tensor.
int launch_keys(const float4* rec, const float* tri, const fr_mesh_table* mesh, unsigned long long* keys, int batch, int nver,
                int ntri, int height, int width, bool dependent, bool early_ok, cudaStream_t st) {
  const bool table = mesh != nullptr;
  const int nt = table ? mesh->hdr.ntri_slots : ntri;
  if (nt == 0) return FR_OK;
  // With a mesh table and enough faces to fill a tile: the shared-memory tile rasterizer (raster_tile.cuh), one block per
  // (cluster, 16 faces).  FR_TILE_KEYS=0 keeps the gather kernel below (A/B switch).
  if (table && batch >= tile_keys_min_batch() && height <= rt::kMaxExtent && width <= rt::kMaxExtent) {
#ifndef FR_TILE_NF
#define FR_TILE_NF 16
#endif
    constexpr int NF = FR_TILE_NF;
    const size_t smem = sizeof(rt::TileSmem<NF>);
    FR_CUDA(cudaFuncSetAttribute(rt::raster_tile_keys_kernel<NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // (dependent launch only behind the persistent reconstruction kernel: behind the pack pass its blocks, 53 KB of shared
    // memory each, would become resident early and starve the HBM-bound pack kernel -- measured 1.35 ms vs 1.12 ms per
    // batch-256 forward + backward)
    FR_CUDA(launch_pdl(rt::raster_tile_keys_kernel<NF>, dim3(mesh->hdr.nclusters, ceil_div(batch, NF)), dim3(rt::kTileThreads), smem, st,
                       dependent && early_ok, rec, static_cast<const unsigned char*>(mesh->dev), keys, batch, nver, height, width));
    FR_LAUNCHED("raster_tile_keys_kernel");
    return FR_OK;
  }
  const uint4* tv = table ? reinterpret_cast<const uint4*>(mesh->dev + mesh->hdr.off_tri_vid) : nullptr;
  const unsigned gx = (unsigned)ceil_div(nt, kKeysThreads);
#define FR_LAUNCH_KEYS(FPT, TABLE)                                                                                              \
  FR_CUDA(launch_pdl(raster_keys_kernel<FPT, TABLE>, dim3(gx, ceil_div(batch, FPT)), dim3(kKeysThreads), 0, st, dependent, rec, tri, tv, \
                     keys, batch, nver, nt, height, width))
  if (batch >= 8) {
    if (table) FR_LAUNCH_KEYS(8, true); else FR_LAUNCH_KEYS(8, false);
  } else if (batch >= 3) {
    if (table) FR_LAUNCH_KEYS(4, true); else FR_LAUNCH_KEYS(4, false);
  } else {
    if (table) FR_LAUNCH_KEYS(1, true); else FR_LAUNCH_KEYS(1, false);
  }
#undef FR_LAUNCH_KEYS
  FR_LAUNCHED("raster_keys_kernel");
  return FR_OK;
}

// fr_render_depth_forward: pack pass (vertex tensor -> 16-byte records, clears the keys), visibility pass, resolve pass.
// `ready`: what the workspace already holds for this batch -- kNothingReady; kRecordsReady: vertex records and cleared
// visibility keys (written by the fused call's reconstruction kernels); kKeysReady: records AND final visibility keys (the
// reconstruction kernel rasterized in its epilogue).  `vertex` may be null with records unless normals or texture are requested.
enum { kNothingReady = 0, kRecordsReady = 1, kKeysReady = 2 };
int render_depth_forward_impl(const float* vertex, const float* tri, const float* texture, long long texture_batch_stride,
                              float* depth, float* texture_image, float* normal, float* tri_ind, int batch, int nver,
                              int ntri, int height, int width, const fr_mesh_table* mesh, void* workspace, size_t workspace_bytes,
                              void* stream, int ready, LayerOut layer = LayerOut{nullptr, nullptr, nullptr, false},
                              cudaEvent_t after_keys = nullptr) {
  const bool records_ready = ready >= kRecordsReady;
  if (int rc = check_render_dims(batch, nver, ntri, height, width)) return rc;
  if (int rc = check_mesh(mesh, nver, ntri)) return rc;
  if (batch == 0) return FR_OK;
  FR_REQUIRE((vertex || records_ready) && (tri || ntri == 0) && depth && tri_ind, "null pointer argument");
  FR_REQUIRE(vertex || (texture_image == nullptr && normal == nullptr), "normals / texture need the planar vertex tensor");
  FR_REQUIRE(texture_image == nullptr || texture != nullptr, "texture_image requested without a texture");
  FR_REQUIRE(texture_batch_stride == 0 || texture_batch_stride >= 3ll * nver, "texture_batch_stride must be 0 or >= 3*nver");
  const size_t need = fr_render_workspace_bytes(batch, nver, height, width);
  if (int rc = check_workspace(workspace, workspace_bytes, need)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned long long* keys = static_cast<unsigned long long*>(workspace);
  const int npix = height * width;
  float4* rec = reinterpret_cast<float4*>(static_cast<char*>(workspace) + key_bytes(batch, height, width));
  const bool pdl = pdl_enabled();   // dependent launches: the kernels call pdl_wait() before touching their predecessor's output
  const bool draws = ntri > 0 && (mesh == nullptr || mesh->hdr.ntri_slots > 0);
  if (draws) {
    if (!records_ready) {   // the pack pass also clears the visibility keys
      raster_pack_kernel<<<dim3(ceil_div(nver, kRasterThreads * kSnapPerThread), batch), kRasterThreads, 0, st>>>(
          vertex, rec, keys, mesh_vert_rank(mesh), nver, npix, width, height);
      FR_LAUNCHED("raster_pack_kernel");
    }
    if (ready < kKeysReady)
      if (int rc = launch_keys(rec, tri, mesh, keys, batch, nver, ntri, height, width, pdl, records_ready, st)) return rc;
    if (after_keys != nullptr) FR_CUDA(cudaEventRecord(after_keys, st));
  } else if (!records_ready) {
    FR_CUDA(cudaMemsetAsync(keys, 0, sizeof(unsigned long long) * (size_t)batch * npix, st));
  }
  // (the records hold this batch's vertices whenever something was drawn; normals then gather them instead of the planar tensor;
  // with a mesh table the winner's vertex ranks come from one 16-byte gather, and a texture shared by all faces is repacked
  // once per call as float4 by rank)
  const uint4* tri_rank4 = (mesh != nullptr && draws) ? reinterpret_cast<const uint4*>(mesh->dev + mesh_off_tri_rank4(mesh->hdr)) : nullptr;
  float4* tex4 = nullptr;
  if (tri_rank4 != nullptr && texture_image != nullptr && texture_batch_stride == 0) {
    tex4 = reinterpret_cast<float4*>(static_cast<char*>(workspace) + key_bytes(batch, height, width) +
                                     align_up(sizeof(float4) * (size_t)batch * nver, kAlign));
    raster_pack_texture_kernel<<<ceil_div(nver, kRasterThreads), kRasterThreads, 0, st>>>(texture, mesh_vert_rank(mesh), tex4, nver);
    FR_LAUNCHED("raster_pack_texture_kernel");
  }
  return launch_resolve(keys, vertex, draws ? rec : nullptr, mesh_vert_rank(mesh), tri_rank4, tex4, tri, texture, texture_batch_stride,
                        depth, texture_image, normal, tri_ind, batch, nver, ntri, npix, layer,
                        pdl && draws && after_keys == nullptr && tex4 == nullptr, st);
}

}  // namespace

extern "C" {

const char* fr_last_error(void) { return error_buffer(); }
int fr_version(void) { return FR_VERSION; }
unsigned long long fr_launch_count(void) { return launch_counter().load(); }

// ------------------------------------------------------------------------------------------------ mesh table
static int mesh_finish(fr_mesh_table* m, int device, fr_mesh_table** out) {
  std::memcpy(&m->hdr, m->host.data(), sizeof(m->hdr));
  m->dev = nullptr;
  m->device = device;
  if (device >= 0) {
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc(&m->dev, m->host.size());
    if (e == cudaSuccess) e = cudaMemcpy(m->dev, m->host.data(), m->host.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      if (m->dev) cudaFree(m->dev);
      delete m;
      return fail(FR_ERR_CUDA, "mesh table upload failed: %s", cudaGetErrorString(e));
    }
  }
  *out = m;
  return FR_OK;
}

int fr_mesh_table_create(const float* tri, int ntri, int nver, const float* positions, int positions_interleaved, int device,
                         fr_mesh_table** out) {
  FR_REQUIRE(out != nullptr, "out is null");
  *out = nullptr;
  FR_REQUIRE(nver > 0 && nver <= (1 << 24) && ntri >= 0 && ntri < 10 * 1000 * 1000 && (tri != nullptr || ntri == 0),
             "bad mesh dimensions nver=%d ntri=%d", nver, ntri);
  fr_mesh_table* m = new (std::nothrow) fr_mesh_table();
  if (!m) return fail(FR_ERR_CUDA, "out of host memory");
  try {
    MeshTableBuilder builder(tri, ntri, nver, positions, positions_interleaved != 0);
    m->host = builder.build();
  } catch (const std::exception& e) {
    delete m;
    return fail(FR_ERR_CUDA, "mesh table build failed: %s", e.what());
  }
  return mesh_finish(m, device, out);
}

int fr_mesh_table_from_blob(const void* blob, size_t bytes, int device, fr_mesh_table** out) {
  FR_REQUIRE(out != nullptr, "out is null");
  *out = nullptr;
  FR_REQUIRE(blob != nullptr && bytes >= sizeof(MeshTableHeader), "mesh table blob too small");
  MeshTableHeader h;
  std::memcpy(&h, blob, sizeof(h));
  FR_REQUIRE(h.magic == kMeshMagic && h.version == kMeshVersion, "not a mesh table (magic %08x version %u)", h.magic, h.version);
  FR_REQUIRE(h.total_bytes == bytes && h.nclusters >= 0 && h.ntri_slots >= 0 && h.off_vert == sizeof(MeshTableHeader) &&
                 (size_t)h.off_vert + (size_t)h.nclusters * kClusterVerts * 4 <= h.off_tri_begin &&
                 (size_t)h.off_tri_begin + ((size_t)h.nclusters + 1) * 4 <= h.off_tri &&
                 (size_t)h.off_tri + (size_t)h.ntri_slots * 8 <= h.off_tri_vid &&
                 (size_t)h.off_tri_vid + (size_t)h.ntri_slots * 16 <= h.off_rank_vert && h.off_rank_vert <= h.off_vert_rank &&
                 h.nver > 0 && h.ntri >= 0 && (size_t)mesh_off_tri_rank4(h) + (size_t)h.ntri * 16 <= bytes,
             "mesh table blob is inconsistent");
  const unsigned char* p = static_cast<const unsigned char*>(blob);
  uint32_t hash = 2166136261u;
  for (size_t i = sizeof(MeshTableHeader); i < bytes; ++i) hash = (hash ^ p[i]) * 16777619u;
  FR_REQUIRE(hash == h.hash, "mesh table blob is corrupt (hash mismatch)");
  fr_mesh_table* m = new (std::nothrow) fr_mesh_table();
  if (!m) return fail(FR_ERR_CUDA, "out of host memory");
  m->host.assign(p, p + bytes);
  return mesh_finish(m, device, out);
}

void fr_mesh_table_destroy(fr_mesh_table* m) {
  if (!m) return;
  if (m->dev) {
    cudaSetDevice(m->device);
    cudaFree(m->dev);
  }
  delete m;
}

const void* fr_mesh_table_blob(const fr_mesh_table* m, size_t* bytes) {
  if (!m) return nullptr;
  if (bytes) *bytes = m->host.size();
  return m->host.data();
}
int fr_mesh_table_clusters(const fr_mesh_table* m) { return m ? m->hdr.nclusters : 0; }
int fr_mesh_table_vertex_slots(const fr_mesh_table* m) { return m ? m->hdr.nvert_slots : 0; }

// ------------------------------------------------------------------------------------------------ packing
size_t fr_packed_basis_bytes(int nver, int ndim_shape, int ndim_exp, unsigned layout_flags, const fr_mesh_table* mesh) {
  if (nver <= 0 || ndim_shape < 0 || ndim_exp < 0) return 0;
  return basis_geom(nver, ndim_shape, ndim_exp, tile_clusters(mesh, layout_flags)).bytes();
}

int fr_pack_basis(const float* mu, const float* pc_shape, const float* pc_exp, int nver, int ndim_shape, int ndim_exp,
                  unsigned layout_flags, const fr_mesh_table* mesh, float* packed, void* stream) {
  if (int rc = check_model_dims(0, nver, ndim_shape, ndim_exp)) return rc;
  if (int rc = check_mesh(mesh, nver, -1)) return rc;
  FR_REQUIRE(!(layout_flags & FR_CLUSTER_TILES) || mesh != nullptr, "FR_CLUSTER_TILES needs a mesh table");
  FR_REQUIRE(mu && packed && (pc_shape || ndim_shape == 0) && (pc_exp || ndim_exp == 0), "null model pointer");
  FR_REQUIRE(reinterpret_cast<uintptr_t>(packed) % 16 == 0, "packed basis must be 16-byte aligned");
  const BasisGeom g = basis_geom(nver, ndim_shape, ndim_exp, tile_clusters(mesh, layout_flags));
  const size_t total = (size_t)g.ntiles * 3 * g.kg * kTileVerts;
  pack_basis_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      mu, pc_shape, pc_exp, nver, ndim_shape, ndim_exp, g.kg, g.ntiles, layout_flags, reinterpret_cast<float4*>(packed));
  FR_LAUNCHED("pack_basis_kernel");
  // fp16-pair section: column maxima -> power-of-two column scales -> hi/lo operand tiles
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned char* base = reinterpret_cast<unsigned char*>(packed);
  float* scale = reinterpret_cast<float*>(base + g.scale_offset());
  FR_CUDA(cudaMemsetAsync(scale, 0, sizeof(float) * g.kpad16, st));
  f16::basis_colmax_kernel<<<(unsigned)(((size_t)3 * nver + 511) / 512), 256, 0, st>>>(mu, pc_shape, pc_exp, nver, ndim_shape, ndim_exp,
                                                                                     reinterpret_cast<unsigned*>(scale));
  FR_LAUNCHED("basis_colmax_kernel");
  f16::basis_colscale_kernel<<<ceil_div(g.kpad16, 256), 256, 0, st>>>(scale, g.kreal, g.kpad16);
  FR_LAUNCHED("basis_colscale_kernel");
  const size_t pieces = (size_t)g.ntiles * 3 * g.nch16 * 2 * kTileVerts;          // tiles of 128 consecutive vertices / ranks
  f16::pack_basis_f16_kernel<<<(unsigned)((pieces + 255) / 256), 256, 0, st>>>(mu, pc_shape, pc_exp, scale, nver, ndim_shape, ndim_exp,
                                                                              g.nch16, g.ntiles, layout_flags, mesh_rank_vert(mesh),
                                                                              reinterpret_cast<uint4*>(base + g.f16_offset()));
  FR_LAUNCHED("pack_basis_f16_kernel");
  if (g.nclusters > 0) {                                                         // ... and one 128-row tile per cluster (mesh_table.h)
    const size_t cpieces = (size_t)g.nclusters * 3 * g.nch16 * 2 * kTileVerts;
    f16::pack_basis_f16_kernel<<<(unsigned)((cpieces + 255) / 256), 256, 0, st>>>(mu, pc_shape, pc_exp, scale, nver, ndim_shape, ndim_exp,
                                                                                 g.nch16, g.nclusters, layout_flags, mesh_cluster_vert(mesh),
                                                                                 reinterpret_cast<uint4*>(base + g.f16c_offset()));
    FR_LAUNCHED("pack_basis_f16_kernel");
  }
  // ... and the same pairs transposed for the backward contraction over the vertices
  const size_t bpieces = (size_t)g.ntiles * 3 * (kTileVerts / 16) * g.mtiles() * 2 * kTileVerts;
  b16::pack_basis_bwd_kernel<<<(unsigned)((bpieces + 255) / 256), 256, 0, st>>>(pc_shape, pc_exp, scale, nver, ndim_shape, ndim_exp,
                                                                               g.mtiles(), g.ntiles, layout_flags,
                                                                               reinterpret_cast<uint4*>(base + g.bwd_offset()));
  FR_LAUNCHED("pack_basis_bwd_kernel");
  b16::pack_mean_kernel<<<ceil_div(3 * g.ntiles * kTileVerts, 256), 256, 0, st>>>(mu, nver, g.ntiles, layout_flags,
                                                                               reinterpret_cast<float*>(base + g.mean_offset()));
  FR_LAUNCHED("pack_mean_kernel");
  return FR_OK;
}

// ------------------------------------------------------------------------------------------------ recon
size_t fr_recon_workspace_bytes(int batch, int nver, int ndim_shape, int ndim_exp) {
  if (batch <= 0 || nver <= 0 || ndim_shape < 0 || ndim_exp < 0) return 0;
  return carve_recon(nullptr, batch, basis_geom(nver, ndim_shape, ndim_exp)).bytes;
}

int fr_recon_project_forward(const float* params, const float* packed, const fr_mesh_table* mesh, float* vertex_proj, int batch,
                             int nver, int ndim_shape, int ndim_exp, float im_size, unsigned flags, void* workspace,
                             size_t workspace_bytes, void* stream) {
  FR_REQUIRE(batch == 0 || vertex_proj != nullptr, "null pointer argument");
  const ReconOut out = {vertex_proj, nullptr, 0, 0, nullptr};
  return recon_project_forward_impl(params, packed, mesh, out, nullptr, nullptr, 0, batch, nver, ndim_shape, ndim_exp, im_size, flags,
                                    workspace, workspace_bytes, stream);
}

int fr_recon_project_backward(const float* params, const float* packed, const float* vertex_grad, float* params_grad,
                              int batch, int nver, int ndim_shape, int ndim_exp, float im_size, unsigned flags, void* workspace,
                              size_t workspace_bytes, void* stream) {
  if (int rc = check_model_dims(batch, nver, ndim_shape, ndim_exp)) return rc;
  if (batch == 0) return FR_OK;
  FR_REQUIRE(params && packed && vertex_grad && params_grad, "null pointer argument");
  const BasisGeom g = basis_geom(nver, ndim_shape, ndim_exp);
  const ReconWorkspace w = carve_recon(workspace, batch, g);
  if (int rc = check_workspace(workspace, workspace_bytes, w.bytes)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int bpad = batch_padded(batch);
  const int dparam = FR_NDIM_POSE + ndim_shape + ndim_exp;

  // recomputed rather than trusted from a previous forward: the workspace is the caller's scratch
  recon_prep_kernel<<<ceil_div(g.kpad * bpad, 256), 256, 0, st>>>(params, dparam, batch, bpad, ndim_shape, ndim_exp,
                                                                 g.kpad, flags, im_size, w.coefT, w.pose);
  FR_LAUNCHED("recon_prep_kernel");
  FR_CUDA(cudaMemsetAsync(w.G, 0, (size_t)(reinterpret_cast<char*>(w.gmean64 + bpad) - reinterpret_cast<char*>(w.G)), st));
  // dispatch: tcgen05 contraction above 8 faces (FR_RECON_PATH=simt forces the FFMA kernel)
  const uint32_t bwd_smem = b16::smem_bytes(g.mtiles(), b16::faces_per_tile(batch));
  const bool use_tc = recon_bwd_f16_fits(g) && bwd_smem <= 227u * 1024u && batch > 8 && recon_path_override() != 1;
  recon_bwd_dt_kernel<<<dim3(batch, 3), 256, 0, st>>>(vertex_grad, nver, flags, w.dt, use_tc ? w.gmax : nullptr);
  FR_LAUNCHED("recon_bwd_dt_kernel");
  if (use_tc) {
    const unsigned char* base = reinterpret_cast<const unsigned char*>(packed);
    const float* inv_scale = reinterpret_cast<const float*>(base + g.scale_offset());
    const int nb = b16::faces_per_tile(batch), nbt = ceil_div(batch, nb);
    // (launched over whole batch tiles: the operand tiles of the faces between batch and nbt * nb are zero-filled by the
    // blocks that own them, so the contraction never reads unwritten workspace)
    b16::recon_bwd_pack_grad_kernel<<<dim3(ceil_div(g.ntiles * (kTileVerts / 8), 32), ceil_div(nbt * nb, 8)), 256, 0, st>>>(
        vertex_grad, w.pose, w.gmax, reinterpret_cast<const float*>(base + g.mean_offset()), batch, nver, g.ntiles, nb, flags, w.gtiles,
        w.gscale, w.gmean64);
    FR_LAUNCHED("recon_bwd_pack_grad_kernel");
    int ctas = sm_count() / nbt;
    if (ctas < 1) ctas = 1;
    if (ctas > g.ntiles) ctas = g.ntiles;
    FR_CUDA(cudaFuncSetAttribute(b16::recon_bwd_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem));
    b16::recon_bwd_f16_kernel<<<dim3(ctas, nbt), b16::kThreads, bwd_smem, st>>>(base + g.bwd_offset(), w.gtiles, inv_scale, w.gscale,
                                                                              w.G, batch, nb, g.mtiles(), g.ntiles, g.kpad);
    FR_LAUNCHED("recon_bwd_f16_kernel");
    recon_bwd_finalize_kernel<<<batch, 256, 0, st>>>(w.G, w.coefT, w.pose, w.dt, bpad, ndim_shape, ndim_exp, g.kpad, dparam,
                                                    params_grad, w.gmean64, params, flags, im_size);
    FR_LAUNCHED("recon_bwd_finalize_kernel");
    return FR_OK;
  }

  const int gy = batch <= 16 ? 1 : (batch <= 32 ? 2 : 4);
  const int fbt = kBwdFB * gy;
  const size_t smem = sizeof(float) * ((size_t)3 * kBwdKG * kBwdRow * 4 + (size_t)3 * kBwdVC * (fbt + 4));
  FR_CUDA(cudaFuncSetAttribute(recon_bwd_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int ygroups = ceil_div(batch, fbt), zgroups = ceil_div(g.kg, kBwdKG);
  int ctas = (3 * sm_count()) / (ygroups * zgroups);
  if (ctas < 1) ctas = 1;
  if (ctas > g.ntiles) ctas = g.ntiles;
  recon_bwd_simt_kernel<<<dim3(ctas, ygroups, zgroups), dim3(kBwdKG, gy), smem, st>>>(
      reinterpret_cast<const float4*>(packed), w.pose, vertex_grad, w.G, batch, nver, g.kg, g.ntiles, flags);
  FR_LAUNCHED("recon_bwd_simt_kernel");
  recon_bwd_finalize_kernel<<<batch, 256, 0, st>>>(w.G, w.coefT, w.pose, w.dt, bpad, ndim_shape, ndim_exp, g.kpad, dparam,
                                                  params_grad, nullptr, params, flags, im_size);
  FR_LAUNCHED("recon_bwd_finalize_kernel");
  return FR_OK;
}

// ------------------------------------------------------------------------------------------------ render
size_t fr_render_workspace_bytes(int batch, int nver, int height, int width) {
  if (batch <= 0 || nver <= 0 || height <= 0 || width <= 0) return 0;
  return key_bytes(batch, height, width) +                                   // visibility keys
         align_up(sizeof(float4) * (size_t)batch * nver, kAlign) +           // vertex records
         align_up(sizeof(float4) * (size_t)nver, kAlign) +                   // a shared texture repacked by vertex rank
         kAlign;                                                             // schedule counters of the fused forward kernel
}

int fr_render_depth_forward(const float* vertex, const float* tri, const float* texture, long long texture_batch_stride,
                            float* depth, float* texture_image, float* normal, float* tri_ind, int batch, int nver,
                            int ntri, int height, int width, const fr_mesh_table* mesh, void* workspace, size_t workspace_bytes,
                            void* stream) {
  return render_depth_forward_impl(vertex, tri, texture, texture_batch_stride, depth, texture_image, normal, tri_ind, batch, nver,
                                   ntri, height, width, mesh, workspace, workspace_bytes, stream, kNothingReady);
}

int fr_render_depth_backward(const float* depth_grad, const float* tri, const float* tri_ind, float* vertex_grad,
                             int batch, int nver, int ntri, int height, int width, void* stream) {
  FR_REQUIRE(batch >= 0 && nver > 0 && ntri >= 0 && height > 0 && width > 0,
             "bad dimensions batch=%d nver=%d ntri=%d height=%d width=%d", batch, nver, ntri, height, width);
  FR_REQUIRE((long long)height * width < (1ll << 31), "image too large");
  FR_REQUIRE(batch <= 65535, "batch %d exceeds 65535 faces per call: split the batch", batch);
  if (batch == 0) return FR_OK;
  FR_REQUIRE(depth_grad && (tri || ntri == 0) && tri_ind && vertex_grad, "null pointer argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int npix = height * width;
  FR_CUDA(cudaMemsetAsync(vertex_grad, 0, sizeof(float) * (size_t)batch * 3 * nver, st));
  render_backward_kernel<<<dim3(ceil_div(npix, kRasterThreads), batch), kRasterThreads, 0, st>>>(depth_grad, tri, tri_ind,
                                                                                              vertex_grad, nver, ntri, npix,
                                                                                              nullptr, nullptr, nullptr);
  FR_LAUNCHED("render_backward_kernel");
  return FR_OK;
}

// ------------------------------------------------------------------------------------------------ rendering layer
int fr_rendering_layer_forward(const float* vertex, const float* tri, const float* texture, long long texture_batch_stride,
                               const float* im_gray, float* pncc, float* normalimg, float* maskimg, float* depthimg,
                               float* raw_depth, float* tri_ind, int batch, int nver, int ntri, int height, int width,
                               const fr_mesh_table* mesh, void* workspace, size_t workspace_bytes, void* stream) {
  FR_REQUIRE(batch <= 0 || (pncc && normalimg && maskimg && depthimg && texture), "null pointer argument");
  const LayerOut layer = {maskimg, im_gray, raw_depth, true};
  return render_depth_forward_impl(vertex, tri, texture, texture_batch_stride, depthimg, pncc, normalimg, tri_ind, batch, nver, ntri,
                                   height, width, mesh, workspace, workspace_bytes, stream, kNothingReady, layer);
}

int fr_rendering_layer_backward(const float* depthimg_grad, const float* maskimg_grad, const float* im_gray, const float* raw_depth,
                                const float* tri, const float* tri_ind, float* vertex_grad, int batch, int nver, int ntri,
                                int height, int width, void* stream) {
  FR_REQUIRE(batch >= 0 && nver > 0 && ntri >= 0 && height > 0 && width > 0,
             "bad dimensions batch=%d nver=%d ntri=%d height=%d width=%d", batch, nver, ntri, height, width);
  FR_REQUIRE((long long)height * width < (1ll << 31), "image too large");
  FR_REQUIRE(batch <= 65535, "batch %d exceeds 65535 faces per call: split the batch", batch);
  if (batch == 0) return FR_OK;
  FR_REQUIRE(raw_depth && (tri || ntri == 0) && tri_ind && vertex_grad && (depthimg_grad || maskimg_grad), "null pointer argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int npix = height * width;
  FR_CUDA(cudaMemsetAsync(vertex_grad, 0, sizeof(float) * (size_t)batch * 3 * nver, st));
  render_backward_kernel<<<dim3(ceil_div(npix, kRasterThreads), batch), kRasterThreads, 0, st>>>(
      depthimg_grad, tri, tri_ind, vertex_grad, nver, ntri, npix, maskimg_grad, im_gray, raw_depth);
  FR_LAUNCHED("render_backward_kernel");
  return FR_OK;
}

// ------------------------------------------------------------------------------------------------ fused
// Workspace of fr_recon_render_forward: [reconstruction | visibility keys | vertex records].
// Does the call run with the rasterizer inside the reconstruction epilogue?  (cluster tiles, tensor-core batch size)
static bool fused_raster(const fr_mesh_table* mesh, unsigned flags, int batch, int nver, int ndim_shape, int ndim_exp, int height,
                         int width) {
  if (mesh == nullptr || !(flags & FR_CLUSTER_TILES) || mesh->hdr.ntri_slots == 0) return false;
  if (height > rt::kMaxExtent || width > rt::kMaxExtent) return false;     // the tile rasterizer's cull needs code fields < 2^15
  return use_f16_forward(basis_geom(nver, ndim_shape, ndim_exp, mesh->hdr.nclusters), batch, true);
}

size_t fr_pipeline_workspace_bytes(int batch, int nver, int ndim_shape, int ndim_exp, int height, int width) {
  if (batch <= 0 || nver <= 0) return 0;
  return fr_recon_workspace_bytes(batch, nver, ndim_shape, ndim_exp) + fr_render_workspace_bytes(batch, nver, height, width);
}

// fr_recon_render_forward / fr_recon_render_forward_all: texture_image / normal null = depth + tri_ind only.
static int recon_render_forward_impl(const float* params, const float* packed, const float* tri, const fr_mesh_table* mesh,
                                     const float* texture, long long texture_batch_stride, float* vertex_proj, float* depth,
                                     float* texture_image, float* normal, float* tri_ind, int batch, int nver, int ntri,
                                     int ndim_shape, int ndim_exp, int height, int width, float im_size, unsigned flags,
                                     void* workspace, size_t workspace_bytes, void* stream, void* const* stage_events) {
  if (int rc = check_model_dims(batch, nver, ndim_shape, ndim_exp)) return rc;
  if (int rc = check_render_dims(batch, nver, ntri, height, width)) return rc;
  if (int rc = check_mesh(mesh, nver, ntri)) return rc;
  if (batch == 0) return FR_OK;
  FR_REQUIRE(params && packed && (tri || ntri == 0) && depth && tri_ind, "null pointer argument");
  const bool all = texture_image != nullptr || normal != nullptr;
  FR_REQUIRE(!all || vertex_proj != nullptr, "normals / texture need the vertex_proj output");
  FR_REQUIRE(texture_image == nullptr || texture != nullptr, "texture_image requested without a texture");
  const size_t rb = fr_recon_workspace_bytes(batch, nver, ndim_shape, ndim_exp);
  const size_t vb = fr_render_workspace_bytes(batch, nver, height, width);
  if (int rc = check_workspace(workspace, workspace_bytes, rb + vb)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* rws = static_cast<char*>(workspace) + rb;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(rws);
  float4* rec = reinterpret_cast<float4*>(rws + key_bytes(batch, height, width));
  const size_t kbytes = sizeof(unsigned long long) * (size_t)batch * height * width;
  const bool timed = stage_events != nullptr && (stage_events[0] != nullptr || stage_events[1] != nullptr);
  auto record = [&](int i) -> cudaError_t {
    return (stage_events != nullptr && stage_events[i] != nullptr) ? cudaEventRecord(static_cast<cudaEvent_t>(stage_events[i]), st)
                                                                   : cudaSuccess;
  };
  if (fused_raster(mesh, flags, batch, nver, ndim_shape, ndim_exp, height, width)) {
    // prep kernel (clears the keys and the schedule counters at the end of the render workspace) -> tensor-core
    // reconstruction with the tile rasterizer in its epilogue -> resolve.  With normals / texture requested the epilogue also
    // leaves the 16-byte vertex records (by rank) the resolve pass gathers them from: no repack pass, no visibility kernel.
    const f16::RasterTarget target = {static_cast<const unsigned char*>(mesh->dev), keys, width, height,
                                      reinterpret_cast<unsigned*>(rws + vb - kAlign)};
    const ReconOut out = {vertex_proj, all ? rec : nullptr, width, height, nullptr};
    if (int rc = recon_project_forward_impl(params, packed, mesh, out, &target, keys, kbytes, batch, nver, ndim_shape, ndim_exp, im_size,
                                            flags, workspace, rb, stream))
      return rc;
    FR_CUDA(record(0));
    FR_CUDA(record(1));
    if (all) {
      if (int rc = render_depth_forward_impl(vertex_proj, tri, texture, texture_batch_stride, depth, texture_image, normal, tri_ind, batch,
                                             nver, ntri, height, width, mesh, rws, vb, stream, kKeysReady))
        return rc;
    } else {
      const LayerOut layer = {nullptr, nullptr, nullptr, false};
      if (int rc = launch_resolve(keys, nullptr, nullptr, nullptr, nullptr, nullptr, tri, nullptr, 0, depth, nullptr, nullptr, tri_ind, batch, nver, ntri,
                                  height * width, layer, pdl_enabled() && !timed, st))
        return rc;
    }
    FR_CUDA(record(2));
    return FR_OK;
  }
  // The reconstruction kernels write the rasterizer's vertex records straight into the render workspace (same carve-up as
  // render_depth_forward_impl: keys first, records after) and clear its keys, so the repack pass over vertex_proj disappears;
  // vertex_proj itself is optional here.
  const ReconOut out = {vertex_proj, rec, width, height, mesh_vert_rank(mesh)};
  if (int rc = recon_project_forward_impl(params, packed, mesh, out, nullptr, keys, kbytes, batch, nver, ndim_shape, ndim_exp, im_size,
                                          flags, workspace, rb, stream))
    return rc;
  FR_CUDA(record(0));
  const cudaEvent_t after_keys = (stage_events != nullptr) ? static_cast<cudaEvent_t>(stage_events[1]) : nullptr;
  if (int rc = render_depth_forward_impl(vertex_proj, tri, texture, texture_batch_stride, depth, texture_image, normal, tri_ind, batch, nver,
                                         ntri, height, width, mesh, rws, vb, stream, kRecordsReady,
                                         LayerOut{nullptr, nullptr, nullptr, false}, after_keys))
    return rc;
  FR_CUDA(record(2));
  return FR_OK;
}

int fr_recon_render_forward(const float* params, const float* packed, const float* tri, const fr_mesh_table* mesh,
                            float* vertex_proj, float* depth, float* tri_ind, int batch, int nver, int ntri, int ndim_shape,
                            int ndim_exp, int height, int width, float im_size, unsigned flags, void* workspace,
                            size_t workspace_bytes, void* stream, void* const* stage_events) {
  return recon_render_forward_impl(params, packed, tri, mesh, nullptr, 0, vertex_proj, depth, nullptr, nullptr, tri_ind, batch, nver, ntri,
                                   ndim_shape, ndim_exp, height, width, im_size, flags, workspace, workspace_bytes, stream, stage_events);
}

int fr_recon_render_forward_all(const float* params, const float* packed, const float* tri, const fr_mesh_table* mesh,
                                const float* texture, long long texture_batch_stride, float* vertex_proj, float* depth,
                                float* texture_image, float* normal, float* tri_ind, int batch, int nver, int ntri, int ndim_shape,
                                int ndim_exp, int height, int width, float im_size, unsigned flags, void* workspace,
                                size_t workspace_bytes, void* stream) {
  FR_REQUIRE(vertex_proj && texture_image && normal, "null pointer argument");
  return recon_render_forward_impl(params, packed, tri, mesh, texture, texture_batch_stride, vertex_proj, depth, texture_image, normal,
                                   tri_ind, batch, nver, ntri, ndim_shape, ndim_exp, height, width, im_size, flags, workspace,
                                   workspace_bytes, stream, nullptr);
}

// ------------------------------------------------------------------------------------------------ session
// Device state of one in-flight batch.  Two slots let the device->host copy of one batch overlap the kernels of the
// next (fr_session_submit / fr_session_wait); fr_session_forward is submit + wait on slot 0.
struct fr_slot {
  cudaStream_t stream;
  float *params, *vertex, *depth, *tri_ind;
  void* ws;
  int batch;          // faces of the batch submitted last on this slot (0 = none)
  float im_size;
  bool pending;       // submitted and not yet waited for
};

struct fr_session {
  int device, nver, ntri, ks, ke, height, width, max_batch;
  unsigned flags;
  float *packed, *tri, *depth_grad, *vgrad, *pgrad;
  fr_mesh_table* mesh;
  size_t ws_bytes;
  fr_slot slot[FR_SESSION_SLOTS];
};

static void session_free(fr_session* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  for (fr_slot& sl : s->slot) {
    if (sl.stream) cudaStreamSynchronize(sl.stream);
    float* bufs[] = {sl.params, sl.vertex, sl.depth, sl.tri_ind};
    for (float* p : bufs)
      if (p) cudaFree(p);
    if (sl.ws) cudaFree(sl.ws);
    if (sl.stream) cudaStreamDestroy(sl.stream);
  }
  float* bufs[] = {s->packed, s->tri, s->depth_grad, s->vgrad, s->pgrad};
  for (float* p : bufs)
    if (p) cudaFree(p);
  fr_mesh_table_destroy(s->mesh);
  delete s;
}

int fr_session_create(const float* mu, const float* pc_shape, const float* pc_exp, const float* tri, int nver, int ntri,
                      int ndim_shape, int ndim_exp, int height, int width, int max_batch, unsigned flags, int device,
                      fr_session** out) {
  FR_REQUIRE(out != nullptr, "out is null");
  *out = nullptr;
  if (int rc = check_model_dims(max_batch, nver, ndim_shape, ndim_exp)) return rc;
  FR_REQUIRE(mu && tri && ntri > 0 && height > 0 && width > 0 && max_batch > 0, "bad session arguments");
  FR_CUDA(cudaSetDevice(device));
  fr_session* s = new (std::nothrow) fr_session();
  if (!s) return fail(FR_ERR_CUDA, "out of host memory");
  std::memset(s, 0, sizeof(*s));
  s->device = device; s->nver = nver; s->ntri = ntri; s->ks = ndim_shape; s->ke = ndim_exp;
  s->height = height; s->width = width; s->max_batch = max_batch; s->flags = flags;
  const size_t n3 = (size_t)3 * nver, npix = (size_t)height * width;
  const int d = FR_NDIM_POSE + ndim_shape + ndim_exp;
  float *d_mu = nullptr, *d_ps = nullptr, *d_pe = nullptr;
  int rc = FR_OK;
  auto chk = [&](cudaError_t e, const char* what) {
    if (e != cudaSuccess && rc == FR_OK) rc = fail(FR_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
    return e == cudaSuccess;
  };
  // mesh table from the mean shape's geometry (one-off, host side)
  if (int mrc = fr_mesh_table_create(tri, ntri, nver, mu, (flags & FR_MEAN_INTERLEAVED) ? 1 : 0, device, &s->mesh)) {
    session_free(s);
    return mrc;
  }
  s->ws_bytes = fr_pipeline_workspace_bytes(max_batch, nver, ndim_shape, ndim_exp, height, width);
  chk(cudaMalloc(&s->packed, fr_packed_basis_bytes(nver, ndim_shape, ndim_exp, flags, s->mesh)), "cudaMalloc(packed)");
  chk(cudaMalloc(&s->tri, sizeof(float) * 3 * (size_t)ntri), "cudaMalloc(tri)");
  for (fr_slot& sl : s->slot) {
    chk(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking), "cudaStreamCreate");
    chk(cudaMalloc(&sl.params, sizeof(float) * (size_t)max_batch * d), "cudaMalloc(params)");
    chk(cudaMalloc(&sl.vertex, sizeof(float) * (size_t)max_batch * n3), "cudaMalloc(vertex)");
    chk(cudaMalloc(&sl.depth, sizeof(float) * (size_t)max_batch * npix), "cudaMalloc(depth)");
    chk(cudaMalloc(&sl.tri_ind, sizeof(float) * (size_t)max_batch * npix), "cudaMalloc(tri_ind)");
    chk(cudaMalloc(&sl.ws, s->ws_bytes), "cudaMalloc(workspace)");
  }
  // backward staging is allocated on first use (fr_session_backward): forward-only callers never pay for it
  chk(cudaMalloc(&d_mu, sizeof(float) * n3), "cudaMalloc(mu)");
  if (ndim_shape) chk(cudaMalloc(&d_ps, sizeof(float) * n3 * ndim_shape), "cudaMalloc(pc_shape)");
  if (ndim_exp) chk(cudaMalloc(&d_pe, sizeof(float) * n3 * ndim_exp), "cudaMalloc(pc_exp)");
  cudaStream_t st = s->slot[0].stream;
  if (rc == FR_OK) {
    chk(cudaMemcpyAsync(d_mu, mu, sizeof(float) * n3, cudaMemcpyHostToDevice, st), "copy mu");
    if (ndim_shape) chk(cudaMemcpyAsync(d_ps, pc_shape, sizeof(float) * n3 * ndim_shape, cudaMemcpyHostToDevice, st), "copy pc_shape");
    if (ndim_exp) chk(cudaMemcpyAsync(d_pe, pc_exp, sizeof(float) * n3 * ndim_exp, cudaMemcpyHostToDevice, st), "copy pc_exp");
    chk(cudaMemcpyAsync(s->tri, tri, sizeof(float) * 3 * (size_t)ntri, cudaMemcpyHostToDevice, st), "copy tri");
  }
  if (rc == FR_OK) rc = fr_pack_basis(d_mu, d_ps, d_pe, nver, ndim_shape, ndim_exp, flags, s->mesh, s->packed, st);
  if (rc == FR_OK) chk(cudaStreamSynchronize(st), "cudaStreamSynchronize");
  if (d_mu) cudaFree(d_mu);
  if (d_ps) cudaFree(d_ps);
  if (d_pe) cudaFree(d_pe);
  if (rc != FR_OK) {
    session_free(s);
    return rc;
  }
  *out = s;
  return FR_OK;
}

void fr_session_destroy(fr_session* s) { session_free(s); }

int fr_session_submit(fr_session* s, int slot, const float* params, int batch, float im_size, float* depth, float* tri_ind,
                      float* vertex_proj) {
  FR_REQUIRE(s && params && depth, "null pointer argument");
  FR_REQUIRE(slot >= 0 && slot < FR_SESSION_SLOTS, "slot %d outside [0, %d)", slot, FR_SESSION_SLOTS);
  FR_REQUIRE(batch > 0 && batch <= s->max_batch, "batch %d outside (0, %d]", batch, s->max_batch);
  fr_slot& sl = s->slot[slot];
  FR_REQUIRE(!sl.pending, "slot %d still has a batch in flight: call fr_session_wait first", slot);
  FR_CUDA(cudaSetDevice(s->device));
  const int d = FR_NDIM_POSE + s->ks + s->ke;
  const size_t npix = (size_t)s->height * s->width;
  FR_CUDA(cudaMemcpyAsync(sl.params, params, sizeof(float) * (size_t)batch * d, cudaMemcpyHostToDevice, sl.stream));
  if (int rc = fr_recon_render_forward(sl.params, s->packed, s->tri, s->mesh, vertex_proj ? sl.vertex : nullptr, sl.depth, sl.tri_ind, batch,
                                       s->nver, s->ntri, s->ks, s->ke, s->height, s->width, im_size, s->flags, sl.ws, s->ws_bytes, sl.stream,
                                       nullptr))
    return rc;
  FR_CUDA(cudaMemcpyAsync(depth, sl.depth, sizeof(float) * batch * npix, cudaMemcpyDeviceToHost, sl.stream));
  if (tri_ind) FR_CUDA(cudaMemcpyAsync(tri_ind, sl.tri_ind, sizeof(float) * batch * npix, cudaMemcpyDeviceToHost, sl.stream));
  if (vertex_proj)
    FR_CUDA(cudaMemcpyAsync(vertex_proj, sl.vertex, sizeof(float) * (size_t)batch * 3 * s->nver, cudaMemcpyDeviceToHost, sl.stream));
  sl.batch = batch;
  sl.im_size = im_size;
  sl.pending = true;
  return FR_OK;
}

int fr_session_wait(fr_session* s, int slot) {
  FR_REQUIRE(s != nullptr, "null pointer argument");
  FR_REQUIRE(slot >= 0 && slot < FR_SESSION_SLOTS, "slot %d outside [0, %d)", slot, FR_SESSION_SLOTS);
  fr_slot& sl = s->slot[slot];
  if (!sl.pending) return FR_OK;
  sl.pending = false;
  FR_CUDA(cudaSetDevice(s->device));
  FR_CUDA(cudaStreamSynchronize(sl.stream));
  return FR_OK;
}

int fr_session_forward(fr_session* s, const float* params, int batch, float im_size, float* depth, float* tri_ind,
                       float* vertex_proj) {
  FR_REQUIRE(s != nullptr, "null pointer argument");
  if (int rc = fr_session_wait(s, 0)) return rc;
  if (int rc = fr_session_submit(s, 0, params, batch, im_size, depth, tri_ind, vertex_proj)) return rc;
  return fr_session_wait(s, 0);
}

int fr_session_backward(fr_session* s, const float* depth_grad, int batch, float* params_grad) {
  FR_REQUIRE(s && depth_grad && params_grad, "null pointer argument");
  fr_slot& sl = s->slot[0];
  FR_REQUIRE(!sl.pending, "slot 0 still has a batch in flight: call fr_session_wait first");
  FR_REQUIRE(batch > 0 && batch == sl.batch, "batch %d does not match the last forward on slot 0 (%d)", batch, sl.batch);
  FR_CUDA(cudaSetDevice(s->device));
  const int d = FR_NDIM_POSE + s->ks + s->ke;
  const size_t npix = (size_t)s->height * s->width;
  if (s->depth_grad == nullptr) {
    FR_CUDA(cudaMalloc(&s->depth_grad, sizeof(float) * (size_t)s->max_batch * npix));
    FR_CUDA(cudaMalloc(&s->vgrad, sizeof(float) * (size_t)s->max_batch * 3 * s->nver));
    FR_CUDA(cudaMalloc(&s->pgrad, sizeof(float) * (size_t)s->max_batch * d));
  }
  FR_CUDA(cudaMemcpyAsync(s->depth_grad, depth_grad, sizeof(float) * batch * npix, cudaMemcpyHostToDevice, sl.stream));
  if (int rc = fr_render_depth_backward(s->depth_grad, s->tri, sl.tri_ind, s->vgrad, batch, s->nver, s->ntri, s->height,
                                        s->width, sl.stream))
    return rc;
  if (int rc = fr_recon_project_backward(sl.params, s->packed, s->vgrad, s->pgrad, batch, s->nver, s->ks, s->ke, sl.im_size, s->flags,
                                         sl.ws, s->ws_bytes, sl.stream))
    return rc;
  FR_CUDA(cudaMemcpyAsync(params_grad, s->pgrad, sizeof(float) * (size_t)batch * d, cudaMemcpyDeviceToHost, sl.stream));
  FR_CUDA(cudaStreamSynchronize(sl.stream));
  return FR_OK;
}

}  // extern "C"

#ifdef FR_TIMELINE
extern "C" int fr_debug_timeline(long long* out) {   // developer build only (tools/timeline.py): clock64 stamps of the last forward kernel
  return cudaMemcpyFromSymbol(out, fr::f16::g_timeline, sizeof(long long) * 160 * 16) == cudaSuccess ? 0 : 1;
}
extern "C" int fr_debug_marks(unsigned long long* out, int reset) {   // [8] globaltimer marks (fr_common.cuh); reset: min slots = ~0, max slots = 0
  if (reset) {
    const unsigned long long init[8] = {~0ull, 0ull, ~0ull, 0ull, ~0ull, 0ull, ~0ull, 0ull};
    return cudaMemcpyToSymbol(fr::g_marks, init, sizeof(init)) == cudaSuccess ? 0 : 1;
  }
  return cudaMemcpyFromSymbol(out, fr::g_marks, sizeof(unsigned long long) * 8) == cudaSuccess ? 0 : 1;
}
extern "C" int fr_debug_cluster_cost(float* out) {   // [1024] cycles per octet of every cluster in the last fused kernel
  return cudaMemcpyFromSymbol(out, fr::f16::g_cluster_cost, sizeof(float) * 1024) == cudaSuccess ? 0 : 1;
}
#endif
