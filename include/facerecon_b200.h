/* facerecon_b200 -- C ABI of the B200-native 3DMM reconstruction + depth-rendering hot path.
 *
 * This library replaces the TensorFlow custom-op shared object the reference builds from
 * rendering_layer/ops_src/ (rendering_layer/ops.py:23-72 compiles and loads render_depth_op.so)
 * and the in-graph geometry of nets/network.py:140-171.  Each entry point below cites the
 * reference interface it stands in for.  There is no CPU fallback: every compute entry point
 * launches CUDA kernels built for sm_100a on the caller's stream.
 *
 * Conventions
 *   - All data pointers are DEVICE pointers owned by the caller (PyTorch / the framework's
 *     allocator); the library never allocates on the hot path (the reference's GPU op calls
 *     cudaMalloc/cudaFree six times per launch, render_depth_op.cu.cc:272-277,335-340).
 *   - Calls are stream-ordered on `stream` (a cudaStream_t passed as void*), never synchronise,
 *     and are CUDA-graph capturable (the reference launches on the legacy default stream,
 *     render_depth_op.cu.cc:284,302,321).
 *   - Return value: FR_OK or an FR_ERR_* code; fr_last_error() returns a thread-local message
 *     (the reference printf()s and returns with uninitialised outputs, render_depth_op.cc:161-172).
 *   - Tensors are dense row-major float32 with the reference's layouts.
 *
 * The fr_session_* group at the end is the HOST-buffer flavour (inputs/outputs in host memory,
 * copies done by the library) for callers that are not on the GPU already.
 */
#ifndef FACERECON_B200_H_
#define FACERECON_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FR_VERSION 202

/* status codes */
#define FR_OK 0
#define FR_ERR_INVALID_ARGUMENT 1 /* shape rule violated (TF InvalidArgument, render_depth_op.cc:408-418) */
#define FR_ERR_CUDA 2             /* a CUDA runtime call or launch failed */
#define FR_ERR_WORKSPACE 3        /* workspace pointer null / too small / misaligned */
#define FR_ERR_UNSUPPORTED 4

/* convention flags.  Run-time flags of fr_recon_project_*: */
#define FR_ROT_XYZ 0x0u           /* R = Rx.Ry.Rz   nets/network.py:288 (default) */
#define FR_ROT_ZYX 0x1u           /* R = Rz.Ry.Rx   rendering_layer/sample_test.py:71 */
#define FR_YFLIP_S_Y_1 0x0u       /* y' = S - y - 1 nets/network.py:168 (default) */
#define FR_YFLIP_S_Y 0x2u         /* y' = S - y     rendering_layer/sample_test.py:105 */
#define FR_YFLIP_NONE 0x4u        /* no flip        prepare_data/Project2D.m:12 */
#define FR_PARAMS_RAW 0x100u      /* params are the regressor's raw outputs: FaceRecNet.set_constraints (nets/network.py:204-218:
                                   * sigmoid, then angles*3-1.5 | t_xy*im_size | t_z*0 | f*1e-3 | shape*1e4 | exp*3-1.5) is applied
                                   * inside the prep kernels, and the backward returns the gradient w.r.t. the raw values */
/* Pack-time flags of fr_pack_basis (memory layout of the 3N-long axis): */
#define FR_MEAN_PLANAR 0x0u       /* mu[c*N+n]      nets/network.py:157 (default) */
#define FR_MEAN_INTERLEAVED 0x10u /* mu[3*n+c]      rendering_layer/sample_test.py:101 */
#define FR_BASIS_PLANAR 0x0u      /* pc[c*N+n,k]    nets/network.py:154,156 (default) */
#define FR_BASIS_INTERLEAVED 0x20u/* pc[3*n+c,k]    prepare_data/Project2D.m:8-9 */
/* Pack-time AND run-time flag.  At pack time: the packed basis also carries the tensor-core forward operands with one row
 * tile per CLUSTER of the mesh table (+20 % of that section: border vertices are repeated in every member cluster).  At run
 * time: fr_recon_render_forward streams that section and rasterizes every cluster inside the reconstruction epilogue from
 * shared memory (raster_tile.cuh), so the projected vertices of the fused call never touch global memory and the rasterizer
 * runs in the issue slots the HBM-bound basis stream leaves idle.  Needs a basis packed with the flag and the same mesh
 * table; every other entry point ignores the flag.  Same results either way (bit-identical depth maps). */
#define FR_CLUSTER_TILES 0x40u

/* number of pose parameters in front of the shape/expression coefficients (utils/parser_3dmm.py:49) */
#define FR_NDIM_POSE 7

const char* fr_last_error(void);
int fr_version(void);

/* ---- mesh table: the triangle list of a model, clustered once for the rasterizer ------------------
 * The reference walks `tri` [3,ntri] triangle by triangle, converting three float indices and gathering nine vertex
 * floats per triangle and face (render_depth_op.cc:204-213; render_depth_op.cu.cc:84-92).  A mesh table partitions the
 * triangles ONCE per model (host side) into clusters of <= 128 unique vertices / <= 256 triangles with pre-validated
 * 8-bit local indices; the rasterizer then stages a cluster's vertices in shared memory once per face, and the
 * tensor-core reconstruction uses the clusters as its row tiles, so that in the fused params -> depth-map call the
 * vertices never pass through global memory (3dfacerecon_b200/csrc/mesh_table.h, raster_tile.cuh).
 *   tri        HOST pointer, [3,ntri] float 0-based indices as rendering_layer/ops.py:78 takes them; triangles with an
 *              index outside [0,nver) are dropped (the reference reads out of bounds)
 *   positions  HOST pointer or NULL: any vertex positions that reflect the mesh's locality -- the mean shape `mu`
 *              ([3,nver] planar, or [nver,3] with positions_interleaved != 0) or one face of a vertex tensor; only the
 *              quality of the partition depends on them, never a result
 *   device     CUDA device that receives a copy of the table, or -1 for a host-only table (inspection, tests)
 * The table replaces nothing in the results: every entry point below gives bit-identical outputs with and without it. */
typedef struct fr_mesh_table fr_mesh_table;
int fr_mesh_table_create(const float* tri, int ntri, int nver, const float* positions, int positions_interleaved, int device,
                         fr_mesh_table** out);
/* Re-creates a table from the bytes fr_mesh_table_blob returned earlier (an on-disk cache); validates header and hash. */
int fr_mesh_table_from_blob(const void* blob, size_t bytes, int device, fr_mesh_table** out);
void fr_mesh_table_destroy(fr_mesh_table* mesh);
const void* fr_mesh_table_blob(const fr_mesh_table* mesh, size_t* bytes);   /* host copy (layout: mesh_table.h) */
int fr_mesh_table_clusters(const fr_mesh_table* mesh);
int fr_mesh_table_vertex_slots(const fr_mesh_table* mesh);   /* vertices counted once per member cluster */

/* ---- model packing: utils/parser_3dmm.py dict -> one device buffer ------------------------------
 * Packs [pc_shape | pc_exp | mu | 0-pad] into the layouts the kernels stream (DESIGN.md "Packed basis"):
 * an fp32 float4-tiled section (FFMA kernels), the column scales, fp16 hi/lo tcgen05 operand tiles of the
 * column-scaled basis transposed for the backward contraction, an fp32 copy of the mean, and the fp16 hi/lo
 * operand tiles of the forward pass with ONE ROW TILE PER CLUSTER of `mesh` (NULL: tiles of 128 consecutive
 * vertices).  mu [3N], pc_shape [3N,ndim_shape], pc_exp [3N,ndim_exp] are device pointers in the reference's layouts
 * (nets/network.py:41-43).  One-off, at model load.  Every later call that takes this packed basis must be given the
 * same `mesh` (or NULL if it was packed with NULL). */
size_t fr_packed_basis_bytes(int nver, int ndim_shape, int ndim_exp, unsigned layout_flags, const fr_mesh_table* mesh);
int fr_pack_basis(const float* mu, const float* pc_shape, const float* pc_exp, int nver, int ndim_shape, int ndim_exp,
                  unsigned layout_flags, const fr_mesh_table* mesh, float* packed, void* stream);

/* ---- FaceRecNet.vertices_transform (nets/network.py:140-171) ------------------------------------
 * params [batch, 7+ndim_shape+ndim_exp] (layout nets/network.py:143-145,258-262) -> vertex_proj [batch,3,nver].
 * Rotation (network.py:266-297, a host py_func in the reference) is computed on the device. */
size_t fr_recon_workspace_bytes(int batch, int nver, int ndim_shape, int ndim_exp);
int fr_recon_project_forward(const float* params, const float* packed, const fr_mesh_table* mesh, float* vertex_proj, int batch,
                             int nver, int ndim_shape, int ndim_exp, float im_size, unsigned flags, void* workspace,
                             size_t workspace_bytes, void* stream);

/* Gradient of the above as TF autodiff produces it: vertex_grad [batch,3,nver]
 * (gradient w.r.t. vertex_proj) -> params_grad [batch, 7+ndim_shape+ndim_exp]; the three angle entries
 * are 0 because tf.py_func (network.py:150) has no gradient.  im_size only matters with FR_PARAMS_RAW. */
int fr_recon_project_backward(const float* params, const float* packed, const float* vertex_grad, float* params_grad,
                              int batch, int nver, int ndim_shape, int ndim_exp, float im_size, unsigned flags, void* workspace,
                              size_t workspace_bytes, void* stream);

/* ---- TF op "RenderDepth" (render_depth_op.cc:378-458, :535-569; functor :132-322) ---------------
 * vertex [batch,3,nver], tri [3,ntri] FLOAT 0-based indices, texture [batch,3,nver] addressed with
 * texture_batch_stride floats between faces (0 = one [3,nver] texture shared by all faces, what
 * network.py:179 materialises with tf.tile).  (batch,height,width) are what the reference takes from
 * its `image` input (:397-403), whose values it never reads.
 * Outputs: depth [batch,H,W,1], texture_image [batch,H,W,3], normal [batch,H,W,3], tri_ind [batch,H,W,1]
 * (float, -1 = background).  texture_image and normal may be NULL to skip them (texture may then be NULL).
 * Triangles whose indices fall outside [0,nver) are skipped (the reference reads out of bounds).
 * mesh: the mesh table of `tri` (shared-memory tile rasterizer, raster_tile.cuh) or NULL (generic per-triangle gather kernel); same outputs either way.
 * At most 65535 faces per call. */
size_t fr_render_workspace_bytes(int batch, int nver, int height, int width);
int fr_render_depth_forward(const float* vertex, const float* tri, const float* texture, long long texture_batch_stride,
                            float* depth, float* texture_image, float* normal, float* tri_ind, int batch, int nver,
                            int ntri, int height, int width, const fr_mesh_table* mesh, void* workspace, size_t workspace_bytes,
                            void* stream);

/* ---- TF op "RenderDepthGrad" (render_depth_op.cc:470-528, :571-589; functor :325-368) -----------
 * depth_grad [batch,H,W,1], tri [3,ntri], tri_ind [batch,H,W,1] -> vertex_grad [batch,3,nver], fully
 * written: rows 0,1 are zero, row 2 receives (g*1.0f)/3.0f per covered pixel and vertex.  Unlike the
 * reference the output is zero-filled and background pixels are skipped.
 * The reference's `vertex`, `depth` and `image` inputs are not needed (never read, :329,349-353). */
int fr_render_depth_backward(const float* depth_grad, const float* tri, const float* tri_ind, float* vertex_grad,
                             int batch, int nver, int ntri, int height, int width, void* stream);

/* ---- FaceRecNet.rendering_layer (nets/network.py:174-201): render_depth + its four elementwise post-processing passes
 * in the resolve kernel.  pncc [B,H,W,3] = clip(texture_image, 1e-6, 1); normalimg [B,H,W,3] = normals flipped
 * to +z and divided by sqrt(mag)+1e-6 (mag <= 1e-6 -> 1); maskimg [B,H,W,1] = clip(depth, 1e-6, 1) * im_gray (im_gray NULL:
 * the mask alone); depthimg [B,H,W,1] = max(depth, 1e-6); raw_depth [B,H,W,1] = the op's depth (optional, needed by the
 * backward's gates).  Same workspace as fr_render_depth_forward. */
int fr_rendering_layer_forward(const float* vertex, const float* tri, const float* texture, long long texture_batch_stride,
                               const float* im_gray, float* pncc, float* normalimg, float* maskimg, float* depthimg,
                               float* raw_depth, float* tri_ind, int batch, int nver, int ntri, int height, int width,
                               const fr_mesh_table* mesh, void* workspace, size_t workspace_bytes, void* stream);
/* Its gradient as autodiff composes it: depth_grad of the op = depthimg_grad where depth >= 1e-6, plus maskimg_grad * im_gray
 * where 1e-6 <= depth <= 1 (either gradient may be NULL), then RenderDepthGrad. */
int fr_rendering_layer_backward(const float* depthimg_grad, const float* maskimg_grad, const float* im_gray, const float* raw_depth,
                                const float* tri, const float* tri_ind, float* vertex_grad, int batch, int nver, int ntri,
                                int height, int width, void* stream);

/* ---- fused: params -> depth map (the north-star path in one call) -------------------------------
 * Same results as fr_recon_project_forward followed by fr_render_depth_forward (depth + tri_ind only), in four kernels:
 * parameter prep (which also clears the visibility keys), the tensor-core reconstruction whose epilogue writes the
 * rasterizer's 16-byte vertex records (by vertex rank when a mesh table is given), the visibility pass and the resolve
 * pass -- the repack pass over the vertex tensor disappears.  vertex_proj [batch,3,nver] is optional (NULL = do not
 * materialise it).  With FR_CLUSTER_TILES (and more than 8 faces) the reconstruction epilogue instead rasterizes each cluster
 * of the mesh table from shared memory: no records, three kernels.
 * stage_events: NULL, or three cudaEvent_t (any may be NULL) recorded on `stream` after the reconstruction kernels, after
 * the visibility kernel and after the resolve kernel -- lets a benchmark split the device time of one real call (the
 * kernels then launch fully serialised instead of programmatically dependent). */
size_t fr_pipeline_workspace_bytes(int batch, int nver, int ndim_shape, int ndim_exp, int height, int width);
int fr_recon_render_forward(const float* params, const float* packed, const float* tri, const fr_mesh_table* mesh,
                            float* vertex_proj, float* depth, float* tri_ind, int batch, int nver, int ntri, int ndim_shape,
                            int ndim_exp, int height, int width, float im_size, unsigned flags, void* workspace,
                            size_t workspace_bytes, void* stream, void* const* stage_events);

/* The same call with ALL FOUR outputs of render_depth (rendering_layer/ops.py:78-81) plus vertex_proj: what the training path
 * (nets/network.py:300-308, depth_rendering_layer) needs from one batch of parameters.  Bit-identical to
 * fr_recon_project_forward + fr_render_depth_forward; with FR_CLUSTER_TILES the rasterizing epilogue also leaves the 16-byte
 * vertex records the resolve pass gathers normals from, so neither the repack pass nor the visibility kernel runs.
 * texture: [3,nver] (texture_batch_stride 0) or per face, as in fr_render_depth_forward.  Gradients: fr_render_depth_backward
 * followed by fr_recon_project_backward. */
int fr_recon_render_forward_all(const float* params, const float* packed, const float* tri, const fr_mesh_table* mesh,
                                const float* texture, long long texture_batch_stride, float* vertex_proj, float* depth,
                                float* texture_image, float* normal, float* tri_ind, int batch, int nver, int ntri, int ndim_shape,
                                int ndim_exp, int height, int width, float im_size, unsigned flags, void* workspace,
                                size_t workspace_bytes, void* stream);

/* ---- host-buffer session (what a non-GPU caller binds) ----------------------
 * A session owns the device copy of the model, device staging for `max_batch` faces and one stream.
 * Every pointer below is a HOST pointer (pinned memory makes the copies asynchronous). */
typedef struct fr_session fr_session;
int fr_session_create(const float* mu, const float* pc_shape, const float* pc_exp, const float* tri, int nver, int ntri,
                      int ndim_shape, int ndim_exp, int height, int width, int max_batch, unsigned flags, int device,
                      fr_session** out);
void fr_session_destroy(fr_session* s);
/* params [batch,d] -> depth [batch,H,W,1] and (optional, may be NULL) tri_ind [batch,H,W,1], vertex_proj [batch,3,nver].
 * Copies in, runs recon + projection + render, copies out, and waits for completion. */
int fr_session_forward(fr_session* s, const float* params, int batch, float im_size, float* depth, float* tri_ind,
                       float* vertex_proj);
/* Pipelined flavour: a session has FR_SESSION_SLOTS independent slots (device staging + stream each).
 * fr_session_submit enqueues copy-in, recon + projection + render and the copies out of one batch on `slot` and returns
 * without waiting; fr_session_wait blocks until that batch's outputs are in the host buffers.  Alternating the slots
 * overlaps the device->host copy of one batch with the kernels of the next (pinned host buffers required for the
 * overlap; the host buffers of a slot must stay untouched until its wait).  fr_session_forward == submit + wait on slot 0. */
#define FR_SESSION_SLOTS 3
int fr_session_submit(fr_session* s, int slot, const float* params, int batch, float im_size, float* depth, float* tri_ind,
                      float* vertex_proj);
int fr_session_wait(fr_session* s, int slot);
/* depth_grad [batch,H,W,1] for the faces of the last batch run on slot 0 (fr_session_forward) -> params_grad [batch,d]. */
int fr_session_backward(fr_session* s, const float* depth_grad, int batch, float* params_grad);
/* counters: kernels launched by this library since load (for bench.py's gpu_launches) */
unsigned long long fr_launch_count(void);
#ifdef __cplusplus
}
#endif
#endif /* FACERECON_B200_H_ */
