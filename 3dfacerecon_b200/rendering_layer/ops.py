"""``render_depth`` -- drop-in for the reference's ``rendering_layer/ops.py`` on ``torch.Tensor``s.

Same name, argument order, layouts and dtypes as ``rendering_layer/ops.py:78-81``:

    depth, texture_image, normal, tri_ind = render_depth(ver, tri, texture, image)

* ``ver``     [B,3,N] float32 projected vertices (x = column, y = row, z = depth, larger = nearer)
* ``tri``     [3,T]   float32 vertex indices (0-based)
* ``texture`` [B,3,N] float32 per-vertex attribute (an ``expand``-ed [3,N] is used without materialising it)
* ``image``   [B,H,W,C] -- only its shape is used (``render_depth_op.cc:397-403``)

The gradient matches ``_RenderDepthGrad`` (``ops.py:86-95``): ``[d_ver, None, None, None]`` computed from
``depth_grad`` alone.  The kernels run on ``ver``'s device and the current CUDA stream; there is no CPU path --
a CPU tensor raises.
"""
from __future__ import annotations

import torch

from .. import mesh as _mesh
from .._lib import check, lib

OP_NAMES = ["render_depth"]          # rendering_layer/ops.py:13

_workspaces = {}
_MAX_WORKSPACES = 16


def _workspace(device, nbytes: int) -> torch.Tensor:
    """Scratch per (device, stream), grown on demand; the library itself never allocates (include/facerecon_b200.h).
    The cache is bounded: the least recently used entry goes when more than ``_MAX_WORKSPACES`` streams have asked, and
    ``release_workspaces()`` drops everything (e.g. after a one-off large batch)."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.pop(key, None)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
    _workspaces[key] = ws                       # (re-)inserted last: dict order is the LRU order
    while len(_workspaces) > _MAX_WORKSPACES:
        _workspaces.pop(next(iter(_workspaces)))
    return ws


def release_workspaces() -> None:
    _workspaces.clear()


def _mesh_handle(tri, nver, ver, mesh):
    """``mesh`` keyword of render_depth -> fr_mesh_table handle or None: "auto" (registry, csrc/mesh_table.h), a
    ``MeshTable``, or None / False for the generic kernels."""
    if mesh is None or mesh is False:
        return None
    if isinstance(mesh, _mesh.MeshTable):
        return mesh.handle
    table = _mesh.table_for(tri, nver, ver, build="now" if mesh == "now" else "auto")
    return None if table is None else table.handle


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s is on %s: render_depth has no CPU path, move it to a CUDA device" % (name, t.device))
    if t.dtype != torch.float32:
        t = t.float()                                  # tf.to_float at nets/network.py:177
    return t


class _RenderDepth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ver, tri, texture, image, mesh):
        ver, tri, texture = _f32c(ver, "ver"), _f32c(tri, "tri"), _f32c(texture, "texture")
        if ver.dim() != 3 or tri.dim() != 2 or texture.dim() != 3 or image.dim() != 4:
            raise ValueError("render_depth expects ver [B,3,N], tri [3,T], texture [B,3,N], image [B,H,W,C]")
        B, H, W = int(image.shape[0]), int(image.shape[1]), int(image.shape[2])
        # render_depth_op.cc:408-418 (same messages)
        if ver.shape[0] != B:
            raise ValueError("The vertex's batch is not the same as image batch")
        if ver.shape[1] != 3:
            raise ValueError("The vertex is not Batch x 3 x nver")
        if tri.shape[0] != 3:
            raise ValueError("The tri is not 3 x ntri")
        if texture.shape[1] != 3:
            raise ValueError("The texture channel must be equal to image channel namely 3")
        N, T = int(ver.shape[2]), int(tri.shape[1])
        if texture.shape[0] != B or texture.shape[2] != N:
            raise ValueError("texture must be [B,3,N] like ver")
        dev = ver.device
        ver, tri = ver.contiguous(), tri.contiguous()
        mesh_h = _mesh_handle(tri, N, ver, mesh)
        if texture.stride(0) == 0 and texture[0].is_contiguous():
            tex_ptr, tex_stride = texture.data_ptr(), 0          # tiled texture (network.py:179) without the copy
        else:
            texture = texture.contiguous()
            tex_ptr, tex_stride = texture.data_ptr(), 3 * N
        depth = torch.empty((B, H, W, 1), dtype=torch.float32, device=dev)
        texture_image = torch.empty((B, H, W, 3), dtype=torch.float32, device=dev)
        normal = torch.empty((B, H, W, 3), dtype=torch.float32, device=dev)
        tri_ind = torch.empty((B, H, W, 1), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            nbytes = lib().fr_render_workspace_bytes(B, N, H, W)
            ws = _workspace(dev, nbytes)
            check(lib().fr_render_depth_forward(ver.data_ptr(), tri.data_ptr(), tex_ptr, tex_stride, depth.data_ptr(),
                                                texture_image.data_ptr(), normal.data_ptr(), tri_ind.data_ptr(), B, N, T,
                                                H, W, mesh_h, ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
        ctx.save_for_backward(tri, tri_ind)
        ctx.dims = (B, N, T, H, W)
        ctx.mark_non_differentiable(texture_image, normal, tri_ind)   # ops.py:95: only `ver` gets a gradient
        return depth, texture_image, normal, tri_ind

    @staticmethod
    def backward(ctx, depth_grad, *_unused):
        tri, tri_ind = ctx.saved_tensors
        B, N, T, H, W = ctx.dims
        dev = tri_ind.device
        depth_grad = _f32c(depth_grad, "depth_grad").contiguous()
        vertex_grad = torch.empty((B, 3, N), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib().fr_render_depth_backward(depth_grad.data_ptr(), tri.data_ptr(), tri_ind.data_ptr(),
                                                 vertex_grad.data_ptr(), B, N, T, H, W,
                                                 torch.cuda.current_stream(dev).cuda_stream))
        return vertex_grad, None, None, None, None


def render_depth(ver, tri, texture, image, mesh="auto", **kwargs):
    """``rendering_layer/ops.py:78-81``.  Extra keyword arguments (TF's ``name=``) are accepted and ignored.

    ``mesh`` selects the rasterizer's side table for ``tri`` (same outputs either way): "auto" (default) uses the table
    registered for this ``tri`` tensor -- ``DeviceModel`` registers one at model load, any other tensor gets one the
    second time it is seen --, "now" builds one immediately, a ``MeshTable`` is used as given, None = generic kernels."""
    return _RenderDepth.apply(ver, tri, texture, image, mesh)


def render_depth_grad(depth_grad, vertex, tri, depth, tri_ind, image):
    """The reference's second op, ``RenderDepthGrad`` (``ops.py:89-92``, ``render_depth_op.cc:571-589``), callable directly.
    ``vertex``, ``depth`` and ``image`` only donate shapes (the reference never reads their values, ``.cc:329,349-353``)."""
    depth_grad, tri, tri_ind = _f32c(depth_grad, "depth_grad").contiguous(), _f32c(tri, "tri").contiguous(), _f32c(tri_ind, "tri_ind").contiguous()
    B, H, W = int(image.shape[0]), int(image.shape[1]), int(image.shape[2])
    if vertex.shape[0] != B:
        raise ValueError("The vertex's batch is not the same as image batch")
    if vertex.shape[1] != 3:
        raise ValueError("The vertex is not Batch x 3 x nver")
    if tri.shape[0] != 3:
        raise ValueError("The tri is not 3 x ntri")
    N, T = int(vertex.shape[2]), int(tri.shape[1])
    dev = depth_grad.device
    vertex_grad = torch.empty((B, 3, N), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib().fr_render_depth_backward(depth_grad.data_ptr(), tri.data_ptr(), tri_ind.data_ptr(), vertex_grad.data_ptr(),
                                             B, N, T, H, W, torch.cuda.current_stream(dev).cuda_stream))
    return vertex_grad
