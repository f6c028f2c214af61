"""Geometry part of the reference's ``FaceRecNet`` (``nets/network.py``) on PyTorch + the CUDA library.

Only the methods on the params -> depth-map path are mirrored, with the reference's names and argument meaning:

=============================  =====================================  ==========================================
here                           reference                              what runs
=============================  =====================================  ==========================================
``vertices_transform``         ``nets/network.py:140-171``            ``fr_recon_project_forward/backward`` (CUDA)
``rendering_layer``            ``nets/network.py:174-201``            ``fr_rendering_layer_forward/backward`` (CUDA)
``depth_rendering_layer``      ``nets/network.py:300-308``            both of the above
``set_constraints``            ``nets/network.py:204-218``            torch elementwise; fused: ``vertices_transform_raw``
``geometry_loss``              ``nets/network.py:346-355``            Gram-matrix form (no basis pass)
``parse_pose_params``          ``nets/network.py:253-263``            slicing
``rotation_matrix(_batch)``    ``nets/network.py:266-297``            numpy, host-side mirror for inspection only
=============================  =====================================  ==========================================

The CNN regressors, losses and checkpoint plumbing of ``FaceRecNet`` are out of scope.
"""
from __future__ import annotations

from math import cos, sin

import numpy as np
import torch

from .. import _lib
from .._lib import check, lib
from ..model import DeviceModel
from ..rendering_layer.ops import _mesh_handle, _workspace, render_depth


class _ReconProject(torch.autograd.Function):
    """params [B,d] -> vertex_proj [B,3,N]; gradient as TF autodiff gives it (no gradient to the three angles)."""

    @staticmethod
    def forward(ctx, params, model: DeviceModel, im_size: float, flags: int):
        if not params.is_cuda:
            raise RuntimeError("params is on %s: vertices_transform has no CPU path" % params.device)
        if params.dim() != 2 or params.shape[1] != model.ndim:
            raise ValueError("params must be [B, %d] (pose 7 | shape %d | expression %d)" %
                             (model.ndim, model.ndim_shape, model.ndim_exp))
        if params.device != model.device:
            raise ValueError("params is on %s but the model lives on %s" % (params.device, model.device))
        params = params.float().contiguous()
        B = int(params.shape[0])
        dev = params.device
        out = torch.empty((B, 3, model.nver), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            nbytes = lib().fr_recon_workspace_bytes(B, model.nver, model.ndim_shape, model.ndim_exp)
            ws = _workspace(dev, nbytes)
            check(lib().fr_recon_project_forward(params.data_ptr(), model.packed.data_ptr(), model.mesh.handle, out.data_ptr(), B, model.nver,
                                                 model.ndim_shape, model.ndim_exp, float(im_size), flags, ws.data_ptr(),
                                                 ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
        ctx.save_for_backward(params)
        ctx.model, ctx.flags, ctx.im_size = model, flags, float(im_size)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (params,) = ctx.saved_tensors
        model, flags = ctx.model, ctx.flags
        B = int(params.shape[0])
        dev = params.device
        grad_out = grad_out.float().contiguous()
        dparams = torch.empty_like(params)
        with torch.cuda.device(dev):
            nbytes = lib().fr_recon_workspace_bytes(B, model.nver, model.ndim_shape, model.ndim_exp)
            ws = _workspace(dev, nbytes)
            check(lib().fr_recon_project_backward(params.data_ptr(), model.packed.data_ptr(), grad_out.data_ptr(),
                                                  dparams.data_ptr(), B, model.nver, model.ndim_shape, model.ndim_exp,
                                                  ctx.im_size, flags, ws.data_ptr(), ws.numel(),
                                                  torch.cuda.current_stream(dev).cuda_stream))
        return dparams, None, None, None


class _RenderingLayer(torch.autograd.Function):
    """``FaceRecNet.rendering_layer`` (``nets/network.py:174-201``) with the four post-processing passes inside the resolve
    kernel (``fr_rendering_layer_forward``).  Gradients flow to ``vertex_proj`` from ``depthimg`` and ``maskimg`` only, as in
    the reference (the op registers no gradient for texture / normal, ``rendering_layer/ops.py:95``)."""

    @staticmethod
    def forward(ctx, vertex_proj, tri, texture, im_gray, height, width, mesh):
        if not vertex_proj.is_cuda:
            raise RuntimeError("vertex_proj is on %s: rendering_layer has no CPU path" % vertex_proj.device)
        if vertex_proj.dim() != 3 or vertex_proj.shape[1] != 3:
            raise ValueError("The vertex is not Batch x 3 x nver")                # render_depth_op.cc:411
        if tri.dim() != 2 or tri.shape[0] != 3:
            raise ValueError("The tri is not 3 x ntri")                           # :414
        ver, tri = vertex_proj.float().contiguous(), tri.float().contiguous()
        B, N, T = int(ver.shape[0]), int(ver.shape[2]), int(tri.shape[1])
        dev = ver.device
        if texture.dim() not in (2, 3) or texture.shape[-2] != 3 or texture.shape[-1] != N or (texture.dim() == 3 and texture.shape[0] != B):
            raise ValueError("colors must be [3,N] or [B,3,N] like vertex_proj (got %s)" % (tuple(texture.shape),))
        if im_gray is not None and tuple(im_gray.shape) != (B, height, width, 1):
            raise ValueError("im_gray must be [B,%d,%d,1] (got %s)" % (height, width, tuple(im_gray.shape)))
        if not tri.is_cuda or not texture.is_cuda or (im_gray is not None and not im_gray.is_cuda):
            raise RuntimeError("rendering_layer has no CPU path: tri, colors and im_gray must be CUDA tensors")
        if texture.dim() == 2 or (texture.stride(0) == 0 and texture[0].is_contiguous()):
            tex = (texture if texture.dim() == 2 else texture[0]).float().contiguous()
            tex_ptr, tex_stride, keep = tex.data_ptr(), 0, tex
        else:
            tex = texture.float().contiguous()
            tex_ptr, tex_stride, keep = tex.data_ptr(), 3 * N, tex
        gray = None if im_gray is None else im_gray.float().contiguous()
        new = lambda c: torch.empty((B, height, width, c), dtype=torch.float32, device=dev)
        pncc, normalimg, maskimg, depthimg, raw, tri_ind = new(3), new(3), new(1), new(1), new(1), new(1)
        mesh_h = _mesh_handle(tri, N, ver, mesh)
        with torch.cuda.device(dev):
            ws = _workspace(dev, lib().fr_render_workspace_bytes(B, N, height, width))
            check(lib().fr_rendering_layer_forward(ver.data_ptr(), tri.data_ptr(), tex_ptr, tex_stride,
                                                   None if gray is None else gray.data_ptr(), pncc.data_ptr(), normalimg.data_ptr(),
                                                   maskimg.data_ptr(), depthimg.data_ptr(), raw.data_ptr(), tri_ind.data_ptr(), B, N, T,
                                                   height, width, mesh_h, ws.data_ptr(), ws.numel(),
                                                   torch.cuda.current_stream(dev).cuda_stream))
        del keep
        ctx.save_for_backward(tri, tri_ind, raw, gray if gray is not None else raw.new_empty(0))
        ctx.dims = (B, N, T, height, width)
        ctx.mark_non_differentiable(pncc, normalimg)
        return pncc, normalimg, maskimg, depthimg

    @staticmethod
    def backward(ctx, _g_pncc, _g_normal, g_mask, g_depth):
        tri, tri_ind, raw, gray = ctx.saved_tensors
        B, N, T, H, W = ctx.dims
        dev = raw.device
        g_mask = None if g_mask is None else g_mask.float().contiguous()
        g_depth = None if g_depth is None else g_depth.float().contiguous()
        vertex_grad = torch.empty((B, 3, N), dtype=torch.float32, device=dev)
        if g_mask is None and g_depth is None:
            return vertex_grad.zero_(), None, None, None, None, None, None
        with torch.cuda.device(dev):
            check(lib().fr_rendering_layer_backward(None if g_depth is None else g_depth.data_ptr(),
                                                    None if g_mask is None else g_mask.data_ptr(),
                                                    gray.data_ptr() if gray.numel() else None, raw.data_ptr(), tri.data_ptr(),
                                                    tri_ind.data_ptr(), vertex_grad.data_ptr(), B, N, T, H, W,
                                                    torch.cuda.current_stream(dev).cuda_stream))
        return vertex_grad, None, None, None, None, None, None


class _ReconRenderDepth(torch.autograd.Function):
    """params [B,d] -> (vertex_proj, depth, texture_image, normal, tri_ind) in ONE library call (``fr_recon_render_forward_all``):
    ``vertices_transform`` followed by ``render_depth`` (``nets/network.py:300-308``) without the repack pass and -- with a
    cluster-tile model -- without a separate visibility kernel.  Bit-identical to the two separate ops; gradients as they compose
    (``fr_render_depth_backward`` into ``fr_recon_project_backward``; none for texture / normal, ``rendering_layer/ops.py:95``)."""

    @staticmethod
    def forward(ctx, params, model: DeviceModel, texture, height, width, im_size, flags):
        if not params.is_cuda:
            raise RuntimeError("params is on %s: there is no CPU path" % params.device)
        if params.dim() != 2 or params.shape[1] != model.ndim:
            raise ValueError("params must be [B, %d] (pose 7 | shape %d | expression %d)" % (model.ndim, model.ndim_shape, model.ndim_exp))
        if params.device != model.device:
            raise ValueError("params is on %s but the model lives on %s" % (params.device, model.device))
        params = params.float().contiguous()
        B, N, dev = int(params.shape[0]), model.nver, params.device
        if texture.dim() not in (2, 3) or texture.shape[-2] != 3 or texture.shape[-1] != N or (texture.dim() == 3 and texture.shape[0] != B):
            raise ValueError("colors must be [3,N] or [B,3,N] (got %s)" % (tuple(texture.shape),))
        if texture.dim() == 2 or (texture.stride(0) == 0 and texture[0].is_contiguous()):
            tex, tex_stride = (texture if texture.dim() == 2 else texture[0]).float().contiguous(), 0
        else:
            tex, tex_stride = texture.float().contiguous(), 3 * N
        vertex = torch.empty((B, 3, N), dtype=torch.float32, device=dev)
        new = lambda c: torch.empty((B, height, width, c), dtype=torch.float32, device=dev)
        depth, teximg, normal, tri_ind = new(1), new(3), new(3), new(1)
        with torch.cuda.device(dev):
            ws = _workspace(dev, lib().fr_pipeline_workspace_bytes(B, N, model.ndim_shape, model.ndim_exp, height, width))
            check(lib().fr_recon_render_forward_all(params.data_ptr(), model.packed.data_ptr(), model.tri.data_ptr(), model.mesh.handle,
                                                    tex.data_ptr(), tex_stride, vertex.data_ptr(), depth.data_ptr(), teximg.data_ptr(),
                                                    normal.data_ptr(), tri_ind.data_ptr(), B, N, model.ntri, model.ndim_shape,
                                                    model.ndim_exp, height, width, float(im_size), flags, ws.data_ptr(), ws.numel(),
                                                    torch.cuda.current_stream(dev).cuda_stream))
        ctx.save_for_backward(params, tri_ind)
        ctx.model, ctx.flags, ctx.im_size, ctx.hw = model, flags, float(im_size), (height, width)
        ctx.mark_non_differentiable(teximg, normal, tri_ind)
        ctx.set_materialize_grads(False)                 # an unused vertex_proj output must not cost a [B,3,N] tensor of zeros
        return vertex, depth, teximg, normal, tri_ind

    @staticmethod
    def backward(ctx, g_vertex, g_depth, _g_tex, _g_normal, _g_tri):
        params, tri_ind = ctx.saved_tensors
        model, (H, W) = ctx.model, ctx.hw
        B, N, dev = int(params.shape[0]), model.nver, params.device
        vertex_grad = torch.empty((B, 3, N), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            sp = torch.cuda.current_stream(dev).cuda_stream
            if g_depth is None and g_vertex is None:
                return None, None, None, None, None, None, None
            if g_depth is None:
                vertex_grad.zero_()
            else:
                check(lib().fr_render_depth_backward(g_depth.float().contiguous().data_ptr(), model.tri.data_ptr(), tri_ind.data_ptr(),
                                                     vertex_grad.data_ptr(), B, N, model.ntri, H, W, sp))
            if g_vertex is not None:
                vertex_grad += g_vertex.float()
            dparams = torch.empty_like(params)
            ws = _workspace(dev, lib().fr_recon_workspace_bytes(B, N, model.ndim_shape, model.ndim_exp))
            check(lib().fr_recon_project_backward(params.data_ptr(), model.packed.data_ptr(), vertex_grad.data_ptr(), dparams.data_ptr(), B, N,
                                                  model.ndim_shape, model.ndim_exp, ctx.im_size, ctx.flags, ws.data_ptr(), ws.numel(), sp))
        return dparams, None, None, None, None, None, None


def recon_render_depth(params, model: DeviceModel, texture, height=200, width=200, im_size=200, flags=None, raw=False):
    """``vertices_transform`` + ``render_depth`` for a [B,d] parameter tensor in one call: returns (vertex_proj [B,3,N], depth,
    texture_image, normal, tri_ind) exactly as ``recon_project`` followed by ``render_depth`` does, gradients included."""
    flags = model.run_flags if flags is None else int(flags)
    return _ReconRenderDepth.apply(params, model, texture, int(height), int(width), float(im_size),
                                   flags | (_lib.FR_PARAMS_RAW if raw else 0))


def recon_project(params, model: DeviceModel, im_size=200, flags=None, raw=False):
    """Functional form of ``vertices_transform`` for a [B,d] parameter tensor.  ``raw=True``: ``params`` are the regressor's raw
    outputs and ``set_constraints`` (``nets/network.py:204-218``) is applied inside the prep kernels; the
    gradient then lands on the raw values."""
    flags = model.run_flags if flags is None else int(flags)
    return _ReconProject.apply(params, model, float(im_size), flags | (_lib.FR_PARAMS_RAW if raw else 0))


class FaceRecNet:
    """The geometry/rendering slice of the reference's ``FaceRecNet`` (constructor arguments as ``network.py:19``)."""

    def __init__(self, im_gray=None, params_label=None, mesh_data=None, nIter=4, batch_size=64, im_size=200,
                 device="cuda:0", convention="network"):
        self.im_gray = im_gray
        self.params_label = params_label
        self.nIter = nIter
        self.batch_size = batch_size
        self.im_size = im_size                  # kept as given (the y flip uses the float value, network.py:168)
        self._hw = int(im_size)                 # image height == width in pixels
        if self._hw != im_size or self._hw <= 0:
            raise ValueError("im_size must be a positive whole number of pixels (got %r)" % (im_size,))
        self.model = mesh_data if isinstance(mesh_data, DeviceModel) else DeviceModel(mesh_data, device, convention)
        self.vertex_code = self.model.vertex_code          # network.py:39
        self.tri = self.model.tri                           # network.py:40
        self.mu_tex = self.model.mu_tex
        self.ndim_shape, self.ndim_exp, self.ndim_pose = self.model.ndim_shape, self.model.ndim_exp, self.model.ndim_pose
        self.ndim = self.model.ndim
        self.nvert = self.model.nver
        # network.py:57-61: initial prediction = neutral face in the image centre
        init = torch.zeros((batch_size, self.ndim), dtype=torch.float32, device=self.model.device)
        init[:, 3] = im_size / 2.0
        init[:, 4] = im_size / 2.0
        init[:, 6] = 0.001
        self.pred_params = init[:, None, None, :]

    # ------------------------------------------------------------------ network.py:140-171
    def vertices_transform(self, pred_params):
        """pred_params [B,1,1,d] (or [B,d]) -> vertex_proj [B,3,N]."""
        p = pred_params
        if p.dim() == 4:
            p = p.squeeze(2).squeeze(1)                      # tf.squeeze(pred_params, [1, 2]), network.py:142
        return recon_project(p, self.model, self.im_size)

    def vertices_transform_raw(self, raw_params):
        """``vertices_transform(set_constraints(raw_params))`` with the constraints fused into the prep kernels."""
        p = raw_params.squeeze(2).squeeze(1) if raw_params.dim() == 4 else raw_params
        return recon_project(p, self.model, self.im_size, raw=True)

    # ------------------------------------------------------------------ network.py:174-201
    def rendering_layer(self, vertex_proj, triangles, colors):
        """network.py:174-201 in one kernel pass after the rasterizer; ``rendering_layer_unfused`` is the
        literal transcription it is tested against."""
        return _RenderingLayer.apply(vertex_proj, triangles, colors, self.im_gray, self._hw, self._hw, "auto")

    def rendering_layer_unfused(self, vertex_proj, triangles, colors):
        B = vertex_proj.shape[0]
        texture = colors if colors.dim() == 3 else colors.unsqueeze(0).expand(B, -1, -1)      # tf.tile, network.py:179
        im_gray = self.im_gray
        image = vertex_proj.new_empty((B, self._hw, self._hw, 3)) if im_gray is None else im_gray.expand(-1, -1, -1, 3)
        tf_depth, tf_tex, tf_normal, _ = render_depth(ver=vertex_proj, tri=triangles, texture=texture, image=image)
        pncc_batch = torch.clamp(tf_tex, 1e-6, 1.0)                                            # :185
        flip = (tf_normal[..., 2:3] < 0)                                                       # :188
        tf_normal = torch.where(flip, -1.0 * tf_normal, tf_normal)                             # :189
        mag = (tf_normal * tf_normal).sum(dim=-1)                                              # :190
        mag = torch.where(mag > 1e-6, mag, torch.ones_like(mag))                               # :191
        normalimg_batch = tf_normal / (torch.sqrt(mag) + 1e-6).unsqueeze(-1)                   # :192
        mask = torch.clamp(tf_depth, 1e-6, 1.0)                                                # :195
        maskimg_batch = mask * im_gray if im_gray is not None else mask                        # :196
        depthimg_batch = torch.clamp_min(tf_depth, 1e-6)                                       # :199
        return pncc_batch, normalimg_batch, maskimg_batch, depthimg_batch

    # ------------------------------------------------------------------ network.py:300-308
    def depth_rendering_layer(self):
        self.vertices_proj = self.vertices_transform(self.pred_params)
        self.pncc_batch, self.normal_batch, self.maskimg_batch, self.coarse_depth_map = \
            self.rendering_layer(self.vertices_proj, self.tri, self.vertex_code)
        return self.coarse_depth_map

    # ------------------------------------------------------------------ network.py:346-355
    def geometry_loss(self, pred_params, params_label):
        """``tf.losses.mean_squared_error(Basis . label, Basis . pred)`` over the 3N x B reconstructed coordinates, without
        touching the basis: with d = label - pred it equals  sum_b d_b^T (Basis^T Basis) d_b / (3N B);  the 228 x 228 Gram
        matrix is computed once at model load (``DeviceModel.gram``, float64)."""
        def geo(p):
            p = p.squeeze(2).squeeze(1) if p.dim() == 4 else p
            return p[:, self.ndim_pose:self.ndim_pose + self.ndim_shape + self.ndim_exp]
        d = (geo(params_label) - geo(pred_params)).double()
        quad = ((d @ self.model.gram) * d).sum()
        return (quad / (3.0 * self.model.nver * d.shape[0])).float()

    # ------------------------------------------------------------------ network.py:204-218
    def set_constraints(self, pred_params):
        s = torch.sigmoid(pred_params)
        kp, ks = self.ndim_pose, self.ndim_shape
        self.pred_params = torch.cat([s[..., 0:3] * 3.0 - 1.5,
                                      s[..., 3:5] * self.im_size,
                                      s[..., 5:6] * 0.0,
                                      s[..., 6:7] * 1e-3,
                                      s[..., kp:kp + ks] * 1e4,
                                      s[..., kp + ks:self.ndim] * 3.0 - 1.5], dim=-1)
        return self.pred_params

    # ------------------------------------------------------------------ network.py:253-263
    @staticmethod
    def parse_pose_params(pose_params):
        return (pose_params[:, 0:1], pose_params[:, 1:2], pose_params[:, 2:3], pose_params[:, 3:6], pose_params[:, 6:7])

    # ------------------------------------------------------------------ network.py:266-297 (host mirror; the device
    # computes the same matrices inside recon_prep_kernel, no py_func round trip)
    @staticmethod
    def rotation_matrix(angles):
        phi, gamma, theta = [float(a) for a in angles]
        r_pitch = np.array([[1, 0, 0], [0, cos(phi), sin(phi)], [0, -sin(phi), cos(phi)]])
        r_yaw = np.array([[cos(gamma), 0, -sin(gamma)], [0, 1, 0], [sin(gamma), 0, cos(gamma)]])
        r_roll = np.array([[cos(theta), sin(theta), 0], [-sin(theta), cos(theta), 0], [0, 0, 1]])
        return np.dot(np.dot(r_pitch, r_yaw), r_roll).astype(np.float32)

    @classmethod
    def rotation_matrix_batch(cls, angles_batch):
        return np.stack([cls.rotation_matrix(a) for a in np.asarray(angles_batch)]).astype(np.float32)
