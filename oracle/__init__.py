"""TEST INFRASTRUCTURE ONLY -- CPU checkers for the CUDA path; never imported by the product.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference arm
may import this package.  It exposes two independent CPU implementations of the reference's
z-buffer op (``/root/reference/rendering_layer/ops_src/render_depth_op.cc``):

* ``ref_*``    -- the UNMODIFIED reference translation unit, compiled where it lies against the
                  stub headers in ``oracle/tf_shim`` into ``oracle/_ref/libref_render_depth.so``
                  and driven through the reference's own ``OpKernel::Compute`` methods.
* ``oracle_*`` -- our plain-C restatement (``oracle/render_depth_oracle.c``), pinned against the
                  former by ``tests/test_oracle_vs_ref.py`` and ``tests/golden/``.

plus ``oracle.recon`` (numpy float64 restatement of ``nets/network.py:140-171`` and
``rendering_layer/sample_test.py:95-105``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "liboracle_render_depth.so")
_REF_SO = os.path.join(_HERE, "_ref", "libref_render_depth.so")
REFERENCE_ROOT = os.environ.get("FR_REFERENCE_ROOT", "/root/reference")

_f32p = ctypes.POINTER(ctypes.c_float)
_i64p = ctypes.POINTER(ctypes.c_longlong)

#: float value the reference initialises the depth buffer with (render_depth_op.cc:187)
BACKGROUND_DEPTH = np.float32(-99999999999999)


def build(force: bool = False) -> None:
    """Compile the C restatement and, when the reference tree is present, ``oracle/_ref``."""
    have_ref_src = os.path.exists(os.path.join(REFERENCE_ROOT, "rendering_layer", "ops_src", "render_depth_op.cc"))
    need = force or not os.path.exists(_ORACLE_SO) or (have_ref_src and not os.path.exists(_REF_SO))
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("render_depth_oracle.c", "ref_harness.cc", "Makefile"))
    if os.path.exists(_ORACLE_SO) and os.path.getmtime(_ORACLE_SO) < src_m:
        need = True
    if not need:
        return
    cmd = ["make", "-C", _HERE, "REF=" + REFERENCE_ROOT]
    if force:
        cmd.insert(1, "-B")
    subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


def _as_f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_f32p)


def _dims(shape):
    return (ctypes.c_longlong * len(shape))(*[int(s) for s in shape])


# --------------------------------------------------------------------------- restatement
_oracle_lib = None


def _oracle():
    global _oracle_lib
    if _oracle_lib is None:
        build()
        lib = ctypes.CDLL(_ORACLE_SO)
        lib.fr_oracle_render_depth_forward.restype = ctypes.c_int
        lib.fr_oracle_render_depth_forward.argtypes = [_f32p, _f32p, _f32p, ctypes.c_long] + [ctypes.c_int] * 6 + [_f32p] * 4
        lib.fr_oracle_render_depth_backward.restype = ctypes.c_int
        lib.fr_oracle_render_depth_backward.argtypes = [_f32p, _f32p, _f32p] + [ctypes.c_int] * 5 + [_f32p]
        _oracle_lib = lib
    return _oracle_lib


def oracle_render_depth_forward(vertex, tri, texture, height: int, width: int):
    """C restatement of ``RenderDepth(CPUDevice)`` (render_depth_op.cc:132-322).

    vertex [B,3,N], tri [3,T] float, texture [B,3,N] or [3,N] (shared by all faces).
    Returns depth [B,H,W,1], texture_image [B,H,W,3], normal [B,H,W,3], tri_ind [B,H,W,1].
    """
    vertex, tri, texture = _as_f32(vertex), _as_f32(tri), _as_f32(texture)
    B, three, N = vertex.shape
    assert three == 3 and tri.shape[0] == 3
    T = tri.shape[1]
    if texture.ndim == 2:
        ch, stride = texture.shape[0], 0
    else:
        assert texture.shape[0] == B
        ch, stride = texture.shape[1], texture.shape[1] * texture.shape[2]
    assert texture.shape[-1] == N
    depth = np.empty((B, height, width, 1), np.float32)
    teximg = np.empty((B, height, width, ch), np.float32)
    normal = np.empty((B, height, width, 3), np.float32)
    tri_ind = np.empty((B, height, width, 1), np.float32)
    rc = _oracle().fr_oracle_render_depth_forward(_ptr(vertex), _ptr(tri), _ptr(texture), stride, B, N, T, height, width,
                                                  ch, _ptr(depth), _ptr(teximg), _ptr(normal), _ptr(tri_ind))
    if rc != 0:
        raise RuntimeError("oracle forward failed (ntri >= 10M)")
    return depth, teximg, normal, tri_ind


def oracle_render_depth_backward(depth_grad, tri, tri_ind, nver: int):
    """C restatement of ``RenderDepthGrad(CPUDevice)`` (render_depth_op.cc:325-368) with the two
    documented fixes (zero-fill, skip background).  Returns vertex_grad [B,3,N]."""
    depth_grad, tri, tri_ind = _as_f32(depth_grad), _as_f32(tri), _as_f32(tri_ind)
    B, H, W = depth_grad.shape[:3]
    T = tri.shape[1]
    out = np.empty((B, 3, nver), np.float32)
    _oracle().fr_oracle_render_depth_backward(_ptr(depth_grad), _ptr(tri), _ptr(tri_ind), B, nver, T, H, W, _ptr(out))
    return out


# --------------------------------------------------------------------------- compiled reference
_ref_lib = None


def ref_available() -> bool:
    """True when ``oracle/_ref/libref_render_depth.so`` exists or can be built here."""
    if os.path.exists(_REF_SO):
        return True
    try:
        build()
    except Exception:
        return False
    return os.path.exists(_REF_SO)


def _ref():
    global _ref_lib
    if _ref_lib is None:
        if not ref_available():
            raise RuntimeError("oracle/_ref is not built and %s is absent" % REFERENCE_ROOT)
        lib = ctypes.CDLL(_REF_SO)
        lib.ref_render_depth_op.restype = ctypes.c_int
        lib.ref_render_depth_op.argtypes = [_f32p, _i64p, _f32p, _i64p, _f32p, _i64p, _i64p, _f32p, _f32p, _f32p, _f32p,
                                            ctypes.c_char_p, ctypes.c_int]
        lib.ref_render_depth_grad_op.restype = ctypes.c_int
        lib.ref_render_depth_grad_op.argtypes = [_f32p, _i64p, _f32p, _i64p, _f32p, _i64p, _f32p, _f32p, _i64p, _f32p,
                                                 ctypes.c_char_p, ctypes.c_int]
        lib.ref_infer_shapes.restype = ctypes.c_int
        lib.ref_infer_shapes.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int), _i64p,
                                         ctypes.POINTER(ctypes.c_int), _i64p]
        lib.ref_op_signature.restype = ctypes.c_int
        lib.ref_op_signature.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
        _ref_lib = lib
    return _ref_lib


class ReferenceInvalidArgument(ValueError):
    """The reference op's OP_REQUIRES failed (TF would raise InvalidArgumentError)."""


def ref_render_depth(vertex, tri, texture, image_shape):
    """Run the reference's ``RenderDepthOp<CPUDevice>::Compute`` (render_depth_op.cc:378-458).

    ``image_shape`` = (B,H,W,C); image values are never read by the reference.
    Not re-entrant (static scratch, render_depth_op.cc:125-131): call from one thread.
    """
    vertex, tri, texture = _as_f32(vertex), _as_f32(tri), _as_f32(texture)
    B, H, W, _ = [int(s) for s in image_shape]
    ch = texture.shape[1]
    depth = np.empty((B, H, W, 1), np.float32)
    teximg = np.empty((B, H, W, ch), np.float32)
    normal = np.empty((B, H, W, 3), np.float32)
    tri_ind = np.empty((B, H, W, 1), np.float32)
    err = ctypes.create_string_buffer(512)
    rc = _ref().ref_render_depth_op(_ptr(vertex), _dims(vertex.shape), _ptr(tri), _dims(tri.shape), _ptr(texture),
                                    _dims(texture.shape), _dims(image_shape), _ptr(depth), _ptr(teximg), _ptr(normal),
                                    _ptr(tri_ind), err, 512)
    if rc == 1:
        raise ReferenceInvalidArgument(err.value.decode())
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return depth, teximg, normal, tri_ind


def ref_render_depth_grad(depth_grad, vertex, tri, depth, tri_ind, image_shape, sanitize: bool = True):
    """Run the reference's ``RenderDepthOpGrad<CPUDevice>::Compute`` (render_depth_op.cc:470-528).

    With ``sanitize`` (default) the harness removes the reference's two undefined behaviours before
    calling it: ``vertex_grad`` starts at zero, and background pixels
    (``tri_ind < 0``) get ``tri_ind = 0`` and ``depth_grad = 0`` so they contribute exactly +0.0.
    """
    depth_grad, vertex, tri = _as_f32(depth_grad).copy(), _as_f32(vertex), _as_f32(tri)
    depth, tri_ind = _as_f32(depth), _as_f32(tri_ind).copy()
    if sanitize:
        bg = tri_ind < 0
        tri_ind[bg] = 0.0
        depth_grad[bg] = 0.0
    vertex_grad = np.zeros(vertex.shape, np.float32)
    err = ctypes.create_string_buffer(512)
    rc = _ref().ref_render_depth_grad_op(_ptr(depth_grad), _dims(depth_grad.shape), _ptr(vertex), _dims(vertex.shape),
                                         _ptr(tri), _dims(tri.shape), _ptr(depth), _ptr(tri_ind), _dims(image_shape),
                                         _ptr(vertex_grad), err, 512)
    if rc == 1:
        raise ReferenceInvalidArgument(err.value.decode())
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return vertex_grad


def ref_infer_shapes(op: str, in_shapes):
    """Output shapes from the reference's registered shape function (render_depth_op.cc:544-589)."""
    ranks = (ctypes.c_int * len(in_shapes))(*[len(s) for s in in_shapes])
    flat = [int(d) for s in in_shapes for d in s]
    dims = (ctypes.c_longlong * len(flat))(*flat)
    out_ranks = (ctypes.c_int * 4)()
    out_dims = (ctypes.c_longlong * 16)()
    n = _ref().ref_infer_shapes(op.encode(), len(in_shapes), ranks, dims, out_ranks, out_dims)
    if n < 0:
        raise KeyError(op)
    return [tuple(out_dims[4 * i + j] for j in range(out_ranks[i])) for i in range(n)]


def ref_op_signature(op: str) -> str:
    buf = ctypes.create_string_buffer(1024)
    if _ref().ref_op_signature(op.encode(), buf, 1024) < 0:
        raise KeyError(op)
    return buf.value.decode()
