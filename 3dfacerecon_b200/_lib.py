"""ctypes binding of the C-ABI library ``lib3dfacerecon_b200.so`` (``include/facerecon_b200.h``).

This is the only bridge between the Python host code and the CUDA kernels -- the role
``tf.load_op_library`` plays in the reference (``rendering_layer/ops.py:63-72``).  Unlike the reference
it never compiles on import and never falls back: if the shared library is missing, loading fails loudly
and tells the caller to run ``python __graft_entry__.py build`` (or ``make -C 3dfacerecon_b200/csrc``).
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FR_LIB_PATH: developer switch for A/B runs of differently built libraries (tools/); the product always loads the in-tree one
LIB_PATH = os.environ.get("FR_LIB_PATH") or os.path.join(_HERE, "lib3dfacerecon_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "facerecon_b200.h")

FR_OK, FR_ERR_INVALID_ARGUMENT, FR_ERR_CUDA, FR_ERR_WORKSPACE, FR_ERR_UNSUPPORTED = 0, 1, 2, 3, 4
FR_ROT_XYZ, FR_ROT_ZYX = 0x0, 0x1
FR_YFLIP_S_Y_1, FR_YFLIP_S_Y, FR_YFLIP_NONE = 0x0, 0x2, 0x4
FR_MEAN_PLANAR, FR_MEAN_INTERLEAVED = 0x0, 0x10
FR_BASIS_PLANAR, FR_BASIS_INTERLEAVED = 0x0, 0x20
FR_CLUSTER_TILES = 0x40
FR_PARAMS_RAW = 0x100
FR_NDIM_POSE = 7
FR_SESSION_SLOTS = 3

_vp, _sz, _i, _f, _u, _ll = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_float, ctypes.c_uint, ctypes.c_longlong

#: name -> (restype, argtypes); one entry per symbol declared in include/facerecon_b200.h
SIGNATURES = {
    "fr_last_error": (ctypes.c_char_p, []),
    "fr_version": (_i, []),
    "fr_launch_count": (ctypes.c_ulonglong, []),
    "fr_mesh_table_create": (_i, [_vp, _i, _i, _vp, _i, _i, ctypes.POINTER(_vp)]),
    "fr_mesh_table_from_blob": (_i, [_vp, _sz, _i, ctypes.POINTER(_vp)]),
    "fr_mesh_table_destroy": (None, [_vp]),
    "fr_mesh_table_blob": (_vp, [_vp, ctypes.POINTER(_sz)]),
    "fr_mesh_table_clusters": (_i, [_vp]),
    "fr_mesh_table_vertex_slots": (_i, [_vp]),
    "fr_packed_basis_bytes": (_sz, [_i, _i, _i, _u, _vp]),
    "fr_pack_basis": (_i, [_vp, _vp, _vp, _i, _i, _i, _u, _vp, _vp, _vp]),
    "fr_recon_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "fr_recon_project_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _u, _vp, _sz, _vp]),
    "fr_recon_project_backward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _u, _vp, _sz, _vp]),
    "fr_render_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "fr_render_depth_forward": (_i, [_vp, _vp, _vp, _ll, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "fr_render_depth_backward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "fr_rendering_layer_forward": (_i, [_vp, _vp, _vp, _ll, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "fr_rendering_layer_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "fr_pipeline_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "fr_recon_render_forward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _u, _vp, _sz, _vp, _vp]),
    "fr_recon_render_forward_all": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _u, _vp, _sz,
                                         _vp]),
    "fr_session_create": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _u, _i, ctypes.POINTER(_vp)]),
    "fr_session_destroy": (None, [_vp]),
    "fr_session_forward": (_i, [_vp, _vp, _i, _f, _vp, _vp, _vp]),
    "fr_session_submit": (_i, [_vp, _i, _vp, _i, _f, _vp, _vp, _vp]),
    "fr_session_wait": (_i, [_vp, _i]),
    "fr_session_backward": (_i, [_vp, _vp, _i, _vp]),
}

_lib = None


class FaceReconError(RuntimeError):
    """A CUDA / workspace failure reported by the library (FR_ERR_CUDA, FR_ERR_WORKSPACE, FR_ERR_UNSUPPORTED)."""


def lib():
    """The loaded library.  Raises ``ImportError`` (never falls back to a CPU path) when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: the CUDA extension is not built and there is no CPU fallback. "
                "Run `python __graft_entry__.py build` or `make -C 3dfacerecon_b200/csrc`." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error() -> str:
    msg = lib().fr_last_error()
    return msg.decode() if msg else ""


def check(rc: int) -> None:
    """Map a status code to the exception the reference's caller would see: shape-rule violations are TF
    ``InvalidArgument`` there (``render_depth_op.cc:408-418``) and ``ValueError`` here."""
    if rc == FR_OK:
        return
    msg = last_error()
    if rc == FR_ERR_INVALID_ARGUMENT:
        raise ValueError(msg)
    raise FaceReconError("facerecon_b200 error %d: %s" % (rc, msg))
