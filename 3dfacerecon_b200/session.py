"""Host-buffer session: numpy / pinned-host arrays in, numpy arrays out, through ``fr_session_*`` of the C ABI.

This is the flavour a caller that is not already on the GPU binds: the library owns the device copy of
the model and the staging buffers, copies ``params`` in, runs recon + projection + render, and copies the depth map out.
``bench.py`` measures its end-to-end number through it.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import check, lib

_CONV = {"network": 0x0, "sample_test": _lib.FR_MEAN_INTERLEAVED | _lib.FR_ROT_ZYX | _lib.FR_YFLIP_S_Y,
         "matlab": _lib.FR_MEAN_INTERLEAVED | _lib.FR_BASIS_INTERLEAVED | _lib.FR_YFLIP_NONE}


def _fptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


class Session:
    def __init__(self, model: dict, height=200, width=200, max_batch=64, device=0, convention="network", cluster_tiles=True):
        mu = np.ascontiguousarray(np.asarray(model["mu"], np.float32).reshape(-1))
        ps = np.ascontiguousarray(model["pc_shape"], np.float32)
        pe = np.ascontiguousarray(model["pc_exp"], np.float32)
        tri = np.ascontiguousarray(model["tri"], np.float32)
        self.nver, self.ntri = mu.size // 3, tri.shape[1]
        self.ks, self.ke = ps.shape[1], pe.shape[1]
        self.d = _lib.FR_NDIM_POSE + self.ks + self.ke
        self.height, self.width, self.max_batch = int(height), int(width), int(max_batch)
        self._h = ctypes.c_void_p()
        self._inflight = {}
        check(lib().fr_session_create(_fptr(mu), _fptr(ps), _fptr(pe), _fptr(tri), self.nver, self.ntri, self.ks, self.ke,
                                      self.height, self.width, self.max_batch,
                                      _CONV[convention] | (_lib.FR_CLUSTER_TILES if cluster_tiles else 0), int(device),
                                      ctypes.byref(self._h)))

    def forward(self, params, im_size=200.0, depth=None, tri_ind=None, vertex_proj=None, want_tri_ind=True):
        """params [B,d] float32 host array -> (depth [B,H,W,1], tri_ind [B,H,W,1] or None).  Output arrays may be
        passed in (e.g. views of pinned torch tensors) to avoid allocations."""
        params = np.ascontiguousarray(params, np.float32)
        B = params.shape[0]
        if depth is None:
            depth = np.empty((B, self.height, self.width, 1), np.float32)
        if tri_ind is None and want_tri_ind:
            tri_ind = np.empty((B, self.height, self.width, 1), np.float32)
        check(lib().fr_session_forward(self._h, _fptr(params), B, float(im_size), _fptr(depth), _fptr(tri_ind), _fptr(vertex_proj)))
        return depth, tri_ind

    def submit(self, slot, params, im_size=200.0, depth=None, tri_ind=None, vertex_proj=None):
        """Enqueue one batch on ``slot`` (0 .. FR_SESSION_SLOTS-1) and return immediately; ``wait(slot)`` blocks until
        ``depth`` (and the optional outputs) are filled.  ``params`` and the output arrays must be C-contiguous float32
        and stay alive and untouched until the wait -- pass pinned memory so that the copies really overlap."""
        for a in (params, depth, tri_ind, vertex_proj):
            if a is not None and not (a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]):
                raise ValueError("submit() needs C-contiguous float32 arrays (no hidden copies on the pipelined path)")
        if depth is None:
            raise ValueError("submit() needs the output array `depth`")
        self._inflight[slot] = (params, depth, tri_ind, vertex_proj)       # keep the buffers alive until wait()
        check(lib().fr_session_submit(self._h, int(slot), _fptr(params), params.shape[0], float(im_size), _fptr(depth),
                                      _fptr(tri_ind), _fptr(vertex_proj)))

    def wait(self, slot):
        check(lib().fr_session_wait(self._h, int(slot)))
        return self._inflight.pop(slot, None)

    def backward(self, depth_grad, params_grad=None):
        depth_grad = np.ascontiguousarray(depth_grad, np.float32)
        B = depth_grad.shape[0]
        if params_grad is None:
            params_grad = np.empty((B, self.d), np.float32)
        check(lib().fr_session_backward(self._h, _fptr(depth_grad), B, _fptr(params_grad)))
        return params_grad

    def close(self):
        if self._h:
            lib().fr_session_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
