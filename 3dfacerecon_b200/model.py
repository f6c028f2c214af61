"""Device-resident 3DMM: the ``read_3dmm_model`` dict packed once for the kernels.

The reference keeps ``mu``, ``pc_shape`` and ``pc_exp`` as three TF constants and multiplies them separately
(``nets/network.py:41-43,153-159``).  Here they are packed ONCE into a single tiled, K-contiguous matrix
``[pc_shape | pc_exp | mu | pad]`` (DESIGN.md "Packed basis") that every forward/backward launch streams exactly once.
"""
from __future__ import annotations

import hashlib
import os
import warnings

import numpy as np
import torch

from . import _lib
from . import mesh as _mesh
from ._lib import check, lib
from .utils.parser_3dmm import tri_is_one_based

_CONVENTIONS = {
    # name: (pack flags, run flags)            SURVEY.md App. A.2
    "network": (_lib.FR_MEAN_PLANAR | _lib.FR_BASIS_PLANAR, _lib.FR_ROT_XYZ | _lib.FR_YFLIP_S_Y_1),
    "sample_test": (_lib.FR_MEAN_INTERLEAVED | _lib.FR_BASIS_PLANAR, _lib.FR_ROT_ZYX | _lib.FR_YFLIP_S_Y),
    "matlab": (_lib.FR_MEAN_INTERLEAVED | _lib.FR_BASIS_INTERLEAVED, _lib.FR_ROT_XYZ | _lib.FR_YFLIP_NONE),
}


def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


class DeviceModel:
    """Packed basis + triangles + textures of one 3DMM on one GPU."""

    def __init__(self, model: dict, device="cuda:0", convention: str = "network", validate_tri: bool = True,
                 tri_base: int | None = None, cache_dir: str | None = None, cluster_tiles: bool = True):
        if convention not in _CONVENTIONS:
            raise ValueError("convention must be one of %s" % sorted(_CONVENTIONS))
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("DeviceModel needs a CUDA device: there is no CPU path")
        self.convention = convention
        self.pack_flags, self.run_flags = _CONVENTIONS[convention]
        if cluster_tiles:       # FR_CLUSTER_TILES (include/facerecon_b200.h): the fused call rasterizes inside the reconstruction epilogue
                                # (default; costs a second, cluster-ordered copy of the forward operand tiles: +184 MB for BFM)
            self.pack_flags |= _lib.FR_CLUSTER_TILES
            self.run_flags |= _lib.FR_CLUSTER_TILES
        self.cluster_tiles = bool(cluster_tiles)
        mu = np.ascontiguousarray(np.asarray(model["mu"], np.float32).reshape(-1))
        pc_shape = np.ascontiguousarray(model["pc_shape"], np.float32)
        pc_exp = np.ascontiguousarray(model["pc_exp"], np.float32)
        if mu.size % 3 or pc_shape.shape[0] != mu.size or pc_exp.shape[0] != mu.size:
            raise ValueError("mu [3N,1], pc_shape [3N,Ks] and pc_exp [3N,Ke] disagree on 3N")
        self.nver = mu.size // 3
        self.ndim_shape, self.ndim_exp = int(pc_shape.shape[1]), int(pc_exp.shape[1])
        self.ndim_pose = _lib.FR_NDIM_POSE
        self.ndim = self.ndim_pose + self.ndim_shape + self.ndim_exp
        tri = np.ascontiguousarray(model["tri"], np.float32)
        if tri.ndim != 2 or tri.shape[0] != 3:
            raise ValueError("The tri is not 3 x ntri")                           # render_depth_op.cc:417
        # The real Model_Shape.mat stores MATLAB 1-based indices and the reference uses them unshifted (one vertex off,
        # and index nver reads past the row: SURVEY App. B-7).  tri_base=None detects that case and shifts with a
        # warning; tri_base=0 / 1 states the base explicitly (1 -> tri - 1).
        if tri_base is None:
            tri_base = 1 if (tri.size and tri_is_one_based(tri, self.nver)) else 0
            if tri_base == 1:
                warnings.warn("tri looks 1-based (min >= 1, max == nver): using tri - 1 (the reference indexes one vertex off, "
                              "SURVEY App. B-7); pass tri_base=0 to keep the values as they are", stacklevel=2)
        if tri_base not in (0, 1):
            raise ValueError("tri_base must be None, 0 or 1")
        if tri_base == 1:
            tri = tri - 1.0
        if validate_tri and tri.size and (tri.min() < 0 or tri.max() >= self.nver):
            raise ValueError("tri must hold vertex indices in [0, %d) after the base shift (tri_base=%d)" % (self.nver, tri_base))
        self.ntri = int(tri.shape[1])
        self.tri_base = tri_base
        with torch.cuda.device(self.device):
            self.tri = torch.from_numpy(tri).to(self.device)
            self.vertex_code = torch.from_numpy(np.ascontiguousarray(model["vertex"], np.float32)).to(self.device)
            self.mu_tex = torch.from_numpy(np.ascontiguousarray(model["mu_tex"], np.float32)).to(self.device)
            # mesh table (csrc/mesh_table.h): clusters of the triangle list, partitioned along the mean shape's geometry.
            # The partition takes a few seconds of host time for a BFM-sized mesh, so its blob is cached on disk when a
            # cache_dir is given (the packed basis is NOT cached: re-packing 146 MB of source on the GPU is faster than
            # reading the 500 MB packed image back from disk).
            self.mesh = self._mesh_table(tri, mu, cache_dir)
            _mesh.register(self.tri, self.mesh)
            nbytes = lib().fr_packed_basis_bytes(self.nver, self.ndim_shape, self.ndim_exp, self.pack_flags, self.mesh.handle)
            self.packed = torch.empty(nbytes // 4, dtype=torch.float32, device=self.device)
            d_mu = torch.from_numpy(mu).to(self.device)
            d_ps = torch.from_numpy(pc_shape).to(self.device)
            d_pe = torch.from_numpy(pc_exp).to(self.device)
            check(lib().fr_pack_basis(d_mu.data_ptr(), d_ps.data_ptr() if d_ps.numel() else None,
                                      d_pe.data_ptr() if d_pe.numel() else None, self.nver, self.ndim_shape,
                                      self.ndim_exp, self.pack_flags, self.mesh.handle, self.packed.data_ptr(),
                                      _stream_ptr(self.device)))
            # Gram matrix of [pc_shape | pc_exp] for the geometry loss: one plain float64 library GEMM at
            # model load replaces the reference's two 146 MB basis contractions per training step (network.py:348-355)
            d_basis = torch.cat([d_ps, d_pe], dim=1).double()
            self.gram = (d_basis.t() @ d_basis)
            del d_basis
            torch.cuda.current_stream(self.device).synchronize()   # d_mu/d_ps/d_pe die here
        self.basis_bytes = int(nbytes)

    def _mesh_table(self, tri, mu, cache_dir):
        interleaved = bool(self.pack_flags & _lib.FR_MEAN_INTERLEAVED)
        path = None
        if cache_dir is not None:
            digest = hashlib.sha1()
            for part in (tri.tobytes(), mu.tobytes(), b"interleaved" if interleaved else b"planar", b"v%d" % lib().fr_version()):
                digest.update(part)
            path = os.path.join(cache_dir, "mesh_%s.bin" % digest.hexdigest()[:20])
            if os.path.exists(path):
                try:
                    return _mesh.MeshTable(blob=np.fromfile(path, np.uint8), device=self.device.index)
                except (ValueError, RuntimeError):
                    pass                                    # stale / corrupt cache entry: rebuild below
        table = _mesh.MeshTable(tri, self.nver, mu, interleaved=interleaved, device=self.device.index)
        if path is not None:
            os.makedirs(cache_dir, exist_ok=True)
            tmp = path + ".tmp%d" % os.getpid()
            table.blob().tofile(tmp)
            os.replace(tmp, path)
        return table
