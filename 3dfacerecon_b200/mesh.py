"""Mesh tables: the per-model side table of the rasterizer (``fr_mesh_table_*`` of the C ABI, ``csrc/mesh_table.h``).

The reference passes ``tri`` [3,T] to every ``render_depth`` call and its op re-reads and re-converts the float indices
per triangle and face (``render_depth_op.cc:204-213``).  A mesh table partitions the triangle list ONCE into clusters of
at most 128 vertices that the CUDA rasterizer stages in shared memory; results are bit-identical with and without it.

``table_for(tri)`` is what ``render_depth`` uses: a small registry keyed by the identity of the ``tri`` tensor (storage
address + version counter, with the tensor kept alive so the address cannot be recycled).  ``DeviceModel`` registers the
table it builds at model load; for any other ``tri`` tensor a table is built the second time the same tensor is seen
(one-shot triangle lists go through the generic kernels and never pay for a build).
"""
from __future__ import annotations

import ctypes
from collections import OrderedDict

import numpy as np

from ._lib import check, lib

_MAX_TABLES = 8
_registry = OrderedDict()      # key -> (tri tensor kept alive, MeshTable or None when seen only once)


class MeshTable:
    """Handle of one ``fr_mesh_table`` (host blob + its copy on ``device``; ``device=None`` keeps it host-only)."""

    def __init__(self, tri=None, nver=None, positions=None, interleaved=False, device=None, blob=None):
        self._h = ctypes.c_void_p()
        dev = -1 if device is None else int(device)
        if blob is not None:
            blob = np.ascontiguousarray(blob, np.uint8)
            check(lib().fr_mesh_table_from_blob(ctypes.c_void_p(blob.ctypes.data), blob.size, dev, ctypes.byref(self._h)))
        else:
            tri = np.ascontiguousarray(tri, np.float32)
            if tri.ndim != 2 or tri.shape[0] != 3:
                raise ValueError("The tri is not 3 x ntri")                       # render_depth_op.cc:417
            pos = None if positions is None else np.ascontiguousarray(positions, np.float32).reshape(-1)
            if pos is not None and pos.size != 3 * int(nver):
                raise ValueError("positions must hold 3 * nver floats")
            check(lib().fr_mesh_table_create(ctypes.c_void_p(tri.ctypes.data), int(tri.shape[1]), int(nver),
                                             None if pos is None else ctypes.c_void_p(pos.ctypes.data), int(bool(interleaved)),
                                             dev, ctypes.byref(self._h)))
        self.device = device
        self.nclusters = int(lib().fr_mesh_table_clusters(self._h))
        self.vertex_slots = int(lib().fr_mesh_table_vertex_slots(self._h))

    @property
    def handle(self):
        return self._h

    def blob(self) -> np.ndarray:
        """The serialised table (uint8 copy): header, cluster vertex lists, triangle entries -- ``csrc/mesh_table.h``."""
        n = ctypes.c_size_t()
        p = lib().fr_mesh_table_blob(self._h, ctypes.byref(n))
        return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_ubyte)), shape=(n.value,)).copy()

    def parsed(self) -> dict:
        """Header fields and the three arrays of the table as numpy views of a copy (tests, inspection)."""
        b = self.blob()
        h = b[:64].view(np.int32)
        ncl, nslots = int(h[4]), int(h[5])
        out = {"nver": int(h[2]), "ntri": int(h[3]), "nclusters": ncl, "ntri_slots": nslots, "max_cluster_tris": int(h[6]),
               "nvert_slots": int(h[7])}
        off_v, off_b, off_t = int(h[8]), int(h[9]), int(h[10])
        out["cluster_vert"] = b[off_v:off_v + ncl * 128 * 4].view(np.int32).reshape(ncl, 128)
        out["tri_begin"] = b[off_b:off_b + (ncl + 1) * 4].view(np.int32)
        out["tri_entry"] = b[off_t:off_t + nslots * 8].view(np.uint32).reshape(nslots, 2)
        off_q, off_rv, off_vr = int(h[13]), int(h[14]), int(h[15])
        out["tri_vid"] = b[off_q:off_q + nslots * 16].view(np.uint32).reshape(nslots, 4)
        nrank = (out["nver"] + 127) // 128 * 128
        out["rank_vert"] = b[off_rv:off_rv + nrank * 4].view(np.int32)
        out["vert_rank"] = b[off_vr:off_vr + out["nver"] * 4].view(np.int32)
        off_cr = (off_vr + out["nver"] * 4 + 15) // 16 * 16
        out["cluster_rank"] = b[off_cr:off_cr + ncl * 128 * 4].view(np.int32).reshape(ncl, 128)
        off_t4 = off_cr + ncl * 128 * 4
        out["tri_rank4"] = b[off_t4:off_t4 + out["ntri"] * 16].view(np.uint32).reshape(out["ntri"], 4)
        return out

    def close(self):
        if self._h:
            lib().fr_mesh_table_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _key(tri):
    return (tri.device.index, tri.data_ptr(), tri._version, tuple(tri.shape))


def register(tri, table: MeshTable) -> None:
    """Associate ``table`` with the tensor ``tri`` (kept alive by the registry)."""
    _registry[_key(tri)] = (tri, table)
    _registry.move_to_end(_key(tri))
    while len(_registry) > _MAX_TABLES:
        _registry.popitem(last=False)


def table_for(tri, nver: int, ver=None, build: str = "auto"):
    """The mesh table registered for the float32 CUDA tensor ``tri`` [3,T], or None (-> generic kernels).

    ``build``: "auto" builds a table the second time the same tensor is seen, "now" builds it immediately, "never"
    only looks it up.  ``ver`` [B,3,N] (optional) donates the positions of its first face to the partitioner."""
    key = _key(tri)
    hit = _registry.get(key)
    if hit is not None and hit[1] is not None:
        _registry.move_to_end(key)
        return hit[1]
    if build == "never" or tri.shape[1] == 0:
        return None
    if build == "auto" and hit is None:
        _registry[key] = (tri, None)             # first sighting: remember it, do not pay for a build yet
        while len(_registry) > _MAX_TABLES:
            _registry.popitem(last=False)
        return None
    pos = None if ver is None else ver[0].detach().float().cpu().numpy()
    table = MeshTable(tri.detach().cpu().numpy(), nver, pos, device=tri.device.index)
    register(tri, table)
    return table
