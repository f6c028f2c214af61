"""3dfacerecon_b200 -- B200-native 3DMM reconstruction + depth rendering (the 235-d params -> depth-map path of
Cogito2012/3DFaceRecon), PyTorch host code over hand-written sm_100a CUDA behind a C ABI.

The directory name is the reference's; it is not a Python identifier, so import it with::

    import importlib; fr = importlib.import_module("3dfacerecon_b200")

Public surface (mirrors the reference's modules on this path):

* ``fr.render_depth(ver, tri, texture, image)``           <- ``rendering_layer/ops.py:78-81``
* ``fr.FaceRecNet(...).vertices_transform / rendering_layer / depth_rendering_layer / set_constraints``
                                                           <- ``nets/network.py:140-218, 300-308``
* ``fr.read_3dmm_model(path)`` / ``fr.synthetic_3dmm_model()``  <- ``utils/parser_3dmm.py:36-61``
* ``fr.DeviceModel``, ``fr.recon_project``, ``fr.recon_render_depth``, ``fr.Session`` (host-buffer C-ABI session), ``fr.shard_batch``
"""
from . import _lib  # noqa: F401
from .utils.parser_3dmm import read_3dmm_model, synthetic_3dmm_model  # noqa: F401
from .sharding import shard_batch  # noqa: F401

__all__ = ["render_depth", "render_depth_grad", "FaceRecNet", "DeviceModel", "recon_project", "recon_render_depth", "Session", "shard_batch",
           "read_3dmm_model", "synthetic_3dmm_model", "library_path"]


def library_path() -> str:
    return _lib.LIB_PATH


def __getattr__(name):
    # torch-dependent pieces are imported lazily so that the CPU-only bits (synth, parser, sharding, ABI checks)
    # import fast; nothing here falls back to a CPU implementation.
    if name in ("render_depth", "render_depth_grad"):
        from .rendering_layer import ops
        return getattr(ops, name)
    if name in ("FaceRecNet", "recon_project", "recon_render_depth"):
        from .nets import network
        return getattr(network, name)
    if name == "DeviceModel":
        from .model import DeviceModel
        return DeviceModel
    if name == "Session":
        from .session import Session
        return Session
    raise AttributeError(name)
