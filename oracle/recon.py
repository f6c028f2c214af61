"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's 3DMM reconstruction + projection.

Restates, step by step, what the reference computes with TensorFlow 1.2 / numpy:

* variant A (canonical, what ``trainval.py`` runs): ``FaceRecNet.vertices_transform``
  (``/root/reference/nets/network.py:140-171``) with ``parse_pose_params`` (:253-263) and
  ``rotation_matrix`` (:266-290).
* variant B (config 1): the numpy block of ``rendering_layer/sample_test.py:95-110`` with its own
  ``rotation_matrix`` (:48-72).
* variant C (cross-check only): MATLAB ``prepare_data/Project2D.m:1-12`` + ``RotationMatrix.m:8-12``.

``tf.matmul`` itself (TF 1.2.0, named only in prose at ``README.md:13``) is third-party arithmetic that
cannot run offline; it is restated here as a numpy GEMM.  Pinned by ``tests/golden/recon_*.npz``: outputs of
the reference's OWN ``vertices_transform`` source executed through a numpy stand-in for the dozen TF ops it
calls, and of ``sample_test.py``'s own functions (``tests/golden/make_golden.py``).
"""
from __future__ import annotations

from dataclasses import dataclass
from math import cos, sin

import numpy as np


@dataclass(frozen=True)
class Convention:
    """The reference's three mutually inconsistent conventions."""
    mean_layout: str = "planar"     # planar: mu[c*N+n]; interleaved: mu[3*n+c]
    basis_layout: str = "planar"
    rot_order: str = "xyz"          # xyz: R = Rx.Ry.Rz (network.py:288); zyx: R = Rz.Ry.Rx (sample_test.py:71)
    yflip: str = "S-y-1"            # network.py:168 | "S-y" sample_test.py:105 | "none" Project2D.m


VARIANT_A = Convention()
VARIANT_B = Convention(mean_layout="interleaved", basis_layout="planar", rot_order="zyx", yflip="S-y")
VARIANT_C = Convention(mean_layout="interleaved", basis_layout="interleaved", rot_order="xyz", yflip="none")


def rotation_matrix(angles, rot_order="xyz"):
    """float64 elementary rotations, product cast to float32 (network.py:277-290 / sample_test.py:59-72)."""
    phi, gamma, theta = [float(a) for a in angles]
    r_pitch = np.array([[1, 0, 0], [0, cos(phi), sin(phi)], [0, -sin(phi), cos(phi)]])
    r_yaw = np.array([[cos(gamma), 0, -sin(gamma)], [0, 1, 0], [sin(gamma), 0, cos(gamma)]])
    r_roll = np.array([[cos(theta), sin(theta), 0], [-sin(theta), cos(theta), 0], [0, 0, 1]])
    if rot_order == "xyz":
        r = np.dot(np.dot(r_pitch, r_yaw), r_roll)          # network.py:288
    elif rot_order == "zyx":
        r = r_roll.dot(r_yaw.dot(r_pitch))                  # sample_test.py:71
    else:
        raise ValueError(rot_order)
    return r.astype(np.float32)


def rotation_matrix_batch(angles_batch, rot_order="xyz"):
    """network.py:292-297."""
    return np.stack([rotation_matrix(a, rot_order) for a in np.asarray(angles_batch)]).astype(np.float32)


def _as_rows(vec3n, n, layout, dtype):
    """[3N] or [3N,B] -> [3,N,...] coordinate-major view according to the memory layout."""
    a = np.asarray(vec3n, dtype=dtype)
    tail = a.shape[1:]
    if layout == "planar":
        return a.reshape((3, n) + tail)
    if layout == "interleaved":
        return np.swapaxes(a.reshape((n, 3) + tail), 0, 1)
    raise ValueError(layout)


def reconstruct_vertices(params, model, conv: Convention = VARIANT_A, dtype=np.float64):
    """v[b] = mu + pc_shape.alpha_b + pc_exp.eps_b as [B,3,N] (network.py:153-159 / sample_test.py:97-102)."""
    params = np.atleast_2d(np.asarray(params)).astype(dtype)
    ks, ke = model["pc_shape"].shape[1], model["pc_exp"].shape[1]
    n = model["mu"].shape[0] // 3
    alpha = params[:, 7:7 + ks]
    eps = params[:, 7 + ks:7 + ks + ke]
    shapes = np.asarray(model["pc_shape"], dtype) @ alpha.T       # [3N,B]   network.py:153
    exps = np.asarray(model["pc_exp"], dtype) @ eps.T             # [3N,B]   network.py:155
    geo = _as_rows(shapes, n, conv.basis_layout, dtype) + _as_rows(exps, n, conv.basis_layout, dtype)   # [3,N,B]
    mu = _as_rows(np.asarray(model["mu"]).reshape(-1), n, conv.mean_layout, dtype)                       # [3,N]
    return np.transpose(mu[:, :, None] + geo, (2, 0, 1))          # [B,3,N]  network.py:159


def project(vertex, params, im_size, conv: Convention = VARIANT_A, dtype=np.float64):
    """(f.R).v + t, then the y flip (network.py:163-169 / sample_test.py:104-105)."""
    params = np.atleast_2d(np.asarray(params))
    rot = rotation_matrix_batch(params[:, 0:3], conv.rot_order)                     # float32 values
    f = params[:, 6].astype(np.float32 if dtype == np.float32 else dtype)
    m = (f[:, None, None] * rot.astype(f.dtype)).astype(dtype)                      # network.py:165 f_expand * R
    t = params[:, 3:6].astype(dtype)
    vp = np.einsum("brc,bcn->brn", m, vertex.astype(dtype)) + t[:, :, None]
    if conv.yflip == "S-y-1":
        vp[:, 1, :] = dtype(im_size) - vp[:, 1, :] - dtype(1)                       # network.py:168
    elif conv.yflip == "S-y":
        vp[:, 1, :] = dtype(im_size) - vp[:, 1, :]                                  # sample_test.py:105
    elif conv.yflip != "none":
        raise ValueError(conv.yflip)
    return vp


def vertices_transform(params, model, im_size=200, conv: Convention = VARIANT_A, dtype=np.float64):
    """params [B,d] -> vertex_proj [B,3,N] (network.py:140-171)."""
    return project(reconstruct_vertices(params, model, conv, dtype), params, im_size, conv, dtype)


def vertices_transform_backward(params, model, grad_vertex_proj, conv: Convention = VARIANT_A, dtype=np.float64):
    """Gradient of ``vertices_transform`` w.r.t. params as TF autodiff produces it.

    ``tf.py_func`` (network.py:150) has no gradient, so d/d(phi,gamma,theta) = 0.
    Returns dparams [B,d].
    """
    params = np.atleast_2d(np.asarray(params))
    g = np.array(grad_vertex_proj, dtype=dtype)                                     # [B,3,N]
    if conv.yflip in ("S-y-1", "S-y"):
        g[:, 1, :] = -g[:, 1, :]
    ks, ke = model["pc_shape"].shape[1], model["pc_exp"].shape[1]
    n = model["mu"].shape[0] // 3
    rot = rotation_matrix_batch(params[:, 0:3], conv.rot_order).astype(dtype)
    f = params[:, 6].astype(dtype)
    v = reconstruct_vertices(params, model, conv, dtype)                            # [B,3,N]
    out = np.zeros((params.shape[0], 7 + ks + ke), dtype)
    out[:, 3:6] = g.sum(axis=2)                                                     # d t3d
    rv = np.einsum("brc,bcn->brn", rot, v)
    out[:, 6] = (g * rv).sum(axis=(1, 2))                                           # d f
    dv = f[:, None, None] * np.einsum("brc,brn->bcn", rot, g)                       # [B,3,N] = (f R)^T g
    if conv.basis_layout == "planar":
        dflat = dv.reshape(dv.shape[0], 3 * n)
    else:
        dflat = np.transpose(dv, (0, 2, 1)).reshape(dv.shape[0], 3 * n)
    out[:, 7:7 + ks] = dflat @ np.asarray(model["pc_shape"], dtype)
    out[:, 7 + ks:] = dflat @ np.asarray(model["pc_exp"], dtype)
    return out
