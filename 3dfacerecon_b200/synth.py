"""Synthetic BFM-shaped 3DMM ("SYN-BFM") and parameter samplers.

The Basel Face Model ``.mat`` files the reference loads (``utils/parser_3dmm.py:6-33``) are not
redistributable and not available offline, so benchmarks and tests run on a synthetic model with
the TRUE dimensions (53 215 vertices, 105 840 triangles, 199 shape + 29 expression components)
and the same dict layout ``read_3dmm_model`` returns (``utils/parser_3dmm.py:50-60``), so byte
counts and access patterns equal the real model's.

Geometry: a 145 x 367 vertex grid on an ellipsoidal cap in BFM-like micrometre units
(x in [-8e4, 8e4], y in [-1e5, 1e5]); a scale f ~ 1e-3 maps it to ~160 x 200 px, i.e. roughly
1.1 px / 0.55 px vertex spacing, like the real mesh rendered at 200 x 200.  Bases are spatially
smooth cosine modes (like PCA modes) plus a little white noise so that no two entries repeat.
"""
from __future__ import annotations

import numpy as np

BFM_GRID = (145, 367)      # 145 * 367 = 53 215 vertices
BFM_NTRI = 105840          # 2 * 144 * 366 = 105 408 grid triangles + 432 duplicates
NDIM_POSE = 7              # utils/parser_3dmm.py:49
NDIM_SHAPE = 199
NDIM_EXP = 29


def _smooth_basis(rng, u, w, ncomp, rms, noise):
    """[3N, ncomp] float32: column k = low-order 2-D cosine mode per coordinate, RMS ``rms``, + N(0, noise)."""
    n = u.size
    out = np.empty((3 * n, ncomp), np.float32)
    for k in range(ncomp):
        for c in range(3):
            fu, fw = rng.uniform(-4.0, 4.0, 2)
            phase = rng.uniform(0.0, 2.0 * np.pi)
            amp = rng.uniform(0.5, 1.5)
            col = amp * np.cos(np.pi * (fu * u + fw * w) + phase)
            col *= rms / max(float(np.sqrt(np.mean(col * col))), 1e-12)
            out[c * n:(c + 1) * n, k] = col
    out += rng.normal(0.0, noise, out.shape).astype(np.float32)
    return out


def make_synthetic_model(grid=BFM_GRID, ndim_shape=NDIM_SHAPE, ndim_exp=NDIM_EXP, ntri=None, seed=0, jitter=0.2, permute=False):
    """Return a model dict with the keys of ``read_3dmm_model`` (``utils/parser_3dmm.py:50-60``).

    ``mu`` [3N,1] is PLANAR (x block, y block, z block) as ``nets/network.py:157`` reads it.
    ``tri`` [3,T] float32 holds 0-based vertex indices.
    ``jitter`` displaces each mean vertex by U(-jitter, jitter) grid cells (0 = exact grid, the
    stress case where many pixel centres fall exactly on shared edges).
    ``permute`` renumbers the vertices with a random permutation (applied consistently to ``mu``, the bases, the
    per-vertex textures and ``tri``) and shuffles the triangle order: the same surface, but without the generator's
    grid order, in which triangle i touches vertices ~ i/2 (best-case gather locality that a real BFM file does not have).
    """
    gx, gy = int(grid[0]), int(grid[1])
    n = gx * gy
    rng = np.random.default_rng(seed)
    col, row = np.meshgrid(np.arange(gx, dtype=np.float64), np.arange(gy, dtype=np.float64))  # [gy, gx]
    if jitter > 0.0:
        col = col + rng.uniform(-jitter, jitter, col.shape)
        row = row + rng.uniform(-jitter, jitter, row.shape)
    u = (2.0 * col / max(gx - 1, 1) - 1.0).ravel()
    w = (2.0 * row / max(gy - 1, 1) - 1.0).ravel()
    x = 8.0e4 * u
    y = 1.0e5 * w
    z = 6.0e4 * np.sqrt(np.maximum(0.0, 1.2 - u * u - w * w))
    z += 4.0e3 * np.cos(3.0 * np.pi * u) * np.cos(2.0 * np.pi * w) + 2.5e4 * np.exp(-((u * 4.0) ** 2 + (w * 3.0 + 0.3) ** 2))
    mu = np.concatenate([x, y, z]).astype(np.float32).reshape(3 * n, 1)

    # two triangles per grid quad, consistent winding; vertex index = row * gx + col
    r, c = np.meshgrid(np.arange(gy - 1), np.arange(gx - 1), indexing="ij")
    v00 = (r * gx + c).ravel()
    v01 = v00 + 1
    v10 = v00 + gx
    v11 = v10 + 1
    tri = np.concatenate([np.stack([v00, v10, v01]), np.stack([v01, v10, v11])], axis=1)  # [3, 2*quads]
    order = np.argsort(np.concatenate([2 * np.arange(v00.size), 2 * np.arange(v00.size) + 1]), kind="stable")
    tri = tri[:, order]
    if ntri is None:
        ntri = BFM_NTRI if (gx, gy) == BFM_GRID else tri.shape[1]
    if ntri > tri.shape[1]:
        # pad with duplicates of existing triangles: a duplicate has a higher index and equal depth,
        # so it never wins a tie (render_depth_op.cc:295) and the output is unchanged
        extra = rng.integers(0, tri.shape[1], ntri - tri.shape[1])
        tri = np.concatenate([tri, tri[:, extra]], axis=1)
    tri = tri[:, :ntri].astype(np.float32)

    pc_shape = _smooth_basis(rng, u, w, ndim_shape, 2.5e-2, 2.5e-4)
    pc_exp = _smooth_basis(rng, u, w, ndim_exp, 1.0e3, 10.0)
    vertex_code = rng.uniform(0.0, 1.0, (3, n)).astype(np.float32)
    mu_tex = rng.uniform(0.0, 255.0, (3, n)).astype(np.float32)
    if permute:
        prng = np.random.default_rng(seed + 7919)
        new_id = prng.permutation(n)                    # new_id[old vertex] = its new number
        old_of = np.argsort(new_id)                     # old_of[new vertex] = old number
        rows = (np.arange(3)[:, None] * n + old_of[None, :]).ravel()      # planar 3N rows in the new numbering
        mu, pc_shape, pc_exp = mu[rows], pc_shape[rows], pc_exp[rows]
        vertex_code, mu_tex = vertex_code[:, old_of], mu_tex[:, old_of]
        tri = new_id[tri.astype(np.int64)][:, prng.permutation(tri.shape[1])].astype(np.float32)
    return {
        "vertex": vertex_code,                                           # PNCC code
        "tri": tri,
        "mu": mu,
        "mu_tex": mu_tex,
        "pc_tex": np.zeros((3 * n, 1), np.float32),                      # not on the hot path
        "param_tex": np.zeros((1, 1), np.float32),
        "pc_shape": pc_shape,
        "pc_exp": pc_exp,
        "ndim_shape": int(ndim_shape),
        "ndim_exp": int(ndim_exp),
        "ndim_pose": NDIM_POSE,
    }


def sample_params_constrained(batch, ndim_shape=NDIM_SHAPE, ndim_exp=NDIM_EXP, im_size=200, seed=2, angle_scale=0.3,
                              full_range=False):
    """[B, 7+ndim_shape+ndim_exp] float32 in the ranges ``set_constraints`` produces (``nets/network.py:210-217``).

    Layout (``nets/network.py:143-145, 258-262``): phi, gamma, theta | t3d x,y,z | f | shape | expression.
    By default angles are scaled by ``angle_scale`` and t_x,t_y in [0.3,0.7]*im_size, f in [6e-4,1e-3] so most
    of the face stays in frame; ``full_range`` draws from the whole constraint box.
    """
    rng = np.random.default_rng(seed)
    p = np.zeros((batch, NDIM_POSE + ndim_shape + ndim_exp), np.float32)
    if full_range:
        p[:, 0:3] = rng.uniform(-1.5, 1.5, (batch, 3))
        p[:, 3:5] = rng.uniform(0.0, im_size, (batch, 2))
        p[:, 6] = rng.uniform(0.0, 1e-3, batch)
    else:
        p[:, 0:3] = rng.uniform(-1.5, 1.5, (batch, 3)) * angle_scale
        p[:, 3:5] = rng.uniform(0.3 * im_size, 0.7 * im_size, (batch, 2))
        p[:, 6] = rng.uniform(6e-4, 1e-3, batch)
    p[:, 5] = 0.0                                                       # nets/network.py:213
    p[:, 7:7 + ndim_shape] = rng.uniform(0.0, 1e4, (batch, ndim_shape))
    p[:, 7 + ndim_shape:] = rng.uniform(-1.5, 1.5, (batch, ndim_exp))
    return p


def sample_params_sample_test(ndim_shape=NDIM_SHAPE, ndim_exp=NDIM_EXP, im_size=200, seed=1):
    """One parameter vector with ``sample_test.get_random_params(im_size, ., ., beta=1.0)`` semantics
    (``rendering_layer/sample_test.py:23-38``): pose is exactly [0,0,0,S/2,S/2,0,1e-3]; shape ~ U[0,1e4);
    expression ~ U[-1.5,1.5).  Returned as float64 [1, d] like the reference's float64 numpy values."""
    rng = np.random.default_rng(seed)
    p = np.zeros((1, NDIM_POSE + ndim_shape + ndim_exp), np.float64)
    p[0, 0:7] = np.array([0, 0, 0, im_size / 2, im_size / 2, 0, 0.001], np.float32)
    p[0, 7:7 + ndim_shape] = rng.random(ndim_shape) * 1e4
    p[0, 7 + ndim_shape:] = -1.5 + 3 * rng.random(ndim_exp)
    return p
