#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE (run in the authoring container only).

    python tests/golden/make_golden.py          # needs /root/reference; rewrites the .npz files

Nothing under /root/reference is copied: its code is executed where it lies.

* ``render_*.npz``     -- outputs of the UNMODIFIED reference CPU op (render_depth_op.cc compiled in place as
                          oracle/_ref/libref_render_depth.so and driven through its OpKernel::Compute).
* ``recon_A_*.npz``    -- outputs of the reference's own ``FaceRecNet.vertices_transform`` /
                          ``rotation_matrix`` source (nets/network.py:140-171, 253-297), extracted with ``ast``
                          and executed against ``FakeTF`` below -- a float32 numpy stand-in for the dozen
                          TensorFlow-1.2 ops that method calls (TensorFlow itself cannot be installed offline).
* ``recon_B_*.npz``    -- outputs of ``rendering_layer/sample_test.py``'s own ``get_random_params`` /
                          ``rotation_matrix`` functions and of the numpy statements of its ``main()``
                          (:93-113), executed verbatim from the reference file.
* ``layer_cases.npz``  -- outputs of the reference's own ``FaceRecNet.rendering_layer`` source (nets/network.py:172-201),
                          executed against ``FakeTF`` with ``render_depth`` bound to the compiled reference op.
* ``constraints.npz``  -- outputs of the reference's own ``FaceRecNet.set_constraints`` source (:204-218).
* ``geometry_loss.npz``-- value of the geometry-loss statements of ``FaceRecNet.get_loss`` (:346-355), executed verbatim.
"""
from __future__ import annotations

import ast
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("FR_REFERENCE_ROOT", "/root/reference")

import oracle  # noqa: E402

synth = importlib.import_module("3dfacerecon_b200.synth")


# ----------------------------------------------------------------------------- numpy stand-in for TF 1.2
class _T(np.ndarray):
    def set_shape(self, shape):  # tf.Tensor.set_shape: static-shape hint only
        assert tuple(self.shape) == tuple(int(s) for s in shape)


def _t(a):
    return np.asarray(a).view(_T)


class FakeTF:
    """float32 numpy semantics for exactly the ops ``vertices_transform`` uses."""
    float32 = np.float32

    @staticmethod
    def constant(v, dtype=None, name=None):
        return _t(np.asarray(v, dtype=dtype))

    @staticmethod
    def squeeze(x, axis=None, name=None):
        return _t(np.squeeze(x, axis=tuple(axis)))

    @staticmethod
    def slice(x, begin, size, name=None):
        idx = tuple(slice(b, None if s == -1 else b + s) for b, s in zip(begin, size))
        return _t(np.asarray(x)[idx])

    @staticmethod
    def concat(values, axis, name=None):
        return _t(np.concatenate([np.asarray(v) for v in values], axis=axis))

    @staticmethod
    def py_func(func, inp, Tout, name=None):
        return _t(np.asarray(func(*[np.asarray(i) for i in inp]), dtype=Tout))

    @staticmethod
    def transpose(x, perm=None, name=None):
        return _t(np.transpose(np.asarray(x), perm))

    @staticmethod
    def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
        a, b = np.asarray(a), np.asarray(b)
        if transpose_a:
            a = np.swapaxes(a, -1, -2)
        if transpose_b:
            b = np.swapaxes(b, -1, -2)
        assert a.dtype == np.float32 and b.dtype == np.float32
        return _t(np.matmul(a, b))

    @staticmethod
    def reshape(x, shape, name=None):
        return _t(np.reshape(np.asarray(x), [int(s) for s in shape]))

    @staticmethod
    def tile(x, multiples, name=None):
        return _t(np.tile(np.asarray(x), [int(m) for m in multiples]))

    @staticmethod
    def expand_dims(x, axis, name=None):
        return _t(np.expand_dims(np.asarray(x), axis))

    @staticmethod
    def shape(x, name=None):
        return np.asarray(np.shape(x))

    # ---- the further ops rendering_layer / set_constraints / the geometry loss use (float32 elementwise numpy)
    @staticmethod
    def to_float(x, name=None):
        return _t(np.asarray(x, np.float32))

    @staticmethod
    def clip_by_value(x, lo, hi, name=None):
        return _t(np.minimum(np.maximum(np.asarray(x), np.float32(lo)), np.float32(hi)))

    @staticmethod
    def where(cond, a, b, name=None):
        return _t(np.where(np.asarray(cond), np.asarray(a), np.asarray(b)))

    @staticmethod
    def reduce_sum(x, axis=None, name=None):
        x = np.asarray(x)
        assert x.dtype == np.float32 and axis == -1 and x.shape[-1] == 3
        return _t((x[..., 0] + x[..., 1]) + x[..., 2])        # Eigen reduces a 3-long inner axis in order

    @staticmethod
    def square(x, name=None):
        x = np.asarray(x)
        return _t(x * x)

    @staticmethod
    def zeros_like(x, name=None):
        return _t(np.zeros_like(np.asarray(x)))

    @staticmethod
    def sqrt(x, name=None):
        return _t(np.sqrt(np.asarray(x)))

    @staticmethod
    def maximum(x, y, name=None):
        return _t(np.maximum(np.asarray(x), np.float32(y)))

    class nn:
        @staticmethod
        def sigmoid(x, name=None):
            x = np.asarray(x, np.float32)
            return _t(np.float32(1) / (np.float32(1) + np.exp(-x)))

    class losses:
        @staticmethod
        def mean_squared_error(labels, predictions, scope=None):
            d = np.asarray(predictions, np.float32) - np.asarray(labels, np.float32)
            return np.float32(np.sum((d * d).astype(np.float64)) / d.size)     # TF: float32 sum of squares / count; summed in
                                                                               # float64 here so the golden is order-free


def _extract_methods(path, class_name, names):
    tree = ast.parse(open(path).read(), path)
    out = []
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == class_name:
            out = [n for n in node.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(out) == len(names), [n.name for n in out]
    return out


def reference_vertices_transform(model, params, im_size):
    """Run the reference's FaceRecNet.vertices_transform on (model, params[B,d] float32)."""
    from math import cos, sin
    path = os.path.join(REF, "nets", "network.py")
    methods = _extract_methods(path, "FaceRecNet",
                               ["vertices_transform", "parse_pose_params", "rotation_matrix", "rotation_matrix_batch"])
    cls = ast.ClassDef(name="RefGeometry", bases=[], keywords=[], body=methods, decorator_list=[])
    if hasattr(ast, "TypeVar"):
        cls.type_params = []
    mod = ast.fix_missing_locations(ast.Module(body=[cls], type_ignores=[]))
    ns = {"np": np, "tf": FakeTF, "cos": cos, "sin": sin}
    exec(compile(mod, path, "exec"), ns)
    g = ns["RefGeometry"]()
    # attributes FaceRecNet.__init__ sets (network.py:19-55)
    g.batch_size = params.shape[0]
    g.im_size = im_size
    g.mu = FakeTF.constant(model["mu"], np.float32)
    g.pc_shape = FakeTF.constant(model["pc_shape"], np.float32)
    g.pc_exp = FakeTF.constant(model["pc_exp"], np.float32)
    g.ndim_shape, g.ndim_exp, g.ndim_pose = model["ndim_shape"], model["ndim_exp"], model["ndim_pose"]
    pred = FakeTF.constant(params.astype(np.float32)[:, None, None, :])            # (B,1,1,d) network.py:61
    vp = np.asarray(g.vertices_transform(pred), np.float32)
    rot = np.asarray(g.rotation_matrix_batch(params[:, 0:3].astype(np.float32)), np.float32)
    return vp, rot


def reference_sample_test(model, im_size, seed):
    """Run sample_test.py's own functions and the numpy statements of its main() (:93-113)."""
    from math import cos, sin
    path = os.path.join(REF, "rendering_layer", "sample_test.py")
    tree = ast.parse(open(path).read(), path)
    funcs = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("get_random_params", "parse_pose_params", "rotation_matrix")]
    main = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "main"][0]

    def targets(stmt):
        return [t.id for t in getattr(stmt, "targets", []) if isinstance(t, ast.Name)] + \
               [e.id for t in getattr(stmt, "targets", []) if isinstance(t, ast.Tuple) for e in t.elts if isinstance(e, ast.Name)]

    start = next(i for i, s in enumerate(main.body) if "pose_param" in targets(s))
    stop = next(i for i, s in enumerate(main.body) if "abedo_code" in targets(s))
    block = main.body[start:stop + 1]
    mod = ast.fix_missing_locations(ast.Module(body=funcs + block, type_ignores=[]))
    nvert = model["mu"].shape[0] // 3
    ns = {"np": np, "rand": np.random.rand, "cos": cos, "sin": sin,
          "im_size": im_size, "ndim_shape": model["ndim_shape"], "ndim_exp": model["ndim_exp"],
          "pc_shape": model["pc_shape"], "pc_exp": model["pc_exp"], "mu": model["mu"], "nvert": nvert,
          "tri": model["tri"], "vertex_code": model["vertex"], "mu_tex": model["mu_tex"]}
    np.random.seed(seed)
    exec(compile(mod, path, "exec"), ns)
    params = np.concatenate([ns["pose_param"], ns["shape_param"], ns["exp_param"]]).reshape(1, -1)
    angles = np.random.uniform(-1.5, 1.5, (5, 3))
    rots = np.stack([ns["rotation_matrix"](list(a)) for a in angles])
    return params, ns["vertex_proj"], angles, rots


def _ref_class(names, extra_ns=None):
    """FaceRecNet methods `names`, extracted from the reference file and compiled against FakeTF."""
    from math import cos, sin
    path = os.path.join(REF, "nets", "network.py")
    methods = _extract_methods(path, "FaceRecNet", names)
    cls = ast.ClassDef(name="RefNet", bases=[], keywords=[], body=methods, decorator_list=[])
    if hasattr(ast, "TypeVar"):
        cls.type_params = []
    mod = ast.fix_missing_locations(ast.Module(body=[cls], type_ignores=[]))
    ns = {"np": np, "tf": FakeTF, "cos": cos, "sin": sin}
    ns.update(extra_ns or {})
    exec(compile(mod, path, "exec"), ns)
    return ns["RefNet"]()


def reference_rendering_layer(vertex_proj, tri, colors, im_gray):
    """Run the reference's FaceRecNet.rendering_layer (network.py:172-201); its render_depth is the compiled reference op."""
    def render_depth(ver, tri, texture, image):
        return tuple(_t(a) for a in oracle.ref_render_depth(np.asarray(ver), np.asarray(tri), np.asarray(texture), np.shape(image)))
    g = _ref_class(["rendering_layer"], {"render_depth": render_depth})
    g.batch_size = vertex_proj.shape[0]
    g.im_gray = _t(im_gray.astype(np.float32))
    return [np.asarray(a, np.float32) for a in g.rendering_layer(vertex_proj.astype(np.float32), tri, colors)]


def reference_set_constraints(raw, im_size, ks, ke):
    """Run the reference's FaceRecNet.set_constraints (network.py:204-218) on raw [B,1,1,d]."""
    g = _ref_class(["set_constraints"])
    g.im_size, g.ndim_pose, g.ndim_shape, g.ndim_exp, g.ndim = im_size, 7, ks, ke, 7 + ks + ke
    return np.asarray(g.set_constraints(_t(raw.astype(np.float32))), np.float32)


def reference_geometry_loss(model, pred, label):
    """Execute the geometry-loss statements of FaceRecNet.get_loss (network.py:346-355) verbatim."""
    path = os.path.join(REF, "nets", "network.py")
    get_loss = _extract_methods(path, "FaceRecNet", ["get_loss"])[0]
    wanted = ("geometry_pred", "geometry_label", "geometry_basis", "loss_geometry")
    stmts = []
    for node in ast.walk(get_loss):
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name) and node.targets[0].id in wanted:
            stmts.append(node)
    stmts.sort(key=lambda n: n.lineno)
    assert [n.targets[0].id for n in stmts] == list(wanted), [n.targets[0].id for n in stmts]

    class Self:
        pass
    me = Self()
    me.pred_params, me.params_label = _t(pred.astype(np.float32)[:, None, None, :]), _t(label.astype(np.float32)[:, None, None, :])
    me.pc_shape, me.pc_exp = _t(model["pc_shape"].astype(np.float32)), _t(model["pc_exp"].astype(np.float32))
    me.ndim_pose, me.ndim_shape, me.ndim_exp = 7, model["ndim_shape"], model["ndim_exp"]
    ns = {"np": np, "tf": FakeTF, "self": me}
    exec(compile(ast.fix_missing_locations(ast.Module(body=stmts, type_ignores=[])), path, "exec"), ns)
    return np.float32(ns["loss_geometry"])


def small_model(grid, ks, ke, seed, jitter):
    return synth.make_synthetic_model(grid=grid, ndim_shape=ks, ndim_exp=ke, seed=seed, jitter=jitter)


def model_arrays(m):
    return {k: m[k] for k in ("mu", "pc_shape", "pc_exp", "tri", "vertex", "mu_tex")}


def main():
    assert os.path.isdir(REF), "the reference tree is required to regenerate goldens"
    oracle.build()
    assert oracle.ref_available()
    out = {}

    # ---------------- recon, variant A (network.py) : tiny dims and true K dims
    for tag, grid, ks, ke, B, seed in (("tiny", (7, 9), 5, 3, 4, 11), ("truek", (6, 8), 199, 29, 3, 12)):
        m = small_model(grid, ks, ke, seed, 0.2)
        p = synth.sample_params_constrained(B, ks, ke, 200, seed=seed, full_range=(tag == "tiny"))
        vp, rot = reference_vertices_transform(m, p, 200)
        np.savez_compressed(os.path.join(HERE, "recon_A_%s.npz" % tag), params=p, vertex_proj=vp, rot=rot,
                            im_size=200, grid=np.array(grid), **model_arrays(m))
        out["recon_A_" + tag] = vp.shape

    # ---------------- recon, variant B (sample_test.py)
    m = small_model((6, 8), 199, 29, 13, 0.2)
    params, vp, angles, rots = reference_sample_test(m, 200, seed=1)
    np.savez_compressed(os.path.join(HERE, "recon_B_sample_test.npz"), params=params, vertex_proj=vp, angles=angles,
                        rots=rots, im_size=200, **model_arrays(m))
    out["recon_B"] = vp.shape

    # ---------------- render forward/backward through the compiled reference op
    rng = np.random.default_rng(5)
    cases = {}
    # (1) grid mesh, exact grid (pixel centres on shared edges), several poses incl. partly off-screen
    for tag, jitter, full in (("grid_exact", 0.0, False), ("grid_jitter", 0.2, True)):
        m = small_model((19, 25), 6, 4, 21, jitter)
        p = synth.sample_params_constrained(4, 6, 4, 48, seed=22, full_range=full)
        p[:, 6] = p[:, 6] * 0.25 + 1.5e-4                         # fit a 48x48 image
        if tag == "grid_exact":                                    # integer-aligned vertices: edges hit pixel centres
            p[0, 0:3] = 0
            p[0, 3:5] = 24
            p[0, 6] = 2.5e-4
            p[0, 7:] = 0
        from oracle import recon
        vp = recon.vertices_transform(p, m, 48, dtype=np.float32).astype(np.float32)
        cases[tag] = (vp, m["tri"], np.broadcast_to(m["vertex"], (4,) + m["vertex"].shape).copy(), (4, 48, 48, 3))
    # (2) random triangle soup with degenerate, duplicate, off-screen and huge/NaN-depth triangles
    nv, nt = 60, 90
    v = np.empty((2, 3, nv), np.float32)
    v[:, 0:2] = rng.uniform(-4, 36, (2, 2, nv))
    v[:, 0:2] = np.where(rng.random((2, 2, nv)) < 0.5, np.round(v[:, 0:2]), v[:, 0:2])   # many integer coords
    v[:, 2] = np.round(rng.uniform(-3, 3, (2, nv)))                                      # many depth ties
    v[0, 2, 0] = -2e14
    v[0, 2, 1] = np.inf
    v[1, 2, 2] = np.nan
    v[1, 2, 3] = -0.0
    t = rng.integers(0, nv, (3, nt)).astype(np.float32)
    t[:, 10] = t[:, 9]                      # duplicate triangle
    t[:, 11] = [5, 5, 7]                    # degenerate (repeated vertex)
    cases["soup"] = (v, t, rng.uniform(0, 1, (2, 3, nv)).astype(np.float32), (2, 32, 32, 3))
    # (3) SURVEY App. C known answers T1..T6 on an 8x8 image
    kat_v = np.array([[[0, 4, 0], [0, 0, 4], [1, 2, 4]]], np.float32)
    cases["kat_T5"] = (kat_v, np.array([[0], [1], [2]], np.float32), np.arange(9, dtype=np.float32).reshape(1, 3, 3), (1, 8, 8, 3))
    cases["kat_T2"] = (kat_v, np.array([[0, 0], [1, 1], [2, 2]], np.float32), np.zeros((1, 3, 3), np.float32), (1, 8, 8, 3))
    cases["kat_T3"] = (np.array([[[1, 3, 2], [1, 3, 2], [1, 1, 1]]], np.float32), np.array([[0], [1], [2]], np.float32),
                       np.zeros((1, 3, 3), np.float32), (1, 8, 8, 3))
    for name, xs in (("a", [-1.5, 2.5, -1.5]), ("b", [-0.5, 3.5, -0.5]), ("c", [3.5, 7.5, 3.5]), ("d", [4.0, 8.0, 4.0])):
        cases["kat_T4" + name] = (np.array([[xs, [0, 0, 4], [1, 1, 1]]], np.float32), np.array([[0], [1], [2]], np.float32),
                                  np.zeros((1, 3, 3), np.float32), (1, 8, 8, 3))
    cases["kat_T6"] = (np.array([[[0, 4, 0], [0, 0, 4], [-2e14, -2e14, -2e14]]], np.float32),
                       np.array([[0], [1], [2]], np.float32), np.zeros((1, 3, 3), np.float32), (1, 8, 8, 3))

    blob = {}
    for name, (vertex, tri, tex, ishape) in cases.items():
        depth, teximg, normal, tri_ind = oracle.ref_render_depth(vertex, tri, tex, ishape)
        dgrad = rng.normal(0, 1, depth.shape).astype(np.float32)
        vgrad = oracle.ref_render_depth_grad(dgrad, vertex, tri, depth, tri_ind, ishape, sanitize=True)
        for k, a in (("vertex", vertex), ("tri", tri), ("texture", tex), ("image_shape", np.array(ishape)),
                     ("depth", depth), ("texture_image", teximg), ("normal", normal), ("tri_ind", tri_ind),
                     ("depth_grad", dgrad), ("vertex_grad", vgrad)):
            blob[name + "/" + k] = a
        out["render_" + name] = int((tri_ind >= 0).sum())
    np.savez_compressed(os.path.join(HERE, "render_cases.npz"), **blob)

    # ---------------- (f) rows: rendering_layer post-processing, set_constraints, geometry loss -- reference source executed
    from oracle import recon
    m = small_model((23, 31), 12, 5, 3, 0.2)                         # == tests/conftest.py small_model
    B, S = 5, 64
    p = synth.sample_params_constrained(B, 12, 5, S, seed=11)
    vp = recon.vertices_transform(p, m, S, dtype=np.float32).astype(np.float32)
    zs = vp[:, 2, :]
    vp[:, 2, :] = ((zs - zs.mean()) / (zs.std() + 1e-9) * 0.6 + 0.5).astype(np.float32)   # depths straddle 1e-6 and 1: all clip regimes
    gray = rng.uniform(0, 1, (B, S, S, 1)).astype(np.float32)
    pncc, normalimg, maskimg, depthimg = reference_rendering_layer(vp, m["tri"], m["vertex"], gray)
    np.savez_compressed(os.path.join(HERE, "layer_cases.npz"), vertex_proj=vp, tri=m["tri"], colors=m["vertex"], im_gray=gray,
                        pncc=pncc, normalimg=normalimg, maskimg=maskimg, depthimg=depthimg)
    out["layer"] = (int((depthimg > 1e-6).sum()), int((normalimg != 0).any(-1).sum()))

    raw = rng.normal(scale=1.5, size=(7, 1, 1, 235)).astype(np.float32)
    raw[0, 0, 0, :8] = [0.0, 40.0, -40.0, 100.0, -100.0, 3.0, -3.0, 0.5]      # saturated / extreme sigmoid arguments
    np.savez_compressed(os.path.join(HERE, "constraints.npz"), raw=raw, im_size=200,
                        constrained=reference_set_constraints(raw, 200, 199, 29))
    out["constraints"] = raw.shape

    mg = small_model((9, 11), 199, 29, 14, 0.2)
    pred = synth.sample_params_constrained(6, seed=70)
    label = synth.sample_params_constrained(6, seed=71)
    np.savez_compressed(os.path.join(HERE, "geometry_loss.npz"), pred=pred, label=label, loss=reference_geometry_loss(mg, pred, label),
                        grid=np.array((9, 11)), **model_arrays(mg))
    out["geometry_loss"] = float(reference_geometry_loss(mg, pred, label))
    for k, v_ in out.items():
        print(k, v_)


if __name__ == "__main__":
    main()
