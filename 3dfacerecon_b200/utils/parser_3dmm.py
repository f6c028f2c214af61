"""3DMM model I/O with the reference's dict layout (``utils/parser_3dmm.py:6-61``).

``read_3dmm_model(path)`` returns exactly the keys the reference returns, so ``trainval.py`` /
``sample_test.py`` style callers keep working:
``vertex`` (PNCC code [3,N]), ``tri`` [3,T], ``mu`` [3N,1] (= mu_shape + mu_exp), ``mu_tex`` [3,N], ``pc_tex``,
``param_tex``, ``pc_shape`` [3N,199], ``pc_exp`` [3N,29], ``ndim_shape``, ``ndim_exp``, ``ndim_pose`` (= 7).
The BFM ``.mat`` files are not shipped with the reference (``3dmm/.gitignore``); ``synthetic_3dmm_model`` builds a
stand-in of identical shapes for offline runs.
"""
from __future__ import annotations

import os

import numpy as np

from ..synth import NDIM_POSE, make_synthetic_model

MODEL_FILES = ("Model_Shape.mat", "Model_Expression.mat", "vertex_code.mat")


def parse_3dmm_files(model_shapefile, model_expfile, vertex_codefile):
    """Load the three Matlab files (``utils/parser_3dmm.py:6-33``); returns the same 8-tuple."""
    import scipy.io as sio
    for f in (model_shapefile, model_expfile, vertex_codefile):
        if not os.path.exists(f):
            raise FileNotFoundError("File %s does not exist!" % f)
    shape = sio.loadmat(model_shapefile)
    exp = sio.loadmat(model_expfile)
    code = sio.loadmat(vertex_codefile)
    mu = shape["mu_shape"] + exp["mu_exp"]                       # utils/parser_3dmm.py:32
    return code["vertex_code"], shape["tri"], mu, shape["w"], exp["w_exp"], shape["tex"], shape["w_tex"], shape["alpha_tex"]


def read_3dmm_model(model_path):
    """``utils/parser_3dmm.py:36-61``."""
    vertex_code, tri, mu, pc_shape, pc_exp, mu_tex, pc_tex, param_tex = parse_3dmm_files(
        *[os.path.join(model_path, f) for f in MODEL_FILES])
    return {"vertex": vertex_code, "tri": tri, "mu": mu, "mu_tex": mu_tex, "pc_tex": pc_tex, "param_tex": param_tex,
            "pc_shape": pc_shape, "pc_exp": pc_exp, "ndim_shape": int(np.shape(pc_shape)[1]),
            "ndim_exp": int(np.shape(pc_exp)[1]), "ndim_pose": NDIM_POSE}


def synthetic_3dmm_model(**kwargs):
    """Offline stand-in with the true BFM dimensions (see ``synth.make_synthetic_model``)."""
    return make_synthetic_model(**kwargs)


def tri_is_one_based(tri, nver) -> bool:
    """The real ``Model_Shape.mat`` stores MATLAB 1-based indices which the reference uses unshifted
    (SURVEY.md App. B-7); callers should pass ``tri - 1`` in that case."""
    t = np.asarray(tri)
    return bool(t.min() >= 1 and t.max() == nver)
