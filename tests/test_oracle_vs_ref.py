"""CPU: our C restatement side by side with the compiled, unmodified reference op (oracle/_ref).

Runs wherever oracle/_ref/libref_render_depth.so exists (built here from /root/reference; the built file
travels to the GPU box).  Everything is compared bit for bit."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle
from oracle import recon
from conftest import fr

pytestmark = pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built and reference tree absent")


def _same(a, b):
    return all(x.tobytes() == y.tobytes() for x, y in zip(a, b))


def test_reference_op_contract():
    """Registered signature and shape functions (render_depth_op.cc:535-589)."""
    assert oracle.ref_op_signature("RenderDepth") == ("vertex: float|tri: float|texture: float|image: float->"
                                                      "depth: float|texture_image: float|normal: float|tri_ind: float")
    assert oracle.ref_op_signature("RenderDepthGrad").endswith("->vertex_grad: float")
    assert oracle.ref_infer_shapes("RenderDepth", [(2, 3, 10), (3, 7), (2, 3, 10), (2, 20, 30, 3)]) == \
        [(2, 20, 30, 1), (2, 20, 30, 3), (2, 20, 30, 3), (2, 20, 30, 1)]
    assert oracle.ref_infer_shapes("RenderDepthGrad", [(2, 20, 30, 1), (2, 3, 10), (3, 7), (2, 20, 30, 1), (2, 20, 30, 1),
                                                       (2, 20, 30, 3)]) == [(2, 3, 10)]


@pytest.mark.parametrize("bad", ["batch", "vdim", "tdim", "texch"])
def test_reference_validation(bad):
    """OP_REQUIRES at render_depth_op.cc:408-418."""
    v = np.zeros((2, 4 if bad == "vdim" else 3, 5), np.float32)
    t = np.zeros((2 if bad == "tdim" else 3, 1), np.float32)
    x = np.zeros((2, 4 if bad == "texch" else 3, 5), np.float32)
    with pytest.raises(oracle.ReferenceInvalidArgument):
        oracle.ref_render_depth(v, t, x, (1 if bad == "batch" else 2, 8, 8, 3))


@pytest.mark.parametrize("jitter,full", [(0.0, False), (0.2, False), (0.2, True)])
def test_bfm_like_mesh(jitter, full):
    synth = fr("synth")
    m = synth.make_synthetic_model(grid=(49, 61), ndim_shape=10, ndim_exp=4, seed=5, jitter=jitter)
    p = synth.sample_params_constrained(3, 10, 4, 100, seed=6, full_range=full)
    p[:, 6] *= 0.5
    vp = recon.vertices_transform(p, m, 100, dtype=np.float32).astype(np.float32)
    tex = np.broadcast_to(m["vertex"], (3,) + m["vertex"].shape).copy()
    ref = oracle.ref_render_depth(vp, m["tri"], tex, (3, 100, 100, 3))
    mine = oracle.oracle_render_depth_forward(vp, m["tri"], tex, 100, 100)
    assert _same(ref, mine)
    shared = oracle.oracle_render_depth_forward(vp, m["tri"], m["vertex"], 100, 100)      # stride-0 texture
    assert _same(ref, shared)
    assert (ref[3] >= 0).sum() > 1000
    g = np.random.default_rng(1).normal(size=ref[0].shape).astype(np.float32)
    rg = oracle.ref_render_depth_grad(g, vp, m["tri"], ref[0], ref[3], (3, 100, 100, 3))
    og = oracle.oracle_render_depth_backward(g, m["tri"], ref[3], vp.shape[2])
    assert rg.tobytes() == og.tobytes()


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(3, 40), st.integers(1, 60), st.sampled_from([(8, 8), (16, 11), (5, 23)]),
       st.sampled_from([0.0, 0.5, 1.0]))
def test_random_soup(seed, nv, nt, hw, p_int):
    """Random triangle soups: integer-aligned vertices (edge hits), depth ties, off-screen triangles,
    degenerate triangles, +-inf / NaN / -0.0 depths."""
    H, W = hw
    rng = np.random.default_rng(seed)
    v = np.empty((2, 3, nv), np.float32)
    v[:, 0] = rng.uniform(-3, W + 2, (2, nv))
    v[:, 1] = rng.uniform(-3, H + 2, (2, nv))
    snap = rng.random((2, 2, nv)) < p_int
    v[:, 0:2] = np.where(snap, np.round(v[:, 0:2]), v[:, 0:2])
    v[:, 2] = np.round(rng.uniform(-2, 2, (2, nv)) * 2) / 2
    specials = np.array([np.inf, -np.inf, np.nan, -0.0, -2e14, 3e38], np.float32)
    k = rng.integers(0, nv, 3)
    v[rng.integers(0, 2, 3), 2, k] = specials[rng.integers(0, len(specials), 3)]
    t = rng.integers(0, nv, (3, nt)).astype(np.float32)
    tex = rng.uniform(0, 1, (2, 3, nv)).astype(np.float32)
    ref = oracle.ref_render_depth(v, t, tex, (2, H, W, 3))
    mine = oracle.oracle_render_depth_forward(v, t, tex, H, W)
    assert _same(ref, mine)
    g = rng.normal(size=ref[0].shape).astype(np.float32)
    assert oracle.ref_render_depth_grad(g, v, t, ref[0], ref[3], (2, H, W, 3)).tobytes() == \
        oracle.oracle_render_depth_backward(g, t, ref[3], nv).tobytes()


def test_extreme_coordinates():
    """Coordinates beyond int range and NaN xy: the x86 cvttsd2si behaviour of (int)ceil()
    (render_depth_op.cc:276-280) culls them; the restatement spells that out."""
    v = np.array([[[0, 4, 0, 1e20, -1e20, np.nan, 3e9, 2, 1],
                   [0, 0, 4, 1, 2, 3, 1, -3e9, np.inf],
                   [1, 1, 1, 1, 1, 1, 1, 1, 1]]], np.float32)
    t = np.array([[0, 0, 0, 3, 5, 0, 6, 8], [1, 1, 3, 1, 1, 7, 1, 1], [2, 4, 2, 2, 2, 2, 2, 2]], np.float32)
    tex = np.zeros_like(v)
    ref = oracle.ref_render_depth(v, t, tex, (1, 8, 8, 3))
    assert _same(ref, oracle.oracle_render_depth_forward(v, t, tex, 8, 8))
    assert set(np.unique(ref[3]).tolist()) == {-1.0, 0.0}
