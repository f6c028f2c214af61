"""CPU: the oracle restatements against the committed golden vectors (outputs of the reference itself,
see tests/golden/make_golden.py).  Bit-exact for the rasterizer, 1e-6 relative for the float GEMM part."""
import os

import numpy as np
import pytest

import oracle
from oracle import recon
from conftest import GOLDEN


def test_render_forward_bit_exact(render_golden):
    assert len(render_golden) >= 10
    for name, c in render_golden.items():
        B, H, W, _ = [int(x) for x in c["image_shape"]]
        got = oracle.oracle_render_depth_forward(c["vertex"], c["tri"], c["texture"], H, W)
        for g, key in zip(got, ("depth", "texture_image", "normal", "tri_ind")):
            assert g.shape == c[key].shape, (name, key)
            assert g.tobytes() == c[key].tobytes(), (name, key)


def test_render_backward_matches_reference(render_golden):
    for name, c in render_golden.items():
        nver = c["vertex"].shape[2]
        got = oracle.oracle_render_depth_backward(c["depth_grad"], c["tri"], c["tri_ind"], nver)
        # same pixel order and same float adds as render_depth_op.cc:346-363 => bit-exact
        assert got.tobytes() == c["vertex_grad"].tobytes(), name
        assert not got[:, 0:2].any()


def test_known_answers(render_golden):
    """SURVEY.md App. C T1-T6, produced by the reference code."""
    t5 = render_golden["kat_T5"]
    rows = ["".join("0" if v == 0 else "-" for v in r) for r in t5["tri_ind"][0, :, :, 0]]
    assert rows[:5] == ["0000----", "000-----", "00------", "0-------", "--------"]       # T1: hypotenuse, P2, P3 excluded
    assert t5["depth"][0, 0, 0, 0] == np.float32(2.33333325)                               # T5
    assert t5["normal"][0, 0, 0].tolist() == [-4.0, -12.0, 16.0]
    assert t5["texture_image"][0, 0, 0].tolist() == [1.0, 4.0, 7.0]
    assert t5["depth"][0, 7, 7, 0].view(np.uint32) == 0xD6B5E621                           # background depth bits
    assert (render_golden["kat_T2"]["tri_ind"] <= 0).all()                                 # T2: lowest index wins ties
    t3 = render_golden["kat_T3"]["tri_ind"][0, :, :, 0]
    assert (t3[1:4, 1:4] == 0).all() and (t3 >= 0).sum() == 9                               # T3: degenerate paints bbox
    assert (render_golden["kat_T4a"]["tri_ind"] == -1).all()                               # T4: whole-triangle cull
    assert (render_golden["kat_T4b"]["tri_ind"] >= 0).sum() == 10
    assert (render_golden["kat_T4c"]["tri_ind"] >= 0).sum() == 10
    assert (render_golden["kat_T4d"]["tri_ind"] == -1).all()
    assert (render_golden["kat_T6"]["tri_ind"] == -1).all()                                # T6: below background depth


@pytest.mark.parametrize("tag", ["tiny", "truek"])
def test_recon_variant_A(tag):
    g = np.load(os.path.join(GOLDEN, "recon_A_%s.npz" % tag))
    model = {k: g[k] for k in ("mu", "pc_shape", "pc_exp")}
    ref = g["vertex_proj"]
    scale = np.abs(ref).max()
    for dtype in (np.float64, np.float32):
        vp = recon.vertices_transform(g["params"], model, int(g["im_size"]), dtype=dtype)
        assert np.abs(vp - ref).max() <= 1e-6 * scale
    assert np.array_equal(recon.rotation_matrix_batch(g["params"][:, :3]), g["rot"])


def test_recon_variant_B():
    g = np.load(os.path.join(GOLDEN, "recon_B_sample_test.npz"))
    model = {k: g[k] for k in ("mu", "pc_shape", "pc_exp")}
    vp = recon.vertices_transform(g["params"], model, int(g["im_size"]), conv=recon.VARIANT_B)[0]
    assert np.abs(vp - g["vertex_proj"]).max() <= 1e-6 * np.abs(g["vertex_proj"]).max()
    assert np.array_equal(recon.rotation_matrix_batch(g["angles"], "zyx"), g["rots"])
    assert g["params"][0, :7].tolist() == np.array([0, 0, 0, 100, 100, 0, 0.001], np.float32).astype(np.float64).tolist()


def test_recon_backward_matches_autograd(small_model):
    """App. A.4 closed form == torch autograd through a float64 re-host of network.py:140-171."""
    import torch
    from conftest import fr
    m = small_model
    ks, ke = m["ndim_shape"], m["ndim_exp"]
    p = fr("synth").sample_params_constrained(3, ks, ke, 64, seed=9, full_range=True).astype(np.float64)
    n = m["mu"].shape[0] // 3
    g = np.random.default_rng(0).normal(size=(3, 3, n))
    pt = torch.tensor(p, requires_grad=True)
    rot = torch.tensor(recon.rotation_matrix_batch(p[:, :3]).astype(np.float64))        # constant: py_func has no grad
    v = (torch.tensor(m["mu"].astype(np.float64)).reshape(1, 3, n)
         + (pt[:, 7:7 + ks] @ torch.tensor(m["pc_shape"].astype(np.float64)).T).reshape(3, 3, n)
         + (pt[:, 7 + ks:] @ torch.tensor(m["pc_exp"].astype(np.float64)).T).reshape(3, 3, n))
    vp = pt[:, 6].reshape(3, 1, 1) * rot @ v + pt[:, 3:6].reshape(3, 3, 1)
    vp = torch.cat([vp[:, 0:1], 64 - vp[:, 1:2] - 1, vp[:, 2:3]], dim=1)
    assert np.allclose(vp.detach().numpy(), recon.vertices_transform(p, m, 64), rtol=1e-12, atol=1e-9)
    (vp * torch.tensor(g)).sum().backward()
    mine = recon.vertices_transform_backward(p, m, g)
    assert np.allclose(mine, pt.grad.numpy(), rtol=1e-9, atol=1e-9)
    assert not mine[:, 0:3].any()
