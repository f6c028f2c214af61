"""GPU parity: the CUDA rasterizer (through the public render_depth API -> C ABI) against the CPU oracle and the
golden vectors produced by the reference.  Forward outputs are compared BIT FOR BIT; gradients within 1e-4."""
import numpy as np
import pytest

import oracle
from oracle import recon
from conftest import fr

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

DEV = "cuda:0"
NAMES = ("depth", "texture_image", "normal", "tri_ind")


MESH_MODES = ["now", None]       # triangles from a mesh table built on the spot (integer ids, cluster order) / from the float index tensor


def _gpu_render(vertex, tri, texture, H, W, expand_texture=False, mesh="now"):
    ops = fr("rendering_layer.ops")
    v = torch.from_numpy(np.ascontiguousarray(vertex, np.float32)).to(DEV)
    t = torch.from_numpy(np.ascontiguousarray(tri, np.float32)).to(DEV)
    x = torch.from_numpy(np.ascontiguousarray(texture, np.float32)).to(DEV)
    if expand_texture:
        x = x.unsqueeze(0).expand(v.shape[0], -1, -1)
    image = torch.empty((v.shape[0], H, W, 3), device=DEV)
    out = ops.render_depth(v, t, x, image, mesh=mesh)
    torch.cuda.synchronize()
    return [o.cpu().numpy() for o in out]


def _assert_same(got, want, tag=""):
    for g, w, n in zip(got, want, NAMES):
        assert g.shape == w.shape, (tag, n)
        if g.tobytes() != w.tobytes():
            bad = np.argwhere(g.view(np.uint32) != w.view(np.uint32))
            raise AssertionError("%s %s: %d mismatching elements, first at %s: got %r want %r" %
                                 (tag, n, len(bad), bad[0], g[tuple(bad[0])], w[tuple(bad[0])]))


def test_library_is_the_cuda_build():
    lib = fr("_lib").lib()
    assert lib.fr_version() == 202
    before = lib.fr_launch_count()
    _gpu_render(np.zeros((1, 3, 3), np.float32), np.array([[0], [1], [2]], np.float32), np.zeros((1, 3, 3), np.float32), 8, 8)
    assert lib.fr_launch_count() >= before + 2


@pytest.mark.parametrize("mesh", MESH_MODES)
def test_golden_cases_bit_exact(render_golden, mesh):
    for name, c in render_golden.items():
        B, H, W, _ = [int(x) for x in c["image_shape"]]
        got = _gpu_render(c["vertex"], c["tri"], c["texture"], H, W, mesh=mesh)
        _assert_same(got, [c[k] for k in NAMES], name)


def test_golden_backward(render_golden):
    ops = fr("rendering_layer.ops")
    for name, c in render_golden.items():
        B, H, W, _ = [int(x) for x in c["image_shape"]]
        image = torch.empty((B, H, W, 3), device=DEV)
        got = ops.render_depth_grad(torch.from_numpy(c["depth_grad"]).to(DEV), torch.from_numpy(c["vertex"]).to(DEV),
                                    torch.from_numpy(c["tri"]).to(DEV), torch.from_numpy(c["depth"]).to(DEV),
                                    torch.from_numpy(c["tri_ind"]).to(DEV), image).cpu().numpy()
        want = c["vertex_grad"]
        assert not got[:, 0:2].any()
        assert np.abs(got - want).max() <= 1e-4 * max(1.0, np.abs(want).max()), name


def _bfm_vertices(B, seed, jitter=0.2, im=200, full_range=False, permute=False):
    synth = fr("synth")
    m = synth.make_synthetic_model(ndim_shape=6, ndim_exp=3, seed=0, jitter=jitter, permute=permute)     # true N and T, small K (CPU recon)
    p = synth.sample_params_constrained(B, 6, 3, im, seed=seed, full_range=full_range)
    p[:, 7:13] *= 10.0                                                                   # keep ~2 px deformations with 6 modes
    vp = recon.vertices_transform(p, m, im, dtype=np.float32).astype(np.float32)
    return m, vp


@pytest.mark.parametrize("jitter,full_range,B,permute", [(0.2, False, 4, False), (0.0, False, 3, False), (0.2, True, 5, False),
                                                         (0.2, False, 1, False), (0.2, False, 3, True), (0.0, False, 2, True)])
def test_bfm_size_forward_bit_exact(jitter, full_range, B, permute):
    """53 215 vertices / 105 840 triangles / 200x200 (BASELINE configs 1-2 geometry) on a shared float32 vertex buffer, in the
    generator's grid order and with randomly renumbered vertices / shuffled triangles; both rasterizers (cluster / generic)."""
    m, vp = _bfm_vertices(B, seed=4, jitter=jitter, full_range=full_range, permute=permute)
    assert m["tri"].shape == (3, 105840) and vp.shape == (B, 3, 53215)
    want = oracle.oracle_render_depth_forward(vp, m["tri"], m["vertex"], 200, 200)
    for mesh in MESH_MODES:
        got = _gpu_render(vp, m["tri"], m["vertex"], 200, 200, expand_texture=True, mesh=mesh)
        _assert_same(got, want, "bfm mesh=%r" % (mesh,))
        got2 = _gpu_render(vp, m["tri"], m["vertex"], 200, 200, expand_texture=True, mesh=mesh)         # run-to-run determinism
        _assert_same(got2, got, "determinism")
    if not full_range:
        assert (want[3] >= 0).mean() > 0.3


def test_bfm_size_backward_and_autograd():
    ops = fr("rendering_layer.ops")
    m, vp = _bfm_vertices(3, seed=6)
    v = torch.from_numpy(vp).to(DEV).requires_grad_(True)
    tri = torch.from_numpy(m["tri"]).to(DEV)
    tex = torch.from_numpy(m["vertex"]).to(DEV).unsqueeze(0).expand(3, -1, -1)
    image = torch.empty((3, 200, 200, 3), device=DEV)
    depth, teximg, normal, tri_ind = ops.render_depth(v, tri, tex, image)
    assert depth.requires_grad and not tri_ind.requires_grad and not normal.requires_grad
    g = torch.from_numpy(np.random.default_rng(0).normal(size=(3, 200, 200, 1)).astype(np.float32)).to(DEV)
    g = g * (tri_ind >= 0)                                            # tf.maximum gating, nets/network.py:199
    (depth * g).sum().backward()
    want = oracle.oracle_render_depth_backward(g.cpu().numpy(), m["tri"], tri_ind.cpu().numpy(), vp.shape[2])
    got = v.grad.cpu().numpy()
    assert not got[:, 0:2].any()
    assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max()
    # checksum property: sum of vertex_grad == sum of depth_grad over covered pixels
    assert abs(got.astype(np.float64).sum() - g.cpu().numpy().astype(np.float64).sum()) <= 1e-3 * np.abs(g.cpu().numpy()).sum()


def test_validation_errors():
    ops = fr("rendering_layer.ops")
    z = lambda *s: torch.zeros(s, device=DEV)
    with pytest.raises(ValueError, match="vertex's batch"):
        ops.render_depth(z(2, 3, 5), z(3, 1), z(2, 3, 5), z(1, 8, 8, 3))
    with pytest.raises(ValueError, match="Batch x 3 x nver"):
        ops.render_depth(z(2, 4, 5), z(3, 1), z(2, 3, 5), z(2, 8, 8, 3))
    with pytest.raises(ValueError, match="3 x ntri"):
        ops.render_depth(z(2, 3, 5), z(2, 1), z(2, 3, 5), z(2, 8, 8, 3))
    with pytest.raises(ValueError, match="texture channel"):
        ops.render_depth(z(2, 3, 5), z(3, 1), z(2, 4, 5), z(2, 8, 8, 3))
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.render_depth(torch.zeros(2, 3, 5), z(3, 1), z(2, 3, 5), z(2, 8, 8, 3))


def test_edge_shapes():
    """Empty triangle list, out-of-range indices (skipped), non-square image, huge coordinates."""
    v = np.array([[[0, 4, 0, 1e20, np.nan], [0, 0, 4, 1, 2], [1, 1, 1, 1, 1]]], np.float32)
    tex = np.zeros_like(v)
    got = _gpu_render(v, np.zeros((3, 0), np.float32), tex, 5, 9)
    assert (got[3] == -1).all() and (got[0].view(np.uint32) == 0xD6B5E621).all() and not got[1].any() and not got[2].any()
    tri = np.array([[0, 0, 0, 7], [1, 3, 4, 1], [2, 2, 2, 2]], np.float32)            # last column: index 7 >= nver
    want = oracle.oracle_render_depth_forward(v, tri[:, :3], tex, 5, 9)
    for mesh in MESH_MODES:
        _assert_same(_gpu_render(v, tri, tex, 5, 9, mesh=mesh), want, "edge")
        got = _gpu_render(v, np.zeros((3, 0), np.float32), tex, 5, 9, mesh=mesh)
        assert (got[3] == -1).all() and (got[0].view(np.uint32) == 0xD6B5E621).all()


def test_signed_zero_depths_and_zero_params():
    """Depth +-0.0: the key folds the two for ordering (they tie, lowest index wins) but keeps the sign for the output, so the
    resolve pass never needs the vertices (round-1 advisor finding: a null vertex pointer was read on this path).  Includes
    the fused call with vertex_proj = NULL and all-zero parameters, where every vertex collapses onto one point at depth 0
    and every (degenerate) triangle paints that pixel."""
    v = np.array([[[0, 4, 0, 0, 4, 0], [0, 0, 4, 0, 0, 4], [-0.0, -0.0, -0.0, 0.0, 0.0, 0.0]]], np.float32)
    for tri in (np.array([[0, 3], [1, 4], [2, 5]], np.float32), np.array([[3, 0], [4, 1], [5, 2]], np.float32)):
        want = oracle.oracle_render_depth_forward(v, tri, np.zeros_like(v), 8, 8)
        for mesh in MESH_MODES:
            _assert_same(_gpu_render(v, tri, np.zeros_like(v), 8, 8, mesh=mesh), want, "signed zero")
    lib, check = fr("_lib").lib(), fr("_lib").check
    model = fr("synth").make_synthetic_model(grid=(23, 31), ndim_shape=12, ndim_exp=5, seed=3, jitter=0.2)
    for tiles, B in ((False, 3), (False, 20), (True, 20)):                 # FFMA path / tcgen05 + records / tcgen05 + tile rasterizer in the epilogue
        dm = fr("model").DeviceModel(model, DEV, cluster_tiles=tiles)
        S = 32
        params = torch.zeros((B, dm.ndim), device=DEV)
        ws = torch.empty(lib.fr_pipeline_workspace_bytes(B, dm.nver, dm.ndim_shape, dm.ndim_exp, S, S), dtype=torch.uint8, device=DEV)
        d, t = torch.empty((B, S, S, 1), device=DEV), torch.empty((B, S, S, 1), device=DEV)
        check(lib.fr_recon_render_forward(params.data_ptr(), dm.packed.data_ptr(), dm.tri.data_ptr(), dm.mesh.handle, None, d.data_ptr(),
                                          t.data_ptr(), B, dm.nver, dm.ntri, dm.ndim_shape, dm.ndim_exp, S, S, float(S), dm.run_flags,
                                          ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream, None))
        torch.cuda.synchronize()
        vp = recon.vertices_transform(params.cpu().numpy(), model, S, dtype=np.float32).astype(np.float32)
        want = oracle.oracle_render_depth_forward(vp, model["tri"], np.zeros_like(vp), S, S)
        assert (want[3] >= 0).sum() == B                                    # one painted pixel per face: (0, S-1)
        assert d.cpu().numpy().tobytes() == want[0].tobytes() and t.cpu().numpy().tobytes() == want[3].tobytes()


def test_full_batch_properties():
    """BASELINE config-2 size (B = 64): properties that need no oracle run per face + oracle on a sample of faces."""
    B = 64
    m, vp = _bfm_vertices(B, seed=2)
    got = _gpu_render(vp, m["tri"], m["vertex"], 200, 200, expand_texture=True)
    _assert_same(_gpu_render(vp, m["tri"], m["vertex"], 200, 200, expand_texture=True, mesh=None), got, "table order vs reference order")
    depth, teximg, normal, tri_ind = got
    covered = tri_ind[..., 0] >= 0
    assert (depth[..., 0][~covered].view(np.uint32) == 0xD6B5E621).all()
    assert np.isfinite(depth[..., 0][covered]).all() and (tri_ind[..., 0] < m["tri"].shape[1]).all()
    # depth of a covered pixel is the flat depth of its triangle
    b_idx, y_idx, x_idx = np.nonzero(covered)
    t = tri_ind[b_idx, y_idx, x_idx, 0].astype(np.int64)
    idx = m["tri"].astype(np.int64)[:, t]
    z = vp[b_idx[None, :], 2, idx]
    assert np.array_equal(depth[b_idx, y_idx, x_idx, 0], ((z[0] + z[1]) + z[2]) / np.float32(3.0))
    # batch permutation invariance and independence of faces
    perm = np.random.default_rng(1).permutation(B)
    got_p = _gpu_render(vp[perm], m["tri"], m["vertex"], 200, 200, expand_texture=True)
    _assert_same(got_p, [g[perm] for g in got], "perm")
    for b in (0, 17, 63):
        want = oracle.oracle_render_depth_forward(vp[b:b + 1], m["tri"], m["vertex"], 200, 200)
        _assert_same([g[b:b + 1] for g in got], want, "face %d" % b)


def test_unsure_queue_overflow():
    """A large degenerate (collinear) triangle paints its whole bbox (render_depth_op.cc:105-109) and makes every pixel of it
    'unsure' for the float pre-filter: > 1024 unsure pixels in one block overflow the deferred FP64 queue and exercise
    the redo path of raster_keys_kernel; big slivers and pixel centres exactly on edges ride along."""
    rng = np.random.default_rng(7)
    nv = 40
    v = np.empty((9, 3, nv), np.float32)                       # 9 faces: exercises the 8-faces-per-thread grouping + a tail
    v[:, 0] = rng.uniform(2, 60, (9, nv))
    v[:, 1] = rng.uniform(2, 60, (9, nv))
    v[:, 2] = rng.uniform(-1, 1, (9, nv))
    v[:, 0:2, 0] = [3.0, 3.0]
    v[:, 0:2, 1] = [58.0, 58.0]
    v[:, 0:2, 2] = [30.0, 30.0]                                 # collinear with vertices 0 and 1 -> 56x56 painted bbox
    v[:, 2, 0:3] = 5.0                                          # ... nearest, so it survives the z-buffer
    v[:, 0:2, 3] = [3.0, 50.0]
    v[:, 0:2, 4] = [60.5, 50.000004]                            # long sliver
    v[:, 0:2, 5] = [20.0, 50.00001]
    v[:, 0:2, 6:12] = np.round(v[:, 0:2, 6:12])                 # integer vertices: edges through pixel centres
    tri = rng.integers(0, nv, (3, 300)).astype(np.float32)
    tri[:, 0] = [0, 1, 2]
    tri[:, 1] = [3, 4, 5]
    tex = rng.uniform(0, 1, (9, 3, nv)).astype(np.float32)
    want = oracle.oracle_render_depth_forward(v, tri, tex, 64, 64)
    assert (want[3][0] == 0).sum() > 1024                        # the degenerate triangle really covers > queue capacity
    for mesh in MESH_MODES:
        _assert_same(_gpu_render(v, tri, tex, 64, 64, mesh=mesh), want, "overflow")


def test_rendering_layer_fused_matches_unfused(small_model):
    """FaceRecNet.rendering_layer with the post-processing fused into the resolve kernel against the literal
    torch transcription of network.py:184-199 -- outputs and the gradient w.r.t. the vertices (depthimg and maskimg paths,
    clip gates included: the face is pushed through z = 0 and z = 1 so that all three regimes occur)."""
    net_mod = fr("nets.network")
    B, S = 5, 64
    ks, ke = small_model["ndim_shape"], small_model["ndim_exp"]
    gray = torch.rand((B, S, S, 1), device=DEV)
    net = net_mod.FaceRecNet(im_gray=gray, mesh_data=small_model, batch_size=B, im_size=S, device=DEV)
    p = fr("synth").sample_params_constrained(B, ks, ke, S, seed=11)
    vp0 = net.vertices_transform(torch.from_numpy(p).to(DEV)[:, None, None, :]).detach()
    zs = vp0[:, 2, :]
    vp0[:, 2, :] = (zs - zs.mean()) / (zs.std() + 1e-9) * 0.6 + 0.5       # depths straddle 1e-6 and 1
    outs, grads = [], []
    for fused in (True, False):
        vp = vp0.clone().requires_grad_(True)
        layer = net.rendering_layer if fused else net.rendering_layer_unfused
        pncc, normalimg, maskimg, depthimg = layer(vp, net.tri, net.vertex_code)
        w1 = torch.linspace(0.5, 1.5, depthimg.numel(), device=DEV).view_as(depthimg)
        w2 = torch.linspace(-1.0, 2.0, maskimg.numel(), device=DEV).view_as(maskimg)
        ((depthimg * w1).sum() + (maskimg * w2).sum()).backward()
        outs.append([t.detach().cpu().numpy() for t in (pncc, normalimg, maskimg, depthimg)])
        grads.append(vp.grad.cpu().numpy())
    covered = outs[1][3] > 1e-6
    assert covered.mean() > 0.02 and (outs[1][2] >= 1.0 * 0.0).all()
    for a, b, name in zip(outs[0], outs[1], ("pncc", "normalimg", "maskimg", "depthimg")):
        if name == "normalimg":
            assert np.abs(a - b).max() <= 2e-6, name                      # sqrt / divide / 3-term sum order
        else:
            assert a.tobytes() == b.tobytes(), name
    assert np.abs(grads[0]).max() > 0
    assert np.abs(grads[0] - grads[1]).max() <= 1e-6 * np.abs(grads[1]).max()


def test_rendering_layer_pinned_to_reference_source(small_model):
    """SURVEY 8f-1 pinned: tests/golden/layer_cases.npz holds the four outputs of the reference's own
    FaceRecNet.rendering_layer source (nets/network.py:172-201) executed by make_golden.py with render_depth bound to the
    compiled reference op.  The fused layer (post-processing inside the resolve kernel) and the un-fused torch mirror must
    reproduce them: bit for bit where only selections / one multiply are involved, 2e-6 for the normalised normals."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "layer_cases.npz"))
    B, S = g["im_gray"].shape[0], g["im_gray"].shape[1]
    assert np.array_equal(g["tri"], small_model["tri"])                 # the golden was made on this very model
    net = fr("nets.network").FaceRecNet(im_gray=torch.from_numpy(g["im_gray"]).to(DEV), mesh_data=small_model, batch_size=B, im_size=S,
                                        device=DEV)
    vp = torch.from_numpy(g["vertex_proj"]).to(DEV)
    colors = torch.from_numpy(g["colors"]).to(DEV)
    assert (g["depthimg"] > 1e-6).mean() > 0.02 and (g["depthimg"] == np.float32(1e-6)).any() and (g["maskimg"] > 0).any()
    for layer in (net.rendering_layer, net.rendering_layer_unfused):
        got = [t.cpu().numpy() for t in layer(vp, net.tri, colors)]
        for a, name in zip(got, ("pncc", "normalimg", "maskimg", "depthimg")):
            w = g[name]
            assert a.shape == w.shape, name
            if name == "normalimg":
                assert np.abs(a - w).max() <= 2e-6, (layer.__name__, name)
            else:
                assert a.tobytes() == w.tobytes(), (layer.__name__, name)


def test_full_size_properties_batch_256():
    """BASELINE config sizes (53 215 vertices, 105 840 triangles, 200 x 200), 256 faces -- too many for the CPU oracle in a
    test, so size-independent properties instead: (a) the 256-face call equals four 64-face calls byte for byte,
    (b) a repeated face gives repeated maps, (c) every covered pixel's depth is exactly the float mean of its winner's three
    vertex depths and every background pixel holds the reference's background value, (d) the winner really covers the pixel
    centre (its bounding box contains it), (e) coverage is plausible."""
    lib, check = fr("_lib").lib(), fr("_lib").check
    synth = fr("synth")
    model = synth.make_synthetic_model(seed=0, jitter=0.2)
    dm = fr("model").DeviceModel(model, DEV)
    B, S = 256, 200
    p = synth.sample_params_constrained(B, seed=12)
    p[200] = p[7]                                                          # (b)
    pt = torch.from_numpy(p).to(DEV)
    sp = torch.cuda.current_stream().cuda_stream

    def fused(params, nb):
        ws = torch.empty(lib.fr_pipeline_workspace_bytes(nb, dm.nver, dm.ndim_shape, dm.ndim_exp, S, S), dtype=torch.uint8, device=DEV)
        vp = torch.empty((nb, 3, dm.nver), device=DEV)
        d, t = torch.empty((nb, S, S, 1), device=DEV), torch.empty((nb, S, S, 1), device=DEV)
        check(lib.fr_recon_render_forward(params.data_ptr(), dm.packed.data_ptr(), dm.tri.data_ptr(), dm.mesh.handle, vp.data_ptr(),
                                          d.data_ptr(), t.data_ptr(), nb, dm.nver, dm.ntri, dm.ndim_shape, dm.ndim_exp, S, S, float(S),
                                          dm.run_flags, ws.data_ptr(), ws.numel(), sp, None))
        torch.cuda.synchronize()
        return vp, d, t

    vp, depth, tri_ind = fused(pt, B)
    for q in range(4):                                                     # (a)
        vq, dq, tq = fused(pt[64 * q:64 * (q + 1)].contiguous(), 64)
        assert torch.equal(dq, depth[64 * q:64 * (q + 1)]) and torch.equal(tq, tri_ind[64 * q:64 * (q + 1)])
        assert torch.equal(vq, vp[64 * q:64 * (q + 1)])
    assert torch.equal(depth[200], depth[7]) and torch.equal(tri_ind[200], tri_ind[7])
    covered = tri_ind[..., 0] >= 0
    frac = covered.float().mean().item()
    assert 0.15 < frac < 0.95, frac                                        # (e)
    bidx = torch.arange(B, device=DEV)[:, None, None].expand(-1, S, S)[covered]
    t = tri_ind[..., 0][covered].long()
    i1, i2, i3 = (dm.tri[k].long()[t] for k in range(3))
    z = vp[:, 2, :]
    zsum = (z[bidx, i1] + z[bidx, i2]) + z[bidx, i3]
    want = zsum / torch.full_like(zsum, 3.0)                               # (c) render_depth_op.cc:217 -- tensor / tensor: an IEEE divide
                                                                           # (tensor / python scalar multiplies by the reciprocal)
    assert torch.equal(depth[..., 0][covered], want)
    assert (depth[..., 0][~covered] == -99999999999999.0).all()
    ys, xs = torch.meshgrid(torch.arange(S, device=DEV), torch.arange(S, device=DEV), indexing="ij")
    px = xs[None].expand(B, -1, -1)[covered].float()
    py = ys[None].expand(B, -1, -1)[covered].float()
    x = torch.stack([vp[:, 0, :][bidx, i] for i in (i1, i2, i3)])
    y = torch.stack([vp[:, 1, :][bidx, i] for i in (i1, i2, i3)])
    assert bool(((x.min(0).values <= px) & (px <= x.max(0).values) & (y.min(0).values <= py) & (py <= y.max(0).values)).all())   # (d)


@pytest.mark.parametrize("B", [9, 17, 33])
def test_tile_rasterizer_on_golden_cases(render_golden, B):
    """The stand-alone tile rasterizer (raster_tile.cuh) only runs from 8 faces up: the golden cases (degenerate, duplicate,
    off-screen, NaN / inf-depth triangles, -0.0 depths, known answers T1-T6) replicated to B faces in rotated order, every face
    bit-identical to the reference op's output for that face; B = 9 / 17 / 33 leave a ragged last tile."""
    for name, c in render_golden.items():
        b0, H, W, _ = [int(x) for x in c["image_shape"]]
        sel = (np.arange(B) * 3 + 1) % b0
        got = _gpu_render(c["vertex"][sel], c["tri"], c["texture"][sel] if c["texture"].shape[0] == b0 else c["texture"], H, W, mesh="now")
        _assert_same(got, [c[k][sel] for k in NAMES], "%s B=%d" % (name, B))


def test_tile_rasterizer_large_boxes_and_exact_grid():
    """Close-up faces: every triangle covers tens of pixels (the plane-equation path of the certified inside test with large
    boxes and its growing thresholds), vertices on integer coordinates (pixel centres on shared edges and vertices: everything
    goes through the literal PointInTri pass) and a mix with sub-pixel slivers, at 12 faces through the tile rasterizer."""
    synth = fr("synth")
    m = synth.make_synthetic_model(grid=(19, 23), ndim_shape=4, ndim_exp=2, seed=5, jitter=0.3)
    B, S = 12, 96
    p = synth.sample_params_constrained(B, 4, 2, S, seed=9)
    p[:, 6] *= np.linspace(0.6, 1.4, B)                                 # ~4 .. 10 px between grid vertices
    p[:, 7:11] *= 3.0
    vp = recon.vertices_transform(p, m, S, dtype=np.float32).astype(np.float32)
    vp[3] = np.round(vp[3])                                             # integer vertices
    vp[4, 0:2] = np.round(vp[4, 0:2] * 2.0) / 2.0                       # half-integer
    vp[5, 1] = vp[5, 1] * 1e-3 + 40.0                                   # squashed: slivers along a pixel row
    vp[6, 0:2] += 1e4                                                   # entirely off-screen
    vp[7, 2, ::7] = np.nan                                              # NaN depths never draw
    want = oracle.oracle_render_depth_forward(vp, m["tri"], m["vertex"], S, S)
    assert (want[3][0] >= 0).mean() > 0.2 and (want[3][6] < 0).all()
    for mesh in ("now", None):
        got = _gpu_render(vp, m["tri"], m["vertex"], S, S, expand_texture=True, mesh=mesh)
        _assert_same(got, want, "close-up mesh=%r" % (mesh,))


def test_images_beyond_the_tile_rasterizer_limit():
    """The tile rasterizer's packed cull needs snap codes below 2^15 (images up to 16 000 px a side); a 16 500 x 2 image at 9
    faces with a mesh table must take the gather kernel and still match the reference bit for bit."""
    rng = np.random.default_rng(11)
    nv, B, H, W = 60, 9, 2, 16500
    v = np.empty((B, 3, nv), np.float32)
    v[:, 0] = rng.uniform(16380.0, 16499.0, (B, nv))            # around and beyond x = 16384
    v[:, 1] = rng.uniform(-0.5, 1.9, (B, nv))
    v[:, 2] = rng.uniform(-1, 1, (B, nv))
    tri = rng.integers(0, nv, (3, 200)).astype(np.float32)
    tex = rng.uniform(0, 1, (B, 3, nv)).astype(np.float32)
    want = oracle.oracle_render_depth_forward(v, tri, tex, H, W)
    assert (want[3] >= 0).sum() > 50
    for mesh in MESH_MODES:
        _assert_same(_gpu_render(v, tri, tex, H, W, mesh=mesh), want, "wide image mesh=%r" % (mesh,))
