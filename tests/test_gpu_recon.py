"""GPU parity: 3DMM reconstruction + projection kernels (recon_project / FaceRecNet.vertices_transform -> C ABI) against
the numpy float64 restatement of nets/network.py:140-171 and the golden vectors made by executing the reference.
Tolerances (BASELINE.json north_star): vertices 1e-5 relative, gradients 1e-4 relative (norm-wise)."""
import os

import numpy as np
import pytest

import oracle
from oracle import recon
from conftest import GOLDEN, fr

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
DEV = "cuda:0"
VERT_TOL = 1e-5
GRAD_TOL = 1e-4


def _model_from_golden(g):
    n = g["mu"].shape[0] // 3
    return {"mu": g["mu"], "pc_shape": g["pc_shape"], "pc_exp": g["pc_exp"], "tri": g["tri"], "vertex": g["vertex"],
            "mu_tex": g["mu_tex"], "ndim_shape": g["pc_shape"].shape[1], "ndim_exp": g["pc_exp"].shape[1], "ndim_pose": 7}


def _gpu_vertices(model, params, im_size, convention="network", cluster_tiles=False):
    dm = fr("model").DeviceModel(model, DEV, convention, cluster_tiles=cluster_tiles)
    out = fr("nets.network").recon_project(torch.from_numpy(np.asarray(params, np.float32)).to(DEV), dm, im_size)
    torch.cuda.synchronize()
    return dm, out.cpu().numpy()


@pytest.mark.parametrize("tag", ["tiny", "truek"])
def test_golden_variant_A(tag):
    g = np.load(os.path.join(GOLDEN, "recon_A_%s.npz" % tag))
    _, got = _gpu_vertices(_model_from_golden(g), g["params"], int(g["im_size"]))
    assert np.abs(got - g["vertex_proj"]).max() <= VERT_TOL * np.abs(g["vertex_proj"]).max()


def test_golden_variant_B():
    g = np.load(os.path.join(GOLDEN, "recon_B_sample_test.npz"))
    _, got = _gpu_vertices(_model_from_golden(g), g["params"], int(g["im_size"]), "sample_test")
    assert np.abs(got[0] - g["vertex_proj"]).max() <= VERT_TOL * np.abs(g["vertex_proj"]).max()


@pytest.fixture(scope="module")
def bfm():
    return fr("synth").make_synthetic_model(seed=0, jitter=0.2)          # true dims: 53 215 / 105 840 / 199 / 29


@pytest.mark.parametrize("B,full,tiles", [(1, False, False), (3, True, False), (8, False, False), (16, True, False), (20, False, False),
                                          (64, False, False), (70, True, False), (3, False, True), (20, True, True), (70, False, True)])
def test_bfm_size_forward(bfm, B, full, tiles):
    """All forward kernels (FFMA for <= 8 faces, tcgen05 above) with consecutive row tiles and with the mesh table's clusters
    as row tiles (FR_CLUSTER_TILES: border vertices are computed by several clusters, written by their owner)."""
    p = fr("synth").sample_params_constrained(B, seed=2 + B, full_range=full)
    _, got = _gpu_vertices(bfm, p, 200, cluster_tiles=tiles)
    want = recon.vertices_transform(p, bfm, 200)                           # float64 ground truth
    assert got.shape == want.shape == (B, 3, 53215)
    err = np.abs(got - want).max(axis=(1, 2)) / np.abs(want).max(axis=(1, 2))
    assert err.max() <= VERT_TOL, err


@pytest.mark.parametrize("convention,conv", [("sample_test", recon.VARIANT_B), ("matlab", recon.VARIANT_C)])
def test_other_conventions(small_model, convention, conv):
    p = fr("synth").sample_params_constrained(5, small_model["ndim_shape"], small_model["ndim_exp"], 64, seed=5, full_range=True)
    _, got = _gpu_vertices(small_model, p, 64, convention)
    want = recon.vertices_transform(p, small_model, 64, conv=conv)
    assert np.abs(got - want).max() <= VERT_TOL * np.abs(want).max()


@pytest.mark.parametrize("B", [2, 20, 64, 300])
def test_bfm_size_backward(bfm, B):
    """d params of sum(vertex_proj * g) for a full [B,3,N] upstream gradient (the TF autodiff chain, App. A.4)."""
    p = fr("synth").sample_params_constrained(B, seed=40 + B)
    rng = np.random.default_rng(B)
    g = rng.normal(size=(B, 3, 53215)).astype(np.float32)
    dm = fr("model").DeviceModel(bfm, DEV)
    pt = torch.from_numpy(p).to(DEV).requires_grad_(True)
    vp = fr("nets.network").recon_project(pt, dm, 200)
    (vp * torch.from_numpy(g).to(DEV)).sum().backward()
    got = pt.grad.cpu().numpy().astype(np.float64)
    want = recon.vertices_transform_backward(p, bfm, g)
    assert not got[:, 0:3].any()
    for sl in (slice(3, 6), slice(6, 7), slice(7, 206), slice(206, 235)):
        scale = np.abs(want[:, sl]).max(axis=1, keepdims=True)
        assert (np.abs(got[:, sl] - want[:, sl]) <= GRAD_TOL * scale).all(), sl


def test_depth_only_gradient_chain(bfm):
    """BASELINE config 4 shape of the computation: params -> depth, d depth -> d params (z row only)."""
    net = fr("nets.network")
    B = 4
    p = fr("synth").sample_params_constrained(B, seed=77)
    dm = fr("model").DeviceModel(bfm, DEV)
    pt = torch.from_numpy(p).to(DEV).requires_grad_(True)
    vp = net.recon_project(pt, dm, 200)
    image = torch.empty((B, 200, 200, 3), device=DEV)
    depth, _, _, tri_ind = fr("rendering_layer.ops").render_depth(vp, dm.tri, dm.vertex_code.unsqueeze(0).expand(B, -1, -1), image)
    gd = torch.from_numpy(np.random.default_rng(3).normal(size=(B, 200, 200, 1)).astype(np.float32)).to(DEV) * (tri_ind >= 0)
    (depth * gd).sum().backward()
    vgrad = oracle.oracle_render_depth_backward(gd.cpu().numpy(), bfm["tri"], tri_ind.cpu().numpy(), 53215)
    want = recon.vertices_transform_backward(p, bfm, vgrad)
    got = pt.grad.cpu().numpy()
    for sl in (slice(3, 6), slice(6, 7), slice(7, 206), slice(206, 235)):
        scale = np.abs(want[:, sl]).max(axis=1, keepdims=True) + 1e-30
        assert (np.abs(got[:, sl] - want[:, sl]) <= GRAD_TOL * scale).all(), sl


def test_pipeline_matches_reference_path(bfm):
    """recon (GPU) -> render (GPU) vs the reference CPU op fed the SAME float32 vertex buffer: bit-exact; and vs the
    oracle fed float64-restated vertices: tri_ind may differ only where <= 1-ulp vertex differences move an edge."""
    B = 6
    p = fr("synth").sample_params_constrained(B, seed=11)
    dm, vp = _gpu_vertices(bfm, p, 200)
    image = torch.empty((B, 200, 200, 3), device=DEV)
    out = fr("rendering_layer.ops").render_depth(torch.from_numpy(vp).to(DEV), dm.tri, dm.vertex_code.unsqueeze(0).expand(B, -1, -1), image)
    got = [o.cpu().numpy() for o in out]
    want = oracle.oracle_render_depth_forward(vp, bfm["tri"], bfm["vertex"], 200, 200)
    for g, w in zip(got, want):
        assert g.tobytes() == w.tobytes()
    vp64 = recon.vertices_transform(p, bfm, 200).astype(np.float32)
    ref = oracle.oracle_render_depth_forward(vp64, bfm["tri"], bfm["vertex"], 200, 200)
    mismatch = (ref[3] != got[3]).mean()
    assert mismatch < 2e-3, mismatch                                      # documented budget
    same = ref[3] == got[3]
    cov = same & (ref[3] >= 0)
    assert np.abs(ref[0][cov] - got[0][cov]).max() <= 1e-5 * np.abs(ref[0][cov]).max()     # norm-wise, like the vertices


def test_session_host_buffers_match_tensor_path(bfm):
    B = 5
    p = fr("synth").sample_params_constrained(B, seed=21)
    dm, vp = _gpu_vertices(bfm, p, 200)
    image = torch.empty((B, 200, 200, 3), device=DEV)
    out = fr("rendering_layer.ops").render_depth(torch.from_numpy(vp).to(DEV), dm.tri, dm.vertex_code.unsqueeze(0).expand(B, -1, -1), image)
    sess = fr("session").Session(bfm, 200, 200, max_batch=8, device=0)
    vps = np.empty((B, 3, 53215), np.float32)
    depth, tri_ind = sess.forward(p, 200.0, vertex_proj=vps)
    assert vps.tobytes() == vp.tobytes()
    assert depth.tobytes() == out[0].cpu().numpy().tobytes() and tri_ind.tobytes() == out[3].cpu().numpy().tobytes()
    gd = (np.random.default_rng(5).normal(size=depth.shape).astype(np.float32)) * (tri_ind >= 0)
    pg = sess.backward(gd)
    vgrad = oracle.oracle_render_depth_backward(gd, bfm["tri"], tri_ind, 53215)
    want = recon.vertices_transform_backward(p, bfm, vgrad)
    for sl in (slice(3, 6), slice(6, 7), slice(7, 206), slice(206, 235)):
        scale = np.abs(want[:, sl]).max(axis=1, keepdims=True) + 1e-30
        assert (np.abs(pg[:, sl] - want[:, sl]) <= GRAD_TOL * scale).all(), sl
    sess.close()


@pytest.fixture(scope="module")
def mid_model():
    """~31 000 vertices: ~300 clusters on 148 SMs -- a shared first round, ONE full round and a pool of ~80 clusters (the
    BFM-sized model has two full rounds and a pool of 140)."""
    return fr("synth").make_synthetic_model(grid=(150, 208), ndim_shape=40, ndim_exp=10, seed=5, jitter=0.2)


@pytest.mark.parametrize("B,tiles,which", [(3, False, "bfm"), (20, False, "bfm"), (3, True, "bfm"), (20, True, "bfm"), (70, True, "bfm"),
                                           (9, True, "bfm"), (33, True, "bfm"), (64, True, "bfm"), (64, True, "mid"), (41, True, "mid")])
def test_fused_call_matches_separate_calls(bfm, mid_model, B, tiles, which):
    """fr_recon_render_forward == fr_recon_project_forward + fr_render_depth_forward bit for bit, with and without the vertex
    tensor: FFMA path (B = 3) and tcgen05 path writing the rasterizer's records (default), and the FR_CLUSTER_TILES
    flavour whose tcgen05 epilogue rasterizes each cluster from shared memory: two batch tiles with the static schedule
    (B = 70), one batch tile with 2 / 3 / 5 / 6 / 8 octets per cluster -- the schedule with dynamic octets and the dynamic
    pool of half-clusters (ItemWalk) -- on two pool geometries."""
    lib, check = fr("_lib").lib(), fr("_lib").check
    bfm = bfm if which == "bfm" else mid_model
    S = 200
    p = fr("synth").sample_params_constrained(B, bfm["ndim_shape"], bfm["ndim_exp"], S, seed=60 + B)
    dm, vp = _gpu_vertices(bfm, p, 200, cluster_tiles=tiles)
    image = torch.empty((B, 200, 200, 3), device=DEV)
    want = fr("rendering_layer.ops").render_depth(torch.from_numpy(vp).to(DEV), dm.tri, dm.vertex_code.unsqueeze(0).expand(B, -1, -1), image)
    pt = torch.from_numpy(p).to(DEV)
    ws = torch.empty(lib.fr_pipeline_workspace_bytes(B, dm.nver, dm.ndim_shape, dm.ndim_exp, 200, 200), dtype=torch.uint8, device=DEV)
    sp = torch.cuda.current_stream().cuda_stream
    for with_vertex in (True, False):
        ws.fill_(0xAB)                                                    # stale workspace contents must not matter
        vertex = torch.full((B, 3, dm.nver), float("nan"), device=DEV)
        depth = torch.empty((B, 200, 200, 1), device=DEV)
        tri_ind = torch.empty((B, 200, 200, 1), device=DEV)
        check(lib.fr_recon_render_forward(pt.data_ptr(), dm.packed.data_ptr(), dm.tri.data_ptr(), dm.mesh.handle,
                                          vertex.data_ptr() if with_vertex else None, depth.data_ptr(), tri_ind.data_ptr(), B,
                                          dm.nver, dm.ntri, dm.ndim_shape, dm.ndim_exp, 200, 200, 200.0, dm.run_flags,
                                          ws.data_ptr(), ws.numel(), sp, None))
        torch.cuda.synchronize()
        assert depth.cpu().numpy().tobytes() == want[0].cpu().numpy().tobytes(), with_vertex
        assert tri_ind.cpu().numpy().tobytes() == want[3].cpu().numpy().tobytes(), with_vertex
        if with_vertex:
            assert vertex.cpu().numpy().tobytes() == vp.tobytes()
        else:
            assert bool(torch.isnan(vertex).all())                         # untouched


@pytest.mark.parametrize("B,tiles,which", [(5, True, "bfm"), (20, True, "bfm"), (70, True, "bfm"), (20, False, "bfm"), (41, True, "mid")])
def test_recon_render_depth_all_outputs(bfm, mid_model, B, tiles, which):
    """recon_render_depth (fr_recon_render_forward_all: the training path's vertices + all four render_depth outputs from one
    call -- FFMA path, records pipeline, rasterizing epilogue that also leaves the vertex records) against recon_project +
    render_depth: five outputs bit for bit, parameter gradients of a loss on depth and vertices equal to 1e-5."""
    net, ops = fr("nets.network"), fr("rendering_layer.ops")
    m = bfm if which == "bfm" else mid_model
    S = 200
    p = fr("synth").sample_params_constrained(B, m["ndim_shape"], m["ndim_exp"], S, seed=80 + B)
    dm = fr("model").DeviceModel(m, DEV, cluster_tiles=tiles)
    rng = np.random.default_rng(B)
    gd = torch.from_numpy(rng.normal(size=(B, S, S, 1)).astype(np.float32)).to(DEV)
    gv = torch.from_numpy((1e-3 * rng.normal(size=(B, 3, dm.nver))).astype(np.float32)).to(DEV)
    tex = dm.vertex_code                                                   # [3,N], shared by all faces
    # separate ops
    p1 = torch.from_numpy(p).to(DEV).requires_grad_(True)
    v1 = net.recon_project(p1, dm, S)
    image = torch.empty((B, S, S, 3), device=DEV)
    d1, t1, n1, i1 = ops.render_depth(v1, dm.tri, tex.unsqueeze(0).expand(B, -1, -1), image)
    ((d1 * gd).sum() + (v1 * gv).sum()).backward()
    # one call
    p2 = torch.from_numpy(p).to(DEV).requires_grad_(True)
    v2, d2, t2, n2, i2 = net.recon_render_depth(p2, dm, tex, S, S, S)
    ((d2 * gd).sum() + (v2 * gv).sum()).backward()
    torch.cuda.synchronize()
    for a, b, name in ((v1, v2, "vertex_proj"), (d1, d2, "depth"), (t1, t2, "texture_image"), (n1, n2, "normal"), (i1, i2, "tri_ind")):
        assert a.detach().cpu().numpy().tobytes() == b.detach().cpu().numpy().tobytes(), name
    g1, g2 = p1.grad.cpu().numpy().astype(np.float64), p2.grad.cpu().numpy().astype(np.float64)
    assert int((i2 >= 0).sum()) > 1000 and np.abs(g1).max() > 0
    scale = np.abs(g1).max(axis=1, keepdims=True)
    assert (np.abs(g1 - g2) <= 1e-5 * scale).all()
    # per-face texture through the same call
    texb = (tex.unsqueeze(0) * torch.linspace(0.5, 1.0, B, device=DEV).view(B, 1, 1)).contiguous()
    _, _, t3, _, _ = net.recon_render_depth(torch.from_numpy(p).to(DEV), dm, texb, S, S, S)
    _, t4, _, _ = ops.render_depth(v1.detach(), dm.tri, texb, image)
    assert t3.cpu().numpy().tobytes() == t4.cpu().numpy().tobytes()


@pytest.mark.parametrize("tiles", [False, True])
@pytest.mark.parametrize("B,H,W", [(70, 33, 31), (9, 48, 64), (130, 20, 20), (1500, 16, 16)])
def test_fused_call_small_model_odd_shapes(small_model, B, H, W, tiles):
    """The fused call on a small model (K = 18: one 16-k chunk pair, one M tile), several 64-face batch tiles, non-square images
    and an odd pixel count (keys cleared by memset instead of the reconstruction epilogue): bit-identical to the two calls.
    1500 faces = 24 batch tiles: the reconstruction kernel goes out in several launches that each fill the GPU."""
    lib, check = fr("_lib").lib(), fr("_lib").check
    ks, ke = small_model["ndim_shape"], small_model["ndim_exp"]
    p = fr("synth").sample_params_constrained(B, ks, ke, max(H, W), seed=7 + B)
    dm = fr("model").DeviceModel(small_model, DEV, cluster_tiles=tiles)
    pt = torch.from_numpy(p).to(DEV)
    sp = torch.cuda.current_stream().cuda_stream
    ws = torch.empty(lib.fr_pipeline_workspace_bytes(B, dm.nver, ks, ke, H, W), dtype=torch.uint8, device=DEV)
    rb = lib.fr_recon_workspace_bytes(B, dm.nver, ks, ke)
    vp = torch.empty((B, 3, dm.nver), device=DEV)
    want_d, want_t = torch.empty((B, H, W, 1), device=DEV), torch.empty((B, H, W, 1), device=DEV)
    check(lib.fr_recon_project_forward(pt.data_ptr(), dm.packed.data_ptr(), dm.mesh.handle, vp.data_ptr(), B, dm.nver, ks, ke,
                                       float(max(H, W)), dm.run_flags, ws.data_ptr(), rb, sp))
    ws2 = torch.empty(lib.fr_render_workspace_bytes(B, dm.nver, H, W), dtype=torch.uint8, device=DEV)
    check(lib.fr_render_depth_forward(vp.data_ptr(), dm.tri.data_ptr(), None, 0, want_d.data_ptr(), None, None, want_t.data_ptr(), B,
                                      dm.nver, dm.ntri, H, W, None, ws2.data_ptr(), ws2.numel(), sp))      # generic rasterizer
    ws.fill_(0x5C)
    got_d, got_t = torch.empty_like(want_d), torch.empty_like(want_t)
    check(lib.fr_recon_render_forward(pt.data_ptr(), dm.packed.data_ptr(), dm.tri.data_ptr(), dm.mesh.handle, None, got_d.data_ptr(),
                                      got_t.data_ptr(), B, dm.nver, dm.ntri, ks, ke, H, W, float(max(H, W)), dm.run_flags, ws.data_ptr(),
                                      ws.numel(), sp, None))
    torch.cuda.synchronize()
    assert (want_t >= 0).sum() > 0.05 * want_t.numel()                    # the faces are in frame
    assert got_d.cpu().numpy().tobytes() == want_d.cpu().numpy().tobytes()
    assert got_t.cpu().numpy().tobytes() == want_t.cpu().numpy().tobytes()
    want_vp = recon.vertices_transform(p, small_model, max(H, W))
    assert np.abs(vp.cpu().numpy() - want_vp).max() <= VERT_TOL * np.abs(want_vp).max()


@pytest.mark.parametrize("B", [4, 20, 70, 260])
def test_small_model_backward_all_paths(small_model, B):
    """Reconstruction backward on the small model: FFMA kernel (B = 4) and the tcgen05 contraction with one M tile, one and
    several 64-face operand groups and two 256-face batch tiles (B = 260), against the float64 restatement."""
    ks, ke = small_model["ndim_shape"], small_model["ndim_exp"]
    p = fr("synth").sample_params_constrained(B, ks, ke, 64, seed=90 + B, full_range=True)
    nver = small_model["mu"].size // 3
    g = np.random.default_rng(B).normal(size=(B, 3, nver)).astype(np.float32)
    g[B // 2] *= 1e-3                                                      # per-face operand scales must not leak between faces
    dm = fr("model").DeviceModel(small_model, DEV)
    pt = torch.from_numpy(p).to(DEV).requires_grad_(True)
    vp = fr("nets.network").recon_project(pt, dm, 64)
    (vp * torch.from_numpy(g).to(DEV)).sum().backward()
    got = pt.grad.cpu().numpy().astype(np.float64)
    want = recon.vertices_transform_backward(p, small_model, g)
    assert not got[:, 0:3].any()
    for sl in (slice(3, 6), slice(6, 7), slice(7, 7 + ks), slice(7 + ks, 7 + ks + ke)):
        scale = np.abs(want[:, sl]).max(axis=1, keepdims=True)
        assert (np.abs(got[:, sl] - want[:, sl]) <= GRAD_TOL * scale).all(), sl


@pytest.mark.parametrize("tiles", [False, True])
def test_fused_call_is_cuda_graph_capturable(small_model, tiles):
    """include/facerecon_b200.h promises stream-ordered, capturable calls (the reference launches on the legacy default
    stream and mallocs per call): capture the fused call -- prep kernel, programmatic dependent launches, rasterizer -- in a
    CUDA graph, replay it on new parameters, compare with the eager call."""
    lib, check = fr("_lib").lib(), fr("_lib").check
    ks, ke = small_model["ndim_shape"], small_model["ndim_exp"]
    B, S = 20, 48
    dm = fr("model").DeviceModel(small_model, DEV, cluster_tiles=tiles)
    pa = torch.from_numpy(fr("synth").sample_params_constrained(B, ks, ke, S, seed=1)).to(DEV)
    pb = torch.from_numpy(fr("synth").sample_params_constrained(B, ks, ke, S, seed=2)).to(DEV)
    params = pa.clone()
    ws = torch.empty(lib.fr_pipeline_workspace_bytes(B, dm.nver, ks, ke, S, S), dtype=torch.uint8, device=DEV)
    depth, tri_ind = torch.empty((B, S, S, 1), device=DEV), torch.empty((B, S, S, 1), device=DEV)

    def call(stream):
        check(lib.fr_recon_render_forward(params.data_ptr(), dm.packed.data_ptr(), dm.tri.data_ptr(), dm.mesh.handle, None,
                                          depth.data_ptr(), tri_ind.data_ptr(), B, dm.nver, dm.ntri, ks, ke, S, S, float(S), dm.run_flags,
                                          ws.data_ptr(), ws.numel(), stream.cuda_stream, None))
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        call(side)                                                        # warm-up outside the capture (function attributes)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        call(torch.cuda.current_stream())
    for src in (pb, pa, pb):
        params.copy_(src)
        graph.replay()
        torch.cuda.synchronize()
        got = (depth.clone(), tri_ind.clone())
        call(torch.cuda.current_stream())
        torch.cuda.synchronize()
        assert got[0].cpu().numpy().tobytes() == depth.cpu().numpy().tobytes()
        assert got[1].cpu().numpy().tobytes() == tri_ind.cpu().numpy().tobytes()
    assert (tri_ind >= 0).any()


def test_session_pipelined_slots_match_synchronous(bfm):
    """fr_session_submit / fr_session_wait: batches alternating over the two slots (different sizes, outputs in pinned
    memory) give bit-identical results to the synchronous call, and a busy slot refuses a second submit."""
    sess = fr("session").Session(bfm, 200, 200, max_batch=6, device=0)
    batches = [fr("synth").sample_params_constrained(b, seed=40 + i) for i, b in enumerate((6, 3, 5, 1))]
    want = [sess.forward(p, 200.0)[0].copy() for p in batches]
    pins = [torch.empty((p.shape[0], 200, 200, 1), dtype=torch.float32).pin_memory() for p in batches]
    pps = [torch.from_numpy(p.copy()).pin_memory() for p in batches]
    for i, p in enumerate(pps):
        slot = i % 2
        if i >= 2:
            sess.wait(slot)
        sess.submit(slot, p.numpy(), 200.0, depth=pins[i].numpy())
    with pytest.raises(ValueError):
        sess.submit(0, pps[0].numpy(), 200.0, depth=pins[0].numpy())      # slot 0 is still in flight
    sess.wait(0)
    sess.wait(1)
    sess.wait(1)                                                          # waiting on an idle slot is a no-op
    for i in range(len(batches)):
        assert pins[i].numpy().tobytes() == want[i].tobytes(), i
    sess.close()


@pytest.mark.parametrize("B", [5, 24])
def test_set_constraints_fused_into_prep(bfm, B):
    """FR_PARAMS_RAW -- set_constraints inside the prep kernels (FFMA path at B = 5, tcgen05 path at B = 24) ==
    the torch transcription of network.py:204-218 followed by vertices_transform; same for the gradient w.r.t. the raw values."""
    net = fr("nets.network").FaceRecNet(mesh_data=bfm, batch_size=B, im_size=200, device=DEV)
    rng = np.random.default_rng(B)
    raw_np = rng.normal(scale=1.5, size=(B, 235)).astype(np.float32)
    raw_np[:, 0:3] *= 0.2                                                   # keep the faces roughly frontal
    g = torch.from_numpy(rng.normal(size=(B, 3, 53215)).astype(np.float32)).to(DEV)
    outs, grads = [], []
    for fused in (True, False):
        raw = torch.from_numpy(raw_np).to(DEV).requires_grad_(True)
        if fused:
            vp = net.vertices_transform_raw(raw[:, None, None, :])
        else:
            vp = net.vertices_transform(net.set_constraints(raw[:, None, None, :]))
        (vp * g).sum().backward()
        outs.append(vp.detach().cpu().numpy())
        grads.append(raw.grad.cpu().numpy().astype(np.float64))
    assert np.abs(outs[0] - outs[1]).max() <= 2e-6 * np.abs(outs[1]).max()
    for sl in (slice(3, 5), slice(6, 7), slice(7, 206), slice(206, 235)):
        scale = np.abs(grads[1][:, sl]).max(axis=1, keepdims=True) + 1e-30
        assert (np.abs(grads[0][:, sl] - grads[1][:, sl]) <= GRAD_TOL * scale).all(), sl
    assert not grads[0][:, 0:3].any() and not grads[0][:, 5].any()


def test_geometry_loss_gram_form(bfm):
    """the geometry loss through the 228 x 228 Gram matrix == the literal mean squared difference of the two
    basis contractions (network.py:346-355), value and gradient w.r.t. the predicted coefficients."""
    B = 6
    net = fr("nets.network").FaceRecNet(mesh_data=bfm, batch_size=B, im_size=200, device=DEV)
    rng = np.random.default_rng(3)
    pred = fr("synth").sample_params_constrained(B, seed=70)
    label = fr("synth").sample_params_constrained(B, seed=71)
    pt = torch.from_numpy(pred).to(DEV).requires_grad_(True)
    loss = net.geometry_loss(pt[:, None, None, :], torch.from_numpy(label).to(DEV)[:, None, None, :])
    loss.backward()
    basis = np.concatenate([bfm["pc_shape"], bfm["pc_exp"]], axis=1).astype(np.float64)
    d = (label[:, 7:] - pred[:, 7:]).astype(np.float64)
    diff = basis @ d.T                                                    # [3N, B]
    want = (diff ** 2).mean()
    want_grad = -2.0 * (basis.T @ diff).T / diff.size
    assert abs(float(loss.detach()) - want) <= 1e-6 * want
    got_grad = pt.grad.cpu().numpy()
    assert not got_grad[:, :7].any()
    assert np.abs(got_grad[:, 7:] - want_grad).max() <= 1e-5 * np.abs(want_grad).max()


def test_set_constraints_pinned_to_reference_source():
    """SURVEY 8f-2 pinned: tests/golden/constraints.npz holds the output of the reference's own FaceRecNet.set_constraints
    source (nets/network.py:204-218, executed by make_golden.py).  The torch mirror must reproduce it, and the prep kernels'
    fused constraints (FR_PARAMS_RAW) must give the vertices of the reference-constrained parameters."""
    g = np.load(os.path.join(GOLDEN, "constraints.npz"))
    raw, want = g["raw"], g["constrained"]
    model = fr("synth").make_synthetic_model(grid=(9, 11), seed=14, jitter=0.2)           # true K = 199 + 29
    net = fr("nets.network").FaceRecNet(mesh_data=model, batch_size=raw.shape[0], im_size=int(g["im_size"]), device=DEV)
    got = net.set_constraints(torch.from_numpy(raw).to(DEV)).cpu().numpy()
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max()
    assert (got[..., 5] == 0).all() and (want[..., 5] == 0).all()                          # network.py:213
    vp_ref = net.vertices_transform(torch.from_numpy(want).to(DEV)).cpu().numpy()          # vertices of the reference-constrained params
    vp_fused = net.vertices_transform_raw(torch.from_numpy(raw).to(DEV)).cpu().numpy()     # constraints inside the prep kernel
    assert np.abs(vp_fused - vp_ref).max() <= 2e-6 * np.abs(vp_ref).max()


def test_geometry_loss_pinned_to_reference_source():
    """SURVEY 8f-3 pinned: tests/golden/geometry_loss.npz holds the value of the geometry-loss statements of the reference's
    FaceRecNet.get_loss (nets/network.py:346-355, executed verbatim by make_golden.py)."""
    g = np.load(os.path.join(GOLDEN, "geometry_loss.npz"))
    model = _model_from_golden(g)
    net = fr("nets.network").FaceRecNet(mesh_data=model, batch_size=g["pred"].shape[0], im_size=200, device=DEV)
    loss = net.geometry_loss(torch.from_numpy(g["pred"]).to(DEV)[:, None, None, :], torch.from_numpy(g["label"]).to(DEV)[:, None, None, :])
    want = float(g["loss"])
    assert abs(float(loss) - want) <= 1e-5 * want, (float(loss), want)


def test_facerecnet_mirror(bfm):
    """The FaceRecNet geometry slice end to end: default pred_params -> depth_rendering_layer (network.py:300-308)."""
    net = fr("nets.network").FaceRecNet(mesh_data=bfm, batch_size=2, im_size=200, device=DEV)
    depth = net.depth_rendering_layer()
    assert depth.shape == (2, 200, 200, 1) and float(depth.min()) >= 1e-6 * 0.999
    vp = net.vertices_proj.cpu().numpy()
    want = recon.vertices_transform(net.pred_params[:, 0, 0, :].cpu().numpy(), bfm, 200)
    assert np.abs(vp - want).max() <= VERT_TOL * np.abs(want).max()
    assert net.pncc_batch.shape == (2, 200, 200, 3) and net.normal_batch.shape == (2, 200, 200, 3)
    nrm = (net.normal_batch ** 2).sum(-1)
    cov = net.coarse_depth_map[..., 0] > 1e-6
    assert torch.allclose(nrm[cov], torch.ones_like(nrm[cov]), atol=1e-3)


def test_device_model_tri_base_and_mesh_cache(small_model, tmp_path):
    """The drop-in flow read_3dmm_model(path) -> FaceRecNet(mesh_data=...) with the BFM files' 1-based ``tri``: detected and
    shifted with a warning (tri_base=None), stated explicitly (tri_base=1), or rejected when out of range; and the mesh
    table's on-disk cache (cache_dir) gives the same table and the same results as a fresh build."""
    model_mod = fr("model")
    one_based = dict(small_model)
    one_based["tri"] = small_model["tri"] + 1.0
    assert one_based["tri"].max() == small_model["mu"].size // 3            # the last vertex is referenced
    with pytest.warns(UserWarning, match="1-based"):
        dm_auto = model_mod.DeviceModel(one_based, DEV)
    dm_explicit = model_mod.DeviceModel(one_based, DEV, tri_base=1)
    dm_zero = model_mod.DeviceModel(small_model, DEV, cache_dir=str(tmp_path))
    assert dm_auto.tri_base == 1 and torch.equal(dm_auto.tri, dm_zero.tri) and torch.equal(dm_explicit.tri, dm_zero.tri)
    with pytest.raises(ValueError, match="tri_base"):
        model_mod.DeviceModel(one_based, DEV, tri_base=0)                   # index nver is out of range when taken as 0-based
    files = list(tmp_path.iterdir())
    assert len(files) == 1 and files[0].name.startswith("mesh_")
    dm_cached = model_mod.DeviceModel(small_model, DEV, cache_dir=str(tmp_path))
    assert dm_cached.mesh.blob().tobytes() == dm_zero.mesh.blob().tobytes()
    files[0].write_bytes(files[0].read_bytes()[:-7])                        # a damaged cache entry is rebuilt, not trusted
    dm_rebuilt = model_mod.DeviceModel(small_model, DEV, cache_dir=str(tmp_path))
    assert dm_rebuilt.mesh.blob().tobytes() == dm_zero.mesh.blob().tobytes()
    ks, ke = small_model["ndim_shape"], small_model["ndim_exp"]
    p = torch.from_numpy(fr("synth").sample_params_constrained(12, ks, ke, 64, seed=5)).to(DEV)
    want = fr("nets.network").recon_project(p, dm_zero, 64)
    for dm in (dm_auto, dm_cached, dm_rebuilt):
        assert torch.equal(fr("nets.network").recon_project(p, dm, 64), want)


@pytest.mark.parametrize("tiles", [False, True])
def test_fused_call_on_adversarial_geometry(tiles):
    """The fused params -> depth-map call on geometry that lands EXACTLY on pixel centres and edges (integer mean grid, f = 1 and
    f = 0.5, integer translations: the certified fast inside test must hand every such pixel to the literal PointInTri pass),
    partly off-screen faces, and NaN / inf / huge mean entries (culled like the reference culls them) -- through both fused
    flavours (records pipeline / tile rasterizer inside the reconstruction epilogue), against the reference semantics
    evaluated on the vertices the planar reconstruction call returns for the same parameters."""
    lib, check = fr("_lib").lib(), fr("_lib").check
    synth = fr("synth")
    m = synth.make_synthetic_model(grid=(23, 31), ndim_shape=2, ndim_exp=1, seed=3, jitter=0.0)
    n = 23 * 31
    col, row = np.meshgrid(np.arange(23, dtype=np.float32), np.arange(31, dtype=np.float32))
    mu = np.stack([col.ravel() + 2.0, row.ravel() + 1.0, ((col + 2 * row) % 5).ravel()]).astype(np.float32)      # integer grid
    mu[0, 40], mu[1, 77], mu[2, 200] = np.nan, np.inf, np.nan
    mu[0, 300], mu[1, 333] = 3.0e30, -3.0e30
    m["mu"] = mu.reshape(3 * n, 1)
    m["pc_shape"] = np.zeros_like(m["pc_shape"])
    m["pc_exp"] = np.zeros_like(m["pc_exp"])
    B, S = 12, 40
    p = np.zeros((B, 7 + 2 + 1), np.float32)
    p[:, 6] = 1.0
    p[6:, 6] = 0.5                                                        # half-integer coordinates
    p[:, 3] = np.arange(B) % 6 - 2                                        # integer shifts, some faces partly off-screen
    p[:, 4] = (np.arange(B) * 3) % 7
    p[10, 3] = 1000.0                                                     # one face entirely off-screen
    dm = fr("model").DeviceModel(m, DEV, cluster_tiles=tiles)
    pt = torch.from_numpy(p).to(DEV)
    sp = torch.cuda.current_stream().cuda_stream
    ws = torch.empty(lib.fr_pipeline_workspace_bytes(B, dm.nver, 2, 1, S, S), dtype=torch.uint8, device=DEV)
    rb = lib.fr_recon_workspace_bytes(B, dm.nver, 2, 1)
    vp = torch.empty((B, 3, dm.nver), device=DEV)
    check(lib.fr_recon_project_forward(pt.data_ptr(), dm.packed.data_ptr(), dm.mesh.handle, vp.data_ptr(), B, dm.nver, 2, 1, float(S),
                                       dm.run_flags, ws.data_ptr(), rb, sp))
    torch.cuda.synchronize()
    v = vp.cpu().numpy()
    ok = np.isfinite(mu).all(axis=0) & (np.abs(mu) < 1e29).all(axis=0)
    assert (v[0, 0, ok] == mu[0, ok] + p[0, 3]).all() and (v[0, 1, ok] == S - (mu[1, ok] + p[0, 4]) - 1).all()   # exact integers
    assert np.isnan(v[:, 0, 40]).all() and not np.isfinite(v[:, 1, 77]).any()
    want = oracle.oracle_render_depth_forward(v, m["tri"], np.zeros_like(v), S, S)
    assert (want[3][0] >= 0).sum() > 300 and (want[3][10] < 0).all()
    d, t = torch.empty((B, S, S, 1), device=DEV), torch.empty((B, S, S, 1), device=DEV)
    ws.fill_(0xA5)
    check(lib.fr_recon_render_forward(pt.data_ptr(), dm.packed.data_ptr(), dm.tri.data_ptr(), dm.mesh.handle, None, d.data_ptr(),
                                      t.data_ptr(), B, dm.nver, dm.ntri, 2, 1, S, S, float(S), dm.run_flags, ws.data_ptr(), ws.numel(),
                                      sp, None))
    torch.cuda.synchronize()
    assert t.cpu().numpy().tobytes() == want[3].tobytes()
    assert d.cpu().numpy().tobytes() == want[0].tobytes()
