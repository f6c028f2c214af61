"""CPU: the arithmetic the CUDA rasterizer runs (3dfacerecon_b200/csrc/raster_core.h: bbox/cull, FP64 inside test,
(depth, ~index) key packing, resolve) re-enacted on the host in a scrambled triangle order and compared bit for bit
with the oracle and the golden vectors.  This checks the kernel LOGIC without a GPU; the kernels themselves are
checked on the B200 by tests/test_gpu_*.py."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle
from conftest import ROOT, fr

SRC = os.path.join(ROOT, "tests", "host_emul", "raster_emul.cc")
SO = os.path.join(ROOT, "tests", "host_emul", "libraster_emul.so")
_f32p = ctypes.POINTER(ctypes.c_float)


@pytest.fixture(scope="module")
def emul():
    hdrs = [os.path.join(ROOT, "3dfacerecon_b200", "csrc", h) for h in ("raster_core.h", "mesh_table.h")]
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(f) for f in [SRC] + hdrs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", SO, SRC], check=True)
    lib = ctypes.CDLL(SO)
    lib.fr_emul_render_forward.restype = ctypes.c_int
    lib.fr_emul_render_forward.argtypes = [_f32p, _f32p, _f32p, ctypes.c_longlong] + [ctypes.c_int] * 5 + [ctypes.c_uint] + [_f32p] * 4

    def run(vertex, tri, texture, H, W, seed=12345):
        vertex, tri, texture = [np.ascontiguousarray(a, np.float32) for a in (vertex, tri, texture)]
        B, _, N = vertex.shape
        outs = [np.empty((B, H, W, c), np.float32) for c in (1, 3, 3, 1)]
        p = lambda a: a.ctypes.data_as(_f32p)
        rc = lib.fr_emul_render_forward(p(vertex), p(tri), p(texture), 3 * N, B, N, tri.shape[1], H, W, seed,
                                        p(outs[0]), p(outs[1]), p(outs[2]), p(outs[3]))
        assert rc == 0, "emulation self-check %d failed" % rc
        return outs
    run.lib = lib
    return run


def test_emulation_matches_golden(emul, render_golden):
    for name, c in render_golden.items():
        B, H, W, _ = [int(x) for x in c["image_shape"]]
        for seed in (0, 7):
            got = emul(c["vertex"], c["tri"], c["texture"], H, W, seed)
            for g, key in zip(got, ("depth", "texture_image", "normal", "tri_ind")):
                assert g.tobytes() == c[key].tobytes(), (name, key, seed)


def test_emulation_matches_oracle_on_mesh(emul):
    synth = fr("synth")
    from oracle import recon
    m = synth.make_synthetic_model(grid=(61, 75), ndim_shape=8, ndim_exp=4, seed=8, jitter=0.0)
    p = synth.sample_params_constrained(2, 8, 4, 120, seed=3)
    p[:, 6] *= 0.6
    vp = recon.vertices_transform(p, m, 120, dtype=np.float32).astype(np.float32)
    tex = np.broadcast_to(m["mu_tex"], (2,) + m["mu_tex"].shape).copy()
    want = oracle.oracle_render_depth_forward(vp, m["tri"], tex, 120, 120)
    got = emul(vp, m["tri"], tex, 120, 120)
    for g, w in zip(got, want):
        assert g.tobytes() == w.tobytes()
    assert (want[3] >= 0).sum() > 2000


def test_key_order_properties():
    """fr_float_order_bits is monotone, folds -0.0 onto +0.0, and every drawable depth has a non-zero key."""
    def order_bits(h):
        u = np.float32(h).view(np.uint32).item()
        if (u & 0x7FFFFFFF) == 0:
            u = 0
        return (~u & 0xFFFFFFFF) if (u & 0x80000000) else (u | 0x80000000)
    vals = np.array([-9.9e13, -1e5, -1.0, -1e-30, -0.0, 0.0, 1e-30, 1.0, 3e38, np.inf], np.float32)
    bits = [order_bits(v) for v in vals]
    assert all(a <= b for a, b in zip(bits, bits[1:]))
    assert order_bits(-0.0) == order_bits(0.0)
    assert all(b > 0 for b in bits)
    assert order_bits(oracle.BACKGROUND_DEPTH) < order_bits(np.float32(-9.9e13))


def test_emulation_extreme_coordinates(emul):
    """NaN / inf / beyond-int-range xy: the float fast path of the bbox must agree with the literal double path
    (self-checked inside the emulation) and with the oracle."""
    v = np.array([[[0, 4, 0, 1e20, -1e20, np.nan, 3e9, 2, 1, np.inf, -np.inf, 7.5, -0.5],
                   [0, 0, 4, 1, 2, 3, 1, -3e9, np.inf, 1, 2, 7.0, -0.0],
                   [1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1]]], np.float32)
    rng = np.random.default_rng(3)
    t = rng.integers(0, v.shape[2], (3, 400)).astype(np.float32)
    tex = np.zeros_like(v)
    want = oracle.oracle_render_depth_forward(v, t, tex, 8, 8)
    for seed in (0, 5):
        got = emul(v, t, tex, 8, 8, seed)
        for g, w in zip(got, want):
            assert g.tobytes() == w.tobytes()


def test_snap_code_short_form_matches_definition(emul):
    """The few-instruction snap code the kernels compute == its literal definition (per-vertex biased ceil/floor, clamped),
    over every 509-th float bit pattern (NaNs, infinities, denormals, huge values included) and around every integer."""
    lib = emul.lib
    lib.fr_emul_check_snap_code.restype = ctypes.c_longlong
    lib.fr_emul_check_snap_code.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_uint]
    for w, h in ((200, 200), (1, 7), (31999, 640)):
        assert lib.fr_emul_check_snap_code(w, h, 509) == 0, (w, h)


def test_clustered_emulation_matches_oracle(emul, render_golden):
    """The data flow of the tile rasterizer (mesh table -> staged vertices -> raw min/max cull on local slots -> certified fast inside
    test with literal fallback -> packed-key maximum ->
    depth / index decoded from the key alone), re-enacted on the host with the real table builder, equals the oracle bit for
    bit: golden cases (degenerate / duplicate / off-screen / NaN-depth triangles, -0.0 depths), a grid mesh and a soup."""
    lib = emul.lib
    u8p = ctypes.POINTER(ctypes.c_ubyte)
    lib.fr_emul_render_forward_clustered.restype = ctypes.c_int
    lib.fr_emul_render_forward_clustered.argtypes = [_f32p, u8p] + [ctypes.c_int] * 4 + [_f32p, _f32p]
    mesh = fr("mesh")

    def run(vertex, tri, H, W, positions):
        vertex = np.ascontiguousarray(vertex, np.float32)
        B, _, N = vertex.shape
        blob = mesh.MeshTable(tri, N, positions).blob()
        d, t = np.empty((B, H, W, 1), np.float32), np.empty((B, H, W, 1), np.float32)
        rc = lib.fr_emul_render_forward_clustered(vertex.ctypes.data_as(_f32p), blob.ctypes.data_as(u8p), B, N, H, W,
                                                  d.ctypes.data_as(_f32p), t.ctypes.data_as(_f32p))
        assert rc == 0
        return d, t

    for name, c in render_golden.items():
        B, H, W, _ = [int(x) for x in c["image_shape"]]
        for pos in (None, np.nan_to_num(c["vertex"][0], nan=0.0, posinf=0.0, neginf=0.0)):
            d, t = run(c["vertex"], c["tri"], H, W, pos)
            assert d.tobytes() == c["depth"].tobytes() and t.tobytes() == c["tri_ind"].tobytes(), name
    synth = fr("synth")
    from oracle import recon
    m = synth.make_synthetic_model(grid=(61, 75), ndim_shape=8, ndim_exp=4, seed=8, jitter=0.0, permute=True)
    p = synth.sample_params_constrained(2, 8, 4, 120, seed=3)
    p[:, 6] *= 0.6
    vp = recon.vertices_transform(p, m, 120, dtype=np.float32).astype(np.float32)
    want = oracle.oracle_render_depth_forward(vp, m["tri"], np.zeros_like(vp), 120, 120)
    d, t = run(vp, m["tri"], 120, 120, m["mu"].reshape(3, -1))
    assert d.tobytes() == want[0].tobytes() and t.tobytes() == want[3].tobytes()
    # signed zeros: two coplanar triangles at depth -0.0 and +0.0 tie (lowest index wins) and keep their own sign
    v = np.array([[[0, 4, 0, 0, 4, 0], [0, 0, 4, 0, 0, 4], [-0.0, -0.0, -0.0, 0.0, 0.0, 0.0]]], np.float32)
    for tri in (np.array([[0, 3], [1, 4], [2, 5]], np.float32), np.array([[3, 0], [4, 1], [5, 2]], np.float32)):
        want = oracle.oracle_render_depth_forward(v, tri, np.zeros_like(v), 8, 8)
        d, t = run(v, tri, 8, 8, None)
        assert d.tobytes() == want[0].tobytes() and t.tobytes() == want[3].tobytes()
        assert (np.signbit(want[0][want[3] >= 0]) == (tri[0, 0] == 0)).all()


def test_tile_cull_and_certified_fast_inside_test(emul):
    """The tile rasterizer's arithmetic (raster_core.h, used by raster_tile.cuh): the cull on the raw packed min / max of the
    snap codes equals fr_code_keep / fr_code_box, and the float fast inside test never contradicts the literal PointInTri
    -- over random and adversarial triangles (integer vertices, pixel centres on edges, slivers, degenerate triangles, large
    boxes).  It must also decide nearly every test of generic sub-pixel triangles, or it would not be a fast path."""
    lib = emul.lib
    lib.fr_emul_check_tile_arith.restype = ctypes.c_longlong
    lib.fr_emul_check_tile_arith.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint, ctypes.POINTER(ctypes.c_longlong)]
    for (w, h, n, seed) in ((200, 200, 2_000_000, 1), (7, 5, 300_000, 2), (16000, 9000, 300_000, 3)):
        total = ctypes.c_longlong(0)
        decided = lib.fr_emul_check_tile_arith(w, h, n, seed, ctypes.byref(total))
        assert decided >= 0, (w, h, decided)
        assert total.value > n // 10 and decided > 0.5 * total.value, (w, h, decided, total.value)
