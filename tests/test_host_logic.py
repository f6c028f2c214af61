"""CPU: host-side logic -- C-ABI surface, sharding, the multi-process plumbing (gloo, world_size 2), synthetic model."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, fr


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "facerecon_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fr_[a-z_0-9]+)\s*\(", text)))


def test_abi_library_exports_every_declared_symbol():
    lib_mod = fr("_lib")
    names = _declared_symbols()
    assert len(names) >= 17 and "fr_render_depth_forward" in names and "fr_session_forward" in names
    assert sorted(lib_mod.SIGNATURES) == names                      # the ctypes table covers the header exactly
    handle = ctypes.CDLL(lib_mod.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), n
    lib = lib_mod.lib()
    assert lib.fr_version() == 202
    # size queries are pure host arithmetic: safe without a GPU
    assert lib.fr_packed_basis_bytes(53215, 199, 29, 0, None) == 416 * 3 * 58 * 128 * 16 + 416 * 3 * 15 * 8192 + 1024 + 416 * 3 * 8 * 16384 + 3 * 416 * 128 * 4   # fp32 + fp16-pair tiles + column scales + backward tiles + fp32 mean
    assert lib.fr_render_workspace_bytes(64, 53215, 200, 200) == 64 * 200 * 200 * 8 + 64 * 53215 * 16 + (53215 * 16 + 255) // 256 * 256 + 256   # keys + 16-byte vertex records + a shared texture repacked by rank + schedule counters
    assert lib.fr_recon_workspace_bytes(64, 53215, 199, 29) >= 232 * 64 * 4
    assert lib.fr_packed_basis_bytes(0, 199, 29, 0, None) == 0


def test_abi_rejects_bad_arguments_without_a_gpu():
    """Validation happens before any CUDA call, so it is checkable here; messages follow the reference's."""
    lib_mod = fr("_lib")
    lib = lib_mod.lib()
    rc = lib.fr_render_depth_forward(None, None, None, 0, None, None, None, None, 1, 10, 10 * 1000 * 1000, 8, 8, None, None, 0, None)
    assert rc == lib_mod.FR_ERR_INVALID_ARGUMENT and "Too many triangular" in lib_mod.last_error()
    rc = lib.fr_render_depth_forward(None, None, None, 0, None, None, None, None, 1, 10, 5, 8, 8, None, None, 0, None)
    assert rc == lib_mod.FR_ERR_INVALID_ARGUMENT and "null pointer" in lib_mod.last_error()
    rc = lib.fr_render_depth_forward(None, None, None, 0, None, None, None, None, 70000, 10, 5, 8, 8, None, None, 0, None)
    assert rc == lib_mod.FR_ERR_INVALID_ARGUMENT and "65535" in lib_mod.last_error()
    rc = lib.fr_recon_project_forward(None, None, None, None, 1, 10, 0, 0, 200.0, 0, None, 0, None)
    assert rc == lib_mod.FR_ERR_INVALID_ARGUMENT
    with pytest.raises(ValueError):
        lib_mod.check(rc)
    assert lib.fr_render_depth_forward(None, None, None, 0, None, None, None, None, 0, 10, 5, 8, 8, None, None, 0, None) == 0   # empty batch
    # the all-outputs call: needs every output pointer, validates the model / image dimensions first
    args_all = [None] * 5 + [0] + [None] * 5
    rc = lib.fr_recon_render_forward_all(*args_all, 4, 10, 5, 3, 2, 8, 8, 8.0, 0, None, 0, None)
    assert rc == lib_mod.FR_ERR_INVALID_ARGUMENT and "null pointer" in lib_mod.last_error()


def test_one_call_op_has_no_cpu_path():
    """recon_render_depth refuses CPU tensors before it touches the library (no silent fallback)."""
    import torch
    net = fr("nets.network")

    class FakeModel:
        run_flags, ndim, ndim_shape, ndim_exp, nver = 0, 235, 199, 29, 10
    with pytest.raises(RuntimeError, match="no CPU path"):
        net.recon_render_depth(torch.zeros(2, 235), FakeModel(), torch.zeros(3, 10), 8, 8, 8)


def test_product_has_no_oracle_or_cpu_path():
    """The shipped package must not import the oracle or numpy-compute its results (DESIGN.md 'no fallback')."""
    pkg = os.path.join(ROOT, "3dfacerecon_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "libref_render_depth" not in text, f


def test_shard_batch():
    sh = fr("sharding")
    for B in (0, 1, 7, 64, 4096, 4099):
        for W in (1, 2, 4, 8):
            parts = sh.shard_slices(B, W)
            assert sum(c for _, c in parts) == B
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(W - 1)) and parts[0][0] == 0
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    with pytest.raises(ValueError):
        sh.shard_batch(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import importlib
    import sys
    sys.path.insert(0, ROOT)
    d = importlib.import_module("3dfacerecon_b200.distributed")
    sh = importlib.import_module("3dfacerecon_b200.sharding")
    r, lr, w = d.init_from_env(backend="gloo")
    start, count = sh.shard_batch(4099, w, r)
    d.barrier()
    total = d.reduce_scalar(count, "sum")
    slowest = d.reduce_scalar(10.0 + r, "max")
    q.put((r, start, count, total, slowest))
    d.shutdown()


def test_two_rank_gloo_plumbing():
    """What bench.py does at N > 1: shard units over ranks, barrier, sum the units, take the max of the timings."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert [r[1:3] for r in res] == [(0, 2050), (2050, 2049)]
    assert all(r[3] == 4099.0 and r[4] == 11.0 for r in res)


def test_synthetic_model_shapes():
    m = fr("synth").make_synthetic_model(grid=(9, 11), ndim_shape=4, ndim_exp=2, seed=1)
    n = 99
    assert m["mu"].shape == (3 * n, 1) and m["pc_shape"].shape == (3 * n, 4) and m["pc_exp"].shape == (3 * n, 2)
    assert m["tri"].shape == (3, 2 * 8 * 10) and m["tri"].dtype == np.float32
    assert m["tri"].min() == 0 and m["tri"].max() == n - 1
    assert set(m) == {"vertex", "tri", "mu", "mu_tex", "pc_tex", "param_tex", "pc_shape", "pc_exp", "ndim_shape", "ndim_exp", "ndim_pose"}
    p = fr("synth").sample_params_constrained(5, 4, 2)
    assert p.shape == (5, 13) and (p[:, 5] == 0).all() and (np.abs(p[:, :3]) <= 1.5).all()


def test_label_wire_format_round_trip(tmp_path):
    """the 235-line %.6f label files (utils/data_process.py:38-60): batch shape, precision, error behaviour."""
    dp = fr("utils.data_process")
    params = fr("synth").sample_params_constrained(3, seed=9)
    files = []
    for i, p in enumerate(params):
        files.append(str(tmp_path / ("%d.txt" % i)))
        dp.write_label(files[-1], p)
    got = dp.prepare_input_label(files, 3, 235)
    assert got.shape == (3, 1, 1, 235) and got.dtype == np.float64
    assert np.abs(got[:, 0, 0, :] - params).max() <= 5.1e-7 + 1e-7 * np.abs(params).max()    # %.6f quantisation
    with pytest.raises(FileNotFoundError):
        dp.prepare_input_label([files[0], str(tmp_path / "missing.txt"), files[2]], 3, 235)
    short = str(tmp_path / "short.txt")
    dp.write_label(short, params[0][:100])
    with pytest.raises(IOError):
        dp.prepare_input_label([short], 1, 235)


def test_fp16_pair_operand_arithmetic_budget():
    """Host emulation of the operand arithmetic of recon_f16.cuh (no GPU): column-scaled basis split into fp16 hi + lo,
    per-face scaled coefficients split into fp16 b0 + b1, the three products the kernel issues (hi.b0 + lo.b0 + hi.b1),
    exact accumulation -- the representation alone must stay far inside the 1e-5 vertex tolerance, also for a model whose
    columns differ by many orders of magnitude and for coefficients of very different sizes."""
    synth = fr("synth")
    m = synth.make_synthetic_model(grid=(23, 31), ndim_shape=12, ndim_exp=5, seed=3, jitter=0.2)
    ks, ke = 12, 5
    P = np.concatenate([m["pc_shape"], m["pc_exp"], m["mu"].reshape(-1, 1)], axis=1).astype(np.float64)    # [3N, K]
    rng = np.random.default_rng(0)
    P[:, 3] *= 1e-6                                                        # a tiny column and a huge one
    P[:, 7] *= 1e5
    params = synth.sample_params_constrained(9, ks, ke, 64, seed=5, full_range=True)
    coef = np.concatenate([params[:, 7:], np.ones((9, 1), np.float32)], axis=1).astype(np.float64)         # mean coefficient 1
    coef[2, :5] *= 1e-4
    coef[5, 9] *= 300.0
    # pack time: 2^s_k with colmax 2^s in [2^14, 2^15)  (basis_colscale_kernel)
    colmax = np.abs(P).max(axis=0)
    s = 15 - np.frexp(colmax)[1]
    Ps = (P * np.exp2(s)).astype(np.float32)
    hi = Ps.astype(np.float16)
    lo = (Ps - hi.astype(np.float32)).astype(np.float16)
    assert np.isfinite(hi.astype(np.float64)).all() and np.abs(hi.astype(np.float64)).max() < 2.0 ** 15 + 16
    # prep kernel: c = coef 2^-s 2^t with max_k |c| in [2^13, 2^14)  (recon_prep_f16_kernel)
    c = (coef * np.exp2(-s)).astype(np.float32)
    t = 14 - np.frexp(np.abs(c).max(axis=1))[1]
    cs = (c * np.exp2(t)[:, None]).astype(np.float32)
    b0 = cs.astype(np.float16)
    b1 = (cs - b0.astype(np.float32)).astype(np.float16)
    H, L, B0, B1 = (a.astype(np.float64) for a in (hi, lo, b0, b1))
    got = (H @ B0.T + L @ B0.T + H @ B1.T) * np.exp2(-t)[None, :]          # the kernel's three MMAs, exact accumulation
    want = P @ coef.T
    err = np.abs(got - want).max(axis=0) / np.abs(want).max(axis=0)
    assert err.max() < 2e-6, err                                          # representation: ~2^-21 relative, tolerance 1e-5


def test_backward_operand_arithmetic_budget():
    """Host emulation of the tensor-core backward's operand arithmetic (recon_bwd_f16.cuh): the rotated vertex gradient scaled
    per face by a power of two from its own maximum and split into fp16 hi + lo, the column-scaled basis split likewise, the
    three products the kernel issues, the mean column kept in float32/64 -- against the float64 contraction.  Faces with
    gradients eight orders of magnitude apart must not disturb each other."""
    synth = fr("synth")
    m = synth.make_synthetic_model(grid=(23, 31), ndim_shape=12, ndim_exp=5, seed=3, jitter=0.2)
    P = np.concatenate([m["pc_shape"], m["pc_exp"]], axis=1).astype(np.float64)                            # [3N, K], mean handled apart
    rng = np.random.default_rng(1)
    B = 6
    dv = rng.normal(size=(B, P.shape[0])).astype(np.float32).astype(np.float64)
    dv[1] *= 1e-6
    dv[4] *= 1e3
    colmax = np.abs(np.concatenate([P, m["mu"].reshape(-1, 1).astype(np.float64)], axis=1)).max(axis=0)
    s = (15 - np.frexp(colmax)[1])[:P.shape[1]]
    Ps = (P * np.exp2(s)).astype(np.float32)
    hi = Ps.astype(np.float16)
    lo = (Ps - hi.astype(np.float32)).astype(np.float16)
    gm = 2.0 * np.abs(dv).max(axis=1)                                      # the kernel bounds |dv| by 2 max|g|; here dv itself
    u = 14 - np.frexp(gm)[1]
    dvs = (dv * np.exp2(u)[:, None]).astype(np.float32)
    ghi = dvs.astype(np.float16)
    glo = (dvs - ghi.astype(np.float32)).astype(np.float16)
    assert np.isfinite(ghi.astype(np.float64)).all()
    H, L, GH, GL = (a.astype(np.float64) for a in (hi, lo, ghi, glo))
    got = (GH @ H + GH @ L + GL @ H) * np.exp2(-u)[:, None] * np.exp2(-s)[None, :]   # D 2^-u_b 2^-s_k, exact accumulation
    want = dv @ P
    err = np.abs(got - want).max(axis=1) / np.abs(want).max(axis=1)
    assert err.max() < 5e-6, err                                          # gradients are judged at 1e-4


def test_read_3dmm_model_from_mat_files(tmp_path):
    """utils/parser_3dmm.py:6-61 on real .mat files (written here with scipy.io.savemat in the BFM layout): same dict keys,
    mu = mu_shape + mu_exp (:32), ndim_pose = 7 (:49), and the MATLAB 1-based ``tri`` is returned unshifted like the
    reference does -- tri_is_one_based flags it for DeviceModel(tri_base=None)."""
    sio = pytest.importorskip("scipy.io")
    parser = fr("utils.parser_3dmm")
    rng = np.random.default_rng(0)
    n, t, ks, ke = 40, 60, 199, 29
    shape = {"mu_shape": rng.normal(size=(3 * n, 1)).astype(np.float32), "w": rng.normal(size=(3 * n, ks)).astype(np.float32),
             "tri": rng.integers(1, n + 1, (3, t)).astype(np.float64), "tex": rng.uniform(0, 255, (3, n)).astype(np.float32),
             "w_tex": rng.normal(size=(3 * n, 5)).astype(np.float32), "alpha_tex": rng.normal(size=(5, 1)).astype(np.float32)}
    shape["tri"][0, 0], shape["tri"][1, 0] = 1, n                     # both ends of the 1-based range occur
    exp = {"mu_exp": rng.normal(size=(3 * n, 1)).astype(np.float32), "w_exp": rng.normal(size=(3 * n, ke)).astype(np.float32)}
    code = {"vertex_code": rng.uniform(0, 1, (3, n)).astype(np.float32)}
    sio.savemat(str(tmp_path / "Model_Shape.mat"), shape)
    sio.savemat(str(tmp_path / "Model_Expression.mat"), exp)
    sio.savemat(str(tmp_path / "vertex_code.mat"), code)
    m = parser.read_3dmm_model(str(tmp_path))
    assert sorted(m) == sorted(["vertex", "tri", "mu", "mu_tex", "pc_tex", "param_tex", "pc_shape", "pc_exp", "ndim_shape", "ndim_exp",
                                "ndim_pose"])
    assert np.array_equal(m["mu"], shape["mu_shape"] + exp["mu_exp"])
    assert np.array_equal(m["pc_shape"], shape["w"]) and np.array_equal(m["pc_exp"], exp["w_exp"])
    assert np.array_equal(m["vertex"], code["vertex_code"]) and np.array_equal(m["mu_tex"], shape["tex"])
    assert (m["ndim_shape"], m["ndim_exp"], m["ndim_pose"]) == (ks, ke, 7)
    assert np.array_equal(m["tri"], shape["tri"])                     # unshifted, as the reference returns it
    assert parser.tri_is_one_based(m["tri"], n) and not parser.tri_is_one_based(m["tri"] - 1, n)
    with pytest.raises(FileNotFoundError):
        parser.read_3dmm_model(str(tmp_path / "missing"))


def test_forward_kernel_schedule_covers_every_cluster_step_once():
    """csrc/schedule.h on the host: for every launch geometry (clusters, CTAs per batch tile, epilogue steps, pair-shared first
    round, dynamic pool) the static walks of all CTAs plus the pool items visit every (cluster, step) exactly once; the static
    load is balanced to within two clusters; pool items never reach into the static part."""
    import ctypes
    import subprocess
    src = os.path.join(ROOT, "tests", "host_emul", "schedule_emul.cc")
    so = os.path.join(ROOT, "tests", "host_emul", "libschedule_emul.so")
    hdr = os.path.join(ROOT, "3dfacerecon_b200", "csrc", "schedule.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src], check=True)
    lib = ctypes.CDLL(so)
    lib.fr_emul_schedule_cover.restype = ctypes.c_int
    stats = (ctypes.c_int * 4)()
    seen_pool = 0
    for nclusters in (1, 2, 3, 7, 74, 75, 147, 148, 149, 160, 195, 221, 222, 223, 296, 304, 370, 371, 416, 510, 511, 998, 2000):
        for g in (1, 2, 3, 9, 37, 74, 132, 147, 148):
            if g > nclusters:
                continue                                               # the launcher never uses more CTAs than row tiles
            for nsteps in (1, 2, 3, 5, 8):
                for share in (0, 1):
                    for pool in (0, 1):
                        rc = lib.fr_emul_schedule_cover(nclusters, g, nsteps, share, pool, stats)
                        assert rc == 0, (rc, nclusters, g, nsteps, share, pool)
                        lo, hi, npool, nitems = list(stats)
                        assert hi - lo <= 2 * nsteps, (nclusters, g, nsteps, share, pool, lo, hi)
                        if pool and nsteps >= 2 and npool:
                            seen_pool += 1
                            assert nitems == npool * min(2, nsteps)
    assert seen_pool > 100
    # the benchmark's geometry: 510 clusters on 148 CTAs, 8 octets -> 74 shared + 2 x 148 static + a pool of 140 clusters in halves
    assert lib.fr_emul_schedule_cover(510, 148, 8, 1, 1, stats) == 0 and list(stats)[2:] == [140, 280]
