#!/usr/bin/env python
"""Benchmark of the params -> depth-map hot path (BASELINE.json metric: faces/sec, 3DMM recon + 200x200 depth render).

    python bench.py [--gpus N --steps K --warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                      # the reference's CPU path on the host cores

Workload of the headline (BASELINE.json configs[1]): 64 synthetic 235-d parameter vectors per GPU -> BFM-sized
reconstruction (53 215 vertices, 199+29 components), pose projection, 200x200 z-buffer depth render (depth + tri_ind),
forward only.  At N > 1 every rank runs the same per-GPU batch on its own parameter shard (weak scaling, no data-path
collective: faces are independent, the basis is replicated -- SURVEY.md 8e).

One step = one pass of the hot path over one batch.  Prints ONE JSON line (rank 0):
  value      faces/s with inputs resident in HBM: EXACTLY K calls back to back on one stream between two CUDA events
             (barrier + synchronize on both sides, max over ranks); no L2 flush between them -- every call streams 184 MB
             of operand tiles through a 126 MB L2 (inputs larger than L2).  `isolated_call` is the same call timed one at a
             time with the L2 flushed before it (the round-1 methodology); all `roofline` figures are taken that way
  e2e        the same metric through the host-buffer C-ABI session: params copied in from pinned host memory and the depth
             maps copied back inside the timed region; next to it a D2H-only loop over the same pinned buffers (the host
             link's ceiling for this result size)
  roofline   dominant kernel (tensor-core reconstruction with the tile rasterizer in its epilogue): algorithmic bytes / its
             device time, measured live with CUDA events the library records between its kernels, vs the measured HBM copy
             bandwidth; the whole step under three byte conventions beside it
  config3    BASELINE configs[2]: 4096 faces sharded by batch over the N ranks (shard_batch), the same fused call
  cpu_baseline  one P-thread float32 GEMM for the whole batch + the reference CPU op (oracle/_ref) on P processes
  extras     (N = 1) configs[0] (batch 1, sample_test conventions, HBM GB/s), configs[3] (batch 256 forward + backward),
             the same step on a mesh with shuffled vertex / triangle numbering, and the records pipeline (no FR_CLUSTER_TILES)
"""
from __future__ import annotations

import argparse
import ctypes
import importlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "3dfacerecon_b200"

H = W = 200
IM_SIZE = 200.0
METRIC = "faces_per_sec_recon_plus_200x200_depth_render"
UNIT = "faces/s"
CONFIG3_FACES = 4096
WORKLOAD = ("BASELINE configs[1]: batch-64 synthetic 235-d params -> 3DMM recon + pose projection + 200x200 depth render "
            "(depth + tri_ind), forward only, per GPU")


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own path on the host cores -- numpy float32 recon as nets/network.py:153-169 (ONE GEMM over the
# whole batch with all BLAS threads) + the reference CPU op on P processes (it is non-reentrant: static scratch,
# render_depth_op.cc:125-131 => processes, not threads; they read the vertices from a shared buffer)
# ----------------------------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_worker_init(model_small, use_ref, shared, shape):
    try:
        from threadpoolctl import threadpool_limits
        _CPU["limit"] = threadpool_limits(1)
    except Exception:
        pass
    import numpy as np
    import oracle  # noqa: F401  (test/bench infrastructure: the CPU checker doubles as the CPU baseline)
    _CPU["model"] = model_small
    _CPU["use_ref"] = use_ref
    _CPU["vp"] = np.frombuffer(shared, dtype=np.float32).reshape(shape)


def _cpu_worker_render(span):
    """Reference CPU render of faces [lo, hi) of the shared vertex buffer; returns a checksum."""
    import numpy as np
    import oracle
    lo, hi = span
    if hi <= lo:
        return 0.0
    m = _CPU["model"]
    vp = np.ascontiguousarray(_CPU["vp"][lo:hi])
    B = vp.shape[0]
    if _CPU["use_ref"]:
        tex = np.ascontiguousarray(np.broadcast_to(m["vertex"], (B,) + m["vertex"].shape))
        depth = oracle.ref_render_depth(vp, m["tri"], tex, (B, H, W, 3))[0]
    else:
        depth = oracle.oracle_render_depth_forward(vp, m["tri"], m["vertex"], H, W)[0]
    return float(depth[depth > -1e13].sum())


class CpuArm:
    def __init__(self, model, batch, cores=None):
        import multiprocessing as mp
        import oracle
        self.model = model
        self.batch = batch
        self.use_ref = oracle.ref_available()
        self.cores = cores or max(1, len(os.sched_getaffinity(0)))
        nver = model["mu"].size // 3
        self.shape = (batch, 3, nver)
        ctx = mp.get_context("fork")
        self.shared = ctx.RawArray(ctypes.c_float, batch * 3 * nver)
        small = {"tri": model["tri"], "vertex": model["vertex"]}
        self.pool = ctx.Pool(self.cores, initializer=_cpu_worker_init, initargs=(small, self.use_ref, self.shared, self.shape))
        import numpy as np
        self.nver = nver
        self.ks = int(model["pc_shape"].shape[1])
        self.basis = np.ascontiguousarray(np.concatenate([model["pc_shape"], model["pc_exp"]], axis=1), np.float32)   # network.py:351 does the same concat
        self.mu = np.ascontiguousarray(np.asarray(model["mu"], np.float32).reshape(3 * nver, 1))

    def recon(self, params, out=None):
        """nets/network.py:153-169 in float32 numpy, laid out for speed: ONE [3N,228] x [228,B] GEMM on all BLAS threads (the
        reference issues two, :153,155), mean added, then the batched (f.R) v + t and the y flip.  Same values as the checker
        oracle.recon.vertices_transform(dtype=float32) up to float32 summation order (verified once in self_check)."""
        import numpy as np
        from oracle import recon
        p = np.asarray(params, np.float32)
        v = p[:, 7:] @ self.basis.T                                               # [B, 3N]: one GEMM (BLAS reads the basis transposed in place)
        v += self.mu.T
        vb = v.reshape(len(p), 3, self.nver)                                      # [B, 3, N] planar (network.py:154)
        rot = recon.rotation_matrix_batch(p[:, 0:3])                              # float64 sin/cos -> float32, network.py:292-297
        m = p[:, 6, None, None] * rot                                             # network.py:165
        vp = np.matmul(m, vb, out=out)
        vp += p[:, 3:6, None]
        vp[:, 1, :] = np.float32(IM_SIZE) - vp[:, 1, :] - np.float32(1)           # network.py:168
        return vp

    def self_check(self, params):
        import numpy as np
        from oracle import recon
        want = recon.vertices_transform(params[:2], self.model, IM_SIZE, dtype=np.float64)
        got = self.recon(params[:2])
        err = float(np.abs(got - want).max() / np.abs(want).max())
        assert err < 1e-5, err
        return err

    def step(self, params):
        import numpy as np
        self.recon(params, out=np.frombuffer(self.shared, dtype=np.float32).reshape(self.shape)[:len(params)])
        bounds = np.linspace(0, len(params), self.cores + 1).astype(int)
        return sum(self.pool.map(_cpu_worker_render, list(zip(bounds[:-1], bounds[1:])), chunksize=1))

    def time_steps(self, params, steps, warmup):
        self.self_check(params)
        for _ in range(warmup):
            self.step(params)
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step(params)
        return (time.perf_counter() - t0) / steps

    def close(self):
        self.pool.close()
        self.pool.join()

    @property
    def kind(self):
        return "reference" if self.use_ref else "port"

    def describe(self, steps):
        return ("%d steps of the same %d-face batch: numpy float32 recon+projection (network.py:153-169) as one %d-thread GEMM + "
                "%s, faces split over %d processes" % (steps, self.batch, self.cores,
                                                      "the reference CPU op (oracle/_ref, render_depth_op.cc compiled in place)"
                                                      if self.use_ref else "the oracle C port", self.cores))


# ----------------------------------------------------------------------------------------------------------------------
def _clock_sampler_start(gpu_index):
    """nvidia-smi sampling in the background (B200_PROFILING.md clocks line, plus a timestamp)."""
    q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    try:
        f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "50"],
                             stdout=f, stderr=subprocess.DEVNULL)
        return p, f
    except Exception:
        return None, None


def _clock_sampler_stop(p, f, t_begin, t_end):
    """Summarise the samples whose timestamp falls inside [t_begin, t_end] (time.time() seconds)."""
    import datetime
    if p is None:
        return None
    p.terminate()
    try:
        p.wait(5)
    except Exception:
        p.kill()
    f.flush()
    f.seek(0)
    sm, mx, reasons, total = [], [], set(), 0
    for line in f.read().splitlines():
        c = [x.strip() for x in line.split(",")]
        if len(c) < 9:
            continue
        total += 1
        try:
            ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            if ts < t_begin - 0.05 or ts > t_end + 0.05:
                continue
            sm.append(float(c[1]))
            mx.append(float(c[2]))
        except ValueError:
            continue
        for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
            if v.lower().startswith("active"):
                reasons.add(name)
    f.close()
    try:
        os.unlink(f.name)
    except OSError:
        pass
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "samples_total": total}
    return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
            "window": "timed loops + %.1f s soak of the same step" % SOAK_SECONDS}


SOAK_SECONDS = 1.5


def _peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _traffic(kernel):
    """dram bytes per launch from the committed ncu --set full capture (profiles/roofline_traffic.json), or None."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        return json.load(open(path)).get(kernel)
    except Exception:
        return None


def algorithmic_bytes(B, nver, ntri, K):
    """compulsory HBM bytes per launch of B faces, split into the recon part (basis + mean + params read,
    vertex_proj written) and the render part (tri + vertex_proj read, depth + tri_ind written)."""
    n3 = 3 * nver
    recon = 4 * (n3 * K + n3 + B * (7 + K) + B * n3)
    render = 4 * (3 * ntri + B * n3 + 2 * B * H * W)
    return recon, render


def fused_compulsory_bytes(B, nver, ntri, K):
    """What the fused call cannot avoid: basis + mean + params + tri read, depth + tri_ind written (no vertex tensor)."""
    n3 = 3 * nver
    return 4 * (n3 * K + n3 + B * (7 + K) + 3 * ntri + 2 * B * H * W)


def backward_bytes(B, nver, ntri, K):
    """SURVEY.md 8(d) bytes_bwd(B)."""
    n3 = 3 * nver
    return 4 * (n3 * K + 3 * ntri + B * (2 * H * W + 2 * n3 + n3 + 2 * (7 + K)))


def _numa_bind(local_rank):
    """Best effort: pin this process (and so its pinned allocations' first touch) to the CPUs of the GPU's NUMA node."""
    info = {"gpu_numa_node": None, "bound": False}
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=10).stdout.strip().lower()
        bus = out[4:] if out.startswith("0000") and len(out) > 12 else out          # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        info["gpu_numa_node"] = node
        if node >= 0:
            cpus = set()
            for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                info["bound"] = True
                info["cpus"] = len(cpus)
    except Exception as exc:     # virtualised hosts often expose no topology: report it, carry on
        info["error"] = type(exc).__name__
    return info


def run_ours(args):
    import numpy as np
    import torch
    pkg = importlib.import_module(PKG)
    synth = importlib.import_module(PKG + ".synth")
    dist = importlib.import_module(PKG + ".distributed")
    sharding = importlib.import_module(PKG + ".sharding")
    rank, local_rank, world = dist.init_from_env()
    if world != args.gpus and rank == 0:
        print("note: WORLD_SIZE=%d but --gpus %d; using WORLD_SIZE" % (world, args.gpus), file=sys.stderr)
    B = args.batch
    model = synth.make_synthetic_model(seed=0, jitter=0.2)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = CpuArm(model, B)                                         # fork the pool before CUDA is initialised / the affinity changes
    numa = _numa_bind(local_rank)

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = pkg._lib.lib()
    check = pkg._lib.check
    dm = pkg.DeviceModel(model, dev)
    nver, ntri, ks, ke = dm.nver, dm.ntri, dm.ndim_shape, dm.ndim_exp
    K = ks + ke
    params_all = synth.sample_params_constrained(B * world, seed=2)
    params_host = params_all[rank * B:(rank + 1) * B]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream

    class Fused:
        """Buffers + the fused call for one (model, batch)."""

        def __init__(self, dmodel, params_np):
            self.dm = dmodel
            self.B = int(params_np.shape[0])
            self.params = torch.from_numpy(np.ascontiguousarray(params_np, np.float32)).to(dev)
            self.depth = torch.empty((self.B, H, W, 1), dtype=torch.float32, device=dev)
            self.tri_ind = torch.empty((self.B, H, W, 1), dtype=torch.float32, device=dev)
            self.rb = lib.fr_recon_workspace_bytes(self.B, dmodel.nver, ks, ke)
            self.ws = torch.empty(lib.fr_pipeline_workspace_bytes(self.B, dmodel.nver, ks, ke, H, W), dtype=torch.uint8, device=dev)

        def __call__(self, vertex_ptr=None, events=None, stream_ptr=None):
            d = self.dm
            check(lib.fr_recon_render_forward(self.params.data_ptr(), d.packed.data_ptr(), d.tri.data_ptr(), d.mesh.handle, vertex_ptr,
                                              self.depth.data_ptr(), self.tri_ind.data_ptr(), self.B, d.nver, d.ntri, ks, ke, H, W,
                                              IM_SIZE, d.run_flags, self.ws.data_ptr(), self.ws.numel(),
                                              sp if stream_ptr is None else stream_ptr, events))

    main = Fused(dm, params_host)
    vertex = torch.empty((B, 3, nver), dtype=torch.float32, device=dev)

    def step_recon():
        check(lib.fr_recon_project_forward(main.params.data_ptr(), dm.packed.data_ptr(), dm.mesh.handle, vertex.data_ptr(), B, nver,
                                           ks, ke, IM_SIZE, dm.run_flags, main.ws.data_ptr(), main.rb, sp))

    def step_render():
        check(lib.fr_render_depth_forward(vertex.data_ptr(), dm.tri.data_ptr(), None, 0, main.depth.data_ptr(), None, None,
                                          main.tri_ind.data_ptr(), B, nver, ntri, H, W, dm.mesh.handle, main.ws.data_ptr() + main.rb,
                                          main.ws.numel() - main.rb, sp))

    def timed(fn, steps, warmup, barrier=True):
        """Per-step CUDA-event timing on the launching stream, L2 flushed (outside the events) before every step."""
        for _ in range(warmup):
            flush.zero_()
            fn()
        if barrier:
            dist.barrier()
        torch.cuda.synchronize(dev)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in evs:
            flush.zero_()
            a.record(stream)
            fn()
            b.record(stream)
        torch.cuda.synchronize(dev)
        if barrier:
            dist.barrier()
        return sum(a.elapsed_time(b) for a, b in evs) / steps          # ms per step

    def timed_back_to_back(fn, steps, warmup):
        """EXACTLY `steps` calls issued back to back on the launching stream between two CUDA events (barrier + synchronize on
        both sides), no L2 flush in between: every call streams the model's 184 MB operand section (evict-first), more than
        the 126 MB L2 holds, so no step finds its large input cached; what does stay warm is what stays warm in any serving
        loop (mesh table, parameters, the call's own keys)."""
        for _ in range(warmup):
            fn()
        dist.barrier()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(steps):
            fn()
        b.record(stream)
        torch.cuda.synchronize(dev)
        dist.barrier()
        return a.elapsed_time(b) / steps

    def timed_parts(call, steps, warmup):
        """The fused step with events recorded by the library between its kernels (stage_events): device time of the
        reconstruction kernels, the visibility kernel and the resolve kernel of the same real step, L2 flushed before each."""
        for _ in range(warmup):
            flush.zero_()
            call()
        torch.cuda.synchronize(dev)
        rows = []
        for _ in range(steps):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            for e in ev[1:]:
                e.record(stream)                                       # creates the underlying cudaEvent
            flush.zero_()
            ev[0].record(stream)
            call(None, (ctypes.c_void_p * 3)(ev[1].cuda_event, ev[2].cuda_event, ev[3].cuda_event))
            rows.append(ev)
        torch.cuda.synchronize(dev)
        return [sum(r[i].elapsed_time(r[i + 1]) for r in rows) / steps for i in range(3)]

    sampler = _clock_sampler_start(local_rank) if rank == 0 else (None, None)
    time.sleep(0.3 if rank == 0 else 0.0)                             # let nvidia-smi come up
    t_begin = time.time()
    launches0 = lib.fr_launch_count()
    ms_b2b = timed_back_to_back(main, args.steps, max(args.warmup, 3))         # the headline: K steps back to back
    launches_timed = (lib.fr_launch_count() - launches0) * args.steps // (args.steps + max(args.warmup, 3))
    ms_full = timed(main, args.steps, args.warmup)                              # the same call in isolation, L2 flushed before it
    ms_full_vertex = timed(lambda: main(vertex.data_ptr()), args.steps, args.warmup)   # same call, vertex_proj materialised too
    ms_parts = timed_parts(main, args.steps, args.warmup)
    ms_recon = timed(step_recon, args.steps, args.warmup)
    ms_render = timed(step_render, args.steps, args.warmup)

    # BASELINE configs[2]: 4096 faces sharded by batch over the ranks, the same fused call on each shard
    c3_start, c3_count = sharding.shard_batch(CONFIG3_FACES, world, rank)
    c3 = None
    if not args.no_config3:
        p3 = synth.sample_params_constrained(CONFIG3_FACES, seed=3)[c3_start:c3_start + c3_count]
        shard = Fused(dm, p3)
        ms_c3 = timed(shard, max(3, min(args.steps, 10)), 3)
        c3 = {"ms": dist.reduce_scalar(ms_c3, "max"), "faces": int(dist.reduce_scalar(c3_count, "sum"))}
        del shard

    # the timed loops last only milliseconds: keep the same step running so the 50 ms clock sampler sees it under load
    t_soak = time.time()
    while time.time() - t_soak < SOAK_SECONDS:
        for _ in range(20):
            main()
        torch.cuda.synchronize(dev)
    t_end = time.time()
    clocks = _clock_sampler_stop(*sampler, t_begin, t_end) if rank == 0 else None
    red = lambda v, op="max": dist.reduce_scalar(v, op)
    ms_full_max, ms_full_vertex_max = red(ms_full), red(ms_full_vertex)
    ms_b2b_max = red(ms_b2b)
    ms_parts_max = [red(v) for v in ms_parts]
    ms_recon_max, ms_render_max = red(ms_recon), red(ms_render)
    faces_total = red(B, "sum")
    launches_total = int(red(launches_timed, "sum"))

    # parity gate that travels with every measurement: faces of this very output against the CPU oracle
    parity = None
    if rank == 0 and not args.no_parity:
        import oracle
        from oracle import recon as orecon
        nchk = min(8, B)
        main()
        torch.cuda.synchronize(dev)
        depth_timed = main.depth.clone()
        main(vertex.data_ptr())                                           # same call, this time materialising vertex_proj
        torch.cuda.synchronize(dev)
        assert depth_timed.cpu().numpy().tobytes() == main.depth.cpu().numpy().tobytes()
        sel = np.linspace(0, B - 1, nchk).astype(int)
        vp = vertex[torch.from_numpy(sel).to(dev)].cpu().numpy()
        want_vp = orecon.vertices_transform(params_host[sel], model, IM_SIZE)
        want = oracle.oracle_render_depth_forward(vp, model["tri"], model["vertex"], H, W)
        got_t = main.tri_ind.cpu().numpy()[sel]
        got_d = main.depth.cpu().numpy()[sel]
        parity = {"faces_checked": int(nchk),
                  "vertex_rel_err": float(np.abs(vp - want_vp).max() / np.abs(want_vp).max()),
                  "tri_ind_bit_exact": bool(got_t.tobytes() == want[3].tobytes()),
                  "depth_bit_exact": bool(got_d.tobytes() == want[0].tobytes())}

    peak, peak_src = _peak_hbm()

    def roof(nbytes, ms):
        gbs = nbytes / (ms * 1e-3) / 1e9
        return {"algorithmic_bytes": int(nbytes), "ms": ms, "achieved": gbs, "frac": gbs / peak}

    # other BASELINE configs and flavours, reported as extras (not the headline)
    extras = None
    if rank == 0 and world == 1 and not args.no_extras:
        net = importlib.import_module(PKG + ".nets.network")
        ops = importlib.import_module(PKG + ".rendering_layer.ops")
        extras = {}

        def time_torch(fn, reps, flush_l2=False):
            for _ in range(3):
                fn()
            torch.cuda.synchronize(dev)
            tot = 0.0
            for _ in range(reps):
                if flush_l2:
                    flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                fn()
                b.record(stream)
                torch.cuda.synchronize(dev)
                tot += a.elapsed_time(b)
            return tot / reps

        # configs[0]: rendering_layer/sample_test.py -- one face, sample_test conventions (interleaved mean, Rz.Ry.Rx, y' = S - y),
        # pose exactly [0,0,0,S/2,S/2,0,1e-3] (sample_test.py:23-38), all four render outputs, L2 flushed
        model_b = dict(model)                     # the same surface with the mean stored interleaved, as sample_test.py:101 reads it
        model_b["mu"] = np.ascontiguousarray(model["mu"].reshape(3, nver).T).reshape(3 * nver, 1)
        dm_b = pkg.DeviceModel(model_b, dev, convention="sample_test")
        p1 = torch.from_numpy(synth.sample_params_sample_test(seed=1).astype(np.float32)).to(dev)
        img1 = torch.empty((1, H, W, 3), device=dev)

        def fwd1():
            with torch.no_grad():
                vp1 = net.recon_project(p1, dm_b, IM_SIZE)
                ops.render_depth(vp1, dm_b.tri, dm_b.vertex_code.unsqueeze(0), img1)

        ms_1 = time_torch(fwd1, 20, flush_l2=True)
        rb1, nb1 = algorithmic_bytes(1, nver, ntri, K)
        extras["config1_b1_sample_test"] = dict(roof(rb1 + nb1 + 4 * (3 * nver + 6 * H * W), ms_1),
                                                api="recon_project + render_depth (all four outputs), torch API, sample_test convention, "
                                                    "L2 flushed before every call; bytes = SURVEY 8(d) bytes_fwd(1) + texture read + "
                                                    "texture_image / normal written")
        del dm_b, model_b

        # configs[3]: batch-256 forward + backward of the whole path (d depth -> d params), torch API, all four render outputs
        B4 = 256
        p4 = torch.from_numpy(synth.sample_params_constrained(B4, seed=4)).to(dev).requires_grad_(True)
        img4 = torch.empty((B4, H, W, 3), device=dev)
        tex4 = dm.vertex_code.unsqueeze(0).expand(B4, -1, -1)
        gd4 = torch.randn((B4, H, W, 1), device=dev)

        def fwd_bwd():
            p4.grad = None
            vp4 = net.recon_project(p4, dm, IM_SIZE)
            d4, _, _, ti4 = ops.render_depth(vp4, dm.tri, tex4, img4)
            (d4 * (gd4 * (ti4 >= 0))).sum().backward()

        def fwd_only4():
            with torch.no_grad():
                vp4 = net.recon_project(p4, dm, IM_SIZE)
                ops.render_depth(vp4, dm.tri, tex4, img4)

        def fwd_bwd_one_call():                                          # the same outputs and gradient from recon_render_depth
            p4.grad = None
            _, d4, _, _, ti4 = net.recon_render_depth(p4, dm, dm.vertex_code, H, W, IM_SIZE)
            (d4 * (gd4 * (ti4 >= 0))).sum().backward()

        def fwd_only4_one_call():
            with torch.no_grad():
                net.recon_render_depth(p4, dm, dm.vertex_code, H, W, IM_SIZE)

        ms_fb = time_torch(fwd_bwd, 10, flush_l2=True)
        ms_f4 = time_torch(fwd_only4, 10, flush_l2=True)
        ms_fb1 = time_torch(fwd_bwd_one_call, 10, flush_l2=True)
        ms_f41 = time_torch(fwd_only4_one_call, 10, flush_l2=True)
        rb4, nb4 = algorithmic_bytes(B4, nver, ntri, K)
        b4 = rb4 + nb4 + backward_bytes(B4, nver, ntri, K)
        extras["config4_b256_fwd_bwd"] = dict(roof(b4, ms_fb), faces_per_s=B4 / (ms_fb * 1e-3), fwd_only_ms=ms_f4,
                                              with_extra_outputs=roof(b4 + B4 * 4 * (6 * H * W + 3 * nver), ms_fb),
                                              one_call=dict(roof(b4, ms_fb1), faces_per_s=B4 / (ms_fb1 * 1e-3), fwd_only_ms=ms_f41,
                                                            api="recon_render_depth (fr_recon_render_forward_all: vertices + all four "
                                                                "outputs from the rasterizing reconstruction kernel, no repack / visibility "
                                                                "kernels) + the same autograd backward; bit-identical outputs"),
                                              api="recon_project + render_depth (all four outputs) + autograd backward, torch API, "
                                                  "L2 flushed; bytes = SURVEY 8(d) bytes_fwd(256) + bytes_bwd(256)")
        del p4, img4, gd4

        # the headline step on the same surface with randomly renumbered vertices and shuffled triangles
        model_p = synth.make_synthetic_model(seed=0, jitter=0.2, permute=True)
        dm_p = pkg.DeviceModel(model_p, dev)
        fp = Fused(dm_p, params_host)
        ms_perm = timed(fp, 10, 3, barrier=False)
        extras["shuffled_mesh_b64"] = {"ms": ms_perm, "slowdown_vs_grid_order": ms_perm / ms_full_max,
                                       "note": "same fused call; vertex / triangle numbering of the mesh randomly permuted -- the "
                                               "mesh table's rank order supplies the locality"}
        del fp, dm_p, model_p

        # serving-style throughput: independent 64-face calls issued back to back on S streams (own workspace and outputs per
        # stream, no L2 flush: the basis stream alone, 184 MB per call, is larger than L2).  One stream = no launch / drain
        # gaps between the steps; several = the next call's CTAs fill the SMs the previous call's slowest CTAs leave idle.
        b2b = {}
        for S in (1, 3):
            streams = [torch.cuda.Stream(dev) for _ in range(S)]
            calls = [Fused(dm, synth.sample_params_constrained(B, seed=20 + i)) for i in range(S)]
            nrun = 120
            for i in range(2 * S):
                calls[i % S](stream_ptr=streams[i % S].cuda_stream)
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for st_ in streams:
                st_.wait_event(a)
            for i in range(nrun):
                calls[i % S](stream_ptr=streams[i % S].cuda_stream)
            for st_ in streams:
                stream.wait_stream(st_)
            b.record(stream)
            torch.cuda.synchronize(dev)
            ms_b = a.elapsed_time(b) / nrun
            b2b["streams_%d" % S] = {"ms_per_call": ms_b, "faces_per_s": B / (ms_b * 1e-3)}
            del calls, streams
        b2b["note"] = ("independent batch-64 calls back to back on S streams (own workspace / outputs per stream), no L2 flush: "
                       "streams_1 repeats the headline's methodology on a side stream; with several calls in flight the next call's "
                       "CTAs fill the SMs the previous call's slowest CTAs leave idle")
        extras["back_to_back_b64"] = b2b

        # the records pipeline (a basis packed without FR_CLUSTER_TILES): reconstruction kernel writing 16-byte vertex
        # records, stand-alone tile rasterizer reading them back from L2, resolve -- the step split by the library's events
        dm_r = pkg.DeviceModel(model, dev, cluster_tiles=False)
        fr_ = Fused(dm_r, params_host)
        ms_rec_pipe = timed(fr_, 10, 3, barrier=False)
        parts_r = timed_parts(fr_, 10, 3)
        torch.cuda.synchronize(dev)
        main()
        torch.cuda.synchronize(dev)
        same = bool(torch.equal(fr_.depth, main.depth) and torch.equal(fr_.tri_ind, main.tri_ind))
        rb_, nb_ = algorithmic_bytes(B, nver, ntri, K)
        extras["records_pipeline_b64"] = {"ms": ms_rec_pipe, "vs_default": ms_rec_pipe / ms_full_max, "bit_identical_to_default": same,
                                          "recon_kernels": roof(rb_, parts_r[0]),
                                          "tile_keys_kernel": roof(nb_ - 8 * B * H * W, parts_r[1]),
                                          "resolve_kernel": roof(16 * B * H * W, parts_r[2]),
                                          "clusters": dm_r.mesh.nclusters, "vertex_slots": dm_r.mesh.vertex_slots,
                                          "note": "kernels serialised by the events; fr::rt::raster_tile_keys_kernel<16> is the "
                                                  "visibility pass of fr_render_depth_forward as well"}
        del fr_, dm_r

    # end to end through the host-buffer C-ABI session: every step copies its params in from pinned host memory and its
    # depth maps back to pinned host memory inside the timed region.  The session's slots are used round-robin
    # (fr_session_submit / fr_session_wait) so the copy-out of step i overlaps the kernels of the following steps; the
    # synchronous single-call latency (fr_session_forward) and a D2H-only loop over the same buffers are reported next to it.
    main()
    torch.cuda.synchronize(dev)
    depth_ref = main.depth.cpu().numpy().tobytes() if rank == 0 else None
    del main, vertex, flush
    sess = pkg.Session(model, H, W, max_batch=B, device=local_rank)
    nslots = pkg._lib.FR_SESSION_SLOTS
    pin_params = [torch.from_numpy(params_host.copy()).pin_memory() for _ in range(nslots)]
    pin_depth = [torch.zeros((B, H, W, 1), dtype=torch.float32).pin_memory() for _ in range(nslots)]
    pp, pd = [t.numpy() for t in pin_params], [t.numpy() for t in pin_depth]

    def e2e_pipelined(steps):
        for i in range(steps):
            slot = i % nslots
            sess.wait(slot)                                             # no-op while the slot is idle
            sess.submit(slot, pp[slot], IM_SIZE, depth=pd[slot])
        for slot in range(nslots):
            sess.wait(slot)

    def e2e_sync(steps):
        for _ in range(steps):
            sess.forward(pp[0], IM_SIZE, depth=pd[0], want_tri_ind=False)   # returns when depth is on the host

    dsrc = torch.zeros((B, H, W, 1), dtype=torch.float32, device=dev)
    copy_streams = [torch.cuda.Stream(dev) for _ in range(nslots)]

    def d2h_only(steps):
        """The result copy alone: the same bytes, the same pinned buffers, one stream per slot."""
        for i in range(steps):
            slot = i % nslots
            copy_streams[slot].synchronize()
            with torch.cuda.stream(copy_streams[slot]):
                pin_depth[slot].copy_(dsrc, non_blocking=True)
        for s in copy_streams:
            s.synchronize()

    def wall(fn, steps):
        fn(max(3, args.warmup))
        dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        fn(steps)
        ms = (time.perf_counter() - t0) * 1e3 / steps
        dist.barrier()
        return dist.reduce_scalar(ms, "max")

    e2e_sync_ms_max = wall(e2e_sync, args.steps)
    e2e_ms_max = wall(e2e_pipelined, args.steps)
    e2e_ok = all(t.numpy().tobytes() == depth_ref for t in pin_depth) if rank == 0 else None
    d2h_ms_max = wall(d2h_only, args.steps)
    sess.close()

    cpu_baseline = None
    if cpu is not None:
        cpu_steps = 3
        sec = cpu.time_steps(params_host, cpu_steps, 1)
        cpu_baseline = {"value": B / sec, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind, "sample": cpu.describe(cpu_steps)}
        cpu.close()

    if rank == 0:
        rb, nb = algorithmic_bytes(B, nver, ntri, K)
        out_bytes = 8 * B * H * W
        ms_rec, ms_keys, ms_res = ms_parts_max
        d2h_bytes = int(pin_depth[0].numel() * 4)
        line = {
            "metric": METRIC, "value": faces_total / (ms_b2b_max * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_b2b_max, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "arithmetic": "f32 results; reconstruction on tcgen05 with fp16 hi/lo operand pairs (22 significant bits) and "
                                     "fp32 accumulation; rasterizer cull in packed integers, inside tests by a certified float filter "
                                     "that falls back to the reference's separately rounded f64 sequence whenever rounding could "
                                     "matter (bit-exact with the reference)",
                       "batch_per_gpu": B, "nver": nver, "ntri": ntri, "ndim_shape": ks, "ndim_exp": ke, "image": [H, W],
                       "l2": "timed region = K calls back to back on one stream, no flush between them: inputs larger than L2 (every "
                             "call streams the 184 MB cluster-tile section of the packed basis, evict-first, through a 126 MB L2); "
                             "the same call in isolation with the L2 flushed before it (512 MiB write, per-step events) is "
                             "`isolated_call`, and every `roofline` figure is taken that way",
                       "call": "fr_recon_render_forward with the model's mesh table and FR_CLUSTER_TILES (rasterizer inside the "
                               "reconstruction epilogue), depth + tri_ind out; the optional vertex_proj output is not requested in the "
                               "timed loop (the parity check re-runs the call with it)",
                       "parallelism": "batch-sharded x%d, basis replicated, no collective" % world},
            "e2e": {"value": faces_total / (e2e_ms_max * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms_max,
                    "h2d_bytes_per_step": int(pin_params[0].numel() * 4), "d2h_bytes_per_step": d2h_bytes,
                    "api": "fr_session_submit/fr_session_wait round-robin over %d slots (host buffers, pinned; the copy-out of "
                           "step i overlaps the kernels of the following steps)" % nslots,
                    "sync_call_ms": e2e_sync_ms_max, "sync_call_value": faces_total / (e2e_sync_ms_max * 1e-3),
                    "d2h_only_ms": d2h_ms_max, "d2h_only_gbs_per_gpu": d2h_bytes / (d2h_ms_max * 1e-3) / 1e9,
                    "pcie_frac": d2h_ms_max / e2e_ms_max,
                    "note": "d2h_only = the result copy alone (same bytes, same pinned buffers, max over ranks); pcie_frac = its share of "
                            "the end-to-end step: near 1 means the host link, not the library, bounds e2e",
                    "numa": numa, "matches_device_path": e2e_ok},
            "isolated_call": {"ms": ms_full_max, "value": faces_total / (ms_full_max * 1e-3), "unit": UNIT,
                              "note": "one call at a time, L2 flushed before each, per-step CUDA events (the round-1 methodology: "
                                      "includes ~7 us of launch / completion latency per step that back-to-back calls hide)"},
            "gpu_launches": launches_total,
            "roofline": dict(roof(rb + nb - out_bytes, ms_rec + ms_keys), bound="hbm",
                             kernel="fr::f16::recon_fwd_f16_kernel<true> (tcgen05 reconstruction + projection with the tile "
                                    "rasterizer's cull / draw in its epilogue; the 4.5 us prep kernel in front of it is inside the time)",
                             peak=peak, unit="GB/s", traffic=_traffic("recon_fwd_f16_kernel_raster"), peak_source=peak_src,
                             how="device time between the CUDA events the library records around its kernels inside one real fused "
                                 "call (stage_events), L2 flushed before the call; bytes = SURVEY 8(d) bytes_fwd(B) minus the resolve "
                                 "pass's outputs (basis + mean + params + tri read, projected vertices written and read once)",
                             limiter="instruction issue of the rasterizer half (DESIGN.md 6): ~39 M warp instructions per step, 24 "
                                     "rasterizing warps per SM at ~56 % issue utilisation -- 12 us of projection / cull / draw per "
                                     "cluster and SM against 7.5 us for the cluster's basis tile; the 200 MB stream alone would take "
                                     "~31 us at the measured peak",
                             resolve_kernel=roof(2 * out_bytes, ms_res),
                             whole_step={"back_to_back": roof(rb + nb, ms_b2b_max),
                                         "survey_8d_bytes": roof(rb + nb, ms_full_max),
                                         "with_vertex_proj_materialised": roof(rb + nb, ms_full_vertex_max),
                                         "fused_compulsory_bytes": roof(fused_compulsory_bytes(B, nver, ntri, K), ms_full_max)},
                             separate_entry_points_ms={"fr_recon_project_forward": ms_recon_max,
                                                       "fr_render_depth_forward": ms_render_max}),
            "config3": None if c3 is None else {
                "workload": "BASELINE configs[2]: %d faces sharded by batch (shard_batch) over %d GPU(s), fused call per shard" % (c3["faces"], world),
                "ms": c3["ms"], "value": c3["faces"] / (c3["ms"] * 1e-3), "unit": UNIT, "scaling": "strong",
                "roofline_whole_step": roof(sum(algorithmic_bytes(-(-CONFIG3_FACES // world), nver, ntri, K)), c3["ms"])},
            "cpu_baseline": cpu_baseline, "clocks": clocks, "parity": parity, "extras": extras,
        }
        print(json.dumps(line))
    dist.shutdown()


def run_reference(args):
    """The reference's CPU implementation of the path on the host cores, same config/metric (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import importlib as il
    synth = il.import_module(PKG + ".synth")
    B = args.batch
    model = synth.make_synthetic_model(seed=0, jitter=0.2)
    params = synth.sample_params_constrained(B, seed=2)
    cpu = CpuArm(model, B)
    sec = cpu.time_steps(params, args.steps, args.warmup)
    value = B / sec
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD.replace("per GPU", "host CPU"), "batch_per_gpu": B, "image": [H, W]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind,
                             "sample": "each step = the full %d-face batch: " % B + cpu.describe(args.steps).split(": ", 1)[1]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    cpu.close()
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="faces per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-config3", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        args.warmup = max(args.warmup, 1)
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args)


if __name__ == "__main__":
    main()
