#!/usr/bin/env python
"""Benchmark of the params -> depth-map hot path (BASELINE.json metric: faces/sec, 3DMM recon + 200x200 depth render).

    python bench.py [--gpus N --steps K --warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                      # the reference's CPU path on the host cores

Workload (BASELINE.json configs[1]): 64 synthetic 235-d parameter vectors per GPU -> BFM-sized reconstruction
(53 215 vertices, 199+29 components), pose projection, 200x200 z-buffer depth render (depth + tri_ind), forward only.
At N > 1 every rank runs the same per-GPU batch on its own parameter shard (weak scaling, no data-path collective:
faces are independent, the basis is replicated -- SURVEY.md 8e).

One step = one pass of the hot path over one batch.  Prints ONE JSON line (rank 0):
  value      faces/s with inputs resident in HBM, device-timed with CUDA events (max over ranks), L2 flushed
             between timed iterations
  e2e        the same metric through the host-buffer C-ABI session (fr_session_forward): params copied in from
             pinned host memory and the depth maps copied back inside the timed region
  roofline   dominant kernel group: algorithmic bytes / measured device time vs the measured HBM copy bandwidth
  cpu_baseline  the reference CPU op (oracle/_ref, or the oracle port) + numpy recon timed on the host cores
"""
from __future__ import annotations

import argparse
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


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own path on the host cores (numpy recon as nets/network.py:153-169 + the reference CPU op)
# ----------------------------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_worker_init(model, use_ref):
    try:
        from threadpoolctl import threadpool_limits
        _CPU["limit"] = threadpool_limits(1)
    except Exception:
        pass
    import oracle  # noqa: F401  (test/bench infrastructure: the CPU checker doubles as the CPU baseline)
    _CPU["model"] = model
    _CPU["use_ref"] = use_ref


def _cpu_worker_step(params_shard):
    """recon + projection (float32 numpy, as TF would) + reference CPU render for a shard of faces; returns a checksum."""
    import numpy as np
    import oracle
    from oracle import recon
    m = _CPU["model"]
    if len(params_shard) == 0:
        return 0.0
    vp = recon.vertices_transform(params_shard, m, IM_SIZE, dtype=np.float32).astype(np.float32)
    B = vp.shape[0]
    if _CPU["use_ref"]:
        tex = np.broadcast_to(m["vertex"], (B,) + m["vertex"].shape)
        depth = oracle.ref_render_depth(vp, m["tri"], np.ascontiguousarray(tex), (B, H, W, 3))[0]
    else:
        depth = oracle.oracle_render_depth_forward(vp, m["tri"], m["vertex"], H, W)[0]
    return float(depth[depth > -1e13].sum())


class CpuArm:
    """Process pool (the reference op is non-reentrant: static scratch, render_depth_op.cc:125-131 => processes)."""

    def __init__(self, model, cores=None):
        import multiprocessing as mp
        import oracle
        self.use_ref = oracle.ref_available()
        self.cores = cores or max(1, len(os.sched_getaffinity(0)))
        ctx = mp.get_context("fork")
        self.pool = ctx.Pool(self.cores, initializer=_cpu_worker_init, initargs=(model, self.use_ref))

    def step(self, params):
        import numpy as np
        shards = [s for s in np.array_split(params, self.cores) if len(s)]
        return sum(self.pool.map(_cpu_worker_step, shards, chunksize=1))

    def time_steps(self, params, steps, warmup):
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


def _traffic(kernel_group):
    """dram bytes per launch from the committed ncu --set full capture (profiles/roofline_traffic.json), or None."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        return json.load(open(path)).get(kernel_group)
    except Exception:
        return None


def algorithmic_bytes(B, nver, ntri, K):
    """compulsory HBM bytes per launch of B faces, split into the two kernel groups."""
    n3 = 3 * nver
    recon = 4 * (n3 * K + n3 + B * (7 + K) + B * n3)                 # basis + mean + params read, vertex_proj written
    render = 4 * (3 * ntri + B * n3 + 2 * B * H * W)                 # tri + vertex_proj read, depth + tri_ind written
    return recon, render


def run_ours(args):
    import numpy as np
    import torch
    pkg = importlib.import_module(PKG)
    synth = importlib.import_module(PKG + ".synth")
    dist = importlib.import_module(PKG + ".distributed")
    rank, local_rank, world = dist.init_from_env()
    if world != args.gpus and rank == 0:
        print("note: WORLD_SIZE=%d but --gpus %d; using WORLD_SIZE" % (world, args.gpus), file=sys.stderr)
    B = args.batch
    model = synth.make_synthetic_model(seed=0, jitter=0.2)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = CpuArm(model)                                            # fork the pool before CUDA is initialised

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = pkg._lib.lib()
    check = pkg._lib.check
    dm = pkg.DeviceModel(model, dev)
    nver, ntri, ks, ke = dm.nver, dm.ntri, dm.ndim_shape, dm.ndim_exp
    K = ks + ke
    params_all = synth.sample_params_constrained(B * world, seed=2)
    params_host = params_all[rank * B:(rank + 1) * B]
    params = torch.from_numpy(params_host).to(dev)
    vertex = torch.empty((B, 3, nver), dtype=torch.float32, device=dev)
    depth = torch.empty((B, H, W, 1), dtype=torch.float32, device=dev)
    tri_ind = torch.empty((B, H, W, 1), dtype=torch.float32, device=dev)
    rbytes = lib.fr_recon_workspace_bytes(B, nver, ks, ke)
    mesh = dm.mesh.handle
    ws = torch.empty(max(lib.fr_pipeline_workspace_bytes(B, nver, ks, ke, H, W),
                         rbytes + lib.fr_render_workspace_bytes(B, nver, H, W)), dtype=torch.uint8, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream

    def step_full(vertex_ptr=None, events=None):
        """The north-star call: params -> depth + tri_ind.  The intermediate vertex tensor is an optional output of the fused
        call (the reconstruction epilogue hands the vertices to its rasterizer stage in shared memory); the timed loop does
        not ask for it."""
        check(lib.fr_recon_render_forward(params.data_ptr(), dm.packed.data_ptr(), dm.tri.data_ptr(), mesh, vertex_ptr,
                                          depth.data_ptr(), tri_ind.data_ptr(), B, nver, ntri, ks, ke, H, W, IM_SIZE,
                                          dm.run_flags, ws.data_ptr(), ws.numel(), sp, events))

    def step_recon():
        check(lib.fr_recon_project_forward(params.data_ptr(), dm.packed.data_ptr(), mesh, vertex.data_ptr(), B, nver, ks, ke, IM_SIZE,
                                           dm.run_flags, ws.data_ptr(), rbytes, sp))

    def step_render():
        check(lib.fr_render_depth_forward(vertex.data_ptr(), dm.tri.data_ptr(), None, 0, depth.data_ptr(), None, None,
                                          tri_ind.data_ptr(), B, nver, ntri, H, W, mesh, ws.data_ptr() + rbytes, ws.numel() - rbytes, sp))

    def timed_parts(steps, warmup):
        """The fused step with an event recorded by the library between its reconstruction(+rasterizer) kernels and the resolve
        pass (stage_events argument): device time of the two parts of the same real step, L2 flushed before every step."""
        import ctypes
        for _ in range(warmup):
            flush.zero_()
            step_full()
        torch.cuda.synchronize(dev)
        mids = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for (a, b), m in zip(evs, mids):
            m.record(stream)                                           # creates the underlying cudaEvent
            flush.zero_()
            a.record(stream)
            step_full(None, (ctypes.c_void_p * 2)(m.cuda_event, None))
            b.record(stream)
        torch.cuda.synchronize(dev)
        recon = sum(a.elapsed_time(m) for (a, _), m in zip(evs, mids)) / steps
        render = sum(m.elapsed_time(b) for (_, b), m in zip(evs, mids)) / steps
        return recon, render

    def timed(fn, steps, warmup):
        """Per-step CUDA-event timing on the launching stream, L2 flushed (outside the events) before every step."""
        for _ in range(warmup):
            flush.zero_()
            fn()
        dist.barrier()
        torch.cuda.synchronize(dev)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in evs:
            flush.zero_()
            a.record(stream)
            fn()
            b.record(stream)
        torch.cuda.synchronize(dev)
        dist.barrier()
        return sum(a.elapsed_time(b) for a, b in evs) / steps          # ms per step

    sampler = _clock_sampler_start(local_rank) if rank == 0 else (None, None)
    time.sleep(0.3 if rank == 0 else 0.0)                             # let nvidia-smi come up
    t_begin = time.time()
    launches0 = lib.fr_launch_count()
    ms_full = timed(step_full, args.steps, args.warmup)
    launches = (lib.fr_launch_count() - launches0)
    launches_timed = launches * args.steps // (args.steps + args.warmup)
    ms_full_vertex = timed(lambda: step_full(vertex.data_ptr()), args.steps, args.warmup)   # same call, vertex_proj materialised too
    ms_part_recon, ms_part_render = timed_parts(args.steps, args.warmup)
    ms_recon = timed(step_recon, args.steps, args.warmup)
    ms_render = timed(step_render, args.steps, args.warmup)
    # the timed loops last only milliseconds: keep the same step running so the 50 ms clock sampler sees it under load
    t_soak = time.time()
    while time.time() - t_soak < SOAK_SECONDS:
        for _ in range(20):
            step_full()
        torch.cuda.synchronize(dev)
    t_end = time.time()
    clocks = _clock_sampler_stop(*sampler, t_begin, t_end) if rank == 0 else None
    ms_full_max = dist.reduce_scalar(ms_full, "max")
    ms_recon_max = dist.reduce_scalar(ms_recon, "max")
    ms_render_max = dist.reduce_scalar(ms_render, "max")
    ms_full_vertex_max = dist.reduce_scalar(ms_full_vertex, "max")
    ms_part_recon_max = dist.reduce_scalar(ms_part_recon, "max")
    ms_part_render_max = dist.reduce_scalar(ms_part_render, "max")
    faces_total = dist.reduce_scalar(B, "sum")
    launches_total = int(dist.reduce_scalar(launches_timed, "sum"))

    # parity gate that travels with every measurement: a few faces of this very output against the CPU oracle
    parity = None
    if rank == 0 and not args.no_parity:
        import oracle
        from oracle import recon as orecon
        torch.cuda.synchronize(dev)
        depth_timed = depth.clone()
        step_full(vertex.data_ptr())                                      # same call, this time materialising vertex_proj
        torch.cuda.synchronize(dev)
        assert depth_timed.cpu().numpy().tobytes() == depth.cpu().numpy().tobytes()
        vp = vertex[:2].cpu().numpy()
        want_vp = orecon.vertices_transform(params_host[:2], model, IM_SIZE)
        want = oracle.oracle_render_depth_forward(vp, model["tri"], model["vertex"], H, W)
        parity = {"faces_checked": 2,
                  "vertex_rel_err": float(np.abs(vp - want_vp).max() / np.abs(want_vp).max()),
                  "tri_ind_bit_exact": bool(tri_ind[:2].cpu().numpy().tobytes() == want[3].tobytes()),
                  "depth_bit_exact": bool(depth[:2].cpu().numpy().tobytes() == want[0].tobytes())}

    # other BASELINE configs, reported as extras (not the headline): config 4 = batch-256 forward + backward of the whole
    # path (d depth -> d params through render_depth's backward and the recon backward), config 1 = batch-1 forward latency
    extras = None
    if rank == 0 and not args.no_extras:
        net = importlib.import_module(PKG + ".nets.network")
        ops = importlib.import_module(PKG + ".rendering_layer.ops")

        def time_torch(fn, reps):
            for _ in range(3):
                fn()
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(reps):
                fn()
            b.record(stream)
            torch.cuda.synchronize(dev)
            return a.elapsed_time(b) / reps

        B4 = 256
        p4 = torch.from_numpy(synth.sample_params_constrained(B4, seed=4)).to(dev).requires_grad_(True)
        img4 = torch.empty((B4, H, W, 3), device=dev)
        tex4 = dm.vertex_code.unsqueeze(0).expand(B4, -1, -1)
        gd4 = torch.randn((B4, H, W, 1), device=dev)

        def fwd_bwd():
            p4.grad = None
            vp = net.recon_project(p4, dm, IM_SIZE)
            d, _, _, ti = ops.render_depth(vp, dm.tri, tex4, img4)
            (d * (gd4 * (ti >= 0))).sum().backward()

        def fwd_only4():
            with torch.no_grad():
                vp = net.recon_project(p4, dm, IM_SIZE)
                ops.render_depth(vp, dm.tri, tex4, img4)

        ms_fb = time_torch(fwd_bwd, 10)
        ms_f4 = time_torch(fwd_only4, 10)
        p1 = torch.from_numpy(synth.sample_params_constrained(1, seed=1)).to(dev)
        img1 = torch.empty((1, H, W, 3), device=dev)

        def fwd1():
            with torch.no_grad():
                vp = net.recon_project(p1, dm, IM_SIZE)
                ops.render_depth(vp, dm.tri, dm.vertex_code.unsqueeze(0), img1)

        ms_1 = time_torch(fwd1, 20)
        # config 3's per-GPU share at 8 GPUs (512 faces) and the whole 4096-face batch on this one GPU, fused call, depth + tri_ind
        big = {}
        for B3 in (512, 4096):
            p3 = torch.from_numpy(synth.sample_params_constrained(B3, seed=3)).to(dev)
            d3 = torch.empty((B3, H, W, 1), dtype=torch.float32, device=dev)
            t3 = torch.empty((B3, H, W, 1), dtype=torch.float32, device=dev)
            ws3 = torch.empty(lib.fr_pipeline_workspace_bytes(B3, nver, ks, ke, H, W), dtype=torch.uint8, device=dev)

            def fwd3():
                check(lib.fr_recon_render_forward(p3.data_ptr(), dm.packed.data_ptr(), dm.tri.data_ptr(), mesh, None, d3.data_ptr(),
                                                  t3.data_ptr(), B3, nver, ntri, ks, ke, H, W, IM_SIZE, dm.run_flags, ws3.data_ptr(),
                                                  ws3.numel(), sp, None))
            ms3 = time_torch(fwd3, 5)
            big["config3_b%d_fwd" % B3] = {"ms": ms3, "faces_per_s": B3 / (ms3 * 1e-3), "api": "fr_recon_render_forward, one call"}
            del p3, d3, t3, ws3
        extras = {"config4_b256_fwd_bwd": {"ms": ms_fb, "faces_per_s": B4 / (ms_fb * 1e-3), "fwd_only_ms": ms_f4,
                                           "api": "recon_project + render_depth (all four outputs) + autograd backward, torch API, L2 warm"},
                  "config1_b1_fwd": {"ms": ms_1, "api": "recon_project + render_depth (all four outputs), torch API, L2 warm"}}
        extras.update(big)
        del p4, img4, gd4

    # end to end through the host-buffer C-ABI session: every step copies its params in from pinned host memory and its
    # depth maps back to pinned host memory inside the timed region.  The session's two slots are alternated
    # (fr_session_submit / fr_session_wait) so the copy-out of step i overlaps the kernels of step i+1; the synchronous
    # single-call latency (fr_session_forward) is reported next to it.
    del ws, flush
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
    want_depth = depth.cpu().numpy().tobytes()
    e2e_ok = all(t.numpy().tobytes() == want_depth for t in pin_depth)
    sess.close()

    cpu_baseline = None
    if cpu is not None:
        cpu_steps = 3
        sec = cpu.time_steps(params_host, cpu_steps, 1)
        cpu_baseline = {"value": B / sec, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind,
                        "sample": "%d steps of the same %d-face batch: numpy float32 recon+projection (network.py:153-169) + "
                                  "%s, faces split over %d processes" %
                                  (cpu_steps, B, "reference CPU op oracle/_ref" if cpu.use_ref else "oracle C port", cpu.cores)}
        cpu.close()

    if rank == 0:
        peak, peak_src = _peak_hbm()
        rb, nb = algorithmic_bytes(B, nver, ntri, K)
        # dominant part of the timed (fused) step: reconstruction (prep + tcgen05 kernel) vs rasterizer (keys + resolve)
        parts = {"recon_raster_part_of_step (recon_prep_f16 + recon_fwd_f16<raster> kernels)": (rb + nb - 8 * B * H * W, ms_part_recon_max),
                 "render_resolve_part_of_step (raster_resolve kernel)": (8 * B * H * W, ms_part_render_max)}
        dom = max(parts, key=lambda k: parts[k][1])
        dbytes, dms = parts[dom]
        achieved = dbytes / (dms * 1e-3) / 1e9
        other = [k for k in parts if k != dom][0]
        line = {
            "metric": METRIC, "value": faces_total / (ms_full_max * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_full_max, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: batch-64 synthetic 235-d params -> 3DMM recon + pose projection + "
                                   "200x200 depth render (depth + tri_ind), forward only, per GPU",
                       "arithmetic": "f32 results; reconstruction on tcgen05 with fp16 hi/lo operand pairs (22 significant bits) and "
                                     "fp32 accumulation, rasterizer cull in packed integers and inside tests in separately rounded f64 "
                                     "(bit-exact with the reference)",
                       "batch_per_gpu": B, "nver": nver, "ntri": ntri, "ndim_shape": ks, "ndim_exp": ke, "image": [H, W],
                       "l2": "flushed before every timed step (512 MiB write, outside the events)",
                       "call": "fr_recon_render_forward, depth + tri_ind out; the optional vertex_proj output is not requested in "
                               "the timed loop (the parity check re-runs the call with it); groups_ms times the two separate "
                               "entry points, which do materialise and re-read it",
                       "parallelism": "batch-sharded x%d, basis replicated, no collective" % world},
            "e2e": {"value": faces_total / (e2e_ms_max * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms_max,
                    "h2d_bytes_per_step": int(pin_params[0].numel() * 4), "d2h_bytes_per_step": int(pin_depth[0].numel() * 4),
                    "api": "fr_session_submit/fr_session_wait alternating over %d slots (host buffers, pinned; the copy-out "
                           "of step i overlaps the kernels of step i+1)" % nslots,
                    "sync_call_ms": e2e_sync_ms_max, "sync_call_value": faces_total / (e2e_sync_ms_max * 1e-3),
                    "matches_device_path": e2e_ok},
            "gpu_launches": launches_total,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": _traffic("render_part" if dom.startswith("render") else "recon_part"),
                         "peak_source": peak_src, "algorithmic_bytes": dbytes, "ms": dms,
                         "how": "device time between CUDA events of the same real step (the library records one between the "
                                "two parts), algorithmic bytes per SURVEY.md 8(d)",
                         "limiter": "the render part is bound by instruction issue and dependent gathers (profiles/r01_summary.md), "
                                    "the recon part by HBM" if dom.startswith("render") else "HBM (basis stream + vertex records)",
                         "other_part": {"kernel": other, "algorithmic_bytes": parts[other][0], "ms": parts[other][1],
                                        "achieved": parts[other][0] / (parts[other][1] * 1e-3) / 1e9,
                                        "frac": parts[other][0] / (parts[other][1] * 1e-3) / 1e9 / peak,
                                        "traffic": _traffic("recon_part" if dom.startswith("render") else "render_part")},
                         "whole_step": {"algorithmic_bytes": rb + nb, "ms": ms_full_max,
                                        "achieved": (rb + nb) / (ms_full_max * 1e-3) / 1e9,
                                        "frac": (rb + nb) / (ms_full_max * 1e-3) / 1e9 / peak},
                         "fused_call_with_vertex_proj_output_ms": ms_full_vertex_max,
                         "separate_entry_points_ms": {"fr_recon_project_forward": ms_recon_max,
                                                      "fr_render_depth_forward": ms_render_max}},
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
    cpu = CpuArm(model)
    steps, warmup = min(args.steps, 10), min(args.warmup, 2)
    sec = cpu.time_steps(params, steps, max(1, warmup))
    value = B / sec
    sample = ("each step = the full %d-face batch: numpy float32 recon+projection + %s, faces split over %d processes"
              % (B, "reference CPU op (oracle/_ref, render_depth_op.cc compiled in place)" if cpu.use_ref else "oracle C port", cpu.cores))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": max(1, warmup), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: batch-64 synthetic 235-d params -> 3DMM recon + pose projection + "
                                   "200x200 depth render, forward only, host CPU", "batch_per_gpu": B, "image": [H, W]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind, "sample": sample},
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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
