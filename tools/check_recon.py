#!/usr/bin/env python
"""Developer check (GPU): recon_project forward of the current FR_RECON_PATH against the float64 oracle, per batch size.
    FR_RECON_PATH=tc python tools/check_recon.py 64 20 70
"""
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import recon  # noqa: E402

fr = importlib.import_module("3dfacerecon_b200")
synth = importlib.import_module("3dfacerecon_b200.synth")
net = importlib.import_module("3dfacerecon_b200.nets.network")

small = os.environ.get("FR_CHECK_SMALL") == "1"
model = synth.make_synthetic_model(grid=(29, 41), seed=0) if small else synth.make_synthetic_model(seed=0)
dm = fr.DeviceModel(model, "cuda:0")
print("path override:", os.environ.get("FR_RECON_PATH"), "nver", dm.nver, flush=True)
for B in [int(a) for a in sys.argv[1:]] or [64]:
    p = synth.sample_params_constrained(B, seed=100 + B)
    pt = torch.from_numpy(p).cuda()
    t0 = time.time()
    out = net.recon_project(pt, dm, 200)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    want = recon.vertices_transform(p, model, 200)
    err = np.abs(got - want).max(axis=(1, 2)) / np.abs(want).max(axis=(1, 2))
    print("B=%d  max rel err %.3e  (worst face %d)  nan=%d  %.1f ms" % (B, err.max(), int(err.argmax()), int(np.isnan(got).sum()),
                                                                        1e3 * (time.time() - t0)), flush=True)
    if not (err.max() <= 1e-5):
        bad = np.argwhere(np.abs(got - want) > 1e-5 * np.abs(want).max())
        print("   first bad entries (b, c, n):", bad[:8].tolist(), " n mod 128:", (bad[:8, 2] % 128).tolist())
        print("   got", got[tuple(bad[0])], "want", want[tuple(bad[0])])

if os.environ.get("FR_TC_DEBUG") == "1":
    import ctypes
    lib = fr._lib.lib()
    out = (ctypes.c_ulonglong * 32)()
    lib.fr_debug_tc_counters.argtypes = [ctypes.c_void_p, ctypes.c_int]
    # timing pass: reset, run the bench batch a few times, read
    p = synth.sample_params_constrained(64, seed=2)
    pt = torch.from_numpy(p).cuda()
    for _ in range(3):
        net.recon_project(pt, dm, 200)
    torch.cuda.synchronize()
    lib.fr_debug_tc_counters(None, 1)
    reps = 10
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(reps):
        net.recon_project(pt, dm, 200)
    ev1.record()
    torch.cuda.synchronize()
    lib.fr_debug_tc_counters(out, 0)
    c = np.array(list(out), dtype=np.float64).reshape(4, 8) / reps
    print("avg ms per call (incl. prep + python):", ev0.elapsed_time(ev1) / reps)
    nchunk = max(c[0, 7], 1)
    print("converter warp0 per call: chunks %.0f | cycles/chunk: wait raw_full %.0f, lds+split %.0f, wait a_empty %.0f, sttm+wait::st %.0f, (unused %.0f)"
          % (nchunk, c[0, 0] / nchunk, c[0, 1] / nchunk, c[0, 2] / nchunk, c[0, 3] / nchunk, c[0, 4] / nchunk))
    print("mma thread per call: wait d_empty %.0f, wait a_full total %.0f, issue total %.0f" % (c[1, 0], c[1, 1], c[1, 2]))
    print("epilogue warp8 per call: wait d_full %.0f, drain %.0f" % (c[2, 0], c[2, 1]))

if os.environ.get("FR_TC_DEBUG") in ("2", "3", "4") and hasattr(fr._lib.lib(), "fr_debug_tc_trace"):
    import ctypes
    lib = fr._lib.lib()
    p = synth.sample_params_constrained(64, seed=2)
    pt = torch.from_numpy(p).cuda()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        flush.zero_()
        net.recon_project(pt, dm, 200)
    torch.cuda.synchronize()
    tr = (ctypes.c_uint * (8 * 160))()
    lib.fr_debug_tc_trace.argtypes = [ctypes.c_void_p]
    lib.fr_debug_tc_trace(tr)
    t = np.array(list(tr), dtype=np.int64).reshape(8, 160)
    names = ["producer issue", "conv raw acquired", "conv a_empty acquired", "conv a_full arrived", "mma a_full acquired",
             "mma issue done", "epi d_full acquired", "epi done"]
    np.set_printoptions(linewidth=200)
    for i, nme in enumerate(names):
        n = 48 if i < 4 else (30 if i < 6 else 3)
        print("%-24s" % nme, t[i, :n].tolist())
