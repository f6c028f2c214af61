#!/usr/bin/env python
"""Developer check (GPU): recon_project forward of the current FR_RECON_PATH against the float64 oracle, per batch size.
    FR_RECON_PATH=f16 python tools/check_recon.py 64 20 70
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
