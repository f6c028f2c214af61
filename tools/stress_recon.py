"""Developer tool: run-to-run determinism of the reconstruction forward (and the fused call): the same call N times on one
model, every output compared bit for bit with the first.   python tools/stress_recon.py [N] [fresh]   (fresh: new workspace +
output tensors for every call, so that stale contents differ)"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dfacerecon_b200"); synth = importlib.import_module("3dfacerecon_b200.synth")
net = importlib.import_module("3dfacerecon_b200.nets.network")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 500
fresh = len(sys.argv) > 2
dev = torch.device("cuda:0")
model = synth.make_synthetic_model(seed=0, jitter=0.2)
for tiles in (False, True):
    dm = pkg.DeviceModel(model, dev, cluster_tiles=tiles)
    for B, full in ((16, True), (20, False), (70, True)):
        p = torch.from_numpy(synth.sample_params_constrained(B, seed=2 + B, full_range=full)).to(dev)
        ref = net.recon_project(p, dm, 200).clone()
        bad = 0
        junk = []
        for i in range(N):
            if fresh:
                net._workspaces.clear() if hasattr(net, "_workspaces") else None
                junk.append(torch.full((1 << 20,), float(i), device=dev))      # perturb the allocator / stale contents
                if len(junk) > 4: junk.pop(0)
            out = net.recon_project(p, dm, 200)
            if not torch.equal(out, ref):
                bad += 1
                if bad <= 3:
                    d = (out - ref).abs()
                    print("  mismatch run %d: max abs %.3g at faces %s" % (i, d.max().item(), torch.nonzero(d.amax(dim=(1, 2)) > 0).flatten().tolist()[:8]))
        torch.cuda.synchronize()
        print("tiles=%s B=%d full=%s: %d / %d runs differ from the first" % (tiles, B, full, bad, N))
