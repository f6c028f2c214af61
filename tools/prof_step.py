"""Developer tool: runs the fused params -> depth-map call and the two stand-alone entry points a few times at batch B on the
BFM-sized synthetic model -- a short command line for `ncu` captures (bench.py does much more).  Prints CUDA-event times.

    python tools/prof_step.py [B] [reps] [grid|permute] [tiles]
"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dfacerecon_b200")
synth = importlib.import_module("3dfacerecon_b200.synth")

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
permute = len(sys.argv) > 3 and sys.argv[3] == "permute"
tiles = len(sys.argv) > 4 and sys.argv[4] == "tiles"          # FR_CLUSTER_TILES flavour (rasterizer inside the reconstruction epilogue)
H = W = 200
dev = torch.device("cuda:0")
lib, check = pkg._lib.lib(), pkg._lib.check
model = synth.make_synthetic_model(seed=0, jitter=0.2, permute=permute)
dm = pkg.DeviceModel(model, dev, cluster_tiles=tiles)
print("B", B, "permute", permute, "cluster_tiles", tiles, "clusters", dm.mesh.nclusters, "vertex slots", dm.mesh.vertex_slots)
params = torch.from_numpy(synth.sample_params_constrained(B, seed=2)).to(dev)
nver, ntri, ks, ke = dm.nver, dm.ntri, dm.ndim_shape, dm.ndim_exp
mesh = dm.mesh.handle
rb = lib.fr_recon_workspace_bytes(B, nver, ks, ke)
ws = torch.empty(max(lib.fr_pipeline_workspace_bytes(B, nver, ks, ke, H, W), rb + lib.fr_render_workspace_bytes(B, nver, H, W)),
                 dtype=torch.uint8, device=dev)
vertex = torch.empty((B, 3, nver), device=dev)
depth, tri_ind = torch.empty((B, H, W, 1), device=dev), torch.empty((B, H, W, 1), device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
sp = torch.cuda.current_stream().cuda_stream


def fused():
    check(lib.fr_recon_render_forward(params.data_ptr(), dm.packed.data_ptr(), dm.tri.data_ptr(), mesh, None, depth.data_ptr(),
                                      tri_ind.data_ptr(), B, nver, ntri, ks, ke, H, W, 200.0, dm.run_flags, ws.data_ptr(), ws.numel(), sp, None))


def recon():
    check(lib.fr_recon_project_forward(params.data_ptr(), dm.packed.data_ptr(), mesh, vertex.data_ptr(), B, nver, ks, ke, 200.0,
                                       dm.run_flags, ws.data_ptr(), rb, sp))


def render():
    check(lib.fr_render_depth_forward(vertex.data_ptr(), dm.tri.data_ptr(), None, 0, depth.data_ptr(), None, None, tri_ind.data_ptr(), B,
                                      nver, ntri, H, W, mesh, ws.data_ptr() + rb, ws.numel() - rb, sp))


for name, fn in (("fused", fused), ("recon", recon), ("render", render)):
    ms = []
    for i in range(reps + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            ms.append(a.elapsed_time(b))
    print("%-7s %.1f us (min %.1f)" % (name, 1e3 * sum(ms) / len(ms), 1e3 * min(ms)))
