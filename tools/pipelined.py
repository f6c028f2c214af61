"""Developer tool: throughput of back-to-back fused calls on S streams (own workspace / outputs per stream), no L2 flush
(the basis stream, 184 MB per step, is larger than L2).   python tools/pipelined.py [B] [K]"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dfacerecon_b200"); synth = importlib.import_module("3dfacerecon_b200.synth")
lib, check = pkg._lib.lib(), pkg._lib.check
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = int(sys.argv[2]) if len(sys.argv) > 2 else 300
dev = torch.device("cuda:0")
H = W = 200
model = synth.make_synthetic_model(seed=0, jitter=0.2)
dm = pkg.DeviceModel(model, dev, cluster_tiles=True)
nver, ntri, ks, ke = dm.nver, dm.ntri, dm.ndim_shape, dm.ndim_exp
nbytes = lib.fr_pipeline_workspace_bytes(B, nver, ks, ke, H, W)
for S in (1, 2, 3, 4):
    streams = [torch.cuda.Stream(dev) for _ in range(S)]
    ps = [torch.from_numpy(synth.sample_params_constrained(B, seed=2 + i)).to(dev) for i in range(S)]
    wss = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(S)]
    outs = [(torch.empty((B, H, W, 1), device=dev), torch.empty((B, H, W, 1), device=dev)) for _ in range(S)]
    def step(i):
        s = i % S
        check(lib.fr_recon_render_forward(ps[s].data_ptr(), dm.packed.data_ptr(), dm.tri.data_ptr(), dm.mesh.handle, None, outs[s][0].data_ptr(),
                                          outs[s][1].data_ptr(), B, nver, ntri, ks, ke, H, W, 200.0, dm.run_flags, wss[s].data_ptr(), nbytes,
                                          streams[s].cuda_stream, None))
    for i in range(20): step(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(torch.cuda.current_stream())
    for s in streams: s.wait_event(a)
    for i in range(K): step(i)
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    b.record(torch.cuda.current_stream())
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / K
    print("streams=%d: %.2f us per step, %.0f faces/s" % (S, us, B / us * 1e6), flush=True)
