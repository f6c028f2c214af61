"""Developer tool: BASELINE configs[3] (batch-256 forward + backward through the torch API, all four render outputs) a few
times -- a short command line for an ncu launch list.   python tools/prof_config4.py [B] [reps]"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dfacerecon_b200"); synth = importlib.import_module("3dfacerecon_b200.synth")
net = importlib.import_module("3dfacerecon_b200.nets.network"); ops = importlib.import_module("3dfacerecon_b200.rendering_layer.ops")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
model = synth.make_synthetic_model(seed=0, jitter=0.2)
dm = pkg.DeviceModel(model, dev)
p = torch.from_numpy(synth.sample_params_constrained(B, seed=4)).to(dev).requires_grad_(True)
img = torch.empty((B, 200, 200, 3), device=dev)
tex = dm.vertex_code.unsqueeze(0).expand(B, -1, -1)
gd = torch.randn((B, 200, 200, 1), device=dev)


def fwd_bwd():
    p.grad = None
    vp = net.recon_project(p, dm, 200.0)
    d, _, _, ti = ops.render_depth(vp, dm.tri, tex, img)
    (d * (gd * (ti >= 0))).sum().backward()


for i in range(reps + 2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fwd_bwd(); b.record(); torch.cuda.synchronize()
    if i >= 2:
        print("fwd+bwd %.1f us" % (a.elapsed_time(b) * 1e3))
