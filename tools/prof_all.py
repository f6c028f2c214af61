"""Developer tool: one forward + backward of recon_render_depth (fr_recon_render_forward_all: vertices + all four render_depth
outputs from one call) at batch B on the BFM-sized synthetic model -- a short command line for compute-sanitizer / ncu.
    python tools/prof_all.py [B]"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dfacerecon_b200"); synth = importlib.import_module("3dfacerecon_b200.synth")
net = importlib.import_module("3dfacerecon_b200.nets.network")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
dm = pkg.DeviceModel(synth.make_synthetic_model(seed=0, jitter=0.2), dev)
p = torch.from_numpy(synth.sample_params_constrained(B, seed=2)).to(dev).requires_grad_(True)
v, d, t, n, i = net.recon_render_depth(p, dm, dm.vertex_code, 200, 200, 200)
(d * (i >= 0)).sum().backward()
torch.cuda.synchronize()
print("B=%d covered %d px, |grad| max %.3g" % (B, int((i >= 0).sum()), p.grad.abs().max().item()))
