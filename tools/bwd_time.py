import importlib, os, sys, torch
sys.path.insert(0, os.getcwd())
pkg = importlib.import_module("3dfacerecon_b200"); synth = importlib.import_module("3dfacerecon_b200.synth")
net = importlib.import_module("3dfacerecon_b200.nets.network"); ops = importlib.import_module("3dfacerecon_b200.rendering_layer.ops")
B=256; dev=torch.device("cuda:0")
model = synth.make_synthetic_model(seed=0, jitter=0.2); dm = pkg.DeviceModel(model, dev)
p = torch.from_numpy(synth.sample_params_constrained(B, seed=4)).to(dev)
vp = net.recon_project(p, dm, 200.0).detach().requires_grad_(True)
img = torch.empty((B,200,200,3), device=dev); tex = dm.vertex_code.unsqueeze(0).expand(B,-1,-1)
d,_,_,ti = ops.render_depth(vp, dm.tri, tex, img)
g = torch.randn((B,200,200,1), device=dev)*(ti>=0)
for i in range(5):
    vp.grad=None
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record(); d.backward(g, retain_graph=True); b.record(); torch.cuda.synchronize()
    print("render backward %.1f us"%(a.elapsed_time(b)*1e3))
