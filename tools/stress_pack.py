"""Developer tool: is the model build (mesh table + fr_pack_basis) reproducible?  Builds the DeviceModel N times and compares the
packed basis and one reconstruction output bit for bit with the first build.   python tools/stress_pack.py [N]"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dfacerecon_b200"); synth = importlib.import_module("3dfacerecon_b200.synth")
net = importlib.import_module("3dfacerecon_b200.nets.network")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda:0")
model = synth.make_synthetic_model(seed=0, jitter=0.2)
p = torch.from_numpy(synth.sample_params_constrained(16, seed=18, full_range=True)).to(dev)
ref_packed = ref_out = None
bad_p = bad_o = 0
for i in range(N):
    dm = pkg.DeviceModel(model, dev, cluster_tiles=False)
    out = net.recon_project(p, dm, 200)
    torch.cuda.synchronize()
    if ref_packed is None:
        ref_packed, ref_out = dm.packed.clone(), out.clone()
        continue
    same_p = torch.equal(dm.packed.view(torch.int32), ref_packed.view(torch.int32))
    same_o = torch.equal(out, ref_out)
    if not same_p:
        bad_p += 1
        diff = torch.nonzero(dm.packed.view(torch.int32) != ref_packed.view(torch.int32)).flatten()
        print("build %d: packed differs in %d words, first at float index %d" % (i, diff.numel(), int(diff[0])))
    if not same_o:
        bad_o += 1
        d = (out - ref_out).abs()
        print("build %d: output differs, max abs %.3g, faces %s" % (i, d.max().item(), torch.nonzero(d.amax(dim=(1, 2)) > 0).flatten().tolist()))
    del dm
print("%d builds: packed differs %d times, output differs %d times" % (N, bad_p, bad_o))
