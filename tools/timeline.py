"""Developer tool: start-up timeline of the fused forward kernel (library built with -DFR_TIMELINE, FR_LIB_PATH pointing at it).
Per-CTA clock64 stamps, printed in microseconds after the CTA's entry (1.965 GHz assumed).   python tools/timeline.py [B] [seed] [tag]"""
import ctypes, importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dfacerecon_b200"); synth = importlib.import_module("3dfacerecon_b200.synth")
lib, check = pkg._lib.lib(), pkg._lib.check
raw = ctypes.CDLL(pkg._lib.LIB_PATH)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
SEED = int(sys.argv[2]) if len(sys.argv) > 2 else 2
TAG = sys.argv[3] if len(sys.argv) > 3 else ""
dev = torch.device("cuda:0")
H = W = 200
model = synth.make_synthetic_model(seed=0, jitter=0.2)
dm = pkg.DeviceModel(model, dev, cluster_tiles=True)
nver, ntri, ks, ke = dm.nver, dm.ntri, dm.ndim_shape, dm.ndim_exp
p = torch.from_numpy(synth.sample_params_constrained(B, seed=SEED)).to(dev)
ws = torch.empty(lib.fr_pipeline_workspace_bytes(B, nver, ks, ke, H, W), dtype=torch.uint8, device=dev)
depth, tri_ind = torch.empty((B, H, W, 1), device=dev), torch.empty((B, H, W, 1), device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
sp = torch.cuda.current_stream().cuda_stream
names = {1: "setup done", 2: "inputs staged", 3: "operands built", 4: "prep grid complete (pdl_wait)", 5: "MMA: operands ready", 6: "MMA: first stage landed",
         10: "MMA: tile 0 committed", 11: "MMA: tile 1 committed", 12: "MMA: tile 2 committed", 13: "MMA: tile 3 committed",
         8: "epilogue: first accumulators", 9: "thread 0 leaves the loop", 14: "CTA done"}
for rep in range(4):
    if hasattr(raw, "fr_debug_marks"):
        torch.cuda.synchronize()
        raw.fr_debug_marks(None, 1)
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    check(lib.fr_recon_render_forward(p.data_ptr(), dm.packed.data_ptr(), dm.tri.data_ptr(), dm.mesh.handle, None, depth.data_ptr(),
                                      tri_ind.data_ptr(), B, nver, ntri, ks, ke, H, W, 200.0, dm.run_flags, ws.data_ptr(), ws.numel(), sp, None))
    b.record()
    torch.cuda.synchronize()
    print("call %d: %.1f us" % (rep, a.elapsed_time(b) * 1e3))
if hasattr(raw, "fr_debug_marks"):
    m = np.zeros(8, np.uint64)
    assert raw.fr_debug_marks(ctypes.c_void_p(m.ctypes.data), 0) == 0
    m = m.astype(np.int64)
    t0 = m[0]
    for i, name in ((0, "prep first start"), (1, "prep last key-clearing block done"), (2, "forward first CTA entry"), (3, "forward last CTA done"),
                    (6, "resolve first CTA entry"), (4, "resolve first CTA past its dependency wait"), (5, "resolve last block done")):
        print("  mark %-44s %7.2f us" % (name, (m[i] - t0) / 1e3))
tl = np.zeros((160, 16), np.int64)
assert raw.fr_debug_timeline(ctypes.c_void_p(tl.ctypes.data)) == 0
g0 = tl[:148, 15].min()
print("CTA entry spread (globaltimer): %.2f us" % ((tl[:148, 15].max() - g0) / 1e3))
for slot in sorted(names, key=lambda s: (s not in (1, 2, 3, 4, 5, 6), [1, 2, 3, 4, 5, 6, 10, 8, 11, 12, 13, 9, 14].index(s))):
    d = (tl[:148, slot] - tl[:148, 0]) / 1965.0
    d = d[tl[:148, slot] != 0]
    if d.size:
        print("  %-34s min %7.2f  median %7.2f  max %7.2f us   (CTA 0: %.2f)" % (names[slot], d.min(), np.median(d), d.max(), (tl[0, slot] - tl[0, 0]) / 1965.0))
if hasattr(raw, "fr_debug_cluster_cost"):
    cost = np.zeros(1024, np.float32)
    assert raw.fr_debug_cluster_cost(ctypes.c_void_p(cost.ctypes.data)) == 0
    pm = dm.mesh.parsed()
    cv = pm["cluster_vert"]
    mu = np.asarray(model["mu"], np.float32).reshape(3, -1)
    cen = np.zeros((pm["nclusters"], 3), np.float32)
    ntri_c = np.diff(pm["tri_begin"])
    for c in range(pm["nclusters"]):
        ids = cv[c][cv[c] >= 0] & 0x3FFFFFFF
        ids = ids[ids < mu.shape[1]]
        cen[c] = mu[:, ids].mean(axis=1)
    cc = cost[:pm["nclusters"]] / 1965.0
    print("cluster cost per octet (us): min %.2f median %.2f mean %.2f max %.2f" % (cc.min(), np.median(cc), cc.mean(), cc.max()))
    np.savez(os.path.join(ROOT, "gpurun_out", "cluster_cost_b%d_s%d%s.npz" % (B, SEED, TAG)), cost_us=cc, centroid=cen, ntri=ntri_c, done=tl[:148, 14] - tl[:148, 0],
             params=p.cpu().numpy())
