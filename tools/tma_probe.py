"""Diagnostic: the TMA-staged gradient kernel on one small scene, error text and pixel mismatches printed."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
b2 = importlib.import_module("mvs-texturing_b200")
scene = importlib.import_module("mvs-texturing_b200.scene")
par = importlib.import_module("mvs-texturing_b200.sharded")
import oracle as O
name = sys.argv[1] if len(sys.argv) > 1 else "tiny"
s = scene.config(name)
c = b2.Context(0)
c.set_scene(s)
try:
    c.data_costs_run()
except Exception as e:
    print("TMA_PROBE error:", str(e)[:400]); sys.exit(1)
ptr, n = c.device_ptr("grad")
K, H, W = s.num_views, s.height, s.width
g = torch.as_tensor(par._DevArray(ptr, K * H * W, "|u1"), device="cuda").cpu().numpy().reshape(K, H, W)
bad = 0
for v in range(K):
    ref = O.gradient_magnitude(s.images[v])
    d = g[v] != ref
    if d.any():
        ys, xs = np.nonzero(d)
        print(f"TMA_PROBE view {v}: {int(d.sum())} wrong pixels, rows {ys.min()}..{ys.max()} cols {xs.min()}..{xs.max()}; first {ys[0]},{xs[0]} got {g[v][ys[0], xs[0]]} want {ref[ys[0], xs[0]]}")
        bad += int(d.sum())
print("TMA_PROBE", name, f"{W}x{H}x{K}", "mode", os.environ.get("B2TEX_TMA_MODE", "1"), "wrong pixels:", bad)
