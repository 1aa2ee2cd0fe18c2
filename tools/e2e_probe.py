import importlib, os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
b2 = importlib.import_module("mvs-texturing_b200"); scene = importlib.import_module("mvs-texturing_b200.scene")
s = scene.config("C3"); ap, ai = scene.face_adjacency(s.faces); rings = scene.vertex_rings(s.faces, s.verts.shape[0])
cap = s.num_faces*64
import numpy as np
ot=[torch.empty(s.num_faces+1,dtype=torch.int64).pin_memory(), torch.empty(cap,dtype=torch.int16).pin_memory(), torch.empty(cap,dtype=torch.float32).pin_memory()]
out=(ot[0].numpy().view(np.uint64), ot[1].numpy().view(np.uint16), ot[2].numpy())
imgs = torch.from_numpy(s.images).pin_memory(); s.images = imgs.numpy()
for rep in range(2):
    t0=time.perf_counter(); dc = b2.calculate_data_costs(s, out=out); t1=time.perf_counter()
    labels, mi = b2.view_selection(dc, ap, ai); t2=time.perf_counter()
    g = b2.global_seam_leveling(s, rings, labels); t3=time.perf_counter()
    print(f"e2e rep {rep}: dc {1e3*(t1-t0):.0f} ms  vs {1e3*(t2-t1):.0f} ms  seam {1e3*(t3-t2):.0f} ms  total {1e3*(t3-t0):.0f}", flush=True)
