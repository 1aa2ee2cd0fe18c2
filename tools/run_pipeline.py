"""Dev tool: run the resident pipeline on a named scene and print stage times + per-kernel events."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
b2 = importlib.import_module("mvs-texturing_b200")
scene = importlib.import_module("mvs-texturing_b200.scene")

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
with_patches = "--patches" in sys.argv   # also run texture patches + adjust_colors + local seam leveling
t = time.time(); s = scene.config(name); print(f"scene {name}: F={s.num_faces} K={s.num_views} gen {time.time()-t:.1f}s", flush=True)
t = time.time(); ap, ai = scene.face_adjacency(s.faces); rings = scene.vertex_rings(s.faces, s.verts.shape[0]); print(f"graph {time.time()-t:.1f}s", flush=True)
c = b2.Context(0)
t = time.time(); c.set_scene(s); c.set_adjacency(ap, ai); c.set_vertex_rings(*rings); print(f"upload {time.time()-t:.2f}s", flush=True)
for rep in range(reps):
    c.profile(True)
    t0 = time.time(); info = c.data_costs_run(); t1 = time.time()
    minfo, trace = c.view_selection_run(); t2 = time.time()
    sinfo = c.seam_run(); t3 = time.time()
    print(f"rep {rep}: dc {1e3*(t1-t0):.1f} ms (nnz={info.nnz} cand={info.candidates} rays={info.rays})  "
          f"mrf {1e3*(t2-t1):.1f} ms (it={minfo.iterations} E0={minfo.energy_initial:.1f} E={minfo.energy_final:.1f})  "
          f"seam {1e3*(t3-t2):.1f} ms (R={sinfo.num_rows} nnzL={sinfo.nnz_full} A={sinfo.num_a_rows} it={list(sinfo.iterations)} cg_ms={sinfo.cg_ms:.2f})  "
          f"total {1e3*(t3-t0):.1f} ms -> {s.num_faces/(t3-t0):.0f} faces/s", flush=True)
    if with_patches:
        t4 = time.time(); pinfo = c.texture_patches_run(apply_adjust=True); t5 = time.time()
        linfo = c.local_seam_leveling_run(); t6 = time.time()
        print(f"        patches {1e3*(t5-t4):.1f} ms (n={pinfo.num_patches} px={pinfo.num_pixels})  local seam {1e3*(t6-t5):.1f} ms "
              f"(edges={linfo.num_seam_edges} unknowns={linfo.num_unknowns} it={list(linfo.iterations)} res={[f'{r:.1e}' for r in linfo.residual]})", flush=True)
    rep_ = c.profile_report()
    agg = {}
    for n_, ms, by in rep_:
        a = agg.setdefault(n_, [0, 0.0, 0.0]); a[0] += 1; a[1] += ms; a[2] += by
    for n_, (cnt, ms, by) in agg.items():
        print(f"    {n_:22s} n={cnt:3d} {ms:9.3f} ms  {by/1e6:10.1f} MB  {by/ms/1e6 if ms else 0:8.1f} GB/s")
    c.profile(False)
print("trace", np.round(trace, 1))
