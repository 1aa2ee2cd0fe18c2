"""torchrun --nproc-per-node P tools/check_sharded.py [scene]: the P-GPU sharded pipeline must give
the labels of the oracle run with num_parts=P (and the same data costs as the single-GPU run)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
b2 = importlib.import_module("mvs-texturing_b200"); scene = importlib.import_module("mvs-texturing_b200.scene")
par = importlib.import_module("mvs-texturing_b200.sharded")
name = sys.argv[1] if len(sys.argv) > 1 else "C1d"
s = scene.config(name); adj = scene.face_adjacency(s.faces); rings = scene.vertex_rings(s.faces, s.verts.shape[0])
p = par.ShardedPipeline(b2, s, adj, rings, rank, world, lr)
res = p.step()
nnz = p.total_nnz() or int(res["dc"].nnz)
res2 = p.step()          # a second pass on the same peers (barrier epochs continue, blocks are reused)
labels = p.ctx.labels_download()
x = p.ctx.seam_download(res2["seam"])["x"]
if rank == 0:
    import oracle as O
    o = O.data_costs(s)
    om = O.view_selection(adj[0], adj[1], o["face_ptr"], o["view"], o["cost"], threads=1, num_parts=world)
    og = O.global_seam_leveling(s, rings, om["labels"])
    ok = (np.array_equal(labels, om["labels"]) and res["mrf"].iterations == om["iterations"] and nnz == len(o["view"])
          and res2["mrf"].iterations == om["iterations"] and abs(res2["mrf"].energy_final - om["energy"]) <= 1e-6 * max(1.0, om["energy"]))
    rel = np.linalg.norm(x - og["x"]) / np.linalg.norm(og["x"])
    print(f"SHARDED world={world} scene={name} labels_equal={np.array_equal(labels, om['labels'])} iters {res['mrf'].iterations}/{om['iterations']} "
          f"nnz {nnz}/{len(o['view'])} cg {list(res2['seam'].iterations)}/{list(og['iterations'])} E={res['mrf'].energy_final:.3f}/{om['energy']:.3f} seam_rel={rel:.2e} -> {'OK' if ok and rel < 5e-3 else 'FAIL'}", flush=True)
dist.destroy_process_group()
