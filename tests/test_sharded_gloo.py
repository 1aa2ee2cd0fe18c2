"""world_size-2 gloo test (CPU) of the host-side logic of the sharded path: label-range all-gather
with a ragged last range, energy all-reduce and the StopWhenReturnsDiminish rule."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, F, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    par = importlib.import_module("mvs-texturing_b200.sharded")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    psz = (F + world - 1) // world
    fb, fe = min(F, rank * psz), min(F, (rank + 1) * psz)
    truth = (np.arange(F) * 7 + 3).astype(np.int32)
    labels = torch.zeros(F, dtype=torch.int32)
    labels[fb:fe] = torch.from_numpy(truth[fb:fe])        # every rank only knows its own range
    mine = torch.zeros(psz, dtype=torch.int32)
    gathered = torch.zeros(world * psz, dtype=torch.int32)
    par.gather_label_ranges(dist, labels, mine, gathered, fb, fe, F)
    ok = bool(np.array_equal(labels.numpy(), truth))
    e = torch.tensor([1000 + rank], dtype=torch.int64)
    dist.all_reduce(e)
    ok = ok and int(e.item()) == sum(1000 + r for r in range(world))
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("1" if ok else "0")
    dist.destroy_process_group()


@pytest.mark.parametrize("F", [10, 11, 257])
def test_label_exchange_two_ranks(tmp_path, F):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() + F) % 2000
    mp.spawn(_worker, args=(2, port, F, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok0").read() == "1" and open(tmp_path / "ok1").read() == "1"


def test_returns_diminish_rule():
    par = importlib.import_module("mvs-texturing_b200.sharded")
    ratio = float(np.float32(0.01))
    e = [int(x * 2**32) for x in [100.0, 90.0, 85.0, 84.0, 83.5, 83.2, 83.1, 83.05]]
    assert not par.returns_diminish(e, 4, 5, ratio)           # t < window
    assert not par.returns_diminish(e, 5, 5, ratio)           # (100-83.2)/100 = 16.8 %
    assert not par.returns_diminish(e, 6, 5, ratio)           # (90-83.1)/90
    e2 = e + [int(83.0 * 2**32)] * 6
    assert par.returns_diminish(e2, 12, 5, ratio)             # < 1 % over five iterations
    assert par.returns_diminish([0, 0, 0, 0, 0, 0], 5, 5, ratio)  # zero energy stops


def test_sharded_e2e_drives_the_context_from_pinned_copies():
    """ShardedPipeline.e2e (the N > 1 `e2e` leg of bench.py): every step re-uploads mesh, images, graph and rings from host
    copies made once, then runs the step and reads labels and adjust values back.  Driven here with a recording context and
    a torch stand-in (no CUDA in this container), so that the host logic of the multi-GPU bench leg is exercised on CPU."""
    import importlib
    import types
    import torch
    par = importlib.import_module("mvs-texturing_b200.sharded")
    scene = importlib.import_module("mvs-texturing_b200.scene")
    s = scene.config("tiny")
    adj = scene.face_adjacency(s.faces)
    rings = scene.vertex_rings(s.faces, s.verts.shape[0])

    class Ctx:
        def __init__(self, *a):
            self.calls = []

        def set_scene(self, sc):
            self.calls.append(("scene", sc.images.shape, sc.verts.shape))

        def set_adjacency(self, a, b):
            self.calls.append(("adj", len(a), len(b)))

        def set_vertex_rings(self, a, b, c, d):
            self.calls.append(("rings", len(a), len(c)))

        def labels_download(self):
            return np.zeros(s.num_faces, np.uint32)

        def seam_download(self, info):
            return {"x": np.zeros((5, 3), np.float32)}

    class B2:
        Context = Ctx

    cuda = types.SimpleNamespace(synchronize=lambda: None, current_device=lambda: 0)
    fake_torch = types.SimpleNamespace(from_numpy=torch.from_numpy, float64=torch.float64, cuda=cuda,
                                       tensor=lambda v, dtype=None, device=None: torch.tensor(v, dtype=dtype))
    p = par.ShardedPipeline(B2, s, adj, rings, 0, 1, 0, upload=False)
    p.step = lambda: {"seam": None}
    r = p.e2e(fake_torch, steps=2, warmup=1)
    assert r["value"] > 0 and r["d2h_bytes_per_step"] == 4 * s.num_faces + 60
    assert r["h2d_bytes_per_step"] == (s.verts.nbytes + s.faces.nbytes + s.face_normals.nbytes + s.images.nbytes + adj[0].nbytes
                                       + adj[1].nbytes + sum(x.nbytes for x in rings))
    assert len(p.ctx.calls) == 9 and p.ctx.calls[0] == ("scene", s.images.shape, s.verts.shape)   # 3 steps x (scene, adjacency, rings)
    assert p.ctx.calls[1] == ("adj", len(adj[0]), len(adj[1])) and p.ctx.calls[2] == ("rings", len(rings[0]), len(rings[2]))
