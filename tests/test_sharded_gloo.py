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
