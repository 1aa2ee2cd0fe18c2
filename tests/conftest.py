import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def b2():
    """The product package (directory name has a hyphen, so importlib)."""
    return importlib.import_module("mvs-texturing_b200")


@pytest.fixture(scope="session")
def scene_mod():
    return importlib.import_module("mvs-texturing_b200.scene")


@pytest.fixture(scope="session")
def orc():
    import oracle as O  # oracle/oracle.py -- the checker
    O.lib()
    return O


_cache = {}


@pytest.fixture(scope="session")
def get_scene(scene_mod):
    def _get(name):
        if name not in _cache:
            _cache[name] = scene_mod.config(name)
        return _cache[name]
    return _get


_oracle_cache = {}


@pytest.fixture(scope="session")
def oracle_pipeline(orc, scene_mod, get_scene):
    """Oracle results for a named scene: data costs, adjacency, labels, seam system (cached)."""
    def _get(name, stages=("dc", "mrf", "seam")):
        key = name
        r = _oracle_cache.setdefault(key, {})
        s = get_scene(name)
        if "dc" in stages and "dc" not in r:
            r["dc"] = orc.data_costs(s)
        if ("mrf" in stages or "seam" in stages) and "adj" not in r:
            r["adj"] = scene_mod.face_adjacency(s.faces)
        if "mrf" in stages and "mrf" not in r:
            dc = r["dc"]
            r["mrf"] = orc.view_selection(r["adj"][0], r["adj"][1], dc["face_ptr"], dc["view"], dc["cost"], threads=1)
        if "seam" in stages and "seam" not in r:
            r["rings"] = scene_mod.vertex_rings(s.faces, s.verts.shape[0])
            r["seam"] = orc.global_seam_leveling(s, r["rings"], r["mrf"]["labels"])
        return r
    return _get
