"""The C++ tex:: veneer (mvs-texturing_b200/tex) compiles, links against libb2tex.so and -- on a GPU --
runs the texrecon hot-path slice; without a GPU it must fail loudly (no CPU fallback)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_texrecon_hotpath")


def _build(b2):
    b2.lib()
    pkg = os.path.join(ROOT, "mvs-texturing_b200")
    cmd = ["/usr/bin/g++", "-std=c++11", "-O2", "-Wall", "-o", EXE, os.path.join(ROOT, "tests", "cpp", "texrecon_hotpath.cpp"),
           os.path.join(pkg, "tex", "texturing.cpp"), "-L" + pkg, "-lb2tex", "-Wl,-rpath," + pkg]
    subprocess.check_call(cmd)


def test_veneer_compiles_links_and_fails_loudly_without_gpu(b2):
    _build(b2)
    r = subprocess.run([EXE, "--link-only"], capture_output=True, text=True)
    assert r.returncode == 0 and "adjacency edges: 6" in r.stdout
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([EXE], capture_output=True, text=True)
        assert r.returncode == 2 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_veneer_runs_hot_path_on_gpu(b2):
    _build(b2)
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    line = [l for l in r.stdout.splitlines() if l.startswith("nnz=")][0]
    labels = [int(x) for x in line.split("labels=")[1].split()]
    assert all(1 <= l <= 4 for l in labels)          # every face of the tetrahedron is seen by some view
