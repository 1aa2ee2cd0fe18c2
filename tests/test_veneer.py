"""The C++ tex:: veneer (mvs-texturing_b200/tex): the literal call sequence of apps/texrecon/texrecon.cpp:92-189 --
build_adjacency_graph, calculate_data_costs, view_selection, generate_texture_patches, global_seam_leveling (or the
zero-adjust loop), local_seam_leveling, with the reference's signatures (libs/tex/texturing.h:59-106) -- compiles, links
against libb2tex.so and, on a GPU, runs; without a GPU it must fail loudly (no CPU fallback)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_texrecon_hotpath")


def _build(b2):
    b2.lib()
    pkg = os.path.join(ROOT, "mvs-texturing_b200")
    cmd = ["/usr/bin/g++", "-std=c++11", "-O2", "-Wall", "-Werror", "-o", EXE, os.path.join(ROOT, "tests", "cpp", "texrecon_hotpath.cpp"),
           os.path.join(pkg, "tex", "texturing.cpp"), "-L" + pkg, "-lb2tex", "-Wl,-rpath," + pkg]
    subprocess.check_call(cmd)


def test_veneer_compiles_links_and_fails_loudly_without_gpu(b2):
    _build(b2)
    r = subprocess.run([EXE, "--link-only"], capture_output=True, text=True)
    assert r.returncode == 0 and "adjacency edges: 6" in r.stdout
    r = subprocess.run([EXE, "--sphere", "3", "--link-only"], capture_output=True, text=True)
    assert r.returncode == 0 and "adjacency edges: 768" in r.stdout      # 512 faces, closed manifold: 3 F / 2 edges
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([EXE], capture_output=True, text=True)
        assert r.returncode == 2 and "no CPU fallback" in r.stdout


def test_veneer_signatures_match_the_reference_header():
    """Every hot-path function of libs/tex/texturing.h:59-106 is declared in the veneer with the same parameter types."""
    hdr = open(os.path.join(ROOT, "mvs-texturing_b200", "tex", "texturing.h")).read()
    norm = lambda s: re.sub(r"\s+", "", s.replace("const &", "const&").replace(" const", "const"))
    want = {
        "build_adjacency_graph": "(mve::TriangleMesh::ConstPtr mesh, mve::MeshInfo const & mesh_info, UniGraph * graph)",
        "calculate_data_costs": "(mve::TriangleMesh::ConstPtr mesh, TextureViews * texture_views, Settings const & settings, DataCosts * data_costs)",
        "postprocess_face_infos": "(Settings const & settings, FaceProjectionInfos * projected_face_infos, DataCosts * data_costs)",
        "view_selection": "(DataCosts const & data_costs, UniGraph * graph, Settings const & settings)",
        "generate_texture_patches": "(UniGraph const & graph, mve::TriangleMesh::ConstPtr mesh, mve::MeshInfo const & mesh_info, "
                                    "TextureViews * texture_views, Settings const & settings, VertexProjectionInfos * vertex_projection_infos, "
                                    "TexturePatches * texture_patches)",
        "global_seam_leveling": "(UniGraph const & graph, mve::TriangleMesh::ConstPtr mesh, mve::MeshInfo const & mesh_info, "
                                "VertexProjectionInfos const & vertex_projection_infos, TexturePatches * texture_patches)",
        "local_seam_leveling": "(UniGraph const & graph, mve::TriangleMesh::ConstPtr mesh, VertexProjectionInfos const & vertex_projection_infos, "
                               "TexturePatches * texture_patches)",
    }
    flat = norm(hdr)
    for name, params in want.items():
        assert norm("void " + name + params) in flat, name


@pytest.mark.gpu
@pytest.mark.parametrize("args", [[], ["--sphere", "4"], ["--sphere", "4", "--no-global"], ["--sphere", "3", "--labels-from-host", "--no-local"]])
def test_veneer_runs_texrecon_sequence_on_gpu(b2, args):
    _build(b2)
    r = subprocess.run([EXE] + args, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    line = [l for l in r.stdout.splitlines() if l.startswith("patches=")][0]
    f = {k: float(v) for k, v in re.findall(r"(\w+)=([0-9.]+)", line)}
    assert f["patches"] >= 1 and f["valid_pixels"] > 0 and f["vertex_infos"] > 0 and 0.0 < f["mean_red"] < 1.5
    assert "unseen=0" in r.stdout                        # every face of these closed convex meshes is seen by some view
    if "--no-global" not in args:
        assert "adjust values:" in r.stdout
