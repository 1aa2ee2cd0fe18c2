"""CPU-side check of CUDA kernel LOGIC without a GPU: the kernel source of csrc/bvh.cu, csrc/bvh.cuh and
csrc/datacosts.cu is compiled unchanged by g++ against tests/cpp/cuda_emul.h (threads run one after the other)
and compared with the oracle.

Why: none of the smooth sphere / terrain scenes produces a single occluded (face, view) pair, so on those the
visibility rays only ever prove "no false hits".  The `occ` scenes (floating plates in front of a displaced
sphere) make ~30 % of the candidates fail the geometric visibility test; this file checks the device LBVH build,
the any-hit traversal and the candidate / ray-bitmap / quality / compaction kernels on them.  It proves nothing
about races or memory ordering -- tests/test_zz_gpu_occlusion.py runs the same scenes on the real device.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mvs-texturing_b200", "csrc")
OUT = os.path.join(ROOT, "tests", "cpp", "_emul")
CUDA_INC = "/usr/local/cuda/include"

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")),
                                reason="CUDA headers not installed")


def _kernel_part(cu_file, host_entry, drop=None):
    """Text of a .cu file up to its first host entry point (the part that holds the kernels), unchanged except for
    the CUB include and, optionally, a host-only span [drop[0], drop[1]) that calls CUB."""
    src = open(os.path.join(CSRC, cu_file)).read()
    head = src.split(host_entry)[0].replace("#include <cub/cub.cuh>", "")
    if drop:
        a, b = head.index(drop[0]), head.index(drop[1])
        head = head[:a] + head[b:]
    return head + "}  // namespace b2\n"


@pytest.fixture(scope="module")
def emul():
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "bvh_kernels.inc"), "w") as f:
        f.write(_kernel_part("bvh.cu", "int build_bvh("))
    with open(os.path.join(OUT, "datacosts_kernels.inc"), "w") as f:
        f.write(_kernel_part("datacosts.cu", "int data_costs_qualities(", ("int cub_exclusive_sum_u64", "namespace {")))
    libs = {}
    for name in ("emul_bvh", "emul_datacosts"):
        so = os.path.join(OUT, name + ".so")
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-w",
                               "-I" + CUDA_INC, "-I" + CSRC, "-I" + os.path.join(ROOT, "oracle"), "-I" + OUT,
                               os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-o", so])
        libs[name] = C.CDLL(so)
    return libs


@pytest.mark.parametrize("name", ["tiny", "occ", "C2s"])
def test_device_bvh_build_and_traversal(emul, orc, scene_mod, name):
    """k_morton / k_hierarchy / k_refit give a well-formed tree (every triangle in exactly one leaf, boxes nest) and
    bvh_occluded() answers every vertex->camera ray like the oracle's BVH (itself checked against brute force)."""
    s = scene_mod.config(name, with_images=False)
    L, OL = emul["emul_bvh"], orc.lib()
    nv, K = s.verts.shape[0], s.num_views
    occ = np.zeros((K, nv), np.uint8)
    rays = np.zeros((K, nv, 8), np.float32)
    stats = np.zeros(4, np.int32)
    rc = L.emul_bvh_trace(orc._p(s.verts), C.c_uint32(nv), orc._p(s.faces), C.c_uint32(s.num_faces),
                          orc._p(np.ascontiguousarray(s.pos)), C.c_uint32(K), orc._p(occ), orc._p(rays), orc._p(stats))
    assert rc == 0, f"malformed tree (code {rc})"
    assert 0 < stats[0] <= 64                                  # traversal stack holds 100 entries
    OL.orc_bvh_build.restype = C.c_void_p
    b = C.c_void_p(OL.orc_bvh_build(orc._p(s.verts), orc._p(s.faces), C.c_uint32(s.num_faces)))
    fl = rays.reshape(-1, 8)
    base = fl.ctypes.data
    ref = np.zeros(len(fl), np.uint8)
    for i in range(len(fl)):
        p = base + 32 * i
        ref[i] = OL.orc_bvh_occluded(b, C.c_void_p(p), C.c_void_p(p + 12), C.c_float(fl[i, 6]), C.c_float(fl[i, 7]))
    OL.orc_bvh_free(b)
    assert ref.sum() > 0 and np.array_equal(ref, occ.ravel())


@pytest.mark.parametrize("name,data_term,vis", [("tiny", 1, True), ("occ", 1, True), ("occ", 0, True), ("occ", 1, False),
                                                ("occ2", 1, True), ("C2s", 1, True)])
def test_device_data_cost_kernels(emul, orc, get_scene, name, data_term, vis):
    """cull -> ray bitmaps -> rays -> quality -> compaction (csrc/datacosts.cu) vs orc_data_costs: identical
    (face, view) set and bit-identical qualities, with real occlusion in the `occ` scenes."""
    s = get_scene(name)
    L = emul["emul_datacosts"]
    views, keep = orc.make_views(s)
    grads = [orc.gradient_magnitude(s.images[k]) for k in range(s.num_views)]
    gp = (C.c_void_p * s.num_views)(*[g.ctypes.data for g in grads])
    F = s.num_faces
    ptr = np.zeros(F + 1, np.uint64)
    vw, ql = C.c_void_p(), C.c_void_p()
    stats = np.zeros(4, np.int64)
    rc = L.emul_data_costs(orc._p(s.verts), C.c_uint32(s.verts.shape[0]), orc._p(s.faces), orc._p(s.face_normals), C.c_uint32(F),
                           views, C.c_uint32(s.num_views), gp, data_term, 1 if vis else 0, orc._p(ptr), C.byref(vw), C.byref(ql),
                           orc._p(stats))
    assert rc == 0
    n = int(ptr[-1])
    view = np.ctypeslib.as_array(C.cast(vw, C.POINTER(C.c_uint16)), (max(n, 1),))[:n].copy()
    qual = np.ctypeslib.as_array(C.cast(ql, C.POINTER(C.c_float)), (max(n, 1),))[:n].copy()
    L.emul_free(vw); L.emul_free(ql)
    o = orc.data_costs(s, data_term=data_term, visibility=vis)
    assert np.array_equal(ptr, o["face_ptr"]) and np.array_equal(view, o["view"])
    assert np.array_equal(qual.view(np.uint32), o["quality"].view(np.uint32))
    if name.startswith("occ") and vis:
        assert stats[2] > 500                                  # occluded (vertex, view) rays
        assert n < stats[0] * 0.9                              # >10 % of the candidates are occluded
    if not vis:
        assert stats[1] == 0
