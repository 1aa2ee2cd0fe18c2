"""CPU-side check of CUDA kernel LOGIC without a GPU: the kernel source of csrc/bvh.cu, csrc/bvh.cuh and
csrc/datacosts.cu is compiled unchanged by g++ against tests/cpp/cuda_emul.h (threads run one after the other)
and compared with the oracle.

Why: none of the smooth sphere / terrain scenes produces a single occluded (face, view) pair, so on those the
visibility rays only ever prove "no false hits".  The `occ` scenes (floating plates in front of a displaced
sphere) make ~30 % of the candidates fail the geometric visibility test; this file checks the device LBVH build,
the any-hit traversal and the candidate / ray-bitmap / quality / compaction kernels on them.  It proves nothing
about races or memory ordering -- tests/test_zz_gpu_occlusion.py runs the same scenes on the real device.

Kernels that synchronise (csrc/mrf.cu: cooperative k_forest with grid.sync, the k_up / k_down dataflow sweeps over
per-node flags; csrc/seam.cu: the persistent cooperative PCG) run on tests/cpp/cuda_fiber.h: one ucontext fiber per
CUDA thread, barriers / shuffles / grid.sync as yield points, deterministic round-robin scheduling.  `occ` adds what
the smooth scenes lack there too: 37 faces no view sees (label 0) and ten separate mesh components.
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


def _kernel_part(cu_file, host_entry, drop=None, remove_line=None, close=1):
    """Text of a .cu file up to its first host entry point (the part that holds the kernels), unchanged except for
    the CUB include and, optionally, a span [drop[0], drop[1]) that cannot be compiled for the host (CUB calls, inline
    PTX) and one line (`extern __shared__`).  `close` = number of namespaces still open at the cut."""
    src = open(os.path.join(CSRC, cu_file)).read()
    head = src.split(host_entry)[0].replace("#include <cub/cub.cuh>", "")
    if drop:
        a, b = head.index(drop[0]), head.index(drop[1])
        head = head[:a] + head[b:]
    if remove_line:
        assert remove_line in head
        head = head.replace(remove_line, "")
    return head + "}  // namespace\n" * close


@pytest.fixture(scope="module")
def emul():
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "bvh_kernels.inc"), "w") as f:
        f.write(_kernel_part("bvh.cu", "int build_bvh("))
    with open(os.path.join(OUT, "datacosts_kernels.inc"), "w") as f:
        f.write(_kernel_part("datacosts.cu", "int data_costs_qualities(", ("int cub_exclusive_sum_u64", "namespace {")))
    with open(os.path.join(OUT, "mrf_kernels.inc"), "w") as f:
        f.write(_kernel_part("mrf.cu", "Mrf make_mrf(b2tex_ctx",
                             ("__device__ __forceinline__ uint32_t ld_acquire", "template <int G>\n__global__ void __launch_bounds__(256) k_init_labels"),
                             "    extern __shared__ uint32_t sm[];  // [rounds+1] level counts\n", close=2))
    with open(os.path.join(OUT, "seam_kernels.inc"), "w") as f:
        f.write(_kernel_part("seam.cu", "int seam_run(b2tex_ctx"))
    libs = {}
    cpp = os.path.join(ROOT, "tests", "cpp")
    for name in ("emul_bvh", "emul_datacosts", "emul_mrf", "emul_seam"):
        so = os.path.join(OUT, name + ".so")
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-w",
                               "-I" + os.path.join(cpp, "emul_include"), "-I" + cpp, "-I" + CUDA_INC, "-I" + CSRC,
                               "-I" + os.path.join(ROOT, "oracle"), "-I" + OUT, os.path.join(cpp, name + ".cpp"), "-o", so])
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


@pytest.mark.parametrize("name,kw", [("tiny", {}), ("occ", {}), ("occ", dict(root_div=0, rounds=200)), ("occ", dict(num_parts=2)),
                                     ("occ", dict(group=32))])
def test_device_view_selection_kernels(emul, orc, scene_mod, get_scene, name, kw):
    """csrc/mrf.cu on fibers vs orc_view_selection: identical forest levels in iteration 1, identical iteration count,
    identical labels (=> identical energy).  `occ`: unseen faces (label 0, excluded from the graph), ten components,
    single-root mode, two partitions, 32 lanes per node instead of alloc_mrf's choice."""
    s = get_scene(name)
    adj = scene_mod.face_adjacency(s.faces)
    dc = orc.data_costs(s)
    P = dict(orc.DEFAULT_MRF)
    okw = {k: v for k, v in kw.items() if k != "group"}
    P.update(okw)
    o = orc.view_selection(adj[0], adj[1], dc["face_ptr"], dc["view"], dc["cost"], threads=1, **okw)
    F = s.num_faces
    params = np.array([P["max_iterations"], P["rounds"], P["root_div"], P["seed"], P["window"], P["num_parts"], kw.get("group", 0), 2, 2],
                      np.uint32)
    labels = np.zeros(F, np.uint32)
    trace = np.full(P["max_iterations"] + 1, np.nan)
    lvl = np.zeros(F, np.uint32)
    it = emul["emul_mrf"].emul_view_selection(C.c_uint32(F), C.c_uint32(s.num_views), orc._p(adj[0]), orc._p(adj[1]), orc._p(dc["face_ptr"]),
                                              orc._p(dc["view"]), orc._p(dc["cost"]), orc._p(params), C.c_float(P["ratio"]), orc._p(labels),
                                              orc._p(trace), orc._p(lvl))
    assert it >= 0, "a kernel launch did not terminate (protocol hang)" if it == -1 else "bad parameters"
    assert np.array_equal(lvl, orc.mrf_sample_forest(adj[0], adj[1], dc["face_ptr"], 1, **okw))
    assert it == o["iterations"]
    assert np.array_equal(labels, o["labels"])
    assert abs(trace[it] - o["energy"]) <= 1e-6 * max(1.0, o["energy"])
    if name == "occ":
        assert (o["labels"] == 0).sum() > 10


@pytest.mark.parametrize("name", ["tiny", "occ"])
def test_device_seam_leveling_kernels(emul, orc, scene_mod, get_scene, name):
    """csrc/seam.cu on fibers vs orc_global_seam_leveling: identical unknown numbering, identical Laplacian, bit-identical
    right-hand side, same CG iteration counts, solution within 1e-4 relative (the reductions are ordered differently)."""
    import scipy.sparse as sp
    s = get_scene(name)
    adj = scene_mod.face_adjacency(s.faces)
    rings = scene_mod.vertex_rings(s.faces, s.verts.shape[0])
    dc = orc.data_costs(s)
    labels = orc.view_selection(adj[0], adj[1], dc["face_ptr"], dc["view"], dc["cost"], threads=1)["labels"]
    o = orc.global_seam_leveling(s, rings, labels)
    views, keep = orc.make_views(s)
    Vn = s.verts.shape[0]
    row_ptr = np.zeros(Vn + 1, np.uint32)
    sizes = np.zeros(3, np.uint32)
    status = np.zeros(8, np.uint32)
    ptrs = [C.c_void_p() for _ in range(6)]
    L = emul["emul_seam"]
    rc = L.emul_seam(orc._p(s.verts), C.c_uint32(Vn), orc._p(s.faces), C.c_uint32(s.num_faces), orc._p(rings[0]), orc._p(rings[1]),
                     orc._p(rings[2]), orc._p(rings[3]), orc._p(np.ascontiguousarray(labels, np.uint32)), views, C.c_uint32(s.num_views),
                     orc._p(row_ptr), *[C.byref(p) for p in ptrs], orc._p(sizes), orc._p(status))
    assert rc == 0
    R, A, nnz = (int(v) for v in sizes)
    def grab(p, ct, n):
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), (max(n, 1),))[:n].copy()
        L.emul_seam_free(p)
        return a
    row_label, cp, cc, cv = grab(ptrs[0], C.c_uint32, R), grab(ptrs[1], C.c_uint32, R + 1), grab(ptrs[2], C.c_uint32, nnz), grab(ptrs[3], C.c_float, nnz)
    rhs, x = grab(ptrs[4], C.c_float, 3 * R).reshape(R, 3), grab(ptrs[5], C.c_float, 3 * R).reshape(R, 3)
    assert np.array_equal(row_ptr, o["row_ptr"]) and np.array_equal(row_label, o["row_label"]) and A == o["num_a_rows"]
    G = sp.csr_matrix((cv, cc, cp), shape=(R, R)); G.sum_duplicates()
    ocp, occ_, ocv = o["csr"]
    Om = sp.csr_matrix((ocv, occ_, ocp), shape=(R, R)); Om.sum_duplicates()
    assert (G != Om).nnz == 0
    assert np.array_equal(rhs.view(np.uint32), o["rhs"].view(np.uint32))
    assert status[:3].tolist() == list(o["iterations"])
    assert np.linalg.norm(x - o["x"]) / np.linalg.norm(o["x"]) < 1e-4
