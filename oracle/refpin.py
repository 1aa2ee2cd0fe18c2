"""ctypes binding of oracle/_ref/libtexref.so -- TEST INFRASTRUCTURE.

libtexref.so = the reference's own translation units (compiled unmodified from /root/reference/libs/tex)
+ the dependency shims in oracle/refshim/ + oracle/ref_glue.cpp.  It exists to pin the oracle: tests compare
oracle results with the code the oracle restates.  /root/reference is only present in the build container;
on the GPU box the prebuilt .so (git-ignored, not gpurun-ignored) is used as is.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libtexref.so")
REFERENCE = os.environ.get("B2TEX_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.exists(_SO) or os.path.isdir(os.path.join(REFERENCE, "libs", "tex"))


def build() -> str | None:
    """(Re)build when the reference sources are here; otherwise use the prebuilt library if there is one."""
    if os.path.isdir(os.path.join(REFERENCE, "libs", "tex")):
        O.build()
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref", f"REF={REFERENCE}"])
    return _SO if os.path.exists(_SO) else None


_lib = None


def lib():
    global _lib
    if _lib is None:
        so = build()
        if so is None:
            raise RuntimeError("oracle/_ref/libtexref.so missing and no reference checkout to build it from")
        O.lib()
        _lib = C.CDLL(so)
        _lib.ref_tri_area.restype = C.c_float
        _lib.ref_histogram_percentile.restype = C.c_float
    return _lib


_p = O._p


def data_costs(scene, data_term=1, visibility=True, outlier_removal=0, images=None):
    L = lib()
    views, keep = O.make_views(scene, images)
    st = O.Settings(data_term, outlier_removal, 1 if visibility else 0, 0, 0)
    F = scene.num_faces
    face_ptr = np.zeros(F + 1, np.uint64)
    vw, cs = C.c_void_p(), C.c_void_p()
    rc = L.ref_data_costs(_p(scene.verts), C.c_uint32(scene.verts.shape[0]), _p(scene.faces), _p(scene.face_normals),
                          C.c_uint32(F), views, C.c_uint32(scene.num_views), C.byref(st), _p(face_ptr),
                          C.byref(vw), C.byref(cs))
    if rc:
        raise RuntimeError(f"ref_data_costs rc={rc}")
    n = int(face_ptr[-1])
    view = np.ctypeslib.as_array(C.cast(vw, C.POINTER(C.c_uint16)), (max(n, 1),))[:n].copy()
    cost = np.ctypeslib.as_array(C.cast(cs, C.POINTER(C.c_float)), (max(n, 1),))[:n].copy()
    L.ref_free(vw); L.ref_free(cs)
    return dict(face_ptr=face_ptr, view=view, cost=cost)


def _one_view(scene, k, image=None):
    v = O.View()
    v.pos[:] = scene.pos[k].tolist(); v.viewdir[:] = scene.viewdir[k].tolist()
    v.proj[:] = scene.proj[k].tolist(); v.w2c[:] = scene.w2c[k].tolist()
    img = np.ascontiguousarray(scene.images[k] if image is None else image)
    v.height, v.width = img.shape[0], img.shape[1]
    v.rgb = img.ctypes.data
    return v, img


def validity_mask(rgb, erode=False):
    h, w, _ = rgb.shape
    v = O.View()
    rgb = np.ascontiguousarray(rgb)
    v.width, v.height, v.rgb = w, h, rgb.ctypes.data
    m = np.empty((h, w), np.uint8)
    lib().ref_validity_mask(C.byref(v), 1 if erode else 0, _p(m))
    return m


def face_infos(scene, k, tris, data_term=1, outlier_removal=0, image=None):
    v, keep = _one_view(scene, k, image)
    tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
    q = np.empty(len(tris), np.float32)
    mc = np.zeros((len(tris), 3), np.float32)
    lib().ref_face_infos(C.byref(v), data_term, outlier_removal, _p(tris), C.c_uint32(len(tris)), _p(q), _p(mc))
    return q, mc


def pixel_coords(scene, k, x):
    v, keep = _one_view(scene, k)
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty(2, np.float32)
    lib().ref_pixel_coords(C.byref(v), _p(x), _p(out))
    return out


def tri_area(p1, p2, p3):
    a = [np.ascontiguousarray(p, np.float32) for p in (p1, p2, p3)]
    return float(lib().ref_tri_area(_p(a[0]), _p(a[1]), _p(a[2])))


def tri_inside(p1, p2, p3, x, y):
    a = [np.ascontiguousarray(p, np.float32) for p in (p1, p2, p3)]
    return int(lib().ref_tri_inside(_p(a[0]), _p(a[1]), _p(a[2]), C.c_float(x), C.c_float(y)))


def histogram_percentile(values, vmax, bins=10000, p=0.995):
    v = np.ascontiguousarray(values, np.float32)
    return float(lib().ref_histogram_percentile(_p(v), C.c_uint64(len(v)), C.c_float(vmax), bins, C.c_float(p)))


def _grab(ptr, ctype, n):
    a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), (max(n, 1),))[:n].copy()
    lib().ref_free(ptr)
    return a


def build_adjacency(faces, num_verts, rings):
    """tex::build_adjacency_graph on the reference's UniGraph -> CSR in adjacency-list order"""
    vf_ptr, vf_idx, vv_ptr, vv_idx = rings
    F = faces.shape[0]
    adj_ptr = np.zeros(F + 1, np.uint32)
    idx = C.c_void_p()
    rc = lib().ref_build_adjacency(_p(np.ascontiguousarray(faces, np.uint32)), C.c_uint32(F), C.c_uint32(num_verts), _p(vf_ptr), _p(vf_idx),
                                   _p(vv_ptr), _p(vv_idx), _p(adj_ptr), C.byref(idx))
    if rc:
        raise RuntimeError(f"ref_build_adjacency rc={rc}")
    return adj_ptr, _grab(idx, C.c_uint32, int(adj_ptr[-1]))


PARAM_NAMES = ["potts", "window", "ratio", "seed", "deterministic", "tree_algorithm", "use_multilevel", "use_spanning_tree", "use_acyclic",
               "multilevel_after", "force_acyclic", "min_acyclic_iterations", "relax_acyclic_maximal", "components_updated", "compress",
               "model_complete"]


def view_selection_model(adj, face_ptr, view, cost, num_views):
    """tex::view_selection with the recording mapMAP shim: the MRF model it builds + the labels it decodes from the
    shim's trivial solution (cheapest label per node)."""
    adj_ptr, adj_idx = adj
    F = len(face_ptr) - 1
    ne = C.c_uint64()
    edges, ll, lc = C.c_void_p(), C.c_void_p(), C.c_void_p()
    ls_ptr = np.zeros(F + 1, np.uint64)
    labels = np.zeros(F, np.uint32)
    params = np.zeros(16, np.float64)
    rc = lib().ref_view_selection_model(C.c_uint32(F), C.c_uint32(num_views), _p(adj_ptr), _p(adj_idx), _p(face_ptr),
                                        _p(np.ascontiguousarray(view, np.uint16)), _p(np.ascontiguousarray(cost, np.float32)),
                                        C.byref(ne), C.byref(edges), _p(ls_ptr), C.byref(ll), C.byref(lc), _p(labels), _p(params))
    if rc:
        raise RuntimeError(f"ref_view_selection_model rc={rc}")
    n = int(ls_ptr[-1])
    return dict(edges=_grab(edges, C.c_uint32, 2 * int(ne.value)).reshape(-1, 2), ls_ptr=ls_ptr, ls_label=_grab(ll, C.c_int32, n),
                ls_cost=_grab(lc, C.c_float, n), labels=labels, params=dict(zip(PARAM_NAMES, params.tolist())))


class RefPatch:
    pass


def seam_leveling(scene, rings, adj, labels, do_global=True, do_local=False):
    """texrecon.cpp:160-190 on the reference TUs: texture patches (+ vertex projection infos) after
    generate_texture_patches -> global_seam_leveling | zero adjust -> local_seam_leveling."""
    L = lib()
    views, keep = O.make_views(scene)
    vf_ptr, vf_idx, vv_ptr, vv_idx = rings
    n = C.c_uint32()
    labels = np.ascontiguousarray(labels, np.uint32)
    rc = L.ref_seam_leveling(_p(scene.verts), C.c_uint32(scene.verts.shape[0]), _p(scene.faces), C.c_uint32(scene.num_faces),
                             _p(vf_ptr), _p(vf_idx), _p(vv_ptr), _p(vv_idx), _p(adj[0]), _p(adj[1]), _p(labels), views,
                             C.c_uint32(scene.num_views), 1 if do_global else 0, 1 if do_local else 0, C.byref(n))
    if rc:
        raise RuntimeError(f"ref_seam_leveling rc={rc}")
    patches = []
    for pid in range(n.value):
        info = np.zeros(4, np.int32)
        L.ref_patch_info(C.c_uint32(pid), _p(info))
        p = RefPatch()
        p.label, w, h, nf = (int(v) for v in info)
        p.image = np.zeros((h, w, 3), np.float32)
        p.validity = np.zeros((h, w), np.uint8)
        p.blending = np.zeros((h, w), np.uint8)
        faces = np.zeros(nf, np.uint32)
        p.texcoords = np.zeros((3 * nf, 2), np.float32)
        L.ref_patch_data(C.c_uint32(pid), _p(p.image), _p(p.validity), _p(p.blending), _p(faces), _p(p.texcoords))
        p.faces = faces.tolist()
        patches.append(p)
    L.ref_vertex_projection_count.restype = C.c_uint32
    vpi = []
    for v in range(scene.verts.shape[0]):
        k = L.ref_vertex_projection_count(C.c_uint32(v))
        ids = np.zeros(k, np.uint32)
        xy = np.zeros((k, 2), np.float32)
        if k:
            L.ref_vertex_projections(C.c_uint32(v), _p(ids), _p(xy))
        vpi.append({int(i): xy[j].copy() for j, i in enumerate(ids)})
    return patches, vpi
