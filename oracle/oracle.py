"""ctypes binding of oracle/_build/liboracle.so -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module (see oracle/oracle.h).  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class View(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("viewdir", C.c_float * 3), ("proj", C.c_float * 9),
                ("w2c", C.c_float * 16), ("width", C.c_int32), ("height", C.c_int32),
                ("rgb", C.c_void_p)]


class Settings(C.Structure):
    _fields_ = [("data_term", C.c_int32), ("outlier_removal", C.c_int32),
                ("geometric_visibility_test", C.c_int32), ("face_begin", C.c_uint32),
                ("face_end", C.c_uint32)]


class DcInfo(C.Structure):
    _fields_ = [("nnz", C.c_uint64), ("max_quality", C.c_float), ("percentile", C.c_float)]


class MrfParams(C.Structure):
    _fields_ = [("max_iterations", C.c_uint32), ("rounds", C.c_uint32), ("root_div", C.c_uint32),
                ("seed", C.c_uint32), ("window", C.c_uint32), ("ratio", C.c_float),
                ("num_parts", C.c_uint32)]


class MrfInfo(C.Structure):
    _fields_ = [("iterations", C.c_uint32), ("energy_initial", C.c_double),
                ("energy_final", C.c_double), ("unseen", C.c_uint64)]


class SeamInfo(C.Structure):
    _fields_ = [("num_rows", C.c_uint32), ("num_a_rows", C.c_uint32), ("num_gamma_rows", C.c_uint32),
                ("nnz_full", C.c_uint64), ("iterations", C.c_uint32 * 3), ("residual", C.c_float * 3)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_tri_area.restype = C.c_float
        _lib.orc_face_quality.restype = C.c_float
        _lib.orc_histogram_percentile.restype = C.c_float
        _lib.orc_mrf_energy.restype = C.c_double
        _lib.orc_mrf_energy_fixed.restype = C.c_int64
        _lib.orc_mrf_brute_force.restype = C.c_double
        _lib.orc_bvh_build.restype = C.c_void_p
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def make_views(scene, images=None):
    """Array of orc_view (keeps the numpy image alive through the returned tuple)."""
    K = scene.num_views
    arr = (View * K)()
    imgs = scene.images if images is None else images
    for k in range(K):
        arr[k].pos[:] = scene.pos[k].tolist()
        arr[k].viewdir[:] = scene.viewdir[k].tolist()
        arr[k].proj[:] = scene.proj[k].tolist()
        arr[k].w2c[:] = scene.w2c[k].tolist()
        arr[k].width = scene.width
        arr[k].height = scene.height
        arr[k].rgb = imgs[k].ctypes.data if imgs.shape[1] else None
    return arr, imgs


DEFAULT_MRF = dict(max_iterations=100, rounds=16, root_div=64, seed=548923723, window=5,
                   ratio=0.01, num_parts=1)


def mrf_params(**kw):
    d = dict(DEFAULT_MRF)
    d.update(kw)
    return MrfParams(**d)


def validity_mask(rgb):
    h, w, _ = rgb.shape
    m = np.empty((h, w), np.uint8)
    lib().orc_validity_mask(_p(np.ascontiguousarray(rgb)), w, h, _p(m))
    return m


def erode(mask):
    m = np.ascontiguousarray(mask.copy())
    lib().orc_erode_validity_mask(_p(m), m.shape[1], m.shape[0])
    return m


def gradient_magnitude(rgb):
    h, w, _ = rgb.shape
    g = np.empty((h, w), np.uint8)
    lib().orc_gradient_magnitude(_p(np.ascontiguousarray(rgb)), w, h, _p(g))
    return g


def data_costs(scene, data_term=1, visibility=True, threads=0, images=None, face_range=(0, 0), outlier_removal=0):
    L = lib()
    views, keep = make_views(scene, images)
    st = Settings(data_term, outlier_removal, 1 if visibility else 0, face_range[0], face_range[1])
    F = scene.num_faces
    face_ptr = np.zeros(F + 1, np.uint64)
    vw, cs, ql = C.c_void_p(), C.c_void_p(), C.c_void_p()
    info = DcInfo()
    rc = L.orc_data_costs(_p(scene.verts), C.c_uint32(scene.verts.shape[0]), _p(scene.faces),
                          _p(scene.face_normals), C.c_uint32(F), views, C.c_uint32(scene.num_views),
                          C.byref(st), C.c_int(threads), _p(face_ptr), C.byref(vw), C.byref(cs),
                          C.byref(ql), C.byref(info))
    if rc:
        raise RuntimeError(f"orc_data_costs rc={rc}")
    n = int(info.nnz)
    view = np.ctypeslib.as_array(C.cast(vw, C.POINTER(C.c_uint16)), (max(n, 1),))[:n].copy()
    cost = np.ctypeslib.as_array(C.cast(cs, C.POINTER(C.c_float)), (max(n, 1),))[:n].copy()
    qual = np.ctypeslib.as_array(C.cast(ql, C.POINTER(C.c_float)), (max(n, 1),))[:n].copy()
    L.orc_free(vw); L.orc_free(cs); L.orc_free(ql)
    return dict(face_ptr=face_ptr, view=view, cost=cost, quality=qual,
                max_quality=float(info.max_quality), percentile=float(info.percentile))


def view_selection(adj_ptr, adj_idx, face_ptr, view, cost, threads=0, **kw):
    L = lib()
    pr = mrf_params(**kw)
    F = len(face_ptr) - 1
    labels = np.zeros(F, np.uint32)
    trace = np.full(pr.max_iterations + 1, np.nan)
    info = MrfInfo()
    view = np.ascontiguousarray(view, np.uint16)
    cost = np.ascontiguousarray(cost, np.float32)
    rc = L.orc_view_selection(C.c_uint32(F), _p(adj_ptr), _p(adj_idx), _p(face_ptr), _p(view), _p(cost),
                              C.byref(pr), C.c_int(threads), _p(labels), _p(trace), C.byref(info))
    if rc:
        raise RuntimeError(f"orc_view_selection rc={rc}")
    return dict(labels=labels, iterations=int(info.iterations), energy=float(info.energy_final),
                energy_initial=float(info.energy_initial), unseen=int(info.unseen),
                trace=trace[:info.iterations + 1].copy())


def mrf_energy(adj_ptr, adj_idx, face_ptr, view, cost, labels):
    F = len(face_ptr) - 1
    return float(lib().orc_mrf_energy(C.c_uint32(F), _p(adj_ptr), _p(adj_idx), _p(face_ptr),
                                      _p(np.ascontiguousarray(view, np.uint16)),
                                      _p(np.ascontiguousarray(cost, np.float32)),
                                      _p(np.ascontiguousarray(labels, np.uint32))))


def mrf_energy_fixed(adj_ptr, adj_idx, face_ptr, view, cost, labels):
    F = len(face_ptr) - 1
    return int(lib().orc_mrf_energy_fixed(C.c_uint32(F), _p(adj_ptr), _p(adj_idx), _p(face_ptr),
                                          _p(np.ascontiguousarray(view, np.uint16)),
                                          _p(np.ascontiguousarray(cost, np.float32)),
                                          _p(np.ascontiguousarray(labels, np.uint32))))


def mrf_brute_force(adj_ptr, adj_idx, face_ptr, view, cost):
    F = len(face_ptr) - 1
    labels = np.zeros(F, np.uint32)
    e = lib().orc_mrf_brute_force(C.c_uint32(F), _p(adj_ptr), _p(adj_idx), _p(face_ptr),
                                  _p(np.ascontiguousarray(view, np.uint16)),
                                  _p(np.ascontiguousarray(cost, np.float32)), _p(labels))
    return float(e), labels


def mrf_sample_forest(adj_ptr, adj_idx, face_ptr, iteration, **kw):
    pr = mrf_params(**kw)
    F = len(face_ptr) - 1
    level = np.zeros(F, np.uint32)
    lib().orc_mrf_sample_forest(C.c_uint32(F), _p(adj_ptr), _p(adj_idx), _p(face_ptr), C.byref(pr),
                                C.c_uint32(iteration), _p(level))
    return level


def global_seam_leveling(scene, rings, labels, images=None):
    L = lib()
    views, keep = make_views(scene, images)
    vf_ptr, vf_idx, vv_ptr, vv_idx = rings
    Vn = scene.verts.shape[0]
    row_ptr = np.zeros(Vn + 1, np.uint32)
    rl, x, rhs = C.c_void_p(), C.c_void_p(), C.c_void_p()
    info = SeamInfo()
    labels = np.ascontiguousarray(labels, np.uint32)
    rc = L.orc_global_seam_leveling(_p(scene.verts), C.c_uint32(Vn), _p(scene.faces),
                                    C.c_uint32(scene.num_faces), _p(vf_ptr), _p(vf_idx), _p(vv_ptr),
                                    _p(vv_idx), _p(labels), views, C.c_uint32(scene.num_views),
                                    C.c_int(0), _p(row_ptr), C.byref(rl), C.byref(x), C.byref(rhs),
                                    C.byref(info))
    if rc:
        raise RuntimeError(f"orc_global_seam_leveling rc={rc}")
    R = int(info.num_rows)
    def grab(ptr, ctype, n, shape):
        a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), (max(n, 1),))[:n].copy().reshape(shape)
        L.orc_free(ptr)
        return a
    row_label = grab(rl, C.c_uint32, R, (R,))
    xv = grab(x, C.c_float, 3 * R, (R, 3))
    rv = grab(rhs, C.c_float, 3 * R, (R, 3))
    cp, cc, cv = C.c_void_p(), C.c_void_p(), C.c_void_p()
    L.orc_seam_last_matrix(C.byref(cp), C.byref(cc), C.byref(cv))
    nz = int(info.nnz_full)
    csr_ptr = np.ctypeslib.as_array(C.cast(cp, C.POINTER(C.c_uint32)), (R + 1,)).copy()
    csr_col = np.ctypeslib.as_array(C.cast(cc, C.POINTER(C.c_uint32)), (max(nz, 1),))[:nz].copy()
    csr_val = np.ctypeslib.as_array(C.cast(cv, C.POINTER(C.c_float)), (max(nz, 1),))[:nz].copy()
    return dict(row_ptr=row_ptr, row_label=row_label, x=xv, rhs=rv, csr=(csr_ptr, csr_col, csr_val),
                num_a_rows=int(info.num_a_rows), num_gamma_rows=int(info.num_gamma_rows),
                iterations=list(info.iterations), residual=list(info.residual))
