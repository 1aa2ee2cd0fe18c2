"""CPU tests (-m "not gpu"): the oracle against known answers derived from the reference's code
semantics (SURVEY.md section 4 -- the reference ships no tests or golden vectors), the host logic,
and the C-ABI export table."""
import ctypes as C
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _f(*a):
    return (C.c_float * len(a))(*a)


def _view(orc, w=64, h=48, f=50.0, pos=(0, 0, 0)):
    v = orc.View()
    v.pos[:] = list(pos)
    v.viewdir[:] = [0, 0, 1]
    v.proj[:] = [f, 0, w / 2, 0, f, h / 2, 0, 0, 1]
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = -np.asarray(pos, np.float32)
    v.w2c[:] = m.ravel().tolist()
    v.width, v.height = w, h
    return v


# ---- 1. Tri (tri.h:50-84) ----------------------------------------------------------------------
def test_tri_area_and_inside(orc):
    L = orc.lib()
    assert L.orc_tri_area(_f(0, 0), _f(4, 0), _f(0, 3)) == 6.0
    assert L.orc_tri_area(_f(0, 0), _f(0, 3), _f(4, 0)) == 6.0       # orientation independent
    assert L.orc_tri_area(_f(1, 1), _f(2, 2), _f(3, 3)) == 0.0       # degenerate
    t = (_f(0, 0), _f(4, 0), _f(0, 4))
    assert L.orc_tri_inside(*t, C.c_float(1), C.c_float(1)) == 1
    assert L.orc_tri_inside(*t, C.c_float(2), C.c_float(2)) == 1     # on the hypotenuse: alpha+beta == 1
    assert L.orc_tri_inside(*t, C.c_float(2.01), C.c_float(2.01)) == 0
    assert L.orc_tri_inside(*t, C.c_float(-0.01), C.c_float(1)) == 0


# ---- 2. Histogram (histogram.cpp:27-63) ---------------------------------------------------------
def test_histogram_percentile_returns_previous_upper_bound(orc):
    L = orc.lib()
    vals = np.arange(1, 101, dtype=np.float32)          # 1..100, max 100, 11 bins -> width 10
    p = L.orc_histogram_percentile(vals.ctypes.data_as(C.c_void_p), C.c_uint64(100), C.c_float(100.0),
                                   C.c_int(11), C.c_float(0.5))
    # bins: idx=floor(v/100*10): bin0 holds 1..9 (9), bin1 10..19 (10) ... cumulative 9,19,...,59
    # loop: before adding bin i, if num/100 > .5 return upper bound computed for bin i-1.
    # num after bins 0..5 = 59 > 50 -> detected at i=6, returns upper(i=5) = 5/10*100 = 50
    assert p == 50.0
    # never exceeded -> max
    p = L.orc_histogram_percentile(vals.ctypes.data_as(C.c_void_p), C.c_uint64(100), C.c_float(100.0),
                                   C.c_int(11), C.c_float(1.0))
    assert p == 100.0


# ---- 3. projection (texture_view.h:161-166) -----------------------------------------------------
def test_pixel_coords_identity_extrinsics(orc):
    v = _view(orc)
    out = (C.c_float * 2)()
    orc.lib().orc_pixel_coords(C.byref(v), _f(0.2, -0.1, 2.0), out)
    assert out[0] == np.float32(np.float32(50 * np.float32(0.2) + 32 * 2.0) / np.float32(2.0) - np.float32(0.5))
    assert out[1] == np.float32(np.float32(50 * np.float32(-0.1) + 24 * 2.0) / np.float32(2.0) - np.float32(0.5))


# ---- 4. footprint integral (texture_view.cpp:134-251) -------------------------------------------
def test_face_quality_constant_gradient(orc):
    L = orc.lib()
    v = _view(orc, 64, 48, 50.0)
    g = 77
    grad = np.full((48, 64), g, np.uint8)
    rng = np.random.RandomState(0)
    for _ in range(50):
        pts = rng.uniform([-0.5, -0.4, 1.5], [0.5, 0.4, 2.5], size=(3, 3)).astype(np.float32)
        px = []
        for p in pts:
            o = (C.c_float * 2)()
            L.orc_pixel_coords(C.byref(v), _f(*p), o)
            px.append((o[0], o[1]))
        if not all(0 <= x < 63 and 0 <= y < 47 for x, y in px):
            continue
        area = L.orc_tri_area(_f(*px[0]), _f(*px[1]), _f(*px[2]))
        q = L.orc_face_quality(C.byref(v), grad.ctypes.data_as(C.c_void_p), _f(*pts[0]), _f(*pts[1]), _f(*pts[2]), 1)
        assert q == pytest.approx(area * g / 255.0, rel=1e-6)        # GMI = area * mean gradient
        qa = L.orc_face_quality(C.byref(v), grad.ctypes.data_as(C.c_void_p), _f(*pts[0]), _f(*pts[1]), _f(*pts[2]), 0)
        assert qa == area                                             # DATA_TERM_AREA


def test_face_quality_fast_and_slow_paths_agree(orc):
    """Axis-aligned edges force the slow (barycentric) path (m2 == 0 or non-finite slopes); a tiny
    rotation takes the fast scanline path.  On a linear-ramp gradient both must be close."""
    L = orc.lib()
    v = _view(orc, 128, 96, 100.0)
    ramp = np.tile((np.arange(128) * 1.5).astype(np.uint8), (96, 1))
    a = np.array([[-0.3, -0.2, 2], [0.3, -0.2, 2], [-0.3, 0.25, 2]], np.float32)   # right angle, axis aligned
    q_slow = L.orc_face_quality(C.byref(v), ramp.ctypes.data_as(C.c_void_p), _f(*a[0]), _f(*a[1]), _f(*a[2]), 1)
    th = 1e-3
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]], np.float32)
    b = (a @ R.T).astype(np.float32)
    q_fast = L.orc_face_quality(C.byref(v), ramp.ctypes.data_as(C.c_void_p), _f(*b[0]), _f(*b[1]), _f(*b[2]), 1)
    assert q_slow > 0 and q_fast > 0
    assert q_fast == pytest.approx(q_slow, rel=0.05)


def test_face_quality_tiny_triangle_uses_vertex_samples(orc):
    L = orc.lib()
    v = _view(orc, 64, 48, 50.0)
    grad = np.full((48, 64), 255, np.uint8)
    a = np.array([[0.0, 0.0, 2], [0.02, 0.0, 2], [0.0, 0.02, 2]], np.float32)      # 0.5 x 0.5 px -> area 0.125
    q = L.orc_face_quality(C.byref(v), grad.ctypes.data_as(C.c_void_p), _f(*a[0]), _f(*a[1]), _f(*a[2]), 1)
    assert q == pytest.approx(0.125, rel=1e-5)                         # area <= 0.5: mean of 3 linear_at * area
    z = np.zeros((48, 64), np.uint8)
    assert L.orc_face_quality(C.byref(v), z.ctypes.data_as(C.c_void_p), _f(*a[0]), _f(*a[1]), _f(*a[2]), 1) == 0.0


# ---- 5./6. cull rules and normalisation (calculate_data_costs.cpp:183-188, 277-302) --------------
def _single_face_scene(scene_mod, normal_angle_deg, flip=False, behind=False):
    """one triangle at the origin, camera on +z looking down -z, face normal tilted by angle"""
    th = np.deg2rad(normal_angle_deg)
    Rm = np.array([[1, 0, 0], [0, np.cos(th), -np.sin(th)], [0, np.sin(th), np.cos(th)]])
    tri = np.array([[-0.1, -0.1, 0], [0.1, -0.1, 0], [0.0, 0.1, 0]]) @ Rm.T
    F = np.array([[0, 1, 2]], np.uint32) if not flip else np.array([[0, 2, 1]], np.uint32)
    V = tri.astype(np.float32)
    target = (0, 0, 0) if not behind else (0, 0, 10)
    cam = scene_mod.look_at_camera((0, 0, 3), target, 300.0, 320, 240, up=(0, 1, 0))
    s = scene_mod._assemble("one", V, F, [cam], 320, 240)
    return s


@pytest.mark.parametrize("angle,flip,behind,expect", [(0, False, False, 1), (60, False, False, 1),
                                                      (74, False, False, 1), (76, False, False, 0),
                                                      (0, True, False, 0), (0, False, True, 0)])
def test_cull_rules(orc, scene_mod, angle, flip, behind, expect):
    s = _single_face_scene(scene_mod, angle, flip, behind)
    r = orc.data_costs(s)
    assert len(r["view"]) == expect


def test_cost_normalisation(orc, get_scene):
    r = orc.data_costs(get_scene("small"))
    q, c = r["quality"], r["cost"]
    p = np.float32(r["percentile"])
    expect = np.float32(1.0) - np.minimum(np.float32(1.0), (q / p).astype(np.float32))
    assert np.array_equal(c, expect.astype(np.float32))
    assert c.min() >= 0.0 and c.max() < 1.0
    assert r["max_quality"] == q.max()
    assert np.all(q > 0)                                            # quality 0 candidates are dropped (:222)
    fp = r["face_ptr"].astype(np.int64)
    for f in range(0, len(fp) - 1, 97):                             # ascending view ids per face (:272)
        assert np.all(np.diff(r["view"][fp[f]:fp[f + 1]].astype(int)) > 0)


# ---- image preparation (texture_view.cpp:42-132) -------------------------------------------------
def test_validity_mask_flood_fill_and_erosion_quirk(orc):
    rgb = np.full((20, 30, 3), 9, np.uint8)
    rgb[:, :4] = 0                       # black column band touching the corners -> invalid
    rgb[10:13, 10:13] = 0                # interior blob not connected to a corner -> stays valid
    rgb[0, 29] = 0                       # isolated black corner pixel -> invalid
    m = orc.validity_mask(rgb)
    assert not m[:, :4].any() and m[:, 4:29].all() and m[10:13, 10:13].all() and m[0, 29] == 0
    e = orc.erode(m)
    assert not e[1:-1, 4].any()          # interior invalid column 3 dilates into column 4
    assert e[0, 4] == 0 and e[19, 4] == 0  # (1,3) / (18,3) are interior and invalid: their 3x3 covers row 0 / 19
    assert e[:, 5:29].all()              # nothing further
    assert e[0, 28] == 1 and e[1, 28] == 1 and e[1, 29] == 1  # the invalid BORDER pixel (0,29) does not dilate


def test_validity_mask_erosion_exact(orc):
    rgb = np.full((20, 30, 3), 9, np.uint8)
    rgb[:, :4] = 0
    m = orc.validity_mask(rgb)
    e = orc.erode(m)
    ref = m.copy()
    for y in range(1, 19):
        for x in range(1, 29):
            if not m[y, x]:
                ref[y - 1:y + 2, x - 1:x + 2] = 0
    assert np.array_equal(e, ref)
    # the quirk: an invalid pixel ON the image border does not dilate
    rgb2 = np.full((20, 30, 3), 9, np.uint8)
    rgb2[0, 0] = 0
    e2 = orc.erode(orc.validity_mask(rgb2))
    assert e2[0, 0] == 0 and e2.sum() == 20 * 30 - 1


def test_gradient_magnitude_against_numpy(orc):
    rng = np.random.RandomState(3)
    rgb = rng.randint(0, 256, size=(40, 50, 3)).astype(np.uint8)
    g = orc.gradient_magnitude(rgb)
    f = rgb.astype(np.float32)
    lum = ((f[..., 0] * np.float32(0.21) + f[..., 1] * np.float32(0.72)) + f[..., 2] * np.float32(0.07)
           + np.float32(0.5)).astype(np.uint8).astype(np.float64)
    gx = (lum[:-2, 2:] - lum[:-2, :-2]) + 2 * (lum[1:-1, 2:] - lum[1:-1, :-2]) + (lum[2:, 2:] - lum[2:, :-2])
    gy = (lum[2:, :-2] - lum[:-2, :-2]) + 2 * (lum[2:, 1:-1] - lum[:-2, 1:-1]) + (lum[2:, 2:] - lum[:-2, 2:])
    ref = np.zeros((40, 50), np.uint8)
    ref[1:-1, 1:-1] = np.minimum(255.0, np.sqrt(gx * gx + gy * gy)).astype(np.uint8)
    assert np.array_equal(g, ref)
    assert not g[0].any() and not g[-1].any() and not g[:, 0].any() and not g[:, -1].any()


# ---- BVH any-hit vs brute force -------------------------------------------------------------------
def test_bvh_matches_brute_force(orc, get_scene):
    s = get_scene("small")
    L = orc.lib()
    L.orc_bvh_build.restype = C.c_void_p
    b = C.c_void_p(L.orc_bvh_build(s.verts.ctypes.data_as(C.c_void_p), s.faces.ctypes.data_as(C.c_void_p),
                                   C.c_uint32(s.num_faces)))
    rng = np.random.RandomState(1)
    hits = 0
    for i in range(400):
        o = s.verts[rng.randint(len(s.verts))]
        tgt = s.pos[rng.randint(s.num_views)] if i % 2 else rng.normal(size=3).astype(np.float32) * 2
        d = (tgt - o).astype(np.float32)
        tmax = np.float32(np.linalg.norm(d))
        d = (d / tmax).astype(np.float32)
        args = (_f(*o), _f(*d), C.c_float(tmax * 1e-4), C.c_float(tmax))
        a = L.orc_bvh_occluded(b, *args)
        bf = L.orc_brute_occluded(s.verts.ctypes.data_as(C.c_void_p), s.faces.ctypes.data_as(C.c_void_p),
                                  C.c_uint32(s.num_faces), *args)
        assert a == bf
        hits += a
    assert 0 < hits < 400
    L.orc_bvh_free(b)


# ---- 7. MRF -----------------------------------------------------------------------------------------
def _random_mrf(rng, n, edges, max_labels=4, num_views=6, unseen=()):
    ptr = [0]
    view, cost = [], []
    for i in range(n):
        k = 0 if i in unseen else rng.randint(1, max_labels + 1)
        vs = np.sort(rng.choice(num_views, size=k, replace=False))
        view += vs.tolist()
        cost += rng.uniform(0, 1, size=k).astype(np.float32).tolist()
        ptr.append(len(view))
    adj = [[] for _ in range(n)]
    for a, b in edges:
        adj[a].append(b)
        adj[b].append(a)
    ap = np.cumsum([0] + [len(x) for x in adj]).astype(np.uint32)
    ai = np.array([y for x in adj for y in x], np.uint32)
    return ap, ai, np.array(ptr, np.uint64), np.array(view, np.uint16), np.array(cost, np.float32)


def test_mrf_exact_on_trees_single_root(orc):
    rng = np.random.RandomState(5)
    for trial in range(20):
        n = 10
        edges = [(i, rng.randint(0, i)) for i in range(1, n)]        # random tree
        ap, ai, ptr, view, cost = _random_mrf(rng, n, edges)
        e_bf, lab_bf = orc.mrf_brute_force(ap, ai, ptr, view, cost)
        r = orc.view_selection(ap, ai, ptr, view, cost, threads=1, root_div=0, rounds=64, max_iterations=1, window=1)
        assert r["energy"] == pytest.approx(e_bf, abs=1e-5)            # one exact DP sweep = global optimum


def test_mrf_monotone_and_bounded_on_loopy_graphs(orc):
    rng = np.random.RandomState(6)
    for trial in range(20):
        n = 11
        edges = {(i, (i + 1) % n) for i in range(n)} | {(rng.randint(n), rng.randint(n)) for _ in range(6)}
        edges = [(a, b) for a, b in edges if a != b]
        edges = list({(min(a, b), max(a, b)) for a, b in edges})
        ap, ai, ptr, view, cost = _random_mrf(rng, n, edges, unseen=(3,))
        e_bf, _ = orc.mrf_brute_force(ap, ai, ptr, view, cost)
        r = orc.view_selection(ap, ai, ptr, view, cost, threads=1, window=10, ratio=0.0, max_iterations=30)
        tr = r["trace"]
        assert np.all(np.diff(tr) <= 1e-9)                             # energy never increases
        assert r["energy"] >= e_bf - 1e-5
        assert r["energy"] <= tr[0] + 1e-9
        assert r["labels"][3] == 0 and r["unseen"] == 1                # unseen face: label 0 (view_selection.cpp:50-51)
        assert r["energy"] == pytest.approx(orc.mrf_energy(ap, ai, ptr, view, cost, r["labels"]))
        # BCD over induced forests ends close to the optimum on these tiny problems
        assert r["energy"] <= e_bf + 1.0 + 1e-5


def test_mrf_forest_is_induced_forest(orc, oracle_pipeline):
    r = oracle_pipeline("C1d", ("dc", "mrf"))
    ap, ai = r["adj"]
    dc = r["dc"]
    for t in (1, 2, 3):
        for parts in (1, 3):
            lvl = orc.mrf_sample_forest(ap, ai, dc["face_ptr"], t, num_parts=parts)
            inS = lvl <= 32
            assert 0.3 < inS.mean() < 0.9
            F = len(lvl)
            psz = (F + parts - 1) // parts
            # edges inside S (same partition): every node has exactly one neighbour with smaller level,
            # none with equal level => |E_S| = |S| - #roots and acyclic
            nodes = np.flatnonzero(inS)
            n_edges = 0
            for v in nodes:
                nb = ai[ap[v]:ap[v + 1]]
                nb = nb[(nb // psz) == (v // psz)]
                nbS = nb[inS[nb]]
                assert not np.any(lvl[nbS] == lvl[v])
                lower = np.sum(lvl[nbS] < lvl[v])
                assert lower == (0 if lvl[v] == 0 else 1)
                n_edges += lower
            assert n_edges == len(nodes) - np.sum(lvl[nodes] == 0)
            if parts > 1:   # no two adjacent nodes of different partitions are both free to move
                for v in nodes:
                    nb = ai[ap[v]:ap[v + 1]]
                    rem = nb[(nb // psz) != (v // psz)]
                    assert not np.any(inS[rem])


def test_mrf_beats_unary_argmin_and_partitions_stay_monotone(orc, oracle_pipeline):
    r = oracle_pipeline("C1d", ("dc", "mrf"))
    ap, ai = r["adj"]
    dc = r["dc"]
    base = r["mrf"]
    assert base["energy"] < 0.7 * base["energy_initial"]
    for parts in (2, 8):
        rp = orc.view_selection(ap, ai, dc["face_ptr"], dc["view"], dc["cost"], threads=1, num_parts=parts)
        assert np.all(np.diff(rp["trace"]) <= 1e-9)
        assert rp["energy"] < 1.03 * base["energy"]
    # thread count does not change the result (deterministic, order independent)
    r8 = orc.view_selection(ap, ai, dc["face_ptr"], dc["view"], dc["cost"], threads=4)
    assert np.array_equal(r8["labels"], base["labels"])


# ---- 8. seam system -----------------------------------------------------------------------------------
def test_seam_system_is_weighted_laplacian_and_pcg_matches_scipy(oracle_pipeline):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    o = oracle_pipeline("C1d")["seam"]
    cp, cc, cv = o["csr"]
    R = len(cp) - 1
    A = sp.csr_matrix((cv.astype(np.float64), cc.astype(np.int64), cp.astype(np.int64)), shape=(R, R))
    assert abs(A - A.T).max() == 0
    assert abs(A.sum(axis=1)).max() < 1e-5                            # row sums 0
    d = A.diagonal()
    off = A - sp.diags(d)
    assert off.max() <= 0 and d.min() >= 0                            # Laplacian sign pattern => PSD
    assert o["num_a_rows"] > 0 and o["num_gamma_rows"] > 0
    for ch in range(3):
        assert o["residual"][ch] < 1e-4 and 0 < o["iterations"][ch] < 1000
        rhs = o["rhs"][:, ch].astype(np.float64)
        assert abs(rhs.sum()) < 1e-3                                  # A^T b is orthogonal to constants
        x = o["x"][:, ch].astype(np.float64)
        assert abs(x.mean()) < 1e-6                                   # centred (:277)
        assert np.linalg.norm(A @ x - rhs) / np.linalg.norm(rhs) < 2e-4
        xs, _ = spl.cg(A, rhs, rtol=1e-10, maxiter=20000, M=sp.diags(1.0 / np.maximum(d, 1e-30)))
        xs -= xs.mean()
        # The Laplacian mixes weights 1 and 0.01 and is singular, so stopping at |r|/|b| < 1e-4 (as the
        # reference does) leaves a visible smooth error against the exact minimiser: sanity bound only.
        assert np.linalg.norm(x - xs) / np.linalg.norm(xs) < 0.3


def test_seam_two_label_strip_closed_form(orc, scene_mod):
    """Flat strip textured from two views whose images differ by a constant offset d: the leveled
    colours must meet in the middle, i.e. x(l2) - x(l1) ~ -(c2 - c1) summed over the seam."""
    n = 8
    V, F = scene_mod.terrain(n, amplitude=0.0)
    cams = [scene_mod.look_at_camera((0.0, 0.0, 4.0), (0, 0, 0), 120.0, 320, 240, up=(0, 1, 0)) for _ in range(2)]
    s = scene_mod._assemble("strip", V, F, cams, 320, 240, with_images=False)
    img = np.empty((2, 240, 320, 3), np.uint8)
    img[0] = 100
    img[1] = 151                                                       # +51/255 = +0.2 everywhere
    s.images = img
    cx = V[F].mean(axis=1)[:, 0]
    labels = np.where(cx < 0, 1, 2).astype(np.uint32)
    rings = scene_mod.vertex_rings(s.faces, len(V))
    o = orc.global_seam_leveling(s, rings, labels)
    assert o["num_a_rows"] == n + 1                                    # one A row per seam vertex
    rp, rl, x = o["row_ptr"], o["row_label"], o["x"]
    seam_v = [v for v in range(len(V)) if rp[v + 1] - rp[v] == 2]
    assert len(seam_v) == n + 1
    for v in seam_v:
        x1, x2 = x[rp[v]], x[rp[v] + 1]
        # g_l1 - g_l2 = b = c2 - c1 = +0.2 per channel, up to the 0.01 regulariser pull
        assert np.allclose(x1 - x2, 0.2, atol=0.03)
    assert np.allclose(x.mean(axis=0), 0, atol=1e-6)


# ---- C-ABI export table ---------------------------------------------------------------------------------
def test_c_abi_library_exports_every_declared_symbol(b2):
    import re
    hdr = open(os.path.join(ROOT, "include", "b2tex.h")).read()
    declared = sorted(set(re.findall(r"\b(b2tex_[a-z_0-9]+)\s*\(", hdr)))
    assert sorted(declared) == sorted(b2.EXPORTS)
    L = b2.lib()                       # dlopen only; no compute call without a GPU
    for name in declared:
        assert hasattr(L, name), name


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "mvs-texturing_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "liboracle" not in txt and "import oracle" not in txt and "orc_" not in txt, f


# ---- golden fixtures (oracle regression snapshots; generator: tests/golden/make_golden.py) -------------
def test_oracle_matches_golden_snapshots(orc, scene_mod, oracle_pipeline):
    import zlib
    path = os.path.join(ROOT, "tests", "golden", "oracle_snapshots.json")
    gold = json.load(open(path))
    for name, g in gold.items():
        r = oracle_pipeline(name)
        dc, m, sm = r["dc"], r["mrf"], r["seam"]
        assert len(dc["view"]) == g["nnz"]
        assert zlib.crc32(dc["face_ptr"].tobytes()) == g["crc_face_ptr"]
        assert zlib.crc32(dc["view"].tobytes()) == g["crc_view"]
        assert zlib.crc32(dc["cost"].tobytes()) == g["crc_cost"]
        assert m["iterations"] == g["mrf_iterations"]
        assert zlib.crc32(m["labels"].tobytes()) == g["crc_labels"]
        assert m["energy"] == pytest.approx(g["mrf_energy"], rel=1e-12)
        assert len(sm["row_label"]) == g["seam_rows"]
        assert list(sm["iterations"]) == g["cg_iterations"]
        if "patches" in g:                                  # texture patches + zero adjust_colors (oracle/patches.py)
            import patches as P
            s = scene_mod.config(name)
            pp, _ = P.generate_texture_patches(orc, s, r["adj"], m["labels"])
            crc = dict(tex=0, img=0, val=0, bl=0)
            for q in pp:
                img, val, bl = P.adjust_colors(q, np.zeros((3 * len(q.faces), 3), np.float32))
                crc["tex"] = zlib.crc32(np.ascontiguousarray(q.texcoords, np.float32).tobytes(), crc["tex"])
                crc["img"] = zlib.crc32(np.ascontiguousarray(img).tobytes(), crc["img"])
                crc["val"] = zlib.crc32(np.ascontiguousarray(val).tobytes(), crc["val"])
                crc["bl"] = zlib.crc32(np.ascontiguousarray(bl).tobytes(), crc["bl"])
            assert (len(pp), sum(len(q.faces) for q in pp)) == (g["patches"], g["patch_faces"])
            assert (crc["tex"], crc["img"], crc["val"], crc["bl"]) == (g["crc_patch_texcoords"], g["crc_patch_images"],
                                                                       g["crc_patch_validity"], g["crc_patch_blending"])


# ---- texture patches: the reference's patch-relative colour sampling vs the stage-isolated shortcut ----
@pytest.mark.parametrize("name", ["tiny", "small", "C1d"])
def test_patch_sampling_equals_view_sampling(orc, scene_mod, oracle_pipeline, get_scene, name):
    """oracle/seam.c and the CUDA path sample seam colours from the whole view of a label; the reference
    samples cropped float patches at patch-relative coordinates (generate_texture_patches.cpp:78-138,
    seam_leveling.cpp:61-91).  Same image content, different coordinate frame: Rhs = A^T b must agree."""
    import patches as P   # oracle/patches.py
    s = get_scene(name)
    r = oracle_pipeline(name)
    o = r["seam"]
    rhs_p, patches, vpi = P.seam_rhs_from_patches(orc, s, r["adj"], r["rings"], r["mrf"]["labels"],
                                                   o["row_ptr"], o["row_label"])
    # structure: every seen face is in exactly one patch of its own label; texcoords stay inside the patch
    owner = {}
    for pid, p in enumerate(patches):
        h, w, _ = p.image.shape
        assert p.texcoords.min() >= 0 and p.texcoords[:, 0].max() <= w - 1 and p.texcoords[:, 1].max() <= h - 1
        for f in p.faces:
            assert f not in owner and r["mrf"]["labels"][f] == p.label
            owner[f] = pid
    assert len(owner) == int(np.sum(r["mrf"]["labels"] != 0))
    # every (vertex, label) unknown of the seam system corresponds to >= 1 (vertex, patch) projection
    for v in range(s.verts.shape[0]):
        labs = {patches[pid].label for pid in vpi[v]}
        assert labs == set(o["row_label"][o["row_ptr"][v]:o["row_ptr"][v + 1]].tolist())
    scale = np.abs(o["rhs"]).max()
    assert scale > 0
    assert np.abs(rhs_p - o["rhs"]).max() < 2e-6 * max(1.0, scale)   # measured 1.5e-7 .. 2.7e-7


def test_adjust_colors_known_answers(orc):
    """TexturePatch::adjust_colors (texture_patch.cpp:41-116): constant offsets and a linear ramp."""
    import patches as P
    tc = np.array([[3.2, 3.1], [12.7, 4.4], [5.9, 11.6], [12.7, 4.4], [13.8, 12.9], [5.9, 11.6]], np.float32)
    img = np.full((18, 18, 3), 0.25, np.float32)
    p = P.Patch(1, [0, 1], tc, img, [0, 0, 17, 17])
    out, validity, blending = P.adjust_colors(p, np.full((6, 3), 0.125, np.float32))
    inside = blending == 255
    assert inside.sum() > 40 and np.all(validity[inside] == 255)
    assert np.allclose(out[validity == 255], 0.375, atol=1e-6)       # image + constant everywhere it is valid
    assert np.all(out[validity == 0] == 0)                           # untouched pixels are zeroed (:110-114)
    ring = blending == 64                                             # <= sqrt(2) px outside: extrapolated
    assert ring.sum() > 10 and np.all(validity[ring] == 255)
    assert not validity[0, 0] and not validity[17, 17]
    # linear ramp a(x,y) = 0.01 x - 0.02 y + 0.3 is reproduced exactly by barycentric interpolation,
    # also on the extrapolated ring
    ramp = lambda q: (0.01 * q[:, 0] - 0.02 * q[:, 1] + 0.3).astype(np.float32)
    adj = np.repeat(ramp(tc)[:, None], 3, 1).astype(np.float32)
    out, validity, blending = P.adjust_colors(p, adj)
    ys, xs = np.nonzero(validity)
    expect = 0.25 + 0.01 * xs - 0.02 * ys + 0.3
    assert np.allclose(out[ys, xs, 0], expect, atol=2e-5)


def test_global_seam_leveling_reduces_seam_differences(orc, scene_mod, oracle_pipeline, get_scene):
    """End to end on the oracle: apply the solved offsets to the texture patches (adjust_colors) and
    re-measure the colour differences across the seams -- the quantity the solve minimises."""
    import patches as P
    name = "small"
    s = get_scene(name)
    r = oracle_pipeline(name)
    o, labels = r["seam"], r["mrf"]["labels"]
    rhs0, patches, vpi = P.seam_rhs_from_patches(orc, s, r["adj"], r["rings"], labels, o["row_ptr"], o["row_label"])
    adjusted = P.apply_adjust_values(s, patches, o["row_ptr"], o["row_label"], o["x"])
    for q in adjusted:   # every face corner lies in the valid area of its adjusted patch
        xi = np.clip(np.floor(q.texcoords[:, 0]).astype(int), 0, q.image.shape[1] - 1)
        yi = np.clip(np.floor(q.texcoords[:, 1]).astype(int), 0, q.image.shape[0] - 1)
        assert np.all(q.validity[yi, xi] == 255)
    # same seam measurement on the adjusted patch images
    P_generate = P.generate_texture_patches
    try:
        P.generate_texture_patches = lambda *a, **k: (adjusted, vpi)
        rhs1, _, _ = P.seam_rhs_from_patches(orc, s, r["adj"], r["rings"], labels, o["row_ptr"], o["row_label"])
    finally:
        P.generate_texture_patches = P_generate
    n0, n1 = np.linalg.norm(rhs0), np.linalg.norm(rhs1)
    assert n1 < 0.35 * n0, (n0, n1)


def test_poisson_blend_known_answers():
    """poisson_blending.cpp:49-138: with alpha = 1 the interior takes the source's Laplacian; Dirichlet
    pixels (mask 128/64) keep the destination's values."""
    import patches as P
    rng = np.random.RandomState(2)
    src = rng.uniform(0.2, 0.8, size=(14, 16, 3)).astype(np.float32)
    mask = np.zeros((14, 16), np.uint8)
    mask[2:12, 2:14] = 128
    mask[3:11, 3:13] = 255
    dest = src.copy()
    P.poisson_blend(src, mask, dest, 1.0)
    assert np.allclose(dest, src, atol=2e-5)                       # same boundary, same Laplacian -> same image
    dest = src.copy()
    ring = mask == 128
    dest[ring] += np.float32(0.1)                                   # shift the Dirichlet ring by a constant
    P.poisson_blend(src, mask, dest, 1.0)
    assert np.allclose(dest[mask == 255], src[mask == 255] + 0.1, atol=5e-5)
    assert np.allclose(dest[mask == 0], src[mask == 0])            # pixels outside the mask are untouched


def test_local_seam_leveling_makes_patches_agree_on_seams(orc, scene_mod, oracle_pipeline, get_scene):
    """local_seam_leveling.cpp:105-204 after global leveling: on every seam edge the two patches hold the
    same colours along the projected edge (they are overwritten with the mean), and the blend only touches
    the 20-pixel strip."""
    import patches as P
    name = "tiny"
    s = get_scene(name)
    r = oracle_pipeline(name)
    o, labels = r["seam"], r["mrf"]["labels"]
    patches, vpi = P.generate_texture_patches(orc, s, r["adj"], labels)
    patches = P.apply_adjust_values(s, patches, o["row_ptr"], o["row_label"], o["x"])
    before = [p.image.copy() for p in patches]
    seam_edges = P.find_seam_edges(s, r["adj"], labels)
    assert len(seam_edges) > 10

    def seam_gap(ps):
        gaps = []
        for v1, v2 in seam_edges:
            infos = P.find_mesh_edge_projections(vpi, v1, v2)
            if len(infos) != 2:
                continue
            cols = []
            for pid, p1, p2 in infos:   # colour at the stamped pixel of the edge midpoint
                mid = ((p1 + p2) * np.float32(0.5)).astype(np.float32) + np.float32(0.5)
                cols.append(ps[pid].image[int(mid[1]), int(mid[0])])
            gaps.append(np.abs(cols[0] - cols[1]).max())
        return np.array(gaps)

    g0 = seam_gap(patches)
    P.local_seam_leveling(s, r["adj"], labels, patches, vpi)
    g1 = seam_gap(patches)
    assert len(g1) > 10 and np.median(g1) < 0.25 * np.median(g0) and np.median(g1) < 0.02
    for p, b in zip(patches, before):
        changed = np.abs(p.image - b).max(axis=2) > 1e-6
        assert not np.any(changed & (p.blending == 0))             # only masked pixels are re-solved
        assert np.isfinite(p.image).all()


def test_spanning_forest_bound_brackets_the_solver(orc, scene_mod):
    """tools/mrf_quality.py: the exact optimum of the model restricted to a spanning forest is a lower bound of every labeling's
    energy (dropped Potts terms are >= 0), and the labeling that attains it is a valid labeling whose full energy lies above the
    solver's result on this scene."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("mrf_quality", os.path.join(os.path.dirname(__file__), "..", "tools", "mrf_quality.py"))
    mq = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mq)
    s = scene_mod.config("tiny")
    ap, ai = scene_mod.face_adjacency(s.faces)
    dc = orc.data_costs(s)
    fp, view, cost = dc["face_ptr"], dc["view"], dc["cost"]
    m = orc.view_selection(ap, ai, fp, view, cost, threads=1)
    lb, tree_labels = mq.tree_lower_bound(ap, ai, fp, view, cost)
    e_tree = orc.mrf_energy(ap, ai, fp, view, cost, tree_labels)
    assert lb <= m["energy"] + 1e-6 and lb <= e_tree + 1e-6
    assert m["energy"] <= e_tree + 1e-6 <= m["energy_initial"] + 1e-6 or e_tree <= m["energy_initial"] + 1e-6
    for f in range(s.num_faces):   # every label is one of the face's own views
        if fp[f + 1] > fp[f]:
            assert tree_labels[f] - 1 in view[fp[f]:fp[f + 1]]


def test_image_synthesis_is_thread_order_independent(scene_mod):
    """make_images spreads large view sets over threads; every view is seeded by its index, so the bytes do not depend on it."""
    import zlib
    a = scene_mod.make_images(9, 4096, 2048)       # above the threading threshold
    ref = np.empty_like(a)
    T = scene_mod.base_texture(4096, 2048)
    tint = np.array([1.0, 0.9, 0.8], dtype=np.float32)
    for v in range(9):
        rs = np.random.RandomState(1000 + v)
        gain, bias = rs.uniform(0.8, 1.2), rs.uniform(-20.0, 20.0)
        dx, dy = int(rs.randint(0, 4096)), int(rs.randint(0, 2048))
        base = np.float32(gain * 200.0) * np.roll(T, (dy, dx), axis=(0, 1)) + np.float32(bias + 25.0)
        for c in range(3):
            ref[v, :, :, c] = np.clip(np.rint(base * tint[c]), 1, 255).astype(np.uint8)
    assert zlib.crc32(a.tobytes()) == zlib.crc32(ref.tobytes())
