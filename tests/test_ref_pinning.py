"""Pins the oracle against the reference's own translation units.

oracle/_ref/libtexref.so is built from /root/reference/libs/tex/*.cpp (unmodified, compiled where they lie) and
the dependency shims in oracle/refshim/ (MVE / rayint / Eigen / mapMAP are not vendored in the reference, see
oracle/refshim/README.md).  Everything written in libs/tex -- cull rules, projection, validity mask, footprint
integral, histogram, normalisation -- runs as the reference wrote it; the oracle has to agree bit for bit.
"""
import numpy as np
import pytest

refpin = pytest.importorskip("refpin")
if not refpin.available():
    pytest.skip("no libtexref.so and no reference checkout", allow_module_level=True)


@pytest.fixture(scope="module")
def ref():
    refpin.lib()
    return refpin


@pytest.mark.parametrize("name", ["tiny", "small", "occ", "messy"])
@pytest.mark.parametrize("data_term", [1, 0])
def test_data_costs_match_reference_tu(ref, orc, get_scene, name, data_term):
    """tex::calculate_data_costs (calculate_data_costs.cpp:131-323) vs orc_data_costs: same (face, view) set,
    bit-identical costs."""
    s = get_scene(name)
    r = ref.data_costs(s, data_term=data_term)
    o = orc.data_costs(s, data_term=data_term)
    assert int(r["face_ptr"][-1]) > s.num_faces
    assert np.array_equal(r["face_ptr"], o["face_ptr"])
    assert np.array_equal(r["view"], o["view"])
    assert np.array_equal(r["cost"].view(np.uint32), o["cost"].view(np.uint32))


@pytest.mark.parametrize("name", ["tiny", "occ"])
def test_data_costs_without_visibility_test(ref, orc, get_scene, name):
    s = get_scene(name)
    r = ref.data_costs(s, visibility=False)
    o = orc.data_costs(s, visibility=False)
    assert np.array_equal(r["face_ptr"], o["face_ptr"]) and np.array_equal(r["view"], o["view"])
    assert np.array_equal(r["cost"].view(np.uint32), o["cost"].view(np.uint32))
    with_test = ref.data_costs(s)
    assert int(r["face_ptr"][-1]) >= int(with_test["face_ptr"][-1])
    if name == "occ":                                                    # floating plates occlude
        assert int(r["face_ptr"][-1]) > int(with_test["face_ptr"][-1])
        o2 = orc.data_costs(s)
        assert np.array_equal(with_test["face_ptr"], o2["face_ptr"]) and np.array_equal(with_test["view"], o2["view"])
        assert np.array_equal(with_test["cost"].view(np.uint32), o2["cost"].view(np.uint32))


@pytest.mark.parametrize("mode", [1, 2])
def test_outlier_removal_matches_reference_tu(ref, orc, get_scene, mode, monkeypatch):
    """photometric_outlier_detection (calculate_data_costs.cpp:35-129).  The reference's per-face info order
    before detection depends on the OpenMP merge (:241-249; one thread -> descending view id), the oracle uses
    ascending view id, so sums may differ in the last bits: same survivors, costs within 1e-5."""
    s = get_scene("tiny")
    r = ref.data_costs(s, outlier_removal=mode)
    o = orc.data_costs(s, outlier_removal=mode)
    assert np.array_equal(r["face_ptr"], o["face_ptr"]) and np.array_equal(r["view"], o["view"])
    assert np.allclose(r["cost"], o["cost"], rtol=0, atol=1e-5)
    base = orc.data_costs(s)
    assert int(o["face_ptr"][-1]) <= int(base["face_ptr"][-1])


def test_validity_mask_and_erosion_quirk(ref, orc):
    """texture_view.cpp:42-94 (corner flood fill over black pixels) and :109-132 (erosion that leaves the image
    border untouched because the border write lands in the array that is swapped away)."""
    rng = np.random.RandomState(5)
    img = rng.randint(1, 255, size=(40, 56, 3)).astype(np.uint8)
    img[:6, :] = 0; img[:, :4] = 0; img[30:, 50:] = 0          # black frame pieces connected to corners
    img[15:18, 20:23] = 0                                     # an interior black blob: stays valid
    m_ref, m_orc = ref.validity_mask(img), orc.validity_mask(img)
    assert np.array_equal(m_ref, m_orc) and m_ref[16, 21] == 1 and m_ref[2, 10] == 0
    e_ref, e_orc = ref.validity_mask(img, erode=True), orc.erode(m_orc)
    assert np.array_equal(e_ref, e_orc)
    assert e_ref[6, 10] == 0 and e_ref[7, 10] == 1            # one ring eaten next to the black frame
    img2 = rng.randint(1, 255, size=(20, 20, 3)).astype(np.uint8)
    assert ref.validity_mask(img2, erode=True).all()          # the quirk: border pixels stay valid


def test_face_info_matches_reference_tu(ref, orc, get_scene):
    """TextureView::get_face_info (texture_view.cpp:134-251) on random triangles of every size class: sub-pixel
    (vertex sampling), slivers (slow path / skipped scan lines) and large footprints (fast scan line path)."""
    import ctypes as C
    s = get_scene("small")
    k = 3
    rng = np.random.RandomState(11)
    v, keep = ref._one_view(s, k)
    grad = orc.gradient_magnitude(s.images[k])
    # triangles around mesh vertices seen by the view: pick faces, then shrink / stretch them
    dc = orc.data_costs(s)
    faces = [f for f in range(s.num_faces) if k in dc["view"][int(dc["face_ptr"][f]):int(dc["face_ptr"][f + 1])]]
    tris = []
    for f in faces[:300]:
        t = s.verts[s.faces[f]].astype(np.float32)
        c = t.mean(axis=0)
        for scale in (1.0, 0.05, 3.0):
            tris.append((c + (t - c) * np.float32(scale)).astype(np.float32))
        sl = t.copy(); sl[2] = (sl[0] + (sl[1] - sl[0]) * np.float32(0.5) + rng.normal(0, 1e-4, 3)).astype(np.float32)
        tris.append(sl)
    tris = np.array(tris, np.float32)
    for data_term in (1, 0):
        q_ref, _ = ref.face_infos(s, k, tris, data_term=data_term)
        q_orc = np.empty(len(tris), np.float32)
        L = orc.lib()
        for i, t in enumerate(tris):
            q_orc[i] = L.orc_face_quality(C.byref(v), orc._p(grad), orc._p(t[0].copy()), orc._p(t[1].copy()), orc._p(t[2].copy()), data_term)
        ok = ~np.isnan(q_ref)                                  # NaN = projects outside the valid area
        assert ok.sum() > len(tris) // 2
        assert np.array_equal(q_ref[ok].view(np.uint32), q_orc[ok].view(np.uint32))
        if data_term == 1:
            assert (q_ref[ok] > 0).sum() > ok.sum() // 2


def test_tri_and_histogram_match_reference_tu(ref, orc):
    import ctypes as C
    rng = np.random.RandomState(3)
    L = orc.lib()
    for _ in range(300):
        p = rng.uniform(0, 50, size=(3, 2)).astype(np.float32)
        assert ref.tri_area(*p) == L.orc_tri_area(orc._p(p[0].copy()), orc._p(p[1].copy()), orc._p(p[2].copy()))
        x, y = rng.uniform(0, 50, size=2).astype(np.float32)
        assert ref.tri_inside(p[0], p[1], p[2], x, y) == L.orc_tri_inside(orc._p(p[0].copy()), orc._p(p[1].copy()), orc._p(p[2].copy()), C.c_float(x), C.c_float(y))
    for n in (1, 7, 1000, 200000):
        v = (rng.gamma(2.0, 3.0, size=n)).astype(np.float32)
        vmax = float(v.max())
        r = ref.histogram_percentile(v, vmax)
        o = float(L.orc_histogram_percentile(orc._p(v), C.c_uint64(n), C.c_float(vmax), 10000, C.c_float(0.995)))
        assert r == o


def test_pixel_coords_match_reference_tu(ref, orc, get_scene):
    import ctypes as C
    s = get_scene("tiny")
    L = orc.lib()
    for k in (0, 5):
        v, keep = ref._one_view(s, k)
        for i in range(0, s.verts.shape[0], 37):
            x = s.verts[i].copy()
            out = np.empty(2, np.float32)
            L.orc_pixel_coords(C.byref(v), orc._p(x), orc._p(out))
            assert np.array_equal(ref.pixel_coords(s, k, x).view(np.uint32), out.view(np.uint32))


# ---- adjacency graph and MRF model (build_adjacency_graph.cpp, view_selection.cpp) ------------------------------
@pytest.mark.parametrize("name", ["tiny", "small", "occ", "C2s", "messy"])
def test_adjacency_matches_reference_tu(ref, scene_mod, get_scene, name):
    """tex::build_adjacency_graph (:16-53) on the reference's UniGraph vs scene.face_adjacency (what the oracle and the
    C ABI are fed): same neighbours in the same adjacency-list order, borders (C2s) and separate components (occ) included."""
    s = get_scene(name)
    rings = scene_mod.vertex_rings(s.faces, s.verts.shape[0])
    r_ptr, r_idx = ref.build_adjacency(s.faces, s.verts.shape[0], rings)
    a_ptr, a_idx = scene_mod.face_adjacency(s.faces)
    assert np.array_equal(r_ptr, a_ptr) and np.array_equal(r_idx, a_idx)


@pytest.mark.parametrize("name", ["tiny", "occ"])
def test_mrf_model_and_label_decoding_match_reference_tu(ref, orc, scene_mod, get_scene, name):
    """view_selection.cpp:26-82 builds the model, :120-131 decodes the solution.  With the recording mapMAP shim: edges
    only between seen faces (i < j, weight 1), label set = view id + 1 in DataCosts column order, unary = data cost, unseen
    faces get the single label 0 with cost 1, Potts weight 1, StopWhenReturnsDiminish(5, 0.01), deterministic seed
    548923723.  The oracle's energy function must be the energy of exactly this model."""
    s = get_scene(name)
    dc = orc.data_costs(s)
    if name == "tiny":                                           # make a few faces unseen
        keep = np.ones(len(dc["view"]), bool)
        ptr = dc["face_ptr"].astype(np.int64)
        for f in (3, 17, 18, 200):
            keep[ptr[f]:ptr[f + 1]] = False
        cnt = np.add.reduceat(keep.astype(np.int64), ptr[:-1]) * (ptr[1:] > ptr[:-1])
        dc = dict(face_ptr=np.r_[0, np.cumsum(cnt)].astype(np.uint64), view=dc["view"][keep], cost=dc["cost"][keep])
    adj = scene_mod.face_adjacency(s.faces)
    m = ref.view_selection_model(adj, dc["face_ptr"], dc["view"], dc["cost"], s.num_views)
    F = s.num_faces
    ptr = dc["face_ptr"].astype(np.int64)
    seen = ptr[1:] > ptr[:-1]
    assert (~seen).sum() >= (4 if name == "tiny" else 0)
    # edges
    exp = [(i, int(j)) for i in range(F) if seen[i] for j in adj[1][adj[0][i]:adj[0][i + 1]] if i < j and seen[j]]
    assert [tuple(e) for e in m["edges"].tolist()] == exp
    # label sets and unaries
    lp = m["ls_ptr"].astype(np.int64)
    for f in range(F):
        ll, lc = m["ls_label"][lp[f]:lp[f + 1]], m["ls_cost"][lp[f]:lp[f + 1]]
        if seen[f]:
            assert np.array_equal(ll, dc["view"][ptr[f]:ptr[f + 1]].astype(np.int32) + 1)
            assert np.array_equal(lc.view(np.uint32), dc["cost"][ptr[f]:ptr[f + 1]].view(np.uint32))
        else:
            assert ll.tolist() == [0] and lc.tolist() == [1.0]
    p = m["params"]
    assert p["potts"] == 1.0 and p["window"] == 5 and abs(p["ratio"] - 0.01) < 1e-12
    assert p["seed"] == 548923723 and p["deterministic"] == 1 and p["model_complete"] == 1 and p["components_updated"] == 1
    assert p["use_multilevel"] == 1 and p["use_spanning_tree"] == 1 and p["use_acyclic"] == 1 and p["force_acyclic"] == 1
    # decoding of the shim's solution (cheapest label per node, first minimum)
    exp_labels = np.zeros(F, np.uint32)
    for f in range(F):
        if seen[f]:
            c = dc["cost"][ptr[f]:ptr[f + 1]]
            exp_labels[f] = int(dc["view"][ptr[f] + int(np.argmin(c))]) + 1
    assert np.array_equal(m["labels"], exp_labels)
    # the oracle's objective = energy of the recorded model, for the greedy and for the optimised labeling
    o = orc.view_selection(adj[0], adj[1], dc["face_ptr"], dc["view"], dc["cost"], threads=1)
    def model_energy(labels):
        e = 0.0
        for f in range(F):
            ll = m["ls_label"][lp[f]:lp[f + 1]]
            k = int(np.flatnonzero(ll == labels[f])[0])
            e += float(m["ls_cost"][lp[f] + k])
        e += sum(1.0 for a, b in m["edges"] if labels[a] != labels[b])
        return e
    for lab in (exp_labels, o["labels"]):
        e_model = model_energy(lab)
        e_orc = orc.mrf_energy(adj[0], adj[1], dc["face_ptr"], dc["view"], dc["cost"], lab)
        assert abs(e_model - e_orc) < 1e-6 * max(1.0, e_model)
    assert model_energy(o["labels"]) < model_energy(exp_labels)


# ---- texture patches, global and local seam leveling ---------------------------------------------------------------
@pytest.fixture(scope="module")
def seam_inputs(orc, scene_mod, get_scene):
    cache = {}
    def _get(name):
        if name not in cache:
            s = get_scene(name)
            adj = scene_mod.face_adjacency(s.faces)
            rings = scene_mod.vertex_rings(s.faces, s.verts.shape[0])
            dc = orc.data_costs(s)
            labels = orc.view_selection(adj[0], adj[1], dc["face_ptr"], dc["view"], dc["cost"], threads=1)["labels"]
            cache[name] = (s, adj, rings, labels)
        return cache[name]
    return _get


@pytest.mark.parametrize("name", ["tiny", "occ", "messy"])
def test_texture_patches_match_reference_tu(ref, orc, seam_inputs, name):
    """generate_texture_patches.cpp:453-538 (+ generate_candidate :78-138, merge_vertex_projection_infos :40-65) and the
    zero-adjust pass texrecon.cpp:174-183 (TexturePatch::adjust_colors, texture_patch.cpp:41-116) vs oracle/patches.py:
    same patches, faces, bit-identical texcoords, vertex projections, images, validity and blending masks."""
    import patches as P
    s, adj, rings, labels = seam_inputs(name)
    rp, rvpi = ref.seam_leveling(s, rings, adj, labels, do_global=False)
    pp, pvpi = P.generate_texture_patches(orc, s, adj, labels)
    rp = [p for p in rp if p.label != 0]                       # label 0 = hole-filling patches (not restated)
    assert len(rp) == len(pp) >= s.num_views // 2
    for a, b in zip(rp, pp):
        assert a.label == b.label and a.faces == b.faces
        assert np.array_equal(a.texcoords.view(np.uint32), np.asarray(b.texcoords, np.float32).view(np.uint32))
        img, validity, blending = P.adjust_colors(b, np.zeros((3 * len(b.faces), 3), np.float32))
        assert np.array_equal(a.validity, validity) and np.array_equal(a.blending, blending)
        assert np.array_equal(a.image.view(np.uint32), img.view(np.uint32))
    for v in range(s.verts.shape[0]):
        mine = {pid: proj for pid, (proj, _f) in pvpi[v].items()}
        theirs = {pid: xy for pid, xy in rvpi[v].items() if pid < len(pp)}
        assert set(mine) == set(theirs)
        for pid in mine:
            assert np.array_equal(np.asarray(mine[pid], np.float32).view(np.uint32), theirs[pid].view(np.uint32))


@pytest.mark.parametrize("name", ["tiny", "occ", "messy"])
def test_global_seam_leveling_matches_reference_tu(ref, orc, seam_inputs, name):
    """tex::global_seam_leveling (global_seam_leveling.cpp:140-324: unknown numbering, Gamma, A, b from patch-relative
    edge samples, Lhs, CG per channel, mean subtraction, adjust_colors per patch) vs orc_global_seam_leveling +
    oracle/patches.apply_adjust_values.  The CG of both sides is the same restatement of Eigen's (shim / seam.c); the
    right-hand sides are sampled in patch vs view coordinates, so the adjusted images agree to rounding (2e-5)."""
    import patches as P
    s, adj, rings, labels = seam_inputs(name)
    seam = orc.global_seam_leveling(s, rings, labels)
    rp, _ = ref.seam_leveling(s, rings, adj, labels, do_global=True)
    pp, _ = P.generate_texture_patches(orc, s, adj, labels)
    pa = P.apply_adjust_values(s, pp, seam["row_ptr"], seam["row_label"], seam["x"])
    rp = [p for p in rp if p.label != 0]
    assert len(rp) == len(pa)
    moved = 0.0
    for a, b, raw in zip(rp, pa, pp):
        assert np.array_equal(a.validity, b.validity) and np.array_equal(a.blending, b.blending)
        assert np.abs(a.image - b.image).max() < 2e-5
        moved = max(moved, float(np.abs(a.image - raw.image)[a.validity != 0].max()))
    assert moved > 0.02                                        # the leveling really changed the colours


def test_local_seam_leveling_matches_reference_tu(ref, orc, seam_inputs):
    """tex::local_seam_leveling (local_seam_leveling.cpp:105-204, draw_line :39-92, prepare_blending_mask
    texture_patch.cpp:197-297, poisson_blend poisson_blending.cpp:49-138) vs oracle/patches.local_seam_leveling.
    Both solve the same fp32 systems with a direct solver in double (shim elimination / scipy splu): 2e-5."""
    import patches as P
    s, adj, rings, labels = seam_inputs("tiny")
    seam = orc.global_seam_leveling(s, rings, labels)
    rp, _ = ref.seam_leveling(s, rings, adj, labels, do_global=True, do_local=True)
    pp, pvpi = P.generate_texture_patches(orc, s, adj, labels)
    pa = P.apply_adjust_values(s, pp, seam["row_ptr"], seam["row_label"], seam["x"])
    before = [p.image.copy() for p in pa]
    P.local_seam_leveling(s, adj, labels, pa, pvpi)
    changed = 0.0
    for a, b, b0 in zip(rp, pa, before):
        assert np.array_equal(a.validity, b.validity)
        assert np.abs(a.image - b.image).max() < 2e-5
        changed = max(changed, float(np.abs(b.image - b0).max()))
    assert changed > 0.01


def test_adjacency_of_non_manifold_mesh_matches_reference_tu(ref, scene_mod, get_scene):
    """An edge shared by three faces (a fin glued onto `tiny`): tex::build_adjacency_graph links all of them; the order of
    the adjacency lists is what scene.face_adjacency has to reproduce (face graphs with degree > 3 take the generic paths of
    the MRF kernels)."""
    s = get_scene("tiny")
    verts = np.concatenate([s.verts, (s.verts[s.faces[5]].mean(0) * 1.3)[None].astype(np.float32),
                            (s.verts[s.faces[40]].mean(0) * 1.3)[None].astype(np.float32)], 0)
    nv = verts.shape[0]
    fins = np.array([[s.faces[5][0], s.faces[5][1], nv - 2], [s.faces[5][1], s.faces[5][2], nv - 2],
                     [s.faces[40][2], s.faces[40][0], nv - 1]], np.uint32)
    faces = np.ascontiguousarray(np.concatenate([s.faces, fins], 0))
    rings = scene_mod.vertex_rings(faces, nv)
    r_ptr, r_idx = ref.build_adjacency(faces, nv, rings)
    a_ptr, a_idx = scene_mod.face_adjacency(faces)
    assert int(np.diff(r_ptr).max()) > 3
    assert np.array_equal(r_ptr, a_ptr) and np.array_equal(r_idx, a_idx)
