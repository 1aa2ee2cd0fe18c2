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


@pytest.mark.parametrize("name", ["tiny", "small", "occ"])
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
