"""File-level boundary (INTEGRATION.md route B): .spt and .vec byte formats of the reference
(libs/tex/sparse_table.h:112-187, libs/tex/util.h:104-131)."""
import importlib
import struct

import numpy as np
import pytest


@pytest.fixture(scope="module")
def ic():
    return importlib.import_module("mvs-texturing_b200.interchange")


def test_spt_byte_layout_matches_reference_writer(ic, tmp_path):
    # 3 faces, 4 views: face 0 -> {(1, .25)}, face 1 -> {}, face 2 -> {(0, .5), (3, .75)}
    fp = np.array([0, 1, 1, 3], np.uint64)
    p = tmp_path / "t_data_costs.spt"
    ic.save_data_costs(p, fp, np.array([1, 0, 3], np.uint16), np.array([.25, .5, .75], np.float32), 4)
    raw = p.read_bytes()
    header, body = raw.split(b"\n", 1)
    assert header == b"SPT 0.2 3 4 3"                     # "SPT" " " "0.2" " " cols " " rows " " nnz endl
    assert len(body) == 3 * 10                            # sizeof(u32)+sizeof(u16)+sizeof(f32), packed
    assert struct.unpack("<IHf", body[:10]) == (0, 1, 0.25)
    assert struct.unpack("<IHf", body[10:20]) == (2, 0, 0.5)
    assert struct.unpack("<IHf", body[20:30]) == (2, 3, 0.75)


def test_spt_round_trip_and_dimension_check(ic, orc, get_scene, tmp_path):
    s = get_scene("small")
    dc = orc.data_costs(s)
    p = tmp_path / "small_data_costs.spt"
    ic.save_data_costs(p, dc["face_ptr"], dc["view"], dc["cost"], s.num_views)
    fp, vw, cs, rows = ic.load_data_costs(p, s.num_faces, s.num_views)
    assert rows == s.num_views
    assert np.array_equal(fp, dc["face_ptr"]) and np.array_equal(vw, dc["view"])
    assert np.array_equal(cs.view(np.uint32), dc["cost"].view(np.uint32))
    with pytest.raises(ic.FileException, match="different dimension"):
        ic.load_data_costs(p, s.num_faces + 1, s.num_views)
    bad = tmp_path / "bad.spt"
    bad.write_bytes(b"XYZ 0.2 1 1 0\n")
    with pytest.raises(ic.FileException, match="Not a SparseTable"):
        ic.load_data_costs(bad)
    bad.write_bytes(b"SPT 0.1 1 1 0\n")
    with pytest.raises(ic.FileException, match="Incompatible version"):
        ic.load_data_costs(bad)


def test_spt_row_major_input_is_regrouped(ic, tmp_path):
    # a writer that calls set_value in view-major order still loads into CSR by face
    p = tmp_path / "rm.spt"
    with open(p, "wb") as f:
        f.write(b"SPT 0.2 2 2 3\n")
        for col, row, val in [(1, 0, .1), (0, 1, .2), (1, 1, .3)]:
            f.write(struct.pack("<IHf", col, row, val))
    fp, vw, cs, rows = ic.load_data_costs(p, 2, 2)
    assert fp.tolist() == [0, 1, 3] and vw.tolist() == [1, 0, 1]
    assert np.allclose(cs, [.2, .1, .3])


def test_labeling_vec_format(ic, tmp_path):
    p = tmp_path / "x_labeling.vec"
    ic.save_labeling(p, np.array([0, 3, 1, 2], np.uint32))
    raw = p.read_bytes()
    assert len(raw) == 4 * 8 and struct.unpack("<4Q", raw) == (0, 3, 1, 2)   # raw std::size_t[F]
    assert ic.load_labeling(p, 4, 3).tolist() == [0, 3, 1, 2]
    with pytest.raises(ic.FileException):
        ic.load_labeling(p, 5, 3)                          # wrong face count
    with pytest.raises(ic.FileException):
        ic.load_labeling(p, 4, 2)                          # label 3 > 2 views


def test_timings_csv_format(tmp_path):
    """Timer::write_to_file (libs/tex/timer.cpp:42-61) with texrecon's event names (texrecon.cpp:86-211)"""
    ic = importlib.import_module("mvs-texturing_b200.interchange")
    log = ic.TimingLog(header="b200")
    for i, name in enumerate(ic.TIMING_EVENTS):
        log.measure(name, abs_ms=10 * (i + 1) ** 2, abs_clocks=1000 * (i + 1))
    p = str(tmp_path / "OUT_timings.csv")
    log.write_to_file(p)
    lines = open(p).read().splitlines()
    assert lines[0] == "#b200"
    assert lines[1] == "Event, Absolute clocks, Absolute milliseconds, Relative clocks, Relative milliseconds"
    assert lines[2] == "Loading, 1000, 10, 1000, 10"
    assert lines[3] == "Calculating data costs, 2000, 40, 1000, 30"
    assert len(lines) == 2 + len(ic.TIMING_EVENTS) and lines[-1].startswith("Total, ")
    assert [r[0] for r in ic.load_timings(p)] == list(ic.TIMING_EVENTS)


def test_cam_file_round_trip(tmp_path, scene_mod):
    """.cam as generate_texture_views.cpp:118-151 reads it; a synthetic camera survives save -> load and projects a point
    to the same pixel as the scene's own TextureView record."""
    ic = importlib.import_module("mvs-texturing_b200.interchange")
    s = scene_mod.config("tiny", with_images=False)
    k = 2
    w2c = s.w2c[k].reshape(4, 4)
    proj = s.proj[k].reshape(3, 3)
    flen = float(proj[0, 0]) / max(s.width, s.height)          # MVE normalises by the larger side
    p = str(tmp_path / "view.cam")
    ic.save_cam(p, w2c, flen)
    cam = ic.load_cam(p, s.width, s.height)
    assert np.allclose(cam["w2c"], s.w2c[k], atol=1e-6) and np.allclose(cam["pos"], s.pos[k], atol=1e-5)
    assert np.allclose(cam["viewdir"], s.viewdir[k], atol=1e-6) and np.allclose(cam["proj"], s.proj[k], rtol=1e-5)
    with open(p, "w") as f:
        f.write("1 2 3\n0.9\n")
    with pytest.raises(ic.FileException):
        ic.load_cam(p, 640, 480)


@pytest.mark.parametrize("binary", [True, False])
def test_ply_round_trip(tmp_path, scene_mod, binary):
    ic = importlib.import_module("mvs-texturing_b200.interchange")
    s = scene_mod.config("tiny", with_images=False)
    p = str(tmp_path / "mesh.ply")
    ic.save_ply(p, s.verts, s.faces, binary=binary)
    v, f = ic.load_ply(p)
    assert np.array_equal(f, s.faces)
    assert np.array_equal(v, s.verts) if binary else np.allclose(v, s.verts, atol=1e-7)
    with open(p, "wb") as fh:
        fh.write(b"plx\n")
    with pytest.raises(ic.FileException):
        ic.load_ply(p)
