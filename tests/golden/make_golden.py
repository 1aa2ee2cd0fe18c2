"""Generates tests/golden/oracle_snapshots.json: regression snapshots of the oracle on the seeded
synthetic scenes.  The reference ships no golden vectors and cannot be built here,
so these pin the ORACLE (and, through the -m gpu parity tests, the CUDA path) against drift; the
analytic known-answer tests in tests/test_oracle_cpu.py pin the oracle to the reference's semantics.

    python tests/golden/make_golden.py
"""
import importlib
import json
import os
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402

scene = importlib.import_module("mvs-texturing_b200.scene")
out = {}
for name in ["tiny", "small", "C1", "C1d", "C2s", "C3s", "occ", "occ2", "messy", "C5s"]:   # occ*: real occlusion, unseen faces, several
    # components; messy: non-manifold fins, zero-area faces, a sliver, a detached triangle; C5s: 92 candidate views per face
    # (32 lanes per node in the MRF sweeps)
    s = scene.config(name)
    dc = O.data_costs(s)
    ap, ai = scene.face_adjacency(s.faces)
    m = O.view_selection(ap, ai, dc["face_ptr"], dc["view"], dc["cost"], threads=1)
    rings = scene.vertex_rings(s.faces, s.verts.shape[0])
    sm = O.global_seam_leveling(s, rings, m["labels"])
    out[name] = dict(nnz=len(dc["view"]), crc_face_ptr=zlib.crc32(dc["face_ptr"].tobytes()),
                     crc_view=zlib.crc32(dc["view"].tobytes()), crc_cost=zlib.crc32(dc["cost"].tobytes()),
                     max_quality=dc["max_quality"], percentile=dc["percentile"],
                     mrf_iterations=m["iterations"], crc_labels=zlib.crc32(m["labels"].tobytes()),
                     mrf_energy=m["energy"], mrf_energy_fixed=O.mrf_energy_fixed(ap, ai, dc["face_ptr"], dc["view"], dc["cost"], m["labels"]),
                     seam_rows=len(sm["row_label"]), seam_a_rows=sm["num_a_rows"],
                     cg_iterations=list(sm["iterations"]))
    if name in ("tiny", "occ", "messy"):   # texture patches + adjust_colors with zero offsets (oracle/patches.py, pinned to the reference TUs)
        import numpy as np
        import patches as P
        pp, _ = P.generate_texture_patches(O, s, (ap, ai), m["labels"])
        crc = dict(tex=0, img=0, val=0, bl=0)
        for q in pp:
            img, val, bl = P.adjust_colors(q, np.zeros((3 * len(q.faces), 3), np.float32))
            crc["tex"] = zlib.crc32(np.ascontiguousarray(q.texcoords, np.float32).tobytes(), crc["tex"])
            crc["img"] = zlib.crc32(np.ascontiguousarray(img).tobytes(), crc["img"])
            crc["val"] = zlib.crc32(np.ascontiguousarray(val).tobytes(), crc["val"])
            crc["bl"] = zlib.crc32(np.ascontiguousarray(bl).tobytes(), crc["bl"])
        out[name].update(patches=len(pp), patch_faces=sum(len(q.faces) for q in pp), crc_patch_texcoords=crc["tex"],
                         crc_patch_images=crc["img"], crc_patch_validity=crc["val"], crc_patch_blending=crc["bl"])
    print(name, out[name])
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "oracle_snapshots.json"), "w"), indent=1)
