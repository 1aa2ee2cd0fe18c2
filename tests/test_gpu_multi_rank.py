"""Multi-rank paths on ONE GPU: `ranks` contexts on device 0, each driven by its own host thread, run the same kernels and
the same peer-memory protocol as one process per GPU does (mvs-texturing_b200/sharded.py) -- boundary-label halo pushes,
epoch-flag barriers, energy slots (csrc/mrf.cu) and the fused PCG with the search-direction exchange inside the kernel
(csrc/seam_mg.cu).  The peers live in one process here, so they attach each other's blocks by raw device pointer
(b2tex_peer_attach) instead of a cudaIpc handle; everything behind that is identical.  The scenes are small, so the
persistent kernels of all ranks are co-resident (a spinning kernel never keeps a peer's kernel off the SMs), and at most
four ranks, so that every rank's stream has a hardware work queue of its own (CUDA_DEVICE_MAX_CONNECTIONS defaults to 8:
with eight rank streams plus the framework's own, two ranks share a queue and a barrier kernel at its head keeps the
peer's kernel behind it from ever starting -- one process per GPU, the real deployment, has no such coupling).

Bars: labels, iteration count and fixed-point energy equal the oracle run with num_parts = ranks (bit exact); every rank
holds every label after the final all-gather; all ranks end with the bit-identical seam solution, within 5e-3 of the
oracle's (same bar as the single-GPU PCG)."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run_threads(fns):
    out, err = [None] * len(fns), [None] * len(fns)

    def wrap(i):
        try:
            out[i] = fns[i]()
        except BaseException as e:  # noqa: BLE001
            err[i] = e

    th = [threading.Thread(target=wrap, args=(i,)) for i in range(len(fns))]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in th), "a rank did not return (barrier protocol hang)"
    for e in err:
        if e is not None:
            raise e
    return out


@pytest.mark.parametrize("name,ranks", [("occ", 2), ("C2s", 4), ("C1d", 3)])
def test_sharded_pipeline_on_one_gpu(b2, orc, scene_mod, get_scene, name, ranks):
    s = get_scene(name)
    adj = scene_mod.face_adjacency(s.faces)
    rings = scene_mod.vertex_rings(s.faces, s.verts.shape[0])
    F = s.num_faces
    psz = (F + ranks - 1) // ranks
    ctxs = [b2.Context(0) for _ in range(ranks)]
    try:
        for r, c in enumerate(ctxs):
            c.set_scene(s)
            c.set_adjacency(*adj)
            c.set_vertex_rings(*rings)
            c.set_face_range(min(F, r * psz), min(F, (r + 1) * psz))
        # ---- data costs: own faces, global normalisation (calculate_data_costs.cpp:277-302) ----
        infos = [c.data_costs_qualities() for c in ctxs]
        gmax = max(float(i.max_quality) for i in infos)
        bins = np.zeros(10000, np.uint64)
        for c in ctxs:
            bins += c.data_costs_histogram(gmax)
        for c in ctxs:
            c.data_costs_normalize(gmax, bins.astype(np.uint32))
        o = orc.data_costs(s)
        om = orc.view_selection(adj[0], adj[1], o["face_ptr"], o["view"], o["cost"], threads=1, num_parts=ranks)
        # ---- view selection: peers attached once, then every rank runs the single-GPU entry point ----
        for r, c in enumerate(ctxs):
            c.mrf_mg_export(r, ranks)
        blocks = [c.peer_block(0) for c in ctxs]
        for r, c in enumerate(ctxs):
            for k in range(ranks):
                if k != r:
                    c.peer_attach(0, k, blocks[k])
        for c in ctxs:      # all device allocations before any rank can sit in a barrier kernel (one process, one device)
            c.view_selection_prepare(num_parts=ranks)
        res = _run_threads([(lambda c=c: c.view_selection_run(num_parts=ranks)) for c in ctxs])
        for (info, trace), c in zip(res, ctxs):
            assert info.iterations == om["iterations"]
            assert np.array_equal(c.labels_download(), om["labels"])        # own range, halo and the final all-gather
            assert abs(info.energy_final - om["energy"]) <= 1e-6 * max(1.0, om["energy"])
            assert np.array_equal(trace, res[0][1])                          # identical sums on every rank
        assert int(orc.mrf_energy_fixed(adj[0], adj[1], o["face_ptr"], o["view"], o["cost"], om["labels"])) == \
            int(round(res[0][0].energy_final * 4294967296.0))
        # a second run on the same peers (epochs continue, blocks are reused)
        res2 = _run_threads([(lambda c=c: c.view_selection_run(num_parts=ranks)) for c in ctxs])
        assert all(r2[0].iterations == om["iterations"] for r2 in res2)
        assert np.array_equal(ctxs[-1].labels_download(), om["labels"])
        # ---- global seam leveling: replicated assembly, rows of the PCG split, exchange inside the kernel ----
        og = orc.global_seam_leveling(s, rings, om["labels"])
        seams = [c.seam_assemble() for c in ctxs]
        assert all(int(si.num_rows) == len(og["row_label"]) for si in seams)
        for r, c in enumerate(ctxs):
            c.seam_mg_export(r, ranks)
        blocks = [c.peer_block(1) for c in ctxs]
        for r, c in enumerate(ctxs):
            for k in range(ranks):
                if k != r:
                    c.peer_attach(1, k, blocks[k])
        _run_threads([(lambda c=c, si=si: c.seam_mg_solve(si)) for c, si in zip(ctxs, seams)])
        xs = [c.seam_download(si)["x"] for c, si in zip(ctxs, seams)]
        for x, si in zip(xs, seams):
            assert np.array_equal(x.view(np.uint32), xs[0].view(np.uint32))
            assert list(si.iterations) == list(seams[0].iterations)
        rel = np.linalg.norm(xs[0] - og["x"]) / max(1e-30, np.linalg.norm(og["x"]))
        assert rel < 5e-3, rel
        assert max(abs(int(a) - int(b)) for a, b in zip(seams[0].iterations, og["iterations"])) <= 3
    finally:
        for c in ctxs:
            c.close()
