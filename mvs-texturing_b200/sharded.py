"""Host-side driver of the resident pipeline, one process per GPU.

world == 1 : data costs -> view selection -> seam leveling on one context.
world  > 1 : faces are split into `world` contiguous ranges (rank r owns [r*psz, (r+1)*psz)):
    data costs : each rank evaluates its own faces; the normalisation is global, so
                 qualities -> all_reduce(MAX) -> histogram -> all_reduce(SUM, 10 000 bins) -> normalise
                 (calculate_data_costs.cpp:277-302)
    MRF        : each rank runs the forest-BCD iteration on its own nodes with cut edges conditioned
                 (the schedule of oracle/mrf.c with num_parts = world); after every iteration the
                 label ranges are all-gathered and the 32.32 fixed-point energy is all-reduced for the
                 StopWhenReturnsDiminish test (view_selection.cpp:84)
    seam       : replicated on every rank (row-partitioned PCG with a halo exchange is not built yet)
torch.distributed (NCCL) is plumbing only; every kernel is in libb2tex.so.
"""
from __future__ import annotations

import os
import time

import numpy as np


def returns_diminish(efix, t, window, ratio):
    """StopWhenReturnsDiminish(window, ratio) (view_selection.cpp:84) on 32.32 fixed-point energies;
    same double arithmetic as b2tex_view_selection_run and oracle/mrf.c."""
    if t < window:
        return False
    e0, e1 = float(efix[t - window]), float(efix[t])
    return e0 <= 0.0 or (e0 - e1) / e0 < ratio


def gather_label_ranges(dist, labels, mine, gathered, fb, fe, F):
    """All-gather the owned label range of every rank into the full label array (in place).
    `mine` (psz) and `gathered` (world*psz) are scratch tensors on the same device as `labels`."""
    mine[: fe - fb].copy_(labels[fb:fe])
    dist.all_gather_into_tensor(gathered, mine)
    labels.copy_(gathered[:F])
    return labels


class _DevArray:
    """Expose a raw device pointer of the library to torch (zero copy)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def describe_parallelism(world):
    if world == 1:
        return "1 GPU, whole scene resident"
    return (f"{world} GPUs, one process each: faces split in {world} contiguous ranges; data costs: NCCL all-reduce of the "
            f"maximum and of the 10 000 histogram bins; view selection: boundary-label halo + energy all-reduce through "
            f"NVLink peer memory inside the device loop (no host round trip per iteration); seam leveling: assembly "
            f"replicated, PCG rows split, search direction halo + dot products through peer memory inside one "
            f"persistent kernel per GPU")


def _exchange_handles(torch, dist, dev, world, rank, mine: bytes):
    """All-gather of one 64-byte cudaIpc handle per rank."""
    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(dev)
    allh = [torch.empty(64, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(allh, t)
    return [bytes(h.cpu().numpy().tobytes()) for h in allh]


class ShardedPipeline:
    def __init__(self, b2, scene, adj, rings, rank=0, world=1, local_rank=0, upload=True, settings=None):
        self.b2, self.scene, self.adj, self.rings = b2, scene, adj, rings
        self.rank, self.world = rank, world
        # tex::Settings fields the data-cost stage reads: dict(data_term=, visibility=, outlier_removal=); None = reference defaults
        self.settings = settings
        self.F = scene.num_faces
        self.psz = (self.F + world - 1) // world
        self.fb = min(self.F, rank * self.psz)
        self.fe = min(self.F, (rank + 1) * self.psz)
        self.ctx = b2.Context(local_rank)
        # B2TEX_MRF_NCCL=1 / B2TEX_SEAM_P2P=0: the round-1 host-driven paths (label all-gather over NCCL, replicated solve)
        self.mrf_peer = world > 1 and world <= 8 and os.environ.get("B2TEX_MRF_NCCL", "0") != "1"
        self.seam_peer = world > 1 and world <= 8 and os.environ.get("B2TEX_SEAM_P2P", "1") != "0"
        self._mrf_peers_ready = False
        if upload:
            self.upload()

    def upload(self, scene=None, adj=None, rings=None):
        c = self.ctx
        c.set_scene(scene if scene is not None else self.scene)
        c.set_adjacency(*(adj if adj is not None else self.adj))
        c.set_vertex_rings(*(rings if rings is not None else self.rings))
        if self.world > 1:
            c.set_face_range(self.fb, self.fe)
            if self.mrf_peer and not self._mrf_peers_ready:
                self._attach_mrf_peers()

    def _attach_mrf_peers(self):
        """One-time: every rank allocates the peer-visible label block and maps those of its peers (cudaIpc)."""
        import torch
        import torch.distributed as dist
        dev = torch.device("cuda", torch.cuda.current_device())
        handles = _exchange_handles(torch, dist, dev, self.world, self.rank, self.ctx.mrf_mg_export(self.rank, self.world))
        for k in range(self.world):
            if k != self.rank:
                self.ctx.mrf_mg_import(k, handles[k])
        dist.barrier()      # every block is mapped before the first store into it
        self._mrf_peers_ready = True

    def describe(self):
        return describe_parallelism(self.world)

    # ---- one pass of the hot path ---------------------------------------------------------------
    def step(self):
        if self.world == 1:
            return self._step_single()
        return self._step_sharded()

    def _step_single(self):
        c = self.ctx
        t0 = time.perf_counter()
        dc = c.data_costs_run(**(self.settings or {}))
        t1 = time.perf_counter()
        mrf, trace = c.view_selection_run()
        t2 = time.perf_counter()
        if os.environ.get("B2TEX_SEAM_MG1", "0") == "1":   # experiment: the multi-GPU kernel on one rank
            seam = c.seam_assemble()
            if getattr(self, "_mg_rows", None) != int(seam.num_rows):
                c.seam_mg_export(0, 1)
                self._mg_rows = int(seam.num_rows)
            c.seam_mg_solve(seam)
        else:
            seam = c.seam_run()
        t3 = time.perf_counter()
        return dict(dc=dc, mrf=mrf, seam=seam, trace=trace,
                    stage_s=dict(data_costs=t1 - t0, view_selection=t2 - t1, seam_leveling=t3 - t2))

    def _step_sharded(self):
        import torch
        import torch.distributed as dist
        c, b2 = self.ctx, self.b2
        dev = torch.device("cuda", torch.cuda.current_device())
        t0 = time.perf_counter()
        info = c.data_costs_qualities(**(self.settings or {}))
        gmax = torch.tensor([info.max_quality], dtype=torch.float32, device=dev)
        dist.all_reduce(gmax, op=dist.ReduceOp.MAX)
        gmax_f = float(gmax.item())
        c.data_costs_histogram_device(gmax_f)
        c.synchronize()
        hp, hn = c.device_ptr("hist")
        hist = torch.as_tensor(_DevArray(hp, hn, "<i4"), device=dev)
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
        bins = hist.cpu().numpy().astype(np.uint32)
        cand, rays = info.candidates, info.rays
        info = c.data_costs_normalize(gmax_f, bins)
        info.candidates, info.rays = cand, rays
        self._own_nnz = int(info.nnz)
        t1 = time.perf_counter()

        P, F = self.world, self.F
        if self.mrf_peer:
            # the whole loop runs on the device; the ranks meet inside the kernels (csrc/mrf.cu: k_halo_push, k_mg_sync)
            mrf, trace = c.view_selection_run(num_parts=P)
        else:
            mrf, trace = self._mrf_nccl(torch, dist, dev)
        t2 = time.perf_counter()
        if self.seam_peer:
            seam = self._seam_p2p(torch, dist, dev)
        else:
            seam = c.seam_run()
        t3 = time.perf_counter()
        return dict(dc=info, mrf=mrf, seam=seam, trace=trace,
                    stage_s=dict(data_costs=t1 - t0, view_selection=t2 - t1, seam_leveling=t3 - t2))

    def total_nnz(self):
        """DataCosts entries over all ranks (one all-reduce; outside the timed step)."""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return None
        counts = torch.tensor([self._own_nnz], dtype=torch.int64, device="cuda")
        dist.all_reduce(counts)
        return int(counts.item())

    def _mrf_nccl(self, torch, dist, dev):
        """Round-1 path (B2TEX_MRF_NCCL=1): host-driven loop, the whole label array all-gathered every iteration."""
        c, b2 = self.ctx, self.b2
        P, psz, F = self.world, self.psz, self.F
        p = b2.mrf_params(num_parts=P)
        ratio = float(np.float32(p.ratio))
        c.mrf_init(num_parts=P)
        lp, ln = c.device_ptr("labels")
        labels = torch.as_tensor(_DevArray(lp, ln, "<i4"), device=dev)
        gathered = torch.zeros(P * psz, dtype=torch.int32, device=dev)
        mine = torch.zeros(psz, dtype=torch.int32, device=dev)

        def exchange():
            gather_label_ranges(dist, labels, mine, gathered, self.fb, self.fe, F)
            torch.cuda.current_stream().synchronize()

        # the energy of cut edges needs the neighbours' NEW labels: evaluate it after every exchange
        exchange()
        efix = [self._allreduce_energy(c.mrf_energy())]
        t = 1
        while t <= p.max_iterations:
            c.mrf_iterate(t)
            exchange()
            efix.append(self._allreduce_energy(c.mrf_energy()))
            if returns_diminish(efix, t, p.window, ratio):
                break
            t += 1
        t = min(t, p.max_iterations)
        mrf = b2.B2MrfInfo()
        mrf.iterations = t
        mrf.energy_initial = efix[0] / 4294967296.0
        mrf.energy_final = efix[t] / 4294967296.0
        mrf.sweep_bytes = 14 * self._own_nnz + 20 * F
        return mrf, np.array(efix) / 4294967296.0

    def _seam_p2p(self, torch, dist, dev):
        """Row-partitioned PCG with the exchange inside the kernel (csrc/seam_mg.cu): every rank assembles, the cudaIpc
        handles of the peer blocks go round once per system size (all-gather of 64 bytes), then each GPU runs one fused
        compute + exchange kernel."""
        c = self.ctx
        seam = c.seam_assemble()
        if getattr(self, "_mg_rows", None) != int(seam.num_rows):   # peer blocks are kept while the system size stays
            handles = _exchange_handles(torch, dist, dev, self.world, self.rank, c.seam_mg_export(self.rank, self.world))
            for k in range(self.world):
                if k != self.rank:
                    c.seam_mg_import(k, handles[k])
            self._mg_rows = int(seam.num_rows)
            dist.barrier()       # every peer block is mapped before anybody stores into it
        c.seam_mg_solve(seam)
        return seam

    def _allreduce_energy(self, e):
        import torch
        import torch.distributed as dist
        t = torch.tensor([e], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        return int(t.item())

    # ---- e2e for the sharded configuration: upload from host every step ----------------------------
    def e2e(self, torch, steps=1, warmup=1):
        import torch.distributed as dist
        s = self.scene
        h2d = (s.verts.nbytes + s.faces.nbytes + s.face_normals.nbytes + s.images.nbytes + self.adj[0].nbytes
               + self.adj[1].nbytes + sum(r.nbytes for r in self.rings))
        # host buffers pinned once (like e2e_host_path): what travels per step is the data, not the page-locking
        if getattr(self, "_pinned", None) is None:
            import copy

            def pin(a):
                t = torch.from_numpy(np.ascontiguousarray(a))
                try:
                    t = t.pin_memory()
                except Exception:
                    pass
                return t.numpy(), t
            keep, sp = [], copy.copy(s)
            for name in ("verts", "faces", "face_normals", "images"):
                arr, t = pin(getattr(s, name)); setattr(sp, name, arr); keep.append(t)
            padj, prings = [], []
            for a in self.adj:
                arr, t = pin(a); padj.append(arr); keep.append(t)
            for a in self.rings:
                arr, t = pin(a); prings.append(arr); keep.append(t)
            self._pinned = (sp, tuple(padj), tuple(prings), keep)
        sp, padj, prings, _ = self._pinned
        times = []
        d2h = 0
        for i in range(warmup + steps):
            torch.cuda.synchronize()
            if self.world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            self.upload(sp, padj, prings)
            res = self.step()
            labels = self.ctx.labels_download()
            x = self.ctx.seam_download(res["seam"])["x"]
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if self.world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            if i >= warmup:
                times.append(float(dt.item()))
            d2h = labels.nbytes + x.nbytes
        t = sum(times) / len(times)
        return {"value": self.F / t, "unit": "faces/s", "h2d_bytes_per_step": int(h2d) * self.world,
                "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * t,
                "path": "resident API, scene re-uploaded from pinned host memory every step on every rank"}


def e2e_host_path(b2, torch, s, adj, rings, steps=1, warmup=1):
    """The reference-shaped call sequence with pinned HOST buffers (what a texrecon drop-in does):
    b2tex_calculate_data_costs -> b2tex_view_selection -> b2tex_global_seam_leveling."""
    import copy

    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        try:
            t = t.pin_memory()
        except Exception:
            pass
        return t.numpy(), t

    keep = []
    sp = copy.copy(s)
    for name in ("verts", "faces", "face_normals", "images"):
        arr, t = pin(getattr(s, name))
        setattr(sp, name, arr)
        keep.append(t)
    ap, t = pin(adj[0]); keep.append(t)
    ai, t = pin(adj[1]); keep.append(t)
    pr = []
    for r in rings:
        a, t = pin(r); keep.append(t); pr.append(a)
    # result buffers are allocated (pinned) once, like an application that textures many scenes
    cap = int(s.num_faces) * 64
    out_t = [torch.empty(s.num_faces + 1, dtype=torch.int64).pin_memory(), torch.empty(cap, dtype=torch.int16).pin_memory(),
             torch.empty(cap, dtype=torch.float32).pin_memory()]
    out = (out_t[0].numpy().view(np.uint64), out_t[1].numpy().view(np.uint16), out_t[2].numpy())
    times, h2d, d2h, stages = [], 0, 0, None
    for i in range(warmup + steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dc = b2.calculate_data_costs(sp, out=out)
        t1 = time.perf_counter()
        labels, minfo = b2.view_selection(dc, ap, ai)
        t2 = time.perf_counter()
        g = b2.global_seam_leveling(sp, pr, labels)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            stages = (stages or []) + [(round(1e3 * (t1 - t0)), round(1e3 * (t2 - t1)), round(1e3 * (time.perf_counter() - t2)))]
        mesh_b = sp.verts.nbytes + sp.faces.nbytes + sp.face_normals.nbytes
        dc_b = dc.face_ptr.nbytes + dc.view.nbytes + dc.cost.nbytes
        h2d = (mesh_b + sp.images.nbytes) + (dc_b + ap.nbytes + ai.nbytes) + \
              (mesh_b + sp.images.nbytes + sum(r.nbytes for r in pr) + labels.nbytes)
        d2h = dc_b + 2 * labels.nbytes + g["row_ptr"].nbytes + g["row_label"].nbytes + g["x"].nbytes
    t = sum(times) / len(times)
    three = {"value": s.num_faces / t, "unit": "faces/s", "h2d_bytes_per_step": int(h2d),
             "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * t, "stage_ms_per_step": stages,
             "path": "b2tex_calculate_data_costs_into -> b2tex_view_selection -> b2tex_global_seam_leveling "
                     "(DataCosts cross PCIe twice, images are uploaded twice)"}
    # ---- the same three stages on one upload (b2tex_texture_hot_path): the headline e2e ----
    ftimes = []
    for i in range(warmup + steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = b2.texture_hot_path(sp, (ap, ai), pr)
        torch.cuda.synchronize()
        if i >= warmup:
            ftimes.append(time.perf_counter() - t0)
    tf = sum(ftimes) / len(ftimes)
    # the same call on PAGEABLE host buffers (what an unmodified texrecon holds: mve images live in ordinary heap memory)
    ptimes = []
    for i in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        b2.texture_hot_path(s, adj, rings)
        torch.cuda.synchronize()
        if i >= 1:
            ptimes.append(time.perf_counter() - t0)
    tp = sum(ptimes) / len(ptimes)
    mesh_b = sp.verts.nbytes + sp.faces.nbytes + sp.face_normals.nbytes
    return {"value": s.num_faces / tf, "unit": "faces/s",
            "h2d_bytes_per_step": int(mesh_b + sp.images.nbytes + ap.nbytes + ai.nbytes + sum(a.nbytes for a in pr)),
            "d2h_bytes_per_step": int(r["labels"].nbytes + r["row_ptr"].nbytes + r["row_label"].nbytes + r["x"].nbytes),
            "ms_per_step": 1e3 * tf,
            "pageable_host": {"value": s.num_faces / tp, "unit": "faces/s", "ms_per_step": 1e3 * tp,
                              "note": "same call, host buffers in ordinary (pageable) memory"},
            "path": "b2tex_texture_hot_path: pinned host mesh/images/graph in, labels + adjust values out "
                    "(what texrecon does with --no_intermediate_results); the image upload runs on a copy stream under "
                    "the BVH build, culling and visibility rays",
            "three_call_path": three}
    return {"value": s.num_faces / t, "unit": "faces/s", "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * t, "stage_ms_per_step": stages,
            "path": "b2tex_calculate_data_costs_into -> b2tex_view_selection -> b2tex_global_seam_leveling, pinned host buffers"}
