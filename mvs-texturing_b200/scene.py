"""Deterministic synthetic scenes for the hot path.

The reference ships no sample data (no tests/, no fixtures), so every input the
oracle, the parity tests and bench.py consume is generated here:

  C1  geodesic icosphere nu=22  (9 680 faces),  6 axis cameras, 640x480
  C2  heightfield terrain 500x500 quads (500 000 faces), 50 cameras, 1920x1080
  C3  displaced icosphere nu=316 (1 997 120 faces), 200 cameras, 1920x1080
  C5  heightfield terrain 2236x2236 quads (9 999 392 faces), 1000 cameras, 3840x2160 (C5s: 1 % scale)

Conventions follow the reference's TextureView (texture_view.h:43-48, 161-166):
  pos, viewdir (world), proj (3x3 row major, pixels), world_to_cam (4x4 row major).
Images are RGB u8 in [1,255]: never 0, so the corner flood fill of
texture_view.cpp:42-94 marks nothing invalid unless a test paints black borders.

This module is harness code (numpy only); it is not part of the product path.
"""
from __future__ import annotations

import dataclasses
import numpy as np


@dataclasses.dataclass
class Scene:
    verts: np.ndarray         # (Vn,3) f32
    faces: np.ndarray         # (F,3)  u32
    face_normals: np.ndarray  # (F,3)  f32
    pos: np.ndarray           # (K,3)  f32
    viewdir: np.ndarray       # (K,3)  f32
    proj: np.ndarray          # (K,9)  f32 row major
    w2c: np.ndarray           # (K,16) f32 row major
    width: int
    height: int
    images: np.ndarray        # (K,H,W,3) u8
    name: str = ""

    @property
    def num_faces(self) -> int:
        return int(self.faces.shape[0])

    @property
    def num_views(self) -> int:
        return int(self.pos.shape[0])


# ----------------------------------------------------------------------------
# meshes
# ----------------------------------------------------------------------------
def _icosahedron():
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0],
                  [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
                  [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
                  [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
                  [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    return v, f


def icosphere(nu: int, displace: float = 0.0, seed: int = 42):
    """Geodesic icosphere of frequency nu: F = 20 nu^2 faces, Vn = 10 nu^2 + 2."""
    bv, bf = _icosahedron()
    # barycentric lattice of one base triangle
    ij = [(i, j) for i in range(nu + 1) for j in range(nu + 1 - i)]
    ij = np.array(ij, dtype=np.int64)
    idx = -np.ones((nu + 1, nu + 1), dtype=np.int64)
    idx[ij[:, 0], ij[:, 1]] = np.arange(len(ij))
    tris = []
    for i in range(nu):
        j = np.arange(nu - i)
        tris.append(np.stack([idx[i, j], idx[i + 1, j], idx[i, j + 1]], 1))
        j2 = np.arange(nu - i - 1)
        if len(j2):
            tris.append(np.stack([idx[i + 1, j2], idx[i + 1, j2 + 1], idx[i, j2 + 1]], 1))
    tris = np.concatenate(tris, 0)
    a = ij[:, 0:1] / nu
    b = ij[:, 1:2] / nu
    allv, allf = [], []
    for k in range(20):
        A, B, C = bv[bf[k, 0]], bv[bf[k, 1]], bv[bf[k, 2]]
        P = A[None] * (1 - a - b) + B[None] * a + C[None] * b
        P /= np.linalg.norm(P, axis=1, keepdims=True)
        allf.append(tris + k * len(ij))
        allv.append(P)
    V = np.concatenate(allv, 0)
    Fc = np.concatenate(allf, 0)
    # weld duplicate vertices along base edges via a quantised key
    q = np.round((V + 1.5) * (1 << 19)).astype(np.int64)
    key = (q[:, 0] << 42) | (q[:, 1] << 21) | q[:, 2]
    _, first, inv = np.unique(key, return_index=True, return_inverse=True)
    V = V[first]
    Fc = inv[Fc]
    if displace > 0.0:
        rs = np.random.RandomState(seed)
        r = np.zeros(len(V))
        for octave in range(3):
            freq = 3.0 * (2 ** octave)
            for _ in range(3):
                d = rs.normal(size=3)
                d /= np.linalg.norm(d)
                r += (0.5 ** octave) * np.sin(freq * (V @ d) + rs.uniform(0, 6.28))
        r /= np.abs(r).max()
        V = V * (1.0 + displace * r)[:, None]
    return V.astype(np.float32), Fc.astype(np.uint32)


def terrain(n: int, amplitude: float = 0.1, seed: int = 1234):
    """Heightfield over [-1,1]^2 with n x n quads -> 2 n^2 faces."""
    g = np.linspace(-1.0, 1.0, n + 1)
    X, Y = np.meshgrid(g, g, indexing="xy")
    rs = np.random.RandomState(seed)
    Z = np.zeros_like(X)
    for octave in range(4):
        freq = 2.0 * (2 ** octave)
        for _ in range(2):
            th = rs.uniform(0, 6.28)
            Z += (0.5 ** octave) * np.sin(freq * (X * np.cos(th) + Y * np.sin(th)) + rs.uniform(0, 6.28))
    Z *= amplitude * 2.0 / np.abs(Z).max()
    V = np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1).astype(np.float32)
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    v00 = (j * (n + 1) + i).ravel()
    v10 = v00 + 1
    v01 = v00 + (n + 1)
    v11 = v01 + 1
    # two CCW (seen from +z) triangles per quad, interleaved so neighbours stay close
    F = np.stack([np.stack([v00, v10, v11], 1), np.stack([v00, v11, v01], 1)], 1).reshape(-1, 3)
    return V, F.astype(np.uint32)


def face_normals(verts, faces):
    """MVE ensure_normals [UPSTREAM-RECALL]: normalised cross(b-a, c-a), zero if degenerate."""
    a = verts[faces[:, 0]]
    b = verts[faces[:, 1]]
    c = verts[faces[:, 2]]
    n = np.cross((b - a).astype(np.float32), (c - a).astype(np.float32)).astype(np.float32)
    l = np.sqrt((n * n).sum(1, dtype=np.float32)).astype(np.float32)
    ok = l > 0
    n[ok] = n[ok] / l[ok, None]
    n[~ok] = 0
    return n.astype(np.float32)


# ----------------------------------------------------------------------------
# cameras
# ----------------------------------------------------------------------------
def look_at_camera(pos, target, flen_px, width, height, up=(0.0, 0.0, 1.0)):
    pos = np.asarray(pos, np.float64)
    z = np.asarray(target, np.float64) - pos
    z /= np.linalg.norm(z)
    upv = np.asarray(up, np.float64)
    if abs(np.dot(upv, z)) > 0.99:
        upv = np.array([0.0, 1.0, 0.0])
    x = np.cross(z, upv)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z], 0)           # rows = camera axes in world coordinates
    t = -R @ pos
    w2c = np.eye(4)
    w2c[:3, :3] = R
    w2c[:3, 3] = t
    proj = np.array([[flen_px, 0, width / 2.0], [0, flen_px, height / 2.0], [0, 0, 1.0]])
    return (pos.astype(np.float32), z.astype(np.float32),
            proj.astype(np.float32).ravel(), w2c.astype(np.float32).ravel())


def fibonacci_dirs(k: int, hemisphere: bool = False):
    i = np.arange(k) + 0.5
    phi = i * (np.pi * (3.0 - 5.0 ** 0.5))
    if hemisphere:
        z = 0.25 + 0.7 * i / k          # elevations between ~15 and ~72 degrees
    else:
        z = 1.0 - 2.0 * i / k
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    return np.stack([r * np.cos(phi), r * np.sin(phi), z], 1)


# ----------------------------------------------------------------------------
# images
# ----------------------------------------------------------------------------
def _hash_lattice(nx, ny, seed):
    xs = np.arange(nx, dtype=np.uint32)[None, :]
    ys = np.arange(ny, dtype=np.uint32)[:, None]
    h = (xs * np.uint32(0x9E3779B1)) ^ (ys * np.uint32(0x85EBCA77)) ^ np.uint32(seed)
    h ^= h >> np.uint32(16)
    h = h * np.uint32(0x7FEB352D)
    h ^= h >> np.uint32(15)
    h = h * np.uint32(0x846CA68B)
    h ^= h >> np.uint32(16)
    return (h >> np.uint32(8)).astype(np.float32) / np.float32(1 << 24)


def base_texture(width, height, seed=7):
    """Band-limited texture in [0,1]: 4 sinusoids + bilinear value noise (8 px cells)."""
    x = np.arange(width, dtype=np.float32)[None, :]
    y = np.arange(height, dtype=np.float32)[:, None]
    T = (np.sin(0.071 * x + 0.013 * y) + np.sin(0.023 * x - 0.067 * y + 1.3)
         + 0.7 * np.sin(0.19 * x + 0.11 * y + 0.4) + 0.7 * np.sin(0.13 * x - 0.23 * y + 2.1))
    cell = 8
    nx, ny = width // cell + 2, height // cell + 2
    L = _hash_lattice(nx, ny, seed)
    fx = (np.arange(width) % cell) / cell
    fy = (np.arange(height) % cell) / cell
    ix = np.arange(width) // cell
    iy = np.arange(height) // cell
    a = L[iy][:, ix]
    b = L[iy][:, ix + 1]
    c = L[iy + 1][:, ix]
    d = L[iy + 1][:, ix + 1]
    N = (a * (1 - fx)[None, :] + b * fx[None, :]) * (1 - fy)[:, None] + \
        (c * (1 - fx)[None, :] + d * fx[None, :]) * fy[:, None]
    T = (T - T.min()) / (T.max() - T.min())
    return (0.55 * T + 0.45 * N).astype(np.float32)


def make_images(k, width, height, out=None):
    T = base_texture(width, height)
    if out is None:
        out = np.empty((k, height, width, 3), dtype=np.uint8)
    tint = np.array([1.0, 0.9, 0.8], dtype=np.float32)

    def one(v):   # every view from its own seed: the result does not depend on how the views are spread over threads
        rs = np.random.RandomState(1000 + v)
        gain = rs.uniform(0.8, 1.2)
        bias = rs.uniform(-20.0, 20.0)
        dx, dy = int(rs.randint(0, width)), int(rs.randint(0, height))
        Tv = np.roll(T, (dy, dx), axis=(0, 1))
        base = (np.float32(gain * 200.0) * Tv + np.float32(bias + 25.0))
        for c in range(3):
            out[v, :, :, c] = np.clip(np.rint(base * tint[c]), 1, 255).astype(np.uint8)

    if k * width * height >= (1 << 26):   # large view sets (C3, C5): numpy releases the GIL in these kernels
        import os
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:
            list(ex.map(one, range(k)))
    else:
        for v in range(k):
            one(v)
    return out


# ----------------------------------------------------------------------------
# configs
# ----------------------------------------------------------------------------
def _assemble(name, V, F, cams, width, height, with_images=True):
    N = face_normals(V, F)
    pos = np.stack([c[0] for c in cams]).astype(np.float32)
    vd = np.stack([c[1] for c in cams]).astype(np.float32)
    proj = np.stack([c[2] for c in cams]).astype(np.float32)
    w2c = np.stack([c[3] for c in cams]).astype(np.float32)
    imgs = make_images(len(cams), width, height) if with_images else \
        np.zeros((len(cams), 0, 0, 3), np.uint8)
    return Scene(np.ascontiguousarray(V), np.ascontiguousarray(F), N, pos, vd, proj, w2c,
                 width, height, imgs, name)


def sphere_scene(nu, k, width, height, displace=0.0, dist=3.0, fill=0.8, axis_cams=False,
                 name="sphere", with_images=True):
    V, F = icosphere(nu, displace)
    if axis_cams:
        dirs = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], float)[:k]
    else:
        dirs = fibonacci_dirs(k)
    # sphere of radius ~1 seen from `dist` subtends asin(1/dist); fill ~80 % of the short side
    half = fill * min(width, height) / 2.0
    flen = half / np.tan(np.arcsin((1.0 + displace) / dist))
    cams = [look_at_camera(d * dist, (0, 0, 0), flen, width, height) for d in dirs]
    return _assemble(name, V, F, cams, width, height, with_images)


def occluder_scene(nu, k, width, height, plates=5, displace=0.05, dist=3.0, name="occ", with_images=True):
    """Displaced sphere plus `plates` floating square plates (2 large triangles each, facing outwards) at
    radius 1.7: every plate hides part of the sphere from the cameras behind it, so the geometric
    visibility test (calculate_data_costs.cpp:194-215) rejects faces that pass all other culls, and the
    BVH holds triangles of very different sizes."""
    V, F = icosphere(nu, displace)
    V = V.astype(np.float64)
    Vs, Fs = [V], [F.astype(np.int64)]
    base = len(V)
    for d in fibonacci_dirs(plates) * np.array([1.0, 1.0, 0.6]):
        d = d / np.linalg.norm(d)
        a = np.cross(d, [0.0, 0.0, 1.0] if abs(d[2]) < 0.9 else [1.0, 0.0, 0.0])
        a /= np.linalg.norm(a)
        b = np.cross(d, a)
        c = d * 1.7
        h = 0.4
        quad = np.stack([c - h * a - h * b, c + h * a - h * b, c + h * a + h * b, c - h * a + h * b])
        Vs.append(quad)
        Fs.append(np.array([[0, 1, 2], [0, 2, 3]], np.int64) + base)   # normal = a x b = +d (outwards)
        base += 4
    V = np.concatenate(Vs, 0).astype(np.float32)
    F = np.concatenate(Fs, 0).astype(np.uint32)
    half = 0.8 * min(width, height) / 2.0
    flen = half / np.tan(np.arcsin(min(0.99, 2.0 / dist)))
    cams = [look_at_camera(dd * dist, (0, 0, 0), flen, width, height) for dd in fibonacci_dirs(k)]
    return _assemble(name, V, F, cams, width, height, with_images)


def messy_scene(nu, k, width, height, name="messy", with_images=True):
    """Input hygiene cases real scans have and the clean generators lack: non-manifold edges (fins glued onto existing
    edges: three faces per edge, face-graph degree 4), zero-area faces (coincident vertices: zero normal, dropped because their
    quality is 0), an unreferenced vertex, a sliver, and a small detached component."""
    V, F = icosphere(nu, 0.04)
    V = V.astype(np.float64)
    F = F.astype(np.int64)
    extra_v, extra_f = [], []
    nv = len(V)
    for f in (7, 101, 350):                                   # fins
        a, b, c = F[f]
        tip = V[[a, b, c]].mean(0) * 1.25
        extra_v.append(tip); extra_f += [[a, b, nv], [b, c, nv]]; nv += 1
    a, b, c = F[33]                                           # zero-area faces: coincident POSITIONS, distinct indices
    extra_v += [V[a].copy(), V[c].copy(), V[c].copy()]
    extra_f += [[a, nv, b], [c, nv + 1, nv + 2]]; nv += 3
    extra_v.append(np.array([0.0, 0.0, 2.0])); nv += 1        # unreferenced vertex
    a, b, c = F[60]
    extra_v.append(V[a] * 0.999 + V[b] * 0.001 + 1e-5); extra_f.append([a, nv, b]); nv += 1   # sliver on an edge
    base = nv
    d = np.array([0.3, -0.8, 0.52]); d /= np.linalg.norm(d)
    e1 = np.cross(d, [0, 0, 1.0]); e1 /= np.linalg.norm(e1); e2 = np.cross(d, e1)
    extra_v += [d * 1.5 - 0.1 * e1 - 0.1 * e2, d * 1.5 + 0.1 * e1 - 0.1 * e2, d * 1.5 + 0.1 * e2]
    extra_f.append([base, base + 1, base + 2])                # detached triangle facing outwards
    V = np.concatenate([V, np.array(extra_v)], 0).astype(np.float32)
    F = np.concatenate([F, np.array(extra_f, np.int64)], 0).astype(np.uint32)
    half = 0.8 * min(width, height) / 2.0
    flen = half / np.tan(np.arcsin(min(0.99, 1.7 / 3.0)))
    cams = [look_at_camera(dd * 3.0, (0, 0, 0), flen, width, height) for dd in fibonacci_dirs(k)]
    return _assemble(name, V, F, cams, width, height, with_images)


def terrain_scene(n, k, width, height, dist=3.2, name="terrain", with_images=True):
    V, F = terrain(n)
    dirs = fibonacci_dirs(k, hemisphere=True)
    flen = 0.42 * min(width, height) / np.tan(np.arcsin(min(0.99, 1.5 / dist)))
    cams = [look_at_camera(d * dist, (0, 0, 0), flen, width, height) for d in dirs]
    return _assemble(name, V, F, cams, width, height, with_images)


def config(name: str, with_images=True) -> Scene:
    """Named workloads of BASELINE.json `configs` (plus tiny ones for unit tests)."""
    if name == "tiny":      # 320 faces, 6 views: brute-force sized
        return sphere_scene(4, 6, 160, 120, axis_cams=True, name=name, with_images=with_images)
    if name == "small":     # 2 000 faces, 12 views
        return sphere_scene(10, 12, 320, 240, displace=0.05, name=name, with_images=with_images)
    if name == "occ":       # 2 010 faces, 12 views, floating plates: real occlusion
        return occluder_scene(10, 12, 320, 240, plates=9, name=name, with_images=with_images)
    if name == "messy":     # 1 290 faces / 653 vertices: non-manifold fins, zero-area faces, sliver, detached triangle, unreferenced vertex
        return messy_scene(8, 10, 320, 240, name=name, with_images=with_images)
    if name == "occ2":      # 32 030 faces, 24 views
        return occluder_scene(40, 24, 640, 480, plates=15, name=name, with_images=with_images)
    if name == "C1":
        return sphere_scene(22, 6, 640, 480, axis_cams=True, name=name, with_images=with_images)
    if name == "C1d":       # C1 with displacement + more views: exercises occlusion
        return sphere_scene(22, 16, 640, 480, displace=0.08, name=name, with_images=with_images)
    if name == "C2":
        return terrain_scene(500, 50, 1920, 1080, name=name, with_images=with_images)
    if name == "C2s":       # scaled-down terrain for CI-sized parity
        return terrain_scene(60, 10, 480, 270, name=name, with_images=with_images)
    if name == "C3":
        return sphere_scene(316, 200, 1920, 1080, displace=0.05, name=name, with_images=with_images)
    if name == "C5":        # BASELINE.json configs[4]: 10M-face terrain, 1000 views at 4K (24.9 GB of images: 8-GPU runs only)
        return terrain_scene(2236, 1000, 3840, 2160, name=name, with_images=with_images)
    if name == "C5s":       # C5 scaled to 1 %: 100 352 faces, 100 views, same camera layout
        return terrain_scene(224, 100, 768, 432, name=name, with_images=with_images)
    if name == "C3s":       # 1/16-size C3
        return sphere_scene(79, 50, 960, 540, displace=0.05, name=name, with_images=with_images)
    raise KeyError(name)


# ----------------------------------------------------------------------------
# mesh connectivity (MeshInfo / build_adjacency_graph restated; harness side)
# ----------------------------------------------------------------------------
def face_adjacency(faces: np.ndarray):
    """CSR of faces sharing an edge (build_adjacency_graph.cpp:16-53).

    Order inside a row follows the reference: edges (v1,v2),(v2,v3),(v3,v1), duplicates dropped.
    """
    F = faces.shape[0]
    f = faces.astype(np.int64)
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0)
    lo = np.minimum(e[:, 0], e[:, 1])
    hi = np.maximum(e[:, 0], e[:, 1])
    key = lo * (int(f.max()) + 1) + hi
    fid = np.tile(np.arange(F, dtype=np.int64), 3)
    slot = np.repeat(np.arange(3, dtype=np.int64), F)
    order = np.argsort(key, kind="stable")
    ks, fs, ss = key[order], fid[order], slot[order]
    # group boundaries
    start = np.flatnonzero(np.r_[True, ks[1:] != ks[:-1]])
    cnt = np.diff(np.r_[start, len(ks)])
    grp = np.repeat(np.arange(len(start)), cnt)
    # pairs inside each group (manifold: groups of 2)
    pairs_a, pairs_b, pairs_slot = [], [], []
    maxc = int(cnt.max())
    pos_in = np.arange(len(ks)) - start[grp]
    for d in range(1, maxc):
        ok = pos_in + d < cnt[grp]
        i = np.flatnonzero(ok)
        a, b = fs[i], fs[i + d]
        pairs_a += [a, b]
        pairs_b += [b, a]
        pairs_slot += [ss[i], ss[i + d]]
    if not pairs_a:
        return np.zeros(F + 1, np.uint32), np.zeros(0, np.uint32)
    a = np.concatenate(pairs_a)
    b = np.concatenate(pairs_b)
    s = np.concatenate(pairs_slot)
    keep = a != b
    a, b, s = a[keep], b[keep], s[keep]
    # UniGraph::add_edge (uni_graph.h:80-88) appends to BOTH lists when the lower face is
    # visited, so a row holds lower-id neighbours first (ascending), then higher ones by slot.
    hi_side = (b > a).astype(np.int64)
    # several higher faces on one (non-manifold) edge: MeshInfo::get_faces_for_edge order = ascending face id here
    o = np.lexsort((b, np.where(b < a, b, s), hi_side, a))
    a, b, s = a[o], b[o], s[o]
    # drop duplicate (a,b) keeping first occurrence in (slot) order
    ab = a * F + b
    _, first = np.unique(ab, return_index=True)
    first.sort()
    a, b = a[first], b[first]
    ptr = np.zeros(F + 1, np.int64)
    np.add.at(ptr, a + 1, 1)
    ptr = np.cumsum(ptr)
    return ptr.astype(np.uint32), b.astype(np.uint32)


def vertex_rings(faces: np.ndarray, num_verts: int):
    """Per-vertex incident faces and 1-ring vertices as CSR (MeshInfo restated, ascending ids)."""
    f = faces.astype(np.int64)
    F = f.shape[0]
    v = f.ravel()
    fid = np.repeat(np.arange(F, dtype=np.int64), 3)
    o = np.lexsort((fid, v))
    vf_idx = fid[o]
    vf_ptr = np.zeros(num_verts + 1, np.int64)
    np.add.at(vf_ptr, v + 1, 1)
    vf_ptr = np.cumsum(vf_ptr)
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]], f[:, [1, 0]], f[:, [2, 1]], f[:, [0, 2]]], 0)
    key = np.unique(e[:, 0] * num_verts + e[:, 1])
    a = key // num_verts
    b = key % num_verts
    vv_ptr = np.zeros(num_verts + 1, np.int64)
    np.add.at(vv_ptr, a + 1, 1)
    vv_ptr = np.cumsum(vv_ptr)
    return (vf_ptr.astype(np.uint32), vf_idx.astype(np.uint32),
            vv_ptr.astype(np.uint32), b.astype(np.uint32))
