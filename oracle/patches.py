"""oracle/patches.py -- TEST INFRASTRUCTURE (see oracle/oracle.h).

Pure-Python restatement (small scenes only) of the step that sits between view selection and seam
leveling in the reference, and of the patch-based colour sampling that global_seam_leveling really does:

  generate_texture_patches   libs/tex/generate_texture_patches.cpp:453-538 (no hole filling: every face
                             of the test scenes is seen), generate_candidate :78-138,
                             UniGraph::get_subgraphs uni_graph.cpp:21-55,
                             merge_vertex_projection_infos :40-65
  find_mesh_edge_projections libs/tex/seam_leveling.cpp:61-91
  sample_edge / calculate_difference   libs/tex/global_seam_leveling.cpp:26-43, 86-138
  TexturePatch::get_pixel_value        libs/tex/texture_patch.cpp:162-169 (FloatImage::linear_at)

  TexturePatch::adjust_colors          libs/tex/texture_patch.cpp:41-116, applied per patch as global_seam_leveling.cpp:293-323
  tex::local_seam_leveling             libs/tex/local_seam_leveling.cpp:105-204 with draw_line :39-92, find_seam_edges
                                       seam_leveling.cpp:16-59, prepare_blending_mask texture_patch.cpp:197-297,
                                       poisson_blend poisson_blending.cpp:49-138 (scipy splu for Eigen::SparseLU)
All of it is held to the reference's own translation units by tests/test_ref_pinning.py (patches, masks, zero-adjust images
bit for bit; leveled images to 2e-5) and is the checker of csrc/patches.cu / csrc/localseam.cu.

Purpose of the first part: pin the "stage-isolated" shortcut used by oracle/seam.c and by the CUDA path (colours sampled
from the whole view of a label at get_pixel_coords(vertex)) against the patch-relative sampling of the
reference.  tests/test_oracle_cpu.py::test_patch_sampling_equals_view_sampling compares Rhs = A^T b.
Patch ids are assigned in ascending label order (the reference's order depends on OpenMP scheduling,
generate_texture_patches.cpp:469,514-518; ids only name patches, they do not enter the arithmetic).
"""
from __future__ import annotations

import collections
import ctypes as C

import numpy as np

f32 = np.float32
BORDER = 1  # texture_patch.h:21


def _pixel_coords(O, view, x):
    out = (C.c_float * 2)()
    O.lib().orc_pixel_coords(C.byref(view), (C.c_float * 3)(*[float(v) for v in x]), out)
    return f32(out[0]), f32(out[1])


def get_subgraphs(adj_ptr, adj_idx, labels, label):
    """uni_graph.cpp:21-55: BFS components of one label, in the reference's visiting order"""
    used = np.zeros(len(labels), bool)
    comps = []
    for i in np.flatnonzero(labels == label):
        if used[i]:
            continue
        comp, queue = [], collections.deque([int(i)])
        used[i] = True
        while queue:
            node = queue.popleft()
            comp.append(node)
            for a in adj_idx[adj_ptr[node]:adj_ptr[node + 1]]:
                if labels[a] == label and not used[a]:
                    queue.append(int(a))
                    used[a] = True
        comps.append(comp)
    return comps


class Patch:
    def __init__(self, label, faces, texcoords, image, bbox):
        self.label, self.faces, self.texcoords, self.image, self.bbox = label, faces, texcoords, image, bbox


def generate_candidate(O, scene, views, label, faces_list):
    """generate_texture_patches.cpp:78-138"""
    view = views[label - 1]
    W, H = scene.width, scene.height
    min_x, min_y, max_x, max_y = W, H, 0, 0
    tex = []
    for f in faces_list:
        for j in range(3):
            px, py = _pixel_coords(O, view, scene.verts[scene.faces[f, j]])
            tex.append([px, py])
            min_x = min(int(np.floor(px)), min_x); min_y = min(int(np.floor(py)), min_y)
            max_x = max(int(np.ceil(px)), max_x); max_y = max(int(np.ceil(py)), max_y)
    width, height = max_x - min_x + 1 + 2 * BORDER, max_y - min_y + 1 + 2 * BORDER
    min_x -= BORDER; min_y -= BORDER
    tex = np.array(tex, f32) - np.array([min_x, min_y], f32)          # relative texcoords (:121-124)
    # mve::image::crop with (255,0,255) outside the view, then byte_to_float_image (:126-128)
    img = np.empty((height, width, 3), np.uint8)
    img[:] = np.array([255, 0, 255], np.uint8)
    x0, y0 = max(min_x, 0), max(min_y, 0)
    x1, y1 = min(min_x + width, W), min(min_y + height, H)
    if x1 > x0 and y1 > y0:
        img[y0 - min_y:y1 - min_y, x0 - min_x:x1 - min_x] = scene.images[label - 1][y0:y1, x0:x1]
    fimg = (img.astype(f32) / f32(255.0)).astype(f32)
    return Patch(label, list(faces_list), tex, fimg, [min_x, min_y, max_x, max_y])


def generate_texture_patches(O, scene, adj, labels):
    """returns (patches, vertex_projection_infos) -- generate_texture_patches.cpp:453-538"""
    adj_ptr, adj_idx = adj
    views, _keep = O.make_views(scene)
    patches = []
    vpi = [dict() for _ in range(scene.verts.shape[0])]   # vertex -> {patch_id: (projection, [faces])}
    for label in range(1, scene.num_views + 1):
        cands = [generate_candidate(O, scene, views, label, comp)
                 for comp in get_subgraphs(adj_ptr, adj_idx, labels, label)]
        # merge candidates whose bounding box lies inside another one (:484-508)
        i = 0
        while i < len(cands):
            it = cands[i]
            j = 0
            while j < len(cands):
                sit = cands[j]
                b, ob = sit.bbox, it.bbox
                if sit is not it and b[0] >= ob[0] and b[2] <= ob[2] and b[1] >= ob[1] and b[3] <= ob[3]:
                    it.faces += sit.faces
                    off = np.array([b[0] - ob[0], b[1] - ob[1]], f32)
                    it.texcoords = np.concatenate([it.texcoords, (sit.texcoords + off).astype(f32)], 0)
                    del cands[j]
                    if j < i:
                        i -= 1
                else:
                    j += 1
            i += 1
        for cand in cands:
            pid = len(patches)
            patches.append(cand)
            for k, f in enumerate(cand.faces):
                for j in range(3):
                    v = int(scene.faces[f, j])
                    proj = cand.texcoords[3 * k + j]
                    if pid not in vpi[v]:                     # merge_vertex_projection_infos (:40-65):
                        vpi[v][pid] = (proj, [f])             # first projection wins, faces are appended
                    else:
                        vpi[v][pid][1].append(f)
    return patches, vpi


def _linear_at(img, x, y):
    """mve::FloatImage::linear_at [UPSTREAM-RECALL], fp32"""
    h, w, _ = img.shape
    x = max(f32(0.0), min(f32(w - 1), x)); y = max(f32(0.0), min(f32(h - 1), y))
    fx, fy = int(x), int(y)
    fx1, fy1 = min(fx + 1, w - 1), min(fy + 1, h - 1)
    w1 = f32(x - f32(fx)); w0 = f32(f32(1.0) - w1)
    w3 = f32(y - f32(fy)); w2 = f32(f32(1.0) - w3)
    return ((img[fy, fx] * f32(w0 * w2) + img[fy, fx1] * f32(w1 * w2)).astype(f32)
            + img[fy1, fx] * f32(w0 * w3) + img[fy1, fx1] * f32(w1 * w3)).astype(f32)


def sample_edge(patch, p1, p2):
    """global_seam_leveling.cpp:26-43"""
    p12 = (p2 - p1).astype(f32)
    nrm = f32(np.sqrt(f32(f32(p12[0] * p12[0]) + f32(p12[1] * p12[1]))))
    n = int(f32(max(nrm, f32(1.0)) * f32(2.0)))
    acc, wsum = np.zeros(3, f32), f32(0.0)
    for s in range(n):
        fraction = f32(f32(s) / f32(n - 1))
        sp = (p1 + (p12 * fraction).astype(f32)).astype(f32)
        col = _linear_at(patch.image, sp[0], sp[1])
        wgt = f32(f32(1.0) - fraction)
        acc = (acc + (col * wgt).astype(f32)).astype(f32)
        wsum = f32(wsum + wgt)
    return (acc / wsum).astype(f32)


def find_mesh_edge_projections(vpi, v1, v2):
    """seam_leveling.cpp:61-91: one entry per patch that holds both vertices through a common face"""
    out = {}
    for pid, (p1, faces1) in vpi[v1].items():
        if pid in vpi[v2]:
            p2, faces2 = vpi[v2][pid]
            if set(faces1) & set(faces2):
                out[pid] = (p1, p2)
    return [(pid, *out[pid]) for pid in sorted(out)]


def seam_rhs_from_patches(O, scene, adj, rings, labels, row_ptr, row_label):
    """Rhs = A^T b (global_seam_leveling.cpp:211-237,266-270) with colours sampled from texture patches."""
    patches, vpi = generate_texture_patches(O, scene, adj, labels)
    vf_ptr, vf_idx, vv_ptr, vv_idx = rings
    R = len(row_label)
    rhs = np.zeros((R, 3), f32)
    for i in range(scene.verts.shape[0]):
        rows = range(int(row_ptr[i]), int(row_ptr[i + 1]))
        for j in rows:
            for k in rows:
                l1, l2 = int(row_label[j]), int(row_label[k])
                if not l1 < l2:
                    continue
                c1, c2, w1, w2, any_edge = np.zeros(3, f32), np.zeros(3, f32), f32(0), f32(0), False
                for adjv in vv_idx[vv_ptr[i]:vv_ptr[i + 1]]:
                    adjv = int(adjv)
                    if adjv == i:
                        continue
                    ef = [int(f) for f in vf_idx[vf_ptr[i]:vf_ptr[i + 1]] if adjv in scene.faces[f]]
                    for x in range(len(ef)):
                        for y in range(x + 1, len(ef)):
                            fl = sorted((int(labels[ef[x]]), int(labels[ef[y]])))
                            if fl[0] != l1 or fl[1] != l2:
                                continue
                            d = (scene.verts[adjv] - scene.verts[i]).astype(f32)
                            length = f32(np.sqrt(f32(f32(f32(d[0] * d[0]) + f32(d[1] * d[1])) + f32(d[2] * d[2]))))
                            if length == 0:
                                continue
                            any_edge = True
                            n_used = 0
                            for pid, p1, p2 in find_mesh_edge_projections(vpi, i, adjv):
                                pl = patches[pid].label
                                if pl == l1:
                                    c1 = (c1 + (sample_edge(patches[pid], p1, p2) * length).astype(f32)).astype(f32); w1 = f32(w1 + length); n_used += 1
                                if pl == l2:
                                    c2 = (c2 + (sample_edge(patches[pid], p1, p2) * length).astype(f32)).astype(f32); w2 = f32(w2 + length); n_used += 1
                            assert n_used == 2                       # global_seam_leveling.cpp:124
                if not any_edge:
                    continue
                b = ((c2 / w2).astype(f32) - (c1 / w1).astype(f32)).astype(f32)
                rhs[j] = (rhs[j] + b).astype(f32)
                rhs[k] = (rhs[k] - b).astype(f32)
    return rhs, patches, vpi


# ---- TexturePatch::adjust_colors (libs/tex/texture_patch.cpp:41-116) --------------------------------
SQRT2 = f32(np.sqrt(2))


def _bary(v1, v2, v3, detT, x, y):
    """Tri::get_barycentric_coords (tri.h:50-56), fp32"""
    x, y = f32(x), f32(y)
    alpha = f32(f32(f32(f32(v2[1] - v3[1]) * f32(x - v3[0])) + f32(f32(v3[0] - v2[0]) * f32(y - v3[1]))) / detT)
    beta = f32(f32(f32(f32(v3[1] - v1[1]) * f32(x - v3[0])) + f32(f32(v1[0] - v3[0]) * f32(y - v3[1]))) / detT)
    gamma = f32(f32(f32(1.0) - alpha) - beta)
    return alpha, beta, gamma


def adjust_colors(patch, adjust_values):
    """Rasterises barycentric-interpolated per-vertex adjust values into the patch (pixels within
    sqrt(2) of a triangle are extrapolated and marked 64 in the blending mask), adds them to the image
    and zeroes pixels no triangle reaches.  adjust_values: (3 * num_faces, 3) fp32, one row per
    face corner in texcoord order.  Returns (image, validity_mask, blending_mask)."""
    h, w, _ = patch.image.shape
    validity = np.zeros((h, w), np.uint8)
    blending = np.zeros((h, w), np.uint8)
    iadj = np.zeros((h, w, 3), f32)
    tc = patch.texcoords
    for i in range(0, len(tc), 3):
        v1, v2, v3 = tc[i], tc[i + 1], tc[i + 2]
        detT = f32(f32(f32(v1[0] - v3[0]) * f32(v2[1] - v3[1])) - f32(f32(v1[1] - v3[1]) * f32(v2[0] - v3[0])))
        u, v = (v2 - v1).astype(f32), (v3 - v1).astype(f32)
        area = f32(f32(0.5) * abs(f32(f32(u[0] * v[1]) - f32(u[1] * v[0]))))
        if area < np.finfo(np.float32).eps:
            continue
        min_x = int(np.floor(min(v1[0], v2[0], v3[0]))) - BORDER
        min_y = int(np.floor(min(v1[1], v2[1], v3[1]))) - BORDER
        max_x = int(np.ceil(max(v1[0], v2[0], v3[0]))) + BORDER
        max_y = int(np.ceil(max(v1[1], v2[1], v3[1]))) + BORDER
        assert 0 <= min_x and max_x <= w and 0 <= min_y and max_y <= h          # texture_patch.cpp:63-64
        n23 = f32(np.sqrt(f32(f32((v2 - v3)[0] ** 2) + f32((v2 - v3)[1] ** 2))))
        n13 = f32(np.sqrt(f32(f32((v1 - v3)[0] ** 2) + f32((v1 - v3)[1] ** 2))))
        n12 = f32(np.sqrt(f32(f32((v1 - v2)[0] ** 2) + f32((v1 - v2)[1] ** 2))))
        a0, a1, a2 = adjust_values[i], adjust_values[i + 1], adjust_values[i + 2]
        for y in range(min_y, max_y):
            for x in range(min_x, max_x):
                b0, b1, b2 = _bary(v1, v2, v3, detT, x, y)
                inside = min(b0, b1, b2) >= 0
                if not inside:
                    if validity[y, x] == 255:
                        continue
                    ha = f32(f32(f32(f32(2.0) * -b0) * area) / n23)
                    hb = f32(f32(f32(f32(2.0) * -b1) * area) / n13)
                    hc = f32(f32(f32(f32(2.0) * -b2) * area) / n12)
                    if ha > SQRT2 or hb > SQRT2 or hc > SQRT2:
                        continue
                iadj[y, x] = ((a0 * b0).astype(f32) + (a1 * b1).astype(f32) + (a2 * b2).astype(f32)).astype(f32)
                validity[y, x] = 255
                blending[y, x] = 255 if inside else 64
    img = patch.image.copy()
    valid = validity != 0
    img[valid] = (img[valid] + iadj[valid]).astype(f32)
    img[~valid] = 0
    return img, validity, blending


def apply_adjust_values(scene, patches, row_ptr, row_label, x):
    """global_seam_leveling.cpp:293-323: gather the per-(vertex,label) offsets of every face corner of a
    patch and call adjust_colors; returns new Patch objects with adjusted images."""
    out = []
    for p in patches:
        adj = np.zeros((3 * len(p.faces), 3), f32)
        for k, f in enumerate(p.faces):
            for j in range(3):
                v = int(scene.faces[f, j])
                rows = range(int(row_ptr[v]), int(row_ptr[v + 1]))
                r = next(r for r in rows if int(row_label[r]) == p.label)   # adjust_values[vertex].find(label)
                adj[3 * k + j] = x[r]
        img, validity, blending = adjust_colors(p, adj)
        q = Patch(p.label, p.faces, p.texcoords, img, p.bbox)
        q.validity, q.blending = validity, blending
        out.append(q)
    return out


# =====================================================================================================
# local seam leveling (libs/tex/local_seam_leveling.cpp:105-204, poisson_blending.cpp:49-138,
# texture_patch.cpp:171-178,180-192,197-297, seam_leveling.cpp:16-59)
# Eigen::SparseLU<.., COLAMDOrdering> is absent: scipy.sparse.linalg.splu stands in (a direct solve of the
# same fp32 system; results agree up to rounding).  Small scenes only (pure Python loops).
# =====================================================================================================
STRIP_SIZE = 20  # local_seam_leveling.cpp:18


def find_seam_edges(scene, adj, labels):
    """seam_leveling.cpp:16-59: one MeshEdge (v1 < v2) per pair of adjacent faces with different labels"""
    adj_ptr, adj_idx = adj
    out = []
    for node in range(scene.num_faces):
        for a in adj_idx[adj_ptr[node]:adj_ptr[node + 1]]:
            a = int(a)
            if node > a or labels[node] == labels[a]:
                continue
            shared = [int(v) for v in scene.faces[node] if v in scene.faces[a]]
            assert len(shared) == 2 and shared[0] != shared[1]
            out.append((min(shared), max(shared)))
    return out


def _get_pixel_value(patch, p):
    return _linear_at(patch.image, f32(p[0]), f32(p[1]))


def _set_pixel_value(patch, x, y, color):
    """TexturePatch::set_pixel_value (texture_patch.cpp:171-178)"""
    patch.image[y, x] = color
    patch.blending[y, x] = 128


def draw_line(patch, p1, p2, edge_color):
    """local_seam_leveling.cpp:39-92 (Bresenham with colours interpolated along the edge samples)"""
    x0, y0 = int(np.floor(f32(p1[0] + f32(0.5)))), int(np.floor(f32(p1[1] + f32(0.5))))
    x1, y1 = int(np.floor(f32(p2[0] + f32(0.5)))), int(np.floor(f32(p2[1] + f32(0.5))))
    tdx, tdy = f32(x1 - x0), f32(y1 - y0)
    length = f32(np.sqrt(f32(f32(tdx * tdx) + f32(tdy * tdy))))
    dx, dy = abs(x1 - x0), abs(y1 - y0)
    sx, sy = (1 if x0 < x1 else -1), (1 if y0 < y1 else -1)
    err = dx - dy
    x, y = x0, y0
    while True:
        tdx, tdy = f32(x1 - x), f32(y1 - y)
        t = f32(np.sqrt(f32(f32(tdx * tdx) + f32(tdy * tdy))) / length) if length != 0 else f32(0.5)
        if t < 1.0 and len(edge_color) > 1:
            idx = int(np.floor(f32(t * f32(len(edge_color) - 1))))
            color = (f32(f32(1.0) - t) * edge_color[idx] + t * edge_color[idx + 1]).astype(f32)
        else:
            color = edge_color[-1]
        _set_pixel_value(patch, x, y, color)
        if x == x1 and y == y1:
            break
        e2 = 2 * err
        if e2 > -dy:
            err -= dy; x += sx
        if e2 < dx:
            err += dx; y += sy


def prepare_blending_mask(patch, strip_width=STRIP_SIZE):
    """texture_patch.cpp:197-297: keep a strip of `strip_width` pixels along the valid border"""
    validity, blending = patch.validity, patch.blending
    h, w = validity.shape
    border = set()
    for y in range(h):
        for x in range(w):
            if validity[y, x] == 0:
                continue
            if x == 0 or x == w - 1 or y == 0 or y == h - 1:
                border.add((x, y)); continue
            if np.any(validity[y - 1:y + 2, x - 1:x + 2] == 0):
                border.add((x, y))
    inner = validity.copy()
    for _ in range(strip_width):
        new_invalid = sorted(border)
        border = set()
        for x, y in new_invalid:
            inner[y, x] = 0
        for x, y in new_invalid:
            for j in (-1, 0, 1):
                for i in (-1, 0, 1):
                    nx, ny = x + i, y + j
                    if 0 <= nx < w and 0 <= ny < h and inner[ny, nx] == 255:
                        border.add((nx, ny))
    for y in range(1, h - 1):          # sanitize: a 128 pixel surrounded by 255 becomes 255
        for x in range(1, w - 1):
            if blending[y, x] == 128 and all(v == 255 for v in (blending[y, x - 1], blending[y, x + 1],
                                                                 blending[y - 1, x], blending[y + 1, x])):
                blending[y, x] = 255
    blending[inner == 255] = 0
    for x, y in border:
        blending[y, x] = 128


def poisson_blend(src, mask, dest, alpha=1.0):
    """poisson_blending.cpp:49-138: unknown per mask != 0 pixel; identity rows for 128/64, 5-point
    Laplacian rows for 255; rhs = alpha * lap(src) + (1 - alpha) * lap(dest); solved per channel"""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    h, w, _ = dest.shape
    idx = -np.ones(h * w, np.int64)
    nzpix = np.flatnonzero(mask.ravel() != 0)
    idx[nzpix] = np.arange(len(nzpix))
    n = len(nzpix)
    if n == 0:
        return
    m = mask.ravel()
    d = dest.reshape(-1, 3)
    s_ = src.reshape(-1, 3)
    rows, cols, vals = [], [], []
    b = np.zeros((n, 3), f32)
    alpha = f32(alpha)
    for i in nzpix:
        r = idx[i]
        if m[i] in (128, 64):
            rows.append(r); cols.append(r); vals.append(1.0)
            b[r] = d[i]
        if m[i] == 255:
            nb = [i - w, i - 1, i, i + 1, i + w]
            assert all(idx[k] != -1 for k in nb)                     # poisson_blending.cpp:98
            for k, v in zip(nb, (1.0, 1.0, -4.0, 1.0, 1.0)):
                rows.append(r); cols.append(idx[k]); vals.append(v)
            lap = lambda im: (f32(-4.0) * im[i] + im[i - w] + im[i - 1] + im[i + 1] + im[i + w]).astype(f32)
            b[r] = (alpha * lap(s_) + f32(f32(1.0) - alpha) * lap(d)).astype(f32)
    A = sp.csc_matrix((np.array(vals, f32), (np.array(rows), np.array(cols))), shape=(n, n))
    lu = spl.splu(A.astype(np.float64))
    for ch in range(3):
        x = lu.solve(b[:, ch].astype(np.float64)).astype(f32)
        d[nzpix, ch] = x


def local_seam_leveling(scene, adj, labels, patches, vpi):
    """local_seam_leveling.cpp:105-204 on patches that went through adjust_colors (they carry .validity
    and .blending).  Mutates the patch images; returns the list of seam edges."""
    seam_edges = find_seam_edges(scene, adj, labels)
    lines = [[] for _ in patches]
    pixels = [[] for _ in patches]
    for (v1, v2) in seam_edges:
        infos = find_mesh_edge_projections(vpi, v1, v2)
        max_length = f32(1.0)
        for pid, p1, p2 in infos:
            d = (p1 - p2).astype(f32)
            max_length = max(max_length, f32(np.sqrt(f32(f32(d[0] * d[0]) + f32(d[1] * d[1])))))
        n = int(np.ceil(f32(max_length * f32(2.0))))
        edge_color = []
        for j in range(n):
            t = f32(f32(j) / f32(n - 1))
            acc, wsum = np.zeros(3, f32), f32(0)
            for pid, p1, p2 in infos:                                    # mean_color_of_edge_point :20-37
                if patches[pid].label == 0:
                    continue
                pix = ((p1 * t).astype(f32) + (f32(f32(1.0) - t) * p2).astype(f32)).astype(f32)
                acc = (acc + _get_pixel_value(patches[pid], pix)).astype(f32); wsum = f32(wsum + 1)
            edge_color.append((acc / wsum).astype(f32))
        for pid, p1, p2 in infos:
            lines[pid].append(((p1 + f32(0.5)).astype(f32), (p2 + f32(0.5)).astype(f32), edge_color))
    vertex_colors = {}
    for v in range(scene.verts.shape[0]):                                # :155-176
        if len(vpi[v]) <= 1:
            continue
        acc, wsum = np.zeros(3, f32), f32(0)
        for pid, (proj, _faces) in sorted(vpi[v].items()):
            if patches[pid].label == 0:
                continue
            acc = (acc + _get_pixel_value(patches[pid], proj)).astype(f32); wsum = f32(wsum + 1)
        if wsum == 0:
            continue
        vertex_colors[v] = (acc / wsum).astype(f32)
        for pid, (proj, _faces) in sorted(vpi[v].items()):
            q = (proj + f32(0.5)).astype(f32)
            pixels[pid].append((int(q[0]), int(q[1]), vertex_colors[v]))   # math::Vec2i(Vec2f): truncation
    for pid, patch in enumerate(patches):                                # :179-203
        orig = patch.image.copy()
        for x, y, col in pixels[pid]:
            _set_pixel_value(patch, x, y, col)
        for (a, b_, col) in lines[pid]:
            # Line.from/to are Vec2i: the +0.5 shifted projections are truncated, draw_line then takes
            # floor(p + 0.5) of those integers (identity)
            draw_line(patch, np.array([int(a[0]), int(a[1])], f32), np.array([int(b_[0]), int(b_[1])], f32), col)
        if patch.label != 0:
            prepare_blending_mask(patch, STRIP_SIZE)
        poisson_blend(orig, patch.blending, patch.image, 1.0)             # TexturePatch::blend :180-192
        patch.validity[patch.blending == 64] = 0
    return seam_edges
