// patches_host.h -- host-side bookkeeping of the texture-patch stage (plain C++, no CUDA): which faces form a
// patch and in which order.  Follows UniGraph::get_subgraphs (libs/tex/uni_graph.cpp:21-55) and the candidate
// merge of tex::generate_texture_patches (libs/tex/generate_texture_patches.cpp:468-508).  The order matters:
// texcoords are stored in it and TexturePatch::adjust_colors lets later triangles overwrite earlier ones.
// Shared by patches.cu and by the host-emulation harness (tests/cpp/emul_patches.cpp).
#pragma once
#include <stdint.h>

#include <math.h>

#include <algorithm>
#include <cmath>
#include <deque>
#include <vector>

namespace b2 {

constexpr int PATCH_BORDER = 1;  // texture_patch.h:21

struct PatchComponent {
    uint32_t label;
    uint32_t begin, end;  // range in comp_faces (BFS order)
};

// Connected components of equal non-zero label, each in the reference's BFS order (FIFO queue, adjacency-list
// order), components ordered by (label, smallest face id) -- the order in which generate_texture_patches meets
// them when it calls get_subgraphs(label) for label = 1..K (:472-480).
inline void label_components(uint32_t F, const uint32_t *adj_ptr, const uint32_t *adj_idx, const uint32_t *labels,
                             std::vector<uint32_t> &comp_faces, std::vector<PatchComponent> &comps)
{
    comp_faces.clear(); comps.clear();
    comp_faces.reserve(F);
    std::vector<uint8_t> used(F, 0);
    std::deque<uint32_t> queue;
    for (uint32_t i = 0; i < F; ++i) {
        if (labels[i] == 0 || used[i]) continue;
        const uint32_t label = labels[i];
        PatchComponent c; c.label = label; c.begin = (uint32_t)comp_faces.size();
        queue.clear(); queue.push_back(i); used[i] = 1;
        while (!queue.empty()) {
            const uint32_t node = queue.front(); queue.pop_front();
            comp_faces.push_back(node);
            for (uint32_t a = adj_ptr[node]; a < adj_ptr[node + 1]; ++a) {
                const uint32_t w = adj_idx[a];
                if (labels[w] == label && !used[w]) { queue.push_back(w); used[w] = 1; }
            }
        }
        c.end = (uint32_t)comp_faces.size();
        comps.push_back(c);
    }
    // group by label, keeping the ascending-first-face order inside a label; rebuild the face list in that order
    std::vector<uint32_t> order(comps.size());
    for (size_t k = 0; k < order.size(); ++k) order[k] = (uint32_t)k;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return comps[a].label < comps[b].label; });
    std::vector<uint32_t> faces2; faces2.reserve(comp_faces.size());
    std::vector<PatchComponent> comps2; comps2.reserve(comps.size());
    for (uint32_t k : order) {
        PatchComponent c = comps[k];
        const uint32_t b = (uint32_t)faces2.size();
        faces2.insert(faces2.end(), comp_faces.begin() + c.begin, comp_faces.begin() + c.end);
        c.begin = b; c.end = (uint32_t)faces2.size();
        comps2.push_back(c);
    }
    comp_faces.swap(faces2); comps.swap(comps2);
}

struct PatchPlan {
    // per patch, 8 ints: label, min_x, min_y (both already minus the border), width, height, first slot, faces, 0
    std::vector<int32_t> desc;
    std::vector<uint64_t> pix_off;       // [patches + 1] pixel offset of every patch image
    // per final slot (patch major: the absorbing candidate's faces, then the absorbed ones in merge order)
    std::vector<uint32_t> slot_src;      // slot in component order (where k_project wrote the pixel coordinates)
    std::vector<uint32_t> slot_comp;     // component of the slot
    std::vector<uint32_t> slot_patch;
    // per component: its own candidate origin (min - border) and the chain of merge offsets applied afterwards
    std::vector<int32_t> comp_min;       // [comps][2]
    std::vector<uint32_t> comp_chain;    // [comps][2]: begin, count into chain
    std::vector<float> chain;            // [][2] offsets, in the order they are added to the texcoords
    uint32_t num_patches() const { return (uint32_t)(desc.size() / 8); }
    uint32_t num_slots() const { return (uint32_t)slot_src.size(); }
};

// bbox: per component min_x, min_y, max_x, max_y of floor/ceil of the projected corners (:96-99), not yet
// border adjusted.  Candidates = components; inside one label a candidate whose (border adjusted) box lies inside
// another one's is absorbed by it (:484-508), absorbed texcoords get the difference of the origins added.
inline void plan_patches(const std::vector<PatchComponent> &comps, const int32_t *bbox, PatchPlan &plan)
{
    struct Cand { int32_t min_x, min_y, max_x, max_y; std::vector<uint32_t> members; bool erased; };
    const size_t C = comps.size();
    plan = PatchPlan();
    plan.comp_min.assign(2 * C, 0);
    std::vector<std::vector<float> > chains(C);
    plan.pix_off.push_back(0);
    size_t g0 = 0;
    while (g0 < C) {
        size_t g1 = g0;
        while (g1 < C && comps[g1].label == comps[g0].label) ++g1;
        std::vector<Cand> cands;
        for (size_t c = g0; c < g1; ++c) {
            Cand k;
            k.min_x = bbox[4 * c + 0] - PATCH_BORDER; k.min_y = bbox[4 * c + 1] - PATCH_BORDER;   // :113-117
            k.max_x = bbox[4 * c + 2]; k.max_y = bbox[4 * c + 3];                                 // Rect(min, max) :135
            k.members.push_back((uint32_t)c); k.erased = false;
            plan.comp_min[2 * c] = k.min_x; plan.comp_min[2 * c + 1] = k.min_y;
            cands.push_back(k);
        }
        // std::list iteration with erase == index iteration that skips erased entries
        for (size_t it = 0; it < cands.size(); ++it) {
            if (cands[it].erased) continue;
            for (size_t sit = 0; sit < cands.size(); ++sit) {
                if (sit == it || cands[sit].erased) continue;
                const Cand &s = cands[sit];
                Cand &o = cands[it];
                // Rect::is_inside: within or on the border of the other rectangle
                if (s.min_x >= o.min_x && s.max_x <= o.max_x && s.min_y >= o.min_y && s.max_y <= o.max_y) {
                    const float dx = (float)(s.min_x - o.min_x), dy = (float)(s.min_y - o.min_y);   // :495-497
                    for (uint32_t m : s.members) { chains[m].push_back(dx); chains[m].push_back(dy); o.members.push_back(m); }
                    cands[sit].erased = true;
                }
            }
        }
        for (size_t k = 0; k < cands.size(); ++k) {
            if (cands[k].erased) continue;
            const Cand &o = cands[k];
            const uint32_t patch = plan.num_patches();
            const int32_t width = o.max_x - (o.min_x + PATCH_BORDER) + 1 + 2 * PATCH_BORDER;   // :109-114
            const int32_t height = o.max_y - (o.min_y + PATCH_BORDER) + 1 + 2 * PATCH_BORDER;
            const uint32_t first = plan.num_slots();
            for (uint32_t m : o.members)
                for (uint32_t s = comps[m].begin; s < comps[m].end; ++s) {
                    plan.slot_src.push_back(s); plan.slot_comp.push_back(m); plan.slot_patch.push_back(patch);
                }
            const int32_t d[8] = {(int32_t)comps[g0].label, o.min_x, o.min_y, width, height, (int32_t)first,
                                  (int32_t)(plan.num_slots() - first), 0};
            plan.desc.insert(plan.desc.end(), d, d + 8);
            plan.pix_off.push_back(plan.pix_off.back() + (uint64_t)(width > 0 ? width : 0) * (uint64_t)(height > 0 ? height : 0));
        }
        g0 = g1;
    }
    plan.comp_chain.assign(2 * C, 0);
    for (size_t c = 0; c < C; ++c) {
        plan.comp_chain[2 * c] = (uint32_t)(plan.chain.size() / 2);
        plan.comp_chain[2 * c + 1] = (uint32_t)(chains[c].size() / 2);
        plan.chain.insert(plan.chain.end(), chains[c].begin(), chains[c].end());
    }
}

// ---- local seam leveling bookkeeping (libs/tex/seam_leveling.cpp, local_seam_leveling.cpp:113-176) -------------
struct VertexProj {          // VertexProjectionInfo after merge_vertex_projection_infos (:40-65)
    uint32_t patch;
    float x, y;              // the FIRST projection of the vertex met in the patch
    std::vector<uint32_t> faces;
};

// find_seam_edges (seam_leveling.cpp:16-59): one edge (v1 < v2) per pair of adjacent faces with different labels,
// in face-major order
inline void find_seam_edges(uint32_t F, const uint32_t *adj_ptr, const uint32_t *adj_idx, const uint32_t *labels,
                            const uint32_t *mesh_faces, std::vector<uint32_t> &edges /* pairs */)
{
    edges.clear();
    for (uint32_t node = 0; node < F; ++node)
        for (uint32_t a = adj_ptr[node]; a < adj_ptr[node + 1]; ++a) {
            const uint32_t adj = adj_idx[a];
            if (node > adj || labels[node] == labels[adj]) continue;
            uint32_t shared[4]; int ns = 0;
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j)
                    if (mesh_faces[3 * (size_t)node + i] == mesh_faces[3 * (size_t)adj + j] && ns < 4) shared[ns++] = mesh_faces[3 * (size_t)node + i];
            if (ns != 2 || shared[0] == shared[1]) continue;   // the reference asserts this
            uint32_t v1 = shared[0], v2 = shared[1];
            if (v1 > v2) std::swap(v1, v2);
            edges.push_back(v1); edges.push_back(v2);
        }
}

// generate_texture_patches.cpp:520-535 + merge: per vertex one entry per patch (ascending patch id), first projection
// wins, faces appended.  slot order is patch major, so entries arrive in ascending patch order.  Only the vertices local
// seam leveling looks at get entries: those lying in more than one patch (:157) and the end points of seam edges (:117);
// for the two million faces of the C3 workload that is 3 % of the vertices (400 ms -> 85 ms on one host core, incl.
// find_seam_edges).
// Entries exist only for the vertices local seam leveling looks at; `index[v]` is the position of vertex v in `entries`
// or 0xFFFFFFFF.  (A vector per vertex of the mesh costs more than the rest of the bookkeeping at two million faces.)
struct VertexProjections {
    std::vector<uint32_t> index;                       // [num_verts]
    std::vector<uint32_t> vertex;                      // [entries] vertex id, ascending
    std::vector<std::vector<VertexProj> > entries;
    const std::vector<VertexProj> *find(uint32_t v) const { return index[v] == 0xFFFFFFFFu ? nullptr : &entries[index[v]]; }
};

inline void vertex_projections(uint32_t num_verts, const uint32_t *mesh_faces, const PatchPlan &plan, const uint32_t *slot_face,
                               const float *tex /* [slots][3][2] */, const std::vector<uint32_t> &seam_edges, VertexProjections &vpi)
{
    const uint32_t T = plan.num_slots();
    std::vector<uint32_t> first(num_verts, 0xFFFFFFFFu);
    std::vector<uint8_t> need(num_verts, 0);
    for (uint32_t t = 0; t < T; ++t) {
        const uint32_t f = slot_face[t], q = plan.slot_patch[t];
        for (int j = 0; j < 3; ++j) {
            const uint32_t v = mesh_faces[3 * (size_t)f + j];
            if (first[v] == 0xFFFFFFFFu) first[v] = q;
            else if (first[v] != q) need[v] = 1;
        }
    }
    for (uint32_t v : seam_edges) need[v] = 1;
    vpi.index.assign(num_verts, 0xFFFFFFFFu);
    vpi.vertex.clear();
    for (uint32_t v = 0; v < num_verts; ++v)
        if (need[v]) { vpi.index[v] = (uint32_t)vpi.vertex.size(); vpi.vertex.push_back(v); }
    vpi.entries.assign(vpi.vertex.size(), std::vector<VertexProj>());
    for (uint32_t t = 0; t < T; ++t) {
        const uint32_t f = slot_face[t], q = plan.slot_patch[t];
        for (int j = 0; j < 3; ++j) {
            const uint32_t v = mesh_faces[3 * (size_t)f + j];
            if (!need[v]) continue;
            std::vector<VertexProj> &e = vpi.entries[vpi.index[v]];
            if (e.empty() || e.back().patch != q) {
                VertexProj p; p.patch = q; p.x = tex[6 * (size_t)t + 2 * j]; p.y = tex[6 * (size_t)t + 2 * j + 1];
                e.push_back(p);
            }
            e.back().faces.push_back(f);
        }
    }
}

struct SeamLines {
    // per seam edge: [proj_begin, proj_count, sample_begin, sample_count]
    std::vector<uint32_t> edge_info;
    std::vector<float> edge_proj;        // per projection: p1.x, p1.y, p2.x, p2.y
    std::vector<uint32_t> proj_patch;    // per projection
    std::vector<uint32_t> sample_edge;   // per colour sample: its edge
    // per multi-patch vertex: [proj_begin, proj_count]; projections share the layout of edge projections (x, y only)
    std::vector<uint32_t> vert_info;
    std::vector<float> vert_proj;        // per vertex projection: x, y
    std::vector<uint32_t> vert_proj_patch;
    uint32_t num_edges() const { return (uint32_t)(edge_info.size() / 4); }
    uint32_t num_samples() const { return (uint32_t)sample_edge.size(); }
    uint32_t num_verts() const { return (uint32_t)(vert_info.size() / 2); }
};

// find_mesh_edge_projections (seam_leveling.cpp:61-91) per seam edge, the sampling density of
// local_seam_leveling.cpp:131-140 and the vertex list of :155-176
inline void plan_seam_lines(const std::vector<uint32_t> &seam_edges, const VertexProjections &vpi, SeamLines &out)
{
    out = SeamLines();
    for (size_t ei = 0; ei + 1 < seam_edges.size(); ei += 2) {
            const uint32_t v1 = seam_edges[ei], v2 = seam_edges[ei + 1];
            const uint32_t pb = (uint32_t)out.proj_patch.size();
            float max_length = 1.0f;
            const std::vector<VertexProj> *e1 = vpi.find(v1), *e2 = vpi.find(v2);
            if (e1 && e2) for (const VertexProj &p1 : *e1)
                for (const VertexProj &p2 : *e2) {
                    if (p1.patch != p2.patch) continue;
                    bool common = false;
                    for (uint32_t f1 : p1.faces) { for (uint32_t f2 : p2.faces) if (f1 == f2) { common = true; break; } if (common) break; }
                    if (!common) continue;
                    out.proj_patch.push_back(p1.patch);
                    out.edge_proj.push_back(p1.x); out.edge_proj.push_back(p1.y); out.edge_proj.push_back(p2.x); out.edge_proj.push_back(p2.y);
                    const float dx = p1.x - p2.x, dy = p1.y - p2.y;
                    const float length = sqrtf((0.0f + dx * dx) + dy * dy);
                    max_length = std::max(max_length, length);
                }
            const uint32_t n = (uint32_t)std::ceil(max_length * 2.0f);   // :139
            const uint32_t sb = (uint32_t)out.sample_edge.size();
            const uint32_t e = out.num_edges();
            for (uint32_t j = 0; j < n; ++j) out.sample_edge.push_back(e);
            const uint32_t info[4] = {pb, (uint32_t)out.proj_patch.size() - pb, sb, n};
            out.edge_info.insert(out.edge_info.end(), info, info + 4);
        }
    for (size_t k = 0; k < vpi.entries.size(); ++k) {                // ascending vertex id
        const std::vector<VertexProj> &e = vpi.entries[k];
        if (e.size() <= 1) continue;                                 // :157
        const uint32_t info[2] = {(uint32_t)out.vert_proj_patch.size(), (uint32_t)e.size()};
        out.vert_info.insert(out.vert_info.end(), info, info + 2);
        for (const VertexProj &p : e) { out.vert_proj_patch.push_back(p.patch); out.vert_proj.push_back(p.x); out.vert_proj.push_back(p.y); }
    }
}

}  // namespace b2
