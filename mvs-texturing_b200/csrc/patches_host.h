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
#include <thread>
#include <vector>

namespace b2 {

constexpr int PATCH_BORDER = 1;  // texture_patch.h:21

// fn(chunk, begin, end) over [0, n) in `chunks` contiguous pieces on up to 16 host threads; results that depend on the
// order are collected per chunk and concatenated in chunk order by the callers, so they equal the sequential ones
template <typename Fn>
inline void host_parallel_chunks(size_t n, unsigned chunks, Fn fn)
{
    if (chunks <= 1 || n < 4096) { fn(0u, (size_t)0, n); return; }
    std::vector<std::thread> th;
    for (unsigned k = 0; k < chunks; ++k)
        th.emplace_back([=] { fn(k, n * k / chunks, n * (k + 1) / chunks); });
    for (std::thread &t : th) t.join();
}
inline unsigned host_threads()
{
    unsigned h = std::thread::hardware_concurrency();
    return h < 1 ? 1u : (h > 16 ? 16u : h);
}

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
    // faces grouped by label (counting sort, ascending face id inside a label); labels are independent of each other:
    // a breadth-first search never leaves its label, so the labels are searched in parallel and concatenated in order
    uint32_t maxl = 0;
    for (uint32_t i = 0; i < F; ++i) maxl = std::max(maxl, labels[i]);
    std::vector<uint32_t> lptr((size_t)maxl + 2, 0);
    for (uint32_t i = 0; i < F; ++i) lptr[labels[i] + 1]++;
    for (uint32_t l = 0; l <= maxl; ++l) lptr[l + 1] += lptr[l];
    std::vector<uint32_t> by_label(F), cur(lptr.begin(), lptr.end() - 1);
    for (uint32_t i = 0; i < F; ++i) by_label[cur[labels[i]]++] = i;
    std::vector<uint8_t> used(F, 0);
    std::vector<std::vector<uint32_t> > faces_of((size_t)maxl + 1);
    std::vector<std::vector<PatchComponent> > comps_of((size_t)maxl + 1);
    const unsigned nthreads = F > 50000 ? host_threads() : 1u;
    auto search = [&](unsigned k) {
        std::deque<uint32_t> queue;
        for (uint32_t label = 1 + k; label <= maxl; label += nthreads) {   // interleaved: balances the threads
            std::vector<uint32_t> &out = faces_of[label];
            out.reserve(lptr[label + 1] - lptr[label]);
            for (uint32_t p = lptr[label]; p < lptr[label + 1]; ++p) {
                const uint32_t i = by_label[p];
                if (used[i]) continue;
                PatchComponent c; c.label = label; c.begin = (uint32_t)out.size();
                queue.clear(); queue.push_back(i); used[i] = 1;
                while (!queue.empty()) {
                    const uint32_t node = queue.front(); queue.pop_front();
                    out.push_back(node);
                    for (uint32_t a = adj_ptr[node]; a < adj_ptr[node + 1]; ++a) {
                        const uint32_t w = adj_idx[a];
                        if (labels[w] == label && !used[w]) { queue.push_back(w); used[w] = 1; }
                    }
                }
                c.end = (uint32_t)out.size();
                comps_of[label].push_back(c);
            }
        }
    };
    if (nthreads == 1) search(0);
    else {
        std::vector<std::thread> th;
        for (unsigned k = 0; k < nthreads; ++k) th.emplace_back(search, k);
        for (std::thread &t : th) t.join();
    }
    comp_faces.reserve(F - (lptr[1] - lptr[0]));
    for (uint32_t label = 1; label <= maxl; ++label) {
        const uint32_t base = (uint32_t)comp_faces.size();
        comp_faces.insert(comp_faces.end(), faces_of[label].begin(), faces_of[label].end());
        for (PatchComponent c : comps_of[label]) { c.begin += base; c.end += base; comps.push_back(c); }
    }
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
    // faces of the patch around the vertex: a handful, kept inline (a heap vector per entry dominated the bookkeeping)
    uint32_t num_faces = 0;
    uint32_t inline_faces[7];
    std::vector<uint32_t> more_faces;   // beyond 7 (high-valence vertices)
    void add_face(uint32_t f) { if (num_faces < 7) inline_faces[num_faces] = f; else more_faces.push_back(f); ++num_faces; }
    uint32_t face(uint32_t i) const { return i < 7 ? inline_faces[i] : more_faces[i - 7]; }
};

// find_seam_edges (seam_leveling.cpp:16-59): one edge (v1 < v2) per pair of adjacent faces with different labels,
// in face-major order
inline void find_seam_edges(uint32_t F, const uint32_t *adj_ptr, const uint32_t *adj_idx, const uint32_t *labels,
                            const uint32_t *mesh_faces, std::vector<uint32_t> &edges /* pairs */)
{
    const unsigned chunks = F > 50000 ? host_threads() : 1u;
    std::vector<std::vector<uint32_t> > part(chunks);
    host_parallel_chunks(F, chunks, [&](unsigned k, size_t f0, size_t f1) {
        std::vector<uint32_t> &out = part[k];
        for (uint32_t node = (uint32_t)f0; node < (uint32_t)f1; ++node)
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
                out.push_back(v1); out.push_back(v2);
            }
    });
    edges.clear();
    for (unsigned k = 0; k < chunks; ++k) edges.insert(edges.end(), part[k].begin(), part[k].end());   // face-major order
}

// generate_texture_patches.cpp:520-535 + merge: per vertex one entry per patch (ascending patch id), first projection
// wins, faces appended.  slot order is patch major, so entries arrive in ascending patch order.  Only the vertices local
// seam leveling looks at get entries: those lying in more than one patch (:157) and the end points of seam edges (:117);
// for the two million faces of the C3 workload that is 3 % of the vertices.
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
    const unsigned chunks = T > 50000 ? host_threads() : 1u;
    // pass 1 (parallel): which vertices are touched by more than one patch.  first[v] is claimed by compare-and-swap, so
    // whichever patch gets there first, a second DIFFERENT patch always notices: need[] does not depend on the order.
    std::vector<uint32_t> first(num_verts, 0xFFFFFFFFu);
    std::vector<uint8_t> need(num_verts, 0);
    host_parallel_chunks(T, chunks, [&](unsigned, size_t t0, size_t t1) {
        for (size_t t = t0; t < t1; ++t) {
            const uint32_t f = slot_face[t], q = plan.slot_patch[t];
            for (int j = 0; j < 3; ++j) {
                const uint32_t v = mesh_faces[3 * (size_t)f + j];
                uint32_t seen = __atomic_load_n(&first[v], __ATOMIC_RELAXED);
                if (seen == 0xFFFFFFFFu) {
                    uint32_t expected = 0xFFFFFFFFu;
                    if (__atomic_compare_exchange_n(&first[v], &expected, q, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) continue;
                    seen = expected;
                }
                if (seen != q) __atomic_store_n(&need[v], (uint8_t)1, __ATOMIC_RELAXED);
            }
        }
    });
    for (uint32_t v : seam_edges) need[v] = 1;
    vpi.index.assign(num_verts, 0xFFFFFFFFu);
    vpi.vertex.clear();
    for (uint32_t v = 0; v < num_verts; ++v)
        if (need[v]) { vpi.index[v] = (uint32_t)vpi.vertex.size(); vpi.vertex.push_back(v); }
    vpi.entries.assign(vpi.vertex.size(), std::vector<VertexProj>());
    // pass 2 (parallel): the slots that touch such a vertex, per chunk, concatenated in slot order
    std::vector<std::vector<uint32_t> > hit(chunks);
    host_parallel_chunks(T, chunks, [&](unsigned k, size_t t0, size_t t1) {
        for (size_t t = t0; t < t1; ++t) {
            const uint32_t f = slot_face[t];
            if (need[mesh_faces[3 * (size_t)f]] | need[mesh_faces[3 * (size_t)f + 1]] | need[mesh_faces[3 * (size_t)f + 2]]) hit[k].push_back((uint32_t)t);
        }
    });
    // pass 3 (sequential, a few per cent of the slots): entries in slot order = ascending patch id, first projection wins
    for (unsigned k = 0; k < chunks; ++k)
        for (uint32_t t : hit[k]) {
            const uint32_t f = slot_face[t], q = plan.slot_patch[t];
            for (int j = 0; j < 3; ++j) {
                const uint32_t v = mesh_faces[3 * (size_t)f + j];
                if (!need[v]) continue;
                std::vector<VertexProj> &e = vpi.entries[vpi.index[v]];
                if (e.empty()) e.reserve(3);
                if (e.empty() || e.back().patch != q) {
                    VertexProj p; p.patch = q; p.x = tex[6 * (size_t)t + 2 * j]; p.y = tex[6 * (size_t)t + 2 * j + 1];
                    e.push_back(p);
                }
                e.back().add_face(f);
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
                    for (uint32_t i1 = 0; i1 < p1.num_faces && !common; ++i1)
                        for (uint32_t i2 = 0; i2 < p2.num_faces; ++i2) if (p1.face(i1) == p2.face(i2)) { common = true; break; }
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
