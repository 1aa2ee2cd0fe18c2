/* oracle/ref_glue.cpp -- TEST INFRASTRUCTURE (see oracle.h, refshim/README.md).
 *
 * C entry points over the reference's own translation units (compiled unmodified from
 * /root/reference/libs/tex against the dependency shims in oracle/refshim/), so that tests can pin
 * the oracle's restatement against the code it restates.  Built by `make -C oracle ref` into
 * oracle/_ref/libtexref.so; never linked or loaded by the product path.
 */
#include <cstdio>
#include <omp.h>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>
#include <string>

#include "oracle.h"

#include "tex/texturing.h"
#include "tex/histogram.h"
#include "tex/tri.h"
#include "mapmap/full.h"

namespace mapmap { Capture& refshim_capture() { static Capture c; return c; } }

namespace mve { namespace image {
std::map<std::string, ByteImage::Ptr>& refshim_registry() {
    static std::map<std::string, ByteImage::Ptr> reg;
    return reg;
}
} }

namespace {

std::string view_name(uint32_t j) { std::ostringstream s; s << "refshim://view/" << j; return s.str(); }

/* tex::TextureView from the flat view record (texture_view.cpp:20-40 via the CameraInfo shim) */
void make_views(const orc_view* views, uint32_t num_views, tex::TextureViews* out) {
    out->clear();
    out->reserve(num_views);
    for (uint32_t j = 0; j < num_views; ++j) {
        const orc_view& v = views[j];
        mve::ByteImage::Ptr img = mve::ByteImage::create(v.width, v.height, 3);
        std::memcpy(img->get_data_pointer(), v.rgb, static_cast<std::size_t>(v.width) * v.height * 3);
        mve::image::refshim_registry()[view_name(j)] = img;
        mve::CameraInfo cam;
        std::memcpy(cam.calibration, v.proj, sizeof(cam.calibration));
        std::memcpy(cam.position, v.pos, sizeof(cam.position));
        std::memcpy(cam.viewdir, v.viewdir, sizeof(cam.viewdir));
        std::memcpy(cam.world_to_cam, v.w2c, sizeof(cam.world_to_cam));
        out->push_back(tex::TextureView(j, cam, view_name(j)));
    }
}

mve::TriangleMesh::Ptr make_mesh(const float* verts, uint32_t num_verts, const uint32_t* faces,
                                 const float* face_normals, uint32_t num_faces) {
    mve::TriangleMesh::Ptr mesh = mve::TriangleMesh::create();
    mesh->get_vertices().resize(num_verts);
    for (uint32_t i = 0; i < num_verts; ++i) mesh->get_vertices()[i] = math::Vec3f(verts + 3 * static_cast<std::size_t>(i));
    mesh->get_faces().assign(faces, faces + 3 * static_cast<std::size_t>(num_faces));
    if (face_normals) {
        mesh->get_face_normals().resize(num_faces);
        for (uint32_t i = 0; i < num_faces; ++i) mesh->get_face_normals()[i] = math::Vec3f(face_normals + 3 * static_cast<std::size_t>(i));
    }
    return mesh;
}

tex::Settings make_settings(const orc_settings* st) {
    tex::Settings s;
    s.data_term = st->data_term == 0 ? tex::DATA_TERM_AREA : tex::DATA_TERM_GMI;
    s.outlier_removal = st->outlier_removal == 0 ? tex::OUTLIER_REMOVAL_NONE
        : (st->outlier_removal == 1 ? tex::OUTLIER_REMOVAL_GAUSS_DAMPING : tex::OUTLIER_REMOVAL_GAUSS_CLAMPING);
    s.geometric_visibility_test = st->geometric_visibility_test != 0;
    return s;
}

/* stdout of the reference (progress counters, timings) is noise for the tests */
struct Quiet {
    std::streambuf* old;
    std::ostringstream sink;
    Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~Quiet() { std::cout.rdbuf(old); }
};

}  // namespace

extern "C" {

void ref_free(void* p) { std::free(p); }

/* tex::calculate_data_costs (calculate_data_costs.cpp:308-323), result flattened to the CSR the
 * oracle returns: per face, entries in DataCosts::col(face) order (ascending view after :273). */
int ref_data_costs(const float* verts, uint32_t num_verts, const uint32_t* faces, const float* face_normals,
                   uint32_t num_faces, const orc_view* views, uint32_t num_views, const orc_settings* st,
                   uint64_t* face_ptr, uint16_t** view_out, float** cost_out)
{
    try {
        Quiet q;
        mve::TriangleMesh::Ptr mesh = make_mesh(verts, num_verts, faces, face_normals, num_faces);
        tex::TextureViews tvs;
        make_views(views, num_views, &tvs);
        tex::Settings settings = make_settings(st);
        tex::DataCosts data_costs(num_faces, num_views);
        tex::calculate_data_costs(mesh, &tvs, settings, &data_costs);
        std::size_t nnz = data_costs.get_nnz();
        uint16_t* vw = static_cast<uint16_t*>(std::malloc(sizeof(uint16_t) * (nnz ? nnz : 1)));
        float* cs = static_cast<float*>(std::malloc(sizeof(float) * (nnz ? nnz : 1)));
        uint64_t o = 0;
        for (uint32_t f = 0; f < num_faces; ++f) {
            face_ptr[f] = o;
            tex::DataCosts::Column const& col = data_costs.col(f);
            for (std::size_t k = 0; k < col.size(); ++k) { vw[o] = col[k].first; cs[o] = col[k].second; ++o; }
        }
        face_ptr[num_faces] = o;
        *view_out = vw; *cost_out = cs;
        return 0;
    } catch (std::exception& e) {
        std::fprintf(stderr, "ref_data_costs: %s\n", e.what());
        return 1;
    }
}

/* TextureView::generate_validity_mask [+ erode_validity_mask] (texture_view.cpp:42-94,109-132), read
 * back through export_validity_mask (:305-313) */
int ref_validity_mask(const orc_view* view, int erode, uint8_t* mask_out)
{
    Quiet q;
    tex::TextureViews tvs;
    make_views(view, 1, &tvs);
    tvs[0].load_image();
    tvs[0].generate_validity_mask();
    if (erode) tvs[0].erode_validity_mask();
    tvs[0].export_validity_mask("refshim://mask");
    mve::ByteImage::Ptr m = mve::image::refshim_registry()["refshim://mask"];
    for (int i = 0; i < m->get_pixel_amount(); ++i) mask_out[i] = m->at(i) ? 1 : 0;
    return 0;
}

/* TextureView::get_pixel_coords (texture_view.h:161-166) */
void ref_pixel_coords(const orc_view* view, const float x[3], float out[2])
{
    tex::TextureViews tvs;
    make_views(view, 1, &tvs);
    math::Vec2f p = tvs[0].get_pixel_coords(math::Vec3f(x));
    out[0] = p[0]; out[1] = p[1];
}

/* TextureView::get_face_info (texture_view.cpp:134-251) for a batch of world-space triangles of one
 * view; with_mask = generate (and for GMI erode) the validity mask first, as calculate_data_costs does.
 * quality_out[i] = NaN when the triangle does not project inside the view (get_face_info asserts). */
int ref_face_infos(const orc_view* view, int data_term, int outlier_removal, const float* tris /* n x 9 */, uint32_t n,
                   float* quality_out, float* mean_color_out /* n x 3 or NULL */)
{
    Quiet q;
    tex::TextureViews tvs;
    make_views(view, 1, &tvs);
    tex::TextureView& tv = tvs[0];
    orc_settings os = { data_term, outlier_removal, 0, 0, 0 };
    tex::Settings settings = make_settings(&os);
    tv.load_image();
    tv.generate_validity_mask();
    if (settings.data_term == tex::DATA_TERM_GMI) { tv.generate_gradient_magnitude(); tv.erode_validity_mask(); }
    for (uint32_t i = 0; i < n; ++i) {
        math::Vec3f v1(tris + 9 * i), v2(tris + 9 * i + 3), v3(tris + 9 * i + 6);
        if (!tv.inside(v1, v2, v3)) { quality_out[i] = std::numeric_limits<float>::quiet_NaN(); continue; }
        tex::FaceProjectionInfo info = { 0, 0.0f, math::Vec3f(0.0f, 0.0f, 0.0f) };
        tv.get_face_info(v1, v2, v3, &info, settings);
        quality_out[i] = info.quality;
        if (mean_color_out) for (int k = 0; k < 3; ++k) mean_color_out[3 * i + k] = info.mean_color[k];
    }
    return 0;
}

/* Tri (tri.h:50-84, tri.cpp:12-24) */
float ref_tri_area(const float p1[2], const float p2[2], const float p3[2])
{
    return Tri(math::Vec2f(p1), math::Vec2f(p2), math::Vec2f(p3)).get_area();
}
int ref_tri_inside(const float p1[2], const float p2[2], const float p3[2], float x, float y)
{
    return Tri(math::Vec2f(p1), math::Vec2f(p2), math::Vec2f(p3)).inside(x, y) ? 1 : 0;
}

/* Histogram (histogram.cpp:22-63) used as calculate_data_costs.cpp:283-288 does */
float ref_histogram_percentile(const float* values, uint64_t n, float vmax, int bins, float p)
{
    Histogram h(0.0f, vmax, static_cast<std::size_t>(bins));
    for (uint64_t i = 0; i < n; ++i) h.add_value(values[i]);
    return h.get_approx_percentile(p);
}

/* mve::MeshInfo from the per-vertex rings (the arrays b2tex_set_vertex_rings takes) */
static void fill_mesh_info(uint32_t num_verts, const uint32_t* vf_ptr, const uint32_t* vf_idx, const uint32_t* vv_ptr,
                           const uint32_t* vv_idx, mve::MeshInfo* mi)
{
    mi->resize(num_verts);
    for (uint32_t v = 0; v < num_verts; ++v) {
        (*mi)[v].vclass = mve::MeshInfo::VERTEX_CLASS_SIMPLE;
        (*mi)[v].faces.assign(vf_idx + vf_ptr[v], vf_idx + vf_ptr[v + 1]);
        (*mi)[v].verts.assign(vv_idx + vv_ptr[v], vv_idx + vv_ptr[v + 1]);
    }
}

/* tex::build_adjacency_graph (build_adjacency_graph.cpp:16-53): adjacency lists flattened to CSR in
 * UniGraph order.  adj_idx_out is malloc'd. */
int ref_build_adjacency(const uint32_t* faces, uint32_t num_faces, uint32_t num_verts, const uint32_t* vf_ptr, const uint32_t* vf_idx,
                        const uint32_t* vv_ptr, const uint32_t* vv_idx, uint32_t* adj_ptr, uint32_t** adj_idx_out)
{
    Quiet q;
    mve::TriangleMesh::Ptr mesh = mve::TriangleMesh::create();
    mesh->get_faces().assign(faces, faces + 3 * static_cast<std::size_t>(num_faces));
    mve::MeshInfo mi;
    fill_mesh_info(num_verts, vf_ptr, vf_idx, vv_ptr, vv_idx, &mi);
    tex::Graph graph(num_faces);
    tex::build_adjacency_graph(mesh, mi, &graph);
    adj_ptr[0] = 0;
    for (uint32_t f = 0; f < num_faces; ++f) adj_ptr[f + 1] = adj_ptr[f] + static_cast<uint32_t>(graph.get_adj_nodes(f).size());
    uint32_t* idx = static_cast<uint32_t*>(std::malloc(sizeof(uint32_t) * (adj_ptr[num_faces] ? adj_ptr[num_faces] : 1)));
    for (uint32_t f = 0; f < num_faces; ++f) {
        std::vector<std::size_t> const& a = graph.get_adj_nodes(f);
        for (std::size_t k = 0; k < a.size(); ++k) idx[adj_ptr[f] + k] = static_cast<uint32_t>(a[k]);
    }
    *adj_idx_out = idx;
    return 0;
}

/* tex::view_selection (view_selection.cpp:18-133) with the recording mapMAP shim: returns the model it
 * built and the labels it decoded from the shim's trivial solution.  All out arrays are malloc'd.
 * params_out: [potts, window, ratio, seed, deterministic, tree_algorithm, use_multilevel, use_spanning_tree,
 *              use_acyclic, multilevel_after, force_acyclic, min_acyclic_iterations, relax_acyclic_maximal,
 *              components_updated, compress, all unaries set exactly once] */
int ref_view_selection_model(uint32_t num_faces, uint32_t num_views, const uint32_t* adj_ptr, const uint32_t* adj_idx,
                             const uint64_t* face_ptr, const uint16_t* view, const float* cost,
                             uint64_t* num_edges_out, uint32_t** edges_out, uint64_t* ls_ptr /* F+1 */, int32_t** ls_label_out,
                             float** ls_cost_out, uint32_t* labels_out /* F */, double* params_out /* 16 */)
{
    try {
        Quiet q;
        tex::Graph graph(num_faces);
        for (uint32_t f = 0; f < num_faces; ++f)        // UniGraph::add_edge appends to both lists: replay in list order
            for (uint32_t k = adj_ptr[f]; k < adj_ptr[f + 1]; ++k) if (f < adj_idx[k]) graph.add_edge(f, adj_idx[k]);
        tex::DataCosts data_costs(num_faces, static_cast<uint16_t>(num_views));
        for (uint32_t f = 0; f < num_faces; ++f)
            for (uint64_t k = face_ptr[f]; k < face_ptr[f + 1]; ++k) data_costs.set_value(f, view[k], cost[k]);
        tex::Settings settings;
        tex::view_selection(data_costs, &graph, settings);
        mapmap::Capture const& c = mapmap::refshim_capture();
        *num_edges_out = c.edge_weight.size();
        uint32_t* e = static_cast<uint32_t*>(std::malloc(sizeof(uint32_t) * (c.edges.size() ? c.edges.size() : 1)));
        std::copy(c.edges.begin(), c.edges.end(), e);
        *edges_out = e;
        ls_ptr[0] = 0;
        for (uint32_t f = 0; f < num_faces; ++f) ls_ptr[f + 1] = ls_ptr[f] + c.labels[f].size();
        int32_t* ll = static_cast<int32_t*>(std::malloc(sizeof(int32_t) * (ls_ptr[num_faces] ? ls_ptr[num_faces] : 1)));
        float* lc = static_cast<float*>(std::malloc(sizeof(float) * (ls_ptr[num_faces] ? ls_ptr[num_faces] : 1)));
        bool unaries_ok = c.unary_set.size() == num_faces;
        bool weights_ok = true;
        for (std::size_t i = 0; i < c.edge_weight.size(); ++i) weights_ok = weights_ok && c.edge_weight[i] == 1.0f;
        for (uint32_t f = 0; f < num_faces; ++f) {
            if (c.costs[f].size() != c.labels[f].size()) return 2;
            for (std::size_t k = 0; k < c.labels[f].size(); ++k) { ll[ls_ptr[f] + k] = c.labels[f][k]; lc[ls_ptr[f] + k] = c.costs[f][k]; }
            unaries_ok = unaries_ok && c.unary_set[f] == 1;
            labels_out[f] = static_cast<uint32_t>(graph.get_label(f));
        }
        *ls_label_out = ll; *ls_cost_out = lc;
        double p[16] = { c.potts, double(c.window), c.ratio, double(c.ctr.initial_seed), double(c.ctr.sample_deterministic),
                         double(c.ctr.tree_algorithm), double(c.ctr.use_multilevel), double(c.ctr.use_spanning_tree), double(c.ctr.use_acyclic),
                         double(c.ctr.spanning_tree_multilevel_after_n_iterations), double(c.ctr.force_acyclic),
                         double(c.ctr.min_acyclic_iterations), double(c.ctr.relax_acyclic_maximal), double(c.components_updated),
                         double(c.compress), double(unaries_ok && weights_ok) };
        std::copy(p, p + 16, params_out);
        return 0;
    } catch (std::exception& e) {
        std::fprintf(stderr, "ref_view_selection_model: %s\n", e.what());
        return 1;
    }
}

/* texrecon.cpp:160-190: generate_texture_patches -> global_seam_leveling (or the zero-adjust pass) ->
 * local_seam_leveling, on a given labeling.  Results are kept in a static cache and read back with the
 * ref_patches_* getters (two-call protocol: first the counts, then the copies). */
namespace {
struct PatchCache {
    tex::TexturePatches patches;
    tex::VertexProjectionInfos vpi;
    bool has_blending;   /* local_seam_leveling releases the blending masks (:201) */
};
PatchCache g_patches;
}

int ref_seam_leveling(const float* verts, uint32_t num_verts, const uint32_t* faces, uint32_t num_faces,
                      const uint32_t* vf_ptr, const uint32_t* vf_idx, const uint32_t* vv_ptr, const uint32_t* vv_idx,
                      const uint32_t* adj_ptr, const uint32_t* adj_idx, const uint32_t* labels,
                      const orc_view* views, uint32_t num_views, int do_global, int do_local,
                      uint32_t* num_patches_out)
{
    try {
        Quiet q;
        /* one thread: patch ids follow the view order (with more they depend on scheduling, :469,514-518) */
        struct OneThread { int n; OneThread() : n(omp_get_max_threads()) { omp_set_num_threads(1); } ~OneThread() { omp_set_num_threads(n); } } one;
        mve::TriangleMesh::Ptr mesh = make_mesh(verts, num_verts, faces, NULL, num_faces);
        mve::MeshInfo mi;
        fill_mesh_info(num_verts, vf_ptr, vf_idx, vv_ptr, vv_idx, &mi);
        tex::Graph graph(num_faces);
        for (uint32_t f = 0; f < num_faces; ++f)
            for (uint32_t k = adj_ptr[f]; k < adj_ptr[f + 1]; ++k) if (f < adj_idx[k]) graph.add_edge(f, adj_idx[k]);
        for (uint32_t f = 0; f < num_faces; ++f) graph.set_label(f, labels[f]);
        tex::TextureViews tvs;
        make_views(views, num_views, &tvs);
        tex::Settings settings;
        /* fill_hole (:140-451) walks MeshInfo's RING-ORDERED adjacency; the shim hands over ascending ids, and hole
         * filling is not restated by the oracle anyway: switched off, as texrecon --skip_hole_filling does */
        settings.hole_filling = false;
        g_patches.patches.clear(); g_patches.vpi.clear();
        tex::generate_texture_patches(graph, mesh, mi, &tvs, settings, &g_patches.vpi, &g_patches.patches);
        if (do_global) {
            tex::global_seam_leveling(graph, mesh, mi, g_patches.vpi, &g_patches.patches);
        } else {                                               /* texrecon.cpp:174-183 */
            for (std::size_t i = 0; i < g_patches.patches.size(); ++i) {
                TexturePatch::Ptr tp = g_patches.patches[i];
                std::vector<math::Vec3f> zero(tp->get_faces().size() * 3, math::Vec3f(0.0f));
                tp->adjust_colors(zero);
            }
        }
        if (do_local) tex::local_seam_leveling(graph, mesh, g_patches.vpi, &g_patches.patches);
        g_patches.has_blending = !do_local;
        *num_patches_out = static_cast<uint32_t>(g_patches.patches.size());
        return 0;
    } catch (std::exception& e) {
        std::fprintf(stderr, "ref_seam_leveling: %s\n", e.what());
        return 1;
    }
}

/* info[4] = label, width, height, number of faces */
void ref_patch_info(uint32_t id, int32_t* info)
{
    TexturePatch::Ptr tp = g_patches.patches[id];
    info[0] = tp->get_label(); info[1] = tp->get_width(); info[2] = tp->get_height(); info[3] = static_cast<int32_t>(tp->get_faces().size());
}

/* image: h*w*3 floats, validity / blending: h*w bytes, faces: n, texcoords: n*3*2 floats */
void ref_patch_data(uint32_t id, float* image, uint8_t* validity, uint8_t* blending, uint32_t* faces, float* texcoords)
{
    TexturePatch::Ptr tp = g_patches.patches[id];
    std::size_t const px = static_cast<std::size_t>(tp->get_width()) * tp->get_height();
    if (image) std::memcpy(image, tp->get_image()->get_data_pointer(), px * 3 * sizeof(float));
    if (validity) std::memcpy(validity, tp->get_validity_mask()->get_data_pointer(), px);
    if (blending && g_patches.has_blending) std::memcpy(blending, tp->get_blending_mask()->get_data_pointer(), px);
    for (std::size_t i = 0; faces && i < tp->get_faces().size(); ++i) faces[i] = static_cast<uint32_t>(tp->get_faces()[i]);
    for (std::size_t i = 0; texcoords && i < tp->get_texcoords().size(); ++i) { texcoords[2 * i] = tp->get_texcoords()[i][0]; texcoords[2 * i + 1] = tp->get_texcoords()[i][1]; }
}

/* vertex projection infos (after merge_vertex_projection_infos): count, then (patch id, x, y) triples */
uint32_t ref_vertex_projection_count(uint32_t vertex) { return static_cast<uint32_t>(g_patches.vpi[vertex].size()); }
void ref_vertex_projections(uint32_t vertex, uint32_t* patch_ids, float* xy)
{
    std::vector<tex::VertexProjectionInfo> const& v = g_patches.vpi[vertex];
    for (std::size_t i = 0; i < v.size(); ++i) { patch_ids[i] = static_cast<uint32_t>(v[i].texture_patch_id); xy[2 * i] = v[i].projection[0]; xy[2 * i + 1] = v[i].projection[1]; }
}

}  // extern "C"
