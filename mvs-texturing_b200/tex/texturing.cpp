// texturing.cpp -- C++ veneer: flattens the tex:: types to the C ABI of include/b2tex.h.
#include "texturing.h"

#include <cmath>
#include <cstring>
#include <limits>

#include "../../include/b2tex.h"

namespace mve {

void TriangleMesh::ensure_face_normals()
{
    std::size_t nf = faces.size() / 3;
    face_normals.resize(nf);
    for (std::size_t f = 0; f < nf; ++f) {
        math::Vec3f const &a = vertices[faces[3 * f]], &b = vertices[faces[3 * f + 1]], &c = vertices[faces[3 * f + 2]];
        float u[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, v[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
        float n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
        float l = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        for (int k = 0; k < 3; ++k) face_normals[f][k] = l > 0.0f ? n[k] / l : 0.0f;
    }
}

void MeshInfo::initialize(TriangleMesh::ConstPtr mesh)
{
    faces = &mesh->get_faces();
    infos.assign(mesh->get_vertices().size(), VertexInfo());
    std::size_t nf = faces->size() / 3;
    for (std::size_t f = 0; f < nf; ++f)
        for (int k = 0; k < 3; ++k) {
            std::size_t v = (*faces)[3 * f + k];
            infos[v].faces.push_back(f);
            for (int j = 0; j < 3; ++j) {
                std::size_t w = (*faces)[3 * f + j];
                if (w == v) continue;
                bool have = false;
                for (std::size_t x : infos[v].verts) have = have || x == w;
                if (!have) infos[v].verts.push_back(w);
            }
        }
}

void MeshInfo::get_faces_for_edge(std::size_t v1, std::size_t v2, std::vector<std::size_t> *out) const
{
    for (std::size_t f : infos[v1].faces)
        if ((*faces)[3 * f] == v2 || (*faces)[3 * f + 1] == v2 || (*faces)[3 * f + 2] == v2) out->push_back(f);
}

}  // namespace mve

namespace tex {

namespace {
void check(int rc)
{
    if (rc != B2TEX_OK) throw std::runtime_error(b2tex_last_error());
}

std::vector<b2tex_view> flatten_views(TextureViews const &tv)
{
    std::vector<b2tex_view> v(tv.size());
    for (std::size_t i = 0; i < tv.size(); ++i) {
        std::memcpy(v[i].pos, tv[i].pos, sizeof(v[i].pos));
        std::memcpy(v[i].viewdir, tv[i].viewdir, sizeof(v[i].viewdir));
        std::memcpy(v[i].proj, tv[i].projection, sizeof(v[i].proj));
        std::memcpy(v[i].w2c, tv[i].world_to_cam, sizeof(v[i].w2c));
        v[i].width = tv[i].width;
        v[i].height = tv[i].height;
        v[i].rgb = tv[i].rgb;
    }
    return v;
}

void flatten_graph(UniGraph const &g, std::vector<std::uint32_t> *ptr, std::vector<std::uint32_t> *idx)
{
    ptr->assign(g.num_nodes() + 1, 0);
    for (std::size_t i = 0; i < g.num_nodes(); ++i) (*ptr)[i + 1] = (*ptr)[i] + (std::uint32_t)g.get_adj_nodes(i).size();
    idx->resize((*ptr)[g.num_nodes()]);
    for (std::size_t i = 0; i < g.num_nodes(); ++i) {
        std::uint32_t o = (*ptr)[i];
        for (std::size_t a : g.get_adj_nodes(i)) (*idx)[o++] = (std::uint32_t)a;
    }
}
}  // namespace

/* build_adjacency_graph.cpp:16-53 */
void build_adjacency_graph(mve::TriangleMesh::ConstPtr mesh, mve::MeshInfo const &mesh_info, UniGraph *graph)
{
    mve::TriangleMesh::FaceList const &faces = mesh->get_faces();
    for (std::size_t f = 0; f < faces.size() / 3; ++f) {
        std::vector<std::size_t> nb;
        mesh_info.get_faces_for_edge(faces[3 * f], faces[3 * f + 1], &nb);
        mesh_info.get_faces_for_edge(faces[3 * f + 1], faces[3 * f + 2], &nb);
        mesh_info.get_faces_for_edge(faces[3 * f + 2], faces[3 * f], &nb);
        for (std::size_t g : nb)
            if (g != f) graph->add_edge(f, g);
    }
}

void calculate_data_costs(mve::TriangleMesh::ConstPtr mesh, TextureViews *texture_views,
                          Settings const &settings, DataCosts *data_costs)
{
    std::size_t const num_faces = mesh->get_faces().size() / 3;
    std::size_t const num_views = texture_views->size();
    if (num_faces > std::numeric_limits<std::uint32_t>::max()) throw std::runtime_error("Exeeded maximal number of faces");
    if (num_views > std::numeric_limits<std::uint16_t>::max()) throw std::runtime_error("Exeeded maximal number of views");
    std::vector<b2tex_view> views = flatten_views(*texture_views);
    b2tex_settings st = {(int)settings.data_term, (int)settings.outlier_removal, settings.geometric_visibility_test ? 1 : 0};
    std::uint64_t *fp = nullptr;
    std::uint16_t *vw = nullptr;
    float *cs = nullptr;
    b2tex_dc_info info;
    check(b2tex_calculate_data_costs(*mesh->get_vertices()[0], (std::uint32_t)mesh->get_vertices().size(),
                                     mesh->get_faces().data(), *mesh->get_face_normals()[0], (std::uint32_t)num_faces,
                                     views.data(), (std::uint32_t)num_views, &st, &fp, &vw, &cs, &info));
    for (std::uint32_t f = 0; f < num_faces; ++f)
        for (std::uint64_t k = fp[f]; k < fp[f + 1]; ++k) data_costs->set_value(f, vw[k], cs[k]);
    b2tex_free(fp); b2tex_free(vw); b2tex_free(cs);
}

void view_selection(DataCosts const &data_costs, UniGraph *graph, Settings const &)
{
    std::uint32_t const F = data_costs.cols();
    std::vector<std::uint64_t> fp(F + 1, 0);
    for (std::uint32_t i = 0; i < F; ++i) fp[i + 1] = fp[i] + data_costs.col(i).size();
    std::vector<std::uint16_t> vw(fp[F]);
    std::vector<float> cs(fp[F]);
    for (std::uint32_t i = 0; i < F; ++i) {
        std::uint64_t o = fp[i];
        for (auto const &e : data_costs.col(i)) { vw[o] = e.first; cs[o] = e.second; ++o; }
    }
    std::vector<std::uint32_t> ap, ai, labels(F);
    flatten_graph(*graph, &ap, &ai);
    b2tex_mrf_info info;
    check(b2tex_view_selection(F, ap.data(), ai.data(), fp.data(), vw.data(), cs.data(), nullptr, labels.data(), &info));
    std::size_t const num_labels = (std::size_t)data_costs.rows() + 1;  // view_selection.cpp:121-131
    for (std::uint32_t i = 0; i < F; ++i) {
        if (num_labels <= labels[i]) throw std::runtime_error("Incorrect labeling");
        graph->set_label(i, labels[i]);
    }
}

void global_seam_leveling(UniGraph const &graph, mve::TriangleMesh::ConstPtr mesh, mve::MeshInfo const &mesh_info,
                          TextureViews const &texture_views, AdjustValues *adjust_values)
{
    std::uint32_t const Vn = (std::uint32_t)mesh->get_vertices().size();
    std::uint32_t const F = (std::uint32_t)(mesh->get_faces().size() / 3);
    std::vector<std::uint32_t> vf_ptr(Vn + 1, 0), vv_ptr(Vn + 1, 0), vf_idx, vv_idx, labels(F), row_ptr(Vn + 1);
    for (std::uint32_t v = 0; v < Vn; ++v) {
        vf_ptr[v + 1] = vf_ptr[v] + (std::uint32_t)mesh_info[v].faces.size();
        vv_ptr[v + 1] = vv_ptr[v] + (std::uint32_t)mesh_info[v].verts.size();
        for (std::size_t f : mesh_info[v].faces) vf_idx.push_back((std::uint32_t)f);
        for (std::size_t w : mesh_info[v].verts) vv_idx.push_back((std::uint32_t)w);
    }
    for (std::uint32_t f = 0; f < F; ++f) labels[f] = (std::uint32_t)graph.get_label(f);
    std::vector<b2tex_view> views = flatten_views(texture_views);
    std::uint32_t *row_label = nullptr;
    float *x = nullptr;
    b2tex_seam_info info;
    check(b2tex_global_seam_leveling(*mesh->get_vertices()[0], Vn, mesh->get_faces().data(), F, vf_ptr.data(),
                                     vf_idx.data(), vv_ptr.data(), vv_idx.data(), labels.data(), views.data(),
                                     (std::uint32_t)views.size(), row_ptr.data(), &row_label, &x, &info));
    adjust_values->assign(Vn, std::map<std::size_t, math::Vec3f>());
    for (std::uint32_t v = 0; v < Vn; ++v)
        for (std::uint32_t r = row_ptr[v]; r < row_ptr[v + 1]; ++r) {
            math::Vec3f a;
            a[0] = x[3 * r]; a[1] = x[3 * r + 1]; a[2] = x[3 * r + 2];
            (*adjust_values)[v][row_label[r]] = a;
        }
    b2tex_free(row_label); b2tex_free(x);
}

void seam_leveling(UniGraph const &graph, mve::TriangleMesh::ConstPtr mesh, mve::MeshInfo const &mesh_info,
                   TextureViews const &texture_views, Settings const &settings, TexturePatches *texture_patches)
{
    std::uint32_t const Vn = (std::uint32_t)mesh->get_vertices().size();
    std::uint32_t const F = (std::uint32_t)(mesh->get_faces().size() / 3);
    std::vector<std::uint32_t> vf_ptr(Vn + 1, 0), vv_ptr(Vn + 1, 0), vf_idx, vv_idx, labels(F), ap, ai;
    for (std::uint32_t v = 0; v < Vn; ++v) {
        vf_ptr[v + 1] = vf_ptr[v] + (std::uint32_t)mesh_info[v].faces.size();
        vv_ptr[v + 1] = vv_ptr[v] + (std::uint32_t)mesh_info[v].verts.size();
        for (std::size_t f : mesh_info[v].faces) vf_idx.push_back((std::uint32_t)f);
        for (std::size_t w : mesh_info[v].verts) vv_idx.push_back((std::uint32_t)w);
    }
    for (std::uint32_t f = 0; f < F; ++f) labels[f] = (std::uint32_t)graph.get_label(f);
    flatten_graph(graph, &ap, &ai);
    std::vector<b2tex_view> views = flatten_views(texture_views);
    std::int32_t *desc = nullptr;
    std::uint32_t *faces = nullptr;
    float *tex = nullptr, *img = nullptr;
    std::uint8_t *val = nullptr;
    b2tex_patch_info pinfo;
    check(b2tex_seam_leveling_patches(*mesh->get_vertices()[0], Vn, mesh->get_faces().data(), F, ap.data(), ai.data(), vf_ptr.data(),
                                      vf_idx.data(), vv_ptr.data(), vv_idx.data(), labels.data(), views.data(),
                                      (std::uint32_t)views.size(), settings.global_seam_leveling ? 1 : 0,
                                      settings.local_seam_leveling ? 1 : 0, &desc, &faces, &tex, &img, &val, &pinfo, nullptr, nullptr));
    texture_patches->clear();
    texture_patches->resize(pinfo.num_patches);
    std::size_t off = 0;
    for (std::uint32_t q = 0; q < pinfo.num_patches; ++q) {
        std::int32_t const *d = desc + 8 * (std::size_t)q;
        TexturePatch &p = (*texture_patches)[q];
        p.label = d[0]; p.min_x = d[1]; p.min_y = d[2]; p.width = d[3]; p.height = d[4];
        std::size_t const first = (std::size_t)d[5], n = (std::size_t)d[6], px = (std::size_t)d[3] * (std::size_t)d[4];
        p.faces.assign(faces + first, faces + first + n);
        p.texcoords.resize(3 * n);
        for (std::size_t i = 0; i < 3 * n; ++i) { p.texcoords[i][0] = tex[2 * (3 * first + i)]; p.texcoords[i][1] = tex[2 * (3 * first + i) + 1]; }
        p.image.assign(img + 3 * off, img + 3 * (off + px));
        p.validity_mask.assign(val + off, val + off + px);
        off += px;
    }
    b2tex_free(desc); b2tex_free(faces); b2tex_free(tex); b2tex_free(img); b2tex_free(val);
}

}  // namespace tex
