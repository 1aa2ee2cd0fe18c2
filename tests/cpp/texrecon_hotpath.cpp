// texrecon_hotpath.cpp -- the call sequence of apps/texrecon/texrecon.cpp:92-189 written against the tex:: veneer
// (mvs-texturing_b200/tex/texturing.h): build_adjacency_graph (:92), calculate_data_costs (:100), view_selection (:121),
// generate_texture_patches (:166), global_seam_leveling (:171) or the zero-adjust loop (:174-183), local_seam_leveling
// (:188) -- same function names, same argument lists, same types.  What texrecon does before (mesh / scene loading) and
// after (atlas packing, OBJ) is replaced by a procedural scene and a few prints.
//
//   texrecon_hotpath --link-only        no GPU needed: the signatures compile and link against libb2tex.so
//   texrecon_hotpath [--sphere N] [--no-global] [--no-local] [--labels-from-host]
// Without a GPU the first device call reports the library's error (no CPU fallback) and the program exits with 2.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../mvs-texturing_b200/tex/texturing.h"

namespace {

struct Arguments {   // apps/texrecon/arguments.h:17-35, the fields the slice uses
    tex::Settings settings;
    bool write_intermediate_results = false;
};

double now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// a camera at `pos` looking at the origin, focal length f (pixels)
void look_at(tex::TextureView &tv, float const pos[3], float f, int W, int H)
{
    float n = std::sqrt(pos[0] * pos[0] + pos[1] * pos[1] + pos[2] * pos[2]);
    float zc[3] = {-pos[0] / n, -pos[1] / n, -pos[2] / n};
    float up[3] = {0, 0, 1};
    if (std::fabs(zc[2]) > 0.9f) { up[0] = 0; up[1] = 1; up[2] = 0; }
    float xc[3] = {zc[1] * up[2] - zc[2] * up[1], zc[2] * up[0] - zc[0] * up[2], zc[0] * up[1] - zc[1] * up[0]};
    float xn = std::sqrt(xc[0] * xc[0] + xc[1] * xc[1] + xc[2] * xc[2]);
    for (float &c : xc) c /= xn;
    float yc[3] = {zc[1] * xc[2] - zc[2] * xc[1], zc[2] * xc[0] - zc[0] * xc[2], zc[0] * xc[1] - zc[1] * xc[0]};
    float R[3][3] = {{xc[0], xc[1], xc[2]}, {yc[0], yc[1], yc[2]}, {zc[0], zc[1], zc[2]}};
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) tv.world_to_cam[4 * r + c] = R[r][c];
        tv.world_to_cam[4 * r + 3] = -(R[r][0] * pos[0] + R[r][1] * pos[1] + R[r][2] * pos[2]);
    }
    for (int c = 0; c < 3; ++c) { tv.pos[c] = pos[c]; tv.viewdir[c] = zc[c]; }
    float P[9] = {f, 0, W / 2.0f, 0, f, H / 2.0f, 0, 0, 1};
    std::memcpy(tv.projection, P, sizeof(P));
    tv.width = W; tv.height = H;
}

// icosphere-like mesh: an octahedron subdivided `level` times, projected onto the unit sphere
mve::TriangleMesh::Ptr make_sphere(int level)
{
    mve::TriangleMesh::Ptr mesh = mve::TriangleMesh::create();
    std::vector<math::Vec3f> &V = mesh->get_vertices();
    std::vector<unsigned int> &F = mesh->get_faces();
    float const o[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    for (auto const &p : o) V.push_back(math::Vec3f(p[0], p[1], p[2]));
    unsigned int const f0[8][3] = {{0, 2, 4}, {2, 1, 4}, {1, 3, 4}, {3, 0, 4}, {2, 0, 5}, {1, 2, 5}, {3, 1, 5}, {0, 3, 5}};
    for (auto const &f : f0) for (unsigned int k : f) F.push_back(k);
    for (int l = 0; l < level; ++l) {
        std::map<std::pair<unsigned, unsigned>, unsigned> mid;
        auto midpoint = [&](unsigned a, unsigned b) {
            std::pair<unsigned, unsigned> key(std::min(a, b), std::max(a, b));
            auto it = mid.find(key);
            if (it != mid.end()) return it->second;
            math::Vec3f m((V[a][0] + V[b][0]) / 2, (V[a][1] + V[b][1]) / 2, (V[a][2] + V[b][2]) / 2);
            float n = std::sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
            for (int k = 0; k < 3; ++k) m[k] /= n;
            V.push_back(m);
            return mid[key] = (unsigned)V.size() - 1;
        };
        std::vector<unsigned int> G;
        for (std::size_t f = 0; f < F.size() / 3; ++f) {
            unsigned a = F[3 * f], b = F[3 * f + 1], c = F[3 * f + 2];
            unsigned ab = midpoint(a, b), bc = midpoint(b, c), ca = midpoint(c, a);
            unsigned const t[4][3] = {{a, ab, ca}, {ab, b, bc}, {ca, bc, c}, {ab, bc, ca}};
            for (auto const &q : t) for (unsigned k : q) G.push_back(k);
        }
        F.swap(G);
    }
    mesh->ensure_face_normals();
    return mesh;
}

}  // namespace

int main(int argc, char **argv)
{
    bool link_only = false, labels_from_host = false;
    int level = 0;
    Arguments conf;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--link-only")) link_only = true;
        else if (!std::strcmp(argv[i], "--sphere") && i + 1 < argc) level = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--no-global")) conf.settings.global_seam_leveling = false;
        else if (!std::strcmp(argv[i], "--no-local")) conf.settings.local_seam_leveling = false;
        else if (!std::strcmp(argv[i], "--labels-from-host")) labels_from_host = true;
        else if (!std::strcmp(argv[i], "--write-intermediate")) conf.write_intermediate_results = true;
    }

    mve::TriangleMesh::Ptr mesh;
    if (level > 0) mesh = make_sphere(level);
    else {   // a tetrahedron
        mesh = mve::TriangleMesh::create();
        float V[4][3] = {{1, 1, 1}, {-1, -1, 1}, {-1, 1, -1}, {1, -1, -1}};
        unsigned int Fc[4][3] = {{0, 2, 1}, {0, 1, 3}, {0, 3, 2}, {1, 2, 3}};  // outward normals
        for (auto &v : V) mesh->get_vertices().push_back(math::Vec3f(v[0], v[1], v[2]));
        for (auto &f : Fc) for (unsigned int k : f) mesh->get_faces().push_back(k);
        mesh->ensure_face_normals();
    }
    mve::MeshInfo mesh_info(mesh);                                            /* texrecon.cpp:78 */
    std::size_t const num_faces = mesh->get_faces().size() / 3;

    /* texture views (texrecon.cpp:83 generate_texture_views): procedural images, cameras around the object */
    int const W = level > 0 ? 640 : 160, H = level > 0 ? 480 : 120;
    std::size_t const num_views = level > 0 ? 14 : 4;
    std::vector<std::vector<unsigned char> > images(num_views, std::vector<unsigned char>((std::size_t)W * H * 3));
    tex::TextureViews texture_views(num_views);
    for (std::size_t k = 0; k < num_views; ++k) {
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x)
                for (int c = 0; c < 3; ++c) {
                    float v = 120.0f + 60.0f * std::sin(0.11f * x + 0.3f * c + k) * std::cos(0.07f * y - 0.2f * k) +
                              40.0f * std::sin(0.31f * (x + y)) + 6.0f * (float)k;
                    images[k][((std::size_t)y * W + x) * 3 + c] = (unsigned char)std::min(255.0f, std::max(1.0f, v));
                }
        float pos[3];
        if (level > 0) {   // Fibonacci sphere, radius 3
            float t = (k + 0.5f) / num_views, phi = 2.399963f * k, z = 1.0f - 2.0f * t, r = std::sqrt(std::max(0.0f, 1.0f - z * z));
            pos[0] = 3 * r * std::cos(phi); pos[1] = 3 * r * std::sin(phi); pos[2] = 3 * z;
        } else {
            float const V[4][3] = {{1, 1, 1}, {-1, -1, 1}, {-1, 1, -1}, {1, -1, -1}};
            float n = std::sqrt(3.0f);
            for (int c = 0; c < 3; ++c) pos[c] = -3 * V[k][c] / n;   // sees the face opposite to vertex k
        }
        look_at(texture_views[k], pos, level > 0 ? 520.0f : 55.0f, W, H);
        texture_views[k].rgb = images[k].data();
        texture_views[k].id = k;
    }

    tex::Graph graph(num_faces);                                               /* texrecon.cpp:91-92 */
    tex::build_adjacency_graph(mesh, mesh_info, &graph);
    std::printf("adjacency edges: %zu\n", graph.num_edges());
    if (link_only) return graph.num_edges() == (level > 0 ? 3 * num_faces / 2 : 6) ? 0 : 1;

    try {
        double t0 = now();
        {                                                                      /* texrecon.cpp:97-127 */
            tex::DataCosts data_costs(num_faces, texture_views.size());
            tex::calculate_data_costs(mesh, &texture_views, conf.settings, &data_costs);
            double t1 = now();
            std::printf("data costs: %zu entries, %.1f ms\n", data_costs.get_nnz(), 1e3 * (t1 - t0));
            if (conf.write_intermediate_results || labels_from_host) {         /* :102-106 reads every column */
                std::size_t n = 0;
                for (std::uint32_t f = 0; f < data_costs.cols(); ++f) n += data_costs.col(f).size();
                std::printf("data costs on the host: %zu entries\n", n);
            }
            if (labels_from_host) {   // what -D file does (texrecon.cpp:107-117): a table filled through set_value
                tex::DataCosts copy(num_faces, texture_views.size());
                for (std::uint32_t f = 0; f < data_costs.cols(); ++f)
                    for (auto const &e : data_costs.col(f)) copy.set_value(f, e.first, e.second);
                tex::view_selection(copy, &graph, conf.settings);
            } else {
                tex::view_selection(data_costs, &graph, conf.settings);        /* :121 */
            }
            std::printf("view selection: %.1f ms\n", 1e3 * (now() - t1));
        }
        std::size_t unseen = 0;
        std::vector<std::size_t> hist(texture_views.size() + 1, 0);
        for (std::size_t f = 0; f < num_faces; ++f) { ++hist[graph.get_label(f)]; unseen += graph.get_label(f) == 0; }
        std::printf("labels:");
        for (std::size_t f = 0; f < std::min<std::size_t>(num_faces, 8); ++f) std::printf(" %zu", graph.get_label(f));
        std::printf("  unseen=%zu\n", unseen);

        double t2 = now();
        tex::TexturePatches texture_patches;                                   /* texrecon.cpp:162-166 */
        tex::VertexProjectionInfos vertex_projection_infos;
        tex::generate_texture_patches(graph, mesh, mesh_info, &texture_views, conf.settings, &vertex_projection_infos,
                                      &texture_patches);
        if (conf.settings.global_seam_leveling) {                              /* :168-172 */
            tex::global_seam_leveling(graph, mesh, mesh_info, vertex_projection_infos, &texture_patches);
        } else {                                                               /* :173-183 */
            for (std::size_t i = 0; i < texture_patches.size(); ++i) {
                TexturePatch::Ptr texture_patch = texture_patches[i];
                std::vector<math::Vec3f> patch_adjust_values(texture_patch->get_faces().size() * 3, math::Vec3f(0.0f));
                texture_patch->adjust_colors(patch_adjust_values);
            }
        }
        if (conf.settings.local_seam_leveling)                                 /* :186-189 */
            tex::local_seam_leveling(graph, mesh, vertex_projection_infos, &texture_patches);
        double t3 = now();

        /* what generate_texture_atlases reads (texture_atlas.cpp): sizes, images, validity masks */
        std::size_t faces = 0, valid = 0, pixels = 0, vinfos = 0;
        double sum = 0.0;
        for (TexturePatch::Ptr const &p : texture_patches) {
            faces += p->get_faces().size();
            pixels += (std::size_t)p->get_size();
            mve::ByteImage::ConstPtr mask = p->get_validity_mask();
            mve::FloatImage::ConstPtr img = p->get_image();
            for (int i = 0; i < p->get_size(); ++i) {
                valid += mask->at(i) == 255;
                sum += img->at((std::size_t)3 * i);
            }
        }
        for (auto const &v : vertex_projection_infos) vinfos += v.size();
        std::printf("patches=%zu faces=%zu pixels=%zu valid_pixels=%zu vertex_infos=%zu mean_red=%.4f  (%.1f ms, total %.1f ms)\n",
                    texture_patches.size(), faces, pixels, valid, vinfos, pixels ? sum / pixels : 0.0, 1e3 * (t3 - t2), 1e3 * (t3 - t0));
        if (faces + unseen != num_faces) { std::printf("patch faces do not cover the seen faces\n"); return 3; }
        if (conf.settings.global_seam_leveling) {
            tex::AdjustValues adjust;
            tex::get_adjust_values(texture_patches, &adjust);
            std::size_t rows = 0;
            for (auto const &m : adjust) rows += m.size();
            std::printf("adjust values: %zu (vertex, label) pairs\n", rows);
        }
        tex::release_device_session();
    } catch (std::exception const &e) {
        std::printf("tex:: call failed: %s\n", e.what());
        return 2;
    }
    return 0;
}
