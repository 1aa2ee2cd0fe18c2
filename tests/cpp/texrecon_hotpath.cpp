// texrecon_hotpath.cpp -- the hot-path slice of apps/texrecon/texrecon.cpp:92-121,171 written against
// the tex:: veneer: proves that the four signatures compile and link against libb2tex.so.  With a GPU
// it runs a tetrahedron seen by four procedural views and prints labels; without one it reports the
// library's error (no CPU fallback) and exits 0 when invoked with --link-only.
#include <cmath>
#include <cstdio>
#include <string>
#include <cstring>
#include <vector>

#include "../../mvs-texturing_b200/tex/texturing.h"

int main(int argc, char **argv)
{
    bool link_only = argc > 1 && !std::strcmp(argv[1], "--link-only");
    mve::TriangleMesh::Ptr mesh = mve::TriangleMesh::create();
    float V[4][3] = {{1, 1, 1}, {-1, -1, 1}, {-1, 1, -1}, {1, -1, -1}};
    unsigned int Fc[4][3] = {{0, 2, 1}, {0, 1, 3}, {0, 3, 2}, {1, 2, 3}};  // outward normals
    for (auto &v : V) { math::Vec3f p; p[0] = v[0]; p[1] = v[1]; p[2] = v[2]; mesh->get_vertices().push_back(p); }
    for (auto &f : Fc) for (unsigned int k : f) mesh->get_faces().push_back(k);
    mesh->ensure_face_normals();
    mve::MeshInfo mesh_info(mesh);
    tex::Graph graph(4);
    tex::build_adjacency_graph(mesh, mesh_info, &graph);
    std::printf("adjacency edges: %zu\n", graph.num_edges());
    if (link_only) return graph.num_edges() == 6 ? 0 : 1;

    const int W = 160, H = 120;
    std::vector<std::vector<unsigned char> > images(4, std::vector<unsigned char>(W * H * 3));
    tex::TextureViews views(4);
    for (int k = 0; k < 4; ++k) {
        for (int i = 0; i < W * H * 3; ++i) images[k][i] = (unsigned char)(1 + (i * 37 + k * 11 + (i / 3 / W) * 5) % 250);
        tex::TextureView &tv = views[k];
        float n = std::sqrt(3.0f);
        // camera at -3*V[k]/|V| looking at the origin: sees the face opposite to vertex k
        float pos[3] = {-3 * V[k][0] / n, -3 * V[k][1] / n, -3 * V[k][2] / n};
        float zc[3] = {-pos[0] / 3, -pos[1] / 3, -pos[2] / 3};
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
        float P[9] = {55, 0, W / 2.0f, 0, 55, H / 2.0f, 0, 0, 1};
        std::memcpy(tv.projection, P, sizeof(P));
        tv.width = W; tv.height = H; tv.rgb = images[k].data(); tv.id = k;
    }
    try {
        tex::Settings settings;
        tex::DataCosts data_costs(4, 4);
        tex::calculate_data_costs(mesh, &views, settings, &data_costs);
        tex::view_selection(data_costs, &graph, settings);
        tex::AdjustValues adjust;
        tex::global_seam_leveling(graph, mesh, mesh_info, views, &adjust);
        std::printf("nnz=%zu labels=%zu %zu %zu %zu\n", data_costs.get_nnz(), graph.get_label(0), graph.get_label(1),
                    graph.get_label(2), graph.get_label(3));
        if (argc > 1 && std::string(argv[1]) == "--patches") {          // texrecon.cpp:160-189
            tex::TexturePatches patches;
            tex::seam_leveling(graph, mesh, mesh_info, views, settings, &patches);
            std::size_t faces = 0, valid = 0;
            for (tex::TexturePatch const &p : patches) {
                faces += p.get_faces().size();
                for (std::uint8_t v : p.validity_mask) valid += v == 255;
            }
            std::printf("patches=%zu faces=%zu valid_pixels=%zu\n", patches.size(), faces, valid);
        }
    } catch (std::exception const &e) {
        std::printf("tex:: call failed: %s\n", e.what());
        return 2;
    }
    return 0;
}
