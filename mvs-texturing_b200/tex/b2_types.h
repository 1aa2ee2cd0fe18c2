// b2_types.h -- host-side types of the tex:: veneer.
//
// The reference's boundary types come from MVE (absent) and libs/tex; this header provides the
// minimal equivalents the four hot-path signatures need, with the SAME member names texrecon uses
// (apps/texrecon/texrecon.cpp:78-121,166-171), so that a maintainer can either include this header
// stand-alone or replace the shim namespaces by the real MVE headers.
#pragma once

#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace math {
template <typename T, int N>
struct Vector {
    T v[N];
    T &operator[](int i) { return v[i]; }
    T const &operator[](int i) const { return v[i]; }
    T *operator*() { return v; }
    T const *operator*() const { return v; }
};
typedef Vector<float, 2> Vec2f;
typedef Vector<float, 3> Vec3f;
}  // namespace math

namespace mve {
/* mve::TriangleMesh: the three lists the path reads (calculate_data_costs.cpp:136-138) */
class TriangleMesh {
public:
    typedef std::shared_ptr<TriangleMesh> Ptr;
    typedef std::shared_ptr<TriangleMesh const> ConstPtr;
    typedef std::vector<math::Vec3f> VertexList;
    typedef std::vector<unsigned int> FaceList;
    typedef std::vector<math::Vec3f> NormalList;
    static Ptr create() { return Ptr(new TriangleMesh()); }
    VertexList const &get_vertices() const { return vertices; }
    VertexList &get_vertices() { return vertices; }
    FaceList const &get_faces() const { return faces; }
    FaceList &get_faces() { return faces; }
    NormalList const &get_face_normals() const { return face_normals; }
    NormalList &get_face_normals() { return face_normals; }
    /* MVE ensure_normals(face=true): normalised cross(b-a, c-a), zero for degenerate faces */
    void ensure_face_normals();
private:
    VertexList vertices;
    FaceList faces;
    NormalList face_normals;
};

/* mve::MeshInfo: per-vertex incident faces and 1-ring (global_seam_leveling.cpp:55,61,161,187) */
class MeshInfo {
public:
    struct VertexInfo {
        std::vector<std::size_t> verts;
        std::vector<std::size_t> faces;
    };
    MeshInfo() {}
    explicit MeshInfo(TriangleMesh::ConstPtr mesh) { initialize(mesh); }
    void initialize(TriangleMesh::ConstPtr mesh);
    VertexInfo const &operator[](std::size_t i) const { return infos[i]; }
    std::size_t size() const { return infos.size(); }
    /* appends (build_adjacency_graph.cpp:31-34 relies on that) */
    void get_faces_for_edge(std::size_t v1, std::size_t v2, std::vector<std::size_t> *adjacent_faces) const;
private:
    std::vector<VertexInfo> infos;
    std::vector<unsigned int> const *faces = nullptr;
};
}  // namespace mve

namespace tex {

/* libs/tex/settings.h:58-95 */
enum DataTerm { DATA_TERM_AREA = 0, DATA_TERM_GMI = 1 };
enum SmoothnessTerm { SMOOTHNESS_TERM_POTTS = 0 };
enum OutlierRemoval { OUTLIER_REMOVAL_NONE = 0, OUTLIER_REMOVAL_GAUSS_DAMPING = 1, OUTLIER_REMOVAL_GAUSS_CLAMPING = 2 };
enum ToneMapping { TONE_MAPPING_NONE = 0, TONE_MAPPING_GAMMA = 1 };
struct Settings {
    bool verbose = false;
    DataTerm data_term = DATA_TERM_GMI;
    SmoothnessTerm smoothness_term = SMOOTHNESS_TERM_POTTS;
    OutlierRemoval outlier_removal = OUTLIER_REMOVAL_NONE;
    ToneMapping tone_mapping = TONE_MAPPING_NONE;
    bool geometric_visibility_test = true;
    bool global_seam_leveling = true;
    bool local_seam_leveling = true;
    bool hole_filling = true;
    bool keep_unseen_faces = false;
};

/* DataCosts = SparseTable<uint32 face, uint16 view, float> (texturing.h:36, sparse_table.h:29-66):
 * same accessors; stored column-wise only (the row-wise copy is never read on the path). */
template <typename C, typename R, typename T>
class SparseTable {
public:
    typedef std::vector<std::pair<R, T> > Column;
    SparseTable() : nrows(0), nnz(0) {}
    SparseTable(C cols, R rows) : columns(cols), nrows(rows), nnz(0) {}
    C cols() const { return static_cast<C>(columns.size()); }
    R rows() const { return nrows; }
    Column const &col(C id) const { return columns[id]; }
    void set_value(C col, R row, T value) { columns[col].push_back(std::pair<R, T>(row, value)); ++nnz; }
    std::size_t get_nnz() const { return nnz; }
private:
    std::vector<Column> columns;
    R nrows;
    std::size_t nnz;
};
typedef SparseTable<std::uint32_t, std::uint16_t, float> DataCosts;

/* face adjacency graph + labels (libs/tex/uni_graph.h:20-78) */
class UniGraph {
public:
    explicit UniGraph(std::size_t nodes) : adj(nodes), labels(nodes, 0), edges(0) {}
    std::size_t num_nodes() const { return adj.size(); }
    std::size_t num_edges() const { return edges; }
    bool has_edge(std::size_t a, std::size_t b) const {
        for (std::size_t x : adj[a]) if (x == b) return true;
        return false;
    }
    void add_edge(std::size_t a, std::size_t b) {
        if (has_edge(a, b)) return;
        adj[a].push_back(b); adj[b].push_back(a); ++edges;
    }
    std::vector<std::size_t> const &get_adj_nodes(std::size_t n) const { return adj[n]; }
    void set_label(std::size_t n, std::size_t l) { labels[n] = l; }
    std::size_t get_label(std::size_t n) const { return labels[n]; }
private:
    std::vector<std::vector<std::size_t> > adj;
    std::vector<std::size_t> labels;
    std::size_t edges;
};
typedef UniGraph Graph;

/* camera + image of one view (libs/tex/texture_view.h:39-52); the image is borrowed, not loaded:
 * image IO/undistortion (generate_texture_views.cpp) is out of scope */
struct TextureView {
    std::size_t id = 0;
    float pos[3] = {0, 0, 0};
    float viewdir[3] = {0, 0, 1};
    float projection[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // row major
    float world_to_cam[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    int width = 0, height = 0;
    std::uint8_t const *rgb = nullptr;  // H x W x 3
    std::size_t get_id() const { return id; }
    int get_width() const { return width; }
    int get_height() const { return height; }
};
typedef std::vector<TextureView> TextureViews;

/* libs/tex/texture_patch.h:23-69 reduced to what texture atlas packing reads: label, faces, texcoords, the float image and
 * the validity mask.  min_x / min_y = view pixel of patch pixel (0, 0). */
struct TexturePatch {
    int label = 0;
    int min_x = 0, min_y = 0, width = 0, height = 0;
    std::vector<std::size_t> faces;
    std::vector<math::Vec2f> texcoords;       // 3 per face
    std::vector<float> image;                 // height x width x 3
    std::vector<std::uint8_t> validity_mask;  // height x width, 255 = valid
    int get_label() const { return label; }
    int get_width() const { return width; }
    int get_height() const { return height; }
    std::vector<std::size_t> const &get_faces() const { return faces; }
    std::vector<math::Vec2f> const &get_texcoords() const { return texcoords; }
};
typedef std::vector<TexturePatch> TexturePatches;

/* per-(vertex,label) colour adjustment, the product of global_seam_leveling.cpp:251,283-289 */
typedef std::vector<std::map<std::size_t, math::Vec3f> > AdjustValues;

}  // namespace tex
