// refshim: MVE mve::MeshInfo stand-in.  The per-vertex rings (adjacent faces / vertices, in MVE's
// order) are inputs handed over by the glue -- the same arrays the oracle and the C ABI take
// (b2tex_set_vertex_rings).  get_faces_for_edge restates MVE: faces of v1, in ring order, that also
// contain v2.  (see ../README.md)
#pragma once
#include <algorithm>
#include <vector>
#include "mve/mesh.h"

namespace mve {

class MeshInfo {
public:
    enum VertexClass { VERTEX_CLASS_SIMPLE, VERTEX_CLASS_COMPLEX, VERTEX_CLASS_BORDER, VERTEX_CLASS_UNREF };
    typedef std::vector<std::size_t> AdjacentVertices;
    typedef std::vector<std::size_t> AdjacentFaces;
    struct VertexInfo { VertexClass vclass; AdjacentVertices verts; AdjacentFaces faces; };

    MeshInfo() {}
    explicit MeshInfo(TriangleMesh::ConstPtr) {}
    void initialize(TriangleMesh::ConstPtr) {}
    void clear() { vertex_info.clear(); }
    std::size_t size() const { return vertex_info.size(); }
    VertexInfo& operator[](std::size_t i) { return vertex_info[i]; }
    VertexInfo const& operator[](std::size_t i) const { return vertex_info[i]; }
    VertexInfo& at(std::size_t i) { return vertex_info.at(i); }
    VertexInfo const& at(std::size_t i) const { return vertex_info.at(i); }
    void resize(std::size_t n) { vertex_info.resize(n); }

    void get_faces_for_edge(std::size_t v1, std::size_t v2, std::vector<std::size_t>* adjacent_faces) const {
        AdjacentFaces const& f1 = vertex_info[v1].faces;
        AdjacentFaces const& f2 = vertex_info[v2].faces;
        for (std::size_t i = 0; i < f1.size(); ++i)
            if (std::find(f2.begin(), f2.end(), f1[i]) != f2.end()) adjacent_faces->push_back(f1[i]);
    }

private:
    std::vector<VertexInfo> vertex_info;
};

}  // namespace mve
