// refshim: MVE mve::TriangleMesh stand-in: plain containers, filled by the glue (see ../README.md)
#pragma once
#include <memory>
#include <vector>
#include "math/vector.h"

namespace mve {

class TriangleMesh {
public:
    typedef std::shared_ptr<TriangleMesh> Ptr;
    typedef std::shared_ptr<TriangleMesh const> ConstPtr;
    typedef unsigned int VertexID;
    typedef std::vector<math::Vec3f> VertexList;
    typedef std::vector<math::Vec3f> NormalList;
    typedef std::vector<math::Vec4f> ColorList;
    typedef std::vector<math::Vec2f> TexCoordList;
    typedef std::vector<VertexID> FaceList;

    static Ptr create() { return Ptr(new TriangleMesh()); }
    VertexList& get_vertices() { return vertices; }
    VertexList const& get_vertices() const { return vertices; }
    FaceList& get_faces() { return faces; }
    FaceList const& get_faces() const { return faces; }
    NormalList& get_face_normals() { return face_normals; }
    NormalList const& get_face_normals() const { return face_normals; }
    NormalList& get_vertex_normals() { return vertex_normals; }
    NormalList const& get_vertex_normals() const { return vertex_normals; }
    ColorList& get_vertex_colors() { return vertex_colors; }
    ColorList const& get_vertex_colors() const { return vertex_colors; }
    TexCoordList& get_vertex_texcoords() { return vertex_texcoords; }
    TexCoordList const& get_vertex_texcoords() const { return vertex_texcoords; }
    bool has_vertex_colors() const { return !vertices.empty() && vertex_colors.size() == vertices.size(); }

private:
    VertexList vertices;
    FaceList faces;
    NormalList face_normals, vertex_normals;
    ColorList vertex_colors;
    TexCoordList vertex_texcoords;
};

}  // namespace mve
