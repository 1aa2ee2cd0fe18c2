// refshim: MVE mve/mesh_io_ply.h stand-in: debug PLY output is a no-op (see ../README.md)
#pragma once
#include <string>
#include "mve/mesh.h"

namespace mve { namespace geom {
struct SavePLYOptions { bool write_vertex_colors, write_vertex_normals, write_face_colors, write_face_normals, format_binary;
    SavePLYOptions() : write_vertex_colors(false), write_vertex_normals(false), write_face_colors(false), write_face_normals(false), format_binary(true) {} };
inline void save_ply_mesh(TriangleMesh::ConstPtr, std::string const&, SavePLYOptions const& = SavePLYOptions()) {}
} }
