// texturing.h -- the tex:: hot-path API of libs/tex/texturing.h:59-101, backed by libb2tex.so.
//
//   build_adjacency_graph   texturing.h:59-61   (host, feeds view_selection)
//   calculate_data_costs    texturing.h:66-69   -> b2tex_calculate_data_costs
//   view_selection          texturing.h:79-80   -> b2tex_view_selection
//   global_seam_leveling    texturing.h:97-101  -> b2tex_global_seam_leveling (up to adjust_values)
// Errors: std::runtime_error with the reference's messages (calculate_data_costs.cpp:315-318,
// view_selection.cpp:126-128); CUDA failures also surface as std::runtime_error.  No CPU fallback.
#pragma once
#include "b2_types.h"

namespace tex {

void build_adjacency_graph(mve::TriangleMesh::ConstPtr mesh, mve::MeshInfo const &mesh_info, UniGraph *graph);

void calculate_data_costs(mve::TriangleMesh::ConstPtr mesh, TextureViews *texture_views,
                          Settings const &settings, DataCosts *data_costs);

void view_selection(DataCosts const &data_costs, UniGraph *graph, Settings const &settings);

/* The reference mutates TexturePatches (global_seam_leveling.cpp:293-323); patch generation is not
 * on this path yet, so the veneer stops at the adjust values the patches consume and
 * samples colours from the views directly. */
void global_seam_leveling(UniGraph const &graph, mve::TriangleMesh::ConstPtr mesh,
                          mve::MeshInfo const &mesh_info, TextureViews const &texture_views,
                          AdjustValues *adjust_values);

/* texrecon.cpp:160-189 in one call: tex::generate_texture_patches (seen faces; hole filling is not built),
 * tex::global_seam_leveling (settings.global_seam_leveling, else the zero-offset validity pass) and
 * tex::local_seam_leveling (settings.local_seam_leveling) -> b2tex_seam_leveling_patches.  The reference threads
 * TexturePatches through three calls; on the device the patches stay resident between them, so the veneer offers the
 * sequence as one function and returns the finished patches. */
void seam_leveling(UniGraph const &graph, mve::TriangleMesh::ConstPtr mesh, mve::MeshInfo const &mesh_info,
                   TextureViews const &texture_views, Settings const &settings, TexturePatches *texture_patches);

}  // namespace tex
