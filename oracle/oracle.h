/*
 * oracle/ -- CPU restatement of the mvs-texturing hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker (or as the timed CPU baseline).  The product path
 * (mvs-texturing_b200/) never links, imports or calls anything in this directory.
 *
 * PARITY PINNING: the reference ships no tests, golden vectors or fixtures, and its build cannot run here
 * (MVE, rayint, Eigen 3.3.2, mapMAP are network-fetched).  What CAN be done is done:
 * the reference's own translation units (libs/tex/*.cpp, unmodified, compiled where they lie) are built
 * against hand-written type shims of those four libraries (oracle/refshim/, `make ref` ->
 * oracle/_ref/libtexref.so) and tests/test_ref_pinning.py holds this oracle to them: data costs, adjacency,
 * MRF model + label decoding, texture patches bit for bit; global / local seam leveling to 2e-5.
 * STILL UNPINNED (parity "partial" for these): everything the shims themselves provide -- vector/matrix
 * operation order, bilinear sampling, Sobel/desaturate, ray/triangle intersection, 3x3 LU, CG, sparse LU --
 * and the MRF solver (mapMAP is absent; the forest block-coordinate-descent solver is this repo's own).
 * Those pieces are restated from the libraries' published behaviour and marked [UPSTREAM-RECALL].
 * Every function cites the reference file:line it follows.
 *
 * All arithmetic that decides a result is fp32 with FMA contraction disabled (-ffp-contract=off),
 * mirroring the operation order of the reference source.
 */
#ifndef B2TEX_ORACLE_H
#define B2TEX_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    float pos[3];
    float viewdir[3];
    float proj[9];   /* row major 3x3 */
    float w2c[16];   /* row major 4x4 */
    int32_t width, height;
    const uint8_t *rgb; /* H x W x 3 interleaved */
} orc_view;

typedef struct {
    int32_t data_term;        /* 0 AREA, 1 GMI (settings.h:59-62) */
    int32_t outlier_removal;  /* 0 NONE, 1 GAUSS_DAMPING, 2 GAUSS_CLAMPING (settings.h:70-74) */
    int32_t geometric_visibility_test;
    uint32_t face_begin, face_end; /* bounded sample for CPU timing: only faces in [begin,end) are
                                      evaluated (occlusion still against the whole mesh); 0,0 = all */
} orc_settings;

/* ---- image preparation (texture_view.cpp:42-132) ---- */
void orc_validity_mask(const uint8_t *rgb, int w, int h, uint8_t *mask);
void orc_erode_validity_mask(uint8_t *mask, int w, int h);
void orc_gradient_magnitude(const uint8_t *rgb, int w, int h, uint8_t *grad);

/* ---- BVH (rayint acc::BVHTree stand-in) ---- */
typedef struct orc_bvh orc_bvh;
orc_bvh *orc_bvh_build(const float *verts, const uint32_t *faces, uint32_t num_faces);
void orc_bvh_free(orc_bvh *b);
/* any hit with tmin <= t <= tmax; dir must be normalised by the caller */
int orc_bvh_occluded(const orc_bvh *b, const float o[3], const float d[3], float tmin, float tmax);
int orc_brute_occluded(const float *verts, const uint32_t *faces, uint32_t num_faces,
                       const float o[3], const float d[3], float tmin, float tmax);

/* ---- data costs (calculate_data_costs.cpp:131-323) ---- */
/* Two-call protocol: pass NULL out arrays to get nnz, then call again with buffers.
 * Results are cached between the two calls (same thread). */
typedef struct {
    uint64_t nnz;
    float max_quality;
    float percentile;
} orc_dc_info;

int orc_data_costs(const float *verts, uint32_t num_verts, const uint32_t *faces,
                   const float *face_normals, uint32_t num_faces, const orc_view *views,
                   uint32_t num_views, const orc_settings *settings, int num_threads,
                   uint64_t *face_ptr /* F+1 */, uint16_t **view_out, float **cost_out,
                   float **quality_out, orc_dc_info *info);
void orc_free(void *p);

/* pieces exposed for known-answer tests */
void orc_pixel_coords(const orc_view *v, const float x[3], float out[2]);
float orc_tri_area(const float p1[2], const float p2[2], const float p3[2]);
int orc_tri_inside(const float p1[2], const float p2[2], const float p3[2], float x, float y);
float orc_face_quality(const orc_view *v, const uint8_t *grad, const float v1[3],
                       const float v2[3], const float v3[3], int data_term);
float orc_histogram_percentile(const float *values, uint64_t n, float vmax, int bins, float p);

/* ---- MRF (view_selection.cpp:18-133 model; forest block-coordinate-descent solver) ---- */
typedef struct {
    uint32_t max_iterations;   /* hard cap */
    uint32_t rounds;           /* forest growth rounds D */
    uint32_t root_div;         /* one root candidate per root_div nodes */
    uint32_t seed;             /* view_selection.cpp:115 initial_seed */
    uint32_t window;           /* StopWhenReturnsDiminish(window, ratio) view_selection.cpp:84 */
    float ratio;
    uint32_t num_parts;        /* logical face partitions (multi-GPU emulation), >=1 */
} orc_mrf_params;

typedef struct {
    uint32_t iterations;
    double energy_initial;
    double energy_final;
    uint64_t unseen;
} orc_mrf_info;

int orc_view_selection(uint32_t num_faces, const uint32_t *adj_ptr, const uint32_t *adj_idx,
                       const uint64_t *face_ptr, const uint16_t *view, const float *cost,
                       const orc_mrf_params *params, int num_threads, uint32_t *labels_out,
                       double *energy_trace /* max_iterations+1 or NULL */, orc_mrf_info *info);
double orc_mrf_energy(uint32_t num_faces, const uint32_t *adj_ptr, const uint32_t *adj_idx,
                      const uint64_t *face_ptr, const uint16_t *view, const float *cost,
                      const uint32_t *labels);
int64_t orc_mrf_energy_fixed(uint32_t num_faces, const uint32_t *adj_ptr, const uint32_t *adj_idx,
                             const uint64_t *face_ptr, const uint16_t *view, const float *cost,
                             const uint32_t *labels);
/* exhaustive minimum for tiny problems (<= ~16 nodes) */
double orc_mrf_brute_force(uint32_t num_faces, const uint32_t *adj_ptr, const uint32_t *adj_idx,
                           const uint64_t *face_ptr, const uint16_t *view, const float *cost,
                           uint32_t *labels_out);
/* forest sampling of one iteration, exposed for structural tests: level[v] = join round,
 * 0xFFFFFFFF untouched, 0xFFFFFFFE excluded */
void orc_mrf_sample_forest(uint32_t num_faces, const uint32_t *adj_ptr, const uint32_t *adj_idx,
                           const uint64_t *face_ptr, const orc_mrf_params *params,
                           uint32_t iteration, uint32_t *level_out);

/* ---- global seam leveling (global_seam_leveling.cpp:140-291) ---- */
typedef struct {
    uint32_t num_rows;        /* x_rows */
    uint32_t num_a_rows;
    uint32_t num_gamma_rows;
    uint64_t nnz_full;        /* nnz of the full symmetric Lhs */
    uint32_t iterations[3];
    float residual[3];        /* Eigen cg.error(): sqrt(|r|^2/|rhs|^2) */
} orc_seam_info;

/* Output: row_ptr[Vn+1] (rows of vertex v are row_ptr[v]..), row_label[R], x[R*3] (centred).
 * Two-call protocol like orc_data_costs (row_label/x NULL -> only info + row_ptr). */
int orc_global_seam_leveling(const float *verts, uint32_t num_verts, const uint32_t *faces,
                             uint32_t num_faces, const uint32_t *vf_ptr, const uint32_t *vf_idx,
                             const uint32_t *vv_ptr, const uint32_t *vv_idx,
                             const uint32_t *labels, const orc_view *views, uint32_t num_views,
                             int num_threads, uint32_t *row_ptr, uint32_t **row_label_out,
                             float **x_out, float **rhs_out, orc_seam_info *info);
/* full symmetric CSR of Lhs for tests (call after orc_global_seam_leveling on the same thread) */
int orc_seam_last_matrix(uint32_t **csr_ptr, uint32_t **csr_col, float **csr_val);

#ifdef __cplusplus
}
#endif
#endif
