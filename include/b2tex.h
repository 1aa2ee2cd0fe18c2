/*
 * b2tex.h -- C ABI of the B200-native mvs-texturing hot path (libb2tex.so).
 *
 * Drop-in boundary: the reference exposes this path as four C++ free functions in
 * libs/tex/texturing.h; each entry point below replaces one of them and is what a maintainer's
 * binding (mvs-texturing_b200/tex/ for the C++ veneer that keeps the tex::
 * signatures) calls:
 *
 *   tex::calculate_data_costs   libs/tex/texturing.h:66-69  -> b2tex_calculate_data_costs
 *   tex::view_selection         libs/tex/texturing.h:79-80  -> b2tex_view_selection
 *   tex::global_seam_leveling   libs/tex/texturing.h:97-101 -> b2tex_global_seam_leveling
 *   (build_adjacency_graph      libs/tex/texturing.h:59-61  input of view_selection, passed as CSR)
 *
 * Plain pointers and sizes only; host buffers are owned by the caller, device memory by the
 * library.  Every function returns 0 on success and a non-zero status otherwise;
 * b2tex_last_error() returns a thread-local message (the C++ veneer turns it into the
 * std::runtime_error the reference throws, calculate_data_costs.cpp:315-318,
 * view_selection.cpp:126-128).  There is no CPU fallback: without a CUDA device every entry
 * point fails with B2TEX_ERR_CUDA.
 */
#ifndef B2TEX_H
#define B2TEX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2TEX_OK 0
#define B2TEX_ERR_CUDA 1      /* CUDA runtime error / no device */
#define B2TEX_ERR_LIMITS 2    /* "Exeeded maximal number of faces/views" (calculate_data_costs.cpp:315-318) */
#define B2TEX_ERR_ARG 3       /* bad argument / stage prerequisites missing */
#define B2TEX_ERR_LABELING 4  /* "Incorrect labeling" (view_selection.cpp:126-128) */
#define B2TEX_ERR_UNSUPPORTED 5

/* tex::TextureView camera + image (libs/tex/texture_view.h:39-52) as POD */
typedef struct {
    float pos[3];        /* camera centre, world */
    float viewdir[3];    /* viewing direction, world */
    float proj[9];       /* calibration, row major 3x3 (pixels) */
    float w2c[16];       /* world_to_cam, row major 4x4 */
    int32_t width, height;
    const uint8_t *rgb;  /* H x W x 3 interleaved, host memory */
} b2tex_view;

/* the fields of tex::Settings the path reads (libs/tex/settings.h:82-94) */
typedef struct {
    int32_t data_term;                 /* 0 DATA_TERM_AREA, 1 DATA_TERM_GMI */
    int32_t outlier_removal;           /* 0 NONE, 1 GAUSS_DAMPING, 2 GAUSS_CLAMPING (settings.h:70-74) */
    int32_t geometric_visibility_test; /* bool */
} b2tex_settings;

typedef struct {
    uint64_t nnz;          /* DataCosts entries */
    uint64_t candidates;   /* (face,view) pairs that passed culling + projection */
    uint64_t rays;         /* distinct (vertex,view) visibility rays traced */
    float max_quality;     /* calculate_data_costs.cpp:304 */
    float percentile;      /* calculate_data_costs.cpp:305 */
} b2tex_dc_info;

/* mapMAP control as configured at view_selection.cpp:84,103-115, mapped onto the forest-BCD solver */
typedef struct {
    uint32_t max_iterations;
    uint32_t rounds;       /* forest growth rounds per iteration */
    uint32_t root_div;     /* one root candidate per root_div nodes; 0 = single root */
    uint32_t seed;         /* initial_seed */
    uint32_t window;       /* StopWhenReturnsDiminish(window, ratio) */
    float ratio;
    uint32_t num_parts;    /* logical face partitions (>=1); >1 reproduces the multi-GPU schedule */
    uint32_t num_views;    /* DataCosts::rows() for the one-shot call; 0 = derive from the entries */
} b2tex_mrf_params;

typedef struct {
    uint32_t iterations;
    double energy_initial;
    double energy_final;
    uint64_t unseen;       /* "faces have not been seen" view_selection.cpp:132 */
    uint64_t sweep_bytes;  /* algorithmic bytes of one sweep: 14 nnz + 20 F */
} b2tex_mrf_info;

typedef struct {
    uint32_t num_rows;       /* Lhs dimensionality, global_seam_leveling.cpp:253 */
    uint32_t num_a_rows;
    uint32_t num_gamma_rows;
    uint64_t nnz_full;       /* non-zeros of the full symmetric Lhs */
    uint32_t iterations[3];  /* cg.iterations() per colour channel, :280 */
    float residual[3];       /* cg.error() per colour channel, :281 */
    uint32_t cg_launch_iterations; /* iterations executed by the batched kernel = max over channels */
    float cg_ms;             /* device time of the PCG kernel (CUDA events) */
} b2tex_seam_info;

typedef struct {
    uint32_t num_patches;   /* "texture patches." generate_texture_patches.cpp:597 (seen faces only) */
    uint32_t num_faces;     /* faces with a label != 0 = sum of the patches' face counts */
    uint64_t num_pixels;    /* sum of width * height over the patches */
} b2tex_patch_info;

typedef struct {
    uint32_t num_seam_edges;    /* find_seam_edges, seam_leveling.cpp:16-59 */
    uint32_t num_edge_samples;  /* sum of ceil(2 * max projected length), local_seam_leveling.cpp:139 */
    uint32_t num_vertices;      /* vertices projected into more than one patch, :157 */
    uint32_t num_unknowns;      /* blending mask == 255 pixels of all patches (interior of the 20 px strips) */
    uint32_t iterations[3];     /* CG iterations per colour channel (one batched solve for all patches) */
    float residual[3];          /* |r| / |b| per colour channel at exit */
} b2tex_local_seam_info;

typedef struct b2tex_ctx b2tex_ctx;

/* ---- lifetime ---- */
int b2tex_create(int device, b2tex_ctx **out);
void b2tex_destroy(b2tex_ctx *ctx);
const char *b2tex_last_error(void);
void b2tex_free(void *host_ptr);                 /* frees buffers returned by one-shot calls */
int b2tex_device_synchronize(b2tex_ctx *ctx);
uint64_t b2tex_stream(b2tex_ctx *ctx);            /* the cudaStream_t every kernel is launched on */
uint64_t b2tex_launch_count(void);                /* kernels of this library launched by this process so far (CUB's not counted) */
/* per-kernel CUDA-event timing: enable, run stages, read "name ms algorithmic_bytes" lines */
int b2tex_profile(b2tex_ctx *ctx, int enable);
int b2tex_profile_report(b2tex_ctx *ctx, char *buf, uint64_t cap);
void b2tex_default_mrf_params(b2tex_mrf_params *p);

/* ---- resident API: upload once, run stages on the device, download results ---- */
int b2tex_set_mesh(b2tex_ctx *ctx, const float *verts, uint32_t num_verts, const uint32_t *faces,
                   const float *face_normals, uint32_t num_faces);
int b2tex_set_views(b2tex_ctx *ctx, const b2tex_view *views, uint32_t num_views);
int b2tex_set_adjacency(b2tex_ctx *ctx, const uint32_t *adj_ptr, const uint32_t *adj_idx);
int b2tex_set_vertex_rings(b2tex_ctx *ctx, const uint32_t *vf_ptr, const uint32_t *vf_idx,
                           const uint32_t *vv_ptr, const uint32_t *vv_idx);
int b2tex_set_data_costs(b2tex_ctx *ctx, const uint64_t *face_ptr, const uint16_t *view,
                         const float *cost);
int b2tex_set_labels(b2tex_ctx *ctx, const uint32_t *labels);
/* restrict this context to faces [face_begin, face_end) for the data-cost stage (multi-GPU shard);
 * default is all faces */
int b2tex_set_face_range(b2tex_ctx *ctx, uint32_t face_begin, uint32_t face_end);

int b2tex_data_costs_run(b2tex_ctx *ctx, const b2tex_settings *settings, b2tex_dc_info *info);
/* split form of the normalisation for sharded runs (calculate_data_costs.cpp:277-302):
 * qualities -> [allreduce max] -> histogram -> [allreduce sum] -> normalise */
int b2tex_data_costs_qualities(b2tex_ctx *ctx, const b2tex_settings *settings, b2tex_dc_info *info);
int b2tex_data_costs_histogram(b2tex_ctx *ctx, float global_max, uint32_t *bins10000_device_or_host,
                               int to_host);
int b2tex_data_costs_normalize(b2tex_ctx *ctx, float global_max, const uint32_t *bins10000_host,
                               b2tex_dc_info *info);
int b2tex_data_costs_download(b2tex_ctx *ctx, uint64_t *face_ptr, uint16_t *view, float *cost,
                              float *quality_or_null);

int b2tex_view_selection_run(b2tex_ctx *ctx, const b2tex_mrf_params *params, b2tex_mrf_info *info,
                             double *energy_trace_or_null);
/* optional: performs every device allocation b2tex_view_selection_run(params) will need and nothing else.  Only a caller
 * that drives several peer ranks from ONE process needs it (all ranks prepare before the first one runs: cudaMalloc waits
 * for the whole device, which a rank already spinning in a cross-rank barrier kernel would block for good). */
int b2tex_view_selection_prepare(b2tex_ctx *ctx, const b2tex_mrf_params *params);
int b2tex_labels_download(b2tex_ctx *ctx, uint32_t *labels);
/* Multi-GPU view selection (one process per GPU, at most 8): rank r owns the faces [r * ceil(F / P), (r + 1) * ceil(F / P))
 * (b2tex_set_face_range) and runs b2tex_view_selection_run with params->num_parts = P like a single GPU would.  Before
 * that, once per mesh size: every rank calls b2tex_mrf_mg_export (allocates a peer-visible block holding its full-length
 * label array, the energy slots and the barrier flags; the context's labels live inside it from then on), the 64-byte
 * cudaIpc handles are exchanged by the caller (e.g. one all-gather over torch.distributed) and imported with
 * b2tex_mrf_mg_import.  Inside the run the ranks exchange only the labels of their boundary faces -- stored straight
 * into the label arrays of the ranks that own a neighbouring face -- and their partial energies, through NVLink peer
 * memory with epoch-flag barriers; every rank takes the identical stop decision (view_selection.cpp:84) from the
 * identical fixed-point sum.  No NCCL call and no host round trip per iteration. */
int b2tex_mrf_mg_export(b2tex_ctx *ctx, uint32_t rank, uint32_t num_ranks, void *ipc_handle_64_bytes);
int b2tex_mrf_mg_import(b2tex_ctx *ctx, uint32_t peer_rank, const void *ipc_handle_64_bytes);
/* Peers that live in the SAME process (several contexts driven by threads) cannot open each other's IPC handles: after
 * the export they attach the raw device pointer of the peer's block instead.  which: 0 = view selection block
 * (b2tex_mrf_mg_export), 1 = seam solve block (b2tex_seam_mg_export). */
uint64_t b2tex_peer_block(b2tex_ctx *ctx, int which);
int b2tex_peer_attach(b2tex_ctx *ctx, int which, uint32_t peer_rank, uint64_t peer_block_device_ptr);
/* building blocks of one solver iteration, exposed so that a sharded run can exchange boundary
 * labels between iterations; view_selection_run = init + loop(iterate, energy) */
int b2tex_mrf_init(b2tex_ctx *ctx, const b2tex_mrf_params *params, int64_t *energy_fixed);
int b2tex_mrf_iterate(b2tex_ctx *ctx, uint32_t iteration, int64_t *energy_fixed);
/* energy of the owned faces under the labels currently on the device (call after a label exchange) */
int b2tex_mrf_energy(b2tex_ctx *ctx, int64_t *energy_fixed);
int b2tex_mrf_sample_forest(b2tex_ctx *ctx, const b2tex_mrf_params *params, uint32_t iteration,
                            uint32_t *level_out_host);

int b2tex_seam_run(b2tex_ctx *ctx, b2tex_seam_info *info);
/* Multi-GPU solve of the same system (one process per GPU): every rank assembles (b2tex_seam_assemble = b2tex_seam_run
 * without the PCG), allocates a peer-visible exchange block and exports its 64-byte cudaIpc handle, imports the handles
 * of all peers (exchanged by the caller, e.g. an all-gather over torch.distributed), then every rank calls
 * b2tex_seam_mg_solve: one persistent cooperative kernel per GPU that runs the PCG on its slice of the rows and exchanges
 * the search direction and the dot products with its peers through NVLink peer memory inside the kernel.  Every rank ends
 * with the complete solution (b2tex_seam_download).  At most 8 ranks. */
int b2tex_seam_assemble(b2tex_ctx *ctx, b2tex_seam_info *info);
int b2tex_seam_mg_export(b2tex_ctx *ctx, uint32_t rank, uint32_t num_ranks, void *ipc_handle_64_bytes);
int b2tex_seam_mg_import(b2tex_ctx *ctx, uint32_t peer_rank, const void *ipc_handle_64_bytes);
int b2tex_seam_mg_solve(b2tex_ctx *ctx, b2tex_seam_info *info);
int b2tex_seam_download(b2tex_ctx *ctx, uint32_t *row_ptr, uint32_t *row_label, float *x,
                        float *rhs_or_null);
/* full symmetric CSR of Lhs (for tests / inspection); arrays sized from b2tex_seam_info */
int b2tex_seam_matrix_download(b2tex_ctx *ctx, uint32_t *csr_ptr, uint32_t *csr_col, float *csr_val);
/* tex::generate_texture_patches for the seen faces (generate_texture_patches.cpp:78-138,453-538; hole filling
 * :140-451 is not built: faces with label 0 get no patch) followed by TexturePatch::adjust_colors per patch
 * (texture_patch.cpp:41-116) with the solved offsets (apply_adjust = 1, global_seam_leveling.cpp:293-323; needs
 * b2tex_seam_run) or with zeros (apply_adjust = 0, texrecon.cpp:174-183).  Needs mesh, views, adjacency, labels.
 * Patch ids follow ascending labels (the reference's ids depend on OpenMP scheduling, :469,514-518). */
int b2tex_texture_patches_run(b2tex_ctx *ctx, int apply_adjust, b2tex_patch_info *info);
/* desc[num_patches][8] = label, min_x, min_y (view pixel of patch pixel (0,0)), width, height, first face slot,
 * number of faces, 0; faces[num_faces]; texcoords[num_faces][3][2]; per pixel (patch after patch, row major):
 * images[num_pixels][3] float, validity[num_pixels], blending[num_pixels] (255 inside, 64 within sqrt(2), 0).
 * Any pointer may be NULL. */
int b2tex_texture_patches_download(b2tex_ctx *ctx, int32_t *desc, uint32_t *faces, float *texcoords, float *images,
                                   uint8_t *validity, uint8_t *blending);
/* tex::local_seam_leveling (local_seam_leveling.cpp:105-204) on the patches b2tex_texture_patches_run left on the device:
 * mean colours along the seam edges and at shared vertices are stamped into every adjacent patch, the blending mask keeps
 * a 20 pixel strip (TexturePatch::prepare_blending_mask), the strip is Poisson-blended towards the stamped colours
 * (poisson_blend, alpha = 1; one batched CG over all patches instead of one SparseLU per patch: same linear systems,
 * agreement to the CG tolerance) and pixels outside the patch boundary are invalidated (TexturePatch::blend).
 * Download the result with b2tex_texture_patches_download (the blending mask is the one used for blending; the
 * reference releases it at :201). */
int b2tex_local_seam_leveling_run(b2tex_ctx *ctx, b2tex_local_seam_info *info);
/* raw device pointers of resident results (torch / NCCL plumbing); 0 if absent */
uint64_t b2tex_device_ptr(b2tex_ctx *ctx, const char *name, uint64_t *num_elements);

/* ---- one-shot host-buffer entry points (what the reference-side binding calls) ----
 * They run on the calling thread's current CUDA device (cudaGetDevice) and keep up to two finished contexts (stream +
 * device buffers) cached for the next call; b2tex_release_cached_contexts() frees them. */
void b2tex_release_cached_contexts(void);
/* tex::calculate_data_costs: out arrays are malloc'ed by the library (b2tex_free), CSR by face:
 * face_ptr[F+1], view[nnz] ascending per face, cost[nnz]. */
int b2tex_calculate_data_costs(const float *verts, uint32_t num_verts, const uint32_t *faces,
                               const float *face_normals, uint32_t num_faces,
                               const b2tex_view *views, uint32_t num_views,
                               const b2tex_settings *settings, uint64_t **face_ptr_out,
                               uint16_t **view_out, float **cost_out, b2tex_dc_info *info);
/* same, into caller-owned (ideally pinned) buffers of `capacity` entries: no allocation, no extra copy;
 * fails with B2TEX_ERR_ARG (info->nnz = required size) when the capacity is too small */
int b2tex_calculate_data_costs_into(const float *verts, uint32_t num_verts, const uint32_t *faces,
                                    const float *face_normals, uint32_t num_faces,
                                    const b2tex_view *views, uint32_t num_views,
                                    const b2tex_settings *settings, uint64_t *face_ptr, uint16_t *view,
                                    float *cost, uint64_t capacity, b2tex_dc_info *info);
/* tex::postprocess_face_infos (libs/tex/texturing.h:71-74, calculate_data_costs.cpp:253-306) for qualities the caller
 * computed: per face (CSR face_ptr[F+1]) the infos in ascending view order -- view[n], quality[n] and, with outlier
 * removal, the mean YCbCr colour mean_color_ycbcr[n][3] (NULL otherwise).  Photometric outlier removal, quality == 0
 * entries dropped, 99.5 % percentile of the 10 000-bin histogram, cost = 1 - min(1, quality / percentile).  Out arrays are
 * caller allocated: face_ptr_out[F+1], view_out / cost_out with room for n entries. */
int b2tex_postprocess_face_infos(uint32_t num_faces, const uint64_t *face_ptr, const uint16_t *view, const float *quality,
                                 const float *mean_color_ycbcr, const b2tex_settings *settings, uint64_t *face_ptr_out,
                                 uint16_t *view_out, float *cost_out, b2tex_dc_info *info);
/* tex::view_selection: labels_out[F] (0 = unseen, else view index + 1) */
int b2tex_view_selection(uint32_t num_faces, const uint32_t *adj_ptr, const uint32_t *adj_idx,
                         const uint64_t *face_ptr, const uint16_t *view, const float *cost,
                         const b2tex_mrf_params *params_or_null, uint32_t *labels_out,
                         b2tex_mrf_info *info);
/* tex::global_seam_leveling up to the per-(vertex,label) adjust values (:283-289):
 * row_ptr_out[Vn+1] caller allocated; row_label/x (R and R*3, centred) malloc'ed by the library. */
int b2tex_global_seam_leveling(const float *verts, uint32_t num_verts, const uint32_t *faces,
                               uint32_t num_faces, const uint32_t *vf_ptr, const uint32_t *vf_idx,
                               const uint32_t *vv_ptr, const uint32_t *vv_idx,
                               const uint32_t *labels, const b2tex_view *views, uint32_t num_views,
                               uint32_t *row_ptr_out, uint32_t **row_label_out, float **x_out,
                               b2tex_seam_info *info);

/* The three stages back to back on one upload -- what texrecon does between texrecon.cpp:100 and :171
 * when it writes no intermediate results (--no_intermediate_results, arguments.cpp:88-89): the mesh and
 * the images cross PCIe once, DataCosts stay on the device, only labels[F] and the per-(vertex,label)
 * adjust values come back.  Optional outputs may be NULL.  row_label/x are malloc'ed (b2tex_free). */
int b2tex_texture_hot_path(const float *verts, uint32_t num_verts, const uint32_t *faces,
                           const float *face_normals, uint32_t num_faces, const b2tex_view *views,
                           uint32_t num_views, const uint32_t *adj_ptr, const uint32_t *adj_idx,
                           const uint32_t *vf_ptr, const uint32_t *vf_idx, const uint32_t *vv_ptr,
                           const uint32_t *vv_idx, const b2tex_settings *settings,
                           const b2tex_mrf_params *mrf_params_or_null, uint32_t *labels_out,
                           uint32_t *row_ptr_out, uint32_t **row_label_out, float **x_out,
                           b2tex_dc_info *dc_info, b2tex_mrf_info *mrf_info, b2tex_seam_info *seam_info);

/* Everything texrecon does between texrecon.cpp:160 and :189 on one upload: texture patches for the seen faces,
 * global seam leveling (do_global; else the zero-offset validity pass of :174-183), local seam leveling (do_local).
 * Outputs are malloc'ed (b2tex_free) and laid out as b2tex_texture_patches_download describes; info structs may be NULL. */
int b2tex_seam_leveling_patches(const float *verts, uint32_t num_verts, const uint32_t *faces, uint32_t num_faces,
                                const uint32_t *adj_ptr, const uint32_t *adj_idx, const uint32_t *vf_ptr,
                                const uint32_t *vf_idx, const uint32_t *vv_ptr, const uint32_t *vv_idx,
                                const uint32_t *labels, const b2tex_view *views, uint32_t num_views, int do_global,
                                int do_local, int32_t **desc_out, uint32_t **faces_out, float **texcoords_out,
                                float **images_out, uint8_t **validity_out, b2tex_patch_info *patch_info,
                                b2tex_seam_info *seam_info, b2tex_local_seam_info *local_info);

#ifdef __cplusplus
}
#endif
#endif /* B2TEX_H */
