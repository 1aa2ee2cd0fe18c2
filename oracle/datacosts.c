/* oracle/datacosts.c -- TEST INFRASTRUCTURE (see oracle.h).
 * Restates libs/tex/calculate_data_costs.cpp:131-323, texture_view.h:153-166,
 * texture_view.cpp:134-281, tri.h:50-84, tri.cpp:12-24, histogram.cpp:22-63.
 * Parallel structure = the reference's: OpenMP over views, schedule(dynamic). */
#include "oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void orc_free(void *p) { free(p); }

/* texture_view.h:161-166 with MVE's Matrix::mult operation order [UPSTREAM-RECALL]:
 * inner_product from T(0), then "+ w * m[i][3]". */
void orc_pixel_coords(const orc_view *v, const float x[3], float out[2])
{
    float cam[3], pix[3];
    for (int i = 0; i < 3; ++i) {
        const float *m = v->w2c + 4 * i;
        cam[i] = (((0.0f + m[0] * x[0]) + m[1] * x[1]) + m[2] * x[2]) + 1.0f * m[3];
    }
    for (int i = 0; i < 3; ++i) {
        const float *p = v->proj + 3 * i;
        pix[i] = ((0.0f + p[0] * cam[0]) + p[1] * cam[1]) + p[2] * cam[2];
    }
    out[0] = pix[0] / pix[2] - 0.5f;
    out[1] = pix[1] / pix[2] - 0.5f;
}

/* texture_view.cpp:253-281 */
static int valid_pixel(const orc_view *v, const uint8_t *mask, float x, float y)
{
    int w = v->width, h = v->height;
    int valid = (x >= 0.0f && x < (float)(w - 1) && y >= 0.0f && y < (float)(h - 1));
    if (valid && mask) {
        float cx = fmaxf(0.0f, fminf((float)(w - 1), x));
        float cy = fmaxf(0.0f, fminf((float)(h - 1), y));
        int fx = (int)cx, fy = (int)cy;
        int fx1 = fx + 1 < w - 1 ? fx + 1 : w - 1;
        int fy1 = fy + 1 < h - 1 ? fy + 1 : h - 1;
        valid = mask[fx + (size_t)fy * w] && mask[fx + (size_t)fy1 * w]
            && mask[fx1 + (size_t)fy * w] && mask[fx1 + (size_t)fy1 * w];
    }
    return valid;
}

/* tri.h:78-84 */
float orc_tri_area(const float p1[2], const float p2[2], const float p3[2])
{
    float u0 = p2[0] - p1[0], u1 = p2[1] - p1[1];
    float v0 = p3[0] - p1[0], v1 = p3[1] - p1[1];
    return 0.5f * fabsf(u0 * v1 - u1 * v0);
}

typedef struct { float v1[2], v2[2], v3[2], detT, min_x, min_y, max_x, max_y; } tri2;

/* tri.cpp:12-24 */
static void tri_init(tri2 *t, const float p1[2], const float p2[2], const float p3[2])
{
    memcpy(t->v1, p1, 8); memcpy(t->v2, p2, 8); memcpy(t->v3, p3, 8);
    float T0 = p1[0] - p3[0], T1 = p2[0] - p3[0];
    float T2 = p1[1] - p3[1], T3 = p2[1] - p3[1];
    t->detT = T0 * T3 - T2 * T1;
    t->min_x = fminf(p1[0], fminf(p2[0], p3[0]));
    t->min_y = fminf(p1[1], fminf(p2[1], p3[1]));
    t->max_x = fmaxf(p1[0], fmaxf(p2[0], p3[0]));
    t->max_y = fmaxf(p1[1], fmaxf(p2[1], p3[1]));
}

/* tri.h:58-76 */
static int tri_inside(const tri2 *t, float x, float y)
{
    float dx = x - t->v3[0], dy = y - t->v3[1];
    float alpha = ((t->v2[1] - t->v3[1]) * dx + (t->v3[0] - t->v2[0]) * dy) / t->detT;
    if (alpha < 0.0f || alpha > 1.0f) return 0;
    float beta = ((t->v3[1] - t->v1[1]) * dx + (t->v1[0] - t->v3[0]) * dy) / t->detT;
    if (beta < 0.0f || beta > 1.0f) return 0;
    if (alpha + beta > 1.0f) return 0;
    return 1;
}

int orc_tri_inside(const float p1[2], const float p2[2], const float p3[2], float x, float y)
{
    tri2 t;
    tri_init(&t, p1, p2, p3);
    return tri_inside(&t, x, y);
}

/* mve::Image<uint8_t>::linear_at + math::interpolate<unsigned char> [UPSTREAM-RECALL] */
static uint8_t linear_at_u8(const uint8_t *img, int w, int h, float x, float y)
{
    x = fmaxf(0.0f, fminf((float)(w - 1), x));
    y = fmaxf(0.0f, fminf((float)(h - 1), y));
    int fx = (int)x, fy = (int)y;
    int fx1 = fx + 1 < w - 1 ? fx + 1 : w - 1;
    int fy1 = fy + 1 < h - 1 ? fy + 1 : h - 1;
    float w1 = x - (float)fx, w0 = 1.0f - w1;
    float w3 = y - (float)fy, w2 = 1.0f - w3;
    float r = (float)img[fx + (size_t)fy * w] * (w0 * w2) + (float)img[fx1 + (size_t)fy * w] * (w1 * w2)
        + (float)img[fx + (size_t)fy1 * w] * (w0 * w3) + (float)img[fx1 + (size_t)fy1 * w] * (w1 * w3)
        + 0.5f;
    return (uint8_t)r;
}

/* mve::Image<uint8_t>::linear_at on one channel of the interleaved rgb image [UPSTREAM-RECALL] */
static uint8_t linear_at_rgb(const uint8_t *img, int w, int h, float x, float y, int ch)
{
    x = fmaxf(0.0f, fminf((float)(w - 1), x));
    y = fmaxf(0.0f, fminf((float)(h - 1), y));
    int fx = (int)x, fy = (int)y;
    int fx1 = fx + 1 < w - 1 ? fx + 1 : w - 1;
    int fy1 = fy + 1 < h - 1 ? fy + 1 : h - 1;
    float w1 = x - (float)fx, w0 = 1.0f - w1;
    float w3 = y - (float)fy, w2 = 1.0f - w3;
    float r = (float)img[3 * (fx + (size_t)fy * w) + ch] * (w0 * w2) + (float)img[3 * (fx1 + (size_t)fy * w) + ch] * (w1 * w2)
        + (float)img[3 * (fx + (size_t)fy1 * w) + ch] * (w0 * w3) + (float)img[3 * (fx1 + (size_t)fy1 * w) + ch] * (w1 * w3)
        + 0.5f;
    return (uint8_t)r;
}

/* texture_view.cpp:134-251 on already projected points; mean_color != NULL <=> outlier removal on
 * (colours are then sampled even for DATA_TERM_AREA, :159) */
static float face_quality_px(const orc_view *v, const uint8_t *grad, float p1[2], float p2[2],
                             float p3[2], int data_term, float *mean_color)
{
    tri2 tri;
    tri_init(&tri, p1, p2, p3);
    float area = orc_tri_area(p1, p2, p3);
    if (area < FLT_EPSILON) return 0.0f;

    size_t num_samples = 0;
    double gmi = 0.0;
    double colors[3] = {0.0, 0.0, 0.0};
    int sampling_necessary = data_term != 0 || mean_color != NULL;
    int w = v->width;

    if (sampling_necessary && area > 0.5f) {
        /* texture_view.cpp:163-167 */
        for (;;) {
            if (p1[1] <= p2[1]) {
                if (p2[1] <= p3[1]) break;
                float t0 = p2[0], t1 = p2[1]; p2[0] = p3[0]; p2[1] = p3[1]; p3[0] = t0; p3[1] = t1;
            } else {
                float t0 = p1[0], t1 = p1[1]; p1[0] = p2[0]; p1[1] = p2[1]; p2[0] = t0; p2[1] = t1;
            }
        }
        float m1 = (p1[1] - p3[1]) / (p1[0] - p3[0]);
        float b1 = p1[1] - m1 * p1[0];
        float m2 = (p1[1] - p2[1]) / (p1[0] - p2[0]);
        float b2 = p1[1] - m2 * p1[0];
        float m3 = (p2[1] - p3[1]) / (p2[0] - p3[0]);
        float b3 = p2[1] - m3 * p2[0];
        int fast = isfinite(m1) && m2 != 0.0f && isfinite(m2) && m3 != 0.0f && isfinite(m3);

        int y0 = (int)floorf(tri.min_y);
        float y_end = ceilf(tri.max_y);
        for (int y = y0; (float)y < y_end; ++y) {
            float min_x = tri.min_x - 0.5f;
            float max_x = tri.max_x + 0.5f;
            if (fast) {
                float cy = (float)y + 0.5f;
                min_x = (cy - b1) / m1;
                if (cy <= p2[1]) max_x = (cy - b2) / m2;
                else max_x = (cy - b3) / m3;
                if (min_x >= max_x) { float t = min_x; min_x = max_x; max_x = t; }
                if (min_x < tri.min_x || min_x > tri.max_x) continue;
                if (max_x < tri.min_x || max_x > tri.max_x) continue;
            }
            int x0 = (int)floorf(min_x + 0.5f);
            float x_end = ceilf(max_x - 0.5f);
            for (int x = x0; (float)x < x_end; ++x) {
                float cx = (float)x + 0.5f;
                float cy = (float)y + 0.5f;
                if (!fast && !tri_inside(&tri, cx, cy)) continue;
                if (mean_color) /* :207-212 */
                    for (int i = 0; i < 3; ++i) colors[i] += (double)v->rgb[3 * (x + (size_t)y * w) + i] / 255.0;
                if (data_term == 1) gmi += (double)grad[x + (size_t)y * w] / 255.0;
                ++num_samples;
            }
        }
    }

    if (mean_color) { /* :233-245 */
        if (num_samples > 0) {
            for (int i = 0; i < 3; ++i) mean_color[i] = (float)(colors[i] / (double)num_samples);
        } else {
            for (int i = 0; i < 3; ++i) {
                double c1 = (double)linear_at_rgb(v->rgb, v->width, v->height, p1[0], p1[1], i) / 255.0;
                double c2 = (double)linear_at_rgb(v->rgb, v->width, v->height, p2[0], p2[1], i) / 255.0;
                double c3 = (double)linear_at_rgb(v->rgb, v->width, v->height, p3[0], p3[1], i) / 255.0;
                mean_color[i] = (float)((c1 + c2 + c3) / 3.0);
            }
        }
    }
    if (data_term == 1) {
        if (num_samples > 0) {
            gmi = (gmi / (double)num_samples) * (double)area;
        } else {
            double g1 = (double)linear_at_u8(grad, v->width, v->height, p1[0], p1[1]) / 255.0;
            double g2 = (double)linear_at_u8(grad, v->width, v->height, p2[0], p2[1]) / 255.0;
            double g3 = (double)linear_at_u8(grad, v->width, v->height, p3[0], p3[1]) / 255.0;
            gmi = ((g1 + g2 + g3) / 3.0) * (double)area;
        }
        return (float)gmi;
    }
    return area;
}

float orc_face_quality(const orc_view *v, const uint8_t *grad, const float v1[3],
                       const float v2[3], const float v3[3], int data_term)
{
    float p1[2], p2[2], p3[2];
    orc_pixel_coords(v, v1, p1);
    orc_pixel_coords(v, v2, p2);
    orc_pixel_coords(v, v3, p3);
    return face_quality_px(v, grad, p1, p2, p3, data_term, NULL);
}

static inline float norm3(const float a[3]) { return sqrtf(((0.0f + a[0] * a[0]) + a[1] * a[1]) + a[2] * a[2]); }
static inline float dot3(const float a[3], const float b[3]) { return ((0.0f + a[0] * b[0]) + a[1] * b[1]) + a[2] * b[2]; }

/* histogram.cpp:22-63 on a flat value array */
float orc_histogram_percentile(const float *values, uint64_t n, float vmax, int nbins, float p)
{
    float vmin = 0.0f;
    unsigned int *bins = (unsigned int *)calloc((size_t)nbins, sizeof(unsigned int));
    int num_values = 0;
    for (uint64_t i = 0; i < n; ++i) {
        float c = fmaxf(vmin, fminf(vmax, values[i]));
        size_t index = (size_t)floorf(((c - vmin) / (vmax - vmin)) * (float)(nbins - 1));
        bins[index]++;
        ++num_values;
    }
    int num = 0;
    float upper = vmin;
    float result = vmax;
    int found = 0;
    for (int i = 0; i < nbins; ++i) {
        if ((float)num / (float)num_values > p) { result = upper; found = 1; break; }
        num += (int)bins[i];
        upper = ((float)i / (float)(nbins - 1)) * (vmax - vmin) + vmin;
    }
    (void)found;
    free(bins);
    return result;
}

/* ---- photometric outlier detection, calculate_data_costs.cpp:35-129 ------------------------------
 * Eigen pieces restated [UPSTREAM-RECALL]: colwise().mean(), (C^T C)/(n-1) with sequential sums,
 * FullPivLU<Matrix3d> (full pivoting, rank threshold eps*3*|max pivot|, inverse by P/L/U/Q solves),
 * multi_gauss_unnormalized (util.h:60-66) = exp((-0.5*d) * Cinv * d^T), evaluated left to right.
 * Per-face infos are processed in ascending view order (the reference's order depends on the OpenMP
 * merge, calculate_data_costs.cpp:241-249, i.e. is not deterministic there). */
static int lu3_inverse(const double A[9], double inv[9])
{
    double lu[9];
    memcpy(lu, A, sizeof(lu));
    int rt[3], ct[3];
    double maxpivot = 0.0;
    int nonzero = 3;
    for (int k = 0; k < 3; ++k) {
        int br = k, bc = k;
        double biggest = -1.0;
        for (int cc = k; cc < 3; ++cc)      /* column-major visiting order, first strict maximum */
            for (int rr = k; rr < 3; ++rr) {
                double a = fabs(lu[rr * 3 + cc]);
                if (a > biggest) { biggest = a; br = rr; bc = cc; }
            }
        if (biggest == 0.0) { nonzero = k; for (int i = k; i < 3; ++i) { rt[i] = i; ct[i] = i; } break; }
        if (biggest > maxpivot) maxpivot = biggest;
        rt[k] = br; ct[k] = bc;
        if (br != k) for (int cc = 0; cc < 3; ++cc) { double t = lu[k * 3 + cc]; lu[k * 3 + cc] = lu[br * 3 + cc]; lu[br * 3 + cc] = t; }
        if (bc != k) for (int rr = 0; rr < 3; ++rr) { double t = lu[rr * 3 + k]; lu[rr * 3 + k] = lu[rr * 3 + bc]; lu[rr * 3 + bc] = t; }
        for (int rr = k + 1; rr < 3; ++rr) lu[rr * 3 + k] = lu[rr * 3 + k] / lu[k * 3 + k];
        for (int rr = k + 1; rr < 3; ++rr)
            for (int cc = k + 1; cc < 3; ++cc) lu[rr * 3 + cc] = lu[rr * 3 + cc] - lu[rr * 3 + k] * lu[k * 3 + cc];
    }
    int rank = 0;
    double thr = maxpivot * (2.220446049250313e-16 * 3.0);
    for (int i = 0; i < nonzero; ++i) if (fabs(lu[i * 3 + i]) > thr) ++rank;
    if (rank != 3) return 0;
    /* inverse = solve(Identity): c = P*I; L c = ..; U c = ..; result = Q c */
    for (int col = 0; col < 3; ++col) {
        double c[3] = {0.0, 0.0, 0.0};
        c[col] = 1.0;
        for (int k = 0; k < 3; ++k) if (rt[k] != k) { double t = c[k]; c[k] = c[rt[k]]; c[rt[k]] = t; }
        for (int i = 1; i < 3; ++i) for (int j = 0; j < i; ++j) c[i] = c[i] - lu[i * 3 + j] * c[j];
        for (int i = 2; i >= 0; --i) {
            for (int j = i + 1; j < 3; ++j) c[i] = c[i] - lu[i * 3 + j] * c[j];
            c[i] = c[i] / lu[i * 3 + i];
        }
        for (int k = 2; k >= 0; --k) if (ct[k] != k) { double t = c[k]; c[k] = c[ct[k]]; c[ct[k]] = t; }
        for (int i = 0; i < 3; ++i) inv[i * 3 + col] = c[i];
    }
    return 1;
}

static double gauss3(const float x[3], const double mu[3], const double ci[9])
{
    double d[3], t[3], r[3];
    for (int i = 0; i < 3; ++i) { d[i] = (double)x[i] - mu[i]; t[i] = -0.5 * d[i]; }
    for (int j = 0; j < 3; ++j) r[j] = (t[0] * ci[0 * 3 + j] + t[1] * ci[1 * 3 + j]) + t[2] * ci[2 * 3 + j];
    return exp((r[0] * d[0] + r[1] * d[1]) + r[2] * d[2]);
}

/* returns like the reference's bool; q and ycc hold the n infos of one face */
static int photometric_outlier_detection(float *q, const float *ycc, uint32_t n, int mode, uint8_t *is_inlier)
{
    if (n == 0) return 1;
    const double gauss_rejection_threshold = 6e-3, minimal_covariance = 5e-4;
    const int iterations = 10, minimal_num_inliers = 4;
    double factor = mode == 2 ? 1.0 : 0.2; /* (float)0.2f in the reference: outlier_removal_factor is float */
    if (mode == 1) factor = (double)0.2f;
    for (uint32_t r = 0; r < n; ++r) is_inlier[r] = 1;
    uint32_t rows = n;
    double mean[3], cov[9], cinv[9];
    for (int it = 0; it < iterations; ++it) {
        if (rows < (uint32_t)minimal_num_inliers) return 0;
        for (int i = 0; i < 3; ++i) mean[i] = 0.0;
        for (uint32_t r = 0; r < n; ++r) if (is_inlier[r]) for (int i = 0; i < 3; ++i) mean[i] += (double)ycc[3 * r + i];
        for (int i = 0; i < 3; ++i) mean[i] = mean[i] / (double)rows;
        for (int i = 0; i < 9; ++i) cov[i] = 0.0;
        for (uint32_t r = 0; r < n; ++r) if (is_inlier[r]) {
            double c[3];
            for (int i = 0; i < 3; ++i) c[i] = (double)ycc[3 * r + i] - mean[i];
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) cov[i * 3 + j] += c[i] * c[j];
        }
        double maxabs = 0.0;
        for (int i = 0; i < 9; ++i) { cov[i] = cov[i] / (double)(rows - 1); if (fabs(cov[i]) > maxabs) maxabs = fabs(cov[i]); }
        if (maxabs < minimal_covariance) {
            for (uint32_t r = 0; r < n; ++r) if (!is_inlier[r]) q[r] = 0.0f;
            return 1;
        }
        if (!lu3_inverse(cov, cinv)) return 0;
        rows = 0;
        for (uint32_t r = 0; r < n; ++r) {
            is_inlier[r] = gauss3(ycc + 3 * r, mean, cinv) >= gauss_rejection_threshold ? 1 : 0;
            rows += is_inlier[r];
        }
    }
    for (int i = 0; i < 9; ++i) cinv[i] = cinv[i] * factor;
    for (uint32_t r = 0; r < n; ++r) {
        double g = gauss3(ycc + 3 * r, mean, cinv);
        if (mode == 1) q[r] = (float)((double)q[r] * g);           /* info.quality *= gauss_value */
        else if (g < gauss_rejection_threshold) q[r] = 0.0f;
    }
    return 1;
}

typedef struct { uint32_t face; float q; float ycc[3]; } fq;
typedef struct { fq *data; size_t n, cap; } fqvec;

static void fq_push(fqvec *v, uint32_t f, float q, const float *ycc)
{
    if (v->n == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 1024;
        v->data = (fq *)realloc(v->data, v->cap * sizeof(fq));
    }
    v->data[v->n].face = f;
    v->data[v->n].q = q;
    for (int i = 0; i < 3; ++i) v->data[v->n].ycc[i] = ycc ? ycc[i] : 0.0f;
    v->n++;
}

int orc_data_costs(const float *verts, uint32_t num_verts, const uint32_t *faces,
                   const float *face_normals, uint32_t num_faces, const orc_view *views,
                   uint32_t num_views, const orc_settings *settings, int num_threads,
                   uint64_t *face_ptr, uint16_t **view_out, float **cost_out,
                   float **quality_out, orc_dc_info *info)
{
    (void)num_verts;
    /* calculate_data_costs.cpp:315-318 */
    if (num_views > 65535u) return 2;
    if (settings->outlier_removal < 0 || settings->outlier_removal > 2) return 3;
    const int outlier = settings->outlier_removal;

    orc_bvh *bvh = orc_bvh_build(verts, faces, num_faces); /* :144 */
    fqvec *per_view = (fqvec *)calloc(num_views ? num_views : 1, sizeof(fqvec));
    const double thr75 = 75.0f * (3.14159265358979323846264338327950288 / 180.0); /* MATH_DEG2RAD */
#ifdef _OPENMP
    if (num_threads > 0) omp_set_num_threads(num_threads);
#else
    (void)num_threads;
#endif

    #pragma omp parallel for schedule(dynamic)
    for (int jj = 0; jj < (int)num_views; ++jj) {
        const orc_view *tv = &views[jj];
        size_t npx = (size_t)tv->width * tv->height;
        uint8_t *mask = (uint8_t *)malloc(npx);
        uint8_t *grad = NULL;
        orc_validity_mask(tv->rgb, tv->width, tv->height, mask); /* :158 */
        if (settings->data_term == 1) {                             /* :160-163 */
            grad = (uint8_t *)malloc(npx);
            orc_gradient_magnitude(tv->rgb, tv->width, tv->height, grad);
            orc_erode_validity_mask(mask, tv->width, tv->height);
        }
        const float *view_pos = tv->pos;
        const float *viewing_direction = tv->viewdir;

        uint32_t f_lo = 0, f_hi = num_faces;
        if (settings->face_end > settings->face_begin) { f_lo = settings->face_begin; f_hi = settings->face_end < num_faces ? settings->face_end : num_faces; }
        for (uint32_t face_id = f_lo; face_id < f_hi; ++face_id) { /* :168 */
            const float *v1 = verts + 3 * (size_t)faces[3 * (size_t)face_id];
            const float *v2 = verts + 3 * (size_t)faces[3 * (size_t)face_id + 1];
            const float *v3 = verts + 3 * (size_t)faces[3 * (size_t)face_id + 2];
            const float *face_normal = face_normals + 3 * (size_t)face_id;
            float c[3], vtf[3], ftv[3];
            for (int k = 0; k < 3; ++k) c[k] = ((v1[k] + v2[k]) + v3[k]) / 3.0f; /* :175 */
            for (int k = 0; k < 3; ++k) { vtf[k] = c[k] - view_pos[k]; ftv[k] = view_pos[k] - c[k]; }
            float n1 = norm3(vtf), n2 = norm3(ftv);
            for (int k = 0; k < 3; ++k) { vtf[k] = vtf[k] / n1; ftv[k] = ftv[k] / n2; } /* :179-180 */

            float viewing_angle = dot3(ftv, face_normal); /* :183 */
            if (viewing_angle < 0.0f || dot3(viewing_direction, vtf) < 0.0f) continue;
            if ((double)acosf(viewing_angle) > thr75) continue; /* :187 */

            float p1[2], p2[2], p3[2];
            orc_pixel_coords(tv, v1, p1);
            orc_pixel_coords(tv, v2, p2);
            orc_pixel_coords(tv, v3, p3);
            if (!(valid_pixel(tv, mask, p1[0], p1[1]) && valid_pixel(tv, mask, p2[0], p2[1])
                  && valid_pixel(tv, mask, p3[0], p3[1])))
                continue; /* :191 */

            if (settings->geometric_visibility_test) { /* :194-215 */
                int visible = 1;
                const float *samples[3] = {v1, v2, v3};
                for (int k = 0; k < 3; ++k) {
                    float dir[3];
                    for (int a = 0; a < 3; ++a) dir[a] = view_pos[a] - samples[k][a];
                    float tmax = norm3(dir);
                    float tmin = tmax * 0.0001f;
                    float nn = norm3(dir);
                    for (int a = 0; a < 3; ++a) dir[a] = dir[a] / nn;
                    if (orc_bvh_occluded(bvh, samples[k], dir, tmin, tmax)) { visible = 0; break; }
                }
                if (!visible) continue;
            }
            float mc[3] = {0.0f, 0.0f, 0.0f}, ycc[3];
            float q = face_quality_px(tv, grad, p1, p2, p3, settings->data_term, outlier ? mc : NULL); /* :220 */
            if (q == 0.0f) continue;                                                /* :222 */
            /* mve::image::color_rgb_to_ycbcr<float> (:225) [UPSTREAM-RECALL] */
            ycc[0] = (mc[0] * 0.299f + mc[1] * 0.587f) + mc[2] * 0.114f;
            ycc[1] = ((mc[0] * -0.168736f + mc[1] * -0.331264f) + mc[2] * 0.5f) + 0.5f;
            ycc[2] = ((mc[0] * 0.5f + mc[1] * -0.418688f) + mc[2] * -0.081312f) + 0.5f;
            fq_push(&per_view[jj], face_id, q, ycc);
        }
        free(mask);
        free(grad);
    }
    orc_bvh_free(bvh);

    /* postprocess_face_infos :253-306: per-face lists sorted by view id */
    uint64_t *cnt = (uint64_t *)calloc((size_t)num_faces + 1, sizeof(uint64_t));
    uint64_t nnz = 0;
    for (uint32_t j = 0; j < num_views; ++j) {
        for (size_t i = 0; i < per_view[j].n; ++i) cnt[per_view[j].data[i].face + 1]++;
        nnz += per_view[j].n;
    }
    face_ptr[0] = 0;
    for (uint32_t f = 0; f < num_faces; ++f) face_ptr[f + 1] = face_ptr[f] + cnt[f + 1];
    uint16_t *vw = (uint16_t *)malloc(sizeof(uint16_t) * (nnz ? nnz : 1));
    float *ql = (float *)malloc(sizeof(float) * (nnz ? nnz : 1));
    float *cs = (float *)malloc(sizeof(float) * (nnz ? nnz : 1));
    float *yc = outlier ? (float *)malloc(sizeof(float) * 3 * (nnz ? nnz : 1)) : NULL;
    uint64_t *pos = (uint64_t *)malloc(sizeof(uint64_t) * ((size_t)num_faces + 1));
    memcpy(pos, face_ptr, sizeof(uint64_t) * ((size_t)num_faces + 1));
    for (uint32_t j = 0; j < num_views; ++j) {
        for (size_t i = 0; i < per_view[j].n; ++i) {
            uint64_t p = pos[per_view[j].data[i].face]++;
            vw[p] = (uint16_t)j;
            ql[p] = per_view[j].data[i].q;
            if (yc) for (int c = 0; c < 3; ++c) yc[3 * p + c] = per_view[j].data[i].ycc[c];
        }
        free(per_view[j].data);
    }
    free(per_view);
    free(pos);
    free(cnt);
    if (outlier) { /* :265-271: detection per face, then drop quality == 0 */
        uint8_t *flags = (uint8_t *)malloc(num_views ? num_views : 1);
        uint64_t o = 0;
        for (uint32_t f = 0; f < num_faces; ++f) {
            uint64_t a = face_ptr[f], b = face_ptr[f + 1];
            photometric_outlier_detection(ql + a, yc + 3 * a, (uint32_t)(b - a), outlier, flags);
            face_ptr[f] = o;
            for (uint64_t i = a; i < b; ++i)
                if (ql[i] != 0.0f) { vw[o] = vw[i]; ql[o] = ql[i]; ++o; }
        }
        face_ptr[num_faces] = o;
        nnz = o;
        free(flags);
        free(yc);
    }

    float max_quality = 0.0f; /* :278-281 */
    for (uint64_t i = 0; i < nnz; ++i) max_quality = fmaxf(max_quality, ql[i]);
    float percentile = orc_histogram_percentile(ql, nnz, max_quality, 10000, 0.995f); /* :283-288 */
    for (uint64_t i = 0; i < nnz; ++i) { /* :291-298 */
        float normalized = fminf(1.0f, ql[i] / percentile);
        cs[i] = 1.0f - normalized;
    }
    info->nnz = nnz;
    info->max_quality = max_quality;
    info->percentile = percentile;
    *view_out = vw;
    *cost_out = cs;
    if (quality_out) *quality_out = ql; else free(ql);
    return 0;
}
