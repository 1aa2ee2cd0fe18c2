/* oracle/seam.c -- TEST INFRASTRUCTURE (see oracle.h).
 * Restates libs/tex/global_seam_leveling.cpp:26-291 (system assembly + Jacobi-PCG) and the
 * pieces of seam_leveling.cpp:61-91 / texture_patch.cpp:162-169 it needs.
 *
 * Colour source ("stage-isolated"): the reference samples the float patch image,
 * which is crop(view bytes)/255 with texcoord = pixel - patch_min (generate_texture_patches.cpp:
 * 117-128).  Until generate_texture_patches is in scope the patch of label l is taken to be the
 * whole view l-1, i.e. p = get_pixel_coords(view, vertex) and the image is bytes/255.
 *
 * Eigen 3.3.2 ConjugateGradient<SparseMatrix<float>, Lower> + DiagonalPreconditioner is absent;
 * the loop below restates Eigen's conjugate_gradient() [UPSTREAM-RECALL]:
 *   r=b; if |r|^2 < tol^2|b|^2 stop; p=M^-1 r; absNew=r.p;
 *   loop{ t=Ap; a=absNew/(p.t); x+=a p; r-=a t; if |r|^2<thr break; z=M^-1 r;
 *         absOld=absNew; absNew=r.z; p=z+(absNew/absOld) p; ++i }
 *   M = diag(A), zero diagonal -> 1.  error() = sqrt(|r|^2/|b|^2).
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static __thread uint32_t *g_csr_ptr, *g_csr_col;
static __thread float *g_csr_val;
static __thread uint32_t g_rows;

/* mve FloatImage::linear_at on bytes/255 [UPSTREAM-RECALL] */
static void sample_view(const orc_view *v, float x, float y, float out[3])
{
    int w = v->width, h = v->height;
    x = fmaxf(0.0f, fminf((float)(w - 1), x));
    y = fmaxf(0.0f, fminf((float)(h - 1), y));
    int fx = (int)x, fy = (int)y;
    int fx1 = fx + 1 < w - 1 ? fx + 1 : w - 1;
    int fy1 = fy + 1 < h - 1 ? fy + 1 : h - 1;
    float w1 = x - (float)fx, w0 = 1.0f - w1;
    float w3 = y - (float)fy, w2 = 1.0f - w3;
    const uint8_t *a = v->rgb + 3 * ((size_t)fx + (size_t)fy * w);
    const uint8_t *b = v->rgb + 3 * ((size_t)fx1 + (size_t)fy * w);
    const uint8_t *c = v->rgb + 3 * ((size_t)fx + (size_t)fy1 * w);
    const uint8_t *d = v->rgb + 3 * ((size_t)fx1 + (size_t)fy1 * w);
    for (int ch = 0; ch < 3; ++ch)
        out[ch] = (((float)a[ch] / 255.0f) * (w0 * w2) + ((float)b[ch] / 255.0f) * (w1 * w2))
            + ((float)c[ch] / 255.0f) * (w0 * w3) + ((float)d[ch] / 255.0f) * (w1 * w3);
}

/* global_seam_leveling.cpp:26-43 */
static void sample_edge(const orc_view *v, const float p1[2], const float p2[2], float out[3])
{
    float p12[2] = {p2[0] - p1[0], p2[1] - p1[1]};
    float nrm = sqrtf((0.0f + p12[0] * p12[0]) + p12[1] * p12[1]);
    size_t num_samples = (size_t)(fmaxf(nrm, 1.0f) * 2.0f);
    float acc[3] = {0, 0, 0}, wsum = 0.0f;
    for (size_t s = 0; s < num_samples; ++s) {
        float fraction = (float)s / (float)(num_samples - 1);
        float sp[2] = {p1[0] + p12[0] * fraction, p1[1] + p12[1] * fraction};
        float col[3];
        sample_view(v, sp[0], sp[1], col);
        float wgt = 1.0f - fraction;
        for (int c = 0; c < 3; ++c) acc[c] += col[c] * wgt;
        wsum += wgt;
    }
    for (int c = 0; c < 3; ++c) out[c] = acc[c] / wsum;
}

extern void orc_pixel_coords(const orc_view *v, const float x[3], float out[2]);

typedef struct { uint32_t r, c; float v; } coo;
static int coo_cmp(const void *a, const void *b)
{
    const coo *x = (const coo *)a, *y = (const coo *)b;
    if (x->r != y->r) return x->r < y->r ? -1 : 1;
    if (x->c != y->c) return x->c < y->c ? -1 : 1;
    return 0;
}

int orc_seam_last_matrix(uint32_t **csr_ptr, uint32_t **csr_col, float **csr_val)
{
    *csr_ptr = g_csr_ptr; *csr_col = g_csr_col; *csr_val = g_csr_val;
    return (int)g_rows;
}

static int face_has_vertex(const uint32_t *faces, uint32_t f, uint32_t v)
{
    return faces[3 * (size_t)f] == v || faces[3 * (size_t)f + 1] == v || faces[3 * (size_t)f + 2] == v;
}

int orc_global_seam_leveling(const float *verts, uint32_t Vn, const uint32_t *faces, uint32_t F,
                             const uint32_t *vf_ptr, const uint32_t *vf_idx,
                             const uint32_t *vv_ptr, const uint32_t *vv_idx,
                             const uint32_t *labels, const orc_view *views, uint32_t num_views,
                             int num_threads, uint32_t *row_ptr, uint32_t **row_label_out,
                             float **x_out, float **rhs_out, orc_seam_info *info)
{
    (void)F; (void)num_views; (void)num_threads;
    /* :156-176 unknown numbering: vertex major, labels ascending, label 0 skipped */
    row_ptr[0] = 0;
    for (uint32_t i = 0; i < Vn; ++i) {
        uint32_t tmp[64]; uint32_t n = 0;
        for (uint32_t a = vf_ptr[i]; a < vf_ptr[i + 1]; ++a) {
            uint32_t l = labels[vf_idx[a]];
            if (l == 0) continue;
            uint32_t k = 0; while (k < n && tmp[k] != l) ++k;
            if (k == n && n < 64) tmp[n++] = l;
        }
        row_ptr[i + 1] = row_ptr[i] + n;
    }
    uint32_t R = row_ptr[Vn];
    uint32_t *row_label = (uint32_t *)malloc(sizeof(uint32_t) * (R ? R : 1));
    for (uint32_t i = 0; i < Vn; ++i) {
        uint32_t *dst = row_label + row_ptr[i]; uint32_t n = 0;
        for (uint32_t a = vf_ptr[i]; a < vf_ptr[i + 1]; ++a) {
            uint32_t l = labels[vf_idx[a]];
            if (l == 0) continue;
            uint32_t k = 0; while (k < n && dst[k] != l) ++k;
            if (k < n) continue;
            uint32_t p = n++;
            while (p > 0 && dst[p - 1] > l) { dst[p] = dst[p - 1]; --p; }
            dst[p] = l;
        }
    }
#define ROW_OF(v, l, out) do { out = 0xFFFFFFFFu; for (uint32_t _k = row_ptr[v]; _k < row_ptr[(v) + 1]; ++_k) \
        if (row_label[_k] == (l)) { out = _k; break; } } while (0)

    const float lambda = 0.1f; /* :179 */
    const float lam2 = lambda * lambda;
    size_t cap = (size_t)R * 10 + 16, ncoo = 0;
    coo *tr = (coo *)malloc(sizeof(coo) * cap);
    float *diag = (float *)calloc(R ? R : 1, sizeof(float));
    float *gdiag = (float *)calloc(R ? R : 1, sizeof(float));
    uint32_t *adiag = (uint32_t *)calloc(R ? R : 1, sizeof(uint32_t));
#define PUSH(rr, cc, vv) do { if (ncoo == cap) { cap *= 2; tr = (coo *)realloc(tr, sizeof(coo) * cap); } \
        tr[ncoo].r = (rr); tr[ncoo].c = (cc); tr[ncoo].v = (vv); ++ncoo; } while (0)

    /* Gamma :182-208 -> Gamma^T Gamma contributions */
    uint32_t gamma_rows = 0;
    for (uint32_t i = 0; i < Vn; ++i)
        for (uint32_t j = row_ptr[i]; j < row_ptr[i + 1]; ++j)
            for (uint32_t k = vv_ptr[i]; k < vv_ptr[i + 1]; ++k) {
                uint32_t adj = vv_idx[k];
                for (uint32_t l = row_ptr[adj]; l < row_ptr[adj + 1]; ++l)
                    if (i < adj && row_label[j] == row_label[l]) {
                        PUSH(j, l, -lam2); PUSH(l, j, -lam2);
                        gdiag[j] += lam2; gdiag[l] += lam2;
                        ++gamma_rows;
                    }
            }

    /* A and b :211-237 */
    size_t bcap = (size_t)R + 16, A_rows = 0;
    float *bvec = (float *)malloc(sizeof(float) * 3 * bcap);
    uint32_t *arow_r1 = (uint32_t *)malloc(sizeof(uint32_t) * bcap);
    uint32_t *arow_r2 = (uint32_t *)malloc(sizeof(uint32_t) * bcap);
    for (uint32_t i = 0; i < Vn; ++i)
        for (uint32_t j = row_ptr[i]; j < row_ptr[i + 1]; ++j)
            for (uint32_t k = row_ptr[i]; k < row_ptr[i + 1]; ++k) {
                uint32_t label1 = row_label[j], label2 = row_label[k];
                if (!(label1 < label2)) continue;
                /* find_seam_edges_for_vertex_label_combination :46-84 + calculate_difference :86-138 */
                float c1[3] = {0, 0, 0}, c2[3] = {0, 0, 0}, w1 = 0.0f, w2 = 0.0f;
                int any = 0;
                for (uint32_t a = vv_ptr[i]; a < vv_ptr[i + 1]; ++a) {
                    uint32_t adj = vv_idx[a];
                    if (adj == i) continue;
                    uint32_t ef[16]; uint32_t nef = 0; /* MeshInfo::get_faces_for_edge */
                    for (uint32_t q = vf_ptr[i]; q < vf_ptr[i + 1] && nef < 16; ++q)
                        if (face_has_vertex(faces, vf_idx[q], adj)) ef[nef++] = vf_idx[q];
                    for (uint32_t x = 0; x < nef; ++x)
                        for (uint32_t y = x + 1; y < nef; ++y) {
                            uint32_t fl1 = labels[ef[x]], fl2 = labels[ef[y]];
                            if (!(fl1 < fl2)) { uint32_t t = fl1; fl1 = fl2; fl2 = t; }
                            if (fl1 != label1 || fl2 != label2) continue;
                            const float *v1 = verts + 3 * (size_t)i, *v2 = verts + 3 * (size_t)adj;
                            float d[3] = {v2[0] - v1[0], v2[1] - v1[1], v2[2] - v1[2]};
                            float length = sqrtf(((0.0f + d[0] * d[0]) + d[1] * d[1]) + d[2] * d[2]);
                            if (length == 0.0f) continue;
                            any = 1;
                            float pa[2], pb[2], col[3];
                            const orc_view *va = &views[label1 - 1], *vb = &views[label2 - 1];
                            orc_pixel_coords(va, v1, pa); orc_pixel_coords(va, v2, pb);
                            sample_edge(va, pa, pb, col);
                            for (int c = 0; c < 3; ++c) c1[c] += col[c] * length;
                            w1 += length;
                            orc_pixel_coords(vb, v1, pa); orc_pixel_coords(vb, v2, pb);
                            sample_edge(vb, pa, pb, col);
                            for (int c = 0; c < 3; ++c) c2[c] += col[c] * length;
                            w2 += length;
                        }
                }
                if (!any) continue;
                if (A_rows == bcap) {
                    bcap *= 2;
                    bvec = (float *)realloc(bvec, sizeof(float) * 3 * bcap);
                    arow_r1 = (uint32_t *)realloc(arow_r1, sizeof(uint32_t) * bcap);
                    arow_r2 = (uint32_t *)realloc(arow_r2, sizeof(uint32_t) * bcap);
                }
                for (int c = 0; c < 3; ++c) bvec[3 * A_rows + c] = c2[c] / w2 - c1[c] / w1; /* :131 */
                arow_r1[A_rows] = j; arow_r2[A_rows] = k;
                PUSH(j, k, -1.0f); PUSH(k, j, -1.0f);
                adiag[j]++; adiag[k]++;
                ++A_rows;
            }

    /* Lhs = A^T A + Gamma^T Gamma :245 (full symmetric CSR) */
    for (uint32_t r = 0; r < R; ++r) { diag[r] = (float)adiag[r] + gdiag[r]; PUSH(r, r, diag[r]); }
    qsort(tr, ncoo, sizeof(coo), coo_cmp);
    uint32_t *cp = (uint32_t *)calloc((size_t)R + 1, sizeof(uint32_t));
    uint32_t *cc = (uint32_t *)malloc(sizeof(uint32_t) * (ncoo ? ncoo : 1));
    float *cv = (float *)malloc(sizeof(float) * (ncoo ? ncoo : 1));
    size_t nz = 0;
    for (size_t t = 0; t < ncoo; ++t) {
        if (nz > 0 && t > 0 && tr[t].r == tr[t - 1].r && tr[t].c == tr[t - 1].c) { cv[nz - 1] += tr[t].v; continue; }
        cc[nz] = tr[t].c; cv[nz] = tr[t].v; cp[tr[t].r + 1]++; ++nz;
    }
    for (uint32_t r = 0; r < R; ++r) cp[r + 1] += cp[r];
    free(tr);

    /* Rhs = A^T b :266-270 */
    float *rhs = (float *)calloc(3 * (size_t)(R ? R : 1), sizeof(float));
    for (size_t a = 0; a < A_rows; ++a)
        for (int c = 0; c < 3; ++c) {
            rhs[3 * (size_t)arow_r1[a] + c] += bvec[3 * a + c];
            rhs[3 * (size_t)arow_r2[a] + c] -= bvec[3 * a + c];
        }

    float *x = (float *)calloc(3 * (size_t)(R ? R : 1), sizeof(float));
    /* :257 one thread per colour channel, each CG single threaded */
    #pragma omp parallel for num_threads(3)
    for (int ch = 0; ch < 3; ++ch) {
        float *xr = (float *)calloc(R ? R : 1, sizeof(float));
        float *res = (float *)malloc(sizeof(float) * (R ? R : 1));
        float *p = (float *)malloc(sizeof(float) * (R ? R : 1));
        float *z = (float *)malloc(sizeof(float) * (R ? R : 1));
        float *tmp = (float *)malloc(sizeof(float) * (R ? R : 1));
        const float tol = 0.0001f; const uint32_t max_iters = 1000;
        double acc = 0.0;
        for (uint32_t r = 0; r < R; ++r) { res[r] = rhs[3 * (size_t)r + ch]; acc += (double)res[r] * res[r]; }
        float rhsNorm2 = (float)acc;
        uint32_t it = 0; float tol_error = 0.0f;
        if (rhsNorm2 != 0.0f) {
            float threshold = tol * tol * rhsNorm2;
            float residualNorm2 = rhsNorm2;
            if (!(residualNorm2 < threshold)) {
                acc = 0.0;
                for (uint32_t r = 0; r < R; ++r) {
                    float inv = diag[r] != 0.0f ? 1.0f / diag[r] : 1.0f;
                    p[r] = inv * res[r]; acc += (double)res[r] * p[r];
                }
                float absNew = (float)acc;
                while (it < max_iters) {
                    acc = 0.0;
                    for (uint32_t r = 0; r < R; ++r) {
                        float s = 0.0f;
                        for (uint32_t e = cp[r]; e < cp[r + 1]; ++e) s += cv[e] * p[cc[e]];
                        tmp[r] = s; acc += (double)p[r] * s;
                    }
                    float alpha = absNew / (float)acc;
                    acc = 0.0;
                    for (uint32_t r = 0; r < R; ++r) {
                        xr[r] += alpha * p[r]; res[r] -= alpha * tmp[r]; acc += (double)res[r] * res[r];
                    }
                    residualNorm2 = (float)acc;
                    if (residualNorm2 < threshold) break;
                    acc = 0.0;
                    for (uint32_t r = 0; r < R; ++r) {
                        float inv = diag[r] != 0.0f ? 1.0f / diag[r] : 1.0f;
                        z[r] = inv * res[r]; acc += (double)res[r] * z[r];
                    }
                    float absOld = absNew; absNew = (float)acc;
                    float beta = absNew / absOld;
                    for (uint32_t r = 0; r < R; ++r) p[r] = z[r] + beta * p[r];
                    ++it;
                }
            }
            tol_error = sqrtf(residualNorm2 / rhsNorm2);
        }
        /* :277 subtract the global mean */
        acc = 0.0;
        for (uint32_t r = 0; r < R; ++r) acc += xr[r];
        float mean = R ? (float)(acc / (double)R) : 0.0f;
        for (uint32_t r = 0; r < R; ++r) x[3 * (size_t)r + ch] = xr[r] - mean;
        info->iterations[ch] = it; info->residual[ch] = tol_error;
        free(xr); free(res); free(p); free(z); free(tmp);
    }

    info->num_rows = R; info->num_a_rows = (uint32_t)A_rows; info->num_gamma_rows = gamma_rows;
    info->nnz_full = nz;
    free(g_csr_ptr); free(g_csr_col); free(g_csr_val);
    g_csr_ptr = cp; g_csr_col = cc; g_csr_val = cv; g_rows = R;
    free(diag); free(gdiag); free(adiag); free(bvec); free(arow_r1); free(arow_r2);
    if (row_label_out) *row_label_out = row_label; else free(row_label);
    if (x_out) *x_out = x; else free(x);
    if (rhs_out) *rhs_out = rhs; else free(rhs);
    return 0;
}
