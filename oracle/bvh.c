/* oracle/bvh.c -- TEST INFRASTRUCTURE (see oracle.h).
 * Stand-in for rayint's acc::BVHTree<unsigned, Vec3f> (absent dependency; used at
 * libs/tex/calculate_data_costs.cpp:23,144,201-209).  The reference only asks "is there ANY
 * triangle with tmin <= t <= tmax along the ray", so any exact BVH gives the same answer provided
 * (a) box tests are conservative and (b) the triangle test is the one below.  The triangle test
 * (Moeller-Trumbore, fp32, no FMA) IS the definition of visibility for this repo; the CUDA path
 * restates it operation by operation. */
#include "oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    float lo[3], hi[3];
    uint32_t left;   /* inner: index of left child (right = left+1); leaf: first triangle slot */
    uint32_t count;  /* 0 for inner nodes, otherwise number of triangles */
} bvh_node;

struct orc_bvh {
    bvh_node *nodes;
    uint32_t num_nodes;
    uint32_t *tri_ids;
    const float *verts;
    const uint32_t *faces;
    uint32_t num_faces;
    float pad;
};

/* Moeller-Trumbore.  Returns 1 and *t when the ray hits the triangle (any t). */
static inline int tri_hit(const float *a, const float *b, const float *c, const float o[3],
                          const float d[3], float *t_out)
{
    float e1x = b[0] - a[0], e1y = b[1] - a[1], e1z = b[2] - a[2];
    float e2x = c[0] - a[0], e2y = c[1] - a[1], e2z = c[2] - a[2];
    float px = d[1] * e2z - d[2] * e2y;
    float py = d[2] * e2x - d[0] * e2z;
    float pz = d[0] * e2y - d[1] * e2x;
    float det = (e1x * px + e1y * py) + e1z * pz;
    if (det == 0.0f) return 0;
    float inv = 1.0f / det;
    float tx = o[0] - a[0], ty = o[1] - a[1], tz = o[2] - a[2];
    float u = ((tx * px + ty * py) + tz * pz) * inv;
    if (!(u >= 0.0f && u <= 1.0f)) return 0;
    float qx = ty * e1z - tz * e1y;
    float qy = tz * e1x - tx * e1z;
    float qz = tx * e1y - ty * e1x;
    float v = ((d[0] * qx + d[1] * qy) + d[2] * qz) * inv;
    if (!(v >= 0.0f && u + v <= 1.0f)) return 0;
    *t_out = ((e2x * qx + e2y * qy) + e2z * qz) * inv;
    return 1;
}

static inline int tri_occludes(const orc_bvh *b, uint32_t f, const float o[3], const float d[3],
                               float tmin, float tmax)
{
    const uint32_t *idx = b->faces + 3 * (size_t)f;
    float t;
    if (!tri_hit(b->verts + 3 * (size_t)idx[0], b->verts + 3 * (size_t)idx[1],
                 b->verts + 3 * (size_t)idx[2], o, d, &t))
        return 0;
    return t >= tmin && t <= tmax;
}

int orc_brute_occluded(const float *verts, const uint32_t *faces, uint32_t num_faces,
                       const float o[3], const float d[3], float tmin, float tmax)
{
    orc_bvh b;
    b.verts = verts;
    b.faces = faces;
    for (uint32_t f = 0; f < num_faces; ++f)
        if (tri_occludes(&b, f, o, d, tmin, tmax)) return 1;
    return 0;
}

typedef struct { float c[3]; float lo[3], hi[3]; uint32_t id; } prim;

static void bounds_of(const prim *p, uint32_t n, float lo[3], float hi[3], float clo[3], float chi[3])
{
    for (int k = 0; k < 3; ++k) { lo[k] = clo[k] = FLT_MAX; hi[k] = chi[k] = -FLT_MAX; }
    for (uint32_t i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) {
            if (p[i].lo[k] < lo[k]) lo[k] = p[i].lo[k];
            if (p[i].hi[k] > hi[k]) hi[k] = p[i].hi[k];
            if (p[i].c[k] < clo[k]) clo[k] = p[i].c[k];
            if (p[i].c[k] > chi[k]) chi[k] = p[i].c[k];
        }
}

/* quickselect partition of prims around the median centroid on `axis` */
static void select_median(prim *p, uint32_t n, int axis, uint32_t k)
{
    uint32_t lo = 0, hi = n - 1;
    while (lo < hi) {
        float pivot = p[lo + (hi - lo) / 2].c[axis];
        uint32_t i = lo, j = hi;
        while (i <= j) {
            while (p[i].c[axis] < pivot) ++i;
            while (p[j].c[axis] > pivot) { if (j == 0) break; --j; }
            if (i <= j) {
                prim t = p[i]; p[i] = p[j]; p[j] = t;
                ++i;
                if (j == 0) break;
                --j;
            }
        }
        if (k <= j) hi = j; else if (k >= i) lo = i; else return;
    }
}

#define LEAF_SIZE 4

static void build_rec(orc_bvh *b, prim *p, uint32_t first, uint32_t n, uint32_t node, uint32_t *next)
{
    float lo[3], hi[3], clo[3], chi[3];
    bounds_of(p + first, n, lo, hi, clo, chi);
    bvh_node *nd = &b->nodes[node];
    for (int k = 0; k < 3; ++k) { nd->lo[k] = lo[k] - b->pad; nd->hi[k] = hi[k] + b->pad; }
    int axis = 0;
    float ext = chi[0] - clo[0];
    for (int k = 1; k < 3; ++k) if (chi[k] - clo[k] > ext) { ext = chi[k] - clo[k]; axis = k; }
    if (n <= LEAF_SIZE || ext <= 0.0f) {
        nd->left = first;
        nd->count = n;
        return;
    }
    uint32_t mid = n / 2;
    select_median(p + first, n, axis, mid);
    uint32_t l = *next;
    *next += 2;
    nd->left = l;
    nd->count = 0;
    build_rec(b, p, first, mid, l, next);
    build_rec(b, p, first + mid, n - mid, l + 1, next);
}

orc_bvh *orc_bvh_build(const float *verts, const uint32_t *faces, uint32_t num_faces)
{
    orc_bvh *b = (orc_bvh *)calloc(1, sizeof(orc_bvh));
    b->verts = verts;
    b->faces = faces;
    b->num_faces = num_faces;
    prim *p = (prim *)malloc(sizeof(prim) * (num_faces ? num_faces : 1));
    float slo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, shi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (uint32_t f = 0; f < num_faces; ++f) {
        p[f].id = f;
        for (int k = 0; k < 3; ++k) { p[f].lo[k] = FLT_MAX; p[f].hi[k] = -FLT_MAX; }
        for (int c = 0; c < 3; ++c) {
            const float *v = verts + 3 * (size_t)faces[3 * (size_t)f + c];
            for (int k = 0; k < 3; ++k) {
                if (v[k] < p[f].lo[k]) p[f].lo[k] = v[k];
                if (v[k] > p[f].hi[k]) p[f].hi[k] = v[k];
            }
        }
        for (int k = 0; k < 3; ++k) {
            p[f].c[k] = 0.5f * (p[f].lo[k] + p[f].hi[k]);
            if (p[f].lo[k] < slo[k]) slo[k] = p[f].lo[k];
            if (p[f].hi[k] > shi[k]) shi[k] = p[f].hi[k];
        }
    }
    float diag = 0.0f;
    if (num_faces)
        diag = sqrtf((shi[0] - slo[0]) * (shi[0] - slo[0]) + (shi[1] - slo[1]) * (shi[1] - slo[1])
                     + (shi[2] - slo[2]) * (shi[2] - slo[2]));
    b->pad = 1e-5f * diag; /* conservative boxes: never cull what the triangle test accepts */
    b->nodes = (bvh_node *)malloc(sizeof(bvh_node) * (2 * (size_t)num_faces + 2));
    b->tri_ids = (uint32_t *)malloc(sizeof(uint32_t) * (num_faces ? num_faces : 1));
    uint32_t next = 1;
    if (num_faces) build_rec(b, p, 0, num_faces, 0, &next);
    else { b->nodes[0].count = 0; b->nodes[0].left = 0; next = 0; }
    b->num_nodes = next;
    for (uint32_t f = 0; f < num_faces; ++f) b->tri_ids[f] = p[f].id;
    free(p);
    return b;
}

void orc_bvh_free(orc_bvh *b)
{
    if (!b) return;
    free(b->nodes);
    free(b->tri_ids);
    free(b);
}

static inline int box_hit(const bvh_node *n, const float o[3], const float inv[3], float tmin,
                          float tmax)
{
    float t0 = tmin, t1 = tmax;
    for (int k = 0; k < 3; ++k) {
        float a = (n->lo[k] - o[k]) * inv[k];
        float c = (n->hi[k] - o[k]) * inv[k];
        if (a != a || c != c) continue; /* 0 * inf: origin on the slab plane, parallel ray */
        float near = a < c ? a : c;
        float far = a < c ? c : a;
        if (near > t0) t0 = near;
        if (far < t1) t1 = far;
    }
    return t0 <= t1 * 1.00001f + 1e-30f;
}

int orc_bvh_occluded(const orc_bvh *b, const float o[3], const float d[3], float tmin, float tmax)
{
    if (!b->num_nodes) return 0;
    float inv[3];
    for (int k = 0; k < 3; ++k) inv[k] = 1.0f / d[k];
    uint32_t stack[128];
    int sp = 0;
    stack[sp++] = 0;
    /* widen the parametric interval a little: box culling must stay conservative */
    float bt0 = tmin * 0.999f, bt1 = tmax * 1.001f;
    while (sp) {
        const bvh_node *n = &b->nodes[stack[--sp]];
        if (!box_hit(n, o, inv, bt0, bt1)) continue;
        if (n->count) {
            for (uint32_t i = 0; i < n->count; ++i)
                if (tri_occludes(b, b->tri_ids[n->left + i], o, d, tmin, tmax)) return 1;
        } else {
            stack[sp++] = n->left;
            stack[sp++] = n->left + 1;
        }
    }
    return 0;
}
