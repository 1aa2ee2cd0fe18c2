// bvh.cuh -- any-hit traversal of the two-child LBVH (device side, included by datacosts.cu).
//
// Replaces acc::BVHTree::intersect as used at calculate_data_costs.cpp:201-209: the caller only
// needs "is there any triangle with tmin <= t <= tmax".  Box tests are conservative (padded boxes,
// widened interval), the triangle test restates oracle/bvh.c tri_hit() operation by operation, so
// the answer is independent of the tree shape.  Requires -fmad=false.
#pragma once
#include "common.cuh"

namespace b2 {

__device__ __forceinline__ bool tri_occludes(const float *__restrict__ t9, float ox, float oy, float oz,
                                             float dx, float dy, float dz, float tmin, float tmax)
{
    const float ax = t9[0], ay = t9[1], az = t9[2];
    const float e1x = t9[3] - ax, e1y = t9[4] - ay, e1z = t9[5] - az;
    const float e2x = t9[6] - ax, e2y = t9[7] - ay, e2z = t9[8] - az;
    const float px = dy * e2z - dz * e2y;
    const float py = dz * e2x - dx * e2z;
    const float pz = dx * e2y - dy * e2x;
    const float det = (e1x * px + e1y * py) + e1z * pz;
    if (det == 0.0f) return false;
    const float inv = 1.0f / det;
    const float tx = ox - ax, ty = oy - ay, tz = oz - az;
    const float u = ((tx * px + ty * py) + tz * pz) * inv;
    if (!(u >= 0.0f && u <= 1.0f)) return false;
    const float qx = ty * e1z - tz * e1y;
    const float qy = tz * e1x - tx * e1z;
    const float qz = tx * e1y - ty * e1x;
    const float v = ((dx * qx + dy * qy) + dz * qz) * inv;
    if (!(v >= 0.0f && u + v <= 1.0f)) return false;
    const float t = ((e2x * qx + e2y * qy) + e2z * qz) * inv;
    return t >= tmin && t <= tmax;
}

// Slab test.  Box culling only has to be conservative, not bit-reproducible, so it uses explicit
// FMAs ((lo - o) * inv == lo * inv - o * inv up to rounding; boxes are padded and the interval is
// widened by the caller) and min/max that drop NaNs (0 * inf when the origin lies on a slab plane of
// an axis-parallel ray): fminf/fmaxf return the non-NaN operand.
__device__ __forceinline__ bool box_hit(const float *lo, const float *hi, float nox, float noy, float noz,
                                        float ix, float iy, float iz, float t0, float t1, float sx, float sy,
                                        float sz)
{
    const float ax = __fmaf_rn(lo[0], ix, nox), cx = __fmaf_rn(hi[0], ix, nox);
    const float ay = __fmaf_rn(lo[1], iy, noy), cy = __fmaf_rn(hi[1], iy, noy);
    const float az = __fmaf_rn(lo[2], iz, noz), cz = __fmaf_rn(hi[2], iz, noz);
    // per-axis slack s* covers the rounding of -o*inv: an axis the ray is (almost) parallel to gets a
    // huge slack, i.e. stops culling on that axis only; the other two axes keep culling
    t0 = fmaxf(t0, fmaxf(fmaxf(fminf(ax, cx) - sx, fminf(ay, cy) - sy), fminf(az, cz) - sz));
    t1 = fminf(t1, fminf(fminf(fmaxf(ax, cx) + sx, fmaxf(ay, cy) + sy), fmaxf(az, cz) + sz));
    return t0 <= t1 * 1.0001f + 1e-30f;
}

// nodes: N-1 internal nodes (root = 0); tri: 9 floats per Morton-ordered triangle.
__device__ __forceinline__ bool bvh_occluded(const BvhNode *__restrict__ nodes,
                                             const float *__restrict__ tri, uint32_t num_tris,
                                             float ox, float oy, float oz, float dx, float dy, float dz,
                                             float tmin, float tmax, uint32_t *overflow)
{
    if (num_tris == 0) return false;
    if (num_tris == 1) return tri_occludes(tri, ox, oy, oz, dx, dy, dz, tmin, tmax);
    const float ix = 1.0f / dx, iy = 1.0f / dy, iz = 1.0f / dz;
    const float bt0 = tmin * 0.999f, bt1 = tmax * 1.001f;
    const float nox = -ox * ix, noy = -oy * iy, noz = -oz * iz;
    float sx = 1e-6f * fabsf(nox), sy = 1e-6f * fabsf(noy), sz = 1e-6f * fabsf(noz);
    if (!(sx == sx)) sx = INFINITY;  // NaN (0 * inf): no culling on that axis
    if (!(sy == sy)) sy = INFINITY;
    if (!(sz == sz)) sz = INFINITY;
    int stack[100];
    int sp = 0;
    int node = 0;
    for (;;) {
        const float4 *n4 = reinterpret_cast<const float4 *>(nodes + node);
        const float4 q0 = __ldg(n4), q1 = __ldg(n4 + 1), q2 = __ldg(n4 + 2), q3 = __ldg(n4 + 3);
        const float lo0[3] = {q0.x, q0.y, q0.z}, hi0[3] = {q0.w, q1.x, q1.y};
        const float lo1[3] = {q1.z, q1.w, q2.x}, hi1[3] = {q2.y, q2.z, q2.w};
        const int left = __float_as_int(q3.x), right = __float_as_int(q3.y);
        const bool h0 = box_hit(lo0, hi0, nox, noy, noz, ix, iy, iz, bt0, bt1, sx, sy, sz);
        const bool h1 = box_hit(lo1, hi1, nox, noy, noz, ix, iy, iz, bt0, bt1, sx, sy, sz);
        int next = -1;
        if (h0) {
            if (left < 0) {
                if (tri_occludes(tri + 9 * (size_t)(~left), ox, oy, oz, dx, dy, dz, tmin, tmax)) return true;
            } else next = left;
        }
        if (h1) {
            if (right < 0) {
                if (tri_occludes(tri + 9 * (size_t)(~right), ox, oy, oz, dx, dy, dz, tmin, tmax)) return true;
            } else if (next < 0) next = right;
            else if (sp < 100) stack[sp++] = right;
            else if (overflow) atomicOr(overflow, 1u);   // a subtree would be dropped: reported (B2TEX_ERR_LIMITS), never silent
        }
        if (next >= 0) { node = next; continue; }
        if (sp == 0) return false;
        node = stack[--sp];
    }
}

}  // namespace b2
