// refshim: rayint acc::BVHTree stand-in: forwards to the oracle's BVH (oracle/bvh.c), so that the
// ray/triangle arithmetic is the oracle's restatement; only the call protocol is rayint's
// (see ../README.md).  intersect() reports "some hit in [tmin, tmax]" -- the only thing
// calculate_data_costs.cpp:207-211 looks at.
#pragma once
#include <limits>
#include <vector>
#include "oracle.h"

namespace acc {

template <typename IdxType, typename Vec3fType>
class BVHTree {
public:
    struct Ray { Vec3fType origin; Vec3fType dir; float tmin; float tmax; };
    struct Hit { float t; IdxType idx; Vec3fType bcoords; };

    BVHTree(std::vector<IdxType> const& faces, std::vector<Vec3fType> const& vertices, int /*max_threads*/ = 0) {
        f.assign(faces.begin(), faces.end());        // orc_bvh keeps pointers into these
        v.resize(vertices.size() * 3);
        for (std::size_t i = 0; i < vertices.size(); ++i) for (int k = 0; k < 3; ++k) v[3 * i + k] = vertices[i][k];
        bvh = orc_bvh_build(v.data(), f.data(), static_cast<uint32_t>(faces.size() / 3));
    }
    ~BVHTree() { orc_bvh_free(bvh); }
    bool intersect(Ray ray, Hit* hit) const {
        float o[3] = { ray.origin[0], ray.origin[1], ray.origin[2] };
        float d[3] = { ray.dir[0], ray.dir[1], ray.dir[2] };
        bool const any = orc_bvh_occluded(bvh, o, d, ray.tmin, ray.tmax) != 0;
        if (any && hit) { hit->t = ray.tmin; hit->idx = IdxType(0); }
        return any;
    }
private:
    BVHTree(BVHTree const&);
    BVHTree& operator=(BVHTree const&);
    std::vector<uint32_t> f;
    std::vector<float> v;
    orc_bvh* bvh;
};

}  // namespace acc
