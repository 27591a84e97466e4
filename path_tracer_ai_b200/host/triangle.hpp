// triangle.hpp — host mirror of the reference's Triangle record (include/triangle.hpp:8-21).
// Intersection is not done on the host: Triangle::intersect's replacement is tri_test in
// csrc/exact.cuh.
#pragma once
#include "vec.hpp"

namespace b2pt {

struct Triangle {
    vec3 v0, v1, v2;
    vec3 n0, n1, n2;
    vec2 uv0, uv1, uv2;
    int materialId = 0;

    Triangle() = default;
    Triangle(const vec3& a, const vec3& b, const vec3& c, const vec3& na, const vec3& nb, const vec3& nc,
             const vec2& ta, const vec2& tb, const vec2& tc, int matId)
        : v0(a), v1(b), v2(c), n0(na), n1(nb), n2(nc), uv0(ta), uv1(tb), uv2(tc), materialId(matId) {}
};

}  // namespace b2pt
