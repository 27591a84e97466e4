// traverse.cuh — device traversal building blocks shared by the query kernels (trace.cu) and the
// wavefront renderer (render.cu).  Replaces BVH::intersect / intersectNode (reference
// include/bvh.hpp:37-39, :74-116).
//
//   closest_exact_dfs  the reference recursion itself, flattened: left-then-right DFS over the
//                      implicit reference tree, global tMax shrink, reference tie rules.  Slow
//                      (no ordering, binary tree) but exact by construction; used for the rays the
//                      fast kernels cannot certify, and alone under B2PT_FLAG_EXACT_ONLY.
//   node_test4         the reference slab test (aabb.hpp:13-25) of four children of a wide node
//   box_pass           the same test for one box: the certificate, the hoisted leaves, the exact recursion
//
// Exactness.  A triangle X is a CANDIDATE for a ray iff the exact box of X's reference leaf
// passes the reference slab test at the ray's initial tMax (T0) and Triangle::intersect accepts X in
// [tMin, T0].  The traversal tree only has to be CONSERVATIVE above the leaves: its inner boxes are exact min/max
// unions of exact leaf boxes, every box is tested with the reference's own arithmetic (node_test4), and that
// arithmetic is monotone — a superset box passes whenever a box inside it does — so a visible leaf is never culled,
// whatever the shape of the tree, and a visited leaf is visible by the very test the reference applies.
//
// (Measured and dropped in round 2: a one-FMA-per-plane slab test, fma(plane, 1/d, -o/d -+ slack), made provably
// conservative with per-ray slack constants and re-checked exactly per accepted triangle.  21 instead of 32
// arithmetic instructions per child, but six more live registers per ray: closest hit 1760 -> 1690, any-hit
// 2750 -> 2230 Mrays/s, 1M-triangle render 216 -> 208 Msamples/s.  profiles/r02_experiments.md.)
#pragma once
#include "ctx.cuh"

namespace b2pt {

struct RayQ {
    V3 o, d;      // d normalised (ray.hpp:12)
    V3 invD;      // 1.0f / d, IEEE (aabb.hpp:15)
    float T0;     // initial ray.tMax
};

__device__ __forceinline__ RayQ make_rayq_normalised(V3 o, V3 d, float T0) {
    RayQ r;
    r.o = o; r.d = d; r.T0 = T0;
    r.invD = mk3(B2PT_DIV(1.0f, r.d.x), B2PT_DIV(1.0f, r.d.y), B2PT_DIV(1.0f, r.d.z));
    return r;
}

__device__ __forceinline__ RayQ make_rayq(V3 o, V3 d_unnormalised, float T0) {
    return make_rayq_normalised(o, vnormalize(d_unnormalised), T0);
}

// Distance cull of closest-hit queries: once a candidate with distance t is known, subtrees entered later than
// t * (1 + 2^-10) + 2^-12 * (R + |o|) are skipped (R = S.coord_bound, the largest |coordinate| in the scene).  The
// computed t of a Möller–Trumbore hit is off by an ABSOLUTE amount that grows with the distance between the origin and
// the triangle, not with t, so the slack has a part proportional to the coordinate range as well as a relative one.
__device__ __forceinline__ float cull_after_hit(const DeviceScene& S, const RayQ& r, float t) {
    const float reach = S.coord_bound + fmaxf(fmaxf(fabsf(r.o.x), fabsf(r.o.y)), fabsf(r.o.z));
    return fminf(r.T0, __fmaf_rn(t, 0.0009765625f, t) + reach * 2.44140625e-4f);
}

// A ray with a non-finite component never hits anything in the reference.
//  * NaN component: every slab comparison against a NaN is false, so that axis (for a NaN origin: every axis)
//    constrains nothing (aabb.hpp:18-21) and the DFS walks into leaves; there Triangle::intersect computes
//    t = NaN — every `<`/`>` rejection is false for a NaN, so it "accepts" (triangle.hpp:34-58) — and
//    bvh.hpp:88 discards it because `NaN < isect.t` is false.
//  * infinite origin component: s = o - v0 is infinite, every product in q = cross(s, e1) and dot(e2, q) that
//    touches it is +-inf or NaN, so t is never finite; t = +inf passes `t > tMax` (tMax = +inf) but fails
//    `t < isect.t` (= +inf).
//  * infinite direction (normalising a vector whose squared length underflows): invD = 0, every t0/t1 is 0 or
//    NaN, the root box ends with tMax = 0 <= tMin.
// The answer is a miss / "not occluded" — in the first two cases after walking most of the tree.  The kernels
// return it at once: such a ray would also pass the boxes of EMPTY child slots (NaN slabs), which are not valid
// triangle ranges, and would hold its warp for a full-scene walk.  These rays do occur: a zero interpolated normal
// makes the dielectric branch's refract() return vec3(0) (renderer.hpp:233), whose Ray ctor normalisation is NaN
// (ray.hpp:12).
__device__ __forceinline__ bool ray_has_nan(const RayQ& r) {
    const float inf = B2PT_INF;
    return !((fabsf(r.o.x) < inf) & (fabsf(r.o.y) < inf) & (fabsf(r.o.z) < inf) &
             (fabsf(r.d.x) < inf) & (fabsf(r.d.y) < inf) & (fabsf(r.d.z) < inf));
}

// The reference slab test (aabb.hpp:13-25), bit for bit.
__device__ __forceinline__ bool box_pass(float4 lo, float4 hi, V3 o, V3 invD, float T) {
    float tmin = B2PT_TMIN, tmax = T;
    slab_axis(lo.x, hi.x, o.x, invD.x, tmin, tmax);
    slab_axis(lo.y, hi.y, o.y, invD.y, tmin, tmax);
    slab_axis(lo.z, hi.z, o.z, invD.z, tmin, tmax);
    return tmax > tmin;
}

// Entry distance of a box in the reference's arithmetic: the running tMin after the three axes (aabb.hpp:13-25).
__device__ __forceinline__ float box_entry(float4 lo, float4 hi, V3 o, V3 invD) {
    float tmin = B2PT_TMIN, tmax = B2PT_INF;
    slab_axis(lo.x, hi.x, o.x, invD.x, tmin, tmax);
    slab_axis(lo.y, hi.y, o.y, invD.y, tmin, tmax);
    slab_axis(lo.z, hi.z, o.z, invD.z, tmin, tmax);
    return tmin;
}

// ---- the certificate ---------------------------------------------------------------------------------------------
// The fast traversals see a set of CANDIDATES (triangles of visited = visible-at-T0 leaves that Möller–Trumbore accepts
// in [tMin, T0]); subtrees they skipped were entered later than the final cull distance, so their candidates are
// farther still.  Let X be the UNIQUE candidate with the smallest t = m, L its reference leaf, e = L's entry distance,
// and s2 the smallest t among the other candidates seen (+inf if none).  A box that passes at T0 passes at T iff
// T > e (the running tMax is min(T, far planes) and far > e already).  While the reference recursion runs, ray.tMax is
// T0 or the t of an accepted candidate — i.e. T0, some value >= s2, or some value beyond the cull distance — until X is
// accepted.  So if e < min(T0, s2) (and e is not absurdly beyond m, see below) L is entered whenever the recursion
// reaches it, X is accepted there (m <= ray.tMax), nothing else can be accepted with t <= m afterwards, and the
// combine picks the smaller t: the reference returns X.  The usual case is e < m; e in [m, s2) is a hit on the entry
// face of its own leaf box (round 1 sent those to the exact recursion).  e <= m (1 + 2^-12) keeps e well inside the
// cull slack, so "unseen candidates are beyond e" needs no more than the slack already assumes.  A miss (no candidate)
// is certified as it is; bit-equal ties at the minimum are not (closest_careful resolves most of them).
__device__ __forceinline__ bool certify_unique(const DeviceScene& S, const RayQ& r, float m, int leaf, float s2) {
    const float e = box_entry(__ldg(&S.leaf_lo[leaf]), __ldg(&S.leaf_hi[leaf]), r.o, r.invD);
    return (e < fminf(s2, r.T0)) & (e <= __fmaf_rn(m, 0.000244140625f, m));
}

// Does the exact box of reference leaf `leaf` pass the reference slab test with ray.tMax == T?
__device__ __forceinline__ bool leaf_visible(const DeviceScene& S, int leaf, const RayQ& r, float T) {
    return box_pass(__ldg(&S.leaf_lo[leaf]), __ldg(&S.leaf_hi[leaf]), r.o, r.invD, T);
}

// Children 4k..4k+3 of a wide node, planes ordered along the ray: WideNode is lox|loy|loz|hix|hiy|hiz (8 floats
// each = two float4), so the near plane of axis a is float4 index 2a + (invD[a] < 0 ? 6 : 0) + k and the far plane
// the other one.  Selecting the PLANE by the ray's sign once per node (an address offset) instead of swapping the two
// products per child (aabb.hpp:17) gives bit-identical t values.  Slab-tests the four boxes at T0 with the reference's
// arithmetic; pass[s] / tmin[s] per child.
struct Node4 { float tmin[4]; bool pass[4]; uint32_t code[4]; };
__device__ __forceinline__ void node_test4(const WideNode* nd, int k, const RayQ& r, Node4& out) {
    const float4* p = reinterpret_cast<const float4*>(nd);
    const int nx = r.invD.x < 0.0f ? 6 : 0, ny = r.invD.y < 0.0f ? 6 : 0, nz = r.invD.z < 0.0f ? 6 : 0;
    const float4 a = __ldg(p + nx + k), b = __ldg(p + 2 + ny + k), c = __ldg(p + 4 + nz + k);
    const float4 d = __ldg(p + 6 - nx + k), e = __ldg(p + 8 - ny + k), f = __ldg(p + 10 - nz + k);
    const uint4 cc = __ldg(reinterpret_cast<const uint4*>(nd->child) + k);
    const float nxs[4] = {a.x, a.y, a.z, a.w}, nys[4] = {b.x, b.y, b.z, b.w}, nzs[4] = {c.x, c.y, c.z, c.w};
    const float fxs[4] = {d.x, d.y, d.z, d.w}, fys[4] = {e.x, e.y, e.z, e.w}, fzs[4] = {f.x, f.y, f.z, f.w};
    out.code[0] = cc.x; out.code[1] = cc.y; out.code[2] = cc.z; out.code[3] = cc.w;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        float tmin = B2PT_TMIN, tmax = r.T0;
        slab_axis_nf(nxs[s], fxs[s], r.o.x, r.invD.x, tmin, tmax);
        slab_axis_nf(nys[s], fys[s], r.o.y, r.invD.y, tmin, tmax);
        slab_axis_nf(nzs[s], fzs[s], r.o.z, r.invD.z, tmin, tmax);
        out.tmin[s] = tmin;
        out.pass[s] = tmax > tmin;
    }
}

// Triangle i (reference id): Möller–Trumbore exactly as the reference computes it; `leaf` = its reference leaf.
__device__ __forceinline__ bool tri_fetch_test(const DeviceScene& S, int i, const RayQ& r, float tmax, float& t, float& u, float& v, int& leaf) {
    float4 a = __ldg(&S.tri[3ll * i + 0]);
    float4 b = __ldg(&S.tri[3ll * i + 1]);
    float4 c = __ldg(&S.tri[3ll * i + 2]);
    leaf = __float_as_int(a.w);
    return tri_test(mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(c.x, c.y, c.z), r.o, r.d, tmax, t, u, v);
}

// ---- the reference recursion, flattened ----------------------------------------------------------
// bvh.hpp:74-116.  The recursive combine `left.t < right.t ? left : right` (ties -> right) over
// fresh per-child Intersections is the associative fold "a later leaf's result replaces the current
// one iff its t <= current t", applied to leaf results in DFS order; a leaf's own result starts from
// t = +inf and takes a triangle iff its t is strictly smaller (first wins inside a leaf, :88).
__device__ __forceinline__ void closest_exact_dfs(const DeviceScene& S, const RayQ& r, HitRec& out) {
    out.t = B2PT_INF; out.tri = -1; out.u = 0.0f; out.v = 0.0f;
    if (S.nnodes == 0 || ray_has_nan(r)) return;
    float tmax = r.T0;          // ray.tMax, shrunk globally (bvh.hpp:90)
    int stack[48];
    int sp = 0;
    int node = 0;
    while (true) {
        bool descend = false;
        if (box_pass(__ldg(&S.node_lo[node]), __ldg(&S.node_hi[node]), r.o, r.invD, tmax)) {
            int4 info = __ldg(&S.node_info[node]);
            if (info.w >= 0) {
                float lt = B2PT_INF, lu = 0.0f, lv = 0.0f; int ltri = -1;
                for (int i = info.x; i < info.y; ++i) {
                    float t, u, v; int leaf;
                    if (tri_fetch_test(S, i, r, tmax, t, u, v, leaf)) {
                        if (t < lt) { lt = t; lu = u; lv = v; ltri = i; tmax = t; }
                    }
                }
                if (ltri >= 0 && lt <= out.t) { out.t = lt; out.tri = ltri; out.u = lu; out.v = lv; }
            } else {
                stack[sp++] = info.z;   // right child, visited after the whole left subtree
                node = node + 1;        // left child
                descend = true;
            }
        }
        if (!descend) {
            if (sp == 0) break;
            node = stack[--sp];
        }
    }
}

}  // namespace b2pt
