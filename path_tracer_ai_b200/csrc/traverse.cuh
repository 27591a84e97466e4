// traverse.cuh — device traversal routines shared by the query kernels (trace.cu) and the
// wavefront renderer (render.cu).  Replaces BVH::intersect / intersectNode (reference
// include/bvh.hpp:37-39, :74-116) with three routines:
//
//   closest_exact_dfs  the reference recursion itself, flattened: left-then-right DFS over the
//                      implicit reference tree, global tMax shrink, reference tie rules.  Slow
//                      (no ordering, binary tree) but exact by construction; used for the rays the
//                      fast kernel cannot certify, and alone under B2PT_FLAG_EXACT_ONLY.
//   closest_octet      ordered, warp-cooperative (8 lanes per ray) traversal of the 8-wide collapse.
//                      Returns the candidate plus a "certified" verdict: certified
//                      results are provably what the reference returns; the rest go to
//                      closest_exact_dfs.
//   any_octet          boolean occlusion query (shadow rays, renderer.hpp:274-278), exact without
//                      any fallback because its answer does not depend on traversal history.
#pragma once
#include "ctx.cuh"

namespace b2pt {

struct RayQ {
    V3 o, d;      // d normalised (ray.hpp:12)
    V3 invD;      // 1.0f / d, IEEE (aabb.hpp:15)
    float T0;     // initial ray.tMax
};

__device__ __forceinline__ RayQ make_rayq(V3 o, V3 d_unnormalised, float T0) {
    RayQ r;
    r.o = o;
    r.d = vnormalize(d_unnormalised);
    r.invD = mk3(B2PT_DIV(1.0f, r.d.x), B2PT_DIV(1.0f, r.d.y), B2PT_DIV(1.0f, r.d.z));
    r.T0 = T0;
    return r;
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
// return it at once: such a ray would also pass the boxes of EMPTY child slots (lo = +inf, hi = -inf give NaN
// slabs), which are not valid triangle ranges, and would hold its warp for a full-scene walk.  These rays do
// occur: a zero interpolated normal makes the dielectric branch's refract() return vec3(0) (renderer.hpp:233),
// whose Ray ctor normalisation is NaN (ray.hpp:12).
__device__ __forceinline__ bool ray_has_nan(const RayQ& r) {
    const float inf = B2PT_INF;
    return !((fabsf(r.o.x) < inf) & (fabsf(r.o.y) < inf) & (fabsf(r.o.z) < inf) &
             (fabsf(r.d.x) < inf) & (fabsf(r.d.y) < inf) & (fabsf(r.d.z) < inf));
}

__device__ __forceinline__ bool box_pass(float4 lo, float4 hi, const RayQ& r, float T, float& entry) {
    float tmin = B2PT_TMIN, tmax = T;
    slab_axis(lo.x, hi.x, r.o.x, r.invD.x, tmin, tmax);
    slab_axis(lo.y, hi.y, r.o.y, r.invD.y, tmin, tmax);
    slab_axis(lo.z, hi.z, r.o.z, r.invD.z, tmin, tmax);
    entry = tmin;
    return tmax > tmin;
}

// Latency hints (ncu: long-scoreboard is the top stall of every traversal kernel on the 1M-triangle scene;
// a leaf's triangles span 3-4 cache lines that the test loop would otherwise miss on one after the other).
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// Lines 2.. of a leaf's triangle run (the loop's first load fetches line 1 itself).
__device__ __forceinline__ void prefetch_leaf_rest(const DeviceScene& S, int first, int cnt) {
    const char* b = reinterpret_cast<const char*>(S.tri + 3ll * first);
    const int bytes = cnt * 48;
    if (bytes > 128) prefetch_l1(b + 128);
    if (bytes > 256) prefetch_l1(b + 256);
    prefetch_l1(b + bytes - 16);
}
// What a stack entry will touch first when it is popped: an inner node's 224 B or a leaf's first triangles.
__device__ __forceinline__ void prefetch_child(const DeviceScene& S, uint32_t code) {
    if (code & B2PT_CHILD_LEAF) {
        prefetch_l1(S.tri + 3ll * (code & 0x0FFFFFFF));
    } else {
        const char* b = reinterpret_cast<const char*>(S.wide + code);
        prefetch_l1(b);
        prefetch_l1(b + 208);
    }
}

// Children 4k..4k+3 of a wide node, planes ordered along the ray: WideNode is lox|loy|loz|hix|hiy|hiz (8 floats
// each = two float4), so the near plane of axis a is float4 index 2a + (invD[a] < 0 ? 6 : 0) + k and the far plane
// the other one.  Slab-tests the four boxes at T0; pass[s] / tmin[s] per child.
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

__device__ __forceinline__ bool tri_fetch_test(const DeviceScene& S, int i, const RayQ& r, float tmax, float& t, float& u, float& v) {
    float4 a = __ldg(&S.tri[3ll * i + 0]);
    float4 b = __ldg(&S.tri[3ll * i + 1]);
    float4 c = __ldg(&S.tri[3ll * i + 2]);
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
        float entry;
        bool descend = false;
        if (box_pass(__ldg(&S.node_lo[node]), __ldg(&S.node_hi[node]), r, tmax, entry)) {
            int4 info = __ldg(&S.node_info[node]);
            if (info.w >= 0) {
                float lt = B2PT_INF, lu = 0.0f, lv = 0.0f; int ltri = -1;
                for (int i = info.x; i < info.y; ++i) {
                    float t, u, v;
                    if (tri_fetch_test(S, i, r, tmax, t, u, v)) {
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

// =====================================================================================================
// Warp-cooperative traversal: 8 lanes ("an octet") per ray.
//
// The per-thread routines above execute with ~4 of 32 lanes active (ncu: profiles/r01_ncu_closest_v1_*):
// every lane walks its own tree, so node steps, leaf steps, sorting and popping all diverge.  Here a ray
// is owned by 8 consecutive lanes that always do the same thing: in a wide node each lane slab-tests ONE
// child box (a coalesced 32-byte read per plane), in a reference leaf each lane runs Möller–Trumbore on
// ONE triangle (leaves hold <= 8).  Hit ordering, stack pushes and the closest-hit reduction are
// __ballot_sync / __shfl_sync operations inside the octet; the traversal stack is one shared-memory array
// per octet.  The four octets of a warp are independent rays (sub-warp masks), so divergence is only
// "octet A is in a node while octet B is in a leaf".
//
// Closest hit — candidate set: triangles whose reference leaf box passes the reference slab test at T0
// (exact: the leaf's box IS the wide child's box, tested with the reference arithmetic) and that
// Triangle::intersect accepts in [tMin, T0].  Subtrees are culled when their box fails at T0 (exact and
// monotone: a superset box passes whenever a leaf inside it passes) or when their entry distance exceeds
// the current best by more than a relative 2^-10.  The result is CERTIFIED to be the reference's answer
// when it is a miss (no candidate), or a unique minimum-t candidate whose leaf box still passes the slab
// test at T = t; everything else is re-run by closest_exact_dfs.
//
// Occlusion (renderer.hpp:274-278 asks only whether Scene::intersect returns true): before the first
// accepted triangle ray.tMax still has its initial value, so the answer is "does a triangle exist whose
// reference leaf box passes at T0 and which Triangle::intersect accepts in [tMin, T0]" — independent of
// traversal order, exact without any fallback.
// =====================================================================================================
#define B2PT_STACK 64            // entries per octet; deepest push chain is 7 per wide level, <= 9 levels
#define B2PT_STACK_PITCH 65      // +1 entry of padding: octets' stacks start in different banks

struct OctetCtx {
    unsigned gmask;      // the octet's 8 lanes within the warp
    int gl;              // lane within the octet, 0..7
    int gbase;           // first lane of the octet within the warp
    uint2* stack;        // shared-memory stack of this octet
};

__device__ __forceinline__ OctetCtx make_octet(uint2* block_stacks) {
    OctetCtx c;
    int lane = threadIdx.x & 31;
    c.gl = lane & 7;
    c.gbase = lane & ~7;
    c.gmask = 0xffu << c.gbase;
    c.stack = block_stacks + (threadIdx.x >> 3) * B2PT_STACK_PITCH;
    return c;
}

// One lane's slab test of child `gl` of a wide node.
__device__ __forceinline__ bool octet_child_test(const WideNode* nd, int gl, const RayQ& r, float& tmin, uint32_t& code) {
    float lx = __ldg(&nd->lox[gl]), ly = __ldg(&nd->loy[gl]), lz = __ldg(&nd->loz[gl]);
    float hx = __ldg(&nd->hix[gl]), hy = __ldg(&nd->hiy[gl]), hz = __ldg(&nd->hiz[gl]);
    code = __ldg(&nd->child[gl]);
    float tmax = r.T0;
    tmin = B2PT_TMIN;
    slab_axis(lx, hx, r.o.x, r.invD.x, tmin, tmax);
    slab_axis(ly, hy, r.o.y, r.invD.y, tmin, tmax);
    slab_axis(lz, hz, r.o.z, r.invD.z, tmin, tmax);
    return tmax > tmin;
}

// Closest hit, cooperative.  All 8 lanes of the octet call this with the same ray and get the same result.
// Returns the certificate (true = provably the reference's answer).
template <bool COUNT>
__device__ __forceinline__ bool closest_octet(const DeviceScene& S, const OctetCtx& g, const RayQ& r, HitRec& out,
                                              unsigned& n_nodes, unsigned& n_tris) {
    out.t = B2PT_INF; out.tri = -1; out.u = 0.0f; out.v = 0.0f;
    if (S.nwide == 0 || ray_has_nan(r)) return true;
    bool tie = false, overflow = false;
    float cull = r.T0;
    int sp = 0;
    uint32_t cur = 0;
    while (true) {
        if (!(cur & B2PT_CHILD_LEAF)) {
            if (COUNT && g.gl == 0) ++n_nodes;
            float tmin; uint32_t code;
            bool hit = octet_child_test(&S.wide[cur], g.gl, r, tmin, code) && tmin <= cull;
            unsigned hm = (__ballot_sync(g.gmask, hit) >> g.gbase) & 0xffu;
            int n = __popc(hm);
            if (n > 0) {
                // rank = number of hit children that must sit BELOW me on the stack (farther first)
                int rank = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float tj = __shfl_sync(g.gmask, tmin, g.gbase + j);
                    bool below = (tj > tmin) || (tj == tmin && j < g.gl);
                    rank += (((hm >> j) & 1u) && below) ? 1 : 0;
                }
                if (sp + n - 1 > B2PT_STACK) { overflow = true; break; }
                if (hit && rank < n - 1) g.stack[sp + rank] = make_uint2(code, __float_as_uint(tmin));
                // the nearest child (rank n-1) is visited next without going through the stack
                unsigned nm = (__ballot_sync(g.gmask, hit && rank == n - 1) >> g.gbase) & 0xffu;
                cur = __shfl_sync(g.gmask, code, g.gbase + __ffs(nm) - 1);
                sp += n - 1;
                __syncwarp(g.gmask);
                continue;
            }
        } else {
            int first = cur & 0x0FFFFFFF, cnt = ((cur >> 28) & 7) + 1;
            float t = B2PT_INF, u = 0.0f, v = 0.0f;
            bool acc = false;
            if (g.gl < cnt) {
                if (COUNT) ++n_tris;
                acc = tri_fetch_test(S, first + g.gl, r, r.T0, t, u, v);
                if (!acc) t = B2PT_INF;
            }
            unsigned am = (__ballot_sync(g.gmask, acc) >> g.gbase) & 0xffu;
            if (am) {
                // octet minimum of t
                float m = t;
                m = fminf(m, __shfl_xor_sync(g.gmask, m, 1));
                m = fminf(m, __shfl_xor_sync(g.gmask, m, 2));
                m = fminf(m, __shfl_xor_sync(g.gmask, m, 4));
                unsigned wm = (__ballot_sync(g.gmask, acc && t == m) >> g.gbase) & 0xffu;
                if (m < out.t) {
                    int w = __ffs(wm) - 1;   // first triangle of the leaf with the minimum t (bvh.hpp:88)
                    out.t = m;
                    out.tri = first + w;
                    out.u = __shfl_sync(g.gmask, u, g.gbase + w);
                    out.v = __shfl_sync(g.gmask, v, g.gbase + w);
                    tie = __popc(wm) > 1;
                    cull = fminf(r.T0, __fmaf_rn(m, 0.0009765625f, m));
                } else if (m == out.t) {
                    tie = true;
                }
            }
        }
        // pop the nearest pending subtree that can still matter
        bool got = false;
        while (sp > 0) {
            --sp;
            uint2 e = g.stack[sp];
            if (__uint_as_float(e.y) <= cull) { cur = e.x; got = true; break; }
        }
        if (!got) break;
    }
    if (overflow) return false;
    if (out.tri < 0) return true;
    if (tie) return false;
    int leaf = __float_as_int(__ldg(&S.tri[3ll * out.tri]).w);
    float entry;
    return box_pass(__ldg(&S.leaf_lo[leaf]), __ldg(&S.leaf_hi[leaf]), r, out.t, entry);
}

// Occlusion query, cooperative.  Returns 1 occluded, 0 free, -1 stack overflow (caller must use the exact path).
template <bool COUNT>
__device__ __forceinline__ int any_octet(const DeviceScene& S, const OctetCtx& g, const RayQ& r, unsigned& n_nodes, unsigned& n_tris) {
    if (S.nwide == 0 || ray_has_nan(r)) return 0;
    int sp = 0;
    uint32_t cur = 0;
    while (true) {
        if (!(cur & B2PT_CHILD_LEAF)) {
            if (COUNT && g.gl == 0) ++n_nodes;
            float tmin; uint32_t code;
            bool hit = octet_child_test(&S.wide[cur], g.gl, r, tmin, code);
            unsigned hm = (__ballot_sync(g.gmask, hit) >> g.gbase) & 0xffu;
            int n = __popc(hm);
            if (n > 0) {
                if (sp + n - 1 > B2PT_STACK) return -1;
                // leaves first: they can end the query at once (top of stack = visited next)
                bool leaf = (code & B2PT_CHILD_LEAF) != 0;
                unsigned lm = (__ballot_sync(g.gmask, hit && leaf) >> g.gbase) & 0xffu;
                unsigned below_me = (1u << g.gl) - 1u;
                int nl = __popc(lm);
                int pos = leaf ? (n - nl) + __popc(lm & below_me) : __popc((hm & ~lm) & below_me);
                if (hit && pos < n - 1) g.stack[sp + pos] = make_uint2(code, 0u);
                unsigned nm = (__ballot_sync(g.gmask, hit && pos == n - 1) >> g.gbase) & 0xffu;
                cur = __shfl_sync(g.gmask, code, g.gbase + __ffs(nm) - 1);
                sp += n - 1;
                __syncwarp(g.gmask);
                continue;
            }
        } else {
            int first = cur & 0x0FFFFFFF, cnt = ((cur >> 28) & 7) + 1;
            bool acc = false;
            if (g.gl < cnt) {
                float t, u, v;
                if (COUNT) ++n_tris;
                acc = tri_fetch_test(S, first + g.gl, r, r.T0, t, u, v);
            }
            if (__ballot_sync(g.gmask, acc) & g.gmask) return 1;
        }
        if (sp == 0) return 0;
        cur = g.stack[--sp].x;
    }
}

}  // namespace b2pt
