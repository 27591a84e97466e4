// traverse_rtc.cuh — run-to-completion traversal, one ray per thread: the variant for COHERENT ray batches
// (camera rays, their shadow rays, any ray in a scene of a few dozen triangles).
//
// When the lanes of a warp carry similar rays they walk the tree in lock-step anyway, so the cheapest code
// wins: a tight per-thread loop with the stack in (coalesced) local memory and no refill bookkeeping.
// Exactness as in DESIGN.md §2 / traverse.cuh.
#pragma once
#include "traverse.cuh"

namespace b2pt {

// Stack entries per ray.  A node step pushes at most 8 and the next pop takes one back, so `sp <= STACK - 8` before
// a node step is all that is checked; a ray that would need more (a very deep, unbalanced tree) is handed to the
// exact reference recursion instead — never truncated.
#define B2PT_RTC_STACK 64

// ---- fast ordered traversal of the wide BVH --------------------------------------------------------
// Candidate set: triangles whose reference leaf box passes the reference slab test at T0 (exact: a leaf's box is
// its parent's child box, tested with the reference arithmetic) and that Triangle::intersect accepts in [tMin, T0].
// Subtrees are skipped when their box fails at T0 (exact, monotone) or when their entry distance exceeds the current
// best by more than the cull slack (traverse.cuh, cull_after_hit).
// Returns true when the result is certified to be the reference's answer (traverse.cuh, certify_unique):
//   * miss (no candidate), or
//   * a unique minimum-t candidate whose leaf the reference recursion is bound to enter.
template <bool COUNT>
__device__ __forceinline__ bool closest_rtc(const DeviceScene& S, const RayQ& r, HitRec& out,
                                             unsigned& n_nodes, unsigned& n_tris) {
    out.t = B2PT_INF; out.tri = -1; out.u = 0.0f; out.v = 0.0f;
    if ((S.nwide == 0 && S.nhoist == 0) || ray_has_nan(r)) return true;   // NaN ray: a miss in the reference (traverse.cuh)
    bool tie = false;
    int best_leaf = -1;
    float best = B2PT_INF;        // t of the current best candidate
    float second = B2PT_INF;      // smallest t among the other candidates seen (certify_unique)
    float cull = r.T0;            // entry distances above this cannot matter
    // Hoisted leaves (boxes covering most of the scene): tested by every ray, here, while the warp is converged.
    for (int h = 0; h < S.nhoist; ++h) {
        if (!leaf_visible(S, S.hoist_leaf[h], r, r.T0)) continue;
        const int first = S.hoist_code[h] & 0x0FFFFFFF, cnt = ((S.hoist_code[h] >> 28) & 7) + 1;
        for (int i = first; i < first + cnt; ++i) {
            float t, u, v; int leaf;
            if (COUNT) ++n_tris;
            if (tri_fetch_test(S, i, r, r.T0, t, u, v, leaf)) {
                if (t < best) {
                    second = best;
                    best = t; out.t = t; out.tri = i; out.u = u; out.v = v; tie = false; best_leaf = leaf;
                    cull = cull_after_hit(S, r, t);
                } else if (t == best) {
                    tie = true;
                } else {
                    second = fminf(second, t);
                }
            }
        }
    }
    // stack of (child code, entry distance bits): one 8-byte local-memory access per entry
    uint2 stk[B2PT_RTC_STACK];
    int sp = 0;
    uint32_t cur = 0;             // root wide node
    while (S.nwide > 0) {
        if (!(cur & B2PT_CHILD_LEAF)) {
            if (sp > B2PT_RTC_STACK - 8) return false;   // deeper than the stack: the exact recursion decides
            const WideNode* nd = &S.wide[cur];
            if (COUNT) ++n_nodes;
            // 8 children in two halves; push hits in insertion-sorted order (farthest deepest)
            int base = sp;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                Node4 n4;
                node_test4(nd, k, r, n4);
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const float tmin = n4.tmin[s];
                    if (n4.pass[s] && tmin <= cull) {
                        // insert so that entries in [base, sp) are sorted by decreasing entry distance
                        int j = sp++;
                        while (j > base) {
                            const uint2 prev = stk[j - 1];
                            if (!(__uint_as_float(prev.y) < tmin)) break;
                            stk[j] = prev; --j;
                        }
                        stk[j] = make_uint2(n4.code[s], __float_as_uint(tmin));
                    }
                }
            }
        } else {
            int first = cur & 0x0FFFFFFF, cnt = ((cur >> 28) & 7) + 1;
            auto accept = [&](int i, float t, float u, float v, int leaf) {
                if (t < best) {
                    second = best;
                    best = t; out.t = t; out.tri = i; out.u = u; out.v = v; tie = false; best_leaf = leaf;
                    cull = cull_after_hit(S, r, t);
                } else if (t == best) {
                    tie = true;
                } else {
                    second = fminf(second, t);
                }
            };
            for (int i = first; i < first + cnt; ++i) {
                float t, u, v; int leaf;
                if (COUNT) ++n_tris;
                if (tri_fetch_test(S, i, r, r.T0, t, u, v, leaf)) accept(i, t, u, v, leaf);
            }
        }
        // pop
        bool got = false;
        while (sp > 0) {
            const uint2 e = stk[--sp];
            if (__uint_as_float(e.y) <= cull) { cur = e.x; got = true; break; }
        }
        if (!got) break;
    }
    if (out.tri < 0) return true;
    if (tie) return false;
    return certify_unique(S, r, best, best_leaf, second);
}

// ---- occlusion query -------------------------------------------------------------------------------
// renderer.hpp:274-278 asks only whether Scene::intersect returns true.  Before the first accepted
// triangle ray.tMax still has its initial value, so the answer is: does any triangle exist whose
// reference leaf box passes the slab test at T0 and which Triangle::intersect accepts in [tMin, T0] —
// independent of the order in which leaves are tried.  Passing children are visited last slot first.
template <bool COUNT>
__device__ __forceinline__ bool any_rtc(const DeviceScene& S, const RayQ& r, unsigned& n_nodes, unsigned& n_tris) {
    if ((S.nwide == 0 && S.nhoist == 0) || ray_has_nan(r)) return false;
    for (int h = 0; h < S.nhoist; ++h) {   // hoisted leaves first, warp still converged
        if (!leaf_visible(S, S.hoist_leaf[h], r, r.T0)) continue;
        const int first = S.hoist_code[h] & 0x0FFFFFFF, cnt = ((S.hoist_code[h] >> 28) & 7) + 1;
        for (int i = first; i < first + cnt; ++i) {
            float t, u, v; int leaf;
            if (COUNT) ++n_tris;
            if (tri_fetch_test(S, i, r, r.T0, t, u, v, leaf)) return true;
        }
    }
    if (S.nwide == 0) return false;
    uint32_t scode[B2PT_RTC_STACK];
    int sp = 0;
    uint32_t cur = 0;
    while (true) {
        if (!(cur & B2PT_CHILD_LEAF)) {
            if (sp > B2PT_RTC_STACK - 8) { HitRec h; closest_exact_dfs(S, r, h); return h.tri >= 0; }
            const WideNode* nd = &S.wide[cur];
            if (COUNT) ++n_nodes;
            uint32_t pend = B2PT_CHILD_EMPTY;   // the last passing child so far: entered next, without a stack round trip
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                Node4 n4;
                node_test4(nd, k, r, n4);
#pragma unroll
                for (int s = 0; s < 4; ++s)
                    if (n4.pass[s]) { if (pend != B2PT_CHILD_EMPTY) scode[sp++] = pend; pend = n4.code[s]; }
            }
            if (pend != B2PT_CHILD_EMPTY) { cur = pend; continue; }
        } else {
            int first = cur & 0x0FFFFFFF, cnt = ((cur >> 28) & 7) + 1;
            for (int i = first; i < first + cnt; ++i) {
                float t, u, v; int leaf;
                if (COUNT) ++n_tris;
                if (tri_fetch_test(S, i, r, r.T0, t, u, v, leaf)) return true;
            }
        }
        if (sp == 0) return false;
        cur = scode[--sp];
    }
}

// ---- occlusion query with visiting statistics --------------------------------------------------------
// The same query as any_rtc (same candidate set, same answer), instrumented for the occluder-aware child order
// (build.cu learn_child_order): it counts how often each (wide node, slot) passes the box test of a shadow ray
// (`visits`) and in which slot's leaf a ray was found occluded (`hits`).  Run for ONE wavefront batch per scene and
// only by the warps that sample (`on`); the production kernel is any_rtc, untouched by this.
__device__ __forceinline__ bool any_rtc_learn(const DeviceScene& S, const RayQ& r, bool on, unsigned* __restrict__ visits,
                                              unsigned* __restrict__ hits) {
    if ((S.nwide == 0 && S.nhoist == 0) || ray_has_nan(r)) return false;
    for (int h = 0; h < S.nhoist; ++h) {
        if (!leaf_visible(S, S.hoist_leaf[h], r, r.T0)) continue;
        const int first = S.hoist_code[h] & 0x0FFFFFFF, cnt = ((S.hoist_code[h] >> 28) & 7) + 1;
        for (int i = first; i < first + cnt; ++i) {
            float t, u, v; int leaf;
            if (tri_fetch_test(S, i, r, r.T0, t, u, v, leaf)) return true;
        }
    }
    if (S.nwide == 0) return false;
    uint32_t scode[B2PT_RTC_STACK];
    uint32_t sslot[B2PT_RTC_STACK];   // (wide node << 3 | slot) each entry came from
    int sp = 0;
    uint32_t cur = 0, cur_slot = 0;
    while (true) {
        if (!(cur & B2PT_CHILD_LEAF)) {
            if (sp > B2PT_RTC_STACK - 8) { HitRec h; closest_exact_dfs(S, r, h); return h.tri >= 0; }
            const WideNode* nd = &S.wide[cur];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                Node4 n4;
                node_test4(nd, k, r, n4);
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    if (n4.pass[s]) {
                        const uint32_t slot = (cur << 3) | (uint32_t)(4 * k + s);
                        if (on) atomicAdd(&visits[slot], 1u);
                        scode[sp] = n4.code[s]; sslot[sp] = slot; ++sp;
                    }
                }
            }
        } else {
            int first = cur & 0x0FFFFFFF, cnt = ((cur >> 28) & 7) + 1;
            for (int i = first; i < first + cnt; ++i) {
                float t, u, v; int leaf;
                if (tri_fetch_test(S, i, r, r.T0, t, u, v, leaf)) {
                    if (on) atomicAdd(&hits[cur_slot], 1u);
                    return true;
                }
            }
        }
        if (sp == 0) return false;
        --sp;
        cur = scode[sp]; cur_slot = sslot[sp];
    }
}

}  // namespace b2pt
