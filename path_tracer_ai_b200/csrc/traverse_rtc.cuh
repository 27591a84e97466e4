// traverse_rtc.cuh — run-to-completion traversal, one ray per thread: the variant for COHERENT ray batches
// (camera rays, their shadow rays, any ray in a scene of a few dozen triangles).
//
// When the lanes of a warp carry similar rays they walk the tree in lock-step anyway, so the cheapest code
// wins: a tight per-thread loop with the stack in (coalesced) local memory and no refill bookkeeping.  On the
// 50-triangle Cornell frame this is 2.5x faster than the persistent kernels of traverse_thread.cuh, which in
// turn are 2x faster on incoherent rays in a 1M-triangle scene (profiles/r01_*).  Exactness as in DESIGN.md §2.
#pragma once
#include "traverse.cuh"

namespace b2pt {

#define B2PT_RTC_STACK 96   // 7 pending siblings per wide level; <= 11 levels below 2^28 triangles even with misaligned subtrees
// Explicit prefetch.global.L1 of a leaf's later cache lines and of the next stack entry: measured SLOWER (Cornell
// 744 -> 475 Msamples/s, 1M mesh 185 -> 175): the hints are LSU instructions in kernels that are issue-bound.
// Order in which an occlusion query visits the passing children of a node: 0 = last slot first, 1 = first slot
// first, 2 = nearest child first (see profiles/r01_experiments.md).
static_assert(7 * B2PT_MAX_WIDE_LEVEL + 8 <= B2PT_RTC_STACK, "rtc stack must cover the deepest tree build_scene accepts");
#ifndef B2PT_ANY_ORDER
#define B2PT_ANY_ORDER 0
#endif
#ifndef B2PT_PREFETCH
#define B2PT_PREFETCH 0
#endif

// ---- fast ordered traversal of the wide BVH --------------------------------------------------------
// Candidate set: triangles whose reference leaf box passes the reference slab test at T0 (exact —
// the leaf's box is the wide child's box, tested with the reference arithmetic) and that
// Triangle::intersect accepts in [tMin, T0].  Subtrees are culled when their box fails at T0 (exact,
// monotone) or when their entry distance exceeds the current best by more than a relative 2^-10.
// Returns true when the result is certified to be the reference's answer:
//   * miss (no candidate), or
//   * a unique minimum-t candidate whose leaf box still passes the slab test at T = t.
template <bool COUNT>
__device__ __forceinline__ bool closest_rtc(const DeviceScene& S, const RayQ& r, HitRec& out,
                                             unsigned& n_nodes, unsigned& n_tris) {
    out.t = B2PT_INF; out.tri = -1; out.u = 0.0f; out.v = 0.0f;
    if (S.nwide == 0 || ray_has_nan(r)) return true;   // NaN ray: a miss in the reference (traverse.cuh)
    bool tie = false;
    float best = B2PT_INF;        // t of the current best candidate
    float cull = r.T0;            // entry distances above this cannot matter
    // stack of (child code, entry)
    uint32_t scode[B2PT_RTC_STACK];
    float sent[B2PT_RTC_STACK];
    int sp = 0;
    uint32_t cur = 0;             // root wide node
    while (true) {
        if (!(cur & B2PT_CHILD_LEAF)) {
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
                        while (j > base && sent[j - 1] < tmin) { sent[j] = sent[j - 1]; scode[j] = scode[j - 1]; --j; }
                        sent[j] = tmin; scode[j] = n4.code[s];
                    }
                }
            }
        } else {
            // reference leaf: its box passed at T0, so its triangles are candidates
            int first = cur & 0x0FFFFFFF, cnt = ((cur >> 28) & 7) + 1;
            if (B2PT_PREFETCH) prefetch_leaf_rest(S, first, cnt);
            for (int i = first; i < first + cnt; ++i) {
                float t, u, v;
                if (COUNT) ++n_tris;
                if (tri_fetch_test(S, i, r, r.T0, t, u, v)) {
                    if (t < best) {
                        best = t; out.t = t; out.tri = i; out.u = u; out.v = v; tie = false;
                        cull = fminf(r.T0, __fmaf_rn(t, 0.0009765625f, t));
                    } else if (t == best) {
                        tie = true;
                    }
                }
            }
        }
        // pop
        bool got = false;
        while (sp > 0) {
            --sp;
            if (sent[sp] <= cull) { cur = scode[sp]; got = true; break; }
        }
        if (!got) break;
        if (B2PT_PREFETCH && sp > 0) prefetch_child(S, scode[sp - 1]);   // what the next pop will need
    }
    if (out.tri < 0) return true;
    if (tie) return false;
    // certify: the winner's reference leaf must still be visible with ray.tMax == t
    int leaf = __float_as_int(__ldg(&S.tri[3ll * out.tri]).w);
    float entry;
    return box_pass(__ldg(&S.leaf_lo[leaf]), __ldg(&S.leaf_hi[leaf]), r, out.t, entry);
}

// ---- occlusion query -------------------------------------------------------------------------------
// renderer.hpp:274-278 asks only whether Scene::intersect returns true.  Before the first accepted
// triangle ray.tMax still has its initial value, so the answer is: does any triangle exist whose
// reference leaf box passes the slab test at T0 and which Triangle::intersect accepts in [tMin, T0].
template <bool COUNT>
__device__ __forceinline__ bool any_rtc(const DeviceScene& S, const RayQ& r, unsigned& n_nodes, unsigned& n_tris) {
    if (S.nwide == 0 || ray_has_nan(r)) return false;
    uint32_t scode[B2PT_RTC_STACK];
    int sp = 0;
    uint32_t cur = 0;
    while (true) {
        if (!(cur & B2PT_CHILD_LEAF)) {
            const WideNode* nd = &S.wide[cur];
            if (COUNT) ++n_nodes;
#if B2PT_ANY_ORDER == 2
            // nearest child next, the others in slot order
            float near_t = B2PT_INF;
            uint32_t near_code = B2PT_CHILD_EMPTY;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                Node4 n4;
                node_test4(nd, k, r, n4);
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    if (n4.pass[s]) {
                        if (n4.tmin[s] < near_t) {
                            if (near_code != B2PT_CHILD_EMPTY) scode[sp++] = near_code;
                            near_t = n4.tmin[s]; near_code = n4.code[s];
                        } else {
                            scode[sp++] = n4.code[s];
                        }
                    }
                }
            }
            if (near_code != B2PT_CHILD_EMPTY) { cur = near_code; continue; }
#else
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                Node4 n4;
                node_test4(nd, B2PT_ANY_ORDER == 1 ? 1 - k : k, r, n4);
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const int ss = B2PT_ANY_ORDER == 1 ? 3 - s : s;
                    if (n4.pass[ss]) scode[sp++] = n4.code[ss];
                }
            }
#endif
        } else {
            int first = cur & 0x0FFFFFFF, cnt = ((cur >> 28) & 7) + 1;
            if (B2PT_PREFETCH) prefetch_leaf_rest(S, first, cnt);
            for (int i = first; i < first + cnt; ++i) {
                float t, u, v;
                if (COUNT) ++n_tris;
                if (tri_fetch_test(S, i, r, r.T0, t, u, v)) return true;
            }
        }
        if (sp == 0) return false;
        cur = scode[--sp];
        if (B2PT_PREFETCH && sp > 0) prefetch_child(S, scode[sp - 1]);
    }
}


// ---- occlusion query with visiting statistics --------------------------------------------------------
// The same query as any_rtc (same candidate set, same answer), instrumented for the occluder-aware child order
// (build.cu learn_child_order): it counts how often each (wide node, slot) passes the box test of a shadow ray
// (`visits`) and in which slot's leaf a ray was found occluded (`hits`).  Run for ONE wavefront batch per scene and
// only by the warps that sample (`on`); the production kernel is any_rtc, untouched by this.
__device__ __forceinline__ bool any_rtc_learn(const DeviceScene& S, const RayQ& r, bool on, unsigned* __restrict__ visits,
                                              unsigned* __restrict__ hits) {
    if (S.nwide == 0 || ray_has_nan(r)) return false;
    uint32_t scode[B2PT_RTC_STACK];
    uint32_t sslot[B2PT_RTC_STACK];   // (wide node << 3 | slot) each entry came from
    int sp = 0;
    uint32_t cur = 0, cur_slot = 0;
    while (true) {
        if (!(cur & B2PT_CHILD_LEAF)) {
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
                float t, u, v;
                if (tri_fetch_test(S, i, r, r.T0, t, u, v)) {
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
