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

#define B2PT_RTC_STACK 72
// Explicit prefetch.global.L1 of a leaf's later cache lines and of the next stack entry: measured SLOWER (Cornell
// 744 -> 475 Msamples/s, 1M mesh 185 -> 175): the hints are LSU instructions in kernels that are issue-bound.
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
            // 8 children, SoA: 16-byte loads
            float lox[8], loy[8], loz[8], hix[8], hiy[8], hiz[8];
            uint32_t code[8];
            {
                const float4* p = reinterpret_cast<const float4*>(nd);
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    float4 a = __ldg(p + 0 + k), b = __ldg(p + 2 + k), c = __ldg(p + 4 + k);
                    float4 d = __ldg(p + 6 + k), e = __ldg(p + 8 + k), f = __ldg(p + 10 + k);
                    lox[4 * k] = a.x; lox[4 * k + 1] = a.y; lox[4 * k + 2] = a.z; lox[4 * k + 3] = a.w;
                    loy[4 * k] = b.x; loy[4 * k + 1] = b.y; loy[4 * k + 2] = b.z; loy[4 * k + 3] = b.w;
                    loz[4 * k] = c.x; loz[4 * k + 1] = c.y; loz[4 * k + 2] = c.z; loz[4 * k + 3] = c.w;
                    hix[4 * k] = d.x; hix[4 * k + 1] = d.y; hix[4 * k + 2] = d.z; hix[4 * k + 3] = d.w;
                    hiy[4 * k] = e.x; hiy[4 * k + 1] = e.y; hiy[4 * k + 2] = e.z; hiy[4 * k + 3] = e.w;
                    hiz[4 * k] = f.x; hiz[4 * k + 1] = f.y; hiz[4 * k + 2] = f.z; hiz[4 * k + 3] = f.w;
                }
                const uint4* q = reinterpret_cast<const uint4*>(nd->child);
                uint4 c0 = __ldg(q), c1 = __ldg(q + 1);
                code[0] = c0.x; code[1] = c0.y; code[2] = c0.z; code[3] = c0.w;
                code[4] = c1.x; code[5] = c1.y; code[6] = c1.z; code[7] = c1.w;
            }
            // test all 8, push hits in insertion-sorted order (farthest deepest)
            int base = sp;
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                float tmin = B2PT_TMIN, tmax = r.T0;
                slab_axis(lox[s], hix[s], r.o.x, r.invD.x, tmin, tmax);
                slab_axis(loy[s], hiy[s], r.o.y, r.invD.y, tmin, tmax);
                slab_axis(loz[s], hiz[s], r.o.z, r.invD.z, tmin, tmax);
                if (tmax > tmin && tmin <= cull) {
                    // insert so that entries in [base, sp) are sorted by decreasing entry distance
                    int j = sp++;
                    while (j > base && sent[j - 1] < tmin) { sent[j] = sent[j - 1]; scode[j] = scode[j - 1]; --j; }
                    sent[j] = tmin; scode[j] = code[s];
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
            const float4* p = reinterpret_cast<const float4*>(nd);
            const uint4* q = reinterpret_cast<const uint4*>(nd->child);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                float4 a = __ldg(p + 0 + k), b = __ldg(p + 2 + k), c = __ldg(p + 4 + k);
                float4 d = __ldg(p + 6 + k), e = __ldg(p + 8 + k), f = __ldg(p + 10 + k);
                uint4 cc = __ldg(q + k);
                const float lx[4] = {a.x, a.y, a.z, a.w}, ly[4] = {b.x, b.y, b.z, b.w}, lz[4] = {c.x, c.y, c.z, c.w};
                const float hx[4] = {d.x, d.y, d.z, d.w}, hy[4] = {e.x, e.y, e.z, e.w}, hz[4] = {f.x, f.y, f.z, f.w};
                const uint32_t cd[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    float tmin = B2PT_TMIN, tmax = r.T0;
                    slab_axis(lx[s], hx[s], r.o.x, r.invD.x, tmin, tmax);
                    slab_axis(ly[s], hy[s], r.o.y, r.invD.y, tmin, tmax);
                    slab_axis(lz[s], hz[s], r.o.z, r.invD.z, tmin, tmax);
                    if (tmax > tmin) scode[sp++] = cd[s];
                }
            }
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


}  // namespace b2pt
