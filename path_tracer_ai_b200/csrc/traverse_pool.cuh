// traverse_pool.cuh — traversal for INCOHERENT ray batches: persistent warps, two rays per lane, phase-split steps.
//
// What bounds every one-ray-per-lane traversal of incoherent rays is SIMT divergence, not arithmetic: ncu on the
// run-to-completion kernels shows 5.5-7.7 of 32 lanes active per instruction on bounce rays in a 1M-triangle scene
// (profiles/r02_ncu_c3_rtc_v1.txt).  A lane's ray is either expanding a wide node (8 slab tests) or testing a leaf's
// triangles (Möller–Trumbore), the two bodies cannot overlap inside a warp, and a lane whose ray has finished idles
// until the slowest ray of its warp is done.
//
// Here every lane owns B2PT_PR rays.  One iteration of the warp runs
//     a NODE phase   each lane advances ONE of its rays that is waiting on a wide node,
//     a TRI phase    each lane advances ONE of its rays that is inside a leaf by up to B2PT_PTPS triangles,
//     a FINISH phase results of finished rays are written out,
// and empty ray slots are refilled from a warp-local pool of work indices once enough of them have accumulated.  A
// phase that too few lanes want is skipped for an iteration (its rays wait while the lane's other ray moves on).  With
// two rays per lane the probability that a lane has something to do in a phase rises from p to 1 - (1 - p)^2, and
// nothing waits for a slow neighbour.
//
// Everything inside a phase is written to be the SAME instruction stream for every lane: the 8 children of a node are
// ordered by a fixed 19-comparator sorting network on 32-bit keys (entry-distance bits with the slot number in the
// three lowest mantissa bits), pushed with predicated stores, and a stack entry is (key, node) — the child code is
// read from the node when the entry is popped.  The ray's constants stay in registers (selected by the ray number),
// its small mutable state lives in shared memory indexed by the lane (no bank conflicts, no cross-lane traffic, no
// synchronisation) and its stack in local memory.  The leaves hoisted out of the tree (ctx.cuh) are pushed on a new
// ray's stack like any other leaf, so their triangles are tested in ordinary TRI phases.
//
// Exactness is that of traverse.cuh / DESIGN.md §2: same candidate set, same certificate, same exact fallback.  Dropping
// the three low bits of an entry distance only makes a key SMALLER, i.e. the distance cull more permissive.
#pragma once
#include "traverse_thread.cuh"

namespace b2pt {

#define B2PT_PR 2            // rays per lane
#define B2PT_PBLOCK 128      // threads per block
#define B2PT_PSTACK 40       // stack entries per ray (local memory); a ray that needs more goes to the exact recursion
#ifndef B2PT_PTPS
#define B2PT_PTPS 4          // triangles per TRI phase
#endif
#define B2PT_PREFILL 8       // empty ray slots in the warp that trigger a refill
#define B2PT_PCHUNK 256      // work indices a warp claims per global atomic
#ifndef B2PT_PMINB
#define B2PT_PMINB 6         // resident blocks per SM the kernels are compiled for (register cap)
#endif
#ifndef B2PT_PVOTE
#define B2PT_PVOTE 12
#endif
// PVOTE: a phase wanted by fewer lanes than this waits, unless it is the busier of the two

enum { PS_EMPTY = 0, PS_NODE = 1, PS_TRI = 2, PS_DONE = 3 };
// mutable per-ray state; CUR = wide node to expand (PS_NODE) / next triangle of the current leaf (PS_TRI), TEND = its end
enum { PF_CUR, PF_TEND, PF_SP, PF_IDX, PF_ANY_COUNT,
       PF_CULL = PF_ANY_COUNT, PF_BT, PF_BU, PF_BV, PF_BTRI, PF_BLEAF, PF_TIE, PF_SECOND, PF_CLOSEST_COUNT };

template <bool ANY>
struct PoolSmem {
    static constexpr int NF = ANY ? PF_ANY_COUNT : PF_CLOSEST_COUNT;
    float f[NF][B2PT_PR][B2PT_PBLOCK];
    uint4 codes[2][B2PT_PBLOCK];   // the 8 child codes of the node a lane is expanding: looked up by (dynamic) slot number
};

#ifndef B2PT_PPREFETCH
#define B2PT_PPREFETCH 0
#endif
#ifndef B2PT_PDEFER
#define B2PT_PDEFER 0      // 1: both phases pick their ray from the states at the start of the iteration
#endif
#ifndef B2PT_PLUT
#define B2PT_PLUT 0        // 1: stack entries carry child codes (looked up in a shared-memory copy of the node's 8 codes); 0: (key, node), code loaded at pop
#endif
__device__ __forceinline__ void pool_prefetch(const void* p) {
#if B2PT_PPREFETCH
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#endif
}

#define B2PT_PKEY_NONE 0xffffffffu
__device__ __forceinline__ void pool_cswap(unsigned& a, unsigned& b) { const unsigned lo = min(a, b), hi = max(a, b); a = lo; b = hi; }
// ascending sort of 8 keys (optimal 19-comparator network)
__device__ __forceinline__ void pool_sort8(unsigned (&k)[8]) {
    pool_cswap(k[0], k[1]); pool_cswap(k[2], k[3]); pool_cswap(k[4], k[5]); pool_cswap(k[6], k[7]);
    pool_cswap(k[0], k[2]); pool_cswap(k[1], k[3]); pool_cswap(k[4], k[6]); pool_cswap(k[5], k[7]);
    pool_cswap(k[1], k[2]); pool_cswap(k[5], k[6]); pool_cswap(k[0], k[4]); pool_cswap(k[3], k[7]);
    pool_cswap(k[1], k[5]); pool_cswap(k[2], k[6]);
    pool_cswap(k[1], k[4]); pool_cswap(k[3], k[6]);
    pool_cswap(k[2], k[4]); pool_cswap(k[3], k[5]);
    pool_cswap(k[3], k[4]);
}

// The planes of the 8 children of a wide node, ordered along the ray (node_test4 without the child codes).
__device__ __forceinline__ void pool_node_test8(const WideNode* nd, V3 o, V3 invD, float T0, float (&tmin)[8], unsigned& passmask, uint4& cc0, uint4& cc1) {
    const float4* p = reinterpret_cast<const float4*>(nd);
#if B2PT_PLUT
    cc0 = __ldg(reinterpret_cast<const uint4*>(nd->child));
    cc1 = __ldg(reinterpret_cast<const uint4*>(nd->child) + 1);
#endif
    const int nx = invD.x < 0.0f ? 6 : 0, ny = invD.y < 0.0f ? 6 : 0, nz = invD.z < 0.0f ? 6 : 0;
    passmask = 0u;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float4 a = __ldg(p + nx + k), b = __ldg(p + 2 + ny + k), c = __ldg(p + 4 + nz + k);
        const float4 d = __ldg(p + 6 - nx + k), e = __ldg(p + 8 - ny + k), f = __ldg(p + 10 - nz + k);
        const float nxs[4] = {a.x, a.y, a.z, a.w}, nys[4] = {b.x, b.y, b.z, b.w}, nzs[4] = {c.x, c.y, c.z, c.w};
        const float fxs[4] = {d.x, d.y, d.z, d.w}, fys[4] = {e.x, e.y, e.z, e.w}, fzs[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            float t0 = B2PT_TMIN, t1 = T0;
            slab_axis_nf(nxs[s], fxs[s], o.x, invD.x, t0, t1);
            slab_axis_nf(nys[s], fys[s], o.y, invD.y, t0, t1);
            slab_axis_nf(nzs[s], fzs[s], o.z, invD.z, t0, t1);
            tmin[4 * k + s] = t0;
            if (t1 > t0) passmask |= 1u << (4 * k + s);
        }
    }
}

// One warp's share of a batch: claims work indices, runs the phases until the batch is exhausted.
//   IO::load(S, k, o, d, T0, tag) -> work item k as a ray (d already normalised as the Ray ctor does) and the int the
//                                    result is stored under; false if the item needs no traversal
//   IO::store_closest(tag, HitRec, certified) / IO::store_any(tag, occluded)
template <bool ANY, bool COUNT, class IO>
__device__ __forceinline__ void pool_traverse(const DeviceScene& S, PoolSmem<ANY>& sm, IO& io, unsigned long long* __restrict__ work_counter,
                                              long long total, TraceCounters* __restrict__ tc) {
    // closest: (key, node) — key = entry-distance bits with the child slot in bits 0..2; any-hit: (node, mask of children still to visit)
    const int tid = threadIdx.x;
    uint2 stk[B2PT_PR][B2PT_PSTACK];
    int st0 = PS_EMPTY, st1 = PS_EMPTY;
    // ray constants: origin, direction, 1/direction, T0
    float c0[10], c1[10];
    WarpPool pool{0, 0, false};
    unsigned n_nodes = 0, n_tris = 0;
    int flip = 0;

#define PSF(F, r) sm.f[F][r][tid]
#define PSI(F, r) __float_as_int(sm.f[F][r][tid])
#define PSET(F, r, v) sm.f[F][r][tid] = __int_as_float(v)
#define PC(i, r) ((r) ? c1[i] : c0[i])
    auto get_state = [&](int r) { return r == 0 ? st0 : st1; };
    auto set_state = [&](int r, int v) { if (r == 0) st0 = v; else st1 = v; };
    // child `code` becomes the ray's current work item
    // (its data is requested now: the phase that consumes it runs an iteration later at the earliest)
    auto enter = [&](int r, uint32_t code) {
        if (code & B2PT_CHILD_LEAF) {
            const int first = (int)(code & 0x0FFFFFFF), cnt = (int)((code >> 28) & 7) + 1;
            PSET(PF_CUR, r, first); PSET(PF_TEND, r, first + cnt);
            set_state(r, PS_TRI);
            const char* b = reinterpret_cast<const char*>(S.tri + 3ll * first);
            pool_prefetch(b); pool_prefetch(b + 128);
            if (cnt > B2PT_PTPS) pool_prefetch(b + cnt * 48 - 16);
        } else {
            PSET(PF_CUR, r, (int)code);
            set_state(r, PS_NODE);
            const char* b = reinterpret_cast<const char*>(S.wide + code);
            pool_prefetch(b); pool_prefetch(b + 208);
        }
    };
    // Stack entries — closest: (key, child code); any-hit: (node, mask of its children still to visit), or (code, 0) for
    // a code to enter directly (the root, the hoisted leaves).
    // next stack entry that can still matter -> the ray's new state
    auto pop = [&](int r, int sp) {
        bool found = false;
        uint32_t code = 0u;
        if constexpr (ANY) {
            if (sp > 0) {
                uint2 e = stk[r][sp - 1];
                if (e.y == 0u) {
                    code = e.x;
                    --sp;
                } else {   // (any-hit entries always name the node: one entry covers all its passing children)
                    const int slot = 31 - __clz(e.y);          // last slot first (the learned occlusion order, build.cu)
                    e.y &= ~(1u << slot);
                    if (e.y) stk[r][sp - 1].y = e.y; else --sp;
                    code = __ldg(&S.wide[e.x].child[slot]);
                }
                found = true;
            }
        } else {
            const float cull = PSF(PF_CULL, r);
            while (sp > 0) {
                const uint2 e = stk[r][--sp];
                if (!(__uint_as_float(e.x & ~7u) <= cull)) continue;
#if B2PT_PLUT
                code = e.y;
#else
                code = (e.y & 0xC0000000u) ? ((e.y & 0x80000000u) ? e.y : (e.y & 0x3FFFFFFFu)) : __ldg(&S.wide[e.y].child[e.x & 7u]);
#endif
                found = true;
                break;
            }
        }
        PSET(PF_SP, r, sp);
        if (found) enter(r, code); else set_state(r, PS_DONE);
    };

    while (true) {
        // ---- refill -----------------------------------------------------------------------------------------
        const unsigned e0 = __ballot_sync(0xffffffffu, st0 == PS_EMPTY), e1 = __ballot_sync(0xffffffffu, st1 == PS_EMPTY);
        const bool all_idle = (e0 & e1) == 0xffffffffu;
        if (!pool.exhausted && (__popc(e0) + __popc(e1) >= B2PT_PREFILL || all_idle)) {
#pragma unroll
            for (int r = 0; r < B2PT_PR; ++r) {
                const bool want = get_state(r) == PS_EMPTY;
                long long k = warp_pool_take<B2PT_PCHUNK>(pool, work_counter, total, want);
                if (want && k >= 0) {
                    V3 o, d; float T0; int tag;
                    if (io.load(S, k, o, d, T0, tag)) {
                        RayQ q = make_rayq_normalised(o, d, T0);
                        float* c = r ? c1 : c0;
                        c[0] = q.o.x; c[1] = q.o.y; c[2] = q.o.z; c[3] = q.d.x; c[4] = q.d.y; c[5] = q.d.z;
                        c[6] = q.invD.x; c[7] = q.invD.y; c[8] = q.invD.z; c[9] = T0;
                        PSET(PF_IDX, r, tag);
                        if constexpr (!ANY) {
                            PSF(PF_CULL, r) = T0; PSF(PF_BT, r) = B2PT_INF; PSF(PF_BU, r) = 0.0f; PSF(PF_BV, r) = 0.0f;
                            PSET(PF_BTRI, r, -1); PSET(PF_BLEAF, r, -1); PSET(PF_TIE, r, 0); PSF(PF_SECOND, r) = B2PT_INF;
                        }
                        int sp = 0;
                        if (!ray_has_nan(q)) {   // a NaN ray is a miss in the reference (traverse.cuh): nothing to traverse
                            // the root, and on top of it the visible hoisted leaves (popped first: an early closest hit culls the tree)
                            // closest: (key 0, code) — without B2PT_PLUT the root is marked by bit 30; any-hit: (code, 0) = direct
                            if (S.nwide > 0) { if constexpr (ANY) stk[r][sp++] = make_uint2(0u, 0u); else stk[r][sp++] = make_uint2(0u, B2PT_PLUT ? 0u : 0x40000000u); }
                            for (int h = 0; h < S.nhoist; ++h)
                                if (leaf_visible(S, S.hoist_leaf[h], q, T0)) {
                                    if constexpr (ANY) stk[r][sp++] = make_uint2(S.hoist_code[h], 0u); else stk[r][sp++] = make_uint2(0u, S.hoist_code[h]);
                                }
                        }
                        pop(r, sp);
                    }
                }
            }
        }
        const unsigned wantN = __ballot_sync(0xffffffffu, st0 == PS_NODE || st1 == PS_NODE);
        const unsigned wantT = __ballot_sync(0xffffffffu, st0 == PS_TRI || st1 == PS_TRI);
        const unsigned busy = __ballot_sync(0xffffffffu, st0 != PS_EMPTY || st1 != PS_EMPTY);
        if (busy == 0) {
            if (pool.exhausted) break;
            continue;
        }
        const int cN = __popc(wantN), cT = __popc(wantT);
        const bool runN = cN > 0 && (cN >= B2PT_PVOTE || cN >= cT);
        const bool runT = cT > 0 && (cT >= B2PT_PVOTE || cT > cN);
        // Both phases choose their ray from the states as they are NOW: a ray that changes state in the NODE phase is
        // not touched again in this iteration, so the data its next step needs (requested by enter()) has an
        // iteration's time to arrive.  The preferred ray alternates.
        flip ^= 1;
        const int pa = flip, pb = flip ^ 1;
        const int rN = get_state(pa) == PS_NODE ? pa : (get_state(pb) == PS_NODE ? pb : -1);
        const int rT = get_state(pa) == PS_TRI ? pa : (get_state(pb) == PS_TRI ? pb : -1);

        // ---- NODE phase: expand one wide node -------------------------------------------------------------------
        if (runN) {
            const int r = rN;
            if (r >= 0) {
                int sp = PSI(PF_SP, r);
                if (sp > B2PT_PSTACK - 8) {
                    // deeper than the stack: the exact recursion decides (closest: via the fallback list)
                    if constexpr (ANY) {
                        RayQ q;
                        q.o = mk3(PC(0, r), PC(1, r), PC(2, r)); q.d = mk3(PC(3, r), PC(4, r), PC(5, r));
                        q.invD = mk3(PC(6, r), PC(7, r), PC(8, r)); q.T0 = PC(9, r);
                        HitRec h; closest_exact_dfs(S, q, h);
                        io.store_any(PSI(PF_IDX, r), h.tri >= 0);
                        set_state(r, PS_EMPTY);
                    } else {
                        PSET(PF_TIE, r, 2);
                        set_state(r, PS_DONE);
                    }
                } else {
                    const V3 o = mk3(PC(0, r), PC(1, r), PC(2, r)), invD = mk3(PC(6, r), PC(7, r), PC(8, r));
                    const float T0 = PC(9, r);
                    const uint32_t node = (uint32_t)PSI(PF_CUR, r);
                    const WideNode* nd = &S.wide[node];
                    if (COUNT) ++n_nodes;
                    float tmin[8];
                    unsigned passmask;
                    uint4 cc0, cc1;
                    pool_node_test8(nd, o, invD, T0, tmin, passmask, cc0, cc1);
#if B2PT_PLUT
                    sm.codes[0][tid] = cc0; sm.codes[1][tid] = cc1;   // read back below by slot number (same thread: program order)
                    const uint32_t* lut = reinterpret_cast<const uint32_t*>(&sm.codes[0][tid]);
                    auto code_of = [&](unsigned slot) { return lut[(slot >> 2) * (B2PT_PBLOCK * 4) + (slot & 3u)]; };
#else
                    auto code_of = [&](unsigned slot) { return __ldg(&nd->child[slot]); };
#endif
                    if constexpr (ANY) {
                        if (passmask) {
                            // one entry for all passing children; the last slot is visited at once
                            const int slot = 31 - __clz(passmask);
                            const unsigned rest = passmask & ~(1u << slot);
                            if (rest) stk[r][sp++] = make_uint2(node, rest);
                            PSET(PF_SP, r, sp);
                            enter(r, code_of((unsigned)slot));
                        } else {
                            pop(r, sp);
                        }
                    } else {
                        const float cull = PSF(PF_CULL, r);
                        unsigned key[8];
#pragma unroll
                        for (int s = 0; s < 8; ++s)
                            key[s] = (((passmask >> s) & 1u) && tmin[s] <= cull) ? ((__float_as_uint(tmin[s]) & ~7u) | (unsigned)s) : B2PT_PKEY_NONE;
                        pool_sort8(key);
                        int n = 0;
#pragma unroll
                        for (int s = 0; s < 8; ++s) n += key[s] != B2PT_PKEY_NONE ? 1 : 0;
                        if (n > 0) {
                            // farthest first, so that the nearest pending child is on top; the nearest of all is entered at once
#pragma unroll
                            for (int j = 7; j >= 1; --j)
                                if (j < n) stk[r][sp + (n - 1 - j)] = make_uint2(key[j], B2PT_PLUT ? code_of(key[j] & 7u) : node);
                            sp += n - 1;
                            PSET(PF_SP, r, sp);
                            enter(r, code_of(key[0] & 7u));
                        } else {
                            pop(r, sp);
                        }
                    }
                }
            }
        }

        // ---- TRI phase: up to B2PT_PTPS triangles of the current leaf ----------------------------------------------
        if (runT) {
#if B2PT_PDEFER
            const int r = rT;
#else
            const int r = get_state(pa) == PS_TRI ? pa : (get_state(pb) == PS_TRI ? pb : -1);
            (void)rT;
#endif
            if (r >= 0) {
                RayQ q;
                q.o = mk3(PC(0, r), PC(1, r), PC(2, r)); q.d = mk3(PC(3, r), PC(4, r), PC(5, r));
                q.T0 = PC(9, r);
                int i = PSI(PF_CUR, r);
                const int end = PSI(PF_TEND, r);
                bool occluded = false;
                float best = 0.0f;
                if constexpr (!ANY) best = PSF(PF_BT, r);
#pragma unroll
                for (int k = 0; k < B2PT_PTPS; ++k) {
                    if (i < end && !occluded) {
                        float t, u, v; int leaf;
                        if (COUNT) ++n_tris;
                        const int tri = i++;
                        if (tri_fetch_test(S, tri, q, q.T0, t, u, v, leaf)) {
                            if constexpr (ANY) {
                                occluded = true;
                            } else if (t > best) {
                                PSF(PF_SECOND, r) = fminf(PSF(PF_SECOND, r), t);
                            } else {
                                if (t < best) {
                                    PSF(PF_SECOND, r) = best;
                                    best = t;
                                    PSF(PF_BT, r) = t; PSF(PF_BU, r) = u; PSF(PF_BV, r) = v; PSET(PF_BTRI, r, tri); PSET(PF_BLEAF, r, leaf);
                                    if (PSI(PF_TIE, r) == 1) PSET(PF_TIE, r, 0);
                                    PSF(PF_CULL, r) = cull_after_hit(S, q, t);
                                } else if (PSI(PF_TIE, r) == 0) {
                                    PSET(PF_TIE, r, 1);
                                }
                            }
                        }
                    }
                }
                if (occluded) {
                    io.store_any(PSI(PF_IDX, r), true);
                    set_state(r, PS_EMPTY);
                } else if (i < end) {
                    PSET(PF_CUR, r, i);
                } else {
                    pop(r, PSI(PF_SP, r));
                }
            }
        }

        // ---- FINISH phase --------------------------------------------------------------------------------------------
        if (__ballot_sync(0xffffffffu, st0 == PS_DONE || st1 == PS_DONE)) {
#pragma unroll
            for (int r = 0; r < B2PT_PR; ++r) {
                if (get_state(r) == PS_DONE) {
                    if constexpr (ANY) {
                        io.store_any(PSI(PF_IDX, r), false);
                    } else {
                        HitRec h;
                        h.t = PSF(PF_BT, r); h.tri = PSI(PF_BTRI, r); h.u = PSF(PF_BU, r); h.v = PSF(PF_BV, r);
                        bool certified = PSI(PF_TIE, r) == 0;
                        if (certified && h.tri >= 0) {   // the reference is bound to enter the winner's leaf (traverse.cuh, certify_unique)
                            RayQ q;
                            q.o = mk3(PC(0, r), PC(1, r), PC(2, r)); q.invD = mk3(PC(6, r), PC(7, r), PC(8, r)); q.T0 = PC(9, r);
                            certified = certify_unique(S, q, h.t, PSI(PF_BLEAF, r), PSF(PF_SECOND, r));
                        }
                        io.store_closest(PSI(PF_IDX, r), h, certified);
                    }
                    set_state(r, PS_EMPTY);
                }
            }
        }
    }
#undef PSF
#undef PSI
#undef PSET
#undef PC
    if (COUNT) {
        for (int off = 16; off > 0; off >>= 1) {
            n_nodes += __shfl_down_sync(0xffffffffu, n_nodes, off);
            n_tris += __shfl_down_sync(0xffffffffu, n_tris, off);
        }
        if ((threadIdx.x & 31) == 0) { atomicAdd(&tc->node_fetches, (unsigned long long)n_nodes); atomicAdd(&tc->tri_fetches, (unsigned long long)n_tris); }
    }
}

}  // namespace b2pt
