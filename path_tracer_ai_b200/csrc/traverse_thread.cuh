// traverse_thread.cuh — one-ray-per-lane traversal of the wide BVH as a RESUMABLE state machine, for
// persistent kernels that refill finished lanes with new rays.
//
// ncu on the first (run-to-completion, one thread per ray) kernel showed 4.3 of 32 lanes active per
// instruction (profiles/r01_ncu_closest_v1_perthread.txt).  Two causes multiply: lanes whose ray has
// finished idle until the slowest ray of the warp ends (no refill), and a lane in a leaf (8 triangle
// tests) serialises against lanes in a node (8 box tests).  Here a lane advances its ray by ONE bounded
// step per loop iteration — one wide node, or up to TPS triangles of the current leaf — and the kernel
// loop refills idle lanes from a warp-local pool of ray indices between steps.
//
// Exactness is that of DESIGN.md §2 / traverse.cuh: same candidate set, same certificate as the run-to-completion variant.
#pragma once
#include "traverse.cuh"

namespace b2pt {

#define B2PT_TSTACK 64   // deeper pushes set `overflow`: the ray is then answered by the exact reference recursion
#define B2PT_SSTACK 16   // entries of each lane's stack that live in shared memory (the rest: local memory)
#define B2PT_TBLOCK 128  // block size of the per-lane kernels (stride of the shared stack layout)

// Per-lane stack: the first B2PT_SSTACK entries are in shared memory, laid out entry-major
// (entry e of thread t at smem[e * B2PT_TBLOCK + t]): the 8-byte accesses of a warp hit distinct banks
// whatever depth each lane is at.  ncu on the all-local version: 4.2 KB of local-memory traffic per ray,
// long-scoreboard the top stall, L1 at 76 % (profiles/r01_ncu_closest_v3_persistent.txt).  Deeper
// entries (rare) overflow into a small local array.
struct LaneStack {
    uint2* sm;                                    // &smem[threadIdx.x]
    uint2 ovf[B2PT_TSTACK - B2PT_SSTACK];
    __device__ __forceinline__ uint2 get(int i) const { return i < B2PT_SSTACK ? sm[i * B2PT_TBLOCK] : ovf[i - B2PT_SSTACK]; }
    __device__ __forceinline__ void set(int i, uint2 v) { if (i < B2PT_SSTACK) sm[i * B2PT_TBLOCK] = v; else ovf[i - B2PT_SSTACK] = v; }
};
#define B2PT_LANE_SMEM_BYTES (B2PT_SSTACK * B2PT_TBLOCK * 8)

struct LaneState {
    RayQ r;
    HitRec best;
    int best_leaf;
    float second;          // smallest t among the other candidates seen (certify_unique)
    float cull;
    int sp;
    uint32_t cur;          // inner wide node to expand (valid when tri_next == tri_end)
    int tri_next, tri_end; // triangles of the current leaf still to test
    bool tie;
    bool overflow;
    bool dead;             // NaN ray: finished before it starts, result = miss (traverse.cuh, ray_has_nan)
    LaneStack stack;       // (child code, entry distance bits)
};

// Starts a query.  ANY: occlusion query — returns true when a hoisted leaf already occludes the ray.  The hoisted
// leaves (ctx.cuh) are tested here, by all the lanes a refill starts together.
template <bool ANY, bool COUNT>
__device__ __forceinline__ bool lane_begin(const DeviceScene& S, LaneState& st, const RayQ& r, unsigned& n_tris) {
    st.r = r;
    st.best.t = B2PT_INF; st.best.tri = -1; st.best.u = 0.0f; st.best.v = 0.0f;
    st.best_leaf = -1;
    st.second = B2PT_INF;
    st.cull = r.T0;
    st.sp = 0;
    st.cur = 0;
    st.tri_next = st.tri_end = 0;
    st.tie = false;
    st.overflow = false;
    st.dead = ray_has_nan(r);
    if (st.dead) return false;
    for (int h = 0; h < S.nhoist; ++h) {
        if (!leaf_visible(S, S.hoist_leaf[h], r, r.T0)) continue;
        const int first = S.hoist_code[h] & 0x0FFFFFFF, cnt = ((S.hoist_code[h] >> 28) & 7) + 1;
        for (int i = first; i < first + cnt; ++i) {
            float t, u, v; int leaf;
            if (COUNT) ++n_tris;
            if (tri_fetch_test(S, i, r, r.T0, t, u, v, leaf)) {
                if (ANY) return true;
                if (t < st.best.t) {
                    st.second = st.best.t;
                    st.best.t = t; st.best.tri = i; st.best.u = u; st.best.v = v; st.tie = false; st.best_leaf = leaf;
                    st.cull = cull_after_hit(S, r, t);
                } else if (t == st.best.t) {
                    st.tie = true;
                } else {
                    st.second = fminf(st.second, t);
                }
            }
        }
    }
    if (S.nwide == 0) st.dead = true;   // nothing below the hoisted leaves
    return false;
}

// Expands wide node st.cur: slab-tests the 8 children at T0 and pushes the survivors sorted by entry
// distance (farthest deepest).  ANY = occlusion query: no ordering, no distance culling.
template <bool ANY, bool COUNT>
__device__ __forceinline__ void lane_node_step(const DeviceScene& S, LaneState& st, unsigned& n_nodes) {
    const WideNode* nd = &S.wide[st.cur];
    if (COUNT) ++n_nodes;
    if (st.sp + 8 > B2PT_TSTACK) { st.overflow = true; return; }
    const int base = st.sp;
    int sp = st.sp;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        Node4 n4;
        node_test4(nd, k, st.r, n4);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const float tmin = n4.tmin[s];
            if (ANY) {
                if (n4.pass[s]) st.stack.set(sp++, make_uint2(n4.code[s], 0u));
            } else if (n4.pass[s] && tmin <= st.cull) {
                int j = sp++;
                while (j > base) {
                    uint2 prev = st.stack.get(j - 1);
                    if (!(__uint_as_float(prev.y) < tmin)) break;
                    st.stack.set(j, prev);
                    --j;
                }
                st.stack.set(j, make_uint2(n4.code[s], __float_as_uint(tmin)));
            }
        }
    }
    st.sp = sp;
}

// Pops the next subtree that can still matter.  Returns false when the stack is exhausted.
template <bool ANY>
__device__ __forceinline__ bool lane_pop(LaneState& st) {
    while (st.sp > 0) {
        uint2 e = st.stack.get(--st.sp);
        if (ANY || __uint_as_float(e.y) <= st.cull) {
            if (e.x & B2PT_CHILD_LEAF) {
                st.tri_next = e.x & 0x0FFFFFFF;
                st.tri_end = st.tri_next + (int)((e.x >> 28) & 7) + 1;
            } else {
                st.cur = e.x;
            }
            return true;
        }
    }
    return false;
}

// One bounded step of a closest-hit query.  Returns true when the traversal is finished.
template <bool COUNT, int TPS>
__device__ __forceinline__ bool lane_closest_step(const DeviceScene& S, LaneState& st, unsigned& n_nodes, unsigned& n_tris) {
    if (st.dead) return true;
    if (st.tri_next < st.tri_end) {
#pragma unroll
        for (int k = 0; k < TPS; ++k) {
            if (st.tri_next < st.tri_end) {
                float t, u, v; int leaf;
                if (COUNT) ++n_tris;
                int i = st.tri_next++;
                if (tri_fetch_test(S, i, st.r, st.r.T0, t, u, v, leaf)) {
                    if (t < st.best.t) {
                        st.second = st.best.t;
                        st.best.t = t; st.best.tri = i; st.best.u = u; st.best.v = v; st.tie = false; st.best_leaf = leaf;
                        st.cull = cull_after_hit(S, st.r, t);
                    } else if (t == st.best.t) {
                        st.tie = true;
                    } else {
                        st.second = fminf(st.second, t);
                    }
                }
            }
        }
        if (st.tri_next < st.tri_end) return false;
    } else {
        lane_node_step<false, COUNT>(S, st, n_nodes);
        if (st.overflow) return true;
    }
    return !lane_pop<false>(st);
}

// After lane_closest_step returned true: is the result certified to be the reference's answer?
__device__ __forceinline__ bool lane_certify(const DeviceScene& S, const LaneState& st) {
    if (st.overflow) return false;
    if (st.best.tri < 0) return true;
    if (st.tie) return false;
    return certify_unique(S, st.r, st.best.t, st.best_leaf, st.second);
}

// One bounded step of an occlusion query.  Returns 0 = keep going, 1 = finished & occluded, 2 = finished
// & free, 3 = stack overflow (caller must use the exact recursion).
template <bool COUNT, int TPS>
__device__ __forceinline__ int lane_any_step(const DeviceScene& S, LaneState& st, unsigned& n_nodes, unsigned& n_tris) {
    if (st.dead) return st.overflow ? 3 : 2;
    if (st.tri_next < st.tri_end) {
#pragma unroll
        for (int k = 0; k < TPS; ++k) {
            if (st.tri_next < st.tri_end) {
                float t, u, v; int leaf;
                if (COUNT) ++n_tris;
                if (tri_fetch_test(S, st.tri_next++, st.r, st.r.T0, t, u, v, leaf)) return 1;
            }
        }
        if (st.tri_next < st.tri_end) return 0;
    } else {
        lane_node_step<true, COUNT>(S, st, n_nodes);
        if (st.overflow) return 3;
    }
    return lane_pop<true>(st) ? 0 : 2;
}

// Warp-local pool of work indices: the warp claims CHUNK indices per global atomic and hands them to idle
// lanes by ballot rank.  All 32 lanes must call refill() together.
struct WarpPool {
    long long next, end;   // warp-uniform
    bool exhausted;
};

template <int CHUNK>
__device__ __forceinline__ long long warp_pool_take(WarpPool& pool, unsigned long long* counter, long long total, bool want) {
    // returns a work index for lanes with want == true (or -1)
    const int lane = threadIdx.x & 31;
    long long mine = -1;
#pragma unroll 1
    for (int round = 0; round < 2; ++round) {
        unsigned need = __ballot_sync(0xffffffffu, want && mine < 0);
        if (need == 0) break;
        if (pool.next >= pool.end) {
            if (pool.exhausted) break;
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(counter, (unsigned long long)CHUNK);
            base = __shfl_sync(0xffffffffu, base, 0);
            pool.next = (long long)base;
            pool.end = min((long long)base + CHUNK, total);
            if (pool.next >= total) { pool.exhausted = true; pool.next = pool.end = 0; break; }
        }
        long long avail = pool.end - pool.next;
        int rank = __popc(need & ((1u << lane) - 1u));
        if (want && mine < 0 && rank < avail) mine = pool.next + rank;
        long long taken = min((long long)__popc(need), avail);
        pool.next += taken;
    }
    return mine;
}

}  // namespace b2pt
