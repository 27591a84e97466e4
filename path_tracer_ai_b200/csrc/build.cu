// build.cu — scene upload and acceleration-structure build, all on the device.
//
// Replaces OptixRenderer::uploadScene / buildAccelerationStructure (reference
// src/gpu/optix_renderer.cu:383-409, :233-353) and the box computation of BVH::buildRecursive
// (include/bvh.hpp:44-52).
//
// Two structures are built per upload:
//
//  1. The REFERENCE tree.  The triangles arrive in the reference's post-build order, so the reference tree is
//     implicit in the array: node = [start,end), mid = start + count/2, leaf iff count <= 8 (bvh.hpp:55-61).  Its
//     topology depends only on the triangle count (laid out on the host once per count, cached); every box is
//     computed on the GPU (leaf boxes from the triangles, inner boxes level by level, bottom-up; min/max are exact
//     so the result equals the reference's sequential fold).  It is what visibility is DEFINED on (a triangle can
//     only be hit through its reference leaf's box) and what the exact fallback kernel walks.
//
//  2. The TRAVERSAL tree: an 8-wide BVH over the reference LEAVES (never over single triangles — a leaf's exact
//     fp32 box is the unit of visibility), built from scratch on the GPU:
//        Morton codes of the leaf-box centroids -> radix sort (CUB) -> PLOC (parallel locally-ordered
//        clustering, Meister & Bittner 2018: every cluster looks R neighbours to each side along the Morton
//        order for the partner with the smallest merged surface area, mutual pairs merge, the array is
//        compacted, repeat) -> binary tree -> greedy surface-area collapse into 8-wide nodes, level by level.
//     Inner boxes are exact min/max unions of exact leaf boxes, so "a leaf passes the reference slab test" still
//     implies that every ancestor passes (monotone arithmetic), whatever the shape of the tree.  Unlike the
//     reference's object-median tree, the agglomerative build keeps a handful of huge leaves (the 16-unit room
//     triangles Scene::loadFromObj adds to every scene, src/scene.cpp:118-209) out of the small clusters: they
//     end up directly under the root instead of inflating every ancestor box on two root-to-leaf paths.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <type_traits>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "ctx.cuh"

namespace b2pt {

namespace {

// scratch slots of the scene buffers, of the upload staging and of the build (persistent, grow-only)
enum { SL_TRI = 16, SL_NRM, SL_NODE_LO, SL_NODE_HI, SL_INFO, SL_WIDE, SL_MATS, SL_POS, SL_NIN, SL_MAT, SL_IDS,
       SL_NB_LO, SL_NB_HI, SL_BN_CHILD, SL_KEYS, SL_VALS, SL_CL_NODE, SL_CL_LO, SL_CL_HI, SL_NN, SL_FLAGS, SL_SCAN,
       SL_CUB_TEMP, SL_WIDE_BIN, SL_COUNTERS, SL_ORDER_STATS, SL_PERM, SL_END };
static_assert(SL_END <= B2PT_SCRATCH_SLOTS, "scratch slots");

#define B2PT_PLOC_RADIUS 16      // neighbours searched to each side along the Morton order
#define B2PT_PLOC_BLOCK 256
#define B2PT_PLOC_FINISH 2048    // at most this many clusters: one block finishes the tree in a single launch

enum { CNT_M = 0, CNT_INNER = 1, CNT_NWIDE = 2, CNT_NHOIST = 3, CNT_BOUNDS = 4 /* centroid lo/hi: 6 encoded floats */,
       CNT_SCENE = 10 /* scene box lo/hi: 6 encoded floats */, CNT_HOIST = 16 /* B2PT_MAX_HOIST leaf ids */, CNT_WORDS = 32 };
#ifndef B2PT_HOIST_AREA_FRACTION
#define B2PT_HOIST_AREA_FRACTION 0.5f
#endif
//  // a leaf is hoisted when its box has at least this share of the scene box's surface area

struct HostNode { int start, end, right, leaf, depth; };

// Host: topology of the implicit reference tree in DFS pre-order.
void layout_tree(int start, int end, int depth, std::vector<HostNode>& nodes, int& nleaves) {
    // Iterative pre-order to keep 10M-triangle scenes off the host call stack.
    struct Item { int start, end, depth, parent; bool is_right; };
    std::vector<Item> stack;
    stack.push_back({start, end, depth, -1, false});
    while (!stack.empty()) {
        Item it = stack.back();
        stack.pop_back();
        int me = static_cast<int>(nodes.size());
        nodes.push_back({it.start, it.end, -1, -1, it.depth});
        if (it.parent >= 0 && it.is_right) nodes[it.parent].right = me;
        int count = it.end - it.start;
        if (count <= 8) {
            nodes[me].leaf = nleaves++;
        } else {
            int mid = it.start + count / 2;
            stack.push_back({mid, it.end, it.depth + 1, me, true});      // right: popped after the whole left subtree
            stack.push_back({it.start, mid, it.depth + 1, me, false});   // left: next index (me + 1)
        }
    }
}

__global__ void k_pack_triangles(const float* __restrict__ pos, const float* __restrict__ nrm,
                                 const int32_t* __restrict__ mat, int ntri,
                                 float4* __restrict__ tri, float4* __restrict__ nout) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntri) return;
    const float* p = pos + 9ll * i;
    V3 v0 = mk3(p[0], p[1], p[2]), v1 = mk3(p[3], p[4], p[5]), v2 = mk3(p[6], p[7], p[8]);
    V3 e1 = vsub(v1, v0), e2 = vsub(v2, v0);   // triangle.hpp:28-29
    tri[3ll * i + 0] = make_float4(v0.x, v0.y, v0.z, 0.0f);   // w = leaf id, patched by k_leaf_boxes
    tri[3ll * i + 1] = make_float4(e1.x, e1.y, e1.z, 0.0f);
    tri[3ll * i + 2] = make_float4(e2.x, e2.y, e2.z, 0.0f);
    float n[9];
    if (nrm) { for (int k = 0; k < 9; ++k) n[k] = nrm[9ll * i + k]; }
    else { for (int k = 0; k < 9; ++k) n[k] = 0.0f; }
    int m = mat ? mat[i] : 0;
    nout[3ll * i + 0] = make_float4(n[0], n[1], n[2], __int_as_float(m));
    nout[3ll * i + 1] = make_float4(n[3], n[4], n[5], 0.0f);
    nout[3ll * i + 2] = make_float4(n[6], n[7], n[8], 0.0f);
}

// One thread per reference node; leaves only: box = fold of Triangle::getAABB (triangle.hpp:73-77) with
// AABB::merge (aabb.hpp:27-32); tags the leaf's triangles with the leaf id.  The box goes into the reference tree
// (node_lo/hi) and, by leaf id, into the first `nleaves` entries of the traversal tree's box arrays (nb_lo/hi), whose
// .w carries the leaf's child code (bit31 | count-1 << 28 | first triangle).
__global__ void k_leaf_boxes(const float* __restrict__ pos, const int4* __restrict__ node_info, int nnodes,
                             float4* __restrict__ node_lo, float4* __restrict__ node_hi,
                             float4* __restrict__ nb_lo, float4* __restrict__ nb_hi, float4* __restrict__ tri) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnodes) return;
    int4 info = node_info[i];
    if (info.w < 0) return;
    float lo[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
    float hi[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    for (int t = info.x; t < info.y; ++t) {
        const float* p = pos + 9ll * t;
        for (int a = 0; a < 3; ++a) {
            float mn = gmin(gmin(p[a], p[3 + a]), p[6 + a]);
            float mx = gmax(gmax(p[a], p[3 + a]), p[6 + a]);
            lo[a] = gmin(lo[a], mn);
            hi[a] = gmax(hi[a], mx);
        }
        tri[3ll * t].w = __int_as_float(info.w);
    }
    const uint32_t code = B2PT_CHILD_LEAF | (static_cast<uint32_t>(info.y - info.x - 1) << 28) | static_cast<uint32_t>(info.x);
    node_lo[i] = make_float4(lo[0], lo[1], lo[2], 0.0f);
    node_hi[i] = make_float4(hi[0], hi[1], hi[2], 0.0f);
    nb_lo[info.w] = make_float4(lo[0], lo[1], lo[2], __uint_as_float(code));
    nb_hi[info.w] = make_float4(hi[0], hi[1], hi[2], 0.0f);
}

// Inner nodes of one depth of the reference tree: box = merge(left, right).
__global__ void k_inner_boxes(const int* __restrict__ ids, int n, const int4* __restrict__ node_info,
                              float4* __restrict__ node_lo, float4* __restrict__ node_hi) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int i = ids[k];
    int l = i + 1, r = node_info[i].z;
    float4 a = node_lo[l], b = node_lo[r], c = node_hi[l], d = node_hi[r];
    node_lo[i] = make_float4(gmin(a.x, b.x), gmin(a.y, b.y), gmin(a.z, b.z), 0.0f);
    node_hi[i] = make_float4(gmax(c.x, d.x), gmax(c.y, d.y), gmax(c.z, d.z), 0.0f);
}

// ---- traversal tree: Morton order --------------------------------------------------------------------------
// order-preserving float <-> uint map, so bounds reduce with integer atomics
__device__ __forceinline__ unsigned f2ord(float f) { unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__global__ void k_init_counters(unsigned* __restrict__ cnt, int m) {
    if (threadIdx.x == 0) {
        cnt[CNT_M] = (unsigned)m; cnt[CNT_INNER] = 0; cnt[CNT_NWIDE] = 1; cnt[CNT_NHOIST] = 0;
        for (int a = 0; a < 3; ++a) {
            cnt[CNT_BOUNDS + a] = 0xffffffffu; cnt[CNT_BOUNDS + 3 + a] = 0u;
            cnt[CNT_SCENE + a] = 0xffffffffu; cnt[CNT_SCENE + 3 + a] = 0u;
        }
    }
}

__global__ void __launch_bounds__(256) k_centroid_bounds(const float4* __restrict__ nb_lo, const float4* __restrict__ nb_hi, int n, unsigned* __restrict__ cnt) {
    const float big = 3.402823466e+38f;
    float lo[3] = {big, big, big}, hi[3] = {-big, -big, -big};      // of the leaf-box centroids (Morton grid)
    float slo[3] = {big, big, big}, shi[3] = {-big, -big, -big};    // of the leaf boxes (scene box)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 a = nb_lo[i], b = nb_hi[i];
        float c[3] = {0.5f * a.x + 0.5f * b.x, 0.5f * a.y + 0.5f * b.y, 0.5f * a.z + 0.5f * b.z};
        float l[3] = {a.x, a.y, a.z}, h[3] = {b.x, b.y, b.z};
        for (int k = 0; k < 3; ++k) {
            lo[k] = fminf(lo[k], c[k]); hi[k] = fmaxf(hi[k], c[k]);
            slo[k] = fminf(slo[k], l[k]); shi[k] = fmaxf(shi[k], h[k]);
        }
    }
    for (int k = 0; k < 3; ++k) {
        for (int off = 16; off > 0; off >>= 1) {
            lo[k] = fminf(lo[k], __shfl_down_sync(0xffffffffu, lo[k], off));
            hi[k] = fmaxf(hi[k], __shfl_down_sync(0xffffffffu, hi[k], off));
            slo[k] = fminf(slo[k], __shfl_down_sync(0xffffffffu, slo[k], off));
            shi[k] = fmaxf(shi[k], __shfl_down_sync(0xffffffffu, shi[k], off));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&cnt[CNT_BOUNDS + k], f2ord(lo[k])); atomicMax(&cnt[CNT_BOUNDS + 3 + k], f2ord(hi[k]));
            atomicMin(&cnt[CNT_SCENE + k], f2ord(slo[k])); atomicMax(&cnt[CNT_SCENE + 3 + k], f2ord(shi[k]));
        }
    }
}

__device__ __forceinline__ unsigned expand_bits10(unsigned v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

// Leaves whose box has at least B2PT_HOIST_AREA_FRACTION of the scene box's surface area are hoisted out of the tree
// (at most B2PT_MAX_HOIST of them; which ones, when more qualify, does not matter for any result).
__global__ void __launch_bounds__(256) k_select_hoist(const float4* __restrict__ nb_lo, const float4* __restrict__ nb_hi, int n, unsigned* __restrict__ cnt,
                                                      int* __restrict__ hoisted) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float sx = ord2f(cnt[CNT_SCENE + 3]) - ord2f(cnt[CNT_SCENE]), sy = ord2f(cnt[CNT_SCENE + 4]) - ord2f(cnt[CNT_SCENE + 1]),
          sz = ord2f(cnt[CNT_SCENE + 5]) - ord2f(cnt[CNT_SCENE + 2]);
    float4 a = nb_lo[i], b = nb_hi[i];
    float dx = b.x - a.x, dy = b.y - a.y, dz = b.z - a.z;
    int h = 0;
    if (n > 1 && dx * dy + dy * dz + dz * dx >= B2PT_HOIST_AREA_FRACTION * (sx * sy + sy * sz + sz * sx)) {
        unsigned slot = atomicAdd(&cnt[CNT_NHOIST], 1u);
        if (slot < B2PT_MAX_HOIST) { cnt[CNT_HOIST + slot] = (unsigned)i; h = 1; }
    }
    hoisted[i] = h;
}

__global__ void __launch_bounds__(256) k_morton(const float4* __restrict__ nb_lo, const float4* __restrict__ nb_hi, int n, const unsigned* __restrict__ cnt,
                                                const int* __restrict__ hoisted, unsigned* __restrict__ keys, int* __restrict__ vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    vals[i] = i;
    if (hoisted[i]) { keys[i] = 0xffffffffu; return; }   // sorts behind every real Morton code (30 bits)
    float4 a = nb_lo[i], b = nb_hi[i];
    float c[3] = {0.5f * a.x + 0.5f * b.x, 0.5f * a.y + 0.5f * b.y, 0.5f * a.z + 0.5f * b.z};
    unsigned q[3];
    for (int k = 0; k < 3; ++k) {
        float lo = ord2f(cnt[CNT_BOUNDS + k]), hi = ord2f(cnt[CNT_BOUNDS + 3 + k]);
        float ext = hi - lo;
        float x = ext > 0.0f ? (c[k] - lo) / ext : 0.0f;
        x = fminf(fmaxf(x * 1024.0f, 0.0f), 1023.0f);
        q[k] = isfinite(x) ? (unsigned)x : 0u;
    }
    keys[i] = (expand_bits10(q[0]) << 2) | (expand_bits10(q[1]) << 1) | expand_bits10(q[2]);
}

__global__ void __launch_bounds__(256) k_ploc_init(const int* __restrict__ sorted_leaf, int n, const float4* __restrict__ nb_lo, const float4* __restrict__ nb_hi,
                                                   int* __restrict__ cl_node, float4* __restrict__ cl_lo, float4* __restrict__ cl_hi) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int leaf = sorted_leaf[i];
    cl_node[i] = leaf;
    cl_lo[i] = nb_lo[leaf];
    cl_hi[i] = nb_hi[leaf];
}

// ---- traversal tree: PLOC ------------------------------------------------------------------------------------
__device__ __forceinline__ float merged_area(float4 alo, float4 ahi, float4 blo, float4 bhi) {
    float dx = fmaxf(ahi.x, bhi.x) - fminf(alo.x, blo.x);
    float dy = fmaxf(ahi.y, bhi.y) - fminf(alo.y, blo.y);
    float dz = fmaxf(ahi.z, bhi.z) - fminf(alo.z, blo.z);
    return dx * dy + dy * dz + dz * dx;
}

// Partner of cluster i among its window: the candidate j minimising the key (merged area, j != i^1, |i-j|, min(i,j)).
// The key is a strict total order on unordered pairs and symmetric in (i, j), so the globally smallest pair is always
// mutual (progress), and a run of identical boxes pairs up (2k, 2k+1) instead of chaining.
__device__ __forceinline__ bool pair_better(float d, int i, int j, float bd, int bj) {
    if (d != bd) return d < bd;
    const int rj = (j == (i ^ 1)) ? 0 : 1, rb = (bj == (i ^ 1)) ? 0 : 1;
    if (rj != rb) return rj < rb;
    const int aj = abs(i - j), ab = abs(i - bj);
    if (aj != ab) return aj < ab;
    return j < bj;
}

__global__ void __launch_bounds__(B2PT_PLOC_BLOCK) k_ploc_nn(int m, int force, const float4* __restrict__ cl_lo, const float4* __restrict__ cl_hi, int* __restrict__ nn) {
    __shared__ float4 s_lo[B2PT_PLOC_BLOCK + 2 * B2PT_PLOC_RADIUS], s_hi[B2PT_PLOC_BLOCK + 2 * B2PT_PLOC_RADIUS];
    const int base = blockIdx.x * B2PT_PLOC_BLOCK - B2PT_PLOC_RADIUS;
    for (int k = threadIdx.x; k < B2PT_PLOC_BLOCK + 2 * B2PT_PLOC_RADIUS; k += B2PT_PLOC_BLOCK) {
        int g = base + k;
        if (g >= 0 && g < m) { s_lo[k] = cl_lo[g]; s_hi[k] = cl_hi[g]; }
    }
    __syncthreads();
    const int i = blockIdx.x * B2PT_PLOC_BLOCK + threadIdx.x;
    if (i >= m) return;
    if (force) {   // safety net against adversarial inputs: pair neighbours unconditionally (halves m)
        int j = i ^ 1;
        nn[i] = j < m ? j : i;
        return;
    }
    const float4 alo = s_lo[threadIdx.x + B2PT_PLOC_RADIUS], ahi = s_hi[threadIdx.x + B2PT_PLOC_RADIUS];
    float bd = 3.402823466e+38f;
    int bj = -1;
    for (int k = -B2PT_PLOC_RADIUS; k <= B2PT_PLOC_RADIUS; ++k) {
        const int j = i + k;
        if (k == 0 || j < 0 || j >= m) continue;
        const float d = merged_area(alo, ahi, s_lo[threadIdx.x + B2PT_PLOC_RADIUS + k], s_hi[threadIdx.x + B2PT_PLOC_RADIUS + k]);
        if (bj < 0 || pair_better(d, i, j, bd, bj)) { bd = d; bj = j; }
    }
    nn[i] = bj < 0 ? i : bj;
}

// Mutual pairs merge into a new binary node (in place, at the smaller index); flags[i] = entry i survives.
__device__ __forceinline__ int ploc_merge_one(int i, int n, const int* __restrict__ nn, int* __restrict__ cl_node, float4* __restrict__ cl_lo,
                                              float4* __restrict__ cl_hi, unsigned* __restrict__ cnt, int2* __restrict__ bn_child,
                                              float4* __restrict__ nb_lo, float4* __restrict__ nb_hi) {
    const int j = nn[i];
    if (j == i || nn[j] != i) return 1;
    if (i > j) return 0;
    const int k = (int)atomicAdd(&cnt[CNT_INNER], 1u);
    const float4 alo = cl_lo[i], ahi = cl_hi[i], blo = cl_lo[j], bhi = cl_hi[j];
    const float4 lo = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.0f);
    const float4 hi = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.0f);
    bn_child[k] = make_int2(cl_node[i], cl_node[j]);
    nb_lo[n + k] = lo; nb_hi[n + k] = hi;
    cl_node[i] = n + k; cl_lo[i] = lo; cl_hi[i] = hi;
    return 1;
}

__global__ void __launch_bounds__(256) k_ploc_merge(int m, int n, const int* __restrict__ nn, int* __restrict__ cl_node, float4* __restrict__ cl_lo,
                                                    float4* __restrict__ cl_hi, unsigned* __restrict__ cnt, int2* __restrict__ bn_child,
                                                    float4* __restrict__ nb_lo, float4* __restrict__ nb_hi, int* __restrict__ flags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    flags[i] = ploc_merge_one(i, n, nn, cl_node, cl_lo, cl_hi, cnt, bn_child, nb_lo, nb_hi);
}

__global__ void __launch_bounds__(256) k_ploc_scatter(int m, const int* __restrict__ flags, const int* __restrict__ pos,
                                                      const int* __restrict__ in_node, const float4* __restrict__ in_lo, const float4* __restrict__ in_hi,
                                                      int* __restrict__ out_node, float4* __restrict__ out_lo, float4* __restrict__ out_hi,
                                                      unsigned* __restrict__ cnt) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    if (flags[i]) {
        int p = pos[i];
        out_node[p] = in_node[i]; out_lo[p] = in_lo[i]; out_hi[p] = in_hi[i];
    }
    if (i == m - 1) cnt[CNT_M] = (unsigned)(pos[i] + flags[i]);
}

// The last <= B2PT_PLOC_FINISH clusters: one block runs every remaining iteration (search, merge, compact) itself.
// Thread t owns entries 2t and 2t+1, so a block-wide exclusive scan of the pair sums gives the compacted positions.
__global__ void __launch_bounds__(B2PT_PLOC_FINISH / 2) k_ploc_finish(int m, int n, int cur, int* node0, int* node1, float4* lo0, float4* lo1, float4* hi0, float4* hi1,
                                                                      int* __restrict__ nn, unsigned* __restrict__ cnt, int2* __restrict__ bn_child,
                                                                      float4* __restrict__ nb_lo, float4* __restrict__ nb_hi) {
    __shared__ int s_warp[B2PT_PLOC_FINISH / 64];
    __shared__ int s_total;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    int iter = 0;
    while (m > 1) {
        int* node = cur ? node1 : node0; float4* lo = cur ? lo1 : lo0; float4* hi = cur ? hi1 : hi0;
        int* onode = cur ? node0 : node1; float4* olo = cur ? lo0 : lo1; float4* ohi = cur ? hi0 : hi1;
        const bool force = iter > 256;
        for (int e = 0; e < 2; ++e) {
            const int i = 2 * t + e;
            if (i >= m) continue;
            if (force) { int j = i ^ 1; nn[i] = j < m ? j : i; continue; }
            const float4 alo = lo[i], ahi = hi[i];
            float bd = 3.402823466e+38f;
            int bj = -1;
            for (int k = -B2PT_PLOC_RADIUS; k <= B2PT_PLOC_RADIUS; ++k) {
                const int j = i + k;
                if (k == 0 || j < 0 || j >= m) continue;
                const float d = merged_area(alo, ahi, lo[j], hi[j]);
                if (bj < 0 || pair_better(d, i, j, bd, bj)) { bd = d; bj = j; }
            }
            nn[i] = bj < 0 ? i : bj;
        }
        __syncthreads();
        int f[2] = {0, 0};
        for (int e = 0; e < 2; ++e) {
            const int i = 2 * t + e;
            if (i < m) f[e] = ploc_merge_one(i, n, nn, node, lo, hi, cnt, bn_child, nb_lo, nb_hi);
        }
        // block exclusive scan of f[0] + f[1]
        int sum = f[0] + f[1], incl = sum;
        for (int off = 1; off < 32; off <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += v; }
        if (lane == 31) s_warp[w] = incl;
        __syncthreads();   // also orders the in-place merges before the scatter reads
        if (w == 0) {
            int v = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0, iv = v;
            for (int off = 1; off < 32; off <<= 1) { int u = __shfl_up_sync(0xffffffffu, iv, off); if (lane >= off) iv += u; }
            if (lane < (int)(blockDim.x >> 5)) s_warp[lane] = iv - v;
            if (lane == 31) s_total = iv;
        }
        __syncthreads();
        int p = s_warp[w] + incl - sum;
        for (int e = 0; e < 2; ++e) {
            const int i = 2 * t + e;
            if (i < m && f[e]) { onode[p] = node[i]; olo[p] = lo[i]; ohi[p] = hi[i]; ++p; }
        }
        m = s_total;
        cur ^= 1;
        ++iter;
        __syncthreads();
    }
    if (t == 0) { cnt[CNT_M] = 1u; node0[0] = (cur ? node1 : node0)[0]; }   // root of the binary tree -> node0[0]
}

// ---- traversal tree: 8-wide collapse -------------------------------------------------------------------------
// Wide node w expands binary node wide_bin[w]: start from its two children and keep replacing one inner child by its
// own two children until there are 8 (or nothing is left to gain).  A ray visits a wide node with a probability
// proportional to its box area, and every visit costs the same 8-slot test however full the node is, so the expected
// number of node visits is the sum of the areas of the binary nodes that END UP as wide nodes.  Expanding child X in
// place removes area(X) from that sum and adds the areas of X's inner children: the child with the largest such
// gain goes first (a small subtree of two leaves is pure gain; a big node whose children overlap badly is not).
// Inner children get the next free wide indices (children always come after their parent: BFS order).
__global__ void __launch_bounds__(128) k_collapse_level(int begin, int end, int n, int* __restrict__ wide_bin, const int2* __restrict__ bn_child,
                                                        const float4* __restrict__ nb_lo, const float4* __restrict__ nb_hi,
                                                        unsigned* __restrict__ cnt, WideNode* __restrict__ wide) {
    int w = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= end) return;
    int ch[8];
    float gain[8];   // what expanding this child would save; leaves: -inf
    int c = 0;
    auto area_of = [&](int id) {
        float4 l = nb_lo[id], h = nb_hi[id];
        float dx = h.x - l.x, dy = h.y - l.y, dz = h.z - l.z;
        return dx * dy + dy * dz + dz * dx;
    };
    auto gain_of = [&](int id) {
        if (id < n) return -3.402823466e+38f;
        int2 k = bn_child[id - n];
        float g = area_of(id);
        if (k.x >= n) g -= area_of(k.x);
        if (k.y >= n) g -= area_of(k.y);
        return g;
    };
    const int b = wide_bin[w];
    if (b < n) {            // the whole tree is one reference leaf
        ch[0] = b; c = 1;
    } else {
        int2 k = bn_child[b - n];
        ch[0] = k.x; ch[1] = k.y; gain[0] = gain_of(k.x); gain[1] = gain_of(k.y); c = 2;
        while (c < 8) {
            int best = -1;
            for (int s = 0; s < c; ++s) if (gain[s] > 0.0f && (best < 0 || gain[s] > gain[best])) best = s;
            if (best < 0) {   // nothing gains: still fill the node from the largest inner child (an empty slot costs the same)
                float ba = -1.0f;
                for (int s = 0; s < c; ++s) if (ch[s] >= n) { float a = area_of(ch[s]); if (a > ba) { ba = a; best = s; } }
                if (best < 0) break;
            }
            int2 kk = bn_child[ch[best] - n];
            ch[best] = kk.x; gain[best] = gain_of(kk.x);
            ch[c] = kk.y; gain[c] = gain_of(kk.y);
            ++c;
        }
    }
    WideNode& nd = wide[w];
    for (int s = 0; s < 8; ++s) {
        if (s >= c) {
            // inverted box: never passes a slab test
            nd.lox[s] = nd.loy[s] = nd.loz[s] = 3.402823466e+38f;
            nd.hix[s] = nd.hiy[s] = nd.hiz[s] = -3.402823466e+38f;
            nd.child[s] = B2PT_CHILD_EMPTY;
            continue;
        }
        float4 l = nb_lo[ch[s]], h = nb_hi[ch[s]];
        nd.lox[s] = l.x; nd.loy[s] = l.y; nd.loz[s] = l.z;
        nd.hix[s] = h.x; nd.hiy[s] = h.y; nd.hiz[s] = h.z;
        if (ch[s] < n) {
            nd.child[s] = __float_as_uint(l.w);   // the leaf's code (k_leaf_boxes)
        } else {
            int wi = (int)atomicAdd(&cnt[CNT_NWIDE], 1u);
            wide_bin[wi] = ch[s];
            nd.child[s] = (uint32_t)wi;
        }
    }
}

// Re-orders the child slots of every wide node: slot s of the new node is slot perm[8*w+s] of the old one.
__global__ void __launch_bounds__(128) k_permute_slots(int nwide, const uint8_t* __restrict__ perm, WideNode* __restrict__ wide) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwide) return;
    WideNode in = wide[w], out;
    for (int s = 0; s < 8; ++s) {
        int k = perm[8 * (size_t)w + s];
        out.lox[s] = in.lox[k]; out.loy[s] = in.loy[k]; out.loz[s] = in.loz[k];
        out.hix[s] = in.hix[k]; out.hiy[s] = in.hiy[k]; out.hiz[s] = in.hiz[k];
        out.child[s] = in.child[k];
    }
    wide[w] = out;
}

}  // namespace

void free_scene(b2pt_ctx* ctx) {
    // scene buffers are persistent scratch slots of the context (freed by b2pt_destroy); only forget the scene
    ctx->has_scene = false;
    ctx->scene = DeviceScene{};
}

int build_scene(b2pt_ctx* ctx, const float* pos, const float* nrm, const int32_t* mat, int64_t ntri64,
                const b2pt_material* mats, int32_t nmat, const b2pt_light* lights, int32_t nlight) {
    if (ntri64 < 0 || ntri64 >= (1ll << 28)) { ctx->err = "b2pt_upload_scene: triangle count out of range (max 2^28-1)"; return B2PT_ERR_INVALID; }
    if (nlight < 0 || nlight > B2PT_MAX_LIGHTS) { ctx->err = "b2pt_upload_scene: at most 16 lights"; return B2PT_ERR_INVALID; }
    if (nmat < 0) { ctx->err = "b2pt_upload_scene: negative material count"; return B2PT_ERR_INVALID; }
    if (ntri64 > 0 && !pos) { ctx->err = "b2pt_upload_scene: pos is NULL"; return B2PT_ERR_INVALID; }
    free_scene(ctx);
    const int ntri = static_cast<int>(ntri64);
    cudaStream_t st = ctx->stream;
    DeviceScene S{};
    S.ntri = ntri;

    // Every device buffer of the scene, of its staging and of the build is a persistent grow-only allocation of the
    // context: re-uploading a scene of the same size — the reference's GPU branch uploads once per run, an interactive
    // caller once per frame — costs no cudaMalloc/cudaFree.

    // The topology of the REFERENCE tree depends on the triangle COUNT only: it is laid out on the host once per
    // count and its device copies (node_info, per-depth id lists) are reused.
    b2pt_ctx::Topology& T = ctx->topo;
    const bool topo_cached = T.valid && T.ntri == ntri;
    if (!topo_cached) {
        T = b2pt_ctx::Topology{};
        T.ntri = ntri;
        std::vector<HostNode> nodes;
        if (ntri > 0) {
            nodes.reserve(static_cast<size_t>(ntri) / 2 + 16);
            layout_tree(0, ntri, 0, nodes, T.nleaves);
        }
        const int nn = static_cast<int>(nodes.size());
        T.info.assign(nn, int4{});
        for (int i = 0; i < nn; ++i) {
            T.info[i] = make_int4(nodes[i].start, nodes[i].end, nodes[i].right, nodes[i].leaf);
            T.maxdepth = std::max(T.maxdepth, nodes[i].depth);
        }
        // inner nodes grouped by depth, deepest first, flattened: the bottom-up box pass launches one kernel per span
        std::vector<std::vector<int>> by_depth(T.maxdepth + 1);
        for (int i = 0; i < nn; ++i) if (nodes[i].leaf < 0) by_depth[nodes[i].depth].push_back(i);
        for (int dpt = T.maxdepth; dpt >= 0; --dpt) {
            T.spans.push_back({T.ids_flat.size(), by_depth[dpt].size()});
            T.ids_flat.insert(T.ids_flat.end(), by_depth[dpt].begin(), by_depth[dpt].end());
        }
        T.nnodes = nn;
        T.valid = true;
    }
    const int nnodes = T.nnodes, nleaves = T.nleaves;
    S.nnodes = nnodes; S.nleaves = nleaves;

    // ---- device buffers ------------------------------------------------------------------------------
    int rc;
    auto reserve = [&](int slot, size_t bytes, auto** out) {
        void* p = nullptr;
        int r = scratch_reserve(ctx, slot, std::max<size_t>(bytes, 16), &p);
        *out = static_cast<std::remove_reference_t<decltype(**out)>*>(p);
        return r;
    };
    float4 *d_tri, *d_nrm, *d_node_lo, *d_node_hi, *d_nb_lo, *d_nb_hi;
    int4* d_info; WideNode* d_wide; DMaterial* d_mats; unsigned* d_cnt;
    const size_t nl = static_cast<size_t>(std::max(nleaves, 1));
    if ((rc = reserve(SL_TRI, sizeof(float4) * 3ull * ntri, &d_tri))) return rc;
    if ((rc = reserve(SL_NRM, sizeof(float4) * 3ull * ntri, &d_nrm))) return rc;
    if ((rc = reserve(SL_NODE_LO, sizeof(float4) * (size_t)nnodes, &d_node_lo))) return rc;
    if ((rc = reserve(SL_NODE_HI, sizeof(float4) * (size_t)nnodes, &d_node_hi))) return rc;
    if ((rc = reserve(SL_NB_LO, sizeof(float4) * 2 * nl, &d_nb_lo))) return rc;   // leaves [0, nleaves), binary inner nodes after them
    if ((rc = reserve(SL_NB_HI, sizeof(float4) * 2 * nl, &d_nb_hi))) return rc;
    if ((rc = reserve(SL_INFO, sizeof(int4) * (size_t)nnodes, &d_info))) return rc;
    if ((rc = reserve(SL_WIDE, sizeof(WideNode) * nl, &d_wide))) return rc;       // a wide node per binary inner node at worst
    if ((rc = reserve(SL_MATS, sizeof(DMaterial) * (size_t)std::max(nmat, 1), &d_mats))) return rc;
    if ((rc = reserve(SL_COUNTERS, sizeof(unsigned) * CNT_WORDS, &d_cnt))) return rc;

    float *d_pos = nullptr, *d_nin = nullptr; int32_t* d_mat = nullptr; int* d_ids = nullptr;
    int nwide = 0;
    int64_t build_launches = 0, ploc_iters = 0, wide_levels = 0;
    float scene_box[6] = {0, 0, 0, 0, 0, 0};
    unsigned h_cnt[CNT_WORDS] = {};
    int nhoist = 0;
#define STAGE(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cuda_fail(ctx, e__, #call, __FILE__, __LINE__); return B2PT_ERR_CUDA; } } while (0)
    if (ntri > 0) {
        if ((rc = reserve(SL_POS, 9ull * ntri * sizeof(float), &d_pos))) return rc;
        STAGE(cudaMemcpyAsync(d_pos, pos, 9ull * ntri * sizeof(float), cudaMemcpyHostToDevice, st));
        if (nrm) {
            if ((rc = reserve(SL_NIN, 9ull * ntri * sizeof(float), &d_nin))) return rc;
            STAGE(cudaMemcpyAsync(d_nin, nrm, 9ull * ntri * sizeof(float), cudaMemcpyHostToDevice, st));
        }
        if (mat) {
            if ((rc = reserve(SL_MAT, 1ull * ntri * sizeof(int32_t), &d_mat))) return rc;
            STAGE(cudaMemcpyAsync(d_mat, mat, 1ull * ntri * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        }
        const size_t ninner = T.ids_flat.size();
        if ((rc = reserve(SL_IDS, sizeof(int) * std::max<size_t>(ninner, 1), &d_ids))) return rc;
        if (!topo_cached || !T.on_device) {
            STAGE(cudaMemcpyAsync(d_info, T.info.data(), sizeof(int4) * nnodes, cudaMemcpyHostToDevice, st));
            if (ninner) STAGE(cudaMemcpyAsync(d_ids, T.ids_flat.data(), sizeof(int) * ninner, cudaMemcpyHostToDevice, st));
            T.on_device = true;
        }
        // build scratch
        unsigned *d_keys; int *d_vals, *d_cl_node, *d_nn, *d_flags, *d_scan, *d_wide_bin; float4 *d_cl_lo, *d_cl_hi; int2* d_bn_child; char* d_cub;
        if ((rc = reserve(SL_KEYS, sizeof(unsigned) * 2 * nl, &d_keys))) return rc;
        if ((rc = reserve(SL_VALS, sizeof(int) * 2 * nl, &d_vals))) return rc;
        if ((rc = reserve(SL_CL_NODE, sizeof(int) * 2 * nl, &d_cl_node))) return rc;
        if ((rc = reserve(SL_CL_LO, sizeof(float4) * 2 * nl, &d_cl_lo))) return rc;
        if ((rc = reserve(SL_CL_HI, sizeof(float4) * 2 * nl, &d_cl_hi))) return rc;
        if ((rc = reserve(SL_NN, sizeof(int) * nl, &d_nn))) return rc;
        if ((rc = reserve(SL_FLAGS, sizeof(int) * nl, &d_flags))) return rc;
        if ((rc = reserve(SL_SCAN, sizeof(int) * nl, &d_scan))) return rc;
        if ((rc = reserve(SL_BN_CHILD, sizeof(int2) * nl, &d_bn_child))) return rc;
        if ((rc = reserve(SL_WIDE_BIN, sizeof(int) * nl, &d_wide_bin))) return rc;
        size_t cub_sort = 0, cub_scan = 0;
        cub::DoubleBuffer<unsigned> kb(d_keys, d_keys + nl);
        cub::DoubleBuffer<int> vb(d_vals, d_vals + nl);
        STAGE(cub::DeviceRadixSort::SortPairs(nullptr, cub_sort, kb, vb, nleaves, 0, 32, st));
        STAGE(cub::DeviceScan::ExclusiveSum(nullptr, cub_scan, d_flags, d_scan, nleaves, st));
        const size_t cub_bytes = std::max(cub_sort, cub_scan);
        if ((rc = reserve(SL_CUB_TEMP, cub_bytes, &d_cub))) return rc;

        // ---- device build, timed -----------------------------------------------------------------
        STAGE(cudaEventRecord(ctx->ev2, st));
        const int B = 256;
        k_pack_triangles<<<(ntri + B - 1) / B, B, 0, st>>>(d_pos, d_nin, d_mat, ntri, d_tri, d_nrm);
        k_leaf_boxes<<<(nnodes + B - 1) / B, B, 0, st>>>(d_pos, d_info, nnodes, d_node_lo, d_node_hi, d_nb_lo, d_nb_hi, d_tri);
        build_launches += 2;
        for (auto& sp : T.spans) {
            if (!sp.second) continue;
            k_inner_boxes<<<(static_cast<int>(sp.second) + B - 1) / B, B, 0, st>>>(d_ids + sp.first, static_cast<int>(sp.second), d_info, d_node_lo, d_node_hi);
            ++build_launches;
        }
        // Morton order of the reference leaves
        k_init_counters<<<1, 32, 0, st>>>(d_cnt, nleaves);
        k_centroid_bounds<<<std::min((nleaves + B - 1) / B, ctx->sm_count * 8), B, 0, st>>>(d_nb_lo, d_nb_hi, nleaves, d_cnt);
        k_select_hoist<<<(nleaves + B - 1) / B, B, 0, st>>>(d_nb_lo, d_nb_hi, nleaves, d_cnt, d_flags);
        k_morton<<<(nleaves + B - 1) / B, B, 0, st>>>(d_nb_lo, d_nb_hi, nleaves, d_cnt, d_flags, kb.Current(), vb.Current());
        size_t tmp = cub_bytes;
        STAGE(cub::DeviceRadixSort::SortPairs(d_cub, tmp, kb, vb, nleaves, 0, 32, st));
        // hoisted leaves sort to the end of the Morton order: the tree is built over the others
        STAGE(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
        STAGE(cudaStreamSynchronize(st));
        nhoist = (int)std::min<unsigned>(h_cnt[CNT_NHOIST], B2PT_MAX_HOIST);
        const int nactive = nleaves - nhoist;
        int cur = 0;   // which half of the cluster arrays is current
        int* cl_node[2] = {d_cl_node, d_cl_node + nl};
        float4* cl_lo[2] = {d_cl_lo, d_cl_lo + nl};
        float4* cl_hi[2] = {d_cl_hi, d_cl_hi + nl};
        if (nactive > 0) k_ploc_init<<<(nactive + B - 1) / B, B, 0, st>>>(vb.Current(), nactive, d_nb_lo, d_nb_hi, cl_node[0], cl_lo[0], cl_hi[0]);
        build_launches += 6;
        // PLOC: search, merge, compact until few enough clusters are left for one block
        int m = nactive;
        const int iter_cap = 96;
        while (m > B2PT_PLOC_FINISH) {
            const int g = (m + B2PT_PLOC_BLOCK - 1) / B2PT_PLOC_BLOCK;
            k_ploc_nn<<<g, B2PT_PLOC_BLOCK, 0, st>>>(m, ploc_iters > iter_cap ? 1 : 0, cl_lo[cur], cl_hi[cur], d_nn);
            k_ploc_merge<<<(m + B - 1) / B, B, 0, st>>>(m, nleaves, d_nn, cl_node[cur], cl_lo[cur], cl_hi[cur], d_cnt, d_bn_child, d_nb_lo, d_nb_hi, d_flags);
            tmp = cub_bytes;
            STAGE(cub::DeviceScan::ExclusiveSum(d_cub, tmp, d_flags, d_scan, m, st));
            k_ploc_scatter<<<(m + B - 1) / B, B, 0, st>>>(m, d_flags, d_scan, cl_node[cur], cl_lo[cur], cl_hi[cur], cl_node[cur ^ 1], cl_lo[cur ^ 1], cl_hi[cur ^ 1], d_cnt);
            unsigned hm = 0;
            STAGE(cudaMemcpyAsync(&hm, d_cnt + CNT_M, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
            STAGE(cudaStreamSynchronize(st));
            if ((int)hm >= m || hm == 0) { ctx->err = "b2pt_upload_scene: BVH clustering made no progress"; return B2PT_ERR_CUDA; }
            m = (int)hm;
            cur ^= 1;
            ++ploc_iters;
            build_launches += 5;
        }
        if (nactive > 0) {
            k_ploc_finish<<<1, B2PT_PLOC_FINISH / 2, 0, st>>>(m, nleaves, cur, cl_node[0], cl_node[1], cl_lo[0], cl_lo[1], cl_hi[0], cl_hi[1], d_nn, d_cnt, d_bn_child, d_nb_lo, d_nb_hi);
            ++build_launches;
            // 8-wide collapse, one launch per level of the wide tree; wide node 0 expands the binary root
            STAGE(cudaMemcpyAsync(d_wide_bin, cl_node[0], sizeof(int), cudaMemcpyDeviceToDevice, st));
        }
        int begin = 0, end = nactive > 0 ? 1 : 0;
        while (begin < end) {
            k_collapse_level<<<(end - begin + 127) / 128, 128, 0, st>>>(begin, end, nleaves, d_wide_bin, d_bn_child, d_nb_lo, d_nb_hi, d_cnt, d_wide);
            unsigned hw = 0;
            STAGE(cudaMemcpyAsync(&hw, d_cnt + CNT_NWIDE, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
            STAGE(cudaStreamSynchronize(st));
            begin = end; end = (int)hw;
            ++wide_levels; ++build_launches;
        }
        nwide = end;
        STAGE(cudaEventRecord(ctx->ev3, st));
        STAGE(cudaGetLastError());
        // hoisted leaves: ids and child codes
        for (int h = 0; h < nhoist; ++h) S.hoist_leaf[h] = (int)h_cnt[CNT_HOIST + h];
        std::sort(S.hoist_leaf, S.hoist_leaf + nhoist);   // deterministic order (results do not depend on it)
        for (int h = 0; h < nhoist; ++h) {
            float4 l{};
            STAGE(cudaMemcpyAsync(&l, d_nb_lo + S.hoist_leaf[h], sizeof(float4), cudaMemcpyDeviceToHost, st));
            STAGE(cudaStreamSynchronize(st));
            std::memcpy(&S.hoist_code[h], &l.w, sizeof(uint32_t));
        }
        // scene box
        for (int k = 0; k < 6; ++k) {
            unsigned u = h_cnt[CNT_SCENE + k];
            unsigned bits = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
            std::memcpy(&scene_box[k], &bits, sizeof(float));
        }
        ctx->stats.kernel_launches = build_launches;
    }
    S.nwide = nwide;
    S.nhoist = nhoist;
    // statistics for the occluder-aware child order: visits and hits per (wide node, slot), zeroed per upload
    unsigned* d_order_stats = nullptr;
    if ((rc = reserve(SL_ORDER_STATS, sizeof(unsigned) * 16 * (size_t)std::max(nwide, 1), &d_order_stats))) return rc;
    STAGE(cudaMemsetAsync(d_order_stats, 0, sizeof(unsigned) * 16 * (size_t)std::max(nwide, 1), st));
    ctx->d_order_stats = d_order_stats;
    ctx->order_state = (ctx->learn_order && nlight > 0 && nwide > 0) ? 0 : 2;
    if (nmat > 0) {
        std::vector<DMaterial> hm(nmat);
        for (int i = 0; i < nmat; ++i)
            hm[i] = DMaterial{mats[i].type, mats[i].albedo[0], mats[i].albedo[1], mats[i].albedo[2], mats[i].roughness, mats[i].metallic, mats[i].ior, 0.0f};
        STAGE(cudaMemcpyAsync(d_mats, hm.data(), sizeof(DMaterial) * nmat, cudaMemcpyHostToDevice, st));
        STAGE(cudaStreamSynchronize(st));   // hm goes out of scope
    }
    STAGE(cudaStreamSynchronize(st));
    if (ntri > 0) {
        float ms = 0.0f;
        STAGE(cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3));
        ctx->stats.build_seconds = ms * 1e-3;
    }
#undef STAGE

    S.tri = d_tri; S.nrm = d_nrm; S.node_lo = d_node_lo; S.node_hi = d_node_hi; S.node_info = d_info;
    S.leaf_lo = d_nb_lo; S.leaf_hi = d_nb_hi; S.wide = d_wide; S.mats = d_mats; S.nmat = nmat; S.nlight = nlight;
    // largest coordinate magnitude of any box plane (conservative slab arithmetic, traverse.cuh)
    float R = 0.0f;
    for (int k = 0; k < 6; ++k) R = std::max(R, std::fabs(scene_box[k]));
    if (!(R >= 1e-30f)) R = 1e-30f;
    if (!(R <= 3.0e38f)) R = 3.0e38f;
    S.coord_bound = R;
    for (int i = 0; i < nlight; ++i) {
        // Light ctor (scene.hpp:27-35): non-positive intensity is replaced by 1.0.
        float inten = lights[i].intensity;
        if (inten <= 0.0f) inten = 1.0f;
        S.lights[i] = DLight{lights[i].position[0], lights[i].position[1], lights[i].position[2],
                             lights[i].color[0], lights[i].color[1], lights[i].color[2], inten, 0.0f};
    }
    ctx->scene = S;
    ctx->has_scene = true;
    ctx->accel_info[0] = nwide; ctx->accel_info[1] = sizeof(WideNode); ctx->accel_info[2] = nleaves;
    ctx->accel_info[3] = nnodes; ctx->accel_info[4] = 48; ctx->accel_info[5] = ploc_iters; ctx->accel_info[6] = wide_levels; ctx->accel_info[7] = nhoist;
    return B2PT_OK;
}

// Occluder-aware child order.  The answer of an occlusion query does not depend on the order in which the passing
// children of a node are tried but its cost does: an occluded ray stops at its first accepted
// triangle.  During the first wavefront batch after an upload the instrumented shadow kernel (any_rtc_learn) counts
// visits and terminal hits per (wide node, slot); here the hits are folded up the wide tree (children follow their
// parent in the BFS order of the collapse) and every node's slots are re-ordered by increasing hits-per-visit — the
// occlusion kernels pop the LAST slot first.  Closest-hit kernels sort by entry distance and do not care about slot
// order.
int learn_child_order(b2pt_ctx* ctx) {
    ctx->order_state = 2;
    const int nwide = ctx->scene.nwide;
    if (!ctx->has_scene || nwide == 0 || !ctx->d_order_stats) return B2PT_OK;
    cudaStream_t st = ctx->stream;
    std::vector<unsigned> stats(16 * (size_t)nwide);
    std::vector<WideNode> nodes((size_t)nwide);
    B2PT_CUDA(ctx, cudaMemcpyAsync(stats.data(), ctx->d_order_stats, sizeof(unsigned) * stats.size(), cudaMemcpyDeviceToHost, st));
    B2PT_CUDA(ctx, cudaMemcpyAsync(nodes.data(), ctx->scene.wide, sizeof(WideNode) * nodes.size(), cudaMemcpyDeviceToHost, st));
    B2PT_CUDA(ctx, cudaStreamSynchronize(st));
    const unsigned* visits = stats.data();
    const unsigned* hits = stats.data() + 8 * (size_t)nwide;
    std::vector<double> subtree_hits(nwide, 0.0);
    std::vector<uint8_t> perm(8 * (size_t)nwide);
    bool any = false;
    for (int w = nwide - 1; w >= 0; --w) {
        double rate[8];
        int order[8];
        for (int s = 0; s < 8; ++s) {
            order[s] = s;
            const uint32_t code = nodes[w].child[s];
            if (code == B2PT_CHILD_EMPTY) { rate[s] = -1.0; continue; }            // empty slots first (never pass a box test)
            const double h = (code & B2PT_CHILD_LEAF) ? (double)hits[8 * (size_t)w + s] : subtree_hits[code];
            subtree_hits[w] += h;
            rate[s] = (h + 0.5) / ((double)visits[8 * (size_t)w + s] + 1.0);
        }
        std::stable_sort(order, order + 8, [&](int a, int b) { return rate[a] < rate[b]; });
        for (int s = 0; s < 8; ++s) { perm[8 * (size_t)w + s] = (uint8_t)order[s]; any |= order[s] != s; }
    }
    if (!any) return B2PT_OK;
    void* d_perm = nullptr;
    int rc = scratch_reserve(ctx, SL_PERM, perm.size(), &d_perm);
    if (rc) return rc;
    B2PT_CUDA(ctx, cudaMemcpyAsync(d_perm, perm.data(), perm.size(), cudaMemcpyHostToDevice, st));
    k_permute_slots<<<(nwide + 127) / 128, 128, 0, st>>>(nwide, static_cast<const uint8_t*>(d_perm), const_cast<WideNode*>(ctx->scene.wide));
    B2PT_CUDA(ctx, cudaGetLastError());
    B2PT_CUDA(ctx, cudaStreamSynchronize(st));   // perm is read by the copy above
    return B2PT_OK;
}

}  // namespace b2pt

// ---- host: reference BVH::build ordering (include/bvh.hpp:27-72) -------------------------------------
// Same algorithm and the same libstdc++ std::nth_element as the reference, run over (key, index)
// pairs: the permutation std::nth_element produces depends only on the comparator's answers, and
// the comparator here returns exactly `a.getCenter()[axis] < b.getCenter()[axis]` (bvh.hpp:64-65).
namespace {
struct KeyIdx { float key; int32_t idx; };

void order_rec(const float* pos, std::vector<KeyIdx>& work, int32_t* order, int64_t start, int64_t end, int par_depth) {
    int64_t count = end - start;
    if (count <= 8) return;
    // bounds of [start,end): fold of per-triangle min/max (bvh.hpp:48-52)
    float lo[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
    float hi[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    for (int64_t i = start; i < end; ++i) {
        const float* p = pos + 9ll * order[i];
        for (int a = 0; a < 3; ++a) {
            float mn = p[a], mx = p[a];
            if (p[3 + a] < mn) mn = p[3 + a];
            if (mx < p[3 + a]) mx = p[3 + a];
            if (p[6 + a] < mn) mn = p[6 + a];
            if (mx < p[6 + a]) mx = p[6 + a];
            if (mn < lo[a]) lo[a] = mn;
            if (hi[a] < mx) hi[a] = mx;
        }
    }
    // AABB::maxExtentAxis (aabb.hpp:34-39)
    float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
    int axis = (ex > ey && ex > ez) ? 0 : ((ey > ez) ? 1 : 2);
    int64_t mid = start + count / 2;
    for (int64_t i = start; i < end; ++i) {
        const float* p = pos + 9ll * order[i];
        // Triangle::getCenter (triangle.hpp:69-71): (v0 + v1 + v2) / 3.0f
        float c = ((p[axis] + p[3 + axis]) + p[6 + axis]) / 3.0f;
        work[i] = KeyIdx{c, order[i]};
    }
    std::nth_element(work.begin() + start, work.begin() + mid, work.begin() + end,
                     [](const KeyIdx& a, const KeyIdx& b) { return a.key < b.key; });
    for (int64_t i = start; i < end; ++i) order[i] = work[i].idx;
    // The two halves touch disjoint ranges of `work` and `order`: the top levels of the recursion run them on
    // separate threads (10M triangles: 2.9 s -> well under a second on 16 threads).  Which thread runs a range has
    // no influence on the result — std::nth_element sees exactly the same input either way.
    if (par_depth > 0 && count >= (1 << 16)) {
        std::thread left([&]() { order_rec(pos, work, order, start, mid, par_depth - 1); });
        order_rec(pos, work, order, mid, end, par_depth - 1);
        left.join();
    } else {
        order_rec(pos, work, order, start, mid, 0);
        order_rec(pos, work, order, mid, end, 0);
    }
}
}  // namespace

extern "C" int b2pt_reference_order(const float* pos, int64_t ntri, int32_t* order) {
    if (ntri < 0 || ntri >= (1ll << 28) || (ntri > 0 && (!pos || !order))) return B2PT_ERR_INVALID;
    for (int64_t i = 0; i < ntri; ++i) order[i] = static_cast<int32_t>(i);
    std::vector<KeyIdx> work(static_cast<size_t>(ntri));
    unsigned hw = std::thread::hardware_concurrency();
    int par_depth = 0;
    while ((1u << par_depth) < std::max(hw, 1u) && par_depth < 6) ++par_depth;
    order_rec(pos, work, order, 0, ntri, par_depth);
    return B2PT_OK;
}
