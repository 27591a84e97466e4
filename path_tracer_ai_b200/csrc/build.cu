// build.cu — scene upload and acceleration-structure build.
//
// Replaces OptixRenderer::uploadScene / buildAccelerationStructure (reference
// src/gpu/optix_renderer.cu:383-409, :233-353) and, on the device, the box computation of
// BVH::buildRecursive (include/bvh.hpp:44-52).
//
// The triangles arrive in the reference's post-build order, so the reference tree is implicit in
// the array: node = [start,end), mid = start + count/2, leaf iff count <= 8 (bvh.hpp:55-61).  Its
// topology depends only on the triangle count and is laid out on the host in O(nodes); every box is
// computed on the GPU (leaf boxes from the triangles, inner boxes level by level, bottom-up; min/max
// are exact so the result equals the reference's sequential fold).  The wide BVH used by the fast
// traversal kernels is an 8-ary collapse of that same tree: each wide node adopts up to 8 descendants
// at most three binary levels down (aligned to the bottom of the tree), with their exact boxes, so "box passes the reference slab
// test" is monotone from any reference leaf up through every wide ancestor.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <type_traits>
#include <vector>

#include "ctx.cuh"

namespace b2pt {

namespace {

#ifndef B2PT_SORT_CHILDREN
#define B2PT_SORT_CHILDREN 0
#endif

// scratch slots of the scene buffers and of the upload staging (persistent, grow-only)
enum { SL_TRI = 16, SL_NRM, SL_NODE_LO, SL_NODE_HI, SL_LEAF_LO, SL_LEAF_HI, SL_INFO, SL_WIDE, SL_MATS,
       SL_POS, SL_NIN, SL_MAT, SL_IDS, SL_WSRC, SL_WCHILD, SL_ORDER_STATS };
static_assert(SL_ORDER_STATS < B2PT_SCRATCH_SLOTS, "scratch slots");

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

// One thread per reference leaf: box = fold of Triangle::getAABB (triangle.hpp:73-77) with
// AABB::merge (aabb.hpp:27-32); tags the leaf's triangles with the leaf id.
__global__ void k_leaf_boxes(const float* __restrict__ pos, const int4* __restrict__ node_info, int nnodes,
                             float4* __restrict__ node_lo, float4* __restrict__ node_hi,
                             float4* __restrict__ leaf_lo, float4* __restrict__ leaf_hi, float4* __restrict__ tri) {
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
    float4 l = make_float4(lo[0], lo[1], lo[2], 0.0f), h = make_float4(hi[0], hi[1], hi[2], 0.0f);
    node_lo[i] = l; node_hi[i] = h;
    leaf_lo[info.w] = l; leaf_hi[info.w] = h;
}

// Inner nodes of one depth: box = merge(left, right).
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

// Fills the wide nodes: wide node w adopts the reference nodes wide_src[8*w .. 8*w+7] (-1 = none).  One thread
// per wide node.  B2PT_SORT_CHILDREN=1 stores the children by decreasing box surface area (slot order is free:
// closest-hit kernels sort by entry distance, occlusion kernels visit the last slot first, i.e. then the most
// compact subtree).  Measured: any-hit on random rays +5 % (5.78 -> 5.53 nodes, 6.31 -> 5.75 triangles per ray),
// renders unchanged within noise (Cornell 793 -> 787, 1M mesh 217 -> 217): off by default.  What the Cornell frame
// responds to is WHICH leaves are tried first (plain slot order reversed: +5 %) — an order learned from where
// occlusions are actually found is the next step.
__global__ void k_fill_wide(const int* __restrict__ wide_src, const uint32_t* __restrict__ wide_child, int nwide,
                            const float4* __restrict__ node_lo, const float4* __restrict__ node_hi,
                            WideNode* __restrict__ wide) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwide) return;
    int src[8];
    uint32_t code[8];
    float area[8];
    for (int s = 0; s < 8; ++s) {
        src[s] = wide_src[8 * w + s];
        code[s] = wide_child[8 * w + s];
        if (src[s] < 0) { area[s] = 3.402823466e+38f; continue; }
        float4 l = node_lo[src[s]], h = node_hi[src[s]];
        float dx = h.x - l.x, dy = h.y - l.y, dz = h.z - l.z;
        area[s] = dx * dy + dy * dz + dz * dx;
    }
#if B2PT_SORT_CHILDREN
    for (int i = 1; i < 8; ++i) {   // insertion sort, decreasing area (stable)
        int si = src[i]; uint32_t ci = code[i]; float ai = area[i];
        int j = i;
        while (j > 0 && area[j - 1] < ai) { src[j] = src[j - 1]; code[j] = code[j - 1]; area[j] = area[j - 1]; --j; }
        src[j] = si; code[j] = ci; area[j] = ai;
    }
#endif
    WideNode& nd = wide[w];
    for (int s = 0; s < 8; ++s) {
        if (src[s] < 0) {
            // inverted box: never passes the slab test
            nd.lox[s] = nd.loy[s] = nd.loz[s] = 3.402823466e+38f;
            nd.hix[s] = nd.hiy[s] = nd.hiz[s] = -3.402823466e+38f;
            nd.child[s] = B2PT_CHILD_EMPTY;
        } else {
            float4 l = node_lo[src[s]], h = node_hi[src[s]];
            nd.lox[s] = l.x; nd.loy[s] = l.y; nd.loz[s] = l.z;
            nd.hix[s] = h.x; nd.hiy[s] = h.y; nd.hiz[s] = h.z;
            nd.child[s] = code[s];
        }
    }
}

}  // namespace

void free_scene(b2pt_ctx* ctx) {
    // scene buffers are persistent scratch slots of the context (freed by b2pt_destroy); only forget the scene
    for (void* p : ctx->scene_allocs) cudaFree(p);
    ctx->scene_allocs.clear();
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

    // Every device buffer of the scene and of its staging is a persistent grow-only allocation of the context
    // (scratch slots 16..31): re-uploading a scene of the same size — the reference's GPU branch uploads once per
    // run, an interactive caller once per frame — costs no cudaMalloc/cudaFree (measured: 0.12-3.7 s per upload of
    // 1M triangles with them, next to a 5 GB wavefront allocation).

    // The topology of the reference tree and of its collapse depends on the triangle COUNT only: it is laid out on
    // the host once per count and its device copies (node_info, wide sources/codes, per-depth id lists) are reused.
    b2pt_ctx::Topology& T = ctx->topo;
    const bool topo_cached = T.valid && T.ntri == ntri;
    if (!topo_cached) {
        T = b2pt_ctx::Topology{};
        T.ntri = ntri;
        std::vector<int4>& info = T.info;
        std::vector<int>& wide_src = T.wide_src;
        std::vector<uint32_t>& wide_child = T.wide_child;
        int& nleaves = T.nleaves;
        int& maxdepth = T.maxdepth;
    // ---- host: topology of the reference tree and of its 8-ary collapse -------------------------
    std::vector<HostNode> nodes;
    if (ntri > 0) {
        nodes.reserve(static_cast<size_t>(ntri) / 2 + 16);
        layout_tree(0, ntri, 0, nodes, nleaves);
    }
    const int nnodes = static_cast<int>(nodes.size());
    info.assign(nnodes, int4{});
    maxdepth = 0;
    for (int i = 0; i < nnodes; ++i) {
        info[i] = make_int4(nodes[i].start, nodes[i].end, nodes[i].right, nodes[i].leaf);
        maxdepth = std::max(maxdepth, nodes[i].depth);
    }
    // inner nodes grouped by depth (deepest first) for the bottom-up box pass
    std::vector<std::vector<int>> by_depth(maxdepth + 1);
    for (int i = 0; i < nnodes; ++i) if (nodes[i].leaf < 0) by_depth[nodes[i].depth].push_back(i);

    // wide collapse: BFS over wide nodes; each adopts descendants three binary levels down.
    std::vector<int> wide_of;          // reference node -> wide node index (for inner children), filled lazily
    if (nnodes > 0) {
        std::vector<int> queue;   // reference node index of each wide node, in wide order
        auto leaf_code = [&](int ref) {
            return B2PT_CHILD_LEAF | (static_cast<uint32_t>(nodes[ref].end - nodes[ref].start - 1) << 28) | static_cast<uint32_t>(nodes[ref].start);
        };
        if (nodes[0].leaf >= 0) {
            // Degenerate tree (<= 8 triangles): a root wide node with a single leaf child.
            wide_src.assign(8, -1); wide_child.assign(8, B2PT_CHILD_EMPTY);
            wide_src[0] = 0; wide_child[0] = leaf_code(0);
        } else {
            // Height (distance to the deepest leaf below) of every reference node; children follow their parent
            // in pre-order, so one reverse sweep suffices.  The collapse is aligned to the BOTTOM of the tree: a
            // descendant is adopted as soon as its height is a multiple of 3 (or after three levels), so the
            // nodes just above the leaves — the most numerous and the most visited — are full 8-wide and any
            // partial fan-out sits at the root.  (Top-aligned, a 1M-triangle tree of depth 17 had 4.5 children
            // per wide node: the whole bottom level tested 4 empty slots per visit.)
            std::vector<int> height(nnodes, 0);
            for (int i = nnodes - 1; i >= 0; --i)
                if (nodes[i].leaf < 0) height[i] = 1 + std::max(height[i + 1], height[nodes[i].right]);
            queue.push_back(0);
            std::vector<int> wlevel(1, 0);   // wide level of each wide node (root = 0)
            for (size_t w = 0; w < queue.size(); ++w) {
                int ref = queue[w];
                T.max_wide_level = std::max(T.max_wide_level, wlevel[w]);
                int cur[8], ncur = 2;
                cur[0] = ref + 1; cur[1] = nodes[ref].right;
                for (int level = 0; level < 2; ++level) {
                    int nxt[8], nn = 0;
                    for (int k = 0; k < ncur; ++k) {
                        int c = cur[k];
                        if (nodes[c].leaf >= 0 || height[c] % 3 == 0) nxt[nn++] = c;
                        else { nxt[nn++] = c + 1; nxt[nn++] = nodes[c].right; }
                    }
                    ncur = nn;
                    std::memcpy(cur, nxt, sizeof(int) * nn);
                }
                for (int s = 0; s < 8; ++s) {
                    if (s >= ncur) { wide_src.push_back(-1); wide_child.push_back(B2PT_CHILD_EMPTY); continue; }
                    int c = cur[s];
                    wide_src.push_back(c);
                    if (nodes[c].leaf >= 0) wide_child.push_back(leaf_code(c));
                    else { wide_child.push_back(static_cast<uint32_t>(queue.size())); queue.push_back(c); wlevel.push_back(wlevel[w] + 1); }
                }
            }
        }
    }
        // inner nodes by depth, deepest first, flattened: the bottom-up box pass launches one kernel per span
        for (int dpt = maxdepth; dpt >= 0; --dpt) {
            T.spans.push_back({T.ids_flat.size(), by_depth[dpt].size()});
            T.ids_flat.insert(T.ids_flat.end(), by_depth[dpt].begin(), by_depth[dpt].end());
        }
        T.nnodes = nnodes;
        T.valid = true;
    }
    if (T.max_wide_level > B2PT_MAX_WIDE_LEVEL) {
        ctx->err = "b2pt_upload_scene: BVH deeper than the traversal stacks allow (wide level " + std::to_string(T.max_wide_level) + ")";
        T.valid = false;
        return B2PT_ERR_INVALID;
    }
    const int nnodes = T.nnodes, nleaves = T.nleaves;
    const std::vector<int4>& info = T.info;
    const std::vector<int>& wide_src = T.wide_src;
    const std::vector<uint32_t>& wide_child = T.wide_child;
    const int nwide = static_cast<int>(wide_src.size() / 8);
    S.nnodes = nnodes; S.nleaves = nleaves; S.nwide = nwide;

    // ---- device buffers ------------------------------------------------------------------------------
    float4 *d_tri, *d_nrm, *d_node_lo, *d_node_hi, *d_leaf_lo, *d_leaf_hi;
    int4* d_info; WideNode* d_wide; DMaterial* d_mats;
    int rc;
    auto reserve = [&](int slot, size_t bytes, auto** out) {
        void* p = nullptr;
        int r = scratch_reserve(ctx, slot, std::max<size_t>(bytes, 16), &p);
        *out = static_cast<std::remove_reference_t<decltype(**out)>*>(p);
        return r;
    };
    if ((rc = reserve(SL_TRI, sizeof(float4) * 3ull * ntri, &d_tri))) return rc;
    if ((rc = reserve(SL_NRM, sizeof(float4) * 3ull * ntri, &d_nrm))) return rc;
    if ((rc = reserve(SL_NODE_LO, sizeof(float4) * (size_t)nnodes, &d_node_lo))) return rc;
    if ((rc = reserve(SL_NODE_HI, sizeof(float4) * (size_t)nnodes, &d_node_hi))) return rc;
    if ((rc = reserve(SL_LEAF_LO, sizeof(float4) * (size_t)nleaves, &d_leaf_lo))) return rc;
    if ((rc = reserve(SL_LEAF_HI, sizeof(float4) * (size_t)nleaves, &d_leaf_hi))) return rc;
    if ((rc = reserve(SL_INFO, sizeof(int4) * (size_t)nnodes, &d_info))) return rc;
    if ((rc = reserve(SL_WIDE, sizeof(WideNode) * (size_t)nwide, &d_wide))) return rc;
    if ((rc = reserve(SL_MATS, sizeof(DMaterial) * (size_t)std::max(nmat, 1), &d_mats))) return rc;
    // statistics for the occluder-aware child order: visits and hits per (wide node, slot), zeroed per upload
    unsigned* d_order_stats = nullptr;
    if ((rc = reserve(SL_ORDER_STATS, sizeof(unsigned) * 16 * (size_t)std::max(nwide, 1), &d_order_stats))) return rc;
    {
        cudaError_t e__ = cudaMemsetAsync(d_order_stats, 0, sizeof(unsigned) * 16 * (size_t)std::max(nwide, 1), st);
        if (e__ != cudaSuccess) { cuda_fail(ctx, e__, "cudaMemsetAsync(order stats)", __FILE__, __LINE__); return B2PT_ERR_CUDA; }
    }
    ctx->d_order_stats = d_order_stats;
    {
        const char* env = std::getenv("B2PT_LEARN_ORDER");
        const bool want = env ? std::atoi(env) != 0 : true;
        ctx->order_state = (want && nlight > 0 && nwide > 0) ? 0 : 2;
    }

    float *d_pos = nullptr, *d_nin = nullptr; int32_t* d_mat = nullptr; int *d_ids = nullptr, *d_wsrc = nullptr; uint32_t* d_wchild = nullptr;
    auto cleanup = [&]() {};
#define STAGE(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); cuda_fail(ctx, e__, #call, __FILE__, __LINE__); return B2PT_ERR_CUDA; } } while (0)
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
        if ((rc = reserve(SL_WSRC, sizeof(int) * wide_src.size(), &d_wsrc))) return rc;
        if ((rc = reserve(SL_WCHILD, sizeof(uint32_t) * wide_child.size(), &d_wchild))) return rc;
        if ((rc = reserve(SL_IDS, sizeof(int) * std::max<size_t>(ninner, 1), &d_ids))) return rc;
        if (!topo_cached || !T.on_device) {
            STAGE(cudaMemcpyAsync(d_info, info.data(), sizeof(int4) * nnodes, cudaMemcpyHostToDevice, st));
            STAGE(cudaMemcpyAsync(d_wsrc, wide_src.data(), sizeof(int) * wide_src.size(), cudaMemcpyHostToDevice, st));
            STAGE(cudaMemcpyAsync(d_wchild, wide_child.data(), sizeof(uint32_t) * wide_child.size(), cudaMemcpyHostToDevice, st));
            if (ninner) STAGE(cudaMemcpyAsync(d_ids, T.ids_flat.data(), sizeof(int) * ninner, cudaMemcpyHostToDevice, st));
            T.on_device = true;
        }

        // ---- device build, timed -----------------------------------------------------------------
        STAGE(cudaEventRecord(ctx->ev2, st));
        const int B = 256;
        k_pack_triangles<<<(ntri + B - 1) / B, B, 0, st>>>(d_pos, d_nin, d_mat, ntri, d_tri, d_nrm);
        k_leaf_boxes<<<(nnodes + B - 1) / B, B, 0, st>>>(d_pos, d_info, nnodes, d_node_lo, d_node_hi, d_leaf_lo, d_leaf_hi, d_tri);
        int launches = 2;
        for (auto& sp : T.spans) {
            if (!sp.second) continue;
            k_inner_boxes<<<(static_cast<int>(sp.second) + B - 1) / B, B, 0, st>>>(d_ids + sp.first, static_cast<int>(sp.second), d_info, d_node_lo, d_node_hi);
            ++launches;
        }
        k_fill_wide<<<(nwide + B - 1) / B, B, 0, st>>>(d_wsrc, d_wchild, nwide, d_node_lo, d_node_hi, d_wide);
        ++launches;
        STAGE(cudaEventRecord(ctx->ev3, st));
        STAGE(cudaGetLastError());
        ctx->stats.kernel_launches = launches;
    }
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
    S.leaf_lo = d_leaf_lo; S.leaf_hi = d_leaf_hi; S.wide = d_wide; S.mats = d_mats; S.nmat = nmat; S.nlight = nlight;
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
    ctx->accel_info[3] = nnodes; ctx->accel_info[4] = 48;
    return B2PT_OK;
}

// Occluder-aware child order.  The answer of an occlusion query does not depend on the order in which the passing
// children of a node are tried but its cost does: an occluded ray stops at its first accepted
// triangle.  During the first wavefront batch after an upload the instrumented shadow kernel (any_rtc_learn) counts
// visits and terminal hits per (wide node, slot); here the hits are folded up the wide tree (children follow their
// parent in the BFS order of the collapse) and every node's slots are re-ordered by increasing hits-per-visit — the
// occlusion kernels pop the LAST slot first.  An offline count on real shadow rays puts the saving at 22 % (Cornell)
// / 11 % (mesh) of the triangle tests (profiles/r01_anyhit_order_study.txt).  Closest-hit kernels sort by entry
// distance and do not care about slot order.
int learn_child_order(b2pt_ctx* ctx) {
    b2pt_ctx::Topology& T = ctx->topo;
    ctx->order_state = 2;
    const int nwide = static_cast<int>(T.wide_src.size() / 8);
    if (!ctx->has_scene || !T.valid || nwide == 0 || !ctx->d_order_stats) return B2PT_OK;
    cudaStream_t st = ctx->stream;
    std::vector<unsigned> stats(16 * (size_t)nwide);
    B2PT_CUDA(ctx, cudaMemcpyAsync(stats.data(), ctx->d_order_stats, sizeof(unsigned) * stats.size(), cudaMemcpyDeviceToHost, st));
    B2PT_CUDA(ctx, cudaStreamSynchronize(st));
    const unsigned* visits = stats.data();
    const unsigned* hits = stats.data() + 8 * (size_t)nwide;
    std::vector<double> subtree_hits(nwide, 0.0);
    bool any = false;
    for (int w = nwide - 1; w >= 0; --w) {
        double rate[8];
        int order[8];
        for (int s = 0; s < 8; ++s) {
            order[s] = s;
            const int src = T.wide_src[8 * w + s];
            const uint32_t code = T.wide_child[8 * w + s];
            if (src < 0) { rate[s] = -1.0; continue; }            // empty slots first (never pass a box test)
            const double h = (code & B2PT_CHILD_LEAF) ? (double)hits[8 * (size_t)w + s] : subtree_hits[code];
            subtree_hits[w] += h;
            rate[s] = (h + 0.5) / ((double)visits[8 * (size_t)w + s] + 1.0);
        }
        std::stable_sort(order, order + 8, [&](int a, int b) { return rate[a] < rate[b]; });
        int src8[8]; uint32_t code8[8];
        for (int s = 0; s < 8; ++s) { src8[s] = T.wide_src[8 * w + order[s]]; code8[s] = T.wide_child[8 * w + order[s]]; any |= order[s] != s; }
        for (int s = 0; s < 8; ++s) { T.wide_src[8 * w + s] = src8[s]; T.wide_child[8 * w + s] = code8[s]; }
    }
    if (!any) return B2PT_OK;
    int* d_wsrc = static_cast<int*>(ctx->scratch[SL_WSRC]);
    uint32_t* d_wchild = static_cast<uint32_t*>(ctx->scratch[SL_WCHILD]);
    B2PT_CUDA(ctx, cudaMemcpyAsync(d_wsrc, T.wide_src.data(), sizeof(int) * T.wide_src.size(), cudaMemcpyHostToDevice, st));
    B2PT_CUDA(ctx, cudaMemcpyAsync(d_wchild, T.wide_child.data(), sizeof(uint32_t) * T.wide_child.size(), cudaMemcpyHostToDevice, st));
    k_fill_wide<<<(nwide + 255) / 256, 256, 0, st>>>(d_wsrc, d_wchild, nwide, static_cast<const float4*>(ctx->scratch[SL_NODE_LO]),
                                                    static_cast<const float4*>(ctx->scratch[SL_NODE_HI]), static_cast<WideNode*>(ctx->scratch[SL_WIDE]));
    B2PT_CUDA(ctx, cudaGetLastError());
    B2PT_CUDA(ctx, cudaStreamSynchronize(st));   // the host vectors are read by the copies above
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
