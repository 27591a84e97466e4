// ctx.cuh — device-side scene layout and the host context shared by the translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <utility>
#include <vector>

#include "../../include/b2pt.h"
#include "exact.cuh"

namespace b2pt {

// ---- HBM layout -----------------------------------------------------------------------------------
// Triangles (reference post-build order, index == reference triangle id):
//   tri[3*i+0] = (v0.xyz, as_float(leaf id))   tri[3*i+1] = (e1.xyz, 0)   tri[3*i+2] = (e2.xyz, 0)
//   nrm[3*i+0] = (n0.xyz, as_float(material))  nrm[3*i+1] = (n1.xyz, 0)   nrm[3*i+2] = (n2.xyz, 0)
// 48 B each, 16-byte vector loads.
//
// Reference tree (include/bvh.hpp:44-72), implicit in the order: node = range [start,end),
// mid = start + count/2, leaf iff count <= 8.  Stored in DFS pre-order (left child = i+1):
//   node_lo[i], node_hi[i] = exact fp32 bounds (float4, w unused)
//   node_info[i] = (start, end, right child index or -1 for leaves, leaf id or -1)
//   leaf_lo/leaf_hi[l]    = the same boxes, indexed by leaf id (what visibility is defined on)
//
// Traversal tree (8-wide BVH over the reference LEAVES, built on the device: build.cu; full-precision child
// boxes, SoA inside the node so that a lane reads four children's planes with one 16-byte load):
struct __align__(32) WideNode {
    float lox[8], loy[8], loz[8];
    float hix[8], hiy[8], hiz[8];
    uint32_t child[8];   // 0xFFFFFFFF empty | leaf: bit31, count-1 in [30:28], first tri in [27:0] | inner: node index
};
static_assert(sizeof(WideNode) == 224, "WideNode layout");

#define B2PT_CHILD_EMPTY 0xFFFFFFFFu
#define B2PT_CHILD_LEAF 0x80000000u

struct DMaterial { int type; float r, g, b, roughness, metallic, ior, pad; };
struct DLight { float px, py, pz, cr, cg, cb, intensity, pad; };

#define B2PT_MAX_LIGHTS 16
// Reference leaves whose box covers most of the scene (the loader's 16-unit room triangles, src/scene.cpp:118-209, sit
// in one or two of them) are kept OUT of the traversal tree: every ray would visit them anyway, so each ray tests
// them once, up front, while the lanes of its warp are still in lock-step (traverse_rtc.cuh).
#define B2PT_MAX_HOIST 4

struct DeviceScene {
    int ntri;
    int nnodes;
    int nleaves;
    int nwide;
    const float4* tri;
    const float4* nrm;
    const float4* node_lo;
    const float4* node_hi;
    const int4* node_info;
    const float4* leaf_lo;
    const float4* leaf_hi;
    const WideNode* wide;
    const DMaterial* mats;
    int nmat;
    int nlight;
    float coord_bound;   // largest |coordinate| in the scene (absolute part of the distance-cull slack, traverse.cuh)
    int nhoist;                              // reference leaves tested by every ray before the traversal
    int hoist_leaf[B2PT_MAX_HOIST];          // their leaf ids ...
    uint32_t hoist_code[B2PT_MAX_HOIST];     // ... and child codes (first triangle, count)
    DLight lights[B2PT_MAX_LIGHTS];
};

struct HitRec { float t; int tri; float u, v; };

// Per-call traversal counters (device): [0] fallback rays, [1] node fetches, [2] tri fetches.
struct TraceCounters { unsigned long long fallback, node_fetches, tri_fetches, pad; };

}  // namespace b2pt

// ---- host context -----------------------------------------------------------------------------------
#define B2PT_SCRATCH_SLOTS 64   // 0..15: per-call scratch (trace / render), 16..: scene buffers, upload staging, build scratch
struct b2pt_ctx {
    int device = 0;
    int flags = 0;
    int sm_count = 0;
    int64_t max_paths = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    // render_frame: the exact recursion of the uncertified rays runs on `side` while the main stream sorts the bounce
    // (fork / join through ev_fork / ev_join); h_count = pinned word the per-bounce active count is read back into
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int* h_count = nullptr;
    std::vector<cudaEvent_t> ev_pool;    // per-bounce timing events of render_frame, created once and reused
    bool debug_sync = false;             // B2PT_DEBUG_SYNC=1 at b2pt_create: synchronise after every render kernel
    std::string err;
    bool has_scene = false;
    // occluder-aware child order (build.cu learn_child_order): 0 = collect statistics during the next wavefront
    // batch, 1 = collected (re-order before the next batch), 2 = done / off
    int order_state = 2;
    bool learn_order = true;             // cleared by B2PT_FLAG_NO_LEARN_ORDER
    unsigned* d_order_stats = nullptr;   // visits[nwide*8] then hits[nwide*8]
    b2pt::DeviceScene scene{};
    // scratch (grown on demand)
    void* scratch[B2PT_SCRATCH_SLOTS] = {};
    size_t scratch_bytes[B2PT_SCRATCH_SLOTS] = {};
    // host topology of the reference tree, cached by triangle count (build.cu)
    struct Topology {
        bool valid = false, on_device = false;
        int ntri = -1, nnodes = 0, nleaves = 0, maxdepth = 0;
        std::vector<int4> info;
        std::vector<int> ids_flat;
        std::vector<std::pair<size_t, size_t>> spans;
    } topo;
    b2pt::TraceCounters* d_counters = nullptr;
    int* d_fallback_count = nullptr;
    b2pt_stats stats{};
    int64_t accel_info[8] = {};
    // output stage: byte thresholds of the last gamma used (render.cu tonemap_thresholds)
    float tonemap_gamma = 0.0f;
    float tonemap_thr[256] = {};
    // the frame of the last b2pt_render / progressive pass stays on the device (scratch slot 7) for b2pt_tonemap_frame
    int last_width = 0, last_height = 0;
    // progressive rendering (b2pt_progressive_begin / _pass)
    bool prog_active = false;
    b2pt_camera prog_cam{};
    b2pt_settings prog_settings{};
    uint64_t prog_seed = 0;
    int prog_done = 0;
};

namespace b2pt {

// Error plumbing: CUDA failures become "<call> failed: <cuda error> (<file>:<line>)", the text the
// reference's CUDA_CHECK would throw (include/gpu/cuda_utils.hpp:16-25).
bool cuda_fail(b2pt_ctx* ctx, cudaError_t e, const char* call, const char* file, int line);
#define B2PT_CUDA(ctx, call)                                                             \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) { b2pt::cuda_fail((ctx), e__, #call, __FILE__, __LINE__); return B2PT_ERR_CUDA; } \
    } while (0)

// Scratch buffer `slot`, at least `bytes` large (contents undefined).
int scratch_reserve(b2pt_ctx* ctx, int slot, size_t bytes, void** out);

// build.cu
int build_scene(b2pt_ctx* ctx, const float* pos, const float* nrm, const int32_t* mat, int64_t ntri,
                const b2pt_material* mats, int32_t nmat, const b2pt_light* lights, int32_t nlight);
void free_scene(b2pt_ctx* ctx);
// Re-orders every wide node's child slots by the occlusion rate (hits per visit) the instrumented shadow kernel
// measured, so that occlusion queries try the likeliest occluders first.  No effect on any result.  Synchronises.
int learn_child_order(b2pt_ctx* ctx);

// trace.cu — all pointers device; launches on ctx->stream; counters accumulate into ctx->d_counters.
int launch_trace_closest(b2pt_ctx* ctx, const float* d_o, const float* d_d, const float* d_tmax, int64_t n,
                         int32_t* d_tri, float* d_t, float* d_uv);
int launch_trace_any(b2pt_ctx* ctx, const float* d_o, const float* d_d, const float* d_tmax, int64_t n,
                     uint8_t* d_occ);

// render.cu
int render_frame(b2pt_ctx* ctx, const b2pt_camera* cam, const b2pt_settings* st, uint64_t seed,
                 const b2pt_partition* part, float* d_rgb, bool keep_accum = false, int divisor = 0);
int tonemap_thresholds(float gamma, float* thr256);
int tonemap_frame(b2pt_ctx* ctx, const float* d_rgb, int width, int height, float gamma, int flip, uint8_t* rgb8_host);

}  // namespace b2pt
