// render.cu — wavefront path tracer: replaces Renderer::render / tracePath /
// calculateDirectLighting (reference include/renderer.hpp:40-102, :129-250, :252-301) and the
// OptiX raygen / closest-hit programs (src/gpu/ptx/optix_kernels.cu:49-257).
//
// One frame = pixel chunks x sample chunks of at most `max_paths` camera paths.  Per chunk:
//   k_raygen                         camera rays (camera.hpp:18-29), T = 1, L = 0
//   for depth in 0 .. maxBounces-1, while paths survive:
//     k_extend_rtc                    closest hit, one ray per thread; writes the hit record and, in scenes whose
//                                     bounces are sorted, the (Morton key of the hit point, path slot) pair
//     k_extend_fallback               uncertified rays -> exact reference recursion (side stream, under the sort)
//     cub::DeviceRadixSort            the bounce in the Morton order of its hit points (large scenes)
//     k_hitinfo                       hit point, shading normal, material; bins the path into its material
//                                     queue (1-pass counting sort on the material type) and into the
//                                     direct-light queue (fused into k_extend_rtc on small trees, nothing sorted)
//     k_shadow_rtc                    one ray per (vertex, light): any-hit traversal -> visibility byte
//     k_shade                         direct light from the visible lights (summed in light order),
//                                     L += T*direct, BSDF sample, T update, next ray -> next queue (grouped by octant)
//     (host)                          reads the number of surviving paths: grids and sort sizes of the next depth
//   k_resolve                        per pixel: samples added in sample order (renderer.hpp:69-72)
// k_finalize divides by spp (renderer.hpp:75-81).
//
// Path state: eight float4 fields per path slot p = s_local * npix_chunk + pixel_local (sample-major: neighbouring
// slots are neighbouring pixels), one 128-byte line per slot in sorted scenes, one array per field in small ones
// (struct PathField); queues hold path slots.  All arithmetic that decides a path (hit, direction, throughput) uses
// the exact.cuh operations, and the RNG is Philox4x32-10 keyed by (seed; pixel, sample, depth, draw): the image is a
// pure function of (scene, camera, settings, seed), independent of chunking, queue order, sorting and GPU count.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdlib>

#include <cub/device/device_radix_sort.cuh>

#include "traverse_pool.cuh"
#include "traverse_rtc.cuh"

namespace b2pt {

namespace {

enum { C_ACTIVE_A = 0, C_ACTIVE_B = 1, C_MAT0 = 2, C_MAT1 = 3, C_MAT2 = 4, C_SHADOW = 5, C_FALLBACK = 6, C_NEXT = 7, C_NCOUNTERS = 8 };

// Path state: eight float4 fields per path slot, in one of two layouts chosen per frame (render_frame):
//  * line layout (scenes whose bounces are sorted): ONE 128-byte line per slot, [ro | rd] [thr | rad] [g0 | g1] [hit | -].
//    Queue order stops following slot order after the first bounce (material binning, then the hit-point sort), so
//    every kernel reaches its paths through an index: with one array per field each 16-byte access pulled in a 32-byte
//    DRAM sector and used half of it; here the fields a kernel reads or writes together share a sector (closest hit:
//    ro+rd; k_shade: rd, thr+rad, g0+g1; k_hitinfo writes g0+g1) and all of a path's sectors share a line
//    (1M-triangle batch 409.7 -> 433.4 Msamples/s);
//  * array layout (small scenes, fused epilogue): one array per field — there the queues stay close to slot order and
//    neighbouring lanes read neighbouring 16-byte elements (Cornell 804 vs 779 Msamples/s with lines).
#define B2PT_PATH_FIELDS 8
struct PathField {
    float4* base;
    int stride;   // field of slot p = base[stride * p]: 8 (line layout) or 1 (array layout)
    __device__ __forceinline__ float4& operator[](long long p) const { return base[stride * p]; }
};
struct Wave {
    PathField ro, rd;      // ray origin / direction (direction already normalised by the Ray ctor rule)
    PathField hit;         // closest hit of the current bounce: (t, tri id, u, v)
    PathField g0, g1;      // (P, material id) / (shading normal, 0)
    uint8_t *vis;          // per (path, light): 1 = shadow ray occluded (own array in both layouts: in the slot's line it measured slower)
    PathField thr, rad;    // throughput T, radiance L
    int *q_active[2];      // active paths, ping-pong by depth parity
    int *q_mat[3];         // per-material queues
    int *q_shadow;         // vertices that need direct light (diffuse + specular)
    int *q_fallback;       // closest-hit queries the fast traversal could not certify
    int *counters;         // C_* above
    unsigned long long* totals;   // [0] extend rays, [1] shadow rays, [2] fallback rays, [4] extend work pool, [5] shadow work pool
};

struct CamConst {
    V3 pos, llc, horizontal, vertical;
};

struct FrameConst {
    int width, height;
    int spp_total;
    int max_bounces;
    uint32_t k0, k1;
    int tile_rank, tile_world, tile_area;
};

__device__ __forceinline__ long long pixel_of(const FrameConst& F, long long j) {
    if (F.tile_world > 1) return ((j / F.tile_area) * F.tile_world + F.tile_rank) * (long long)F.tile_area + (j % F.tile_area);
    return j;
}

__device__ __forceinline__ V3 f4v(float4 a) { return mk3(a.x, a.y, a.z); }

// renderer.hpp:308-319: rejection-sample the cube, NORMALISE the accepted point.
__device__ __forceinline__ V3 random_in_unit_sphere(uint32_t pix, uint32_t smp, uint32_t depth, uint32_t k0, uint32_t k1) {
    for (uint32_t k = 0;; ++k) {
        uint4 r = philox4x32_10(pix, smp, depth, DRAW_SPHERE0 + k, k0, k1);
        V3 p = vsub(vsmul(2.0f, mk3(u01(r.x), u01(r.y), u01(r.z))), mk3(1.0f, 1.0f, 1.0f));
        if (vdot(p, p) < 1.0f) return vnormalize(p);
    }
}

__global__ void __launch_bounds__(256) k_raygen(Wave W, CamConst C, FrameConst F, long long pix_begin, int npc, int s_begin, int P) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    int s_local = p / npc, jl = p - s_local * npc;
    long long i = pixel_of(F, pix_begin + jl);
    int x = (int)(i % F.width), y = (int)(i / F.width);
    uint32_t smp = (uint32_t)(s_begin + s_local);
    uint4 j = philox4x32_10((uint32_t)i, smp, 0u, DRAW_JITTER, F.k0, F.k1);
    float u = B2PT_DIV(B2PT_ADD((float)x, u01(j.x)), (float)(F.width - 1));     // renderer.hpp:63
    float v = B2PT_DIV(B2PT_ADD((float)y, u01(j.y)), (float)(F.height - 1));    // renderer.hpp:64
    // camera.hpp:28: normalize(llc + u*horizontal + v*vertical - position), then the Ray ctor again
    V3 dir = vnormalize(vsub(vadd(vadd(C.llc, vsmul(u, C.horizontal)), vsmul(v, C.vertical)), C.pos));
    dir = vnormalize(dir);
    W.ro[p] = make_float4(C.pos.x, C.pos.y, C.pos.z, 0.0f);
    W.rd[p] = make_float4(dir.x, dir.y, dir.z, 0.0f);
    W.thr[p] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
    W.rad[p] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

// After a closest hit: hit point, shading normal (normalised three times: triangle.hpp:62,
// intersection.hpp:18, renderer.hpp:139), material lookup (renderer.hpp:141-148), binning.
struct Epilogue { bool m0, m1, m2; unsigned lights; };   // lights: bit l = vertex needs a shadow ray towards light l

__device__ __forceinline__ Epilogue hit_epilogue(const DeviceScene& S, const Wave& W, int p, V3 o, V3 d, const HitRec& h) {
    Epilogue e{false, false, false, 0u};
    if (h.tri < 0) return e;   // miss: black background, path ends (renderer.hpp:135-137)
    float4 a = __ldg(&S.nrm[3ll * h.tri + 0]), b = __ldg(&S.nrm[3ll * h.tri + 1]), c = __ldg(&S.nrm[3ll * h.tri + 2]);
    float w = B2PT_SUB(B2PT_SUB(1.0f, h.u), h.v);
    V3 n = vadd(vadd(vsmul(w, f4v(a)), vsmul(h.u, f4v(b))), vsmul(h.v, f4v(c)));
    n = vnormalize(vnormalize(vnormalize(n)));
    V3 P = vadd(o, vmuls(d, h.t));   // ray.hpp:14-16
    int mat = __float_as_int(a.w);
    if (mat < 0 || mat >= S.nmat) {
        // invalid material id: magenta, no further bounce
        float4 T = W.thr[p], L = W.rad[p];
        V3 add = vmul(f4v(T), mk3(1.0f, 0.0f, 1.0f));
        W.rad[p] = make_float4(B2PT_ADD(L.x, add.x), B2PT_ADD(L.y, add.y), B2PT_ADD(L.z, add.z), 0.0f);
        return e;
    }
    int type = S.mats[mat].type;
    W.g0[p] = make_float4(P.x, P.y, P.z, __int_as_float(mat));
    W.g1[p] = make_float4(n.x, n.y, n.z, 0.0f);
    e.m0 = type == B2PT_DIFFUSE; e.m1 = type == B2PT_SPECULAR; e.m2 = type == B2PT_DIELECTRIC;
    if (e.m0 || e.m1) {
        // calculateDirectLighting (renderer.hpp:258-299) traces a shadow ray towards every light and only then
        // multiplies by cosTheta = max(dot(n, l), 0).  A light at or below the horizon (dot <= 0) therefore
        // contributes exactly zero whatever the ray finds — +-0 added to the sum, or a NaN the validity check
        // drops (:295) — and so does a light closer than 1e-4 (:263).  Those rays are not traced: the light is
        // marked "nothing to add" in vis and the image is bit-identical.  Same n, same normalised lightDir, same
        // dot as direct_lighting() computes later, so the two decisions cannot disagree.
        uint8_t* vis = W.vis + (long long)p * S.nlight;
        for (int l = 0; l < S.nlight; ++l) {
            const DLight& lt = S.lights[l];
            V3 lightDir = vsub(mk3(lt.px, lt.py, lt.pz), P);
            float dist = vlength(lightDir);
            bool trace = false;
            if (!(dist < 0.0001f)) trace = !(vdot(n, vnormalize(lightDir)) <= 0.0f);
            if (trace) e.lights |= 1u << l; else vis[l] = 1;
        }
    }
    return e;
}

// Block-aggregated queue appends (256-thread blocks, every thread must call).  One atomic per block and
// destination instead of one per warp: the per-warp version issued ~1M atomics per launch on a single
// address, and same-address atomics retire at roughly one per L2 clock — that alone was ~1 ms per launch.
#define B2PT_BIN_BLOCK 256
#define B2PT_BIN_WARPS (B2PT_BIN_BLOCK / 32)

// Append with the block's entries grouped by `bin` (0..7, -1 = nothing to append): k_shade groups the next
// rays of its 256 neighbouring vertices by direction octant, so a warp of the next closest-hit launch holds rays that
// start close together AND head the same way.
__device__ __forceinline__ void block_append_binned(int* counter, int* queue, int bin, int value) {
    __shared__ int s_cnt[8 * B2PT_BIN_WARPS];
    __shared__ int s_base;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x < 8 * B2PT_BIN_WARPS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const unsigned m = __match_any_sync(0xffffffffu, bin);
    if (bin >= 0 && lane == __ffs(m) - 1) s_cnt[bin * B2PT_BIN_WARPS + w] = __popc(m);
    __syncthreads();
    if (w == 0) {   // exclusive scan of the 64 (bin, warp) counts: two per lane + a warp scan
        static_assert(8 * B2PT_BIN_WARPS == 64, "two entries per lane");
        const int v0 = s_cnt[2 * lane], v1 = s_cnt[2 * lane + 1];
        int incl = v0 + v1;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += t; }
        const int excl = incl - v0 - v1;
        s_cnt[2 * lane] = excl; s_cnt[2 * lane + 1] = excl + v0;
        if (lane == 31) s_base = incl ? atomicAdd(counter, incl) : 0;
    }
    __syncthreads();
    if (bin >= 0) queue[s_base + s_cnt[bin * B2PT_BIN_WARPS + w] + __popc(m & ((1u << lane) - 1u))] = value;
}

// The one-pass counting sort after a closest hit: the path goes into the queue of its material type, and one
// entry p*nlight + l per needed shadow ray goes into the shadow queue.  Shadow entries of a block are written
// light-major — (light, warp, lane) — so consecutive entries are neighbouring vertices aiming at the same
// light (coherent warps in the shadow kernels), and the runs of one vertex group sit next to each other (their
// g0/g1 are fetched from DRAM once).
template <int WARPS>
__device__ __forceinline__ void bin_path(const Wave& W, const Epilogue& e, int p, int nlight) {
    constexpr int MAXROWS = 3 + B2PT_MAX_LIGHTS;
    static_assert(MAXROWS * WARPS <= 160, "the scan below covers 5 entries per lane");
    __shared__ int s_cnt[MAXROWS * WARPS + 1];
    __shared__ int s_base[4];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nrows = 3 + nlight, n = nrows * WARPS;
    const unsigned rowbits = (e.m0 ? 1u : 0u) | (e.m1 ? 2u : 0u) | (e.m2 ? 4u : 0u) | (e.lights << 3);
    for (int r = 0; r < nrows; ++r) {
        unsigned b = __ballot_sync(0xffffffffu, (rowbits >> r) & 1u);
        if (lane == 0) s_cnt[r * WARPS + w] = __popc(b);
    }
    __syncthreads();
    if (w == 0) {
        // exclusive scan of the n <= 152 counts: 5 consecutive entries per lane + a warp scan
        int v[5], sum = 0;
#pragma unroll
        for (int k = 0; k < 5; ++k) { int idx = lane * 5 + k; v[k] = idx < n ? s_cnt[idx] : 0; sum += v[k]; }
        int incl = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += t; }
        int excl = incl - sum;
#pragma unroll
        for (int k = 0; k < 5; ++k) { int idx = lane * 5 + k; if (idx < n) s_cnt[idx] = excl; excl += v[k]; }
        if (lane == 31) s_cnt[n] = incl;
        __syncwarp();
        if (lane < 4) {
            int start = s_cnt[min(lane * WARPS, n)];
            int end = lane < 3 ? s_cnt[min((lane + 1) * WARPS, n)] : s_cnt[n];
            int cnt = end - start;
            int base = cnt ? atomicAdd(&W.counters[lane < 3 ? C_MAT0 + lane : C_SHADOW], cnt) : 0;
            s_base[lane] = base - start;
        }
    }
    __syncthreads();
    for (int r = 0; r < nrows; ++r) {
        unsigned b = __ballot_sync(0xffffffffu, (rowbits >> r) & 1u);
        if ((rowbits >> r) & 1u) {
            int at = s_base[min(r, 3)] + s_cnt[r * WARPS + w] + __popc(b & ((1u << lane) - 1u));
            if (r < 3) W.q_mat[r][at] = p; else W.q_shadow[at] = p * nlight + (r - 3);
        }
    }
}

// ---- hit-point order --------------------------------------------------------------------------------------------------
// From the first bounce on the paths of a batch hit the scene all over the place, and queue neighbours (neighbouring
// pixels) stop being spatial neighbours: the shadow rays and the next bounce rays of a warp then start in 32 unrelated
// corners of the tree.  Before k_hitinfo bins a bounce, its paths are therefore put in the Morton order of their hit
// points (one radix sort of (key, path slot) pairs per bounce, the pairs written by the closest-hit kernel itself; the
// camera rays' hits too — a 3D-compact block of vertices beats a scanline run of pixels): a block of k_hitinfo then holds 256 neighbouring
// vertices, its light-major shadow entries are 32 near-identical rays per warp, and k_shade appends the next rays in
// the same order, so the next closest-hit launch reads origin-sorted rays.  The key is only an ORDER (computed with an
// FMA, quantised to 1024^3 cells of the scene's coordinate range): no result depends on it — every path's arithmetic
// is a function of its own state and k_resolve adds samples in sample order.
// Key: 32-bit Morton code, 11 + 11 + 10 bits (x, y: 2048 cells, z: 1024 cells over the scene's coordinate range), all
// of it sorted (four 8-bit radix passes).  Coarser orders are cheaper to sort and cost more than they save: on one
// 32M-path batch of the 1M-triangle scene the shadow launches take 42.1 / 37.7 / 33.3 / 29.5 ms with 16 / 20 / 24 / 30
// key bits.
__device__ __forceinline__ uint32_t morton_spread11(uint32_t v) {   // bit i -> bit 3i, i < 11
    v &= 0x7FFu;
    v = (v | (v << 16)) & 0x070000FFu;
    v = (v | (v << 8)) & 0x0700F00Fu;
    v = (v | (v << 4)) & 0x430C30C3u;
    v = (v | (v << 2)) & 0x49249249u;
    return v;
}
// Morton key of the hit point o + t d (misses: the largest key).
__device__ __forceinline__ uint32_t hit_point_key(V3 o, V3 d, const HitRec& h, float bound) {
    if (h.tri < 0) return 0xFFFFFFFFu;
    const float sc = 1024.0f / bound;
    const float x = fmaf(fmaf(d.x, h.t, o.x), sc, 1024.0f), y = fmaf(fmaf(d.y, h.t, o.y), sc, 1024.0f), z = fmaf(fmaf(d.z, h.t, o.z), 0.5f * sc, 512.0f);
    const uint32_t xi = (uint32_t)fminf(fmaxf(x, 0.0f), 2047.0f), yi = (uint32_t)fminf(fmaxf(y, 0.0f), 2047.0f), zi = (uint32_t)fminf(fmaxf(z, 0.0f), 1023.0f);
    return morton_spread11(xi) | (morton_spread11(yi) << 1) | (morton_spread11(zi) << 2);
}

// block sizes of the two run-to-completion traversal kernels (64 / 128 / 256 measured: profiles/r01_experiments.md)
#ifndef B2PT_EXT_BLOCK
#define B2PT_EXT_BLOCK 128
#endif
#ifndef B2PT_EXT_BLOCK_FUSED
#define B2PT_EXT_BLOCK_FUSED 256   // fused epilogue: bigger blocks bin longer runs (Cornell +1.6 %); unfused: 128 (mesh -2 % at 256)
#endif
#ifndef B2PT_SHD_BLOCK
#define B2PT_SHD_BLOCK 128
#endif
#ifndef B2PT_EXT_MINB
#define B2PT_EXT_MINB 10
#endif
#ifndef B2PT_SHD_MINB
#define B2PT_SHD_MINB 10   // 48 registers (64 without a bound): sorted 1M-triangle batch 35.0 -> 33.3 ms of shadow rays
#endif
// Closest hit for the active paths (traverse_rtc.cuh): one thread per queue entry.
// With FUSED the kernel also runs the per-vertex epilogue of its certified rays (hit point, shading normal,
// material, light culling, binning — hit_epilogue + bin_path): the hit record never goes through HBM and the
// bandwidth-bound epilogue overlaps other warps' traversal.  Uncertified rays get theirs in k_extend_fallback.
// FUSED = false: the hit record is written and k_hitinfo does the rest.  Measured: fusing wins on scenes whose
// traversal is short and uniform (Cornell 769 -> 790 Msamples/s) and loses where ray costs vary (1M mesh 217 -> 206:
// the block-wide binning makes finished warps wait for the block's slowest ray), so render_frame fuses only small
// trees.
template <bool COUNT, bool FUSED>
__global__ void __launch_bounds__(FUSED ? B2PT_EXT_BLOCK_FUSED : B2PT_EXT_BLOCK, (B2PT_EXT_MINB * 128) / (FUSED ? B2PT_EXT_BLOCK_FUSED : B2PT_EXT_BLOCK)) k_extend_rtc(DeviceScene S, Wave W, const int* __restrict__ list, const int* __restrict__ count_ptr,
                                                    int P, TraceCounters* __restrict__ tc, uint32_t* __restrict__ order_keys, int* __restrict__ order_vals) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    int total = list ? *count_ptr : P;
    if (blockIdx.x * blockDim.x >= total) return;   // whole block beyond the queue (uniform)
    unsigned n_nodes = 0, n_tris = 0;
    Epilogue e{false, false, false, 0u};
    int p = 0;
    if (k < total) {
        p = list ? list[k] : k;
        // path state is read once and written once here: streaming accesses keep L1 for the nodes, the triangles and the stack
        float4 o4 = __ldcs(&W.ro[p]), d4 = __ldcs(&W.rd[p]);
        RayQ r = make_rayq_normalised(f4v(o4), f4v(d4), B2PT_INF);
        HitRec h;
        bool ok = closest_rtc<COUNT>(S, r, h, n_nodes, n_tris);
        if (!FUSED) {
            __stcs(&W.hit[p], make_float4(h.t, __int_as_float(h.tri), h.u, h.v));
            // (key, slot) pair of the hit-point order; an uncertified ray's key comes from its provisional hit — only an order
            if (order_keys) { __stcs(&order_keys[k], hit_point_key(r.o, r.d, h, S.coord_bound)); __stcs(&order_vals[k], p); }
        } else if (ok) e = hit_epilogue(S, W, p, r.o, r.d, h);
        if (!ok) W.q_fallback[atomicAdd(&W.counters[C_FALLBACK], 1)] = p;
    }
    if (FUSED) bin_path<B2PT_EXT_BLOCK_FUSED / 32>(W, e, p, S.nlight);
    if (COUNT) {
        for (int off = 16; off > 0; off >>= 1) {
            n_nodes += __shfl_down_sync(0xffffffffu, n_nodes, off);
            n_tris += __shfl_down_sync(0xffffffffu, n_tris, off);
        }
        if ((threadIdx.x & 31) == 0) { atomicAdd(&tc->node_fetches, (unsigned long long)n_nodes); atomicAdd(&tc->tri_fetches, (unsigned long long)n_tris); }
    }
}

template <bool COUNT>
__global__ void __launch_bounds__(B2PT_SHD_BLOCK, B2PT_SHD_MINB) k_shadow_rtc(DeviceScene S, Wave W, TraceCounters* __restrict__ tc) {
    const int nl = S.nlight;
    const int total = W.counters[C_SHADOW];
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned n_nodes = 0, n_tris = 0;
    if (j < total) {
        int e = W.q_shadow[j];
        int p = e / nl, l = e - p * nl;
        float4 g0 = W.g0[p], g1 = W.g1[p];
        V3 P = f4v(g0), n = f4v(g1);
        const DLight& lt = S.lights[l];
        V3 lightDir = vsub(mk3(lt.px, lt.py, lt.pz), P);
        float dist = vlength(lightDir);
        lightDir = vnormalize(lightDir);
        RayQ r = make_rayq(vadd(P, vmuls(n, 0.001f)), lightDir, B2PT_SUB(dist, 0.001f));   // renderer.hpp:271-275
        W.vis[e] = any_rtc<COUNT>(S, r, n_nodes, n_tris) ? 1 : 0;
    }
    if (COUNT) {
        for (int off = 16; off > 0; off >>= 1) {
            n_nodes += __shfl_down_sync(0xffffffffu, n_nodes, off);
            n_tris += __shfl_down_sync(0xffffffffu, n_tris, off);
        }
        if ((threadIdx.x & 31) == 0) { atomicAdd(&tc->node_fetches, (unsigned long long)n_nodes); atomicAdd(&tc->tri_fetches, (unsigned long long)n_tris); }
    }
}

// ---- incoherent bounces: several rays per lane, phase-split steps (traverse_pool.cuh) -------------------------------
// Same queues, same results as k_extend_rtc<.., false> / k_shadow_rtc; used from the first bounce on in scenes large
// enough for the rays of a warp to diverge.
struct IoExtend {
    Wave W; const int* list;
    __device__ __forceinline__ bool load(const DeviceScene&, long long k, V3& ro, V3& rd, float& T0, int& tag) const {
        const int p = list ? list[k] : (int)k;
        float4 o4 = W.ro[p], d4 = W.rd[p];
        ro = f4v(o4); rd = f4v(d4); T0 = B2PT_INF; tag = p;   // rd is the Ray ctor's normalised direction already
        return true;
    }
    __device__ __forceinline__ void store_closest(int p, const HitRec& h, bool certified) const {
        W.hit[p] = make_float4(h.t, __int_as_float(h.tri), h.u, h.v);
        if (!certified) W.q_fallback[atomicAdd(&W.counters[C_FALLBACK], 1)] = p;
    }
    __device__ __forceinline__ void store_any(int, bool) const {}
};
struct IoShadow {
    Wave W;
    __device__ __forceinline__ bool load(const DeviceScene& S, long long k, V3& ro, V3& rd, float& T0, int& tag) const {
        const int e = W.q_shadow[k], nl = S.nlight;
        const int p = e / nl, l = e - p * nl;
        float4 g0 = W.g0[p], g1 = W.g1[p];
        V3 P = f4v(g0), n = f4v(g1);
        const DLight& lt = S.lights[l];
        V3 lightDir = vsub(mk3(lt.px, lt.py, lt.pz), P);
        float dist = vlength(lightDir);
        ro = vadd(P, vmuls(n, 0.001f));                 // renderer.hpp:271-275
        rd = vnormalize(vnormalize(lightDir));          // :271, then the Ray ctor (ray.hpp:12)
        T0 = B2PT_SUB(dist, 0.001f);
        tag = e;
        return true;
    }
    __device__ __forceinline__ void store_closest(int, const HitRec&, bool) const {}
    __device__ __forceinline__ void store_any(int e, bool occluded) const { W.vis[e] = occluded ? 1 : 0; }
};

template <bool COUNT>
__global__ void __launch_bounds__(B2PT_PBLOCK, B2PT_PMINB) k_extend_pool(DeviceScene S, Wave W, const int* __restrict__ list, const int* __restrict__ count_ptr, int P,
                                                             TraceCounters* __restrict__ tc) {
    __shared__ PoolSmem<false> sm;
    IoExtend io{W, list};
    pool_traverse<false, COUNT>(S, sm, io, &W.totals[4], list ? (long long)*count_ptr : (long long)P, tc);
}
template <bool COUNT>
__global__ void __launch_bounds__(B2PT_PBLOCK, B2PT_PMINB) k_shadow_pool(DeviceScene S, Wave W, TraceCounters* __restrict__ tc) {
    __shared__ PoolSmem<true> sm;
    IoShadow io{W};
    pool_traverse<true, COUNT>(S, sm, io, &W.totals[5], (long long)W.counters[C_SHADOW], tc);
}

// Shadow kernel of the ONE statistics batch per scene (occluder-aware child order, build.cu learn_child_order): same
// visibility bytes as k_shadow_rtc; every 64th warp runs the instrumented query, which also counts box passes and
// terminal hits per (wide node, slot) — the other 63 run the production query, so the batch costs what any batch costs
// (a frame of ONE batch, e.g. BASELINE configs[0] re-uploaded every frame, is all learning batch).
__global__ void __launch_bounds__(B2PT_SHD_BLOCK, B2PT_SHD_MINB) k_shadow_learn(DeviceScene S, Wave W, unsigned* __restrict__ visits, unsigned* __restrict__ hits) {
    const int nl = S.nlight;
    const int total = W.counters[C_SHADOW];
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < total) {
        int e = W.q_shadow[j];
        int p = e / nl, l = e - p * nl;
        float4 g0 = W.g0[p], g1 = W.g1[p];
        V3 P = f4v(g0), n = f4v(g1);
        const DLight& lt = S.lights[l];
        V3 lightDir = vsub(mk3(lt.px, lt.py, lt.pz), P);
        float dist = vlength(lightDir);
        lightDir = vnormalize(lightDir);
        RayQ r = make_rayq(vadd(P, vmuls(n, 0.001f)), lightDir, B2PT_SUB(dist, 0.001f));   // renderer.hpp:271-275
        unsigned n0 = 0, n1 = 0;
        const bool sampled = ((j >> 5) & 63) == 0;   // warp-uniform
        W.vis[e] = (sampled ? any_rtc_learn(S, r, true, visits, hits) : any_rtc<false>(S, r, n0, n1)) ? 1 : 0;
    }
}

// The rays the cooperative kernel could not certify: the flattened reference recursion, one thread per ray.
template <bool FUSED>
__global__ void __launch_bounds__(128) k_extend_fallback(DeviceScene S, Wave W) {
    int total = W.counters[C_FALLBACK];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
        int p = W.q_fallback[k];
        float4 o4 = W.ro[p], d4 = W.rd[p];
        RayQ r = make_rayq_normalised(f4v(o4), f4v(d4), B2PT_INF);
        HitRec h;
        closest_exact_dfs(S, r, h);
        if (!FUSED) {
            W.hit[p] = make_float4(h.t, __int_as_float(h.tri), h.u, h.v);
        } else {
            // the epilogue k_extend_rtc runs for certified rays; a handful of rays per frame, plain atomics
            Epilogue e = hit_epilogue(S, W, p, r.o, r.d, h);
            if (e.m0) W.q_mat[0][atomicAdd(&W.counters[C_MAT0], 1)] = p;
            if (e.m1) W.q_mat[1][atomicAdd(&W.counters[C_MAT1], 1)] = p;
            if (e.m2) W.q_mat[2][atomicAdd(&W.counters[C_MAT2], 1)] = p;
            for (int l = 0; l < S.nlight; ++l)
                if ((e.lights >> l) & 1u) W.q_shadow[atomicAdd(&W.counters[C_SHADOW], 1)] = p * S.nlight + l;
        }
    }
}

// Hit record -> hit point / shading normal / material, and the one-pass counting sort by material type.
__global__ void __launch_bounds__(B2PT_BIN_BLOCK) k_hitinfo(DeviceScene S, Wave W, const int* __restrict__ list, const int* __restrict__ count_ptr, int P) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    int total = list ? *count_ptr : P;
    Epilogue e{false, false, false, 0u};
    int p = 0;
    if (k < total) {
        p = list ? list[k] : k;
        float4 h4 = W.hit[p], o4 = W.ro[p], d4 = W.rd[p];
        HitRec h;
        h.t = h4.x; h.tri = __float_as_int(h4.y); h.u = h4.z; h.v = h4.w;
        e = hit_epilogue(S, W, p, f4v(o4), f4v(d4), h);
    }
    bin_path<B2PT_BIN_WARPS>(W, e, p, S.nlight);
}

// material.hpp:28-42
__device__ __forceinline__ float ggx_distribution(float NdotH, float roughness) {
    if (roughness < 0.0f) roughness = 0.0f;
    if (roughness > 1.0f) roughness = 1.0f;
    float alpha = B2PT_MUL(roughness, roughness);
    float alpha2 = B2PT_MUL(alpha, alpha);
    float NdotH2 = B2PT_MUL(NdotH, NdotH);
    float denom = B2PT_ADD(B2PT_MUL(NdotH2, B2PT_SUB(alpha2, 1.0f)), 1.0f);
    if (denom <= 0.0f) return 0.0f;
    return B2PT_DIV(alpha2, B2PT_MUL(B2PT_MUL(3.14159265358979323846264338327950288f, denom), denom));
}
// material.hpp:21-26
__device__ __forceinline__ float schlick_fresnel(float cosTheta, float F0) {
    float x = B2PT_SUB(1.0f, cosTheta);
    float x2 = B2PT_MUL(x, x);
    float x5 = B2PT_MUL(B2PT_MUL(x2, x2), x);
    return B2PT_ADD(F0, B2PT_MUL(B2PT_SUB(1.0f, F0), x5));
}

// calculateDirectLighting (renderer.hpp:252-301) given the visibility of every light.
__device__ __forceinline__ V3 direct_lighting(const DeviceScene& S, const DMaterial& m, V3 P, V3 n, V3 viewDir, const uint8_t* __restrict__ vis) {
    V3 total = mk3(0.0f, 0.0f, 0.0f);
    for (int l = 0; l < S.nlight; ++l) {
        const DLight& lt = S.lights[l];
        V3 lightDir = vsub(mk3(lt.px, lt.py, lt.pz), P);
        float dist = vlength(lightDir);
        if (dist < 0.0001f) continue;                                     // :263-269
        if (vis[l]) continue;                                             // :278
        lightDir = vnormalize(lightDir);
        float cosTheta = gmax(vdot(n, lightDir), 0.0f);
        float att = B2PT_DIV(lt.intensity, B2PT_MUL(dist, dist));
        V3 brdf;
        if (m.type == B2PT_DIFFUSE) {
            brdf = vdivs(mk3(m.r, m.g, m.b), 3.14159265358979323846264338327950288f);
        } else {
            V3 halfVec = vnormalize(vadd(lightDir, viewDir));
            float NdotH = gmax(vdot(n, halfVec), 0.0f);
            brdf = vmuls(mk3(m.r, m.g, m.b), ggx_distribution(NdotH, m.roughness));
        }
        V3 c = vmuls(vmuls(vmul(mk3(lt.cr, lt.cg, lt.cb), brdf), cosTheta), att);
        if (valid3(c)) total = vadd(total, c);                            // :295-297
    }
    return total;
}

// tracePath's material switch (renderer.hpp:166-247) for one vertex of material type TYPE; returns -1 when the path
// ends, else the direction octant (sign bits) of its next ray, which is then in ro/rd.
__device__ __forceinline__ int octant_of(V3 d) { return (d.x < 0.0f ? 1 : 0) | (d.y < 0.0f ? 2 : 0) | (d.z < 0.0f ? 4 : 0); }
template <int TYPE>
__device__ __forceinline__ int shade_vertex(const DeviceScene& S, const Wave& W, const FrameConst& F, long long pix_begin, int npc, int s_begin,
                                             int depth, int p) {
    int cont = -1;
    {
        float4 g0 = W.g0[p], g1 = W.g1[p], d4 = W.rd[p];
        V3 P = f4v(g0), n = f4v(g1), d = f4v(d4);
        const DMaterial m = S.mats[__float_as_int(g0.w)];
        int s_local = p / npc, jl = p - s_local * npc;
        uint32_t pix = (uint32_t)pixel_of(F, pix_begin + jl);
        uint32_t smp = (uint32_t)(s_begin + s_local);
        const bool last = depth + 1 >= F.max_bounces;
        if (TYPE == B2PT_DIELECTRIC) {
            // renderer.hpp:214-246 (direct light is dropped for dielectrics)
            if (!last) {
                float cosTheta = vdot(vneg(d), n);
                float etai = 1.0f, etat = m.ior;
                V3 normal = n;
                if (cosTheta < 0.0f) { cosTheta = -cosTheta; float s = etai; etai = etat; etat = s; normal = vneg(normal); }
                float sinTheta = B2PT_SQRT(B2PT_SUB(1.0f, B2PT_MUL(cosTheta, cosTheta)));
                float ratio = B2PT_DIV(etai, etat);
                float coin = u01(philox4x32_10(pix, smp, (uint32_t)depth, DRAW_COIN, F.k0, F.k1).x);
                V3 dir;
                if (B2PT_MUL(ratio, sinTheta) > 1.0f ||
                    coin < schlick_fresnel(cosTheta, B2PT_DIV(B2PT_SUB(etai, etat), B2PT_ADD(etai, etat)))) {
                    dir = vreflect(d, normal);
                } else {
                    dir = vrefract(d, normal, ratio);
                }
                float len = vlength(dir);
                if (!(isnan(len) || isinf(len))) {
                    V3 no = vadd(P, vmuls(normal, 0.001f));
                    V3 nd = vnormalize(dir);   // Ray ctor
                    W.ro[p] = make_float4(no.x, no.y, no.z, 0.0f);
                    W.rd[p] = make_float4(nd.x, nd.y, nd.z, 0.0f);
                    cont = octant_of(nd);
                }
            }
        } else {
            V3 direct = direct_lighting(S, m, P, n, vneg(d), W.vis + (long long)p * S.nlight);
            if (valid3(direct)) {                                          // :161-163
                V3 dir;
                if (TYPE == B2PT_DIFFUSE) {
                    dir = random_in_unit_sphere(pix, smp, (uint32_t)depth, F.k0, F.k1);
                    if (vdot(dir, n) < 0.0f) dir = vneg(dir);              // :303-306
                } else {
                    dir = vreflect(d, n);                                  // :191
                    if (m.roughness > 0.0f)
                        dir = vnormalize(vadd(dir, vsmul(m.roughness, random_in_unit_sphere(pix, smp, (uint32_t)depth, F.k0, F.k1))));
                }
                float cosTheta = vdot(dir, n);
                if (!(isnan(cosTheta) || isinf(cosTheta))) {
                    float4 T4 = W.thr[p], L4 = W.rad[p];
                    V3 T = f4v(T4), L = f4v(L4);
                    L = vadd(L, vmul(T, direct));
                    W.rad[p] = make_float4(L.x, L.y, L.z, 0.0f);
                    if (!last) {
                        V3 alb = mk3(m.r, m.g, m.b);
                        V3 wgt;
                        if (TYPE == B2PT_DIFFUSE) {
                            V3 brdf = vdivs(alb, 3.14159265358979323846264338327950288f);
                            wgt = vmuls(vmuls(vmuls(brdf, cosTheta), 2.0f), 3.14159265358979323846264338327950288f);
                        } else {
                            wgt = vmuls(alb, cosTheta);
                        }
                        T = vmul(T, wgt);
                        V3 no = vadd(P, vmuls(n, 0.001f));
                        V3 nd = vnormalize(dir);   // Ray ctor
                        W.thr[p] = make_float4(T.x, T.y, T.z, 0.0f);
                        W.ro[p] = make_float4(no.x, no.y, no.z, 0.0f);
                        W.rd[p] = make_float4(nd.x, nd.y, nd.z, 0.0f);
                        cont = octant_of(nd);
                    }
                }
            }
        }
    }
    return cont;
}

// One launch shades all three material queues: thread k takes entry k of the concatenation diffuse | specular |
// dielectric, so a block is branch-free on the material except where two queues meet.
__global__ void __launch_bounds__(B2PT_BIN_BLOCK) k_shade(DeviceScene S, Wave W, FrameConst F, long long pix_begin, int npc, int s_begin,
                                                           int depth, int next_slot) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int c0 = W.counters[C_MAT0], c1 = W.counters[C_MAT1], c2 = W.counters[C_MAT2];
    if (blockIdx.x * blockDim.x >= c0 + c1 + c2) return;   // whole block beyond the queues (uniform)
    int cont = -1;
    int p = 0;
    if (k < c0) {
        p = W.q_mat[0][k];
        cont = shade_vertex<B2PT_DIFFUSE>(S, W, F, pix_begin, npc, s_begin, depth, p);
    } else if (k < c0 + c1) {
        p = W.q_mat[1][k - c0];
        cont = shade_vertex<B2PT_SPECULAR>(S, W, F, pix_begin, npc, s_begin, depth, p);
    } else if (k < c0 + c1 + c2) {
        p = W.q_mat[2][k - c0 - c1];
        cont = shade_vertex<B2PT_DIELECTRIC>(S, W, F, pix_begin, npc, s_begin, depth, p);
    }
    block_append_binned(&W.counters[next_slot], W.q_active[next_slot], cont, p);
}

// Bookkeeping between bounces (single thread): totals, reset the per-bounce counters.
__global__ void k_begin_bounce(Wave W, int cur_slot, int first, int P, int nlight) {
    // called BEFORE extend of a depth: the active count of this depth is known here
    int active = first ? P : W.counters[cur_slot];
    if (first) W.counters[cur_slot] = P;   // depth 0 has no queue; kernels handed a sorted list read the count from here
    W.totals[0] += (unsigned long long)active;
    W.counters[cur_slot ^ 1] = 0;
    W.counters[C_MAT0] = 0; W.counters[C_MAT1] = 0; W.counters[C_MAT2] = 0;
    W.counters[C_SHADOW] = 0; W.counters[C_FALLBACK] = 0; W.counters[C_NEXT] = 0;
    W.totals[4] = 0; W.totals[5] = 0;
}
__global__ void k_after_extend(Wave W, int nlight) {
    W.totals[1] += (unsigned long long)W.counters[C_SHADOW];
    W.totals[2] += (unsigned long long)W.counters[C_FALLBACK];
}

// renderer.hpp:62-72: samples are added in sample order; invalid (NaN/Inf) samples are skipped.
__global__ void __launch_bounds__(256) k_resolve(Wave W, float4* __restrict__ accum, long long pix_begin, int npc, int ns) {
    int jl = blockIdx.x * blockDim.x + threadIdx.x;
    if (jl >= npc) return;
    float4 a = accum[pix_begin + jl];
    for (int s = 0; s < ns; ++s) {
        float4 L = W.rad[(long long)s * npc + jl];
        if (valid3(f4v(L))) {
            a.x = B2PT_ADD(a.x, L.x); a.y = B2PT_ADD(a.y, L.y); a.z = B2PT_ADD(a.z, L.z);
            a.w = 1.0f;
        }
    }
    accum[pix_begin + jl] = a;
}

// renderer.hpp:75-81
__global__ void __launch_bounds__(256) k_finalize(const float4* __restrict__ accum, FrameConst F, long long nown, int all_samples, int divisor, float* __restrict__ rgb) {
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nown) return;
    long long i = pixel_of(F, j);
    float4 a = accum[j];
    float r, g, b;
    if (a.w != 0.0f) {
        float spp = (float)divisor;
        r = B2PT_DIV(a.x, spp); g = B2PT_DIV(a.y, spp); b = B2PT_DIV(a.z, spp);
    } else if (all_samples) {
        r = 1.0f; g = 0.0f; b = 1.0f;   // debug colour for pixels without a valid sample (:78)
    } else {
        r = g = b = 0.0f;
    }
    rgb[3 * i + 0] = r; rgb[3 * i + 1] = g; rgb[3 * i + 2] = b;
}

// Renderer::saveImage's tonemap (src/renderer.cpp:8-17): clamp -> pow(c, 1/gamma) -> (unsigned char)(c * 255).  The byte a
// value maps to is decided by the HOST's libm powf, the function the reference calls: tonemap_frame finds, for every byte
// value k, the smallest float whose reference byte is >= k (a bisection over float bit patterns with the host's own powf)
// and the device only compares against those 255 thresholds — byte-exact without a device pow.  `flip` writes the rows
// top-down (the reference's image is upside down: renderer.hpp:81 stores row 0 = bottom of the view).
__global__ void __launch_bounds__(256) k_tonemap(const float* __restrict__ rgb, int width, int height, int flip, const float* __restrict__ thr,
                                                  uint8_t* __restrict__ out) {
    __shared__ float s_thr[256];
    s_thr[threadIdx.x] = thr[threadIdx.x];
    __syncthreads();
    const long long n = 3ll * width * height;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float c = rgb[i];
    c = gmin(gmax(c, 0.0f), 1.0f);          // glm::clamp = min(max(x, lo), hi); a NaN stays a NaN and maps to byte 0
    int lo = 0, hi = 255;                   // largest k with thr[k] <= c  (thr[0] = 0, thr non-decreasing)
#pragma unroll
    for (int step = 0; step < 8; ++step) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_thr[mid] <= c) lo = mid; else hi = mid - 1;
    }
    long long o = i;
    if (flip) {
        const long long row = i / (3ll * width), col = i - row * 3ll * width;
        o = (height - 1 - row) * 3ll * width + col;
    }
    out[o] = (uint8_t)lo;
}

}  // namespace

// keep_accum: the per-pixel sums of the previous call stay (progressive rendering: the samples of successive passes are
// added in sample order, exactly as one call over the whole range would add them); divisor: what the sums are divided
// by (0 = settings.samples_per_pixel).
int render_frame(b2pt_ctx* ctx, const b2pt_camera* cam, const b2pt_settings* st, uint64_t seed,
                 const b2pt_partition* part, float* d_rgb, bool keep_accum, int divisor) {
    cudaStream_t stream = ctx->stream;
    const DeviceScene& S = ctx->scene;
    const int W_ = st->width, H_ = st->height;
    const long long npix = (long long)W_ * H_;

    FrameConst F{};
    F.width = W_; F.height = H_; F.spp_total = st->samples_per_pixel; F.max_bounces = st->max_bounces;
    F.k0 = (uint32_t)seed; F.k1 = (uint32_t)(seed >> 32);
    int tile_world = part ? part->tile_world : 0, tile_rank = part ? part->tile_rank : 0;
    int tile_size = (part && part->tile_size > 0) ? part->tile_size : 32;
    if (tile_world > 1 && (tile_rank < 0 || tile_rank >= tile_world)) { ctx->err = "b2pt_render: tile_rank out of range"; return B2PT_ERR_INVALID; }
    F.tile_world = tile_world > 1 ? tile_world : 1; F.tile_rank = tile_world > 1 ? tile_rank : 0; F.tile_area = tile_size * tile_size;
    int s_begin = part ? part->sample_begin : 0;
    int s_count = (part && part->sample_count > 0) ? part->sample_count : st->samples_per_pixel - s_begin;
    if (s_begin < 0 || s_count < 0 || s_begin + s_count > st->samples_per_pixel) { ctx->err = "b2pt_render: sample range out of bounds"; return B2PT_ERR_INVALID; }
    const int all_samples = ((s_begin == 0 || keep_accum) && s_begin + s_count == st->samples_per_pixel) ? 1 : 0;
    if (divisor <= 0) divisor = st->samples_per_pixel;

    // owned pixels: runs of tile_area consecutive pixels (row-major), run k -> rank k % world
    long long nown = npix;
    if (F.tile_world > 1) {
        long long A = F.tile_area, nruns = (npix + A - 1) / A;
        long long mine = (nruns - F.tile_rank + F.tile_world - 1) / F.tile_world;   // runs r, r+w, ...
        nown = mine * A;
        long long last_run = (mine - 1) * F.tile_world + F.tile_rank;
        if (mine > 0 && last_run == nruns - 1) nown -= nruns * A - npix;   // the frame's final, partial run
        if (mine <= 0) nown = 0;
    }

    // camera constants, host fp32 in the order of camera.hpp:19-26
    CamConst C{};
    {
        auto v = [](const float* p) { return mk3(p[0], p[1], p[2]); };
        V3 pos = v(cam->position), fwd = v(cam->forward), right = v(cam->right), up = v(cam->up);
        float theta = cam->fov * 0.01745329251994329576923690768489f;
        float h = std::tan(theta / 2.0f);
        float vh = 2.0f * h;
        float vw = vh * (16.0f / 9.0f);
        V3 hor = mk3(vw * right.x, vw * right.y, vw * right.z);
        V3 ver = mk3(vh * up.x, vh * up.y, vh * up.z);
        V3 llc = mk3(pos.x - hor.x / 2.0f - ver.x / 2.0f + fwd.x, pos.y - hor.y / 2.0f - ver.y / 2.0f + fwd.y,
                     pos.z - hor.z / 2.0f - ver.z / 2.0f + fwd.z);
        C.pos = pos; C.llc = llc; C.horizontal = hor; C.vertical = ver;
    }

    B2PT_CUDA(ctx, cudaMemsetAsync(d_rgb, 0, sizeof(float) * 3 * npix, stream));
    ctx->stats.samples = nown * s_count;
    if (nown == 0 || s_count == 0) return B2PT_OK;

    // chunking
    long long maxp = std::max<long long>(ctx->max_paths, 1024);
    maxp = std::min<long long>(maxp, 0x7fffffffll / std::max(S.nlight, 1));   // shadow-queue entries p*nlight + l are ints
    int npc_max = (int)std::min<long long>(nown, maxp);
    int ns_max = (int)std::max<long long>(1, std::min<long long>(s_count, maxp / npc_max));
    long long Pmax = (long long)npc_max * ns_max;

    // scratch: slot 8 = path state + queues, slot 9 = accumulators, slot 10 = counters/totals
    size_t f4 = sizeof(float4) * (size_t)Pmax, qi = sizeof(int) * (size_t)Pmax;
    void* base = nullptr;
    size_t vb = ((size_t)Pmax * (size_t)std::max(S.nlight, 1) + 255) & ~(size_t)255;
    const size_t qs = qi * (size_t)std::max(S.nlight, 1);   // shadow queue: one entry per (vertex, light)
    int rc = scratch_reserve(ctx, 8, B2PT_PATH_FIELDS * f4 + 6 * qi + qs + vb + 1024, &base);
    if (rc) return rc;
    Wave Wv{};
    {
        char* b = (char*)base;
        float4* state = (float4*)b; b += B2PT_PATH_FIELDS * f4;
        const bool lines = S.nwide > 64;   // == !fused below: the scenes whose bounces are sorted
        PathField* fields[7] = {&Wv.ro, &Wv.rd, &Wv.thr, &Wv.rad, &Wv.g0, &Wv.g1, &Wv.hit};
        for (int f = 0; f < 7; ++f) {
            fields[f]->base = lines ? state + f : state + (size_t)f * (size_t)Pmax;
            fields[f]->stride = lines ? B2PT_PATH_FIELDS : 1;
        }
        Wv.q_active[0] = (int*)b; b += qi; Wv.q_active[1] = (int*)b; b += qi;
        Wv.q_mat[0] = (int*)b; b += qi; Wv.q_mat[1] = (int*)b; b += qi; Wv.q_mat[2] = (int*)b; b += qi;
        Wv.q_shadow = (int*)b; b += qs; Wv.q_fallback = (int*)b; b += qi;
        Wv.vis = (uint8_t*)b; b += vb;
    }
    void* accum = nullptr;
    if ((rc = scratch_reserve(ctx, 9, sizeof(float4) * (size_t)nown, &accum))) return rc;
    void* cnt = nullptr;
    if ((rc = scratch_reserve(ctx, 10, 256, &cnt))) return rc;
    Wv.counters = (int*)cnt;
    Wv.totals = (unsigned long long*)((char*)cnt + 128);
    B2PT_CUDA(ctx, cudaMemsetAsync(cnt, 0, 256, stream));
    if (!keep_accum) B2PT_CUDA(ctx, cudaMemsetAsync(accum, 0, sizeof(float4) * (size_t)nown, stream));

    const bool count = (ctx->flags & B2PT_FLAG_COUNT_FETCHES) != 0;
    // closest-hit kernel with the per-vertex epilogue fused in: small trees only (see k_extend_rtc)
    const bool fused = S.nwide <= 64;
    // hit-point order of the bounces (hit_point_key): scenes whose rays diverge, i.e. the ones that do not fuse
    const bool sort_hits = !fused && !(ctx->flags & B2PT_FLAG_NO_SORT);
    uint32_t *sort_keys[2] = {nullptr, nullptr};
    int* sort_vals[2] = {nullptr, nullptr};
    void* sort_temp = nullptr;
    size_t sort_temp_bytes = 0;
    if (sort_hits) {
        void* sb_ = nullptr;
        if ((rc = scratch_reserve(ctx, 13, 4 * qi, &sb_))) return rc;
        sort_keys[0] = (uint32_t*)sb_; sort_keys[1] = sort_keys[0] + Pmax; sort_vals[0] = (int*)(sort_keys[1] + Pmax); sort_vals[1] = sort_vals[0] + Pmax;
        B2PT_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, sort_temp_bytes, sort_keys[0], sort_keys[1], sort_vals[0], sort_vals[1], (int)Pmax,
                                                       0, 32, stream));   // sized for any bit range
        if ((rc = scratch_reserve(ctx, 14, sort_temp_bytes + 16, &sort_temp))) return rc;
    }
    int64_t launches = 0, n_extend = 0, n_shadow = 0;
    float extend_ms = 0.0f, order_ms = 0.0f, shadow_ms = 0.0f;
    // Traversal time is measured with events around the closest-hit launch, the ordering / epilogue stage and the shadow
    // launch of every bounce; they come from a pool owned by
    // the context (created once, reused by every frame) and are read after the frame's single synchronisation.
    size_t ev_used = 0;
    auto ev = [&]() {
        if (ev_used == ctx->ev_pool.size()) { cudaEvent_t e = nullptr; if (cudaEventCreate(&e) != cudaSuccess) return; ctx->ev_pool.push_back(e); }
        cudaEventRecord(ctx->ev_pool[ev_used++], stream);
    };

    // B2PT_DEBUG_SYNC=1: synchronise after every kernel and name the first one that faults (debugging aid).
    const bool dbg_sync = ctx->debug_sync;
    std::string dbg_fault;
    auto dbg = [&](const char* kernel, long long pb, int sb_, int depth_) {
        ++launches;
        if (!dbg_sync || !dbg_fault.empty()) return;
        cudaError_t e = cudaStreamSynchronize(stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess)
            dbg_fault = std::string(kernel) + " (pixel chunk " + std::to_string(pb) + ", sample chunk " + std::to_string(sb_) + ", depth " +
                        std::to_string(depth_) + "): " + cudaGetErrorString(e);
    };
    for (long long pix_begin = 0; pix_begin < nown; pix_begin += npc_max) {
        int npc = (int)std::min<long long>(npc_max, nown - pix_begin);
        for (int sb = 0; sb < s_count; sb += ns_max) {
            if (ctx->order_state == 1) {   // statistics of the previous batch are in: re-order the children once
                int lrc = learn_child_order(ctx);
                if (lrc) return lrc;
            }
            const bool learn_batch = ctx->order_state == 0 && !count;
            int ns = std::min(ns_max, s_count - sb);
            int P = npc * ns;
            int sabs = s_begin + sb;
            k_raygen<<<(P + 255) / 256, 256, 0, stream>>>(Wv, C, F, pix_begin, npc, sabs, P);
            dbg("k_raygen", pix_begin, sb, -1);
            // A = number of paths the kernels of this depth are launched for: the exact active count, read back after every
            // k_shade, when the bounces are sorted (grids, sort sizes and the end of the loop follow the surviving paths);
            // otherwise P, with the kernels leaving early beyond the device-side count.
            int A = P;
            for (int depth = 0; depth < st->max_bounces && A > 0; ++depth) {
                int cur = depth & 1;
                k_begin_bounce<<<1, 1, 0, stream>>>(Wv, cur, depth == 0, P, S.nlight);
                dbg("k_begin_bounce", pix_begin, sb, depth);
                const int* list = depth == 0 ? nullptr : Wv.q_active[cur];
                ev();
                // Kernel variants.  With the bounces in hit-point order (sort_hits, the default for large scenes) the rays of a
                // warp are near-identical and both ray kinds take the run-to-completion kernels (shadow rays of a 32M-path batch:
                // sorted + rtc 42 ms, sorted + pool 61 ms, unsorted + pool 73 ms, unsorted + rtc 80 ms).  Unsorted bounces
                // (B2PT_FLAG_NO_SORT) keep the phase-split pool kernel for shadow rays, and for closest hit with
                // B2PT_FLAG_POOL_EXTEND (it does not beat the run-to-completion kernel there: 850 vs 801 ms per C3 frame).
                const bool pooled = !fused && depth > 0 && !sort_hits && !(ctx->flags & B2PT_FLAG_LANE_KERNELS);
                const unsigned pgrid = (unsigned)std::min<long long>(((long long)A + B2PT_PBLOCK * B2PT_PR - 1) / (B2PT_PBLOCK * B2PT_PR), (long long)ctx->sm_count * 8);
                if (pooled && (ctx->flags & B2PT_FLAG_POOL_EXTEND)) {
                    if (count) k_extend_pool<true><<<pgrid, B2PT_PBLOCK, 0, stream>>>(S, Wv, list, &Wv.counters[cur], P, ctx->d_counters);
                    else k_extend_pool<false><<<pgrid, B2PT_PBLOCK, 0, stream>>>(S, Wv, list, &Wv.counters[cur], P, ctx->d_counters);
                } else if (fused) {
                    if (count) k_extend_rtc<true, true><<<(A + B2PT_EXT_BLOCK_FUSED - 1) / B2PT_EXT_BLOCK_FUSED, B2PT_EXT_BLOCK_FUSED, 0, stream>>>(S, Wv, list, &Wv.counters[cur], P, ctx->d_counters, nullptr, nullptr);
                    else k_extend_rtc<false, true><<<(A + B2PT_EXT_BLOCK_FUSED - 1) / B2PT_EXT_BLOCK_FUSED, B2PT_EXT_BLOCK_FUSED, 0, stream>>>(S, Wv, list, &Wv.counters[cur], P, ctx->d_counters, nullptr, nullptr);
                } else {
                    if (count) k_extend_rtc<true, false><<<(A + B2PT_EXT_BLOCK - 1) / B2PT_EXT_BLOCK, B2PT_EXT_BLOCK, 0, stream>>>(S, Wv, list, &Wv.counters[cur], P, ctx->d_counters, sort_keys[0], sort_vals[0]);
                    else k_extend_rtc<false, false><<<(A + B2PT_EXT_BLOCK - 1) / B2PT_EXT_BLOCK, B2PT_EXT_BLOCK, 0, stream>>>(S, Wv, list, &Wv.counters[cur], P, ctx->d_counters, sort_keys[0], sort_vals[0]);
                }
                dbg("k_extend", pix_begin, sb, depth);
                ev();
                if (fused) {   // k_extend_rtc has already run the epilogue of its certified rays
                    k_extend_fallback<true><<<ctx->sm_count * 4, 128, 0, stream>>>(S, Wv);
                    dbg("k_extend_fallback", pix_begin, sb, depth);
                } else if (sort_hits) {
                    // The exact recursion of the handful of uncertified rays (a fraction of a millisecond of pure latency) runs
                    // on the side stream while this one sorts the bounce (their keys come from the provisional hits: only an order).
                    B2PT_CUDA(ctx, cudaEventRecord(ctx->ev_fork, stream));
                    B2PT_CUDA(ctx, cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
                    k_extend_fallback<false><<<ctx->sm_count * 4, 128, 0, ctx->side>>>(S, Wv);
                    B2PT_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->side));
                    ++launches;
                    size_t tb = sort_temp_bytes;
                    B2PT_CUDA(ctx, cub::DeviceRadixSort::SortPairs(sort_temp, tb, sort_keys[0], sort_keys[1], sort_vals[0], sort_vals[1], A,
                                                                   0, 32, stream));
                    launches += 6;   // CUB: histogram, exclusive sum, four onesweep passes
                    B2PT_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->ev_join, 0));
                    k_hitinfo<<<(A + B2PT_BIN_BLOCK - 1) / B2PT_BIN_BLOCK, B2PT_BIN_BLOCK, 0, stream>>>(S, Wv, sort_vals[1], &Wv.counters[cur], P);
                    dbg("k_hitinfo", pix_begin, sb, depth);
                } else {
                    k_extend_fallback<false><<<ctx->sm_count * 4, 128, 0, stream>>>(S, Wv);
                    dbg("k_extend_fallback", pix_begin, sb, depth);
                    k_hitinfo<<<(A + B2PT_BIN_BLOCK - 1) / B2PT_BIN_BLOCK, B2PT_BIN_BLOCK, 0, stream>>>(S, Wv, list, &Wv.counters[cur], P);
                    dbg("k_hitinfo", pix_begin, sb, depth);
                }
                k_after_extend<<<1, 1, 0, stream>>>(Wv, S.nlight);
                dbg("k_after_extend", pix_begin, sb, depth);
                ev();
                ++n_extend;
                if (S.nlight > 0) {
                    ++n_shadow;
                    unsigned sg = (unsigned)(((long long)A * S.nlight + B2PT_SHD_BLOCK - 1) / B2PT_SHD_BLOCK);
                    const unsigned spgrid = (unsigned)std::min<long long>(((long long)A * S.nlight + B2PT_PBLOCK * B2PT_PR - 1) / (B2PT_PBLOCK * B2PT_PR), (long long)ctx->sm_count * 8);
                    if (learn_batch) k_shadow_learn<<<sg, B2PT_SHD_BLOCK, 0, stream>>>(S, Wv, ctx->d_order_stats, ctx->d_order_stats + 8 * (size_t)S.nwide);
                    else if (pooled && count) k_shadow_pool<true><<<spgrid, B2PT_PBLOCK, 0, stream>>>(S, Wv, ctx->d_counters);
                    else if (pooled) k_shadow_pool<false><<<spgrid, B2PT_PBLOCK, 0, stream>>>(S, Wv, ctx->d_counters);
                    else if (count) k_shadow_rtc<true><<<sg, B2PT_SHD_BLOCK, 0, stream>>>(S, Wv, ctx->d_counters);
                    else k_shadow_rtc<false><<<sg, B2PT_SHD_BLOCK, 0, stream>>>(S, Wv, ctx->d_counters);
                    dbg("k_shadow", pix_begin, sb, depth);
                }
                ev();
                int nxt = cur ^ 1;
                k_shade<<<(A + B2PT_BIN_BLOCK - 1) / B2PT_BIN_BLOCK, B2PT_BIN_BLOCK, 0, stream>>>(S, Wv, F, pix_begin, npc, sabs, depth, nxt);
                dbg("k_shade", pix_begin, sb, depth);
                if (sort_hits && depth + 1 < st->max_bounces) {
                    B2PT_CUDA(ctx, cudaMemcpyAsync(ctx->h_count, &Wv.counters[nxt], sizeof(int), cudaMemcpyDeviceToHost, stream));
                    B2PT_CUDA(ctx, cudaStreamSynchronize(stream));
                    A = *ctx->h_count;
                }
            }
            if (learn_batch) ctx->order_state = 1;
            k_resolve<<<(npc + 255) / 256, 256, 0, stream>>>(Wv, (float4*)accum, pix_begin, npc, ns);
            dbg("k_resolve", pix_begin, sb, -1);
        }
    }
    k_finalize<<<(unsigned)((nown + 255) / 256), 256, 0, stream>>>((const float4*)accum, F, nown, all_samples, divisor, d_rgb);
    ++launches;
    cudaError_t le = cudaGetLastError();
    unsigned long long totals[3] = {0, 0, 0};
    cudaError_t ce = cudaMemcpyAsync(totals, Wv.totals, sizeof(totals), cudaMemcpyDeviceToHost, stream);
    cudaError_t se = cudaStreamSynchronize(stream);
    for (size_t i = 0; i + 3 < ev_used; i += 4) {   // per bounce: before extend, after extend, after the epilogue, after direct
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, ctx->ev_pool[i], ctx->ev_pool[i + 1]) == cudaSuccess) extend_ms += ms;
        if (cudaEventElapsedTime(&ms, ctx->ev_pool[i + 1], ctx->ev_pool[i + 2]) == cudaSuccess) order_ms += ms;
        if (cudaEventElapsedTime(&ms, ctx->ev_pool[i + 2], ctx->ev_pool[i + 3]) == cudaSuccess) shadow_ms += ms;
    }
    if (!dbg_fault.empty()) { ctx->err = "render kernel fault: " + dbg_fault; return B2PT_ERR_CUDA; }
    if (le != cudaSuccess) { cuda_fail(ctx, le, "render kernels", __FILE__, __LINE__); return B2PT_ERR_CUDA; }
    if (ce != cudaSuccess) { cuda_fail(ctx, ce, "cudaMemcpyAsync(totals)", __FILE__, __LINE__); return B2PT_ERR_CUDA; }
    if (se != cudaSuccess) { cuda_fail(ctx, se, "cudaStreamSynchronize", __FILE__, __LINE__); return B2PT_ERR_CUDA; }
    ctx->stats.extend_rays = (int64_t)totals[0];
    ctx->stats.shadow_rays = (int64_t)totals[1];
    ctx->stats.fallback_rays = (int64_t)totals[2];
    ctx->stats.kernel_launches = launches;
    ctx->stats.extend_seconds = extend_ms * 1e-3;
    ctx->stats.shadow_seconds = shadow_ms * 1e-3;
    ctx->stats.order_seconds = order_ms * 1e-3;
    ctx->stats.trace_seconds = (extend_ms + shadow_ms) * 1e-3;
    ctx->stats.extend_launches = n_extend;
    ctx->stats.shadow_launches = n_shadow;
    return B2PT_OK;
}

namespace {
// (unsigned char)(powf(clamp(c), 1/gamma) * 255) on the host, as src/renderer.cpp:11-16 computes it
inline int reference_byte(float c, float inv_gamma) {
    c = c < 0.0f ? 0.0f : c; c = 1.0f < c ? 1.0f : c;
    return (int)(unsigned char)(std::pow(c, inv_gamma) * 255.0f);
}
}  // namespace

// thr[k] = smallest float in [0, 1] whose reference byte is >= k (thr[0] = 0).  powf is monotone in practice; the
// neighbourhood of every threshold is re-checked and a violation refuses the device path (the caller tonemaps on the host).
int tonemap_thresholds(float gamma, float* thr) {
    const float inv = 1.0f / gamma;
    thr[0] = 0.0f;
    const int top = reference_byte(1.0f, inv);
    for (int k = 1; k < 256; ++k) {
        if (k > top) { thr[k] = __builtin_inff(); continue; }
        uint32_t lo = 0u, hi = 0x3f800000u;   // bit patterns of 0.0f .. 1.0f: ordered like the floats
        while (lo < hi) {
            uint32_t mid = lo + (hi - lo) / 2;
            float x; std::memcpy(&x, &mid, 4);
            if (reference_byte(x, inv) >= k) hi = mid; else lo = mid + 1;
        }
        std::memcpy(&thr[k], &lo, 4);
        for (int d = -48; d <= 48; ++d) {   // monotone around the threshold?
            int64_t b = (int64_t)lo + d;
            if (b < 0 || b > 0x3f800000) continue;
            uint32_t ub = (uint32_t)b; float x; std::memcpy(&x, &ub, 4);
            if ((reference_byte(x, inv) >= k) != (d >= 0)) return B2PT_ERR_INVALID;
        }
    }
    return B2PT_OK;
}

int tonemap_frame(b2pt_ctx* ctx, const float* d_rgb, int width, int height, float gamma, int flip, uint8_t* rgb8_host) {
    if (!(gamma > 0.0f)) { ctx->err = "b2pt_tonemap: gamma must be positive"; return B2PT_ERR_INVALID; }
    const long long n = 3ll * width * height;
    void *d_out = nullptr, *d_thr = nullptr;
    int rc = scratch_reserve(ctx, 11, (size_t)std::max<long long>(n, 1), &d_out);
    if (rc) return rc;
    if ((rc = scratch_reserve(ctx, 12, 256 * sizeof(float), &d_thr))) return rc;
    if (ctx->tonemap_gamma != gamma) {
        if (tonemap_thresholds(gamma, ctx->tonemap_thr)) { ctx->err = "b2pt_tonemap: host powf is not monotone around a byte threshold for this gamma"; return B2PT_ERR_INVALID; }
        ctx->tonemap_gamma = gamma;
    }
    B2PT_CUDA(ctx, cudaMemcpyAsync(d_thr, ctx->tonemap_thr, 256 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    if (n > 0) k_tonemap<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_rgb, width, height, flip, (const float*)d_thr, (uint8_t*)d_out);
    B2PT_CUDA(ctx, cudaGetLastError());
    B2PT_CUDA(ctx, cudaMemcpyAsync(rgb8_host, d_out, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    B2PT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B2PT_OK;
}

}  // namespace b2pt
