// trace.cu — batch ray queries: replaces Scene::intersect (reference include/scene.hpp:96-99) for
// caller-supplied ray batches (BASELINE config 4: the closest-hit microbench).
//
// Two-phase, bit-exact: k_closest_thread traverses the 8-wide BVH with one ray per lane (persistent warps that
// refill idle lanes from a warp-local pool of ray indices) and certifies its answer; rays it cannot certify
// (bit-equal ties, winner on its own leaf box's entry face, stack overflow) are appended to a fallback list and
// re-run by k_closest_exact, the flattened reference recursion.
#include <algorithm>
#include <cstdlib>
#include "traverse_pool.cuh"

namespace b2pt {

namespace {

__device__ __forceinline__ void flush_counters(TraceCounters* c, unsigned n_nodes, unsigned n_tris) {
    // warp-aggregate then one atomic per warp
    for (int off = 16; off > 0; off >>= 1) {
        n_nodes += __shfl_down_sync(0xffffffffu, n_nodes, off);
        n_tris += __shfl_down_sync(0xffffffffu, n_tris, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&c->node_fetches, (unsigned long long)n_nodes);
        atomicAdd(&c->tri_fetches, (unsigned long long)n_tris);
    }
}

__device__ __forceinline__ RayQ load_ray(const DeviceScene& S, const float* __restrict__ o, const float* __restrict__ d,
                                         const float* __restrict__ tmax, long long i) {
    V3 ro = mk3(o[3 * i], o[3 * i + 1], o[3 * i + 2]);
    V3 rd = mk3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
    return make_rayq(ro, rd, tmax ? tmax[i] : B2PT_INF);
}

__device__ __forceinline__ void store_hit(const HitRec& h, long long i, int32_t* tri, float* t, float* uv) {
    tri[i] = h.tri;
    if (t) t[i] = h.t;
    if (uv) { uv[2 * i] = h.u; uv[2 * i + 1] = h.v; }
}

// Exact reference recursion over an index list (fallback) or over the whole batch (list == nullptr).
__global__ void __launch_bounds__(128) k_closest_exact(DeviceScene S, const float* __restrict__ o, const float* __restrict__ d,
                                                       const float* __restrict__ tmax, long long n,
                                                       int32_t* __restrict__ tri, float* __restrict__ t, float* __restrict__ uv,
                                                       const int* __restrict__ list_count, const int* __restrict__ list,
                                                       TraceCounters* __restrict__ counters) {
    long long total = list ? (long long)*list_count : n;
    if (list && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&counters->fallback, (unsigned long long)total);
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        long long i = list ? list[k] : k;
        RayQ r = load_ray(S, o, d, tmax, i);
        HitRec h;
        closest_exact_dfs(S, r, h);
        store_hit(h, i, tri, t, uv);
    }
}

// ---- one ray per lane, persistent warps with lane refill ------------------------------------------------
#define B2PT_POOL_CHUNK 256     // ray indices a warp claims per global atomic
#define B2PT_REFILL_MIN 4       // refill as soon as this many lanes are idle
#define B2PT_TPS 4              // triangles per step
#define B2PT_BLOCKS_PER_SM 12   // persistent blocks per SM

template <bool COUNT, int TPS>
__global__ void __launch_bounds__(B2PT_TBLOCK) k_closest_thread(DeviceScene S, const float* __restrict__ o, const float* __restrict__ d,
                                                        const float* __restrict__ tmax, long long n,
                                                        int32_t* __restrict__ tri, float* __restrict__ t, float* __restrict__ uv,
                                                        unsigned long long* __restrict__ next_ray,
                                                        int* __restrict__ fb_count, int* __restrict__ fb_list,
                                                        TraceCounters* __restrict__ counters) {
    __shared__ uint2 lane_stacks[B2PT_SSTACK * B2PT_TBLOCK];
    LaneState st;
    st.stack.sm = lane_stacks + threadIdx.x;
    WarpPool pool{0, 0, false};
    long long idx = -1;
    unsigned n_nodes = 0, n_tris = 0;
    if (S.nwide == 0 && S.nhoist == 0) {   // empty scene: everything misses
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
            HitRec h{B2PT_INF, -1, 0.0f, 0.0f};
            store_hit(h, i, tri, t, uv);
        }
        return;
    }
    while (true) {
        unsigned idle = __ballot_sync(0xffffffffu, idx < 0);
        if (idle && (__popc(idle) >= B2PT_REFILL_MIN || idle == 0xffffffffu)) {
            long long got = warp_pool_take<B2PT_POOL_CHUNK>(pool, next_ray, n, idx < 0);
            if (idx < 0 && got >= 0) {
                idx = got;
                lane_begin<false, COUNT>(S, st, load_ray(S, o, d, tmax, idx), n_tris);
            }
            if (__ballot_sync(0xffffffffu, idx >= 0) == 0) break;   // nothing left anywhere
        }
        if (idx >= 0) {
            if (lane_closest_step<COUNT, TPS>(S, st, n_nodes, n_tris)) {
                store_hit(st.best, idx, tri, t, uv);
                if (!lane_certify(S, st)) fb_list[atomicAdd(fb_count, 1)] = (int)idx;
                idx = -1;
            }
        }
    }
    if (COUNT) flush_counters(counters, n_nodes, n_tris);
}

template <bool COUNT, int TPS>
__global__ void __launch_bounds__(B2PT_TBLOCK) k_any_thread(DeviceScene S, const float* __restrict__ o, const float* __restrict__ d,
                                                    const float* __restrict__ tmax, long long n, uint8_t* __restrict__ occ,
                                                    unsigned long long* __restrict__ next_ray, TraceCounters* __restrict__ counters) {
    __shared__ uint2 lane_stacks[B2PT_SSTACK * B2PT_TBLOCK];
    LaneState st;
    st.stack.sm = lane_stacks + threadIdx.x;
    WarpPool pool{0, 0, false};
    long long idx = -1;
    unsigned n_nodes = 0, n_tris = 0;
    if (S.nwide == 0 && S.nhoist == 0) {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) occ[i] = 0;
        return;
    }
    while (true) {
        unsigned idle = __ballot_sync(0xffffffffu, idx < 0);
        if (idle && (__popc(idle) >= B2PT_REFILL_MIN || idle == 0xffffffffu)) {
            long long got = warp_pool_take<B2PT_POOL_CHUNK>(pool, next_ray, n, idx < 0);
            if (idx < 0 && got >= 0) {
                idx = got;
                if (lane_begin<true, COUNT>(S, st, load_ray(S, o, d, tmax, idx), n_tris)) { occ[idx] = 1; idx = -1; }
            }
            if (__ballot_sync(0xffffffffu, idx >= 0) == 0 && pool.exhausted) break;
        }
        if (idx >= 0) {
            int res = lane_any_step<COUNT, TPS>(S, st, n_nodes, n_tris);
            if (res) {
                if (res == 3) { HitRec h; closest_exact_dfs(S, st.r, h); res = h.tri >= 0 ? 1 : 2; }
                occ[idx] = res == 1 ? 1 : 0;
                idx = -1;
            }
        }
    }
    if (COUNT) flush_counters(counters, n_nodes, n_tris);
}

// ---- several rays per lane, phase-split steps (traverse_pool.cuh) --------------------------------------------
struct IoTraceClosest {
    const float *o, *d, *tmax;
    int32_t* tri; float* t; float* uv;
    int* fb_count; int* fb_list;
    __device__ __forceinline__ bool load(const DeviceScene&, long long k, V3& ro, V3& rd, float& T0, int& tag) const {
        ro = mk3(o[3 * k], o[3 * k + 1], o[3 * k + 2]);
        rd = vnormalize(mk3(d[3 * k], d[3 * k + 1], d[3 * k + 2]));   // the Ray ctor's normalisation (ray.hpp:12)
        T0 = tmax ? tmax[k] : B2PT_INF;
        tag = (int)k;
        return true;
    }
    __device__ __forceinline__ void store_closest(int k, const HitRec& h, bool certified) const {
        store_hit(h, k, tri, t, uv);
        if (!certified) fb_list[atomicAdd(fb_count, 1)] = k;
    }
    __device__ __forceinline__ void store_any(int, bool) const {}
};
struct IoTraceAny {
    const float *o, *d, *tmax;
    uint8_t* occ;
    __device__ __forceinline__ bool load(const DeviceScene&, long long k, V3& ro, V3& rd, float& T0, int& tag) const {
        ro = mk3(o[3 * k], o[3 * k + 1], o[3 * k + 2]);
        rd = vnormalize(mk3(d[3 * k], d[3 * k + 1], d[3 * k + 2]));
        T0 = tmax ? tmax[k] : B2PT_INF;
        tag = (int)k;
        return true;
    }
    __device__ __forceinline__ void store_closest(int, const HitRec&, bool) const {}
    __device__ __forceinline__ void store_any(int k, bool occluded) const { occ[k] = occluded ? 1 : 0; }
};

template <bool COUNT>
__global__ void __launch_bounds__(B2PT_PBLOCK, B2PT_PMINB) k_closest_pool(DeviceScene S, IoTraceClosest io, long long n, unsigned long long* __restrict__ next_ray,
                                                              TraceCounters* __restrict__ counters) {
    __shared__ PoolSmem<false> sm;
    pool_traverse<false, COUNT>(S, sm, io, next_ray, n, counters);
}
template <bool COUNT>
__global__ void __launch_bounds__(B2PT_PBLOCK, B2PT_PMINB) k_any_pool(DeviceScene S, IoTraceAny io, long long n, unsigned long long* __restrict__ next_ray,
                                                          TraceCounters* __restrict__ counters) {
    __shared__ PoolSmem<true> sm;
    pool_traverse<true, COUNT>(S, sm, io, next_ray, n, counters);
}

}  // namespace

int launch_trace_closest(b2pt_ctx* ctx, const float* d_o, const float* d_d, const float* d_tmax, int64_t n,
                         int32_t* d_tri, float* d_t, float* d_uv) {
    if (n <= 0) return B2PT_OK;
    cudaStream_t st = ctx->stream;
    // Rays are processed in launches of at most 2^30 so the int fallback list can index them.
    const int64_t chunk = 1ll << 30;
    unsigned long long* next_ray = reinterpret_cast<unsigned long long*>(ctx->d_fallback_count + 2);
    for (int64_t off = 0; off < n; off += chunk) {
        int64_t m = std::min(chunk, n - off);
        const float* o = d_o + 3 * off; const float* d = d_d + 3 * off;
        const float* tm = d_tmax ? d_tmax + off : nullptr;
        int32_t* tri = d_tri + off; float* t = d_t ? d_t + off : nullptr; float* uv = d_uv ? d_uv + 2 * off : nullptr;
        if (ctx->flags & B2PT_FLAG_EXACT_ONLY) {
            int grid = (int)std::min<int64_t>((m + 127) / 128, (int64_t)ctx->sm_count * 64);
            k_closest_exact<<<grid, 128, 0, st>>>(ctx->scene, o, d, tm, m, tri, t, uv, nullptr, nullptr, ctx->d_counters);
            ctx->stats.kernel_launches += 1;
        } else {
            void* fb = nullptr;
            int rc = scratch_reserve(ctx, 0, sizeof(int) * (size_t)m, &fb);
            if (rc) return rc;
            B2PT_CUDA(ctx, cudaMemsetAsync(ctx->d_fallback_count, 0, 64, st));
            unsigned tgrid = (unsigned)std::min<int64_t>((m + B2PT_TBLOCK - 1) / B2PT_TBLOCK, (int64_t)ctx->sm_count * B2PT_BLOCKS_PER_SM);
            if (ctx->flags & B2PT_FLAG_POOL_EXTEND) {   // closest hit: the pool kernel only matches the per-lane kernel (1973 vs 1989 Mrays/s)
                IoTraceClosest io{o, d, tm, tri, t, uv, ctx->d_fallback_count, (int*)fb};
                unsigned pgrid = (unsigned)std::min<int64_t>((m + B2PT_PBLOCK * B2PT_PR - 1) / (B2PT_PBLOCK * B2PT_PR), (int64_t)ctx->sm_count * 8);
                if (ctx->flags & B2PT_FLAG_COUNT_FETCHES) k_closest_pool<true><<<pgrid, B2PT_PBLOCK, 0, st>>>(ctx->scene, io, m, next_ray, ctx->d_counters);
                else k_closest_pool<false><<<pgrid, B2PT_PBLOCK, 0, st>>>(ctx->scene, io, m, next_ray, ctx->d_counters);
            } else if (ctx->flags & B2PT_FLAG_COUNT_FETCHES)
                k_closest_thread<true, B2PT_TPS><<<tgrid, B2PT_TBLOCK, 0, st>>>(ctx->scene, o, d, tm, m, tri, t, uv, next_ray, ctx->d_fallback_count, (int*)fb, ctx->d_counters);
            else
                k_closest_thread<false, B2PT_TPS><<<tgrid, B2PT_TBLOCK, 0, st>>>(ctx->scene, o, d, tm, m, tri, t, uv, next_ray, ctx->d_fallback_count, (int*)fb, ctx->d_counters);
            k_closest_exact<<<ctx->sm_count * 8, 128, 0, st>>>(ctx->scene, o, d, tm, m, tri, t, uv, ctx->d_fallback_count, (const int*)fb, ctx->d_counters);
            ctx->stats.kernel_launches += 2;
        }
        B2PT_CUDA(ctx, cudaGetLastError());
    }
    return B2PT_OK;
}

int launch_trace_any(b2pt_ctx* ctx, const float* d_o, const float* d_d, const float* d_tmax, int64_t n, uint8_t* d_occ) {
    if (n <= 0) return B2PT_OK;
    const int64_t chunk = 1ll << 30;
    unsigned long long* next_ray = reinterpret_cast<unsigned long long*>(ctx->d_fallback_count + 2);
    for (int64_t off = 0; off < n; off += chunk) {
        int64_t m = std::min(chunk, n - off);
        const float* tm = d_tmax ? d_tmax + off : nullptr;
        B2PT_CUDA(ctx, cudaMemsetAsync(ctx->d_fallback_count, 0, 64, ctx->stream));
        unsigned tgrid = (unsigned)std::min<int64_t>((m + B2PT_TBLOCK - 1) / B2PT_TBLOCK, (int64_t)ctx->sm_count * B2PT_BLOCKS_PER_SM);
        if (!(ctx->flags & B2PT_FLAG_LANE_KERNELS)) {
            IoTraceAny io{d_o + 3 * off, d_d + 3 * off, tm, d_occ + off};
            unsigned pgrid = (unsigned)std::min<int64_t>((m + B2PT_PBLOCK * B2PT_PR - 1) / (B2PT_PBLOCK * B2PT_PR), (int64_t)ctx->sm_count * 8);
            if (ctx->flags & B2PT_FLAG_COUNT_FETCHES) k_any_pool<true><<<pgrid, B2PT_PBLOCK, 0, ctx->stream>>>(ctx->scene, io, m, next_ray, ctx->d_counters);
            else k_any_pool<false><<<pgrid, B2PT_PBLOCK, 0, ctx->stream>>>(ctx->scene, io, m, next_ray, ctx->d_counters);
        } else if (ctx->flags & B2PT_FLAG_COUNT_FETCHES)
            k_any_thread<true, B2PT_TPS><<<tgrid, B2PT_TBLOCK, 0, ctx->stream>>>(ctx->scene, d_o + 3 * off, d_d + 3 * off, tm, m, d_occ + off, next_ray, ctx->d_counters);
        else
            k_any_thread<false, B2PT_TPS><<<tgrid, B2PT_TBLOCK, 0, ctx->stream>>>(ctx->scene, d_o + 3 * off, d_d + 3 * off, tm, m, d_occ + off, next_ray, ctx->d_counters);
        ctx->stats.kernel_launches += 1;
        B2PT_CUDA(ctx, cudaGetLastError());
    }
    return B2PT_OK;
}

}  // namespace b2pt
