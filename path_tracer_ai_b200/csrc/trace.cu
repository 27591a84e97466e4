// trace.cu — batch ray queries: replaces Scene::intersect (reference include/scene.hpp:96-99) for
// caller-supplied ray batches (BASELINE config 4: the closest-hit microbench).
//
// Two-phase, bit-exact: k_closest_fast traverses the 8-wide BVH and certifies its answer; rays it
// cannot certify (bit-equal ties, winner on its own leaf box's entry face) are appended to a
// fallback list and re-run by k_closest_exact, the flattened reference recursion.
#include <algorithm>
#include "traverse.cuh"

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

__device__ __forceinline__ RayQ load_ray(const float* __restrict__ o, const float* __restrict__ d,
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

template <bool COUNT>
__global__ void __launch_bounds__(128) k_closest_fast(DeviceScene S, const float* __restrict__ o, const float* __restrict__ d,
                                                      const float* __restrict__ tmax, long long n,
                                                      int32_t* __restrict__ tri, float* __restrict__ t, float* __restrict__ uv,
                                                      int* __restrict__ fb_count, int* __restrict__ fb_list, TraceCounters* __restrict__ counters) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned n_nodes = 0, n_tris = 0;
    if (i < n) {
        RayQ r = load_ray(o, d, tmax, i);
        HitRec h;
        bool ok = closest_fast<COUNT>(S, r, h, n_nodes, n_tris);
        store_hit(h, i, tri, t, uv);
        if (!ok) fb_list[atomicAdd(fb_count, 1)] = (int)i;
    }
    if (COUNT) flush_counters(counters, n_nodes, n_tris);
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
        RayQ r = load_ray(o, d, tmax, i);
        HitRec h;
        closest_exact_dfs(S, r, h);
        store_hit(h, i, tri, t, uv);
    }
}

template <bool COUNT>
__global__ void __launch_bounds__(128) k_any_fast(DeviceScene S, const float* __restrict__ o, const float* __restrict__ d,
                                                  const float* __restrict__ tmax, long long n, uint8_t* __restrict__ occ,
                                                  TraceCounters* __restrict__ counters) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned n_nodes = 0, n_tris = 0;
    if (i < n) {
        RayQ r = load_ray(o, d, tmax, i);
        occ[i] = any_fast<COUNT>(S, r, n_nodes, n_tris) ? 1 : 0;
    }
    if (COUNT) flush_counters(counters, n_nodes, n_tris);
}

}  // namespace

int launch_trace_closest(b2pt_ctx* ctx, const float* d_o, const float* d_d, const float* d_tmax, int64_t n,
                         int32_t* d_tri, float* d_t, float* d_uv) {
    if (n <= 0) return B2PT_OK;
    cudaStream_t st = ctx->stream;
    const int B = 128;
    // Rays are processed in launches of at most 2^30 so the int fallback list can index them.
    const int64_t chunk = 1ll << 30;
    for (int64_t off = 0; off < n; off += chunk) {
        int64_t m = std::min(chunk, n - off);
        const float* o = d_o + 3 * off; const float* d = d_d + 3 * off;
        const float* tm = d_tmax ? d_tmax + off : nullptr;
        int32_t* tri = d_tri + off; float* t = d_t ? d_t + off : nullptr; float* uv = d_uv ? d_uv + 2 * off : nullptr;
        if (ctx->flags & B2PT_FLAG_EXACT_ONLY) {
            int grid = (int)std::min<int64_t>((m + B - 1) / B, (int64_t)ctx->sm_count * 64);
            k_closest_exact<<<grid, B, 0, st>>>(ctx->scene, o, d, tm, m, tri, t, uv, nullptr, nullptr, ctx->d_counters);
            ctx->stats.kernel_launches += 1;
        } else {
            void* fb = nullptr;
            int rc = scratch_reserve(ctx, 0, sizeof(int) * (size_t)m, &fb);
            if (rc) return rc;
            B2PT_CUDA(ctx, cudaMemsetAsync(ctx->d_fallback_count, 0, sizeof(int), st));
            unsigned grid = (unsigned)((m + B - 1) / B);
            if (ctx->flags & B2PT_FLAG_COUNT_FETCHES)
                k_closest_fast<true><<<grid, B, 0, st>>>(ctx->scene, o, d, tm, m, tri, t, uv, ctx->d_fallback_count, (int*)fb, ctx->d_counters);
            else
                k_closest_fast<false><<<grid, B, 0, st>>>(ctx->scene, o, d, tm, m, tri, t, uv, ctx->d_fallback_count, (int*)fb, ctx->d_counters);
            k_closest_exact<<<ctx->sm_count * 8, B, 0, st>>>(ctx->scene, o, d, tm, m, tri, t, uv, ctx->d_fallback_count, (const int*)fb, ctx->d_counters);
            ctx->stats.kernel_launches += 2;
        }
        B2PT_CUDA(ctx, cudaGetLastError());
    }
    return B2PT_OK;
}

int launch_trace_any(b2pt_ctx* ctx, const float* d_o, const float* d_d, const float* d_tmax, int64_t n, uint8_t* d_occ) {
    if (n <= 0) return B2PT_OK;
    const int B = 128;
    const int64_t chunk = 1ll << 30;
    for (int64_t off = 0; off < n; off += chunk) {
        int64_t m = std::min(chunk, n - off);
        unsigned grid = (unsigned)((m + B - 1) / B);
        const float* tm = d_tmax ? d_tmax + off : nullptr;
        if (ctx->flags & B2PT_FLAG_COUNT_FETCHES)
            k_any_fast<true><<<grid, B, 0, ctx->stream>>>(ctx->scene, d_o + 3 * off, d_d + 3 * off, tm, m, d_occ + off, ctx->d_counters);
        else
            k_any_fast<false><<<grid, B, 0, ctx->stream>>>(ctx->scene, d_o + 3 * off, d_d + 3 * off, tm, m, d_occ + off, ctx->d_counters);
        ctx->stats.kernel_launches += 1;
        B2PT_CUDA(ctx, cudaGetLastError());
    }
    return B2PT_OK;
}

}  // namespace b2pt
