// api.cu — the C ABI declared in include/b2pt.h.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "ctx.cuh"

namespace {
thread_local std::string g_create_error;
}

namespace b2pt {

bool cuda_fail(b2pt_ctx* ctx, cudaError_t e, const char* call, const char* file, int line) {
    char buf[1024];
    std::snprintf(buf, sizeof(buf), "CUDA call (%s) failed with error: '%s' (%s:%d)", call, cudaGetErrorString(e), file, line);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return false;
}

int scratch_reserve(b2pt_ctx* ctx, int slot, size_t bytes, void** out) {
    if (ctx->scratch_bytes[slot] < bytes) {
        if (ctx->scratch[slot]) { cudaFree(ctx->scratch[slot]); ctx->scratch[slot] = nullptr; ctx->scratch_bytes[slot] = 0; }
        size_t want = bytes + bytes / 8 + 256;
        B2PT_CUDA(ctx, cudaMalloc(&ctx->scratch[slot], want));
        ctx->scratch_bytes[slot] = want;
    }
    *out = ctx->scratch[slot];
    return B2PT_OK;
}

}  // namespace b2pt

using namespace b2pt;

namespace {

void begin_call(b2pt_ctx* ctx) {
    double build = ctx->stats.build_seconds;
    ctx->stats = b2pt_stats{};
    ctx->stats.build_seconds = build;
    cudaMemsetAsync(ctx->d_counters, 0, sizeof(TraceCounters), ctx->stream);
    cudaEventRecord(ctx->ev0, ctx->stream);
}

int end_call(b2pt_ctx* ctx) {
    TraceCounters c{};
    B2PT_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    B2PT_CUDA(ctx, cudaMemcpyAsync(&c, ctx->d_counters, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
    B2PT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0.0f;
    B2PT_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->stats.gpu_seconds = ms * 1e-3;
    ctx->stats.fallback_rays += (int64_t)c.fallback;
    ctx->stats.node_fetches += (int64_t)c.node_fetches;
    ctx->stats.tri_fetches += (int64_t)c.tri_fetches;
    return B2PT_OK;
}

int require_scene(b2pt_ctx* ctx, const char* who) {
    if (!ctx) return B2PT_ERR_INVALID;
    if (!ctx->has_scene) {
        ctx->err = std::string(who) + ": no scene uploaded (call b2pt_upload_scene first)";
        return B2PT_ERR_INVALID;
    }
    return B2PT_OK;
}

}  // namespace

extern "C" {

const char* b2pt_version(void) { return "b2pt 0.1 (sm_100a)"; }

int b2pt_create(const b2pt_config* cfg, b2pt_ctx** out) {
    if (!out) return B2PT_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("b2pt_create: no CUDA device available (") + cudaGetErrorString(e) +
                         "); this engine has no CPU fallback";
        return B2PT_ERR_NO_DEVICE;
    }
    int dev = cfg ? cfg->device : 0;
    if (dev < 0 || dev >= ndev) { g_create_error = "b2pt_create: device ordinal out of range"; return B2PT_ERR_INVALID; }
    cudaDeviceProp prop{};
    if ((e = cudaGetDeviceProperties(&prop, dev)) != cudaSuccess) { cuda_fail(nullptr, e, "cudaGetDeviceProperties", __FILE__, __LINE__); return B2PT_ERR_CUDA; }
    if (prop.major != 10) {
        char buf[512];
        std::snprintf(buf, sizeof(buf), "b2pt_create: device %d (%s) is sm_%d%d; this library is built for sm_100a only", dev, prop.name, prop.major, prop.minor);
        g_create_error = buf;
        return B2PT_ERR_NO_DEVICE;
    }
    b2pt_ctx* ctx = new b2pt_ctx();
    ctx->device = dev;
    ctx->flags = cfg ? cfg->flags : 0;
    ctx->learn_order = (ctx->flags & B2PT_FLAG_NO_LEARN_ORDER) == 0;
    ctx->debug_sync = std::getenv("B2PT_DEBUG_SYNC") != nullptr;   // read once, here
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_paths = (cfg && cfg->max_paths_in_flight > 0) ? cfg->max_paths_in_flight : (int64_t)(32 << 20);   // larger wavefronts amortise the kernels' tails and make the sorted bounces denser: 1M-triangle render 387 (16M) / 412 (32M) / 419 (64M) / 421 (128M) Msamples/s; ~190 B of scratch per path
    auto fail = [&](cudaError_t err, const char* what) {
        cuda_fail(nullptr, err, what, __FILE__, __LINE__);
        b2pt_destroy(ctx);   // releases whatever was created so far
        return B2PT_ERR_CUDA;
    };
    if ((e = cudaSetDevice(dev)) != cudaSuccess) return fail(e, "cudaSetDevice");
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate");
    if ((e = cudaEventCreate(&ctx->ev0)) != cudaSuccess) return fail(e, "cudaEventCreate");
    if ((e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) return fail(e, "cudaEventCreate");
    if ((e = cudaEventCreate(&ctx->ev2)) != cudaSuccess) return fail(e, "cudaEventCreate");
    if ((e = cudaEventCreate(&ctx->ev3)) != cudaSuccess) return fail(e, "cudaEventCreate");
    if ((e = cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate");
    if ((e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return fail(e, "cudaEventCreate");
    if ((e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming)) != cudaSuccess) return fail(e, "cudaEventCreate");
    if ((e = cudaHostAlloc((void**)&ctx->h_count, 64, cudaHostAllocDefault)) != cudaSuccess) return fail(e, "cudaHostAlloc");
    if ((e = cudaMalloc(&ctx->d_counters, sizeof(TraceCounters))) != cudaSuccess) return fail(e, "cudaMalloc");
    if ((e = cudaMalloc(&ctx->d_fallback_count, 64)) != cudaSuccess) return fail(e, "cudaMalloc");
    *out = ctx;
    return B2PT_OK;
}

void b2pt_destroy(b2pt_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->side) cudaStreamSynchronize(ctx->side);
    free_scene(ctx);
    for (int i = 0; i < B2PT_SCRATCH_SLOTS; ++i) if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    if (ctx->d_fallback_count) cudaFree(ctx->d_fallback_count);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev2) cudaEventDestroy(ctx->ev2);
    if (ctx->ev3) cudaEventDestroy(ctx->ev3);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->h_count) cudaFreeHost(ctx->h_count);
    if (ctx->side) cudaStreamDestroy(ctx->side);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* b2pt_last_error(const b2pt_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int b2pt_upload_scene(b2pt_ctx* ctx, const float* pos, const float* nrm, const int32_t* mat, int64_t ntri,
                      const b2pt_material* mats, int32_t nmat, const b2pt_light* lights, int32_t nlight) {
    if (!ctx) return B2PT_ERR_INVALID;
    B2PT_CUDA(ctx, cudaSetDevice(ctx->device));
    if ((nmat > 0 && !mats) || (nlight > 0 && !lights)) { ctx->err = "b2pt_upload_scene: NULL materials / lights"; return B2PT_ERR_INVALID; }
    ctx->stats = b2pt_stats{};
    return build_scene(ctx, pos, nrm, mat, ntri, mats, nmat, lights, nlight);
}

int b2pt_trace_closest_device(b2pt_ctx* ctx, const float* d_o, const float* d_d, const float* d_tmax, int64_t n,
                              int32_t* d_tri, float* d_t, float* d_uv) {
    int rc = require_scene(ctx, "b2pt_trace_closest");
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_o || !d_d || !d_tri))) { ctx->err = "b2pt_trace_closest: bad arguments"; return B2PT_ERR_INVALID; }
    B2PT_CUDA(ctx, cudaSetDevice(ctx->device));
    begin_call(ctx);
    cudaEventRecord(ctx->ev2, ctx->stream);
    if ((rc = launch_trace_closest(ctx, d_o, d_d, d_tmax, n, d_tri, d_t, d_uv))) return rc;
    cudaEventRecord(ctx->ev3, ctx->stream);
    ctx->stats.extend_rays = n;
    if ((rc = end_call(ctx))) return rc;
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3);
    ctx->stats.trace_seconds = ctx->stats.extend_seconds = ms * 1e-3;
    ctx->stats.extend_launches = (n + (1ll << 30) - 1) >> 30;
    return B2PT_OK;
}

int b2pt_trace_any_device(b2pt_ctx* ctx, const float* d_o, const float* d_d, const float* d_tmax, int64_t n, uint8_t* d_occ) {
    int rc = require_scene(ctx, "b2pt_trace_any");
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_o || !d_d || !d_occ))) { ctx->err = "b2pt_trace_any: bad arguments"; return B2PT_ERR_INVALID; }
    B2PT_CUDA(ctx, cudaSetDevice(ctx->device));
    begin_call(ctx);
    cudaEventRecord(ctx->ev2, ctx->stream);
    if ((rc = launch_trace_any(ctx, d_o, d_d, d_tmax, n, d_occ))) return rc;
    cudaEventRecord(ctx->ev3, ctx->stream);
    ctx->stats.shadow_rays = n;
    if ((rc = end_call(ctx))) return rc;
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3);
    ctx->stats.trace_seconds = ctx->stats.shadow_seconds = ms * 1e-3;
    ctx->stats.shadow_launches = (n + (1ll << 30) - 1) >> 30;
    return B2PT_OK;
}

// Host-buffer variants: staged through device scratch in chunks so 100M-ray batches do not need
// 4 GB of extra HBM at once.
int b2pt_trace_closest(b2pt_ctx* ctx, const float* o, const float* d, const float* tmax, int64_t n,
                       int32_t* tri, float* t, float* uv) {
    int rc = require_scene(ctx, "b2pt_trace_closest");
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!o || !d || !tri))) { ctx->err = "b2pt_trace_closest: bad arguments"; return B2PT_ERR_INVALID; }
    B2PT_CUDA(ctx, cudaSetDevice(ctx->device));
    begin_call(ctx);
    const int64_t chunk = 32ll << 20;
    double trace_s = 0.0;
    for (int64_t off = 0; off < n; off += chunk) {
        int64_t m = std::min(chunk, n - off);
        void *d_o, *d_d, *d_tm = nullptr, *d_tri, *d_t, *d_uv;
        if ((rc = scratch_reserve(ctx, 1, sizeof(float) * 3 * m, &d_o))) return rc;
        if ((rc = scratch_reserve(ctx, 2, sizeof(float) * 3 * m, &d_d))) return rc;
        if (tmax && (rc = scratch_reserve(ctx, 3, sizeof(float) * m, &d_tm))) return rc;
        if ((rc = scratch_reserve(ctx, 4, sizeof(int32_t) * m, &d_tri))) return rc;
        if ((rc = scratch_reserve(ctx, 5, sizeof(float) * m, &d_t))) return rc;
        if ((rc = scratch_reserve(ctx, 6, sizeof(float) * 2 * m, &d_uv))) return rc;
        B2PT_CUDA(ctx, cudaMemcpyAsync(d_o, o + 3 * off, sizeof(float) * 3 * m, cudaMemcpyHostToDevice, ctx->stream));
        B2PT_CUDA(ctx, cudaMemcpyAsync(d_d, d + 3 * off, sizeof(float) * 3 * m, cudaMemcpyHostToDevice, ctx->stream));
        if (tmax) B2PT_CUDA(ctx, cudaMemcpyAsync(d_tm, tmax + off, sizeof(float) * m, cudaMemcpyHostToDevice, ctx->stream));
        B2PT_CUDA(ctx, cudaEventRecord(ctx->ev2, ctx->stream));
        if ((rc = launch_trace_closest(ctx, (float*)d_o, (float*)d_d, (float*)d_tm, m, (int32_t*)d_tri, (float*)d_t, (float*)d_uv))) return rc;
        B2PT_CUDA(ctx, cudaEventRecord(ctx->ev3, ctx->stream));
        B2PT_CUDA(ctx, cudaMemcpyAsync(tri + off, d_tri, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, ctx->stream));
        if (t) B2PT_CUDA(ctx, cudaMemcpyAsync(t + off, d_t, sizeof(float) * m, cudaMemcpyDeviceToHost, ctx->stream));
        if (uv) B2PT_CUDA(ctx, cudaMemcpyAsync(uv + 2 * off, d_uv, sizeof(float) * 2 * m, cudaMemcpyDeviceToHost, ctx->stream));
        B2PT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3);
        trace_s += ms * 1e-3;
    }
    ctx->stats.extend_rays = n;
    if ((rc = end_call(ctx))) return rc;
    ctx->stats.trace_seconds = ctx->stats.extend_seconds = trace_s;
    ctx->stats.extend_launches = (n + chunk - 1) / chunk;
    return B2PT_OK;
}

int b2pt_trace_any(b2pt_ctx* ctx, const float* o, const float* d, const float* tmax, int64_t n, uint8_t* occluded) {
    int rc = require_scene(ctx, "b2pt_trace_any");
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!o || !d || !occluded))) { ctx->err = "b2pt_trace_any: bad arguments"; return B2PT_ERR_INVALID; }
    B2PT_CUDA(ctx, cudaSetDevice(ctx->device));
    begin_call(ctx);
    const int64_t chunk = 32ll << 20;
    double trace_s = 0.0;
    for (int64_t off = 0; off < n; off += chunk) {
        int64_t m = std::min(chunk, n - off);
        void *d_o, *d_d, *d_tm = nullptr, *d_occ;
        if ((rc = scratch_reserve(ctx, 1, sizeof(float) * 3 * m, &d_o))) return rc;
        if ((rc = scratch_reserve(ctx, 2, sizeof(float) * 3 * m, &d_d))) return rc;
        if (tmax && (rc = scratch_reserve(ctx, 3, sizeof(float) * m, &d_tm))) return rc;
        if ((rc = scratch_reserve(ctx, 4, m, &d_occ))) return rc;
        B2PT_CUDA(ctx, cudaMemcpyAsync(d_o, o + 3 * off, sizeof(float) * 3 * m, cudaMemcpyHostToDevice, ctx->stream));
        B2PT_CUDA(ctx, cudaMemcpyAsync(d_d, d + 3 * off, sizeof(float) * 3 * m, cudaMemcpyHostToDevice, ctx->stream));
        if (tmax) B2PT_CUDA(ctx, cudaMemcpyAsync(d_tm, tmax + off, sizeof(float) * m, cudaMemcpyHostToDevice, ctx->stream));
        B2PT_CUDA(ctx, cudaEventRecord(ctx->ev2, ctx->stream));
        if ((rc = launch_trace_any(ctx, (float*)d_o, (float*)d_d, (float*)d_tm, m, (uint8_t*)d_occ))) return rc;
        B2PT_CUDA(ctx, cudaEventRecord(ctx->ev3, ctx->stream));
        B2PT_CUDA(ctx, cudaMemcpyAsync(occluded + off, d_occ, m, cudaMemcpyDeviceToHost, ctx->stream));
        B2PT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3);
        trace_s += ms * 1e-3;
    }
    ctx->stats.shadow_rays = n;
    if ((rc = end_call(ctx))) return rc;
    ctx->stats.trace_seconds = ctx->stats.shadow_seconds = trace_s;
    ctx->stats.shadow_launches = (n + chunk - 1) / chunk;
    return B2PT_OK;
}

int b2pt_render_device(b2pt_ctx* ctx, const b2pt_camera* cam, const b2pt_settings* settings, uint64_t seed,
                       const b2pt_partition* part, float* d_rgb) {
    int rc = require_scene(ctx, "b2pt_render");
    if (rc) return rc;
    if (!cam || !settings || !d_rgb) { ctx->err = "b2pt_render: NULL argument"; return B2PT_ERR_INVALID; }
    if (settings->width < 2 || settings->height < 2 || settings->samples_per_pixel < 1 || settings->max_bounces < 0) {
        ctx->err = "b2pt_render: width/height must be >= 2, samples >= 1, bounces >= 0";
        return B2PT_ERR_INVALID;
    }
    B2PT_CUDA(ctx, cudaSetDevice(ctx->device));
    begin_call(ctx);
    if ((rc = render_frame(ctx, cam, settings, seed, part, d_rgb))) return rc;
    return end_call(ctx);
}

int b2pt_render(b2pt_ctx* ctx, const b2pt_camera* cam, const b2pt_settings* settings, uint64_t seed,
                const b2pt_partition* part, float* rgb) {
    if (!ctx) return B2PT_ERR_INVALID;
    if (!settings || !rgb) { ctx->err = "b2pt_render: NULL argument"; return B2PT_ERR_INVALID; }
    B2PT_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t bytes = sizeof(float) * 3ull * (size_t)std::max(settings->width, 0) * (size_t)std::max(settings->height, 0);
    void* d_rgb = nullptr;
    int rc = scratch_reserve(ctx, 7, bytes, &d_rgb);
    if (rc) return rc;
    if ((rc = b2pt_render_device(ctx, cam, settings, seed, part, (float*)d_rgb))) return rc;
    ctx->last_width = settings->width; ctx->last_height = settings->height;
    B2PT_CUDA(ctx, cudaMemcpyAsync(rgb, d_rgb, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    B2PT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B2PT_OK;
}

// ---- progressive rendering ---------------------------------------------------------------------------------------------
int b2pt_progressive_begin(b2pt_ctx* ctx, const b2pt_camera* cam, const b2pt_settings* settings, uint64_t seed) {
    int rc = require_scene(ctx, "b2pt_progressive_begin");
    if (rc) return rc;
    if (!cam || !settings) { ctx->err = "b2pt_progressive_begin: NULL argument"; return B2PT_ERR_INVALID; }
    if (settings->width < 2 || settings->height < 2 || settings->samples_per_pixel < 1 || settings->max_bounces < 0) {
        ctx->err = "b2pt_progressive_begin: width/height must be >= 2, samples >= 1, bounces >= 0";
        return B2PT_ERR_INVALID;
    }
    ctx->prog_active = true; ctx->prog_cam = *cam; ctx->prog_settings = *settings; ctx->prog_seed = seed; ctx->prog_done = 0;
    return B2PT_OK;
}

int b2pt_progressive_pass(b2pt_ctx* ctx, int32_t sample_count, float* rgb, int32_t* samples_done) {
    int rc = require_scene(ctx, "b2pt_progressive_pass");
    if (rc) return rc;
    if (!ctx->prog_active) { ctx->err = "b2pt_progressive_pass: call b2pt_progressive_begin first"; return B2PT_ERR_INVALID; }
    const b2pt_settings& st = ctx->prog_settings;
    const int left = st.samples_per_pixel - ctx->prog_done;
    if (sample_count <= 0 || left <= 0) { ctx->err = "b2pt_progressive_pass: no samples left (or sample_count <= 0)"; return B2PT_ERR_INVALID; }
    const int n = std::min<int>(sample_count, left);
    B2PT_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t bytes = sizeof(float) * 3ull * (size_t)st.width * (size_t)st.height;
    void* d_rgb = nullptr;
    if ((rc = scratch_reserve(ctx, 7, bytes, &d_rgb))) return rc;
    b2pt_partition part{};
    part.sample_begin = ctx->prog_done; part.sample_count = n;
    begin_call(ctx);
    // the per-pixel sums stay on the device between passes; the estimate is the running mean
    if ((rc = render_frame(ctx, &ctx->prog_cam, &st, ctx->prog_seed, &part, (float*)d_rgb, ctx->prog_done > 0, ctx->prog_done + n))) { ctx->prog_active = false; return rc; }
    if ((rc = end_call(ctx))) return rc;
    ctx->prog_done += n;
    ctx->last_width = st.width; ctx->last_height = st.height;
    if (samples_done) *samples_done = ctx->prog_done;
    if (rgb) {
        B2PT_CUDA(ctx, cudaMemcpyAsync(rgb, d_rgb, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        B2PT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return B2PT_OK;
}

int b2pt_tonemap(b2pt_ctx* ctx, const float* d_rgb, int32_t width, int32_t height, float gamma, int32_t flip, uint8_t* rgb8) {
    if (!ctx) return B2PT_ERR_INVALID;
    if (!d_rgb || !rgb8 || width < 0 || height < 0) { ctx->err = "b2pt_tonemap: bad arguments"; return B2PT_ERR_INVALID; }
    B2PT_CUDA(ctx, cudaSetDevice(ctx->device));
    return tonemap_frame(ctx, d_rgb, width, height, gamma, flip, rgb8);
}

int b2pt_tonemap_last(b2pt_ctx* ctx, float gamma, int32_t flip, uint8_t* rgb8) {
    if (!ctx) return B2PT_ERR_INVALID;
    if (!rgb8) { ctx->err = "b2pt_tonemap_last: NULL output"; return B2PT_ERR_INVALID; }
    if (ctx->last_width <= 0 || !ctx->scratch[7]) { ctx->err = "b2pt_tonemap_last: nothing rendered yet"; return B2PT_ERR_INVALID; }
    B2PT_CUDA(ctx, cudaSetDevice(ctx->device));
    return tonemap_frame(ctx, (const float*)ctx->scratch[7], ctx->last_width, ctx->last_height, gamma, flip, rgb8);
}

int b2pt_tonemap_thresholds(float gamma, float* thr256) {
    if (!thr256 || !(gamma > 0.0f)) return B2PT_ERR_INVALID;
    return tonemap_thresholds(gamma, thr256);
}

int b2pt_get_stats(const b2pt_ctx* ctx, b2pt_stats* out) {
    if (!ctx || !out) return B2PT_ERR_INVALID;
    *out = ctx->stats;
    return B2PT_OK;
}

int b2pt_get_accel_info(const b2pt_ctx* ctx, int64_t* out8) {
    if (!ctx || !out8) return B2PT_ERR_INVALID;
    std::memcpy(out8, ctx->accel_info, sizeof(ctx->accel_info));
    return B2PT_OK;
}

void* b2pt_stream(const b2pt_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

}  // extern "C"
