// multi.cu — several GPUs behind ONE renderer object, in one process (the b2pt_multi_* entry points of include/b2pt.h).
//
// The reference's GPU renderer is one object in one process (src/gpu/optix_renderer.cu:439-451); so is this one.  The
// scene is replicated: every device gets its own context (api.cu) driven by its own host thread, the frame is split
// into interleaved runs of 1024 pixels (b2pt_partition: run k -> device k mod N, every pixel computed wholly by one
// device with Philox keyed by (pixel, sample), so the frame is bit-identical for every N), and device 0 then GATHERS
// the runs it does not own straight out of the other devices' frame buffers over NVLink (peer loads inside one kernel;
// a staged cudaMemcpyPeer when peer access is not available).  One exchange per frame, no collective library.
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "ctx.cuh"

#define B2PT_MULTI_MAX 16
#define B2PT_MULTI_TILE 32   // runs of 32*32 pixels

struct b2pt_multi {
    std::vector<b2pt_ctx*> ctx;
    std::vector<int> devices;
    std::vector<char> direct;      // device 0 can load from device i's memory
    std::string err;
    b2pt_stats stats{};
};

namespace {

thread_local std::string g_multi_create_error;

struct PeerFrames { const float* p[B2PT_MULTI_MAX]; };

// Device 0: every pixel of a run owned by device k > 0 is copied from that device's frame (others hold 0 there).
__global__ void __launch_bounds__(256) k_gather_runs(float* __restrict__ dst, PeerFrames src, int world, long long npix, int run) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const int owner = (int)((i / run) % world);
    if (owner == 0) return;
    const float* s = src.p[owner];
    dst[3 * i] = s[3 * i]; dst[3 * i + 1] = s[3 * i + 1]; dst[3 * i + 2] = s[3 * i + 2];
}

// Runs fn(i) for every device on its own host thread; returns the first non-zero status.
template <class F>
int for_each_device(b2pt_multi* m, F fn) {
    const int n = (int)m->ctx.size();
    std::vector<int> rc(n, 0);
    std::vector<std::thread> th;
    for (int i = 1; i < n; ++i) th.emplace_back([&, i]() { rc[i] = fn(i); });
    rc[0] = fn(0);
    for (auto& t : th) t.join();
    for (int i = 0; i < n; ++i)
        if (rc[i]) { m->err = std::string("device ") + std::to_string(m->devices[i]) + ": " + b2pt_last_error(m->ctx[i]); return rc[i]; }
    return B2PT_OK;
}

}  // namespace

extern "C" {

int b2pt_multi_create(const int32_t* devices, int32_t ndev, int32_t flags, int64_t max_paths_in_flight, b2pt_multi** out) {
    if (!out) return B2PT_ERR_INVALID;
    *out = nullptr;
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible == 0) {
        g_multi_create_error = "b2pt_multi_create: no CUDA device available; this engine has no CPU fallback";
        return B2PT_ERR_NO_DEVICE;
    }
    std::vector<int> devs;
    if (ndev <= 0 || !devices) { for (int i = 0; i < visible; ++i) devs.push_back(i); }
    else devs.assign(devices, devices + ndev);
    if ((int)devs.size() > B2PT_MULTI_MAX) { g_multi_create_error = "b2pt_multi_create: at most 16 devices"; return B2PT_ERR_INVALID; }
    b2pt_multi* m = new b2pt_multi();
    m->devices = devs;
    for (int d : devs) {
        b2pt_config cfg{};
        cfg.device = d; cfg.flags = flags; cfg.max_paths_in_flight = max_paths_in_flight;
        b2pt_ctx* c = nullptr;
        int rc = b2pt_create(&cfg, &c);
        if (rc) {
            g_multi_create_error = b2pt_last_error(nullptr);
            for (b2pt_ctx* x : m->ctx) b2pt_destroy(x);
            delete m;
            return rc;
        }
        m->ctx.push_back(c);
    }
    // peer access from the gathering device to every other one
    m->direct.assign(devs.size(), 0);
    cudaSetDevice(devs[0]);
    for (size_t i = 1; i < devs.size(); ++i) {
        if (devs[i] == devs[0]) { m->direct[i] = 1; continue; }
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, devs[0], devs[i]) == cudaSuccess && can) {
            cudaError_t e = cudaDeviceEnablePeerAccess(devs[i], 0);
            if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) m->direct[i] = 1;
            cudaGetLastError();   // clear "already enabled"
        }
    }
    *out = m;
    return B2PT_OK;
}

void b2pt_multi_destroy(b2pt_multi* m) {
    if (!m) return;
    for (b2pt_ctx* c : m->ctx) b2pt_destroy(c);
    delete m;
}

const char* b2pt_multi_last_error(const b2pt_multi* m) { return m ? m->err.c_str() : g_multi_create_error.c_str(); }
int32_t b2pt_multi_device_count(const b2pt_multi* m) { return m ? (int32_t)m->ctx.size() : 0; }
b2pt_ctx* b2pt_multi_ctx(const b2pt_multi* m, int32_t i) { return (m && i >= 0 && i < (int32_t)m->ctx.size()) ? m->ctx[i] : nullptr; }

int b2pt_multi_upload_scene(b2pt_multi* m, const float* pos, const float* nrm, const int32_t* mat, int64_t ntri,
                            const b2pt_material* mats, int32_t nmat, const b2pt_light* lights, int32_t nlight) {
    if (!m) return B2PT_ERR_INVALID;
    return for_each_device(m, [&](int i) { return b2pt_upload_scene(m->ctx[i], pos, nrm, mat, ntri, mats, nmat, lights, nlight); });
}

int b2pt_multi_render(b2pt_multi* m, const b2pt_camera* cam, const b2pt_settings* settings, uint64_t seed, float* rgb) {
    if (!m) return B2PT_ERR_INVALID;
    if (!cam || !settings) { m->err = "b2pt_multi_render: NULL argument"; return B2PT_ERR_INVALID; }
    if (settings->width < 2 || settings->height < 2) { m->err = "b2pt_multi_render: width/height must be >= 2"; return B2PT_ERR_INVALID; }
    const int n = (int)m->ctx.size();
    const long long npix = (long long)settings->width * settings->height;
    const size_t bytes = sizeof(float) * 3ull * (size_t)npix;
    std::vector<float*> frame(n, nullptr);
    int rc = for_each_device(m, [&](int i) {
        b2pt_ctx* c = m->ctx[i];
        cudaSetDevice(c->device);
        void* p = nullptr;
        int r = b2pt::scratch_reserve(c, 7, bytes, &p);
        if (r) return r;
        frame[i] = (float*)p;
        b2pt_partition part{};
        part.tile_rank = i; part.tile_world = n; part.tile_size = B2PT_MULTI_TILE;
        return b2pt_render_device(c, cam, settings, seed, n > 1 ? &part : nullptr, frame[i]);
    });
    if (rc) return rc;
    b2pt_ctx* c0 = m->ctx[0];
    cudaSetDevice(c0->device);
    if (n > 1) {
        PeerFrames src{};
        for (int i = 1; i < n; ++i) {
            if (m->direct[i]) { src.p[i] = frame[i]; continue; }
            // no peer access: stage the frame on device 0 (slots 40.. of its context are free for this)
            void* st = nullptr;
            if ((rc = b2pt::scratch_reserve(c0, 48 + i - 1, bytes, &st))) { m->err = b2pt_last_error(c0); return rc; }
            cudaError_t e = cudaMemcpyPeerAsync(st, c0->device, frame[i], m->ctx[i]->device, bytes, c0->stream);
            if (e != cudaSuccess) { m->err = std::string("cudaMemcpyPeerAsync failed: ") + cudaGetErrorString(e); return B2PT_ERR_CUDA; }
            src.p[i] = (const float*)st;
        }
        k_gather_runs<<<(unsigned)((npix + 255) / 256), 256, 0, c0->stream>>>(frame[0], src, n, npix, B2PT_MULTI_TILE * B2PT_MULTI_TILE);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { m->err = std::string("k_gather_runs launch failed: ") + cudaGetErrorString(e); return B2PT_ERR_CUDA; }
    }
    c0->last_width = settings->width; c0->last_height = settings->height;
    if (rgb) {
        cudaError_t e = cudaMemcpyAsync(rgb, frame[0], bytes, cudaMemcpyDeviceToHost, c0->stream);
        if (e != cudaSuccess) { m->err = std::string("cudaMemcpyAsync(frame) failed: ") + cudaGetErrorString(e); return B2PT_ERR_CUDA; }
    }
    cudaError_t e = cudaStreamSynchronize(c0->stream);
    if (e != cudaSuccess) { m->err = std::string("gather failed: ") + cudaGetErrorString(e); return B2PT_ERR_CUDA; }
    // stats: counts add up, times are the slowest device's
    b2pt_stats s{};
    for (int i = 0; i < n; ++i) {
        const b2pt_stats& t = m->ctx[i]->stats;
        s.extend_rays += t.extend_rays; s.shadow_rays += t.shadow_rays; s.samples += t.samples; s.fallback_rays += t.fallback_rays;
        s.node_fetches += t.node_fetches; s.tri_fetches += t.tri_fetches; s.kernel_launches += t.kernel_launches;
        s.extend_launches += t.extend_launches; s.shadow_launches += t.shadow_launches;
        s.gpu_seconds = std::max(s.gpu_seconds, t.gpu_seconds); s.trace_seconds = std::max(s.trace_seconds, t.trace_seconds);
        s.build_seconds = std::max(s.build_seconds, t.build_seconds); s.extend_seconds = std::max(s.extend_seconds, t.extend_seconds);
        s.shadow_seconds = std::max(s.shadow_seconds, t.shadow_seconds); s.order_seconds = std::max(s.order_seconds, t.order_seconds);
    }
    if (n > 1) s.kernel_launches += 1;
    m->stats = s;
    return B2PT_OK;
}

int b2pt_multi_tonemap_last(b2pt_multi* m, float gamma, int32_t flip, uint8_t* rgb8) {
    if (!m) return B2PT_ERR_INVALID;
    int rc = b2pt_tonemap_last(m->ctx[0], gamma, flip, rgb8);
    if (rc) m->err = b2pt_last_error(m->ctx[0]);
    return rc;
}

int b2pt_multi_get_stats(const b2pt_multi* m, b2pt_stats* out) {
    if (!m || !out) return B2PT_ERR_INVALID;
    *out = m->stats;
    return B2PT_OK;
}

}  // extern "C"
