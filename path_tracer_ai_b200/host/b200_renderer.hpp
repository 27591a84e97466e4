// b200_renderer.hpp — drop-in for the reference's OptixRenderer (include/gpu/optix_renderer.hpp:11-42):
// same nested Settings (same defaults), same lifecycle — ctor, initialize(), uploadScene(const Scene&),
// render(const Camera&), saveImage(const std::string&) — and the same error convention (every failure
// throws std::runtime_error with the failing call's text).  Thin C++ over the C ABI in include/b2pt.h.
//
// Beyond the reference's surface: a list of devices (one object, one process, N GPUs: b2pt_multi_*),
// renderProgressive (resumable accumulation), saveImage's `flip` (the reference writes its PNG upside down; flip
// writes it upright) and a .pfm float dump when the file name ends in ".pfm".
#pragma once
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/b2pt.h"
#include "camera.hpp"
#include "png_writer.hpp"
#include "scene.hpp"

namespace b2pt {

class B200Renderer {
public:
    struct Settings {
        int width;
        int height;
        int samplesPerPixel;
        int maxBounces;
        float gamma;
        Settings() : width(800), height(450), samplesPerPixel(10), maxBounces(3), gamma(2.2f) {}
    };

    explicit B200Renderer(const Settings& settings = Settings(), int device = 0, uint64_t seed = 1234)
        : settings(settings), devices(1, device), seed(seed) {}
    // several GPUs of one box behind the same object (scene replicated, interleaved pixel runs, gather on devices[0])
    B200Renderer(const Settings& settings, const std::vector<int>& devices, uint64_t seed = 1234)
        : settings(settings), devices(devices), seed(seed) {}
    ~B200Renderer() {
        if (multi) b2pt_multi_destroy(multi);
        else if (ctx) b2pt_destroy(ctx);
    }
    B200Renderer(const B200Renderer&) = delete;
    B200Renderer& operator=(const B200Renderer&) = delete;

    void initialize() {
        if (devices.size() > 1) {
            std::vector<int32_t> d(devices.begin(), devices.end());
            if (b2pt_multi_create(d.data(), static_cast<int32_t>(d.size()), 0, 0, &multi) != B2PT_OK)
                throw std::runtime_error(b2pt_multi_last_error(nullptr));
            ctx = b2pt_multi_ctx(multi, 0);
        } else {
            b2pt_config cfg{};
            cfg.device = devices.empty() ? 0 : devices[0];
            if (b2pt_create(&cfg, &ctx) != B2PT_OK) throw std::runtime_error(b2pt_last_error(nullptr));
        }
    }

    void uploadScene(const Scene& scene) {
        requireInit("uploadScene");
        const auto& tris = scene.getTriangles();   // post-build order (optix_renderer.cu:385)
        const size_t n = tris.size();
        std::vector<float> pos(9 * n), nrm(9 * n);
        std::vector<int32_t> mat(n);
        for (size_t i = 0; i < n; ++i) {
            const Triangle& t = tris[i];
            const vec3 v[3] = {t.v0, t.v1, t.v2}, nn[3] = {t.n0, t.n1, t.n2};
            for (int k = 0; k < 3; ++k) {
                pos[9 * i + 3 * k] = v[k].x; pos[9 * i + 3 * k + 1] = v[k].y; pos[9 * i + 3 * k + 2] = v[k].z;
                nrm[9 * i + 3 * k] = nn[k].x; nrm[9 * i + 3 * k + 1] = nn[k].y; nrm[9 * i + 3 * k + 2] = nn[k].z;
            }
            mat[i] = t.materialId;
        }
        std::vector<b2pt_material> mats;
        for (const auto& m : scene.getMaterials()) {
            b2pt_material c{};
            c.type = static_cast<int32_t>(m->type);
            c.albedo[0] = m->albedo.x; c.albedo[1] = m->albedo.y; c.albedo[2] = m->albedo.z;
            c.roughness = m->roughness; c.metallic = m->metallic; c.ior = m->ior;
            mats.push_back(c);
        }
        std::vector<b2pt_light> ls;
        for (const auto& l : scene.getLights()) {
            b2pt_light c{};
            c.position[0] = l.position.x; c.position[1] = l.position.y; c.position[2] = l.position.z;
            c.color[0] = l.color.x; c.color[1] = l.color.y; c.color[2] = l.color.z;
            c.intensity = l.intensity;
            ls.push_back(c);
        }
        if (multi)
            checkMulti(b2pt_multi_upload_scene(multi, pos.data(), nrm.data(), mat.data(), static_cast<int64_t>(n), mats.data(),
                                               static_cast<int32_t>(mats.size()), ls.data(), static_cast<int32_t>(ls.size())));
        else
            check(b2pt_upload_scene(ctx, pos.data(), nrm.data(), mat.data(), static_cast<int64_t>(n), mats.data(),
                                    static_cast<int32_t>(mats.size()), ls.data(), static_cast<int32_t>(ls.size())));
    }

    void render(const Camera& camera) {
        requireInit("render");   // optix_renderer.cu:421-423
        b2pt_camera cam = camera.toC();
        b2pt_settings st = cSettings();
        frame.assign(static_cast<size_t>(settings.width) * settings.height * 3, 0.0f);
        if (multi) checkMulti(b2pt_multi_render(multi, &cam, &st, seed, frame.data()));
        else check(b2pt_render(ctx, &cam, &st, seed, nullptr, frame.data()));
    }

    // Progressive, resumable accumulation: passes of `samplesPerPass` samples per pixel until settings.samplesPerPixel
    // are done; `onPass(samplesDone, frame)` sees the running mean after every pass (return false to stop early).  The
    // last frame is bit-identical to render().  Single device.
    void renderProgressive(const Camera& camera, int samplesPerPass, const std::function<bool(int, const std::vector<float>&)>& onPass = nullptr) {
        requireInit("renderProgressive");
        if (multi) throw std::runtime_error("renderProgressive: one device only");
        b2pt_camera cam = camera.toC();
        b2pt_settings st = cSettings();
        frame.assign(static_cast<size_t>(settings.width) * settings.height * 3, 0.0f);
        check(b2pt_progressive_begin(ctx, &cam, &st, seed));
        int32_t done = 0;
        while (done < settings.samplesPerPixel) {
            check(b2pt_progressive_pass(ctx, samplesPerPass, frame.data(), &done));
            if (onPass && !onPass(done, frame)) break;
        }
    }

    // Renderer::saveImage (src/renderer.cpp:5-21): clamp, pow(1/gamma), truncate to 8 bit — on the GPU, byte-exact
    // (b2pt_tonemap_last) — PNG with rows in framebuffer order (row 0 = bottom of the view: the reference writes it
    // that way too) unless `flip`.  A name ending in ".pfm" writes the linear float frame instead.
    void saveImage(const std::string& filename, bool flip = false) {
        if (frame.empty()) throw std::runtime_error("saveImage: nothing rendered yet");
        if (filename.size() >= 4 && filename.compare(filename.size() - 4, 4, ".pfm") == 0) {
            if (!writePfmRGB(filename, settings.width, settings.height, frame.data())) throw std::runtime_error("saveImage: cannot write " + filename);
            return;
        }
        std::vector<uint8_t> px(frame.size());
        if (multi) checkMulti(b2pt_multi_tonemap_last(multi, settings.gamma, flip ? 1 : 0, px.data()));
        else check(b2pt_tonemap_last(ctx, settings.gamma, flip ? 1 : 0, px.data()));
        if (!writePngRGB8(filename, settings.width, settings.height, px.data()))
            throw std::runtime_error("saveImage: cannot write " + filename);
    }

    const std::vector<float>& frameBuffer() const { return frame; }
    b2pt_stats stats() const {
        b2pt_stats s{};
        if (multi) b2pt_multi_get_stats(multi, &s);
        else if (ctx) b2pt_get_stats(ctx, &s);
        return s;
    }
    b2pt_ctx* handle() const { return ctx; }
    int deviceCount() const { return multi ? b2pt_multi_device_count(multi) : 1; }

private:
    b2pt_settings cSettings() const { return b2pt_settings{settings.width, settings.height, settings.samplesPerPixel, settings.maxBounces, settings.gamma}; }
    void requireInit(const char* who) const {
        if (!ctx) throw std::runtime_error(std::string("B200Renderer::") + who + " called before initialize()");
    }
    void check(int rc) const {
        if (rc != B2PT_OK) throw std::runtime_error(b2pt_last_error(ctx));
    }
    void checkMulti(int rc) const {
        if (rc != B2PT_OK) throw std::runtime_error(b2pt_multi_last_error(multi));
    }

    Settings settings;
    std::vector<int> devices;
    uint64_t seed;
    b2pt_ctx* ctx = nullptr;
    b2pt_multi* multi = nullptr;
    std::vector<float> frame;
};

}  // namespace b2pt
