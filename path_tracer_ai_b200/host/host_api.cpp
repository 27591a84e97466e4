// host_api.cpp — C entry points onto the host-side scene loader (declared in include/b2pt_host.h),
// compiled into libb2pt.so so Python tests and bench.py reach the same Scene::loadFromObj the
// command line uses.
#include <cstring>
#include <string>

#include "../../include/b2pt_host.h"
#include "camera.hpp"
#include "png_writer.hpp"
#include "scene.hpp"

struct b2pt_scene {
    b2pt::Scene scene;
};

extern "C" {

int b2pt_scene_load_obj(const char* path, b2pt_scene** out) {
    if (!path || !out) return B2PT_ERR_INVALID;
    b2pt_scene* s = new b2pt_scene();
    if (!s->scene.loadFromObj(path)) { delete s; *out = nullptr; return B2PT_ERR_INVALID; }
    *out = s;
    return B2PT_OK;
}

int b2pt_scene_load_obj_cached(const char* path, const char* cache_path, b2pt_scene** out) {
    if (!path || !out) return B2PT_ERR_INVALID;
    b2pt_scene* s = new b2pt_scene();
    if (!s->scene.loadFromObjCached(path, cache_path ? cache_path : "")) { delete s; *out = nullptr; return B2PT_ERR_INVALID; }
    *out = s;
    return B2PT_OK;
}

void b2pt_scene_free(b2pt_scene* s) { delete s; }

int b2pt_obj_parser_selfcheck(const char* path, int32_t nthreads, int64_t chunk_bytes) {
    if (!path) return -1;
    b2pt::obj::Mesh a, b;
    bool oka = b2pt::obj::parse_file(path, a, nthreads, chunk_bytes > 0 ? (size_t)chunk_bytes : (size_t)1);
    bool okb = b2pt::obj::parse_file_serial(path, b);
    if (oka != okb) return 1;
    if (!oka) return a.error == b.error ? -1 : 2;
    auto same_idx = [](const b2pt::obj::Index& x, const b2pt::obj::Index& y) {
        return x.vertex_index == y.vertex_index && x.normal_index == y.normal_index && x.texcoord_index == y.texcoord_index;
    };
    auto same_bits = [](const std::vector<float>& x, const std::vector<float>& y) {
        return x.size() == y.size() && (x.empty() || !std::memcmp(x.data(), y.data(), x.size() * sizeof(float)));
    };
    if (!same_bits(a.vertices, b.vertices)) return 3;
    if (!same_bits(a.normals, b.normals)) return 4;
    if (!same_bits(a.texcoords, b.texcoords)) return 5;
    if (a.indices.size() != b.indices.size()) return 6;
    for (size_t i = 0; i < a.indices.size(); ++i) if (!same_idx(a.indices[i], b.indices[i])) return 7;
    if (a.material_ids != b.material_ids) return 8;
    if (a.materials.size() != b.materials.size()) return 9;
    for (size_t i = 0; i < a.materials.size(); ++i)
        if (a.materials[i].name != b.materials[i].name || std::memcmp(a.materials[i].diffuse, b.materials[i].diffuse, 12) || a.materials[i].ior != b.materials[i].ior) return 10;
    return 0;
}

int64_t b2pt_scene_num_triangles(const b2pt_scene* s) { return s ? (int64_t)s->scene.getTriangles().size() : 0; }
int32_t b2pt_scene_num_materials(const b2pt_scene* s) { return s ? (int32_t)s->scene.getMaterials().size() : 0; }
int32_t b2pt_scene_num_lights(const b2pt_scene* s) { return s ? (int32_t)s->scene.getLights().size() : 0; }

int b2pt_scene_get_triangles(const b2pt_scene* s, float* pos, float* nrm, int32_t* mat, int32_t* order) {
    if (!s) return B2PT_ERR_INVALID;
    const auto& tris = s->scene.getTriangles();
    for (size_t i = 0; i < tris.size(); ++i) {
        const b2pt::Triangle& t = tris[i];
        const b2pt::vec3 v[3] = {t.v0, t.v1, t.v2}, n[3] = {t.n0, t.n1, t.n2};
        for (int k = 0; k < 3; ++k) {
            if (pos) { pos[9 * i + 3 * k] = v[k].x; pos[9 * i + 3 * k + 1] = v[k].y; pos[9 * i + 3 * k + 2] = v[k].z; }
            if (nrm) { nrm[9 * i + 3 * k] = n[k].x; nrm[9 * i + 3 * k + 1] = n[k].y; nrm[9 * i + 3 * k + 2] = n[k].z; }
        }
        if (mat) mat[i] = t.materialId;
    }
    if (order) std::memcpy(order, s->scene.getBuildOrder().data(), sizeof(int32_t) * tris.size());
    return B2PT_OK;
}

int b2pt_scene_get_materials(const b2pt_scene* s, b2pt_material* mats) {
    if (!s || !mats) return B2PT_ERR_INVALID;
    const auto& ms = s->scene.getMaterials();
    for (size_t i = 0; i < ms.size(); ++i) {
        mats[i].type = static_cast<int32_t>(ms[i]->type);
        mats[i].albedo[0] = ms[i]->albedo.x; mats[i].albedo[1] = ms[i]->albedo.y; mats[i].albedo[2] = ms[i]->albedo.z;
        mats[i].roughness = ms[i]->roughness; mats[i].metallic = ms[i]->metallic; mats[i].ior = ms[i]->ior; mats[i]._pad = 0.0f;
    }
    return B2PT_OK;
}

int b2pt_scene_get_lights(const b2pt_scene* s, b2pt_light* lights) {
    if (!s || !lights) return B2PT_ERR_INVALID;
    const auto& ls = s->scene.getLights();
    for (size_t i = 0; i < ls.size(); ++i) {
        lights[i].position[0] = ls[i].position.x; lights[i].position[1] = ls[i].position.y; lights[i].position[2] = ls[i].position.z;
        lights[i].color[0] = ls[i].color.x; lights[i].color[1] = ls[i].color.y; lights[i].color[2] = ls[i].color.z;
        lights[i].intensity = ls[i].intensity;
    }
    return B2PT_OK;
}

int b2pt_camera_look_at(const float* position, const float* target, const float* up, float fov, b2pt_camera* out) {
    if (!position || !target || !up || !out) return B2PT_ERR_INVALID;
    b2pt::Camera c(b2pt::vec3(position[0], position[1], position[2]), b2pt::vec3(target[0], target[1], target[2]),
                   b2pt::vec3(up[0], up[1], up[2]), fov);
    *out = c.toC();
    return B2PT_OK;
}

int b2pt_write_png(const char* path, int32_t width, int32_t height, const uint8_t* rgb8) {
    if (!path || !rgb8 || width <= 0 || height <= 0) return B2PT_ERR_INVALID;
    return b2pt::writePngRGB8(path, width, height, rgb8) ? B2PT_OK : B2PT_ERR_INVALID;
}

int b2pt_write_pfm(const char* path, int32_t width, int32_t height, const float* rgb) {
    if (!path || !rgb || width <= 0 || height <= 0) return B2PT_ERR_INVALID;
    return b2pt::writePfmRGB(path, width, height, rgb) ? B2PT_OK : B2PT_ERR_INVALID;
}

}  // extern "C"
