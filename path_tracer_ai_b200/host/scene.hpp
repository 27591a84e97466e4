// scene.hpp — host mirror of the reference's Scene / Light (include/scene.hpp:21-115) and of
// Scene::loadFromObj (src/scene.cpp:8-293).
//
// Same public surface: Scene() installs the four hard-coded lights, loadFromObj() builds the
// triangle list (room + normalised model) and leaves it in the reference's POST-BVH-build order,
// getMaterials()/getLights()/getTriangles() feed the renderer.  What is gone is the CPU
// intersector: Scene::intersect (scene.hpp:96-99) is served by the GPU engine through
// B200Renderer / b2pt_trace_closest, and there is no CPU fallback.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <sys/stat.h>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b2pt.h"
#include "material.hpp"
#include "obj_parser.hpp"
#include "triangle.hpp"
#include "vec.hpp"

namespace b2pt {

struct Light {
    vec3 position;
    vec3 color;
    float intensity;
    Light(const vec3& pos, const vec3& col, float intens) : position(pos), color(col), intensity(intens) {
        if (intensity <= 0.0f) {   // scene.hpp:31-35
            std::fprintf(stderr, "Warning: Invalid light intensity %g, setting to 1.0\n", intensity);
            intensity = 1.0f;
        }
    }
};

// MTL material -> Material.  First the documented name-prefix extension (not in the reference;
// needed to express diffuse and dielectric materials — the reference forces every MTL material to
// SPECULAR and never creates a DIELECTRIC): diffuse* / glass* / mirror* / rough<value>*.  Anything
// else follows the reference rule, src/scene.cpp:74-108.
inline std::shared_ptr<Material> materialFromMtl(const obj::MtlMaterial& m) {
    auto mat = std::make_shared<Material>();
    const vec3 kd(m.diffuse[0], m.diffuse[1], m.diffuse[2]);
    auto has_prefix = [&](const char* p) { return m.name.compare(0, std::strlen(p), p) == 0; };
    auto contains = [&](const char* p) { return m.name.find(p) != std::string::npos; };
    if (has_prefix("diffuse")) {
        mat->type = MaterialType::DIFFUSE; mat->albedo = kd; mat->roughness = 0.95f; mat->metallic = 0.0f;
    } else if (has_prefix("glass")) {
        mat->type = MaterialType::DIELECTRIC; mat->albedo = kd; mat->roughness = 0.0f; mat->metallic = 0.0f;
        mat->ior = m.ior > 1.0f ? m.ior : 1.5f;
    } else if (has_prefix("mirror")) {
        mat->type = MaterialType::SPECULAR; mat->albedo = kd; mat->roughness = 0.0f; mat->metallic = 1.0f;
    } else {
        char* end = nullptr;
        float r = has_prefix("rough") ? std::strtof(m.name.c_str() + 5, &end) : 0.0f;
        if (has_prefix("rough") && end != m.name.c_str() + 5) {
            mat->type = MaterialType::SPECULAR; mat->albedo = kd; mat->roughness = r; mat->metallic = 1.0f;
        } else {
            mat->type = MaterialType::SPECULAR;   // scene.cpp:78-80
            mat->metallic = 1.0f;
            mat->roughness = 0.1f;
            if (contains("red")) { mat->albedo = vec3(0.9f, 0.2f, 0.2f); }
            else if (contains("gold")) { mat->albedo = vec3(1.0f, 0.8f, 0.0f); mat->roughness = 0.05f; }
            else if (contains("silver")) { mat->albedo = vec3(0.95f); mat->roughness = 0.05f; }
            else if (contains("black")) { mat->albedo = vec3(0.02f); }
            else {
                // scene.cpp:103-105: pow(Kd, 0.8) * 1.2 clamped to [0, 1]
                vec3 a(std::pow(kd.x, 0.8f), std::pow(kd.y, 0.8f), std::pow(kd.z, 0.8f));
                a = a * 1.2f;
                mat->albedo = vec3(clamp1(a.x, 0.0f, 1.0f), clamp1(a.y, 0.0f, 1.0f), clamp1(a.z, 0.0f, 1.0f));
            }
        }
    }
    return mat;
}

class Scene {
private:
    std::vector<Triangle> triangles;
    std::vector<std::shared_ptr<Material>> materials;
    std::vector<Light> lights;
    std::vector<int> buildOrder;   // post-build position -> index in the loader's own order

public:
    Scene() {
        // scene.hpp:55-80
        lights.emplace_back(vec3(2.0f, 3.5f, 2.0f), vec3(1.0f, 0.95f, 0.8f), 9.0f);
        lights.emplace_back(vec3(-1.5f, 2.0f, 1.5f), vec3(0.8f, 0.9f, 1.0f), 2.0f);
        lights.emplace_back(vec3(0.0f, 2.0f, -2.0f), vec3(1.0f), 1.0f);
        lights.emplace_back(vec3(0.0f, 0.1f, 0.0f), vec3(0.9f, 0.9f, 1.0f), 2.0f);
    }

    // Not in the reference (its four lights are constants, scene.hpp:55-80): replaces them (command line --lights).
    void setLights(std::vector<Light> ls) { lights = std::move(ls); }

    const std::vector<std::shared_ptr<Material>>& getMaterials() const { return materials; }
    const std::vector<Light>& getLights() const { return lights; }
    const std::vector<Triangle>& getTriangles() const { return triangles; }
    const std::vector<int>& getBuildOrder() const { return buildOrder; }

    // Puts `triangles` into the order the reference's BVH::build leaves them in (bvh.hpp:27-72,
    // called at scene.cpp:290) — that order is the reference tree the GPU engine is exact against.
    void applyReferenceOrder() {
        const size_t n = triangles.size();
        std::vector<float> pos(9 * n);
        for (size_t i = 0; i < n; ++i) {
            const Triangle& t = triangles[i];
            const vec3 v[3] = {t.v0, t.v1, t.v2};
            for (int k = 0; k < 3; ++k) { pos[9 * i + 3 * k] = v[k].x; pos[9 * i + 3 * k + 1] = v[k].y; pos[9 * i + 3 * k + 2] = v[k].z; }
        }
        buildOrder.assign(n, 0);
        b2pt_reference_order(pos.data(), static_cast<int64_t>(n), buildOrder.data());
        std::vector<Triangle> sorted(n);
        for (size_t p = 0; p < n; ++p) sorted[p] = triangles[buildOrder[p]];
        triangles.swap(sorted);
    }

    // Replaces the contents (used by tests / generators that bypass OBJ files).
    void setContents(std::vector<Triangle> tris, std::vector<std::shared_ptr<Material>> mats) {
        triangles = std::move(tris);
        materials = std::move(mats);
        applyReferenceOrder();
    }

    // ---- binary scene cache ------------------------------------------------------------------------------------
    // Parsing a 10M-triangle OBJ and putting it into the reference's BVH::build order (std::nth_element per node,
    // bvh.hpp:63-66) takes seconds on the host; the result depends only on the files.  saveCache writes the finished
    // scene (triangles in post-build order, build order, materials, lights); loadFromObjCached reuses it while the
    // OBJ's size and modification time are the ones recorded in it, and rebuilds + rewrites it otherwise.
    bool saveCache(const std::string& cachePath, const std::string& objPath = std::string()) const {
        static_assert(sizeof(Triangle) == 100, "Triangle is written as it is");
        FILE* f = std::fopen(cachePath.c_str(), "wb");
        if (!f) return false;
        struct stat sb{};
        const bool have = !objPath.empty() && ::stat(objPath.c_str(), &sb) == 0;
        const uint64_t hdr[8] = {0x314e435354503242ull /* "B2PTSCN1" */, static_cast<uint64_t>(triangles.size()), static_cast<uint64_t>(materials.size()),
                                 static_cast<uint64_t>(lights.size()), have ? static_cast<uint64_t>(sb.st_size) : 0ull,
                                 have ? static_cast<uint64_t>(sb.st_mtim.tv_sec) : 0ull, have ? static_cast<uint64_t>(sb.st_mtim.tv_nsec) : 0ull, sizeof(Triangle)};
        bool ok = std::fwrite(hdr, sizeof(hdr), 1, f) == 1;
        if (!triangles.empty()) ok = ok && std::fwrite(triangles.data(), sizeof(Triangle), triangles.size(), f) == triangles.size();
        if (!buildOrder.empty()) ok = ok && std::fwrite(buildOrder.data(), sizeof(int), buildOrder.size(), f) == buildOrder.size();
        for (const auto& m : materials) {
            const float rec[8] = {static_cast<float>(static_cast<int>(m->type)), m->albedo.x, m->albedo.y, m->albedo.z, m->roughness, m->metallic, m->ior, 0.0f};
            ok = ok && std::fwrite(rec, sizeof(rec), 1, f) == 1;
        }
        for (const auto& l : lights) {
            const float rec[7] = {l.position.x, l.position.y, l.position.z, l.color.x, l.color.y, l.color.z, l.intensity};
            ok = ok && std::fwrite(rec, sizeof(rec), 1, f) == 1;
        }
        return (std::fclose(f) == 0) && ok;
    }

    // objPath non-empty: the cache is only accepted if it was written for that file as it is now.
    bool loadCache(const std::string& cachePath, const std::string& objPath = std::string()) {
        FILE* f = std::fopen(cachePath.c_str(), "rb");
        if (!f) return false;
        uint64_t hdr[8] = {};
        bool ok = std::fread(hdr, sizeof(hdr), 1, f) == 1 && hdr[0] == 0x314e435354503242ull && hdr[7] == sizeof(Triangle) &&
                  hdr[1] < (1ull << 28) && hdr[2] < (1ull << 24) && hdr[3] <= 16;
        if (ok && !objPath.empty()) {
            struct stat sb{};
            ok = ::stat(objPath.c_str(), &sb) == 0 && hdr[4] == static_cast<uint64_t>(sb.st_size) &&
                 hdr[5] == static_cast<uint64_t>(sb.st_mtim.tv_sec) && hdr[6] == static_cast<uint64_t>(sb.st_mtim.tv_nsec);
        }
        std::vector<Triangle> tris;
        std::vector<int> order;
        std::vector<std::shared_ptr<Material>> mats;
        std::vector<Light> ls;
        if (ok) {
            tris.resize(hdr[1]); order.resize(hdr[1]);
            if (hdr[1]) ok = std::fread(tris.data(), sizeof(Triangle), tris.size(), f) == tris.size() && std::fread(order.data(), sizeof(int), order.size(), f) == order.size();
            for (uint64_t i = 0; ok && i < hdr[2]; ++i) {
                float rec[8];
                ok = std::fread(rec, sizeof(rec), 1, f) == 1;
                auto m = std::make_shared<Material>();
                m->type = static_cast<MaterialType>(static_cast<int>(rec[0])); m->albedo = vec3(rec[1], rec[2], rec[3]);
                m->roughness = rec[4]; m->metallic = rec[5]; m->ior = rec[6];
                mats.push_back(m);
            }
            for (uint64_t i = 0; ok && i < hdr[3]; ++i) {
                float rec[7];
                ok = std::fread(rec, sizeof(rec), 1, f) == 1;
                ls.emplace_back(vec3(rec[0], rec[1], rec[2]), vec3(rec[3], rec[4], rec[5]), rec[6]);
            }
            ok = ok && std::fgetc(f) == EOF;
        }
        std::fclose(f);
        if (!ok) return false;
        triangles.swap(tris); buildOrder.swap(order); materials.swap(mats); lights.swap(ls);
        return true;
    }

    // loadFromObj through the cache `cachePath` (default: <obj>.b2ptscene next to the OBJ).  Same result either way.
    bool loadFromObjCached(const std::string& objPath, std::string cachePath = std::string()) {
        if (cachePath.empty()) cachePath = objPath + ".b2ptscene";
        const std::vector<Light> keep = lights;
        if (loadCache(cachePath, objPath)) { lights = keep; return true; }
        if (!loadFromObj(objPath)) return false;
        saveCache(cachePath, objPath);   // best effort: a read-only directory just means no cache
        return true;
    }

    bool loadFromObj(const std::string& objPath) {
        obj::Mesh mesh;
        if (!obj::parse_file(objPath, mesh)) {
            if (!mesh.error.empty()) std::fprintf(stderr, "OBJ reader error: %s", mesh.error.c_str());
            return false;
        }
        if (!mesh.warning.empty()) std::fprintf(stdout, "OBJ reader warning: %s", mesh.warning.c_str());

        // scene.cpp:30-52
        vec3 minB(3.402823466e+38f), maxB(-3.402823466e+38f);
        for (size_t i = 0; i + 2 < mesh.vertices.size(); i += 3) {
            vec3 v(mesh.vertices[i], mesh.vertices[i + 1], mesh.vertices[i + 2]);
            minB = vmin(minB, v);
            maxB = vmax(maxB, v);
        }
        vec3 size = maxB - minB;
        float scale = 3.f / max2(max2(size.x, size.y), size.z);
        vec3 centre = (minB + maxB) * 0.5f;

        // scene.cpp:57-114
        materials.clear();
        auto m0 = std::make_shared<Material>();
        m0->type = MaterialType::SPECULAR; m0->albedo = vec3(0.9f, 0.2f, 0.2f); m0->roughness = 0.1f; m0->metallic = 1.0f;
        materials.push_back(m0);
        auto m1 = std::make_shared<Material>();
        m1->type = MaterialType::DIFFUSE; m1->albedo = vec3(0.9f, 0.9f, 0.9f); m1->roughness = 0.95f; m1->metallic = 0.0f;
        materials.push_back(m1);
        for (const auto& m : mesh.materials) materials.push_back(materialFromMtl(m));

        // scene.cpp:118-209 — the room: floor, back, left, right walls; material 1.
        triangles.clear();
        const float R = 8.0f, Hh = 4.0f;
        auto two = [&](const vec3& a, const vec3& b, const vec3& c, const vec2& ta, const vec2& tb, const vec2& tc, const vec3& n) {
            triangles.emplace_back(a, b, c, n, n, n, ta, tb, tc, 1);
        };
        const vec3 up(0, 1, 0), front(0, 0, 1), px(1, 0, 0), nx(-1, 0, 0);
        two(vec3(-R, 0, -R), vec3(R, 0, -R), vec3(R, 0, R), vec2(0.0f), vec2(1, 0), vec2(1.0f), up);
        two(vec3(-R, 0, -R), vec3(R, 0, R), vec3(-R, 0, R), vec2(0.0f), vec2(1.0f), vec2(0, 1), up);
        two(vec3(-R, 0, -R), vec3(-R, Hh, -R), vec3(R, Hh, -R), vec2(0.0f), vec2(0, 1), vec2(1, 1), front);
        two(vec3(-R, 0, -R), vec3(R, Hh, -R), vec3(R, 0, -R), vec2(0.0f), vec2(1, 1), vec2(1, 0), front);
        two(vec3(-R, 0, -R), vec3(-R, 0, R), vec3(-R, Hh, R), vec2(0.0f), vec2(1, 0), vec2(1, 1), px);
        two(vec3(-R, 0, -R), vec3(-R, Hh, R), vec3(-R, Hh, -R), vec2(0.0f), vec2(1, 1), vec2(0, 1), px);
        two(vec3(R, 0, -R), vec3(R, Hh, R), vec3(R, 0, R), vec2(0.0f), vec2(1, 1), vec2(1, 0), nx);
        two(vec3(R, 0, -R), vec3(R, Hh, -R), vec3(R, Hh, R), vec2(0.0f), vec2(0, 1), vec2(1, 1), nx);

        // scene.cpp:215-282
        const size_t nfaces = mesh.material_ids.size();
        triangles.reserve(triangles.size() + nfaces);
        for (size_t f = 0; f < nfaces; ++f) {
            vec3 vs[3], ns[3];
            vec2 ts[3];
            for (int k = 0; k < 3; ++k) {
                const obj::Index idx = mesh.indices[3 * f + k];
                vec3 p(mesh.vertices[3 * idx.vertex_index], mesh.vertices[3 * idx.vertex_index + 1], mesh.vertices[3 * idx.vertex_index + 2]);
                p = (p - centre) * scale;
                p.z = -p.z;
                p.y += 1.8f;
                vs[k] = p;
                if (idx.normal_index >= 0) {
                    vec3 n(mesh.normals[3 * idx.normal_index], mesh.normals[3 * idx.normal_index + 1], mesh.normals[3 * idx.normal_index + 2]);
                    n.z = -n.z;
                    ns[k] = normalize(n);
                } else if (k == 2) {
                    vec3 n = normalize(cross(vs[1] - vs[0], vs[2] - vs[0]));
                    ns[0] = ns[1] = ns[2] = n;
                }
                ts[k] = idx.texcoord_index >= 0 ? vec2(mesh.texcoords[2 * idx.texcoord_index], mesh.texcoords[2 * idx.texcoord_index + 1]) : vec2(0.0f);
            }
            int materialId = mesh.material_ids[f];
            if (materialId < 0) materialId = 0;
            materialId += 2;
            triangles.emplace_back(vs[0], vs[1], vs[2], ns[0], ns[1], ns[2], ts[0], ts[1], ts[2], materialId);
        }
        applyReferenceOrder();   // scene.cpp:290
        return true;
    }
};

}  // namespace b2pt
