// oracle/_ref harness — TEST INFRASTRUCTURE ONLY. Never linked into, imported by or executed
// from the product path (path_tracer_ai_b200/); only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load the library this builds.
//
// What this is: the reference's OWN CPU implementation of the hot path, compiled from the
// UNMODIFIED headers where they lie under /root/reference/include (ray.hpp, intersection.hpp,
// aabb.hpp, triangle.hpp, bvh.hpp, material.hpp, camera.hpp, scene.hpp, renderer.hpp) against
// the glm-compatible shim in oracle/glm_shim (GLM is an absent, un-pinned dependency), exported
// behind a small C API so Python tests can drive it.  Output goes to oracle/_ref/ only.
//
// The two members the reference declares in headers but defines in sources that need absent
// third-party code are defined HERE:
//   * Scene::loadFromObj  (reference src/scene.cpp:8-293 needs tinyobjloader)  -> restated below
//     on top of the repo's OBJ/MTL parser, plus `__arrays__` hooks used to inject geometry.
//   * Renderer::saveImage (reference src/renderer.cpp:5-21 needs stb)          -> dumps the float
//     framebuffer to the address encoded in the "file name" (the tonemap lives in the product host).
// As members they can reach Scene's / Renderer's private state without editing any reference file.
//
// Build: oracle/Makefile (g++ -std=c++17 -O2 -fopenmp, no -march=native / -mfma / -ffast-math).
#include <iostream>
#include <cstdint>
#include <cstring>
#include <chrono>
#include <string>
#include <vector>
#include <memory>

#include "scene.hpp"
#include "camera.hpp"
#include "renderer.hpp"

#include "../path_tracer_ai_b200/host/obj_parser.hpp"

namespace {

// Geometry handed to Scene::loadFromObj through the "__arrays__" pseudo-path.
struct InjectedScene {
    const float* pos = nullptr;   // ntri*9
    const float* nrm = nullptr;   // ntri*9 (may be null -> zero normals)
    const int* mat = nullptr;     // ntri (may be null -> 0)
    int ntri = 0;
    const float* mats = nullptr;  // nmat * 8: type, r, g, b, roughness, metallic, ior, pad
    int nmat = 0;
};
thread_local const InjectedScene* g_injected = nullptr;
// When set, loadFromObj stores a copy of the triangle list as it is just before BVH::build.
thread_local std::vector<Triangle>* g_prebuild_sink = nullptr;

struct CoutMute {
    std::ios_base::iostate saved;
    CoutMute() : saved(std::cout.rdstate()) { std::cout.setstate(std::ios_base::failbit); }
    ~CoutMute() { std::cout.clear(saved); }
};

bool starts_with(const std::string& s, const char* prefix) {
    return s.compare(0, std::strlen(prefix), prefix) == 0;
}

// MTL material -> Material. Reference rule: src/scene.cpp:74-108 (always SPECULAR, name tests in
// the order red/gold/silver|darksilver/black, else pow(Kd,0.8)*1.2 clamped).  The prefix
// extension (NOT in the reference, needed for diffuse/dielectric test scenes)
// is evaluated first: diffuse* / glass* / mirror* / rough<value>*.
std::shared_ptr<Material> material_from_mtl(const b2pt::obj::MtlMaterial& m) {
    auto mat = std::make_shared<Material>();
    const glm::vec3 kd(m.diffuse[0], m.diffuse[1], m.diffuse[2]);
    if (starts_with(m.name, "diffuse")) {
        mat->type = MaterialType::DIFFUSE; mat->albedo = kd; mat->roughness = 0.95f; mat->metallic = 0.0f;
        return mat;
    }
    if (starts_with(m.name, "glass")) {
        mat->type = MaterialType::DIELECTRIC; mat->albedo = kd; mat->roughness = 0.0f; mat->metallic = 0.0f;
        mat->ior = m.ior > 1.0f ? m.ior : 1.5f;
        return mat;
    }
    if (starts_with(m.name, "mirror")) {
        mat->type = MaterialType::SPECULAR; mat->albedo = kd; mat->roughness = 0.0f; mat->metallic = 1.0f;
        return mat;
    }
    if (starts_with(m.name, "rough")) {
        char* end = nullptr;
        float r = std::strtof(m.name.c_str() + 5, &end);
        if (end != m.name.c_str() + 5) {
            mat->type = MaterialType::SPECULAR; mat->albedo = kd; mat->roughness = r; mat->metallic = 1.0f;
            return mat;
        }
    }
    mat->type = MaterialType::SPECULAR;
    mat->metallic = 1.0f;
    mat->roughness = 0.1f;
    if (m.name.find("red") != std::string::npos) {
        mat->albedo = glm::vec3(0.9f, 0.2f, 0.2f);
        mat->roughness = 0.1f;
    } else if (m.name.find("gold") != std::string::npos) {
        mat->albedo = glm::vec3(1.0f, 0.8f, 0.0f);
        mat->roughness = 0.05f;
    } else if (m.name.find("silver") != std::string::npos) {  // "darksilver" contains "silver"
        mat->albedo = glm::vec3(0.95f);
        mat->roughness = 0.05f;
    } else if (m.name.find("black") != std::string::npos) {
        mat->albedo = glm::vec3(0.02f);
        mat->roughness = 0.1f;
    } else {
        mat->albedo = glm::pow(kd, glm::vec3(0.8f));
        mat->albedo = glm::clamp(mat->albedo * 1.2f, 0.0f, 1.0f);
    }
    return mat;
}

// The 8 room triangles of reference src/scene.cpp:118-209 as a table: 3 corners (in units of
// roomSize for x/z and roomHeight for y), the constant normal and the three uv pairs.
struct RoomTri { float v[3][3]; float n[3]; float uv[3][2]; };
const RoomTri kRoom[8] = {
    // floor
    {{{-1, 0, -1}, {1, 0, -1}, {1, 0, 1}}, {0, 1, 0}, {{0, 0}, {1, 0}, {1, 1}}},
    {{{-1, 0, -1}, {1, 0, 1}, {-1, 0, 1}}, {0, 1, 0}, {{0, 0}, {1, 1}, {0, 1}}},
    // back wall
    {{{-1, 0, -1}, {-1, 1, -1}, {1, 1, -1}}, {0, 0, 1}, {{0, 0}, {0, 1}, {1, 1}}},
    {{{-1, 0, -1}, {1, 1, -1}, {1, 0, -1}}, {0, 0, 1}, {{0, 0}, {1, 1}, {1, 0}}},
    // left wall
    {{{-1, 0, -1}, {-1, 0, 1}, {-1, 1, 1}}, {1, 0, 0}, {{0, 0}, {1, 0}, {1, 1}}},
    {{{-1, 0, -1}, {-1, 1, 1}, {-1, 1, -1}}, {1, 0, 0}, {{0, 0}, {1, 1}, {0, 1}}},
    // right wall
    {{{1, 0, -1}, {1, 1, 1}, {1, 0, 1}}, {-1, 0, 0}, {{0, 0}, {1, 1}, {1, 0}}},
    {{{1, 0, -1}, {1, 1, -1}, {1, 1, 1}}, {-1, 0, 0}, {{0, 0}, {0, 1}, {1, 1}}},
};

}  // namespace

// ---------------------------------------------------------------------------------------------
// Scene::loadFromObj — declared at reference include/scene.hpp:94, defined in src/scene.cpp which
// cannot be compiled here.  Restatement of src/scene.cpp:8-293 (same order of operations).
// ---------------------------------------------------------------------------------------------
bool Scene::loadFromObj(const std::string& objPath) {
    if (objPath == "__arrays__") {
        // Test hook: raw triangles + materials, no room, no normalisation; then the reference
        // BVH build exactly as src/scene.cpp:290.
        const InjectedScene* in = g_injected;
        if (!in) return false;
        materials.clear();
        for (int i = 0; i < in->nmat; ++i) {
            const float* m = in->mats + 8 * i;
            auto mat = std::make_shared<Material>();
            mat->type = static_cast<MaterialType>(static_cast<int>(m[0]));
            mat->albedo = glm::vec3(m[1], m[2], m[3]);
            mat->roughness = m[4];
            mat->metallic = m[5];
            mat->ior = m[6];
            materials.push_back(mat);
        }
        triangles.clear();
        triangles.reserve(in->ntri);
        for (int i = 0; i < in->ntri; ++i) {
            const float* p = in->pos + 9 * i;
            glm::vec3 n[3];
            if (in->nrm) {
                const float* q = in->nrm + 9 * i;
                for (int k = 0; k < 3; ++k) n[k] = glm::vec3(q[3 * k], q[3 * k + 1], q[3 * k + 2]);
            }
            triangles.emplace_back(glm::vec3(p[0], p[1], p[2]), glm::vec3(p[3], p[4], p[5]), glm::vec3(p[6], p[7], p[8]),
                                   n[0], n[1], n[2], glm::vec2(0.0f), glm::vec2(0.0f), glm::vec2(0.0f),
                                   in->mat ? in->mat[i] : 0);
        }
        bvh.build(triangles);
        return true;
    }

    b2pt::obj::Mesh mesh;
    if (!b2pt::obj::parse_file(objPath, mesh)) return false;   // scene.cpp:15-20

    // scene.cpp:30-52 — bounds over ALL attrib vertices, scale to 3 units, centre.
    glm::vec3 minBounds(std::numeric_limits<float>::max());
    glm::vec3 maxBounds(-std::numeric_limits<float>::max());
    for (size_t i = 0; i + 2 < mesh.vertices.size(); i += 3) {
        glm::vec3 v(mesh.vertices[i], mesh.vertices[i + 1], mesh.vertices[i + 2]);
        minBounds = glm::min(minBounds, v);
        maxBounds = glm::max(maxBounds, v);
    }
    glm::vec3 modelSize = maxBounds - minBounds;
    float targetSize = 3.f;
    float scaleFactor = targetSize / glm::max(glm::max(modelSize.x, modelSize.y), modelSize.z);
    glm::vec3 centerOffset = (minBounds + maxBounds) * 0.5f;

    // scene.cpp:57-71 — material 0 (red specular default) and 1 (diffuse wall).
    materials.clear();
    {
        auto m0 = std::make_shared<Material>();
        m0->type = MaterialType::SPECULAR; m0->albedo = glm::vec3(0.9f, 0.2f, 0.2f);
        m0->roughness = 0.1f; m0->metallic = 1.0f;
        materials.push_back(m0);
        auto m1 = std::make_shared<Material>();
        m1->type = MaterialType::DIFFUSE; m1->albedo = glm::vec3(0.9f, 0.9f, 0.9f);
        m1->roughness = 0.95f; m1->metallic = 0.0f;
        materials.push_back(m1);
    }
    for (const auto& m : mesh.materials) materials.push_back(material_from_mtl(m));  // scene.cpp:74-114

    // scene.cpp:118-209 — the room (material 1), first in the pre-build order.
    triangles.clear();
    const float roomSize = 8.0f, roomHeight = 4.0f;
    for (const RoomTri& r : kRoom) {
        glm::vec3 v[3]; glm::vec2 uv[3];
        for (int k = 0; k < 3; ++k) {
            v[k] = glm::vec3(r.v[k][0] * roomSize, r.v[k][1] * roomHeight, r.v[k][2] * roomSize);
            uv[k] = glm::vec2(r.uv[k][0], r.uv[k][1]);
        }
        glm::vec3 n(r.n[0], r.n[1], r.n[2]);
        triangles.emplace_back(v[0], v[1], v[2], n, n, n, uv[0], uv[1], uv[2], 1);
    }

    // scene.cpp:215-282 — model faces.
    const size_t nfaces = mesh.material_ids.size();
    for (size_t f = 0; f < nfaces; ++f) {
        glm::vec3 vertices[3], normals[3];
        glm::vec2 uvs[3];
        for (int v = 0; v < 3; ++v) {
            const b2pt::obj::Index idx = mesh.indices[3 * f + v];
            glm::vec3 vertex(mesh.vertices[3 * idx.vertex_index + 0],
                             mesh.vertices[3 * idx.vertex_index + 1],
                             mesh.vertices[3 * idx.vertex_index + 2]);
            vertex = (vertex - centerOffset) * scaleFactor;   // :236
            vertex.z = -vertex.z;                             // :237
            vertex.y += 1.8f;                                 // :238
            vertices[v] = vertex;
            if (idx.normal_index >= 0) {                      // :243-250
                glm::vec3 normal(mesh.normals[3 * idx.normal_index + 0],
                                 mesh.normals[3 * idx.normal_index + 1],
                                 mesh.normals[3 * idx.normal_index + 2]);
                normal.z = -normal.z;
                normals[v] = glm::normalize(normal);
            } else if (v == 2) {                              // :251-256
                glm::vec3 edge1 = vertices[1] - vertices[0];
                glm::vec3 edge2 = vertices[2] - vertices[0];
                glm::vec3 normal = glm::normalize(glm::cross(edge1, edge2));
                normals[0] = normals[1] = normals[2] = normal;
            }
            if (idx.texcoord_index >= 0) {
                uvs[v] = glm::vec2(mesh.texcoords[2 * idx.texcoord_index + 0],
                                   mesh.texcoords[2 * idx.texcoord_index + 1]);
            } else {
                uvs[v] = glm::vec2(0.0f);
            }
        }
        int materialId = mesh.material_ids[f];                // :268-270
        if (materialId < 0) materialId = 0;
        materialId += 2;
        triangles.emplace_back(vertices[0], vertices[1], vertices[2], normals[0], normals[1], normals[2],
                               uvs[0], uvs[1], uvs[2], materialId);
    }
    if (g_prebuild_sink) *g_prebuild_sink = triangles;
    bvh.build(triangles);                                     // :290
    return true;
}

// ---------------------------------------------------------------------------------------------
// Renderer::saveImage — declared at reference include/renderer.hpp:104.  The harness variant
// copies the float framebuffer to the address encoded as "mem:<hex>" (W*H*3 floats).
// ---------------------------------------------------------------------------------------------
// Request block of the "win:" variant (ref_render_window below).
struct WindowJob {
    const Scene* scene; const Camera* camera;
    int x0, y0, x1, y1;      // pixel window [x0,x1) x [y0,y1)
    int spp;                 // samples per pixel taken by THIS Renderer
    unsigned jitter_seed;
    double* sum;             // (y1-y0)*(x1-x0)*3 running sums of the valid samples
    long long* valid;        // (y1-y0)*(x1-x0) number of valid samples
};

void Renderer::saveImage(const std::string& filename) {
    if (filename.compare(0, 4, "win:") == 0) {
        // The per-sample body of Renderer::render (renderer.hpp:62-72) for the pixels of a window, run SERIALLY on this
        // Renderer: jitter from a local mt19937 as in :55-56, tracePath with this object's own `rng`.  No thread shares
        // the generator, so unlike render() under OpenMP the member RNG is not raced.
        auto* job = reinterpret_cast<WindowJob*>(static_cast<uintptr_t>(std::strtoull(filename.c_str() + 4, nullptr, 16)));
        std::mt19937 localRng(job->jitter_seed);
        std::uniform_real_distribution<float> localDist(0.0f, 1.0f);
        const int ww = job->x1 - job->x0;
        for (int y = job->y0; y < job->y1; ++y)
            for (int x = job->x0; x < job->x1; ++x) {
                const size_t k = static_cast<size_t>(y - job->y0) * ww + (x - job->x0);
                for (int s = 0; s < job->spp; ++s) {
                    float u = (x + localDist(localRng)) / (settings.width - 1);
                    float v = (y + localDist(localRng)) / (settings.height - 1);
                    Ray ray = job->camera->getRay(u, v);
                    glm::vec3 sample = tracePath(ray, *job->scene, 0);
                    if (isValidColor(sample, "Sample computation")) {
                        job->sum[3 * k + 0] += sample.x; job->sum[3 * k + 1] += sample.y; job->sum[3 * k + 2] += sample.z;
                        job->valid[k] += 1;
                    }
                }
            }
        return;
    }
    if (filename.compare(0, 4, "mem:") != 0) return;
    float* dst = reinterpret_cast<float*>(static_cast<uintptr_t>(std::strtoull(filename.c_str() + 4, nullptr, 16)));
    for (size_t i = 0; i < frameBuffer.size(); ++i) {
        dst[3 * i + 0] = frameBuffer[i].x;
        dst[3 * i + 1] = frameBuffer[i].y;
        dst[3 * i + 2] = frameBuffer[i].z;
    }
}

// ---------------------------------------------------------------------------------------------
// C API
// ---------------------------------------------------------------------------------------------
namespace {

struct RefScene {
    std::unique_ptr<Scene> scene;   // real material ids (rendering)
    // Second BVH over the same triangles with materialId := original (pre-build) index, so the
    // unmodified intersector reports triangle identity (triangle.hpp:65 passes materialId through).
    std::vector<Triangle> idTris;
    BVH idBvh;
    std::vector<int> order;      // post-build position -> pre-build index
    std::vector<int> position;   // pre-build index -> post-build position
};

void build_id_bvh(RefScene* rs, std::vector<Triangle> preBuild) {
    for (size_t i = 0; i < preBuild.size(); ++i) preBuild[i].materialId = static_cast<int>(i);
    rs->idTris = std::move(preBuild);
    rs->idBvh.build(rs->idTris);   // same comparisons as the scene's own build => same permutation
    const size_t n = rs->idTris.size();
    rs->order.resize(n);
    rs->position.resize(n);
    for (size_t p = 0; p < n; ++p) {
        rs->order[p] = rs->idTris[p].materialId;
        rs->position[rs->order[p]] = static_cast<int>(p);
    }
}

}  // namespace

extern "C" {

// Builds a reference Scene from raw arrays (pre-build order) and runs the reference BVH::build.
void* ref_scene_from_arrays(const float* pos, const float* nrm, const int* mat, int ntri,
                            const float* mats8, int nmat) {
    CoutMute mute;
    auto rs = new RefScene();
    InjectedScene in;
    in.pos = pos; in.nrm = nrm; in.mat = mat; in.ntri = ntri; in.mats = mats8; in.nmat = nmat;
    g_injected = &in;
    rs->scene.reset(new Scene());
    rs->scene->loadFromObj("__arrays__");
    g_injected = nullptr;
    // Pre-build triangles for the id BVH.
    std::vector<Triangle> pre;
    pre.reserve(ntri);
    for (int i = 0; i < ntri; ++i) {
        const float* p = pos + 9 * i;
        pre.emplace_back(glm::vec3(p[0], p[1], p[2]), glm::vec3(p[3], p[4], p[5]), glm::vec3(p[6], p[7], p[8]),
                         glm::vec3(0.0f), glm::vec3(0.0f), glm::vec3(0.0f),
                         glm::vec2(0.0f), glm::vec2(0.0f), glm::vec2(0.0f), 0);
    }
    build_id_bvh(rs, std::move(pre));
    return rs;
}

// Loads an OBJ through the restated reference loader (room, normalisation, material rules).
void* ref_scene_from_obj(const char* path) {
    CoutMute mute;
    auto rs = new RefScene();
    rs->scene.reset(new Scene());
    std::vector<Triangle> pre;
    g_prebuild_sink = &pre;
    bool ok = rs->scene->loadFromObj(path);
    g_prebuild_sink = nullptr;
    if (!ok) { delete rs; return nullptr; }
    build_id_bvh(rs, std::move(pre));
    return rs;
}

void ref_scene_free(void* h) { delete static_cast<RefScene*>(h); }

int ref_scene_ntri(void* h) { return static_cast<int>(static_cast<RefScene*>(h)->scene->getTriangles().size()); }
int ref_scene_nmat(void* h) { return static_cast<int>(static_cast<RefScene*>(h)->scene->getMaterials().size()); }

// Triangles in the reference's post-build order (what OptixRenderer::uploadScene reads,
// reference src/gpu/optix_renderer.cu:385): pos/nrm ntri*9 floats, mat ntri ints.
void ref_scene_get_triangles(void* h, float* pos, float* nrm, int* mat) {
    const auto& tris = static_cast<RefScene*>(h)->scene->getTriangles();
    for (size_t i = 0; i < tris.size(); ++i) {
        const Triangle& t = tris[i];
        const glm::vec3 v[3] = {t.v0, t.v1, t.v2};
        const glm::vec3 n[3] = {t.n0, t.n1, t.n2};
        for (int k = 0; k < 3; ++k) {
            if (pos) { pos[9 * i + 3 * k] = v[k].x; pos[9 * i + 3 * k + 1] = v[k].y; pos[9 * i + 3 * k + 2] = v[k].z; }
            if (nrm) { nrm[9 * i + 3 * k] = n[k].x; nrm[9 * i + 3 * k + 1] = n[k].y; nrm[9 * i + 3 * k + 2] = n[k].z; }
        }
        if (mat) mat[i] = t.materialId;
    }
}

// mats8: nmat * 8 floats (type, r, g, b, roughness, metallic, ior, 0)
void ref_scene_get_materials(void* h, float* mats8) {
    const auto& mats = static_cast<RefScene*>(h)->scene->getMaterials();
    for (size_t i = 0; i < mats.size(); ++i) {
        float* m = mats8 + 8 * i;
        m[0] = static_cast<float>(static_cast<int>(mats[i]->type));
        m[1] = mats[i]->albedo.x; m[2] = mats[i]->albedo.y; m[3] = mats[i]->albedo.z;
        m[4] = mats[i]->roughness; m[5] = mats[i]->metallic; m[6] = mats[i]->ior; m[7] = 0.0f;
    }
}

// order[p] = pre-build index of the triangle at post-build position p (id BVH; for
// ref_scene_from_arrays this equals the scene's own permutation).
void ref_scene_get_order(void* h, int* order) {
    auto rs = static_cast<RefScene*>(h);
    std::memcpy(order, rs->order.data(), rs->order.size() * sizeof(int));
}

// Do the id BVH and the scene's own triangle vector agree position by position?  (They must:
// both builds see the same comparator results.)
int ref_scene_id_bvh_matches(void* h) {
    auto rs = static_cast<RefScene*>(h);
    const auto& tris = rs->scene->getTriangles();
    if (tris.size() != rs->idTris.size()) return 0;
    for (size_t i = 0; i < tris.size(); ++i) {
        if (!(tris[i].v0 == rs->idTris[i].v0 && tris[i].v1 == rs->idTris[i].v1 && tris[i].v2 == rs->idTris[i].v2)) return 0;
    }
    return 1;
}

// Closest-hit queries through the unmodified BVH::intersect (bvh.hpp:37-39).  The Ray ctor
// normalises `d` (ray.hpp:12).  tmax may be null (=> +inf).  Outputs: tri = post-build position
// of the winning triangle in the id BVH (-1 = miss), t = Intersection::t, tmax_after = ray.tMax
// on return (bvh.hpp:90), and optionally position / normal.
// use_scene_bvh=1 traces the Scene's own BVH instead and returns the materialId in `tri`.
void ref_trace_closest(void* h, const float* o, const float* d, const float* tmax, int64_t n,
                       int32_t* tri, float* t, float* pos, float* nrm, int use_scene_bvh, int nthreads) {
    auto rs = static_cast<RefScene*>(h);
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4096) num_threads(nthreads)
    for (int64_t i = 0; i < n; ++i) {
        Ray ray(glm::vec3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), glm::vec3(d[3 * i], d[3 * i + 1], d[3 * i + 2]));
        if (tmax) ray.tMax = tmax[i];
        Intersection isect;
        bool hit = use_scene_bvh ? rs->scene->intersect(ray, isect) : rs->idBvh.intersect(ray, isect);
        if (hit) {
            tri[i] = use_scene_bvh ? isect.materialId : rs->position[isect.materialId];
            if (t) t[i] = isect.t;
            if (pos) { pos[3 * i] = isect.position.x; pos[3 * i + 1] = isect.position.y; pos[3 * i + 2] = isect.position.z; }
            if (nrm) { nrm[3 * i] = isect.normal.x; nrm[3 * i + 1] = isect.normal.y; nrm[3 * i + 2] = isect.normal.z; }
        } else {
            tri[i] = -1;
            if (t) t[i] = std::numeric_limits<float>::infinity();
            if (pos) { pos[3 * i] = pos[3 * i + 1] = pos[3 * i + 2] = 0.0f; }
            if (nrm) { nrm[3 * i] = nrm[3 * i + 1] = nrm[3 * i + 2] = 0.0f; }
        }
    }
}

// Reference Renderer::render (renderer.hpp:40-102) with the reference Camera, timed exactly as
// src/main.cpp:65-70 (wall clock around render() only).  fb: W*H*3 floats, row 0 = v≈0.
// Returns seconds.  Non-deterministic by construction (random_device seeds, shared racy RNG).
double ref_render(void* h, const float* cam_pos, const float* cam_target, const float* cam_up, float fov,
                  int width, int height, int spp, int bounces, float* fb, int nthreads) {
    CoutMute mute;
    auto rs = static_cast<RefScene*>(h);
    Camera camera(glm::vec3(cam_pos[0], cam_pos[1], cam_pos[2]), glm::vec3(cam_target[0], cam_target[1], cam_target[2]),
                  glm::vec3(cam_up[0], cam_up[1], cam_up[2]), fov);
    Renderer::Settings settings;
    settings.width = width; settings.height = height; settings.samplesPerPixel = spp; settings.maxBounces = bounces;
    Renderer renderer(settings);
    int saved = omp_get_max_threads();
    if (nthreads > 0) omp_set_num_threads(nthreads);
    auto t0 = std::chrono::high_resolution_clock::now();
    renderer.render(*rs->scene, camera);
    auto t1 = std::chrono::high_resolution_clock::now();
    omp_set_num_threads(saved);
    if (fb) {
        char name[64];
        std::snprintf(name, sizeof(name), "mem:%llx", static_cast<unsigned long long>(reinterpret_cast<uintptr_t>(fb)));
        renderer.saveImage(name);
    }
    return std::chrono::duration<double>(t1 - t0).count();
}

// The reference's per-sample code (camera ray + tracePath) for a pixel window, `spp` samples per pixel split over
// `nthreads` INDEPENDENT Renderer objects (one per thread, each with its own member RNG): the race-free reading of
// renderer.hpp:53-81.  fb: (y1-y0)*(x1-x0)*3 floats = sum of valid samples / spp.  Returns seconds.
double ref_render_window(void* h, const float* cam_pos, const float* cam_target, const float* cam_up, float fov,
                         int width, int height, int x0, int y0, int x1, int y1, long long spp, int bounces, float* fb,
                         int nthreads) {
    CoutMute mute;
    auto rs = static_cast<RefScene*>(h);
    Camera camera(glm::vec3(cam_pos[0], cam_pos[1], cam_pos[2]), glm::vec3(cam_target[0], cam_target[1], cam_target[2]),
                  glm::vec3(cam_up[0], cam_up[1], cam_up[2]), fov);
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    const size_t npx = static_cast<size_t>(y1 - y0) * (x1 - x0);
    std::vector<std::vector<double>> sums(nthreads, std::vector<double>(npx * 3, 0.0));
    std::vector<std::vector<long long>> valid(nthreads, std::vector<long long>(npx, 0));
    std::random_device rd;
    std::vector<unsigned> seeds(nthreads);
    for (auto& s : seeds) s = rd();
    auto t0 = std::chrono::high_resolution_clock::now();
    #pragma omp parallel for num_threads(nthreads) schedule(static, 1)
    for (int t = 0; t < nthreads; ++t) {
        Renderer::Settings settings;
        settings.width = width; settings.height = height; settings.maxBounces = bounces;
        settings.samplesPerPixel = static_cast<int>(spp / nthreads + (t < spp % nthreads ? 1 : 0));
        Renderer renderer(settings);
        WindowJob job{rs->scene.get(), &camera, x0, y0, x1, y1, settings.samplesPerPixel, seeds[t], sums[t].data(), valid[t].data()};
        char name[64];
        std::snprintf(name, sizeof(name), "win:%llx", static_cast<unsigned long long>(reinterpret_cast<uintptr_t>(&job)));
        renderer.saveImage(name);
    }
    auto t1 = std::chrono::high_resolution_clock::now();
    for (size_t k = 0; k < npx * 3; ++k) {
        double s = 0.0;
        for (int t = 0; t < nthreads; ++t) s += sums[t][k];
        fb[k] = static_cast<float>(s / static_cast<double>(spp));
    }
    return std::chrono::duration<double>(t1 - t0).count();
}

// Reference camera ray (camera.hpp:18-29 + Ray ctor): out = origin(3), direction(3).
void ref_camera_ray(const float* cam_pos, const float* cam_target, const float* cam_up, float fov,
                    const float* uv, int64_t n, float* out) {
    Camera camera(glm::vec3(cam_pos[0], cam_pos[1], cam_pos[2]), glm::vec3(cam_target[0], cam_target[1], cam_target[2]),
                  glm::vec3(cam_up[0], cam_up[1], cam_up[2]), fov);
    for (int64_t i = 0; i < n; ++i) {
        Ray r = camera.getRay(uv[2 * i], uv[2 * i + 1]);
        out[6 * i + 0] = r.origin.x; out[6 * i + 1] = r.origin.y; out[6 * i + 2] = r.origin.z;
        out[6 * i + 3] = r.direction.x; out[6 * i + 4] = r.direction.y; out[6 * i + 5] = r.direction.z;
    }
}

// Camera basis as the reference computes it (camera.hpp:9-16): out = pos, forward, right, up (12).
void ref_camera_basis(const float* cam_pos, const float* cam_target, const float* cam_up, float fov, float* out) {
    Camera camera(glm::vec3(cam_pos[0], cam_pos[1], cam_pos[2]), glm::vec3(cam_target[0], cam_target[1], cam_target[2]),
                  glm::vec3(cam_up[0], cam_up[1], cam_up[2]), fov);
    const glm::vec3 v[4] = {camera.getPosition(), camera.getForward(), camera.getRight(), camera.getUp()};
    for (int k = 0; k < 4; ++k) { out[3 * k] = v[k].x; out[3 * k + 1] = v[k].y; out[3 * k + 2] = v[k].z; }
}

// Walks the reference tree: number of nodes, leaves, depth and nodes whose box is flat on some
// axis (always rejected by aabb.hpp:21).
void ref_tree_stats(void* h, int64_t* out4) {
    auto rs = static_cast<RefScene*>(h);
    int64_t nodes = 0, leaves = 0, flat = 0, depth = 0;
    struct Item { const BVHNode* n; int d; };
    std::vector<Item> stack;
    if (rs->idBvh.root) stack.push_back({rs->idBvh.root, 1});
    while (!stack.empty()) {
        Item it = stack.back(); stack.pop_back();
        ++nodes;
        if (it.d > depth) depth = it.d;
        const AABB& b = it.n->bounds;
        if (b.min.x == b.max.x || b.min.y == b.max.y || b.min.z == b.max.z) ++flat;
        if (it.n->isLeaf()) { ++leaves; continue; }
        stack.push_back({it.n->left, it.d + 1});
        stack.push_back({it.n->right, it.d + 1});
    }
    out4[0] = nodes; out4[1] = leaves; out4[2] = depth; out4[3] = flat;
}

int ref_omp_max_threads() { return omp_get_max_threads(); }

}  // extern "C"
