// OBJ + MTL parser exposing the slice of tinyobjloader's result model that the reference's
// loader consumes (reference src/scene.cpp:11-28, :215-270):
//   attrib.vertices / normals / texcoords (flat float arrays), per-face-vertex index triples
//   {vertex_index, normal_index, texcoord_index} with -1 for "absent", per-face material ids,
//   materials[] with name / diffuse[3] / ior.
// tinyobjloader itself is an un-vendored dependency of the reference (external/, git-ignored)
// and is not available here, so this is written from the OBJ/MTL format, not from its code.
//
// Supported: v (extra w / colour components ignored), vn, vt (1-3 comps), f with v, v/vt, v//vn,
// v/vt/vn, 1-based and negative (relative) indices, polygons (fan triangulation: (0,k-1,k)),
// usemtl, mtllib (searched in the OBJ's own directory, like tinyobj's default mtl_search_path),
// g/o/s (ignored: the reference walks shapes in file order, so one flat face list is equivalent).
// MTL: newmtl, Kd, Ni, Ns, d, illum (others ignored). Defaults follow tinyobj's InitMaterial
// (Kd = 0, Ni = 1, Ns = 1, d = 1).
#pragma once

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace b2pt {
namespace obj {

struct Index {
    int vertex_index = -1;
    int normal_index = -1;
    int texcoord_index = -1;
};

struct MtlMaterial {
    std::string name;
    float diffuse[3] = {0.0f, 0.0f, 0.0f};
    float ior = 1.0f;
    float shininess = 1.0f;
    float dissolve = 1.0f;
    int illum = 0;
};

struct Mesh {
    std::vector<float> vertices;    // 3 per vertex
    std::vector<float> normals;     // 3 per normal
    std::vector<float> texcoords;   // 2 per texcoord
    std::vector<Index> indices;     // 3 per (triangulated) face, file order
    std::vector<int> material_ids;  // 1 per face, -1 = none
    std::vector<MtlMaterial> materials;
    std::string warning;
    std::string error;
};

namespace detail {

inline const char* skip_ws(const char* p) {
    while (*p == ' ' || *p == '\t' || *p == '\r') ++p;
    return p;
}

inline bool is_end(const char* p) { return *p == '\0' || *p == '\n' || *p == '#'; }

// Parses up to `maxn` floats; returns how many were read.
inline int parse_floats(const char* p, float* out, int maxn) {
    int n = 0;
    while (n < maxn) {
        p = skip_ws(p);
        if (is_end(p)) break;
        char* end = nullptr;
        float v = std::strtof(p, &end);
        if (end == p) break;
        out[n++] = v;
        p = end;
    }
    return n;
}

// OBJ index fix-up: 1-based -> 0-based, negative -> relative to the current element count.
inline bool fix_index(int raw, int count, int* out) {
    if (raw > 0) { *out = raw - 1; return true; }
    if (raw < 0) { *out = count + raw; return *out >= 0; }
    return false;  // 0 is not a valid OBJ index
}

inline std::string dirname_of(const std::string& path) {
    size_t pos = path.find_last_of("/\\");
    return pos == std::string::npos ? std::string() : path.substr(0, pos + 1);
}

inline std::string trimmed_rest(const char* p) {
    p = skip_ws(p);
    std::string s(p);
    while (!s.empty() && (s.back() == '\n' || s.back() == '\r' || s.back() == ' ' || s.back() == '\t')) s.pop_back();
    return s;
}

inline bool load_mtl(const std::string& path, std::vector<MtlMaterial>& mats,
                     std::unordered_map<std::string, int>& by_name) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    char line[4096];
    int cur = -1;
    while (std::fgets(line, sizeof(line), f)) {
        const char* p = skip_ws(line);
        if (is_end(p)) continue;
        if (!std::strncmp(p, "newmtl", 6) && (p[6] == ' ' || p[6] == '\t')) {
            MtlMaterial m;
            m.name = trimmed_rest(p + 7);
            mats.push_back(m);
            cur = static_cast<int>(mats.size()) - 1;
            by_name[m.name] = cur;  // later definitions shadow earlier ones, as in tinyobj
            continue;
        }
        if (cur < 0) continue;
        MtlMaterial& m = mats[cur];
        if (p[0] == 'K' && p[1] == 'd' && (p[2] == ' ' || p[2] == '\t')) {
            parse_floats(p + 3, m.diffuse, 3);
        } else if (p[0] == 'N' && p[1] == 'i' && (p[2] == ' ' || p[2] == '\t')) {
            parse_floats(p + 3, &m.ior, 1);
        } else if (p[0] == 'N' && p[1] == 's' && (p[2] == ' ' || p[2] == '\t')) {
            parse_floats(p + 3, &m.shininess, 1);
        } else if (p[0] == 'd' && (p[1] == ' ' || p[1] == '\t')) {
            parse_floats(p + 2, &m.dissolve, 1);
        } else if (!std::strncmp(p, "illum", 5)) {
            m.illum = std::atoi(p + 5);
        }
    }
    std::fclose(f);
    return true;
}

}  // namespace detail

// Returns false (with mesh.error set) only when the OBJ file cannot be opened — a missing MTL is
// a warning, as in tinyobj.
inline bool parse_file(const std::string& path, Mesh& mesh) {
    using namespace detail;
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) {
        mesh.error = "Cannot open file [" + path + "]\n";
        return false;
    }
    const std::string base = dirname_of(path);
    std::unordered_map<std::string, int> mat_by_name;
    int cur_mat = -1;
    std::vector<Index> poly;
    std::vector<char> linebuf(1 << 16);
    while (std::fgets(linebuf.data(), static_cast<int>(linebuf.size()), f)) {
        const char* p = skip_ws(linebuf.data());
        if (is_end(p)) continue;
        if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t')) {
            float v[3] = {0, 0, 0};
            parse_floats(p + 2, v, 3);
            mesh.vertices.insert(mesh.vertices.end(), v, v + 3);
        } else if (p[0] == 'v' && p[1] == 'n' && (p[2] == ' ' || p[2] == '\t')) {
            float v[3] = {0, 0, 0};
            parse_floats(p + 3, v, 3);
            mesh.normals.insert(mesh.normals.end(), v, v + 3);
        } else if (p[0] == 'v' && p[1] == 't' && (p[2] == ' ' || p[2] == '\t')) {
            float v[3] = {0, 0, 0};
            parse_floats(p + 3, v, 3);
            mesh.texcoords.insert(mesh.texcoords.end(), v, v + 2);
        } else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t')) {
            poly.clear();
            p += 2;
            const int nv = static_cast<int>(mesh.vertices.size() / 3);
            const int nn = static_cast<int>(mesh.normals.size() / 3);
            const int nt = static_cast<int>(mesh.texcoords.size() / 2);
            bool ok = true;
            while (true) {
                p = skip_ws(p);
                if (is_end(p)) break;
                Index idx;
                char* end = nullptr;
                long vi = std::strtol(p, &end, 10);
                if (end == p) { ok = false; break; }
                if (!fix_index(static_cast<int>(vi), nv, &idx.vertex_index)) { ok = false; break; }
                p = end;
                if (*p == '/') {
                    ++p;
                    if (*p != '/') {
                        long ti = std::strtol(p, &end, 10);
                        if (end != p) { fix_index(static_cast<int>(ti), nt, &idx.texcoord_index); p = end; }
                    }
                    if (*p == '/') {
                        ++p;
                        long ni = std::strtol(p, &end, 10);
                        if (end != p) { fix_index(static_cast<int>(ni), nn, &idx.normal_index); p = end; }
                    }
                }
                poly.push_back(idx);
            }
            if (!ok || poly.size() < 3) {
                mesh.warning += "Skipping malformed face\n";
                continue;
            }
            for (size_t k = 2; k < poly.size(); ++k) {
                mesh.indices.push_back(poly[0]);
                mesh.indices.push_back(poly[k - 1]);
                mesh.indices.push_back(poly[k]);
                mesh.material_ids.push_back(cur_mat);
            }
        } else if (!std::strncmp(p, "usemtl", 6) && (p[6] == ' ' || p[6] == '\t')) {
            std::string name = trimmed_rest(p + 7);
            auto it = mat_by_name.find(name);
            if (it == mat_by_name.end()) {
                mesh.warning += "material [ '" + name + "' ] not found in .mtl\n";
                cur_mat = -1;
            } else {
                cur_mat = it->second;
            }
        } else if (!std::strncmp(p, "mtllib", 6) && (p[6] == ' ' || p[6] == '\t')) {
            std::string name = trimmed_rest(p + 7);
            if (!load_mtl(base + name, mesh.materials, mat_by_name)) {
                mesh.warning += "Material file [ " + name + " ] not found.\n";
            }
        }
        // g / o / s / anything else: ignored
    }
    std::fclose(f);
    // Out-of-range vertex references would make the caller index past the arrays; report them.
    const int nv = static_cast<int>(mesh.vertices.size() / 3);
    for (const Index& i : mesh.indices) {
        if (i.vertex_index < 0 || i.vertex_index >= nv) {
            mesh.error = "Face references a vertex that does not exist\n";
            return false;
        }
    }
    return true;
}

}  // namespace obj
}  // namespace b2pt
