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
// The file is parsed in line-aligned chunks on all host threads (parse_file); parse_file_serial is the
// line-by-line statement it is tested against.
// MTL: newmtl, Kd, Ni, Ns, d, illum (others ignored). Defaults follow tinyobj's InitMaterial
// (Kd = 0, Ni = 1, Ns = 1, d = 1).
#pragma once

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
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

// Out-of-range references would make the caller index past the arrays.  A face that names a vertex that does not
// exist fails the load; a normal / texture coordinate that does not exist is treated as absent (-1), with a warning.
template <class MeshT>
inline bool validate_indices(MeshT& mesh) {
    const int nv = static_cast<int>(mesh.vertices.size() / 3);
    const int nn = static_cast<int>(mesh.normals.size() / 3);
    const int nt = static_cast<int>(mesh.texcoords.size() / 2);
    bool warned = false;
    for (auto& i : mesh.indices) {
        if (i.vertex_index < 0 || i.vertex_index >= nv) {
            mesh.error = "Face references a vertex that does not exist\n";
            return false;
        }
        if (i.normal_index >= nn || i.normal_index < -1) { i.normal_index = -1; warned = true; }
        if (i.texcoord_index >= nt || i.texcoord_index < -1) { i.texcoord_index = -1; warned = true; }
    }
    if (warned) mesh.warning += "Face references a normal or texture coordinate that does not exist (ignored)\n";
    return true;
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

// ---- chunked parser ---------------------------------------------------------------------------------
// The file is read whole and cut into line-aligned chunks that are parsed concurrently (1M triangles: 1.6 s of
// strtof/strtol on one thread).  Two things in OBJ depend on what came before a line, and both are resolved so
// that the result is identical to a sequential read:
//   * negative (relative) indices count back from the number of v / vt / vn seen so far: a first pass counts
//     those per chunk, the exclusive prefix sums are each chunk's starting counts;
//   * the material of a face is the last `usemtl` before it, and a `usemtl` name is looked up in the `mtllib`s
//     read so far: chunks record their usemtl / mtllib lines as ordered events and tag each face with the event
//     in effect (-1 = inherited from the previous chunk); the events are replayed sequentially at merge time.
namespace detail {

struct Event { bool is_mtllib; std::string name; };

struct Chunk {
    const char* begin = nullptr;
    const char* end = nullptr;
    int nv = 0, nn = 0, nt = 0;        // elements defined inside the chunk
    int nv0 = 0, nn0 = 0, nt0 = 0;     // elements defined before it
    std::vector<float> vertices, normals, texcoords;
    std::vector<Index> indices;
    std::vector<int> face_event;        // per triangulated face: index into events of the usemtl in effect, -1 = none yet
    std::vector<Event> events;
    std::string warning;
};

inline const char* next_line(const char* p, const char* end) {
    while (p < end && *p != '\n') ++p;
    return p < end ? p + 1 : end;
}

inline std::string line_rest(const char* p) {
    p = skip_ws(p);
    const char* e = p;
    while (*e != '\0' && *e != '\n') ++e;
    std::string s(p, e);
    while (!s.empty() && (s.back() == '\r' || s.back() == ' ' || s.back() == '\t')) s.pop_back();
    return s;
}

inline void count_chunk(Chunk& c) {
    for (const char* line = c.begin; line < c.end; line = next_line(line, c.end)) {
        const char* p = skip_ws(line);
        if (p[0] != 'v') continue;
        if (p[1] == ' ' || p[1] == '\t') ++c.nv;
        else if (p[1] == 'n' && (p[2] == ' ' || p[2] == '\t')) ++c.nn;
        else if (p[1] == 't' && (p[2] == ' ' || p[2] == '\t')) ++c.nt;
    }
}

inline void parse_chunk(Chunk& c) {
    std::vector<Index> poly;
    int cur_event = -1;
    c.vertices.reserve(3 * (size_t)c.nv);
    c.normals.reserve(3 * (size_t)c.nn);
    c.texcoords.reserve(2 * (size_t)c.nt);
    for (const char* line = c.begin; line < c.end; line = next_line(line, c.end)) {
        const char* p = skip_ws(line);
        if (is_end(p)) continue;
        if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t')) {
            float v[3] = {0, 0, 0};
            parse_floats(p + 2, v, 3);
            c.vertices.insert(c.vertices.end(), v, v + 3);
        } else if (p[0] == 'v' && p[1] == 'n' && (p[2] == ' ' || p[2] == '\t')) {
            float v[3] = {0, 0, 0};
            parse_floats(p + 3, v, 3);
            c.normals.insert(c.normals.end(), v, v + 3);
        } else if (p[0] == 'v' && p[1] == 't' && (p[2] == ' ' || p[2] == '\t')) {
            float v[3] = {0, 0, 0};
            parse_floats(p + 3, v, 3);
            c.texcoords.insert(c.texcoords.end(), v, v + 2);
        } else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t')) {
            poly.clear();
            p += 2;
            const int nv = c.nv0 + static_cast<int>(c.vertices.size() / 3);
            const int nn = c.nn0 + static_cast<int>(c.normals.size() / 3);
            const int nt = c.nt0 + static_cast<int>(c.texcoords.size() / 2);
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
                        if (end != p) { if (!fix_index(static_cast<int>(ti), nt, &idx.texcoord_index)) idx.texcoord_index = -1; p = end; }
                    }
                    if (*p == '/') {
                        ++p;
                        long ni = std::strtol(p, &end, 10);
                        if (end != p) { if (!fix_index(static_cast<int>(ni), nn, &idx.normal_index)) idx.normal_index = -1; p = end; }
                    }
                }
                poly.push_back(idx);
            }
            if (!ok || poly.size() < 3) {
                c.warning += "Skipping malformed face\n";
                continue;
            }
            for (size_t k = 2; k < poly.size(); ++k) {
                c.indices.push_back(poly[0]);
                c.indices.push_back(poly[k - 1]);
                c.indices.push_back(poly[k]);
                c.face_event.push_back(cur_event);
            }
        } else if (!std::strncmp(p, "usemtl", 6) && (p[6] == ' ' || p[6] == '\t')) {
            c.events.push_back(Event{false, line_rest(p + 7)});
            cur_event = static_cast<int>(c.events.size()) - 1;
        } else if (!std::strncmp(p, "mtllib", 6) && (p[6] == ' ' || p[6] == '\t')) {
            c.events.push_back(Event{true, line_rest(p + 7)});
        }
        // g / o / s / anything else: ignored
    }
}

}  // namespace detail

// Returns false (with mesh.error set) only when the OBJ file cannot be opened — a missing MTL is
// a warning, as in tinyobj.  nthreads <= 0: hardware concurrency; chunk_bytes: smallest chunk worth a thread
// (tests pass a few bytes to put every construct on a chunk boundary).
inline bool parse_file(const std::string& path, Mesh& mesh, int nthreads = 0, size_t chunk_bytes = size_t(1) << 20) {
    using namespace detail;
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) {
        mesh.error = "Cannot open file [" + path + "]\n";
        return false;
    }
    std::vector<char> buf;
    {
        std::fseek(f, 0, SEEK_END);
        long sz = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        buf.resize(static_cast<size_t>(sz > 0 ? sz : 0) + 1);
        size_t got = sz > 0 ? std::fread(buf.data(), 1, static_cast<size_t>(sz), f) : 0;
        buf.resize(got + 1);
        buf[got] = '\0';
        std::fclose(f);
    }
    for (size_t i = 0; i + 1 < buf.size(); ++i) if (buf[i] == '\0') buf[i] = ' ';   // an embedded NUL would end the parse
    const char* begin = buf.data();
    const char* end = buf.data() + buf.size() - 1;

    if (nthreads <= 0) nthreads = static_cast<int>(std::thread::hardware_concurrency());
    if (nthreads <= 0) nthreads = 1;
    size_t nchunks = std::max<size_t>(1, std::min<size_t>(static_cast<size_t>(nthreads) * 4, static_cast<size_t>(end - begin) / std::max<size_t>(chunk_bytes, 1)));
    std::vector<Chunk> chunks(nchunks);
    {
        const char* p = begin;
        for (size_t k = 0; k < nchunks; ++k) {
            chunks[k].begin = p;
            const char* q = k + 1 == nchunks ? end : begin + (static_cast<size_t>(end - begin) * (k + 1)) / nchunks;
            if (q < p) q = p;
            if (k + 1 < nchunks) q = next_line(q, end);   // cut after a newline
            chunks[k].end = q;
            p = q;
        }
    }
    auto run = [&](void (*fn)(Chunk&)) {
        std::atomic<size_t> next{0};
        auto worker = [&]() { for (size_t k; (k = next.fetch_add(1)) < nchunks;) fn(chunks[k]); };
        int nt = static_cast<int>(std::min<size_t>(static_cast<size_t>(nthreads), nchunks));
        std::vector<std::thread> pool;
        for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
        worker();
        for (std::thread& t : pool) t.join();
    };
    run(count_chunk);
    for (size_t k = 1; k < nchunks; ++k) {
        chunks[k].nv0 = chunks[k - 1].nv0 + chunks[k - 1].nv;
        chunks[k].nn0 = chunks[k - 1].nn0 + chunks[k - 1].nn;
        chunks[k].nt0 = chunks[k - 1].nt0 + chunks[k - 1].nt;
    }
    run(parse_chunk);

    // merge, replaying the usemtl / mtllib events in file order
    const std::string base = dirname_of(path);
    std::unordered_map<std::string, int> mat_by_name;
    int cur_mat = -1;
    size_t tv = 0, tn = 0, tt = 0, ti = 0;
    for (const Chunk& c : chunks) { tv += c.vertices.size(); tn += c.normals.size(); tt += c.texcoords.size(); ti += c.indices.size(); }
    mesh.vertices.reserve(tv); mesh.normals.reserve(tn); mesh.texcoords.reserve(tt);
    mesh.indices.reserve(ti); mesh.material_ids.reserve(ti / 3);
    for (Chunk& c : chunks) {
        std::vector<int> ev_mat(c.events.size(), -1);
        const int carry_in = cur_mat;
        // warnings of a chunk: its own (malformed faces) were recorded in line order, the event warnings follow; a
        // sequential read would interleave them by line — only the text order inside `warning` differs.
        mesh.warning += c.warning;
        for (size_t e = 0; e < c.events.size(); ++e) {
            const Event& ev = c.events[e];
            if (ev.is_mtllib) {
                if (!load_mtl(base + ev.name, mesh.materials, mat_by_name)) mesh.warning += "Material file [ " + ev.name + " ] not found.\n";
            } else {
                auto it = mat_by_name.find(ev.name);
                if (it == mat_by_name.end()) {
                    mesh.warning += "material [ '" + ev.name + "' ] not found in .mtl\n";
                    cur_mat = -1;
                } else {
                    cur_mat = it->second;
                }
                ev_mat[e] = cur_mat;
            }
        }
        mesh.vertices.insert(mesh.vertices.end(), c.vertices.begin(), c.vertices.end());
        mesh.normals.insert(mesh.normals.end(), c.normals.begin(), c.normals.end());
        mesh.texcoords.insert(mesh.texcoords.end(), c.texcoords.begin(), c.texcoords.end());
        mesh.indices.insert(mesh.indices.end(), c.indices.begin(), c.indices.end());
        for (int fe : c.face_event) mesh.material_ids.push_back(fe < 0 ? carry_in : ev_mat[fe]);
        c = Chunk{};   // release the chunk's memory
    }
    return detail::validate_indices(mesh);
}

// The same parser reading line by line on one thread (kept as the statement the chunked parser is tested against).
inline bool parse_file_serial(const std::string& path, Mesh& mesh) {
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
                        if (end != p) { if (!fix_index(static_cast<int>(ti), nt, &idx.texcoord_index)) idx.texcoord_index = -1; p = end; }
                    }
                    if (*p == '/') {
                        ++p;
                        long ni = std::strtol(p, &end, 10);
                        if (end != p) { if (!fix_index(static_cast<int>(ni), nn, &idx.normal_index)) idx.normal_index = -1; p = end; }
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
    return detail::validate_indices(mesh);
}

}  // namespace obj
}  // namespace b2pt
