// pt_oracle — CPU restatement of the reference's per-pixel Monte-Carlo hot path.
//
// TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library; the product (path_tracer_ai_b200/) never does.
//
// Pinning: the reference ships no tests, fixtures or golden vectors, so this
// restatement is pinned against the reference ITSELF: oracle/_ref/libref_oracle.so is the
// unmodified reference headers compiled here (oracle/ref_harness.cpp), and
// tests/test_oracle_vs_ref.py checks this file against it — bit-exact for the BVH order, node
// boxes, closest-hit ids / t and camera rays; statistically for the renderer (the reference
// seeds from std::random_device, renderer.hpp:55, so its images are not reproducible).
// Golden vectors generated from _ref are committed under tests/golden/ (tests/golden/make_golden.py).
//
// Arithmetic: plain fp32, left to right, no FMA (build with -ffp-contract=off, no -mfma), in the
// op order of GLM's scalar formulas as the reference uses them.
//
// Differences from the reference, all deliberate and documented in DESIGN.md:
//   * RNG is Philox4x32-10 keyed by (seed; pixel, sample, depth, draw) instead of mt19937 seeded
//     from random_device (renderer.hpp:55-56, :109-110) — same distributions, reproducible.
//   * tracePath (renderer.hpp:129-250) is evaluated iteratively (L = sum_k T_k * direct_k) rather
//     than recursively: same expectation, different rounding.
//   * The dielectric branch's uninitialised `brdf` (renderer.hpp:283) is treated as zero.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#include <omp.h>

namespace {

// ------------------------------------------------------------------------------------------
// glm-equivalent scalar vector math
// ------------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
inline V3 mk(float a, float b, float c) { return V3{a, b, c}; }
inline V3 add(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 sub(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 mul(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 muls(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
inline V3 smul(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
inline V3 divs(V3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
inline V3 neg(V3 a) { return mk(-a.x, -a.y, -a.z); }
inline float dot(V3 a, V3 b) { V3 t = mul(a, b); return t.x + t.y + t.z; }
inline V3 cross(V3 x, V3 y) {
    return mk(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
inline V3 normalize(V3 v) { return muls(v, inversesqrt(dot(v, v))); }
inline float length(V3 v) { return std::sqrt(dot(v, v)); }
inline float gmin(float a, float b) { return (b < a) ? b : a; }
inline float gmax(float a, float b) { return (a < b) ? b : a; }
inline V3 vmin(V3 a, V3 b) { return mk(gmin(a.x, b.x), gmin(a.y, b.y), gmin(a.z, b.z)); }
inline V3 vmax(V3 a, V3 b) { return mk(gmax(a.x, b.x), gmax(a.y, b.y), gmax(a.z, b.z)); }
inline V3 reflect(V3 I, V3 N) { return sub(I, muls(muls(N, dot(N, I)), 2.0f)); }
inline V3 refract(V3 I, V3 N, float eta) {
    float d = dot(N, I);
    float k = 1.0f - eta * eta * (1.0f - d * d);
    if (k >= 0.0f) return sub(smul(eta, I), muls(N, eta * d + std::sqrt(k)));
    return mk(0.0f, 0.0f, 0.0f);
}
inline float comp(V3 v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }
const float kPi = 3.14159265358979323846264338327950288f;
const float kInf = std::numeric_limits<float>::infinity();

// ------------------------------------------------------------------------------------------
// Scene data
// ------------------------------------------------------------------------------------------
struct Tri { V3 v0, v1, v2, n0, n1, n2; int mat; };
struct Box { V3 lo, hi; };
struct Mat { int type; V3 albedo; float roughness, metallic, ior; };
struct LightRec { V3 pos, color; float intensity; };

struct Oracle {
    std::vector<Tri> tris;        // reference post-build order
    std::vector<int> order;       // post-build position -> pre-build index
    std::vector<Mat> mats;
    std::vector<LightRec> lights;
    // Implicit reference tree (bvh.hpp:44-72): node = range [start,end); mid = start + count/2;
    // leaf iff count <= 8.  Nodes stored in DFS pre-order with explicit child links.
    struct Node { Box box; int start, end, left, right; };
    std::vector<Node> nodes;
};

// triangle.hpp:69-71  (v0 + v1 + v2) / 3.0f
inline V3 tri_center(const Tri& t) { return divs(add(add(t.v0, t.v1), t.v2), 3.0f); }
// triangle.hpp:73-77
inline Box tri_box(const Tri& t) { return Box{vmin(vmin(t.v0, t.v1), t.v2), vmax(vmax(t.v0, t.v1), t.v2)}; }
// aabb.hpp:34-39
inline int max_extent_axis(const Box& b) {
    V3 e = sub(b.hi, b.lo);
    if (e.x > e.y && e.x > e.z) return 0;
    else if (e.y > e.z) return 1;
    else return 2;
}

// bvh.hpp:44-72.  The permutation is whatever libstdc++'s std::nth_element produces for the
// comparator `center(a)[axis] < center(b)[axis]`; it depends only on comparison outcomes, so
// running it over (key, index) pairs yields the same permutation as over the 100-byte Triangles.
struct KeyIdx { float key; int idx; };

int build_node(Oracle& o, std::vector<int>& perm, const std::vector<Tri>& pre, int start, int end,
               std::vector<KeyIdx>& scratch) {
    int me = static_cast<int>(o.nodes.size());
    o.nodes.push_back(Oracle::Node());
    Box b{mk(std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()),
          mk(-std::numeric_limits<float>::max(), -std::numeric_limits<float>::max(), -std::numeric_limits<float>::max())};
    for (int i = start; i < end; ++i) {           // bvh.hpp:48-52 (aabb.hpp:27-32 merge)
        Box tb = tri_box(pre[perm[i]]);
        b.lo = vmin(b.lo, tb.lo);
        b.hi = vmax(b.hi, tb.hi);
    }
    o.nodes[me].box = b;
    o.nodes[me].start = start;
    o.nodes[me].end = end;
    o.nodes[me].left = o.nodes[me].right = -1;
    int count = end - start;
    if (count <= 8) return me;                    // bvh.hpp:55
    int axis = max_extent_axis(b);                // :60
    int mid = start + count / 2;                  // :61
    for (int i = start; i < end; ++i) scratch[i] = KeyIdx{comp(tri_center(pre[perm[i]]), axis), perm[i]};
    std::nth_element(scratch.begin() + start, scratch.begin() + mid, scratch.begin() + end,
                     [](const KeyIdx& a, const KeyIdx& c) { return a.key < c.key; });   // :63-66
    for (int i = start; i < end; ++i) perm[i] = scratch[i].idx;
    int l = build_node(o, perm, pre, start, mid, scratch);   // :68
    int r = build_node(o, perm, pre, mid, end, scratch);     // :69
    o.nodes[me].left = l;
    o.nodes[me].right = r;
    return me;
}

// aabb.hpp:13-25
inline bool box_intersect(const Box& b, V3 ro, V3 rd, float& tMin, float& tMax) {
    for (int a = 0; a < 3; ++a) {
        float invD = 1.0f / comp(rd, a);
        float t0 = (comp(b.lo, a) - comp(ro, a)) * invD;
        float t1 = (comp(b.hi, a) - comp(ro, a)) * invD;
        if (invD < 0.0f) std::swap(t0, t1);
        tMin = t0 > tMin ? t0 : tMin;
        tMax = t1 < tMax ? t1 : tMax;
        if (tMax <= tMin) return false;
    }
    return true;
}

struct RayS { V3 o, d; float tMin, tMax; };
// ray.hpp:11-12 — the ctor normalises the direction.
inline RayS make_ray(V3 o, V3 d) { return RayS{o, normalize(d), 0.001f, kInf}; }

struct Hit { float t; int tri; float u, v; bool hit; };

// triangle.hpp:23-58 (decision + t only; attribute interpolation is done by the caller).
inline bool tri_intersect(const Tri& tr, const RayS& ray, float& tOut, float& uOut, float& vOut) {
    const float EPSILON = 0.0000001f;
    V3 edge1 = sub(tr.v1, tr.v0);
    V3 edge2 = sub(tr.v2, tr.v0);
    V3 h = cross(ray.d, edge2);
    float a = dot(edge1, h);
    if (a > -EPSILON && a < EPSILON) return false;
    float f = 1.0f / a;
    V3 s = sub(ray.o, tr.v0);
    float u = f * dot(s, h);
    if (u < 0.0f || u > 1.0f) return false;
    V3 q = cross(s, edge1);
    float v = f * dot(ray.d, q);
    if (v < 0.0f || u + v > 1.0f) return false;
    float t = f * dot(edge2, q);
    if (t < ray.tMin || t > ray.tMax) return false;
    tOut = t; uOut = u; vOut = v;
    return true;
}

// bvh.hpp:74-116, recursion kept as written (fresh Intersection per child, both children always
// visited left then right, global ray.tMax shrink at :90, tie rules at :88 and :101).
bool intersect_node(const Oracle& o, int ni, RayS& ray, Hit& isect) {
    const Oracle::Node& n = o.nodes[ni];
    float tMin = ray.tMin, tMax = ray.tMax;
    if (!box_intersect(n.box, ray.o, ray.d, tMin, tMax)) return false;
    bool hit = false;
    if (n.left < 0) {
        for (int i = n.start; i < n.end; ++i) {
            float t, u, v;
            if (tri_intersect(o.tris[i], ray, t, u, v)) {
                if (t < isect.t) {
                    isect.t = t; isect.tri = i; isect.u = u; isect.v = v; isect.hit = true;
                    ray.tMax = t;
                    hit = true;
                }
            }
        }
    } else {
        Hit l{kInf, -1, 0, 0, false}, r{kInf, -1, 0, 0, false};
        bool hl = intersect_node(o, n.left, ray, l);
        bool hr = intersect_node(o, n.right, ray, r);
        if (hl && hr) { isect = (l.t < r.t) ? l : r; hit = true; }
        else if (hl) { isect = l; hit = true; }
        else if (hr) { isect = r; hit = true; }
    }
    return hit;
}

inline bool scene_intersect(const Oracle& o, RayS& ray, Hit& h) {
    h = Hit{kInf, -1, 0, 0, false};
    if (o.nodes.empty()) return false;
    return intersect_node(o, 0, ray, h);
}

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  Known-answer vectors are checked in tests/.
// ------------------------------------------------------------------------------------------
struct U4 { uint32_t x, y, z, w; };
inline U4 philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c0;
        uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c2;
        uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = static_cast<uint32_t>(p1);
        uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = static_cast<uint32_t>(p0);
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return U4{c0, c1, c2, c3};
}
inline float u01(uint32_t r) { return static_cast<float>(r >> 8) * (1.0f / 16777216.0f); }

struct Rng {
    uint32_t pix, smp, depth, k0, k1;
    U4 draw(uint32_t which) const { return philox(pix, smp, depth, which, k0, k1); }
};
// Draw indices within a (pixel, sample, depth) stream.
enum { DRAW_JITTER = 0, DRAW_COIN = 1, DRAW_SPHERE0 = 2 };

// renderer.hpp:308-319 — rejection sampling in the cube, accepted point NORMALISED.
inline V3 random_in_unit_sphere(const Rng& rng) {
    for (uint32_t k = 0;; ++k) {
        U4 r = rng.draw(DRAW_SPHERE0 + k);
        V3 p = sub(smul(2.0f, mk(u01(r.x), u01(r.y), u01(r.z))), mk(1.0f, 1.0f, 1.0f));
        if (dot(p, p) < 1.0f) return normalize(p);
    }
}

inline bool valid_color(V3 c) {   // renderer.hpp:112-123
    return !(std::isnan(c.x) || std::isnan(c.y) || std::isnan(c.z) || std::isinf(c.x) || std::isinf(c.y) || std::isinf(c.z));
}

// material.hpp:21-26
inline float schlick_fresnel(float cosTheta, float F0) {
    float x = 1.0f - cosTheta;
    float x2 = x * x;
    float x5 = x2 * x2 * x;
    return F0 + (1.0f - F0) * x5;
}
// material.hpp:28-42
inline float ggx_distribution(float NdotH, float roughness) {
    if (roughness < 0.0f) roughness = 0.0f;
    if (roughness > 1.0f) roughness = 1.0f;
    float alpha = roughness * roughness;
    float alpha2 = alpha * alpha;
    float NdotH2 = NdotH * NdotH;
    float denom = NdotH2 * (alpha2 - 1.0f) + 1.0f;
    if (denom <= 0.0f) return 0.0f;
    return alpha2 / (kPi * denom * denom);
}

struct Cam { V3 pos, forward, right, up; float fov; };
// camera.hpp:9-16
inline Cam make_camera(V3 pos, V3 target, V3 up, float fov) {
    Cam c;
    c.pos = pos;
    c.forward = normalize(sub(target, pos));
    V3 upn = normalize(up);
    c.right = normalize(cross(c.forward, upn));
    c.up = cross(c.right, c.forward);
    c.fov = fov;
    return c;
}
// camera.hpp:18-29 (+ the Ray ctor's second normalisation)
inline RayS camera_ray(const Cam& c, float u, float v) {
    float theta = c.fov * 0.01745329251994329576923690768489f;
    float h = std::tan(theta / 2.0f);
    float vh = 2.0f * h;
    float vw = vh * (16.0f / 9.0f);
    V3 horizontal = smul(vw, c.right);
    V3 vertical = smul(vh, c.up);
    V3 llc = add(sub(sub(c.pos, divs(horizontal, 2.0f)), divs(vertical, 2.0f)), c.forward);
    V3 dir = normalize(sub(add(add(llc, smul(u, horizontal)), smul(v, vertical)), c.pos));
    return make_ray(c.pos, dir);
}

struct Counters { int64_t extend = 0, shadow = 0; };

// renderer.hpp:252-301
V3 direct_lighting(const Oracle& o, V3 P, V3 n, const Mat& m, V3 viewDir, Counters& cnt) {
    V3 total = mk(0, 0, 0);
    for (const LightRec& light : o.lights) {
        V3 lightDir = sub(light.pos, P);
        float dist = length(lightDir);
        if (dist < 0.0001f) continue;
        lightDir = normalize(lightDir);
        RayS sray = make_ray(add(P, muls(n, 0.001f)), lightDir);
        sray.tMax = dist - 0.001f;
        Hit sh;
        ++cnt.shadow;
        if (!scene_intersect(o, sray, sh)) {
            float cosTheta = gmax(dot(n, lightDir), 0.0f);
            float att = light.intensity / (dist * dist);
            V3 brdf = mk(0, 0, 0);
            if (m.type == 0) {
                brdf = divs(m.albedo, kPi);
            } else if (m.type == 1) {
                V3 halfVec = normalize(add(lightDir, viewDir));
                float NdotH = gmax(dot(n, halfVec), 0.0f);
                float D = ggx_distribution(NdotH, m.roughness);
                brdf = muls(m.albedo, D);
            }
            V3 contribution = muls(muls(mul(light.color, brdf), cosTheta), att);
            if (valid_color(contribution)) total = add(total, contribution);
        }
    }
    return total;
}

// renderer.hpp:129-250, iterative form: returns L = sum_k T_k * direct_k.
V3 trace_path(const Oracle& o, RayS ray, int maxBounces, Rng rng, Counters& cnt) {
    V3 L = mk(0, 0, 0), T = mk(1, 1, 1);
    for (int depth = 0; depth < maxBounces; ++depth) {
        rng.depth = static_cast<uint32_t>(depth);
        Hit h;
        ++cnt.extend;
        if (!scene_intersect(o, ray, h)) break;                       // :135-137
        const Tri& tr = o.tris[h.tri];
        // triangle.hpp:60-62, intersection.hpp:18, renderer.hpp:139 — normalised three times.
        float w = 1.0f - h.u - h.v;
        V3 n = normalize(add(add(smul(w, tr.n0), smul(h.u, tr.n1)), smul(h.v, tr.n2)));
        n = normalize(n);
        n = normalize(n);
        V3 P = add(ray.o, muls(ray.d, h.t));                           // ray.hpp:14-16
        if (tr.mat < 0 || tr.mat >= static_cast<int>(o.mats.size())) { // :141-148 magenta
            L = add(L, mul(T, mk(1.0f, 0.0f, 1.0f)));
            break;
        }
        const Mat& m = o.mats[tr.mat];
        V3 direct = mk(0, 0, 0);
        if (m.type != 2) direct = direct_lighting(o, P, n, m, neg(ray.d), cnt);   // :160
        else cnt.shadow += static_cast<int64_t>(o.lights.size());  // reference traces them, result unused
        if (!valid_color(direct)) break;                              // :161-163
        if (m.type == 0) {                                            // :167-188
            V3 dir = random_in_unit_sphere(rng);
            if (dot(dir, n) < 0.0f) dir = neg(dir);                    // :303-306
            RayS bounce = make_ray(add(P, muls(n, 0.001f)), dir);
            float cosTheta = dot(dir, n);
            if (std::isnan(cosTheta) || std::isinf(cosTheta)) break;
            V3 brdf = divs(m.albedo, kPi);
            L = add(L, mul(T, direct));
            T = mul(T, muls(muls(muls(brdf, cosTheta), 2.0f), kPi));
            ray = bounce;
        } else if (m.type == 1) {                                     // :190-212
            V3 reflected = reflect(ray.d, n);
            if (m.roughness > 0.0f) reflected = normalize(add(reflected, smul(m.roughness, random_in_unit_sphere(rng))));
            RayS bounce = make_ray(add(P, muls(n, 0.001f)), reflected);
            float cosTheta = dot(reflected, n);
            if (std::isnan(cosTheta) || std::isinf(cosTheta)) break;
            L = add(L, mul(T, direct));
            T = mul(T, muls(m.albedo, cosTheta));
            ray = bounce;
        } else {                                                      // :214-246
            float cosTheta = dot(neg(ray.d), n);
            float etai = 1.0f, etat = m.ior;
            V3 normal = n;
            if (cosTheta < 0.0f) { cosTheta = -cosTheta; std::swap(etai, etat); normal = neg(normal); }
            float sinTheta = std::sqrt(1.0f - cosTheta * cosTheta);
            float ratio = etai / etat;
            V3 direction;
            float coin = u01(rng.draw(DRAW_COIN).x);
            if (ratio * sinTheta > 1.0f || coin < schlick_fresnel(cosTheta, (etai - etat) / (etai + etat))) {
                direction = reflect(ray.d, normal);
            } else {
                direction = refract(ray.d, normal, ratio);
            }
            float len = length(direction);
            if (std::isnan(len) || std::isinf(len)) break;
            ray = make_ray(add(P, muls(normal, 0.001f)), direction);
        }
    }
    return L;
}

}  // namespace

extern "C" {

// pos/nrm: ntri*9 floats in PRE-build order; mat: ntri; mats8: nmat*8 (type,r,g,b,rough,metal,ior,0);
// lights7: nlight*7 (pos, color, intensity).  nlight < 0 => the reference's 4 hard-coded lights
// (scene.hpp:55-80).  Runs the reference BVH build (permutes).
void* pto_create(const float* pos, const float* nrm, const int* mat, int ntri,
                 const float* mats8, int nmat, const float* lights7, int nlight) {
    Oracle* o = new Oracle();
    std::vector<Tri> pre(ntri);
    for (int i = 0; i < ntri; ++i) {
        const float* p = pos + 9 * i;
        Tri t;
        t.v0 = mk(p[0], p[1], p[2]); t.v1 = mk(p[3], p[4], p[5]); t.v2 = mk(p[6], p[7], p[8]);
        if (nrm) {
            const float* q = nrm + 9 * i;
            t.n0 = mk(q[0], q[1], q[2]); t.n1 = mk(q[3], q[4], q[5]); t.n2 = mk(q[6], q[7], q[8]);
        } else {
            t.n0 = t.n1 = t.n2 = mk(0, 0, 0);
        }
        t.mat = mat ? mat[i] : 0;
        pre[i] = t;
    }
    std::vector<int> perm(ntri);
    for (int i = 0; i < ntri; ++i) perm[i] = i;
    std::vector<KeyIdx> scratch(ntri);
    o->nodes.reserve(ntri > 8 ? ntri / 2 : 1);
    if (ntri > 0) build_node(*o, perm, pre, 0, ntri, scratch);
    o->order = perm;
    o->tris.resize(ntri);
    for (int i = 0; i < ntri; ++i) o->tris[i] = pre[perm[i]];
    for (int i = 0; i < nmat; ++i) {
        const float* m = mats8 + 8 * i;
        o->mats.push_back(Mat{static_cast<int>(m[0]), mk(m[1], m[2], m[3]), m[4], m[5], m[6]});
    }
    if (nlight < 0) {
        o->lights.push_back(LightRec{mk(2.0f, 3.5f, 2.0f), mk(1.0f, 0.95f, 0.8f), 9.0f});
        o->lights.push_back(LightRec{mk(-1.5f, 2.0f, 1.5f), mk(0.8f, 0.9f, 1.0f), 2.0f});
        o->lights.push_back(LightRec{mk(0.0f, 2.0f, -2.0f), mk(1.0f, 1.0f, 1.0f), 1.0f});
        o->lights.push_back(LightRec{mk(0.0f, 0.1f, 0.0f), mk(0.9f, 0.9f, 1.0f), 2.0f});
    } else {
        for (int i = 0; i < nlight; ++i) {
            const float* l = lights7 + 7 * i;
            o->lights.push_back(LightRec{mk(l[0], l[1], l[2]), mk(l[3], l[4], l[5]), l[6]});
        }
    }
    return o;
}

void pto_free(void* h) { delete static_cast<Oracle*>(h); }
int pto_ntri(void* h) { return static_cast<int>(static_cast<Oracle*>(h)->tris.size()); }
int pto_nnodes(void* h) { return static_cast<int>(static_cast<Oracle*>(h)->nodes.size()); }
void pto_get_order(void* h, int* order) {
    Oracle* o = static_cast<Oracle*>(h);
    std::memcpy(order, o->order.data(), o->order.size() * sizeof(int));
}
void pto_get_triangles(void* h, float* pos, float* nrm, int* mat) {
    Oracle* o = static_cast<Oracle*>(h);
    for (size_t i = 0; i < o->tris.size(); ++i) {
        const Tri& t = o->tris[i];
        const V3 v[3] = {t.v0, t.v1, t.v2};
        const V3 n[3] = {t.n0, t.n1, t.n2};
        for (int k = 0; k < 3; ++k) {
            if (pos) { pos[9 * i + 3 * k] = v[k].x; pos[9 * i + 3 * k + 1] = v[k].y; pos[9 * i + 3 * k + 2] = v[k].z; }
            if (nrm) { nrm[9 * i + 3 * k] = n[k].x; nrm[9 * i + 3 * k + 1] = n[k].y; nrm[9 * i + 3 * k + 2] = n[k].z; }
        }
        if (mat) mat[i] = t.mat;
    }
}
// Node table in DFS pre-order: boxes nnodes*6 floats, ranges nnodes*4 ints (start,end,left,right).
void pto_get_nodes(void* h, float* boxes, int* ranges) {
    Oracle* o = static_cast<Oracle*>(h);
    for (size_t i = 0; i < o->nodes.size(); ++i) {
        const Oracle::Node& n = o->nodes[i];
        if (boxes) {
            boxes[6 * i] = n.box.lo.x; boxes[6 * i + 1] = n.box.lo.y; boxes[6 * i + 2] = n.box.lo.z;
            boxes[6 * i + 3] = n.box.hi.x; boxes[6 * i + 4] = n.box.hi.y; boxes[6 * i + 5] = n.box.hi.z;
        }
        if (ranges) { ranges[4 * i] = n.start; ranges[4 * i + 1] = n.end; ranges[4 * i + 2] = n.left; ranges[4 * i + 3] = n.right; }
    }
}

// Closest hit.  d is normalised by the Ray ctor rule; tmax may be null (+inf).  tri = post-build
// position or -1; t = hit distance (+inf on miss); uv = barycentrics (may be null).
void pto_trace_closest(void* h, const float* o_, const float* d_, const float* tmax, int64_t n,
                       int32_t* tri, float* t, float* uv, int nthreads) {
    const Oracle& o = *static_cast<Oracle*>(h);
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4096) num_threads(nthreads)
    for (int64_t i = 0; i < n; ++i) {
        RayS ray = make_ray(mk(o_[3 * i], o_[3 * i + 1], o_[3 * i + 2]), mk(d_[3 * i], d_[3 * i + 1], d_[3 * i + 2]));
        if (tmax) ray.tMax = tmax[i];
        Hit hit;
        bool ok = scene_intersect(o, ray, hit);
        tri[i] = ok ? hit.tri : -1;
        if (t) t[i] = ok ? hit.t : kInf;
        if (uv) { uv[2 * i] = ok ? hit.u : 0.0f; uv[2 * i + 1] = ok ? hit.v : 0.0f; }
    }
}

// Boolean query as renderer.hpp:274-278 uses it (a closest-hit query whose result is only tested).
void pto_trace_any(void* h, const float* o_, const float* d_, const float* tmax, int64_t n,
                   uint8_t* occluded, int nthreads) {
    const Oracle& o = *static_cast<Oracle*>(h);
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4096) num_threads(nthreads)
    for (int64_t i = 0; i < n; ++i) {
        RayS ray = make_ray(mk(o_[3 * i], o_[3 * i + 1], o_[3 * i + 2]), mk(d_[3 * i], d_[3 * i + 1], d_[3 * i + 2]));
        if (tmax) ray.tMax = tmax[i];
        Hit hit;
        occluded[i] = scene_intersect(o, ray, hit) ? 1 : 0;
    }
}

void pto_philox(const uint32_t* ctr4, const uint32_t* key2, uint32_t* out4) {
    U4 r = philox(ctr4[0], ctr4[1], ctr4[2], ctr4[3], key2[0], key2[1]);
    out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w;
}

// cam13: position, forward, right, up (as Camera's getters return them) + fov.
void pto_camera_from_lookat(const float* pos, const float* target, const float* up, float fov, float* cam13) {
    Cam c = make_camera(mk(pos[0], pos[1], pos[2]), mk(target[0], target[1], target[2]), mk(up[0], up[1], up[2]), fov);
    const V3 v[4] = {c.pos, c.forward, c.right, c.up};
    for (int k = 0; k < 4; ++k) { cam13[3 * k] = v[k].x; cam13[3 * k + 1] = v[k].y; cam13[3 * k + 2] = v[k].z; }
    cam13[12] = fov;
}

void pto_camera_rays(const float* cam13, const float* uv, int64_t n, float* out6) {
    Cam c{mk(cam13[0], cam13[1], cam13[2]), mk(cam13[3], cam13[4], cam13[5]), mk(cam13[6], cam13[7], cam13[8]),
          mk(cam13[9], cam13[10], cam13[11]), cam13[12]};
    for (int64_t i = 0; i < n; ++i) {
        RayS r = camera_ray(c, uv[2 * i], uv[2 * i + 1]);
        out6[6 * i] = r.o.x; out6[6 * i + 1] = r.o.y; out6[6 * i + 2] = r.o.z;
        out6[6 * i + 3] = r.d.x; out6[6 * i + 4] = r.d.y; out6[6 * i + 5] = r.d.z;
    }
}

// renderer.hpp:40-102 with Philox streams.  fb: W*H*3 floats, row 0 = v≈0 (bottom of the view).
// Pixel window [x0,x1) x [y0,y1) lets bench.py time a bounded crop of a large frame; pixels outside
// are left untouched.  sample indices are s0 .. s0+spp-1 (divisor = spp_total, as :76 divides by
// settings.samplesPerPixel).  stats2 (optional) receives {extend rays, shadow rays}.
double pto_render(void* h, const float* cam13, int width, int height, int spp, int bounces,
                  uint64_t seed, int x0, int y0, int x1, int y1, float* fb, int64_t* stats2, int nthreads) {
    const Oracle& o = *static_cast<Oracle*>(h);
    Cam c{mk(cam13[0], cam13[1], cam13[2]), mk(cam13[3], cam13[4], cam13[5]), mk(cam13[6], cam13[7], cam13[8]),
          mk(cam13[9], cam13[10], cam13[11]), cam13[12]};
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    int64_t extend = 0, shadow = 0;
    auto t0 = std::chrono::high_resolution_clock::now();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) reduction(+ : extend, shadow)
    for (int y = y0; y < y1; ++y) {
        Counters cnt;
        for (int x = x0; x < x1; ++x) {
            V3 color = mk(0, 0, 0);
            bool hasValid = false;
            uint32_t pix = static_cast<uint32_t>(y) * static_cast<uint32_t>(width) + static_cast<uint32_t>(x);
            for (int s = 0; s < spp; ++s) {
                Rng rng{pix, static_cast<uint32_t>(s), 0u, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)};
                U4 j = rng.draw(DRAW_JITTER);
                float u = (static_cast<float>(x) + u01(j.x)) / static_cast<float>(width - 1);    // :63
                float v = (static_cast<float>(y) + u01(j.y)) / static_cast<float>(height - 1);   // :64
                RayS ray = camera_ray(c, u, v);
                V3 sample = trace_path(o, ray, bounces, rng, cnt);
                if (valid_color(sample)) { color = add(color, sample); hasValid = true; }         // :69-72
            }
            if (hasValid) color = divs(color, static_cast<float>(spp));                          // :75-76
            else color = mk(1.0f, 0.0f, 1.0f);                                                    // :78
            float* px = fb + 3 * (static_cast<size_t>(y) * width + x);
            px[0] = color.x; px[1] = color.y; px[2] = color.z;
        }
        extend += cnt.extend;
        shadow += cnt.shadow;
    }
    auto t1 = std::chrono::high_resolution_clock::now();
    if (stats2) { stats2[0] = extend; stats2[1] = shadow; }
    return std::chrono::duration<double>(t1 - t0).count();
}

int pto_omp_max_threads() { return omp_get_max_threads(); }

}  // extern "C"
