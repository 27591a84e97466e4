// anyhit_order_sim — offline (CPU) count of the work an occlusion query does under different child-visiting
// orders.  The answer of an any-hit query does not depend on the order in which the passing children of a node are
// tried, so the order is a free parameter; this tool measures, on real shadow rays of a scene, how
// many wide nodes and triangles each policy touches before the first accepted triangle.  It is an experiment
// aid (tools/anyhit_order_study.py drives it), not part of the product and not a parity reference: plain float
// arithmetic, no claim of bit-exactness.
//
//   anyhit_order_sim <tris.bin> <rays.bin>
//     tris.bin: int32 ntri, then ntri*9 float32 (reference post-build order)
//     rays.bin: int32 nray, then nray*7 float32 (o.xyz, d.xyz normalised, T0)
//
// Policies: slot (last slot first, what the kernels do), reverse (first slot first), nearest (smallest entry
// distance first), area (most compact child first), learned (children ranked by occlusions found per visit,
// statistics from the first half of the rays, evaluated on the second half — all policies are evaluated on that
// second half).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

struct Box { float lo[3], hi[3]; };
struct Node { int start, end, left, right; Box box; };          // reference binary tree
struct Wide { int nchild; int ref[8]; };                         // children = reference node indices

static std::vector<float> P;
static std::vector<Node> nodes;
static std::vector<Wide> wide;
static std::vector<int> wide_of;   // reference node -> wide node (inner children)
static std::vector<int> height_;

static int build(int start, int end) {
    int me = (int)nodes.size();
    nodes.push_back(Node{start, end, -1, -1, {}});
    Box b;
    for (int a = 0; a < 3; ++a) { b.lo[a] = 3.4e38f; b.hi[a] = -3.4e38f; }
    if (end - start <= 8) {
        for (int t = start; t < end; ++t)
            for (int k = 0; k < 3; ++k)
                for (int a = 0; a < 3; ++a) { float v = P[9 * (size_t)t + 3 * k + a]; b.lo[a] = std::min(b.lo[a], v); b.hi[a] = std::max(b.hi[a], v); }
    } else {
        int mid = start + (end - start) / 2;
        int l = build(start, mid), r = build(mid, end);
        nodes[me].left = l; nodes[me].right = r;
        for (int a = 0; a < 3; ++a) { b.lo[a] = std::min(nodes[l].box.lo[a], nodes[r].box.lo[a]); b.hi[a] = std::max(nodes[l].box.hi[a], nodes[r].box.hi[a]); }
    }
    nodes[me].box = b;
    return me;
}
static int height(int n) {
    if (height_[n] >= 0) return height_[n];
    return height_[n] = nodes[n].left < 0 ? 0 : 1 + std::max(height(nodes[n].left), height(nodes[n].right));
}
static void collapse() {   // bottom-aligned 8-ary collapse, as build.cu
    std::vector<int> queue{0};
    wide_of.assign(nodes.size(), -1);
    for (size_t w = 0; w < queue.size(); ++w) {
        int ref = queue[w];
        wide_of[ref] = (int)w;
        std::vector<int> cur;
        if (nodes[ref].left < 0) cur = {ref};
        else {
            cur = {nodes[ref].left, nodes[ref].right};
            for (int level = 0; level < 2; ++level) {
                std::vector<int> nxt;
                for (int c : cur) {
                    if (nodes[c].left < 0 || height(c) % 3 == 0) nxt.push_back(c);
                    else { nxt.push_back(nodes[c].left); nxt.push_back(nodes[c].right); }
                }
                cur = nxt;
            }
        }
        Wide wd{};
        wd.nchild = (int)cur.size();
        for (int s = 0; s < wd.nchild; ++s) { wd.ref[s] = cur[s]; if (nodes[cur[s]].left >= 0 && cur[s] != ref) queue.push_back(cur[s]); }
        wide.push_back(wd);
    }
    // queue order == wide index order
    for (size_t w = 0; w < queue.size(); ++w) wide_of[queue[w]] = (int)w;
}

struct Ray { float o[3], d[3], inv[3], T0; };
static bool slab(const Box& b, const Ray& r, float& tmin_out) {
    float tmin = 0.001f, tmax = r.T0;
    for (int a = 0; a < 3; ++a) {
        float t0 = (b.lo[a] - r.o[a]) * r.inv[a], t1 = (b.hi[a] - r.o[a]) * r.inv[a];
        if (r.inv[a] < 0.0f) std::swap(t0, t1);
        if (t0 > tmin) tmin = t0;
        if (t1 < tmax) tmax = t1;
    }
    tmin_out = tmin;
    return tmax > tmin;
}
static bool tri_hit(int t, const Ray& r) {
    const float* p = &P[9 * (size_t)t];
    float e1[3], e2[3], h[3], s[3], q[3];
    for (int a = 0; a < 3; ++a) { e1[a] = p[3 + a] - p[a]; e2[a] = p[6 + a] - p[a]; }
    h[0] = r.d[1] * e2[2] - e2[1] * r.d[2]; h[1] = r.d[2] * e2[0] - e2[2] * r.d[0]; h[2] = r.d[0] * e2[1] - e2[0] * r.d[1];
    float a = e1[0] * h[0] + e1[1] * h[1] + e1[2] * h[2];
    if (a > -1e-7f && a < 1e-7f) return false;
    float f = 1.0f / a;
    for (int k = 0; k < 3; ++k) s[k] = r.o[k] - p[k];
    float u = f * (s[0] * h[0] + s[1] * h[1] + s[2] * h[2]);
    if (u < 0.0f || u > 1.0f) return false;
    q[0] = s[1] * e1[2] - e1[1] * s[2]; q[1] = s[2] * e1[0] - e1[2] * s[0]; q[2] = s[0] * e1[1] - e1[0] * s[1];
    float v = f * (r.d[0] * q[0] + r.d[1] * q[1] + r.d[2] * q[2]);
    if (v < 0.0f || u + v > 1.0f) return false;
    float tt = f * (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]);
    return !(tt < 0.001f || tt > r.T0);
}

enum Policy { SLOT, REVERSE, NEAREST, AREA, LEARNED, NPOLICY };
static const char* kNames[NPOLICY] = {"slot (kernels today)", "reverse", "nearest first", "compact first", "learned hits/visit"};
struct Counts { double nodes = 0, tris = 0, nodes_occ = 0, tris_occ = 0; long long rays = 0, occ = 0; };
static std::vector<double> visits, hits;   // per reference node (as a child of its wide parent)
static std::vector<float> area_of;

// Returns occluded; order = sequence in which the passing children of each node are tried.
static bool any_hit(const Ray& r, Policy pol, Counts* c, bool learn) {
    struct Item { int ref; };
    std::vector<int> stack{0};              // reference node indices; a wide root or a leaf
    std::vector<int> path;                  // children visited on the way to the current leaf (for crediting)
    double n_nodes = 0, n_tris = 0;
    bool occluded = false;
    std::vector<std::pair<int, int>> trail; // (ref, depth marker) — simple credit: every child ever popped
    std::vector<int> popped;
    while (!stack.empty() && !occluded) {
        int ref = stack.back(); stack.pop_back();
        popped.push_back(ref);
        if (nodes[ref].left < 0) {
            for (int t = nodes[ref].start; t < nodes[ref].end; ++t) { n_tris += 1; if (tri_hit(t, r)) { occluded = true; break; } }
            if (occluded && learn) hits[ref] += 1;
            continue;
        }
        const Wide& wd = wide[wide_of[ref]];
        n_nodes += 1;
        int pass[8]; float ent[8]; int np = 0;
        for (int s = 0; s < wd.nchild; ++s) { float e; if (slab(nodes[wd.ref[s]].box, r, e)) { pass[np] = wd.ref[s]; ent[np] = e; ++np; } }
        if (learn) for (int k = 0; k < np; ++k) visits[pass[k]] += 1;
        // order: the LAST pushed is tried first
        int idx[8];
        for (int k = 0; k < np; ++k) idx[k] = k;
        auto key = [&](int k) -> double {
            switch (pol) {
                case SLOT: return (double)k;                      // last slot first
                case REVERSE: return (double)-k;
                case NEAREST: return (double)-ent[k];             // smallest entry last -> first
                case AREA: return (double)-area_of[pass[k]];      // smallest area last -> first
                case LEARNED: return (hits[pass[k]] + 0.5) / (visits[pass[k]] + 1.0);   // highest rate last -> first
                default: return 0.0;
            }
        };
        std::stable_sort(idx, idx + np, [&](int a, int b) { return key(a) < key(b); });
        for (int k = 0; k < np; ++k) stack.push_back(pass[idx[k]]);
    }
    if (learn && occluded) {
        // credit every inner child on the successful descent: the ancestors of the hit leaf among the popped nodes
        int leaf = popped.back();
        for (int ref : popped)
            if (nodes[ref].left >= 0 && nodes[ref].start <= nodes[leaf].start && nodes[leaf].end <= nodes[ref].end) hits[ref] += 1;
    }
    if (c) {
        c->rays += 1; c->nodes += n_nodes; c->tris += n_tris;
        if (occluded) { c->occ += 1; c->nodes_occ += n_nodes; c->tris_occ += n_tris; }
    }
    return occluded;
}

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s tris.bin rays.bin\n", argv[0]); return 2; }
    FILE* f = std::fopen(argv[1], "rb");
    int ntri = 0;
    if (!f || std::fread(&ntri, 4, 1, f) != 1) return 3;
    P.resize(9 * (size_t)ntri);
    if (std::fread(P.data(), 4, P.size(), f) != P.size()) return 3;
    std::fclose(f);
    f = std::fopen(argv[2], "rb");
    int nray = 0;
    if (!f || std::fread(&nray, 4, 1, f) != 1) return 3;
    std::vector<float> R(7 * (size_t)nray);
    if (std::fread(R.data(), 4, R.size(), f) != R.size()) return 3;
    std::fclose(f);
    build(0, ntri);
    height_.assign(nodes.size(), -1);
    collapse();
    area_of.resize(nodes.size());
    for (size_t i = 0; i < nodes.size(); ++i) {
        const Box& b = nodes[i].box;
        float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
        area_of[i] = dx * dy + dy * dz + dz * dx;
    }
    visits.assign(nodes.size(), 0.0); hits.assign(nodes.size(), 0.0);
    auto ray_at = [&](int i) {
        Ray r;
        for (int a = 0; a < 3; ++a) { r.o[a] = R[7 * (size_t)i + a]; r.d[a] = R[7 * (size_t)i + 3 + a]; r.inv[a] = 1.0f / r.d[a]; }
        r.T0 = R[7 * (size_t)i + 6];
        return r;
    };
    const int half = nray / 2;
    for (int i = 0; i < half; ++i) { Ray r = ray_at(i); any_hit(r, SLOT, nullptr, true); }   // learning pass
    std::printf("%d triangles, %zu wide nodes, %d rays evaluated (second half), statistics from the first %d\n", ntri, wide.size(), nray - half, half);
    std::printf("%-24s %10s %10s %12s %12s %8s\n", "policy", "nodes/ray", "tris/ray", "nodes/occ", "tris/occ", "occ %");
    for (int pol = 0; pol < NPOLICY; ++pol) {
        Counts c;
        for (int i = half; i < nray; ++i) { Ray r = ray_at(i); any_hit(r, (Policy)pol, &c, false); }
        std::printf("%-24s %10.3f %10.3f %12.3f %12.3f %8.2f\n", kNames[pol], c.nodes / c.rays, c.tris / c.rays,
                    c.occ ? c.nodes_occ / c.occ : 0.0, c.occ ? c.tris_occ / c.occ : 0.0, 100.0 * c.occ / c.rays);
    }
    return 0;
}
