// Host harness for tests/test_tri_prefilter.py: the division-free prefilter of tri_test
// (path_tracer_ai_b200/csrc/exact.cuh) must agree with the plain statement of Triangle::intersect
// (reference include/triangle.hpp:23-58) on the decision and on every output bit.
//
//   tri_prefilter_check <n_cases> <seed>   ->  prints "cases accepted mismatches"
//
// Inputs: random triangles/rays over many scales, plus rays aimed at points ON the triangle's edges and
// vertices (u = 0, v = 0, u + v = 1 up to rounding), rays with tMax set to the hit distance +/- a few ulps,
// t near tMin, |a| near the 1e-7 threshold, tiny/huge coordinates (underflowing numerators), and for a share of
// the boundary-aimed rays all 125 neighbours obtained by moving the origin -2..2 ulps per coordinate.
#include <cmath>
using std::isnan;
using std::isinf;
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../path_tracer_ai_b200/csrc/exact.cuh"

using namespace b2pt;

static uint64_t s_state;
static inline uint32_t rnd() {   // splitmix64
    uint64_t z = (s_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)((z ^ (z >> 31)) >> 16);
}
static inline float uni() { return (float)(rnd() >> 8) * (1.0f / 16777216.0f); }
static inline float sym() { return 2.0f * uni() - 1.0f; }
static inline V3 rv(float s) { return mk3(sym() * s, sym() * s, sym() * s); }
static inline uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float ulps(float f, int k) { uint32_t u = fbits(f); u += k; float r; memcpy(&r, &u, 4); return r; }

static long long n_cases = 0, n_acc = 0, n_bad = 0;

static void check(V3 v0, V3 e1, V3 e2, V3 o, V3 d, float tmax) {
    float t0 = -1, u0 = -1, w0 = -1, t1 = -1, u1 = -1, w1 = -1;
    bool a0 = tri_test_plain(v0, e1, e2, o, d, tmax, t0, u0, w0);
    bool a1 = tri_test(v0, e1, e2, o, d, tmax, t1, u1, w1);
    ++n_cases;
    if (a0) ++n_acc;
    if (a0 != a1 || (a0 && (fbits(t0) != fbits(t1) || fbits(u0) != fbits(u1) || fbits(w0) != fbits(w1)))) {
        if (n_bad < 10)
            fprintf(stderr, "mismatch: plain=%d fast=%d t=%a/%a u=%a/%a v=%a/%a tmax=%a\n", a0, a1, t0, t1, u0, u1, w0, w1, tmax);
        ++n_bad;
    }
}

int main(int argc, char** argv) {
    long long n = argc > 1 ? atoll(argv[1]) : 1000000;
    s_state = argc > 2 ? strtoull(argv[2], nullptr, 10) : 1;
    const float scales[] = {1.0f, 1e-3f, 1e3f, 1e-12f, 1e12f, 3e-19f, 1e18f};
    for (long long it = 0; it < n; ++it) {
        float sc = scales[rnd() % 7];
        if (rnd() % 4) sc = 1.0f;
        V3 v0 = rv(sc), v1 = rv(sc), v2 = rv(sc);
        if (rnd() % 8 == 0) { v1 = vadd(v0, rv(sc * 1e-3f)); v2 = vadd(v0, rv(sc * 1e-3f)); }   // small triangle far from the origin
        V3 e1 = vsub(v1, v0), e2 = vsub(v2, v0);
        V3 o = rv(sc * 2.0f);
        int kind = rnd() % 8;
        V3 target;
        float bu = uni(), bv = uni();
        if (kind == 0) { bu = 0.0f; }                       // edge v = free, u = 0
        else if (kind == 1) { bv = 0.0f; }                  // edge v = 0
        else if (kind == 2) { bv = 1.0f - bu; }             // edge u + v = 1
        else if (kind == 3) { bu = (rnd() & 1) ? 1.0f : 0.0f; bv = (bu == 0.0f && (rnd() & 1)) ? 1.0f : 0.0f; }   // a vertex
        else if (kind == 4) { if (bu + bv > 1.0f) { bu = 1.0f - bu; bv = 1.0f - bv; } }   // interior
        // kinds 5..7: anywhere in the parallelogram and beyond
        else { bu = 1.5f * uni() - 0.25f; bv = 1.5f * uni() - 0.25f; }
        target = vadd(vadd(v0, vmuls(e1, bu)), vmuls(e2, bv));
        V3 dir = vsub(target, o);
        if (rnd() % 16 == 0) {
            // grazing: push the origin (almost) into the triangle's plane so that |a| is near the threshold
            V3 nrm = vcross(e1, e2);
            float k = vdot(vsub(o, v0), nrm) / fmaxf(vdot(nrm, nrm), 1e-30f);
            o = vsub(o, vmuls(nrm, k * (1.0f - 1e-6f * uni())));
            dir = vsub(target, o);
        }
        V3 d = vnormalize(dir);   // Ray ctor (ray.hpp:12)
        if (rnd() % 32 == 0) d = dir;   // un-normalised on purpose: the test itself does not care
        float tm = B2PT_INF;
        int tk = rnd() % 6;
        float tt, uu, vv;
        if (tk == 0) tm = 1.5f * sc;
        else if (tk == 1) tm = uni() * vlength(dir) * 2.0f;
        else if (tk >= 2 && tk <= 3 && tri_test_plain(v0, e1, e2, o, d, B2PT_INF, tt, uu, vv)) tm = ulps(tt, (int)(rnd() % 7) - 3);
        check(v0, e1, e2, o, d, tm);
        // dense probe of the decision boundaries: the same ray with its origin moved by -2..2 ulps per coordinate
        // (125 neighbours) — edge- and vertex-aimed rays then straddle u = 0, v = 0, u + v = 1 bit by bit
        if (kind <= 3 && rnd() % 16 == 0) {
            for (int ix = -2; ix <= 2; ++ix)
                for (int iy = -2; iy <= 2; ++iy)
                    for (int iz = -2; iz <= 2; ++iz)
                        check(v0, e1, e2, mk3(ulps(o.x, ix), ulps(o.y, iy), ulps(o.z, iz)), d, tm);
        }
        // the same ray started tMin away from the plane (t lands next to 0.001)
        if (rnd() % 8 == 0 && tri_test_plain(v0, e1, e2, o, d, B2PT_INF, tt, uu, vv)) {
            float back = tt - 0.001f * (1.0f + 4e-7f * sym());
            V3 o2 = vadd(o, vmuls(d, back));
            check(v0, e1, e2, o2, d, tm);
        }
    }
    printf("%lld %lld %lld\n", n_cases, n_acc, n_bad);
    return n_bad ? 1 : 0;
}
