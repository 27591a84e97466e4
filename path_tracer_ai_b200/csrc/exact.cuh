// exact.cuh — fp32 arithmetic in the reference's operation order, bit-reproducible on the device.
//
// The reference does all of its vector maths through GLM's scalar formulas: plain
// fp32, left to right, no FMA.  Bit-exact hit ids need the same bits on sm_100a, so every operation
// below is spelled with the round-to-nearest intrinsics (__fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn/
// __fsqrt_rn), which ptxas never contracts into FFMA and which are IEEE-correct regardless of
// -use_fast_math / -fmad.  Citations are to /root/reference/include.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2pt {

struct V3 { float x, y, z; };

__host__ __device__ __forceinline__ V3 mk3(float a, float b, float c) { V3 v; v.x = a; v.y = b; v.z = c; return v; }

#ifdef __CUDA_ARCH__
#define B2PT_MUL(a, b) __fmul_rn((a), (b))
#define B2PT_ADD(a, b) __fadd_rn((a), (b))
#define B2PT_SUB(a, b) __fsub_rn((a), (b))
#define B2PT_DIV(a, b) __fdiv_rn((a), (b))
#define B2PT_SQRT(a) __fsqrt_rn((a))
#else
#define B2PT_MUL(a, b) ((a) * (b))
#define B2PT_ADD(a, b) ((a) + (b))
#define B2PT_SUB(a, b) ((a) - (b))
#define B2PT_DIV(a, b) ((a) / (b))
#define B2PT_SQRT(a) sqrtf((a))
#endif

__host__ __device__ __forceinline__ V3 vadd(V3 a, V3 b) { return mk3(B2PT_ADD(a.x, b.x), B2PT_ADD(a.y, b.y), B2PT_ADD(a.z, b.z)); }
__host__ __device__ __forceinline__ V3 vsub(V3 a, V3 b) { return mk3(B2PT_SUB(a.x, b.x), B2PT_SUB(a.y, b.y), B2PT_SUB(a.z, b.z)); }
__host__ __device__ __forceinline__ V3 vmul(V3 a, V3 b) { return mk3(B2PT_MUL(a.x, b.x), B2PT_MUL(a.y, b.y), B2PT_MUL(a.z, b.z)); }
__host__ __device__ __forceinline__ V3 vmuls(V3 a, float s) { return mk3(B2PT_MUL(a.x, s), B2PT_MUL(a.y, s), B2PT_MUL(a.z, s)); }
__host__ __device__ __forceinline__ V3 vsmul(float s, V3 a) { return mk3(B2PT_MUL(s, a.x), B2PT_MUL(s, a.y), B2PT_MUL(s, a.z)); }
__host__ __device__ __forceinline__ V3 vdivs(V3 a, float s) { return mk3(B2PT_DIV(a.x, s), B2PT_DIV(a.y, s), B2PT_DIV(a.z, s)); }
__host__ __device__ __forceinline__ V3 vneg(V3 a) { return mk3(-a.x, -a.y, -a.z); }

// glm::dot(vec3): tmp = a*b; (tmp.x + tmp.y) + tmp.z
__host__ __device__ __forceinline__ float vdot(V3 a, V3 b) {
    return B2PT_ADD(B2PT_ADD(B2PT_MUL(a.x, b.x), B2PT_MUL(a.y, b.y)), B2PT_MUL(a.z, b.z));
}
// glm::cross
__host__ __device__ __forceinline__ V3 vcross(V3 x, V3 y) {
    return mk3(B2PT_SUB(B2PT_MUL(x.y, y.z), B2PT_MUL(y.y, x.z)),
               B2PT_SUB(B2PT_MUL(x.z, y.x), B2PT_MUL(y.z, x.x)),
               B2PT_SUB(B2PT_MUL(x.x, y.y), B2PT_MUL(y.x, x.y)));
}
// glm::normalize = v * (1 / sqrt(dot(v, v)))
__host__ __device__ __forceinline__ V3 vnormalize(V3 v) { return vmuls(v, B2PT_DIV(1.0f, B2PT_SQRT(vdot(v, v)))); }
__host__ __device__ __forceinline__ float vlength(V3 v) { return B2PT_SQRT(vdot(v, v)); }
// glm::min / glm::max scalar semantics: (b < a) ? b : a   /   (a < b) ? b : a
__host__ __device__ __forceinline__ float gmin(float a, float b) { return (b < a) ? b : a; }
__host__ __device__ __forceinline__ float gmax(float a, float b) { return (a < b) ? b : a; }
// glm::reflect(I, N) = I - N * dot(N, I) * 2
__host__ __device__ __forceinline__ V3 vreflect(V3 I, V3 N) { return vsub(I, vmuls(vmuls(N, vdot(N, I)), 2.0f)); }
// glm::refract
__host__ __device__ __forceinline__ V3 vrefract(V3 I, V3 N, float eta) {
    float d = vdot(N, I);
    float k = B2PT_SUB(1.0f, B2PT_MUL(B2PT_MUL(eta, eta), B2PT_SUB(1.0f, B2PT_MUL(d, d))));
    if (k >= 0.0f) return vsub(vsmul(eta, I), vmuls(N, B2PT_ADD(B2PT_MUL(eta, d), B2PT_SQRT(k))));
    return mk3(0.0f, 0.0f, 0.0f);
}

__host__ __device__ __forceinline__ bool valid3(V3 c) {   // renderer.hpp:112-123
    return !(isnan(c.x) || isnan(c.y) || isnan(c.z) || isinf(c.x) || isinf(c.y) || isinf(c.z));
}

#define B2PT_TMIN 0.001f   // ray.hpp:8
#ifdef __CUDA_ARCH__
#define B2PT_INF __int_as_float(0x7f800000)
#else
#define B2PT_INF __builtin_inff()
#endif

// ------------------------------------------------------------------------------------------------
// AABB::intersect (aabb.hpp:13-25) for one axis, without the early-out: the caller folds the three
// axes and tests `tmax > tmin` once at the end, which is equivalent because the running tmin only
// grows and the running tmax only shrinks (a rejection at axis k implies one at axis 2).
//   invD = 1/dir (IEEE, computed once per ray); NaN t0/t1 leave the range unchanged exactly as the
//   reference's `t0 > tMin ? t0 : tMin` does — fmaxf/fminf return the non-NaN operand.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void slab_axis(float lo, float hi, float o, float invD, float& tmin, float& tmax) {
    float t0 = B2PT_MUL(B2PT_SUB(lo, o), invD);
    float t1 = B2PT_MUL(B2PT_SUB(hi, o), invD);
    if (invD < 0.0f) { float s = t0; t0 = t1; t1 = s; }
    tmin = fmaxf(tmin, t0);   // == t0 > tmin ? t0 : tmin, incl. NaN (tmin is never NaN)
    tmax = fminf(tmax, t1);
}

// The same axis with the planes already ordered along the ray: `near` is the bound the reference ends up using
// for tMin (lo if invD >= 0, hi if invD < 0 — its swap, aabb.hpp:17) and `far` the one for tMax.  Selecting the
// PLANE by the ray's sign once per node (an address offset) instead of swapping the two products per child gives
// bit-identical t values and removes two selects per axis and child.
__host__ __device__ __forceinline__ void slab_axis_nf(float near_, float far_, float o, float invD, float& tmin, float& tmax) {
    tmin = fmaxf(tmin, B2PT_MUL(B2PT_SUB(near_, o), invD));
    tmax = fminf(tmax, B2PT_MUL(B2PT_SUB(far_, o), invD));
}

// Full reference slab test; returns pass and the entry distance (the running tMin after 3 axes).
__host__ __device__ __forceinline__ bool slab_test(const float lo[3], const float hi[3], V3 o, V3 invD, float T, float& entry) {
    float tmin = B2PT_TMIN, tmax = T;
    slab_axis(lo[0], hi[0], o.x, invD.x, tmin, tmax);
    slab_axis(lo[1], hi[1], o.y, invD.y, tmin, tmax);
    slab_axis(lo[2], hi[2], o.z, invD.z, tmin, tmax);
    entry = tmin;
    return tmax > tmin;   // reference rejects on tMax <= tMin
}

// ------------------------------------------------------------------------------------------------
// Triangle::intersect (triangle.hpp:23-58): decision and t in the reference's op order.  The
// triangle is stored as v0, e1 = v1 - v0, e2 = v2 - v0 (the subtraction the reference performs
// first, done once at upload with the same rounding).  Accepts t in [B2PT_TMIN, tmax].
//
// tri_test_plain is the statement itself.  A hit additionally needs `t < +inf`: the caller in
// bvh.hpp:88 only records `tempIsect.t < closest.t` with closest.t = +inf at leaf entry, so an
// accepted t = +inf or NaN (NaN passes every `<`/`>` rejection above it) never becomes a hit.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ bool tri_test_plain(V3 v0, V3 e1, V3 e2, V3 o, V3 d, float tmax, float& t, float& u, float& v) {
    V3 h = vcross(d, e2);
    float a = vdot(e1, h);
    if (a > -0.0000001f && a < 0.0000001f) return false;
    float f = B2PT_DIV(1.0f, a);
    V3 s = vsub(o, v0);
    u = B2PT_MUL(f, vdot(s, h));
    if (u < 0.0f || u > 1.0f) return false;
    V3 q = vcross(s, e1);
    v = B2PT_MUL(f, vdot(d, q));
    if (v < 0.0f || B2PT_ADD(u, v) > 1.0f) return false;
    t = B2PT_MUL(f, vdot(e2, q));
    if (t < B2PT_TMIN || t > tmax) return false;
    return t < B2PT_INF;
}

// tri_test: the same decision and the same bits with the IEEE division taken off the rejection
// paths.  The three numerators D1 = dot(s,h), D2 = dot(d,q), D3 = dot(e2,q) are computed exactly as
// above; the reference then compares fl(fl(1/a)*D) against 0, 1, tMin, tMax.  Two roundings put the
// computed quotient within a relative 2^-22 of D/a (2^-149 absolute when it underflows), so with
// X = D*sign(a), A = |a|:
//     X < -2^-100*max(A,1)            =>  computed quotient is negative and not flushed to -0  (u<0, v<0)
//     X > A*(1+2^-20)                 =>  computed quotient > 1                               (u>1)
//     X+Y > A*(1+2^-19), X,Y >~ 0     =>  computed u+v > 1
//     Z < A*tMin*(1-2^-20)            =>  computed t < tMin
//     Z > A*tMax*(1+2^-20)            =>  computed t > tMax      (tMax >= 0; never fires for +inf)
// Each left side is a SUFFICIENT condition for the reference's rejection; everything else — hits and
// the thin bands around the thresholds — takes the division and the reference's own comparisons.
// NaN operands fail every prefilter comparison or end in t = NaN, which is never a hit (see above).
// tests/test_tri_prefilter.py pins tri_test == tri_test_plain on random and boundary-aimed inputs.
#define B2PT_PF_UP 1.00000095367431640625f     // 1 + 2^-20
#define B2PT_PF_UP2 1.0000019073486328125f     // 1 + 2^-19
#define B2PT_PF_TMIN_DN 0.00099999899975955486f // tMin * (1 - 2^-20), rounded down
#define B2PT_PF_TINY 7.888609052210118e-31f    // 2^-100
__host__ __device__ __forceinline__ float pf_signed(float x, float a) {
#ifdef __CUDA_ARCH__
    return __int_as_float(__float_as_int(x) ^ (__float_as_int(a) & (int)0x80000000));
#else
    return a < 0.0f ? -x : x;
#endif
}
__host__ __device__ __forceinline__ bool tri_test(V3 v0, V3 e1, V3 e2, V3 o, V3 d, float tmax, float& t, float& u, float& v) {
    V3 h = vcross(d, e2);
    float a = vdot(e1, h);
    if (a > -0.0000001f && a < 0.0000001f) return false;
    V3 s = vsub(o, v0);
    const float D1 = vdot(s, h);
    const float A = fabsf(a);
    const float neg = -B2PT_PF_TINY * fmaxf(A, 1.0f);
    const float X = pf_signed(D1, a);
    if (X < neg || X > A * B2PT_PF_UP) return false;
    V3 q = vcross(s, e1);
    const float D2 = vdot(d, q);
    const float Y = pf_signed(D2, a);
    if (Y < neg || X + Y > A * B2PT_PF_UP2) return false;
    const float D3 = vdot(e2, q);
    const float Z = pf_signed(D3, a);
    if (Z < A * B2PT_PF_TMIN_DN || Z > (A * tmax) * B2PT_PF_UP) return false;
    float f = B2PT_DIV(1.0f, a);
    u = B2PT_MUL(f, D1);
    if (u < 0.0f || u > 1.0f) return false;
    v = B2PT_MUL(f, D2);
    if (v < 0.0f || B2PT_ADD(u, v) > 1.0f) return false;
    t = B2PT_MUL(f, D3);
    if (t < B2PT_TMIN || t > tmax) return false;
    return t < B2PT_INF;
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (counter-based RNG): counter = (pixel, sample, depth, draw), key = seed.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    uint4 out; out.x = c0; out.y = c1; out.z = c2; out.w = c3;
    return out;
}
__host__ __device__ __forceinline__ float u01(uint32_t r) { return B2PT_MUL((float)(r >> 8), 1.0f / 16777216.0f); }

enum { DRAW_JITTER = 0, DRAW_COIN = 1, DRAW_SPHERE0 = 2 };

}  // namespace b2pt
