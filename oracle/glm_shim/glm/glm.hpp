// glm-compatible shim — TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product.
//
// The reference (/root/reference/include/*.hpp) includes <glm/glm.hpp>; GLM itself is an
// un-vendored, un-pinned vcpkg dependency (reference build.ps1:75, CMakeLists.txt:32) and is
// absent from this image.  This header restates the scalar (packed_highp, non-SIMD) GLM
// formulas for exactly the operations the reference's hot path uses, so
// the *unmodified* reference headers compile with g++ -std=c++17 -O2 (no -mfma/-ffast-math).
// Every operation is plain fp32, evaluated left to right, no FMA.
//
// Deliberate deviation: GLM >= 0.9.9 leaves `vec3()` uninitialised; the shim zero-initialises so
// the oracle is deterministic (matters only at reference renderer.hpp:283).
#pragma once
#include <cmath>
#include <limits>
#include <utility>
#include <algorithm>
#include <cstddef>

namespace glm {

struct vec2 {
    union { float x, r, s; };
    union { float y, g, t; };
    vec2() : x(0.0f), y(0.0f) {}
    explicit vec2(float v) : x(v), y(v) {}
    vec2(float a, float b) : x(a), y(b) {}
    float& operator[](int i) { return i == 0 ? x : y; }
    const float& operator[](int i) const { return i == 0 ? x : y; }
};

struct vec3 {
    union { float x, r, s; };
    union { float y, g, t; };
    union { float z, b, p; };
    vec3() : x(0.0f), y(0.0f), z(0.0f) {}
    explicit vec3(float v) : x(v), y(v), z(v) {}
    vec3(float a, float b_, float c) : x(a), y(b_), z(c) {}
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    const float& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    vec3& operator-=(const vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    vec3& operator*=(const vec3& o) { x *= o.x; y *= o.y; z *= o.z; return *this; }
    vec3& operator*=(float s_) { x *= s_; y *= s_; z *= s_; return *this; }
    vec3& operator/=(float s_) { x /= s_; y /= s_; z /= s_; return *this; }
};

// ---- vec3 arithmetic (component-wise; division is a true division, not a reciprocal multiply)
inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(const vec3& a, const vec3& b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator+(const vec3& a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
inline vec3 operator-(const vec3& a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline bool operator==(const vec3& a, const vec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

// ---- vec2 arithmetic
inline vec2 operator+(const vec2& a, const vec2& b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator-(const vec2& a, const vec2& b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator*(const vec2& a, float s) { return vec2(a.x * s, a.y * s); }
inline vec2 operator*(float s, const vec2& a) { return vec2(s * a.x, s * a.y); }

// ---- geometric (GLM detail/func_geometric.inl, scalar path)
inline float dot(const vec3& a, const vec3& b) { vec3 t(a * b); return t.x + t.y + t.z; }
inline float dot(const vec2& a, const vec2& b) { vec2 t(a.x * b.x, a.y * b.y); return t.x + t.y; }
inline vec3 cross(const vec3& x, const vec3& y) {
    return vec3(x.y * y.z - y.y * x.z,
                x.z * y.x - y.z * x.x,
                x.x * y.y - y.x * x.y);
}
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
inline float length(const vec3& v) { return std::sqrt(dot(v, v)); }
inline vec3 normalize(const vec3& v) { return v * inversesqrt(dot(v, v)); }
inline vec3 reflect(const vec3& I, const vec3& N) { return I - N * dot(N, I) * 2.0f; }
inline vec3 refract(const vec3& I, const vec3& N, float eta) {
    float const d = dot(N, I);
    float const k = 1.0f - eta * eta * (1.0f - d * d);
    return (k >= 0.0f) ? (eta * I - (eta * d + std::sqrt(k)) * N) : vec3(0.0f);
}

// ---- common (GLM detail/func_common.inl)
inline float min(float a, float b) { return (b < a) ? b : a; }
inline float max(float a, float b) { return (a < b) ? b : a; }
inline vec3 min(const vec3& a, const vec3& b) { return vec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline vec3 max(const vec3& a, const vec3& b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline vec3 clamp(const vec3& v, float lo, float hi) { return vec3(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)); }
inline float abs(float x) { return std::fabs(x); }

// ---- trigonometric / exponential
inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }
inline float tan(float x) { return std::tan(x); }
inline float sqrt(float x) { return std::sqrt(x); }
inline float pow(float a, float b) { return std::pow(a, b); }
inline vec3 pow(const vec3& a, const vec3& b) { return vec3(std::pow(a.x, b.x), std::pow(a.y, b.y), std::pow(a.z, b.z)); }

}  // namespace glm
