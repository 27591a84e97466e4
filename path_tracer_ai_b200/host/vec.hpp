// vec.hpp — minimal fp32 vector types for the host side, with the operation order of the GLM
// scalar formulas the reference relies on: component-wise ops, true division,
// dot = (x*x' + y*y') + z*z', normalize = v * (1/sqrt(dot)).  Host code is built without FMA
// contraction (-ffp-contract=off), so these produce the bits the reference's glm calls produce.
#pragma once
#include <cmath>

namespace b2pt {

struct vec2 {
    float x = 0.0f, y = 0.0f;
    vec2() = default;
    explicit vec2(float v) : x(v), y(v) {}
    vec2(float a, float b) : x(a), y(b) {}
};

struct vec3 {
    float x = 0.0f, y = 0.0f, z = 0.0f;
    vec3() = default;
    explicit vec3(float v) : x(v), y(v), z(v) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    const float& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};

inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }

inline float dot(const vec3& a, const vec3& b) { vec3 t = a * b; return t.x + t.y + t.z; }
inline vec3 cross(const vec3& a, const vec3& b) {
    return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
inline vec3 normalize(const vec3& v) { return v * (1.0f / std::sqrt(dot(v, v))); }
inline float min2(float a, float b) { return (b < a) ? b : a; }
inline float max2(float a, float b) { return (a < b) ? b : a; }
inline vec3 vmin(const vec3& a, const vec3& b) { return vec3(min2(a.x, b.x), min2(a.y, b.y), min2(a.z, b.z)); }
inline vec3 vmax(const vec3& a, const vec3& b) { return vec3(max2(a.x, b.x), max2(a.y, b.y), max2(a.z, b.z)); }
inline float clamp1(float x, float lo, float hi) { return min2(max2(x, lo), hi); }

}  // namespace b2pt
