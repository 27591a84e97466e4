// glm-compatible shim (oracle/ test infrastructure) — see ../glm.hpp
#pragma once
#include "../glm.hpp"
namespace glm {
template <typename T> inline constexpr T pi() { return static_cast<T>(3.14159265358979323846264338327950288); }
}
