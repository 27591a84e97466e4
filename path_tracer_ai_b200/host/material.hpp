// material.hpp — host mirror of the reference's include/material.hpp:6-18 (types only; the BSDF
// evaluation lives in the shade kernels, csrc/render.cu).
#pragma once
#include "vec.hpp"

namespace b2pt {

enum class MaterialType { DIFFUSE = 0, SPECULAR = 1, DIELECTRIC = 2 };

struct Material {
    MaterialType type = MaterialType::DIFFUSE;
    vec3 albedo = vec3(0.8f);
    float roughness = 0.5f;
    float metallic = 0.5f;
    float ior = 1.5f;
};

}  // namespace b2pt
