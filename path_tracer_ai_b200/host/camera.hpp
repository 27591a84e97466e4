// camera.hpp — host mirror of the reference's Camera (include/camera.hpp:7-44): same constructor,
// same getters.  getRay's per-sample work moved to k_raygen (csrc/render.cu); the basis below is
// what b2pt_camera carries across the C ABI.
#pragma once
#include "vec.hpp"
#include "../../include/b2pt.h"

namespace b2pt {

class Camera {
public:
    Camera(const vec3& position, const vec3& target, const vec3& up, float fov)
        : position(position), forward(normalize(target - position)), up(normalize(up)), fov(fov) {
        right = normalize(cross(forward, this->up));   // camera.hpp:14
        this->up = cross(right, forward);              // camera.hpp:15
    }

    const vec3& getPosition() const { return position; }
    const vec3& getForward() const { return forward; }
    const vec3& getRight() const { return right; }
    const vec3& getUp() const { return up; }
    float getFOV() const { return fov; }

    b2pt_camera toC() const {
        b2pt_camera c{};
        const vec3* src[4] = {&position, &forward, &right, &up};
        float* dst[4] = {c.position, c.forward, c.right, c.up};
        for (int k = 0; k < 4; ++k) { dst[k][0] = src[k]->x; dst[k][1] = src[k]->y; dst[k][2] = src[k]->z; }
        c.fov = fov;
        return c;
    }

private:
    vec3 position, forward, right, up;
    float fov;
};

}  // namespace b2pt
