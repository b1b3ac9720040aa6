// TEST INFRASTRUCTURE (oracle/_ref build only): see mat4x4.hpp.
#pragma once
namespace glm {
struct vec2 {
    float x, y;
};
}  // namespace glm
